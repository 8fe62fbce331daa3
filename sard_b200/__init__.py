"""sard_b200: B200-native PPO update for SARD (symmetry-aware robot design).

Importing the package is cheap and CPU-safe; anything that computes imports ``sard_b200._lib``,
which needs the built ``libsard_b200.so`` and raises otherwise (there is no CPU fallback).
"""
__version__ = '0.1.0'
