"""CPU: the C-ABI library loads, exports every symbol include/sard_b200.h declares, and the ctypes
structures have the C layout.  No compute calls (no GPU here)."""
import ctypes
import os
import re

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'sard_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sard_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from sard_b200 import _lib
    names = _declared()
    assert len(names) >= 28
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f'{n} declared in sard_b200.h but not exported'
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_struct_layouts_match_c():
    from sard_b200 import _lib
    _lib.check_struct_sizes()
    assert _lib.lib.sard_abi_version() == 1
    assert _lib.lib.sard_last_error_string() is not None


def test_product_has_no_oracle_import():
    pkg = os.path.join(ROOT, 'sard_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('oracle/', ''), f'{f} mentions the oracle'


def test_cpu_device_is_refused():
    import numpy as np
    import pytest
    import torch
    from sard_b200 import _lib, synthetic
    from sard_b200.packed import PackedRollout
    from sard_b200.rl_core import estimate_advantages
    roll = synthetic.generate_rollout(synthetic.ENV_SHAPES['point_nav'], 16, seed=1, min_exec=4, max_exec=6)
    with pytest.raises(_lib.SardError):
        PackedRollout(roll, 'cpu')
    with pytest.raises(_lib.SardError):
        estimate_advantages(torch.zeros(4), torch.ones(4), torch.zeros(4, 1), 0.99, 0.95)


def test_binding_argument_counts_match_the_header():
    """Every prototype of include/sard_b200.h has as many parameters as its ctypes binding (sard_b200/_lib.py): a call
    through a stale binding would pass garbage in the missing registers instead of failing."""
    from sard_b200 import _lib
    src = open(os.path.join(ROOT, 'include', 'sard_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = re.findall(r'\b(sard_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;', src, flags=re.S)
    assert len(protos) >= 28
    seen = set()
    for name, params in protos:
        params = ' '.join(params.split())
        n = 0 if params in ('', 'void') else params.count(',') + 1
        res, args = _lib._PROTOS[name]
        assert len(args) == n, f'{name}: header has {n} parameters, the binding {len(args)}'
        seen.add(name)
    assert seen == set(_lib.EXPORTS)
