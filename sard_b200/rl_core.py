"""RL core with the reference's call surface (khrylib/rl/core): GAE on the device, distributions."""
from __future__ import annotations

import torch
from torch.distributions import Categorical as TorchCategorical
from torch.distributions import Normal

from . import _lib
from ._lib import check, lib, ptr, stream


def estimate_advantages(rewards, masks, values, gamma, tau):
    """Drop-in for ``khrylib.rl.core.estimate_advantages`` (common.py:5-25), computed by ``sard_gae``.

    ``rewards[T]``, ``masks[T]``, ``values[T,1]`` on a CUDA device, float32 or float64; returns
    ``(advantages[T,1], returns[T,1])`` on the same device and dtype.  The reverse scan runs in
    fp64 on the GPU instead of a 50 000-iteration Python loop on the CPU.
    """
    _lib.require_cuda(rewards, 'rewards')
    T = rewards.shape[0]
    dt = values.dtype
    if dt not in (torch.float32, torch.float64):
        raise TypeError('estimate_advantages: float32 or float64 tensors expected')
    r = rewards.reshape(-1).to(dt).contiguous()
    m = masks.reshape(-1).to(dt).contiguous()
    v = values.reshape(-1).contiguous()
    adv, ret = torch.empty_like(v), torch.empty_like(v)
    scratch = torch.empty(4 * ((T + 2047) // 2048) + 8, dtype=torch.float64, device=v.device)
    fn = lib.sard_gae if dt == torch.float32 else lib.sard_gae_f64
    with torch.cuda.device(v.device):
        check(fn(ptr(r), ptr(m), ptr(v), T, float(gamma), float(tau), ptr(adv), ptr(ret), ptr(scratch), stream()))
    return adv.unsqueeze(1), ret.unsqueeze(1)


class DiagGaussian(Normal):
    """khrylib/rl/core/distributions.py:6-25."""

    def __init__(self, loc, scale):
        super().__init__(loc, scale)

    def kl(self):
        loc1, scale1 = self.loc, self.scale
        log_scale1 = self.scale.log()
        loc0, scale0, log_scale0 = self.loc.detach(), self.scale.detach(), log_scale1.detach()
        kl = log_scale1 - log_scale0 + (scale0.pow(2) + (loc0 - loc1).pow(2)) / (2.0 * scale1.pow(2)) - 0.5
        return kl.sum(1, keepdim=True)

    def log_prob(self, value):
        return super().log_prob(value).sum(1, keepdim=True)

    def mean_sample(self):
        return self.loc


class Categorical(TorchCategorical):
    """khrylib/rl/core/distributions.py:28-64."""

    def __init__(self, probs=None, logits=None, uniform_prob=0.0):
        super().__init__(probs, logits)
        self.uniform_prob = uniform_prob
        if uniform_prob > 0.0:
            self.uniform = TorchCategorical(logits=torch.zeros_like(self.logits))

    def log_prob(self, value):
        lp = super().log_prob(value).unsqueeze(1)
        if self.uniform_prob == 0.0:
            return lp
        return lp * (1 - self.uniform_prob) + self.uniform.log_prob(value).unsqueeze(1) * self.uniform_prob

    def mean_sample(self):
        return self.probs.argmax(dim=1)

    def sample(self):
        if self.uniform_prob > 0.0 and bool(torch.bernoulli(torch.tensor(self.uniform_prob))):
            return self.uniform.sample()
        return super().sample()
