"""Seeded parity cases shared by oracle/make_golden.py and tests/ (test infrastructure).

A case is fully described by integers: every input and every RNG draw on the path is
regenerated from CPU generators, so fixtures only need to hold the reference's OUTPUTS.
"""
from __future__ import annotations

import os
import sys

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from styletts2_b200.synthetic import synthetic_batch, synthetic_f0  # noqa: E402


_RAND, _RANDN = torch.rand, torch.randn  # originals (make_golden patches torch.rand)


class ReplayRNG:
    """Deterministic stand-in for the RNG draws on the path, in the reference's call
    order: K-1 x randn_like[B,1,256] (sampler.py:509), rand[B,9]
    (istftnet.py:155), randn_like[B,L,9] (:242), randn_like[B,L,1] (:296, unused)."""

    def __init__(self, seed: int):
        self.seed = seed

    def _g(self, idx):
        return torch.Generator().manual_seed(self.seed * 1000 + idx)

    def step_noise(self, i, shape):
        return _RANDN(shape, generator=self._g(i))

    def rand_ini(self, shape):
        return _RAND(shape, generator=self._g(900))

    def sine_noise(self, shape):
        return _RANDN(shape, generator=self._g(901))

    def unused(self, shape):
        return _RANDN(shape, generator=self._g(902))


from styletts2_b200.configs import MODEL_CFGS  # noqa: E402,F401  (single source: the product package)

REF_CONFIG_FILE = {"ljspeech": "config.yml", "libritts": "config_libritts.yml"}

# End-to-end cases (text -> waveform), small enough for the CPU oracle in seconds.
E2E_CASES = {
    "lj_e2e": dict(model="ljspeech", B=2, N=10, steps=5, embedding_scale=1.0, seed=11),
    "lj_e2e_cfg": dict(model="ljspeech", B=1, N=8, steps=3, embedding_scale=1.5, seed=12),
    "libri_e2e": dict(model="libritts", B=2, N=10, steps=4, embedding_scale=1.0, seed=13),
}

# Stage-level decoder cases with synthetic voiced F0.
DECODER_CASES = {
    "lj_dec": dict(model="ljspeech", B=2, T=20, seed=21),
    "libri_dec": dict(model="libritts", B=2, T=12, seed=22),
}


def e2e_inputs(case):
    ms = MODEL_CFGS[case["model"]]["multispeaker"]
    return synthetic_batch(case["B"], case["N"], ms, seed=case["seed"])


def decoder_inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    B, T = case["B"], case["T"]
    asr = torch.randn(B, 512, T, generator=g) * 0.5
    f0 = synthetic_f0(B, 2 * T, seed=case["seed"])
    n = torch.rand(B, 2 * T, generator=g) * 5.0
    s = torch.randn(B, 128, generator=g) * 0.5
    return asr, f0, n, s


def module_shapes_from(sd):
    return {k: tuple(v.shape) for k, v in sd.items()}
