"""Sensitivity of the waveform to the F0 curve IN THE REFERENCE ALGORITHM ITSELF (CPU oracle, pinned against the
unmodified reference): the harmonic source integrates F0 into a phase of 1e4..1e6 rad before sin(), so tiny upstream
differences decorrelate the waveform.  This script measures, on the oracle alone (no GPU, no kernels of this repo):

  1. d wav for a relative F0 perturbation of 1 ulp, 1e-6, 1e-5 and for additive Gaussian noise of 1e-4 .. 3e-3 Hz
     (3e-3 Hz = the error of the round-1 tensor-core predictor, 3e-4 Hz = the fp32-accurate recipe);
  2. the oracle's own rounding noise: F0Ntrain evaluated in float64 vs float32 on identical inputs/weights (what any
     two correct fp32 implementations -- oneDNN vs cuDNN, other thread counts -- differ by), and the waveform change
     when the decoder is fed one or the other.

Written to profiles/r02_f0_sensitivity.json; tests/test_cpu_f0_sensitivity.py asserts the qualitative facts.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
import styletts2_oracle as O  # noqa: E402
from util import oracle_sds  # noqa: E402


def measure(model="libritts", T=160, seed=5, threads=8):
    torch.set_num_threads(threads)
    mcfg = cases.MODEL_CFGS[model]
    sds = oracle_sds(model)
    g = torch.Generator().manual_seed(seed)
    B = 1
    en = torch.randn(B, 640, T, generator=g) * 0.5
    asr = torch.randn(B, 512, T, generator=g) * 0.5
    s = torch.randn(B, 128, generator=g) * 0.5
    ref = torch.randn(B, 128, generator=g) * 0.5
    L = 600 * T
    sine = torch.randn(B, L, 9, generator=g)
    ri = torch.zeros(B, 9)
    pred = sds["predictor"]
    with torch.no_grad():
        f0, n = O.f0n_train(en, s, pred)
        pred64 = {k: v.double() for k, v in pred.items()}
        f0_64, n_64 = O.f0n_train(en.double(), s.double(), pred64)

        def dec(f0_, n_):
            return O.decoder(asr, f0_, n_, ref, sds["decoder"], mcfg["decoder"], ri, sine)
        base = dec(f0, n)
        rows = {}
        ulp = torch.nextafter(f0, f0 + 1) - f0
        pert = {"rel_1ulp": f0 + ulp, "rel_1e-6": f0 * (1 + 1e-6), "rel_1e-5": f0 * (1 + 1e-5)}
        for sig in (1e-4, 3e-4, 1e-3, 3e-3):
            pert[f"gauss_{sig:g}Hz"] = f0 + sig * torch.randn(f0.shape, generator=g)
        for k, f in pert.items():
            w = dec(f, n)
            rows[k] = {"f0_maxabs_hz": float((f - f0).abs().max()), "wav_maxabs": float((w - base).abs().max())}
        w64 = dec(f0_64.float(), n_64.float())
        own = {"f0_fp32_vs_fp64_maxabs_hz": float((f0_64.float() - f0).abs().max()),
               "n_fp32_vs_fp64_maxabs": float((n_64.float() - n).abs().max()),
               "wav_maxabs_when_fed_fp64_curves": float((w64 - base).abs().max())}
    return {"model": model, "T": T, "samples": L, "f0_scale_hz": float(f0.abs().max()), "voiced_fraction": float((f0 > 10).float().mean()),
            "wav_scale": float(base.abs().max()), "perturbations": rows, "reference_own_rounding": own}


if __name__ == "__main__":
    out = [measure("libritts", 160), measure("ljspeech", 160)]
    p = os.path.join(ROOT, "profiles", "r02_f0_sensitivity.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps(out, indent=1))
