"""Deterministic synthetic weights and inputs (no checkpoints or datasets exist offline).

Trained StyleTTS 2 checkpoints are not part of the reference repo (README.md:98-102
points at Hugging Face downloads), so every parity test and the bench run on
*key-seeded* random weights: each tensor is drawn from a CPU generator seeded by
crc32(key) so the SAME state dict can be rebuilt anywhere (build container, GPU
box, the reference's own modules via load_state_dict) from key names and shapes
alone, without shipping 400 MB of weights.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def keyed_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """Value for parameter `key` (reference state-dict name) of `shape`."""
    g = _gen(key, seed)
    shape = tuple(shape)
    leaf = key.split(".")[-1]
    u = lambda: torch.rand(shape, generator=g) * 2 - 1
    if leaf == "weight_g":
        return 0.8 + 0.4 * torch.rand(shape, generator=g)  # rescaled to ||v|| by the caller
    if leaf.startswith("alpha") or ".alpha" in key or key.startswith("alpha"):
        return 1.0 + 0.3 * u()
    if leaf == "weights":  # LearnedPositionalEmbedding (modules.py:663)
        return torch.randn(shape, generator=g)
    if leaf in ("gamma",) or (leaf == "weight" and len(shape) == 1):
        return 1.0 + 0.1 * u()
    if leaf in ("beta",):
        return 0.1 * u()
    if "embedding" in key and leaf == "weight":
        return torch.randn(shape, generator=g)
    if leaf.startswith("weight_ih") or leaf.startswith("weight_hh") or leaf.startswith("bias_ih") or leaf.startswith("bias_hh"):
        hid = shape[0] // 4
        return u() / (hid ** 0.5)
    if leaf.startswith("bias"):
        return 0.05 * u()
    if len(shape) >= 2:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return u() / (fan_in ** 0.5)
    return 0.05 * u()


def keyed_state_dict(shapes: Dict[str, Tuple[int, ...]], prefix: str = "", seed: int = 0,
                     voiced: bool = True) -> Dict[str, torch.Tensor]:
    """Build a full state dict for one top-level module (`prefix` = its name in the
    build_model Munch, e.g. 'decoder').  weight_g is set to a per-row multiple of
    ||weight_v|| (what torch's weight_norm init would give, times 0.8..1.2)."""
    out = {}
    for k, shp in shapes.items():
        # the reference registers the denoiser twice (diffusion.unet is diffusion.diffusion.net,
        # models.py:668-669): both aliases must carry the same values
        canon = (prefix + "." + k).replace("diffusion.unet.", "diffusion.diffusion.net.")
        out[k] = keyed_tensor(canon, tuple(shp), seed).to(torch.float32)
    for k in list(out):
        if k.endswith(".weight_g"):
            v = out[k[: -len("weight_g")] + "weight_v"]
            nrm = v.flatten(1).norm(dim=1).view(out[k].shape)
            out[k] = out[k] * nrm
    for k in list(out):
        if k.endswith(".weight_orig"):
            # spectral_norm buffers (models.py:139-164 StyleEncoder): a trained checkpoint stores the converged
            # power-iteration vectors, so sigma = u.(W v) is the top singular value; reproduce that here
            base = k[: -len("weight_orig")]
            W = out[k].flatten(1).double()
            v = out[base + "weight_v"].double()
            v = v / v.norm().clamp_min(1e-12)
            for _ in range(12):
                u = W @ v
                u = u / u.norm().clamp_min(1e-12)
                v = W.t() @ u
                v = v / v.norm().clamp_min(1e-12)
            out[base + "weight_u"], out[base + "weight_v"] = u.float(), v.float()
    if voiced and prefix == "predictor" and "F0_proj.weight" in out:
        # random nets predict F0 ~ 0 (everything unvoiced); bias the projection so the
        # harmonic source (SineGen) is exercised with voiced and unvoiced spans.
        out["F0_proj.weight"] = out["F0_proj.weight"] * 200.0
        out["F0_proj.bias"] = torch.full_like(out["F0_proj.bias"], 150.0)
    if prefix == "predictor" and "duration_proj.linear_layer.weight" in out:
        # default-scale logits are ~0 -> every token gets round(50*sigmoid(0)) = 25 frames;
        # spread them so durations vary token to token (roughly 2..14 frames).
        out["duration_proj.linear_layer.weight"] = out["duration_proj.linear_layer.weight"] * 10.0
        out["duration_proj.linear_layer.bias"] = out["duration_proj.linear_layer.bias"] * 10.0 - 2.0
    return out


def synthetic_wave(B: int, samples: int, seed: int = 0) -> torch.Tensor:
    """Reference-clip stand-in [B, samples] at 24 kHz: a few harmonics of a gliding pitch plus a noise floor
    (every mel band carries energy, as in recorded speech)."""
    g = _gen("wave", seed)
    t = torch.arange(samples, dtype=torch.float64).unsqueeze(0) / 24000.0
    f0 = 110.0 + 120.0 * torch.rand(B, 1, generator=g).double()
    glide = 1.0 + 0.2 * torch.sin(2 * torch.pi * (0.7 + torch.rand(B, 1, generator=g).double()) * t)
    phase = 2 * torch.pi * torch.cumsum(f0 * glide, dim=1) / 24000.0
    wave = torch.zeros(B, samples, dtype=torch.float64)
    for h in range(1, 9):
        wave += (0.25 / h) * torch.sin(h * phase + 6.28 * torch.rand(B, 1, generator=g).double())
    env = 0.6 + 0.4 * torch.sin(2 * torch.pi * 3.1 * t + 6.28 * torch.rand(B, 1, generator=g).double())
    wave = wave * env + 0.02 * torch.randn(B, samples, generator=g).double()
    return wave.float()


def synthetic_f0(B: int, frames: int, seed: int = 0) -> torch.Tensor:
    """Stage-level F0 curve [B, frames]: 60..400 Hz contours with ~30 % unvoiced spans."""
    g = torch.Generator().manual_seed(1234 + seed)
    t = torch.arange(frames, dtype=torch.float32).unsqueeze(0)
    base = 120 + 180 * torch.rand(B, 1, generator=g)
    f0 = base * (1 + 0.25 * torch.sin(t * (0.02 + 0.03 * torch.rand(B, 1, generator=g)) + 6.28 * torch.rand(B, 1, generator=g)))
    f0 = f0.clamp(60, 400)
    gate = torch.rand(B, (frames + 15) // 16, generator=g).repeat_interleave(16, dim=1)[:, :frames]
    return torch.where(gate < 0.3, torch.zeros_like(f0), f0)


def synthetic_batch(B: int, N: int, multispeaker: bool, seed: int = 1):
    """Seeded inputs of SURVEY section 8(d): tokens (first token 0), equal lengths,
    bert_dur stand-in (PL-BERT is an input producer), initial noise, ref_s."""
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randint(1, 178, (B, N), generator=g)
    tokens[:, 0] = 0
    lengths = torch.full((B,), N, dtype=torch.long)
    bert_dur = torch.randn(B, N, 768, generator=g) * 0.5
    noise = torch.randn(B, 1, 256, generator=g)
    ref_s = torch.randn(B, 256, generator=g) * 0.5 if multispeaker else None
    return tokens, lengths, bert_dur, noise, ref_s
