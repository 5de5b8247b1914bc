"""Checkpoint and wire formats after / before the path.

* safetensors I/O without any pickle: the reference loads `torch.load(...)['net'][module]` state dicts
  (models.py:696-713, Demo/Inference_LJSpeech.ipynb#cell12).  `convert_reference_checkpoint` turns that pickle ONCE
  (weights_only=True) into a `.safetensors` file; `save_model` / `load_model` round-trip a build_model() container
  through the format.  The format itself (8-byte little-endian header length, JSON header {name: {dtype, shape,
  data_offsets}}, raw little-endian tensor bytes) is written and parsed here with the standard library + numpy, so the
  load path needs neither pickle nor the `safetensors` package (files are interchangeable with it; tested).
* folded export: every weight-norm pair (weight_g, weight_v) is stored as the single folded tensor `weight` the kernels
  consume (g * v / ||v||, folded by the library's own kernel).  On load the pair is re-created as (weight_v = w,
  weight_g = st2_row_norm(w)), which folds back to exactly w (the fold and the norm share one reduction), so a folded
  file drives the same bits as the original checkpoint while the modules keep the reference's state-dict schema.
* 16-bit PCM / WAV: `pcm16` runs the conversion kernel on the device; `write_wav` adds the 44-byte RIFF header
  (the notebooks hand the fp32 array to IPython.display.Audio, Demo/Inference_LJSpeech.ipynb#cell19).
"""
from __future__ import annotations

import json
import struct
from typing import Dict, Optional

import numpy as np
import torch

_DT = {"F32": np.float32, "F16": np.float16, "I64": np.int64, "I32": np.int32, "I16": np.int16, "U8": np.uint8, "BOOL": np.bool_,
       "F64": np.float64}
_DT_REV = {np.dtype(v).name: k for k, v in _DT.items()}


def save_safetensors(path: str, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None):
    header, blobs, off = {}, [], 0
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for name in sorted(tensors):
        a = tensors[name].detach().cpu().contiguous().numpy()
        if a.dtype.name not in _DT_REV:
            raise TypeError(f"{name}: dtype {a.dtype} not supported")
        b = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        header[name] = {"dtype": _DT_REV[a.dtype.name], "shape": list(a.shape), "data_offsets": [off, off + len(b)]}
        blobs.append(b)
        off += len(b)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def load_safetensors(path: str, device="cpu"):
    """-> (dict name -> tensor, metadata dict).  Plain byte parsing: no code execution on load."""
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        if n > (100 << 20):
            raise ValueError("safetensors header too large")
        header = json.loads(f.read(n).decode("utf-8"))
        data = np.frombuffer(f.read(), dtype=np.uint8)
    meta = header.pop("__metadata__", {})
    out = {}
    for name, info in header.items():
        lo, hi = info["data_offsets"]
        dt = np.dtype(_DT[info["dtype"]]).newbyteorder("<")
        shape = tuple(info["shape"])
        if hi - lo != int(np.prod(shape, dtype=np.int64)) * dt.itemsize or hi > data.size:
            raise ValueError(f"{name}: inconsistent offsets")
        a = np.frombuffer(data[lo:hi].tobytes(), dtype=dt).reshape(shape)
        out[name] = torch.from_numpy(a.astype(dt.newbyteorder("="), copy=True)).to(device)
    return out, meta


def _wn_prefixes(sd):
    return sorted(k[:-len("weight_g")] for k in sd if k.endswith("weight_g") and (k[:-len("weight_g")] + "weight_v") in sd)


def fold_state_dict(sd: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """reference schema -> folded schema: {p}weight_g, {p}weight_v  ->  {p}weight (folded by st2_weight_norm_fold)"""
    from . import ops
    out = dict(sd)
    for p in _wn_prefixes(sd):
        g, v = out.pop(p + "weight_g"), out.pop(p + "weight_v")
        out[p + "weight"] = ops.fold_weight_norm(v.to(device).float(), g.to(device).float()).cpu()
    return out


def unfold_state_dict(folded: Dict[str, torch.Tensor], schema_keys, device) -> Dict[str, torch.Tensor]:
    """folded schema -> the reference's schema for modules that expect weight_g / weight_v: weight_v = w and
    weight_g = ||w|| (row norms from the fold kernel's own reduction: the re-fold multiplies by exactly 1.0)."""
    from . import ops
    out = {}
    want = set(schema_keys)
    for k, w in folded.items():
        if k in want:
            out[k] = w
            continue
        p = k[:-len("weight")] if k.endswith("weight") else None
        if p is not None and (p + "weight_g") in want and (p + "weight_v") in want:
            wd = w.to(device).float()
            g = ops.row_norm(wd)
            out[p + "weight_v"] = wd
            out[p + "weight_g"] = g.view(w.shape[0], *([1] * (w.dim() - 1)))
        else:
            out[k] = w            # unknown key: left to load_state_dict(strict=...) to report
    return out


SAVED_MODULES = ["bert", "bert_encoder", "predictor", "decoder", "text_encoder", "diffusion", "style_encoder", "predictor_encoder"]


def save_model(model, path: str, folded: bool = True, metadata: Optional[Dict[str, str]] = None):
    """build_model() container -> one .safetensors file, keys '<module>/<state-dict key>'."""
    tensors = {}
    dev = None
    for name in SAVED_MODULES:
        mod = model.get(name)
        if not isinstance(mod, torch.nn.Module) or not list(mod.parameters()):
            continue
        sd = {k: v for k, v in mod.state_dict().items()}
        dev = dev or next(mod.parameters()).device
        if folded:
            sd = fold_state_dict(sd, dev)
        for k, v in sd.items():
            tensors[f"{name}/{k}"] = v
    md = {"format": "styletts2_b200", "schema": "folded" if folded else "reference"}
    md.update(metadata or {})
    save_safetensors(path, tensors, md)
    return len(tensors)


def load_model(model, path: str, strict: bool = True):
    """.safetensors file (either schema) -> the container's modules, in place.  No pickle anywhere on this path."""
    tensors, meta = load_safetensors(path)
    per = {}
    for k, v in tensors.items():
        name, key = k.split("/", 1)
        per.setdefault(name, {})[key] = v
    for name, sd in per.items():
        mod = model.get(name)
        if not isinstance(mod, torch.nn.Module):
            if strict:
                raise KeyError(f"checkpoint has module '{name}' the container lacks")
            continue
        dev = next(mod.parameters()).device
        if meta.get("schema") == "folded":
            sd = unfold_state_dict(sd, mod.state_dict().keys(), dev)
        mod.load_state_dict(sd, strict=strict)
    return meta


def convert_reference_checkpoint(pth_path: str, out_path: str, folded: bool = False, device="cuda"):
    """One-time conversion of a reference `.pth` ({'net': {module: state_dict}}, models.py:696-704) to safetensors.
    Uses torch.load(weights_only=True) -- the ONLY place a pickle is read; serving loads the converted file."""
    params = torch.load(pth_path, map_location="cpu", weights_only=True)
    params = params.get("net", params)
    tensors = {}
    for name, sd in params.items():
        if not isinstance(sd, dict):
            continue
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items() if isinstance(v, torch.Tensor)}
        if folded:
            sd = fold_state_dict(sd, device)
        for k, v in sd.items():
            tensors[f"{name}/{k}"] = v
    save_safetensors(out_path, tensors, {"format": "styletts2_b200", "schema": "folded" if folded else "reference", "source": "reference .pth"})
    return len(tensors)


# ------------------------------------------------------------------ PCM / WAV
def pcm16(wav: torch.Tensor, gain: float = 1.0) -> torch.Tensor:
    """device fp32 waveform -> device int16 PCM (st2_pcm16)"""
    from . import ops
    return ops.pcm16(wav, gain)


def write_wav(path: str, pcm: torch.Tensor, rate: int = 24000):
    """mono int16 PCM (device or host, [L] or [1,L]) -> RIFF/WAVE file"""
    a = pcm.detach().reshape(-1).cpu().numpy().astype("<i2", copy=False)
    n = a.size * 2
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + n) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", n))
        f.write(a.tobytes())
