"""CPU tier: safetensors reader/writer (styletts2_b200/checkpoint.py, row f4) -- round trip, interchange with the
`safetensors` package, malformed-file rejection, and the WAV header."""
import os
import struct
import wave

import numpy as np
import pytest
import torch


def _tensors():
    g = torch.Generator().manual_seed(0)
    return {"decoder/encode.conv1.weight": torch.randn(8, 4, 3, generator=g), "a/b.bias": torch.randn(5, generator=g),
            "tok": torch.arange(7, dtype=torch.int64), "h": torch.randn(3, 2, generator=g).half(), "empty": torch.zeros(0, 4)}


def test_safetensors_round_trip_and_metadata(tmp_path):
    from styletts2_b200.checkpoint import load_safetensors, save_safetensors
    p = str(tmp_path / "w.safetensors")
    t = _tensors()
    save_safetensors(p, t, {"schema": "folded"})
    got, meta = load_safetensors(p)
    assert meta == {"schema": "folded"} and set(got) == set(t)
    for k in t:
        assert got[k].dtype == t[k].dtype and torch.equal(got[k], t[k]), k


def test_interchange_with_the_safetensors_package(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    from styletts2_b200.checkpoint import load_safetensors, save_safetensors
    t = {k: v for k, v in _tensors().items() if v.numel()}
    p1, p2 = str(tmp_path / "ours.safetensors"), str(tmp_path / "theirs.safetensors")
    save_safetensors(p1, t)
    theirs = st.load_file(p1)
    st.save_file(t, p2)
    ours, _ = load_safetensors(p2)
    for k in t:
        assert torch.equal(theirs[k], t[k]) and torch.equal(ours[k], t[k])


def test_malformed_files_are_rejected(tmp_path):
    from styletts2_b200.checkpoint import load_safetensors, save_safetensors
    p = str(tmp_path / "w.safetensors")
    save_safetensors(p, {"x": torch.ones(4)})
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-4])                       # truncated payload
    with pytest.raises(ValueError):
        load_safetensors(p)
    open(p, "wb").write(struct.pack("<Q", 1 << 40) + raw[8:])   # absurd header length
    with pytest.raises(ValueError):
        load_safetensors(p)


def test_write_wav_is_a_valid_riff_file(tmp_path):
    from styletts2_b200.checkpoint import write_wav
    pcm = torch.tensor([0, 1, -1, 32767, -32768, 1234], dtype=torch.int16)
    p = str(tmp_path / "o.wav")
    write_wav(p, pcm, 24000)
    with wave.open(p, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 6)
        assert np.array_equal(np.frombuffer(w.readframes(6), dtype="<i2"), pcm.numpy())
