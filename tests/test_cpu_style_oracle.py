"""CPU tier for SURVEY section 8 row f2 (compute_style): the oracle against the committed reference fixture,
the host-side schema of the style encoders, and the mel filterbank / DFT basis the GPU path uploads."""
import json
import os

import numpy as np
import torch

from oracle import cases, style_oracle as SO
from styletts2_b200.models import build_model, recursive_munch
from styletts2_b200.synthetic import keyed_state_dict, synthetic_wave
from util import GOLD


def _sds():
    shapes = json.load(open(os.path.join(GOLD, "state_shapes_style.json")))
    return {k: keyed_state_dict({n: tuple(s) for n, s in shapes[k].items()}, k) for k in shapes}


def test_style_oracle_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "style_libri.npz"))
    wave = torch.from_numpy(g["wave"])
    assert torch.equal(wave, synthetic_wave(wave.shape[0], wave.shape[1], 11)), "fixture input is the seeded synthetic clip"
    mel = SO.log_mel(wave)
    assert float((mel - torch.from_numpy(g["mel"])).abs().max()) <= 1e-4        # vs torchaudio's MelSpectrogram
    ref_s = SO.compute_style(_sds(), wave)
    ref = torch.from_numpy(g["ref_s"])
    assert float((ref_s - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))   # vs the unmodified StyleEncoder


def test_style_encoder_schema_matches_reference():
    shapes = json.load(open(os.path.join(GOLD, "state_shapes_style.json")))
    model = build_model(recursive_munch(cases.MODEL_CFGS["libritts"]))
    for k in ("style_encoder", "predictor_encoder"):
        ours = {n: list(v.shape) for n, v in model[k].state_dict().items()}
        assert ours == shapes[k], (set(ours) ^ set(shapes[k]))


def test_synthetic_spectral_norm_buffers_are_converged():
    sd = _sds()["style_encoder"]
    for p in ("shared.0", "shared.2.conv1", "shared.4.downsample_res.conv", "shared.6"):
        W = sd[p + ".weight_orig"].flatten(1).double()
        sigma = float(sd[p + ".weight_u"].double() @ (W @ sd[p + ".weight_v"].double()))
        top = float(torch.linalg.svdvals(W)[0])
        assert 0.9 * top <= sigma <= top * (1 + 1e-6), (p, sigma, top)


def test_logmel_host_constants_match_oracle():
    from styletts2_b200.style import LogMel
    lm = LogMel()
    assert torch.equal(lm.fb.weight.detach().t().contiguous(), SO.mel_filterbank())
    assert torch.equal(lm.window, torch.hann_window(1200, periodic=True))
    # the DFT basis reproduces torch.stft's power spectrum on a random frame
    x = torch.randn(1, 4096)
    spec = SO.power_spectrogram(x)[0, :, 4]            # frame 4: samples 4*300-1024 .. +2048 of the padded signal
    start = 4 * 300 - 600
    fr = x[0, start:start + 1200] * lm.window
    y = lm.dft.weight.detach().double() @ fr.double()
    p = y[:1025] ** 2 + y[1025:] ** 2
    assert float((p.float() - spec).abs().max()) <= 2e-4 * float(spec.abs().max())
