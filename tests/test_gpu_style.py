"""GPU parity for SURVEY section 8 row f2 (compute_style): every new C-ABI entry point against torch fp32 / the oracle,
then the whole path against the fixture recorded from the unmodified reference (tests/golden/style_libri.npz)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import style_oracle as SO
from util import golden, gpu_model, maxdiff, record

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return maxdiff(a, b) / max(1e-12, float(b.abs().max()))


@pytest.mark.parametrize("shape", [
    # B, Cin, H, W, Cout, k, pad
    (2, 1, 80, 121, 64, 3, 1),      # shared.0
    (2, 64, 40, 61, 128, 3, 1),     # ResBlk conv2 (odd width, partial tiles)
    (1, 128, 20, 31, 128, 3, 1),
    (2, 64, 16, 23, 128, 1, 0),     # conv1x1 shortcut (no bias)
    (2, 96, 5, 9, 72, 5, 0),        # final 5x5 valid conv; Cin, Cout not multiples of the tiles
    (1, 20, 9, 17, 70, 3, 1),
])
def test_conv2d_matches_torch(shape):
    from styletts2_b200 import ops
    B, Cin, H, W, Cout, k, pad = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    bias = torch.randn(Cout, generator=g) * 0.1 if k != 1 else None
    want = F.conv2d(F.leaky_relu(x, 0.2), w, bias, padding=pad)
    res = torch.randn(want.shape, generator=g)
    want2 = (want + res) / math.sqrt(2)
    wt = w.flatten(1).t().contiguous().cuda()        # [Cin*k*k][Cout]
    got = ops.conv2d(x.cuda(), wt, None if bias is None else bias.cuda(), cout=Cout, kh=k, kw=k, pad=pad, pre_act=True)
    got2 = ops.conv2d(x.cuda(), wt, None if bias is None else bias.cuda(), cout=Cout, kh=k, kw=k, pad=pad, pre_act=True,
                      res=res.cuda(), out_scale=1 / math.sqrt(2))
    assert got.shape == want.shape
    assert _rel(got, want) <= 2e-6 and _rel(got2, want2) <= 2e-6      # fp32 FMA chains of <= 4608 terms
    plain = ops.conv2d(x.cuda(), wt, None, cout=Cout, kh=k, kw=k, pad=pad)
    assert _rel(plain, F.conv2d(x, w, None, padding=pad)) <= 2e-6


@pytest.mark.parametrize("hw", [(80, 121), (40, 61), (10, 16), (5, 7)])
def test_depthwise_downsample_and_pool_match_torch(hw):
    from styletts2_b200 import ops
    H, W = hw
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(2, 48, H, W, generator=g)
    w = torch.randn(48, 1, 3, 3, generator=g)
    b = torch.randn(48, generator=g)
    want = F.conv2d(x, w, b, stride=2, padding=1, groups=48)
    got = ops.dwconv3x3_s2(x.cuda(), w.flatten(1).t().contiguous().cuda(), b.cuda())
    assert got.shape == want.shape and _rel(got, want) <= 1e-6
    if H >= 2:
        wantp = SO.downsample_half(x)
        gotp = ops.avgpool_half(x.cuda())
        assert gotp.shape == wantp.shape and _rel(gotp, wantp) <= 1e-6
    m = ops.mean_hw_lrelu(x.cuda(), 0.2)
    assert _rel(m, F.leaky_relu(x.mean(dim=(2, 3)), 0.2)) <= 1e-5


def test_spectral_norm_fold_matches_eval_mode_rule():
    from styletts2_b200 import ops
    g = torch.Generator().manual_seed(5)
    for shape in [(64, 1, 3, 3), (128, 64, 3, 3), (96, 1, 3, 3), (512, 512, 5, 5)]:
        w = torch.randn(shape, generator=g)
        u = F.normalize(torch.randn(shape[0], generator=g), dim=0)
        v = F.normalize(torch.randn(w[0].numel(), generator=g), dim=0)
        sd = {"c.weight_orig": w, "c.weight_u": u, "c.weight_v": v}
        want = SO.sn_weight(sd, "c")
        wt, sigma = ops.spectral_norm_fold(w.cuda(), u.cuda(), v.cuda())
        want_sigma = float(torch.dot(u.double(), w.flatten(1).double() @ v.double()))
        assert abs(float(sigma) - want_sigma) <= 2e-6 * abs(want_sigma) + 1e-7
        assert _rel(wt.t().reshape(shape), want) <= 1e-5     # torch's fp32 sigma differs from the fp64-accumulated one by ~1e-6


def test_logmel_matches_torchaudio_fixture():
    from styletts2_b200.style import LogMel
    g = golden("style_libri")
    lm = LogMel().cuda()
    mel = lm(torch.from_numpy(g["wave"]).cuda())
    d = maxdiff(mel, torch.from_numpy(g["mel"]))
    record("logmel_vs_torchaudio", max_abs=d)
    assert mel.shape == g["mel"].shape and d <= 1e-4      # normalised log-mel units (range ~0..3.3)
    # and against the oracle on a fresh clip with a different length (even frame count)
    from styletts2_b200.synthetic import synthetic_wave
    w2 = synthetic_wave(3, 30300, seed=3)
    d2 = maxdiff(lm(w2.cuda()), SO.log_mel(w2))
    assert d2 <= 1e-4


def test_compute_style_matches_reference_fixture():
    from styletts2_b200.models import STYLE_MODULES, load_keyed_weights
    from styletts2_b200.style import LogMel, compute_style
    g = golden("style_libri")
    m = gpu_model("libritts")
    load_keyed_weights(m, modules=STYLE_MODULES)
    ref = torch.from_numpy(g["ref_s"])
    got = compute_style(m, torch.from_numpy(g["wave"]).cuda(), LogMel().cuda())
    # the encoder alone, fed the reference's own mel (isolates the conv stack from the front-end)
    mel = torch.from_numpy(g["mel"]).cuda().unsqueeze(1)
    enc = torch.cat([m.style_encoder(mel), m.predictor_encoder(mel)], dim=1)
    d, d_enc, scale = maxdiff(got, ref), maxdiff(enc, ref), float(ref.abs().max())
    record("compute_style_vs_reference", max_abs=d, encoder_only_max_abs=d_enc, ref_absmax=scale)
    assert got.shape == (2, 256)
    assert d_enc <= 2e-5 * scale and d <= 1e-4 * scale       # fp32: relative to the style vector's peak (|ref_s| ~ 0.07)


def test_style_to_waveform_chain_runs():
    """compute_style output feeds the multispeaker synthesizer as `ref_s` (notebook cells 5 -> inference)."""
    import cases
    from styletts2_b200.inference import Synthesizer
    from styletts2_b200.models import STYLE_MODULES, load_keyed_weights
    from styletts2_b200.style import compute_style
    from styletts2_b200.synthetic import synthetic_batch, synthetic_wave
    m = gpu_model("libritts")
    load_keyed_weights(m, modules=STYLE_MODULES)
    ref_s = compute_style(m, synthetic_wave(2, 36000, seed=5).cuda())
    syn = Synthesizer(m, cases.MODEL_CFGS["libritts"])
    tokens, lengths, bert_dur, noise, _ = synthetic_batch(2, 24, multispeaker=True)
    out = syn.synthesize(tokens.cuda(), lengths.cuda(), bert_dur.cuda(), noise.cuda(), ref_s=ref_s, pin_frames_per_token=4)
    wav = out["wav"] if isinstance(out, dict) else out
    assert torch.isfinite(wav).all() and wav.shape[0] == 2
