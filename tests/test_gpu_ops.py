"""GPU tier, per-kernel parity: every C-ABI entry point against plain fp32 torch (CPU) arithmetic
of the same op.  Tolerances are written next to each check (fp32 paths: 1e-4 relative to the
output scale unless the op is bit-exact by construction)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import styletts2_oracle as O
from util import maxdiff


def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def rel(a, b):
    b = b.detach().cpu().float()
    return maxdiff(a, b) / max(1e-6, float(b.abs().max()))


CONV_CASES = [
    # Cin, Cout, K, stride, dil, pad, L
    (22, 256, 12, 6, 1, 3, 2401),
    (1, 64, 60, 30, 1, 15, 6000),
    (128, 22, 7, 1, 1, 3, 777),
    (514, 96, 3, 1, 1, 1, 130),
    (32, 1, 7, 1, 1, 3, 1000),
    (1, 1, 3, 2, 1, 1, 64),
    (64, 64, 11, 1, 5, 25, 600),
    (130, 70, 7, 1, 3, 9, 300),
    (22, 128, 1, 1, 1, 0, 2401),
]


@pytest.mark.parametrize("cfg", CONV_CASES)
def test_conv1d_plain(cfg):
    from styletts2_b200 import ops
    Cin, Cout, K, stride, dil, pad, L = cfg
    x, w, b = rnd(2, Cin, L, seed=1), rnd(Cout, Cin, K, seed=2, scale=1 / math.sqrt(Cin * K)), rnd(Cout, seed=3)
    ref = F.conv1d(x, w, b, stride, pad, dil)
    wt = ops.conv_weight_layout(w.to(dev()))
    y, _ = ops.conv1d(x.to(dev()), wt, b.to(dev()), K=K, stride=stride, dil=dil, pad=pad)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert rel(y, ref) < 1e-5, rel(y, ref)


@pytest.mark.parametrize("act", ["snake", "lrelu"])
def test_conv1d_fused_prologue_epilogue_stats(act):
    """AdaIN affine + Snake/LeakyReLU prologue, bias + residual epilogue, InstanceNorm partial stats."""
    from styletts2_b200 import ops
    from styletts2_b200.lib import ACT_LRELU, ACT_SNAKE
    B, C, K, d, L = 2, 96, 7, 3, 1000
    x, w, bias = rnd(B, C, L, seed=1), rnd(C, C, K, seed=2, scale=1 / math.sqrt(C * K)), rnd(C, seed=3)
    a, b = 1 + 0.3 * rnd(B, C, seed=4), 0.2 * rnd(B, C, seed=5)
    alpha = 1 + 0.3 * torch.rand(1, C, 1, generator=torch.Generator().manual_seed(6))
    res = rnd(B, C, L, seed=7)
    z = a[:, :, None] * x + b[:, :, None]
    z = z + (1 / alpha) * torch.sin(alpha * z) ** 2 if act == "snake" else F.leaky_relu(z, 0.2)
    ref = F.conv1d(z, w, bias, 1, O.get_padding(K, d), d) + res
    D = dev()
    y, st = ops.conv1d(x.to(D), ops.conv_weight_layout(w.to(D)), bias.to(D), K=K, dil=d, pad=O.get_padding(K, d),
                       pre=(a.to(D).contiguous(), b.to(D).contiguous()), pre_act=ACT_SNAKE if act == "snake" else ACT_LRELU,
                       slope=0.2, alpha=alpha.to(D), res=res.to(D), want_stats=True)
    assert rel(y, ref) < 2e-5, rel(y, ref)
    # stats -> AdaIN coefficients against F.instance_norm semantics
    gb = rnd(B, 2 * C, seed=8).to(D)
    ca, cb = ops.adain_coef(st, gb)
    mean, var = ref.mean(-1), ref.var(-1, unbiased=False)
    ea = (1 + gb.cpu()[:, :C]) / torch.sqrt(var + 1e-5)
    eb = gb.cpu()[:, C:] - mean * ea
    assert rel(ca, ea) < 1e-5 and maxdiff(cb, eb) < 1e-4 * max(1.0, float(eb.abs().max())), (rel(ca, ea), maxdiff(cb, eb))


def test_conv1d_mrf_accumulate_and_div():
    from styletts2_b200 import ops
    B, C, K, L = 2, 64, 3, 520
    D = dev()
    x = rnd(B, C, L, seed=1)
    ws = [rnd(C, C, K, seed=10 + i, scale=0.1) for i in range(3)]
    ref = (F.conv1d(x, ws[0], None, 1, 1) + F.conv1d(x, ws[1], None, 1, 1) + F.conv1d(x, ws[2], None, 1, 1)) / 3
    acc = torch.empty(B, C, L, device=D)
    for i in range(3):
        ops.conv1d(x.to(D), ops.conv_weight_layout(ws[i].to(D)), None, K=K, pad=1, out=acc,
                   accum_mode=0 if i == 0 else (2 if i == 2 else 1), accum_div=3.0)
    assert rel(acc, ref) < 1e-5
    # (res + sc)/sqrt(2) with nearest x2 shortcut
    sc = rnd(B, C, L // 2, seed=20)
    ref2 = (F.conv1d(x, ws[0], None, 1, 1) + F.interpolate(sc, scale_factor=2, mode="nearest")) / math.sqrt(2)
    y, _ = ops.conv1d(x.to(D), ops.conv_weight_layout(ws[0].to(D)), None, K=K, pad=1, res=sc.to(D), res_shift=1, out_div=math.sqrt(2))
    assert rel(y, ref2) < 1e-5


CONVT_CASES = [  # Cin, Cout, K, S, P, OP, L, reflect
    (64, 32, 20, 10, 5, 0, 50, False),
    (48, 24, 12, 6, 3, 0, 333, True),
    (32, 16, 10, 5, 3, 1, 100, False),
    (16, 8, 6, 3, 2, 1, 300, False),
    (8, 8, 4, 2, 1, 0, 600, False),
    (512, 256, 20, 10, 5, 0, 40, False),
]


@pytest.mark.parametrize("cfg", CONVT_CASES)
def test_conv_transpose1d_polyphase(cfg):
    from styletts2_b200 import ops
    from styletts2_b200.lib import ACT_LRELU
    Cin, Cout, K, S, P, OP, L, reflect = cfg
    D = dev()
    x, w, b = rnd(2, Cin, L, seed=1), rnd(Cin, Cout, K, seed=2, scale=1 / math.sqrt(Cin * 2)), rnd(Cout, seed=3)
    ref = F.conv_transpose1d(F.leaky_relu(x, 0.1), w, b, stride=S, padding=P, output_padding=OP)
    assert ref.shape[-1] == L * S
    if reflect:
        ref = F.pad(ref, (1, 0), mode="reflect")
    res = rnd(2, Cout, ref.shape[-1], seed=4)
    ref = ref + res
    wp = ops.convT_weight_layout(w.to(D), S, P)
    y, st = ops.conv_transpose1d(x.to(D), wp, b.to(D), K=K, stride=S, padding=P, pre_act=ACT_LRELU, slope=0.1, res=res.to(D),
                                 reflect_left1=reflect, want_stats=True)
    assert y.shape == ref.shape
    assert rel(y, ref) < 1e-5, rel(y, ref)
    gb = torch.zeros(2, 2 * Cout, device=D)
    ca, cb = ops.adain_coef(st, gb)
    ea = 1 / torch.sqrt(ref.var(-1, unbiased=False) + 1e-5)
    assert rel(ca, ea) < 1e-5 and maxdiff(cb, -ref.mean(-1) * ea) < 1e-4


def test_weight_norm_fold_matches_torch():
    from styletts2_b200 import ops
    v, g = rnd(70, 33, 5, seed=1), 0.5 + torch.rand(70, 1, 1, generator=torch.Generator().manual_seed(2))
    ref = torch._weight_norm(v, g, 0)
    w = ops.fold_weight_norm(v.to(dev()), g.to(dev()))
    assert rel(w, ref) < 1e-6


def test_instance_stats_adain_matches_instance_norm():
    from styletts2_b200 import ops
    B, C, L = 3, 37, 1234
    x = rnd(B, C, L, seed=1) * 3 + 5
    gb = rnd(B, 2 * C, seed=2)
    ref = (1 + gb[:, :C, None]) * F.instance_norm(x, eps=1e-5) + gb[:, C:, None]
    D = dev()
    a, b = ops.adain_coef(ops.instance_stats(x.to(D)), gb.to(D))
    y = a[:, :, None] * x.to(D) + b[:, :, None]
    assert maxdiff(y, ref) < 2e-5 * float(ref.abs().max())


def test_adain_lrelu_pool_matches_depthwise_convtranspose():
    from styletts2_b200 import ops
    B, C, L = 2, 45, 77
    D = dev()
    x, a, b = rnd(B, C, L, seed=1), 1 + 0.2 * rnd(B, C, seed=2), 0.1 * rnd(B, C, seed=3)
    pw, pb = rnd(C, 1, 3, seed=4), rnd(C, seed=5)
    z = F.leaky_relu(a[:, :, None] * x + b[:, :, None], 0.2)
    ref = F.conv_transpose1d(z, pw, pb, stride=2, padding=1, output_padding=1, groups=C)
    y = ops.adain_lrelu_pool(x.to(D), a.to(D), b.to(D), pw.to(D).view(C, 3).contiguous(), pb.to(D))
    assert rel(y, ref) < 1e-5


def test_channel_layernorm_lrelu_and_mask():
    from styletts2_b200 import ops
    B, C, L = 2, 512, 50
    D = dev()
    x, g, b = rnd(B, C, L, seed=1), 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    lens = torch.tensor([50, 31], dtype=torch.int32)
    ref = F.leaky_relu(F.layer_norm(x.transpose(1, 2), (C,), g, b, 1e-5).transpose(1, 2), 0.2)
    ref[1, :, 31:] = 0
    y = ops.channel_layernorm_lrelu(x.to(D), g.to(D), b.to(D), lens.to(D))
    assert rel(y, ref) < 1e-5


def test_rows_ln_plain_and_ada():
    from styletts2_b200 import ops
    B, N, C = 3, 17, 1024
    D = dev()
    x, emb, add = rnd(B, 256, seed=1), rnd(B, N, 768, seed=2), rnd(B, C, seed=3)
    g1, b1, g2, b2 = rnd(C, seed=4), rnd(C, seed=5), rnd(C, seed=6), rnd(C, seed=7)
    h_ref = torch.cat([(0.7 * x)[:, None].expand(-1, N, -1), emb], -1) + add[:, None]
    a_ref, c_ref = F.layer_norm(h_ref, (C,), g1, b1), F.layer_norm(h_ref, (C,), g2, b2)
    h, a, c = (torch.empty(B * N, C, device=D) for _ in range(3))
    ops.rows_ln(B=B, N=N, Cw=C, x=x.to(D), xs=0.7, emb=emb.to(D), add=add.to(D), h_out=h, g1=g1.to(D), b1=b1.to(D), g2=g2.to(D),
                b2=b2.to(D), out1=a, out2=c)
    assert rel(h.view(B, N, C), h_ref) < 1e-6 and rel(a.view(B, N, C), a_ref) < 1e-5 and rel(c.view(B, N, C), c_ref) < 1e-5
    # AdaLayerNorm flavour, in-place h, second pass
    gb = rnd(B, 2 * C, seed=8)
    ada_ref = (1 + gb[:, None, :C]) * F.layer_norm(h_ref + add[:, None], (C,)) + gb[:, None, C:]
    gbd = gb.to(D)
    ops.rows_ln(B=B, N=N, Cw=C, h_in=h, add=add.to(D), h_out=h, g1=gbd, b1=gbd[:, C:], gb_bstride=2 * C, ada=True, out1=a)
    assert rel(a.view(B, N, C), ada_ref) < 1e-5


@pytest.mark.parametrize("shape", [(33, 257, 1024), (300, 1024, 512), (64, 640, 2048), (5, 128, 50), (32, 1024, 1024), (1, 768, 7),
                                   (40, 256, 129), (65, 512, 96)])
def test_linear_bias_gelu_residual(shape):
    from styletts2_b200 import ops
    from styletts2_b200.lib import ACT_GELU
    M, K, Nf = shape
    D = dev()
    A, W, b, R = rnd(M, K, seed=1), rnd(Nf, K, seed=2, scale=1 / math.sqrt(K)), rnd(Nf, seed=3), rnd(M, Nf, seed=4)
    ref = F.gelu(F.linear(A, W, b)) + R
    y = ops.linear(A.to(D), W.to(D), b.to(D), act=ACT_GELU, R=R.to(D))
    assert rel(y, ref) < 1e-5, rel(y, ref)


@pytest.mark.parametrize("shape", [(8, 256, 2048), (1, 1024, 1024), (5, 257, 1000), (8, 1024, 130), (3, 64, 18)])
def test_linear_tiny_m(shape):
    """M <= 8 rows (time / feature mapping MLP of the denoiser at 8 utterances per GPU): warp-per-feature-pair kernel when
    the operands are 16-byte aligned (K % 4 == 0), the 32-row small-M kernel otherwise; odd feature counts, tanh, residual,
    strided output rows."""
    from styletts2_b200 import ops
    from styletts2_b200.lib import ACT_TANH
    M, K, Nf = shape
    D = dev()
    A, W, b, R = rnd(M, K, seed=1), rnd(Nf, K, seed=2, scale=1 / math.sqrt(K)), rnd(Nf, seed=3), rnd(M, Nf, seed=4)
    ref = torch.tanh(F.linear(A.double(), W.double(), b.double())).float() + R
    out = torch.zeros(M, Nf + 8, device=D)
    ops.linear(A.to(D), W.to(D), b.to(D), act=ACT_TANH, R=R.to(D), out=out[:, 8:])
    assert rel(out[:, 8:], ref) < 1e-5, rel(out[:, 8:], ref)
    assert float(out[:, :8].abs().max()) == 0.0


def test_linear_small_m_strided_rows():
    """small-M kernel on a column slice of a wider buffer (row stride > K) and writing into a slice"""
    from styletts2_b200 import ops
    D = dev()
    M, K, Nf = 32, 128, 2048
    big = rnd(M, 3 * K, seed=1)
    W, b = rnd(Nf, K, seed=2, scale=1 / math.sqrt(K)), rnd(Nf, seed=3)
    A = big.to(D)[:, K:2 * K]
    out = torch.zeros(M, Nf + 64, device=D)
    ops.linear(A, W.to(D), b.to(D), out=out[:, 64:])
    ref = F.linear(big[:, K:2 * K], W, b)
    assert rel(out[:, 64:], ref) < 1e-5 and float(out[:, :64].abs().max()) == 0.0


def test_linear_conv_layout_input():
    from styletts2_b200 import ops
    B, K, Lr, Nf = 3, 70, 45, 96
    D = dev()
    x, W = rnd(B, K, Lr, seed=1), rnd(Nf, K, seed=2)
    ref = F.linear(x.transpose(1, 2), W).reshape(B * Lr, Nf)
    y = ops.linear_strided(x.to(D), B, Lr, K, K * Lr, 1, Lr, W.to(D))
    assert rel(y, ref) < 1e-5


@pytest.mark.parametrize("N", [10, 128, 200])
def test_attention(N):
    from styletts2_b200 import ops
    B, H, Dh = 2, 8, 64
    D = dev()
    q, kv = rnd(B * N, H * Dh, seed=1), rnd(B * N, 2 * H * Dh, seed=2)
    qh = q.view(B, N, H, Dh).permute(0, 2, 1, 3)
    k, v = kv[:, :H * Dh].view(B, N, H, Dh).permute(0, 2, 1, 3), kv[:, H * Dh:].view(B, N, H, Dh).permute(0, 2, 1, 3)
    att = torch.softmax(qh @ k.transpose(-1, -2) * Dh ** -0.5, -1) @ v
    ref = att.permute(0, 2, 1, 3).reshape(B * N, H * Dh)
    y = ops.attention(q.to(D), kv.to(D), B, N, H, Dh)
    assert rel(y, ref) < 1e-5, rel(y, ref)


@pytest.mark.parametrize("ragged", [False, True])
def test_lstm_bidir(ragged):
    from styletts2_b200.modules import LSTM
    B, Lr, In, H = 3, 23, 40, 32
    m = LSTM(In, H).to(dev())
    sd = {"l." + k: v.detach().cpu() for k, v in m.state_dict().items()}
    x = rnd(B, Lr, In, seed=1)
    lens = torch.tensor([23, 9, 17]) if ragged else None
    ref = O.bilstm(x, sd, "l", lens)
    y, _ = m(x.to(dev()), None if lens is None else lens.to(dev(), torch.int32))
    assert rel(y[:, :ref.shape[1]], ref) < 2e-5, rel(y[:, :ref.shape[1]], ref)
    if ragged:
        assert float(y[1, 9:].abs().max()) == 0.0


@pytest.mark.parametrize("B", [1, 3, 8, 13, 32, 37, 50, 70])
def test_lstm_cluster_kernel_h256(B):
    """H=256 (the size every LSTM of the reference has) takes the 8-CTA cluster / DSMEM kernel: check it against the
    oracle biLSTM and against the cooperative-launch kernel on the same inputs, ragged lengths included."""
    from styletts2_b200 import lib
    from styletts2_b200.modules import LSTM
    Lr, In, H = 41, 48, 256
    m = LSTM(In, H).to(dev())
    sd = {"l." + k: v.detach().cpu() for k, v in m.state_dict().items()}
    x = rnd(B, Lr, In, seed=B)
    g = torch.Generator().manual_seed(B)
    lens = torch.randint(1, Lr + 1, (B,), generator=g)
    lens[0] = Lr
    ref = O.bilstm(x, sd, "l", lens)
    li = lens.to(dev(), torch.int32)
    y, _ = m(x.to(dev()), li)
    lib.call("st2_debug_lstm_cluster", 0)
    try:
        y_coop, _ = m(x.to(dev()), li)
    finally:
        lib.call("st2_debug_lstm_cluster", 1)
    assert rel(y[:, :ref.shape[1]], ref) < 2e-5, rel(y[:, :ref.shape[1]], ref)
    assert rel(y, y_coop) < 1e-5
    for b in range(B):
        if int(lens[b]) < Lr:
            assert float(y[b, int(lens[b]):].abs().max()) == 0.0
    y2, _ = m(x.to(dev()), None)                      # unpadded path
    assert rel(y2, O.bilstm(x, sd, "l", None)) < 2e-5


def test_sine_source_matches_reference_arithmetic():
    """fp64 phase accumulation + PyTorch interpolation rule: |diff| <= 1e-6 on a tanh-bounded signal even
    though the instantaneous phase is ~1e4..1e5 rad."""
    from styletts2_b200 import ops
    from styletts2_b200.synthetic import synthetic_f0
    B, Fr, scale = 2, 64, 300
    f0 = synthetic_f0(B, Fr, seed=3)
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(B, Fr * scale, 9, generator=g)
    sd = {"m_source.l_linear.weight": rnd(1, 9, seed=6) * 0.3, "m_source.l_linear.bias": rnd(1, seed=7) * 0.1}
    ref = O.sine_source(f0, scale, sd, torch.zeros(B, 9), noise)
    D = dev()
    y = ops.sine_source(f0.to(D), scale, noise.to(D), sd["m_source.l_linear.weight"].view(-1).to(D), sd["m_source.l_linear.bias"].to(D))
    assert maxdiff(y, ref) < 2e-6, maxdiff(y, ref)


def test_stft20_and_istft20():
    from styletts2_b200 import ops
    B, L = 2, 3000
    x = rnd(B, L, seed=1) * 0.1
    mag, ph = O.stft_mag_phase(x)
    D = dev()
    har = ops.stft20(x.to(D)).cpu()
    assert maxdiff(har[:, :11], mag) < 1e-6
    # phase compared as a unit phasor weighted by magnitude (angle is ill-conditioned at |X| ~ 0), and
    # directly where |X| is well above rounding noise (interior frames; edge frames are +-pi by noise)
    dphi = torch.remainder(har[:, 11:] - ph + math.pi, 2 * math.pi) - math.pi
    assert float((dphi.abs() * mag).max()) < 1e-6
    good = mag[:, :, 1:-1] > 1e-3
    assert float(dphi[:, :, 1:-1][good].abs().max()) < 1e-4
    # inverse: conv_post tail (exp / sin) + istft
    z = rnd(B, 22, L // 5 + 1, seed=2) * 0.5
    ref = O.istft_from_mag_phase(torch.exp(z[:, :11]), torch.sin(z[:, 11:])).squeeze(1)
    wav = ops.istft20_expsin(z.to(D))
    assert wav.shape == ref.shape and rel(wav, ref) < 1e-5, rel(wav, ref)


def test_durations_alignment_gather():
    from styletts2_b200 import ops
    B, N, J = 3, 12, 50
    D = dev()
    logits = rnd(B, N, J, seed=1) * 3 - 2
    ref = O.predict_durations(logits, 5)
    pred, durf = ops.durations(logits.to(D), 5)
    assert torch.equal(pred.cpu().float(), ref)
    dur = torch.randint(1, 7, (B, N), generator=torch.Generator().manual_seed(2))
    dur[:, -1] += (dur.sum(1).max() - dur.sum(1))
    T = int(dur.sum(1)[0])
    alns = torch.stack([O.alignment_from_durations(dur[b].float()) for b in range(B)])
    d, t_en = rnd(B, N, 40, seed=3), rnd(B, 24, N, seed=4)
    for shift in (False, True):
        tok, total = ops.frame_tokens(dur.to(D, torch.int32), T, shift_right=shift)
        en_ref, asr_ref = d.transpose(1, 2) @ alns, t_en @ alns
        if shift:
            en_ref, asr_ref = O.shift_right_one(en_ref), O.shift_right_one(asr_ref)
        en = ops.expand_rows(d.to(D), tok)
        asr = ops.expand_cl(t_en.to(D), tok)
        assert torch.equal(en.cpu().transpose(1, 2), en_ref) and torch.equal(asr.cpu(), asr_ref)
        assert torch.equal(total.cpu(), dur.sum(1).int())


def test_kdiff_step_and_small_ops():
    from styletts2_b200 import ops
    D = dev()
    xe, xp, xm, xb, eps = (rnd(4, 256, seed=i) for i in range(5))
    c_skip, c_out, sig, dt, up, cfg = 0.3, 0.7, 0.55, -0.2, 0.05, 1.5
    p = xm + (xp - xm) * cfg
    den = c_skip * xe + c_out * p
    ref = xb + ((xe - den) / sig) * dt + eps * up
    y = ops.kdiff_step(xe.to(D), xp.to(D), c_skip, c_out, sig, xb.to(D), dt, eps=eps.to(D), sigma_up=up, x_pred_masked=xm.to(D), cfg_scale=cfg)
    assert maxdiff(y, ref) < 1e-6
    t, w = torch.tensor([0.3, -1.2]), rnd(128, seed=9)
    fr = t[:, None] * w[None] * 2 * math.pi
    ref = torch.cat([t[:, None], fr.sin(), fr.cos()], -1)
    assert maxdiff(ops.time_embedding(t.to(D), w.to(D)), ref) < 1e-6
    tokens = torch.randint(0, 178, (2, 9), generator=torch.Generator().manual_seed(1))
    table = rnd(178, 16, seed=2)
    e = ops.embedding_cl(tokens.to(D), table.to(D), torch.tensor([9, 5], dtype=torch.int32, device=D))
    ref = F.embedding(tokens, table).transpose(1, 2).clone()
    ref[1, :, 5:] = 0
    assert torch.equal(e.cpu(), ref)
    h = rnd(3 * 7, 64, seed=3)
    assert maxdiff(ops.mean_rows(h.to(D), 3, 7), h.view(3, 7, 64).mean(1)) < 1e-6


@pytest.mark.gpu
def test_philox_randn_statistics_and_epoch():
    """st2_randn: N(0,1) moments, stream reproducibility, and a fresh draw after st2_rng_advance."""
    from styletts2_b200 import ops
    ops.manual_seed(1234)
    x = torch.empty(1 << 22, device="cuda")
    a = ops.randn_like(x)
    assert abs(a.mean().item()) < 3e-3 and abs(a.std().item() - 1.0) < 3e-3
    assert abs((a ** 3).mean().item()) < 1e-2 and abs((a ** 4).mean().item() - 3.0) < 3e-2
    # lag-1 / lag-4 correlation of a counter-based stream
    assert abs((a[1:] * a[:-1]).mean().item()) < 3e-3 and abs((a[4:] * a[:-4]).mean().item()) < 3e-3
    ops.manual_seed(1234)
    b = ops.randn_like(x)
    assert torch.equal(a, b)
    ops.manual_seed(1234)
    ops.rng_advance(x.device)
    c = ops.randn_like(x)
    assert not torch.equal(a, c) and abs((a * c).mean().item()) < 3e-3


@pytest.mark.gpu
def test_sine_source_in_kernel_noise_matches_injected_statistics():
    """noise=None draws the 9 normals in the kernel: unvoiced frames give tanh(w . (0.1/3 * n) + b)."""
    from styletts2_b200 import ops
    ops.manual_seed(7)
    B, F, scale = 2, 64, 300
    f0 = torch.zeros(B, F, device="cuda")          # all unvoiced -> pure noise branch
    w = torch.full((9,), 0.5, device="cuda")
    bias = torch.zeros(1, device="cuda")
    y = ops.sine_source(f0, scale, None, w, bias)
    pre = torch.atanh(y.clamp(-0.999999, 0.999999))
    want_std = (0.1 / 3.0) * 0.5 * 3.0               # sqrt(9) * 0.5 * 0.1/3
    assert abs(pre.mean().item()) < 1e-3 and abs(pre.std().item() / want_std - 1.0) < 2e-2
    y2 = ops.sine_source(f0, scale, None, w, bias)   # offset advanced -> different draw
    assert not torch.equal(y, y2)
