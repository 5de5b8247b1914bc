"""Zero-shot reference-style path: `compute_style` of the reference notebooks on
the library's kernels -- log-mel front-end + the two StyleEncoders -> ref_s [B, 256].

Mirrored (same class names, constructor arguments and state-dict keys, so `params['style_encoder']` /
`params['predictor_encoder']` of a StyleTTS 2 checkpoint load as they are):
  models.py  LearnedDownSample :27-46, DownSample :62-79, ResBlk :96-137, StyleEncoder :139-164
  Demo/Inference_LibriTTS.ipynb cell 5  to_mel / preprocess / compute_style
spectral_norm keeps its checkpoint form (weight_orig parameter, weight_u / weight_v buffers) and is folded
once per weight version by st2_spectral_norm_fold (eval-mode semantics: no power iteration).
Boundary: a trimmed 24 kHz waveform already on the device (librosa.load / effects.trim are host file I/O).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .modules import Linear, _Cached


class SNConv2d(_Cached):
    """spectral_norm(nn.Conv2d) holder: weight_orig [Cout, Cin/groups, KH, KW], bias, buffers weight_u, weight_v."""

    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, groups=1, bias=True):
        super().__init__()
        kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
        self.cin, self.cout, self.kh, self.kw, self.stride, self.padding, self.groups = cin, cout, kh, kw, stride, padding, groups
        n = cin // groups * kh * kw
        self.weight_orig = nn.Parameter(torch.empty(cout, cin // groups, kh, kw).uniform_(-1, 1) / math.sqrt(n))
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))
        else:
            self.register_parameter("bias", None)
        self.register_buffer("weight_u", torch.nn.functional.normalize(torch.randn(cout), dim=0))
        self.register_buffer("weight_v", torch.nn.functional.normalize(torch.randn(n), dim=0))

    def _key(self):
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in (self.weight_orig, self.weight_u, self.weight_v))

    def _prepare(self):
        return ops.spectral_norm_fold(self.weight_orig, self.weight_u, self.weight_v)

    def wt(self):
        return self.prepared()[0]

    def sigma(self):
        return self.prepared()[1]

    def forward(self, x, pre_act=False, res=None, out_scale=1.0):
        assert self.groups == 1 and self.stride == 1
        return ops.conv2d(x, self.wt(), self.bias, cout=self.cout, kh=self.kh, kw=self.kw, pad=self.padding, pre_act=pre_act,
                          res=res, out_scale=out_scale)


class LearnedDownSample(nn.Module):
    """models.py:27-46, layer_type 'half' (the only one StyleEncoder builds): depthwise 3x3 stride 2."""

    def __init__(self, layer_type, dim_in):
        super().__init__()
        assert layer_type == "half", "StyleEncoder only builds 'half' downsampling (models.py:147)"
        self.layer_type = layer_type
        self.conv = SNConv2d(dim_in, dim_in, 3, stride=2, padding=1, groups=dim_in)

    def forward(self, x):
        return ops.dwconv3x3_s2(x, self.conv.wt(), self.conv.bias)


class DownSample(nn.Module):
    """models.py:62-79, 'half'"""

    def __init__(self, layer_type):
        super().__init__()
        assert layer_type == "half"
        self.layer_type = layer_type

    def forward(self, x):
        return ops.avgpool_half(x)


class ResBlk(nn.Module):
    """models.py:96-137 with normalize=False, actv=LeakyReLU(0.2), downsample='half'.
    LeakyReLU is applied while the conv stages its input tile; the shortcut add and 1/sqrt(2) ride in conv2's epilogue."""

    def __init__(self, dim_in, dim_out, actv=None, normalize=False, downsample="none"):
        super().__init__()
        assert not normalize and downsample == "half", "StyleEncoder configuration only (models.py:147)"
        self.actv = nn.LeakyReLU(0.2)
        self.normalize = normalize
        self.downsample = DownSample(downsample)
        self.downsample_res = LearnedDownSample(downsample, dim_in)
        self.learned_sc = dim_in != dim_out
        self.conv1 = SNConv2d(dim_in, dim_in, 3, 1, 1)
        self.conv2 = SNConv2d(dim_in, dim_out, 3, 1, 1)
        if self.learned_sc:
            self.conv1x1 = SNConv2d(dim_in, dim_out, 1, 1, 0, bias=False)

    def forward(self, x):
        sc = self.conv1x1(x) if self.learned_sc else x
        sc = self.downsample(sc)
        r = self.conv1(x, pre_act=True)
        r = self.downsample_res(r)
        return self.conv2(r, pre_act=True, res=sc, out_scale=1.0 / math.sqrt(2))


class StyleEncoder(nn.Module):
    """models.py:139-164: forward(mel [B,1,80,F]) -> [B, style_dim]."""

    def __init__(self, dim_in=48, style_dim=48, max_conv_dim=384):
        super().__init__()
        blocks = [SNConv2d(1, dim_in, 3, 1, 1)]
        dim_out = dim_in
        for _ in range(4):
            dim_out = min(dim_in * 2, max_conv_dim)
            blocks.append(ResBlk(dim_in, dim_out, downsample="half"))
            dim_in = dim_out
        blocks += [nn.LeakyReLU(0.2), SNConv2d(dim_out, dim_out, 5, 1, 0), nn.AdaptiveAvgPool2d(1), nn.LeakyReLU(0.2)]
        self.shared = nn.Sequential(*blocks)      # same indices as the reference -> same state-dict keys
        self.unshared = Linear(dim_out, style_dim)

    def forward(self, x):
        assert x.dim() == 4 and x.shape[1] == 1
        assert x.shape[2] >= 80 and x.shape[3] >= 80, "needs >= 80 mel frames (4 halvings then a 5x5 valid conv)"
        h = self.shared[0](x)
        for i in range(1, 5):
            h = self.shared[i](h)
        h = self.shared[6](h, pre_act=True)          # shared[5] LeakyReLU folded into the conv's input staging
        h = ops.mean_hw_lrelu(h, 0.2)                # shared[7] AdaptiveAvgPool2d(1) + shared[8] LeakyReLU
        return self.unshared(h)


class LogMel(nn.Module):
    """to_mel + preprocess of the notebooks (cell 5): torchaudio MelSpectrogram(n_mels=80, n_fft=2048, win_length=1200,
    hop_length=300) with torchaudio's default sample_rate=16000 filterbank (kept as the reference has it), then
    (log(1e-5 + mel) - (-4)) / 4.  The DFT is a GEMM against a [2*1025, 1200] cos/sin basis (only the 1200 window
    taps are non-zero), the filterbank a second GEMM; both run on st2_linear / st2_linear_tc."""

    N_FFT, WIN, HOP, N_MELS, SR = 2048, 1200, 300, 80, 16000
    EPS, MEAN, STD = 1e-5, -4.0, 4.0

    def __init__(self):
        super().__init__()
        nf = self.N_FFT // 2 + 1
        self.register_buffer("window", torch.hann_window(self.WIN, periodic=True), persistent=False)
        k = torch.arange(nf, dtype=torch.int64).unsqueeze(1)
        m = torch.arange(self.WIN, dtype=torch.int64).unsqueeze(0)
        ang = (2.0 * math.pi / self.N_FFT) * ((k * m) % self.N_FFT).double()   # exact argument reduction
        basis = torch.cat([torch.cos(ang), -torch.sin(ang)], dim=0).float()      # rows: re(0..1024), im(0..1024)
        self.dft = Linear(self.WIN, 2 * nf, bias=False)
        self.fb = Linear(nf, self.N_MELS, bias=False)
        with torch.no_grad():
            self.dft.weight.copy_(basis)
            self.fb.weight.copy_(self._melscale_fbanks(nf).t())
        for p in self.parameters():
            p.requires_grad_(False)

    def _melscale_fbanks(self, n_freqs):
        """HTK triangles, norm=None (torchaudio.functional.melscale_fbanks): [n_freqs, n_mels]"""
        all_freqs = torch.linspace(0, self.SR // 2, n_freqs)
        m_max = 2595.0 * math.log10(1.0 + (self.SR // 2) / 700.0)
        m_pts = torch.linspace(0.0, m_max, self.N_MELS + 2)
        f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
        up = slopes[:, 2:] / f_diff[1:]
        return torch.clamp(torch.min(down, up), min=0.0)

    def forward(self, wave):
        """wave [B, L] fp32 on the device -> [B, 80, 1 + L//300]"""
        B = wave.shape[0]
        frames, F = ops.mel_frames(wave, self.window, self.HOP, self.N_FFT)
        y = self.dft(frames)
        p = ops.mel_power(y, self.N_FFT // 2 + 1)
        mel = self.fb(p)
        return ops.logmel(mel, B, F, self.EPS, self.MEAN, self.STD)


@torch.no_grad()
def compute_style(model, wave, to_mel: LogMel = None):
    """compute_style (notebook cell 5) after load/trim: wave [B, L] (24 kHz, device) -> ref_s [B, 256] =
    cat([style_encoder(mel), predictor_encoder(mel)], dim=1)."""
    if to_mel is None:
        to_mel = model.get("to_mel") if isinstance(model, dict) else None
    if to_mel is None:
        to_mel = LogMel().to(wave.device)
    mel = to_mel(wave).unsqueeze(1)
    return torch.cat([model.style_encoder(mel), model.predictor_encoder(mel)], dim=1)
