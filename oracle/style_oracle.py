"""CPU oracle for SURVEY.md section 8 row f2: the zero-shot reference-style path
(`compute_style`): waveform -> log-mel front-end -> StyleEncoder x2 -> ref_s [B,256].

TEST INFRASTRUCTURE -- NOT PRODUCT CODE (same rules as oracle/styletts2_oracle.py: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline legs may import it).

Functional restatement on a flat {key: tensor} state dict with the reference's key names:
  * StyleEncoder / ResBlk / DownSample / LearnedDownSample     models.py:27-164
  * spectral_norm in eval mode (torch.nn.utils.spectral_norm): weight = weight_orig / (u . (W_mat v)),
    no power iteration outside training
  * preprocess / compute_style                                  Demo/Inference_LibriTTS.ipynb cell 5
    (torchaudio.transforms.MelSpectrogram(n_mels=80, n_fft=2048, win_length=1200, hop_length=300);
    note the reference leaves sample_rate at torchaudio's default 16000, so the HTK filterbank
    spans 0..8000 Hz over the 1025 bins -- restated faithfully)
torchaudio is a third-party dependency (present in this image, 2.11.0): its published algorithm
(power spectrogram of a centered reflect-padded STFT with a periodic Hann window zero-padded to n_fft;
melscale_fbanks with the HTK scale and norm=None) is restated below and pinned against torchaudio itself
and against the unmodified reference StyleEncoder by oracle/make_golden_style.py.
librosa.load / librosa.effects.trim (file I/O and silence trimming on the host) are outside the path:
the boundary input is the trimmed 24 kHz waveform.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

N_FFT, WIN, HOP, N_MELS = 2048, 1200, 300, 80
MEL_SAMPLE_RATE = 16000          # torchaudio default kept by the reference (notebook cell 5)
LOG_EPS, MEL_MEAN, MEL_STD = 1e-5, -4.0, 4.0


def sub(sd: SD, prefix: str) -> SD:
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


# --------------------------------------------------------------------------- mel front-end
def mel_filterbank(n_freqs=N_FFT // 2 + 1, n_mels=N_MELS, sample_rate=MEL_SAMPLE_RATE, f_min=0.0, f_max=None):
    """HTK triangular filters, norm=None: [n_freqs, n_mels] (torchaudio.functional.melscale_fbanks)."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def power_spectrogram(wave):
    """wave [B, L] -> |STFT|^2 [B, 1025, 1 + L//300]  (center=True, reflect pad, periodic Hann(1200) centred in 2048)."""
    win = torch.hann_window(WIN, periodic=True)
    spec = torch.stft(wave, N_FFT, hop_length=HOP, win_length=WIN, window=win, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    return spec.real ** 2 + spec.imag ** 2


def log_mel(wave):
    """preprocess() of the notebooks: [B, L] -> normalised log-mel [B, 80, frames]."""
    mel = torch.matmul(power_spectrogram(wave).transpose(-1, -2), mel_filterbank()).transpose(-1, -2)
    return (torch.log(LOG_EPS + mel) - MEL_MEAN) / MEL_STD


# --------------------------------------------------------------------------- StyleEncoder
def sn_weight(sd: SD, p: str) -> torch.Tensor:
    """Eval-mode spectral norm: W / sigma, sigma = u^T W_mat v with the stored u, v."""
    w = sd[p + ".weight_orig"]
    sigma = torch.dot(sd[p + ".weight_u"], torch.mv(w.flatten(1), sd[p + ".weight_v"]))
    return w / sigma


def _conv2d(x, sd, p, stride=1, padding=0, groups=1):
    return F.conv2d(x, sn_weight(sd, p), sd.get(p + ".bias"), stride=stride, padding=padding, groups=groups)


def downsample_half(x):
    """DownSample('half') (models.py:73-78): replicate the last column when the width is odd, then 2x2 mean."""
    if x.shape[-1] % 2 != 0:
        x = torch.cat([x, x[..., -1:]], dim=-1)
    return F.avg_pool2d(x, 2)


def resblk(x, sd: SD, p: str):
    """ResBlk(normalize=False, downsample='half') (models.py:96-137)."""
    sc = x
    if (p + ".conv1x1.weight_orig") in sd:
        sc = _conv2d(sc, sd, p + ".conv1x1")
    sc = downsample_half(sc)
    r = F.leaky_relu(x, 0.2)
    r = _conv2d(r, sd, p + ".conv1", padding=1)
    r = _conv2d(r, sd, p + ".downsample_res.conv", stride=2, padding=1, groups=r.shape[1])
    r = F.leaky_relu(r, 0.2)
    r = _conv2d(r, sd, p + ".conv2", padding=1)
    return (sc + r) / math.sqrt(2)


def style_encoder(mel, sd: SD):
    """StyleEncoder.forward (models.py:139-164): mel [B,1,80,F] -> [B, style_dim]."""
    h = _conv2d(mel, sd, "shared.0", padding=1)
    for i in range(1, 5):
        h = resblk(h, sd, f"shared.{i}")
    h = F.leaky_relu(h, 0.2)
    h = _conv2d(h, sd, "shared.6")                      # 5x5, no padding
    h = F.leaky_relu(h.mean(dim=(2, 3)), 0.2)           # AdaptiveAvgPool2d(1) + LeakyReLU
    return F.linear(h, sd["unshared.weight"], sd["unshared.bias"])


def compute_style(sds: Dict[str, SD], wave):
    """compute_style of the notebooks after load/trim: wave [B, L] (24 kHz) -> ref_s [B, 256] = [acoustic | prosodic]."""
    mel = log_mel(wave).unsqueeze(1)
    return torch.cat([style_encoder(mel, sds["style_encoder"]), style_encoder(mel, sds["predictor_encoder"])], dim=1)
