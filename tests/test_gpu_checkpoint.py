"""GPU tier, row f4: folded-weight safetensors export / import drives the same bits as the original parameters, and the
device PCM writer matches numpy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cases
from util import gpu_model

D = "cuda:0"


@pytest.mark.parametrize("folded", [True, False])
def test_model_safetensors_round_trip_is_bit_exact(tmp_path, folded):
    from styletts2_b200.checkpoint import load_model, load_safetensors, save_model
    from styletts2_b200.models import build_model, recursive_munch
    from styletts2_b200.modules import WNConv1d, WNConvTranspose1d
    src = gpu_model("ljspeech")
    p = str(tmp_path / "m.safetensors")
    n = save_model(src, p, folded=folded)
    tensors, meta = load_safetensors(p)
    assert n == len(tensors) and meta["schema"] == ("folded" if folded else "reference")
    if folded:
        assert not any(k.endswith("weight_g") for k in tensors)      # (spectral-norm `weight_v` buffers of the style encoders stay)
    dst = build_model(recursive_munch(cases.MODEL_CFGS["ljspeech"]))
    for k in dst:
        dst[k].to(D).eval()
    load_model(dst, p)
    pairs = [(a, b) for (na, a), (nb, b) in zip(src.decoder.named_modules(), dst.decoder.named_modules())
             if isinstance(a, (WNConv1d, WNConvTranspose1d))]
    assert len(pairs) > 50
    for a, b in pairs:
        assert torch.equal(a.folded(), b.folded())          # the tensors the kernels consume are identical bit for bit
    x = torch.randn(2, 512, 40, device=D, generator=torch.Generator(device=D).manual_seed(1))
    f0 = torch.rand(2, 80, device=D) * 200 + 80
    nn_ = torch.rand(2, 80, device=D)
    s = torch.randn(2, 128, device=D) * 0.3
    sn = torch.randn(2, 24000, 9, device=D)
    with torch.no_grad():
        assert torch.equal(src.decoder(x, f0, nn_, s, sine_noise=sn), dst.decoder(x, f0, nn_, s, sine_noise=sn))


def test_pcm16_kernel_matches_numpy_and_saturates():
    from styletts2_b200.checkpoint import pcm16
    g = torch.Generator().manual_seed(3)
    for n in (1, 7, 4096, 100003):
        x = torch.randn(n, generator=g) * 0.6
        x[: min(n, 3)] = torch.tensor([1.5, -1.5, 0.5 / 32767])[: min(n, 3)]
        want = np.clip(np.rint(x.numpy().astype(np.float32) * np.float32(32767.0)), -32768, 32767).astype(np.int16)
        got = pcm16(x.to(D)).cpu().numpy()
        assert got.dtype == np.int16 and np.array_equal(got, want), n
    y = pcm16(torch.full((5,), 0.25, device=D), gain=2.0).cpu().numpy()
    assert np.array_equal(y, np.full(5, int(np.rint(0.25 * 2 * 32767)), dtype=np.int16))
