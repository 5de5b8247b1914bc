"""Diagnostic for the tcgen05 conv kernel: prints error structure for small shapes (GPU only)."""
import math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from styletts2_b200 import ops
D = "cuda:0"
def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale
def run(B, Cin, Cout, K, d, L, bias=True):
    x, w = rnd(B, Cin, L, seed=1), rnd(Cout, Cin, K, seed=2, scale=1 / math.sqrt(Cin * K))
    bs = rnd(Cout, seed=3) if bias else None
    pad = (K * d - d) // 2
    ref = F.conv1d(x, w, bs, 1, pad, d)
    wd = w.to(D)
    y, _ = ops.conv1d(x.to(D), ops.conv_weight_layout(wd), None if bs is None else bs.to(D), K=K, dil=d, pad=pad, wtc=ops.conv_tc_weight_layout(wd, int(os.environ.get("ST2_TC_MODE", "0"))))
    torch.cuda.synchronize()
    y = y.cpu()
    err = (y - ref).abs()
    print(f"case B{B} Cin{Cin} Cout{Cout} K{K} d{d} L{L}: max err {err.max():.3e} ref scale {ref.abs().max():.3f} y scale {y.abs().max():.3f} nan {torch.isnan(y).sum().item()}")
    if err.max() > 1e-3 * ref.abs().max():
        e = err[0]
        print("  err by co block of 16:", [f"{e[i:i+16].max():.2e}" for i in range(0, min(Cout, 128), 16)])
        print("  err by t block of 32 :", [f"{e[:, i:i+32].max():.2e}" for i in range(0, min(L, 256), 32)])
        print("  y[0,0,:8]  ", y[0, 0, :8].tolist())
        print("  ref[0,0,:8]", ref[0, 0, :8].tolist())
        print("  y[0,1,:4]  ", y[0, 1, :4].tolist(), " ref[0,1,:4]", ref[0, 1, :4].tolist())
        # does y match ref under some permutation hints?
        yt = y[0, :8, :8]; rt = ref[0, :8, :8]
        print("  corr(y, ref) over tile:", float(torch.corrcoef(torch.stack([y[0].flatten(), ref[0].flatten()]))[0, 1]))
    return float(err.max() / ref.abs().max())
if __name__ == "__main__":
    run(1, 32, 128, 1, 1, 256, bias=False)
    run(1, 32, 128, 1, 1, 256)
    run(1, 16, 128, 1, 1, 256)
    run(1, 32, 128, 3, 1, 256)
    run(1, 64, 128, 3, 2, 512)
    run(2, 128, 256, 7, 3, 1000)
