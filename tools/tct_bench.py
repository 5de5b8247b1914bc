"""Time-major vs channel-major tensor-core conv on the narrow HiFi-GAN / iSTFTNet shapes (CUDA events, warm):
python tools/tct_bench.py > gpurun_out/tct_bench.txt"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from styletts2_b200 import ops
from styletts2_b200.lib import ACT_SNAKE

D = "cuda:0"


def setup(B, Cin, Cout, K, d, L, res, tmax):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Cin, L, generator=g).to(D)
    w = (torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)).to(D)
    bias = torch.randn(Cout, generator=g).to(D)
    a = torch.ones(B, Cin, device=D); b = torch.zeros(B, Cin, device=D); alpha = torch.ones(1, Cin, 1, device=D)
    r = torch.randn(B, Cout, L, generator=g).to(D) if res else None
    ops.TC_TMAJOR_MAX_COUT = tmax
    wt, wtc = ops.conv_weight_layout(w), ops.conv_tc_weight_layout(w, 0)
    pad = (K * d - d) // 2
    out = torch.empty(B, Cout, L, device=D)

    def call():
        return ops.conv1d(x, wt, bias, K=K, dil=d, pad=pad, pre=(a, b), pre_act=ACT_SNAKE, alpha=alpha, res=r, want_stats=True, wtc=wtc, out=out)
    return call


def timeit(call, reps=5):
    call(); call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    shapes = [  # B, Cin, Cout, K, d, L, res
        (8, 32, 32, 11, 1, 307200, True), (8, 32, 32, 7, 1, 307200, True), (8, 32, 32, 3, 1, 307200, True), (8, 32, 32, 11, 5, 307200, False),
        (8, 64, 64, 11, 1, 153600, True), (8, 64, 64, 7, 1, 153600, True), (8, 64, 64, 3, 1, 153600, True),
        (8, 32, 1, 7, 1, 307200, False),
        (32, 128, 22, 7, 1, 61441, False), (32, 22, 128, 1, 1, 61441, False),
        (8, 128, 128, 3, 1, 61441, True), (8, 128, 128, 7, 1, 61441, True), (8, 128, 128, 11, 1, 61441, True),
        (8, 128, 128, 3, 1, 51200, True), (8, 128, 128, 7, 1, 51200, True),
    ]
    if os.environ.get("TCT_TM_ONLY"):   # A/B runs of kernel variants: the time-major column only
        for (B, Cin, Cout, K, d, L, res) in shapes:
            ms = timeit(setup(B, Cin, Cout, K, d, L, res, 128))
            print(f"B{B} ci{Cin} co{Cout} k{K} d{d} L{L} res={int(res)}: {ms:.3f} ms", flush=True)
        sys.exit(0)
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in sys.argv[1:7]) + (True,)]
    for (B, Cin, Cout, K, d, L, res) in shapes:
        fl = 2.0 * B * Cin * Cout * K * L
        by = 4.0 * B * L * (Cin + Cout * (2 if res else 1))
        row = f"B{B} ci{Cin} co{Cout} k{K} d{d} L{L} res={int(res)}  ({fl / 1e9:.1f} GFLOP, {by / 1e6:.0f} MB):"
        for tmax, name in ((0, "channel-major"), (128, "time-major")):
            ms = timeit(setup(B, Cin, Cout, K, d, L, res, tmax))
            row += f"  {name} {ms:.3f} ms ({fl / ms / 1e9:.0f} TF/s fp32-eq, {by / ms / 1e6:.0f} GB/s)"
        print(row, flush=True)
