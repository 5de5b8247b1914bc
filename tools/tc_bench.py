"""Where does the tensor-core conv spend its time?  Times st2_conv1d_tc on representative vocoder shapes for every
precision recipe and with the kernel's timing-experiment switches (st2_debug_set_flags), and prints the per-role cycle
trace of CTA 0.  Run on the GPU box: python tools/tc_bench.py > gpurun_out/tc_bench.txt"""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from styletts2_b200 import lib, ops
from styletts2_b200.lib import ACT_SNAKE

D = "cuda:0"


def setup(B, C, K, d, L, res, mode):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C, L, generator=g).to(D)
    w = (torch.randn(C, C, K, generator=g) / math.sqrt(C * K)).to(D)
    bias = torch.randn(C, generator=g).to(D)
    a = torch.ones(B, C, device=D); b = torch.zeros(B, C, device=D); alpha = torch.ones(1, C, 1, device=D)
    r = torch.randn(B, C, L, generator=g).to(D) if res else None
    wt, wtc = ops.conv_weight_layout(w), ops.conv_tc_weight_layout(w, mode)
    pad = (K * d - d) // 2
    out = torch.empty(B, C, L, device=D)

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


def trace(call, label):
    tr = torch.zeros(4 * 16 * 8, dtype=torch.int64, device=D)
    lib.call("st2_debug_set_trace", ctypes.c_void_p(tr.data_ptr()))
    call(); torch.cuda.synchronize()
    lib.call("st2_debug_set_trace", None)
    t = tr.cpu().view(4, 16, 8)
    base = int(t[0, 0, 0])
    print(f"  trace {label}  (MMA issuer: [start,end] waits=[tmem-empty, act-full, w-full]; EPI warp: [start,end] waits=[tmem-full] cycles=[ld, store, stats])")
    for it in range(2, 8):
        m = [int(v) for v in t[0, it, :5]]
        e = [int(v) for v in t[3, it, :6]]
        print(f"    tile{it}: MMA [{m[0]-base:>7},{m[1]-base:>7}] dur={m[1]-m[0]:>6} w={m[2:5]} | EPI [{e[0]-base:>7},{e[1]-base:>7}] dur={e[1]-e[0]:>6} w_tf={e[2]} ld={e[3]} st={e[4]} ss={e[5]}")


if __name__ == "__main__":
    B = 8
    shapes = [(B, 128, 11, 1, 61441, True), (B, 128, 3, 1, 61441, True), (B, 128, 3, 1, 61441, False), (B, 128, 11, 1, 61441, False), (B, 128, 3, 1, 61440, True), (B, 128, 7, 1, 61440, True), (B, 128, 7, 1, 61441, True), (4 * B, 256, 7, 1, 10240, True),
              (4 * B, 256, 3, 1, 10240, True), (4 * B, 1024, 3, 1, 512, True), (4 * B, 1024, 1, 1, 512, False)]
    for (b, C, K, d, L, res) in shapes:
        fl = 2.0 * C * C * K * L * b
        print(f"== B{b} C{C} K{K} d{d} L{L} res={res}  ({fl / 1e9:.1f} GFLOP)")
        for mode, name in (((0, "FAST"),) if os.environ.get("TC_BENCH_DEEP") else ((0, "FAST"), (2, "F16X3"), (1, "ACC"))):
            call = setup(b, C, K, d, L, res, mode)
            lib.call("st2_debug_set_flags", 0)
            ms = timeit(call)
            row = [f"{name}: {ms:.3f} ms ({fl / ms / 1e9:.0f} TF/s fp32-eq)"]
            variants = ((1, "f16-2nd"), (2, "no-epi-io"), (4, "no-convert"), (8, "no-mma"), (2 | 4, "no-io+no-conv"), (2 | 4 | 8, "skeleton"))
            if os.environ.get("TC_BENCH_QUICK"):
                variants = ()
            if os.environ.get("TC_BENCH_MMA"):      # production issue path with the other roles switched off one by one
                variants = ((4, "no-convert"), (2, "no-epi-io"), (64, "no-epi"), (2 | 4, "no-io+no-conv"), (4 | 64, "no-conv+no-epi"),
                            (4 | 16 | 32 | 64, "mma+handshake only"), (4 | 32 | 64, "mma+weights only"))
            if mode == 0 and os.environ.get("TC_BENCH_DEEP"):
                variants = ((2 | 4 | 8, "skeleton"), (2 | 4 | 8 | 16, "skel-noW"), (2 | 4 | 8 | 32, "skel-noRaw"), (2 | 4 | 8 | 16 | 32, "skel-handshake"),
                            (4 | 8 | 16 | 32 | 64, "handshake-noEpi"), (16, "full-noW"), (64, "full-noEpi"), (16 | 64, "full-noW-noEpi"), (32, "full-noRaw"),
                            (2, "no-epi-io"), (4, "no-convert"), (8, "no-mma"))
            for flags, fn in variants:
                if flags & 1 and mode != 0:
                    continue
                lib.call("st2_debug_set_flags", flags)
                row.append(f"{fn} {timeit(call):.3f}")
            lib.call("st2_debug_set_flags", 0)
            print("  " + " | ".join(row))
            if mode == 0 and not os.environ.get("TC_BENCH_MMA"):
                trace(call, name)
