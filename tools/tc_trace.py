"""Per-role cycle trace of the tcgen05 conv kernel (CTA 0, first 16 tiles) for representative vocoder shapes."""
import math, os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from styletts2_b200 import ops, lib
from styletts2_b200.lib import ACT_SNAKE
D = "cuda:0"
def run(B, C, K, d, L, res, label):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C, L, generator=g).to(D); w = (torch.randn(C, C, K, generator=g) / math.sqrt(C * K)).to(D)
    bias = torch.randn(C, generator=g).to(D)
    a = torch.ones(B, C, device=D); b = torch.zeros(B, C, device=D); alpha = torch.ones(1, C, 1, device=D)
    r = torch.randn(B, C, L, generator=g).to(D) if res else None
    wt, wtc = ops.conv_weight_layout(w), ops.conv_tc_weight_layout(w)
    trace = torch.zeros(4 * 16 * 8, dtype=torch.int64, device=D)
    pad = (K * d - d) // 2
    def call():
        return ops.conv1d(x, wt, bias, K=K, dil=d, pad=pad, pre=(a, b), pre_act=ACT_SNAKE, alpha=alpha, res=r, want_stats=True, wtc=wtc)
    call(); torch.cuda.synchronize()
    lib.call("st2_debug_set_trace", ctypes.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    lib.call("st2_debug_set_trace", None)
    ms = e0.elapsed_time(e1)
    t = trace.cpu().view(4, 16, 8)
    flops = 2.0 * C * C * K * L * B
    print(f"== {label}: B{B} C{C} K{K} d{d} L{L} res={res}: {ms:.3f} ms, {flops / ms / 1e9:.1f} TFLOP/s-equivalent")
    base = int(t[0, 0, 0])
    names = ["MMA ", "WPRD", "STAG", "EPI "]
    for it in range(6):
        row = []
        for role in (0, 3):   # the kernel records the MMA issuer (waits: TMEM-empty, act-full, weights-full) and one epilogue warp
                              # (waits: TMEM-full; cycles in residual+TMEM load, row stores, statistics); producer / stager slots are unused
            s_, e_ = int(t[role, it, 0]) - base, int(t[role, it, 1]) - base
            extra = [int(v) for v in t[role, it, 2:6]]
            row.append(f"{names[role]} [{s_:>7},{e_:>7}] w={extra}")
        print(f"  tile{it}: " + " | ".join(row))
if __name__ == "__main__":
    run(32, 128, 3, 1, 61441, False, "stage1 conv1 K3")
    run(32, 128, 3, 1, 61441, True, "stage1 conv2 K3")
    run(32, 128, 11, 5, 61441, True, "stage1 conv K11 d5")
