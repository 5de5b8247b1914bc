"""One fp32-accurate tensor-core GEMM and one tensor-core attention call at the denoiser's C2 shapes (for ncu --set full):
python tools/gemm_one.py [reps]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from styletts2_b200 import ops

D = "cuda:0"
if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    g = torch.Generator().manual_seed(0)
    M, K, Nf = 4096, 1024, 2048
    A = torch.randn(M, K, generator=g).to(D); W = (torch.randn(Nf, K, generator=g) / math.sqrt(K)).to(D); b = torch.randn(Nf, generator=g).to(D)
    wtc = ops.linear_tc_weight_layout(W)
    B, N, H, Dh = 32, 128, 8, 64
    q = torch.randn(B * N, H * Dh, generator=g).to(D); kv = torch.randn(B * N, 2 * H * Dh, generator=g).to(D)
    for name, fn, fl in (("linear_tc M4096 N2048 K1024", lambda: ops.linear(A, W, b, wtc=wtc), 2.0 * M * K * Nf),
                         ("attention_tc B32 N128 H8", lambda: ops.attention(q, kv, B, N, H, Dh), 4.0 * B * H * N * N * Dh)):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name}: {ms * 1e3:.1f} us, {fl / ms / 1e9:.1f} TF/s fp32-eq")
