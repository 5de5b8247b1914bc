"""Micro-timings of single stages on the GPU (CUDA events, warm): biLSTM recurrence per step, compute_style, one
denoiser evaluation.  Usage: python tools/stage_bench.py  (prints one JSON object)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    import cases
    from styletts2_b200 import ops
    from styletts2_b200.models import STYLE_MODULES, build_model, load_keyed_weights, recursive_munch
    from styletts2_b200.style import LogMel, compute_style
    from styletts2_b200.synthetic import synthetic_wave
    out = {}
    dev = "cuda"
    from styletts2_b200 import lib
    tr = torch.zeros(8, dtype=torch.int64, device=dev)
    for (B, Ls) in [(1, 512), (4, 512), (8, 512), (16, 512), (32, 512), (32, 2000)]:
        H = 256
        gx = torch.randn(B * Ls, 8 * H, device=dev) * 0.3
        whh = torch.randn(2, 4 * H, H, device=dev) / 16
        y = torch.empty(B, Ls, 2 * H, device=dev)
        ms = timeit(lambda: ops.lstm_bidir(gx, whh, y, Ls * 2 * H, 2 * H, 1, B, Ls, H))
        lib.call("st2_debug_lstm_trace", tr.data_ptr())
        ops.lstm_bidir(gx, whh, y, Ls * 2 * H, 2 * H, 1, B, Ls, H)
        torch.cuda.synchronize()
        lib.call("st2_debug_lstm_trace", None)
        t = tr.tolist()
        out[f"lstm_B{B}_L{Ls}"] = {"ms": ms, "us_per_step": 1e3 * ms / Ls,
                                   "cycles_per_step": dict(zip(["wait", "fma", "reduce", "gates", "push"], [round(x / max(1, t[5])) for x in t[:5]]))}
    m = build_model(recursive_munch(cases.MODEL_CFGS["libritts"]))
    for k in m:
        m[k].to(dev).eval()
    load_keyed_weights(m, modules=STYLE_MODULES)
    lm = LogMel().to(dev)
    for (B, sec) in [(1, 5), (8, 5), (1, 20)]:
        w = synthetic_wave(B, 24000 * sec).to(dev)
        out[f"compute_style_B{B}_{sec}s_ms"] = timeit(lambda: compute_style(m, w, lm))
        out[f"logmel_B{B}_{sec}s_ms"] = timeit(lambda: lm(w))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
