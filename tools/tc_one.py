"""One tensor-core conv shape, a few launches (for `ncu --set full -k regex:conv1d_tc`):
python tools/tc_one.py B C K d L mode [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from tc_bench import setup, timeit

if __name__ == "__main__":
    B, C, K, d, L, mode = [int(v) for v in sys.argv[1:7]]
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
    call = setup(B, C, K, d, L, True, mode)
    ms = timeit(call, reps)
    print(f"B{B} C{C} K{K} d{d} L{L} mode{mode}: {ms:.3f} ms, {2.0 * C * C * K * L * B / ms / 1e9:.0f} TF/s fp32-eq")
