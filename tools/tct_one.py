"""One time-major conv shape, a few launches (for `ncu --set full -k regex:conv1d_tct`):
python tools/tct_one.py B Cin Cout K d L [tmax]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tct_bench import setup, timeit

if __name__ == "__main__":
    B, Cin, Cout, K, d, L = [int(v) for v in sys.argv[1:7]]
    tmax = int(sys.argv[7]) if len(sys.argv) > 7 else 128
    ms = timeit(setup(B, Cin, Cout, K, d, L, True, tmax), 3)
    print(f"B{B} ci{Cin} co{Cout} K{K} d{d} L{L} tmax{tmax}: {ms:.3f} ms")
