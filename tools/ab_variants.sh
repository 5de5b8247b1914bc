# A/B of kernel variants built into tools/_variants/lib_<v>.so (same ABI): time-major micro-benchmark + C2 step
set -e
cp styletts2_b200/libstyletts2_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp tools/_variants/lib_$v.so styletts2_b200/libstyletts2_b200.so
  echo "== variant $v"
  TCT_TM_ONLY=1 timeout 200 python tools/tct_bench.py 2>&1 | tr '\n' ';' ; echo
  timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('C2', d['ms_per_step'], d['stages_ms']['decoder'])"
done
cp /tmp/lib_orig.so styletts2_b200/libstyletts2_b200.so
