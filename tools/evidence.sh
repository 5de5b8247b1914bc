# Round-end evidence on one B200 (run through gpurun): GPU test suite, the three bench configurations with launch tables,
# memcheck on the tiny workload.  FULL=1 adds the ncu captures (set full of the time-major kernel on three launch classes,
# launch list of an eager C2 pass).  Outputs under gpurun_out/ev/; the summaries are copied into profiles/ by hand.
set -x
mkdir -p gpurun_out/ev
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ev/pytest_gpu.log 2>&1; tail -3 gpurun_out/ev/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --dump-launches gpurun_out/ev/launches_c2.json > gpurun_out/ev/bench_c2.json 2> gpurun_out/ev/bench_c2.err
timeout 200 python bench.py --workload C3 --no-cpu-baseline --dump-launches gpurun_out/ev/launches_c3.json > gpurun_out/ev/bench_c3.json 2> gpurun_out/ev/bench_c3.err
timeout 200 python bench.py --workload C4 --no-cpu-baseline > gpurun_out/ev/bench_c4.json 2> gpurun_out/ev/bench_c4.err
python -c "
import json
for w in ['c2','c3','c4']:
    d=json.load(open('gpurun_out/ev/bench_%s.json'%w)); print(w, d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_executed'], d['stages_ms'])
"
timeout 300 compute-sanitizer --tool memcheck --report-api-errors no python bench.py --workload tiny --steps 1 --warmup 1 --no-graph --skip-e2e --no-cpu-baseline > gpurun_out/ev/sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/ev/sanitizer_memcheck.txt
if [ -n "$FULL" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tct -s 2 -c 1 -o gpurun_out/ev/tct_c128_k11_B32 -f python tools/tct_one.py 32 128 128 11 1 61441 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tct -s 2 -c 1 -o gpurun_out/ev/tct_c128_k3_B32 -f python tools/tct_one.py 32 128 128 3 1 61441 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tct -s 2 -c 1 -o gpurun_out/ev/tct_c32_k11_B8 -f python tools/tct_one.py 8 32 32 11 1 307200 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ev/launch_list_c2.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e --no-cpu-baseline > gpurun_out/ev/launch_list.log 2>&1; tail -2 gpurun_out/ev/launch_list.log
fi
