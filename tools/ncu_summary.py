"""Key metrics of one kernel launch from an `ncu --set full` report -> markdown (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/r02_conv_k11_B32.ncu-rep "title" [algorithmic_bytes] [algorithmic_flops] > profiles/...md"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration [us]"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed (avg)"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock [GHz]"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active [% of elapsed]"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active [% of active]"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed", "uniform pipe [%]"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active [% of peak]"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput [%]"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput [% of peak]"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs (x32 B)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput [%]"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate [%]"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts (LSU)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    alg_bytes = float(sys.argv[3]) if len(sys.argv) > 3 else None
    alg_flops = float(sys.argv[4]) if len(sys.argv) > 4 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"# {title}\n")
    print(f"`ncu --set full --clock-control none --import-source on` (report `{rep.split('/')[-1]}`, not committed: 9 MB); kernel `{m.get('Kernel Name', ('?',))[0]}`.\n")
    print("| metric | value |\n|---|---|")
    got = {}
    for key, label in WANT:
        if key in m:
            v, u = m[key]
            got[key] = (v, u)
            print(f"| {label} (`{key}`) | {v} {u} |")
    def num(key):
        v, u = got[key]
        x = float(v.replace(",", ""))
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
        return x * mult
    try:
        traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        dur_us = float(got["gpu__time_duration.sum"][0].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(got["gpu__time_duration.sum"][1], 1.0)
        print(f"\nDRAM traffic = {traffic / 1e6:.1f} MB per launch" + (f" vs {alg_bytes / 1e6:.1f} MB algorithmic ({traffic / alg_bytes:.2f}x)" if alg_bytes else "") +
              f"; {traffic / dur_us / 1e3:.0f} GB/s under the profiler.")
        if alg_flops:
            print(f"Algorithmic {alg_flops / 1e9:.1f} GFLOP -> {alg_flops / dur_us / 1e6:.0f} TFLOP/s fp32-equivalent under the profiler (cold caches, serialised).")
        json.dump({"dram_bytes_per_launch": traffic, "duration_us": dur_us, "report": rep.split('/')[-1]}, open(rep + ".traffic.json", "w"))
    except Exception as e:   # noqa
        print(f"\n(traffic summary unavailable: {e})")


if __name__ == "__main__":
    main()
