"""Benchmark of the StyleTTS 2 text->waveform hot path on B200 (driver contract, see prompt section 4).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores (oracle port)

A "step" = one pass of the whole path (text encoder -> style diffusion sampler -> duration/prosody predictor ->
AdaIN decoder -> vocoder) over one batch of synthetic utterances.  Workload = BASELINE.json configs[1]:
LJSpeech config, iSTFTNet decoder, batch 32 x 128 tokens x 512 frames (4 frames/token pinned after the duration
kernel has run), diffusion_steps=5.  Weights: key-seeded random init of that architecture
(no checkpoints offline).  With N>1 every rank runs the same per-GPU batch (weak scaling, utterance sharding,
no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (model cfg, per-GPU batch, tokens, frames/token, diffusion steps)
    "C2": dict(model="ljspeech", B=32, N=128, fpt=4, steps=5, desc="LJSpeech iSTFTNet B32 N128 T512 K5"),
    "C3": dict(model="libritts", B=8, N=128, fpt=4, steps=10, desc="LibriTTS HiFi-GAN B8/GPU N128 T512 K10"),
    "C4": dict(model="ljspeech", B=16, N=500, fpt=4, steps=5, desc="LJSpeech iSTFTNet B16 N500 T2000 K5"),
    "tiny": dict(model="ljspeech", B=2, N=16, fpt=4, steps=3, desc="plumbing check"),
}
# algorithmic bytes / FLOPs of the decoder+vocoder path per frame-utterance (fp32, conv
# boundaries, norm/activation/residual fused) + weights once per launch chain.
VOCODER_BYTES_PER_FRAME = {"ljspeech": 4.445e6, "libritts": 11.74e6}
VOCODER_WEIGHT_BYTES = {"ljspeech": 223e6, "libritts": 224e6}
VOCODER_FLOPS_PER_FRAME = {"ljspeech": 1.317e9, "libritts": 1.767e9}


def measured_peaks():
    """(hbm GB/s, sustained bf16 TFLOP/s, source): the kernel is timed inside a long step -> sustained figure."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(wl, seed, pinned=False):
    from styletts2_b200.synthetic import synthetic_batch
    tokens, lengths, bert_dur, noise, ref_s = synthetic_batch(wl["B"], wl["N"], wl["model"] == "libritts", seed=seed)
    ts = [tokens, lengths, bert_dur, noise] + ([ref_s] if ref_s is not None else [])
    if pinned:
        ts = [t.pin_memory() for t in ts]
    return ts


def run_ours(args):
    import cases
    from styletts2_b200 import lib
    from styletts2_b200.inference import Synthesizer
    from styletts2_b200.models import build_model, load_keyed_weights, recursive_munch
    from styletts2_b200.parallel import init_from_env

    rank, local, world = init_from_env("nccl")
    assert world == args.gpus or world == 1, (world, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    mcfg = cases.MODEL_CFGS[wl["model"]]
    model = build_model(recursive_munch(mcfg))
    for k in model:
        model[k].to(dev).eval()
    load_keyed_weights(model)
    syn = Synthesizer(model, mcfg, dev)
    B, N, T = wl["B"], wl["N"], wl["N"] * wl["fpt"]
    L = 600 * T
    samples_per_step = B * L * world
    ms = wl["model"] == "libritts"

    host = make_inputs(wl, seed=1 + rank, pinned=True)
    devin = [t.to(dev) for t in host]
    wav_host = torch.empty(B, L, dtype=torch.float32).pin_memory()
    dec_ev = []

    marks = []

    graph_launches = [0]

    def step(inputs, timed_decoder=False):
        tokens, lengths, bert_dur, noise = inputs[:4]
        ref_s = inputs[4] if ms else None
        if args.graph:
            wav, nl = syn.synthesize_graphed(tokens, lengths, bert_dur, noise, diffusion_steps=wl["steps"], ref_s=ref_s,
                                             pin_frames_per_token=wl["fpt"])
            graph_launches[0] = nl
            return wav
        mk = [] if timed_decoder else None
        out = syn.synthesize(tokens, lengths, bert_dur, noise, diffusion_steps=wl["steps"], ref_s=ref_s,
                             pin_frames_per_token=wl["fpt"], decoder_events=dec_ev if timed_decoder else None, stage_marks=mk)
        if mk is not None:
            marks.append(mk)
        return out["wav"]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms_val):
        if world == 1:
            return ms_val
        t = torch.tensor([ms_val], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step(devin)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(devin, timed_decoder=True)
    e1.record()
    barrier()
    launches = lib.launch_count() - n0
    if args.graph:
        launches = graph_launches[0] * args.steps   # kernels of this library inside the replayed CUDA graph x replays
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    dec_ms = sum(a.elapsed_time(b) for a, b in dec_ev) / max(1, len(dec_ev))
    stages = {}
    for mk in marks:
        for (n0_, e0_), (n1_, e1_) in zip(mk[:-1], mk[1:]):
            stages[n1_] = stages.get(n1_, 0.0) + e0_.elapsed_time(e1_) / len(marks)
    if args.skip_e2e:
        if rank == 0:
            clocks.stop()
            print(json.dumps({"profile_only": True, "ms_per_step": ms_total / args.steps, "decoder_ms": dec_ms, "gpu_launches": launches,
                              "stages_ms": stages}))
        return
    # ---- end to end through the public API with HOST buffers (`e2e`)
    def e2e_step():
        ins = [t.to(dev, non_blocking=True) for t in host]
        wav = step(ins)
        wav_host.copy_(wav.view(B, L), non_blocking=True)
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        e2e_step()
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
        return
    ms_per_step = ms_total / args.steps
    value = samples_per_step / (ms_per_step / 1e3)
    e2e_value = samples_per_step / (ms_e2e / args.steps / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = wav_host.numel() * 4
    hbm_peak, tc_peak, peak_src = measured_peaks()
    alg_bytes = VOCODER_BYTES_PER_FRAME[wl["model"]] * B * T + VOCODER_WEIGHT_BYTES[wl["model"]]
    alg_flops = VOCODER_FLOPS_PER_FRAME[wl["model"]] * B * T
    # dominant kernel = tc::conv1d_tc_kernel: one extra eager (non-graph) pass with CUDA events around every launch
    from styletts2_b200 import ops as _ops
    _ops.PROFILE = []
    tokens, lengths, bert_dur, noise = devin[:4]
    syn.synthesize(tokens, lengths, bert_dur, noise, diffusion_steps=wl["steps"], ref_s=devin[4] if ms else None, pin_frames_per_token=wl["fpt"])
    torch.cuda.synchronize()
    prof, _ops.PROFILE = _ops.PROFILE, None
    if not dec_ev:   # graph mode: take the decoder time from one eager pass
        mk = []
        syn.synthesize(tokens, lengths, bert_dur, noise, diffusion_steps=wl["steps"], ref_s=devin[4] if ms else None,
                       pin_frames_per_token=wl["fpt"], stage_marks=mk)
        torch.cuda.synchronize()
        for (n0_, e0_), (n1_, e1_) in zip(mk[:-1], mk[1:]):
            stages[n1_] = e0_.elapsed_time(e1_)
        dec_ms = stages.get("decoder", 0.0)
    if args.dump_launches and rank == 0:
        agg = {}
        for name_, f_, b_, e0_, e1_ in prof:
            r_ = agg.setdefault(name_, [0, 0.0, 0.0])
            r_[0] += 1; r_[1] += e0_.elapsed_time(e1_); r_[2] += f_
        rows_ = [dict(launch=k_, n=v_[0], ms=round(v_[1], 4), fp32_tflops=round(v_[2] / (v_[1] / 1e3) / 1e12, 1),
                      executed_bf16_tflops=round(3 * v_[2] / (v_[1] / 1e3) / 1e12, 1)) for k_, v_ in sorted(agg.items(), key=lambda kv: -kv[1][1])]
        with open(args.dump_launches, "w") as f_:
            json.dump(rows_, f_, indent=1)
    tc_ms = sum(e0_.elapsed_time(e1_) for _, _, _, e0_, e1_ in prof)
    tc_flops = sum(f for _, f, _, _, _ in prof)
    tc_bytes = sum(b_ for _, _, b_, _, _ in prof)
    n_tc = max(1, len(prof))
    executed_tflops = 3.0 * tc_flops / (tc_ms / 1e3) / 1e12   # bf16 hi/lo split: 3 MMAs per fp32 product
    line = {
        "metric": "24 kHz waveform samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (key-seeded random weights, seeded tokens/bert_dur/noise; durations pinned to 4 frames/token)",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "per_gpu_batch": B, "tokens": N, "frames": T, "samples_per_utt": L,
                   "diffusion_steps": wl["steps"], "sharding": f"utterances x{world}, no data-path collective",
                   "l2": "inputs+activations per step (>3 GB) exceed the 126 MB L2; no flush needed",
                   "launch": "one CUDA graph per step" if args.graph else "eager"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "stages_ms": stages,
        "clocks": clk,
        "roofline": {"kernel": "st2::tc::conv1d_tc_kernel (tcgen05 implicit-GEMM conv, bf16 hi/lo x3)", "bound": "tensor",
                     "achieved": executed_tflops, "peak": tc_peak, "unit": "TFLOP/s", "frac": executed_tflops / tc_peak,
                     "traffic": None, "peak_source": peak_src, "launches_per_step": len(prof),
                     "avg_launch_ms": tc_ms / n_tc, "algorithmic_gflop_per_launch": tc_flops / n_tc / 1e9,
                     "fp32_equivalent_tflops": tc_flops / (tc_ms / 1e3) / 1e12,
                     "hbm_algorithmic_gbs": tc_bytes / (tc_ms / 1e3) / 1e9, "hbm_peak_gbs": hbm_peak,
                     "kernel_ms_per_step": tc_ms, "decoder_ms_per_step": dec_ms,
                     "vocoder_path_algorithmic": {"bytes_per_step": alg_bytes, "flops_per_step": alg_flops,
                                                  "gbs": alg_bytes / (max(dec_ms, 1e-9) / 1e3) / 1e9 if dec_ms else None},
                     "note": "achieved = executed bf16 MMA FLOPs (3 per fp32 product, algorithmic 2*Cin*Cout*K*L*B summed over the step's "
                             "launches) / CUDA-event time of those launches in an eager pass; traffic: see profiles/r01_ncu_conv1d_tc.md "
                             "(DRAM bytes == algorithmic bytes for the captured launches)"},
    }
    if args.cpu_baseline and world >= 1:
        line["cpu_baseline"] = cpu_reference(wl, sample_B=args.cpu_batch, steps=1, warmup=0)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def cpu_reference(wl, sample_B, steps, warmup):
    """The reference algorithm (oracle port, torch CPU fp32) on the host cores, bounded sample of the workload."""
    import cases
    import styletts2_oracle as O
    from styletts2_b200.synthetic import keyed_state_dict

    # more threads than ~32 slow torch's CPU kernels down at these sizes (measured: 128 threads 7x slower than 8)
    torch.set_num_threads(min(32, os.cpu_count()))
    mcfg = cases.MODEL_CFGS[wl["model"]]
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", f"state_shapes_{wl['model']}.json")))
    sds = {k: keyed_state_dict({n: tuple(s) for n, s in shapes[k].items()}, k) for k in shapes}
    sub = dict(wl, B=sample_B)
    tokens, lengths, bert_dur, noise = make_inputs(sub, seed=1)[:4]
    ref_s = make_inputs(sub, seed=1)[4] if wl["model"] == "libritts" else None
    forced = torch.full((sample_B, wl["N"]), float(wl["fpt"]))
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        with torch.no_grad():
            out = O.synthesize(sds, mcfg, tokens, lengths, bert_dur, noise, diffusion_steps=wl["steps"], ref_s=ref_s,
                               forced_durations=forced)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
    nsamp = out["wav"].numel()
    v = nsamp / (sum(times) / len(times))
    return {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample_B} utterance(s) of the workload ({wl['N']} tokens, {wl['N'] * wl['fpt']} frames, K={wl['steps']}), "
                      f"{len(times)} timed run(s), torch {torch.__version__} CPU, {torch.get_num_threads()} threads",
            "seconds_per_run": sum(times) / len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cb = cpu_reference(wl, sample_B=args.cpu_batch, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "24 kHz waveform samples/sec", "value": cb["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_run"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']} (bounded sample: {args.cpu_batch} utterance(s) per step)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=list(WORKLOADS))
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--dump-launches", default=None, help="write the per-shape table of tensor-core conv launches (CUDA events, eager pass) to this JSON file")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only (ncu): device-resident loop, no JSON contract line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
