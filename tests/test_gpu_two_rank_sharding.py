"""GPU tier: the real engine under the multi-process launch.  Two processes (one per rank; both on cuda:0 when the box has
a single GPU, so the collective runs over gloo on host copies -- NCCL needs one device per rank) each synthesise their
contiguous shard of ONE global batch through Synthesizer, rank 0 gathers the shards (uneven: 3 + 2 utterances) and
compares with its own unsharded run of the whole batch: bitwise (same kernels, per-utterance reduction order)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    from styletts2_b200.configs import MODEL_CFGS
    from styletts2_b200.inference import Synthesizer
    from styletts2_b200.models import build_model, load_keyed_weights, recursive_munch
    from styletts2_b200.parallel import gather_waveforms, init_from_env, shard_range
    from styletts2_b200.synthetic import synthetic_batch
    r, local, w = init_from_env("gloo")
    dev = torch.device("cuda", local % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    cfg = MODEL_CFGS["ljspeech"]
    model = build_model(recursive_munch(cfg))
    for k in model:
        model[k].to(dev).eval()
    load_keyed_weights(model)
    syn = Synthesizer(model, cfg, dev)
    B, N, fpt, K = 5, 128, 2, 3
    T = N * fpt
    tokens, lengths, bert_dur, noise, _ = synthetic_batch(B, N, False, seed=3)
    g = torch.Generator().manual_seed(17)
    steps = [torch.randn(B, 1, 256, generator=g) for _ in range(K - 1)]
    sine = torch.randn(B, 600 * T, 9, generator=g)

    def run(lo, hi):
        inj = dict(step_noises=[s[lo:hi].to(dev) for s in steps], sine_noise=sine[lo:hi].to(dev))
        out = syn.synthesize(tokens[lo:hi].to(dev), lengths[lo:hi].to(dev), bert_dur[lo:hi].to(dev), noise[lo:hi].to(dev),
                             diffusion_steps=K, pin_frames_per_token=fpt, rng=inj)
        return out["wav"].view(hi - lo, -1).cpu()

    lo, hi = shard_range(B, r, w)
    shards = gather_waveforms(run(lo, hi), w, dst=0, batch=B)
    if r == 0:
        full = run(0, B)
        got = torch.cat(shards)
        q.put((bool(torch.equal(got, full)), float((got - full).abs().max()), [int(s.shape[0]) for s in shards]))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_process_sharded_synthesis_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    equal, diff, sizes = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sizes == [3, 2]
    assert equal, diff
