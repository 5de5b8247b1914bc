"""CPU tier: world_size-2 gloo test of the multi-GPU plan (utterance sharding + optional final gather)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from styletts2_b200.parallel import gather_waveforms, init_from_env, shard_range
    r, _, w = init_from_env("gloo")
    assert (r, w) == (rank, world)
    B, L = 6, 50
    full = torch.arange(B * L, dtype=torch.float32).view(B, L)     # stand-in for the per-utterance waveforms
    lo, hi = shard_range(B, r, w)
    local = full[lo:hi] * 1.0                                         # each rank "synthesises" only its utterances
    shards = gather_waveforms(local, w, dst=0)
    if r == 0:
        q.put(torch.equal(torch.cat(shards), full))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
