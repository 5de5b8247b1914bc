"""CPU tier: world_size-2 gloo tests of the multi-GPU plan's HOST logic: contiguous utterance sharding of one global batch
(bench.py --global-batch / parallel.shard_range), the uneven-shard gather, and the serving planner.  The engine itself
needs a GPU: tests/test_gpu_two_rank_sharding.py runs the real Synthesizer in two processes."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, B):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from styletts2_b200.parallel import gather_waveforms, init_from_env, shard_range
    from styletts2_b200.synthetic import synthetic_batch
    r, _, w = init_from_env("gloo")
    assert (r, w) == (rank, world)
    # the same seeded global batch on every rank, each keeps its contiguous shard (what bench.py --global-batch does)
    tokens, lengths, bert_dur, noise, _ = synthetic_batch(B, 12, False, seed=1)
    lo, hi = shard_range(B, r, w)
    L = 40
    # stand-in for the engine: a per-utterance function of that utterance's inputs only (as every op on the path is)
    local = (bert_dur[lo:hi].sum(dim=(1, 2))[:, None] + noise[lo:hi, 0, :L] + tokens[lo:hi, :1].float()).contiguous()
    shards = gather_waveforms(local, w, dst=0, batch=B)
    if r == 0:
        full = bert_dur.sum(dim=(1, 2))[:, None] + noise[:, 0, :L] + tokens[:, :1].float()
        q.put(bool(torch.equal(torch.cat(shards), full)) and [s.shape[0] for s in shards] == [shard_range(B, i, w)[1] - shard_range(B, i, w)[0] for i in range(w)])
    else:
        assert shards is None
    dist.barrier()
    dist.destroy_process_group()


def _run(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, B)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_two_rank_shard_and_gather_gloo_even():
    _run(6)


def test_two_rank_shard_and_gather_gloo_uneven():
    _run(7)      # shards of 4 and 3 utterances: padded gather, sliced on the destination


def test_shard_range_covers_every_utterance_once():
    from styletts2_b200.parallel import shard_range
    for B in (1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            cover = []
            for r in range(w):
                lo, hi = shard_range(B, r, w)
                cover.extend(range(lo, hi))
            assert cover == list(range(B))
