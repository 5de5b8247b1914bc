"""Multi-GPU plan: utterances are independent (no cross-utterance op anywhere on the path:
InstanceNorm is per (b,c), attention is within an utterance, LSTMs are per sequence), so the batch
is split contiguously across ranks with replicated weights and NO collective on the data path
(SURVEY.md section 8e).  The only optional collective is a final gather of the waveforms."""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `batch` utterances; the first (batch % world) ranks get one extra."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def plan_equal_length_batches(lengths: Sequence[int], world: int, max_batch: int) -> List[List[List[int]]]:
    """Serving-side plan for a queue of utterances with DIFFERENT token counts (host part only): the
    batched engine needs equal-length batches (Synthesizer.synthesize), so utterances are bucketed by token count, cut into
    batches of at most `max_batch`, and the batches are dealt to ranks longest-processing-time first (cost ~ tokens x
    utterances) so that every GPU finishes at about the same time.  Deterministic (ties by first utterance index).
    Returns plan[rank] = list of batches, each a list of utterance indices."""
    assert world >= 1 and max_batch >= 1
    buckets: Dict[int, List[int]] = {}
    for i, n in enumerate(lengths):
        buckets.setdefault(int(n), []).append(i)
    batches = []
    for n in sorted(buckets):
        idx = buckets[n]
        for k in range(0, len(idx), max_batch):
            batches.append((n * len(idx[k:k + max_batch]), idx[k:k + max_batch]))
    batches.sort(key=lambda b: (-b[0], b[1][0]))
    load = [0] * world
    plan: List[List[List[int]]] = [[] for _ in range(world)]
    for cost, idx in batches:
        r = min(range(world), key=lambda q: (load[q], q))
        plan[r].append(idx)
        load[r] += cost
    return plan


def init_from_env(backend: str = "nccl"):
    """torchrun environment -> (rank, local_rank, world).  Single process if WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group(backend=backend)
    return rank, local, world


def gather_waveforms(wav_local: torch.Tensor, world: int, dst: int = 0):
    """Optional final collective: gather [B/G, L] fp32 waveforms on `dst` (NCCL over NVLink on GPUs,
    gloo in the CPU tests).  Returns the list of shards on dst, None elsewhere."""
    if world == 1:
        return [wav_local]
    rank = torch.distributed.get_rank()
    if torch.distributed.get_backend() == "nccl":
        out = [torch.empty_like(wav_local) for _ in range(world)]
        torch.distributed.all_gather(out, wav_local.contiguous())
        return out if rank == dst else None
    bufs = [torch.empty_like(wav_local) for _ in range(world)] if rank == dst else None
    torch.distributed.gather(wav_local.contiguous(), bufs, dst=dst)
    return bufs
