"""Utterance-level sharding across the GPUs of one box (one process per GPU).

The reference is single-process; its per-batch / per-utterance loops are independent
(inference_alignment.py:145, utils/alignment.py:140), so the path shards by utterance with NO
data-path collective. The only exchange is the final gather of alignments (and per-utterance
MAE terms) to rank 0 -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .alignment import AlignResult


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split (first n % world ranks get one extra item)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_frames(t_len, world: int) -> List[np.ndarray]:
    """Length-aware split: utterances sorted by frame count, dealt round-robin (snake order) so
    every rank reads about the same number of logit bytes. Returns index arrays per rank."""
    order = np.argsort(-np.asarray(t_len), kind="stable")
    buckets: List[list] = [[] for _ in range(world)]
    for i, u in enumerate(order):
        r = i % (2 * world)
        buckets[r if r < world else 2 * world - 1 - r].append(int(u))
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def gather_alignments(res: AlignResult, device: Optional[torch.device] = None, dst: int = 0,
                      group=None) -> Optional[AlignResult]:
    """Ragged gather of every rank's AlignResult to `dst`, concatenated in rank order.
    Two collectives: an all_gather of the (n_utt, n_labels) counts, then one gather to `dst` of a
    padded int32 payload [first | last_plus1 | status | l_len | score bits]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = device or torch.device("cpu")
    n_u, n_l = len(res.status), len(res.first)
    counts = torch.tensor([n_u, n_l], dtype=torch.int64, device=device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = torch.stack(all_counts).cpu().numpy()
    mu, ml = int(all_counts[:, 0].max()), int(all_counts[:, 1].max())
    width = 2 * ml + 4 * mu
    payload = np.zeros(max(width, 1), np.int32)
    payload[0:n_l] = res.first
    payload[ml:ml + n_l] = res.last_plus1
    payload[2 * ml:2 * ml + n_u] = res.status
    payload[2 * ml + mu:2 * ml + mu + n_u] = res.l_len
    payload[2 * ml + 2 * mu:2 * ml + 2 * mu + 2 * n_u] = np.ascontiguousarray(res.score, np.float64).view(np.int32)
    mine = torch.from_numpy(payload).to(device)
    # only `dst` needs the payloads: a gather moves 1/world of what an all_gather would (170 ms -> ~25 ms for the
    # 24 M labels of the 10^6-clip run at world 8)
    # (one [world, width] buffer on `dst`, so the payloads come back to the host in ONE copy, not one per rank)
    buf = torch.empty((world, mine.numel()), dtype=mine.dtype, device=device) if rank == dst else None
    dist.gather(mine, list(buf.unbind(0)) if rank == dst else None, dst=dst, group=group)
    if rank != dst:
        return None
    # compact on `dst`'s device (slices of the padded rows, one torch.cat), then ONE copy to the host and zero-copy
    # numpy views: unpacking 8 x 4 MB rows with numpy cost rank 0 ~20 ms, this costs the copy
    firsts, lasts, stats, lens, scores = [], [], [], [], []
    for r in range(world):
        row = buf[r]
        u, l = int(all_counts[r, 0]), int(all_counts[r, 1])
        firsts.append(row[0:l]); lasts.append(row[ml:ml + l])
        stats.append(row[2 * ml:2 * ml + u]); lens.append(row[2 * ml + mu:2 * ml + mu + u])
        scores.append(row[2 * ml + 2 * mu:2 * ml + 2 * mu + 2 * u])
    tot_l, tot_u = int(all_counts[:, 1].sum()), int(all_counts[:, 0].sum())
    flat = torch.cat(firsts + lasts + scores + stats + lens).cpu().numpy()     # [first | last | score bits | status | l_len]
    o = 0
    first = flat[o:o + tot_l]; o += tot_l
    last = flat[o:o + tot_l]; o += tot_l
    score = flat[o:o + 2 * tot_u].view(np.float64); o += 2 * tot_u               # 2 tot_l int32 before it: 8-byte aligned
    status = flat[o:o + tot_u]; o += tot_u
    l_len = flat[o:o + tot_u]
    return AlignResult(first, last, score, status, l_len)


def average_mae_in_dataset_order(per_batch_mae: List[float]) -> float:
    """inference_alignment.py:172-177: unweighted mean of per-batch MAEs, summed sequentially in
    dataset order in Python fp64 (so a sharded run prints the same number)."""
    total = 0
    for m in per_batch_mae:
        total += m
    return total / len(per_batch_mae)
