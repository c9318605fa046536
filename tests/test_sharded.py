"""CPU, world_size 2 over gloo: the utterance-sharding host logic (partition + ragged gather of
alignments to rank 0). The data path itself has no collective; this is the only exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lyricalignment_b200.alignment import AlignResult
from lyricalignment_b200 import sharded


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_result(lo, hi, l_len_all):
    """Deterministic stand-in for what a rank's GPU would return for utterances [lo, hi)."""
    l_len = l_len_all[lo:hi]
    n = int(l_len.sum())
    base = int(l_len_all[:lo].sum())
    first = (np.arange(n) + base).astype(np.int32) * 3
    return AlignResult(first, first + 2, np.arange(lo, hi, dtype=np.float64) * -1.25 - 0.1,
                       (np.arange(lo, hi) % 3 == 2).astype(np.int32) * 2, l_len.astype(np.int32))


def _worker(rank, world, port, l_len_all, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharded.shard_bounds(len(l_len_all), world, rank)
    res = _fake_result(lo, hi, l_len_all)
    out = sharded.gather_alignments(res, device=torch.device("cpu"), dst=0)
    if rank == 0:
        q.put((out.first, out.last_plus1, out.score, out.status, out.l_len))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_utt", [7, 2, 1])
def test_gather_alignments_gloo_world2(n_utt):
    rng = np.random.default_rng(n_utt)
    l_len_all = rng.integers(1, 9, size=n_utt).astype(np.int32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, l_len_all, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = _fake_result(0, n_utt, l_len_all)
    for g, w in zip(got, (want.first, want.last_plus1, want.score, want.status, want.l_len)):
        assert np.array_equal(g, w)


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 5, 8, 2000, 1_000_003):
        for world in (1, 2, 4, 8):
            spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_frames_balances_bytes():
    rng = np.random.default_rng(0)
    t = rng.integers(250, 751, size=2000)
    parts = sharded.shard_by_frames(t, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(2000))
    loads = np.array([t[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01


def test_average_mae_is_sequential_python_sum():
    maes = [0.1, 0.2, 0.30000000000000004, 1e-17, 0.7]
    total = 0
    for m in maes:
        total += m
    assert sharded.average_mae_in_dataset_order(maes) == total / len(maes)
