"""The sharded runner (lyricalignment_b200.run_sharded): the reference's align_and_evaluate loop
(inference_alignment.py:127-180) dealt over ranks in whole batches, one ragged gather, MAE averaged in dataset
order. CPU: world_size 2 over gloo with fake logits and the oracle standing in for K2/K3 (host logic only).
GPU: the real kernels, world 1 run as two logical shards == the unsharded run == the oracle."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from lyricalignment_b200 import alignment as A, io as la_io, run_sharded as RS

V = 64


def _dataset(n, seed=0, with_gt=True):
    rng = np.random.default_rng(seed)
    recs, toks, auds = [], {}, {}
    for i in range(n):
        L = int(rng.integers(1, 6))
        T = int(rng.integers(3 * L + 4, 40))
        path = f"song_{i}.npy"
        text = "".join(chr(0x4E00 + int(c)) for c in rng.integers(0, 500, size=L))
        gt = [[round(0.02 * k, 2), round(0.02 * (k + 1), 2)] for k in range(L)] if with_gt else None
        recs.append(la_io.Record(path, text, gt))
        toks[path] = rng.integers(1, V - 2, size=L).tolist()
        auds[path] = np.full(T, float(i), np.float32)           # "audio": its length fixes T, its value seeds the logits
    return recs, toks, auds


def _fake_logits(audios):
    """Deterministic [B, T, V] logits from the clips alone (batch padded to the longest, like the reference)."""
    T = max(len(a) for a in audios)
    out = np.zeros((len(audios), T, V), np.float32)
    for b, a in enumerate(audios):
        rng = np.random.default_rng(int(a[0]) + 1000)
        out[b] = 2.0 * rng.standard_normal((T, V))
    return torch.from_numpy(out)


def _oracle_align(logits, tokens, mode):
    """Stand-in for alignment.align in the CPU tests: the checker produces the AlignResult."""
    z = logits.numpy() if torch.is_tensor(logits) else logits
    firsts, lasts, scores, stats, lens = [], [], [], [], []
    emis = oracle.emission_ctc if mode == A.MODE_CTC else oracle.emission_ce
    for b in range(z.shape[0]):
        lab = tokens[b].numpy()
        lab = lab[lab != -100].astype(np.int64)
        e, s = emis(z[b:b + 1])
        r = oracle.align_one(np.ascontiguousarray(e[0]), np.ascontiguousarray(s[0]), lab)
        firsts.append(r["first"]); lasts.append(r["last_plus1"]); scores.append([r["score"]])
        stats.append([r["status"]]); lens.append([len(lab)])
    c = lambda xs, dt: np.concatenate([np.asarray(x) for x in xs]).astype(dt)
    return A.AlignResult(c(firsts, np.int32), c(lasts, np.int32), c(scores, np.float64), c(stats, np.int32), c(lens, np.int32))


def _run(recs, toks, auds, batch_size, align_fn, device=None, logits_fn=_fake_logits):
    return RS.run(recs, lambda r: auds[r.audio_path], lambda r: toks[r.audio_path], logits_fn, None, True,
                  batch_size, device=device, align_fn=align_fn)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, batch_size, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    recs, toks, auds = _dataset(n)
    out = _run(recs, toks, auds, batch_size, _oracle_align)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,batch_size", [(9, 2), (5, 1), (3, 4)])
def test_runner_gloo_world2_equals_single_process(n, batch_size):
    recs, toks, auds = _dataset(n)
    want = _run(recs, toks, auds, batch_size, _oracle_align)             # world 1, no process group
    assert len(want["alignments"]) == n and want["average_mae"] is not None
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, batch_size, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert got["alignments"] == want["alignments"]
    assert got["batch_mae"] == want["batch_mae"]
    assert got["average_mae"] == want["average_mae"]                      # same fp64 bits: dataset-order sum


def test_runner_without_ground_truth_prints_like_the_reference(tmp_path):
    recs, toks, auds = _dataset(4, seed=3, with_gt=False)
    out = _run(recs, toks, auds, 1, _oracle_align)
    assert out["average_mae"] is None and out["batch_mae"] == [None] * 4
    line = la_io.format_prediction(out["alignments"][0], recs[0].text)    # inference_alignment_nogt.py:175-176
    assert line.startswith("[[") and recs[0].text[0] in line
    path = tmp_path / "al.json"
    la_io.write_alignments(str(path), recs, out["alignments"])
    back = la_io.read_data(str(path))
    assert [r.lyric_onset_offset for r in back] == out["alignments"]


@pytest.mark.gpu
def test_runner_on_the_kernels_sharded_equals_unsharded_equals_oracle():
    """World 1, run once over the whole dataset and once as two logical shards (two half datasets whose batch
    boundaries coincide), with the logits on the GPU: same alignments, same per-batch MAEs, and both equal the
    oracle."""
    n, bs = 12, 2
    recs, toks, auds = _dataset(n, seed=5)
    gpu_logits = lambda audios: _fake_logits(audios).cuda()
    whole = _run(recs, toks, auds, bs, A.align, logits_fn=gpu_logits)
    half = n // 2
    a = _run(recs[:half], toks, auds, bs, A.align, logits_fn=gpu_logits)
    b = _run(recs[half:], toks, auds, bs, A.align, logits_fn=gpu_logits)
    assert a["alignments"] + b["alignments"] == whole["alignments"]
    assert a["batch_mae"] + b["batch_mae"] == whole["batch_mae"]
    want = _run(recs, toks, auds, bs, _oracle_align)
    assert whole["alignments"] == want["alignments"]
    assert whole["batch_mae"] == want["batch_mae"] and whole["average_mae"] == want["average_mae"]
