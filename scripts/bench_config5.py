"""BASELINE.json configs[4]: 1M synthetic Opencpop-shaped clips sharded by utterance across 1/2/4/8 B200 with an
NCCL gather of the alignments.

    python scripts/bench_config5.py [--clips 1000000]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        scripts/bench_config5.py --clips 1000000                                        # N GPUs

10^6 x 42 MB of logits cannot be materialised, so every rank recycles ONE resident pool of 2 000
clips' logits (84.5 GB) and waveforms, and aligns it against FRESH lyrics batch after batch: clips/N clips per
rank in 2 000-clip plans. Per batch: K1 over the waveforms, label flattening + la_plan_create, K2, K3, D2H -- the
public API (log_mel_spectrogram_ragged + align_clips_async), two batches in flight so the host work of batch i+1
hides behind the kernels of batch i. At the end ONE ragged gather of everything to rank 0 (sharded.gather_alignments).
Prints one JSON line on rank 0: audio-s/s over the whole job (wall clock between barriers, max over ranks), the
host-side share (wall time not covered by the kernels), and the gather time."""
import argparse, json, os, sys, time
import numpy as np, torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1_000_000)
    ap.add_argument("--pool-clips", type=int, default=2000)
    ap.add_argument("--label-sets", type=int, default=64)
    ap.add_argument("--depth", type=int, default=2)
    args = ap.parse_args()
    import torch.distributed as dist
    from lyricalignment_b200 import alignment as A, audio as LA, sharded, synth
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = sharded.shard_bounds(args.clips, world, rank)
    n_mine = hi - lo
    pool = synth.opencpop_shaped(args.pool_clips, seed=114514 + rank)
    V = synth.V_HEAD
    logits = synth.planted_logits(pool, V, ctc=True, device=dev, seed=114514 + rank)
    wave, w_off = synth.synthetic_waveforms(pool, device=dev, seed=114514 + rank)
    n_samp = pool.n_samples.astype(np.int32)
    mel_frames = (n_samp // 160).astype(np.int32)
    mel_off = np.concatenate([[0], np.cumsum(80 * mel_frames.astype(np.int64))[:-1]]).astype(np.int64)
    mel_out = torch.empty(int(80 * mel_frames.astype(np.int64).sum()), dtype=torch.float32, device=dev)
    label_sets = [synth.fresh_labels(pool, 9000 + 131 * rank + j) for j in range(args.label_sets)]
    n_batches = (n_mine + args.pool_clips - 1) // args.pool_clips

    def run_batches(nb, collect):
        jobs, out, host_s = [], [], 0.0
        for b in range(nb):
            nclip = min(args.pool_clips, n_mine - b * args.pool_clips)
            t0 = time.perf_counter()
            LA.log_mel_spectrogram_ragged(wave, w_off[:nclip], n_samp[:nclip], out=mel_out, out_offsets=mel_off[:nclip],
                                          out_strides=mel_frames[:nclip])
            rows = int(pool.t_len[:nclip].sum())
            jobs.append(A.align_clips_async(logits[:rows], pool.t_len[:nclip], label_sets[b % len(label_sets)][:nclip]))
            host_s += time.perf_counter() - t0
            if len(jobs) >= args.depth:
                r = jobs.pop(0).result()
                if collect:
                    out.append(r)
        for j in jobs:
            r = j.result()
            if collect:
                out.append(r)
        return out, host_s

    warm, _ = run_batches(min(3, n_batches), True)             # warm-up (kernels, allocator, pinned pool)
    if world > 1:
        sharded.gather_alignments(warm[0], device=dev)         # ... and NCCL's lazily-built point-to-point channels
    del warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    results, host_s = run_batches(n_batches, True)
    e1.record()
    torch.cuda.synchronize()
    t_compute = time.perf_counter() - t0
    cat = lambda f, dt: np.concatenate([getattr(r, f) for r in results]).astype(dt)
    mine = A.AlignResult(cat("first", np.int32), cat("last_plus1", np.int32), cat("score", np.float64),
                         cat("status", np.int32), cat("l_len", np.int32))
    assert int(mine.status.max()) == 0 and len(mine.status) == n_mine
    tg = time.perf_counter()
    allres = sharded.gather_alignments(mine, device=dev) if world > 1 else mine
    torch.cuda.synchronize()
    t_gather = time.perf_counter() - tg
    t_total = time.perf_counter() - t0
    tt = torch.tensor([t_total, t_compute, t_gather, host_s, e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
    audio = torch.tensor([float(sum(pool.durations[:min(args.pool_clips, n_mine - b * args.pool_clips)].sum() for b in range(n_batches)))],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(audio, op=dist.ReduceOp.SUM)
    if rank == 0:
        assert len(allres.status) == args.clips
        t_total, t_compute, t_gather, host_s, t_ev = [float(x) for x in tt.tolist()]
        print(json.dumps({
            "config": f"configs[4]: {args.clips} synthetic Opencpop-shaped clips, utterance-sharded x{world}, "
                      f"{args.pool_clips}-clip plans over a recycled logits pool, fresh labels per batch",
            "n_gpus": world, "clips": args.clips, "clips_per_rank": n_mine, "batches_per_rank": n_batches,
            "audio_s": float(audio.item()), "audio_s_per_s": round(float(audio.item()) / t_total, 1),
            "wall_s": round(t_total, 3), "compute_phase_s": round(t_compute, 3), "cuda_event_s": round(t_ev, 3),
            "gather_s": round(t_gather, 4), "gathered_labels": int(len(allres.first)),
            "host_enqueue_s_per_rank": round(host_s, 3),
            "host_enqueue_share_of_wall": round(host_s / t_total, 4),
            "ms_per_batch": round(1e3 * t_compute / n_batches, 3), "in_flight": args.depth}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
