#!/usr/bin/env python
"""bench.py -- aligned audio-seconds per second of the alignment decode path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): an Opencpop-test-shaped batch of 2,000 synthetic clips of
5-15 s, head width V = 21129 (fp32 logits, 84.5 KB per 20 ms frame), CTC-flavour decode
(perform_viterbi_ctc) against 2.4 char/s pinyin-class lyrics. The Whisper encoder + GRU head between
K1 and K2 are stock PyTorch and outside the product, so the logits are synthetic and resident.

One "step" = one pass of the WHOLE hot path over the batch, as SURVEY.md 8(d) defines it -- from
"waveforms and logits resident on the device" to "on/offset int32 on the host":
    K1 tcgen05 log-mel of every clip's waveform
 -> label flattening + la_plan_create (the replacement of the reference's per-utterance label strip and
    dp/bt allocation, utils/alignment.py:141-152) -- host work, INSIDE the timed region
 -> K2 fused log-softmax + gather over the head logits -> K3 Viterbi + backtrace
 -> D2H of first / last+1 / score / status -> numpy int32 on the host.
At N > 1 the K steps are followed by ONE NCCL gather of every step's alignments to rank 0, inside the timed region.

  value    : whole-job audio-s/s of that step; CUDA-event timed around the K steps, max over ranks.
  kernels  : the same three kernels timed alone (pre-built plan, nothing but launches), for reference.
  roofline : K2, the dominant kernel: algorithmic bytes / its CUDA-event time inside the timed steps.
  e2e      : the same metric through the REFERENCE-SHAPED call with HOST buffers, one clip per call as
             inference_alignment.py does (default --batch-size 1): log_mel_spectrogram(pinned waveform) +
             perform_viterbi_ctc(pinned cpu logits[1,T,V], labels) -> nested Python lists. H2D of every
             clip's logits and D2H of its alignment are inside the timed region. Two more legs are
             reported beside it: the repo's ragged batch API from host buffers, and the one-line-edit
             drop-in where the caller keeps the logits on the GPU (perform_viterbi_ctc(cuda logits)).
  cpu_baseline / --impl reference : the UNMODIFIED reference decode (utils/alignment.py, vendored to the
             git-ignored oracle/_ref/ by __graft_entry__.build()) + whisper's torch.stft log-mel, on the
             box's host cores, on a bounded sample of the same workload.
With N > 1 (torchrun) every rank owns its own 2,000 clips (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLIPS = 2000
POOL_CLIPS = 100          # pinned host pool for the e2e legs (~4.2 GB), recycled 20x per step
CPU_SAMPLE_CLIPS = 64     # cpu_baseline leg: first 64 clips, 5 passes
REF_STEP_CLIPS = 32       # --impl reference: clips per step


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own decode on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_decode_fn():
    """Returns (fn(pred_cpu[1,T,V] tensor, labels) -> onoff, kind). The unmodified reference
    (utils/alignment.py, from /root/reference or the vendored oracle/_ref/) when it can be imported;
    else the oracle port: the reference's torch CPU emission chain (all intra-op threads, as shipped) +
    the C restatement of its DP."""
    from oracle import ref_shim
    if ref_shim.available():
        ref = ref_shim.load()
        return ref.perform_viterbi_ctc, "reference"
    import torch.nn.functional as F
    import oracle

    def port(pred, labels, hop=0.02):
        lp = F.log_softmax(pred[:, :, 1:-1], dim=2)                   # utils/alignment.py:123
        s = torch.sigmoid(pred[:, :, -1:])                            # :125
        emit = torch.clip(lp + torch.log(1.0 - s), min=-1000)         # :126-132
        blank = torch.clip(torch.log(s), min=-1000)                   # :128,134
        out = []
        for i in range(pred.shape[0]):
            lab = np.array([x for x in labels[i] if x != -100], dtype=np.int64)
            r = oracle.align_one(emit[i].numpy(), blank[i].numpy(), lab)
            assert r["status"] == 0
            out.append([[float(int(a)) * hop, float(int(b)) * hop] for a, b in zip(r["first"], r["last_plus1"])])
        return out
    return port, "port"


def host_threads() -> int:
    """All the host threads the reference could use: torchrun exports OMP_NUM_THREADS=1, which is
    an artefact of the launcher, not of the reference (its torch ops use every core by default)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def time_cpu(pool_pred, pool_labels, pool_dur, n_clips, repeats, pool_wave=None):
    from oracle.logmel import log_mel_spectrogram_torch_f32
    torch.set_num_threads(host_threads())
    fn, kind = cpu_decode_fn()
    fn(pool_pred[0][:, :8], [pool_labels[0][:1].tolist()])            # warm numba / libs
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        for i in range(n_clips):
            if pool_wave is not None:
                log_mel_spectrogram_torch_f32(pool_wave[i])           # whisper's CPU formulation
            fn(pool_pred[i], [pool_labels[i].tolist()])
        times.append(time.perf_counter() - t0)
    secs = float(sum(pool_dur[:n_clips]))
    return secs / statistics.median(times), kind, statistics.median(times)


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=N_CLIPS)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-head", action="store_true", help="skip the fused-head (N1) legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import faulthandler
    faulthandler.dump_traceback_later(900, exit=True)     # a hung collective must not eat the whole time budget
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from lyricalignment_b200 import synth

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference_arm(args, synth)
        return

    t_begin = time.perf_counter()
    import torch.distributed as dist
    import lyricalignment_b200 as la
    from lyricalignment_b200 import _lib, alignment as A, audio as LA, sharded
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # ---- workload: every rank owns its own clips (weak scaling) ----------------------------
    batch = synth.opencpop_shaped(args.clips, seed=114514 + rank)
    V = synth.V_HEAD
    total_T = int(batch.t_len.sum())
    logits = synth.planted_logits(batch, V, ctc=True, device=dev, seed=114514 + rank)
    # K1 inputs/outputs: ragged waveforms -> per-clip [80][F] log-mel (each clip its own call/max)
    wave, w_off = synth.synthetic_waveforms(batch, device=dev, seed=114514 + rank)
    n_samp = batch.n_samples.astype(np.int32)
    mel_frames = (n_samp // 160).astype(np.int32)
    mel_off = np.concatenate([[0], np.cumsum(80 * mel_frames.astype(np.int64))[:-1]]).astype(np.int64)
    mel_out = torch.empty(int(80 * mel_frames.astype(np.int64).sum()), dtype=torch.float32, device=dev)

    launches = [0]

    def step(ev=None):
        """The whole hot path through the public API, device-resident inputs -> int32 results on the host."""
        if ev:
            ev[0].record()
        LA.log_mel_spectrogram_ragged(wave, w_off, n_samp, out=mel_out, out_offsets=mel_off, out_strides=mel_frames)   # K1
        if ev:
            ev[1].record()
        job = A.align_clips_async(logits, batch.t_len, batch.labels, timing=(ev[2], ev[3]) if ev else None)   # labels + plan + K2 + K3 + D2H
        launches[0] = 3 + job.plan.num_launches        # K1: init + logmel + floor; K2; one K3 launch per bucket
        res = job.result()                              # first / last+1 int32, score, status: numpy on the host
        if world > 1:
            pending.append(res)                         # gathered to rank 0 at the end of the job (final_gather)
        return res

    pending = []

    def final_gather():
        """NCCL: every rank's alignments of every step to rank 0 -- the job's ONE exchange (north_star: 'NCCL only
        for the final gather of alignments'), inside the timed region, ONE call (two collectives). Round 1 gathered
        inside every step: ~0.65 ms per step on one N = 8 box, ~3 ms on another (eight processes' NCCL proxy threads
        and Python on 32 vCPUs), and all ranks in lock-step."""
        if pending:                                     # one payload per rank: the K steps' results back to back
            cat = A.AlignResult(*(np.concatenate([getattr(r, f) for r in pending])
                                  for f in ("first", "last_plus1", "score", "status", "l_len")))
            sharded.gather_alignments(cat, device=dev)
        pending.clear()

    def note(msg):
        if rank == 0:
            print(f"[bench +{time.perf_counter() - t_begin:6.1f}s] {msg}", file=sys.stderr, flush=True)
    note("inputs ready, warm-up")
    for _ in range(args.warmup):
        res = step()
    final_gather()
    torch.cuda.synchronize()
    assert int(res.status.max()) == 0, "synthetic clips must all be feasible"
    l_len = res.l_len.astype(np.int64)
    n_labels = int(l_len.sum())

    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t_start.record()
    for k in range(args.steps):
        step(ev[k])
    final_gather()
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    note("timed region done")
    ms_total = t_start.elapsed_time(t_end)
    mel_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    emit_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    audio_s = torch.tensor([batch.audio_seconds], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(audio_s, op=dist.ReduceOp.SUM)
    value = float(audio_s.item()) * args.steps / (ms_total / 1e3)
    step_ms = ms_total / args.steps
    # per-rank diagnostics (outside the timed region): the job ends in a gather, so it runs at the pace of the slowest
    # rank -- say which one that was and by how much (K1 start -> K2 end on each rank's own GPU, own wall of the loop)
    per_rank = None
    if world > 1:
        mine = torch.tensor([statistics.mean(e[0].elapsed_time(e[3]) for e in ev), t_start.elapsed_time(t_end) / args.steps],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"k1_to_k2_end_ms": [round(float(x[0]), 3) for x in allr], "step_ms": [round(float(x[1]), 3) for x in allr]}

    # ---- the three kernels alone (pre-built plan, launches only): explains `value`, is not `value` ----
    lens, cols = A._resolve_columns(A._flatten_labels(batch.labels), V - 2)
    plan = A.AlignPlan(A.MODE_CTC, V, batch.t_len, lens, cols, local_rank)
    ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
    first = torch.empty(plan.total_labels, dtype=torch.int32, device=dev)
    last = torch.empty_like(first)
    score = torch.empty(plan.n_utt, dtype=torch.float64, device=dev)
    status = torch.empty(plan.n_utt, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    mel_ws = torch.empty(int(lib.la_logmel_workspace_bytes(len(n_samp), int(n_samp.astype(np.int64).sum()))),
                         dtype=torch.uint8, device=dev)
    kev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(5)]
    for k in range(-2, 5):
        e = kev[max(k, 0)]
        e[0].record()
        _lib.check(lib.la_logmel_ragged(wave.data_ptr(), len(n_samp), w_off.ctypes.data, n_samp.ctypes.data,
                                        mel_out.data_ptr(), mel_off.ctypes.data, mel_frames.ctypes.data,
                                        mel_ws.data_ptr(), stream), "la_logmel_ragged")
        e[1].record()
        _lib.check(lib.la_emit(plan.handle, logits.data_ptr(), V, None, 0, ws.data_ptr(), stream), "la_emit")
        e[2].record()
        _lib.check(lib.la_viterbi(plan.handle, ws.data_ptr(), first.data_ptr(), last.data_ptr(),
                                  score.data_ptr(), status.data_ptr(), stream), "la_viterbi")
        e[3].record()
    torch.cuda.synchronize()
    k1_alone = statistics.mean(e[0].elapsed_time(e[1]) for e in kev)
    k2_alone = statistics.mean(e[1].elapsed_time(e[2]) for e in kev)
    k3_alone = statistics.mean(e[2].elapsed_time(e[3]) for e in kev)
    plan.close()

    # ---- roofline of the dominant kernel (K2), from the events inside the timed steps -------------
    peak, peak_src = measured_peaks()
    tl64 = batch.t_len.astype(np.int64)
    emis_bytes = 4.0 * float(np.sum(tl64 * (l_len + 1)))
    algo_bytes = 4.0 * total_T * V + emis_bytes
    achieved = algo_bytes / (emit_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_frame") * total_T   # ncu, per frame, scaled to this launch
        except Exception:
            traffic = None
    roofline = {"kernel": "la::emit_kernel<CTC> (K2 fused log-softmax + gather)", "bound": "hbm",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": round(emit_ms, 4)}
    # K1: tensor-nominal. 6 fp16 MMAs x 13 k-steps x 2 passes per 128-frame tile, M128 N208 K16.
    tiles = float(np.sum((mel_frames + 127) // 128))
    mel_flops = 2.0 * 128 * 208 * 16 * 6 * 26 * tiles
    mel_bytes = 4.0 * float(n_samp.astype(np.int64).sum()) + 4.0 * 80 * float(mel_frames.astype(np.int64).sum())
    # K3: 4 T (L+1) emissions read + ceil(S/4) T backpointer bytes written + read back + 8 L out.
    k3_bytes = emis_bytes + 2.0 * float(np.sum(tl64 * ((2 * l_len + 1 + 3) // 4))) + 8.0 * n_labels
    kernels = {"k1_logmel_ms": round(mel_ms, 4), "k2_emit_ms": round(emit_ms, 4),
               "host_plan_k3_d2h_ms": round(step_ms - mel_ms - emit_ms, 4), "per_rank": per_rank,
               "alone": {"k1_logmel_ms": round(k1_alone, 4), "k2_emit_ms": round(k2_alone, 4),
                         "k3_viterbi_ms": round(k3_alone, 4),
                         "sum_ms": round(k1_alone + k2_alone + k3_alone, 4),
                         "audio_s_per_s": round(batch.audio_seconds / ((k1_alone + k2_alone + k3_alone) / 1e3), 1)},
               "host_and_copy_share_of_step": round(max(0.0, 1.0 - (k1_alone + k2_alone + k3_alone) / step_ms), 4),
               "k1_f16_tflops": round(mel_flops / (k1_alone / 1e3) / 1e12, 1),
               "k1_algorithmic_gbs": round(mel_bytes / (k1_alone / 1e3) / 1e9, 1),
               "k3_algorithmic_bytes": k3_bytes,
               "k3_algorithmic_gbs": round(k3_bytes / (k3_alone / 1e3) / 1e9, 1),
               "k3_frac_of_hbm_peak": round(k3_bytes / (k3_alone / 1e3) / 1e9 / peak, 4),
               "k3_note": "latency-bound dependent chain (one fp64 add per frame on the critical path), not a streaming kernel"}

    # ---- e2e + cpu baseline on rank 0's pinned pool ----------------------------------------
    e2e, cpu = None, None
    pool_n = min(POOL_CLIPS, args.clips)
    pool_rows = int(batch.t_len[:pool_n].sum())
    if not args.skip_e2e:
        host = torch.empty((pool_rows, V), dtype=torch.float32).pin_memory()
        host.copy_(logits[:pool_rows])
        torch.cuda.synchronize()
        offs = np.concatenate([[0], np.cumsum(batch.t_len[:pool_n])])
        pool_pred = [host[offs[i]:offs[i + 1]].unsqueeze(0) for i in range(pool_n)]
        pool_pred_dev = [logits[offs[i]:offs[i + 1]].unsqueeze(0) for i in range(pool_n)]
        n_calls = args.clips
        lab_sets = [synth.opencpop_shaped(pool_n, seed=7000 + 31 * rank + j).labels for j in range((n_calls + pool_n - 1) // pool_n)]

        # labels drawn for other durations may be too long for this clip: keep them feasible
        def lab_for(i):
            lab = lab_sets[i // pool_n][i % pool_n]
            return lab[:max(1, int(batch.t_len[i % pool_n]) // 3)]

        wave_host = wave.cpu().pin_memory()
        pool_wave = [wave_host[w_off[i]:w_off[i] + int(n_samp[i])] for i in range(pool_n)]
        pool_wave_end = int(w_off[pool_n - 1] + n_samp[pool_n - 1])

        # (a) HEADLINE e2e: the literal drop-in loop of inference_alignment.py (default --batch-size 1)
        def dropin_step():
            tot = 0
            for i in range(n_calls):
                j = i % pool_n
                LA.log_mel_spectrogram(pool_wave[j])                               # B1: host waveform -> device log-mel
                out = la.perform_viterbi_ctc(pool_pred[j], [lab_for(i).tolist()])   # B2: host logits -> on/offsets (lists)
                tot += len(out[0])
            return tot

        # (b) the repo's ragged batch API, one call per 100-clip batch, HOST buffers in and out
        def ragged_step():
            tot = 0
            for b0 in range(0, n_calls, pool_n):
                nb = min(pool_n, n_calls - b0)
                LA.log_mel_spectrogram_ragged(wave_host[:pool_wave_end], w_off[:nb], n_samp[:nb])
                r = la.align_clips(host[:int(offs[nb])], batch.t_len[:nb], [lab_for(b0 + i) for i in range(nb)],
                                   staging_bytes=int(os.environ.get("BENCH_STAGING_MB", "0")) << 20)
                tot += sum(len(u) for u in la.onoff_seconds(r))
            return tot

        # (c) the one-line-edit drop-in: the caller drops its `.cpu()` (inference_alignment.py:161), logits stay on the GPU
        def cuda_dropin_step():
            tot = 0
            for i in range(n_calls):
                j = i % pool_n
                LA.log_mel_spectrogram(pool_wave[j])
                out = la.perform_viterbi_ctc(pool_pred_dev[j], [lab_for(i).tolist()])
                tot += len(out[0])
            return tot

        note("e2e warm-up")
        dropin_step(); ragged_step(); cuda_dropin_step()

        def timed(fn, reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                n = fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item()) / reps, n
        note("e2e timing")
        dt_dropin, n_lab = timed(dropin_step, args.e2e_steps)
        dt_ragged, _ = timed(ragged_step, 1)
        dt_cuda, _ = timed(cuda_dropin_step, 1)
        e2e_audio = float(sum(batch.durations[i % pool_n] for i in range(n_calls))) * world
        wave_bytes = 4.0 * float(sum(int(n_samp[i % pool_n]) for i in range(n_calls)))
        h2d = float(sum(int(batch.t_len[i % pool_n]) for i in range(n_calls))) * V * 4 + wave_bytes
        d2h = float(n_lab * 8 + n_calls * 12)
        e2e = {"value": round(e2e_audio / dt_dropin, 1), "unit": "audio-s/s",
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": args.e2e_steps,
               "call": "per clip, as inference_alignment.py with --batch-size 1: audio.log_mel_spectrogram(pinned "
                       "waveform) + perform_viterbi_ctc(pinned cpu logits[1,T,V], labels) -> nested Python lists",
               "limiter": f"PCIe: {h2d / 1e9:.1f} GB of fp32 logits per step per GPU at {h2d / dt_dropin / 1e9:.1f} GB/s",
               "ragged_api": {"value": round(e2e_audio / dt_ragged, 1), "unit": "audio-s/s",
                              "call": f"repo-only API, per {pool_n}-clip batch: audio.log_mel_spectrogram_ragged(pinned waveforms)"
                                      " + align_clips(pinned cpu logits[sumT,V], t_len, labels) -> onoff_seconds()",
                              "h2d_bytes_per_step": h2d * world},
               "cuda_logits_dropin": {"value": round(e2e_audio / dt_cuda, 1), "unit": "audio-s/s",
                                      "call": "per clip, entry script with its `.cpu()` dropped: log_mel_spectrogram(pinned "
                                              "waveform) + perform_viterbi_ctc(CUDA logits[1,T,V], labels) -> nested Python lists",
                                      "h2d_bytes_per_step": wave_bytes * world}}
        if rank == 0 and world == 1 and not args.skip_cpu:
            note("cpu baseline")
            n_cpu = min(CPU_SAMPLE_CLIPS, pool_n)
            v, kind, secs = time_cpu(pool_pred, batch.labels, batch.durations, n_cpu, 5, pool_wave)
            cpu = {"value": round(v, 1), "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": kind,
                   "sample": f"first {n_cpu} clips of the workload ({sum(batch.durations[:n_cpu]):.0f} audio-s), same logits "
                             f"(host copies), median of 5 passes, {secs:.2f} s per pass, host cpu_count={os.cpu_count()}"}

    # ---- N1: the head's Linear fused in -- hidden states [sum T][768] instead of logits [sum T][21129] -------------
    head_leg = None
    if not args.skip_head:
        from lyricalignment_b200.head import FusedHead
        note("fused head (N1)")
        D = 768
        torch.manual_seed(114514 + rank)
        fc = torch.nn.Linear(D, V).to(dev)                                    # random-init head (module/align_model.py:32-33)
        hidden = torch.nn.functional.mish(torch.randn(total_T, D, device=dev))  # what the Linear sees (:38)
        fh = FusedHead(fc.weight, fc.bias, device=dev)
        hev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

        def head_step(x, timing=None):
            job = fh.align_clips_async(x, batch.t_len, batch.labels, timing=timing)
            r = job.result()
            job.close()
            return r
        head_step(hidden)                                                      # warm-up
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            r = head_step(hidden, hev)
        b.record()
        torch.cuda.synchronize()
        assert int(r.status.max()) == 0
        ms_dev = a.elapsed_time(b) / reps
        ms_emit = hev[0].elapsed_time(hev[1])
        flops = 2.0 * total_T * V * D
        tpeak = None
        try:
            tpeak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            pass
        # host hidden states: 3 KB per frame over PCIe; 200-clip sub-batches, two jobs in flight (copy || GEMM)
        hidden_host = hidden.cpu().pin_memory()
        sub = 200
        t_off = np.concatenate([[0], np.cumsum(batch.t_len)])

        def head_host_step():
            jobs, tot = [], 0
            for b0 in range(0, args.clips, sub):
                b1 = min(args.clips, b0 + sub)
                jobs.append(fh.align_clips_async(hidden_host[int(t_off[b0]):int(t_off[b1])], batch.t_len[b0:b1], batch.labels[b0:b1]))
                if len(jobs) >= 2:
                    j = jobs.pop(0); tot += len(j.result().first); j.close()
            for j in jobs:
                tot += len(j.result().first); j.close()
            return tot

        # per clip, reference-shaped but with the hidden state instead of the logits (what a caller that adopts
        # FusedHead writes instead of `fc(...)` + `.cpu()` + perform_viterbi_ctc)
        wave_host2 = wave.cpu().pin_memory() if args.skip_e2e else wave_host
        def head_dropin_step():
            tot = 0
            for i in range(args.clips):
                LA.log_mel_spectrogram(wave_host2[w_off[i]:w_off[i] + int(n_samp[i])])
                out = fh.perform_viterbi_ctc(hidden_host[int(t_off[i]):int(t_off[i + 1])].unsqueeze(0), [batch.labels[i]])
                tot += len(out[0])
            return tot
        head_host_step(); head_dropin_step()

        def timed_wall(fn):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())
        dt_host = timed_wall(head_host_step)
        dt_drop = timed_wall(head_dropin_step)
        tot_audio = float(audio_s.item())
        head_leg = {
            "what": "N1: Linear(768 -> 21129) + log-softmax + gather fused (logits never materialised) + K3; inputs are the "
                    "Mish output [sum T][768] (random-init head, synthetic hidden states)",
            "device_resident": {"value": round(tot_audio / (ms_dev / 1e3), 1), "unit": "audio-s/s", "ms_per_step": round(ms_dev, 3),
                                "head_emit_ms": round(ms_emit, 3)},
            "roofline": {"kernel": "la::head_lse_kernel (split-fp16 tcgen05 GEMM + online LSE) incl. pack + gather", "bound": "tensor",
                         "achieved": round(3 * flops / (ms_emit / 1e3) / 1e12, 1), "peak": tpeak, "unit": "TFLOP/s",
                         "frac": round(3 * flops / (ms_emit / 1e3) / 1e12 / tpeak, 4) if tpeak else None,
                         "useful_tflops": round(flops / (ms_emit / 1e3) / 1e12, 1),
                         "note": "achieved = issued fp16 MMA FLOPs (3 per useful one: hi*hi + hi*lo + lo*hi) / time of "
                                 "pack + GEMM + gather; peak = MEASURED_PEAKS.json bf16_tflops_sustained"},
            "e2e_host_hidden_batched": {"value": round(tot_audio / dt_host, 1), "unit": "audio-s/s",
                                        "h2d_bytes_per_step": 4.0 * total_T * D * world,
                                        "call": f"per {sub}-clip sub-batch, two in flight: FusedHead.align_clips_async(pinned host hidden "
                                                "[sumT,768], t_len, labels).result()"},
            "e2e_host_hidden_per_clip": {"value": round(tot_audio / dt_drop, 1), "unit": "audio-s/s",
                                         "h2d_bytes_per_step": (4.0 * total_T * D + 4.0 * float(n_samp.astype(np.int64).sum())) * world,
                                         "call": "per clip: log_mel_spectrogram(pinned waveform) + FusedHead.perform_viterbi_ctc(pinned "
                                                 "host hidden[1,T,768], labels) -> nested Python lists"}}
        del hidden, hidden_host, fh

    note("done")
    faulthandler.cancel_dump_traceback_later()
    if rank == 0:
        line = {
            "metric": "aligned audio-sec/sec (alignment decode path)", "value": round(value, 1), "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 emissions / f64 DP (K1: fp16 operand slices, f32 accumulate)", "data": "synthetic",
            "config": {"workload": f"configs[1]: Opencpop-test-shaped batch, {args.clips} clips of 5-15 s per GPU, "
                                   f"V=21129 CTC decode, {n_labels} syllables, {total_T} frames",
                       "timed_region": "K1 -> label flattening + la_plan_create -> K2 -> K3 -> D2H -> int32 numpy on the host"
                                       + (" (x steps) -> ONE NCCL gather of every step's alignments to rank 0, inside the timed region" if world > 1 else "") + " (public API: "
                                       "log_mel_spectrogram_ragged + align_clips_async().result())",
                       "l2_policy": f"inputs larger than L2 ({4.0 * total_T * V / 1e9:.1f} GB of logits resident in HBM per GPU)",
                       "parallelism": f"utterance-sharded x{world}" + (", one NCCL gather of all alignments at the end of the timed region" if world > 1 else "")},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "fused_head": head_leg,
            "gpu_launches": launches[0] * args.steps, "clocks": clk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args, synth):
    """--impl reference: the reference's CPU decode on the host cores, bounded sample per step."""
    torch.manual_seed(0)
    torch.set_num_threads(host_threads())
    n = REF_STEP_CLIPS
    batch = synth.opencpop_shaped(args.clips, seed=114514)
    sub = synth.ClipBatch(batch.durations[:n], batch.n_samples[:n], batch.t_len[:n], batch.labels[:n])
    pred = synth.planted_logits(sub, synth.V_HEAD, ctc=True, device="cpu", seed=114514)
    offs = np.concatenate([[0], np.cumsum(sub.t_len)])
    pool = [pred[offs[i]:offs[i + 1]].unsqueeze(0) for i in range(n)]
    from oracle.logmel import log_mel_spectrogram_torch_f32
    wave, w_off = synth.synthetic_waveforms(sub, device="cpu", seed=114514)
    waves = [wave[w_off[i]:w_off[i] + int(sub.n_samples[i])] for i in range(n)]
    fn, kind = cpu_decode_fn()
    fn(pool[0][:, :8], [sub.labels[0][:1].tolist()])

    def step():
        for i in range(n):
            log_mel_spectrogram_torch_f32(waves[i])                   # whisper's CPU log-mel (align_model.py:84)
            fn(pool[i], [sub.labels[i].tolist()])                     # utils/alignment.py decode
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sub.audio_seconds * args.steps / dt
    sample = (f"each step = first {n} clips of the workload ({sub.audio_seconds:.0f} audio-s), one "
              f"perform_viterbi_ctc call per clip as inference_alignment.py does; host cpu_count={os.cpu_count()}")
    cpu = {"value": round(value, 1), "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample}
    print(json.dumps({
        "impl": "reference", "metric": "aligned audio-sec/sec (alignment decode path)", "value": round(value, 1),
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 emissions / f64 DP", "data": "synthetic",
        "config": {"workload": f"configs[1]: Opencpop-test-shaped batch, {args.clips} clips of 5-15 s, V=21129 CTC decode",
                   "sample": sample},
        "cpu_baseline": cpu,
        "e2e": {"value": round(value, 1), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    main()
