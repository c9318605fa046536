"""Sharded alignment runner: the reference's `align_and_evaluate` loop (inference_alignment.py:127-180,
inference_alignment_nogt.py:131-178) over the GPUs of one box, one process per GPU.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m lyricalignment_b200.run_sharded \\
        -f test.json --pronounce-table bert_base_chinese_pronunce_table.json --use-ctc-loss --out alignments.json

What it does, in the reference's order:
  io.read_data(-f)                                  data_processor/record.py:22-39 (same JSON schema)
  batches of --batch-size CONSECUTIVE records       DataLoader(shuffle=False), inference_alignment.py:199-208
  whole batches dealt to ranks, contiguously        (the front end's global max couples a batch)
  per batch: tokens -> pinyin classes (LUT gather)  inference_alignment.py:149-152
             frame_manual_forward(audios)           module/align_model.py:72-123 (K1 + stock encoder/head)
             perform_viterbi_ctc | perform_viterbi  inference_alignment.py:163-166 (K2 + K3, logits stay on the GPU)
             get_mae(gt, prediction)                inference_alignment.py:168
  ragged gather of all alignments to rank 0         (NCCL; gloo on CPU in the tests) -- the ONLY collective
  rank 0: "Average MAE: x" = unweighted mean of the per-batch MAEs summed in DATASET order in Python fp64
          (inference_alignment.py:172-178), per-record `[[on, off, char], ...]` lines when there is no ground
          truth (inference_alignment_nogt.py:175-176), and the machine-readable alignment file (io.write_alignments).

What the offline image cannot supply is injected: the checkpoint / whisper weights (``--encoder-size``
builds the random-init stock stand-in of pipeline.py), the BERT tokenizer (records may carry a "tokens" list
of BERT ids; otherwise transformers' bert-base-chinese is loaded from the local cache) and librosa
(``song_path`` may be a .npy waveform at 16 kHz). The library entry point `run()` takes these as callables,
which is also how the tests drive it with fake logits.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import alignment as A
from . import io as la_io
from . import sharded
from .labels import load_pinyin_lut, remap_tokens_


def batches_of(n_records: int, batch_size: int) -> List[range]:
    return [range(s, min(s + batch_size, n_records)) for s in range(0, n_records, batch_size)]


def pad_tokens(rows: Sequence[Sequence[int]]) -> torch.Tensor:
    """dataset.py:212-232 leaves a LongTensor [B, Lmax] padded with -100."""
    Lmax = max((len(r) for r in rows), default=0)
    out = torch.full((len(rows), max(Lmax, 1)), -100, dtype=torch.long)
    for i, r in enumerate(rows):
        if len(r):
            out[i, :len(r)] = torch.as_tensor(list(r), dtype=torch.long)
    return out


def run(records: Sequence[la_io.Record], audio_fn: Callable[[la_io.Record], np.ndarray],
        tokens_fn: Callable[[la_io.Record], Sequence[int]], logits_fn: Callable[[List[np.ndarray]], torch.Tensor],
        lut: Optional[torch.Tensor], use_ctc_loss: bool, batch_size: int = 1, device: Optional[torch.device] = None,
        group=None, hop_size_second: float = 0.02, align_fn: Callable = A.align):
    """Runs this rank's share and gathers. Returns on rank 0:
    {"alignments": per-record [[on, off], ...] in dataset order, "batch_mae": per-batch MAE or None,
     "average_mae": float or None}; None on the other ranks. Works without torch.distributed (world 1).
    `align_fn(logits, tokens, mode) -> AlignResult` is alignment.align (K2 + K3); the CPU tests of this host
    logic substitute a checker-backed stand-in, the product never does."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    batches = batches_of(len(records), batch_size)
    lo, hi = sharded.shard_bounds(len(batches), world, rank)
    firsts, lasts, scores, stats, lens, maes = [], [], [], [], [], []
    for b in batches[lo:hi]:
        recs = [records[i] for i in b]
        tokens = pad_tokens([tokens_fn(r) for r in recs])
        if lut is not None:
            remap_tokens_(tokens, lut)                                   # inference_alignment.py:149-152
        logits = logits_fn([audio_fn(r) for r in recs])                  # [B, T, V], stays on its device
        res = align_fn(logits, tokens, A.MODE_CTC if use_ctc_loss else A.MODE_CE)   # == perform_viterbi_ctc | perform_viterbi
        onoff = A.onoff_seconds(res, hop_size_second)                    # raises like the reference does
        firsts.append(res.first); lasts.append(res.last_plus1); scores.append(res.score)
        stats.append(res.status); lens.append(res.l_len)
        gts = [r.lyric_onset_offset for r in recs]
        maes.append(A.get_mae(gts, onoff) if all(g is not None for g in gts) else None)   # :156-157, :168
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    mine = A.AlignResult(cat(firsts, np.int32), cat(lasts, np.int32), cat(scores, np.float64), cat(stats, np.int32),
                         cat(lens, np.int32))
    if world > 1:
        allres = sharded.gather_alignments(mine, device=device or torch.device("cpu"), dst=0, group=group)
        all_maes: list = [None] * world
        dist.all_gather_object(all_maes, maes, group=group)
        maes = [m for part in all_maes for m in part]                    # contiguous shards: rank order == dataset order
    else:
        allres = mine
    if rank != 0:
        return None
    alignments = A.onoff_seconds(allres, hop_size_second)
    scored = [m for m in maes if m is not None]
    return {"alignments": alignments, "batch_mae": maes,
            "average_mae": sharded.average_mae_in_dataset_order(scored) if scored else None}


# ----------------------------------------------------------------------------------------------
# CLI wiring of what the offline image lacks
# ----------------------------------------------------------------------------------------------
def _audio_from_path(rec: la_io.Record) -> np.ndarray:
    if rec.audio_path.endswith(".npy"):
        return np.load(rec.audio_path).astype(np.float32).reshape(-1)
    try:
        import librosa                                                   # utils/audio.py:3-20
    except ImportError as e:
        raise RuntimeError("librosa is not installed: give 16 kHz mono waveforms as .npy files") from e
    return librosa.load(rec.audio_path, sr=16000)[0].astype(np.float32)


def _make_tokens_fn(raw_items):
    by_path = {d["song_path"]: d.get("tokens") for d in raw_items}
    tok = None

    def fn(rec: la_io.Record):
        nonlocal tok
        t = by_path.get(rec.audio_path)
        if t is not None:
            return t
        if tok is None:
            from transformers import AutoTokenizer                        # inference_alignment.py:98
            tok = AutoTokenizer.from_pretrained("bert-base-chinese", local_files_only=True)
        ids = tok(rec.text)["input_ids"][1:]                             # dataset.py:222-226: strip [CLS] ...
        return [i for i in ids if i not in (0, 102)]                     # ... [PAD] / [SEP] become -100 (dropped here)
    return fn


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-f", "--test-data", required=True)
    ap.add_argument("--pronounce-table", default="bert_base_chinese_pronunce_table.json")
    ap.add_argument("--use-ctc-loss", action="store_true")
    ap.add_argument("--batch-size", type=int, default=1)
    ap.add_argument("--encoder-size", default="tiny", help="random-init stock encoder/head stand-in (no checkpoints offline)")
    ap.add_argument("--seed", type=int, default=114514)
    ap.add_argument("--out", default=None, help="write the gathered alignments (io.write_alignments)")
    args = ap.parse_args(argv)

    distributed = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("lyricalignment_b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    records = la_io.read_data(args.test_data)
    with open(args.test_data) as f:
        raw = json.load(f)
    lut = load_pinyin_lut(args.pronounce_table) if os.path.exists(args.pronounce_table) else None
    from .pipeline import AlignPipeline
    pipe = AlignPipeline(args.encoder_size, device=dev, seed=args.seed)
    out = run(records, _audio_from_path, _make_tokens_fn(raw), pipe.frame_manual_forward, lut, args.use_ctc_loss,
              args.batch_size, device=dev)
    if out is not None:
        if out["average_mae"] is not None:
            print("Average MAE:", out["average_mae"])                    # inference_alignment.py:178
        else:
            for rec, al in zip(records, out["alignments"]):
                print(la_io.format_prediction(al, rec.text))              # inference_alignment_nogt.py:175-176
        if args.out:
            per_rec_mae = None
            la_io.write_alignments(args.out, records, out["alignments"], per_rec_mae)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
