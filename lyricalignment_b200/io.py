"""Input / result formats either side of the decode path.

Inputs follow the reference's dataset schema (data_processor/record.py:8-39): a JSON list of
``{"song_path": str, "lyric": str, "on_offset": [[on, off], ...] (optional)}``. The reference only
PRINTS its alignments (inference_alignment_nogt.py:175-176: ``[[onset, offset, char], ...]`` per
record; inference_alignment.py:178: ``Average MAE: x``); the sharded runner needs something a
process can gather and write, so the same content gets a machine-readable form here.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence


@dataclass
class Record:
    """Same fields as the reference's Record (data_processor/record.py:8-19)."""
    audio_path: str
    text: str
    lyric_onset_offset: Optional[list] = None


def read_data(data_path: str) -> List[Record]:
    """data_processor/record.py:22-39."""
    if not os.path.exists(data_path):
        raise AssertionError(data_path)
    with open(data_path, "r") as f:
        items = json.load(f)
    out = []
    for d in items:
        rec = Record(audio_path=d["song_path"], text=d["lyric"])
        if "on_offset" in d:
            rec.lyric_onset_offset = d["on_offset"]
        out.append(rec)
    return out


def prediction_rows(onoff: Sequence[Sequence[float]], text: str) -> list:
    """One record's ``[[onset, offset, char], ...]`` exactly as inference_alignment_nogt.py:175 builds it."""
    return [[onoff[i][0], onoff[i][1], text[i]] for i in range(len(onoff))]


def format_prediction(onoff, text: str) -> str:
    """The line the reference prints for a record (Python repr of the list)."""
    return str(prediction_rows(onoff, text))


def write_alignments(path: str, records: Sequence[Record], alignments: Sequence[Sequence[Sequence[float]]],
                     mae: Optional[Sequence[float]] = None) -> None:
    """JSON list mirroring the input schema: ``song_path``, ``lyric``, ``on_offset`` = the PREDICTED
    [[onset, offset], ...] (so the file can be fed back as a dataset), plus ``mae`` where ground truth
    was available."""
    out = []
    for i, (rec, al) in enumerate(zip(records, alignments)):
        item = {"song_path": rec.audio_path, "lyric": rec.text, "on_offset": [[float(a), float(b)] for a, b in al]}
        if mae is not None and mae[i] is not None:
            item["mae"] = float(mae[i])
        out.append(item)
    with open(path, "w") as f:
        json.dump(out, f, ensure_ascii=False)
