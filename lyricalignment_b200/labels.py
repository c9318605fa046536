"""Label remap on the caller side of the decode path.

The reference maps BERT token ids to pinyin-class ids with a double Python loop over tensor
elements (inference_alignment.py:149-152, inference_alignment_nogt.py:165-168):

    tokens[i][j] = pinyin_lookup_table[token_pinyin[tokens[i][j]]]      (for tokens[i][j] != -100)

Here the two tables are folded once into a 21 128-entry LUT and the remap is one gather.
"""
from __future__ import annotations

import json
from typing import Dict, Sequence

import torch


def build_pinyin_lut(token_pinyin: Sequence[str], pinyin_lookup_table: Dict[str, int]) -> torch.Tensor:
    """lut[token_id] = class id (1..402); the tables are the 1st and 3rd items of
    bert_base_chinese_pronunce_table.json (get_pronunce_table.py:41-47)."""
    return torch.tensor([int(pinyin_lookup_table[p]) for p in token_pinyin], dtype=torch.long)


def load_pinyin_lut(path: str = "bert_base_chinese_pronunce_table.json") -> torch.Tensor:
    with open(path, "r") as f:
        token_pinyin, _pinyin_reverse, pinyin_lookup_table = json.load(f)
    return build_pinyin_lut(token_pinyin, pinyin_lookup_table)


def remap_tokens_(tokens: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """In place, like the reference loop: every entry that is not the -100 padding becomes its class id."""
    keep = tokens != -100
    lut = lut.to(tokens.device)
    tokens[keep] = lut[tokens[keep]]
    return tokens
