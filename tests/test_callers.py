"""CPU: the caller-side rows of the path -- label remap LUT (A10) and the unmodified-entry-script
runner (names re-pointed before the script runs)."""
import json
import os
import sys
import textwrap

import numpy as np
import pytest
import torch

from lyricalignment_b200 import labels as L


def _reference_loop(tokens, token_pinyin, pinyin_lookup_table):
    """inference_alignment.py:149-152, verbatim in meaning."""
    for i in range(len(tokens)):
        for j in range(len(tokens[i])):
            if tokens[i][j] != -100:
                tokens[i][j] = pinyin_lookup_table[token_pinyin[tokens[i][j]]]
    return tokens


def test_remap_matches_reference_loop_synthetic_table():
    rng = np.random.default_rng(0)
    names = ["bad"] + [f"p{i}" for i in range(1, 402)]
    table = {n: i + 1 for i, n in enumerate(names)}
    token_pinyin = [names[int(k)] for k in rng.integers(0, 402, size=21128)]
    tokens = torch.from_numpy(rng.integers(0, 21128, size=(5, 17)))
    tokens[0, 10:] = -100
    tokens[3, 1:] = -100
    want = _reference_loop(tokens.clone(), token_pinyin, table)
    got = L.remap_tokens_(tokens.clone(), L.build_pinyin_lut(token_pinyin, table))
    assert torch.equal(got, want) and got.dtype == torch.long


@pytest.mark.skipif(not os.path.exists("/root/reference/bert_base_chinese_pronunce_table.json"),
                    reason="reference tree not mounted")
def test_remap_with_the_real_table():
    path = "/root/reference/bert_base_chinese_pronunce_table.json"
    token_pinyin, _, table = json.load(open(path))
    lut = L.load_pinyin_lut(path)
    assert lut.shape == (21128,) and int(lut.min()) == 1 and int(lut.max()) == 402
    rng = np.random.default_rng(1)
    tokens = torch.from_numpy(rng.integers(0, 21128, size=(3, 40)))
    tokens[1, 25:] = -100
    assert torch.equal(L.remap_tokens_(tokens.clone(), lut), _reference_loop(tokens.clone(), token_pinyin, table))


def test_runner_repoints_names_before_the_script_runs(tmp_path, monkeypatch, capsys):
    # a miniature "reference tree": utils/alignment.py + an entry script importing from it
    (tmp_path / "utils").mkdir()
    (tmp_path / "utils" / "__init__.py").write_text("")
    (tmp_path / "utils" / "alignment.py").write_text(textwrap.dedent("""
        def perform_viterbi(prediction, labels, hop_size_second=0.02): return "cpu"
        def perform_viterbi_ctc(prediction, labels, hop_size_second=0.02): return "cpu"
        def get_mae(gt, predict): return -1.0
    """))
    (tmp_path / "entry.py").write_text(textwrap.dedent("""
        import sys
        from utils.alignment import perform_viterbi, perform_viterbi_ctc, get_mae
        print("ARGS", sys.argv[1:])
        print("MOD", perform_viterbi_ctc.__module__, perform_viterbi.__module__, get_mae.__module__)
        print("MAE", get_mae([[[0, .5], [.5, 1]]], [[[.02, .48], [.5, 1.02]]]))
    """))
    for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", list(sys.argv))
    monkeypatch.setattr(sys, "path", list(sys.path))
    from lyricalignment_b200 import run_reference
    assert run_reference.main([str(tmp_path / "entry.py"), "-f", "x.json", "--use-ctc-loss"]) == 0
    out = capsys.readouterr().out
    assert "ARGS ['-f', 'x.json', '--use-ctc-loss']" in out
    assert out.count("lyricalignment_b200.alignment") == 3
    assert "MAE 0.01500000000000001" in out


def test_io_schema_roundtrip(tmp_path):
    from lyricalignment_b200 import io as lio
    src = [{"song_path": "/a/1.wav", "lyric": "你好吗", "on_offset": [[0.0, 0.5], [0.5, 1.0], [1.0, 1.5]]},
           {"song_path": "/a/2.wav", "lyric": "好"}]
    p = tmp_path / "d.json"
    p.write_text(json.dumps(src, ensure_ascii=False))
    recs = lio.read_data(str(p))
    assert recs[0].text == "你好吗" and recs[0].lyric_onset_offset == src[0]["on_offset"] and recs[1].lyric_onset_offset is None
    pred = [[[0.02, 0.48], [0.48, 1.0], [1.0, 1.52]], [[0.1, 0.12]]]
    assert lio.format_prediction(pred[0], recs[0].text) == str([[0.02, 0.48, "你"], [0.48, 1.0, "好"], [1.0, 1.52, "吗"]])
    out = tmp_path / "o.json"
    lio.write_alignments(str(out), recs, pred, mae=[0.013, None])
    back = lio.read_data(str(out))
    assert back[0].lyric_onset_offset == pred[0] and back[1].text == "好"
    with pytest.raises(AssertionError):
        lio.read_data(str(tmp_path / "missing.json"))


@pytest.mark.skipif(not os.path.exists("/root/reference/inference_alignment.py"), reason="reference tree not mounted")
def test_runner_repoints_the_real_entry_scripts_import_graph(monkeypatch):
    """run_reference.install() against the REAL inference_alignment.py / inference_alignment_nogt.py: their whole
    import graph (module.align_model, dataset, utils.audio, ...) is imported with the packages this image lacks
    (whisper, librosa, pypinyin) stubbed, and the names the scripts bound at import time must be the CUDA ones."""
    import importlib
    import types
    for m in [k for k in sys.modules if k.split(".")[0] in ("utils", "module", "dataset", "data_processor", "whisper",
                                                             "librosa", "inference_alignment", "inference_alignment_nogt")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.setattr(sys, "path", ["/root/reference"] + list(sys.path))

    import importlib.machinery
    import transformers  # noqa: F401  (its lazy-module probes for optional packages must run before the stubs exist)

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        mod.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, mod)
        return mod
    sentinel_lm = lambda *a, **k: "whisper-cpu-logmel"
    wa = stub("whisper.audio", N_FRAMES=3000, pad_or_trim=lambda *a, **k: "whisper-pad", log_mel_spectrogram=sentinel_lm)
    wt = stub("whisper.tokenizer", Tokenizer=object, get_tokenizer=lambda *a, **k: None)
    stub("whisper", audio=wa, tokenizer=wt, load_model=lambda *a, **k: None, log_mel_spectrogram=sentinel_lm,
         pad_or_trim=wa.pad_or_trim, Whisper=object, DecodingOptions=object)
    stub("librosa", load=lambda *a, **k: (np.zeros(1, np.float32), 16000))
    from lyricalignment_b200 import alignment as la_align, audio as la_audio, run_reference
    run_reference.install("/root/reference")
    for script in ("inference_alignment", "inference_alignment_nogt"):
        mod = importlib.import_module(script)
        assert mod.perform_viterbi_ctc is la_align.perform_viterbi_ctc, script
        assert mod.perform_viterbi is la_align.perform_viterbi, script
        if hasattr(mod, "get_mae"):
            assert mod.get_mae is la_align.get_mae, script
    am = importlib.import_module("module.align_model")
    assert am.log_mel_spectrogram is la_audio.log_mel_spectrogram          # module/align_model.py:9,84
    assert am.pad_or_trim is la_audio.pad_or_trim and am.N_FRAMES == 3000
