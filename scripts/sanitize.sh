#!/bin/bash
# compute-sanitizer over the small-shape parity tests (run on the GPU box); logs in gpurun_out/
mkdir -p gpurun_out
SEL="golden or error or ragged or single_clip_vs_fp64_oracle or batch_shares or small_vocab or planted"
for tool in memcheck synccheck racecheck; do
  echo "== $tool =="
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_decode.py tests/test_gpu_logmel.py tests/test_gpu_head.py -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
