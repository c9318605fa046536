#!/bin/bash
# compute-sanitizer over the small-shape parity tests (run on the GPU box)
for tool in memcheck racecheck synccheck; do
  echo "== $tool =="
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_decode.py tests/test_gpu_logmel.py -x -q -k "golden or error or ragged or single_clip_vs_fp64_oracle or batch_shares" 2>&1 | tail -4
done
