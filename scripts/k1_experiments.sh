#!/bin/bash
# K1 timing experiments (perf triage; the dbg >= 32 knobs break the numerics on purpose)
for dbg in 0 32 64 128 256 416; do
  echo "== LA_LOGMEL_DBG=$dbg"
  LA_LOGMEL_DBG=$dbg timeout 120 python bench.py --steps 5 --warmup 3 --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('k1_ms', d['kernels']['k1_logmel_ms'])"
done
