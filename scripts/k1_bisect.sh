#!/bin/bash
# perf triage of K1: which role bounds the tile time? (results are garbage with flags set)
for f in 0 1 2 4 3 6 7; do
  echo -n "LA_LOGMEL_DBG=$f : "
  LA_LOGMEL_DBG=$f python bench.py --clips 400 --steps 5 --warmup 3 --skip-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernels']['k1_logmel_ms'])"
done
