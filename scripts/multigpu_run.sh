#!/bin/bash
# bench.py + config 5 + H2D scaling at N ranks (run under `gpurun --gpus N`); outputs in gpurun_out/
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 300 $TR scripts/bench_config5.py --clips 1000000 > gpurun_out/config5_n$N.json 2> gpurun_out/config5_n$N.err
timeout 120 $TR scripts/h2d_scaling.py > gpurun_out/h2d_n$N.json 2> gpurun_out/h2d_n$N.err
tail -n 1 gpurun_out/bench_n$N.json | cut -c1-400; cat gpurun_out/config5_n$N.json; cut -c1-600 gpurun_out/h2d_n$N.json
tail -n 3 gpurun_out/bench_n$N.err gpurun_out/config5_n$N.err
