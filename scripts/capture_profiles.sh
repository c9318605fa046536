#!/bin/bash
# Run on the GPU box (gpurun): launch list of one bench step + full ncu captures of K1/K2/K3/N1.
# Outputs land in gpurun_out/; scripts/summarize_profiles.py turns them into profiles/*.
set -x
mkdir -p gpurun_out
# only the product's kernels (the synthetic-input generation launches hundreds of torch kernels first)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"emit_kernel|viterbi_|logmel_kernel|logmel_floor_kernel|logmel_init_kernel|gather_logp_kernel" \
    -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-e2e --skip-head > gpurun_out/launches_bench.log 2>&1
for k in emit_kernel logmel_kernel viterbi_wave_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -o gpurun_out/$k -f \
      python bench.py --clips 400 --steps 2 --warmup 3 --skip-e2e --skip-head > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_lse_kernel -s 1 -c 1 -o gpurun_out/head_lse_kernel -f \
    python scripts/bench_head.py 400 > gpurun_out/ncu_head_lse_kernel.log 2>&1
ls -la gpurun_out
# K3 alone at fine warp-state sampling: the 2 000-clip batch and one 30 s clip per SM (source-level stall profile)
timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:viterbi_wave -s 2 -c 1 \
    -o gpurun_out/k3_wave_batch -f python scripts/k3_single.py opencpop 2000 5 > gpurun_out/ncu_k3_wave_batch.log 2>&1
