#!/bin/bash
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "$F" | tail -25
echo "=== gemm sweep"; timeout 600 python tools/gemm_bench.py --json gpurun_out/gemm_sweep_r1.json 2>&1 | grep -v "$F" | tail -20
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/bench_r1b.json
echo "=== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2460 -c 830 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
echo "=== ncu full (ViT layer-0 GEMMs fwd)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 1204 -c 4 -o gpurun_out/prof_gemm_r1b python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "=== ncu full (attention bwd)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 110 -c 2 -o gpurun_out/prof_attn_r1b python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
