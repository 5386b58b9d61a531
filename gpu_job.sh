#!/bin/bash
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== all tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "$F" | grep -E "^E  |passed|failed|FAILED|Error|informational" | cut -c1-400 | tail -30
echo "=== gemm sweep"; timeout 300 python tools/gemm_bench.py --json gpurun_out/gemm_sweep_r1e.json 2>&1 | tail -40
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_r1e.json | cut -c1-2600
echo "=== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2180 -c 760 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-100
