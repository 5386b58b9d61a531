#!/bin/bash
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== decode tests"; timeout 600 python -m pytest tests/test_decode_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED|informational" | cut -c1-400 | tail -8
echo "=== epilogue variants on ffn-up (contiguous spans)"; for e in bias gelu res gelugrad; do timeout 100 python tools/gemm_bench.py --only "vit ffn-up fwd" --epi $e --cfg 0 2>&1 | tail -1 | cut -c1-70; done
echo "=== gemm tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | grep -v "$F" | grep -E "^E  |passed|failed|FAILED|Error" | cut -c1-300 | tail -5
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_r1k.json | cut -c1-2400
echo "=== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2180 -c 760 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-100
