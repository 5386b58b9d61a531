#!/bin/bash
# one GPU-box visit: tests, bench, profiles (everything under its own timeout)
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== gemm tests (1-CTA, split-K)"; timeout 400 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "not 2cta" 2>&1 | grep -v "$F" | tail -8
echo "=== ops/rrg/models tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_rrg_gpu.py tests/test_models_gpu.py -m gpu -q 2>&1 | grep -v "$F" | tail -25
echo "=== bench (1-CTA)"; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/bench_1cta.json
echo "=== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2460 -c 830 --csv --log-file gpurun_out/launches_r1.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
echo "=== ncu full (gemm)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 300 -c 3 -o gpurun_out/prof_gemm_r1 python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "=== 2-CTA tests"; timeout 240 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "2cta" 2>&1 | grep -v "$F" | tail -8 && \
  (echo "=== bench (2-CTA)"; VLM_GEMM_2CTA=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/bench_2cta.json)
