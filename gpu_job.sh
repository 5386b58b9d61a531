#!/bin/bash
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== gemm tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | grep -v "$F" | grep -E "^E  |passed|failed|FAILED|Error" | cut -c1-300 | tail -5
echo "=== epilogue variants on ffn-up"; for e in none bias gelu res gelugrad; do timeout 100 python tools/gemm_bench.py --only "vit ffn-up fwd" --epi $e --cfg 0 2>&1 | tail -1 | cut -c1-70; done
echo "=== gemm sweep"; timeout 300 python tools/gemm_bench.py --json gpurun_out/gemm_sweep_r1j.json 2>&1 | tail -14 | cut -c1-100
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_r1j.json | cut -c1-2400
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "$F" | grep -E "^E  |passed|failed|FAILED|Error" | cut -c1-300 | tail -12
