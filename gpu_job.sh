#!/bin/bash
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== gloria + gemm tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "gloria or gemm or contrastive" 2>&1 | grep -v "$F" | grep -E "^E  |passed|failed|FAILED|Error" | cut -c1-600 | tail -30
echo "=== gemm sweep (per-thread epilogue restored)"; timeout 300 python tools/gemm_bench.py --json gpurun_out/gemm_sweep_r1f.json 2>&1 | tail -16
echo "=== epilogue variants on ffn-up"; for e in none bias gelu res gelugrad; do timeout 100 python tools/gemm_bench.py --only "vit ffn-up fwd" --epi $e 2>&1 | tail -1; done
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/bench_r1f.json | cut -c1-1800
echo "=== ncu full ffn-up gelu"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/prof_ffnup_gelu python tools/gemm_bench.py --only "vit ffn-up fwd" --cfg 0 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/prof_ffnup_none python tools/gemm_bench.py --only "vit ffn-up fwd" --cfg 0 --epi none > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
