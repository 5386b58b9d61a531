#!/bin/bash
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== tcgen05 attention bwd tests"; timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "tcgen05" 2>&1 | grep -v "$F" | grep -E "^E  |passed|failed|Error|assert" | cut -c1-300 | tail -25
echo "=== all other tests"; timeout 1200 python -m pytest tests -m gpu -q -k "not tcgen05" 2>&1 | grep -v "$F" | grep -E "^E  |passed|failed|FAILED|Error" | cut -c1-300 | tail -25
echo "=== bench (mma.sync attention)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "$F" | tail -1 | cut -c1-1500 | tee gpurun_out/bench_r1c.json
echo "=== bench (tcgen05 attention bwd)"; VLM_ATTN_TC=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>&1 | grep -v "$F" | tail -1 | cut -c1-900 | tee gpurun_out/bench_r1c_tc.json
