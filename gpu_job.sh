#!/bin/bash
mkdir -p gpurun_out
F='loss_type\|Swig\|swig\|Docs:\|^$'
echo "=== decode tests"; timeout 600 python -m pytest tests/test_decode_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED|informational" | cut -c1-400 | tail -8
echo "=== launches + dram bytes"; timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2180 -c 760 --csv --log-file gpurun_out/launches_r1k_dram.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-100
