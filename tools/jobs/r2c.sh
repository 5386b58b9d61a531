#!/bin/bash
# Round-2 GPU job C: full GPU test suite after the epilogue / optimizer / boundary changes, micro-bench, bench.py with the shape table.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2c_status.log
timeout 1200 python -m pytest tests -m gpu -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_status.log
timeout 300 python tools/gemm_bench.py --cfg 0 > $O/r2c_gemm.log 2>&1; echo "gemm bench rc=$?" >> $O/r2c_status.log
VLM_BENCH_SHAPES=$O/r2c_shapes.txt timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2c_bench.log 2>&1; echo "bench.py rc=$?" >> $O/r2c_status.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file $O/r2c_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-graph > $O/r2c_ncu_launch.log 2>&1; echo "launch list rc=$?" >> $O/r2c_status.log
cat $O/r2c_status.log; tail -5 $O/r2c_pytest.log
