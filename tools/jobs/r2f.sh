#!/bin/bash
# Round-2 GPU job F: decode / rrg parity re-run (unscaled 12-layer models), full bench line (decode metric + same-box GPU arm),
# reference-gpu arm, other workloads, launch list of one step.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2f_status.log $O/r2f_report.jsonl
VLM_TEST_REPORT=$O/r2f_report.jsonl timeout 2400 python -m pytest tests/test_decode_gpu.py tests/test_rrg_gpu.py -m gpu -q > $O/r2f_pytest_a.log 2>&1; echo "decode+rrg pytest rc=$?" >> $O/r2f_status.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2f_bench.log 2>&1; echo "bench.py rc=$?" >> $O/r2f_status.log
timeout 300 python bench.py --impl reference-gpu --steps 5 --warmup 3 > $O/r2f_bench_refgpu.log 2>&1; echo "reference-gpu rc=$?" >> $O/r2f_status.log
timeout 300 python bench.py --workload mvqa --steps 5 --warmup 3 > $O/r2f_bench_mvqa.log 2>&1; echo "mvqa rc=$?" >> $O/r2f_status.log
timeout 300 python bench.py --workload convirt --steps 5 --warmup 3 > $O/r2f_bench_convirt.log 2>&1; echo "convirt rc=$?" >> $O/r2f_status.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 1700 --csv --log-file $O/r2f_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-graph > $O/r2f_ncu_launch.log 2>&1; echo "launch list rc=$?" >> $O/r2f_status.log
cat $O/r2f_status.log; grep -E "passed|failed|Error" $O/r2f_pytest_a.log | tail -12 | cut -c1-250; cat $O/r2f_report.jsonl; tail -2 $O/r2f_bench_refgpu.log | cut -c1-400; tail -2 $O/r2f_bench_mvqa.log | cut -c1-600; tail -2 $O/r2f_bench_convirt.log | cut -c1-600
