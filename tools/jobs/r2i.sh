#!/bin/bash
# Round-2 GPU job I: whole suite (DeiT, proto loading, SCST, LN), capture check of the other workloads, their bench lines.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2i_status.log $O/r2i_report.jsonl
VLM_TEST_REPORT=$O/r2i_report.jsonl timeout 2700 python -m pytest tests -m gpu -q > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2i_status.log
timeout 300 python tools/capture_debug.py mvqa 8 > $O/r2i_capture_mvqa.log 2>&1; echo "capture mvqa rc=$?" >> $O/r2i_status.log
timeout 300 python tools/capture_debug.py convirt 8 > $O/r2i_capture_convirt.log 2>&1; echo "capture convirt rc=$?" >> $O/r2i_status.log
timeout 300 python bench.py --workload mvqa --steps 5 --warmup 3 > $O/r2i_bench_mvqa.log 2>&1; echo "mvqa rc=$?" >> $O/r2i_status.log
timeout 300 python bench.py --workload convirt --steps 5 --warmup 3 > $O/r2i_bench_convirt.log 2>&1; echo "convirt rc=$?" >> $O/r2i_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline > $O/r2i_bench.log 2>&1; echo "bench.py rc=$?" >> $O/r2i_status.log
cat $O/r2i_status.log; grep -E "passed|failed|^FAILED" $O/r2i_pytest.log | tail -8 | cut -c1-200; tail -4 $O/r2i_capture_mvqa.log | cut -c1-200; tail -4 $O/r2i_capture_convirt.log | cut -c1-200; tail -1 $O/r2i_bench_mvqa.log | cut -c1-500; tail -1 $O/r2i_bench_convirt.log | cut -c1-500; tail -1 $O/r2i_bench.log | cut -c1-200
