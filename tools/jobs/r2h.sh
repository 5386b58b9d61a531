#!/bin/bash
# Round-2 GPU job H: whole GPU suite (LN changes, SCST fixes), capture debugging with anomaly mode, bench.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2h_status.log $O/r2h_report.jsonl
VLM_TEST_REPORT=$O/r2h_report.jsonl timeout 2700 python -m pytest tests -m gpu -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_status.log
timeout 300 python tools/capture_debug.py mvqa 8 > $O/r2h_capture_mvqa.log 2>&1; echo "capture mvqa rc=$?" >> $O/r2h_status.log
timeout 300 python tools/capture_debug.py convirt 8 > $O/r2h_capture_convirt.log 2>&1; echo "capture convirt rc=$?" >> $O/r2h_status.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline > $O/r2h_bench.log 2>&1; echo "bench.py rc=$?" >> $O/r2h_status.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 1700 --csv --log-file $O/r2h_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-graph > $O/r2h_ncu_launch.log 2>&1; echo "launch list rc=$?" >> $O/r2h_status.log
cat $O/r2h_status.log; grep -E "passed|failed|^E  " $O/r2h_pytest.log | tail -12 | cut -c1-300; grep -B2 -A12 "Traceback of forward\|previous calls" $O/r2h_capture_mvqa.log | head -60 | cut -c1-200; tail -1 $O/r2h_bench.log | cut -c1-300
