#!/bin/bash
# Round-2 GPU job G: decode parity (lock-step cfg#5), benchmarked-config parity, SCST, capture debugging of the other workloads.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2g_status.log $O/r2g_report.jsonl
VLM_TEST_REPORT=$O/r2g_report.jsonl timeout 2400 python -m pytest tests/test_decode_gpu.py tests/test_rrg_gpu.py tests/test_scst_gpu.py -m gpu -q > $O/r2g_pytest_a.log 2>&1; echo "decode+rrg+scst pytest rc=$?" >> $O/r2g_status.log
timeout 300 python tools/capture_debug.py mvqa 8 > $O/r2g_capture_mvqa.log 2>&1; echo "capture mvqa rc=$?" >> $O/r2g_status.log
timeout 300 python tools/capture_debug.py convirt 8 > $O/r2g_capture_convirt.log 2>&1; echo "capture convirt rc=$?" >> $O/r2g_status.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_decode_gpu.py --deselect tests/test_rrg_gpu.py --deselect tests/test_scst_gpu.py > $O/r2g_pytest_rest.log 2>&1; echo "rest pytest rc=$?" >> $O/r2g_status.log
cat $O/r2g_status.log; grep -E "passed|failed|Error" $O/r2g_pytest_a.log | tail -12 | cut -c1-300; cat $O/r2g_report.jsonl; tail -25 $O/r2g_capture_mvqa.log | cut -c1-200; tail -25 $O/r2g_capture_convirt.log | cut -c1-200; tail -3 $O/r2g_pytest_rest.log | cut -c1-200
