#!/bin/bash
# Round-2 GPU job D: device-side decoding — parity tests (measurements to gpurun_out/r2d_decode_report.jsonl), decode bench.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2d_status.log $O/r2d_decode_report.jsonl
VLM_TEST_REPORT=$O/r2d_decode_report.jsonl timeout 1500 python -m pytest tests/test_decode_gpu.py -m gpu -q -x > $O/r2d_pytest_decode.log 2>&1; echo "decode pytest rc=$?" >> $O/r2d_status.log
timeout 600 python tools/decode_bench.py > $O/r2d_decode_bench.log 2>&1; echo "decode bench rc=$?" >> $O/r2d_status.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_decode_gpu.py > $O/r2d_pytest_rest.log 2>&1; echo "rest pytest rc=$?" >> $O/r2d_status.log
cat $O/r2d_status.log; tail -30 $O/r2d_pytest_decode.log | cut -c1-220; cat $O/r2d_decode_report.jsonl; tail -3 $O/r2d_decode_bench.log
