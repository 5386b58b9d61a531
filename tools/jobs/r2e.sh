#!/bin/bash
# Round-2 GPU job E: decode parity (full file), model-level parity at the benchmarked config, whole suite, bench sanity.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2e_status.log $O/r2e_report.jsonl
VLM_TEST_REPORT=$O/r2e_report.jsonl timeout 2400 python -m pytest tests/test_decode_gpu.py tests/test_rrg_gpu.py -m gpu -q > $O/r2e_pytest_a.log 2>&1; echo "decode+rrg pytest rc=$?" >> $O/r2e_status.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_decode_gpu.py --deselect tests/test_rrg_gpu.py > $O/r2e_pytest_rest.log 2>&1; echo "rest pytest rc=$?" >> $O/r2e_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2e_bench.log 2>&1; echo "bench.py rc=$?" >> $O/r2e_status.log
cat $O/r2e_status.log; grep -E "passed|failed|Error|error" $O/r2e_pytest_a.log | tail -15 | cut -c1-250; tail -3 $O/r2e_pytest_rest.log | cut -c1-200; cat $O/r2e_report.jsonl
