#!/bin/bash
# Round-2 GPU job S (1 GPU): backward attention with the straight-line visible-chunk passes; whole suite; bench.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2s_status.log $O/r2s_report.jsonl
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention or attn" > $O/r2s_attn_tests.log 2>&1; echo "attn tests rc=$?" >> $O/r2s_status.log
timeout 300 python tools/attn_bench.py > $O/r2s_attn_bench.log 2>&1; echo "attn bench rc=$?" >> $O/r2s_status.log
VLM_TEST_REPORT=$O/r2s_report.jsonl timeout 2700 python -m pytest tests -m gpu -q > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline > $O/r2s_bench_n1.log 2>&1; echo "bench n1 rc=$?" >> $O/r2s_status.log
cat $O/r2s_status.log; tail -3 $O/r2s_attn_tests.log | cut -c1-200; cat $O/r2s_attn_bench.log | tail -12
grep -E "passed|failed|^FAILED" $O/r2s_pytest.log | tail -8 | cut -c1-200
tail -1 $O/r2s_bench_n1.log | cut -c1-300
