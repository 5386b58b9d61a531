#!/bin/bash
# Round-2 last GPU job (1 GPU): the whole suite on the final tree, the ConVIRT workload with the loss-kernel timing, a quick bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2fin2_*
VLM_TEST_REPORT=$O/r2fin2_report.jsonl timeout 1500 python -m pytest tests -m gpu -q > $O/r2fin2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2fin2_status.log
timeout 300 python bench.py --workload convirt --steps 5 --warmup 3 > $O/r2fin2_convirt.log 2>&1; echo "convirt rc=$?" >> $O/r2fin2_status.log
timeout 300 python bench.py --quick --steps 20 --warmup 3 > $O/r2fin2_bench.log 2>&1; echo "bench rc=$?" >> $O/r2fin2_status.log
cat $O/r2fin2_status.log; grep -E "passed|failed|^FAILED" $O/r2fin2_pytest.log | tail -6 | cut -c1-250
grep -h '^{' $O/r2fin2_convirt.log | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), json.dumps(d["loss_kernels"]))'
grep -h '^{' $O/r2fin2_bench.log | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"],1))'
