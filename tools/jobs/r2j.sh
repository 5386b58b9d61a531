#!/bin/bash
# Round-2 GPU job J (2 GPUs): whole suite incl. the 2-rank DDP test, bench at N=1 and N=2 (per-layer gradient buckets), other workloads.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2j_status.log $O/r2j_report.jsonl
VLM_TEST_REPORT=$O/r2j_report.jsonl timeout 2700 python -m pytest tests -m gpu -q > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline > $O/r2j_bench_n1.log 2>&1; echo "bench n1 rc=$?" >> $O/r2j_status.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/r2j_bench_n2.log 2>&1; echo "bench n2 rc=$?" >> $O/r2j_status.log
timeout 300 python bench.py --workload mvqa --steps 5 --warmup 3 > $O/r2j_bench_mvqa.log 2>&1; echo "mvqa rc=$?" >> $O/r2j_status.log
timeout 300 python bench.py --workload convirt --steps 5 --warmup 3 > $O/r2j_bench_convirt.log 2>&1; echo "convirt rc=$?" >> $O/r2j_status.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload mvqa --steps 5 --warmup 3 > $O/r2j_bench_mvqa_n2.log 2>&1; echo "mvqa n2 rc=$?" >> $O/r2j_status.log
cat $O/r2j_status.log; grep -E "passed|failed|^FAILED" $O/r2j_pytest.log | tail -8 | cut -c1-200
for f in n1 n2 mvqa convirt mvqa_n2; do tail -1 $O/r2j_bench_$f.log | cut -c1-420; done
