#!/bin/bash
# Round-2 last 1-GPU job: e2e leg with the loss read one step behind (vs loss.item() after every replay), graph test.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2fin4_*
timeout 200 python -m pytest tests/test_graph_gpu.py -m gpu -q > $O/r2fin4_graph_test.log 2>&1; echo "graph test rc=$?" >> $O/r2fin4_status.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline"
timeout 200 $B > $O/r2fin4_bench_async.log 2>&1; echo "bench async rc=$?" >> $O/r2fin4_status.log
VLM_BENCH_SYNC_LOSS=1 timeout 200 $B > $O/r2fin4_bench_sync.log 2>&1; echo "bench sync rc=$?" >> $O/r2fin4_status.log
cat $O/r2fin4_status.log; tail -2 $O/r2fin4_graph_test.log | cut -c1-200
for f in async sync; do grep -h '^{' $O/r2fin4_bench_$f.log | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"],1), d["e2e"]["how"][-90:], d["config"]["loss_last"])'; done
