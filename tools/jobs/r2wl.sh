#!/bin/bash
# Round-2 GPU job (1 GPU): the other BASELINE configs through the bench harness with the end-of-round code.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2wl_*
timeout 300 python bench.py --workload convirt --steps 5 --warmup 3 > $O/r2wl_convirt.log 2>&1; echo "convirt rc=$?" >> $O/r2wl_status.log
timeout 300 python bench.py --workload mvqa --steps 5 --warmup 3 > $O/r2wl_mvqa.log 2>&1; echo "mvqa rc=$?" >> $O/r2wl_status.log
cat $O/r2wl_status.log
for f in convirt mvqa; do grep -h '^{' $O/r2wl_$f.log | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(d["metric"][:50], round(d["ms_per_step"],3), "ms", round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), d["config"].get("launch"), d["gpu_launches"])'; grep -h "Traceback\|Error" $O/r2wl_$f.log | head -3 | cut -c1-300; done
