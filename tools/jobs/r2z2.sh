#!/bin/bash
# Round-2 GPU job Z2 (8 GPUs): two-shot peer-memory exchange with four peer loads in flight per thread.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2z2_*
N=${VLM_JOB_GPUS:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VLM_DDP_TRANSPORT=p2p VLM_P2P_TWO_SHOT=1 VLM_DDP_STEPS=3 timeout 200 $TR --master-port 29551 tools/ddp_check.py > $O/r2z2_check_p2p2.log 2>&1; echo "check p2p two-shot rc=$?" >> $O/r2z2_status.log
B="bench.py --gpus $N --quick --steps 15 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline"
VLM_DDP_TRANSPORT=p2p VLM_P2P_TWO_SHOT=1 timeout 240 $TR --master-port 29555 $B > $O/r2z2_bench_p2p2.log 2>&1; echo "bench p2p two-shot rc=$?" >> $O/r2z2_status.log
cat $O/r2z2_status.log; grep -h "ddp_check\|Error\|error\|GradSync" $O/r2z2_check_p2p2.log | tail -6 | cut -c1-300
for f in p2p2; do echo "$f: $(grep -h '^{' $O/r2z2_bench_$f.log | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), d["config"].get("grad_exchange","")[:40], d.get("invalid"))
except Exception as e: print("ERR", e)')"; grep -h "GradSync\|Traceback\|Error" $O/r2z2_bench_$f.log | head -3 | cut -c1-300; done
