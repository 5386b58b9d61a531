#!/bin/bash
# Round-2 GPU job Y (2 GPUs): peer-memory gradient exchange with the publish stream; one-shot vs two-shot (reduce-scatter) vs NCCL.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2y_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
VLM_DDP_TRANSPORT=p2p VLM_P2P_TWO_SHOT=0 VLM_DDP_STEPS=3 timeout 240 $TR --master-port 29551 tools/ddp_check.py > $O/r2y_check_p2p1.log 2>&1; echo "check p2p one-shot rc=$?" >> $O/r2y_status.log
VLM_DDP_TRANSPORT=p2p VLM_P2P_TWO_SHOT=1 VLM_DDP_STEPS=3 timeout 240 $TR --master-port 29552 tools/ddp_check.py > $O/r2y_check_p2p2.log 2>&1; echo "check p2p two-shot rc=$?" >> $O/r2y_status.log
B="bench.py --gpus 2 --quick --steps 20 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline"
timeout 300 $TR --master-port 29553 $B > $O/r2y_bench_nccl.log 2>&1; echo "bench nccl rc=$?" >> $O/r2y_status.log
VLM_DDP_TRANSPORT=p2p VLM_P2P_TWO_SHOT=0 timeout 300 $TR --master-port 29554 $B > $O/r2y_bench_p2p1.log 2>&1; echo "bench p2p one-shot rc=$?" >> $O/r2y_status.log
VLM_DDP_TRANSPORT=p2p VLM_P2P_TWO_SHOT=1 timeout 300 $TR --master-port 29555 $B > $O/r2y_bench_p2p2.log 2>&1; echo "bench p2p two-shot rc=$?" >> $O/r2y_status.log
timeout 300 python bench.py --quick --steps 20 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline > $O/r2y_bench_n1.log 2>&1; echo "bench n1 rc=$?" >> $O/r2y_status.log
cat $O/r2y_status.log; grep -h "ddp_check\|Error\|error\|GradSync" $O/r2y_check_p2p1.log $O/r2y_check_p2p2.log | tail -12 | cut -c1-300
for f in nccl p2p1 p2p2 n1; do echo "$f: $(grep -h '^{' $O/r2y_bench_$f.log | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), d["config"].get("grad_exchange","")[:40], d["config"].get("loss_last"))
except Exception as e: print("ERR", e)')"; grep -h "GradSync\|Traceback\|Error" $O/r2y_bench_$f.log | head -3 | cut -c1-300; done
