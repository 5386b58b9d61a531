#!/bin/bash
# Round-2 GPU job N4 (4 GPUs): the DEFAULT data-parallel path at 4 ranks (peer memory, two-shot): correctness + bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2n4_*
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VLM_DDP_TRANSPORT=p2p VLM_DDP_STEPS=3 timeout 200 $TR --master-port 29551 tools/ddp_check.py > $O/r2n4_check.log 2>&1; echo "check rc=$?" >> $O/r2n4_status.log
timeout 240 $TR --master-port 29555 bench.py --gpus $N --quick --steps 15 --warmup 3 > $O/r2n4_bench.log 2>&1; echo "bench rc=$?" >> $O/r2n4_status.log
cat $O/r2n4_status.log; grep -h "ddp_check\|Error\|error\|GradSync" $O/r2n4_check.log | tail -6 | cut -c1-300
echo "bench: $(grep -h '^{' $O/r2n4_bench.log | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), d["config"].get("grad_exchange","")[:40], d.get("invalid"))
except Exception as e: print("ERR", e)')"; grep -h "GradSync\|Traceback\|Error" $O/r2n4_bench.log | head -3 | cut -c1-300
