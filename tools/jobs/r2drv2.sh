#!/bin/bash
# Round-2 GPU job (2 GPUs): the driver's own N=2 command line (full default bench: resident + e2e legs) on the default exchange path.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2drv2_*
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2drv2_bench.log 2>&1; echo "bench rc=$?" >> $O/r2drv2_status.log
timeout 300 python -m pytest tests/test_ddp_gpu.py -m gpu -q > $O/r2drv2_ddp_tests.log 2>&1; echo "ddp tests rc=$?" >> $O/r2drv2_status.log
cat $O/r2drv2_status.log; tail -3 $O/r2drv2_ddp_tests.log | cut -c1-200
grep -h '^{' $O/r2drv2_bench.log | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"]["how"][:60], d["config"].get("grad_exchange","")[:30], d.get("invalid"), d["gpu_launches"])'
grep -h "GradSync\|Traceback\|Error" $O/r2drv2_bench.log | head -5 | cut -c1-300
