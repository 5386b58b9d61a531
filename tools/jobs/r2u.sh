#!/bin/bash
# Round-2 GPU job U (2 GPUs): SM margin for the persistent kernels x NCCL channel count (bf16 payload).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2u_status.log
run() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/r2u_$name.log 2>&1
  echo "$name rc=$? $(tail -1 $O/r2u_$name.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["e2e"]["ms_per_step"],3), "ms e2e", round(d["value"],1))' 2>&1 | tail -1)" >> $O/r2u_status.log
}
run m4c4 VLM_SM_MARGIN=4 NCCL_MAX_NCHANNELS=4 NCCL_MIN_NCHANNELS=4
run m8c8 VLM_SM_MARGIN=8 NCCL_MAX_NCHANNELS=8 NCCL_MIN_NCHANNELS=8
run m4c4hp VLM_SM_MARGIN=4 NCCL_MAX_NCHANNELS=4 NCCL_MIN_NCHANNELS=4 TORCH_NCCL_HIGH_PRIORITY=1
run m2c2 VLM_SM_MARGIN=2 NCCL_MAX_NCHANNELS=2 NCCL_MIN_NCHANNELS=2
run m16c16 VLM_SM_MARGIN=16 NCCL_MAX_NCHANNELS=16 NCCL_MIN_NCHANNELS=16
run m8c8hp VLM_SM_MARGIN=8 NCCL_MAX_NCHANNELS=8 NCCL_MIN_NCHANNELS=8 TORCH_NCCL_HIGH_PRIORITY=1
run m0hp TORCH_NCCL_HIGH_PRIORITY=1
run m4noex VLM_SM_MARGIN=4 VLM_BENCH_NO_EXCHANGE=1
cat $O/r2u_status.log
