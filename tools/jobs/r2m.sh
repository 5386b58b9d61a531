#!/bin/bash
# Round-2 GPU job M (2 GPUs): where does the N=2 loss (24.4 vs 23.0 ms) come from?  NCCL channel count, bucket size, no exchange.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2m_status.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/r2m_$name.log 2>&1
  echo "$name rc=$? $(tail -1 $O/r2m_$name.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["e2e"]["ms_per_step"],3), "ms e2e")' 2>&1 | tail -1)" >> $O/r2m_status.log
}
run base A=1
run noex VLM_BENCH_NO_EXCHANGE=1
run ch4 NCCL_MAX_NCHANNELS=4
run ch8 NCCL_MAX_NCHANNELS=8
run ch2 NCCL_MAX_NCHANNELS=2
run b128 VLM_DDP_BUCKET_MB=128
run b8 VLM_DDP_BUCKET_MB=8
run ctas8 NCCL_MAX_CTAS=8
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29877 bench.py --gpus 2 --steps 2 --warmup 3 --no-roofline 2>&1 | grep -i -E "channels|nvls|algo|proto|nthreads" | head -30 > $O/r2m_nccl_info.log
cat $O/r2m_status.log; head -12 $O/r2m_nccl_info.log | cut -c1-250
