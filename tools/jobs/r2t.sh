#!/bin/bash
# Round-2 GPU job T (2 GPUs): bf16 gradient payload — 2-rank parity test (both payloads) and bench at N=2 (bf16 vs fp32 payload).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2t_status.log
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_optim_gpu.py -m gpu -q > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2t_status.log
run() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/r2t_$name.log 2>&1
  echo "$name rc=$? $(tail -1 $O/r2t_$name.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", round(d["e2e"]["ms_per_step"],3), "ms e2e", round(d["value"],1))' 2>&1 | tail -1)" >> $O/r2t_status.log
}
run bf16 VLM_DDP_PAYLOAD=bf16
run fp32 VLM_DDP_PAYLOAD=fp32
run noex VLM_BENCH_NO_EXCHANGE=1
cat $O/r2t_status.log; tail -3 $O/r2t_pytest.log | cut -c1-200; grep "ddp_check" $O/r2t_pytest.log | head
