#!/bin/bash
# Round-2 GPU job V (1 GPU): LayerNorm v2 kernels, side-stream bias column sums, optimizer pipelined behind the backward pass.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2v_*
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "layernorm" > $O/r2v_ln_tests.log 2>&1; echo "ln tests rc=$?" >> $O/r2v_status.log
timeout 900 python -m pytest tests/test_rrg_gpu.py tests/test_graph_gpu.py tests/test_optim_gpu.py tests/test_models_gpu.py -m gpu -q > $O/r2v_model_tests.log 2>&1; echo "model tests rc=$?" >> $O/r2v_status.log
timeout 600 python tools/tail_bench.py > $O/r2v_tail_bench.log 2>&1; echo "tail bench rc=$?" >> $O/r2v_status.log
B="python bench.py --quick --steps 20 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline"
VLM_LN_V2=0 VLM_SIDE_COLSUM=0 VLM_PIPELINE_OPTIMIZER=0 timeout 300 $B > $O/r2v_bench_base.log 2>&1; echo "bench base rc=$?" >> $O/r2v_status.log
VLM_LN_V2=1 VLM_SIDE_COLSUM=0 VLM_PIPELINE_OPTIMIZER=0 timeout 300 $B > $O/r2v_bench_ln.log 2>&1; echo "bench ln rc=$?" >> $O/r2v_status.log
VLM_LN_V2=1 VLM_SIDE_COLSUM=1 VLM_PIPELINE_OPTIMIZER=0 timeout 300 $B > $O/r2v_bench_ln_side.log 2>&1; echo "bench ln+side rc=$?" >> $O/r2v_status.log
VLM_LN_V2=1 VLM_SIDE_COLSUM=0 VLM_PIPELINE_OPTIMIZER=1 timeout 300 $B > $O/r2v_bench_ln_pipe.log 2>&1; echo "bench ln+pipe rc=$?" >> $O/r2v_status.log
VLM_LN_V2=1 VLM_SIDE_COLSUM=1 VLM_PIPELINE_OPTIMIZER=1 timeout 300 $B > $O/r2v_bench_all.log 2>&1; echo "bench all rc=$?" >> $O/r2v_status.log
cat $O/r2v_status.log; tail -4 $O/r2v_ln_tests.log | cut -c1-300; grep -E "passed|failed|^FAILED|Error" $O/r2v_model_tests.log | tail -8 | cut -c1-300
cat $O/r2v_tail_bench.log
for f in base ln ln_side ln_pipe all; do echo "$f: $(tail -1 $O/r2v_bench_$f.log | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", "loss", d["config"].get("loss_last"))
except Exception as e: print("ERR", e)')"; done
