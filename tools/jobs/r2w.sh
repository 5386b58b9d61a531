#!/bin/bash
# Round-2 GPU job W (1 GPU): attention backward with the delta pre-pass fused in; weight-gradient GEMMs on a low-priority side stream.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2w_*
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attention or attn" > $O/r2w_attn_tests.log 2>&1; echo "attn tests rc=$?" >> $O/r2w_status.log
timeout 300 python tools/attn_bench.py > $O/r2w_attn_bench.log 2>&1; echo "attn bench rc=$?" >> $O/r2w_status.log
B="python bench.py --quick --steps 20 --warmup 3 --no-cpu-baseline --no-decode --no-gpu-baseline --no-roofline"
VLM_SIDE_WGRAD=0 VLM_PIPELINE_OPTIMIZER=0 timeout 300 $B > $O/r2w_bench_base.log 2>&1; echo "bench base rc=$?" >> $O/r2w_status.log
VLM_SIDE_WGRAD=1 VLM_PIPELINE_OPTIMIZER=0 timeout 300 $B > $O/r2w_bench_wgrad.log 2>&1; echo "bench wgrad rc=$?" >> $O/r2w_status.log
VLM_SIDE_WGRAD=0 VLM_PIPELINE_OPTIMIZER=1 timeout 300 $B > $O/r2w_bench_pipe.log 2>&1; echo "bench pipe rc=$?" >> $O/r2w_status.log
VLM_SIDE_WGRAD=1 VLM_PIPELINE_OPTIMIZER=1 timeout 300 $B > $O/r2w_bench_wgrad_pipe.log 2>&1; echo "bench wgrad+pipe rc=$?" >> $O/r2w_status.log
VLM_SIDE_WGRAD=1 VLM_PIPELINE_OPTIMIZER=0 timeout 900 python -m pytest tests/test_rrg_gpu.py tests/test_graph_gpu.py tests/test_models_gpu.py -m gpu -q > $O/r2w_model_tests.log 2>&1; echo "model tests (side wgrad) rc=$?" >> $O/r2w_status.log
cat $O/r2w_status.log; tail -4 $O/r2w_attn_tests.log | cut -c1-300; cat $O/r2w_attn_bench.log | tail -5; grep -E "passed|failed|^FAILED|Error" $O/r2w_model_tests.log | tail -8 | cut -c1-300
for f in base wgrad pipe wgrad_pipe; do echo "$f: $(tail -1 $O/r2w_bench_$f.log | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", d["config"].get("launch"), d["config"].get("loss_last"))
except Exception as e: print("ERR", e)')"; done
