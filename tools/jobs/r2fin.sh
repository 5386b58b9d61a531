#!/bin/bash
# Round-2 final GPU job (1 GPU): whole suite, smoke, the default bench line, ncu launch list, GEMM DRAM traffic, --set full captures.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2fin_*
VLM_TEST_REPORT=$O/r2fin_report.jsonl timeout 1500 python -m pytest tests -m gpu -q > $O/r2fin_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2fin_status.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2fin_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2fin_status.log
VLM_BENCH_SHAPES=$O/r2fin_shapes.txt timeout 900 python bench.py > $O/r2fin_bench.log 2>&1; echo "bench rc=$?" >> $O/r2fin_status.log
Q="python bench.py --quick --no-graph --steps 1 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 5000 --csv --log-file $O/r2fin_launches.csv $Q > $O/r2fin_ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/r2fin_status.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm|optim_kernel" -s 860 -c 900 --csv --log-file $O/r2fin_gemm_dram.csv $Q > $O/r2fin_ncu_dram.log 2>&1; echo "ncu gemm dram rc=$?" >> $O/r2fin_status.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:layernorm_bwd_v2 -s 2 -c 1 -o $O/r2fin_ln_bwd -f python tools/tail_bench.py --only ln_bwd --iters 1 > $O/r2fin_ncu1.log 2>&1; echo "ncu ln bwd rc=$?" >> $O/r2fin_status.log
timeout 200 $NCU -k regex:layernorm_fwd_v2 -s 2 -c 1 -o $O/r2fin_ln_fwd -f python tools/tail_bench.py --only ln_fwd --iters 1 > $O/r2fin_ncu2.log 2>&1; echo "ncu ln fwd rc=$?" >> $O/r2fin_status.log
timeout 200 $NCU -k regex:attn_bwd_tc -s 1 -c 1 -o $O/r2fin_attn_bwd_vit -f python tools/attn_bench.py --only vit --iters 2 > $O/r2fin_ncu3.log 2>&1; echo "ncu attn bwd rc=$?" >> $O/r2fin_status.log
timeout 200 $NCU -k regex:gemm_bf16 -s 3 -c 1 -o $O/r2fin_gemm_out -f python tools/gemm_bench.py --only "dec out fwd" --cfg 0 > $O/r2fin_ncu4.log 2>&1; echo "ncu gemm rc=$?" >> $O/r2fin_status.log
cat $O/r2fin_status.log; grep -E "passed|failed|^FAILED" $O/r2fin_pytest.log | tail -8 | cut -c1-250; tail -2 $O/r2fin_smoke.log | cut -c1-200
tail -1 $O/r2fin_bench.log | cut -c1-400
