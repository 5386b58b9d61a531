#!/bin/bash
# Round-2 GPU job B: epilogue rewrite (row inputs one span ahead, no local-memory spills), pair kernel staged by default,
# new optimizer + big-shape GEMM parity tests.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2b_status.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_status.log
timeout 300 python tools/gemm_bench.py --json $O/r2b_gemm.json > $O/r2b_gemm.log 2>&1; echo "gemm bench rc=$?" >> $O/r2b_status.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2b_bench.log 2>&1; echo "bench.py rc=$?" >> $O/r2b_status.log
VLM_GEMM_2CTA=1 timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2b_bench_pair.log 2>&1; echo "pair bench.py rc=$?" >> $O/r2b_status.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o $O/r2b_gemm_out -f python tools/gemm_bench.py --only "vit out fwd" --cfg 0 > $O/r2b_ncu_out.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o $O/r2b_gemm_lmhead -f python tools/gemm_bench.py --only "lm head fwd" --cfg 0 > $O/r2b_ncu_lm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 3 -c 1 -o $O/r2b_gemm2_lmhead -f python tools/gemm_bench.py --only "lm head fwd" --cfg 1256 > $O/r2b_ncu_lm2.log 2>&1
cat $O/r2b_status.log; tail -5 $O/r2b_pytest.log
