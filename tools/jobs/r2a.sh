#!/bin/bash
# Round-2 GPU job A: measure the prepared experiments (pair GEMM with uniform issuer + staged epilogue, packed GELU, pipelined
# attention backward) and capture the ncu report of the K = 768 out-projection GEMM.  Every step has its own timeout.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
PAIR=$PWD/vilmedic_b200/libvlmb200_pair.so
F32=$PWD/vilmedic_b200/libvlmb200_f32x2.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2a_pytest_default.log 2>&1; echo "default pytest rc=$?" >> $O/r2a_status.log
VLM_LIB=$PAIR timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k 2cta > $O/r2a_pytest_pair_2cta.log 2>&1; echo "pair 2cta pytest rc=$?" >> $O/r2a_status.log
VLM_LIB=$PAIR VLM_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q > $O/r2a_pytest_pair_all.log 2>&1; echo "pair all-gemm pytest rc=$?" >> $O/r2a_status.log
VLM_LIB=$PAIR timeout 300 python tools/gemm_bench.py --json $O/r2a_gemm_pair.json > $O/r2a_gemm_pair.log 2>&1; echo "pair bench rc=$?" >> $O/r2a_status.log
for epi in none bias res; do
  timeout 120 python tools/gemm_bench.py --only "out fwd" --epi $epi >> $O/r2a_gemm_out_epi.log 2>&1
done
VLM_LIB=$F32 timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q > $O/r2a_pytest_f32x2.log 2>&1; echo "f32x2 pytest rc=$?" >> $O/r2a_status.log
VLM_LIB=$F32 timeout 120 python tools/gemm_bench.py --only "ffn-up fwd" --cfg 0 > $O/r2a_gemm_f32x2.log 2>&1
timeout 120 python tools/gemm_bench.py --only "ffn-up fwd" --cfg 0 >> $O/r2a_gemm_f32x2.log 2>&1
VLM_ATTN_BWD_PIPE=1 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k attention > $O/r2a_pytest_bwdpipe.log 2>&1; echo "bwd pipe pytest rc=$?" >> $O/r2a_status.log
timeout 120 python tools/attn_bench.py > $O/r2a_attn.log 2>&1
VLM_ATTN_BWD_PIPE=1 timeout 120 python tools/attn_bench.py >> $O/r2a_attn.log 2>&1
VLM_LIB=$PAIR VLM_GEMM_2CTA=1 timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2a_bench_pair.log 2>&1; echo "pair bench.py rc=$?" >> $O/r2a_status.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o $O/r2a_gemm_out -f python tools/gemm_bench.py --only "vit out fwd" --cfg 0 > $O/r2a_ncu_out.log 2>&1
VLM_LIB=$PAIR timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 3 -c 1 -o $O/r2a_gemm2_out -f python tools/gemm_bench.py --only "vit out fwd" --cfg 1256 > $O/r2a_ncu_out2.log 2>&1
cat $O/r2a_status.log
