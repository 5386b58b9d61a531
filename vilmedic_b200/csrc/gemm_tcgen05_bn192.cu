// Explicit instantiation of the tcgen05 GEMM launchers for 192-column tiles (see gemm_kernel.cuh / gemm_tcgen05.cu).
#include "gemm_kernel.cuh"

namespace vlm {
template int gemm_launch_bn<192>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, int, int, int, int,
                                 int, int, int, int, long long, long long, long long, const GemmEpilogue&, int, cudaStream_t);
}  // namespace vlm
