// explicit instantiation of the tcgen05 GEMM for BN = 256 (tc_gemm_impl.cuh)
#include "tc_gemm_impl.cuh"

namespace tvit {
template int dispatch_epi<256>(const tvit_gemm_args*, const GemmShape&, const EpiParams&, cudaStream_t);
}  // namespace tvit
