// gemm.cu -- K2: tcgen05 tensor-core GEMMs (placeholder until the tcgen05 path lands in this round).
#include "blocks.cuh"
namespace ggufb200 {
int gemm_fused_supported(int) { return 0; }
int gemm_fused_dispatch(int, const void *, long long, long long, const void *, long long, long long, int, int, const void *, int, void *,
                        long long, cudaStream_t)
{
    return GGUFB200_E_UNSUPPORTED;
}
int gemm_dense_dispatch(const void *, long long, long long, long long, const void *, long long, long long, int, const void *, int, void *,
                        long long, cudaStream_t)
{
    return GGUFB200_E_UNSUPPORTED;
}
}  // namespace ggufb200
