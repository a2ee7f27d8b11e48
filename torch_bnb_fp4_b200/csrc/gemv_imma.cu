// placeholder until the integer tensor-core GEMV lands: reports "unsupported" so capi.cu uses the generic kernel
#include "common.cuh"
namespace fp4b200 {
bool gemv_imma_supported(int, int, int, int, int) { return false; }
int gemv_imma_dispatch(const void*, const uint8_t*, const float*, const fp4_b200_nested_t*,
                       const NestedDev&, const void*, void*, int, int, int, int, int,
                       cudaStream_t) {
    return FP4_B200_ERR_UNSUPPORTED;
}
}  // namespace fp4b200
