// Blockwise FP4 quantiser with the bitsandbytes decision thresholds (kQuantizeBlockwise<FP4> /
// dQuantizeFP4 of bitsandbytes 0.42, the dependency the reference pins: requirements.txt:1).
// The reference reaches it through bitsandbytes (torch_bnb_fp4/__init__.py:736-746, :775); it is the
// step before the hot path, provided so models can be converted without bitsandbytes.
// Not performance critical: one warp per quantisation block, two passes over the block.
#include "common.cuh"

namespace fp4b200 {

__device__ __forceinline__ uint32_t quantize_fp4_nibble(float x) {
    const uint32_t sign = x < 0.f ? 8u : 0u;
    x = fabsf(x);
    if (x > 0.29166667f) {
        if (x > 0.583333f) return (x > 0.8333333f ? 3u : 2u) + sign;
        return (x > 0.4166667f ? 5u : 4u) + sign;
    }
    if (x > 0.0859375f) return (x > 0.20833333f ? 7u : 6u) + sign;
    return (x > 0.00260417f ? 1u : 0u) + sign;
}

template <typename T>
__global__ void __launch_bounds__(256)
quantize_kernel(const T* __restrict__ w, uint8_t* __restrict__ packed, float* __restrict__ absmax,
                const int64_t n, const int blocksize) {
    const int lane = threadIdx.x & 31;
    const int64_t nblocks = (n + blocksize - 1) / blocksize;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    for (int64_t blk = warp0; blk < nblocks; blk += (int64_t)gridDim.x * 8) {
        const int64_t e0 = blk * blocksize;
        const int64_t e1 = (e0 + blocksize < n) ? e0 + blocksize : n;
        float m = 0.f;
        for (int64_t i = e0 + lane; i < e1; i += 32) m = fmaxf(m, fabsf(DT<T>::to_f32(w[i])));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) absmax[blk] = m;
        const float inv = __fdiv_rn(1.0f, m);  // m == 0 -> inf -> 0*inf = NaN -> nibble 0, as bitsandbytes
        for (int64_t i = e0 + 2 * lane; i < e1; i += 64) {
            const uint32_t hi = quantize_fp4_nibble(__fmul_rn(DT<T>::to_f32(w[i]), inv));
            const uint32_t lo =
                (i + 1 < e1) ? quantize_fp4_nibble(__fmul_rn(DT<T>::to_f32(w[i + 1]), inv)) : 0u;
            packed[i >> 1] = (uint8_t)((hi << 4) | lo);
        }
    }
}

int quantize_dispatch(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed,
                      float* absmax, cudaStream_t st) {
    if (!w || !packed || !absmax) return FP4_B200_ERR_NULL;
    if (n < 0) return FP4_B200_ERR_SHAPE;
    if (ilog2_exact(blocksize) < 1) return FP4_B200_ERR_BLOCKSIZE;
    if (n == 0) return FP4_B200_OK;
    const int64_t nblocks = (n + blocksize - 1) / blocksize;
    int64_t blocks = (nblocks + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    switch (dtype) {
        case FP4_B200_F16:
            quantize_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)w, packed,
                                                                      absmax, n, blocksize);
            break;
        case FP4_B200_BF16:
            quantize_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
                (const __nv_bfloat16*)w, packed, absmax, n, blocksize);
            break;
        case FP4_B200_F32:
            quantize_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)w, packed,
                                                                     absmax, n, blocksize);
            break;
        default:
            return FP4_B200_ERR_DTYPE;
    }
    return (int)cudaGetLastError();
}

}  // namespace fp4b200
