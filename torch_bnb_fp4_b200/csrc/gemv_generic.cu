// Generic fused dequant + GEMV (batch 1..8), CUDA-core path.
//
// Used when the caller's codebook is not the bitsandbytes FP4 table, for blocksize 32, and as the
// A/B baseline for the integer tensor-core GEMV in gemv_imma.cu.  Replaces
// gemv_4bit_inference_kernel{,_float} (reference csrc/gemv_fp4_optimized.cu:60-259) with these
// differences: fp32 accumulation throughout (the reference accumulates in T),
// the caller's `code` tensor is honoured (the reference ignores it, :266), batch up to 8, bias fused.
//
// One warp per output row; a lane consumes 16 packed bytes (32 consecutive k, one absmax) per
// iteration with one 128-bit streaming load, factors the absmax out of the 32-term partial sum
//     y[r] = sum_blocks absmax_b * sum_{k in b} code[q_rk] * x[k]
// and the warp reduces with shuffles.  x is read through L1 (it is shared by every warp of the CTA).
#include "common.cuh"

namespace fp4b200 {

template <typename T>
__device__ __forceinline__ void load_x8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load_x8<__half>(const __half* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
template <>
__device__ __forceinline__ void load_x8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
template <>
__device__ __forceinline__ void load_x8<float>(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <typename T, int BATCH, bool NESTED>
__global__ void __launch_bounds__(256)
gemv_generic_kernel(const T* __restrict__ x, const uint8_t* __restrict__ packed,
                    const float* __restrict__ absmax, const NestedDev nd,
                    const float* __restrict__ code, const __grid_constant__ Code16 dflt,
                    const T* __restrict__ bias, T* __restrict__ out, const int N, const int K,
                    const int bs_log2) {
    __shared__ float s_code[16];
    if (threadIdx.x < 16)
        s_code[threadIdx.x] = code ? __ldg(code + threadIdx.x) : dflt.v[threadIdx.x];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= N) return;

    float acc[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) acc[b] = 0.f;

    const int64_t row_elem0 = (int64_t)row * K;
    const uint8_t* wrow = packed + (row_elem0 >> 1);

#pragma unroll 2
    for (int k0 = lane * 32; k0 < K; k0 += 32 * 32) {
        const uint4 q = ldg_stream_u4(wrow + (k0 >> 1));
        const float am = load_absmax<NESTED>(absmax, nd, (row_elem0 + k0) >> bs_log2);
        const uint32_t words[4] = {q.x, q.y, q.z, q.w};
        float part[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) part[b] = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            float c[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t byte = (words[w] >> (8 * j)) & 0xFFu;
                c[2 * j] = s_code[byte >> 4];
                c[2 * j + 1] = s_code[byte & 0xFu];
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                float xv[8];
                load_x8<T>(x + (int64_t)b * K + k0 + 8 * w, xv);
#pragma unroll
                for (int e = 0; e < 8; ++e) part[b] = fmaf(c[e], xv[e], part[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) acc[b] = fmaf(am, part[b], acc[b]);
    }

#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
        float v = acc[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            if (bias) v += DT<T>::to_f32(bias[row]);
            out[(int64_t)b * N + row] = DT<T>::from_f32(v);
        }
    }
}

template <typename T, bool NESTED>
static int launch_generic_b(const T* x, const uint8_t* packed, const float* absmax,
                            const NestedDev& nd, const float* code, const T* bias, T* out,
                            int batch, int N, int K, int bs_log2, cudaStream_t st) {
    const Code16 dflt = {FP4_B200_BNB_CODE_INIT};
    const unsigned blocks = (unsigned)((N + 7) / 8);
#define FP4_CASE(B)                                                                            \
    case B:                                                                                    \
        gemv_generic_kernel<T, B, NESTED><<<blocks, 256, 0, st>>>(x, packed, absmax, nd, code, \
                                                                  dflt, bias, out, N, K,       \
                                                                  bs_log2);                    \
        break;
    switch (batch) {
        FP4_CASE(1) FP4_CASE(2) FP4_CASE(3) FP4_CASE(4)
        FP4_CASE(5) FP4_CASE(6) FP4_CASE(7) FP4_CASE(8)
        default: return FP4_B200_ERR_BATCH;
    }
#undef FP4_CASE
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_generic_t(const void* x, const uint8_t* packed, const float* absmax,
                            const fp4_b200_nested_t* nested, const NestedDev& nd,
                            const float* code, const void* bias, void* out, int batch, int N,
                            int K, int bs_log2, cudaStream_t st) {
    return nested ? launch_generic_b<T, true>((const T*)x, packed, absmax, nd, code,
                                              (const T*)bias, (T*)out, batch, N, K, bs_log2, st)
                  : launch_generic_b<T, false>((const T*)x, packed, absmax, nd, code,
                                               (const T*)bias, (T*)out, batch, N, K, bs_log2, st);
}

int gemv_generic_dispatch(const void* x, const uint8_t* packed, const float* absmax,
                          const fp4_b200_nested_t* nested, const NestedDev& nd, const float* code,
                          const void* bias, void* out, int batch, int N, int K, int bs_log2,
                          int dtype, cudaStream_t st) {
    switch (dtype) {
        case FP4_B200_F16:
            return launch_generic_t<__half>(x, packed, absmax, nested, nd, code, bias, out, batch,
                                            N, K, bs_log2, st);
        case FP4_B200_BF16:
            return launch_generic_t<__nv_bfloat16>(x, packed, absmax, nested, nd, code, bias, out,
                                                   batch, N, K, bs_log2, st);
        case FP4_B200_F32:
            return launch_generic_t<float>(x, packed, absmax, nested, nd, code, bias, out, batch,
                                           N, K, bs_log2, st);
        default:
            return FP4_B200_ERR_DTYPE;
    }
}

}  // namespace fp4b200
