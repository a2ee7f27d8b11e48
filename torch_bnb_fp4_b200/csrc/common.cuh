// Shared device helpers for libfp4_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fp4_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfp4_b200 is written for sm_100a (B200) only"
#endif

namespace fp4b200 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// bitsandbytes FP4 codebook == the literals of the reference tree decoder
// (reference csrc/dequant_fp4_optimized.cu:55-76); index = nibble.
#define FP4_B200_BNB_CODE_INIT                                                                \
    {0.00000000f,  5.208333333e-03f,  0.66666667f,  1.00000000f,  0.33333333f,  0.50000000f,  \
     0.16666667f,  0.25000000f,  -0.00000000f, -5.208333333e-03f, -0.66666667f, -1.00000000f, \
     -0.33333333f, -0.50000000f, -0.16666667f, -0.25000000f}

struct Code16 {
    float v[16];
};

// ---- dtype traits ---------------------------------------------------------------------------
template <typename T>
struct DT;
template <>
struct DT<__half> {
    static constexpr int code = FP4_B200_F16;
    static __device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }
    static __device__ __forceinline__ __half from_f32(float x) { return __float2half_rn(x); }
    // two fp32 -> packed 2x16-bit, round-to-nearest-even; `lo` lands in bits 0..15
    static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
};
template <>
struct DT<__nv_bfloat16> {
    static constexpr int code = FP4_B200_BF16;
    static __device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
    static __device__ __forceinline__ __nv_bfloat16 from_f32(float x) {
        return __float2bfloat16_rn(x);
    }
    static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
};
template <>
struct DT<float> {
    static constexpr int code = FP4_B200_F32;
    static __device__ __forceinline__ float to_f32(float x) { return x; }
    static __device__ __forceinline__ float from_f32(float x) { return x; }
};

// ---- streaming loads / stores ----------------------------------------------------------------
// Weights are read exactly once per call: bypass L1 allocation so x / absmax stay cached.
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// sm_100 256-bit store: one warp instruction writes 1 KiB contiguous.
__device__ __forceinline__ void stg_u8x32(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// nested absmax decode, two separately rounded fp32 ops
struct NestedDev {
    const uint8_t* qabsmax;
    const float* code2;
    const float* absmax2;
    float offset;
    int blocksize2_log2;
};
__device__ __forceinline__ float nested_absmax(const NestedDev& nd, int64_t blk) {
    const float c = __ldg(nd.code2 + __ldg(nd.qabsmax + blk));
    const float a2 = __ldg(nd.absmax2 + (blk >> nd.blocksize2_log2));
    return __fadd_rn(__fmul_rn(c, a2), nd.offset);
}

template <bool NESTED>
__device__ __forceinline__ float load_absmax(const float* absmax, const NestedDev& nd, int64_t blk) {
    if constexpr (NESTED) {
        return nested_absmax(nd, blk);
    } else {
        return __ldg(absmax + blk);
    }
}

// GEMV workspace: per-row-tile unit counters first (must be zero before the first launch; the kernel
// leaves them zero), then the fp32 partial-sum slots.
static inline size_t fp4b200_ws_counter_bytes(int N) {
    return (((size_t)(N + 15) / 16) * 4 + 255) & ~(size_t)255;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a once-flag per device, not per process
struct PerDeviceOnce {
    bool done[64] = {};
    bool& flag() {
        int d = 0;
        cudaGetDevice(&d);
        return done[(d < 0 || d >= 64) ? 0 : d];
    }
};

static inline int ilog2_exact(int64_t v) {  // -1 if v is not a power of two
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((int64_t(1) << l) < v) ++l;
    return l;
}

}  // namespace fp4b200
