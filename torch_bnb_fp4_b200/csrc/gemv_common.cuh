// Device helpers of the fused dequant-GEMV (gemv_stream.cu): x loaders, shared-memory loads by 32-bit address,
// launch-constant division.
#pragma once
#include "common.cuh"

namespace fp4b200 {
namespace gemv {

template <typename T>
struct XLoad;  // 8 consecutive x elements -> fp32
template <>
struct XLoad<__half> {
    static __device__ __forceinline__ void load(const __half* p, float (&f)[8]) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
};
template <>
struct XLoad<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
};
template <>
struct XLoad<float> {
    static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};

__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(saddr));
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(saddr));
    return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t saddr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr));
    return r;
}

// Division by a launch-time constant without the ~40-instruction udiv sequence (Granlund-Montgomery
// round-up method; exact for n < 2^31, which gemv_stream_supported guarantees for unit indices).
struct FastDiv {
    uint32_t d, mul, shift;
    FastDiv() = default;
    explicit FastDiv(uint32_t div) : d(div) {
        if (div <= 1) { mul = 0; shift = 0; return; }
        uint32_t l = 0;
        while ((1ull << l) < div) ++l;   // ceil(log2 d)
        shift = l - 1;
        mul = (uint32_t)(((1ull << (32 + shift)) + div - 1) / div);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        return d <= 1 ? n : (__umulhi(n, mul) >> shift);
    }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};

}  // namespace gemv
}  // namespace fp4b200
