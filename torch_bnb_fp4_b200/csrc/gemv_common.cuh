// Device helpers shared by the two fused dequant-GEMV kernels (gemv_imma.cu: register-streamed,
// gemv_tma.cu: TMA-staged): nibble decode, MMA wrapper, x loaders, launch-constant division.
#pragma once
#include "common.cuh"

namespace fp4b200 {
namespace gemv {

// e5m2 bytes (= high byte of fp16) of 12*|code[i]|, i = 0..7: 0, 1/16, 8, 12 | 4, 6, 2, 3
constexpr uint32_t kTabLo = 0x4A482C00u;
constexpr uint32_t kTabHi = 0x42404644u;

__device__ __forceinline__ void unpack_e5m2x4(uint32_t m, uint32_t& h01, uint32_t& h23) {
    uint16_t lo, hi;
    asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(m));
    asm("cvt.rn.f16x2.e5m2x2 %0, %1;" : "=r"(h01) : "h"(lo));
    asm("cvt.rn.f16x2.e5m2x2 %0, %1;" : "=r"(h23) : "h"(hi));
}

// 8 nibbles (one 32-bit word of packed weights) -> 4 x half2 of 12*code[nibble].
// h[j] holds nibbles (2j, 2j+1) of the word = elements (2*byte+1, 2*byte) of packed byte j.
__device__ __forceinline__ void decode_word(uint32_t w, uint32_t tab_lo, uint32_t (&h)[4]) {
    const uint32_t wm = w & 0x77777777u;           // magnitude index of every nibble
    // the two shifts are done as integer multiplies so they issue on the FMA pipe; the ALU pipe
    // (PRMT / LOP3 / F2FP, half rate) is the one this kernel saturates
    const uint32_t w4 = w * 16u;                   // brings even nibbles' sign bits to byte msbs
    const uint32_t mag_lo = prmt(tab_lo, kTabHi, wm);                    // nibbles 0..3 -> bytes 0..3
    const uint32_t mag_hi = prmt(tab_lo, kTabHi, __umulhi(wm, 65536u));  // nibbles 4..7 (wm >> 16)
    // sign-replicate mode (selector msb): byte = 0xFF if the selected source byte is negative
    const uint32_t sg_lo = prmt(w, w4, 0x9D8Cu);  // signs of nibbles 0,1,2,3
    const uint32_t sg_hi = prmt(w, w4, 0xBFAEu);  // signs of nibbles 4,5,6,7
    const uint32_t m_lo = mag_lo | (sg_lo & 0x80808080u);
    const uint32_t m_hi = mag_hi | (sg_hi & 0x80808080u);
    unpack_e5m2x4(m_lo, h[0], h[1]);
    unpack_e5m2x4(m_hi, h[2], h[3]);
}

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

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

// pack 8 fp32 -> 8 fp16 in B-fragment order: within each packed byte the LOW nibble is the odd
// element, so half2 j = (elem 2j+1, elem 2j)
__device__ __forceinline__ uint4 pack_swapped(const float (&v)[8]) {
    uint4 r;
    r.x = DT<__half>::pack2(v[1], v[0]);
    r.y = DT<__half>::pack2(v[3], v[2]);
    r.z = DT<__half>::pack2(v[5], v[4]);
    r.w = DT<__half>::pack2(v[7], v[6]);
    return r;
}

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
// round-up method; exact for n < 2^31, which gemv_imma_supported guarantees for unit indices).
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

constexpr int kWarps = 8;     // warps per CTA that own work ranges
constexpr size_t kCounterBytes = 256 * 1024;  // fixed-size counter region: 65536 row tiles (N <= 2^20)

// Balanced flat partition of B units over W warps: the first r = B % W warps get q+1 units.
struct Partition {
    uint32_t q, r;
    FastDiv by_q, by_q1;
    __device__ __forceinline__ uint32_t begin(uint32_t w) const { return w * q + (w < r ? w : r); }
    __device__ __forceinline__ uint32_t owner(uint32_t u) const {  // warp whose range holds unit u
        const uint32_t big = r * (q + 1);
        return u < big ? by_q1.div(u) : r + by_q.div(u - big);
    }
};

struct Workspace {
    unsigned* counters;  // [kMaxTiles] units accounted for per row tile; zero between launches
    float* partials;     // [W][2][16][NC] fp32
};

// Everything the (rare, out-of-line) tile combine needs.
template <typename T, int NC>
struct FlushCtx {
    const float* sScale;
    float* sPart;        // [kWarps][2][16][NC]
    unsigned* sCnt;      // [kWarps]
    const T* bias;
    T* out;
    Workspace ws;
    Partition part;
    FastDiv by_nkb;
    uint32_t nkb, wid, first_tile, cta_L0, cta_L1, cta_w0;
    int batch, N;
};

// Combine / store the partial sums `acc` of (tile, k blocks [seg_start_kb, kb_end)).  Called once per
// row tile and warp, so it is kept out of the streaming loop (noinline).
template <int NCOLT>
struct AccV {
    float v[NCOLT][4];
};

template <typename T, int NCOLT>
__device__ __noinline__ void flush_tile(const FlushCtx<T, 8 * NCOLT>& c, const AccV<NCOLT> accv,
                                        uint32_t tile, uint32_t seg_start_kb, uint32_t kb_end) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    constexpr int NC = 8 * NCOLT;
    const float (&acc)[NCOLT][4] = accv.v;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = lane >> 2, t = lane & 3;
    const uint32_t seg_len = kb_end - seg_start_kb;
    const uint32_t row0 = tile * 16;
    auto store_out = [&](float v, int b, uint32_t row) {
        v *= c.sScale[b];
        if (c.bias) v += DT<T>::to_f32(c.bias[row]);
        c.out[(size_t)b * c.N + row] = DT<T>::from_f32(v);
    };
    if (seg_len == c.nkb) {
        // this warp covered the whole tile: finish directly.
        // acc[ct][i]: row g (+8 for i >= 2), column ct*8 + 2t + (i & 1)
        if constexpr (PIECES == 1) {
#pragma unroll
            for (int ct = 0; ct < NCOLT; ++ct)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int col = ct * 8 + 2 * t + (i & 1);
                    if (col < c.batch) store_out(acc[ct][i], col, row0 + g + ((i & 2) ? 8 : 0));
                }
        } else {
            // columns (2b, 2b+1) = (hi, lo) of batch row b live in the same lane
#pragma unroll
            for (int ct = 0; ct < NCOLT; ++ct) {
                const int b = ct * 4 + t;
                if (b < c.batch) {
                    store_out(acc[ct][0] + acc[ct][1] * (1.f / 2048.f), b, row0 + g);
                    store_out(acc[ct][2] + acc[ct][3] * (1.f / 2048.f), b, row0 + g + 8);
                }
            }
        }
        return;
    }
    // shared tile.  Contributors are the consecutive warps wa..wb whose ranges intersect it; each
    // parks its partial in its slot (0 if the tile is where its range starts, else 1).
    const uint32_t u_lo = tile * c.nkb, u_hi = u_lo + c.nkb;
    const uint32_t wa = c.part.owner(u_lo), wb = c.part.owner(u_hi - 1);
    const uint32_t slot = (tile == c.first_tile) ? 0u : 1u;
    const bool local = (u_lo >= c.cta_L0) && (u_hi <= c.cta_L1);  // all contributors in this CTA
    float* dst = local ? c.sPart + ((size_t)warp * 2 + slot) * 16 * NC
                       : c.ws.partials + ((size_t)c.wid * 2 + slot) * 16 * NC;
#pragma unroll
    for (int ct = 0; ct < NCOLT; ++ct) {
        float2* d2 = reinterpret_cast<float2*>(dst + ct * 8 + 2 * t);
        d2[(g * NC) / 2] = make_float2(acc[ct][0], acc[ct][1]);
        d2[((g + 8) * NC) / 2] = make_float2(acc[ct][2], acc[ct][3]);
    }
    if (local) __threadfence_block(); else __threadfence();
    __syncwarp();
    unsigned old = 0;
    if (lane == 0)
        old = local ? atomicAdd(c.sCnt + (wa - c.cta_w0), seg_len)
                    : atomicAdd(c.ws.counters + tile, seg_len);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old + seg_len != c.nkb) return;
    // last arriver: sum the contributors' slots in warp order (deterministic)
    if (local) __threadfence_block(); else __threadfence();
    // contributor wa starts at or before the tile (slot 0 only if it starts inside it); every later
    // contributor starts inside the tile (slot 0)
    const uint32_t slot_a = (c.part.begin(wa) >= u_lo) ? 0u : 1u;
    for (int idx = lane; idx < 16 * c.batch; idx += 32) {
        const int row = idx & 15, b = idx >> 4;
        float v = 0.f;
        for (uint32_t wc = wa; wc <= wb; ++wc) {
            const uint32_t cslot = (wc == wa) ? slot_a : 0u;
            float p0, p1 = 0.f;
            if (local) {
                const float* src = c.sPart + (((size_t)(wc - c.cta_w0) * 2 + cslot) * 16 + row) * NC + b * PIECES;
                p0 = src[0];
                if constexpr (PIECES == 2) p1 = src[1];
            } else {
                const float* src = c.ws.partials + (((size_t)wc * 2 + cslot) * 16 + row) * NC + b * PIECES;
                p0 = __ldcg(src);
                if constexpr (PIECES == 2) p1 = __ldcg(src + 1);
            }
            v += p0 + p1 * (1.f / 2048.f);
        }
        store_out(v, b, row0 + row);
    }
    if (lane == 0 && !local) c.ws.counters[tile] = 0;  // ready for the next launch
}

}  // namespace gemv
}  // namespace fp4b200
