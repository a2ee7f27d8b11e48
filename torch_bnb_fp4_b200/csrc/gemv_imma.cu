// Fused dequant + GEMV for decode (batch 1..8) on sm_100a — the fast path.
//
// Replaces gemv_4bit_inference_kernel{,_float} (reference csrc/gemv_fp4_optimized.cu:60-259).
// Bound: HBM.  Algorithmic bytes per weight: 0.5 (packed) + 4/blocksize (absmax) = 0.5625 at
// blocksize 64.  At the measured 6.56 TB/s one SM must retire ~41 weights per clock, which leaves
// ~3 issue slots per weight in total; a per-nibble shared-memory lookup + FFMA (the reference's
// scheme: 32 LDS + 32 HMUL2 + 32 HFMA2 per 16 bytes) does not fit.  This kernel therefore
//   * decodes FOUR nibbles per instruction with PRMT used as an 8-entry byte table: the
//     bitsandbytes FP4 magnitudes times 12 are {0, 1/16, 8, 12, 4, 6, 2, 3} - exactly representable
//     in e5m2, i.e. one byte each, the high byte of their fp16 encoding;
//   * merges the sign bits in with one more PRMT (sign-replicate mode) + one LOP3 per four nibbles;
//   * widens e5m2 -> fp16 with F2FP (cvt.rn.f16x2.e5m2x2: the byte becomes the high byte, exact);
//   * feeds the fp16 pairs to mma.sync.m16n8k16 (fp32 accumulate) as the A operand: 16 weight rows x
//     16 k per instruction, B = x (8 columns = up to 8 batch rows, so batch 2..8 costs no extra
//     decode work).  This is not a reshaping of the problem into a GEMM: it is the same contraction,
//     with the multiply-adds moved off the issue-limited FMA pipe.
// The absmax is factored out of the inner sum, y[r] = sum_b absmax[r,b] * sum_{k in b} c12[q]*x[k],
// so products are exact (fp16 x fp16 in fp32) and each 64-element block costs 4 FFMA per lane.
// x is staged ONCE per CTA in shared memory as fp16, pre-scaled by a power of two per batch row so
// that bf16/fp32 inputs cannot overflow fp16 (fp32 inputs are split hi + lo into two columns, ~22
// bits), and stored in mma B-fragment order so a lane fetches its operands with two LDS.128 per
// block.  Weights stream with 64-bit ld.global.nc.L1::no_allocate loads, U blocks in flight per lane
// (each warp instruction reads whole 32-byte sectors of 8 rows; a warp's U consecutive loads cover
// whole 128-byte lines), issued before the x staging so the first HBM round trip is overlapped.
//
// Grid: one CTA of 8 warps covers 16*RW rows x all K, warps arranged RW (row tiles) x KW (k split),
// partial sums combined through shared memory; RW/KW are chosen per shape so that small layers still
// put >= ~2 CTAs on each of the 148 SMs.
//
// Requirements (checked by gemv_imma_supported): codebook == bitsandbytes FP4 table, K % 64 == 0,
// blocksize % 64 == 0, N % 16 == 0, shared memory for x <= 200 KB.  Everything else takes
// gemv_generic.cu.
#include "common.cuh"

#include <cstdlib>

namespace fp4b200 {

namespace {

constexpr int kU = 4;  // 64-element blocks in flight per lane (per row)

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
__device__ __forceinline__ void decode_word(uint32_t w, uint32_t (&h)[4]) {
    const uint32_t wm = w & 0x77777777u;           // magnitude index of every nibble
    const uint32_t w4 = w << 4;                    // brings even nibbles' sign bits to byte msbs
    const uint32_t mag_lo = prmt(kTabLo, kTabHi, wm);        // nibbles 0..3 -> bytes 0..3
    const uint32_t mag_hi = prmt(kTabLo, kTabHi, wm >> 16);  // nibbles 4..7
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

// T: activation dtype.  PIECES = 1 (fp16/bf16 x: one fp16 column per batch row) or 2 (fp32 x: hi + lo).
// NCOLT = number of 8-column MMA tiles (1, or 2 when batch*PIECES > 8).
template <typename T, int NCOLT, bool NESTED>
__global__ void __launch_bounds__(256)
gemv_mma_kernel(const T* __restrict__ x, const uint8_t* __restrict__ packed,
                const float* __restrict__ absmax, const NestedDev nd, const T* __restrict__ bias,
                T* __restrict__ out, const int batch, const int N, const int K, const int bs_log2,
                const int kw_log2) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int nkb = K >> 6;
    const int ncols = batch * PIECES;
    uint4* sB = reinterpret_cast<uint4*>(smem_raw);                  // [nkb][2][ncols][4] x 16 B
    float* sRed = reinterpret_cast<float*>(sB + (size_t)nkb * 2 * ncols * 4);  // [8][16][8*NCOLT]
    float* sMax = sRed + 8 * 16 * 8 * NCOLT;                         // [8 warps][8 batch]
    float* sScale = sMax + 64;                                       // [8 batch] inverse scales

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int KW = 1 << kw_log2, RW = 8 >> kw_log2;
    const int rt = warp >> kw_log2, kwi = warp & (KW - 1);
    const int rowbase = (blockIdx.x * RW + rt) * 16;
    const bool tile_valid = rowbase < N;
    const int r0 = tile_valid ? rowbase + g : g;  // clamp: loads stay in bounds, results unused
    const int r1 = r0 + 8;
    const int per = (nkb + KW - 1) >> kw_log2;
    const int kb_begin = kwi * per;
    const int kb_end = (kb_begin + per < nkb) ? kb_begin + per : nkb;

    const uint8_t* p0 = packed + (((int64_t)r0 * K) >> 1) + 8 * t;
    const uint8_t* p1 = packed + (((int64_t)r1 * K) >> 1) + 8 * t;
    const int64_t e0 = (int64_t)r0 * K, e1 = (int64_t)r1 * K;

    // ---- 1. put the first U blocks of the weight stream in flight ------------------------------
    uint2 q0[kU], q1[kU];
    float a0[kU], a1[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        const int kb = kb_begin + u;
        if (kb < kb_end) {
            q0[u] = ldg_stream_u2(p0 + kb * 32);
            q1[u] = ldg_stream_u2(p1 + kb * 32);
            a0[u] = load_absmax<NESTED>(absmax, nd, (e0 + (int64_t)kb * 64) >> bs_log2);
            a1[u] = load_absmax<NESTED>(absmax, nd, (e1 + (int64_t)kb * 64) >> bs_log2);
        }
    }

    // ---- 2. stage x as scaled fp16 in B-fragment order -----------------------------------------
    const int nchunk = K >> 3;  // 8-element chunks per batch row
    {   // pass 1: max |x| per batch row
        float mx[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) mx[b] = 0.f;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < batch) {
                for (int c = tid; c < nchunk; c += 256) {
                    float f[8];
                    XLoad<T>::load(x + (int64_t)b * K + c * 8, f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) mx[b] = fmaxf(mx[b], fabsf(f[i]));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    mx[b] = fmaxf(mx[b], __shfl_xor_sync(0xffffffffu, mx[b], o));
                if (lane == 0) sMax[warp * 8 + b] = mx[b];
            }
        }
        __syncthreads();
    }
    float scale[8];  // power of two bringing max|x| into [2^13, 2^14)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        scale[b] = 0.f;
        if (b < batch) {
            float m = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) m = fmaxf(m, sMax[w * 8 + b]);
            int E = (int)((__float_as_uint(m) >> 23) & 0xFFu);
            E = E < 14 ? 14 : (E > 254 ? 254 : E);
            scale[b] = __uint_as_float((uint32_t)(267 - E) << 23);
            if (tid == 0) sScale[b] = __uint_as_float((uint32_t)(E - 13) << 23);  // 2^-(13-e)
        }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b < batch) {
            for (int c = tid; c < nchunk; c += 256) {
                float f[8];
                XLoad<T>::load(x + (int64_t)b * K + c * 8, f);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] *= scale[b];
                const int kb = c >> 3, qq = c & 7;  // chunk qq of the block: lane t = qq/2, half = qq%2
                uint4* dst = sB + ((size_t)(kb * 2 + (qq & 1)) * ncols) * 4 + (qq >> 1);
                dst[(b * PIECES) * 4] = pack_swapped(f);
                if constexpr (PIECES == 2) {
                    float lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        lo[i] = (f[i] - __half2float(__float2half_rn(f[i]))) * 2048.f;
                    dst[(b * PIECES + 1) * 4] = pack_swapped(lo);
                }
            }
        }
    }
    __syncthreads();

    // ---- 3. main loop: decode + MMA, U blocks in flight ----------------------------------------
    float acc[NCOLT][4];
#pragma unroll
    for (int c = 0; c < NCOLT; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;

    for (int kb0 = kb_begin; kb0 < kb_end; kb0 += kU) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int kb = kb0 + u;
            if (kb < kb_end) {
                const uint2 w0 = q0[u], w1 = q1[u];
                const float am0 = a0[u], am1 = a1[u];
                const int kn = kb + kU;
                if (kn < kb_end) {  // refill this slot
                    q0[u] = ldg_stream_u2(p0 + kn * 32);
                    q1[u] = ldg_stream_u2(p1 + kn * 32);
                    a0[u] = load_absmax<NESTED>(absmax, nd, (e0 + (int64_t)kn * 64) >> bs_log2);
                    a1[u] = load_absmax<NESTED>(absmax, nd, (e1 + (int64_t)kn * 64) >> bs_log2);
                }
                uint32_t ha[2][4], hb[2][4];  // [word][half2] for rows r0 / r1
                decode_word(w0.x, ha[0]);
                decode_word(w0.y, ha[1]);
                decode_word(w1.x, hb[0]);
                decode_word(w1.y, hb[1]);
#pragma unroll
                for (int ct = 0; ct < NCOLT; ++ct) {
                    const int col = ct * 8 + g;
                    uint4 bA = make_uint4(0, 0, 0, 0), bB = make_uint4(0, 0, 0, 0);
                    if (col < ncols) {
                        const uint4* src = sB + ((size_t)(kb * 2) * ncols + col) * 4 + t;
                        bA = src[0];                       // k16 groups 0,1 (x elements 0..7 of the lane's 16)
                        bB = src[(size_t)ncols * 4];       // k16 groups 2,3
                    }
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
                    mma16816(d, ha[0][0], hb[0][0], ha[0][1], hb[0][1], bA.x, bA.y);
                    mma16816(d, ha[0][2], hb[0][2], ha[0][3], hb[0][3], bA.z, bA.w);
                    mma16816(d, ha[1][0], hb[1][0], ha[1][1], hb[1][1], bB.x, bB.y);
                    mma16816(d, ha[1][2], hb[1][2], ha[1][3], hb[1][3], bB.z, bB.w);
                    acc[ct][0] = fmaf(am0, d[0], acc[ct][0]);
                    acc[ct][1] = fmaf(am0, d[1], acc[ct][1]);
                    acc[ct][2] = fmaf(am1, d[2], acc[ct][2]);
                    acc[ct][3] = fmaf(am1, d[3], acc[ct][3]);
                }
            }
        }
    }

    // ---- 4. combine the k-split warps and the pieces, add bias, store ----------------------------
    constexpr int RC = 8 * NCOLT;
#pragma unroll
    for (int ct = 0; ct < NCOLT; ++ct) {
        float* dst = sRed + (warp * 16) * RC + ct * 8 + 2 * t;
        dst[g * RC] = acc[ct][0];
        dst[g * RC + 1] = acc[ct][1];
        dst[(g + 8) * RC] = acc[ct][2];
        dst[(g + 8) * RC + 1] = acc[ct][3];
    }
    __syncthreads();
    // one thread per (row tile, row, batch row)
    for (int idx = tid; idx < RW * 16 * batch; idx += 256) {
        const int b = idx % batch;
        const int row = (idx / batch) & 15;
        const int rtile = idx / (batch * 16);
        const int grow = (blockIdx.x * RW + rtile) * 16 + row;
        if (grow < N) {
            float v = 0.f;
            for (int kw = 0; kw < KW; ++kw) {
                const float* src = sRed + (((rtile << kw_log2) + kw) * 16 + row) * RC + b * PIECES;
                if constexpr (PIECES == 2) v += src[0] + src[1] * (1.f / 2048.f);
                else v += src[0];
            }
            v *= sScale[b] * (1.f / 12.f);
            if (bias) v += DT<T>::to_f32(bias[grow]);
            out[(int64_t)b * N + grow] = DT<T>::from_f32(v);
        }
    }
}

struct Plan {
    int kw_log2;
    int grid;
    size_t smem;
};

static size_t smem_bytes(int batch, int K, int pieces, int ncolt) {
    const size_t nkb = K / 64;
    return nkb * 2 * (size_t)(batch * pieces) * 4 * 16 + (size_t)8 * 16 * 8 * ncolt * 4 + 64 * 4 +
           8 * 4;
}

static Plan make_plan(int batch, int N, int K, int pieces, int ncolt) {
    static const int force_kw = [] {
        const char* e = getenv("FP4_B200_GEMV_KW_LOG2");
        return e ? atoi(e) : -1;
    }();
    const int ntiles = N / 16;
    const int nkb = K / 64;
    int kw_log2 = 0;
    // prefer many rows per CTA (x staging is amortised over rows) while keeping >= ~2 CTAs per SM
    // and at least kU blocks of work per warp
    while (kw_log2 < 3) {
        const int rw = 8 >> kw_log2;
        const int ctas = (ntiles + rw - 1) / rw;
        if (ctas >= 2 * kNumSMs) break;
        if ((nkb >> (kw_log2 + 1)) < kU) break;
        ++kw_log2;
    }
    if (force_kw >= 0 && force_kw <= 3) kw_log2 = force_kw;
    const int rw = 8 >> kw_log2;
    return Plan{kw_log2, (ntiles + rw - 1) / rw, smem_bytes(batch, K, pieces, ncolt)};
}

template <typename T, int NCOLT, bool NESTED>
static int launch(const void* x, const uint8_t* packed, const float* absmax, const NestedDev& nd,
                  const void* bias, void* out, int batch, int N, int K, int bs_log2,
                  cudaStream_t st) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    const Plan p = make_plan(batch, N, K, PIECES, NCOLT);
    auto kern = gemv_mma_kernel<T, NCOLT, NESTED>;
    if (p.smem > 48 * 1024) {
        static size_t configured = 0;  // per template instantiation
        if (p.smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 200 * 1024);
            if (e != cudaSuccess) return (int)e;
            configured = 200 * 1024;
        }
    }
    kern<<<p.grid, 256, p.smem, st>>>((const T*)x, packed, absmax, nd, (const T*)bias, (T*)out,
                                      batch, N, K, bs_log2, p.kw_log2);
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_t(const void* x, const uint8_t* packed, const float* absmax, bool nested,
                    const NestedDev& nd, const void* bias, void* out, int batch, int N, int K,
                    int bs_log2, cudaStream_t st) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    const bool two = batch * PIECES > 8;
    if (nested)
        return two ? launch<T, 2, true>(x, packed, absmax, nd, bias, out, batch, N, K, bs_log2, st)
                   : launch<T, 1, true>(x, packed, absmax, nd, bias, out, batch, N, K, bs_log2, st);
    return two ? launch<T, 2, false>(x, packed, absmax, nd, bias, out, batch, N, K, bs_log2, st)
               : launch<T, 1, false>(x, packed, absmax, nd, bias, out, batch, N, K, bs_log2, st);
}

}  // namespace

bool gemv_imma_supported(int batch, int N, int K, int blocksize, int dtype) {
    if (batch < 1 || batch > 8 || N <= 0 || K <= 0) return false;
    if (K % 64 != 0 || blocksize % 64 != 0 || N % 16 != 0) return false;
    const int pieces = dtype == FP4_B200_F32 ? 2 : 1;
    const int ncolt = batch * pieces > 8 ? 2 : 1;
    return smem_bytes(batch, K, pieces, ncolt) <= 200 * 1024;
}

int gemv_imma_dispatch(const void* x, const uint8_t* packed, const float* absmax,
                       const fp4_b200_nested_t* nested, const NestedDev& nd, const void* bias,
                       void* out, int batch, int N, int K, int bs_log2, int dtype,
                       cudaStream_t st) {
    const bool nst = nested != nullptr;
    switch (dtype) {
        case FP4_B200_F16:
            return launch_t<__half>(x, packed, absmax, nst, nd, bias, out, batch, N, K, bs_log2, st);
        case FP4_B200_BF16:
            return launch_t<__nv_bfloat16>(x, packed, absmax, nst, nd, bias, out, batch, N, K,
                                           bs_log2, st);
        case FP4_B200_F32:
            return launch_t<float>(x, packed, absmax, nst, nd, bias, out, batch, N, K, bs_log2, st);
        default:
            return FP4_B200_ERR_DTYPE;
    }
}

}  // namespace fp4b200
