// Fused dequant + GEMV for decode (batch 1..8) on sm_100a — integer tensor-core formulation.
//
//   out[b, r] = T( sum_k x[b,k] * code[W[r,k]] * absmax[r, k/64] + bias[r] )
// replaces the reference's gemv_4bit_inference kernels (csrc/gemv_fp4_optimized.cu:60-259).
//
// Why integers.  At 6.5 TB/s one SM has to retire ~40 weights per clock, and the half-rate ALU pipe
// (PRMT / LOP3 / SHF) is the first thing that saturates: any per-weight decode to fp16 costs >= 11 ALU
// instructions per 8 weights (lookup, sign merge, byte -> half widening).  On sm_100a mma.sync with
// fp8 operands is emulated (F2FP + HMMA), but IMMA.16832.U8.S8 is native.  So:
//   * weights:  192*|code| = {0, 1, 128, 192, 64, 96, 32, 48} fits a byte.  ONE PRMT against that
//     8-entry table turns four nibbles into four u8 magnitudes ("ALL"); a second PRMT in sign-replicate
//     mode gives the 0x00/0xFF sign masks, ALL & mask = the magnitudes of the negative weights ("NEG").
//     7 ALU-pipe instructions per 8 weights, no widening, nothing else.
//   * x:  per 64-block and batch row a power-of-two scale brings |x| below 64, then
//     x*s = t1 + t2/128 (+ t3/128^2 + t4/128^3 for fp32 inputs) with s8 integers t_j (residual
//     expansion; exact to 2^-14 / 2^-28 of the block maximum).  The terms sit in different MMA columns.
//   * sum_k w*x = (1/192) * absmax * 2^-e * ( IMMA(ALL, t) - 2*IMMA(NEG, t) ): exact integer dot products,
//     converted with the 1.5*2^23 trick and scaled once per (row, 64-block) in fp32.
// The contraction is unchanged (one weight row x one activation vector); the tensor core is only the
// multiply-add engine that takes the FMAs off the issue-limited pipes.
//
// Weight stream: every warp owns a small shared-memory ring fed by its lane 0 with cp.async.bulk.tensor:
// one box = [16 rows x 128 B] packed weights (hardware 128-byte swizzle) + [16 rows x 4] fp32 absmax =
// one "item" (16 rows x 256 k).  The ring starts filling before griddepcontrol.wait, so under
// programmatic dependent launch the weights of layer i+1 stream while layer i drains.
//
// Schedule: the flat item sequence (row-tile major) is cut into equal contiguous ranges per CTA and,
// inside the CTA, per warp.  Row tiles shared between warps are combined deterministically in shared
// memory (slots summed in warp order by the last arriver), tiles shared between CTAs through the
// workspace the same way.
//
// Requirements (gemv_i8_supported): bitsandbytes FP4 codebook, blocksize 64, fp32 absmax, K % 256 == 0.
#include <cuda.h>

#include <cstdlib>

#include "gemv_common.cuh"

namespace fp4b200 {

#ifdef FP4_I8_TIMELINE
long long* g_i8_tl = nullptr;
int g_i8_tl_launch = 0;
#endif

namespace {

using gemv::FastDiv;
using gemv::XLoad;
using gemv::lds_u2;
using gemv::lds_u4;

#ifndef FP4_I8_WARPS
#define FP4_I8_WARPS 16
#endif
#ifndef FP4_I8_STAGES
#define FP4_I8_STAGES 2
#endif
constexpr int kW = FP4_I8_WARPS;  // warps per CTA (power of two)
constexpr int kThreads = kW * 32;
constexpr int kStages = FP4_I8_STAGES;
constexpr uint32_t kWBox = 16 * 128;  // packed weights of one item
constexpr uint32_t kABox = 16 * 16;   // 4 fp32 absmax for each of 16 rows
constexpr uint32_t kItemBytes = kWBox + kABox;
constexpr size_t kCounterBytes = gemv::kCounterBytes;  // shared with gemv_imma.cu: zero between launches
constexpr float kMagic = 12582912.f;  // 1.5 * 2^23
// 192*|code[0..3]| = 0, 1, 128, 192 is 0xC0800100 (loaded into a register in the kernel)
constexpr uint32_t kTabHi = 0x30206040u;  // 192*|code[4..7]| = 64, 96, 32, 48

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(mbar)
        : "memory");
}
__device__ __forceinline__ void imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                     uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// One 32-bit word = 8 nibbles -> u8 magnitudes of nibbles 0..3 / 4..7 (ALL) and the same with the
// non-negative weights zeroed (NEG).  Byte j of *_lo is nibble j, i.e. element (j ^ 1) of the word.
__device__ __forceinline__ void decode_word(uint32_t w, uint32_t tab_lo, uint32_t& all_lo, uint32_t& all_hi,
                                            uint32_t& neg_lo, uint32_t& neg_hi) {
    const uint32_t wm = w & 0x77777777u;
    const uint32_t w4 = w * 16u;  // integer multiply: issues on the FMA pipe, not the saturated ALU pipe
    all_lo = prmt(tab_lo, kTabHi, wm);
    all_hi = prmt(tab_lo, kTabHi, __umulhi(wm, 65536u));
    // sign-replicate mode (selector msb): byte = 0xFF if the selected source byte has its msb set
    neg_lo = all_lo & prmt(w, w4, 0x9D8Cu);  // signs of nibbles 0,1,2,3
    neg_hi = all_hi & prmt(w, w4, 0xBFAEu);  // signs of nibbles 4,5,6,7
}

struct Params {
    const void* x;
    const void* bias;
    void* out;
    unsigned* gcnt;   // [tiles] items accounted for per row tile shared between CTAs; zero between launches
    float* gpart;     // [grid][2][128] fp32
    int batch, N, K;
    uint32_t ipt;     // items per row tile = K / 256
    uint32_t items;   // tiles * ipt
    uint32_t q, r;    // CTA c owns items [c*q + min(c,r), +q + (c<r))
    uint32_t ntl;     // local tile counters per CTA
    FastDiv by_ipt, by_q, by_q1;
    long long* tl;    // debug timeline (FP4_I8_TIMELINE builds): [warp][8] globaltimer ns, else unused
};

#ifdef FP4_I8_TIMELINE
#define TL_STAMP(i)                                                                    \
    do {                                                                               \
        if (p.tl && lane == 0) {                                                       \
            long long gt_;                                                             \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                    \
            p.tl[((size_t)blockIdx.x * kW + warp) * 8 + (i)] = gt_;                    \
        }                                                                              \
    } while (0)
#else
#define TL_STAMP(i) do {} while (0)
#endif

struct Ctx {
    const Params* p;
    float* sPart;     // [ntl][kW][batch*16] per-(tile, warp) partial sums, zero-initialised
    unsigned* sCnt;   // [ntl] items accounted for per local tile
    uint32_t cta_b, cta_e, t_first;
};

template <typename T>
__device__ __forceinline__ void store_out(const Params& p, float v, int b, uint32_t row) {
    if (row >= (uint32_t)p.N) return;
    const T* bias = reinterpret_cast<const T*>(p.bias);
    if (bias) v += DT<T>::to_f32(bias[row]);
    reinterpret_cast<T*>(p.out)[(size_t)b * p.N + row] = DT<T>::from_f32(v);
}

__device__ __forceinline__ uint32_t cta_begin(const Params& p, uint32_t c) {
    return c * p.q + (c < p.r ? c : p.r);
}
__device__ __forceinline__ uint32_t cta_owner(const Params& p, uint32_t item) {
    const uint32_t big = p.r * (p.q + 1);
    return item < big ? p.by_q1.div(item) : p.r + p.by_q.div(item - big);
}

// Park this warp's partial sums of `cnt` items of row tile `tile`; the warp that completes the CTA's
// share of the tile sums the slots in warp order and stores the rows (tile entirely inside the CTA)
// or publishes the CTA partial for the neighbouring CTA (last CTA sums in CTA order).  v[ct][h]: row
// rho + 8h, batch row owned by this lane in column tile ct.  Runs once per (warp, tile): out of line.
template <typename T, int NT, int NCT>
__device__ __noinline__ void flush_tile(const Ctx& c, const float (&v)[NCT][2], uint32_t tile, uint32_t cnt) {
    const Params& p = *c.p;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = lane >> 2, t = lane & 3;
    const uint32_t rho = ((g & 3) << 1) | (g >> 2);
    const uint32_t row0 = tile * 16;
    const int nb = p.batch;
    const bool owner_lane = (NT == 2) || ((t & 1) == 0);
    auto lane_batch = [&](int ct) { return NT == 2 ? ct * 4 + (int)t : ct * 2 + (int)(t >> 1); };
    if (cnt == p.ipt) {  // the warp covered the whole tile by itself
        if (owner_lane) {
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                const int b = lane_batch(ct);
                if (b < nb) {
                    store_out<T>(p, v[ct][0], b, row0 + rho);
                    store_out<T>(p, v[ct][1], b, row0 + rho + 8);
                }
            }
        }
        return;
    }
    const uint32_t tl = tile - c.t_first;
    float* set = c.sPart + (size_t)tl * kW * 16 * nb;
    float* mine = set + (size_t)warp * 16 * nb;
    if (owner_lane) {
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) {
            const int b = lane_batch(ct);
            if (b < nb) {
                mine[b * 16 + rho] = v[ct][0];
                mine[b * 16 + rho + 8] = v[ct][1];
            }
        }
    }
    __threadfence_block();
    __syncwarp();
    const uint32_t tile_b = tile * p.ipt, tile_e = tile_b + p.ipt;
    const uint32_t lo = tile_b > c.cta_b ? tile_b : c.cta_b, hi = tile_e < c.cta_e ? tile_e : c.cta_e;
    unsigned old = 0;
    if (lane == 0) old = atomicAdd(c.sCnt + tl, cnt);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old + cnt != hi - lo) return;
    // last warp of this CTA on the tile: sum the slots in warp order (warps without a share left zeros)
    __threadfence_block();
    const bool whole = (hi - lo == p.ipt);
    const uint32_t gslot = (c.cta_b >= tile_b) ? 0u : 1u;
    float* gmine = p.gpart + ((size_t)blockIdx.x * 2 + gslot) * 128;
    for (int idx = lane; idx < 16 * nb; idx += 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kW; ++w) s += set[(size_t)w * 16 * nb + idx];
        if (whole) store_out<T>(p, s, idx >> 4, row0 + (idx & 15));
        else __stcg(gmine + idx, s);
    }
    if (whole) return;
    // the tile continues in a neighbouring CTA: the last CTA to arrive sums the CTA partials in CTA order
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    __syncwarp();
    if (lane == 0) old = atomicAdd(p.gcnt + tile, hi - lo);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old + (hi - lo) != p.ipt) return;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    const uint32_t ca = cta_owner(p, tile_b), cb = cta_owner(p, tile_e - 1);
    for (int idx = lane; idx < 16 * nb; idx += 32) {
        float s = 0.f;
        for (uint32_t cc = ca; cc <= cb; ++cc) {
            const uint32_t gs = (cta_begin(p, cc) >= tile_b) ? 0u : 1u;
            s += __ldcg(p.gpart + ((size_t)cc * 2 + gs) * 128 + idx);
        }
        store_out<T>(p, s, idx >> 4, row0 + (idx & 15));
    }
    if (lane == 0) p.gcnt[tile] = 0;  // ready for the next launch
}

template <typename T, int NT, int NCT>
__global__ void __launch_bounds__(kThreads, (kW <= 8 ? 4 : 2))
gemv_i8_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
               const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t g = lane >> 2, t = lane & 3;
    const uint32_t rho = ((g & 3) << 1) | (g >> 2);  // MMA row g <-> box row rho: conflict-free LDS.64
    const int batch = p.batch;
    const int ncols = batch * NT;
    const uint32_t nkb = (uint32_t)p.K >> 6;

    // ---- shared memory carve-up -----------------------------------------------------------------
    uint8_t* sp = smem;
    const uint32_t ringW = (uint32_t)__cvta_generic_to_shared(sp) + warp * kStages * kWBox;
    sp += kW * kStages * kWBox;
    const uint32_t ringA = (uint32_t)__cvta_generic_to_shared(sp) + warp * kStages * kABox;
    sp += kW * kStages * kABox;
    uint8_t* sX = sp;                       // [nkb][ncols][64] s8, pairs of k swapped (nibble order)
    sp += (size_t)nkb * ncols * 64;
    float* sXs = reinterpret_cast<float*>(sp);  // [batch][nkb] 2^-e / 192
    sp += (size_t)batch * nkb * 4;
    float* sPart = reinterpret_cast<float*>(sp);
    const uint32_t npart = p.ntl * kW * 16 * batch;
    sp += (size_t)npart * 4;
    unsigned* sCnt = reinterpret_cast<unsigned*>(sp);
    sp += (size_t)p.ntl * 4;
    sp = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sp) + 7) & ~(uintptr_t)7);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(sp) + warp * kStages * 8;

    TL_STAMP(0);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    }
    __syncwarp();

    // ---- this CTA's item range, dealt round-robin to its warps ------------------------------------
    // Sequence position s of the CTA (warp w takes s = w, w + kW, ...) -> item: the items of the two
    // row tiles shared with the neighbouring CTAs come first, so their cross-CTA combine overlaps the
    // rest of the stream; then the tiles owned outright, in memory order.  At any moment the boxes in
    // flight from one CTA cover one contiguous stretch of the weight matrix (DRAM page locality).
    const uint32_t cta = blockIdx.x;
    const uint32_t cta_b = cta_begin(p, cta);
    const uint32_t n_c = p.q + (cta < p.r ? 1u : 0u);
    const uint32_t cta_e = cta_b + n_c;
    uint32_t t_first, head_off;
    p.by_ipt.divmod(cta_b, t_first, head_off);
    uint32_t nh = head_off ? p.ipt - head_off : 0u;
    nh = nh < n_c ? nh : n_c;
    uint32_t nt_q, nt;
    p.by_ipt.divmod(n_c - nh, nt_q, nt);
    const uint32_t n = n_c > (uint32_t)warp ? (n_c - warp + kW - 1) / kW : 0u;  // items of this warp
    auto item_of = [&](uint32_t s) {
        return s < nh ? cta_b + s : (s < nh + nt ? cta_e - nt + (s - nh) : cta_b + (s - nt));
    };

    uint32_t st_tile[kStages], st_kq[kStages];
    uint32_t ld_s = warp, ld_left = n;
    auto issue_box = [&](int slot) {  // slot is a compile-time constant at every call site
        if (ld_left == 0) return;
        uint32_t tile, kq;
        p.by_ipt.divmod(item_of(ld_s), tile, kq);
        st_tile[slot] = tile;
        st_kq[slot] = kq;
        if (lane == 0) {
            const uint32_t bar = bar0 + slot * 8;
#ifdef FP4_I8_NOABSMAX  // experiment: weights only
            mbar_expect_tx(bar, kWBox);
            tma_load_2d(ringW + slot * kWBox, &tmW, (int)(kq * 128), (int)(tile * 16), bar);
#else
            mbar_expect_tx(bar, kItemBytes);
            tma_load_2d(ringW + slot * kWBox, &tmW, (int)(kq * 128), (int)(tile * 16), bar);
            tma_load_2d(ringA + slot * kABox, &tmA, (int)(kq * 4), (int)(tile * 16), bar);
#endif
        }
        ld_s += kW;
        --ld_left;
    };
    // ---- 1. fill the ring: weights never depend on the previous kernel in the stream --------------
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
        st_tile[s] = 0;
        st_kq[s] = 0;
        issue_box(s);
    }
    for (uint32_t i = tid; i < npart; i += kThreads) sPart[i] = 0.f;
    for (uint32_t i = tid; i < p.ntl; i += kThreads) sCnt[i] = 0;
    TL_STAMP(1);

    // x, the workspace and `out` may be products of the previous kernel: wait for it, then let the
    // next kernel start its own prologue (its ring fills while this kernel computes)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    TL_STAMP(2);

    // ---- 2. stage x as s8 residual terms, one power-of-two scale per (batch row, 64-block) ----------
    {
        const T* x = reinterpret_cast<const T*>(p.x);
        const int nchunk = p.K >> 3;
        for (int b = 0; b < batch; ++b) {
            for (int c = tid; c < nchunk; c += kThreads) {
                float f[8];
                XLoad<T>::load(x + (size_t)b * p.K + c * 8, f);
                float mx = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fabsf(f[i]));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
                uint32_t E = (__float_as_uint(mx) >> 23) & 0xFFu;
                E = E < 32u ? 32u : (E > 250u ? 250u : E);
                const float s = __uint_as_float((259u - E) << 23);  // 2^(5 - e): |x * s| < 64, so |rint| <= 64
                const int kb = c >> 3, pos = c & 7;
                if (pos == 0) sXs[b * nkb + kb] = __uint_as_float((E - 5u) << 23) * (1.f / 192.f);
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = f[i] * s;
                uint8_t* dst = sX + ((size_t)(kb * ncols + b * NT) * 64) + pos * 8;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    uint32_t ti[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float rr = y[i] + kMagic;  // round to nearest integer
                        ti[i] = __float_as_uint(rr);
                        y[i] = (y[i] - (rr - kMagic)) * 128.f;  // exact residual, rescaled
                    }
                    // byte order = nibble order of the packed weights: (k+1, k, k+3, k+2)
                    const uint32_t w0 = prmt(prmt(ti[1], ti[0], 0x0040u), prmt(ti[3], ti[2], 0x0040u), 0x5410u);
                    const uint32_t w1 = prmt(prmt(ti[5], ti[4], 0x0040u), prmt(ti[7], ti[6], 0x0040u), 0x5410u);
                    *reinterpret_cast<uint2*>(dst + j * 64) = make_uint2(w0, w1);
                }
            }
        }
    }
    __syncthreads();
    TL_STAMP(3);
    if (n == 0) return;

    // ---- 3. main loop over this warp's items -------------------------------------------------------
    Ctx fc;
    fc.p = &p; fc.sPart = sPart; fc.sCnt = sCnt;
    fc.cta_b = cta_b; fc.cta_e = cta_e; fc.t_first = t_first;

    // TMA SWIZZLE_128B: 16-byte chunk c of box row r lands at chunk c ^ (r & 7).  Unit u of this lane is
    // bytes [32u + 8t, +8) of row rho (row rho + 8 is 1 KiB further: same swizzle phase).
    uint32_t woff[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) woff[u] = rho * 128 + (((2 * u + (t >> 1)) ^ rho) << 4) + 8 * (t & 1);
    const uint32_t aoff = rho * 16;

    uint32_t xoff[NCT], soff[NCT];  // byte offsets of this lane's x fragments / block scales
#pragma unroll
    for (int ct = 0; ct < NCT; ++ct) {
        int col = ct * 8 + (int)g;
        col = col < ncols ? col : ncols - 1;  // lanes without a column re-read the last one (never stored)
        xoff[ct] = (uint32_t)col * 64 + t * 16;
        int b = NT == 2 ? ct * 4 + (int)t : ct * 2 + (int)(t >> 1);
        b = b < batch ? b : batch - 1;
        soff[ct] = (uint32_t)b * nkb * 4;
    }
    const uint32_t sX_a = (uint32_t)__cvta_generic_to_shared(sX);
    const uint32_t sXs_a = (uint32_t)__cvta_generic_to_shared(sXs);
    const uint32_t xstep = (uint32_t)ncols * 64;  // bytes of x fragments per 64-block
    uint32_t tab_lo;
    asm volatile("mov.b32 %0, 0xC0800100;" : "=r"(tab_lo));  // kept in a register: PRMT takes no immediates here
    int magic_i;
    asm volatile("mov.b32 %0, 0x4B400000;" : "=r"(magic_i));

    float acc[NCT][4];
#pragma unroll
    for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[ct][i] = 0.f;

    auto flush = [&](uint32_t tile, uint32_t cnt) {
        float v[NCT][2];
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) {
            if constexpr (NT == 2) {
                v[ct][0] = acc[ct][0] + acc[ct][1] * (1.f / 128.f);
                v[ct][1] = acc[ct][2] + acc[ct][3] * (1.f / 128.f);
            } else {
                const float w0 = (t & 1) ? (1.f / 16384.f) : 1.f, w1 = w0 * (1.f / 128.f);
                const float a = acc[ct][0] * w0 + acc[ct][1] * w1, b2 = acc[ct][2] * w0 + acc[ct][3] * w1;
                v[ct][0] = __shfl_xor_sync(0xffffffffu, a, 1) + a;
                v[ct][1] = __shfl_xor_sync(0xffffffffu, b2, 1) + b2;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[ct][i] = 0.f;
        }
        flush_tile<T, NT, NCT>(fc, v, tile, cnt);
    };

    uint32_t cur_tile = st_tile[0], left = n, phase = 0, cnt = 0;
    while (left) {
#pragma unroll
        for (int slot = 0; slot < kStages; ++slot) {
            if (left == 0) break;
            const uint32_t tile = st_tile[slot], kq = st_kq[slot];
            if (tile != cur_tile) {
                flush(cur_tile, cnt);
                cur_tile = tile;
                cnt = 0;
            }
            const uint32_t wbase = ringW + slot * kWBox, abase = ringA + slot * kABox;
            mbar_wait(bar0 + slot * 8, phase);
#ifdef FP4_I8_TIMELINE
            if (left == n) TL_STAMP(4);
#endif
            const uint4 amA = lds_u4(abase + aoff), amB = lds_u4(abase + aoff + 8 * 16);
            const float am0[4] = {__uint_as_float(amA.x), __uint_as_float(amA.y), __uint_as_float(amA.z), __uint_as_float(amA.w)};
            const float am1[4] = {__uint_as_float(amB.x), __uint_as_float(amB.y), __uint_as_float(amB.z), __uint_as_float(amB.w)};
            uint4 xs4[NCT];
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) xs4[ct] = lds_u4(sXs_a + soff[ct] + kq * 16);
            const uint32_t xbase = sX_a + kq * 4 * xstep;
#ifdef FP4_I8_NOCOMPUTE  // experiment: the weight stream alone (no decode, no MMA)
            acc[0][0] += __uint_as_float(lds_u2(wbase + woff[0]).x & 0x3F000000u) * am0[0] * __uint_as_float(xs4[0].x);
#else
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint2 wa = lds_u2(wbase + woff[u]);
                const uint2 wb = lds_u2(wbase + woff[u] + 8 * 128);
                uint32_t aA[2][2], nA[2][2], aB[2][2], nB[2][2];  // [word][lo/hi], rows rho / rho+8
                decode_word(wa.x, tab_lo, aA[0][0], aA[0][1], nA[0][0], nA[0][1]);
                decode_word(wa.y, tab_lo, aA[1][0], aA[1][1], nA[1][0], nA[1][1]);
                decode_word(wb.x, tab_lo, aB[0][0], aB[0][1], nB[0][0], nB[0][1]);
                decode_word(wb.y, tab_lo, aB[1][0], aB[1][1], nB[1][0], nB[1][1]);
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    const uint4 bx = lds_u4(xbase + u * xstep + xoff[ct]);
                    int dall[4] = {magic_i, magic_i, magic_i, magic_i}, dneg[4] = {magic_i, magic_i, magic_i, magic_i};
                    imma(dall, aA[0][0], aB[0][0], aA[0][1], aB[0][1], bx.x, bx.y);
                    imma(dneg, nA[0][0], nB[0][0], nA[0][1], nB[0][1], bx.x, bx.y);
                    imma(dall, aA[1][0], aB[1][0], aA[1][1], aB[1][1], bx.z, bx.w);
                    imma(dneg, nA[1][0], nB[1][0], nA[1][1], nB[1][1], bx.z, bx.w);
                    const float xsu = __uint_as_float(u == 0 ? xs4[ct].x : u == 1 ? xs4[ct].y : u == 2 ? xs4[ct].z : xs4[ct].w);
                    const float s0 = am0[u] * xsu, s1 = am1[u] * xsu;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // both accumulators started at the bit pattern of 1.5*2^23, so as floats they read
                        // 1.5*2^23 + sum exactly; (M + all) - 2 (M + neg) + M = all - 2 neg, every step exact
                        // (all on the FMA pipe: the ALU pipe is the one the decode saturates)
                        const float f = fmaf(__int_as_float(dneg[i]), -2.f, __int_as_float(dall[i])) + kMagic;
                        acc[ct][i] = fmaf(f, i < 2 ? s0 : s1, acc[ct][i]);
                    }
                }
            }
#endif
            // every lane has consumed the slot (the MMAs above are warp-synchronous): refill it
            __syncwarp();
            issue_box(slot);
            --left;
            ++cnt;
        }
        phase ^= 1;
    }
    TL_STAMP(5);
    flush(cur_tile, cnt);
    TL_STAMP(6);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) != cudaSuccess ||
            r != cudaDriverEntryPointSuccess)
            q = nullptr;
        return (EncodeTiledFn)q;
    }();
    return fn;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int nterms(int dtype) { return dtype == FP4_B200_F32 ? 4 : 2; }

static size_t smem_bytes(int batch, int K, int nt, uint32_t ntl) {
    const size_t nkb = (size_t)K / 64;
    return (size_t)kW * kStages * (kWBox + kABox) + nkb * batch * nt * 64 + (size_t)batch * nkb * 4 +
           (size_t)ntl * kW * 16 * batch * 4 + (size_t)ntl * 4 + 8 + (size_t)kW * kStages * 8;
}

template <typename T, int NT, int NCT>
static int launch(const void* x, const uint8_t* packed, const float* absmax, const void* bias, void* out,
                  void* workspace, size_t workspace_bytes, int batch, int N, int K, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return FP4_B200_ERR_UNSUPPORTED;
    auto kern = gemv_i8_kernel<T, NT, NCT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    static const int ctas_per_sm = env_int("FP4_B200_GEMV_CTAS_PER_SM", 2);
    static const int min_items = env_int("FP4_B200_GEMV_MIN_ITEMS", 4);
    static const int use_pdl = env_int("FP4_B200_GEMV_PDL", 1);

    const uint32_t ipt = (uint32_t)K / 256, tiles = ((uint32_t)N + 15) / 16;
    const uint64_t items64 = (uint64_t)tiles * ipt;
    if (items64 >= (1ull << 31)) return FP4_B200_ERR_UNSUPPORTED;
    const uint32_t items = (uint32_t)items64;
    // grid: as many CTAs per SM as shared memory allows (capped), never fewer than min_items items per CTA
    int occ = ctas_per_sm < 1 ? 1 : ctas_per_sm;
    uint32_t grid = 0, ntl = 0;
    size_t smem = 0;
    for (; occ >= 1; --occ) {
        grid = (uint32_t)kNumSMs * occ;
        const uint32_t cap = (items + min_items - 1) / min_items;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        ntl = (items / grid + 1 + ipt - 1) / ipt + 2;
        smem = smem_bytes(batch, K, NT, ntl);
        if (smem * occ <= 220 * 1024 && smem <= 200 * 1024) break;
    }
    if (occ < 1) return FP4_B200_ERR_UNSUPPORTED;
    const size_t need = kCounterBytes + (size_t)grid * 2 * 128 * 4;
    if (!workspace || workspace_bytes < need) return FP4_B200_ERR_WORKSPACE;

    CUtensorMap tmW, tmA;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K / 2, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)K / 2};
        const cuuint32_t box[2] = {128, 16};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(packed), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FP4_B200_ERR_UNSUPPORTED;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K / 64, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)(K / 64) * 4};
        const cuuint32_t box[2] = {4, 16};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(absmax), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FP4_B200_ERR_UNSUPPORTED;
    }
    Params p;
    p.x = x; p.bias = bias; p.out = out;
    p.gcnt = reinterpret_cast<unsigned*>(workspace);
    p.gpart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCounterBytes);
    p.batch = batch; p.N = N; p.K = K;
    p.ipt = ipt; p.items = items;
    p.q = items / grid; p.r = items % grid;
    p.ntl = ntl;
    p.by_ipt = FastDiv(ipt);
    p.by_q = FastDiv(p.q ? p.q : 1);
    p.by_q1 = FastDiv(p.q + 1);
    p.tl = nullptr;
#ifdef FP4_I8_TIMELINE
    {   // debug: launch i of the process writes its stamps at g_tl + i * kTlStride
        if (g_i8_tl) p.tl = g_i8_tl + (size_t)(g_i8_tl_launch++) * (kNumSMs * 4 * kW * 8);
    }
#endif

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, kern, tmW, tmA, p);
}

template <typename T, int NT>
static int launch_nct(const void* x, const uint8_t* packed, const float* absmax, const void* bias, void* out,
                      void* workspace, size_t workspace_bytes, int batch, int N, int K, cudaStream_t st) {
    const int nct = (batch * NT + 7) / 8;
#define FP4_GO(NCT) launch<T, NT, NCT>(x, packed, absmax, bias, out, workspace, workspace_bytes, batch, N, K, st)
    if (nct <= 1) return FP4_GO(1);
    if (nct <= 2) return FP4_GO(2);
    return FP4_GO(4);
#undef FP4_GO
}

}  // namespace

#ifdef FP4_I8_TIMELINE
extern "C" void fp4_b200_debug_timeline(long long* buf) { g_i8_tl = buf; g_i8_tl_launch = 0; }
#endif

bool gemv_i8_supported(int batch, int N, int K, int blocksize, int dtype, bool nested, const void* packed,
                       const void* absmax) {
    static const int disabled = env_int("FP4_B200_GEMV_NO_I8", 0);
    if (disabled || nested || blocksize != 64) return false;
    if (batch < 1 || batch > 8 || N <= 0 || K <= 0) return false;
    if (K % 256 != 0) return false;
    if ((size_t)((N + 15) / 16) * 4 > kCounterBytes) return false;
    if (reinterpret_cast<uintptr_t>(packed) % 16 || reinterpret_cast<uintptr_t>(absmax) % 16) return false;
    const int nt = nterms(dtype);
    if ((batch * nt + 7) / 8 > 4) return false;
    const uint32_t ipt = (uint32_t)K / 256;
    const uint64_t items = (uint64_t)((N + 15) / 16) * ipt;
    if (items >= (1ull << 31)) return false;
    const uint32_t ntl = (uint32_t)((items / kNumSMs + 1 + ipt - 1) / ipt + 2);
    return smem_bytes(batch, K, nt, ntl) <= 200 * 1024 && encode_fn() != nullptr;
}

int gemv_i8_dispatch(const void* x, const uint8_t* packed, const float* absmax, const void* bias, void* out,
                     void* workspace, size_t workspace_bytes, int batch, int N, int K, int dtype,
                     cudaStream_t st) {
    switch (dtype) {
        case FP4_B200_F16:
            return launch_nct<__half, 2>(x, packed, absmax, bias, out, workspace, workspace_bytes, batch, N, K, st);
        case FP4_B200_BF16:
            return launch_nct<__nv_bfloat16, 2>(x, packed, absmax, bias, out, workspace, workspace_bytes, batch, N, K, st);
        case FP4_B200_F32:
            return launch_nct<float, 4>(x, packed, absmax, bias, out, workspace, workspace_bytes, batch, N, K, st);
        default:
            return FP4_B200_ERR_DTYPE;
    }
}

}  // namespace fp4b200
