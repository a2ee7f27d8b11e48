// Fused dequant + GEMV for decode (batch 1..8) on sm_100a — the fast path.
//
// Replaces gemv_4bit_inference_kernel{,_float} (reference csrc/gemv_fp4_optimized.cu:60-259).
// Bound: HBM.  Algorithmic bytes per weight: 0.5 (packed) + 4/blocksize (absmax) = 0.5625 at
// blocksize 64.  At the measured 6.56 TB/s one SM must retire ~41 weights per clock, i.e. ~4 warp
// instructions per 32 weights in total; a per-nibble shared-memory lookup + FMA (the reference's
// scheme: 32 LDS + 32 HMUL2 + 32 HFMA2 per 16 bytes) cannot fit.  This kernel therefore
//   * decodes FOUR nibbles per instruction with PRMT used as an 8-entry byte table: the
//     bitsandbytes FP4 magnitudes times 12 are {0, 1/16, 8, 12, 4, 6, 2, 3} - exactly representable
//     in e5m2, i.e. one byte each, the high byte of their fp16 encoding;
//   * merges the sign bits in with one more PRMT (sign-replicate mode) + one LOP3 per four nibbles;
//   * widens e5m2 -> fp16 with F2FP (cvt.rn.f16x2.e5m2x2: the byte becomes the high byte, exact);
//   * feeds the fp16 pairs to mma.sync.m16n8k16 (fp32 accumulate) as the A operand: 16 weight rows x
//     16 k per instruction, B = x (8 columns = up to 8 batch rows, so batch 2..8 costs no extra
//     decode work).  The contraction is unchanged; the multiply-adds just leave the FMA pipe.
// The absmax is factored out of the inner sum, y[r] = sum_b absmax[r,b] * sum_{k in b} c12[q]*x[k],
// so products are exact (fp16 x fp16 in fp32) and each 64-element block costs 4 FFMA per lane.
// x is staged ONCE per CTA in shared memory as fp16, pre-scaled by a power of two per batch row so
// that bf16/fp32 inputs cannot overflow fp16 (fp32 inputs are split hi + lo into two columns, ~22
// bits), and stored in mma B-fragment order so a lane fetches its operands with two LDS.128 per
// block.  Weights stream with 64-bit ld.global.nc.L1::no_allocate loads, kU blocks in flight per
// lane and row (a warp instruction reads whole 32-byte sectors of 8 rows; kU consecutive loads cover
// whole 128-byte lines); the first loads are issued before the x staging to overlap the first HBM
// round trip.
//
// Schedule ("stream-K"): the work is the flat sequence of units (row tile of 16 rows, 64-wide k
// block), 1 KiB of packed weights each.  The grid is persistent - (CTAs per SM) x 148 CTAs of 8
// warps - and every WARP owns an equal contiguous range of units (+-1), so there is no wave
// quantisation whatever N and K are, and x is staged once per CTA instead of once per row tile.
// A warp that covers a row tile alone writes the output directly; tiles shared between warps are
// combined deterministically: each contributor stores its fp32 partial in a workspace slot, bumps a
// per-tile counter, and the last arriver sums the slots in warp order, applies scale and bias,
// converts and stores (and resets the counter for the next launch).
//
// Requirements (checked by gemv_imma_supported): codebook == bitsandbytes FP4 table, K % 64 == 0,
// blocksize % 64 == 0, N % 16 == 0, shared memory for x <= 200 KB, a workspace.  Everything else
// takes gemv_generic.cu.
#include "gemv_common.cuh"

#include <cstdlib>
#include <type_traits>

namespace fp4b200 {

namespace {

using namespace gemv;

#ifndef FP4_GEMV_MIN_CTAS
#define FP4_GEMV_MIN_CTAS 2
#endif
constexpr int kThreads = kWarps * 32;

// T: activation dtype.  PIECES = 1 (fp16/bf16 x: one fp16 column per batch row) or 2 (fp32 x: hi + lo).
// NCOLT = number of 8-column MMA tiles (1, or 2 when batch*PIECES > 8).
template <typename T, int NCOLT, bool NESTED>
__global__ void __launch_bounds__(kThreads, FP4_GEMV_MIN_CTAS)
gemv_mma_kernel(const T* __restrict__ x, const uint8_t* __restrict__ packed,
                const float* __restrict__ absmax, const NestedDev nd, const T* __restrict__ bias,
                T* __restrict__ out, const Workspace ws, const int batch, const int N, const int K,
                const int am_shift /* log2(blocksize / 64) */, const Partition part,
                const FastDiv by_nkb) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    constexpr int NC = 8 * NCOLT;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ float sPart[kWarps * 2 * 16 * NC];  // CTA-local partial sums of shared row tiles
    __shared__ unsigned sCnt[kWarps];              // units accounted for, by first contributing warp
    __shared__ float sMax[kWarps * 8];
    __shared__ float sScale[8];
    const uint32_t nkb = (uint32_t)K >> 6;
    const int ncols = batch * PIECES;
    uint4* sB = reinterpret_cast<uint4*>(smem_raw);  // [nkb][2][ncols][4] x 16 B

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t g = lane >> 2, t = lane & 3;

    const uint32_t W = gridDim.x * kWarps;
    const uint32_t wid = blockIdx.x * kWarps + warp;
    const uint32_t L0 = part.begin(wid);
    const uint32_t n = part.begin(wid + 1) - L0;
    if (tid < kWarps) sCnt[tid] = 0;
#ifdef FP4_GEMV_TIMELINE
    long long tl[5];
    tl[0] = clock64();
#endif

    // ---- weight-stream cursor ---------------------------------------------------------------------
    // unit L -> (tile, kb).  Row g of the tile at block kb starts at packed + ((tile*16+g)*nkb + kb)*32;
    // this lane reads 8 bytes at + 8t, and the same for row g+8 (+ row8 bytes).
    uint32_t ld_left = n;                       // units not yet requested
    uint32_t tile0, ld_kb;
    by_nkb.divmod(n ? L0 : 0, tile0, ld_kb);    // first unit -> (row tile, k block)
    const uint32_t ld_kb_first = ld_kb;
    const uint8_t* ld_ptr = packed + ((size_t)((tile0 * 16 + g) * nkb + ld_kb)) * 32 + 8 * t;
    const uint32_t row8 = 8 * nkb * 32;         // bytes from row g to row g+8
    const uint32_t tile_skip = 15 * nkb * 32;   // extra bytes when stepping into the next row tile

    uint2 qa0[4], qa1[4], qb0[4], qb1[4];       // two groups of 4 units in flight, rows g / g+8
    auto load_unit = [&](uint2& d0, uint2& d1) {  // generic: one unit, with tile wrap
        if (ld_left) {
            d0 = ldg_stream_u2(ld_ptr);
            d1 = ldg_stream_u2(ld_ptr + row8);
            --ld_left;
            ld_ptr += 32;
            if (++ld_kb == nkb) {
                ld_kb = 0;
                ld_ptr += tile_skip;
            }
        }
    };
    // absmax: the four lanes of a quad would all fetch the same value; instead lane t fetches the value
    // of unit (group base + t) - one load per lane per 4 units, 4x fewer L1 wavefronts - and the quad
    // exchanges them with shuffles.
    uint32_t am_kb = ld_kb + t, am_u64 = (tile0 * 16 + g) * nkb + ld_kb + t;
    while (am_kb >= nkb) {
        am_kb -= nkb;
        am_u64 += 15 * nkb;
    }
    const uint32_t r8u = 8 * nkb;
    uint32_t am_unit = t;  // index (within the warp's range) of the unit this lane fetches next
    float amA0 = 0.f, amA1 = 0.f, amB0 = 0.f, amB1 = 0.f;  // next group / the one after, rows g / g+8
    auto issue_absmax = [&](float& d0, float& d1) {
        if (am_unit < n) {
            d0 = load_absmax<NESTED>(absmax, nd, (int64_t)(am_u64 >> am_shift));
            d1 = load_absmax<NESTED>(absmax, nd, (int64_t)((am_u64 + r8u) >> am_shift));
        }
        am_unit += 4;
        am_kb += 4;
        am_u64 += 4;
        while (am_kb >= nkb) {  // crossed into the next row tile(s)
            am_kb -= nkb;
            am_u64 += 15 * nkb;
        }
    };
    // ---- 1. put the first 8 units of the weight stream in flight --------------------------------
#pragma unroll
    for (int u = 0; u < 4; ++u) load_unit(qa0[u], qa1[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) load_unit(qb0[u], qb1[u]);
    issue_absmax(amA0, amA1);
    issue_absmax(amB0, amB1);

#ifdef FP4_GEMV_TIMELINE
    tl[1] = clock64();
#endif
    // Programmatic dependent launch: everything above touched only the weights (never written), so it
    // overlaps the tail of the previous kernel in the stream; x, the workspace and `out` may be that
    // kernel's products, so wait for it here.  Then let the NEXT kernel start its own prologue.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    // ---- 2. stage x as scaled fp16 in B-fragment order (once per CTA) --------------------------
    const int nchunk = K >> 3;  // 8-element chunks per batch row
    for (int b = 0; b < batch; ++b) {  // pass 1: max |x| per batch row
        float mx = 0.f;
        for (int c = tid; c < nchunk; c += kThreads) {
            float f[8];
            XLoad<T>::load(x + (size_t)b * K + c * 8, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fabsf(f[i]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) sMax[warp * 8 + b] = mx;
    }
    __syncthreads();
    for (int b = 0; b < batch; ++b) {  // pass 2: scale into [2^13, 2^14), convert, store
        float m = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) m = fmaxf(m, sMax[w * 8 + b]);
        int E = (int)((__float_as_uint(m) >> 23) & 0xFFu);
        E = E < 14 ? 14 : (E > 254 ? 254 : E);
        const float scale = __uint_as_float((uint32_t)(267 - E) << 23);  // 2^(13 - e)
        if (tid == 0) sScale[b] = __uint_as_float((uint32_t)(E - 13) << 23) * (1.f / 12.f);
        for (int c = tid; c < nchunk; c += kThreads) {
            float f[8];
            XLoad<T>::load(x + (size_t)b * K + c * 8, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] *= scale;
            const int kb = c >> 3, qq = c & 7;  // chunk qq of the block: lane t = qq/2, half = qq%2
            uint4* dst = sB + ((size_t)(kb * 2 + (qq & 1)) * ncols + b * PIECES) * 4 + (qq >> 1);
            dst[0] = pack_swapped(f);
            if constexpr (PIECES == 2) {
                float lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    lo[i] = (f[i] - __half2float(__float2half_rn(f[i]))) * 2048.f;
                dst[4] = pack_swapped(lo);
            }
        }
    }
    __syncthreads();
#ifdef FP4_GEMV_TIMELINE
    tl[2] = clock64();
#endif
    if (n == 0) return;

    // ---- 3. main loop over this warp's units -----------------------------------------------------
    FlushCtx<T, NC> fc;
    fc.sScale = sScale; fc.sPart = sPart; fc.sCnt = sCnt; fc.bias = bias; fc.out = out; fc.ws = ws;
    fc.part = part; fc.nkb = nkb; fc.wid = wid; fc.first_tile = tile0; fc.by_nkb = by_nkb;
    fc.cta_w0 = blockIdx.x * kWarps;
    fc.cta_L0 = part.begin(fc.cta_w0); fc.cta_L1 = part.begin(fc.cta_w0 + kWarps);
    fc.batch = batch; fc.N = N;

    AccV<NCOLT> acc;
#pragma unroll
    for (int c = 0; c < NCOLT; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc.v[c][i] = 0.f;

    uint32_t tile = tile0, kb = ld_kb_first;
    uint32_t seg_start_kb = kb;  // first block of the current tile segment
    // shared-space byte address of this lane's B fragments for block kb and column tile 0:
    //   sB[(kb*2 + h)*ncols*4 + col*4 + t];  lanes whose column does not exist read column 0 instead
    //   (their accumulator columns are never stored), so the loads need no predicate
    const uint32_t bstep = 2 * ncols * 4 * 16;   // bytes per k block
    const uint32_t bhalf = ncols * 4 * 16;       // bytes between the two halves of a block
    uint32_t bcol[NCOLT];
#pragma unroll
    for (int ct = 0; ct < NCOLT; ++ct) {
        const int col = ct * 8 + (int)g;
        bcol[ct] = (col < ncols ? col : 0) * 64;
    }
    const uint32_t sB_base = (uint32_t)__cvta_generic_to_shared(sB) + t * 16;
    uint32_t bsaddr = sB_base + kb * bstep;
    uint32_t tab_lo;
    asm volatile("mov.b32 %0, 0x4A482C00;" : "=r"(tab_lo));  // kTabLo, pinned in a register
    const uint32_t quad = lane & ~3u;

    // FULL: the group holds 4 units and its slots are refilled (steady state, no per-unit checks);
    // otherwise `cnt` (< 4 possible) units are consumed and nothing is refilled (tail).
    auto consume_group = [&](uint2 (&q0)[4], uint2 (&q1)[4], auto full_tag, const uint32_t cnt) {
        constexpr bool FULL = decltype(full_tag)::value;
        const float gA0 = amA0, gA1 = amA1;  // this group's absmax, one unit per lane of the quad
        amA0 = amB0;
        amA1 = amB1;
        issue_absmax(amB0, amB1);            // two groups ahead
        // fast refill: the group 8 units ahead lies inside one row tile -> immediate offsets
        const bool fast = FULL && ld_left >= 4 && ld_kb + 4 <= nkb;
        const uint8_t* rp0 = ld_ptr;
        const uint8_t* rp1 = ld_ptr + row8;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (FULL || (uint32_t)u < cnt) {
                uint32_t ha[2][4], hb[2][4];  // [word][half2] for rows g / g+8
                decode_word(q0[u].x, tab_lo, ha[0]);
                decode_word(q0[u].y, tab_lo, ha[1]);
                decode_word(q1[u].x, tab_lo, hb[0]);
                decode_word(q1[u].y, tab_lo, hb[1]);
                if (fast) {  // the slot's registers are dead now: refill them
                    q0[u] = ldg_stream_u2(rp0 + 32 * u);
                    q1[u] = ldg_stream_u2(rp1 + 32 * u);
                }
                const float am0 = __shfl_sync(0xffffffffu, gA0, quad | u);
                const float am1 = __shfl_sync(0xffffffffu, gA1, quad | u);
#pragma unroll
                for (int ct = 0; ct < NCOLT; ++ct) {
                    const uint4 bA = lds_u4(bsaddr + bcol[ct]);          // k16 groups 0,1
                    const uint4 bB = lds_u4(bsaddr + bhalf + bcol[ct]);  // k16 groups 2,3
                    // two independent accumulation chains (halves the dependent HMMA latency)
                    float d[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
                    mma16816(d, ha[0][0], hb[0][0], ha[0][1], hb[0][1], bA.x, bA.y);
                    mma16816(e, ha[1][0], hb[1][0], ha[1][1], hb[1][1], bB.x, bB.y);
                    mma16816(d, ha[0][2], hb[0][2], ha[0][3], hb[0][3], bA.z, bA.w);
                    mma16816(e, ha[1][2], hb[1][2], ha[1][3], hb[1][3], bB.z, bB.w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] += e[i];
                    acc.v[ct][0] = fmaf(am0, d[0], acc.v[ct][0]);
                    acc.v[ct][1] = fmaf(am0, d[1], acc.v[ct][1]);
                    acc.v[ct][2] = fmaf(am1, d[2], acc.v[ct][2]);
                    acc.v[ct][3] = fmaf(am1, d[3], acc.v[ct][3]);
                }
                bsaddr += bstep;
                if (++kb == nkb) {  // row tile complete (for this warp's part of it)
                    flush_tile<T, NCOLT>(fc, acc, tile, seg_start_kb, kb);
#pragma unroll
                    for (int c = 0; c < NCOLT; ++c)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc.v[c][i] = 0.f;
                    kb = 0;
                    seg_start_kb = 0;
                    ++tile;
                    bsaddr = sB_base;
                }
            }
        }
        if (FULL) {
            if (fast) {
                ld_left -= 4;
                ld_ptr += 128;
                if ((ld_kb += 4) == nkb) {
                    ld_kb = 0;
                    ld_ptr += tile_skip;
                }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) load_unit(q0[u], q1[u]);
            }
        }
    };

    uint32_t s = 0;
    for (; s + 8 <= n; s += 8) {
        consume_group(qa0, qa1, std::true_type{}, 4);
        consume_group(qb0, qb1, std::true_type{}, 4);
    }
    const uint32_t rem = n - s;  // < 8 units left, already in the two register groups
    if (rem) consume_group(qa0, qa1, std::false_type{}, rem < 4 ? rem : 4);
    if (rem > 4) consume_group(qb0, qb1, std::false_type{}, rem - 4);
#ifdef FP4_GEMV_TIMELINE
    tl[3] = clock64();
#endif
    if (kb != seg_start_kb) flush_tile<T, NCOLT>(fc, acc, tile, seg_start_kb, kb);
#ifdef FP4_GEMV_TIMELINE
    tl[4] = clock64();
    if (lane == 0) {
        long long* dbg = reinterpret_cast<long long*>(ws.partials + (size_t)W * 2 * 16 * NC) + (size_t)wid * 8;
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        for (int i = 0; i < 5; ++i) dbg[i] = tl[i];
        dbg[5] = smid;
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[6] = gt;
    }
#endif
}

struct Plan {
    int grid;
    size_t smem;
};

static size_t smem_bytes(int batch, int K, int pieces) {
    const size_t nkb = K / 64;
    return nkb * 2 * (size_t)(batch * pieces) * 4 * 16;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <typename T, int NCOLT, bool NESTED>
static int launch(const void* x, const uint8_t* packed, const float* absmax, const NestedDev& nd,
                  const void* bias, void* out, void* workspace, size_t workspace_bytes, int batch,
                  int N, int K, int bs_log2, cudaStream_t st) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    constexpr int NC = 8 * NCOLT;
    auto kern = gemv_mma_kernel<T, NCOLT, NESTED>;
    const size_t smem = smem_bytes(batch, K, PIECES);
    static bool configured = false;  // per template instantiation
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    static const int ctas_per_sm = env_int("FP4_B200_GEMV_CTAS_PER_SM", 0);
    static const int min_units = env_int("FP4_B200_GEMV_MIN_UNITS", 8);
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return FP4_B200_ERR_UNSUPPORTED;
    if (occ > 4) occ = 4;  // the workspace is sized for <= 4 CTAs per SM
    if (ctas_per_sm > 0 && ctas_per_sm < occ) occ = ctas_per_sm;
    const int64_t units = (int64_t)(N / 16) * (K / 64);
    int64_t grid = (int64_t)kNumSMs * occ;
    const int64_t max_by_work = (units + (int64_t)kWarps * min_units - 1) / ((int64_t)kWarps * min_units);
    if (grid > max_by_work) grid = max_by_work;
    if (grid < 1) grid = 1;
    // workspace: counters [ntiles] + partials [W][2][16][NC]
    const size_t need = kCounterBytes + (size_t)grid * kWarps * 2 * 16 * NC * 4;
    if (!workspace || workspace_bytes < need) return FP4_B200_ERR_WORKSPACE;
    Workspace ws;
    ws.counters = reinterpret_cast<unsigned*>(workspace);
    ws.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) +
                                           kCounterBytes);
    const uint32_t Wn = (uint32_t)grid * kWarps, Bn = (uint32_t)units;
    Partition part;
    part.q = Bn / Wn;
    part.r = Bn % Wn;
    part.by_q = FastDiv(part.q ? part.q : 1);
    part.by_q1 = FastDiv(part.q + 1);
    static const int use_pdl = env_int("FP4_B200_GEMV_PDL", 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kern, (const T*)x, packed, absmax, nd, (const T*)bias, (T*)out, ws,
                           batch, N, K, bs_log2 - 6, part, FastDiv((uint32_t)(K / 64)));
    return (int)e;
}

template <typename T>
static int launch_t(const void* x, const uint8_t* packed, const float* absmax, bool nested,
                    const NestedDev& nd, const void* bias, void* out, void* workspace,
                    size_t workspace_bytes, int batch, int N, int K, int bs_log2, cudaStream_t st) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    const bool two = batch * PIECES > 8;
#define FP4_GO(NCT, NST)                                                                         \
    launch<T, NCT, NST>(x, packed, absmax, nd, bias, out, workspace, workspace_bytes, batch, N, K, \
                        bs_log2, st)
    if (nested) return two ? FP4_GO(2, true) : FP4_GO(1, true);
    return two ? FP4_GO(2, false) : FP4_GO(1, false);
#undef FP4_GO
}

}  // namespace

bool gemv_imma_supported(int batch, int N, int K, int blocksize, int dtype) {
    if (batch < 1 || batch > 8 || N <= 0 || K <= 0) return false;
    if (K % 64 != 0 || blocksize % 64 != 0 || N % 16 != 0) return false;
    if ((int64_t)N * K / 64 >= (int64_t)1 << 31) return false;  // 32-bit unit cursors
    if ((size_t)(N / 16) * 4 > kCounterBytes) return false;
    const int pieces = dtype == FP4_B200_F32 ? 2 : 1;
    return smem_bytes(batch, K, pieces) <= 200 * 1024;
}

size_t gemv_imma_workspace_bytes(int /*N*/) {
    // counters + the largest partial area any launch can need: 3 CTAs/SM x 8 warps x 2 slots x 16 x 16 floats
    return kCounterBytes + (size_t)kNumSMs * 4 * kWarps * 2 * 16 * 16 * 4;
}

int gemv_imma_dispatch(const void* x, const uint8_t* packed, const float* absmax,
                       const fp4_b200_nested_t* nested, const NestedDev& nd, const void* bias,
                       void* out, void* workspace, size_t workspace_bytes, int batch, int N, int K,
                       int bs_log2, int dtype, cudaStream_t st) {
    const bool nst = nested != nullptr;
    switch (dtype) {
        case FP4_B200_F16:
            return launch_t<__half>(x, packed, absmax, nst, nd, bias, out, workspace,
                                    workspace_bytes, batch, N, K, bs_log2, st);
        case FP4_B200_BF16:
            return launch_t<__nv_bfloat16>(x, packed, absmax, nst, nd, bias, out, workspace,
                                           workspace_bytes, batch, N, K, bs_log2, st);
        case FP4_B200_F32:
            return launch_t<float>(x, packed, absmax, nst, nd, bias, out, workspace,
                                   workspace_bytes, batch, N, K, bs_log2, st);
        default:
            return FP4_B200_ERR_DTYPE;
    }
}

}  // namespace fp4b200
