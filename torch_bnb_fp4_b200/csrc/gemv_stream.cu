// Fused dequant + GEMV for decode (batch 1..8) on sm_100a — the default kernel for bitsandbytes-FP4
// weights with blocksize 64:
//
//   out[b, r] = T( sum_k x[b,k] * code[W[r,k]] * absmax[r, k/64] + bias[r] )
//
// replaces the reference's gemv_4bit_inference kernels (csrc/gemv_fp4_optimized.cu:60-259).
//
// Memory side (what bounds the kernel).  One persistent CTA per SM owns a contiguous range of 16-row
// tiles; its 16 warps split the range's units (16 rows x 512 k = 4 KiB of packed weights + 512 B of
// absmax) contiguously.  Loaded latency of HBM on B200 is ~2 us, so ~100+ KB per SM must be in flight:
// far more than registers hold.  Every warp therefore owns a private ring of `ring` unit slots in shared
// memory, filled with 16-byte cp.async copies laid out directly in MMA-fragment order (lane l's bytes
// at lane l's slot: each lane reads back only what it copied, so there are no barriers, no bank
// conflicts and no cross-lane visibility to wait for — cp.async.wait_group is the only
// synchronisation; the ALIGNED variant below adds a __syncwarp).  The first ring slot (or half of it, see
// p.pre) is issued BEFORE griddepcontrol.wait: under programmatic dependent launch the next layer's first
// units are already on their way while the current layer finishes, and a layer of <= 1 unit per warp
// (4096x4096) is entirely in flight before its input exists.  The rest of the ring goes out between the loads
// of x and their first use, so its LSU time overlaps the L2 round trip of x.
//
// Arithmetic side (must stay under ~20 issue slots per 8 weights to keep up with HBM):
//   * weights: 192*|code| = {0,1,128,192,64,96,32,48} fits a byte: one PRMT against that 8-entry table
//     turns four nibbles into four u8 magnitudes (ALL); a PRMT in sign-replicate mode gives 0x00/0xFF
//     sign masks, ALL & mask = the magnitudes of the negative weights (NEG).
//   * x: per (batch row, 64-block) a power-of-two scale, then x*s = t1 + t2/128 (+ t3/128^2 + t4/128^3
//     for fp32) with s8 integers t_j (exact residual expansion, 2^-14 / 2^-28 of the block maximum).
//   * sum_k w*x = absmax * 2^-e/192 * (IMMA(ALL,t) - 2 IMMA(NEG,t)): exact integer dot products on the
//     tensor cores (mma.sync m16n8k32 u8 x s8), scaled once per (row, 64-block) in fp32.
//   * one lane holds 16 contiguous bytes of a row (LDG.128), so the four lanes of an MMA k-group span
//     TWO absmax blocks.  The B operand separates them: each (batch row, term) owns two MMA columns,
//     one holding x for the lanes of block A and zeros elsewhere, the other the same for block B.
//   The contraction is unchanged (one weight row x one activation vector); the tensor core is only the
//   multiply-add engine that takes the FMAs off the issue-limited pipes.
//   * ALIGNED variant (more than eight (batch row, term) columns, i.e. batch >= 5 for 16-bit inputs): half of
//     the columns above are structural zeros, and IMMA issue is what batch 3..8 is bound by.  Here the 16-byte
//     chunks of a row are stored transposed in the ring (lane (g,t) writes slot position t*8+g) and a lane reads
//     word t of chunk i with a 32-bit load, so the four lanes of a k-group cover ONE absmax block and every MMA
//     column is a real (batch row, term): half the IMMAs.  x is staged in the matching order with a 16-byte
//     row pad (conflict-free 128-bit reads).
//
// Row tiles are never split between CTAs (no global atomics, fences or workspace): tiles a CTA's warps
// share are summed through shared memory in warp order after one __syncthreads (deterministic).
//
// Requirements (gemv_stream_supported): bitsandbytes FP4 codebook, blocksize 64, fp32 absmax,
// K % 256 == 0, N % 16 == 0, enough row tiles to occupy the GPU, x terms fit in shared memory.
#include <cstdlib>

#include "gemv_common.cuh"

namespace fp4b200 {
#ifdef FP4_STREAM_TIMELINE
long long* g_stream_tl = nullptr;
int g_stream_tl_launch = 0;
#endif
namespace {

using gemv::FastDiv;
using gemv::XLoad;
using gemv::lds_u2;
using gemv::lds_u4;

#ifndef FP4_STREAM_MINB
#define FP4_STREAM_MINB 1
#endif
#ifndef FP4_STREAM_WARPS
#define FP4_STREAM_WARPS 16
#endif
constexpr int kW = FP4_STREAM_WARPS;  // warps per CTA
constexpr int kThreads = kW * 32;
constexpr float kMagic = 12582912.f;     // 1.5 * 2^23: int32 accumulators that start at its bit pattern read as floats
constexpr uint32_t kTabHi = 0x30206040u;  // 192*|code[4..7]| = 64, 96, 32, 48  (low half 0xC0800100 lives in a register)
constexpr uint32_t kZeroBytes = 512;      // zero region read by the lanes of masked MMA columns
#ifndef FP4_STREAM_SMEM_KB
#define FP4_STREAM_SMEM_KB 226
#endif
constexpr uint32_t kMaxSmem = FP4_STREAM_SMEM_KB * 1024; // per CTA: half an SM, so the next launch (PDL) is resident while this one runs
constexpr uint32_t kSlot = 4096 + 512;    // one unit: 4 steps x 2 row halves x 32 lanes x 16 B, then 32 lanes x 16 B of absmax
constexpr uint32_t kMaxRing = 4;

constexpr int kMaxGroup = 4;  // weight matrices sharing one x in a grouped launch (q/k/v, gate/up)

struct Params {
    const void* x;
    // per matrix, rebased so that they index by GLOBAL row (rows of the matrices concatenated):
    // vpacked[m] + grow * K/2, vabsmax[m] + grow * K/64, vout[m][b * Nm[m] + grow], vbias[m][grow]
    const uint8_t* vpacked[kMaxGroup];
    const float* vabsmax[kMaxGroup];
    const void* vbias[kMaxGroup];  // may be NULL
    void* vout[kMaxGroup];
    int Nm[kMaxGroup];             // out_features of matrix m (row pitch of its output)
    // nested (double-quantised) absmax of matrix m, decoded in the kernel: uint8 codes in place of the
    // fp32 absmax (vqabs rebased to global rows like vabsmax), the 256-entry map, fp32 absmax2 per 2^bs2_log2
    // blocks (indexed by the matrix's own block number) and the offset.  nested_mask bit m: matrix m is nested.
    const uint8_t* vqabs[kMaxGroup];
    const float* vcode2[kMaxGroup];
    const float* vabs2[kMaxGroup];
    float voffset[kMaxGroup];
    int vbs2_log2[kMaxGroup];
    uint32_t nested_mask;
    // fused neighbours of the Linear:
    //  gated != 0: matrices 0 / 1 are the gate / up projection of an MLP (same N, same K).  A row tile is then 8 rows
    //    of EACH (MMA rows g = gate row, g + 8 = up row of the same index), so one lane ends up holding both values
    //    of an output element and writes act(gate) * up - one output [batch, N], no intermediate tensors
    //    (gated: 1 = SiLU, 2 = GELU tanh approximation);
    //  vres[m] != NULL: a residual [batch, N_m] of dtype T added to the output (o / down projections).
    int gated;
    const void* vres[kMaxGroup];
    uint32_t tstart[kMaxGroup];    // first global tile of matrix m (tstart[0] = 0; unused entries = UINT32_MAX)
    int batch, K;
    uint32_t upt;     // units per row tile = ceil(K / 512)
    uint32_t tq, tr;  // CTA c owns tiles [c*tq + min(c,tr), +tq + (c<tr))
    uint32_t ring;    // unit slots per warp (1..kMaxRing)
    uint32_t pre;     // steps (1..4) of the first slot issued before griddepcontrol.wait
    FastDiv by_upt;
    fp4_b200_tp_t tp;  // tensor-parallel exchange through peer memory (in_world / out_world <= 1: off)
    long long* tl;    // debug timeline (FP4_STREAM_TIMELINE builds): [cta][warp][8] globaltimer ns
};

#ifdef FP4_STREAM_TIMELINE
#define TL_STAMP(i)                                                              \
    do {                                                                         \
        if (p.tl && lane == 0) {                                                 \
            long long gt_;                                                       \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));              \
            p.tl[((size_t)blockIdx.x * kW + warp) * 8 + (i)] = gt_;              \
        }                                                                        \
    } while (0)
#else
#define TL_STAMP(i) do {} while (0)
#endif

__device__ __forceinline__ void imma_first(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b0, uint32_t b1, int c) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(c));
}
__device__ __forceinline__ void imma_acc(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
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
    const uint32_t w4 = w * 16u;  // integer multiply: issues on the FMA pipe, not the busier ALU pipe
    all_lo = prmt(tab_lo, kTabHi, wm);
    all_hi = prmt(tab_lo, kTabHi, __umulhi(wm, 65536u));
    // sign-replicate mode (selector msb): byte = 0xFF if the selected source byte has its msb set
    neg_lo = all_lo & prmt(w, w4, 0x9D8Cu);  // signs of nibbles 0,1,2,3
    neg_hi = all_hi & prmt(w, w4, 0xBFAEu);  // signs of nibbles 4,5,6,7
}

__device__ __forceinline__ void cp_async_cg16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_ca16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_ca4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(uint32_t pending) {  // uniform across the CTA
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}

// activation of the gate in a gated MLP epilogue: 1 = SiLU (x * sigmoid(x)), 2 = GELU, tanh approximation
__device__ __forceinline__ float gate_act(float v, int kind) {
    if (kind == 2) {
        const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
        return 0.5f * v * (1.f + tanhf(u));
    }
    return v / (1.f + __expf(-v));
}

// ---- tensor-parallel exchange through peer (symmetric) memory -------------------------------------------
// A row-parallel layer pushes its partial output as self-validating 64-bit words {two 16-bit values, tag32 = epoch}
// into every rank's exchange buffer over NVLink (no fences, no flags: an aligned 8-byte store is single-copy atomic);
// the consumer sums the ranks' partials (fp32, fixed rank order) out of its LOCAL buffer while it stages x, re-reading
// words whose tag is not the current epoch yet.  A word carries rows r and r + 8 of one 16-row tile - the two values a
// lane of the m16n8 accumulator owns - so the producer issues ONE store per lane and rank, and the exchange moves
// 4 bytes per element, as many as the {value16, tag16} words of round 1 whose tags came round again after 65536 epochs
// (a word left over from a larger batch could then validate by accident).  32-bit tags do not wrap in the life of a
// process.  Word index of (batch row b, output row r): (b * N + (r & ~15)) / 2 + (r & 7); low half: rows with bit 3 clear.
template <typename T>
__device__ __forceinline__ uint32_t tp_pack2(float lo, float hi) {
    if constexpr (DT<T>::code == FP4_B200_BF16) {
        return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
    } else {
        return (uint32_t)__half_as_ushort(__float2half_rn(lo)) | ((uint32_t)__half_as_ushort(__float2half_rn(hi)) << 16);
    }
}
template <typename T>
__device__ __forceinline__ float tp_lo(uint32_t w) {
    if constexpr (DT<T>::code == FP4_B200_BF16) return __uint_as_float(w << 16);
    else return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu)));
}
template <typename T>
__device__ __forceinline__ float tp_hi(uint32_t w) {
    if constexpr (DT<T>::code == FP4_B200_BF16) return __uint_as_float(w & 0xFFFF0000u);
    else return __half2float(__ushort_as_half((unsigned short)(w >> 16)));
}
__device__ __forceinline__ void st_peer_word(void* p, uint32_t v, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_peer_u4(const void* p) {  // written by other GPUs: never the read-only path
    uint4 r;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
// Four words (a, b: two each) of one rank's partial must carry `tag`; late words are re-read.  A peer that has not
// delivered after kTpTimeoutNs (a dead rank, or call sequences that diverged) is fatal: the flag is raised for the
// host and the kernel traps instead of computing with unvalidated words.
#ifndef FP4_TP_FLY
#define FP4_TP_FLY 4
#endif
constexpr int kTpFly = FP4_TP_FLY;  // ranks whose words a consumer thread loads before it looks at any of them
constexpr unsigned long long kTpTimeoutNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void tp_wait4(const uint8_t* src, uint32_t tag, uint4& a, uint4& b, uint32_t* err) {
    unsigned long long t0 = 0;
    for (uint32_t spins = 0; !(a.y == tag && a.w == tag && b.y == tag && b.w == tag); ++spins) {
        a = ld_peer_u4(src);
        b = ld_peer_u4(src + 16);
        if ((spins & 4095u) == 4095u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) {
                t0 = now;
            } else if (now - t0 > kTpTimeoutNs) {
                *reinterpret_cast<volatile uint32_t*>(err) = 1u;
                __threadfence_system();
                asm volatile("trap;");
            }
        }
    }
}

// HALF: K % 512 == 256, i.e. the last unit of every row tile holds two steps instead of four
// EXTRA: the rarely used features - nested absmax, the gated-MLP / residual epilogues, the tensor-parallel exchange -
// are compiled into a second instantiation, so the plain decode launch carries none of their code or registers
template <typename T, int NCT, bool HALF, bool ALIGNED, bool EXTRA>
__global__ void __launch_bounds__(kThreads, FP4_STREAM_MINB) gemv_stream_kernel(const __grid_constant__ Params p) {
    const uint32_t nested_mask = EXTRA ? p.nested_mask : 0u;
    const int gated_kind = EXTRA ? p.gated : 0;
    constexpr int TERMS = sizeof(T) == 4 ? 4 : 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t g = lane >> 2, t = lane & 3;
    const int batch = p.batch, nq = batch * TERMS;
    const uint32_t K = (uint32_t)p.K, nkb = K >> 6, rowb = K >> 1;  // rowb: packed bytes per weight row

    // ---- shared memory carve-up -----------------------------------------------------------------
    // ring_a: where this lane WRITES (and, unless ALIGNED, reads back) its 16 bytes of a 512-byte half step
    const uint32_t ring_w = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)warp * p.ring * kSlot;
    const uint32_t ring_a = ring_w + (ALIGNED ? (t * 8 + g) * 16 : lane * 16);
    const uint32_t ring_am = ring_w + 4096 + lane * 16;   // absmax: lane-private in both variants
    const uint32_t ring_rd = ring_w + g * 16 + t * 4;     // ALIGNED: word t of chunk i of row g is at + i*128
    const uint32_t XP = ALIGNED ? K + 16 : K;             // pitch of an x row
    uint8_t* sX = smem + (size_t)kW * p.ring * kSlot;     // [nq][XP] s8, pairs of k swapped (nibble order)
    uint8_t* sZero = sX + (size_t)nq * XP;                // kZeroBytes of zeros
    float* sXs = reinterpret_cast<float*>(sZero + kZeroBytes);  // [batch][nkb] 2^-e / 192
    float* sPart = sXs + (size_t)batch * nkb;             // [kW][2][batch*16]
    float* sCode2 = sPart + (size_t)kW * 2 * batch * 16;  // [kMaxGroup][256], nested absmax only

    // ---- this CTA's tiles and this warp's units ----------------------------------------------------
    const uint32_t cta = blockIdx.x;
    const uint32_t tile0 = cta * p.tq + (cta < p.tr ? cta : p.tr);
    const uint32_t ntile = p.tq + (cta < p.tr ? 1u : 0u);
    const uint32_t nU = ntile * p.upt;
    const uint32_t wq = nU / kW, wr = nU % kW;
    const uint32_t ua = (uint32_t)warp * wq + ((uint32_t)warp < wr ? (uint32_t)warp : wr);  // first unit (CTA-local)
    const uint32_t n = wq + ((uint32_t)warp < wr ? 1u : 0u);
    uint32_t tl_a, ku_a;  // CTA-local tile and unit-in-tile of the first unit
    p.by_upt.divmod(ua, tl_a, ku_a);

    // matrix of a global tile (grouped launches: the matrices' row tiles are numbered consecutively)
    auto mat_of = [&](uint32_t gt) {
        return (int)(gt >= p.tstart[1]) + (int)(gt >= p.tstart[2]) + (int)(gt >= p.tstart[3]);
    };
    // loader: lane (g, t) reads 16 B of row g and 16 B of row g + 8 per 128-k step; 4 steps per unit
    // byte distance from a lane's first weight row (MMA row g) to its second (MMA row g + 8): eight rows further
    // down the same matrix, or the same row of the up projection in gated mode
    const bool gated = gated_kind != 0;
    const size_t row8 = gated ? (size_t)(p.vpacked[1] - p.vpacked[0]) : (size_t)8 * rowb;
    const uint32_t trows = gated ? 8u : 16u;  // weight rows of one matrix per row tile
    const uint8_t* wp;
    const float* ap;  // absmax of a unit: 16 rows x 8 blocks; lane (g, t) holds row g + 8 (t & 1), blocks 4 (t >> 1) ..
    const uint8_t* qp = nullptr;  // nested: this lane's four uint8 codes ...
    const float* a2p = nullptr;   // ... and the absmax2 of their 256-group (four aligned blocks share one)
    bool ld_nested = false;
    uint32_t ld_gt = tile0 + tl_a;
    auto loader_at = [&](uint32_t gt, uint32_t kunit) {
        const int m = mat_of(gt);
        const size_t trow = (size_t)gt * trows;
        wp = p.vpacked[m] + (trow + g) * rowb + kunit * 256 + t * 16;
        if (gated) ap = p.vabsmax[t & 1] + (trow + g) * nkb + kunit * 8 + 4 * (t >> 1);
        else ap = p.vabsmax[m] + (trow + g + 8 * (t & 1)) * nkb + kunit * 8 + 4 * (t >> 1);
        ld_nested = (nested_mask >> m) & 1u;
        if (ld_nested) {
            qp = p.vqabs[m] + (trow + g + 8 * (t & 1)) * nkb + kunit * 8 + 4 * (t >> 1);
            const size_t lblk = ((size_t)(gt - p.tstart[m]) * 16 + g + 8 * (t & 1)) * nkb + kunit * 8 + 4 * (t >> 1);
            a2p = p.vabs2[m] + (lblk >> p.vbs2_log2[m]);
        }
    };
    loader_at(ld_gt, ku_a);
    TL_STAMP(0);
    // tensor parallel (see the helpers above): a PRODUCER pushes its output into every rank's exchange buffer, a
    // CONSUMER's x is the sum of the ranks' partials read out of its own exchange buffer
    const int in_world = EXTRA ? p.tp.in_world : 0, out_world = EXTRA ? p.tp.out_world : 0;
    const bool producer = out_world > 1;
    // a consumer that reads nothing but the exchange needs no grid dependency at all: every word it reads validates
    // itself, so it starts polling while the producer (of this and of the other ranks) is still running - no
    // completion -> release latency, no wait for the slowest CTA.  The producer in turn lets its dependents start
    // only AFTER its own wait, so a consumer never overtakes the kernels before its producer.
    const bool free_running = EXTRA && in_world > 1 && !producer && p.x == nullptr;
    // the next kernel of the stream may be scheduled as soon as every CTA of this one is running: its CTAs then
    // start (and request their first ring slot) the moment a CTA of this kernel leaves its SM
    if (!producer) asm volatile("griddepcontrol.launch_dependents;");
    uint32_t ld_ku = ku_a, issued = 0;
    // copy the next unit of this warp's range into ring slot `dst` (lane-private bytes) and advance
    // (steps [j0, j1) of the unit; the absmax rides with step 0, the loader advances after step 3)
    auto issue_steps = [&](uint32_t dst, uint32_t j0, uint32_t j1) {
        const uint32_t nst = (HALF && ld_ku + 1 == p.upt) ? 2u : 4u;  // a tile's last unit may be half a unit
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((uint32_t)j >= j0 && (uint32_t)j < j1 && (uint32_t)j < nst) {
                cp_async_cg16(dst + j * 1024, wp + j * 64);
                cp_async_cg16(dst + j * 1024 + 512, wp + j * 64 + row8);
            }
        }
        if (j0 == 0 && 2 * (t >> 1) < nst) {  // this lane's 4 blocks exist
            if (ld_nested) {  // four codes + the absmax2 of their group ride in the slot (decoded when consumed)
                cp_async_ca4(dst - ring_a + ring_am, qp);
                cp_async_ca4(dst - ring_a + ring_am + 4, a2p);
            } else {
                cp_async_ca16(dst - ring_a + ring_am, ap);
            }
        }
        if (j1 < 4) return;
        if (++ld_ku == p.upt) {
            ld_ku = 0;
            loader_at(++ld_gt, 0);  // next row tile (possibly the next matrix of the group)
        } else {
            wp += 256;
            ap += 8;
            if (ld_nested) loader_at(ld_gt, ld_ku);  // (the 256-group of the codes may change)
        }
        ++issued;
    };
    auto issue_unit = [&](uint32_t dst) { issue_steps(dst, 0, 4); };
    // ---- 1. fill the ring: nothing here depends on the previous kernel in the stream ----------------
    // (only the first slot here: issuing a deep ring costs microseconds of LSU time that would delay the
    // x staging everything else waits for; the other slots are filled right after it)
    // p.pre steps of the first slot go out here; the rest of it follows the x loads (same cp.async group)
    const bool have0 = issued < n;
    if (have0) issue_steps(ring_a, 0, p.pre);
    if (p.pre >= 4) cp_async_commit();  // one group per slot, empty or not: the number of pending groups stays `ring`
    TL_STAMP(1);
    // x and `out` may be products of the previous kernel: wait for it
    if (!free_running) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (producer) asm volatile("griddepcontrol.launch_dependents;");
    TL_STAMP(2);
    // Epoch of the exchange = epochs[1] + 1, where epochs[1] is the last epoch this rank's consumers finished: a
    // producer publishes it, the consumer after it consumes it and advances epochs[1] when it ends.  The word is
    // written by the previous consumer, which completed before the producer passed its wait, i.e. before either
    // kernel of the current pair started.
    uint32_t in_tag = 0, out_tag = 0, e_in = 0, e_out = 0;
    const uint8_t* in_slot = nullptr;
    size_t out_off = 0;
    if (in_world > 1) {
        e_in = __ldcg(p.tp.epochs + 1) + 1u;
        in_tag = e_in;
        in_slot = reinterpret_cast<const uint8_t*>(p.tp.in_base) + (size_t)(e_in & 1u) * in_world * p.tp.slot_bytes;
    }
    if (out_world > 1) {
        e_out = __ldcg(p.tp.epochs + 1) + 1u;
        out_tag = e_out;
        out_off = ((size_t)(e_out & 1u) * out_world + p.tp.out_rank) * p.tp.slot_bytes;
    }

    // ---- 2. stage x as s8 residual terms, one power-of-two scale per (batch row, 64-block) ----------
    for (uint32_t i = tid; i < kZeroBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(sZero)[i] = 0u;
    if (nested_mask) {
        for (uint32_t i = tid; i < kMaxGroup * 256; i += kThreads) {
            const uint32_t m = i >> 8;
            if ((nested_mask >> m) & 1u) sCode2[i] = __ldg(p.vcode2[m] + (i & 255u));
        }
    }
    {
        const T* x = reinterpret_cast<const T*>(p.x);
        const int nchunk = (int)(K >> 3);
        // one chunk = 8 consecutive elements of batch row b; a 64-block = the chunks of 8 consecutive lanes
        // the rest of the ring is issued BETWEEN the loads of x and their first use: the ~0.6 us of LSU time the
        // cp.asyncs take overlap the L2 round trip of x instead of following it
        bool ring_filled = false;
        auto fill_ring = [&]() {
            if (ring_filled) return;
            ring_filled = true;
            if (p.pre < 4) {
                if (have0) issue_steps(ring_a, p.pre, 4);
                cp_async_commit();
            }
            for (uint32_t r = 1; r < p.ring; ++r) {
                if (issued < n) issue_unit(ring_a + r * kSlot);
                cp_async_commit();
            }
        };
        auto stage_chunk = [&](int b, int c, float (&f)[8]) {
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fabsf(f[i]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            uint32_t E = (__float_as_uint(mx) >> 23) & 0xFFu;
            E = E < 32u ? 32u : (E > 250u ? 250u : E);
            const float s = __uint_as_float((259u - E) << 23);  // 2^(5 - e): |x * s| < 64
            if ((c & 7) == 0) sXs[b * nkb + (c >> 3)] = __uint_as_float((E - 5u) << 23) * (1.f / 192.f);
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = f[i] * s;
            // ALIGNED: within a 128-k step the 8-byte chunk (i, t) sits at t*32 + i*8 (a lane's four chunks
            // are contiguous); otherwise natural order
            const uint32_t cpos = ALIGNED ? ((uint32_t)c >> 4) * 128 + ((uint32_t)c & 3) * 32 + (((uint32_t)c >> 2) & 3) * 8
                                          : (uint32_t)c * 8;
            uint8_t* dst = sX + (size_t)(b * TERMS) * XP + cpos;
#pragma unroll
            for (int j = 0; j < TERMS; ++j) {
                uint32_t ti[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float rr = y[i] + kMagic;  // round to nearest integer
                    ti[i] = __float_as_uint(rr);
                    if (j + 1 < TERMS) y[i] = (y[i] - (rr - kMagic)) * 128.f;  // exact residual, rescaled
                }
                // byte order = nibble order of the packed weights: (k+1, k, k+3, k+2)
                const uint32_t w0 = prmt(prmt(ti[1], ti[0], 0x0040u), prmt(ti[3], ti[2], 0x0040u), 0x5410u);
                const uint32_t w1 = prmt(prmt(ti[5], ti[4], 0x0040u), prmt(ti[7], ti[6], 0x0040u), 0x5410u);
                *reinterpret_cast<uint2*>(dst + (size_t)j * XP) = make_uint2(w0, w1);
            }
        };
        if (in_world > 1) {
            if constexpr (sizeof(T) == 2) {
                // thread c reads words 4c..4c+3 of each rank (the loads of up to four ranks in flight together): rows
                // 16*(c/2) + 4*(c&1) + i in the low halves and those 8 further in the high halves; lane pairs then swap
                // what the other needs, so the even lane holds chunk c and the odd lane chunk c of x as usual
                for (int b = 0; b < batch; ++b) {
                    for (int c = tid; c < nchunk; c += kThreads) {
                        const uint8_t* src = in_slot + ((size_t)b * (K >> 1) + (size_t)c * 4) * 8;
                        float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int r0 = 0; r0 < in_world; r0 += kTpFly) {
                            uint4 wa[kTpFly], wb[kTpFly];
#pragma unroll
                            for (int j = 0; j < kTpFly; ++j) {
                                if (r0 + j < in_world) {
                                    wa[j] = ld_peer_u4(src + (size_t)(r0 + j) * p.tp.slot_bytes);
                                    wb[j] = ld_peer_u4(src + (size_t)(r0 + j) * p.tp.slot_bytes + 16);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < kTpFly; ++j) {  // fixed rank order: every rank computes the same x
                                if (r0 + j < in_world) {
                                    tp_wait4(src + (size_t)(r0 + j) * p.tp.slot_bytes, in_tag, wa[j], wb[j], p.tp.err);
                                    lo[0] += tp_lo<T>(wa[j].x); hi[0] += tp_hi<T>(wa[j].x);
                                    lo[1] += tp_lo<T>(wa[j].z); hi[1] += tp_hi<T>(wa[j].z);
                                    lo[2] += tp_lo<T>(wb[j].x); hi[2] += tp_hi<T>(wb[j].x);
                                    lo[3] += tp_lo<T>(wb[j].z); hi[3] += tp_hi<T>(wb[j].z);
                                }
                            }
                        }
                        const bool odd = c & 1;
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float got = __shfl_xor_sync(0xffffffffu, odd ? lo[i] : hi[i], 1);
                            f[i] = odd ? got : lo[i];      // even lane: rows 0..3 its own, 4..7 the odd lane's low halves
                            f[i + 4] = odd ? hi[i] : got;  // odd lane: rows 8..11 the even lane's high halves, 12..15 its own
                        }
                        // x = the all-reduced activation AS A TENSOR OF T would hold it (fp32 sum in rank order, rounded
                        // once): bit-identical to gathering the partials and summing them on the host side
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = DT<T>::to_f32(DT<T>::from_f32(f[i]));
                        stage_chunk(b, c, f);
                    }
                }
            }
        } else {
            // x is [batch][K] contiguous: chunk id = b * nchunk + c.  Four loads in flight per thread - one L2
            // round trip per four chunks instead of one each (8 rows of K = 8192 are 16 chunks per thread)
            constexpr int U = 4;
            const int total = batch * nchunk;  // a multiple of 32: the lanes of a warp are in or out together
            if (total <= kThreads) {           // one chunk per thread (batch 1, K <= 4096): nothing to overlap
                if (tid < total) {
                    float f[8];
                    XLoad<T>::load(x + (size_t)tid * 8, f);
                    fill_ring();
                    const int b = batch == 1 ? 0 : tid / nchunk;
                    stage_chunk(b, tid - b * nchunk, f);
                }
            } else
            for (int base = tid; base < total; base += kThreads * U) {
                float f[U][8];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int id = base + u * kThreads;
                    if (id < total) XLoad<T>::load(x + (size_t)id * 8, f[u]);
                }
                if (batch <= 2) fill_ring();  // (with more rows to stage, x first measured 1 % better)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int id = base + u * kThreads;
                    if (id < total) {
                        const int b = id / nchunk;
                        stage_chunk(b, id - b * nchunk, f[u]);
                    }
                }
            }
        }
        fill_ring();  // threads without a chunk, and the tensor-parallel path
    }
    __syncthreads();
    TL_STAMP(3);

    // ---- 3. main loop over this warp's units --------------------------------------------------------
    // B fragments: MMA column ct*8 + g = ((batch row, term) q, block blk); lanes of the other block's
    // k-group read zeros
    uint32_t xbase[NCT], xstep[NCT], sbase[NCT];
    const uint32_t sX_a = (uint32_t)__cvta_generic_to_shared(sX);
    const uint32_t sZero_a = (uint32_t)__cvta_generic_to_shared(sZero);
    const uint32_t sXs_a = (uint32_t)__cvta_generic_to_shared(sXs);
#pragma unroll
    for (int ct = 0; ct < NCT; ++ct) {
        if constexpr (ALIGNED) {
            const int q = ct * 8 + (int)g;  // MMA column g of tile ct = (batch row, term) q
            const bool valid = q < nq;
            xbase[ct] = valid ? sX_a + (uint32_t)q * XP + t * 32 : sZero_a + 16;
            xstep[ct] = valid ? 512u : 0u;
            int qs = ct * 8 + 2 * (int)t;   // this lane reads back columns 2t, 2t+1: two terms of one batch row
            qs = qs < nq ? qs : 0;
            sbase[ct] = sXs_a + (uint32_t)(qs / TERMS) * nkb * 4;
        } else {
            const int col = ct * 8 + (int)g, q = col >> 1, blk = col & 1;
            const bool valid = q < nq && (int)(t >> 1) == blk;
            xbase[ct] = valid ? sX_a + (uint32_t)q * K + t * 32 : sZero_a + 16;
            xstep[ct] = valid ? 512u : 0u;  // bytes per unit
            int qs = ct * 4 + (int)t;       // the (batch row, term) whose two columns this lane reads back
            qs = qs < nq ? qs : 0;
            sbase[ct] = sXs_a + (uint32_t)(qs / TERMS) * nkb * 4;
        }
    }
    uint32_t tab_lo;
    asm volatile("mov.b32 %0, 0xC0800100;" : "=r"(tab_lo));  // 192*|code[0..3]| = 0, 1, 128, 192
    int magic_i;
    asm volatile("mov.b32 %0, 0x4B400000;" : "=r"(magic_i));
    const uint32_t quad = lane & 28u;

    // !ALIGNED: [0], [1] = rows g, g + 8 of (batch row, term) ct*4 + t.
    //  ALIGNED: [0], [1] = row g of columns ct*8 + 2t, + 1;  [2], [3] = row g + 8
    float acc[NCT][ALIGNED ? 4 : 2];
#pragma unroll
    for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
        for (int i = 0; i < (ALIGNED ? 4 : 2); ++i) acc[ct][i] = 0.f;

    float* myPart = sPart + (size_t)warp * 2 * batch * 16;

    // finish this warp's share (cnt units) of CTA-local tile tl
    auto flush = [&](uint32_t tl, uint32_t cnt) {
        const uint32_t row0 = (tile0 + tl) * trows;
        const bool whole = cnt == p.upt;
        const int mm = mat_of(tile0 + tl);
        const T* bias = reinterpret_cast<const T*>(p.vbias[mm]);
        const T* res = EXTRA ? reinterpret_cast<const T*>(p.vres[mm]) : nullptr;
        T* out = reinterpret_cast<T*>(p.vout[mm]);
        const size_t Nm = (size_t)p.Nm[mm];
        float* part = myPart + (tl == tl_a ? 0 : batch * 16);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) {
            float v0, v1;
            int b;
            bool owner;
            if constexpr (ALIGNED) {
                if constexpr (TERMS == 2) {  // the lane holds both terms of batch row ct*4 + t
                    v0 = fmaf(acc[ct][1], 1.f / 128.f, acc[ct][0]);
                    v1 = fmaf(acc[ct][3], 1.f / 128.f, acc[ct][2]);
                    owner = true;
                    b = ct * 4 + (int)t;
                } else {  // terms 2(t&1), 2(t&1)+1 of batch row ct*2 + (t>>1); the lane t^1 holds the other two
                    const float wa = (t & 1) ? (1.f / 16384.f) : 1.f, wb = wa * (1.f / 128.f);
                    v0 = fmaf(acc[ct][1], wb, acc[ct][0] * wa);
                    v1 = fmaf(acc[ct][3], wb, acc[ct][2] * wa);
                    v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
                    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                    owner = (t & 1) == 0;
                    b = ct * 2 + (int)(t >> 1);
                }
                acc[ct][0] = acc[ct][1] = acc[ct][2] = acc[ct][3] = 0.f;
            } else {
            v0 = acc[ct][0]; v1 = acc[ct][1];
            acc[ct][0] = acc[ct][1] = 0.f;
            if constexpr (TERMS == 2) {
                const float o0 = __shfl_xor_sync(0xffffffffu, v0, 1), o1 = __shfl_xor_sync(0xffffffffu, v1, 1);
                v0 = fmaf(o0, 1.f / 128.f, v0);
                v1 = fmaf(o1, 1.f / 128.f, v1);
                owner = (t & 1) == 0;
                b = ct * 2 + (int)(t >> 1);
            } else {
                const float wgt = t == 0 ? 1.f : t == 1 ? (1.f / 128.f) : t == 2 ? (1.f / 16384.f) : (1.f / 2097152.f);
                v0 *= wgt;
                v1 *= wgt;
                v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
                v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
                v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
                owner = t == 0;
                b = ct;
            }
            }
            if (owner && b < batch) {
                if (whole && gated) {  // v0 = gate row, v1 = up row of output element row0 + g
                    const uint32_t r0 = row0 + g;
                    if (bias) v0 += DT<T>::to_f32(bias[r0]);
                    if (p.vbias[1]) v1 += DT<T>::to_f32(reinterpret_cast<const T*>(p.vbias[1])[r0]);
                    out[(size_t)b * Nm + r0] = DT<T>::from_f32(gate_act(v0, gated_kind) * v1);
                } else if (whole) {
                    const uint32_t r0 = row0 + g, r1 = r0 + 8;
                    if (bias) {
                        v0 += DT<T>::to_f32(bias[r0]);
                        v1 += DT<T>::to_f32(bias[r1]);
                    }
                    if (res) {
                        v0 += DT<T>::to_f32(res[(size_t)b * Nm + r0]);
                        v1 += DT<T>::to_f32(res[(size_t)b * Nm + r1]);
                    }
                    if (out_world > 1) {
                        if constexpr (sizeof(T) == 2) {
                            const uint32_t w2 = tp_pack2<T>(v0, v1);
                            const size_t woff = out_off + ((((size_t)b * Nm + row0) >> 1) + g) * 8;
                            for (int q = 0; q < out_world; ++q)
                                st_peer_word(reinterpret_cast<uint8_t*>(p.tp.out_peer_base[q]) + woff, w2, out_tag);
                        }
                    } else {
                        out[(size_t)b * Nm + r0] = DT<T>::from_f32(v0);
                        out[(size_t)b * Nm + r1] = DT<T>::from_f32(v1);
                    }
                } else {
                    part[b * 16 + g] = v0;
                    part[b * 16 + g + 8] = v1;
                }
            }
        }
    };

    uint32_t tl = tl_a, ku = ku_a, cnt = 0, slot = 0;
    for (uint32_t left = n; left; --left) {
        const bool more = left > 1;
        cp_async_wait_pending(p.ring - 1);  // the oldest group (this unit) has landed
#ifdef FP4_STREAM_TIMELINE
        if (left == n) TL_STAMP(4);
#endif
        const uint32_t sl = ring_a + slot * kSlot;
        if constexpr (ALIGNED) __syncwarp();  // the lanes read each other's copies
        uint4 amc = lds_u4(ring_am + slot * kSlot);
        if (nested_mask) {  // uniform
            const int mm = mat_of(tile0 + tl);
            if ((nested_mask >> mm) & 1u) {
                // absmax = fp32_add(fp32_mul(code2[q], absmax2), offset): two separately rounded operations
                const uint32_t codes = amc.x;
                const float a2 = __uint_as_float(amc.y), off = p.voffset[mm];
                const float* tab = sCode2 + mm * 256;
                amc.x = __float_as_uint(__fadd_rn(__fmul_rn(tab[codes & 255u], a2), off));
                amc.y = __float_as_uint(__fadd_rn(__fmul_rn(tab[(codes >> 8) & 255u], a2), off));
                amc.z = __float_as_uint(__fadd_rn(__fmul_rn(tab[(codes >> 16) & 255u], a2), off));
                amc.w = __float_as_uint(__fadd_rn(__fmul_rn(tab[codes >> 24], a2), off));
            }
        }
        uint32_t xa[NCT], sa[NCT];
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) {
            xa[ct] = xbase[ct] + ku * xstep[ct];
            sa[ct] = sbase[ct] + ku * 32;
        }
        const uint32_t nst_c = (HALF && ku + 1 == p.upt) ? 2u : 4u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (HALF && (uint32_t)j >= nst_c) break;
            if constexpr (ALIGNED) {
                const uint32_t rd = ring_rd + slot * kSlot + j * 1024;
                const uint32_t srcA = quad | (2 * (j >> 1)), srcB = srcA + 1;
                const uint32_t c0 = (j & 1) ? amc.z : amc.x, c1 = (j & 1) ? amc.w : amc.y;
                const float am[2][2] = {  // [row g / g+8][block 2j / 2j+1]
                    {__uint_as_float(__shfl_sync(0xffffffffu, c0, srcA)), __uint_as_float(__shfl_sync(0xffffffffu, c1, srcA))},
                    {__uint_as_float(__shfl_sync(0xffffffffu, c0, srcB)), __uint_as_float(__shfl_sync(0xffffffffu, c1, srcB))}};
                uint4 bq[NCT][2];
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    bq[ct][0] = lds_u4(xa[ct] + j * 128);
                    bq[ct][1] = lds_u4(xa[ct] + j * 128 + 16);
                }
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    int dall[NCT][4], dneg[NCT][4];
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii) {
                        const int i = 2 * blk + ii;
                        const uint32_t wA = gemv::lds_u32(rd + i * 128), wB = gemv::lds_u32(rd + 512 + i * 128);
                        uint32_t aA0, aA1, nA0, nA1, aB0, aB1, nB0, nB1;
                        decode_word(wA, tab_lo, aA0, aA1, nA0, nA1);
                        decode_word(wB, tab_lo, aB0, aB1, nB0, nB1);
#pragma unroll
                        for (int ct = 0; ct < NCT; ++ct) {
                            const uint4 b4 = bq[ct][blk];
                            const uint2 bx = ii == 0 ? make_uint2(b4.x, b4.y) : make_uint2(b4.z, b4.w);
                            if (ii == 0) {
                                imma_first(dall[ct], aA0, aB0, aA1, aB1, bx.x, bx.y, magic_i);
                                imma_first(dneg[ct], nA0, nB0, nA1, nB1, bx.x, bx.y, magic_i);
                            } else {
                                imma_acc(dall[ct], aA0, aB0, aA1, aB1, bx.x, bx.y);
                                imma_acc(dneg[ct], nA0, nB0, nA1, nB1, bx.x, bx.y);
                            }
                        }
                    }
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        const float xs = __uint_as_float(gemv::lds_u32(sa[ct] + j * 8 + blk * 4));
                        const float s0 = am[0][blk] * xs, s1 = am[1][blk] * xs;
                        const float f0 = fmaf(__int_as_float(dneg[ct][0]), -2.f, __int_as_float(dall[ct][0])) + kMagic;
                        const float f1 = fmaf(__int_as_float(dneg[ct][1]), -2.f, __int_as_float(dall[ct][1])) + kMagic;
                        const float f2 = fmaf(__int_as_float(dneg[ct][2]), -2.f, __int_as_float(dall[ct][2])) + kMagic;
                        const float f3 = fmaf(__int_as_float(dneg[ct][3]), -2.f, __int_as_float(dall[ct][3])) + kMagic;
                        acc[ct][0] = fmaf(f0, s0, acc[ct][0]);
                        acc[ct][1] = fmaf(f1, s0, acc[ct][1]);
                        acc[ct][2] = fmaf(f2, s1, acc[ct][2]);
                        acc[ct][3] = fmaf(f3, s1, acc[ct][3]);
                    }
                }
                continue;
            }
            const uint4 wA4 = lds_u4(sl + j * 1024), wB4 = lds_u4(sl + j * 1024 + 512);
            const uint32_t wA[4] = {wA4.x, wA4.y, wA4.z, wA4.w};
            const uint32_t wB[4] = {wB4.x, wB4.y, wB4.z, wB4.w};
            // absmax of blocks (2j, 2j+1) of rows g / g+8 sit in lanes t = 2(j>>1) / 2(j>>1)+1 of the quad
            const uint32_t srcA = quad | (2 * (j >> 1)), srcB = srcA + 1;
            const uint32_t c0 = (j & 1) ? amc.z : amc.x, c1 = (j & 1) ? amc.w : amc.y;
            const float amA0 = __uint_as_float(__shfl_sync(0xffffffffu, c0, srcA));
            const float amA1 = __uint_as_float(__shfl_sync(0xffffffffu, c1, srcA));
            const float amB0 = __uint_as_float(__shfl_sync(0xffffffffu, c0, srcB));
            const float amB1 = __uint_as_float(__shfl_sync(0xffffffffu, c1, srcB));
            int dall[NCT][4], dneg[NCT][4];
            uint4 bx01 = make_uint4(0, 0, 0, 0), bx23 = bx01;
            if constexpr (NCT == 1) {
                bx01 = lds_u4(xa[0] + j * 128);
                bx23 = lds_u4(xa[0] + j * 128 + 16);
            }
            // x fragments: 128-bit loads only.  A quarter-warp (lanes g = 2i, 2i+1) reads ONE x row plus the zero
            // word, so the load is conflict-free; 64-bit loads would put the four rows of a column tile on the
            // same banks (row pitch K = 0 mod 128): a 4-way conflict that made batch 3..8 shared-memory bound
            uint4 bq[NCT];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                uint32_t aA0, aA1, nA0, nA1, aB0, aB1, nB0, nB1;
                decode_word(wA[m], tab_lo, aA0, aA1, nA0, nA1);
                decode_word(wB[m], tab_lo, aB0, aB1, nB0, nB1);
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    uint2 bx;
                    if constexpr (NCT == 1) {
                        bx = m == 0 ? make_uint2(bx01.x, bx01.y) : m == 1 ? make_uint2(bx01.z, bx01.w)
                           : m == 2 ? make_uint2(bx23.x, bx23.y) : make_uint2(bx23.z, bx23.w);
                    } else {
                        if ((m & 1) == 0) bq[ct] = lds_u4(xa[ct] + j * 128 + m * 8);
                        bx = (m & 1) == 0 ? make_uint2(bq[ct].x, bq[ct].y) : make_uint2(bq[ct].z, bq[ct].w);
                    }
                    if (m == 0) {
                        imma_first(dall[ct], aA0, aB0, aA1, aB1, bx.x, bx.y, magic_i);
                        imma_first(dneg[ct], nA0, nB0, nA1, nB1, bx.x, bx.y, magic_i);
                    } else {
                        imma_acc(dall[ct], aA0, aB0, aA1, aB1, bx.x, bx.y);
                        imma_acc(dneg[ct], nA0, nB0, nA1, nB1, bx.x, bx.y);
                    }
                }
            }
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                const uint2 xs = lds_u2(sa[ct] + j * 8);
                const float xs0 = __uint_as_float(xs.x), xs1 = __uint_as_float(xs.y);
                // both accumulators started at the bit pattern of 1.5*2^23, so as floats they read
                // 1.5*2^23 + sum exactly; (M + all) - 2 (M + neg) + M = all - 2 neg, every step exact
                const float f0 = fmaf(__int_as_float(dneg[ct][0]), -2.f, __int_as_float(dall[ct][0])) + kMagic;
                const float f1 = fmaf(__int_as_float(dneg[ct][1]), -2.f, __int_as_float(dall[ct][1])) + kMagic;
                const float f2 = fmaf(__int_as_float(dneg[ct][2]), -2.f, __int_as_float(dall[ct][2])) + kMagic;
                const float f3 = fmaf(__int_as_float(dneg[ct][3]), -2.f, __int_as_float(dall[ct][3])) + kMagic;
                acc[ct][0] = fmaf(f0, amA0 * xs0, acc[ct][0]);
                acc[ct][0] = fmaf(f1, amA1 * xs1, acc[ct][0]);
                acc[ct][1] = fmaf(f2, amB0 * xs0, acc[ct][1]);
                acc[ct][1] = fmaf(f3, amB1 * xs1, acc[ct][1]);
            }
        }
        // every word of the slot has been decoded: refill it with the unit `ring` ahead
        if constexpr (ALIGNED) __syncwarp();  // ... by every lane
        if (issued < n) issue_unit(sl);
        cp_async_commit();
        slot = slot + 1 == p.ring ? 0 : slot + 1;
        ++cnt;
        if (++ku == p.upt || !more) {
            flush(tl, cnt);
            ++tl;
            ku = 0;
            cnt = 0;
        }
    }

    // ---- 4. tiles shared between warps: sum the parked partials in warp order ------------------------
    TL_STAMP(5);
    __syncthreads();
    TL_STAMP(6);
    {
        const uint32_t per = 16u * (uint32_t)batch;
        const uint32_t big = wr * (wq + 1);
        auto unit_owner = [&](uint32_t u) { return u < big ? u / (wq + 1) : wr + (u - big) / (wq ? wq : 1u); };
        for (uint32_t idx = tid; idx < ntile * per; idx += kThreads) {
            const uint32_t tt = idx / per, e = idx - tt * per;
            const uint32_t wa = unit_owner(tt * p.upt), wz = unit_owner(tt * p.upt + p.upt - 1);
            if (wa == wz) continue;  // one warp covered the whole tile and stored it
            float v = 0.f;
            for (uint32_t w = wa; w <= wz; ++w) {
                const uint32_t w_ua = w * wq + (w < wr ? w : wr);
                const uint32_t w_tl = p.by_upt.div(w_ua);
                v += sPart[((size_t)w * 2 + (tt == w_tl ? 0 : 1)) * per + e];
            }
            const uint32_t b = e >> 4;
            // lanes l and l + 8 of a 16-lane group hold the two halves of a pair (gate / up row, or rows r / r + 8 of
            // an exchange word); the group shares tt, so all 16 lanes are here together
            const uint32_t half_mask = (tid & 16u) ? 0xFFFF0000u : 0x0000FFFFu;
            if (gated) {  // e & 15 < 8: the gate partial sums; the up row's are 8 entries further
                const uint32_t row = (tile0 + tt) * 8 + (e & 7);
                const T* bias = reinterpret_cast<const T*>(p.vbias[(e >> 3) & 1u]);
                if (bias) v += DT<T>::to_f32(bias[row]);
                const float u = __shfl_down_sync(half_mask, v, 8);
                if (!(e & 8u))
                    reinterpret_cast<T*>(p.vout[0])[(size_t)b * p.Nm[0] + row] = DT<T>::from_f32(gate_act(v, gated_kind) * u);
                continue;
            }
            const uint32_t row = (tile0 + tt) * 16 + (e & 15);
            const int mm = mat_of(tile0 + tt);
            const T* bias = reinterpret_cast<const T*>(p.vbias[mm]);
            if (bias) v += DT<T>::to_f32(bias[row]);
            if (EXTRA && p.vres[mm]) v += DT<T>::to_f32(reinterpret_cast<const T*>(p.vres[mm])[(size_t)b * p.Nm[mm] + row]);
            if (out_world > 1) {  // one word per row pair (row, row + 8), stored by the lane of the lower row
                if constexpr (sizeof(T) == 2) {
                    const float u = __shfl_down_sync(half_mask, v, 8);
                    if (!(e & 8u)) {
                        const uint32_t w2 = tp_pack2<T>(v, u);
                        const size_t woff = out_off + ((((size_t)b * p.Nm[mm] + (row & ~15u)) >> 1) + (row & 7u)) * 8;
                        for (int q = 0; q < out_world; ++q)
                            st_peer_word(reinterpret_cast<uint8_t*>(p.tp.out_peer_base[q]) + woff, w2, out_tag);
                    }
                }
            } else {
                reinterpret_cast<T*>(p.vout[mm])[(size_t)b * p.Nm[mm] + row] = DT<T>::from_f32(v);
            }
        }
    }
    // epoch bookkeeping for the NEXT kernel of this stream (kernel boundaries order these plain stores)
    if (blockIdx.x == 0 && tid == 0) {
        if (in_world > 1) *reinterpret_cast<volatile uint32_t*>(p.tp.epochs + 1) = e_in;
    }
    TL_STAMP(7);
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int nterms(int dtype) { return dtype == FP4_B200_F32 ? 4 : 2; }

// the ALIGNED variant (one MMA column per (batch row, term)): its inner loop is ~45 % slower per column tile
// (32-bit loads, two-deep IMMA chains, twice the fp32 epilogue), so it pays off where it replaces FOUR column
// tiles by two - 16-bit inputs with batch 5..8, fp32 with batch 3..4 - and it is what makes fp32 batch 5..8 fit
// (measured: profiles/r01_gemv_batch_sweep.log)
static bool use_aligned(int batch, int nt) {
    static const int enabled = env_int("FP4_B200_GEMV_ALIGNED", 1);  // 0: never, 2: always (experiments)
    return enabled == 2 || (enabled && batch * nt > 8);
}
static int column_tiles(int batch, int nt) { return (batch * nt * (use_aligned(batch, nt) ? 1 : 2) + 7) / 8; }

static size_t fixed_smem_bytes(int batch, int K, int nt) {
    return (size_t)batch * nt * (K + (use_aligned(batch, nt) ? 16 : 0)) + kZeroBytes + (size_t)batch * (K / 64) * 4 +
           (size_t)kW * 2 * batch * 16 * 4 + (size_t)kMaxGroup * 256 * 4 /* nested code2 tables */;
}

struct Group {
    const fp4_b200_tp_t* tp;
    int nmat;
    const uint8_t* packed[kMaxGroup];
    const float* absmax[kMaxGroup];
    const fp4_b200_nested_t* nested[kMaxGroup];  // non-NULL: absmax[m] is ignored
    const void* bias[kMaxGroup];
    const void* residual[kMaxGroup];
    void* out[kMaxGroup];
    int N[kMaxGroup];
    int gated;  // 0, or the gate activation: matrices 0 / 1 = gate / up, out[0] = act(gate) * up
};

// how a launch deals the row tiles to CTAs and sizes the per-warp rings
struct Partition {
    uint32_t grid, upt, tq, tr, ring;
};
static Partition partition(uint32_t tiles, int batch, int K, int nt) {
    static const int ctas_per_sm = env_int("FP4_B200_GEMV_CTAS_PER_SM", FP4_STREAM_MINB);
    static const int max_ring = env_int("FP4_B200_GEMV_RING", (int)kMaxRing);
    static const int smem_cap = env_int("FP4_B200_GEMV_SMEM_KB", (int)(kMaxSmem / 1024 / FP4_STREAM_MINB)) * 1024;
    Partition pt;
    const uint32_t max_grid = (uint32_t)(kNumSMs * (ctas_per_sm < 1 ? 1 : ctas_per_sm));
    pt.grid = tiles < max_grid ? tiles : max_grid;
    pt.upt = ((uint32_t)K + 511) / 512;
    pt.tq = tiles / pt.grid;
    pt.tr = tiles % pt.grid;
    // ring depth: no deeper than a warp has units, no larger than shared memory allows
    const size_t fixed = fixed_smem_bytes(batch, K, nt);
    const uint32_t units_per_warp = ((pt.tq + (pt.tr ? 1u : 0u)) * pt.upt + kW - 1) / kW;
    uint32_t ring = (size_t)smem_cap > fixed ? (uint32_t)(((size_t)smem_cap - fixed) / ((size_t)kW * kSlot)) : 1u;
    if (ring > units_per_warp) ring = units_per_warp;
    if (ring > (uint32_t)max_ring) ring = (uint32_t)max_ring;
    if (ring > kMaxRing) ring = kMaxRing;
    if (ring < 1) ring = 1;
    pt.ring = ring;
    return pt;
}

template <typename T, int NCT, bool HALF, bool ALIGNED, bool EXTRA>
static int launch(const void* x, const Group& gr, int batch, int K, cudaStream_t st) {
    auto kern = gemv_stream_kernel<T, NCT, HALF, ALIGNED, EXTRA>;
    // the opt-in to large dynamic shared memory is per device (a process may drive several GPUs)
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return FP4_B200_ERR_UNSUPPORTED;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    static const int use_pdl = env_int("FP4_B200_GEMV_PDL", 1);
    Params p;
    p.x = x;
    p.batch = batch; p.K = K;
    if (gr.tp) p.tp = *gr.tp; else { p.tp = fp4_b200_tp_t(); }
    p.nested_mask = 0;
    for (int m = 0; m < kMaxGroup; ++m) {
        p.vqabs[m] = nullptr; p.vcode2[m] = nullptr; p.vabs2[m] = nullptr; p.voffset[m] = 0.f; p.vbs2_log2[m] = 0;
    }
    p.gated = gr.gated;
    for (int m = 0; m < kMaxGroup; ++m) p.vres[m] = nullptr;
    uint32_t tiles = 0;
    if (gr.gated) {  // one logical matrix of N / 8 tiles; nothing is rebased
        for (int m = 0; m < kMaxGroup; ++m) {
            p.tstart[m] = m == 0 ? 0u : 0xffffffffu;
            p.vpacked[m] = m < 2 ? gr.packed[m] : nullptr;
            p.vabsmax[m] = m < 2 ? gr.absmax[m] : nullptr;
            p.vbias[m] = m < 2 ? gr.bias[m] : nullptr;
            p.vout[m] = m == 0 ? gr.out[0] : nullptr;
            p.Nm[m] = m < 2 ? gr.N[0] : 0;
        }
        tiles = (uint32_t)gr.N[0] / 8;
    } else
    for (int m = 0; m < kMaxGroup; ++m) {
        if (m < gr.nmat) {
            // rebase to global rows (integer arithmetic on addresses; never dereferenced outside the matrix)
            const size_t grow0 = (size_t)tiles * 16;
            p.tstart[m] = tiles;
            p.vpacked[m] = gr.packed[m] - grow0 * ((size_t)K / 2);
            p.vabsmax[m] = gr.absmax[m] ? gr.absmax[m] - grow0 * ((size_t)K / 64) : nullptr;
            if (gr.nested[m]) {
                const fp4_b200_nested_t& nd = *gr.nested[m];
                p.nested_mask |= 1u << m;
                p.vqabs[m] = nd.qabsmax - grow0 * ((size_t)K / 64);
                p.vcode2[m] = nd.code2; p.vabs2[m] = nd.absmax2; p.voffset[m] = nd.offset;
                p.vbs2_log2[m] = ilog2_exact(nd.blocksize2);
                p.vabsmax[m] = reinterpret_cast<const float*>(p.vqabs[m]);  // never dereferenced
            }
            p.vbias[m] = gr.bias[m] ? reinterpret_cast<const uint8_t*>(gr.bias[m]) - grow0 * sizeof(T) : nullptr;
            p.vout[m] = reinterpret_cast<uint8_t*>(gr.out[m]) - grow0 * sizeof(T);
            p.vres[m] = gr.residual[m] ? reinterpret_cast<const uint8_t*>(gr.residual[m]) - grow0 * sizeof(T) : nullptr;
            p.Nm[m] = gr.N[m];
            tiles += (uint32_t)gr.N[m] / 16;
        } else {
            p.tstart[m] = 0xffffffffu;
            p.vpacked[m] = nullptr; p.vabsmax[m] = nullptr; p.vbias[m] = nullptr; p.vout[m] = nullptr; p.Nm[m] = 0;
        }
    }
    constexpr int NT = sizeof(T) == 4 ? 4 : 2;
    const Partition pt = partition(tiles, batch, K, NT);
    const uint32_t grid = pt.grid, ring = pt.ring;
    p.upt = pt.upt; p.tq = pt.tq; p.tr = pt.tr; p.ring = pt.ring;
    p.by_upt = FastDiv(p.upt);
    const size_t fixed = fixed_smem_bytes(batch, K, NT);
    // how much of the first slot goes out before griddepcontrol.wait: the x loads queue behind those bytes.
    // Where x staging is short (one chunk per thread) and every SM streams, half a slot measured 2-3 % faster;
    // small grids and long K (more chunks of x per thread) want the whole slot
    static const int pre_steps = env_int("FP4_B200_GEMV_PRE_STEPS", 0);
    if (pre_steps > 0) p.pre = (uint32_t)(pre_steps > 4 ? 4 : pre_steps);
    else p.pre = (batch * K <= 4096 && tiles >= (uint32_t)kNumSMs) ? 2u : 4u;
    p.tl = nullptr;
#ifdef FP4_STREAM_TIMELINE
    if (g_stream_tl) p.tl = g_stream_tl + (size_t)(g_stream_tl_launch++) * (kNumSMs * kW * 8);
#endif

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = fixed + (size_t)kW * ring * kSlot;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, kern, p);
}

template <typename T, int NT>
static int launch_nct(const void* x, const Group& gr, int batch, int K, cudaStream_t st) {
    const int nct = column_tiles(batch, NT);
    const bool half = K % 512 != 0;
    bool extra = gr.gated != 0 || (gr.tp && (gr.tp->in_world > 1 || gr.tp->out_world > 1));
    for (int m = 0; m < gr.nmat; ++m) extra = extra || gr.nested[m] || gr.residual[m];
#define FP4_GO2(NCT, AL, HF) (extra ? launch<T, NCT, HF, AL, true>(x, gr, batch, K, st) : launch<T, NCT, HF, AL, false>(x, gr, batch, K, st))
#define FP4_GO(NCT, AL) (half ? FP4_GO2(NCT, AL, true) : FP4_GO2(NCT, AL, false))
    if (use_aligned(batch, NT)) {
        if (nct <= 1) return FP4_GO(1, true);
        if (nct <= 2) return FP4_GO(2, true);
        return FP4_GO(4, true);
    }
    if (nct <= 1) return FP4_GO(1, false);
    if (nct <= 2) return FP4_GO(2, false);
    return FP4_GO(4, false);  // only with FP4_B200_GEMV_ALIGNED=0
#undef FP4_GO2
#undef FP4_GO
}

static int dispatch_group(const void* x, const Group& gr, int batch, int K, int dtype, cudaStream_t st) {
    switch (dtype) {
        case FP4_B200_F16: return launch_nct<__half, 2>(x, gr, batch, K, st);
        case FP4_B200_BF16: return launch_nct<__nv_bfloat16, 2>(x, gr, batch, K, st);
        case FP4_B200_F32: return launch_nct<float, 4>(x, gr, batch, K, st);
        default: return FP4_B200_ERR_DTYPE;
    }
}

}  // namespace

#ifdef FP4_STREAM_TIMELINE
extern "C" void fp4_b200_debug_stream_timeline(long long* buf) { g_stream_tl = buf; g_stream_tl_launch = 0; }
#endif

bool gemv_stream_supported(int batch, int N, int K, int blocksize, int dtype, bool nested, const void* packed,
                           const void* absmax) {
    static const int disabled = env_int("FP4_B200_GEMV_NO_STREAM", 0);
    // (round 1 sent layers with fewer than 48 row tiles to stream-K kernels; measured again in round 2 this kernel is
    // faster down to 4 tiles - profiles/r02_gemv_family_sweep.log - and those kernels are gone)
    static const int min_tiles = env_int("FP4_B200_GEMV_STREAM_MIN_TILES", 1);
    (void)nested;  // decoded in the kernel
    if (disabled || blocksize != 64) return false;
    if (batch < 1 || batch > 8 || N <= 0 || K <= 0) return false;
    if (K % 256 != 0 || N % 16 != 0) return false;
    if (N / 16 < min_tiles) return false;
    if ((uint64_t)N * (uint64_t)K >= (1ull << 40)) return false;
    if (reinterpret_cast<uintptr_t>(packed) % 16 || reinterpret_cast<uintptr_t>(absmax) % 16) return false;
    const int nt = nterms(dtype);
    if (column_tiles(batch, nt) > 4) return false;
    return fixed_smem_bytes(batch, K, nt) + (size_t)kW * kSlot <= kMaxSmem;
}

static bool nested_ok(const fp4_b200_nested_t* nd) {
    return !nd || (nd->qabsmax && nd->code2 && nd->absmax2 && ilog2_exact(nd->blocksize2) >= 2 &&
                   reinterpret_cast<uintptr_t>(nd->qabsmax) % 4 == 0);
}

int gemv_stream_dispatch(const void* x, const uint8_t* packed, const float* absmax, const fp4_b200_nested_t* nested,
                         const void* bias, void* out, int batch, int N, int K, int dtype, cudaStream_t st) {
    Group gr = {};
    gr.nmat = 1;
    gr.packed[0] = packed; gr.absmax[0] = absmax; gr.nested[0] = nested; gr.bias[0] = bias; gr.out[0] = out; gr.N[0] = N;
    if (!nested_ok(nested)) return FP4_B200_ERR_UNSUPPORTED;
    return dispatch_group(x, gr, batch, K, dtype, st);
}

// Several weight matrices with the same K applied to the same x in ONE launch (q/k/v, gate/up): their row
// tiles are numbered consecutively and dealt to the CTAs as if they were one matrix.
bool gemv_stream_group_supported(int nmat, int batch, const int* N, int K, int blocksize, int dtype,
                                 const uint8_t* const* packed, const float* const* absmax) {
    if (nmat < 1 || nmat > kMaxGroup) return false;
    long total = 0;
    for (int m = 0; m < nmat; ++m) {
        if (N[m] <= 0 || N[m] % 16 != 0) return false;
        if (reinterpret_cast<uintptr_t>(packed[m]) % 16 || reinterpret_cast<uintptr_t>(absmax[m]) % 16) return false;  // (NULL passes)
        total += N[m];
    }
    if (total > (1 << 24)) return false;
    // same rules as one matrix of `total` rows
    return gemv_stream_supported(batch, (int)total, K, blocksize, dtype, false, packed[0], absmax[0]);
}

int gemv_stream_group_dispatch(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                               const void* const* bias, void* const* out, const int* N, int batch, int K, int dtype,
                               const fp4_b200_tp_t* tp, const fp4_b200_epilogue_t* epi, cudaStream_t st) {
    Group gr = {};
    gr.tp = tp;
    gr.nmat = nmat;
    for (int m = 0; m < nmat; ++m) {
        gr.packed[m] = packed[m]; gr.absmax[m] = absmax[m]; gr.bias[m] = bias ? bias[m] : nullptr;
        gr.out[m] = out[m]; gr.N[m] = N[m];
        gr.residual[m] = (epi && epi->residual) ? epi->residual[m] : nullptr;
        gr.nested[m] = (epi && epi->nested) ? epi->nested[m] : nullptr;
        if (!nested_ok(gr.nested[m])) return FP4_B200_ERR_UNSUPPORTED;
    }
    if (epi && epi->gate_act) {
        const bool tp_on = tp && (tp->in_world > 1 || tp->out_world > 1);
        if (nmat != 2 || N[0] != N[1] || N[0] % 8 || epi->gate_act < 1 || epi->gate_act > 2 || tp_on || epi->residual ||
            gr.nested[0] || gr.nested[1])
            return FP4_B200_ERR_UNSUPPORTED;
        gr.gated = epi->gate_act;
    }
    return dispatch_group(x, gr, batch, K, dtype, st);
}

}  // namespace fp4b200
