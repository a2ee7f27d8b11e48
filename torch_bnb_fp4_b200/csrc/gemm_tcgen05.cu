// Dequant-fused tensor-core GEMM for prefill-sized M on sm_100a:
//
//   out[m, r] = T( sum_k x[m,k] * RN_T(code[W[r,k]] * absmax[r, k/blocksize]) + bias[r] )      T = bf16 / fp16
//
// replaces the reference's dequant + cuBLAS pair (torch_bnb_fp4/__init__.py:423-436;
// csrc/torch_fp4.cpp:64-103): the weight is dequantised tile by tile straight into shared memory in the
// tcgen05 operand layout and never written to HBM (0.5625 B/weight of traffic instead of 0.5625 + 2 + 2).
// The dequantised values are bit-identical to the dequant kernel's (one IEEE fp32 multiply, one
// round-to-nearest-even), so the result equals dequant-then-GEMM up to fp32 summation order.
//
// Operands are swapped with respect to the usual GEMM: the WEIGHT tile is the MMA's A operand
// (M = 128 weight rows) and the activation tile is B (N = up to 256 tokens, any multiple of 16), so
// small token counts waste no tensor-core rows and one dequantised tile serves up to 256 tokens.
// The dequantised weight tile never touches shared memory either: the dequantiser threads write it straight
// into TENSOR MEMORY (tcgen05.st) and the MMA takes its A operand from there (tcgen05.mma with A in TMEM).
// Shared-memory bandwidth is what bounds this kernel (128 B/clk: the MMA alone reads 48 KB of operands per
// 64-k block in its 524 clocks), and this removes 32 of the ~104 KB per block that an A tile staged in shared
// memory costs (16 KB written by the dequantisers + 16 KB read back by the tensor core).
// Accumulators (128 lanes x tokens, fp32) live in TMEM too, double-buffered for token tiles <= 192 so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles of the persistent CTA (one per SM, 320 threads):
//   warp 0      TMA producer: activation tiles [tokens x 64 k] -> shared memory (128-byte swizzle), and the
//               PACKED weight tile [128 rows x 128 B = 4 k blocks] into its own ring (16 KiB boxes keep
//               ~48 KB of the 0.5 B/weight stream in flight per SM; register prefetch cannot)
//   warp 1      MMA issuer (one lane): tcgen05.mma kind::f16, 4 x (128 x tokens x 16) per stage;
//               tcgen05.commit releases the stage / publishes the accumulator
//   warps 2-5   dequantisers: thread r owns weight row r of the tile; per 64-k block it reads its 32 packed
//               bytes from the ring (swizzled: conflict-free), builds the 8
//               possible magnitudes RN_T(|code_i| * absmax) once, then decodes 64 nibbles with byte
//               permutes (PRMT as an 8-entry table, sign bit merged with one LOP3) and stores its row
//               (32 columns of two 16-bit values) into the A ring in tensor memory
//   warps 6-9   epilogue: tcgen05.ld (thread = output feature, registers = tokens), bias, convert, store
//
// Requirements: bitsandbytes FP4 codebook (code == NULL or FP4_B200_FLAG_CODE_IS_BNB_FP4), fp32 absmax,
// K % 64 == 0, blocksize % 64 == 0, 16-byte aligned x / packed.  Other inputs: FP4_B200_ERR_UNSUPPORTED
// (the caller then takes dequant + library GEMM, which is what the reference always does).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace fp4b200 {
namespace {

constexpr int BW = 128;  // weight rows per tile (MMA M)
constexpr int BK = 64;   // k per pipeline stage = one 128-byte swizzle row of 16-bit elements
constexpr int kThreads = 320;
constexpr int kStagesA = 8;                // ring of dequantised weight tiles in tensor memory
constexpr uint32_t kColsA = BK / 2;        // 32 TMEM columns per stage: 128 lanes x 64 k x 16 bit
constexpr uint32_t kColA0 = 512 - kStagesA * kColsA;  // the A ring sits at the top of the 512 columns
constexpr int kWSlots = 3;                 // ring of packed-weight boxes
constexpr uint32_t kWBox = BW * 128;       // 128 rows x 128 packed bytes = 4 k blocks

// 12 * |code| of the bitsandbytes FP4 table is exact in binary; the table itself (fp32) is the one the
// dequant kernel multiplies with
__constant__ float kBnbMag[8] = {0.00000000f, 5.208333333e-03f, 0.66666667f, 1.00000000f,
                                 0.33333333f, 0.50000000f,      0.16666667f, 0.25000000f};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
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
// same with the A operand in tensor memory (128 lanes x 8 columns of packed 16-bit pairs per K = 16 step)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one warp stores 32 lanes x 32 columns: register j of lane l -> (lane base + l, column base + j)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
// ---- CTA-pair (cta_group::2) variant: one MMA of M = 256 spans two SMs; each CTA supplies its 128 weight rows
// (A, tensor memory) and HALF of the activation tile (B, shared memory), which halves the per-SM shared-memory
// traffic of the operand that bounds the 1-CTA kernel.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {  // same offset in CTA `rank`'s smem
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait; gives up (sets *failed) instead of hanging the GPU.  Default (CTA-scope) semantics as in CUTLASS'
// ClusterBarrier: a cluster-scope acquire costs an L1 invalidate (CCTL.IVALL) per poll and is not needed, the
// data handed over lives in tensor memory / the async proxy (tcgen05 fences order it)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t a, uint32_t parity, volatile uint32_t* failed) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (!ok && (++spins > (1u << 22) || *failed)) {
            *failed = 1u;
            return;
        }
    } while (!ok);
}
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t mbar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(mbar), "h"((unsigned short)3)
        : "memory");
}
// mbarrier arrives when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile with 128-byte rows and the 128-byte swizzle: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

struct Params {
    const uint8_t* packed;
    const float* absmax;
    const void* bias;
    void* out;
    int M, N, K;
    int bs_shift;  // log2(blocksize / 64)
    uint32_t nkb;  // K / 64
    uint32_t tiles_t, num_tiles;
};

template <typename T>
struct Pack2;  // two fp32 -> packed 16-bit pair, round to nearest even (lo in bits 0..15)
template <>
struct Pack2<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t go(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};
template <>
struct Pack2<__half> {
    static __device__ __forceinline__ uint32_t go(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};

// BT: tokens per tile (MMA N).  TMEM holds two accumulators of BT columns.
template <typename T, int BT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_fp4_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                        const __grid_constant__ Params p, int stagesB) {
    constexpr uint32_t kStageB = BT * BK * 2;
    constexpr uint32_t kTmemCols = 512;           // accumulators at columns [0, kAcc * BT), A ring at the top
    constexpr uint32_t kAcc = 2 * BT <= kColA0 ? 2 : 1;  // accumulator buffers (BT <= 192: double-buffered)
    static_assert(kAcc * BT <= kColA0, "accumulators and the A ring overlap");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sB = smem_u32(smem);                       // [stagesB][BT rows][128 B] activation tiles
    const uint32_t sW = sB + (uint32_t)stagesB * kStageB;     // [kWSlots][128 rows][128 B] packed weights
    const uint32_t bars = sW + kWSlots * kWBox;
    const uint32_t fullA0 = bars, emptyA0 = fullA0 + kStagesA * 8;
    const uint32_t fullB0 = emptyA0 + kStagesA * 8, emptyB0 = fullB0 + stagesB * 8;
    const uint32_t tfull0 = emptyB0 + stagesB * 8, tempty0 = tfull0 + 16;
    const uint32_t wfull0 = tempty0 + 16, wempty0 = wfull0 + kWSlots * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)stagesB * kStageB + kWSlots * kWBox +
                                                      (kStagesA + stagesB) * 16 + 32 + kWSlots * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesA; ++s) {
            mbar_init(fullA0 + s * 8, 4);       // one arrival per dequantiser warp
            mbar_init(emptyA0 + s * 8, 1);      // tcgen05.commit
        }
        for (int s = 0; s < stagesB; ++s) {
            mbar_init(fullB0 + s * 8, 1);       // TMA producer (with tx bytes)
            mbar_init(emptyB0 + s * 8, 1);      // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + a * 8, 1);       // tcgen05.commit after the last k block
            mbar_init(tempty0 + a * 8, 128);    // epilogue threads
        }
        for (int w = 0; w < kWSlots; ++w) {
            mbar_init(wfull0 + w * 8, 1);       // TMA producer (with tx bytes)
            mbar_init(wempty0 + w * 8, 128);    // dequantiser threads
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    if (warp == 1) {  // one warp allocates TMEM and later frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t nkb = p.nkb;

    if (warp == 0) {
        // ===== TMA producer: packed weight boxes (one per 4 k blocks) and activation tiles =====
        if (lane == 0) {
            uint32_t it = 0, wit = 0;
            for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const uint32_t tt = tile % p.tiles_t, wt = tile / p.tiles_t;
                for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                    if ((kb & 3) == 0) {
                        const uint32_t ws = wit % kWSlots, wph = (wit / kWSlots) & 1;
                        ++wit;
                        mbar_wait(wempty0 + ws * 8, wph ^ 1);
                        mbar_expect_tx(wfull0 + ws * 8, kWBox);
                        tma_load_2d(sW + ws * kWBox, &tmW, (int)(kb * 32), (int)(wt * BW), wfull0 + ws * 8);
                    }
                    const uint32_t s = it % stagesB, ph = (it / stagesB) & 1;
                    mbar_wait(emptyB0 + s * 8, ph ^ 1);
                    mbar_expect_tx(fullB0 + s * 8, kStageB);
                    tma_load_2d(sB + s * kStageB, &tmX, (int)(kb * BK), (int)(tt * BT), fullB0 + s * 8);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t fmt = sizeof(T) == 2 && DT<T>::code == FP4_B200_BF16 ? 1u : 0u;
            constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BT >> 3) << 17) |
                                       ((uint32_t)(BW >> 4) << 24);
            uint32_t it = 0, tcount = 0;
            for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount % kAcc, aph = (tcount / kAcc) & 1;
                mbar_wait(tempty0 + acc * 8, aph ^ 1);  // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BT;
                for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t sa = it % kStagesA, pha = (it / kStagesA) & 1;
                    const uint32_t sb = it % stagesB, phb = (it / stagesB) & 1;
                    mbar_wait(fullA0 + sa * 8, pha);
                    mbar_wait(fullB0 + sb * 8, phb);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + kColA0 + sa * kColsA;
                    const uint64_t db = make_desc(sB + sb * kStageB);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)  // 16 elements along K: 8 TMEM columns of A, 32 bytes of the B row
                        umma_f16_ts(tmem_d, ta + (uint32_t)(k * 8), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(emptyA0 + sa * 8);  // frees the stages once these MMAs have read them
                    umma_commit(emptyB0 + sb * 8);
                }
                umma_commit(tfull0 + acc * 8);  // accumulator complete
            }
        }
    } else if (warp < 6) {
        // ===== dequantisers: thread r = weight row r of the tile =====
        const int r = (warp & 3) * 32 + lane;  // warp w may only touch TMEM lanes 32 (w % 4) .. +31
        const uint32_t srow = (uint32_t)r * 128, sx = (uint32_t)(r & 7);
        const uint32_t ta_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kColA0;
        float mag[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mag[i] = kBnbMag[i];
        uint32_t it = 0, wit = 0;
        constexpr int PD = 4;  // k blocks per packed-weight box; absmax values prefetched one box ahead
        float am[PD];
        // absmax loader runs one box (PD k blocks) ahead of the consumer, across tile boundaries
        uint32_t ld_tile = blockIdx.x, ld_kb0 = 0;
        auto load_box_absmax = [&]() {
            if (ld_tile < p.num_tiles) {
                const uint32_t wt = ld_tile / p.tiles_t;
                uint32_t row = wt * BW + (uint32_t)r;
                row = row < (uint32_t)p.N ? row : (uint32_t)p.N - 1;
                const size_t b0 = (size_t)row * nkb + ld_kb0;
#pragma unroll
                for (int i = 0; i < PD; ++i)
                    if (ld_kb0 + i < nkb) am[i] = __ldg(p.absmax + ((b0 + i) >> p.bs_shift));
                ld_kb0 += PD;
                if (ld_kb0 >= nkb) {
                    ld_kb0 = 0;
                    ld_tile += gridDim.x;
                }
            }
        };
        load_box_absmax();
        for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            for (uint32_t kb0 = 0; kb0 < nkb; kb0 += PD) {
                const uint32_t ws = wit % kWSlots, wph = (wit / kWSlots) & 1;
                ++wit;
                mbar_wait(wfull0 + ws * 8, wph);  // the packed box of k blocks kb0 .. kb0+3 has landed
                const uint32_t wsrc = sW + ws * kWBox + srow;
                float amc[PD];
#pragma unroll
                for (int i = 0; i < PD; ++i) amc[i] = am[i];
                load_box_absmax();  // absmax of the next box
#pragma unroll
                for (int i = 0; i < PD; ++i) {
                    if (kb0 + i < nkb) {
                        const uint32_t s = it % kStagesA, ph = (it / kStagesA) & 1;
                        ++it;
                        // this row's 32 packed bytes of the k block: 16-byte chunks 2i, 2i+1 of the swizzled row
                        uint4 qa, qb;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(qa.x), "=r"(qa.y), "=r"(qa.z), "=r"(qa.w)
                                     : "r"(wsrc + (((uint32_t)(2 * i) ^ sx) << 4)));
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(qb.x), "=r"(qb.y), "=r"(qb.z), "=r"(qb.w)
                                     : "r"(wsrc + (((uint32_t)(2 * i + 1) ^ sx) << 4)));
                        // the 8 magnitudes this (row, block) can take, rounded exactly like the dequant kernel
                        const float a_ = amc[i];
                        const uint32_t p01 = Pack2<T>::go(__fmul_rn(mag[0], a_), __fmul_rn(mag[1], a_));
                        const uint32_t p23 = Pack2<T>::go(__fmul_rn(mag[2], a_), __fmul_rn(mag[3], a_));
                        const uint32_t p45 = Pack2<T>::go(__fmul_rn(mag[4], a_), __fmul_rn(mag[5], a_));
                        const uint32_t p67 = Pack2<T>::go(__fmul_rn(mag[6], a_), __fmul_rn(mag[7], a_));
                        // byte tables: low bytes / high bytes of magnitudes 0..3 and 4..7
                        const uint32_t lo_a = prmt(p01, p23, 0x6420u), hi_a = prmt(p01, p23, 0x7531u);
                        const uint32_t lo_b = prmt(p45, p67, 0x6420u), hi_b = prmt(p45, p67, 0x7531u);
                        const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                        uint32_t o[32];  // the row: column j = elements (2j, 2j+1)
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            // word c = 8 nibbles = elements 8c..8c+7; nibble j of the word is element j ^ 1
                            const uint32_t ww = w[c];
                            const uint32_t wm = ww & 0x77777777u, w4 = ww * 16u;
                            const uint32_t wmh = __umulhi(wm, 65536u);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t sel = h ? wmh : wm;
                                const uint32_t lo4 = prmt(lo_a, lo_b, sel);
                                // sign-replicate mode: 0xFF where the nibble's sign bit is set
                                const uint32_t sg = prmt(ww, w4, h ? 0xBFAEu : 0x9D8Cu);
                                const uint32_t hi4 = prmt(hi_a, hi_b, sel) | (sg & 0x80808080u);
                                o[4 * c + 2 * h] = prmt(lo4, hi4, 0x4051u);      // elements (0,1) of the group: nibbles (1,0)
                                o[4 * c + 2 * h + 1] = prmt(lo4, hi4, 0x6273u);  // elements (2,3): nibbles (3,2)
                            }
                        }
                        mbar_wait(emptyA0 + s * 8, ph ^ 1);  // the MMAs that read this A stage have completed
                        tc_fence_after();
                        tmem_st_32x32(ta_lane + s * kColsA, o);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(fullA0 + s * 8);
                    }
                }
                mbar_arrive(wempty0 + ws * 8);  // this thread has read its row of the box
            }
        }
    } else {
        // ===== epilogue: thread = output feature (TMEM lane), registers = tokens =====
        const int q = warp & 3;  // TMEM lane partition this warp may access
        const int r = q * 32 + lane;
        const T* bias = reinterpret_cast<const T*>(p.bias);
        T* out = reinterpret_cast<T*>(p.out);
        uint32_t tcount = 0;
        for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t acc = tcount % kAcc, aph = (tcount / kAcc) & 1;
            const uint32_t tt = tile % p.tiles_t, wt = tile / p.tiles_t;
            const uint32_t row = wt * BW + (uint32_t)r;
            const bool row_ok = row < (uint32_t)p.N;
            const float bv = (bias && row_ok) ? DT<T>::to_f32(bias[row]) : 0.f;
            mbar_wait(tfull0 + acc * 8, aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BT;
#pragma unroll 1
            for (int c0 = 0; c0 < BT; c0 += 16) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15])
                    : "r"(taddr + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t tok0 = tt * BT + c0;
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t tok = tok0 + j;
                        if (tok < (uint32_t)p.M) out[(size_t)tok * p.N + row] = DT<T>::from_f32(__uint_as_float(v[j]) + bv);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty0 + acc * 8);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}


// CTA-pair kernel: cluster of 2; pair tile = 256 weight rows x 256 tokens.  CTA `rank` owns weight rows
// [256 wt2 + 128 rank, +128) (dequantised into ITS tensor memory) and tokens [256 tt + 128 rank, +128) (TMA into ITS
// shared memory).  The leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256, N = 256); the hardware reads A
// from both CTAs' TMEM and B from both CTAs' shared memory and leaves each CTA's 128 x 256 accumulator in its
// own TMEM.  Cross-CTA signalling: the peer's dequantiser warps arrive on the leader's fullA barriers, its
// idle "MMA" lane relays its activation-tile arrivals to the leader, its epilogue arrives on the leader's
// tempty; tcgen05.commit multicasts to both CTAs' empty / tfull barriers.
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
gemm_fp4_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                             const __grid_constant__ Params p, int stagesB, uint32_t* gfailed) {
    constexpr int BT = 256, BTH = 128;            // tokens per pair tile / per CTA
    constexpr uint32_t kStageB = BTH * BK * 2;    // 16 KiB: this CTA's half of the activation tile
    constexpr uint32_t kTmemCols = 512;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sB = smem_u32(smem);
    const uint32_t sW = sB + (uint32_t)stagesB * kStageB;
    const uint32_t bars = sW + kWSlots * kWBox;
    const uint32_t fullA0 = bars, emptyA0 = fullA0 + kStagesA * 8;
    const uint32_t fullB0 = emptyA0 + kStagesA * 8, emptyB0 = fullB0 + stagesB * 8;
    const uint32_t peerB0 = emptyB0 + stagesB * 8;  // leader only: the peer's half of stage s has landed
    const uint32_t tfull0 = peerB0 + stagesB * 8, tempty0 = tfull0 + 16;
    const uint32_t wfull0 = tempty0 + 16, wempty0 = wfull0 + kWSlots * 8;
    uint8_t* tail = smem + (size_t)stagesB * kStageB + kWSlots * kWBox + kStagesA * 16 + stagesB * 24 + 32 + kWSlots * 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail);
    volatile uint32_t* failed = reinterpret_cast<volatile uint32_t*>(tail + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) {
        *failed = 0u;
        for (int s = 0; s < kStagesA; ++s) {
            mbar_init(fullA0 + s * 8, 8);       // 4 dequantiser warps of each CTA (counted in the leader)
            mbar_init(emptyA0 + s * 8, 1);      // multicast tcgen05.commit
        }
        for (int s = 0; s < stagesB; ++s) {
            mbar_init(fullB0 + s * 8, 1);       // this CTA's TMA (with tx bytes)
            mbar_init(emptyB0 + s * 8, 1);      // multicast tcgen05.commit
            mbar_init(peerB0 + s * 8, 1);       // relay from the peer
        }
        mbar_init(tfull0, 1);                   // multicast tcgen05.commit after the last k block
        mbar_init(tempty0, 256);                // epilogue threads of both CTAs (counted in the leader)
        for (int w = 0; w < kWSlots; ++w) {
            mbar_init(wfull0 + w * 8, 1);
            mbar_init(wempty0 + w * 8, 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    if (warp == 1) {  // collective over the pair: one warp of each CTA
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();  // both CTAs' barriers are initialised before anyone signals across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t nkb = p.nkb;
    const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===== TMA producer: this CTA's packed weight boxes and its half of the activation tile =====
        if (lane == 0) {
            uint32_t it = 0, wit = 0;
            for (uint32_t tile = pair; tile < p.num_tiles && !*failed; tile += npairs) {
                const uint32_t tt = tile % p.tiles_t, wt = (tile / p.tiles_t) * 2 + rank;
                for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                    if ((kb & 3) == 0) {
                        const uint32_t ws = wit % kWSlots, wph = (wit / kWSlots) & 1;
                        ++wit;
                        mbar_wait_cluster(wempty0 + ws * 8, wph ^ 1, failed);
                        mbar_expect_tx(wfull0 + ws * 8, kWBox);
                        tma_load_2d(sW + ws * kWBox, &tmW, (int)(kb * 32), (int)(wt * BW), wfull0 + ws * 8);
                    }
                    const uint32_t s = it % stagesB, ph = (it / stagesB) & 1;
                    mbar_wait_cluster(emptyB0 + s * 8, ph ^ 1, failed);
                    mbar_expect_tx(fullB0 + s * 8, kStageB);
                    tma_load_2d(sB + s * kStageB, &tmX, (int)(kb * BK), (int)(tt * BT + rank * BTH), fullB0 + s * 8);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ===== MMA issuer (leader) =====
            constexpr uint32_t fmt = sizeof(T) == 2 && DT<T>::code == FP4_B200_BF16 ? 1u : 0u;
            constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BT >> 3) << 17) |
                                       ((uint32_t)(256 >> 4) << 24);
            uint32_t it = 0, tcount = 0;
            for (uint32_t tile = pair; tile < p.num_tiles && !*failed; tile += npairs, ++tcount) {
                mbar_wait_cluster(tempty0, (tcount & 1) ^ 1, failed);  // both epilogues have drained the accumulator
                tc_fence_after();
                for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t sa = it % kStagesA, pha = (it / kStagesA) & 1;
                    const uint32_t sb = it % stagesB, phb = (it / stagesB) & 1;
                    mbar_wait_cluster(fullA0 + sa * 8, pha, failed);
                    mbar_wait_cluster(fullB0 + sb * 8, phb, failed);
                    mbar_wait_cluster(peerB0 + sb * 8, phb, failed);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + kColA0 + sa * kColsA;
                    const uint64_t db = make_desc(sB + sb * kStageB);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16_ts_pair(tmem_base, ta + (uint32_t)(k * 8), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit_pair(emptyA0 + sa * 8);
                    umma_commit_pair(emptyB0 + sb * 8);
                }
                umma_commit_pair(tfull0);
            }
        } else if (lane == 0 && rank == 1) {
            // ===== relay (peer): tell the leader that this CTA's half of stage s has landed =====
            uint32_t it = 0;
            for (uint32_t tile = pair; tile < p.num_tiles && !*failed; tile += npairs) {
                for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t sb = it % stagesB, phb = (it / stagesB) & 1;
                    mbar_wait_cluster(fullB0 + sb * 8, phb, failed);
                    mbar_arrive_cluster(map_to_cta(peerB0 + sb * 8, 0));
                }
            }
        }
    } else if (warp < 6) {
        // ===== dequantisers (as in the 1-CTA kernel; arrivals go to the leader's fullA) =====
        const int r = (warp & 3) * 32 + lane;
        const uint32_t srow = (uint32_t)r * 128, sx = (uint32_t)(r & 7);
        const uint32_t ta_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kColA0;
        float mag[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mag[i] = kBnbMag[i];
        uint32_t it = 0, wit = 0;
        constexpr int PD = 4;
        float am[PD];
        uint32_t ld_tile = pair, ld_kb0 = 0;
        auto load_box_absmax = [&]() {
            if (ld_tile < p.num_tiles) {
                const uint32_t wt = (ld_tile / p.tiles_t) * 2 + rank;
                uint32_t row = wt * BW + (uint32_t)r;
                row = row < (uint32_t)p.N ? row : (uint32_t)p.N - 1;
                const size_t b0 = (size_t)row * nkb + ld_kb0;
#pragma unroll
                for (int i = 0; i < PD; ++i)
                    if (ld_kb0 + i < nkb) am[i] = __ldg(p.absmax + ((b0 + i) >> p.bs_shift));
                ld_kb0 += PD;
                if (ld_kb0 >= nkb) {
                    ld_kb0 = 0;
                    ld_tile += npairs;
                }
            }
        };
        load_box_absmax();
        for (uint32_t tile = pair; tile < p.num_tiles && !*failed; tile += npairs) {
            for (uint32_t kb0 = 0; kb0 < nkb; kb0 += PD) {
                const uint32_t ws = wit % kWSlots, wph = (wit / kWSlots) & 1;
                ++wit;
                mbar_wait_cluster(wfull0 + ws * 8, wph, failed);
                const uint32_t wsrc = sW + ws * kWBox + srow;
                float amc[PD];
#pragma unroll
                for (int i = 0; i < PD; ++i) amc[i] = am[i];
                load_box_absmax();
#pragma unroll
                for (int i = 0; i < PD; ++i) {
                    if (kb0 + i < nkb) {
                        const uint32_t s = it % kStagesA, ph = (it / kStagesA) & 1;
                        ++it;
                        uint4 qa, qb;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(qa.x), "=r"(qa.y), "=r"(qa.z), "=r"(qa.w)
                                     : "r"(wsrc + (((uint32_t)(2 * i) ^ sx) << 4)));
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(qb.x), "=r"(qb.y), "=r"(qb.z), "=r"(qb.w)
                                     : "r"(wsrc + (((uint32_t)(2 * i + 1) ^ sx) << 4)));
                        const float a_ = amc[i];
                        const uint32_t p01 = Pack2<T>::go(__fmul_rn(mag[0], a_), __fmul_rn(mag[1], a_));
                        const uint32_t p23 = Pack2<T>::go(__fmul_rn(mag[2], a_), __fmul_rn(mag[3], a_));
                        const uint32_t p45 = Pack2<T>::go(__fmul_rn(mag[4], a_), __fmul_rn(mag[5], a_));
                        const uint32_t p67 = Pack2<T>::go(__fmul_rn(mag[6], a_), __fmul_rn(mag[7], a_));
                        const uint32_t lo_a = prmt(p01, p23, 0x6420u), hi_a = prmt(p01, p23, 0x7531u);
                        const uint32_t lo_b = prmt(p45, p67, 0x6420u), hi_b = prmt(p45, p67, 0x7531u);
                        const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                        uint32_t o[32];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint32_t ww = w[c];
                            const uint32_t wm = ww & 0x77777777u, w4 = ww * 16u;
                            const uint32_t wmh = __umulhi(wm, 65536u);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t sel = h ? wmh : wm;
                                const uint32_t lo4 = prmt(lo_a, lo_b, sel);
                                const uint32_t sg = prmt(ww, w4, h ? 0xBFAEu : 0x9D8Cu);
                                const uint32_t hi4 = prmt(hi_a, hi_b, sel) | (sg & 0x80808080u);
                                o[4 * c + 2 * h] = prmt(lo4, hi4, 0x4051u);
                                o[4 * c + 2 * h + 1] = prmt(lo4, hi4, 0x6273u);
                            }
                        }
                        mbar_wait_cluster(emptyA0 + s * 8, ph ^ 1, failed);
                        tc_fence_after();
                        tmem_st_32x32(ta_lane + s * kColsA, o);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(map_to_cta(fullA0 + s * 8, 0));
                    }
                }
                mbar_arrive(wempty0 + ws * 8);
            }
        }
    } else {
        // ===== epilogue: this CTA's 128 weight rows x 256 tokens =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const T* bias = reinterpret_cast<const T*>(p.bias);
        T* out = reinterpret_cast<T*>(p.out);
        const uint32_t tempty_leader = map_to_cta(tempty0, 0);
        uint32_t tcount = 0;
        for (uint32_t tile = pair; tile < p.num_tiles && !*failed; tile += npairs, ++tcount) {
            const uint32_t tt = tile % p.tiles_t, wt = (tile / p.tiles_t) * 2 + rank;
            const uint32_t row = wt * BW + (uint32_t)r;
            const bool row_ok = row < (uint32_t)p.N;
            const float bv = (bias && row_ok) ? DT<T>::to_f32(bias[row]) : 0.f;
            mbar_wait_cluster(tfull0, tcount & 1, failed);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BT; c0 += 16) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15])
                    : "r"(taddr + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t tok0 = tt * BT + c0;
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t tok = tok0 + j;
                        if (tok < (uint32_t)p.M) out[(size_t)tok * p.N + row] = DT<T>::from_f32(__uint_as_float(v[j]) + bv);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(tempty_leader);
        }
    }

    if (threadIdx.x == 0 && *failed) *gfailed = 1u;
    tc_fence_before();
    __syncthreads();
    cluster_sync();  // neither CTA frees tensor memory (or exits) while the other may still use the pair
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
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

template <typename T, int BT>
static int launch(const void* x, const uint8_t* packed, const float* absmax, const void* bias, void* out, int M,
                  int N, int K, int bs_shift, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return FP4_B200_ERR_UNSUPPORTED;
    CUtensorMap tmX;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BT};
        const cuuint32_t estr[2] = {1, 1};
        const CUtensorMapDataType dt =
            DT<T>::code == FP4_B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        if (enc(&tmX, dt, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FP4_B200_ERR_UNSUPPORTED;
    }
    CUtensorMap tmW;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K / 2, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)K / 2};
        const cuuint32_t box[2] = {128, (cuuint32_t)BW};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(packed), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FP4_B200_ERR_UNSUPPORTED;
    }
    constexpr uint32_t kStageB = BT * BK * 2;
    int stagesB = (int)((220 * 1024 - kWSlots * kWBox) / kStageB);
    if (stagesB > 8) stagesB = 8;
    const size_t smem = (size_t)stagesB * kStageB + kWSlots * kWBox + (kStagesA + stagesB) * 16 + 32 + kWSlots * 16 + 64 + 1024;
    auto kern = gemm_fp4_tcgen05_kernel<T, BT>;
    static PerDeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.flag() = true;
    }
    Params p;
    p.packed = packed; p.absmax = absmax; p.bias = bias; p.out = out;
    p.M = M; p.N = N; p.K = K; p.bs_shift = bs_shift;
    p.nkb = (uint32_t)K / BK;
    p.tiles_t = ((uint32_t)M + BT - 1) / BT;
    p.num_tiles = p.tiles_t * (((uint32_t)N + BW - 1) / BW);
    const uint32_t grid = p.num_tiles < (uint32_t)kNumSMs ? p.num_tiles : (uint32_t)kNumSMs;
    kern<<<grid, kThreads, smem, st>>>(tmX, tmW, p, stagesB);
    return (int)cudaGetLastError();
}


// CTA-pair launch (M > 128 tokens): returns FP4_B200_ERR_UNSUPPORTED when switched off or not applicable
template <typename T>
static int launch_pair(const void* x, const uint8_t* packed, const float* absmax, const void* bias, void* out, int M,
                       int N, int K, int bs_shift, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return FP4_B200_ERR_UNSUPPORTED;
    CUtensorMap tmX, tmW;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BK, 128};
        const cuuint32_t estr[2] = {1, 1};
        const CUtensorMapDataType dt =
            DT<T>::code == FP4_B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        if (enc(&tmX, dt, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FP4_B200_ERR_UNSUPPORTED;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K / 2, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)K / 2};
        const cuuint32_t box[2] = {128, (cuuint32_t)BW};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(packed), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FP4_B200_ERR_UNSUPPORTED;
    }
    constexpr uint32_t kStageB = 128 * BK * 2;
    int stagesB = (int)((200 * 1024 - kWSlots * kWBox) / kStageB);
    if (stagesB > 8) stagesB = 8;
    const size_t smem = (size_t)stagesB * kStageB + kWSlots * kWBox + kStagesA * 16 + stagesB * 24 + 32 + kWSlots * 16 +
                        64 + 1024;
    auto kern = gemm_fp4_tcgen05_pair_kernel<T>;
    // per device: the function attributes and the failure flag live in the device's context
    static PerDeviceOnce configured;
    static uint32_t* gfailed_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return FP4_B200_ERR_UNSUPPORTED;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (cudaMalloc(&gfailed_dev[dev], 4) != cudaSuccess) return FP4_B200_ERR_UNSUPPORTED;
        cudaMemset(gfailed_dev[dev], 0, 4);
        configured.flag() = true;
    }
    uint32_t* gfailed = gfailed_dev[dev];
    Params p;
    p.packed = packed; p.absmax = absmax; p.bias = bias; p.out = out;
    p.M = M; p.N = N; p.K = K; p.bs_shift = bs_shift;
    p.nkb = (uint32_t)K / BK;
    p.tiles_t = ((uint32_t)M + 255) / 256;
    p.num_tiles = p.tiles_t * (((uint32_t)N + 255) / 256);
    const uint32_t pairs = p.num_tiles < (uint32_t)kNumSMs / 2 ? p.num_tiles : (uint32_t)kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pairs * 2);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, tmX, tmW, p, stagesB, gfailed);
}

template <typename T>
static int launch_bt(const void* x, const uint8_t* packed, const float* absmax, const void* bias, void* out, int M,
                     int N, int K, int bs_shift, cudaStream_t st) {
#define FP4_GO(BT) launch<T, BT>(x, packed, absmax, bias, out, M, N, K, bs_shift, st)
    if (M <= 16) return FP4_GO(16);
    if (M <= 32) return FP4_GO(32);
    if (M <= 64) return FP4_GO(64);
    if (M <= 128) return FP4_GO(128);
    static const int pair_min = getenv("FP4_B200_GEMM_PAIR_MIN_M") ? atoi(getenv("FP4_B200_GEMM_PAIR_MIN_M")) : (1 << 30);
    if (M >= pair_min) return launch_pair<T>(x, packed, absmax, bias, out, M, N, K, bs_shift, st);
    // 256-token tiles fill TMEM with ONE accumulator (the epilogue is not overlapped); 192-token tiles with two
    // accumulators measured slower (padding + smaller MMAs), so they are not used
    return FP4_GO(256);
#undef FP4_GO
}

}  // namespace

int gemm_tcgen05_dispatch(const void* x, const uint8_t* packed, const float* absmax, const float* code,
                          const void* bias, void* out, int M, int N, int K, int blocksize, int dtype, unsigned flags,
                          cudaStream_t st) {
    if (code != nullptr && !(flags & FP4_B200_FLAG_CODE_IS_BNB_FP4)) return FP4_B200_ERR_UNSUPPORTED;
    if (K % BK != 0 || blocksize % 64 != 0) return FP4_B200_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(packed) % 16 || (K * 2) % 16)
        return FP4_B200_ERR_ALIGN;
    const int bs_shift = ilog2_exact(blocksize / 64);
    if (bs_shift < 0) return FP4_B200_ERR_BLOCKSIZE;
    switch (dtype) {
        case FP4_B200_BF16:
            return launch_bt<__nv_bfloat16>(x, packed, absmax, bias, out, M, N, K, bs_shift, st);
        case FP4_B200_F16:
            return launch_bt<__half>(x, packed, absmax, bias, out, M, N, K, bs_shift, st);
        default:
            return FP4_B200_ERR_DTYPE;
    }
}

}  // namespace fp4b200
