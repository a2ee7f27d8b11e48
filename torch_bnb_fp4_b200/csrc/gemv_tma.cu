// Fused dequant + GEMV for decode (batch 1..8) on sm_100a — TMA-staged stream-K variant (fp16 tensor-core decode).
// Superseded as the default by gemv_stream.cu; kept as a fallback for shapes outside that kernel's domain.
//
// Same arithmetic and the same warp-granular stream-K schedule / deterministic combine as gemv_imma.cu
// (see there and gemv_common.cuh); what changes is how the weight stream reaches the SM.
// gemv_imma.cu loads the MMA A fragments straight from global memory, which makes every warp-level load
// touch 8 different rows: 8 L1 wavefronts for 256 bytes, 32-byte DRAM bursts per row, 32 registers of
// load buffers, and a ~2500-cycle ramp just to get the first loads through the LSU.  Here every WARP
// owns a small shared-memory ring and lane 0 feeds it with cp.async.bulk.tensor (TMA):
//   * one box = [16 rows x 128 B] of packed weights = 4 units (256 k per row, whole 128-byte lines,
//     hardware 128-byte swizzle so the fragment reads below are nearly conflict free) plus the
//     matching [16 rows x 4] fp32 absmax box, completion on an mbarrier;
//   * kStages boxes in flight per warp (8 KiB), 16 warps per SM -> 128 KiB in flight per SM, no LSU
//     wavefronts and no registers spent on it; the ring starts filling BEFORE griddepcontrol.wait, so
//     under programmatic dependent launch the weight stream of layer i+1 overlaps the tail of layer i
//     (weights never depend on the previous kernel; x, workspace and out do);
//   * the warp that consumed a box refills it itself (program order: no "empty" barriers at all).
// Boxes may start at any k block, so the unit-granular balanced partition of gemv_imma.cu is kept;
// a box is simply cut short at the end of a row tile / of the warp's range.
//
// Requirements (gemv_tma_supported): bitsandbytes FP4 codebook, blocksize 64, fp32 absmax (not nested),
// K % 256 == 0 (16-byte absmax row pitch), N % 16 == 0.  Everything else takes gemv_imma.cu /
// gemv_generic.cu.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "gemv_common.cuh"

namespace fp4b200 {

namespace {

using namespace gemv;

constexpr int kThreads = kWarps * 32;
constexpr int kStages = 4;
constexpr int kBoxUnits = 4;                       // units (64-wide k blocks) per TMA box
constexpr uint32_t kWBox = 16 * 128;               // 2 KiB of packed weights
constexpr uint32_t kABox = 16 * 32;                // 512 B: 8 fp32 absmax for each of 16 rows (TMA box
                                                   // origins are 16-byte granular: fetch from kb & ~3)
constexpr uint32_t kStageBytes = kWBox + kABox;    // bytes one stage's two TMA boxes deliver
// dynamic smem: [kWarps][kStages] weight boxes (2 KiB each, so every box is 1024-byte aligned as the
// 128-byte swizzle pattern assumes), then [kWarps][kStages] absmax boxes, then the x fragments
constexpr uint32_t kRingW = kWarps * kStages * kWBox;
constexpr uint32_t kRingA = kWarps * kStages * kABox;

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes)
                 : "memory");
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
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(saddr));
    return r;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1,
                                            uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(mbar)
        : "memory");
}

template <typename T, int NCOLT>
__global__ void __launch_bounds__(kThreads, 2)
gemv_tma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                const T* __restrict__ x, const T* __restrict__ bias, T* __restrict__ out,
                const Workspace ws, const int batch, const int N, const int K, const Partition part,
                const FastDiv by_nkb) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    constexpr int NC = 8 * NCOLT;
    // dynamic smem: per-warp rings (1024-aligned for the 128-byte swizzle) | x fragments
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[kWarps * kStages];
    __shared__ float sPart[kWarps * 2 * 16 * NC];  // CTA-local partial sums of shared row tiles
    __shared__ unsigned sCnt[kWarps];
    __shared__ float sMax[kWarps * 8];
    __shared__ float sScale[8];
    const uint32_t nkb = (uint32_t)K >> 6;
    const int ncols = batch * PIECES;
    uint4* sB = reinterpret_cast<uint4*>(smem_raw + kRingW + kRingA);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t g = lane >> 2, t = lane & 3;
    const uint32_t wid = blockIdx.x * kWarps + warp;
    const uint32_t L0 = part.begin(wid);
    const uint32_t n = part.begin(wid + 1) - L0;
    if (tid < kWarps) sCnt[tid] = 0;

    const uint32_t ringW = (uint32_t)__cvta_generic_to_shared(smem_raw) + warp * kStages * kWBox;
    const uint32_t ringA = (uint32_t)__cvta_generic_to_shared(smem_raw) + kRingW + warp * kStages * kABox;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars) + warp * kStages * 8;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // ---- box cursor: the warp's units in order, cut into boxes of <= 4 units inside one row tile ----
    uint32_t tile0, kb0;
    by_nkb.divmod(n ? L0 : 0, tile0, kb0);
    uint32_t ld_left = n, ld_tile = tile0, ld_kb = kb0;
    auto issue_box = [&](uint32_t slot) {  // executed by the whole warp, lane 0 talks to the TMA
        if (ld_left == 0) return;
        uint32_t cnt = nkb - ld_kb;
        cnt = cnt < (uint32_t)kBoxUnits ? cnt : (uint32_t)kBoxUnits;
        cnt = cnt < ld_left ? cnt : ld_left;
        if (lane == 0) {
            const uint32_t bar = bar0 + slot * 8;
            mbar_expect_tx(bar, kStageBytes);
            tma_load_2d(ringW + slot * kWBox, &tmW, (int)(ld_kb * 32), (int)(ld_tile * 16), bar);
            tma_load_2d(ringA + slot * kABox, &tmA, (int)(ld_kb & ~3u), (int)(ld_tile * 16), bar);
        }
        ld_left -= cnt;
        if ((ld_kb += cnt) == nkb) {
            ld_kb = 0;
            ++ld_tile;
        }
    };
    // ---- 1. fill the ring: weights never depend on the previous kernel in the stream --------------
    if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    }
#pragma unroll
    for (int s = 0; s < kStages; ++s) issue_box(s);

    // x, the workspace and `out` may be products of the previous kernel: wait for it, then let the
    // next kernel start its own prologue
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");

    // ---- 2. stage x as scaled fp16 in B-fragment order (once per CTA) --------------------------
    const int nchunk = K >> 3;
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
    if (n == 0) return;

    // ---- 3. main loop over this warp's boxes -----------------------------------------------------
    FlushCtx<T, NC> fc;
    fc.sScale = sScale; fc.sPart = sPart; fc.sCnt = sCnt; fc.bias = bias; fc.out = out; fc.ws = ws;
    fc.part = part; fc.nkb = nkb; fc.wid = wid; fc.first_tile = tile0; fc.by_nkb = by_nkb;
    fc.cta_w0 = blockIdx.x * kWarps;
    fc.cta_L0 = part.begin(fc.cta_w0); fc.cta_L1 = part.begin(fc.cta_w0 + kWarps);
    fc.batch = batch; fc.N = N;

    // TMA SWIZZLE_128B: 16-byte chunk c of box row r lands at chunk c ^ (r & 7).  Unit u of this lane is
    // bytes [32u + 8t, +8) of row g (and of row g + 8, 1 KiB further: (g + 8) & 7 == g).
    uint32_t woff[kBoxUnits];
#pragma unroll
    for (int u = 0; u < kBoxUnits; ++u)
        woff[u] = g * 128 + (((2 * u + (t >> 1)) ^ g) << 4) + 8 * (t & 1);
    const uint32_t aoff = g * 32;  // this lane's rows' absmax: 8 floats of row g, row g+8 at +256

    const uint32_t bstep = 2 * ncols * 4 * 16;  // bytes of x fragments per k block
    const uint32_t bhalf = ncols * 4 * 16;
    uint32_t bcol[NCOLT];
#pragma unroll
    for (int ct = 0; ct < NCOLT; ++ct) {
        const int col = ct * 8 + (int)g;
        bcol[ct] = (col < ncols ? col : 0) * 64;  // lanes without a column read column 0 (never stored)
    }
    const uint32_t sB_base = (uint32_t)__cvta_generic_to_shared(sB) + t * 16;
    uint32_t tab_lo;
    asm volatile("mov.b32 %0, 0x4A482C00;" : "=r"(tab_lo));

    AccV<NCOLT> acc;
#pragma unroll
    for (int c = 0; c < NCOLT; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc.v[c][i] = 0.f;

    uint32_t tile = tile0, kb = kb0, seg_start_kb = kb0;
    uint32_t left = n, slot = 0, phase = 0;
    while (left) {
        uint32_t cnt = nkb - kb;
        cnt = cnt < (uint32_t)kBoxUnits ? cnt : (uint32_t)kBoxUnits;
        cnt = cnt < left ? cnt : left;
        const uint32_t sbase = ringW + slot * kWBox, abase = ringA + slot * kABox;
        mbar_wait(bar0 + slot * 8, phase);
        // absmax of unit u: float (kb & 3) + u of this lane's rows in the 8-wide box
        const uint32_t am_addr = abase + aoff + (kb & 3u) * 4;
        uint32_t bsaddr = sB_base + kb * bstep;
#pragma unroll
        for (int u = 0; u < kBoxUnits; ++u) {
            if ((uint32_t)u < cnt) {
                const uint2 w0 = lds_u2(sbase + woff[u]);
                const uint2 w1 = lds_u2(sbase + woff[u] + 8 * 128);
                const float am0 = lds_f32(am_addr + 4 * u), am1 = lds_f32(am_addr + 8 * 32 + 4 * u);
                uint32_t ha[2][4], hb[2][4];  // [word][half2] for rows g / g+8
                decode_word(w0.x, tab_lo, ha[0]);
                decode_word(w0.y, tab_lo, ha[1]);
                decode_word(w1.x, tab_lo, hb[0]);
                decode_word(w1.y, tab_lo, hb[1]);
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
            }
        }
        // every lane has its fragments in registers (the MMAs above consumed them): refill the slot
        __syncwarp();
        issue_box(slot);
        if (++slot == kStages) {
            slot = 0;
            phase ^= 1;
        }
        left -= cnt;
        kb += cnt;
        if (kb == nkb || left == 0) {
            flush_tile<T, NCOLT>(fc, acc, tile, seg_start_kb, kb);
#pragma unroll
            for (int c = 0; c < NCOLT; ++c)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc.v[c][i] = 0.f;
            if (kb == nkb) {
                kb = 0;
                ++tile;
            }
            seg_start_kb = kb;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

static size_t smem_bytes(int batch, int K, int pieces) {
    return (size_t)kRingW + kRingA + (size_t)(K / 64) * 2 * (batch * pieces) * 4 * 16;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <typename T, int NCOLT>
static int launch(const void* x, const uint8_t* packed, const float* absmax, const void* bias,
                  void* out, void* workspace, size_t workspace_bytes, int batch, int N, int K,
                  cudaStream_t st) {
    constexpr int PIECES = (sizeof(T) == 4) ? 2 : 1;
    constexpr int NC = 8 * NCOLT;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return FP4_B200_ERR_UNSUPPORTED;
    auto kern = gemv_tma_kernel<T, NCOLT>;
    const size_t smem = smem_bytes(batch, K, PIECES);
    static bool configured = false;
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return FP4_B200_ERR_UNSUPPORTED;
    if (occ > 4) occ = 4;  // the workspace is sized for <= 4 CTAs per SM
    static const int ctas_per_sm = env_int("FP4_B200_GEMV_CTAS_PER_SM", 0);
    static const int min_units = env_int("FP4_B200_GEMV_MIN_UNITS", 8);
    if (ctas_per_sm > 0 && ctas_per_sm < occ) occ = ctas_per_sm;

    const int64_t units = (int64_t)(N / 16) * (K / 64);
    int64_t grid = (int64_t)kNumSMs * occ;
    const int64_t max_by_work = (units + (int64_t)kWarps * min_units - 1) / ((int64_t)kWarps * min_units);
    if (grid > max_by_work) grid = max_by_work;
    if (grid < 1) grid = 1;

    const size_t need = kCounterBytes + (size_t)grid * kWarps * 2 * 16 * NC * 4;
    if (!workspace || workspace_bytes < need) return FP4_B200_ERR_WORKSPACE;
    Workspace ws;
    ws.counters = reinterpret_cast<unsigned*>(workspace);
    ws.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCounterBytes);

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
        const cuuint32_t box[2] = {8, 16};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(absmax), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FP4_B200_ERR_UNSUPPORTED;
    }
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
    return (int)cudaLaunchKernelEx(&cfg, kern, tmW, tmA, (const T*)x, (const T*)bias, (T*)out, ws, batch, N,
                                   K, part, FastDiv((uint32_t)(K / 64)));
}

}  // namespace

bool gemv_tma_supported(int batch, int N, int K, int blocksize, int dtype, bool nested,
                        const void* packed, const void* absmax) {
    static const int disabled = env_int("FP4_B200_GEMV_NO_TMA", 0);
    if (disabled || nested || blocksize != 64) return false;
    if (batch < 1 || batch > 8 || N <= 0 || K <= 0) return false;
    if (K % 256 != 0 || N % 16 != 0) return false;
    if ((size_t)(N / 16) * 4 > gemv::kCounterBytes) return false;
    if ((int64_t)N * K / 64 >= (int64_t)1 << 31) return false;
    if (reinterpret_cast<uintptr_t>(packed) % 16 || reinterpret_cast<uintptr_t>(absmax) % 16) return false;
    const int pieces = dtype == FP4_B200_F32 ? 2 : 1;
    return smem_bytes(batch, K, pieces) <= 200 * 1024 && encode_fn() != nullptr;
}

int gemv_tma_dispatch(const void* x, const uint8_t* packed, const float* absmax, const void* bias,
                      void* out, void* workspace, size_t workspace_bytes, int batch, int N, int K,
                      int dtype, cudaStream_t st) {
#define FP4_GO(T)                                                                                     \
    ((batch * ((sizeof(T) == 4) ? 2 : 1) > 8)                                                         \
         ? launch<T, 2>(x, packed, absmax, bias, out, workspace, workspace_bytes, batch, N, K, st)   \
         : launch<T, 1>(x, packed, absmax, bias, out, workspace, workspace_bytes, batch, N, K, st))
    switch (dtype) {
        case FP4_B200_F16: return FP4_GO(__half);
        case FP4_B200_BF16: return FP4_GO(__nv_bfloat16);
        case FP4_B200_F32: return FP4_GO(float);
        default: return FP4_B200_ERR_DTYPE;
    }
#undef FP4_GO
}

}  // namespace fp4b200
