// Blockwise FP4 -> fp16/bf16/fp32 dequantisation for sm_100a.
//
// Replaces the reference's dequantize_blockwise_kernel_fp4 / dequantize_blockwise_codebook_kernel_fp4
// (reference csrc/dequant_fp4_optimized.cu:89-171).  Numerical definition:
//     out[i] = RN_T( fp32_mul( code[nibble_i], absmax[i / blocksize] ) )
// with element 2j = high nibble of byte j (reference :117-118).  One IEEE fp32 multiply (__fmul_rn,
// no FTZ, no FMA contraction), then round-to-nearest-even to T.  Bit-exact by construction.
//
// HBM-bound: per element 0.5 B packed + 4/blocksize B absmax read, sizeof(T) written.
// Layout of the work: one thread owns 32 bytes of OUTPUT (16 halves / 8 floats), so a warp reads
// 256 (128) contiguous packed bytes with one 64-bit (32-bit) load per lane and writes 1 KiB
// contiguous with a single 256-bit store per lane (STG.E.ENL2.256) - full 128-byte lines in both
// directions, no shared-memory transpose (the reference goes through cub WARP_TRANSPOSE with byte
// loads and 16-bit stores).  The 16-entry codebook sits in shared memory; its 16 words occupy 16
// different banks, so the per-nibble lookup is conflict-free for any nibble pattern.
#include "common.cuh"

namespace fp4b200 {

template <typename T>
struct Pack32B;  // 32 bytes of output per thread

template <typename T, bool NESTED, int U>
__global__ void __launch_bounds__(256)
dequant_vec_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ absmax,
                   const NestedDev nd, const float* __restrict__ code,
                   const __grid_constant__ Code16 dflt, T* __restrict__ out, const int64_t n,
                   const int bs_log2) {
    constexpr int EPT = 32 / (int)sizeof(T);  // elements per thread-chunk: 16 (16-bit) or 8 (fp32)
    constexpr int WPT = EPT / 8;              // packed 32-bit words per chunk: 2 or 1
    __shared__ float s_code[16];
    if (threadIdx.x < 16)
        s_code[threadIdx.x] = code ? __ldg(code + threadIdx.x) : dflt.v[threadIdx.x];
    __syncthreads();

    const int64_t nfull = n / EPT;
    const int64_t base = (int64_t)blockIdx.x * (256 * U) + threadIdx.x;

    uint32_t q[U][WPT];
    float am[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t c = base + (int64_t)u * 256;
        if (c < nfull) {
            if constexpr (WPT == 2) {
                const uint2 v = ldg_stream_u2(packed + c * 8);
                q[u][0] = v.x;
                q[u][1] = v.y;
            } else {
                q[u][0] = __ldg(reinterpret_cast<const uint32_t*>(packed + c * 4));
            }
            am[u] = load_absmax<NESTED>(absmax, nd, (c * EPT) >> bs_log2);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t c = base + (int64_t)u * 256;
        if (c < nfull) {
            uint32_t r[8];
#pragma unroll
            for (int w = 0; w < WPT; ++w) {
                const uint32_t word = q[u][w];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t byte = (word >> (8 * b)) & 0xFFu;
                    const float hi = __fmul_rn(s_code[byte >> 4], am[u]);   // element 2j
                    const float lo = __fmul_rn(s_code[byte & 0xFu], am[u]); // element 2j+1
                    if constexpr (sizeof(T) == 2) {
                        r[w * 4 + b] = DT<T>::pack2(hi, lo);
                    } else {
                        r[b * 2] = __float_as_uint(hi);
                        r[b * 2 + 1] = __float_as_uint(lo);
                    }
                }
            }
            stg_u8x32(out + c * EPT, r);
        }
    }
    // ragged tail (< EPT elements): one thread, scalar
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        for (int64_t i = nfull * EPT; i < n; ++i) {
            const uint8_t byte = packed[i >> 1];
            const uint32_t nib = (i & 1) ? (byte & 0xFu) : (byte >> 4);
            const float a = load_absmax<NESTED>(absmax, nd, i >> bs_log2);
            out[i] = DT<T>::from_f32(__fmul_rn(s_code[nib], a));
        }
    }
}

// Fallback for blocksize < 16 or unaligned pointers: one thread per packed byte.
template <typename T, bool NESTED>
__global__ void __launch_bounds__(256)
dequant_scalar_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ absmax,
                      const NestedDev nd, const float* __restrict__ code,
                      const __grid_constant__ Code16 dflt, T* __restrict__ out, const int64_t n,
                      const int bs_log2) {
    __shared__ float s_code[16];
    if (threadIdx.x < 16)
        s_code[threadIdx.x] = code ? __ldg(code + threadIdx.x) : dflt.v[threadIdx.x];
    __syncthreads();
    const int64_t nbytes = (n + 1) >> 1;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < nbytes;
         j += (int64_t)gridDim.x * 256) {
        const uint8_t byte = packed[j];
        const int64_t i0 = 2 * j;
        out[i0] = DT<T>::from_f32(
            __fmul_rn(s_code[byte >> 4], load_absmax<NESTED>(absmax, nd, i0 >> bs_log2)));
        if (i0 + 1 < n)
            out[i0 + 1] = DT<T>::from_f32(__fmul_rn(
                s_code[byte & 0xFu], load_absmax<NESTED>(absmax, nd, (i0 + 1) >> bs_log2)));
    }
}

template <typename T, bool NESTED>
static int launch_dequant(const uint8_t* packed, const float* absmax, const NestedDev& nd,
                          const float* code, T* out, int64_t n, int bs_log2, cudaStream_t st) {
    const Code16 dflt = {FP4_B200_BNB_CODE_INIT};
    constexpr int EPT = 32 / (int)sizeof(T);
    constexpr int U = 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(out) % 32 == 0) &&
                         (reinterpret_cast<uintptr_t>(packed) % (EPT / 2) == 0);
    if ((1 << bs_log2) >= EPT && aligned) {
        const int64_t nchunks = (n + EPT - 1) / EPT;
        const int64_t blocks = (nchunks + 256 * U - 1) / (256 * U);
        dequant_vec_kernel<T, NESTED, U><<<(unsigned)blocks, 256, 0, st>>>(
            packed, absmax, nd, code, dflt, out, n, bs_log2);
    } else {
        const int64_t nbytes = (n + 1) / 2;
        int64_t blocks = (nbytes + 255) / 256;
        if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
        dequant_scalar_kernel<T, NESTED><<<(unsigned)blocks, 256, 0, st>>>(
            packed, absmax, nd, code, dflt, out, n, bs_log2);
    }
    return (int)cudaGetLastError();
}

int dequant_dispatch(const uint8_t* packed, const float* absmax, const fp4_b200_nested_t* nested,
                     const float* code, void* out, int64_t n, int blocksize, int out_dtype,
                     cudaStream_t st) {
    if (!packed || !out) return FP4_B200_ERR_NULL;
    if (!nested && !absmax) return FP4_B200_ERR_NULL;
    if (n < 0) return FP4_B200_ERR_SHAPE;
    const int bs_log2 = ilog2_exact(blocksize);
    if (bs_log2 < 1) return FP4_B200_ERR_BLOCKSIZE;
    if (n == 0) return FP4_B200_OK;
    NestedDev nd = {};
    if (nested) {
        if (!nested->qabsmax || !nested->code2 || !nested->absmax2) return FP4_B200_ERR_NULL;
        const int l2 = ilog2_exact(nested->blocksize2);
        if (l2 < 0) return FP4_B200_ERR_BLOCKSIZE;
        nd = NestedDev{nested->qabsmax, nested->code2, nested->absmax2, nested->offset, l2};
    }
#define FP4_DISPATCH(T)                                                                        \
    (nested ? launch_dequant<T, true>(packed, absmax, nd, code, (T*)out, n, bs_log2, st)       \
            : launch_dequant<T, false>(packed, absmax, nd, code, (T*)out, n, bs_log2, st))
    switch (out_dtype) {
        case FP4_B200_F16: return FP4_DISPATCH(__half);
        case FP4_B200_BF16: return FP4_DISPATCH(__nv_bfloat16);
        case FP4_B200_F32: return FP4_DISPATCH(float);
        default: return FP4_B200_ERR_DTYPE;
    }
#undef FP4_DISPATCH
}

// ---- nested absmax -> fp32 (load-time helper) -------------------------------------------------
__global__ void __launch_bounds__(256)
denest_kernel(const NestedDev nd, float* __restrict__ out, const int64_t nblocks) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nblocks;
         i += (int64_t)gridDim.x * 256)
        out[i] = nested_absmax(nd, i);
}

int denest_dispatch(const fp4_b200_nested_t* nested, float* out, int64_t nblocks,
                    cudaStream_t st) {
    if (!nested || !out || !nested->qabsmax || !nested->code2 || !nested->absmax2)
        return FP4_B200_ERR_NULL;
    if (nblocks < 0) return FP4_B200_ERR_SHAPE;
    const int l2 = ilog2_exact(nested->blocksize2);
    if (l2 < 0) return FP4_B200_ERR_BLOCKSIZE;
    if (nblocks == 0) return FP4_B200_OK;
    const NestedDev nd{nested->qabsmax, nested->code2, nested->absmax2, nested->offset, l2};
    int64_t blocks = (nblocks + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    denest_kernel<<<(unsigned)blocks, 256, 0, st>>>(nd, out, nblocks);
    return (int)cudaGetLastError();
}

}  // namespace fp4b200
