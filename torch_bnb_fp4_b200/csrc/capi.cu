// extern "C" surface of libfp4_b200.so (declared in include/fp4_b200.h): argument validation and
// kernel selection only; the kernels live in dequant.cu, gemv_stream.cu, gemv_generic.cu, gemm_tcgen05.cu, quantize.cu.
#include <atomic>
#include <new>

#include "common.cuh"

namespace fp4b200 {
int dequant_dispatch(const uint8_t*, const float*, const fp4_b200_nested_t*, const float*, void*,
                     int64_t, int, int, cudaStream_t);
int denest_dispatch(const fp4_b200_nested_t*, float*, int64_t, cudaStream_t);
int quantize_dispatch(const void*, int, int64_t, int, uint8_t*, float*, cudaStream_t);
int gemv_generic_dispatch(const void*, const uint8_t*, const float*, const fp4_b200_nested_t*,
                          const NestedDev&, const float*, const void*, void*, int, int, int, int,
                          int, cudaStream_t);
bool gemv_stream_supported(int batch, int N, int K, int blocksize, int dtype, bool nested,
                           const void* packed, const void* absmax);
int gemv_stream_dispatch(const void*, const uint8_t*, const float*, const fp4_b200_nested_t*, const void*, void*, int,
                         int, int, int, cudaStream_t);
bool gemv_stream_group_supported(int nmat, int batch, const int* N, int K, int blocksize, int dtype,
                                 const uint8_t* const* packed, const float* const* absmax);
int gemv_stream_group_dispatch(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                               const void* const* bias, void* const* out, const int* N, int batch, int K, int dtype,
                               const fp4_b200_tp_t* tp, const fp4_b200_epilogue_t* epi, cudaStream_t st);
int gemm_tcgen05_dispatch(const void*, const uint8_t*, const float*, const float*, const void*,
                          void*, int, int, int, int, int, unsigned, cudaStream_t);
}  // namespace fp4b200

using namespace fp4b200;

// kernels launched by this library in this process (every entry point below launches exactly one kernel when it
// returns 0): lets a benchmark report its launch count from the library's own bookkeeping
static std::atomic<unsigned long long> g_launches{0};
static inline int counted(int rc) {
    if (rc == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
    return rc;
}

extern "C" {

unsigned long long fp4_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int fp4_b200_abi_version(void) { return FP4_B200_ABI_VERSION; }

const char* fp4_b200_status_string(int s) {
    switch (s) {
        case FP4_B200_OK: return "ok";
        case FP4_B200_ERR_NULL: return "a required pointer is NULL";
        case FP4_B200_ERR_DTYPE: return "unsupported dtype (expected float16, float32 or bfloat16)";
        case FP4_B200_ERR_SHAPE: return "invalid shape";
        case FP4_B200_ERR_BLOCKSIZE: return "blocksize must be a power of two";
        case FP4_B200_ERR_ALIGN: return "pointer or row pitch is not sufficiently aligned";
        case FP4_B200_ERR_BATCH: return "gemv batch must be in 1..8";
        case FP4_B200_ERR_UNSUPPORTED: return "shape not supported by this entry point";
        case FP4_B200_ERR_WORKSPACE: return "workspace missing or too small";
        default: return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown status";
    }
}

int fp4_b200_dequantize(const uint8_t* packed, const float* absmax, const float* code, void* out,
                        int64_t n, int blocksize, int out_dtype, void* stream) {
    if (!absmax) return FP4_B200_ERR_NULL;
    return counted(dequant_dispatch(packed, absmax, nullptr, code, out, n, blocksize, out_dtype,
                            (cudaStream_t)stream));
}

int fp4_b200_dequantize_nested(const uint8_t* packed, const fp4_b200_nested_t* nested,
                               const float* code, void* out, int64_t n, int blocksize,
                               int out_dtype, void* stream) {
    if (!nested) return FP4_B200_ERR_NULL;
    return counted(dequant_dispatch(packed, nullptr, nested, code, out, n, blocksize, out_dtype,
                            (cudaStream_t)stream));
}

int fp4_b200_absmax_denest(const fp4_b200_nested_t* nested, float* absmax_out, int64_t nblocks,
                           void* stream) {
    return counted(denest_dispatch(nested, absmax_out, nblocks, (cudaStream_t)stream));
}

int fp4_b200_gemv(const void* x, const uint8_t* packed, const float* absmax,
                  const fp4_b200_nested_t* nested, const float* code, const void* bias, void* out,
                  int batch, int N, int K, int blocksize, int dtype, unsigned flags,
                  void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !packed || !out) return FP4_B200_ERR_NULL;
    if (!nested && !absmax) return FP4_B200_ERR_NULL;
    if (batch < 1 || batch > 8) return FP4_B200_ERR_BATCH;
    if (N < 0 || K < 0) return FP4_B200_ERR_SHAPE;
    if (dtype != FP4_B200_F16 && dtype != FP4_B200_BF16 && dtype != FP4_B200_F32)
        return FP4_B200_ERR_DTYPE;
    const int bs_log2 = ilog2_exact(blocksize);
    if (bs_log2 < 0) return FP4_B200_ERR_BLOCKSIZE;
    if (blocksize % 32 != 0) return FP4_B200_ERR_UNSUPPORTED;
    if (K % 32 != 0) return FP4_B200_ERR_UNSUPPORTED;  // 16-byte row chunks, as the reference
    if (N == 0) return FP4_B200_OK;
    if (K == 0) return FP4_B200_ERR_SHAPE;
    if (reinterpret_cast<uintptr_t>(packed) % 16 || reinterpret_cast<uintptr_t>(x) % 16)
        return FP4_B200_ERR_ALIGN;
    NestedDev nd = {};
    if (nested) {
        if (!nested->qabsmax || !nested->code2 || !nested->absmax2) return FP4_B200_ERR_NULL;
        const int l2 = ilog2_exact(nested->blocksize2);
        if (l2 < 0) return FP4_B200_ERR_BLOCKSIZE;
        nd = NestedDev{nested->qabsmax, nested->code2, nested->absmax2, nested->offset, l2};
    }
    (void)workspace;
    (void)workspace_bytes;
    const bool std_code = (code == nullptr) || (flags & FP4_B200_FLAG_CODE_IS_BNB_FP4);
    if (std_code && !(flags & (FP4_B200_FLAG_FORCE_GENERIC | FP4_B200_FLAG_NO_STREAM))) {
        const void* aq = nested ? (const void*)nested->qabsmax : (const void*)absmax;
        // the streaming integer tensor-core kernel (bitsandbytes table, blocksize 64, K % 256 == 0, N % 16 == 0)
        if (gemv_stream_supported(batch, N, K, blocksize, dtype, nested != nullptr, packed, aq))
            return counted(gemv_stream_dispatch(x, packed, absmax, nested, bias, out, batch, N, K, dtype,
                                                (cudaStream_t)stream));
        // rows of x (as integer terms) that do not fit its shared memory together - fp32 inputs with batch 5..8 on
        // K = 8192, 16-bit batch 8 on K = 14336 ... - are done in two launches: the weights stream twice, which is
        // still an order of magnitude faster than the generic kernel
        if (batch >= 2) {
            const int b0 = (batch + 1) / 2, b1 = batch - b0;
            if (gemv_stream_supported(b0, N, K, blocksize, dtype, nested != nullptr, packed, aq)) {
                const size_t es = dtype == FP4_B200_F32 ? 4 : 2;
                const int rc = counted(gemv_stream_dispatch(x, packed, absmax, nested, bias, out, b0, N, K, dtype,
                                                            (cudaStream_t)stream));
                if (rc) return rc;
                return counted(gemv_stream_dispatch(static_cast<const uint8_t*>(x) + (size_t)b0 * K * es, packed,
                                                    absmax, nested, bias,
                                                    static_cast<uint8_t*>(out) + (size_t)b0 * N * es, b1, N, K,
                                                    dtype, (cudaStream_t)stream));
            }
        }
    }
    // any codebook, any block size (>= 32), K % 32 == 0: CUDA-core kernel
    return counted(gemv_generic_dispatch(x, packed, absmax, nested, nd, code, bias, out, batch, N, K,
                                 bs_log2, dtype, (cudaStream_t)stream));
}

int fp4_b200_gemv_grouped_ex(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                             const void* const* bias, void* const* out, const int* N, int batch, int K,
                             int blocksize, int dtype, unsigned flags, const fp4_b200_tp_t* tp,
                             const fp4_b200_epilogue_t* epi, void* stream) {
    const bool peer_x = tp && tp->in_world > 1;
    if ((!x && !peer_x) || !packed || !absmax || !out || !N) return FP4_B200_ERR_NULL;
    if (nmat < 1 || nmat > 4) return FP4_B200_ERR_SHAPE;
    if (batch < 1 || batch > 8) return FP4_B200_ERR_BATCH;
    if (dtype != FP4_B200_F16 && dtype != FP4_B200_BF16 && dtype != FP4_B200_F32) return FP4_B200_ERR_DTYPE;
    if (!(flags & FP4_B200_FLAG_CODE_IS_BNB_FP4)) return FP4_B200_ERR_UNSUPPORTED;  // bitsandbytes table only
    const bool gated = epi && epi->gate_act;
    for (int m = 0; m < nmat; ++m) {
        const bool nested_m = epi && epi->nested && epi->nested[m];
        if (!packed[m] || (!absmax[m] && !nested_m) || (!out[m] && !(tp && tp->out_world > 1) && !(gated && m == 1)))
            return FP4_B200_ERR_NULL;
    }
    if (x && reinterpret_cast<uintptr_t>(x) % 16) return FP4_B200_ERR_ALIGN;
    if (tp) {
        if (tp->in_world < 0 || tp->in_world > 8 || tp->out_world < 0 || tp->out_world > 8) return FP4_B200_ERR_SHAPE;
        const bool on = tp->in_world > 1 || tp->out_world > 1;
        if (on && (!tp->epochs || !tp->err || tp->slot_bytes % 16)) return FP4_B200_ERR_NULL;
        if (tp->in_world > 1 && !tp->in_base) return FP4_B200_ERR_NULL;
        if (tp->out_world > 1 && nmat != 1) return FP4_B200_ERR_SHAPE;
        if (on && dtype == FP4_B200_F32) return FP4_B200_ERR_UNSUPPORTED;  // words carry two 16-bit values
        if (tp->in_world > 1 && (size_t)batch * K * 4 > tp->slot_bytes) return FP4_B200_ERR_SHAPE;  // 8-byte words, 2 rows each
        if (tp->out_world > 1 && (size_t)batch * N[0] * 4 > tp->slot_bytes) return FP4_B200_ERR_SHAPE;
    }
    if (!gemv_stream_group_supported(nmat, batch, N, K, blocksize, dtype, packed, absmax))
        return FP4_B200_ERR_UNSUPPORTED;
    return counted(gemv_stream_group_dispatch(x, nmat, packed, absmax, bias, out, N, batch, K, dtype, tp, epi,
                                      (cudaStream_t)stream));
}

int fp4_b200_gemv_grouped_tp(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                             const void* const* bias, void* const* out, const int* N, int batch, int K,
                             int blocksize, int dtype, unsigned flags, const fp4_b200_tp_t* tp, void* stream) {
    return fp4_b200_gemv_grouped_ex(x, nmat, packed, absmax, bias, out, N, batch, K, blocksize, dtype, flags, tp,
                                    nullptr, stream);
}

int fp4_b200_gemv_grouped(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                          const void* const* bias, void* const* out, const int* N, int batch, int K,
                          int blocksize, int dtype, unsigned flags, void* stream) {
    return fp4_b200_gemv_grouped_tp(x, nmat, packed, absmax, bias, out, N, batch, K, blocksize, dtype, flags,
                                    nullptr, stream);
}

size_t fp4_b200_gemv_workspace_bytes(int N) {
    (void)N;
    return 0;  // no kernel of this library needs scratch memory any more (kept for ABI compatibility)
}

struct fp4_b200_layer {
    fp4_b200_nested_t nested[4];
    const fp4_b200_nested_t* pnested[4];  // NULL, or &nested[m]
    int nmat;
    const uint8_t* packed[4];
    const float* absmax[4];
    const void* bias[4];
    int N[4];
    const float* code;
    int K, blocksize, dtype;
    unsigned flags;
};

fp4_b200_layer_t* fp4_b200_layer_create_grouped(int nmat, const uint8_t* const* packed, const float* const* absmax,
                                                const float* code, const void* const* bias, const int* N, int K,
                                                int blocksize, int dtype, unsigned flags) {
    if (nmat < 1 || nmat > 4 || !packed || !absmax || !N || K <= 0 || blocksize <= 0) return nullptr;
    fp4_b200_layer* l = new (std::nothrow) fp4_b200_layer();
    if (!l) return nullptr;
    l->nmat = nmat;
    for (int m = 0; m < nmat; ++m) {
        if (!packed[m] || N[m] <= 0) { delete l; return nullptr; }  // (absmax[m] may be NULL until set_nested)
        l->packed[m] = packed[m]; l->absmax[m] = absmax[m]; l->bias[m] = bias ? bias[m] : nullptr; l->N[m] = N[m];
    }
    l->code = code; l->K = K; l->blocksize = blocksize; l->dtype = dtype; l->flags = flags;
    return l;
}

fp4_b200_layer_t* fp4_b200_layer_create(const uint8_t* packed, const float* absmax, const float* code,
                                        const void* bias, int N, int K, int blocksize, int dtype,
                                        unsigned flags) {
    return fp4_b200_layer_create_grouped(1, &packed, &absmax, code, &bias, &N, K, blocksize, dtype, flags);
}

int fp4_b200_layer_set_nested(fp4_b200_layer_t* l, int m, const fp4_b200_nested_t* nested) {
    if (!l || !nested) return FP4_B200_ERR_NULL;
    if (m < 0 || m >= l->nmat) return FP4_B200_ERR_SHAPE;
    l->nested[m] = *nested;
    l->pnested[m] = &l->nested[m];
    return FP4_B200_OK;
}

int fp4_b200_layer_gemv(const fp4_b200_layer_t* lc, const void* x, void* out, int batch, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!lc) return FP4_B200_ERR_NULL;
    if (lc->nmat != 1) return FP4_B200_ERR_SHAPE;
    const fp4_b200_layer* l = lc;
    return fp4_b200_gemv(x, l->packed[0], l->absmax[0], l->pnested[0], l->code, l->bias[0], out, batch, l->N[0], l->K,
                         l->blocksize, l->dtype, l->flags, workspace, workspace_bytes, stream);
}

int fp4_b200_layer_gemv_grouped(const fp4_b200_layer_t* lc, const void* x, void* const* out, int batch,
                                const fp4_b200_tp_t* tp, void* stream) {
    if (!lc || !out) return FP4_B200_ERR_NULL;
    const fp4_b200_layer* l = lc;
    fp4_b200_epilogue_t epi = {};
    bool any = false;
    for (int m = 0; m < l->nmat; ++m) any = any || l->pnested[m];
    epi.nested = any ? l->pnested : nullptr;
    return fp4_b200_gemv_grouped_ex(x, l->nmat, l->packed, l->absmax, l->bias, out, l->N, batch, l->K, l->blocksize,
                                    l->dtype, l->flags, tp, any ? &epi : nullptr, stream);
}

void fp4_b200_layer_destroy(fp4_b200_layer_t* l) { delete l; }

int fp4_b200_gemm(const void* x, const uint8_t* packed, const float* absmax, const float* code,
                  const void* bias, void* out, int M, int N, int K, int blocksize, int dtype,
                  unsigned flags, void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace;
    (void)workspace_bytes;
    if (!x || !packed || !absmax || !out) return FP4_B200_ERR_NULL;
    if (M < 0 || N < 0 || K <= 0) return FP4_B200_ERR_SHAPE;
    if (dtype != FP4_B200_F16 && dtype != FP4_B200_BF16) return FP4_B200_ERR_DTYPE;
    if (ilog2_exact(blocksize) < 0) return FP4_B200_ERR_BLOCKSIZE;
    if (M == 0 || N == 0) return FP4_B200_OK;
    return counted(gemm_tcgen05_dispatch(x, packed, absmax, code, bias, out, M, N, K, blocksize, dtype,
                                 flags, (cudaStream_t)stream));
}

int fp4_b200_quantize(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed,
                      float* absmax, void* stream) {
    return counted(quantize_dispatch(w, dtype, n, blocksize, packed, absmax, (cudaStream_t)stream));
}

}  // extern "C"
