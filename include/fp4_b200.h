/*
 * fp4_b200.h — C-ABI of libfp4_b200.so: the B200 (sm_100a) implementation of the
 * bitsandbytes-FP4 quantized nn.Linear hot path of aredden/torch-bnb-fp4.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and a CUDA stream handle (cudaStream_t passed as void*); no torch types.
 * The reference binds the same operations through pybind11 in csrc/torch_fp4.cpp;
 * each function below names the reference interface it replaces (file:line are
 * relative to the reference repository).  The Python module `torch_bnb_fp4_ext`
 * shipped in this repo is a thin binding over exactly these symbols.
 *
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device unless noted;
 *   - `stream` is a cudaStream_t (0 = the legacy default stream); every call is
 *     asynchronous on that stream, never synchronises, and is CUDA-graph capturable;
 *   - return value: 0 on success, FP4_B200_ERR_* (<0) for rejected arguments,
 *     or a positive cudaError_t for a launch failure.  Nothing is printed.
 *   - packed layout (bitsandbytes `Params4bit.data`): element 2j is the HIGH nibble
 *     of byte j, element 2j+1 the LOW nibble (reference csrc/dequant_fp4_optimized.cu:117-118);
 *     quantisation blocks run over the FLATTENED [N*K] array, one fp32 absmax per
 *     `blocksize` elements (reference :110).
 *   - nibble -> value: the 16-entry fp32 codebook `code` (bitsandbytes QuantState.code);
 *     passing code == NULL selects the bitsandbytes FP4 constants
 *     {0, 0.0052083333, 0.66666667, 1, 0.33333333, 0.5, 0.16666667, 0.25} (and negatives for
 *     nibble >= 8), which are the literals of the reference's tree decoder
 *     (csrc/dequant_fp4_optimized.cu:55-76).
 */
#ifndef FP4_B200_H_
#define FP4_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP4_B200_ABI_VERSION 1

/* dtype codes (order matches the reference's ScalarTypeEnum, csrc/torch_fp4.cpp:22-26) */
enum {
    FP4_B200_F16 = 0,
    FP4_B200_F32 = 1,
    FP4_B200_BF16 = 2
};

/* status codes */
enum {
    FP4_B200_OK = 0,
    FP4_B200_ERR_NULL = -1,       /* a required pointer is NULL                              */
    FP4_B200_ERR_DTYPE = -2,      /* dtype code not one of FP4_B200_{F16,F32,BF16}            */
    FP4_B200_ERR_SHAPE = -3,      /* negative / inconsistent sizes                            */
    FP4_B200_ERR_BLOCKSIZE = -4,  /* blocksize not a power of two >= 2                        */
    FP4_B200_ERR_ALIGN = -5,      /* pointer or row pitch not aligned as the kernel requires  */
    FP4_B200_ERR_BATCH = -6,      /* gemv batch outside 1..8                                  */
    FP4_B200_ERR_UNSUPPORTED = -7,/* shape outside what this entry point implements           */
    FP4_B200_ERR_WORKSPACE = -8   /* workspace missing or too small                           */
};

/* flags for fp4_b200_gemv / fp4_b200_gemm */
enum {
    /* caller guarantees `code` (if non-NULL) holds the bitsandbytes FP4 constants bit for bit;
       lets the GEMV use its integer tensor-core decode.  Without the flag a non-NULL code
       is honoured entry by entry through the generic kernel. */
    FP4_B200_FLAG_CODE_IS_BNB_FP4 = 1,
    /* force the generic CUDA-core GEMV (testing / A-B timing) */
    FP4_B200_FLAG_FORCE_GENERIC = 2,
    /* accepted and ignored (they selected GEMV kernel families of ABI version 1 that no longer exist) */
    FP4_B200_FLAG_NO_TMA = 4,
    FP4_B200_FLAG_NO_I8 = 8,
    /* do not use the streaming integer tensor-core GEMV; take the generic CUDA-core kernel (testing / A-B timing) */
    FP4_B200_FLAG_NO_STREAM = 16
};

/* nested ("double-quantised") absmax, bitsandbytes QuantState.state2 + offset:
 *   absmax_f32[i] = fp32_add(fp32_mul(code2[qabsmax[i]], absmax2[i / blocksize2]), offset)
 * two separately rounded fp32 operations (no FMA).  SURVEY.md §8 N5. */
typedef struct {
    const uint8_t* qabsmax;  /* [nblocks] uint8 codes                         */
    const float* code2;      /* [256] fp32 map (state2.code)                  */
    const float* absmax2;    /* [ceil(nblocks / blocksize2)] fp32             */
    float offset;            /* quant_state.offset                            */
    int blocksize2;          /* state2.blocksize (256 in bitsandbytes)        */
} fp4_b200_nested_t;

/* Tensor-parallel exchange through peer (symmetric) memory, for fp4_b200_gemv_grouped_tp (16-bit dtypes).
 * A row-parallel layer (o / down projection) produces PARTIAL sums.  Instead of a collective launch it PUSHES
 * them, as 64-bit words {two 16-bit values, tag32 = epoch}, into the exchange buffer of every rank (plain 8-byte
 * NVLink stores, no fences, no flags: each word validates itself, as in NCCL's LL protocol); the consumer - the
 * next column-parallel layer - reads only its LOCAL buffer while it stages x, re-reading words whose tag is not
 * yet the current epoch, and sums the ranks' partials in fp32 in rank order.  The 32-bit tag does not come round
 * again in the life of a process, so a word left over from an earlier, larger batch can never validate.
 * Exchange buffer of a rank: [2 slots (epoch parity)][world][slot_bytes / 8 words]; rank r's partial for epoch e
 * starts at word ((e & 1) * world + r) * slot_bytes / 8; within it the word (b * N + (row & ~15)) / 2 + (row & 7)
 * carries output rows `row` (bit 3 clear, low half) and `row + 8` (high half) of batch row b: N % 16 == 0,
 * batch * N * 4 <= slot_bytes.  Must start zero-filled.
 *   consumer:  in_world > 1, in_base = this rank's exchange buffer;
 *   producer:  out_world > 1, out_peer_base[q] = rank q's exchange buffer as mapped on this GPU (the call's
 *              out[0] is ignored), out_rank = this rank.
 * epochs: two uint32 in local device memory, zero-initialised: [1] = last epoch this rank's consumers finished
 * (the exchange epoch of a producer and of the consumer after it is epochs[1] + 1; [0] is reserved).  Every rank
 * must issue the same sequence of calls, and a published partial must be consumed - by exactly one consumer launch -
 * before the next one is published.
 * A consumer launch with x == NULL does NOT wait for the preceding kernels of its stream (its input validates
 * itself word by word, so it polls while the producers are still running); its bias and output buffers must
 * therefore not be written or read by the kernel directly before it.  A wait that gets no data for 20 s sets err
 * to 1 and traps. */
typedef struct {
    int in_world;
    const void* in_base;
    uint32_t slot_bytes;
    int out_world, out_rank;
    void* out_peer_base[8];
    uint32_t* epochs;
    uint32_t* err;
} fp4_b200_tp_t;

int fp4_b200_abi_version(void);
/* number of kernels this library has launched in this process (diagnostics / benchmark bookkeeping) */
unsigned long long fp4_b200_launch_count(void);
const char* fp4_b200_status_string(int status);

/* Blockwise dequantise n elements:  out[i] = RN_T(fp32_mul(code[nib_i], absmax[i / blocksize])).
 * Replaces dequantize_fp4 (csrc/torch_fp4.cpp:41-50 -> csrc/dequant_fp4_optimized.cu:182-205,
 * tree kernel :89-123) when code == NULL, and dequantize_fp4_codebook (csrc/torch_fp4.cpp:52-62 ->
 * csrc/dequant_fp4_optimized.cu:207-255) when code != NULL.
 * packed: ceil(n/2) bytes, absmax: ceil(n/blocksize) floats, out: n elements of out_dtype.
 * Requires: out 32-byte aligned, packed 8-byte aligned (torch allocations are 512-byte aligned). */
int fp4_b200_dequantize(const uint8_t* packed, const float* absmax, const float* code,
                        void* out, int64_t n, int blocksize, int out_dtype, void* stream);

/* Same, with the absmax itself still double-quantised (decoded in the kernel, never materialised). */
int fp4_b200_dequantize_nested(const uint8_t* packed, const fp4_b200_nested_t* nested,
                               const float* code, void* out, int64_t n, int blocksize,
                               int out_dtype, void* stream);

/* Materialise a nested absmax to fp32 (load-time helper). */
int fp4_b200_absmax_denest(const fp4_b200_nested_t* nested, float* absmax_out, int64_t nblocks,
                           void* stream);

/* Fused dequant + GEMV for decode:  out[b, r] = T( sum_k x[b,k] * code[W[r,k]] * absmax[(r*K+k)/blocksize] + bias[r] )
 * for b < batch (1..8), r < N.  fp32 accumulation.  Replaces gemv_fp4 (csrc/torch_fp4.cpp:105-123 ->
 * csrc/gemv_fp4_optimized.cu:277-368, kernels :60-259), which handles batch 1 only and adds the bias
 * in a separate op (torch_bnb_fp4/__init__.py:608-613).
 * x: [batch, K] dtype (row pitch K), out: [batch, N] dtype, bias: [N] dtype or NULL.
 * Requires K % 32 == 0 (as the reference: 16-byte row chunks), blocksize % 32 == 0, K % blocksize == 0
 * is NOT required (blocks may straddle rows as in bitsandbytes).
 * nested may be NULL (absmax is fp32) or non-NULL (absmax ignored, decoded in the kernel).
 *
 * Requires for the streaming tensor-core kernel: bitsandbytes codebook (code == NULL or
 * FP4_B200_FLAG_CODE_IS_BNB_FP4), blocksize 64, K % 256 == 0, N % 16 == 0, 16-byte aligned buffers; a batch whose
 * activations do not fit shared memory together runs as two launches.  Anything else takes the generic kernel.
 *
 * workspace / workspace_bytes: ignored (no kernel needs scratch memory any more; fp4_b200_gemv_workspace_bytes
 * returns 0).  The parameters remain so that callers of ABI version 1 keep working. */
int fp4_b200_gemv(const void* x, const uint8_t* packed, const float* absmax,
                  const fp4_b200_nested_t* nested, const float* code, const void* bias, void* out,
                  int batch, int N, int K, int blocksize, int dtype, unsigned flags,
                  void* workspace, size_t workspace_bytes, void* stream);
size_t fp4_b200_gemv_workspace_bytes(int N);

/* Prepared layer: the per-layer constants of fp4_b200_gemv bound once, so that a decode step costs the host
 * one short call per layer (FFI argument marshalling is a measurable part of an eager, un-graphed token:
 * the reference pays it per nn.Linear in torch_bnb_fp4/__init__.py:471-492 -> csrc/torch_fp4.cpp:105-123).
 * The handle is host memory holding the pointers and sizes given here; it does not own or copy the device
 * buffers, which must outlive it.  fp4_b200_layer_gemv(h, ...) == fp4_b200_gemv(x, packed, absmax, NULL,
 * code, bias, out, batch, N, K, blocksize, dtype, flags, workspace, workspace_bytes, stream). */
typedef struct fp4_b200_layer fp4_b200_layer_t;
fp4_b200_layer_t* fp4_b200_layer_create(const uint8_t* packed, const float* absmax, const float* code,
                                        const void* bias, int N, int K, int blocksize, int dtype,
                                        unsigned flags);
int fp4_b200_layer_gemv(const fp4_b200_layer_t* layer, const void* x, void* out, int batch,
                        void* workspace, size_t workspace_bytes, void* stream);
void fp4_b200_layer_destroy(fp4_b200_layer_t* layer);

/* Prepared GROUP of 1..4 layers that share the input (fp4_b200_gemv_grouped with its host arrays bound once).
 * fp4_b200_layer_gemv_grouped(h, x, out[], batch, tp, stream) == fp4_b200_gemv_grouped_tp(x, nmat, packed, absmax,
 * bias, out, N, batch, K, blocksize, dtype, flags, tp, stream); `out` is a HOST array of nmat device pointers. */
fp4_b200_layer_t* fp4_b200_layer_create_grouped(int nmat, const uint8_t* const* packed, const float* const* absmax,
                                                const float* code, const void* const* bias, const int* N, int K,
                                                int blocksize, int dtype, unsigned flags);
/* matrix m of a prepared group keeps its absmax double-quantised (the struct is copied; its buffers must outlive
 * the handle); absmax[m] given at creation is then ignored */
int fp4_b200_layer_set_nested(fp4_b200_layer_t* layer, int m, const fp4_b200_nested_t* nested);
int fp4_b200_layer_gemv_grouped(const fp4_b200_layer_t* layer, const void* x, void* const* out, int batch,
                                const fp4_b200_tp_t* tp, void* stream);

/* Grouped fused dequant + GEMV: nmat (1..4) weight matrices with the same K applied to the SAME x in one
 * launch - out[m][b, r] = T( sum_k x[b,k] * W_m[r,k] + bias_m[r] ) - e.g. the q/k/v or gate/up projections
 * of a decoder layer, which the reference issues as separate gemv_fp4 calls
 * (torch_bnb_fp4/__init__.py:471-492 once per nn.Linear).  The arrays are HOST arrays of device pointers /
 * sizes.  bias may be NULL, or hold NULL entries.  Results equal nmat fp4_b200_gemv calls up to fp32 summation order (deterministic).
 * Requires FP4_B200_FLAG_CODE_IS_BNB_FP4, blocksize 64, K % 256 == 0, every N[m] % 16 == 0; otherwise
 * FP4_B200_ERR_UNSUPPORTED (the caller then issues the calls one by one). */
int fp4_b200_gemv_grouped(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                          const void* const* bias, void* const* out, const int* N, int batch, int K,
                          int blocksize, int dtype, unsigned flags, void* stream);

/* Dequant-fused tensor-core GEMM for prefill:  out[m, r] = T( sum_k x[m,k] * W[r,k] + bias[r] ),
 * W dequantised tile by tile in shared memory (never written to HBM) and multiplied with tcgen05.mma,
 * fp32 accumulators in TMEM.  Replaces the reference's dequant + cuBLAS pair
 * (torch_bnb_fp4/__init__.py:423-436; csrc/torch_fp4.cpp:64-103).
 * dtype: FP4_B200_BF16 or FP4_B200_F16.  Requires K % 64 == 0, N % 8 == 0, blocksize % 64 == 0.
 * workspace: reserved (may be NULL). */
int fp4_b200_gemm(const void* x, const uint8_t* packed, const float* absmax, const float* code,
                  const void* bias, void* out, int M, int N, int K, int blocksize, int dtype,
                  unsigned flags, void* workspace, size_t workspace_bytes, void* stream);

/* Blockwise FP4 quantiser with the bitsandbytes thresholds (kQuantizeBlockwise<FP4>; reached by the
 * reference through bitsandbytes: torch_bnb_fp4/__init__.py:775).  w: n elements of dtype,
 * packed: ceil(n/2) bytes, absmax: ceil(n/blocksize) floats. */
int fp4_b200_quantize(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed,
                      float* absmax, void* stream);

/* Fused neighbours of the Linear: what the reference runs as separate elementwise kernels
 * after its gemv_fp4 calls (torch_bnb_fp4/__init__.py:608-613 adds even the bias as a second op).
 *   gate_act != 0 (1 = SiLU, 2 = GELU tanh approximation): the call has nmat == 2 matrices of the same N - the gate
 *     and the up projection of a gated MLP - and writes ONE output, out[0][b, r] =
 *     act(x W_gate^T + bias_gate)[b, r] * (x W_up^T + bias_up)[b, r]; out[1] is ignored.  N % 8 == 0.
 *   residual: NULL, or a HOST array of nmat device pointers (entries may be NULL) to [batch, N_m] tensors of the
 *     call's dtype that are added to out[m] (the residual stream around an o / down projection). */
typedef struct {
    int gate_act;
    const void* const* residual;
    /* NULL, or a HOST array of nmat pointers (entries may be NULL): matrix m carries a double-quantised absmax that
     * is decoded in the kernel (absmax[m] is then ignored and may be NULL) */
    const fp4_b200_nested_t* const* nested;
} fp4_b200_epilogue_t;
int fp4_b200_gemv_grouped_ex(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                             const void* const* bias, void* const* out, const int* N, int batch, int K,
                             int blocksize, int dtype, unsigned flags, const fp4_b200_tp_t* tp,
                             const fp4_b200_epilogue_t* epilogue, void* stream);

/* fp4_b200_gemv_grouped with the tensor-parallel exchange described by fp4_b200_tp_t; tp may be NULL (plain).
 * x may be NULL when tp->in_world > 1 (x is then the sum of the ranks' partials). */
int fp4_b200_gemv_grouped_tp(const void* x, int nmat, const uint8_t* const* packed, const float* const* absmax,
                             const void* const* bias, void* const* out, const int* N, int batch, int K,
                             int blocksize, int dtype, unsigned flags, const fp4_b200_tp_t* tp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FP4_B200_H_ */
