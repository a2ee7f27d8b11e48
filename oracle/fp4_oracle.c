/*
 * fp4_oracle.c — CPU restatement of the reference's FP4 Linear hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (torch_bnb_fp4_b200/, torch_bnb_fp4/,
 * torch_bnb_fp4_ext.py) may import, link or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or the
 * reported CPU baseline.
 *
 * Parity pin: the dequant and GEMV restatements are checked against outputs of the UNMODIFIED
 * reference CUDA extension (built by oracle/build_ref.py from /root/reference/csrc into oracle/_ref/)
 * run on a B200; those outputs are committed as tests/golden/ref_*.npz with the generating script
 * tests/golden/make_ref_golden.py.  The quantiser restates bitsandbytes 0.42 (third-party, pinned by
 * the reference's requirements.txt:1 `bitsandbytes<0.43`, not vendored under /root/reference): its
 * bitwise behaviour is "parity unpinned"; it is pinned statistically by the reference's 0.045-0.065
 * band (sanity_check.py:177-179, README.md:113-115).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (no FMA contraction: the reference's
 * dequant is one fp32 multiply followed by a conversion).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { FP4O_F16 = 0, FP4O_F32 = 1, FP4O_BF16 = 2 }; /* csrc/torch_fp4.cpp:22-26 enum order */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* __float2bfloat16_rn (csrc/dequant_fp4_optimized.cu:79-81): round to nearest even */
uint16_t fp4o_f32_to_bf16(float f) {
    uint32_t u = f2u(f);
    if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return (uint16_t)((u >> 16) | 0x0040u);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float fp4o_bf16_to_f32(uint16_t h) { return u2f((uint32_t)h << 16); }

/* __float2half_rn (csrc/dequant_fp4_optimized.cu:82-84): round to nearest even, IEEE binary16 */
uint16_t fp4o_f32_to_f16(float f) {
    const uint32_t u = f2u(f);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const uint32_t absu = u & 0x7FFFFFFFu;
    if (absu >= 0x7F800000u) return (uint16_t)(sign | (absu > 0x7F800000u ? 0x7E00u : 0x7C00u));
    if (absu >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); /* >= 65520 rounds to inf */
    if (absu < 0x33000001u) return (uint16_t)sign;              /* <= 2^-25 rounds to 0 */
    int32_t exp = (int32_t)(absu >> 23) - 127;
    uint32_t man = (absu & 0x007FFFFFu) | 0x00800000u;
    uint32_t shift, half;
    if (exp < -14) { shift = (uint32_t)(13 + (-14 - exp)); half = 0; }
    else { shift = 13; half = (uint32_t)(exp + 15) << 10; }
    uint32_t q = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u);
    const uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) q++;
    if (exp < -14) return (uint16_t)(sign | q);            /* subnormal (q may carry into exp=1) */
    return (uint16_t)(sign | (half + (q - 0x400u)));       /* carry propagates into the exponent */
}
float fp4o_f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
    if (exp == 0) {
        if (man == 0) return u2f(sign);
        int e = -1;
        do { man <<= 1; e++; } while (!(man & 0x400u));
        return u2f(sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13));
    }
    if (exp == 31) return u2f(sign | 0x7F800000u | (man << 13));
    return u2f(sign | ((exp + 112u) << 23) | (man << 13));
}

static inline float round_to(float v, int dtype) {
    if (dtype == FP4O_F16) return fp4o_f16_to_f32(fp4o_f32_to_f16(v));
    if (dtype == FP4O_BF16) return fp4o_bf16_to_f32(fp4o_f32_to_bf16(v));
    return v;
}
static inline void store_as(void* out, int64_t i, float v, int dtype) {
    if (dtype == FP4O_F16) ((uint16_t*)out)[i] = fp4o_f32_to_f16(v);
    else if (dtype == FP4O_BF16) ((uint16_t*)out)[i] = fp4o_f32_to_bf16(v);
    else ((float*)out)[i] = v;
}

/* bitsandbytes FP4 code == tree literals (csrc/dequant_fp4_optimized.cu:55-76) */
static const float BNB_CODE[16] = {
    0.00000000f,  5.208333333e-03f,  0.66666667f,  1.00000000f,  0.33333333f,  0.50000000f,
    0.16666667f,  0.25000000f,       -0.00000000f, -5.208333333e-03f, -0.66666667f, -1.00000000f,
    -0.33333333f, -0.50000000f,      -0.16666667f, -0.25000000f};
/* the reference's hard-coded kernel codebook (csrc/dequant_fp4_optimized.cu:24-46,
 * csrc/gemv_fp4_optimized.cu:28-50); differs from BNB_CODE by 1 / 12 / 2 ulp at nibbles 1 / 4 / 6 */
static const float REF_CODE_PARAM[16] = {
    0.00000f,  5.208333e-03f,  0.6666667f,  1.000000f,  0.333333f,  0.500000f,  0.1666667f,  0.250000f,
    -0.000000f, -5.208333e-03f, -0.6666667f, -1.000000f, -0.333333f, -0.500000f, -0.1666667f, -0.250000f};
void fp4o_bnb_code(float* out16) { memcpy(out16, BNB_CODE, sizeof(BNB_CODE)); }
void fp4o_ref_code_param(float* out16) { memcpy(out16, REF_CODE_PARAM, sizeof(REF_CODE_PARAM)); }

/* dequantize_fp4_tree, csrc/dequant_fp4_optimized.cu:55-76: literal * absmax * sign */
static float tree_decode(unsigned val, float absmax) {
    const float sign = (val & 8u) ? -1.0f : 1.0f;
    if (val & 4u) {
        if (val & 2u) return ((val & 1u) ? 0.25000000f : 0.16666667f) * absmax * sign;
        return ((val & 1u) ? 0.50000000f : 0.33333333f) * absmax * sign;
    }
    if (val & 2u) return ((val & 1u) ? 1.00000000f : 0.66666667f) * absmax * sign;
    return ((val & 1u) ? 5.208333333e-03f : 0.00000000f) * absmax * sign;
}

/* dequantize_blockwise_kernel_fp4, csrc/dequant_fp4_optimized.cu:107-121:
 * element 2j = high nibble of byte j (:117), 2j+1 = low nibble (:118); absmax index = byte / (blocksize/2)
 * (:110 with the launcher passing blocksize/2, :176); n odd -> (n+1)/2 bytes (:108). */
void fp4o_dequant_tree(const uint8_t* packed, const float* absmax, int64_t n, int blocksize,
                       int dtype, void* out) {
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t b = packed[i >> 1];
        const unsigned nib = (i & 1) ? (b & 0xFu) : (unsigned)(b >> 4);
        store_as(out, i, tree_decode(nib, absmax[i / blocksize]), dtype);
    }
}

/* dequantize_blockwise_codebook_kernel_fp4, csrc/dequant_fp4_optimized.cu:156-170: code[nib]*absmax.
 * (The reference passes its own CODE_PARAM and ignores the tensor argument, :207-255; the oracle
 * takes the table explicitly so both behaviours can be stated.) */
void fp4o_dequant_code(const uint8_t* packed, const float* absmax, const float* code16, int64_t n,
                       int blocksize, int dtype, void* out) {
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t b = packed[i >> 1];
        const unsigned nib = (i & 1) ? (b & 0xFu) : (unsigned)(b >> 4);
        store_as(out, i, code16[nib] * absmax[i / blocksize], dtype);
    }
}

/* nested absmax (bitsandbytes dequantize_blockwise + offset):
 * two separately rounded fp32 operations. */
void fp4o_denest(const uint8_t* qabsmax, const float* code2, const float* absmax2, float offset,
                 int blocksize2, int64_t nblocks, float* out) {
    for (int64_t i = 0; i < nblocks; ++i) {
        const float p = code2[qabsmax[i]] * absmax2[i / blocksize2];
        out[i] = p + offset;
    }
}

/* bitsandbytes 0.42 dQuantizeFP4 thresholds + kQuantizeBlockwise<FP4> packing, reached by the
 * reference at torch_bnb_fp4/__init__.py:775 (BF.quantize_fp4).  [third-party; parity unpinned] */
static unsigned quantize_nibble(float x) {
    const unsigned sign = x < 0.f ? 8u : 0u;
    x = fabsf(x);
    if (x > 0.29166667f) {
        if (x > 0.583333f) return (x > 0.8333333f ? 3u : 2u) + sign;
        return (x > 0.4166667f ? 5u : 4u) + sign;
    }
    if (x > 0.0859375f) return (x > 0.20833333f ? 7u : 6u) + sign;
    return (x > 0.00260417f ? 1u : 0u) + sign;
}
void fp4o_quantize(const float* w, int64_t n, int blocksize, uint8_t* packed, float* absmax) {
    const int64_t nblocks = (n + blocksize - 1) / blocksize;
    memset(packed, 0, (size_t)((n + 1) / 2));
    for (int64_t b = 0; b < nblocks; ++b) {
        const int64_t e0 = b * blocksize, e1 = (e0 + blocksize < n) ? e0 + blocksize : n;
        float m = 0.f;
        for (int64_t i = e0; i < e1; ++i) m = fmaxf(m, fabsf(w[i]));
        absmax[b] = m;
        const float inv = 1.0f / m;
        for (int64_t i = e0; i < e1; ++i) {
            const unsigned q = quantize_nibble(w[i] * inv);
            packed[i >> 1] |= (uint8_t)((i & 1) ? q : (q << 4));
        }
    }
}

/* Ground truth linear: y[b,r] = sum_k x[b,k] * W[r,k] (+ bias[r]) accumulated in fp64.
 * W[r,k] = round_to(code[nib] * absmax, wdtype): wdtype = FP4O_F32 is the exact fp32 product,
 * FP4O_F16/BF16 is the weight the reference's dequant + F.linear path multiplies with
 * (torch_bnb_fp4/__init__.py:423-436).  x and bias are given as fp32 values (already rounded to
 * the compute dtype by the caller). */
void fp4o_linear_f64(const float* x, const uint8_t* packed, const float* absmax,
                     const float* code16, const float* bias, int batch, int N, int K, int blocksize,
                     int wdtype, double* out) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < N; ++r) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int b0 = 0; b0 < batch; b0 += 8) {
            const int nb = batch - b0 < 8 ? batch - b0 : 8;
            for (int j = 0; j < nb; ++j) acc[j] = 0.0;
            for (int k = 0; k < K; ++k) {
                const int64_t i = (int64_t)r * K + k;
                const uint8_t by = packed[i >> 1];
                const unsigned nib = (i & 1) ? (by & 0xFu) : (unsigned)(by >> 4);
                const double w = (double)round_to(code16[nib] * absmax[i / blocksize], wdtype);
                for (int j = 0; j < nb; ++j) acc[j] += w * (double)x[(int64_t)(b0 + j) * K + k];
            }
            for (int j = 0; j < nb; ++j)
                out[(int64_t)(b0 + j) * N + r] = acc[j] + (bias ? (double)bias[r] : 0.0);
        }
    }
}

/* CPU baseline of BASELINE.json config #1: fp32 dequant (code*absmax) + fp32 dot products,
 * one row per OpenMP task; this is the "pure CPU dequantize_fp4 + matmul" leg that bench.py times. */
void fp4o_linear_f32(const float* x, const uint8_t* packed, const float* absmax,
                     const float* code16, int batch, int N, int K, int blocksize, float* out) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < N; ++r) {
        float acc[8];
        for (int b0 = 0; b0 < batch; b0 += 8) {
            const int nb = batch - b0 < 8 ? batch - b0 : 8;
            for (int j = 0; j < nb; ++j) acc[j] = 0.f;
            for (int k = 0; k < K; k += 2) {
                const int64_t i = (int64_t)r * K + k;
                const uint8_t by = packed[i >> 1];
                const float a = absmax[i / blocksize];
                const float w0 = code16[by >> 4] * a, w1 = code16[by & 0xFu] * a;
                for (int j = 0; j < nb; ++j) {
                    const float* xr = x + (int64_t)(b0 + j) * K + k;
                    acc[j] += w0 * xr[0] + w1 * xr[1];
                }
            }
            for (int j = 0; j < nb; ++j) out[(int64_t)(b0 + j) * N + r] = acc[j];
        }
    }
}

/* Emulation of the reference GEMV numerics for T in {fp16, bf16} (SURVEY.md §8 N4;
 * csrc/gemv_fp4_optimized.cu:60-157): quant_map[i] = T(CODE_PARAM[i]) (:95); per lane, k-stride 1024
 * (:99); local_absmax = T(absmax) (:103); w = quant_map[nib] * local_absmax in T (:128-129);
 * local_C += a * w in T (:147; ptxas contracts it to one fused multiply-add, HFMA2, so one rounding);
 * warp sum in fp32 with a shuffle-down tree, offsets 16,8,4,2,1 (:152, cub::WarpReduce); out = T(sum).
 * dtype FP4O_F32 follows gemv_4bit_inference_kernel_float (:159-259) with fp32 fused multiply-adds.
 * x given as fp32 values already rounded to T.  Batch 1, K % 32 == 0. */
void fp4o_gemv_ref_emulate(const float* x, const uint8_t* packed, const float* absmax, int N, int K,
                           int blocksize, int dtype, float* out) {
    float qmap[16];
    for (int i = 0; i < 16; ++i) qmap[i] = round_to(REF_CODE_PARAM[i], dtype);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < N; ++r) {
        float lane_c[32];
        for (int lane = 0; lane < 32; ++lane) {
            float c = 0.f;
            for (int k0 = lane * 32; k0 < K; k0 += 32 * 32) {
                const int64_t e0 = (int64_t)r * K + k0;
                const float am = round_to(absmax[e0 / blocksize], dtype);
                for (int j = 0; j < 32; ++j) {
                    const int64_t i = e0 + j;
                    const uint8_t by = packed[i >> 1];
                    const unsigned nib = (i & 1) ? (by & 0xFu) : (unsigned)(by >> 4);
                    const float w = round_to(qmap[nib] * am, dtype);
                    if (dtype == FP4O_F32) c = fmaf(x[k0 + j], w, c);
                    else c = round_to((float)((double)x[k0 + j] * (double)w + (double)c), dtype);
                }
            }
            lane_c[lane] = c;
        }
        for (int off = 16; off > 0; off >>= 1)
            for (int lane = 0; lane < off; ++lane) lane_c[lane] = lane_c[lane] + lane_c[lane + off];
        out[r] = round_to(lane_c[0], dtype);
    }
}

int fp4o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
