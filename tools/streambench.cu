// Stand-alone HBM read-stream microbenchmark for B200 (test infrastructure, not product code).
// What can a kernel shaped like a decode GEMV (read S bytes once, tiny output, chained launches) reach?
//   ldg   : grid-stride 128-bit loads
//   bulk  : persistent CTAs, cp.async.bulk 1-D chunks into a shared-memory ring (producer lane + consumer warps),
//           static contiguous partition or dynamic chunk grabbing; optional PDL with pre-filled rings
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/streambench.bin tools/streambench.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c));
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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ src, size_t n16, const float* xin,
                                                  float* xout, int pdl) {
    if (pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    uint32_t acc = __float_as_uint(xin[threadIdx.x]);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride),
              d = __ldcs(src + i + 3 * stride);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n16; i += stride) {
        uint4 a = __ldcs(src + i);
        acc ^= a.x ^ a.y ^ a.z ^ a.w;
    }
    if (acc == 0x12345678u) xout[threadIdx.x] = 1.f;  // practically never
    if (blockIdx.x == 0 && threadIdx.x < 32) xout[threadIdx.x] = xin[threadIdx.x];
}


// CTA-contiguous 128-bit loads; optionally every CTA first asks L2 to prefetch its whole range
// (cp.async.bulk.prefetch.L2) BEFORE griddepcontrol.wait, i.e. while the previous launch still runs.
template <int UNROLL>
__global__ void __launch_bounds__(256) ldgc_kernel(const uint8_t* __restrict__ src, size_t bytes, const float* xin,
                                                   float* xout, int pdl, int prefetch, uint32_t pf_piece) {
    const size_t per = ((bytes / gridDim.x) + 4095) & ~(size_t)4095;
    const size_t b0 = (size_t)blockIdx.x * per, b1 = b0 + per < bytes ? b0 + per : bytes;
    // prefetch == 1: the whole range at once; prefetch >= 2: a rolling window of `prefetch` iterations
    const size_t iter_bytes = (size_t)UNROLL * 256 * 16;
    if (prefetch >= 2 && threadIdx.x < 32) {
        const size_t w1 = b0 + iter_bytes * prefetch < b1 ? b0 + iter_bytes * prefetch : b1;
        for (size_t o = b0 + (size_t)threadIdx.x * pf_piece; o < w1; o += (size_t)32 * pf_piece) {
            const uint32_t sz = (uint32_t)((w1 - o) < pf_piece ? (w1 - o) : pf_piece);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + o), "r"(sz) : "memory");
        }
    }
    if (prefetch == 1 && threadIdx.x < 32) {
        for (size_t o = b0 + (size_t)threadIdx.x * pf_piece; o < b1; o += (size_t)32 * pf_piece) {
            const uint32_t sz = (uint32_t)((b1 - o) < pf_piece ? (b1 - o) : pf_piece);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + o), "r"(sz) : "memory");
        }
    }
    if (pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    uint32_t acc = __float_as_uint(__ldcg(xin + threadIdx.x));
    const uint4* p = reinterpret_cast<const uint4*>(src + b0) + threadIdx.x;
    const size_t n16 = b0 < b1 ? (b1 - b0) / 16 : 0;
    size_t i = threadIdx.x;
    for (; i + (UNROLL - 1) * 256 < n16; i += UNROLL * 256) {
        uint4 v[UNROLL];
        if (prefetch >= 2) {  // keep the window `prefetch` iterations ahead: iter_bytes more, split over pf lanes
            const size_t o = b0 + (i - threadIdx.x) * 16 + iter_bytes * prefetch + (size_t)threadIdx.x * pf_piece;
            if ((size_t)threadIdx.x * pf_piece < iter_bytes && o < b1) {
                const uint32_t sz = (uint32_t)((b1 - o) < pf_piece ? (b1 - o) : pf_piece);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + o), "r"(sz) : "memory");
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                         : "l"(p + (i - threadIdx.x) + u * 256));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += 256) {
        uint4 v = __ldcs(p + (i - threadIdx.x));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) xout[threadIdx.x] = 1.f;
    if (blockIdx.x == 0 && threadIdx.x < 32) xout[threadIdx.x] = __ldcg(xin + threadIdx.x);
}


// Cross-launch prefetch: launch i demand-reads region i (which launch i-1 asked L2 to prefetch) and asks L2
// to prefetch region i+1 while it runs.  nextoff = byte distance to the next region (0 = no prefetch).
template <int UNROLL>
__global__ void __launch_bounds__(512) ldgn_kernel(const uint8_t* __restrict__ src, size_t bytes, size_t nextoff,
                                                   const float* xin, float* xout, uint32_t pf_piece, int evict_first) {
    const size_t per = ((bytes / gridDim.x) + 4095) & ~(size_t)4095;
    const size_t b0 = (size_t)blockIdx.x * per, b1 = b0 + per < bytes ? b0 + per : bytes;
    if (nextoff && threadIdx.x < 32) {
        for (size_t o = b0 + (size_t)threadIdx.x * pf_piece; o < b1; o += (size_t)32 * pf_piece) {
            const uint32_t sz = (uint32_t)((b1 - o) < pf_piece ? (b1 - o) : pf_piece);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + nextoff + o), "r"(sz) : "memory");
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    uint32_t acc = __float_as_uint(__ldcg(xin + (threadIdx.x & 255)));
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const uint4* p = reinterpret_cast<const uint4*>(src + b0) + threadIdx.x;
    const size_t n16 = b0 < b1 ? (b1 - b0) / 16 : 0;
    size_t i = threadIdx.x;
    for (; i + (UNROLL - 1) * 512 < n16; i += UNROLL * 512) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (evict_first)
                asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                             : "l"(p + (i - threadIdx.x) + u * 512), "l"(pol));
            else
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                             : "l"(p + (i - threadIdx.x) + u * 512));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += 512) {
        uint4 v = __ldcs(p + (i - threadIdx.x));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) xout[threadIdx.x & 255] = 1.f;
    if (blockIdx.x == 0 && threadIdx.x < 32) xout[threadIdx.x] = __ldcg(xin + threadIdx.x);
}

struct BulkParams {
    const uint8_t* src;
    uint32_t nchunks, chunk, stages, dynamic, pdl, nprod;
    unsigned* counter;
    const float* xin;
    float* xout;
    long long* tl;  // [grid][4]
};

// warp 0 = producers (nprod lanes), warps 1..NW = consumers
template <int NW>
__global__ void __launch_bounds__(32 + NW * 32) bulk_kernel(const __grid_constant__ BulkParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t S = p.stages;
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.chunk);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + S * 8;
    int* sidx = reinterpret_cast<int*>(bars + 2 * S);  // [S] chunk index per slot (dynamic)
    float* sx = reinterpret_cast<float*>(sidx + S);    // 2048 floats
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(full0 + s * 8, 1);
            mbar_init(empty0 + s * 8, NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint32_t G = gridDim.x, c = blockIdx.x;
    const uint32_t q = p.nchunks / G, r = p.nchunks % G;
    const uint32_t my_b = c * q + min(c, r), my_n = q + (c < r ? 1u : 0u);
    long long t0 = 0;
    if (tid == 0) t0 = gtime();

    if (warp == 0) {
        if (lane < (int)p.nprod) {
            // slot sequence n = lane, lane + nprod, ...: each producer lane fills its own slots
            for (uint32_t n = lane;; n += p.nprod) {
                uint32_t idx;
                if (p.dynamic) {
                    // the first `stages` chunks per CTA are static (no atomic latency before the first bytes)
                    idx = n < S ? c * S + n : G * S + atomicAdd(p.counter, 1u);
                    if (n < S && idx >= p.nchunks) idx = 0xffffffffu;
                } else {
                    idx = n < my_n ? my_b + n : 0xffffffffu;
                }
                const uint32_t s = n % S, ph = (n / S) & 1;
                if (n >= S) mbar_wait(empty0 + s * 8, ph ^ 1);
                if (idx >= p.nchunks) {
                    sidx[s] = -1;
                    mbar_arrive(full0 + s * 8);
                    break;
                }
                sidx[s] = (int)idx;
                mbar_expect_tx(full0 + s * 8, p.chunk);
                bulk_load(smem_u32(ring + (size_t)s * p.chunk), p.src + (size_t)idx * p.chunk, p.chunk, full0 + s * 8);
            }
        }
        return;
    }
    // consumers
    if (p.pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    const int ct = tid - 32;
    for (int i = ct; i < 2048; i += NW * 32) sx[i] = __ldcg(p.xin + i);
    asm volatile("bar.sync 1, %0;" ::"r"(NW * 32));
    long long t1 = 0;
    if (ct == 0) t1 = gtime();
    uint32_t acc = __float_as_uint(sx[ct]);
    const uint32_t per_thread = p.chunk / (NW * 32);  // bytes per consumer thread per chunk (multiple of 16)
    uint32_t ended = 0, nend = 0;
    for (uint32_t n = 0;; ++n) {
        const uint32_t pl = n % p.nprod;
        if ((ended >> pl) & 1u) continue;  // that producer lane has stopped filling its slots
        const uint32_t s = n % S, ph = (n / S) & 1;
        mbar_wait(full0 + s * 8, ph);
        if (sidx[s] < 0) {
            ended |= 1u << pl;
            if (++nend == p.nprod) break;
            continue;
        }
        const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)s * p.chunk) + ct;
        for (uint32_t b = 0; b < per_thread / 16; ++b) {
            uint4 v = src[b * NW * 32];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + s * 8);
    }
    if (acc == 0x12345678u) p.xout[ct] = 1.f;
    if (c == 0 && ct < 32) p.xout[ct] = sx[ct];
    if (ct == 0 && p.tl) {
        p.tl[c * 4 + 0] = t0;
        p.tl[c * 4 + 1] = t1;
        p.tl[c * 4 + 2] = gtime();
    }
}

int main(int argc, char** argv) {
    const size_t pool_bytes = (size_t)2 << 30;  // 2 GiB pool >> L2
    uint8_t* pool;
    CK(cudaMalloc(&pool, pool_bytes));
    CK(cudaMemset(pool, 0x5a, pool_bytes));
    float *xa, *xb;
    CK(cudaMalloc(&xa, 8192));
    CK(cudaMalloc(&xb, 8192));
    CK(cudaMemset(xa, 0, 8192));
    CK(cudaMemset(xb, 0, 8192));
    unsigned* counters;
    CK(cudaMalloc(&counters, 4096 * 4));
    long long* tl;
    CK(cudaMalloc(&tl, 4096 * 4 * 8));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const int L = 48;  // launches per graph
    const size_t sizes[] = {2359296 + 0, 9437184, 33030144, 132120576};
    const char* names[] = {"1024x4096", "4096x4096", "14336x4096", "28672x8192"};

    auto run = [&](const char* label, size_t S, auto&& launch_one) {
        // graph of L chained launches over consecutive regions of the pool
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        size_t off = 0;
        for (int i = 0; i < L; ++i) {
            if (off + S > pool_bytes) off = 0;
            launch_one(pool + off, i);
            off += (S + 4095) & ~(size_t)4095;
        }
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaMemsetAsync(counters, 0, 4096 * 4, st));
            CK(cudaEventRecord(e0, st));
            CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep >= 2) best = std::min(best, ms);
        }
        const double us = best * 1e3 / L;
        printf("%-44s %9.2f us/launch  %8.1f GB/s\n", label, us, S / us * 1e-3);
        CK(cudaGraphExecDestroy(ge));
        CK(cudaGraphDestroy(g));
    };

    for (int si = 0; si < 4; ++si) {
        const size_t S = sizes[si];
        printf("== %s (%zu bytes) ==\n", names[si], S);
        for (int pdl = 0; pdl < 2; ++pdl) {
            for (int occ : {2, 4}) {
                char label[128];
                snprintf(label, sizeof label, "ldg grid=148x%d pdl=%d", occ, pdl);
                run(label, S, [&](const uint8_t* src, int i) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(148 * occ);
                    cfg.blockDim = dim3(512);
                    cfg.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = pdl;
                    CK(cudaLaunchKernelEx(&cfg, ldg_kernel, (const uint4*)src, S / 16, (const float*)((i & 1) ? xb : xa),
                                          (i & 1) ? xa : xb, pdl));
                });
            }
        }

        {
            struct C2 { int occ, unroll, pdl, pf; uint32_t piece; };
            const C2 c2[] = {{2, 8, 1, 0, 0}};
            for (const C2& c : c2) {
                char label[128];
                snprintf(label, sizeof label, "ldgc grid=148x%d unroll=%d pdl=%d l2prefetch=%d piece=%u", c.occ, c.unroll, c.pdl,
                         c.pf, c.piece);
                run(label, S, [&](const uint8_t* src, int i) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(148 * c.occ);
                    cfg.blockDim = dim3(256);
                    cfg.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = c.pdl;
                    const float* xi = (i & 1) ? xb : xa;
                    float* xo = (i & 1) ? xa : xb;
                    if (c.unroll == 4) CK(cudaLaunchKernelEx(&cfg, ldgc_kernel<4>, src, S, xi, xo, c.pdl, c.pf, c.piece));
                    else if (c.unroll == 8) CK(cudaLaunchKernelEx(&cfg, ldgc_kernel<8>, src, S, xi, xo, c.pdl, c.pf, c.piece));
                    else CK(cudaLaunchKernelEx(&cfg, ldgc_kernel<16>, src, S, xi, xo, c.pdl, c.pf, c.piece));
                });
            }
        }

        {
            struct C3 { int occ, pf, ef; uint32_t piece; };
            const C3 c3[] = {{1, 0, 0, 0}, {2, 0, 0, 0}, {1, 1, 0, 16384}, {1, 1, 1, 16384}, {2, 1, 0, 16384}, {2, 1, 1, 16384}, {2, 1, 1, 65536}};
            const size_t stride = (S + 4095) & ~(size_t)4095;
            for (const C3& c : c3) {
                char label[128];
                snprintf(label, sizeof label, "ldgn grid=148x%d next-layer-L2-prefetch=%d evict_first=%d piece=%u", c.occ, c.pf, c.ef, c.piece);
                run(label, S, [&](const uint8_t* src, int i) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(148 * c.occ);
                    cfg.blockDim = dim3(512);
                    cfg.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = 1;
                    const float* xi = (i & 1) ? xb : xa;
                    float* xo = (i & 1) ? xa : xb;
                    // the last launch of the graph wraps to the pool start in run(); prefetching past it is harmless
                    const size_t nextoff = (c.pf && (size_t)(src - pool) + 2 * stride <= pool_bytes) ? stride : 0;
                    CK(cudaLaunchKernelEx(&cfg, ldgn_kernel<8>, src, S, nextoff, xi, xo, c.piece, c.ef));
                });
            }
        }
        struct Cfg { uint32_t chunk, stages, occ, dynamic, pdl, nprod; };
        const Cfg cfgs[] = {{32768, 4, 1, 0, 1, 1}};
        for (const Cfg& cf : cfgs) {
            const uint32_t nchunks = (uint32_t)(S / cf.chunk);
            const size_t smem = (size_t)cf.stages * cf.chunk + cf.stages * 16 + cf.stages * 4 + 8192 + 64;
            constexpr int NW = 8;
            CK(cudaFuncSetAttribute(bulk_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            char label[128];
            snprintf(label, sizeof label, "bulk chunk=%u stages=%u occ=%u dyn=%u pdl=%u nprod=%u", cf.chunk, cf.stages, cf.occ,
                     cf.dynamic, cf.pdl, cf.nprod);
            run(label, S, [&](const uint8_t* src, int i) {
                BulkParams p;
                p.src = src;
                p.nchunks = nchunks;
                p.chunk = cf.chunk;
                p.stages = cf.stages;
                p.dynamic = cf.dynamic;
                p.pdl = cf.pdl;
                p.nprod = cf.nprod;
                p.counter = counters + i;
                p.xin = (i & 1) ? xb : xa;
                p.xout = (i & 1) ? xa : xb;
                p.tl = (i == L - 8) ? tl : nullptr;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(148 * cf.occ);
                cfg.blockDim = dim3(32 + NW * 32);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at;
                cfg.numAttrs = cf.pdl;
                CK(cudaLaunchKernelEx(&cfg, bulk_kernel<NW>, p));
            });
            // timeline of one launch: spread of per-CTA durations
            std::vector<long long> h(148 * cf.occ * 4);
            CK(cudaMemcpy(h.data(), tl, h.size() * 8, cudaMemcpyDeviceToHost));
            long long tmin = h[0], emax = 0, emin = 1ll << 62, x1max = 0;
            double emean = 0;
            for (uint32_t cI = 0; cI < 148 * cf.occ; ++cI) tmin = std::min(tmin, h[cI * 4]);
            for (uint32_t cI = 0; cI < 148 * cf.occ; ++cI) {
                const long long e = h[cI * 4 + 2] - tmin;
                emax = std::max(emax, e);
                emin = std::min(emin, e);
                emean += (double)e;
                x1max = std::max(x1max, h[cI * 4 + 1] - tmin);
            }
            printf("      per-CTA end: min %lld mean %.0f max %lld ns after first start; x staged by %lld\n", emin,
                   emean / (148 * cf.occ), emax, x1max);
        }
    }
    return 0;
}
