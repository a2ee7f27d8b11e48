// Throughput of legacy mma.sync flavours on sm_100a (test infrastructure): cycles per MMA per SM sub-partition.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int KIND>
__global__ void k(int iters, long long* out, int* sink) {
    int d[8][4];
    float f[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) { d[i][j] = 0; f[i][j] = 0.f; }
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    __syncthreads();
    long long t0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[i][0]), "+r"(d[i][1]), "+r"(d[i][2]), "+r"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1) : "memory");
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1) : "memory");
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1) : "memory");
            else if (KIND == 3)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(d[i][0]), "+r"(d[i][1]), "+r"(d[i][2]), "+r"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0) : "memory");
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1) : "memory");
        }
    }
    __syncthreads(); long long t1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
    int s = 0; float fs = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) { s += d[i][j]; fs += f[i][j]; }
    if (s + (int)fs == 12345) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
}
// Does an IMMA overlap with ALU-pipe work of the same sub-partition?  One IMMA + NALU PRMTs per iteration slot:
// overlapped -> max(8, 2*NALU) cycles, serialised -> 8 + 2*NALU.
template <int NALU>
__global__ void mix(int iters, long long* out, int* sink) {
    int d[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    uint32_t p[4] = {a0 * 17, a0 * 19, a0 * 23, a0 * 29};
    __syncthreads();
    long long t0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(d[i][0]), "+r"(d[i][1]), "+r"(d[i][2]), "+r"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
#pragma unroll
            for (int q = 0; q < NALU; ++q)  // four independent chains
                asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(p[q & 3]) : "r"(b0));
        }
    }
    __syncthreads(); long long t1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
    int s = p[0] + p[1] + p[2] + p[3];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    if (s == 12345) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
}
template <int NALU>
static void run_mix(long long* out, int* sink) {
    const int iters = 2000;
    for (int warps : {4, 16}) {
        for (int rep = 0; rep < 2; ++rep) { mix<NALU><<<148, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize(); }
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("IMMA + %2d PRMT  warps/SM=%2d  %.2f cycles per (IMMA + PRMTs) per sub-partition\n", NALU, warps,
               (double)h / (iters * 8.0 * (warps / 4)));
    }
}
int main() {
    long long* out; int* sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
    const char* names[] = {"IMMA m16n8k32 u8.s8", "HMMA m16n8k16 f16", "QMMA m16n8k32 e4m3", "IMMA m16n8k16 u8.s8", "HMMA m16n8k16 bf16"};
    const int iters = 2000;
    for (int kind = 0; kind < 5; ++kind) {
        for (int warps : {4, 8, 16}) {  // per CTA (= per SM): 1, 2, 4 warps per sub-partition
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) k<0><<<148, warps * 32>>>(iters, out, sink);
                if (kind == 1) k<1><<<148, warps * 32>>>(iters, out, sink);
                if (kind == 2) k<2><<<148, warps * 32>>>(iters, out, sink);
                if (kind == 3) k<3><<<148, warps * 32>>>(iters, out, sink);
                if (kind == 4) k<4><<<148, warps * 32>>>(iters, out, sink);
                { cudaError_t e = cudaDeviceSynchronize(); if (e) printf("sync err %s\n", cudaGetErrorString(e)); }
            }
            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            double per_smsp = (double)h / (iters * 8.0 * (warps / 4)); printf("raw=%lld ", h);
            printf("%-22s warps/SM=%2d  %.2f cycles per MMA per sub-partition (err=%s)\n", names[kind], warps, per_smsp, cudaGetErrorString(cudaGetLastError()));
        }
    }
    run_mix<0>(out, sink); run_mix<2>(out, sink); run_mix<4>(out, sink); run_mix<8>(out, sink); run_mix<14>(out, sink);
    return 0;
}
