// Issue throughput of the ALU-pipe instructions the GEMV decode uses (test infrastructure).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int KIND>
__global__ void k(int iters, long long* out, uint32_t* sink) {
    uint32_t r[16];
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * (i + 3) + 12345;
    uint32_t a = threadIdx.x * 7 + 1, b = threadIdx.x * 13 + 5;
    __syncthreads();
    long long t0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (KIND == 0) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (KIND == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (KIND == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
            if (KIND == 4) { if (i & 1) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)); }
            if (KIND == 5) { if (i & 1) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b)); }
            if (KIND == 6) asm volatile("prmt.b32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));  // selector in a register, table operands fixed
            if (KIND == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
        }
    }
    __syncthreads(); long long t1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
    uint32_t s = 0;
    for (int i = 0; i < 16; ++i) s ^= r[i];
    if (s == 0x1234567) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
}
int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
    const char* names[] = {"PRMT", "LOP3", "IMAD", "IMAD.HI", "PRMT+IMAD alternating", "PRMT+LOP3 alternating", "PRMT (reg selector)", "SHF"};
    const int iters = 2000;
    for (int kind = 0; kind < 8; ++kind) {
        for (int warps : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (kind) {
                    case 0: k<0><<<148, warps * 32>>>(iters, out, sink); break;
                    case 1: k<1><<<148, warps * 32>>>(iters, out, sink); break;
                    case 2: k<2><<<148, warps * 32>>>(iters, out, sink); break;
                    case 3: k<3><<<148, warps * 32>>>(iters, out, sink); break;
                    case 4: k<4><<<148, warps * 32>>>(iters, out, sink); break;
                    case 5: k<5><<<148, warps * 32>>>(iters, out, sink); break;
                    case 6: k<6><<<148, warps * 32>>>(iters, out, sink); break;
                    case 7: k<7><<<148, warps * 32>>>(iters, out, sink); break;
                }
                cudaDeviceSynchronize();
            }
            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("%-24s warps/SM=%2d  %.2f cycles per warp-instruction per sub-partition\n", names[kind], warps,
                   (double)h / (iters * 16.0 * (warps / 4)));
        }
    }
    return 0;
}
