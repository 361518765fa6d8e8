// Issue-port behaviour of packed fp32x2 instructions on sm_100a, measured in SM cycles (clock64) per instruction per
// SM sub-partition, for 1..16 resident warps per sub-partition:
//   FFMA2 alone, FFMA alone, LOP3 alone (ALU pipe), FFMA2 + LOP3 1:1, FFMA2 + 2 LOP3, FFMA + LOP3 1:1
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o x2co scripts/fp32x2_coissue.cu && ./x2co
#include <cstdio>
#include <cuda_runtime.h>
template <int NF2, int NF1, int NI>
__global__ void k(float *out, long long *cycles, int iters, float a, float b, unsigned m) {
    float2 x[8], y = make_float2(a, a + 1e-7f), z = make_float2(b, -b);
    float s[8];
    unsigned q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = make_float2(threadIdx.x + u, threadIdx.x - u); s[u] = threadIdx.x * 0.5f + u; q[u] = threadIdx.x * 2654435761u + u; }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (NF2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(*reinterpret_cast<unsigned long long *>(&x[u])) : "l"(*reinterpret_cast<unsigned long long *>(&y)), "l"(*reinterpret_cast<unsigned long long *>(&z)));
                if (NF1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[u]) : "f"(a), "f"(b));
#pragma unroll
                for (int j = 0; j < NI; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[(u + j) & 7]) : "r"(m), "r"(q[(u + j + 3) & 7]));
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) r += x[u].x + x[u].y + s[u] + __uint_as_float(q[u]);
    if (r == 123.456f) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int NF2, int NF1, int NI> void run(const char *name) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    printf("%-22s", name);
    for (int warps = 1; warps <= 4; warps *= 2) {   // warps per sub-partition: one CTA of 4*warps warps per SM
        const int iters = 2048;
        for (int rep = 0; rep < 2; ++rep) k<NF2, NF1, NI><<<148, warps * 128>>>(out, cyc, iters, 1.0000001f, 1e-9f, 0x5bd1e995u);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf(" launch failed"); continue; }
        printf("  %2dw: %5.2f", warps, (double)h / ((double)warps * iters * 32.0));
    }
    printf("   cycles per loop slot\n");
}
int main() {
    run<1, 0, 0>("FFMA2"); run<0, 1, 0>("FFMA"); run<0, 0, 1>("LOP3"); run<1, 0, 1>("FFMA2 + LOP3"); run<1, 0, 2>("FFMA2 + 2 LOP3"); run<1, 0, 3>("FFMA2 + 3 LOP3");
    run<0, 1, 1>("FFMA + LOP3"); run<0, 1, 2>("FFMA + 2 LOP3"); run<1, 1, 0>("FFMA2 + FFMA");
    return 0;
}
