#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    float2 x[8], y[8], z[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = make_float2(threadIdx.x + u, threadIdx.x - u); y[u] = make_float2(a + 1e-7f * u, a - 1e-7f * u); z[u] = make_float2(b * (u + 1), b * (u + 2)); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (OP == 0) x[u] = __ffma2_rn(x[u], y[(u + r) & 7], z[(u + 3 * r) & 7]);
                if (OP == 1) x[u] = __fmul2_rn(x[u], y[(u + r) & 7]);
                if (OP == 2) x[u] = __fadd2_rn(x[u], z[(u + 3 * r) & 7]);
                if (OP == 3) { x[u].x = fmaf(x[u].x, y[(u + r) & 7].x, z[(u + 3 * r) & 7].x); x[u].y = fmaf(x[u].y, y[(u + r) & 7].y, z[(u + 3 * r) & 7].y); }
                if (OP == 4) { x[u].x = x[u].x * y[(u + r) & 7].x; x[u].y = x[u].y * y[(u + r) & 7].y; }
                if (OP == 5) { x[u].x = x[u].x + z[(u + 3 * r) & 7].x; x[u].y = x[u].y + z[(u + 3 * r) & 7].y; }
                if (OP == 6) x[u] = __fadd2_rn(make_float2(y[(u + r) & 7].x, y[(u + r) & 7].x), x[u]);   // scalar broadcast + packed (the walk's c - p)
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) r += x[u].x + x[u].y;
    if (r == 123.456f) out[0] = r;
}
template <int OP> void run(const char *name) {
    float *out; cudaMalloc(&out, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * 8, iters = 4096; float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) { cudaEventRecord(a); k<OP><<<grid, 256>>>(out, iters, 1.0000001f, 1e-9f); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms; }
    const double laneops = 2.0 * 64.0 * iters * 256.0 * grid;   // fp32 lane-operations
    printf("%-28s %.3f ms  %.2f T lane-ops/s  (%.1f%% of 37.2)\n", name, best, laneops / (best * 1e-3) / 1e12, 100 * laneops / (best * 1e-3) / 1e12 / 37.2);
}
int main() { run<0>("FFMA2"); run<1>("FMUL2"); run<2>("FADD2"); run<6>("FADD2 scalar-broadcast"); run<3>("2x FFMA (scalar)"); run<4>("2x FMUL (scalar)"); run<5>("2x FADD (scalar)"); return 0; }
