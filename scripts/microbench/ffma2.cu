// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 (sm_100a) issue/throughput.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, -0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            else a[i] = __ffma2_rn(a[i], b, c);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            if (mode == 0) k<0><<<148, threads>>>(out, iters); else k<1><<<148, threads>>>(out, iters);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, threads>>>(out, iters); else k<1><<<148, threads>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * threads * iters * 16.0;
            printf("%s threads/SM=%4d  %.3f ms  %.2f TFMA/s (lane-FMAs)  %.1f FMA/clk/SM @1.9GHz\n", mode ? "FFMA2 " : "FFMA  ",
                   threads, ms, fma / ms / 1e9, fma / ms / 1e6 / 148 / 1.9e3 / 1e3 * 1e3 / 1e3);
        }
    return 0;
}
