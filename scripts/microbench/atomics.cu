// Microbenchmark: throughput of the primitives a per-pixel histogram can be built from on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t rnd(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void k(uint32_t *g, int iters, int nb, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    uint32_t acc = 0;
    uint8_t *sb = reinterpret_cast<uint8_t *>(sm);
    for (int i = 0; i < iters; ++i) {
        uint32_t key = rnd(s) % nb;
        if (MODE == 0) atomicAdd(&sm[key], 1u);                       // smem atomics, spread
        if (MODE == 1) atomicAdd(&g[(size_t)(rnd(s) & 0x7fffff)], 1u); // global REDs over 32 MB
        if (MODE == 2) { sb[key * blockDim.x + threadIdx.x] += 1; }    // thread-private u8 counters (nb*blockDim <= 64K)
        if (MODE == 3) { unsigned m = __match_any_sync(0xffffffffu, key); acc += m; }
        if (MODE == 4) { unsigned m = __match_any_sync(0xffffffffu, key); int leader = __ffs(m) - 1;
                         if ((threadIdx.x & 31) == leader) sm[(threadIdx.x >> 5) * 256 + key] += __popc(m); }
        if (MODE == 5) atomicAdd(&g[blockIdx.x * 4096 + key], 1u);     // global REDs, per-CTA 253 addresses (L2-hot)
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = sm[1] + acc;
}
int main() {
    uint32_t *g, *sink; cudaMalloc(&g, 64 << 20); cudaMemset(g, 0, 64 << 20); cudaMalloc(&sink, 1 << 20);
    const int iters = 4096, grid = 148;
    const char *names[] = {"ATOMS spread(253)", "REDG spread(32MB)", "private u8 RMW", "MATCH.ANY only", "match+warp-private RMW", "REDG per-CTA 253"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int threads : {256, 512, 1024}) {
            if (mode == 2 && threads * 253 > 65536) continue;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            auto run = [&]() {
                switch (mode) {
                case 0: k<0><<<grid, threads, 65536>>>(g, iters, 253, sink); break;
                case 1: k<1><<<grid, threads, 65536>>>(g, iters, 253, sink); break;
                case 2: k<2><<<grid, threads, 65536>>>(g, iters, 253, sink); break;
                case 3: k<3><<<grid, threads, 65536>>>(g, iters, 253, sink); break;
                case 4: k<4><<<grid, threads, 65536>>>(g, iters, 253, sink); break;
                case 5: k<5><<<grid, threads, 65536>>>(g, iters, 253, sink); break;
                }
            };
            cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            run(); cudaDeviceSynchronize();
            cudaEventRecord(a); run(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double ops = (double)grid * threads * iters;
            printf("%-26s threads=%4d  %.3f ms  %.1f Gop/s  (%.2f ns/op/SM)  err=%s\n", names[mode], threads, ms,
                   ops / ms / 1e6, ms * 1e6 / (threads * (double)iters), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
