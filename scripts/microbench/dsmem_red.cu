// Microbenchmark: rate of reductions into the shared memory of the OTHER CTAs of a thread-block cluster
// (red.shared::cluster.add.u32 through mapa), random cells over the whole cluster: the primitive a per-roach
// [253][4096] pulse-height histogram held in distributed shared memory would be built from (the L2 reduction unit,
// which the 4096-bin histogram of K6 is bound by, does 198 G reductions per second: profiles/r01_atomics_microbench.txt).
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dsmem_red.bin dsmem_red.cu && ./dsmem_red.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rnd(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// MODE 0: reductions into random cells of the whole cluster; MODE 1: only into this CTA's own cells (plain ATOMS);
// MODE 2: as 0, but the lanes of a warp all aim at ONE target CTA per instruction (different cells)
template <int MODE>
__global__ void k(int iters, int cells_per_cta, int cs, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < cells_per_cta; i += blockDim.x) sm[i] = 0;
    cluster_sync();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t sw = (threadIdx.x >> 5) * 977u + blockIdx.x * 131u + 7;
    for (int i = 0; i < iters; ++i) {
        const uint32_t cell = rnd(s) % (uint32_t)cells_per_cta;
        uint32_t rank = MODE == 1 ? cluster_ctarank() : rnd(s) % (uint32_t)cs;
        if (MODE == 2) rank = rnd(sw) % (uint32_t)cs;
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + cell * 4u), "r"(rank));
        asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], 1;" ::"r"(ra) : "memory");
    }
    cluster_sync();
    // checksum: every reduction must have landed somewhere in the cluster
    uint32_t t = 0;
    for (int i = threadIdx.x; i < cells_per_cta; i += blockDim.x) t += sm[i];
    atomicAdd(sink, t);
}

template <int MODE>
static void run(const char *name, int cs, int cells_per_cta, int threads, int iters, uint32_t *sink) {
    const size_t smem = (size_t)cells_per_cta * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cs);
    int max_clusters = 0;
    cudaOccupancyMaxActiveClusters(&max_clusters, k<MODE>, &cfg);
    if (max_clusters < 1) { printf("%-34s cluster %2d: not launchable (%s)\n", name, cs, cudaGetErrorString(cudaGetLastError())); return; }
    cfg.gridDim = dim3(cs * max_clusters);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaMemset(sink, 0, 4);
    cudaLaunchKernelEx(&cfg, k<MODE>, iters, cells_per_cta, cs, sink);
    cudaDeviceSynchronize();
    cudaMemset(sink, 0, 4);
    cudaEventRecord(a);
    cudaLaunchKernelEx(&cfg, k<MODE>, iters, cells_per_cta, cs, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    uint32_t got = 0; cudaMemcpy(&got, sink, 4, cudaMemcpyDeviceToHost);
    const double ops = (double)cs * max_clusters * threads * iters;
    printf("%-34s cluster %2d x %2d active, %3d KiB/CTA, %4d thr: %.3f ms  %7.1f Gop/s  (%.2f ns/op/SM)  sum %s  err=%s\n", name, cs,
           max_clusters, (int)(smem >> 10), threads, ms, ops / ms / 1e6, ms * 1e6 / ((double)threads * iters),
           got == (uint32_t)ops ? "ok" : "WRONG", cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *sink; cudaMalloc(&sink, 4);
    const int iters = 2048;
    for (int threads : {512, 1024}) {
        run<1>("ATOMS own CTA (through mapa)", 8, 224 * 256, threads, iters, sink);
        run<0>("RED random CTA of the cluster", 2, 224 * 256, threads, iters, sink);
        run<0>("RED random CTA of the cluster", 4, 224 * 256, threads, iters, sink);
        run<0>("RED random CTA of the cluster", 8, 224 * 256, threads, iters, sink);
        run<0>("RED random CTA of the cluster", 16, 128 * 256, threads, iters, sink);
        run<0>("RED random CTA of the cluster", 16, 192 * 256, threads, iters, sink);
        run<2>("RED one target CTA per warp instr", 8, 224 * 256, threads, iters, sink);
        run<2>("RED one target CTA per warp instr", 16, 128 * 256, threads, iters, sink);
    }
    return 0;
}
