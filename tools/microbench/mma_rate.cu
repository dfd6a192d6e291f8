// mma_rate.cu -- microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128 per CTA) as a function of N,
// cta_group and the A descriptor geometry (aligned SBO=1024 vs halo-style shifted start with SBO=1280).
// Operands are whatever is in shared memory (values do not matter for timing; zero-initialised).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rate tools/microbench/mma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}

template <int CG>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int a_off, int sbo, int n_mma_per_commit, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tptr;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24);
        const uint32_t a_base = smem_u32(smem) + a_off, b_base = smem_u32(smem) + 64 * 1024;
        uint32_t parity = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int m = 0; m < n_mma_per_commit; ++m) {
                // walk 9 shifted A views x 4 K-slices like the conv kernel does
                const int tap = (m >> 2) % 9, ks = m & 3;
                const uint64_t ad = desc(a_base + ((tap / 3) * (sbo / 128) + tap % 3) * 128, sbo) + 2 * ks;
                const uint64_t bd = desc(b_base + (tap * 4 + ks) * 0 + (m % 36) * 0, 1024) + 2 * ks;
                if (CG == 1)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        // one commit for everything issued, then a bounded wait (a bug traps instead of hanging the GPU)
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        while (!try_wait(&bar, parity)) { if (clock64() - t0 > 6000000000ll) __trap(); }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (threadIdx.x < 32) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 200, per = 36;
    printf("cg  N    a_off sbo   grid  cycles/MMA  (MMA = M128/CTA x N x K16)\n");
    for (int cg = 1; cg <= 2; ++cg)
        for (int N : {64, 128, 256})
            for (int geom = 0; geom < 3; ++geom)
                for (int grid : {cg, 148}) {
                    const int a_off = geom == 0 ? 0 : 0, sbo = geom == 0 ? 1024 : (geom == 1 ? 1280 : 2048);
                    long long h = 0;
                    for (int rep = 0; rep < 2; ++rep) {
                        if (cg == 1) k<1><<<grid, 128, smem>>>(N, iters, a_off, sbo, per, d);
                        else {
                            cudaLaunchConfig_t cfg{};
                            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                            cudaLaunchAttribute at[1];
                            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                            cfg.attrs = at; cfg.numAttrs = 1;
                            cudaLaunchKernelEx(&cfg, k<2>, N, iters, a_off, sbo, per, d);
                        }
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    printf("%d   %-4d %-5d %-5d %-5d %8.1f\n", cg, N, a_off, sbo, grid, (double)h / (iters * per)); fflush(stdout);
                }
    return 0;
}
