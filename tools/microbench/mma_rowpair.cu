// mma_rowpair.cu -- microbenchmark of the instruction mix of tc::conv_rowpair_kernel (unet_conv_tc.cuh, kernel 5): per filter
// column and K-slice two N = 128 and two N = 64 tcgen05.mma (M = 128, cta_group::1) on shifted views of one halo buffer with
// SBO = 2560 B, the second N = 64 one into accumulator columns [64, 128).  Variants isolate what an instruction of the mix costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rowpair tools/microbench/mma_rowpair.cu
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
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_n(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// variant: 0 = the kernel's mix; 1 = same descriptors, all four N = 128; 2 = the mix, but the last one also into columns [0, 64);
// 3 = per column all eight N = 128 first, then the eight N = 64; 4 = plain N = 128, one A view, SBO 2560; 5 = plain N = 64;
// 6 = the mix with all four A views identical (no shifts); 7 = the mix with SBO 1280; 8 = three N = 128 (a = 0, +1 and a zero-padded third)
__global__ void __launch_bounds__(128, 1) k(int variant, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        const uint32_t i128 = idesc_n(128), i64 = idesc_n(64);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 64 * 1024;
        const uint32_t sbo = variant == 7 ? 1280 : 2560;
        const int pitch = 10 * 8;                 // one halo row in descriptor units
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
                const uint64_t a0 = desc(a_base, sbo) + dxi * 8;
                const uint64_t w0 = desc(b_base + dxi * 24576, 1024);
                if (variant == 3) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { mma(tm, a0 + pitch + 2 * k, w0 + 512 + 2 * k, i128); mma(tm, a0 + 2 * pitch + 2 * k, w0 + 2 * k, i128); }
#pragma unroll
                    for (int k = 0; k < 4; ++k) { mma(tm, a0 + 2 * k, w0 + 1024 + 2 * k, i64); mma(tm + 64, a0 + 3 * pitch + 2 * k, w0 + 2 * k, i64); }
                    continue;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ar = a0 + 2 * k, wk = w0 + 2 * k;
                    if (variant == 4) { for (int q = 0; q < 4; ++q) mma(tm, ar, wk, i128); continue; }
                    if (variant == 5) { for (int q = 0; q < 4; ++q) mma(tm, ar, wk, i64); continue; }
                    if (variant == 8) { mma(tm, ar + pitch, wk + 512, i128); mma(tm, ar + 2 * pitch, wk, i128); mma(tm, ar, wk + 1024, i128); continue; }
                    const int s1 = variant == 6 ? 0 : pitch, s2 = variant == 6 ? 0 : 2 * pitch, s3 = variant == 6 ? 0 : 3 * pitch;
                    mma(tm, ar + s1, wk + 512, i128);
                    mma(tm, ar + s2, wk, i128);
                    mma(tm, ar, wk + 1024, variant == 1 ? i128 : i64);
                    mma(variant == 2 || variant == 1 ? tm : tm + 64, ar + s3, wk, variant == 1 ? i128 : i64);
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        while (!try_wait(&bar, 0)) { if (clock64() - t0 > 6000000000ll) __trap(); }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 300;
    const char* names[] = {"kernel mix (128,128,64,64@+64)", "all four N=128", "mix, last into cols [0,64)", "grouped: 8 x N=128 then 8 x N=64",
                           "plain N=128, one view", "plain N=64, one view", "mix without A shifts", "mix with SBO 1280", "three N=128 (zero-padded third)"};
    printf("variant                                  grid  cycles per filter column and K-slice   cycles per instruction\n");
    for (int v = 0; v < 9; ++v)
        for (int grid : {1, 148}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                k<<<grid, 128, smem>>>(v, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double per_slice = (double)h / (iters * 12.0);
            printf("%-40s %-5d %10.1f %28.1f\n", names[v], grid, per_slice, per_slice / (v == 8 ? 3 : 4)); fflush(stdout);
        }
    return 0;
}
