// mma_rowpair_order.cu -- which ORDER of the 16 instructions of one filter column (4 A views x 4 K-slices; views a = 0, +1 are
// N = 128, a = -1, +2 are N = 64) does the tensor pipe run fastest?  Same descriptors as tc::conv_rowpair_kernel; see mma_rowpair.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rowpair_order tools/microbench/mma_rowpair_order.cu
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

// instruction v (view): 0: a = 0 (N128, A row +1, B tile 1), 1: a = +1 (N128, A row +2, B tile 0), 2: a = -1 (N64, A row 0, B tile 2),
// 3: a = +2 (N64 into cols 64.., A row +3, B tile 0)
struct Ins { int v, k; };
constexpr int NORD = 8;
__host__ __device__ constexpr Ins order_at(int ord, int i) {
    switch (ord) {
        case 0: return Ins{i & 3, i >> 2};                                   // k outer, views inner (the first kernel version)
        case 1: return Ins{i < 8 ? (i & 1) : 2 + (i & 1), (i >> 1) & 3};    // N128 pair over k, then N64 pair over k
        case 2: return Ins{i >> 2, i & 3};                                   // view outer, k inner
        case 3: return Ins{(i >> 2) == 0 ? 1 : (i >> 2) == 1 ? 3 : (i >> 2) == 2 ? 0 : 2, i & 3};   // view outer, B-sharing views adjacent
        case 4: return Ins{i < 8 ? 2 + (i & 1) : (i & 1), (i >> 1) & 3};    // N64 pair over k first, then the N128 pair
        case 5: return Ins{i < 8 ? (i >> 2) : 2 + ((i >> 2) & 1), i & 3};    // = order 2 (control)
        case 6: return Ins{(i & 1) ? ((i >> 1) & 1 ? 3 : 2) : ((i >> 1) & 1 ? 1 : 0), i >> 2};   // k outer: 128, 64, 128, 64
        default: return Ins{i < 8 ? (i & 1) : 2 + (i & 1), i < 8 ? (i >> 1) : ((i - 8) >> 1)};   // = order 1
    }
}

template <int ORD>
__global__ void __launch_bounds__(128, 1) k(int iters, int sbo, long long* out) {
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
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 64 * 1024;
        const int pitch = 10 * 8;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
                const uint64_t a0 = desc(a_base, sbo) + dxi * 8;
                const uint64_t w0 = desc(b_base + dxi * 24576, 1024);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    constexpr int dummy = 0; (void)dummy;
                    const Ins in = order_at(ORD, i);
                    const int arow = in.v == 0 ? 1 : in.v == 1 ? 2 : in.v == 2 ? 0 : 3;
                    const int btile = in.v == 0 ? 1 : in.v == 2 ? 2 : 0;
                    mma(in.v == 3 ? tm + 64 : tm, a0 + arow * pitch + 2 * in.k, w0 + btile * 512 + 2 * in.k, in.v < 2 ? idesc_n(128) : idesc_n(64));
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

template <int ORD>
void run(long long* d, const char* name) {
    const int smem = 200 * 1024, iters = 300;
    cudaFuncSetAttribute(k<ORD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int sbo : {2560, 1280, 2048})
        for (int grid : {1, 148}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                k<ORD><<<grid, 128, smem>>>(iters, sbo, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
            }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("%-52s sbo %-5d grid %-4d %8.1f cycles per filter column (16 instr) = %6.1f per tile chunk of 48\n", name, sbo, grid,
                   (double)h / (iters * 3.0), (double)h / iters);
            fflush(stdout);
        }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    run<0>(d, "0: k outer, views (0,+1,-1,+2) inner");
    run<1>(d, "1: (0,+1) over k, then (-1,+2) over k");
    run<2>(d, "2: view outer (0,+1,-1,+2), k inner");
    run<3>(d, "3: view outer (+1,+2,0,-1), k inner");
    run<4>(d, "4: (-1,+2) over k, then (0,+1) over k");
    run<6>(d, "6: k outer, (0,-1,+1,+2) inner");
    return 0;
}
