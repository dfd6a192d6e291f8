// preprocess.cu -- K1: RAW u16 slice -> 512x512 u8 (+ optional bf16 = u8/255).
//
// Replaces Preprocess::compute_minmax + the resample loop of Preprocess::preprocess_raw
// (/root/reference/src/preprocess.cpp:65-74, 81-118) and MedicalSeg::preprocess_image
// (/root/reference/src/process.cpp:36-39).  Two launches per *batch* (the global min/max is a
// grid-wide dependency, SURVEY.md hard part H7):
//   minmax_kernel   : one streaming pass, 16-byte loads, packed-u16 SIMD min/max, one atomic pair
//                     per block.  Algorithmic bytes: 2*w*h per slice (read once from HBM).
//   resample_kernel : 4 output pixels per thread, bilinear in fp64 with *explicitly unfused*
//                     multiplies/adds (__dmul_rn/__dadd_rn) so the result is bit-identical to the
//                     reference's double arithmetic, u8 written as one 32-bit store, bf16 as one
//                     64-bit store.  The source rows it gathers were just streamed by minmax_kernel
//                     and sit in L2 (a 32-slice batch of 512x512 is 16 MiB << 126 MB).
#include "common.cuh"
#include <cstdlib>

namespace ms {

namespace {

// mm[2*b] = 0xFFFF - min, mm[2*b+1] = max  (both grow under atomicMax, so one memset(0) initialises)
__global__ void __launch_bounds__(256) minmax_kernel(const uint16_t* __restrict__ src, size_t n_per_slice,
                                                      uint32_t* __restrict__ mm) {
    const int b = blockIdx.y;
    const uint16_t* s = src + (size_t)b * n_per_slice;
    uint32_t vmin = 0xFFFFFFFFu, vmax = 0u;  // two packed u16 lanes
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * blockDim.x;

    // head (unaligned prefix), vector body, tail
    const uintptr_t addr = reinterpret_cast<uintptr_t>(s);
    size_t head = ((16 - (addr & 15)) & 15) / 2;
    if (head > n_per_slice) head = n_per_slice;
    const size_t nvec = (n_per_slice - head) / 8;
    const uint4* v = reinterpret_cast<const uint4*>(s + head);
    for (size_t i = tid; i < nvec; i += nthr) {
        uint4 q = __ldg(v + i);
        vmin = __vminu2(vmin, __vminu2(__vminu2(q.x, q.y), __vminu2(q.z, q.w)));
        vmax = __vmaxu2(vmax, __vmaxu2(__vmaxu2(q.x, q.y), __vmaxu2(q.z, q.w)));
    }
    uint32_t mn = min(vmin & 0xFFFFu, vmin >> 16), mx = max(vmax & 0xFFFFu, vmax >> 16);
    const size_t tail0 = head + nvec * 8;
    for (size_t i = tid; i < head; i += nthr) { uint32_t x = s[i]; mn = min(mn, x); mx = max(mx, x); }
    for (size_t i = tail0 + tid; i < n_per_slice; i += nthr) { uint32_t x = s[i]; mn = min(mn, x); mx = max(mx, x); }

    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    __shared__ uint32_t smn[8], smx[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        mn = lane < (blockDim.x >> 5) ? smn[lane] : 0xFFFFu;
        mx = lane < (blockDim.x >> 5) ? smx[lane] : 0u;
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if (lane == 0) {
            atomicMax(&mm[2 * b], 0xFFFFu - mn);
            atomicMax(&mm[2 * b + 1], mx);
        }
    }
}

// One thread = 4 consecutive output pixels of one row.
__global__ void __launch_bounds__(128) resample_kernel(const uint16_t* __restrict__ src, int w, int h, int out_w, int out_h,
                                                        const uint32_t* __restrict__ mm, uint8_t* __restrict__ out_u8,
                                                        __nv_bfloat16* __restrict__ out_bf16) {
    const int b = blockIdx.z;
    const int y = blockIdx.y;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= out_w) return;
    const uint16_t* s = src + (size_t)b * w * h;

    int mn = 0xFFFF - (int)mm[2 * b], mx = (int)mm[2 * b + 1];
    if (mn == mx) mx = mn + 1;                                   // preprocess.cpp:92
    const double scale8 = __ddiv_rn(255.0, (double)(mx - mn));   // :93
    const double step_x = __ddiv_rn((double)w, (double)out_w);   // :82
    const double step_y = __ddiv_rn((double)h, (double)out_h);   // :83

    const double fy = __dmul_rn((double)y, step_y);              // :100
    const int iy = (int)fy;                                      // :102
    const int iy1 = min(iy + 1, h - 1);                          // :104
    const double dy = __dsub_rn(fy, (double)iy);                 // :105
    const double ody = __dsub_rn(1.0, dy);
    const uint16_t* r0 = s + (size_t)iy * w;
    const uint16_t* r1 = s + (size_t)iy1 * w;

    uint32_t packed = 0;
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = x0 + j;
        uint32_t q = 0;
        if (x < out_w) {
            const double fx = __dmul_rn((double)x, step_x);
            const int ix = (int)fx;                              // :101
            const int ix1 = min(ix + 1, w - 1);                  // :103
            const double dx = __dsub_rn(fx, (double)ix);
            const double odx = __dsub_rn(1.0, dx);
            const double v00 = (double)__ldg(r0 + ix), v01 = (double)__ldg(r0 + ix1);
            const double v10 = (double)__ldg(r1 + ix), v11 = (double)__ldg(r1 + ix1);
            // :112-115, evaluated left to right without contraction
            double v = __dmul_rn(__dmul_rn(odx, ody), v00);
            v = __dadd_rn(v, __dmul_rn(__dmul_rn(dx, ody), v01));
            v = __dadd_rn(v, __dmul_rn(__dmul_rn(odx, dy), v10));
            v = __dadd_rn(v, __dmul_rn(__dmul_rn(dx, dy), v11));
            const double qd = __dadd_rn(__dmul_rn(__dsub_rn(v, (double)mn), scale8), 0.5);  // :116
            q = (uint32_t)__double2int_rz(qd) & 0xFFu;           // static_cast<uchar>
        }
        packed |= q << (8 * j);
        f[j] = __fdiv_rn((float)q, 255.0f);                      // process.cpp:38
    }
    const size_t o = ((size_t)b * out_h + y) * out_w + x0;
    if (x0 + 3 < out_w && (out_w & 3) == 0) {
        *reinterpret_cast<uint32_t*>(out_u8 + o) = packed;
        if (out_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]);
            __nv_bfloat162 hi = __floats2bfloat162_rn(f[2], f[3]);
            uint2 st;
            st.x = *reinterpret_cast<uint32_t*>(&lo);
            st.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(out_bf16 + o) = st;
        }
    } else {
        for (int j = 0; j < 4 && x0 + j < out_w; ++j) {
            out_u8[o + j] = (uint8_t)(packed >> (8 * j));
            if (out_bf16) out_bf16[o + j] = __float2bfloat16_rn(f[j]);
        }
    }
}

// Fast path for the identity geometry (w == out_w, h == out_h; BASELINE cfg1-3): step = 1.0, so dx = dy = 0 and the
// reference's bilinear expression collapses to v00 exactly (the other three products are +0.0).  One thread =
// 8 consecutive pixels: one 16-byte load, one 8-byte u8 store, one 16-byte bf16 store.
__global__ void __launch_bounds__(256) normalise_identity_kernel(const uint16_t* __restrict__ src, size_t n_per_slice,
                                                                  const uint32_t* __restrict__ mm, uint8_t* __restrict__ out_u8,
                                                                  __nv_bfloat16* __restrict__ out_bf16) {
    const int b = blockIdx.y;
    const size_t i8 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i8 * 8 >= n_per_slice) return;
    int mn = 0xFFFF - (int)mm[2 * b], mx = (int)mm[2 * b + 1];
    if (mn == mx) mx = mn + 1;
    const double scale8 = __ddiv_rn(255.0, (double)(mx - mn));
    const size_t o = (size_t)b * n_per_slice + i8 * 8;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + o));
    const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
    uint32_t lo = 0, hi = 0;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int v = (int)((wds[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
        const double qd = __dadd_rn(__dmul_rn((double)(v - mn), scale8), 0.5);   // preprocess.cpp:116 with v == v00
        const uint32_t r = (uint32_t)__double2int_rz(qd) & 0xFFu;
        if (j < 4) lo |= r << (8 * j); else hi |= r << (8 * (j - 4));
        f[j] = __fdiv_rn((float)r, 255.0f);
    }
    *reinterpret_cast<uint2*>(out_u8 + o) = make_uint2(lo, hi);
    if (out_bf16) {
        uint4 st;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(f[0], f[1]); st.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(f[2], f[3]); st.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(f[4], f[5]); st.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(f[6], f[7]); st.w = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint4*>(out_bf16 + o) = st;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One kernel for the identity geometry (BASELINE cfg1-3): a CLUSTER of 8 CTAs per slice.  Every thread loads its share of
// the slice ONCE into registers (16 x 16 bytes, all loads in flight together), the slice's min / max is reduced through the
// warps, the CTA and then the cluster's distributed shared memory, and the normalised pixels are written from the same
// registers: the u16 slice crosses HBM exactly once (the two-kernel form reads it twice, the second time from L2, and pays
// a memset + two launches).  Algorithmic bytes 2 w h + 2 * 512^2 per slice; actual traffic 2 w h + 512^2 (u8 out).
constexpr int kK1Cluster = 8;     // 8 CTAs x THREADS x VEC x 8 px = 262,144 px
__device__ __forceinline__ uint32_t k1_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k1_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t k1_ld_peer(uint32_t smem_addr, uint32_t rank) {
    uint32_t a, v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_addr), "r"(rank));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
template <int kK1Threads, int kK1Vec>
__global__ void __cluster_dims__(kK1Cluster, 1, 1) __launch_bounds__(kK1Threads)
normalise_cluster_kernel(const uint16_t* __restrict__ src, int n_vec_per_slice /* uint4 per slice */, uint8_t* __restrict__ out_u8,
                         __nv_bfloat16* __restrict__ out_bf16) {
    __shared__ uint32_t s_warp[2][kK1Threads / 32];
    __shared__ uint32_t s_mm[2];
    const int b = blockIdx.x / kK1Cluster, rank = blockIdx.x % kK1Cluster;
    const uint4* v = reinterpret_cast<const uint4*>(src) + (size_t)b * n_vec_per_slice;
    // thread t of CTA r owns vectors r * 256 + t + k * 2048, k = 0 .. 15 (warp-contiguous 512-byte segments)
    const int first = rank * kK1Threads + threadIdx.x;
    uint4 q[kK1Vec];
    uint32_t vmin = 0xFFFFFFFFu, vmax = 0u;
#pragma unroll
    for (int k = 0; k < kK1Vec; ++k) {
        const int i = first + k * kK1Cluster * kK1Threads;
        q[k] = i < n_vec_per_slice ? __ldg(v + i) : make_uint4(0xFFFF0000u, 0xFFFF0000u, 0xFFFF0000u, 0xFFFF0000u);   // neutral: {0, 65535}
    }
#pragma unroll
    for (int k = 0; k < kK1Vec; ++k) {
        const int i = first + k * kK1Cluster * kK1Threads;
        if (i < n_vec_per_slice) {
            vmin = __vminu2(vmin, __vminu2(__vminu2(q[k].x, q[k].y), __vminu2(q[k].z, q[k].w)));
            vmax = __vmaxu2(vmax, __vmaxu2(__vmaxu2(q[k].x, q[k].y), __vmaxu2(q[k].z, q[k].w)));
        }
    }
    uint32_t mn = min(vmin & 0xFFFFu, vmin >> 16), mx = max(vmax & 0xFFFFu, vmax >> 16);
    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_warp[0][warp] = mn; s_warp[1][warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        mn = lane < kK1Threads / 32 ? s_warp[0][lane] : 0xFFFFu;
        mx = lane < kK1Threads / 32 ? s_warp[1][lane] : 0u;
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if (lane == 0) { s_mm[0] = mn; s_mm[1] = mx; }
    }
    k1_cluster_sync();                         // every CTA's {min, max} is published
    {
        const uint32_t a = k1_smem_u32(s_mm);
        uint32_t m0 = 0xFFFFu, m1 = 0u;
#pragma unroll
        for (int r = 0; r < kK1Cluster; ++r) {
            m0 = min(m0, k1_ld_peer(a, r));
            m1 = max(m1, k1_ld_peer(a + 4, r));
        }
        mn = m0;
        mx = m1;
    }
    k1_cluster_sync();                         // no CTA leaves (or reuses s_mm) while a peer may still read it
    int imn = (int)mn, imx = (int)mx;
    if (imn == imx) imx = imn + 1;                                  // preprocess.cpp:92
    const double scale8 = __ddiv_rn(255.0, (double)(imx - imn));   // :93
#pragma unroll
    for (int k = 0; k < kK1Vec; ++k) {
        const int i = first + k * kK1Cluster * kK1Threads;
        if (i >= n_vec_per_slice) continue;
        const uint32_t wds[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
        uint32_t lo = 0, hi = 0;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int px = (int)((wds[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
            const double qd = __dadd_rn(__dmul_rn((double)(px - imn), scale8), 0.5);   // :116 with v == v00 (identity geometry)
            const uint32_t r = (uint32_t)__double2int_rz(qd) & 0xFFu;
            if (j < 4) lo |= r << (8 * j); else hi |= r << (8 * (j - 4));
            f[j] = __fdiv_rn((float)r, 255.0f);                                        // process.cpp:38
        }
        const size_t o = ((size_t)b * n_vec_per_slice + i) * 8;
        *reinterpret_cast<uint2*>(out_u8 + o) = make_uint2(lo, hi);
        if (out_bf16) {
            uint4 st;
            __nv_bfloat162 t;
            t = __floats2bfloat162_rn(f[0], f[1]); st.x = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(f[2], f[3]); st.y = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(f[4], f[5]); st.z = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(f[6], f[7]); st.w = *reinterpret_cast<uint32_t*>(&t);
            *reinterpret_cast<uint4*>(out_bf16 + o) = st;
        }
    }
}

}  // namespace

void preprocess_launch(PreprocessWs& ws, const uint16_t* d_src, int w, int h, int batch, int out_w, int out_h,
                       uint8_t* d_out_u8, __nv_bfloat16* d_out_bf16, cudaStream_t st) {
    MS_REQUIRE(w > 0 && h > 0 && batch > 0 && out_w > 0 && out_h > 0, MS_ERR_ARG, "preprocess: bad shape");
    MS_REQUIRE((int64_t)w * h < (int64_t)1 << 31, MS_ERR_ARG, "preprocess: slice too large");
    const size_t n = (size_t)w * h;
    {
        const bool aligned = (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out_u8) & 7) == 0 &&
                             (reinterpret_cast<uintptr_t>(d_out_bf16) & 15) == 0 && n % 8 == 0;
        const char* e = std::getenv("MEDSEG_K1_CLUSTER");
        if (w == out_w && h == out_h && aligned && n / 8 <= (size_t)kK1Cluster * 256 * 16 && !(e && e[0] == '0')) {
            // 512 threads x 8 vectors: 3.98 TB/s of algorithmic bytes at batch 256 (256 x 16: 3.52; two kernels: 3.22)
            normalise_cluster_kernel<512, 8><<<kK1Cluster * batch, 512, 0, st>>>(d_src, (int)(n / 8), d_out_u8, d_out_bf16);
            MS_LAUNCH_CHECK();
            return;
        }
    }
    ws.minmax.reserve((size_t)batch * 2 * sizeof(uint32_t));
    uint32_t* mm = ws.minmax.as<uint32_t>();
    MS_CUDA(cudaMemsetAsync(mm, 0, (size_t)batch * 2 * sizeof(uint32_t), st));
    // enough blocks per slice to fill 148 SMs across the batch, 16 B per thread per iteration
    int bx = (int)std::min<size_t>((n / 8 + 255) / 256, (size_t)std::max(1, (148 * 8 + batch - 1) / batch));
    bx = std::max(bx, 1);
    minmax_kernel<<<dim3(bx, batch), 256, 0, st>>>(d_src, n, mm);
    MS_LAUNCH_CHECK();
    const bool aligned = (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out_u8) & 7) == 0 &&
                         (reinterpret_cast<uintptr_t>(d_out_bf16) & 15) == 0 && n % 8 == 0;
    if (w == out_w && h == out_h && aligned) {
        normalise_identity_kernel<<<dim3((unsigned)cdiv64((int64_t)(n / 8), 256), batch), 256, 0, st>>>(d_src, n, mm, d_out_u8, d_out_bf16);
    } else {
        dim3 grid(cdiv(cdiv(out_w, 4), 128), out_h, batch);
        resample_kernel<<<grid, 128, 0, st>>>(d_src, w, h, out_w, out_h, mm, d_out_u8, d_out_bf16);
    }
    MS_LAUNCH_CHECK();
}

}  // namespace ms
