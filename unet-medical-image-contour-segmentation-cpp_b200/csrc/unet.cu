// unet.cu -- UNet forward: weight preparation, activation arena, tensor maps, layer schedule.
//
// Replaces the TensorRT engine the reference deserialises and replays
// (/root/reference/src/initialize.cpp:48-60, src/process.cpp:45-120,147).  Architecture: the
// canonical UNet fixed in SURVEY.md section 8(a) row P3 (1->64->128->256->512->1024, two
// conv3x3+BN+ReLU per level, 2x2 max-pool down, ConvTranspose 2x2 s2 up, cat([skip, up]), 1x1 head).
//
// HBM layout (all activations NHWC bf16, sized for max_batch):
//   e1a                         enc1a output (direct CUDA-core conv: Cin = 1, K = 9 is pure bandwidth)
//   cat1..cat4  [.., 2C]        channels [0,C) written by the encoder's second conv (skip), [C,2C) by
//                               the ConvT of the level below -> the concat is never materialised twice
//   p1..p4                      pooled encoder outputs, written by the fused max-pool epilogue
//   e2a,e3a,e4a,ba,bb,d4a..d1a  intermediate feature maps
// The last feature map (dec1b) never reaches HBM: the head runs in the epilogue.
#include "unet.hpp"
#include "unet_conv_tc.cuh"
#include "json_min.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

namespace ms {

namespace {

// ------------------------------------------------------------------ first conv (Cin = 1): direct
// K = 9 is pure bandwidth (SURVEY.md hard part H5), so this layer stays on the CUDA cores.  One block =
// eight image rows.  The three input rows are staged once in shared memory as x = float(u8) / 255.0f
// (exactly src/process.cpp:38); every thread keeps the 9 x 8 weights of its 8 output channels in
// registers and walks the row 4 pixels at a time, so the inner loop is FMAs and 16-byte NHWC stores
// (a warp writes four fully used 128-byte lines per store instruction).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// kFirstRows = image rows per block: 8 amortises the weights-to-registers prologue and the halo rows (batch >= ~5); 2 gives a
// single slice 256 blocks instead of 64 (batch-1 latency: 29 -> ~12 us)
template <bool VEC16, int kFirstRows>
__global__ void __launch_bounds__(256) first_conv_kernel(const uint8_t* __restrict__ in, int H, int W,
                                                          const float* __restrict__ w /*[64][9]*/, const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ out /*NHWC 64*/) {
    extern __shared__ float srow[];           // [kFirstRows + 2][W + 2], column 0 <-> x = -1
    __shared__ float sw[64 * 9 + 64];
    const int blocks_per_img = H / kFirstRows;
    const int y0 = (blockIdx.x % blocks_per_img) * kFirstRows;
    const size_t img = (size_t)(blockIdx.x / blocks_per_img) * H * W;
    const int pitch = W + 2;
    for (int i = threadIdx.x; i < 64 * 9 + 64; i += 256) sw[i] = i < 576 ? w[i] : bias[i - 576];
    pdl_launch_dependents();
    pdl_wait();                   // `in` is the preprocess kernel's output
    if (VEC16) {
        // 16 pixels per load, all of a thread's loads in flight together (the scalar form below spends a fifth of the
        // kernel waiting on dependent byte loads: profiles/r1_ncu_full_forward_b32.json, enc1a)
        const int cpr = W / 16;
        for (int i = threadIdx.x; i < (kFirstRows + 2) * cpr; i += 256) {
            const int r = i / cpr, c = i - r * cpr;
            const int yy = y0 + r - 1;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (yy >= 0 && yy < H) v = __ldg(reinterpret_cast<const uint4*>(in + img + (size_t)yy * W) + c);
            float* dst = srow + r * pitch + 1 + c * 16;
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j] = __fdiv_rn((float)((wds[j >> 2] >> (8 * (j & 3))) & 0xFFu), 255.0f);   // src/process.cpp:33
        }
        for (int r = threadIdx.x; r < kFirstRows + 2; r += 256) srow[r * pitch] = srow[r * pitch + W + 1] = 0.0f;
    } else {
        for (int i = threadIdx.x; i < (kFirstRows + 2) * pitch; i += 256) {
            const int r = i / pitch, c = i % pitch;
            const int yy = y0 + r - 1, xx = c - 1;
            srow[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __fdiv_rn((float)in[img + (size_t)yy * W + xx], 255.0f) : 0.0f;
        }
    }
    __syncthreads();
    const int cg = (threadIdx.x & 7) * 8, g = threadIdx.x >> 3;
    float wr[8][9], br[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        br[j] = sw[576 + cg + j];
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[j][t] = sw[(cg + j) * 9 + t];
    }
    for (int ry = 0; ry < kFirstRows; ++ry) {
        const float* rows = srow + ry * pitch;
        for (int x0 = g * 4; x0 < W; x0 += 128) {
            float v[3][6];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) v[r][c] = rows[r * pitch + x0 + c];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // two output channels per packed FMA (sm_100 fma.rn.f32x2): same rounding as two scalar fmaf
                    float2 acc = make_float2(br[2 * j], br[2 * j + 1]);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const float x = v[t / 3][p + t % 3];
                        acc = ffma2(make_float2(x, x), make_float2(wr[2 * j][t], wr[2 * j + 1][t]), acc);
                    }
                    pk[j] = tc::pack_bf16(fmaxf(acc.x, 0.0f), fmaxf(acc.y, 0.0f));
                }
                *reinterpret_cast<uint4*>(out + (img + (size_t)(y0 + ry) * W + x0 + p) * 64 + cg) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

// ------------------------------------------------------------------ CUDA-core reference kernels
// Debug / validation only (MEDSEG_NAIVE_CONV=1): same operands, same epilogue semantics as the
// tcgen05 kernel, one thread per GEMM output element.  Never used by default.
__global__ void naive_gemm_conv_kernel(const __nv_bfloat16* __restrict__ src, int src_c, const __nv_bfloat16* __restrict__ wgt,
                                       tc::ConvArgs a, int epi) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)a.batch * a.H * a.W * a.n_total;
    if (gid >= total) return;
    const int n = (int)(gid % a.n_total);
    const size_t pix = gid / a.n_total;
    const int x = (int)(pix % a.W), y = (int)((pix / a.W) % a.H), b = (int)(pix / ((size_t)a.W * a.H));
    const int K = a.taps * a.Cin;
    float acc = 0.0f;
    for (int tap = 0; tap < a.taps; ++tap) {
        const int yy = y + (a.taps == 9 ? tap / 3 - 1 : 0), xx = x + (a.taps == 9 ? tap % 3 - 1 : 0);
        if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W) continue;
        const __nv_bfloat16* s = src + (((size_t)b * a.H + yy) * a.W + xx) * src_c;
        const __nv_bfloat16* wr = wgt + (size_t)n * K + (size_t)tap * a.Cin;
        for (int ci = 0; ci < a.Cin; ++ci) acc = fmaf(__bfloat162float(s[ci]), __bfloat162float(wr[ci]), acc);
    }
    const int co = n % a.Cout;
    acc += a.bias[co];
    if (epi == tc::EPI_STORE) {
        acc = fmaxf(acc, 0.0f);
        a.out[(((size_t)b * a.H + y) * a.W + x) * a.out_cstride + a.out_coff + n] = __float2bfloat16_rn(acc);
    } else {
        const int q = n / a.Cout;
        const int oy = 2 * y + (q >> 1), ox = 2 * x + (q & 1);
        a.out[(((size_t)b * 2 * a.H + oy) * 2 * a.W + ox) * a.out_cstride + a.out_coff + co] = __float2bfloat16_rn(acc);
    }
}
__global__ void naive_pool_kernel(const __nv_bfloat16* __restrict__ src, int cstride, int H, int W, int C, int batch,
                                  __nv_bfloat16* __restrict__ dst) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)batch * (H / 2) * (W / 2) * C;
    if (gid >= total) return;
    const int c = (int)(gid % C);
    size_t p = gid / C;
    const int x = (int)(p % (W / 2)), y = (int)((p / (W / 2)) % (H / 2)), b = (int)(p / ((size_t)(W / 2) * (H / 2)));
    float m = -INFINITY;
    for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx)
            m = fmaxf(m, __bfloat162float(src[(((size_t)b * H + 2 * y + dy) * W + 2 * x + dx) * cstride + c]));
    dst[gid] = __float2bfloat16_rn(m);
}
__global__ void naive_head_conv_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ wgt, tc::ConvArgs a) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (size_t)a.batch * a.H * a.W) return;
    const int x = (int)(pix % a.W), y = (int)((pix / a.W) % a.H), b = (int)(pix / ((size_t)a.W * a.H));
    float logit[8];
    for (int c = 0; c < a.n_classes; ++c) logit[c] = a.head_b[c];
    for (int n = 0; n < 64; ++n) {
        float acc = 0.0f;
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
            if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W) continue;
            const __nv_bfloat16* s = src + (((size_t)b * a.H + yy) * a.W + xx) * 64;
            const __nv_bfloat16* wr = wgt + (size_t)n * 576 + tap * 64;
            for (int ci = 0; ci < 64; ++ci) acc = fmaf(__bfloat162float(s[ci]), __bfloat162float(wr[ci]), acc);
        }
        acc = fmaxf(acc + a.bias[n], 0.0f);
        for (int c = 0; c < a.n_classes; ++c) logit[c] = fmaf(acc, a.head_w[c * 64 + n], logit[c]);
    }
    float best = -3.402823466e+38f;
    int bc = 0;
    const size_t plane = (size_t)a.H * a.W;
    for (int c = 0; c < a.n_classes; ++c) {
        if (a.logits) a.logits[((size_t)b * a.n_classes + c) * plane + (size_t)y * a.W + x] = logit[c];
        if (logit[c] > best) { best = logit[c]; bc = c; }
    }
    a.mask[pix] = a.n_classes == 1 ? (uint8_t)(best > 0.0f ? a.fg_value : 0) : (uint8_t)bc;
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        MS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        MS_REQUIRE(p && q == cudaDriverEntryPointSuccess, MS_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// activations [B][H][W][C] bf16, box {64, 16, 8, 1}, 128-byte swizzle, OOB -> zero (= conv padding)
void make_act_map(CUtensorMap* m, const __nv_bfloat16* base, int B, int H, int W, int C, int box_w = tc::TILE_W, int box_h = tc::TILE_H) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)tc::BLOCK_K, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MS_REQUIRE(r == CUDA_SUCCESS, MS_ERR_CUDA, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
}
// destination of EPI_STORE: [B][H][W][C] bf16, box {64 ch, tw px, 32/tw rows, 1} (one epilogue warp's slab)
void make_out_map(CUtensorMap* m, const __nv_bfloat16* base, int B, int H, int W, int C, int tw) {
    make_act_map(m, base, B, H, W, C, tw, 32 / tw);
}
// destination of EPI_CONVT: the (2H x 2W) image viewed as (C, dx:2, W, dy:2, B*H); box {64, 1, 16, 1, 2}
void make_convt_out_map(CUtensorMap* m, const __nv_bfloat16* base, int B, int H_in, int W_in, int C, int tw = tc::TILE_W) {
    cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)W_in, 2, (cuuint64_t)B * H_in};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)2 * C * 2, (cuuint64_t)2 * W_in * C * 2, (cuuint64_t)4 * W_in * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)tc::BLOCK_K, 1, (cuuint32_t)tw, 1, (cuuint32_t)(32 / tw)};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MS_REQUIRE(r == CUDA_SUCCESS, MS_ERR_CUDA, "cuTensorMapEncodeTiled(convT output) failed: " + std::to_string((int)r));
}
// destination of the row-pair kernel: [B][H][W][C] viewed as (C, W, row parity, H / 2, B); box {64 ch, 8 px, 1, 4, 1} = the four
// rows of one parity of an epilogue warp's eight
void make_rowpair_out_map(CUtensorMap* m, const __nv_bfloat16* base, int B, int H, int W, int C) {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 2, (cuuint64_t)H / 2, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)tc::BLOCK_K, (cuuint32_t)tc::RP_TW, 1, 4, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MS_REQUIRE(r == CUDA_SUCCESS, MS_ERR_CUDA, "cuTensorMapEncodeTiled(row-pair output) failed: " + std::to_string((int)r));
}
// weights [N][K] bf16, box {64, block_n}
void make_wgt_map(CUtensorMap* m, const __nv_bfloat16* base, int N, int K, int block_n) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BLOCK_K, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MS_REQUIRE(r == CUDA_SUCCESS, MS_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
}

template <int BN, int EPI>
void launch_tc(const UNetLayer& L, const tc::ConvArgs& a, int sm_count, cudaStream_t st) {
    using C = tc::Cfg<BN, EPI>;
    set_max_dynamic_smem(tc::conv_gemm_kernel<BN, EPI>, C::SMEM_BYTES);
    const int total = a.batch * (a.H / tc::TILE_H) * (a.W / tc::TILE_W) * (a.n_total / BN);
    const int grid = std::min(total, sm_count);
    launch_kernel(tc::conv_gemm_kernel<BN, EPI>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, true, L.map_a, L.map_b, L.map_out, a);
    MS_LAUNCH_CHECK();
}

template <int BN, int EPI, int RKC, int PITCH>
void launch_halo_p(const UNetLayer& L, const tc::ConvArgs& a, int sm_count, cudaStream_t st) {
    using C = tc::HaloCfg<BN, RKC, PITCH>;
    set_max_dynamic_smem(tc::conv_halo_kernel<BN, EPI, RKC, PITCH>, C::SMEM_BYTES);
    const int total = a.batch * (a.H / tc::HALO_TH) * (a.W / tc::HALO_TW);
    const int grid = std::min(total, sm_count);
    launch_kernel(tc::conv_halo_kernel<BN, EPI, RKC, PITCH>, dim3(grid), dim3(tc::NUM_THREADS), C::SMEM_BYTES, st, true, L.map_a_row, L.map_b,
                  L.map_out, a);
    MS_LAUNCH_CHECK();
}
template <int BN, int EPI, int RKC>
void launch_halo(const UNetLayer& L, const tc::ConvArgs& a, int sm_count, cudaStream_t st) {
    launch_halo_p<BN, EPI, RKC, 10>(L, a, sm_count, st);   // dense halo box (the 2048-byte pitch variant measured the same)
}

template <int EPI, int RKC>
void launch_rowpair(const UNetLayer& L, const tc::ConvArgs& a, int sm_count, cudaStream_t st) {
    using C = tc::RowPairCfg<EPI, RKC>;
    set_max_dynamic_smem(tc::conv_rowpair_kernel<EPI, RKC>, C::SMEM_BYTES);
    const int total = a.batch * (a.H / tc::RP_TH) * (a.W / tc::RP_TW);
    const int grid = std::min(total, sm_count);
    launch_kernel(tc::conv_rowpair_kernel<EPI, RKC>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, true, L.map_a_row, L.map_b, L.map_out,
                  L.map_pool, a);
    MS_LAUNCH_CHECK();
}

template <int EPI, int RKC>
void launch_rowpair2(const UNetLayer& L, const tc::ConvArgs& a, int sm_count, cudaStream_t st) {
    using C = tc::RowPair2Cfg<EPI, RKC>;
    set_max_dynamic_smem(tc::conv_rowpair2_kernel<EPI, RKC>, C::SMEM_BYTES);
    const int pairs = a.batch * (a.H / tc::RP_TH) * (a.W / tc::RP_TW) / 2;
    const int grid = 2 * std::min(pairs, sm_count / 2);
    launch_kernel(tc::conv_rowpair2_kernel<EPI, RKC>, dim3(grid), dim3(tc::HALO2_THREADS), C::SMEM_BYTES, st, true, L.map_a_row, L.map_b,
                  L.map_b_half, L.map_out, L.map_pool, a);
    MS_LAUNCH_CHECK();
}

template <int BN>
void launch_convt_pair(const UNetLayer& L, const tc::ConvArgs& a, int sm_count, cudaStream_t st) {
    using C = tc::ConvT2Cfg<BN>;
    set_max_dynamic_smem(tc::convt_pair_kernel<BN>, C::SMEM_BYTES);
    const int pairs = a.batch * (a.H / tc::HALO_TH) * (a.W / tc::HALO_TW) / 2 * (a.n_total / BN);
    const int grid = 2 * std::min(pairs, sm_count / 2);
    launch_kernel(tc::convt_pair_kernel<BN>, dim3(grid), dim3(tc::NUM_THREADS), C::SMEM_BYTES, st, true, L.map_a_row, L.map_b_half, L.map_out, a);
    MS_LAUNCH_CHECK();
}

template <int BN, int EPI, int RKC>
void launch_halo2(const UNetLayer& L, const tc::ConvArgs& a, int sm_count, cudaStream_t st, const CUtensorMap* b_half = nullptr) {
    using C = tc::Halo2Cfg<BN, RKC>;
    set_max_dynamic_smem(tc::conv_halo2_kernel<BN, EPI, RKC>, C::SMEM_BYTES);
    const int pairs = a.batch * (a.H / tc::HALO_TH) * (a.W / tc::HALO_TW) / 2 * (a.n_total / BN);
    const int grid = 2 * std::min(pairs, sm_count / 2);
    launch_kernel(tc::conv_halo2_kernel<BN, EPI, RKC>, dim3(grid), dim3(tc::HALO2_THREADS), C::SMEM_BYTES, st, true, L.map_a_row,
                  b_half ? *b_half : L.map_b_half, L.map_out, a);
    MS_LAUNCH_CHECK();
}

struct Blob {
    std::vector<char> raw;
    size_t base = 0;
    std::map<std::string, std::pair<const float*, std::vector<int64_t>>> t;
    int n_classes = 0;
    double eps = 1e-5;
    const float* get(const std::string& name, size_t expect) const {
        auto it = t.find(name);
        MS_REQUIRE(it != t.end(), MS_ERR_FORMAT, "weight blob: missing tensor " + name);
        size_t n = 1;
        for (auto d : it->second.second) n *= (size_t)d;
        MS_REQUIRE(n == expect, MS_ERR_FORMAT, "weight blob: tensor " + name + " has wrong size");
        return it->second.first;
    }
};
Blob read_blob(const std::string& path) {
    Blob b;
    std::ifstream f(path, std::ios::binary);
    MS_REQUIRE(f.good(), MS_ERR_IO, "cannot open weight blob: " + path);
    f.seekg(0, std::ios::end);
    const size_t sz = (size_t)f.tellg();
    f.seekg(0);
    b.raw.resize(sz);
    f.read(b.raw.data(), (std::streamsize)sz);
    MS_REQUIRE(sz > 12 && std::memcmp(b.raw.data(), "MSEGW001", 8) == 0, MS_ERR_FORMAT, "not a MSEGW001 weight blob: " + path);
    uint32_t hl;
    std::memcpy(&hl, b.raw.data() + 8, 4);
    MS_REQUIRE(12 + (size_t)hl <= sz, MS_ERR_FORMAT, "weight blob: truncated header");
    json::Value h;
    try {
        h = json::parse(std::string(b.raw.data() + 12, hl));
    } catch (const std::exception& e) {
        fail(MS_ERR_FORMAT, std::string("weight blob header: ") + e.what());
    }
    b.base = (12 + (size_t)hl + 63) / 64 * 64;
    b.n_classes = (int)h.at("arch").integer("n_classes", 3);
    b.eps = h.at("arch").number("bn_eps", 1e-5);
    for (const auto& tv : h.at("tensors").arr) {
        std::vector<int64_t> shape;
        size_t n = 1;
        for (const auto& d : tv.at("shape").arr) { shape.push_back((int64_t)d.num); n *= (size_t)d.num; }
        const size_t off = b.base + (size_t)tv.at("offset").num;
        MS_REQUIRE(off + n * 4 <= sz, MS_ERR_FORMAT, "weight blob: tensor out of range: " + tv.at("name").str);
        b.t[tv.at("name").str] = {reinterpret_cast<const float*>(b.raw.data() + off), shape};
    }
    return b;
}

}  // namespace

UNet::~UNet() {
    for (cudaEvent_t e : prof_events_) cudaEventDestroy(e);
    for (void* p : allocs_) cudaFree(p);
}

void UNet::load(const std::string& blob_path, int net_h, int net_w, int n_classes_cfg, int max_batch, int fg_value, int sm_count) {
    MS_REQUIRE(!loaded_, MS_ERR_STATE, "UNet already loaded");
    MS_REQUIRE(net_h % 128 == 0 && net_w % 256 == 0 && net_h > 0 && net_w > 0, MS_ERR_ARG,
               "net size must be a multiple of 128 (height) x 256 (width): 16x8-pixel tiles at 1/16 resolution");
    MS_REQUIRE(max_batch >= 1 && max_batch <= 4096, MS_ERR_ARG, "max_batch out of range");
    H_ = net_h; W_ = net_w; max_batch_ = max_batch; fg_value_ = fg_value; sm_count_ = sm_count;
    const char* nv = std::getenv("MEDSEG_NAIVE_CONV");
    naive_ = nv && nv[0] == '1';
    const char* hv = std::getenv("MEDSEG_HALO");
    halo_enabled_ = !(hv && hv[0] == '0');
    const char* cv = std::getenv("MEDSEG_CTA2");
    cta2_enabled_ = !(cv && cv[0] == '0');
    cta2_force_ = cv && cv[0] == '2';
    cta2_single_pref_ = cv && cv[0] == '1';
    const char* d2 = std::getenv("MEDSEG_DEEP2");
    deep2_enabled_ = !(d2 && d2[0] == '0');
    const char* rb = std::getenv("MEDSEG_RES_BIG");
    res_big_ = !(rb && rb[0] == '0');
    const char* sv = std::getenv("MEDSEG_STREAM2");
    stream2_enabled_ = !(sv && sv[0] == '0');
    const char* rp = std::getenv("MEDSEG_ROWPAIR");
    rowpair_enabled_ = !(rp && rp[0] == '0');
    rowpair_stream_ = !(rp && rp[0] == '1');
    const char* rp2 = std::getenv("MEDSEG_ROWPAIR2");
    rowpair2_enabled_ = !(rp2 && rp2[0] == '0');   // MEDSEG_CTA2=2: prefer the pair kernel wherever it applies (A/B measurements)
    const char* pv = std::getenv("MEDSEG_HALO_PITCH");
    halo_pitch_ = 10;
    (void)pv;
    {   // publish the watchdog word to this device's copy of the symbol (mapped host memory: same address under UVA)
        unsigned* wd = watchdog_host_word();
        MS_CUDA(cudaMemcpyToSymbol(tc::g_watchdog_dev, &wd, sizeof(wd)));
    }
    Blob blob = read_blob(blob_path);
    n_classes_ = blob.n_classes;
    MS_REQUIRE(n_classes_cfg <= 0 || n_classes_cfg == n_classes_, MS_ERR_FORMAT, "config n_classes does not match the weight blob");
    MS_REQUIRE(n_classes_ >= 1 && n_classes_ <= 8, MS_ERR_ARG, "n_classes must be 1..8");

    auto dmalloc = [&](size_t bytes) { void* p = nullptr; MS_CUDA(cudaMalloc(&p, bytes)); allocs_.push_back(p); return p; };
    auto upload_f32 = [&](const std::vector<float>& v) { float* d = (float*)dmalloc(v.size() * 4); MS_CUDA(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice)); return d; };
    auto upload_bf16 = [&](const std::vector<__nv_bfloat16>& v) { auto* d = (__nv_bfloat16*)dmalloc(v.size() * 2); MS_CUDA(cudaMemcpy(d, v.data(), v.size() * 2, cudaMemcpyHostToDevice)); return d; };

    // ---- activation arena
    auto add_buf = [&](const char* name, int level, int C) {
        ActBuf b; b.name = name; b.level = level; b.C = C;
        const size_t n = (size_t)max_batch * (H_ >> level) * (W_ >> level) * C;
        b.p = (__nv_bfloat16*)dmalloc(n * 2);
        bufs_.push_back(b);
        return (int)bufs_.size() - 1;
    };
    const int e1a = add_buf("e1a", 0, 64), cat1 = add_buf("cat1", 0, 128), p1 = add_buf("p1", 1, 64);
    const int e2a = add_buf("e2a", 1, 128), cat2 = add_buf("cat2", 1, 256), p2 = add_buf("p2", 2, 128);
    const int e3a = add_buf("e3a", 2, 256), cat3 = add_buf("cat3", 2, 512), p3 = add_buf("p3", 3, 256);
    const int e4a = add_buf("e4a", 3, 512), cat4 = add_buf("cat4", 3, 1024), p4 = add_buf("p4", 4, 512);
    const int ba = add_buf("ba", 4, 1024), bb = add_buf("bb", 4, 1024);
    const int d4a = add_buf("d4a", 3, 512), d4b = add_buf("d4b", 3, 512);
    const int d3a = add_buf("d3a", 2, 256), d3b = add_buf("d3b", 2, 256);
    const int d2a = add_buf("d2a", 1, 128), d2b = add_buf("d2b", 1, 128);
    const int d1a = add_buf("d1a", 0, 64);
    scratch_mask_.reserve((size_t)max_batch * H_ * W_);

    // ---- weights
    n_params_ = 0;
    flops_ = 0;
    const double eps = blob.eps;
    // conv3x3 + BN folded: returns K-major bf16 [Cout][9*Cin] and fp32 bias
    auto fold_conv = [&](const std::string& conv, const std::string& bn, int cin, int cout, std::vector<float>& wf, std::vector<float>& bias) {
        const float* w = blob.get(conv + ".weight", (size_t)cout * cin * 9);
        const float* g = blob.get(bn + ".weight", cout);
        const float* be = blob.get(bn + ".bias", cout);
        const float* mu = blob.get(bn + ".running_mean", cout);
        const float* var = blob.get(bn + ".running_var", cout);
        wf.assign((size_t)cout * 9 * cin, 0.f);
        bias.assign(cout, 0.f);
        for (int co = 0; co < cout; ++co) {
            const float s = g[co] / std::sqrt(var[co] + (float)eps);
            bias[co] = be[co] - mu[co] * s;
            for (int ci = 0; ci < cin; ++ci)
                for (int t = 0; t < 9; ++t) wf[((size_t)co * 9 + t) * cin + ci] = w[((size_t)co * cin + ci) * 9 + t] * s;
        }
        n_params_ += (int64_t)cout * cin * 9 + 2 * cout;
    };
    auto to_bf16 = [](const std::vector<float>& v) {
        std::vector<__nv_bfloat16> o(v.size());
        for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
        return o;
    };
    auto pick_bn = [](int n_total) { return n_total % 256 == 0 ? 256 : (n_total % 128 == 0 ? 128 : 64); };
    auto add_conv = [&](const char* name, const std::string& conv, const std::string& bn, int level, int cin, int cout, int src, int dst,
                        int coff, int pool_dst, bool head) {
        UNetLayer L;
        L.name = name; L.kind = head ? 3 : 1; L.level = level; L.Cin = cin; L.Cout = cout; L.taps = 9; L.n_total = cout;
        L.block_n = head ? 64 : pick_bn(cout);
        L.src = src; L.dst = dst; L.dst_coff = coff; L.pool_dst = pool_dst;
        std::vector<float> wf, bias;
        fold_conv(conv, bn, cin, cout, wf, bias);
        L.w = upload_bf16(to_bf16(wf));
        L.bias = upload_f32(bias);
        const int h = H_ >> level, w = W_ >> level;
        L.flops_per_slice = 2.0 * h * w * (double)cout * 9 * cin;
        make_act_map(&L.map_a, bufs_[src].p, max_batch, h, w, bufs_[src].C);
        make_wgt_map(&L.map_b, L.w, cout, 9 * cin, L.block_n);
        // Halo-stationary kernels for the narrow-N, large-grid layers, used only when the layer's weights can
        // stay resident in shared memory (measured A/B, profiles/r1_ab_kernel_variants.log): single CTA when
        // all 9 * kc weight tiles fit in 144 KiB, the cta_group::2 pair when half of them do; otherwise the
        // per-tap streaming kernel is faster.
        if (halo_enabled_ && cout == L.block_n && L.block_n <= 128 && h % tc::HALO_TH == 0 && w % tc::HALO_TW == 0) {
            const int kc = cin / tc::BLOCK_K;
            // single CTA: all weights resident.  With two chunks a 144 KiB weight set leaves only two halo stages and the
            // pair kernel (half the weights per CTA, four stages) is faster; with one chunk per tile the pair kernel's
            // cross-CTA hand-offs are not amortised and it loses (profiles/r1_ab_pair_vs_single.log).
            const int bytes1 = 9 * kc * L.block_n * 128;
            const bool fits1 = bytes1 <= 144 * 1024 && (kc == 1 || bytes1 <= 96 * 1024);
            const bool pair_ok = ((h / tc::HALO_TH) * (w / tc::HALO_TW)) % 2 == 0;
            // a resident half-weight set that leaves room for only two halo stages loses to streaming (measured)
            const bool fits2 = 9 * kc * (L.block_n / 2) * 128 <= (res_big_ ? 144 : 96) * 1024 && pair_ok;
            // N = 128 with one chunk (enc2a): since the pair's remote arrive lost its cluster-scope release the pair kernel wins
            // here too (0.216 -> 0.203 ms); MEDSEG_CTA2=1 keeps the single-CTA kernel
            const bool prefer_pair = cta2_enabled_ && !cta2_single_pref_ && fits2 && L.block_n == 128;
            if (fits1 && !prefer_pair && !(cta2_force_ && fits2)) {
                L.halo = 1;
                L.resident_kc = kc;
            } else if (fits2 && cta2_enabled_) {
                L.halo = 2;
                L.resident_kc = kc;
                make_wgt_map(&L.map_b_half, L.w, cout, 9 * cin, L.block_n / 2);
            } else if (cta2_enabled_ && stream2_enabled_ && L.block_n == 128 && ((h / tc::HALO_TH) * (w / tc::HALO_TW)) % 2 == 0) {
                L.halo = 2;             // weights too large to stay resident: streamed by their own producer warp
                L.resident_kc = 0;
                make_wgt_map(&L.map_b_half, L.w, cout, 9 * cin, L.block_n / 2);
            }
            if (L.halo) make_act_map(&L.map_a_row, bufs_[src].p, max_batch, h, w, bufs_[src].C, tc::HALO_TW + 2, tc::HALO_TH + 2);
            // Cout = 64 with one 64-channel chunk (enc1b, dec1b + head): the row-pair kernel packs two output rows into one
            // N = 128 accumulator, 4 MMA issue slots per 256 pixels and filter column instead of 6 (the N = 64 layers sit on
            // the ~72-cycle issue floor of an M = 128 instruction)
            // (two chunks -- dec1a -- stream their weights: 144 KiB do not fit beside two 43 KiB halo stages; MEDSEG_ROWPAIR=1
            // keeps that layer on the pair kernel)
            if (rowpair_enabled_ && L.halo && L.block_n == 64 && (kc == 1 || (kc == 2 && rowpair_stream_ && !head)) && h % tc::RP_TH == 0 &&
                w % tc::RP_TW == 0) {
                L.halo = 3;
                L.resident_kc = kc == 1 ? 1 : 0;
                // its cta_group::2 form: every CTA reads half of each B operand (the kernel is shared-memory-bandwidth bound)
                if (cta2_enabled_ && rowpair2_enabled_ && ((h / tc::RP_TH) * (w / tc::RP_TW)) % 2 == 0) {
                    L.halo = 4;
                    make_wgt_map(&L.map_b_half, L.w, cout, 9 * cin, 32);
                }
                make_act_map(&L.map_a_row, bufs_[src].p, max_batch, h, w, bufs_[src].C, tc::RP_TW + 2, tc::RP_TH + 2);
            }
        }
        // deep layers (N tile = 256): streaming pair kernel -- each CTA fetches the halo once per chunk and half of every
        // weight tile, ~2.6x less L2 -> SM traffic than the per-tap kernel
        if (!L.halo && halo_enabled_ && cta2_enabled_ && deep2_enabled_ && L.block_n == 256 && h % tc::HALO_TH == 0 &&
            w % tc::HALO_TW == 0 && ((h / tc::HALO_TH) * (w / tc::HALO_TW)) % 2 == 0) {
            L.halo = 2;
            L.resident_kc = 0;
            make_wgt_map(&L.map_b_half, L.w, cout, 9 * cin, 128);
            make_act_map(&L.map_a_row, bufs_[src].p, max_batch, h, w, bufs_[src].C, tc::HALO_TW + 2, tc::HALO_TH + 2);
            // small batches: the same kernel with 128-wide N tiles doubles the number of work units when the 256-wide
            // tiling cannot fill one wave of CTA pairs (batch 1: bott_b 49 -> 29 us, whole forward 0.54 -> 0.46 ms)
            L.block_n_alt = 128;
            make_wgt_map(&L.map_b_half_alt, L.w, cout, 9 * cin, 64);
        }
        if (dst >= 0 && L.halo >= 3) make_rowpair_out_map(&L.map_out, bufs_[dst].p, max_batch, h, w, bufs_[dst].C);
        else if (dst >= 0) make_out_map(&L.map_out, bufs_[dst].p, max_batch, h, w, bufs_[dst].C, L.halo ? tc::HALO_TW : tc::TILE_W);
        else L.map_out = L.map_b;  // head layer: no bf16 output
        L.map_pool = L.map_out;
        if (L.halo >= 3 && pool_dst >= 0)      // pooled map of the row-pair kernels: box {64 ch, 4 px, 4 rows} = one epilogue warp's share
            make_act_map(&L.map_pool, bufs_[pool_dst].p, max_batch, h / 2, w / 2, bufs_[pool_dst].C, 4, 4);
        flops_ += L.flops_per_slice;
        layers_.push_back(L);
    };
    auto add_convt = [&](const char* name, const std::string& up, int level_in, int cin, int cout, int src, int dst, int coff) {
        UNetLayer L;
        L.name = name; L.kind = 2; L.level = level_in; L.Cin = cin; L.Cout = cout; L.taps = 1; L.n_total = 4 * cout;
        L.block_n = pick_bn(4 * cout);
        L.src = src; L.dst = dst; L.dst_coff = coff;
        const float* w = blob.get(up + ".weight", (size_t)cin * cout * 4);  // torch layout [Cin][Cout][2][2]
        const float* b = blob.get(up + ".bias", cout);
        std::vector<float> wf((size_t)4 * cout * cin);
        for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
                for (int q = 0; q < 4; ++q) wf[((size_t)q * cout + co) * cin + ci] = w[((size_t)ci * cout + co) * 4 + q];
        L.w = upload_bf16(to_bf16(wf));
        L.bias = upload_f32(std::vector<float>(b, b + cout));
        const int h = H_ >> level_in, wd = W_ >> level_in;
        L.flops_per_slice = 2.0 * h * wd * 4.0 * cout * cin;
        make_act_map(&L.map_a, bufs_[src].p, max_batch, h, wd, bufs_[src].C);
        make_wgt_map(&L.map_b, L.w, 4 * cout, cin, L.block_n);
        make_convt_out_map(&L.map_out, bufs_[dst].p, max_batch, h, wd, bufs_[dst].C);
        // cta_group::2 GEMM: half the weight bytes per CTA.  Measured in-step at batch 32 (profiles/r2_ab_convt_pair.log):
        // up4 (K = 1024) 0.115 -> 0.097 ms; up3 / up2 / up1 (K <= 512: one to four k-steps per tile, the epilogue and the
        // 5-D scatter dominate and the pair's hand-offs are not amortised) lose 13 - 34 %, so they keep the per-tap kernel.
        // MEDSEG_CONVT_PAIR=0 / =all force either form everywhere.
        const char* cp = std::getenv("MEDSEG_CONVT_PAIR");
        const bool pair_all = cp && cp[0] == 'a';
        if (cta2_enabled_ && !(cp && cp[0] == '0') && (cin >= 1024 || pair_all) && L.block_n == 256 && h % tc::HALO_TH == 0 && wd % tc::HALO_TW == 0 &&
            ((h / tc::HALO_TH) * (wd / tc::HALO_TW)) % 2 == 0) {
            L.convt_pair = true;
            make_act_map(&L.map_a_row, bufs_[src].p, max_batch, h, wd, bufs_[src].C, tc::HALO_TW, tc::HALO_TH);
            make_wgt_map(&L.map_b_half, L.w, 4 * cout, cin, 128);
            make_convt_out_map(&L.map_out, bufs_[dst].p, max_batch, h, wd, bufs_[dst].C, tc::HALO_TW);
        }
        n_params_ += (int64_t)cin * cout * 4 + cout;
        flops_ += L.flops_per_slice;
        layers_.push_back(L);
    };

    {   // enc1a: direct conv, fp32 weights [64][9]
        UNetLayer L;
        L.name = "enc1a"; L.kind = 0; L.level = 0; L.Cin = 1; L.Cout = 64; L.taps = 9; L.n_total = 64; L.dst = e1a;
        std::vector<float> wf, bias;
        fold_conv("inc.double_conv.0", "inc.double_conv.1", 1, 64, wf, bias);  // [co][t][ci=1] == [64][9]
        L.w_f32 = upload_f32(wf);
        L.bias = upload_f32(bias);
        L.flops_per_slice = 2.0 * H_ * W_ * 64 * 9;
        flops_ += L.flops_per_slice;
        layers_.push_back(L);
    }
    add_conv("enc1b", "inc.double_conv.3", "inc.double_conv.4", 0, 64, 64, e1a, cat1, 0, p1, false);
    const int enc_src[4] = {p1, p2, p3, p4}, enc_mid[4] = {e2a, e3a, e4a, ba}, enc_dst[4] = {cat2, cat3, cat4, bb};
    const int enc_pool[4] = {p2, p3, p4, -1};
    static const char* enc_a[4] = {"enc2a", "enc3a", "enc4a", "bott_a"};
    static const char* enc_b[4] = {"enc2b", "enc3b", "enc4b", "bott_b"};
    for (int i = 0; i < 4; ++i) {
        const int cin = 64 << i, cout = 128 << i;
        const std::string pre = "down" + std::to_string(i + 1) + ".maxpool_conv.1.double_conv";
        add_conv(enc_a[i], pre + ".0", pre + ".1", i + 1, cin, cout, enc_src[i], enc_mid[i], 0, -1, false);
        add_conv(enc_b[i], pre + ".3", pre + ".4", i + 1, cout, cout, enc_mid[i], enc_dst[i], 0, enc_pool[i], false);
    }
    const int up_src[4] = {bb, d4b, d3b, d2b}, up_cat[4] = {cat4, cat3, cat2, cat1}, dec_mid[4] = {d4a, d3a, d2a, d1a};
    const int dec_out[4] = {d4b, d3b, d2b, -1};
    static const char* up_n[4] = {"up4", "up3", "up2", "up1"};
    static const char* dec_a[4] = {"dec4a", "dec3a", "dec2a", "dec1a"};
    static const char* dec_b[4] = {"dec4b", "dec3b", "dec2b", "dec1b_head"};
    for (int i = 0; i < 4; ++i) {
        const int cin = 1024 >> i, cout = cin / 2, level = 3 - i;
        const std::string pre = "up" + std::to_string(i + 1);
        add_convt(up_n[i], pre + ".up", level + 1, cin, cout, up_src[i], up_cat[i], cout);
        add_conv(dec_a[i], pre + ".conv.double_conv.0", pre + ".conv.double_conv.1", level, cin, cout, up_cat[i], dec_mid[i], 0, -1, false);
        add_conv(dec_b[i], pre + ".conv.double_conv.3", pre + ".conv.double_conv.4", level, cout, cout, dec_mid[i], dec_out[i], 0, -1, i == 3);
    }
    {   // 1x1 head (fp32, applied in the dec1b epilogue)
        const float* w = blob.get("outc.conv.weight", (size_t)n_classes_ * 64);
        const float* b = blob.get("outc.conv.bias", n_classes_);
        head_w_ = upload_f32(std::vector<float>(w, w + (size_t)n_classes_ * 64));
        head_b_ = upload_f32(std::vector<float>(b, b + n_classes_));
        n_params_ += (int64_t)n_classes_ * 64 + n_classes_;
        flops_ += 2.0 * H_ * W_ * 64 * n_classes_;
    }
    loaded_ = true;
}

void UNet::run_layer(int li, const uint8_t* d_in_u8, int batch, uint8_t* d_mask, float* d_logits, cudaStream_t st) {
    MS_REQUIRE(loaded_, MS_ERR_STATE, "UNet weights not loaded");
    MS_REQUIRE(li >= 0 && li < (int)layers_.size(), MS_ERR_ARG, "layer index out of range");
    MS_REQUIRE(batch >= 1 && batch <= max_batch_, MS_ERR_ARG, "batch exceeds max_batch");
    const UNetLayer& L = layers_[li];
    const int h = H_ >> L.level, w = W_ >> L.level;
    if (L.kind == 0) {
        const bool small = batch * (h / 8) < 2 * sm_count_ && h % 2 == 0;     // fewer than two waves of 8-row blocks
        const int rows = small ? 2 : 8;
        const unsigned grid = (unsigned)(batch * h / rows);
        const size_t smem = (size_t)(rows + 2) * (w + 2) * sizeof(float);
        MS_REQUIRE(smem <= 200 * 1024, MS_ERR_ARG, "network width too large for the first-conv row buffer");
        auto go = [&](auto kernel) {
            if (smem > 48 * 1024) set_max_dynamic_smem(kernel, 200 * 1024);   // nets wider than ~1200 px
            launch_kernel(kernel, dim3(grid), dim3(256), smem, st, true, d_in_u8, h, w, (const float*)L.w_f32, (const float*)L.bias, bufs_[L.dst].p);
        };
        if (w % 16 == 0 && (reinterpret_cast<uintptr_t>(d_in_u8) & 15) == 0) { if (small) go(first_conv_kernel<true, 2>); else go(first_conv_kernel<true, 8>); }
        else { if (small) go(first_conv_kernel<false, 2>); else go(first_conv_kernel<false, 8>); }
        MS_LAUNCH_CHECK();
        return;
    }
    tc::ConvArgs a{};
    a.H = h; a.W = w; a.batch = batch; a.Cin = L.Cin; a.taps = L.taps; a.n_total = L.n_total; a.Cout = L.Cout; a.bias = L.bias;
    if (L.dst >= 0) { a.out = bufs_[L.dst].p; a.out_cstride = bufs_[L.dst].C; a.out_coff = L.dst_coff; }
    if (L.pool_dst >= 0) { a.pool = bufs_[L.pool_dst].p; a.pool_cstride = bufs_[L.pool_dst].C; }
    a.head_w = head_w_; a.head_b = head_b_; a.n_classes = n_classes_; a.fg_value = fg_value_;
    a.mask = d_mask ? d_mask : scratch_mask_.as<uint8_t>();
    a.logits = d_logits;
    if (naive_) {
        const __nv_bfloat16* src = bufs_[L.src].p;
        if (L.kind == 3) {
            const size_t npix = (size_t)batch * h * w;
            naive_head_conv_kernel<<<(unsigned)cdiv64((int64_t)npix, 128), 128, 0, st>>>(src, L.w, a);
            MS_LAUNCH_CHECK();
        } else {
            const size_t total = (size_t)batch * h * w * L.n_total;
            naive_gemm_conv_kernel<<<(unsigned)cdiv64((int64_t)total, 256), 256, 0, st>>>(src, bufs_[L.src].C, L.w, a,
                                                                                         L.kind == 2 ? tc::EPI_CONVT : tc::EPI_STORE);
            MS_LAUNCH_CHECK();
            if (L.pool_dst >= 0) {
                const size_t pt = (size_t)batch * (h / 2) * (w / 2) * L.Cout;
                naive_pool_kernel<<<(unsigned)cdiv64((int64_t)pt, 256), 256, 0, st>>>(a.out + a.out_coff, a.out_cstride, h, w, L.Cout, batch, a.pool);
                MS_LAUNCH_CHECK();
            }
        }
        return;
    }
    if (L.halo == 2 && L.block_n_alt == 128 && L.resident_kc == 0 &&
        batch * (h / tc::HALO_TH) * (w / tc::HALO_TW) / 2 * (L.n_total / L.block_n) < sm_count_ / 2) {
        launch_halo2<128, tc::EPI_STORE, 0>(L, a, sm_count_, st, &L.map_b_half_alt);   // less than one wave of pairs at N = 256
        return;
    }
    if (L.halo == 4) {
        if (L.kind == 3) launch_rowpair2<tc::EPI_HEAD, 1>(L, a, sm_count_, st);
        else if (L.resident_kc == 1) launch_rowpair2<tc::EPI_STORE, 1>(L, a, sm_count_, st);
        else launch_rowpair2<tc::EPI_STORE, 0>(L, a, sm_count_, st);
    } else if (L.halo == 3) {
        if (L.kind == 3) launch_rowpair<tc::EPI_HEAD, 1>(L, a, sm_count_, st);
        else if (L.resident_kc == 1) launch_rowpair<tc::EPI_STORE, 1>(L, a, sm_count_, st);
        else launch_rowpair<tc::EPI_STORE, 0>(L, a, sm_count_, st);
    } else if (L.halo == 2) {
        const int rk = L.resident_kc;
        if (L.kind == 3 && rk == 1) launch_halo2<64, tc::EPI_HEAD, 1>(L, a, sm_count_, st);
        else if (L.kind != 3 && L.block_n == 64 && rk == 1) launch_halo2<64, tc::EPI_STORE, 1>(L, a, sm_count_, st);
        else if (L.kind != 3 && L.block_n == 64 && rk == 2) launch_halo2<64, tc::EPI_STORE, 2>(L, a, sm_count_, st);
        else if (L.kind != 3 && L.block_n == 128 && rk == 1) launch_halo2<128, tc::EPI_STORE, 1>(L, a, sm_count_, st);
        else if (L.kind != 3 && L.block_n == 128 && rk == 2) launch_halo2<128, tc::EPI_STORE, 2>(L, a, sm_count_, st);
        else if (L.kind != 3 && L.block_n == 128 && rk == 0) launch_halo2<128, tc::EPI_STORE, 0>(L, a, sm_count_, st);
        else if (L.kind != 3 && L.block_n == 256 && rk == 0) launch_halo2<256, tc::EPI_STORE, 0>(L, a, sm_count_, st);
        else fail(MS_ERR_INTERNAL, "no halo2 kernel instantiation for layer " + L.name);
    } else if (L.halo) {
        if (L.kind == 3) launch_halo<64, tc::EPI_HEAD, 1>(L, a, sm_count_, st);
        else if (L.block_n == 64 && L.resident_kc == 1) launch_halo<64, tc::EPI_STORE, 1>(L, a, sm_count_, st);
        else if (L.block_n == 64 && L.resident_kc == 2) launch_halo<64, tc::EPI_STORE, 2>(L, a, sm_count_, st);
        else if (L.block_n == 128 && L.resident_kc == 1) launch_halo<128, tc::EPI_STORE, 1>(L, a, sm_count_, st);
        else fail(MS_ERR_INTERNAL, "no halo kernel instantiation for layer " + L.name);
    } else if (L.kind == 3) {
        launch_tc<64, tc::EPI_HEAD>(L, a, sm_count_, st);
    } else if (L.kind == 2) {
        if (L.convt_pair) launch_convt_pair<256>(L, a, sm_count_, st);
        else if (L.block_n == 256) launch_tc<256, tc::EPI_CONVT>(L, a, sm_count_, st);
        else if (L.block_n == 128) launch_tc<128, tc::EPI_CONVT>(L, a, sm_count_, st);
        else launch_tc<64, tc::EPI_CONVT>(L, a, sm_count_, st);
    } else {
        if (L.block_n == 256) launch_tc<256, tc::EPI_STORE>(L, a, sm_count_, st);
        else if (L.block_n == 128) launch_tc<128, tc::EPI_STORE>(L, a, sm_count_, st);
        else launch_tc<64, tc::EPI_STORE>(L, a, sm_count_, st);
    }
}

void UNet::forward(const uint8_t* d_in_u8, int batch, uint8_t* d_mask, float* d_logits, cudaStream_t st) {
    const int n = (int)layers_.size();
    cudaEvent_t* ev = prof_n_ < prof_cap_ ? prof_events_.data() + (size_t)prof_n_ * (n + 1) : nullptr;
    for (int i = 0; i < n; ++i) {
        if (ev) MS_CUDA(cudaEventRecord(ev[i], st));
        run_layer(i, d_in_u8, batch, d_mask, d_logits, st);
    }
    if (ev) {
        MS_CUDA(cudaEventRecord(ev[n], st));
        ++prof_n_;
    }
}

void UNet::profile_begin(int max_forwards) {
    const size_t need = (size_t)std::max(0, max_forwards) * (layers_.size() + 1);
    while (prof_events_.size() < need) {
        cudaEvent_t e;
        MS_CUDA(cudaEventCreate(&e));
        prof_events_.push_back(e);
    }
    prof_cap_ = std::max(0, max_forwards);
    prof_n_ = 0;
}

int UNet::profile_read(std::vector<float>& ms_per_layer) {
    const int n = (int)layers_.size(), passes = prof_n_;
    ms_per_layer.assign(n, 0.0f);
    for (int f = 0; f < passes; ++f) {
        cudaEvent_t* ev = prof_events_.data() + (size_t)f * (n + 1);
        MS_CUDA(cudaEventSynchronize(ev[n]));
        for (int i = 0; i < n; ++i) {
            float ms = 0;
            MS_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            ms_per_layer[i] += ms;
        }
    }
    for (float& v : ms_per_layer) v = passes ? v / passes : 0.0f;
    prof_cap_ = 0;
    prof_n_ = 0;
    return passes;
}

}  // namespace ms
