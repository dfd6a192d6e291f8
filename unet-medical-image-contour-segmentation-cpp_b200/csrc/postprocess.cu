// postprocess.cu -- K5: hole fill + 3x3 open + 8-connected area filter on the class mask.
//
// Replaces postprocess_mask / fill_holes_inside_foreground (/root/reference/src/postprocess.cpp:13-79):
//   1. bin = (mask == FG); inv = ~bin; 8-connected components of inv with stats          (:18-26)
//   2. an inv component is a hole iff its bbox stays off the image border and its area is below
//      int(w*h*0.06f); holes become FG                                                   (:30-43)
//   3. bin = (mask == FG); morphologyEx(OPEN, 3x3 rect): erode then dilate, with OpenCV's default
//      border (outside pixels never win the min / max)                                   (:57-60)
//   4. 8-connected components of the opened image; keep area >= int(w*h*0.06f)           (:64-72)
//   5. output 0 / FG                                                                     (:75-76)
// The reference pays O(n_components * H * W) for `labels == i` scans; here every step is one
// streaming pass.  "bbox off the border" == "no pixel on the border", so a 1-byte flag replaces the
// bbox.  Algorithmic bytes: 2 B/px (mask in, mask out); the label scratch is extra traffic.
#include "ccl.cuh"

namespace ms {

namespace {

struct PredNe {  // inverse foreground
    int v;
    __device__ bool operator()(uint8_t m) const { return m != v; }
};

constexpr int TW = 32, TH = 8;  // one warp per tile row -> matches the ccl segment layout

// Fused: hole fill -> erode -> dilate -> labels/area init of the opened image.
// grid = (ceil(W/32), ceil(H/8), batch), block = 256.
__global__ void __launch_bounds__(256) fill_open_init_kernel(const uint8_t* __restrict__ mask, int H, int W, int fg_value,
                                                              int min_area, const int* __restrict__ inv_labels,
                                                              const int* __restrict__ inv_area,
                                                              const uint8_t* __restrict__ inv_flag,
                                                              int* __restrict__ out_labels, int* __restrict__ out_area) {
    __shared__ uint8_t F[TH + 4][TW + 4];  // filled foreground, halo 2; outside the image = 1 (never wins erode's min)
    __shared__ uint8_t E[TH + 2][TW + 2];  // eroded, halo 1; outside the image = 0 (never wins dilate's max)
    const size_t slice = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    for (int i = threadIdx.x; i < (TH + 4) * (TW + 4); i += 256) {
        const int ly = i / (TW + 4), lx = i % (TW + 4);
        const int x = x0 + lx - 2, y = y0 + ly - 2;
        uint8_t f = 1;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const int p = y * W + x;
            if (mask[slice + p] == fg_value) {
                f = 1;
            } else {
                const int r = inv_labels[slice + p];  // root of the inverse component (>= 0 here)
                f = (r >= 0 && inv_flag[slice + r] == 0 && inv_area[slice + r] < min_area) ? 1 : 0;
            }
        }
        F[ly][lx] = f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += 256) {
        const int ly = i / (TW + 2), lx = i % (TW + 2);
        const int x = x0 + lx - 1, y = y0 + ly - 1;
        uint8_t e = 0;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            e = 1;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) e &= F[ly + dy][lx + dx];
        }
        E[ly][lx] = e;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    const bool in = x < W && y < H;
    uint8_t o = 0;
    if (in) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) o |= E[ly + dy][lx + dx];
    }
    // ccl init of the opened image (same run-start rule as ccl::init_kernel)
    const unsigned bits = __ballot_sync(0xFFFFFFFFu, o != 0);
    if (!in) return;
    const int p = y * W + x;
    int lab = -1;
    if (o) {
        const unsigned zeros_below = ~bits & ((1u << lx) - 1u);
        const int start = zeros_below ? 32 - __clz(zeros_below) : 0;
        lab = p - lx + start;
    }
    out_labels[slice + p] = lab;
    out_area[slice + p] = 0;
}

__global__ void __launch_bounds__(256) keep_kernel(const int* __restrict__ labels, const int* __restrict__ area, size_t n_per_slice,
                                                    int min_area, int fg_value, uint8_t* __restrict__ out) {
    const size_t slice = (size_t)blockIdx.y * n_per_slice;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_per_slice) return;
    const int r = labels[slice + i];
    out[slice + i] = (r >= 0 && area[slice + r] >= min_area) ? (uint8_t)fg_value : (uint8_t)0;
}

}  // namespace

void postprocess_launch(PostprocessWs& ws, const uint8_t* d_in, uint8_t* d_out, int h, int w, int batch, int fg_value,
                        float min_area_ratio, cudaStream_t st) {
    MS_REQUIRE(h > 0 && w > 0 && batch > 0 && h <= 65535 && batch <= 65535, MS_ERR_ARG, "postprocess: bad shape");
    MS_REQUIRE((int64_t)h * w < ((int64_t)1 << 31), MS_ERR_ARG, "postprocess: slice too large");
    const size_t n = (size_t)h * w, nb = n * batch;
    // src/postprocess.cpp:30,66: static_cast<int>(w * h * MIN_AREA_RATIO) -- int product, float multiply, truncate
    const int min_area = static_cast<int>(static_cast<float>(w * h) * min_area_ratio);
    ws.ccl.labels.reserve(nb * 4);
    ws.ccl.area.reserve(nb * 4);
    ws.ccl.flag.reserve(nb);
    ws.bin_a.reserve(nb * 4);  // second label plane
    ws.bin_b.reserve(nb * 4);  // second area plane
    int* L1 = ws.ccl.labels.as<int>();
    int* A1 = ws.ccl.area.as<int>();
    uint8_t* F1 = ws.ccl.flag.as<uint8_t>();
    int* L2 = ws.bin_a.as<int>();
    int* A2 = ws.bin_b.as<int>();
    const dim3 g = ccl::grid_for(h, w, batch);

    ccl::init_kernel<<<g, ccl::kThreads, 0, st>>>(d_in, h, w, PredNe{fg_value}, L1, A1, F1);
    MS_LAUNCH_CHECK();
    ccl::merge_kernel<8><<<g, ccl::kThreads, 0, st>>>(L1, h, w);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<<<g, ccl::kThreads, 0, st>>>(L1, h, w, A1, F1);
    MS_LAUNCH_CHECK();
    fill_open_init_kernel<<<dim3(cdiv(w, TW), cdiv(h, TH), batch), 256, 0, st>>>(d_in, h, w, fg_value, min_area, L1, A1, F1, L2, A2);
    MS_LAUNCH_CHECK();
    ccl::merge_kernel<8><<<g, ccl::kThreads, 0, st>>>(L2, h, w);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<<<g, ccl::kThreads, 0, st>>>(L2, h, w, A2, nullptr);
    MS_LAUNCH_CHECK();
    keep_kernel<<<dim3((unsigned)cdiv64(n, 256), batch), 256, 0, st>>>(L2, A2, n, min_area, fg_value, d_out);
    MS_LAUNCH_CHECK();
}

}  // namespace ms
