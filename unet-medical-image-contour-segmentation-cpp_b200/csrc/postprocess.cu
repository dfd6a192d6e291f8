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
// The reference pays O(n_components * H * W) for `labels == i` scans.  Here the u8 mask is read once (-> one bit per
// pixel) and written once; everything in between works on bit-packed rows: run-based union-find (ccl.cuh), hole fill
// as "OR the run masks of hole components", and the 3x3 open as shifts / ANDs / ORs of 36-bit row windows.
// "bbox off the border" == "no pixel on the border", so a 1-byte flag per root replaces the bbox.
// Algorithmic bytes: 2 B/px (mask in, mask out).
#include "ccl.cuh"

namespace ms {

namespace {

// mask -> bits of (mask == fg), one word per warp; also initialises the heads of the INVERSE image's runs
// grid = (ceil(W / 256), H, batch), block = 256 (8 words)
__global__ void __launch_bounds__(256) fg_bits_kernel(const uint8_t* __restrict__ mask, int H, int W, int wpitch, const FgSpec fgs,
                                                       uint32_t* __restrict__ bits, int* __restrict__ L_all, int* __restrict__ area_all,
                                                       uint8_t* __restrict__ flag_all) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    const bool in = x < W;
    const bool fg = in && mask[((size_t)fgs.src(blockIdx.z) * H + y) * W + x] == fgs.fg(blockIdx.z);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, fg);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < wpitch) {
        bits[((size_t)blockIdx.z * H + y) * wpitch + (x >> 5)] = b;
        const size_t slice = (size_t)blockIdx.z * H * W;
        ccl::init_heads<true>(b, W, y, x >> 5, L_all + slice, area_all + slice, flag_all + slice);   // runs of the inverse image
    }
}

// filled = fg | (runs of inverse components that are holes).  One thread per word.
__global__ void __launch_bounds__(ccl::kThreads) fill_kernel(const uint32_t* __restrict__ fg_bits, int H, int W, int wpitch, int min_area,
                                                              const int* __restrict__ L_all, const int* __restrict__ area_all,
                                                              const uint8_t* __restrict__ flag_all, uint32_t* __restrict__ filled) {
    MS_CCL_WORD_COORDS();
    const size_t slice = (size_t)sl * H * W;
    const size_t wi = ((size_t)sl * H + y) * wpitch + wx;
    const uint32_t f = fg_bits[wi];
    const uint32_t inv = ~f & ccl::valid_mask(W, wx);
    uint32_t out = f;
    uint32_t h = ccl::head_mask(inv);
    while (h) {
        const int x = __ffs((int)h) - 1;
        h &= h - 1;
        const int r = L_all[slice + (size_t)y * W + wx * 32 + x];             // root (resolve_kernel ran)
        if (flag_all[slice + r] == 0 && area_all[slice + r] < min_area) out |= ccl::run_mask(inv, x);   // postprocess.cpp:40-41
    }
    filled[wi] = out;
}

// 3x3 open on bits + heads of the opened image.  One thread per word; window = positions -2 .. 33 of the word.
__global__ void __launch_bounds__(ccl::kThreads) open_kernel(const uint32_t* __restrict__ filled, int H, int W, int wpitch,
                                                              uint32_t* __restrict__ opened, int* __restrict__ L_all, int* __restrict__ area_all) {
    MS_CCL_WORD_COORDS();
    const uint32_t* F = filled + (size_t)sl * H * wpitch;
    // validity of window bit k (position wx*32 + k - 2)
    unsigned long long valid = 0;
    {
        const long long lo = -(long long)wx * 32 + 2, hi = (long long)W - (long long)wx * 32 + 2;   // k in [lo, hi)
        const int a = (int)(lo < 0 ? 0 : lo), b = (int)(hi > 36 ? 36 : hi);
        if (b > a) valid = ((b >= 64 ? ~0ull : ((1ull << b) - 1ull)) & ~((1ull << a) - 1ull));
    }
    auto window = [&](int yy) -> unsigned long long {     // fill bits, outside the image = 1 (never wins erode's min)
        if (yy < 0 || yy >= H) return ~0ull;
        const uint32_t c = __ldg(F + (size_t)yy * wpitch + wx);
        const uint32_t l = wx > 0 ? __ldg(F + (size_t)yy * wpitch + wx - 1) : 0u;
        const uint32_t r = wx + 1 < wpitch ? __ldg(F + (size_t)yy * wpitch + wx + 1) : 0u;
        const unsigned long long v = (unsigned long long)(l >> 30) | ((unsigned long long)c << 2) | ((unsigned long long)(r & 3u) << 34);
        return v | ~valid;
    };
    unsigned long long eh[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const unsigned long long v = window(y - 2 + i);
        eh[i] = v & (v << 1) & (v >> 1);                   // horizontal erode
    }
    unsigned long long op = 0;
#pragma unroll
    for (int i = 1; i <= 3; ++i) {                         // eroded rows y-1, y, y+1
        const int yy = y - 2 + i;
        unsigned long long er = 0;                         // rows outside the image never win dilate's max
        if (yy >= 0 && yy < H) er = eh[i - 1] & eh[i] & eh[i + 1] & valid;
        op |= er | (er << 1) | (er >> 1);                  // horizontal dilate, accumulated vertically
    }
    const uint32_t o = (uint32_t)(op >> 2) & ccl::valid_mask(W, wx);
    opened[((size_t)sl * H + y) * wpitch + wx] = o;
    const size_t slice = (size_t)sl * H * W;
    uint32_t h = ccl::head_mask(o);
    while (h) {
        const int x = __ffs((int)h) - 1;
        h &= h - 1;
        const size_t p = slice + (size_t)y * W + wx * 32 + x;
        L_all[p] = y * W + wx * 32 + x;
        area_all[p] = 0;
    }
}

// kept = runs of opened components with area >= min_area, expanded to the u8 output.  One thread per word (32 bytes out).
__global__ void __launch_bounds__(ccl::kThreads) keep_kernel(const uint32_t* __restrict__ opened, int H, int W, int wpitch, int min_area,
                                                              const FgSpec fgs, const int* __restrict__ L_all, const int* __restrict__ area_all,
                                                              uint8_t* __restrict__ out) {
    MS_CCL_WORD_COORDS();
    const int fg_value = fgs.fg(sl);
    const size_t slice = (size_t)sl * H * W;
    const uint32_t o = opened[((size_t)sl * H + y) * wpitch + wx];
    uint32_t keep = 0;
    uint32_t h = ccl::head_mask(o);
    while (h) {
        const int x = __ffs((int)h) - 1;
        h &= h - 1;
        const int r = L_all[slice + (size_t)y * W + wx * 32 + x];
        if (area_all[slice + r] >= min_area) keep |= ccl::run_mask(o, x);      // postprocess.cpp:70-76
    }
    uint8_t* dst = out + slice + (size_t)y * W + wx * 32;
    const uint32_t v = (uint32_t)fg_value;
    if (wx * 32 + 32 <= W && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        uint32_t wds[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t nib = (keep >> (4 * j)) & 15u;
            wds[j] = ((nib & 1u) ? v : 0u) | ((nib & 2u) ? v << 8 : 0u) | ((nib & 4u) ? v << 16 : 0u) | ((nib & 8u) ? v << 24 : 0u);
        }
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
    } else {
        for (int j = 0; j < 32 && wx * 32 + j < W; ++j) dst[j] = (keep >> j) & 1u ? (uint8_t)fg_value : (uint8_t)0;
    }
}

}  // namespace

void postprocess_launch(PostprocessWs& ws, const uint8_t* d_in, uint8_t* d_out, int h, int w, int batch, int fg_value,
                        float min_area_ratio, cudaStream_t st, const FgSpec* multi) {
    const FgSpec fgs = multi ? *multi : FgSpec::single(fg_value, batch);
    MS_REQUIRE(h > 0 && w > 0 && batch > 0 && h <= 65535 && batch <= 65535, MS_ERR_ARG, "postprocess: bad shape");
    MS_REQUIRE((int64_t)h * w < ((int64_t)1 << 31), MS_ERR_ARG, "postprocess: slice too large");
    if (slice_fused_supported(h, w)) {     // one CTA per slice, everything on chip (slice_fused.cuh)
        slice_fused_launch(ws.fused, nullptr, nullptr, d_in, d_out, h, w, batch, true, false, fg_value, min_area_ratio, 0, st, multi);
        return;
    }
    const size_t n = (size_t)h * w, nb = n * batch;
    const int wpitch = cdiv(w, 32);
    const size_t nwords = (size_t)batch * h * wpitch;
    // src/postprocess.cpp:30,66: static_cast<int>(w * h * MIN_AREA_RATIO) -- int product, float multiply, truncate
    const int min_area = static_cast<int>(static_cast<float>(w * h) * min_area_ratio);
    ws.ccl.labels.reserve(nb * 4);
    ws.ccl.area.reserve(nb * 4);
    ws.ccl.flag.reserve(nb);
    ws.bin_a.reserve(nwords * 4);   // fg bits, later opened bits
    ws.bin_b.reserve(nwords * 4);   // filled bits
    int* L = ws.ccl.labels.as<int>();
    int* A = ws.ccl.area.as<int>();
    uint8_t* F = ws.ccl.flag.as<uint8_t>();
    uint32_t* Bfg = ws.bin_a.as<uint32_t>();
    uint32_t* Bfill = ws.bin_b.as<uint32_t>();
    const dim3 gw = ccl::grid_for(h, wpitch, batch);

    fg_bits_kernel<<<dim3(cdiv(w, 256), h, batch), 256, 0, st>>>(d_in, h, w, wpitch, fgs, Bfg, L, A, F);
    MS_LAUNCH_CHECK();
    ccl::merge_kernel<8, true><<<gw, ccl::kThreads, 0, st>>>(Bfg, h, w, wpitch, L);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<true><<<gw, ccl::kThreads, 0, st>>>(Bfg, h, w, wpitch, L, A, F);
    MS_LAUNCH_CHECK();
    fill_kernel<<<gw, ccl::kThreads, 0, st>>>(Bfg, h, w, wpitch, min_area, L, A, F, Bfill);
    MS_LAUNCH_CHECK();
    uint32_t* Bopen = Bfg;   // the fg bits are dead after fill_kernel
    open_kernel<<<gw, ccl::kThreads, 0, st>>>(Bfill, h, w, wpitch, Bopen, L, A);
    MS_LAUNCH_CHECK();
    ccl::merge_kernel<8, false><<<gw, ccl::kThreads, 0, st>>>(Bopen, h, w, wpitch, L);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<false><<<gw, ccl::kThreads, 0, st>>>(Bopen, h, w, wpitch, L, A, nullptr);
    MS_LAUNCH_CHECK();
    keep_kernel<<<gw, ccl::kThreads, 0, st>>>(Bopen, h, w, wpitch, min_area, fgs, L, A, d_out);
    MS_LAUNCH_CHECK();
}

}  // namespace ms
