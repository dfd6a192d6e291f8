// mask2polygon.cu -- K6: mask -> external contours (bit-exact with cv::findContours).
//
// Replaces Mask2Polygon::extract_contours + map_contour_points
// (/root/reference/src/mask2polygon.cpp:29-36, 41-63):
//     threshold(mask, 127) ; findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) ; (int)(pt * scale)
// following the closed form in SURVEY.md section 8(c):
//   (1) 8-connected foreground components  -> ccl (root = raster-first pixel = contour start)
//   (2) external test: the pixel left of the start must belong to the 4-connected background
//       component that reaches the image frame                       -> second ccl + border flag
//   (3) contours ordered by descending start index                   -> count / scan / scatter
//   (4)-(6) border following + CHAIN_APPROX_SIMPLE                   -> contour_trace.cuh, one
//       thread per contour on the bit-packed foreground (a whole slice of it in shared memory when it
//       fits), two passes (count, then emit at scanned offsets)
//   (7) coordinate mapping (int)(x * (double)orig_w / w)             -> fused into the emit pass
// No host round trip happens between these launches: all sizes live in `header` on the device, so
// the whole stage is CUDA-graph capturable and only the caller decides when to synchronise.
// The u8 mask is read once (-> one bit per pixel); labelling is the run-based union-find of ccl.cuh.
// Algorithmic bytes: H*W (mask read) + 8 B per vertex + 4 B per contour offset.
#include "ccl.cuh"
#include "contour_trace.cuh"

namespace ms {

namespace {

constexpr size_t kTraceSmemMax = 200 * 1024;   // slices up to ~1264 x 1264 trace out of shared memory

// mask -> bits of (mask > thr), one word per warp.  grid = (ceil(W / 256), H, batch), block = 256 (8 words)
__global__ void __launch_bounds__(256) thr_bits_kernel(const uint8_t* __restrict__ mask, int H, int W, int wpitch, int thr,
                                                        uint32_t* __restrict__ bits) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    const bool fg = x < W && mask[((size_t)blockIdx.z * H + y) * W + x] > thr;   // src/mask2polygon.cpp:31 threshold(127)
    const unsigned b = __ballot_sync(0xFFFFFFFFu, fg);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < wpitch) bits[((size_t)blockIdx.z * H + y) * wpitch + (x >> 5)] = b;
}

// number of external contour starts among the run heads of word `wx` of row `y`, and (optionally) their pixel indices.
// A component's root is its raster-first pixel, always a run head.  External <=> the pixel left of it belongs to a
// background component that reaches the image frame (x == 0: the frame itself).  SURVEY.md section 8(c) clause (2).
__device__ __forceinline__ int external_starts(const uint32_t* __restrict__ B, const int* __restrict__ Lfg, const int* __restrict__ Lbg,
                                               const uint8_t* __restrict__ bg_flag, int W, int wpitch, int y, int wx, int* out /*<= 16*/) {
    const uint32_t fgw = __ldg(B + (size_t)y * wpitch + wx);
    uint32_t h = ccl::head_mask(fgw);
    int n = 0;
    while (h) {
        const int x = __ffs((int)h) - 1;
        h &= h - 1;
        const int X = wx * 32 + x, p = y * W + X;
        if (Lfg[p] != p) continue;                       // not a component root
        bool ext = X == 0;
        if (!ext) {
            // left neighbour is background (p is raster-first): root of its background run
            const int wl = (X - 1) >> 5;
            const uint32_t bgw = ~__ldg(B + (size_t)y * wpitch + wl) & ccl::valid_mask(W, wl);
            const int hp = y * W + (wl << 5) + ccl::run_head_bit(bgw, (X - 1) & 31);
            ext = bg_flag[Lbg[hp]] != 0;
        }
        if (ext) {
            if (out) out[n] = p;
            ++n;
        }
    }
    return n;
}

// word-parallel: block b covers words [256 b, 256 b + 256) of its slice.  grid = (blocks_per_slice, batch)
__global__ void __launch_bounds__(256) count_starts_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ Lfg,
                                                            const int* __restrict__ Lbg, const uint8_t* __restrict__ bg_flag, int H, int W,
                                                            int wpitch, int* __restrict__ block_counts) {
    __shared__ int wsum[8];
    const int widx = blockIdx.x * 256 + threadIdx.x;
    const size_t slice = (size_t)blockIdx.y * H * W;
    int c = 0;
    if (widx < H * wpitch)
        c = external_starts(bits + (size_t)blockIdx.y * H * wpitch, Lfg + slice, Lbg + slice, bg_flag + slice, W, wpitch, widx / wpitch,
                            widx % wpitch, nullptr);
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += wsum[i];
        block_counts[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// Block-wide exclusive scan of `v` (blockDim.x == 1024); returns the exclusive prefix, *total gets the sum.
__device__ __forceinline__ int block_exscan_1024(int v, int* total) {
    __shared__ int wsum[32];
    __shared__ int tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += t;
        }
        wsum[lane] = winc - w;
        if (lane == 31) tot = winc;
    }
    __syncthreads();
    const int r = inc - v + wsum[warp];
    *total = tot;
    __syncthreads();
    return r;
}

// grid = batch, block = 1024: in-place exclusive scan of each slice's block counts; slice_total[b] = sum
__global__ void __launch_bounds__(1024) scan_blocks_kernel(int* __restrict__ block_counts, int blocks_per_slice,
                                                            int* __restrict__ slice_total) {
    int* c = block_counts + (size_t)blockIdx.x * blocks_per_slice;
    int carry = 0;
    for (int base = 0; base < blocks_per_slice; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < blocks_per_slice ? c[i] : 0;
        int tot;
        const int ex = block_exscan_1024(v, &tot);
        if (i < blocks_per_slice) c[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) slice_total[blockIdx.x] = carry;
}

// 1 block of 1024: slice_start = exclusive scan of slice totals (slice_start[batch] = n_contours)
__global__ void __launch_bounds__(1024) scan_slices_kernel(const int* __restrict__ slice_total, int batch,
                                                            int* __restrict__ slice_start, long long* __restrict__ header) {
    int carry = 0;
    for (int base = 0; base < batch; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < batch ? slice_total[i] : 0;
        int tot;
        const int ex = block_exscan_1024(v, &tot);
        if (i < batch) slice_start[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        slice_start[batch] = carry;
        header[0] = carry;  // n_contours
        header[1] = 0;      // n_points (set by scan_points_kernel)
        header[2] = 0;      // overflow flags
        header[3] = 0;      // trace errors
    }
}

// Scatter start pixels in *descending* raster order per slice (SURVEY.md section 8(c) clause (3)).
__global__ void __launch_bounds__(256) write_starts_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ Lfg,
                                                            const int* __restrict__ Lbg, const uint8_t* __restrict__ bg_flag, int H, int W,
                                                            int wpitch, const int* __restrict__ block_offsets,
                                                            const int* __restrict__ slice_start, int cap_contours,
                                                            int* __restrict__ starts, int* __restrict__ start_slice) {
    __shared__ int wsum[8];
    const int b = blockIdx.y;
    const int widx = blockIdx.x * 256 + threadIdx.x;
    const size_t slice = (size_t)b * H * W;
    int found[16];
    int c = 0;
    if (widx < H * wpitch)
        c = external_starts(bits + (size_t)b * H * wpitch, Lfg + slice, Lbg + slice, bg_flag + slice, W, wpitch, widx / wpitch, widx % wpitch, found);
    // exclusive prefix of c over the block (raster order = thread order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int rank = inc - c + block_offsets[(size_t)b * gridDim.x + blockIdx.x];
    for (int i = 0; i < warp; ++i) rank += wsum[i];
    const int first = slice_start[b], total = slice_start[b + 1] - first;
    for (int i = 0; i < c; ++i) {
        const int pos = first + (total - 1 - (rank + i));
        if (pos < cap_contours) {
            starts[pos] = found[i];
            start_slice[pos] = b;
        }
    }
}

struct CountEmit {
    __device__ void operator()(int, int) const {}
};
struct WriteEmit {
    int2* dst;
    double sx, sy;
    mutable int i;
    __device__ void operator()(int x, int y) const {
        // src/mask2polygon.cpp:54-55: static_cast<int>(pt.x * scale_x)
        dst[i++] = make_int2((int)__dmul_rn((double)x, sx), (int)__dmul_rn((double)y, sy));
    }
};

// 8-neighbour foreground code of a pixel from bit-packed rows.  PADDED: `bits` has a zero word / zero row on every
// side (shared-memory copy); otherwise bounds are checked (global memory, large slices).
// Codes: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE ; bit0 of a 3-bit row window = x-1, bit1 = x, bit2 = x+1
__device__ __forceinline__ unsigned code_from_rows(unsigned up, unsigned cu, unsigned dn) {
    return ((cu >> 2) & 1u) | (((up >> 2) & 1u) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) | ((cu & 1u) << 4) |
           ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | (((dn >> 2) & 1u) << 7);
}
struct PaddedBitsCode {
    const uint32_t* bits;   // (H + 2) rows x pitch words, row 0 / word 0 are the zero frame
    int pitch;
    __device__ __forceinline__ unsigned operator()(int, int x, int y) const {
        const int X = x + 31, w = X >> 5, sh = X & 31;              // pixel x lives at bit x + 32 of the padded row
        const uint32_t* r = bits + (size_t)y * pitch + w;            // padded row y <-> image row y - 1
        return code_from_rows(__funnelshift_r(r[0], r[1], sh) & 7u, __funnelshift_r(r[pitch], r[pitch + 1], sh) & 7u,
                              __funnelshift_r(r[2 * pitch], r[2 * pitch + 1], sh) & 7u);
    }
};
// Large slices (the bit image does not fit in shared memory): one CTA per contour keeps a 256 x 256-pixel WINDOW of the
// bit image in shared memory, centred on the walk; thread 0 follows the border (six LDS per step) until it leaves the
// window's interior, then the CTA re-centres the window and the walk resumes (contour_trace.cuh: trace_run).  A long
// contour (the stress masks have ~10^5-step borders) therefore walks at shared-memory latency, not at one dependent
// L2 access per step, and all contours of all slices walk concurrently.
constexpr int kWinW = 256, kWinH = 256, kWinPitch = kWinW / 32 + 1;   // +1 word: funnel shifts read one word ahead
struct WindowCode {
    const uint32_t* win;   // kWinH rows x kWinPitch words
    int x0, y0;            // image coordinates of the window's first pixel (x0 a multiple of 32, may be negative)
    __device__ __forceinline__ unsigned operator()(int, int x, int y) const {
        const int cx = x - x0 - 1, w = cx >> 5, sh = cx & 31;          // bits cx .. cx+2 = x-1 .. x+1
        const uint32_t* r = win + (y - y0 - 1) * kWinPitch + w;
        return code_from_rows(__funnelshift_r(r[0], r[1], sh) & 7u, __funnelshift_r(r[kWinPitch], r[kWinPitch + 1], sh) & 7u,
                              __funnelshift_r(r[2 * kWinPitch], r[2 * kWinPitch + 1], sh) & 7u);
    }
};
struct WindowInside {
    int x0, y0;
    __device__ __forceinline__ bool operator()(int x, int y) const {
        return x - 1 >= x0 && x + 1 < x0 + kWinW && y - 1 >= y0 && y + 1 < y0 + kWinH;
    }
};

template <bool EMIT>
__global__ void __launch_bounds__(64) trace_window_kernel(const uint32_t* __restrict__ fgbits, int H, int W, int wpitch,
                                                           const int* __restrict__ starts, const int* __restrict__ start_slice,
                                                           long long* __restrict__ header, int cap_contours, int* __restrict__ npts,
                                                           long long cap_points, double sx, double sy, int2* __restrict__ xy) {
    __shared__ uint32_t win[kWinH * kWinPitch];
    __shared__ int s_org[2], s_status;
    const int c = blockIdx.x;
    const long long n = header[0] < cap_contours ? header[0] : cap_contours;
    if (c >= n) return;
    if (EMIT && header[1] > cap_points) {   // caller's buffer too small: write nothing, flag it
        if (c == 0 && threadIdx.x == 0) header[2] |= 2;
        return;
    }
    const uint32_t* B = fgbits + (size_t)start_slice[c] * H * wpitch;
    TraceState st;
    trace_begin(st, W, starts[c]);
    CountEmit count_emit;
    WriteEmit write_emit{xy + (EMIT ? npts[c] : 0), sx, sy, 0};
    for (int round = 0; round < (1 << 20); ++round) {
        if (threadIdx.x == 0) {
            s_org[0] = ((st.x - kWinW / 2) >> 5) << 5;
            s_org[1] = st.y - kWinH / 2;
        }
        __syncthreads();
        const int x0 = s_org[0], y0 = s_org[1];
        for (int i = threadIdx.x; i < kWinH * kWinPitch; i += 64) {
            const int ry = i / kWinPitch, wq = i - ry * kWinPitch;
            const int gy = y0 + ry, gw = (x0 >> 5) + wq;
            win[i] = (gy >= 0 && gy < H && gw >= 0 && gw < wpitch) ? __ldg(B + (size_t)gy * wpitch + gw) : 0u;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const WindowCode code{win, x0, y0};
            const WindowInside inside{x0, y0};
            s_status = EMIT ? trace_run(code, W, st, 8 * H * W + 8, write_emit, inside) : trace_run(code, W, st, 8 * H * W + 8, count_emit, inside);
        }
        __syncthreads();
        if (s_status != 0) break;
    }
    if (threadIdx.x == 0 && !EMIT) {
        if (s_status != 1) atomicAdd((unsigned long long*)&header[3], 1ull);
        npts[c] = s_status == 1 ? st.n : 0;
    }
}

// Shared-memory variant: one CTA per slice keeps the slice's bit-packed foreground (H*W/8 bytes, padded by a zero word /
// zero row on every side) in shared memory, so a border-following step costs six LDS instead of dependent global
// loads.  Thread t traces contours slice_start[b] + t, + blockDim, ...
template <bool EMIT>
__global__ void __launch_bounds__(128) trace_smem_kernel(const uint32_t* __restrict__ fgbits, int H, int W, int wpitch,
                                                          const int* __restrict__ starts, const int* __restrict__ slice_start,
                                                          long long* __restrict__ header, int cap_contours, int* __restrict__ npts,
                                                          long long cap_points, double sx, double sy, int2* __restrict__ xy) {
    extern __shared__ uint32_t sbits[];
    const int b = blockIdx.x;
    const int pitch = wpitch + 2;
    const int c_lo = slice_start[b], c_hi = min(slice_start[b + 1], cap_contours);
    if (c_lo >= c_hi) return;
    if (EMIT && header[1] > cap_points) {
        if (threadIdx.x == 0) header[2] |= 2;
        return;
    }
    for (int i = threadIdx.x; i < (H + 2) * pitch; i += blockDim.x) {
        const int r = i / pitch, c = i % pitch;
        sbits[i] = (r >= 1 && r <= H && c >= 1 && c <= wpitch) ? fgbits[((size_t)b * H + (r - 1)) * wpitch + (c - 1)] : 0u;
    }
    __syncthreads();
    const PaddedBitsCode code{sbits, pitch};
    for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
        if (EMIT) {
            trace_contour_fn(code, W, starts[c], 8 * H * W + 8, WriteEmit{xy + npts[c], sx, sy, 0});
        } else {
            const int cnt = trace_contour_fn(code, W, starts[c], 8 * H * W + 8, CountEmit{});
            if (cnt < 0) atomicAdd((unsigned long long*)&header[3], 1ull);
            npts[c] = cnt < 0 ? 0 : cnt;
        }
    }
}

// 1 block of 1024: in-place exclusive scan of npts[0..n) ; npts[n] = header[1] = total
__global__ void __launch_bounds__(1024) scan_points_kernel(int* __restrict__ npts, int cap_contours, long long* __restrict__ header) {
    const long long n64 = header[0];
    const int n = (int)(n64 < cap_contours ? n64 : cap_contours);
    long long carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? npts[i] : 0;
        int tot;
        const int ex = block_exscan_1024(v, &tot);
        if (i < n) npts[i] = (int)(carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0) {
        npts[n] = (int)carry;
        header[1] = carry;
        if (n64 > cap_contours) header[2] |= 1;
        if (carry > 0x7FFFFFFFll) header[2] |= 4;
    }
}

template <bool EMIT>
void launch_trace(M2pWs& ws, PolyDev& P, int h, int w, int batch, double sx, double sy, cudaStream_t st) {
    const int wpitch = cdiv(w, 32);
    const size_t trace_smem = (size_t)(h + 2) * (wpitch + 2) * 4;
    long long* header = P.header.as<long long>();
    if (trace_smem <= kTraceSmemMax) {
        static bool attr = false;
        if (!attr) {
            MS_CUDA(cudaFuncSetAttribute(trace_smem_kernel<EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTraceSmemMax));
            attr = true;
        }
        trace_smem_kernel<EMIT><<<batch, 128, trace_smem, st>>>(ws.fgbits.as<uint32_t>(), h, w, wpitch, P.starts.as<int>(),
                                                               P.slice_start.as<int>(), header, (int)P.cap_contours, P.npts.as<int>(),
                                                               (long long)P.cap_points, sx, sy, P.xy.as<int2>());
    } else {
        trace_window_kernel<EMIT><<<(unsigned)P.cap_contours, 64, 0, st>>>(ws.fgbits.as<uint32_t>(), h, w, wpitch, P.starts.as<int>(),
                                                                          P.start_slice.as<int>(), header, (int)P.cap_contours,
                                                                          P.npts.as<int>(), (long long)P.cap_points, sx, sy, P.xy.as<int2>());
    }
    MS_LAUNCH_CHECK();
}

}  // namespace

void m2p_phase_a(M2pWs& ws, PolyDev& P, const uint8_t* d_mask, int h, int w, int batch, int threshold, cudaStream_t st) {
    MS_REQUIRE(h > 0 && w > 0 && batch > 0 && h <= 65535 && batch <= 32767, MS_ERR_ARG, "mask2polygon: bad shape");
    MS_REQUIRE((int64_t)h * w <= (int64_t)(0x7FFFFFFF - 8) / 8, MS_ERR_ARG, "mask2polygon: slice too large");
    const int n = h * w;
    const size_t nb = (size_t)n * batch;
    if (P.cap_contours == 0) P.cap_contours = std::max<int64_t>(1024, 64 * (int64_t)batch);
    if (P.cap_points == 0) P.cap_points = std::max<int64_t>(65536, 4096 * (int64_t)batch);
    const int wpitch = cdiv(w, 32);
    ws.fg.labels.reserve(nb * 4);
    ws.bg.labels.reserve(nb * 4);
    ws.bg.flag.reserve(nb);
    ws.fgbits.reserve((size_t)batch * h * wpitch * 4);
    const int bps = cdiv(h * wpitch, 256);
    P.block_counts.reserve(((size_t)bps * batch + batch + 1) * 4);
    P.slice_start.reserve(((size_t)batch + 1) * 4);
    P.starts.reserve((size_t)P.cap_contours * 4);
    P.start_slice.reserve((size_t)P.cap_contours * 4);
    P.npts.reserve(((size_t)P.cap_contours + 1) * 4);
    P.xy.reserve((size_t)P.cap_points * 8);
    P.header.reserve(4 * sizeof(long long));
    int* Lfg = ws.fg.labels.as<int>();
    int* Lbg = ws.bg.labels.as<int>();
    uint8_t* flag = ws.bg.flag.as<uint8_t>();
    uint32_t* B = ws.fgbits.as<uint32_t>();
    int* bc = P.block_counts.as<int>();
    int* slice_total = bc + (size_t)bps * batch;
    long long* header = P.header.as<long long>();
    const dim3 gw = ccl::grid_for(h, wpitch, batch);

    thr_bits_kernel<<<dim3(cdiv(w, 256), h, batch), 256, 0, st>>>(d_mask, h, w, wpitch, threshold, B);
    MS_LAUNCH_CHECK();
    ccl::heads_kernel<false><<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lfg, nullptr, nullptr);      // 8-connected foreground
    MS_LAUNCH_CHECK();
    ccl::heads_kernel<true><<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lbg, nullptr, flag);          // 4-connected background
    MS_LAUNCH_CHECK();
    ccl::merge_kernel<8, false><<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lfg);
    MS_LAUNCH_CHECK();
    ccl::merge_kernel<4, true><<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lbg);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<false><<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lfg, nullptr, nullptr);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<true><<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lbg, nullptr, flag);
    MS_LAUNCH_CHECK();
    count_starts_kernel<<<dim3(bps, batch), 256, 0, st>>>(B, Lfg, Lbg, flag, h, w, wpitch, bc);
    MS_LAUNCH_CHECK();
    scan_blocks_kernel<<<batch, 1024, 0, st>>>(bc, bps, slice_total);
    MS_LAUNCH_CHECK();
    scan_slices_kernel<<<1, 1024, 0, st>>>(slice_total, batch, P.slice_start.as<int>(), header);
    MS_LAUNCH_CHECK();
    write_starts_kernel<<<dim3(bps, batch), 256, 0, st>>>(B, Lfg, Lbg, flag, h, w, wpitch, bc, P.slice_start.as<int>(), (int)P.cap_contours,
                                                          P.starts.as<int>(), P.start_slice.as<int>());
    MS_LAUNCH_CHECK();
    launch_trace<false>(ws, P, h, w, batch, 1.0, 1.0, st);
    scan_points_kernel<<<1, 1024, 0, st>>>(P.npts.as<int>(), (int)P.cap_contours, header);
    MS_LAUNCH_CHECK();
}

void m2p_phase_b(M2pWs& ws, PolyDev& P, int h, int w, int batch, int orig_w, int orig_h, cudaStream_t st) {
    // src/mask2polygon.cpp:199-200
    const double sx = static_cast<double>(orig_w) / w;
    const double sy = static_cast<double>(orig_h) / h;
    launch_trace<true>(ws, P, h, w, batch, sx, sy, st);
}

}  // namespace ms
