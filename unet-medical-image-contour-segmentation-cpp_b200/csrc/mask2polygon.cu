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
//       thread per contour, two passes (count, then emit at scanned offsets)
//   (7) coordinate mapping (int)(x * (double)orig_w / w)             -> fused into the emit pass
// No host round trip happens between these launches: all sizes live in `header` on the device, so
// the whole stage is CUDA-graph capturable and only the caller decides when to synchronise.
// Algorithmic bytes: H*W (mask read) + 8 B per vertex + 4 B per contour offset.
#include "ccl.cuh"
#include "contour_trace.cuh"

namespace ms {

namespace {

constexpr int TW = 32, TH = 8;
constexpr size_t kTraceSmemMax = 200 * 1024;   // slices up to ~1264 x 1264 trace out of shared memory

// Fused: threshold -> 8-neighbour code, foreground label init, background label init + flag clear.
// grid = (ceil(W/32), ceil(H/8), batch), block = 256.
__global__ void __launch_bounds__(256) m2p_init_kernel(const uint8_t* __restrict__ mask, int H, int W, int thr,
                                                        int* __restrict__ Lfg, int* __restrict__ Lbg,
                                                        uint8_t* __restrict__ bg_flag, uint8_t* __restrict__ nb,
                                                        uint32_t* __restrict__ fgbits, int wpitch) {
    __shared__ uint8_t S[TH + 2][TW + 2];
    const size_t slice = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += 256) {
        const int ly = i / (TW + 2), lx = i % (TW + 2);
        const int x = x0 + lx - 1, y = y0 + ly - 1;
        S[ly][lx] = (x >= 0 && x < W && y >= 0 && y < H) ? (uint8_t)(mask[slice + (size_t)y * W + x] > thr) : (uint8_t)0;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    const bool in = x < W && y < H;
    const bool fg = in && S[ly + 1][lx + 1];
    const bool bg = in && !fg;
    const unsigned fbits = __ballot_sync(0xFFFFFFFFu, fg);
    const unsigned bbits = __ballot_sync(0xFFFFFFFFu, bg);
    // bit-packed foreground, one word per 32-pixel segment (the trace kernels keep a whole slice of it in smem)
    if (lx == 0 && y < H) fgbits[((size_t)blockIdx.z * H + y) * wpitch + blockIdx.x] = fbits;
    if (!in) return;
    const int p = y * W + x;
    const unsigned below = (1u << lx) - 1u;
    unsigned code = 0;
    int lf = -1, lb = -1;
    if (fg) {
        const unsigned z = ~fbits & below;
        lf = p - lx + (z ? 32 - __clz(z) : 0);
        code = (unsigned)S[ly + 1][lx + 2] | ((unsigned)S[ly][lx + 2] << 1) | ((unsigned)S[ly][lx + 1] << 2) |
               ((unsigned)S[ly][lx] << 3) | ((unsigned)S[ly + 1][lx] << 4) | ((unsigned)S[ly + 2][lx] << 5) |
               ((unsigned)S[ly + 2][lx + 1] << 6) | ((unsigned)S[ly + 2][lx + 2] << 7);
    } else {
        const unsigned z = ~bbits & below;
        lb = p - lx + (z ? 32 - __clz(z) : 0);
    }
    Lfg[slice + p] = lf;
    Lbg[slice + p] = lb;
    bg_flag[slice + p] = 0;
    nb[slice + p] = (uint8_t)code;
}

__device__ __forceinline__ bool is_external_start(const int* Lfg, const int* Lbg, const uint8_t* bg_flag, int W, int p) {
    if (Lfg[p] != p) return false;                 // not a component root
    if (p % W == 0) return true;                    // left neighbour is the frame itself
    return bg_flag[Lbg[p - 1]] != 0;                // left neighbour is background (p is raster-first)
}

// grid = (blocks_per_slice, batch), block = 256: block b covers pixels [256 b, 256 b + 256) of its slice
__global__ void __launch_bounds__(256) count_starts_kernel(const int* __restrict__ Lfg, const int* __restrict__ Lbg,
                                                            const uint8_t* __restrict__ bg_flag, int W, int n_per_slice,
                                                            int* __restrict__ block_counts) {
    const size_t slice = (size_t)blockIdx.y * n_per_slice;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const bool s = p < n_per_slice && is_external_start(Lfg + slice, Lbg + slice, bg_flag + slice, W, p);
    const int c = __syncthreads_count(s);
    if (threadIdx.x == 0) block_counts[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = c;
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

// Scatter start pixels in *descending* raster order per slice.
__global__ void __launch_bounds__(256) write_starts_kernel(const int* __restrict__ Lfg, const int* __restrict__ Lbg,
                                                            const uint8_t* __restrict__ bg_flag, int W, int n_per_slice,
                                                            const int* __restrict__ block_offsets,
                                                            const int* __restrict__ slice_start, int cap_contours,
                                                            int* __restrict__ starts, int* __restrict__ start_slice) {
    __shared__ int wcount[8];
    const int b = blockIdx.y;
    const size_t slice = (size_t)b * n_per_slice;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const bool s = p < n_per_slice && is_external_start(Lfg + slice, Lbg + slice, bg_flag + slice, W, p);
    const unsigned bits = __ballot_sync(0xFFFFFFFFu, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wcount[warp] = __popc(bits);
    __syncthreads();
    if (!s) return;
    int rank = __popc(bits & ((1u << lane) - 1u));
    for (int i = 0; i < warp; ++i) rank += wcount[i];
    rank += block_offsets[(size_t)b * gridDim.x + blockIdx.x];
    const int first = slice_start[b], total = slice_start[b + 1] - first;
    const int pos = first + (total - 1 - rank);
    if (pos < cap_contours) {
        starts[pos] = p;
        start_slice[pos] = b;
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

// One thread per contour.  npts[c] <- number of kept vertices.
__global__ void __launch_bounds__(128) trace_count_kernel(const uint8_t* __restrict__ nb, int W, int n_per_slice,
                                                           const int* __restrict__ starts, const int* __restrict__ start_slice,
                                                           long long* __restrict__ header, int cap_contours,
                                                           int* __restrict__ npts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = header[0] < cap_contours ? header[0] : cap_contours;
    if (c >= n) return;
    const int cnt = trace_contour(nb + (size_t)start_slice[c] * n_per_slice, W, starts[c], 8 * n_per_slice + 8, CountEmit{});
    if (cnt < 0) atomicAdd((unsigned long long*)&header[3], 1ull);
    npts[c] = cnt < 0 ? 0 : cnt;
}

// ---- shared-memory variant: one CTA per slice keeps the slice's bit-packed foreground (H*W/8 bytes, padded by a
// zero word / zero row on every side) in shared memory, so a border-following step costs six LDS instead of a
// dependent global load.  Thread t traces contours slice_start[b] + t, + blockDim, ...
struct BitsCode {
    const uint32_t* bits;   // (H + 2) rows x pitch words, row 0 / word 0 are the zero frame
    int pitch;
    __device__ __forceinline__ unsigned operator()(int, int x, int y) const {
        // bits x-1 .. x+1 of rows y-1, y, y+1; pixel x lives at bit position x + 32 of the padded row
        const int X = x + 31, w = X >> 5, sh = X & 31;
        const uint32_t* r = bits + (size_t)y * pitch + w;          // padded row y   <-> image row y-1
        const unsigned up = __funnelshift_r(r[0], r[1], sh) & 7u;
        const unsigned cu = __funnelshift_r(r[pitch], r[pitch + 1], sh) & 7u;
        const unsigned dn = __funnelshift_r(r[2 * pitch], r[2 * pitch + 1], sh) & 7u;
        // 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE ; bit0 = x-1, bit1 = x, bit2 = x+1
        return ((cu >> 2) & 1u) | (((up >> 2) & 1u) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) | ((cu & 1u) << 4) |
               ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | (((dn >> 2) & 1u) << 7);
    }
};

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
    const BitsCode code{sbits, pitch};
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

__global__ void __launch_bounds__(128) trace_emit_kernel(const uint8_t* __restrict__ nb, int W, int n_per_slice,
                                                          const int* __restrict__ starts, const int* __restrict__ start_slice,
                                                          const int* __restrict__ offsets, long long* __restrict__ header,
                                                          int cap_contours, long long cap_points, double sx, double sy,
                                                          int2* __restrict__ xy) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = header[0] < cap_contours ? header[0] : cap_contours;
    if (c >= n) return;
    if (header[1] > cap_points) {  // caller's buffer too small: write nothing, flag it
        if (c == 0) header[2] |= 2;
        return;
    }
    trace_contour(nb + (size_t)start_slice[c] * n_per_slice, W, starts[c], 8 * n_per_slice + 8,
                  WriteEmit{xy + offsets[c], sx, sy, 0});
}

}  // namespace

void m2p_phase_a(M2pWs& ws, PolyDev& P, const uint8_t* d_mask, int h, int w, int batch, int threshold, cudaStream_t st) {
    MS_REQUIRE(h > 0 && w > 0 && batch > 0 && h <= 65535 && batch <= 32767, MS_ERR_ARG, "mask2polygon: bad shape");
    MS_REQUIRE((int64_t)h * w <= (int64_t)(0x7FFFFFFF - 8) / 8, MS_ERR_ARG, "mask2polygon: slice too large");
    const int n = h * w;
    const size_t nb = (size_t)n * batch;
    if (P.cap_contours == 0) P.cap_contours = std::max<int64_t>(1024, 64 * (int64_t)batch);
    if (P.cap_points == 0) P.cap_points = std::max<int64_t>(65536, 4096 * (int64_t)batch);
    ws.fg.labels.reserve(nb * 4);
    ws.bg.labels.reserve(nb * 4);
    ws.bg.flag.reserve(nb);
    ws.nb.reserve(nb);
    const int wpitch = cdiv(w, 32);
    ws.fgbits.reserve((size_t)batch * h * wpitch * 4);
    const int bps = cdiv(n, 256);
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
    uint8_t* nbc = ws.nb.as<uint8_t>();
    int* bc = P.block_counts.as<int>();
    int* slice_total = bc + (size_t)bps * batch;
    long long* header = P.header.as<long long>();

    m2p_init_kernel<<<dim3(cdiv(w, TW), cdiv(h, TH), batch), 256, 0, st>>>(d_mask, h, w, threshold, Lfg, Lbg, flag, nbc,
                                                                          ws.fgbits.as<uint32_t>(), wpitch);
    MS_LAUNCH_CHECK();
    const dim3 g = ccl::grid_for(h, w, batch);
    ccl::merge_kernel<8><<<g, ccl::kThreads, 0, st>>>(Lfg, h, w);
    MS_LAUNCH_CHECK();
    ccl::merge_kernel<4><<<g, ccl::kThreads, 0, st>>>(Lbg, h, w);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<<<g, ccl::kThreads, 0, st>>>(Lfg, h, w, nullptr, nullptr);
    MS_LAUNCH_CHECK();
    ccl::resolve_kernel<<<g, ccl::kThreads, 0, st>>>(Lbg, h, w, nullptr, flag);
    MS_LAUNCH_CHECK();
    count_starts_kernel<<<dim3(bps, batch), 256, 0, st>>>(Lfg, Lbg, flag, w, n, bc);
    MS_LAUNCH_CHECK();
    scan_blocks_kernel<<<batch, 1024, 0, st>>>(bc, bps, slice_total);
    MS_LAUNCH_CHECK();
    scan_slices_kernel<<<1, 1024, 0, st>>>(slice_total, batch, P.slice_start.as<int>(), header);
    MS_LAUNCH_CHECK();
    write_starts_kernel<<<dim3(bps, batch), 256, 0, st>>>(Lfg, Lbg, flag, w, n, bc, P.slice_start.as<int>(), (int)P.cap_contours,
                                                          P.starts.as<int>(), P.start_slice.as<int>());
    MS_LAUNCH_CHECK();
    const size_t trace_smem = (size_t)(h + 2) * (wpitch + 2) * 4;
    if (trace_smem <= kTraceSmemMax) {
        static bool attr = false;
        if (!attr) {
            MS_CUDA(cudaFuncSetAttribute(trace_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTraceSmemMax));
            MS_CUDA(cudaFuncSetAttribute(trace_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTraceSmemMax));
            attr = true;
        }
        trace_smem_kernel<false><<<batch, 128, trace_smem, st>>>(ws.fgbits.as<uint32_t>(), h, w, wpitch, P.starts.as<int>(),
                                                                P.slice_start.as<int>(), header, (int)P.cap_contours, P.npts.as<int>(),
                                                                0, 1.0, 1.0, nullptr);
    } else {
        trace_count_kernel<<<cdiv((int)P.cap_contours, 128), 128, 0, st>>>(nbc, w, n, P.starts.as<int>(), P.start_slice.as<int>(), header,
                                                                         (int)P.cap_contours, P.npts.as<int>());
    }
    MS_LAUNCH_CHECK();
    scan_points_kernel<<<1, 1024, 0, st>>>(P.npts.as<int>(), (int)P.cap_contours, header);
    MS_LAUNCH_CHECK();
}

void m2p_phase_b(M2pWs& ws, PolyDev& P, int h, int w, int batch, int orig_w, int orig_h, cudaStream_t st) {
    const int n = h * w;
    // src/mask2polygon.cpp:199-200
    const double sx = static_cast<double>(orig_w) / w;
    const double sy = static_cast<double>(orig_h) / h;
    const int wpitch = cdiv(w, 32);
    const size_t trace_smem = (size_t)(h + 2) * (wpitch + 2) * 4;
    if (trace_smem <= kTraceSmemMax) {
        trace_smem_kernel<true><<<batch, 128, trace_smem, st>>>(ws.fgbits.as<uint32_t>(), h, w, wpitch, P.starts.as<int>(),
                                                               P.slice_start.as<int>(), P.header.as<long long>(), (int)P.cap_contours,
                                                               P.npts.as<int>(), (long long)P.cap_points, sx, sy, P.xy.as<int2>());
    } else {
        trace_emit_kernel<<<cdiv((int)P.cap_contours, 128), 128, 0, st>>>(ws.nb.as<uint8_t>(), w, n, P.starts.as<int>(),
                                                                        P.start_slice.as<int>(), P.npts.as<int>(),
                                                                        P.header.as<long long>(), (int)P.cap_contours,
                                                                        (long long)P.cap_points, sx, sy, P.xy.as<int2>());
    }
    MS_LAUNCH_CHECK();
}

}  // namespace ms
