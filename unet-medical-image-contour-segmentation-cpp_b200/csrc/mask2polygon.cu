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
//   (4)-(6) contour ordering + CHAIN_APPROX_SIMPLE, four interchangeable bit-exact variants (MEDSEG_TRACE):
//       rank   : (default when the slice fits in shared memory) one CTA per slice ranks the border cracks on chip
//       smem   : one thread per contour follows the border (contour_trace.cuh) on the slice's bit image in shared
//                memory, ONE walk: kept vertices -> 64-vertex chunks, counts -> scan -> parallel gather
//       crack  : slices too large for shared memory: list ranking on directed pixel edges, nothing walks (below)
//       window : fallback, the walk runs inside a re-centred 256 x 256 shared-memory window
//   (7) coordinate mapping (int)(x * (double)orig_w / w)             -> fused into the gather / emit pass
// No host round trip happens between these launches: all sizes live in `header` on the device, so
// the whole stage is CUDA-graph capturable and only the caller decides when to synchronise.
// The u8 mask is read once (-> one bit per pixel); labelling is the run-based union-find of ccl.cuh.
// Algorithmic bytes: H*W (mask read) + 8 B per vertex + 4 B per contour offset.
#include "ccl.cuh"
#include "contour_trace.cuh"
#include <cstdlib>
#include <cstring>

namespace ms {

namespace {

constexpr size_t kTraceSmemMax = 200 * 1024;   // slices up to ~1264 x 1264 trace out of shared memory (+ the 8 KiB step table)

// mask -> bits of (mask > thr), one word per warp, and the run heads of both labellings (8-connected foreground,
// 4-connected background + frame flag).  grid = (ceil(W / 256), H, batch), block = 256 (8 words)
__global__ void __launch_bounds__(256) thr_bits_kernel(const uint8_t* __restrict__ mask, int H, int W, int wpitch, int thr,
                                                        uint32_t* __restrict__ bits, int* __restrict__ L_fg, int* __restrict__ L_bg,
                                                        uint8_t* __restrict__ bg_flag) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    const bool fg = x < W && mask[((size_t)blockIdx.z * H + y) * W + x] > thr;   // src/mask2polygon.cpp:31 threshold(127)
    const unsigned b = __ballot_sync(0xFFFFFFFFu, fg);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < wpitch) {
        bits[((size_t)blockIdx.z * H + y) * wpitch + (x >> 5)] = b;
        const size_t slice = (size_t)blockIdx.z * H * W;
        ccl::init_heads<false>(b, W, y, x >> 5, L_fg + slice, nullptr, nullptr);
        ccl::init_heads<true>(b, W, y, x >> 5, L_bg + slice, nullptr, bg_flag + slice);
    }
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
        header[4] = 0;      // vertex chunks handed out by the border followers
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

// The border followers walk every contour ONCE.  Kept vertices go, unmapped, into 64-vertex chunks handed out by an atomic
// counter (a chunk is owned by one contour and filled completely before the next one is taken); after the per-contour
// counts are scanned, gather_kernel copies chunk (contour c, sequence s) to offset[c] + 64 s with the coordinate
// mapping applied -- a parallel copy instead of a second walk, and re-runnable with another mapping (the overlay wants
// network-space coordinates, the JSON original-space ones).
constexpr int kChunk = 64;
struct ChunkEmit {
    int2* tmp;                       // [cap_chunks][kChunk]
    int2* meta;                      // [cap_chunks] {contour, sequence}
    unsigned long long* counter;     // header[4]
    long long cap_chunks;
    int contour;
    int k;
    int2* cur;
    __device__ void operator()(int x, int y) {
        const int i = k & (kChunk - 1);
        if (i == 0) {
            const unsigned long long j = atomicAdd(counter, 1ull);
            cur = nullptr;
            if ((long long)j < cap_chunks) {   // else: capacities too small, the caller grows them and re-runs (counts stay exact)
                meta[j] = make_int2(contour, k / kChunk);
                cur = tmp + j * kChunk;
            }
        }
        if (cur) cur[i] = make_int2(x, y);
        ++k;
    }
};
__device__ __forceinline__ int2 map_point(int2 p, double sx, double sy) {
    // src/mask2polygon.cpp:54-55: static_cast<int>(pt.x * scale_x)
    return make_int2((int)__dmul_rn((double)p.x, sx), (int)__dmul_rn((double)p.y, sy));
}
struct WriteEmit {   // direct mapped write at a known offset (crack path)
    int2* dst;
    double sx, sy;
    mutable int i;
    __device__ void operator()(int x, int y) const { dst[i++] = map_point(make_int2(x, y), sx, sy); }
};
inline long long chunk_capacity(const PolyDev& P) { return P.cap_points / kChunk + P.cap_contours + 1; }

// one thread per chunk slot: copy + map.  grid covers cap_chunks * kChunk threads
__global__ void __launch_bounds__(256) gather_kernel(const int2* __restrict__ tmp, const int2* __restrict__ meta, long long* __restrict__ header,
                                                      long long cap_chunks, const int* __restrict__ offsets, int cap_contours,
                                                      long long cap_points, double sx, double sy, int2* __restrict__ xy) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long j = idx / kChunk;
    const long long n_chunks = header[4] < cap_chunks ? header[4] : cap_chunks;
    if (j >= n_chunks) return;
    if (header[1] > cap_points || header[0] > cap_contours) {   // caller's buffers too small: write nothing, flag it
        if (idx == 0) header[2] |= 2;
        return;
    }
    const int2 m = meta[j];
    const int k = m.y * kChunk + (int)(idx % kChunk);
    const int base = offsets[m.x];
    if (k < offsets[m.x + 1] - base) xy[base + k] = map_point(tmp[idx], sx, sy);
}

// 3x3 window of a pixel as 9 raw bits: up3 | cu3 << 3 | dn3 << 6, bit 0 of each triple = x - 1
struct PaddedBitsWindow {
    const uint32_t* bits;   // (H + 2) rows x pitch words, row 0 / word 0 are the zero frame
    int pitch;
    __device__ __forceinline__ unsigned operator()(int x, int y) const {
        const int X = x + 31, w = X >> 5, sh = X & 31;              // pixel x lives at bit x + 32 of the padded row
        const uint32_t* r = bits + y * pitch + w;                    // padded row y <-> image row y - 1
        return (__funnelshift_r(r[0], r[1], sh) & 7u) | ((__funnelshift_r(r[pitch], r[pitch + 1], sh) & 7u) << 3) |
               ((__funnelshift_r(r[2 * pitch], r[2 * pitch + 1], sh) & 7u) << 6);
    }
};
__device__ void build_trace_lut(uint16_t* lut) {   // contour_trace.cuh: the step table, one copy per CTA
    for (int i = threadIdx.x; i < kTraceLutEntries; i += blockDim.x) lut[i] = trace_lut_entry(i);
}
// Large slices (the bit image does not fit in shared memory): one CTA per contour keeps a 256 x 256-pixel WINDOW of the
// bit image in shared memory, centred on the walk; thread 0 follows the border (six LDS per step) until it leaves the
// window's interior, then the CTA re-centres the window and the walk resumes (contour_trace.cuh: trace_run).  A long
// contour (the stress masks have ~10^5-step borders) therefore walks at shared-memory latency, not at one dependent
// L2 access per step, and all contours of all slices walk concurrently.
constexpr int kWinW = 256, kWinH = 256, kWinPitch = kWinW / 32 + 1;   // +1 word: funnel shifts read one word ahead
struct WindowBits {
    const uint32_t* win;   // kWinH rows x kWinPitch words
    int x0, y0;            // image coordinates of the window's first pixel (x0 a multiple of 32, may be negative)
    __device__ __forceinline__ unsigned operator()(int x, int y) const {
        const int cx = x - x0 - 1, w = cx >> 5, sh = cx & 31;          // bits cx .. cx+2 = x-1 .. x+1
        const uint32_t* r = win + (y - y0 - 1) * kWinPitch + w;
        return (__funnelshift_r(r[0], r[1], sh) & 7u) | ((__funnelshift_r(r[kWinPitch], r[kWinPitch + 1], sh) & 7u) << 3) |
               ((__funnelshift_r(r[2 * kWinPitch], r[2 * kWinPitch + 1], sh) & 7u) << 6);
    }
};
struct WindowInside {
    int x0, y0;
    __device__ __forceinline__ bool operator()(int x, int y) const {
        return x - 1 >= x0 && x + 1 < x0 + kWinW && y - 1 >= y0 && y + 1 < y0 + kWinH;
    }
};

__global__ void __launch_bounds__(64) trace_window_kernel(const uint32_t* __restrict__ fgbits, int H, int W, int wpitch,
                                                           const int* __restrict__ starts, const int* __restrict__ start_slice,
                                                           long long* __restrict__ header, int cap_contours, int* __restrict__ npts,
                                                           int2* __restrict__ chunk_tmp, int2* __restrict__ chunk_meta, long long cap_chunks) {
    __shared__ uint32_t win[kWinH * kWinPitch];
    __shared__ uint16_t lut[kTraceLutEntries];
    build_trace_lut(lut);
    __shared__ int s_org[2], s_status;
    const int c = blockIdx.x;
    const long long n = header[0] < cap_contours ? header[0] : cap_contours;
    if (c >= n) return;
    const uint32_t* B = fgbits + (size_t)start_slice[c] * H * wpitch;
    TraceState st;
    trace_begin(st, W, starts[c]);
    ChunkEmit emit{chunk_tmp, chunk_meta, (unsigned long long*)&header[4], cap_chunks, c, 0, nullptr};
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
            const WindowBits bits{win, x0, y0};
            const WindowInside inside{x0, y0};
            s_status = trace_run(bits, lut, W, st, 8 * H * W + 8, emit, inside);
        }
        __syncthreads();
        if (s_status != 0) break;
    }
    if (threadIdx.x == 0) {
        if (s_status != 1) atomicAdd((unsigned long long*)&header[3], 1ull);
        npts[c] = s_status == 1 ? st.n : 0;
    }
}

// Shared-memory variant: one CTA per slice keeps the slice's bit-packed foreground (H*W/8 bytes, padded by a zero word /
// zero row on every side) in shared memory, so a border-following step costs six LDS instead of dependent global
// loads.  Thread t traces contours slice_start[b] + t, + blockDim, ...
__global__ void __launch_bounds__(128) trace_smem_kernel(const uint32_t* __restrict__ fgbits, int H, int W, int wpitch,
                                                          const int* __restrict__ starts, const int* __restrict__ slice_start,
                                                          long long* __restrict__ header, int cap_contours, int* __restrict__ npts,
                                                          int2* __restrict__ chunk_tmp, int2* __restrict__ chunk_meta, long long cap_chunks) {
    extern __shared__ uint32_t sbits[];
    __shared__ uint16_t lut[kTraceLutEntries];
    const int b = blockIdx.x;
    const int pitch = wpitch + 2;
    const int c_lo = slice_start[b], c_hi = min(slice_start[b + 1], cap_contours);
    if (c_lo >= c_hi) return;
    for (int i = threadIdx.x; i < (H + 2) * pitch; i += blockDim.x) {
        const int r = i / pitch, c = i % pitch;
        sbits[i] = (r >= 1 && r <= H && c >= 1 && c <= wpitch) ? fgbits[((size_t)b * H + (r - 1)) * wpitch + (c - 1)] : 0u;
    }
    build_trace_lut(lut);
    __syncthreads();
    const PaddedBitsWindow bits{sbits, pitch};
    for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
        TraceState ts;
        trace_begin(ts, W, starts[c]);
        ChunkEmit emit{chunk_tmp, chunk_meta, (unsigned long long*)&header[4], cap_chunks, c, 0, nullptr};
        const int cnt = trace_run(bits, lut, W, ts, 8 * H * W + 8, emit, TraceAlwaysInside{}) == 1 ? ts.n : -1;
        if (cnt < 0) atomicAdd((unsigned long long*)&header[3], 1ull);
        npts[c] = cnt < 0 ? 0 : cnt;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Parallel contour ordering (slices whose bit image does not fit in shared memory).
//
// A *crack* is a directed pixel edge with foreground on its left and background on its right: crack (p, s) is side s of
// foreground pixel p (s: 0 = W side walked southwards, 1 = S eastwards, 2 = E northwards, 3 = N westwards).  At the end
// of a crack the walk turns towards the diagonal pixel if that is foreground (8-connectivity), else runs straight on
// along the next pixel, else turns around p itself.  succ() is a permutation of the cracks; its cycle through the W
// side of a component's raster-first pixel is that component's outer border, and the owners of the cracks along the
// cycle -- consecutive repeats removed -- are exactly the pixels cv::findContours visits, in its order (checked against
// cv2 by the pure-Python model in tests/test_crack_model.py and, for these kernels, by the GPU stress tests).  Border following, inherently sequential, becomes
// list ranking:
//   init   every crack: {succ, 1}
//   cut    the predecessor of each external start crack ends its list and carries the contour number
//   jump   log2(longest border) rounds of pointer jumping, in place: {next, rank} is one 64-bit word, and every value a
//          racing reader can observe is a true statement "next is rank cracks ahead", so no double buffering
//   flags  every crack now knows its contour and position.  The first crack of each pixel visit applies the
//          CHAIN_APPROX_SIMPLE rule (direction to the next visited pixel != direction from the previous one) from the
//          3x3 neighbourhood alone and drops {pixel, kept} at its position
//   scan   kept flags -> output offsets;  emit maps the coordinates.
// Every kernel is word-parallel over the bit image or position-parallel; nothing walks a contour.
namespace crack {

typedef unsigned long long u64;
constexpr uint32_t kKept = 0x80000000u;
__device__ __forceinline__ u64 pack(int next, uint32_t rank) { return ((u64)rank << 32) | (uint32_t)next; }
__device__ __forceinline__ int next_of(u64 v) { return (int)(uint32_t)v; }
__device__ __forceinline__ uint32_t rank_of(u64 v) { return (uint32_t)(v >> 32); }

// 34-bit row windows around word (y, wx): bit k <-> x = wx * 32 + k - 1; outside the image = background
struct Nbhd {
    u64 up, cu, dn;
    uint32_t border;   // foreground pixels of the word with a background 4-neighbour (the pixels that own cracks)
};
__device__ __forceinline__ Nbhd load_nbhd(const uint32_t* __restrict__ B, int H, int wpitch, int y, int wx) {
    auto window = [&](int yy) -> u64 {
        if (yy < 0 || yy >= H) return 0ull;
        const uint32_t* r = B + (size_t)yy * wpitch + wx;
        const uint32_t c = __ldg(r), l = wx > 0 ? __ldg(r - 1) : 0u, rr = wx + 1 < wpitch ? __ldg(r + 1) : 0u;
        return (u64)(l >> 31) | ((u64)c << 1) | ((u64)(rr & 1u) << 33);
    };
    Nbhd n;
    n.cu = window(y);
    const uint32_t fg = (uint32_t)(n.cu >> 1);
    n.up = n.dn = 0;
    n.border = 0;
    if (fg == 0) return n;
    n.up = window(y - 1);
    n.dn = window(y + 1);
    n.border = fg & ~((uint32_t)(n.up >> 1) & (uint32_t)(n.dn >> 1) & (uint32_t)n.cu & (uint32_t)(n.cu >> 2));
    return n;
}
__device__ __forceinline__ unsigned code_at(const Nbhd& n, int j) {
    return code_from_rows((unsigned)(n.up >> j) & 7u, (unsigned)(n.cu >> j) & 7u, (unsigned)(n.dn >> j) & 7u);
}
__device__ __forceinline__ bool has(unsigned code, int d) { return (code >> (d & 7)) & 1u; }
__device__ __forceinline__ int step(int p, int d, int W) { return p + trace_dy(d & 7) * W + trace_dx(d & 7); }
// side s exists as a crack iff the 4-neighbour across it (W, S, E, N = directions 4, 6, 0, 2) is background
__device__ __forceinline__ bool side_open(unsigned code, int s) { return !has(code, 4 + 2 * s); }
// successor: diagonal pixel (directions SW, SE, NE, NW for s = 0..3) -> its side s+3; straight pixel (S, E, N, W) -> its side
// s; else the next side of p
__device__ __forceinline__ int succ(unsigned code, int p, int s, int W) {
    if (has(code, 5 + 2 * s)) return step(p, 5 + 2 * s, W) * 4 + ((s + 3) & 3);
    if (has(code, 6 + 2 * s)) return step(p, 6 + 2 * s, W) * 4 + s;
    return p * 4 + ((s + 1) & 3);
}
// predecessor (the inverse): diagonal pixel (NW, SW, SE, NE) -> its side s+1; straight pixel (N, W, S, E) -> its side s;
// else the previous side of p
__device__ __forceinline__ int pred_dir(unsigned code, int s) {   // direction of the predecessor's owner, -1 = p itself
    if (has(code, 3 + 2 * s)) return (3 + 2 * s) & 7;
    if (has(code, 2 + 2 * s)) return (2 + 2 * s) & 7;
    return -1;
}
__device__ __forceinline__ int pred(unsigned code, int p, int s, int W) {
    if (has(code, 3 + 2 * s)) return step(p, 3 + 2 * s, W) * 4 + ((s + 1) & 3);
    if (has(code, 2 + 2 * s)) return step(p, 2 + 2 * s, W) * 4 + s;
    return p * 4 + ((s + 3) & 3);
}
template <class F>
__device__ __forceinline__ void for_each_crack(const Nbhd& n, int row0, F&& f) {   // row0 = pixel index of the word's bit 0
    uint32_t m = n.border;
    while (m) {
        const int j = __ffs((int)m) - 1;
        m &= m - 1;
        const unsigned code = code_at(n, j);
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (side_open(code, s)) f(row0 + j, s, code);
    }
}

__global__ void __launch_bounds__(ccl::kThreads) init_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                              u64* __restrict__ pair_all) {
    MS_CCL_WORD_COORDS();
    const Nbhd n = load_nbhd(bits_all + (size_t)sl * H * wpitch, H, wpitch, y, wx);
    u64* pair = pair_all + (size_t)sl * H * W * 4;
    for_each_crack(n, y * W + wx * 32, [&](int p, int s, unsigned code) { __stcg(&pair[p * 4 + s], pack(succ(code, p, s, W), 1u)); });
}

// one thread per contour: end the list at the predecessor of the start crack; rotation of the start pixel's visit
__global__ void __launch_bounds__(256) cut_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                   const int* __restrict__ starts, const int* __restrict__ start_slice,
                                                   const long long* __restrict__ header, int cap_contours, u64* __restrict__ pair_all,
                                                   int* __restrict__ rot, int* __restrict__ npts) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const long long n = header[0] < cap_contours ? header[0] : cap_contours;
    if (c >= n) return;
    const int sl = start_slice[c], p = starts[c], x = p % W, y = p / W;
    const Nbhd nb = load_nbhd(bits_all + (size_t)sl * H * wpitch, H, wpitch, y, x >> 5);
    const unsigned code = code_at(nb, x & 31);
    // the start pixel's visit may begin up to three cracks before its W side (N, E, S sides walked just before)
    int back = 0, s = 0;
    while (back < 3 && pred_dir(code, s) < 0) {
        s = (s + 3) & 3;
        ++back;
    }
    rot[c] = back;
    npts[c] = 0;
    __stcg(&pair_all[(size_t)sl * H * W * 4 + pred(code, p, 0, W)], pack(-(c + 2), 1u));
}

__global__ void __launch_bounds__(ccl::kThreads) jump_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                              u64* __restrict__ pair_all, const int* __restrict__ busy, int round) {
    if (round > 0 && busy[round - 1] == 0) return;   // every external border already ranked
    MS_CCL_WORD_COORDS();
    const Nbhd n = load_nbhd(bits_all + (size_t)sl * H * wpitch, H, wpitch, y, wx);
    u64* pair = pair_all + (size_t)sl * H * W * 4;
    for_each_crack(n, y * W + wx * 32, [&](int p, int s, unsigned) {
        u64* me = &pair[p * 4 + s];
        const u64 v = __ldcg(me);
        const int nx = next_of(v);
        if (nx < 0) return;
        const u64 t = __ldcg(&pair[nx]);
        __stcg(me, pack(next_of(t), rank_of(v) + rank_of(t)));
    });
}

// busy[round] = some external border is not completely ranked after rounds 0 .. round.  After k rounds every crack is
// either ranked or points at least 2^k cracks ahead (the in-place update can only jump further than the synchronous one),
// so a border of length L is completely ranked once its start crack is ranked and 2^k >= L; the start crack alone being
// ranked is not enough, it may have got there through already-updated successors while a crack behind it has not.
__global__ void __launch_bounds__(256) check_kernel(int H, int W, const int* __restrict__ starts, const int* __restrict__ start_slice,
                                                     const long long* __restrict__ header, int cap_contours,
                                                     const u64* __restrict__ pair_all, int* __restrict__ busy, int round) {
    if (round > 0 && busy[round - 1] == 0) return;
    const int c = blockIdx.x * 256 + threadIdx.x;
    const long long n = header[0] < cap_contours ? header[0] : cap_contours;
    if (c >= n) return;
    const u64 t = __ldcg(&pair_all[(size_t)start_slice[c] * H * W * 4 + (size_t)starts[c] * 4]);
    if (next_of(t) >= 0 || (u64)rank_of(t) > (1ull << (round + 1))) busy[round] = 1;
}

// 1 block of 1024: border length of every contour -> exclusive position bases; meta[0] = total positions
__global__ void __launch_bounds__(1024) length_kernel(int H, int W, const int* __restrict__ starts, const int* __restrict__ start_slice,
                                                       long long* __restrict__ header, int cap_contours, const u64* __restrict__ pair_all,
                                                       int* __restrict__ cbase, long long* __restrict__ meta) {
    const long long n64 = header[0];
    const int n = (int)(n64 < cap_contours ? n64 : cap_contours);
    long long carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        int v = 0;
        if (i < n) {
            const u64 t = __ldcg(&pair_all[(size_t)start_slice[i] * H * W * 4 + (size_t)starts[i] * 4]);
            if (next_of(t) == -(i + 2)) v = (int)rank_of(t);
            else atomicAdd((unsigned long long*)&header[3], 1ull);   // unranked border: cannot happen
        }
        int tot;
        const int ex = block_exscan_1024(v, &tot);
        if (i < n) cbase[i] = (int)(carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0) {
        cbase[n] = (int)carry;
        meta[0] = carry;
    }
}

__global__ void __launch_bounds__(ccl::kThreads) flags_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                               const u64* __restrict__ pair_all, const int* __restrict__ cbase,
                                                               const int* __restrict__ rot, uint32_t* __restrict__ pos_px,
                                                               int* __restrict__ npts) {
    MS_CCL_WORD_COORDS();
    const Nbhd n = load_nbhd(bits_all + (size_t)sl * H * wpitch, H, wpitch, y, wx);
    const u64* pair = pair_all + (size_t)sl * H * W * 4;
    for_each_crack(n, y * W + wx * 32, [&](int p, int s, unsigned code) {
        const u64 v = __ldcg(&pair[p * 4 + s]);
        const int nx = next_of(v);
        if (nx >= 0) return;                               // hole border or border of a non-external component
        const int c = -nx - 2;
        const int base = cbase[c], len = cbase[c + 1] - base;
        int pos = len - (int)rank_of(v) + rot[c];          // position along the border, the start pixel's visit first
        if (pos >= len) pos -= len;
        bool kept = false;
        const int pd = pred_dir(code, s);
        if (pd >= 0) {                                     // first crack of a visit of p
            const int d_prev = (pd + 4) & 7;               // direction previous pixel -> p
            int d_out = -1;                                // direction p -> next visited pixel
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int ss = (s + t) & 3;
                if (d_out < 0 && has(code, 5 + 2 * ss)) d_out = (5 + 2 * ss) & 7;
                if (d_out < 0 && has(code, 6 + 2 * ss)) d_out = (6 + 2 * ss) & 7;
            }
            kept = d_out != d_prev;
        } else if (code == 0 && s == 0) {
            kept = true;                                   // isolated pixel: a single vertex
        }
        pos_px[(size_t)base + pos] = (uint32_t)p | (kept ? kKept : 0u);
        if (kept) atomicAdd(&npts[c], 1);
    });
}

// position-parallel: kept vertices per 1024 positions
__global__ void __launch_bounds__(1024) pos_count_kernel(const uint32_t* __restrict__ pos_px, const long long* __restrict__ meta,
                                                          int* __restrict__ blocks) {
    const long long T = meta[0], i = (long long)blockIdx.x * 1024 + threadIdx.x;
    if ((long long)blockIdx.x * 1024 >= T) return;
    const int cnt = __syncthreads_count(i < T && (pos_px[i] & kKept));
    if (threadIdx.x == 0) blocks[blockIdx.x] = cnt;
}
__global__ void __launch_bounds__(1024) pos_scan_kernel(int* __restrict__ blocks, const long long* __restrict__ meta) {
    const int nb = (int)((meta[0] + 1023) / 1024);
    int carry = 0;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? blocks[i] : 0;
        int tot;
        const int ex = block_exscan_1024(v, &tot);
        if (i < nb) blocks[i] = carry + ex;
        carry += tot;
    }
}
__global__ void __launch_bounds__(1024) pos_emit_kernel(const uint32_t* __restrict__ pos_px, const long long* __restrict__ meta,
                                                         const int* __restrict__ blocks, long long* __restrict__ header,
                                                         long long cap_points, int W, double sx, double sy, int2* __restrict__ xy) {
    const long long T = meta[0], i = (long long)blockIdx.x * 1024 + threadIdx.x;
    if ((long long)blockIdx.x * 1024 >= T) return;
    if (header[1] > cap_points) {
        if (blockIdx.x == 0 && threadIdx.x == 0) header[2] |= 2;
        return;
    }
    const uint32_t v = i < T ? pos_px[i] : 0u;
    const bool kept = v & kKept;
    int tot;
    const int ex = block_exscan_1024(kept ? 1 : 0, &tot);
    if (kept) {
        const int p = (int)(v & ~kKept), x = p % W, y = p / W;
        WriteEmit{xy + blocks[blockIdx.x] + ex, sx, sy, 0}(x, y);
    }
}

}  // namespace crack

// ---------------------------------------------------------------------------------------------------------------------
// The same list ranking for slices whose bit image fits in shared memory (the pipeline's case): ONE CTA per slice does
// everything on chip -- bits, crack table, pointer jumping, CHAIN_APPROX_SIMPLE flags, prefix sums -- and hands the kept
// vertices to the chunk store that gather_kernel reads.  A CT-like slice has one to three borders of ~1,500 pixels; walking
// them costs ~170 ns per step with one thread busy per border (0.24 ms per batch), ranking them takes ~12 rounds over a few
// thousand table entries with 1,024 threads.
// Crack slots are indexed by border pixel: slot = (rank of the pixel among the slice's border pixels) * 4 + side, the
// pixel's rank = pix_off[word] + popc(border bits below it).  A slot is 32 bits, {next : 16, rank : 16}; next >= kMark
// is a marker: kInvalidNext (side not open) or kMark + contour (list end).  A slice with more border pixels than the
// table holds (or > 16 k contours) is walked sequentially by this same kernel, exactly as trace_smem_kernel does.
namespace srank {

constexpr uint32_t kMark = 0xC000u, kInvalidNext = 0xFFFFu;
constexpr int kMaxBorder = kMark / 4;            // slot ids stay below kMark
constexpr int kBytesPerBorder = 16 + 16 + 4;     // slots (later: prefix sums), positions, pixel index
constexpr int kInfo = 6;                         // per-contour scratch arrays in global memory
__device__ __forceinline__ uint32_t pack(uint32_t next, uint32_t rank) { return (rank << 16) | (next & 0xFFFFu); }
__device__ __forceinline__ uint32_t next_of(uint32_t v) { return v & 0xFFFFu; }
__device__ __forceinline__ uint32_t rank_of(uint32_t v) { return v >> 16; }

// 34-bit row windows of word (y, wx) out of the zero-framed shared-memory copy (row y+1, word wx+1)
__device__ __forceinline__ crack::Nbhd nbhd(const uint32_t* sb, int P, int y, int wx) {
    auto window = [&](int ry) -> crack::u64 {
        const uint32_t* r = sb + ry * P + wx;
        return (crack::u64)(r[0] >> 31) | ((crack::u64)r[1] << 1) | ((crack::u64)(r[2] & 1u) << 33);
    };
    crack::Nbhd n;
    n.cu = window(y + 1);
    n.up = n.dn = 0;
    n.border = 0;
    const uint32_t fg = (uint32_t)(n.cu >> 1);
    if (fg == 0) return n;
    n.up = window(y);
    n.dn = window(y + 2);
    n.border = fg & ~((uint32_t)(n.up >> 1) & (uint32_t)(n.dn >> 1) & (uint32_t)n.cu & (uint32_t)(n.cu >> 2));
    return n;
}
struct Table {
    const uint32_t* sb;
    const uint16_t* pix_off;
    int P, wpitch;
    // slot of side s of border pixel (x, y)
    __device__ __forceinline__ uint32_t slot(int x, int y, int s) const {
        const int wx = x >> 5;
        const crack::Nbhd n = nbhd(sb, P, y, wx);
        return ((uint32_t)pix_off[y * wpitch + wx] + (uint32_t)__popc(n.border & ((1u << (x & 31)) - 1u))) * 4u + (uint32_t)s;
    }
    __device__ __forceinline__ unsigned code(int x, int y) const { return crack::code_at(nbhd(sb, P, y, x >> 5), x & 31); }
};

__global__ void __launch_bounds__(1024) rank_smem_kernel(const uint32_t* __restrict__ fgbits, int H, int W, int wpitch,
                                                          const int* __restrict__ starts, const int* __restrict__ slice_start,
                                                          long long* __restrict__ header, int cap_contours, int* __restrict__ npts,
                                                          int2* __restrict__ chunk_tmp, int2* __restrict__ chunk_meta, long long cap_chunks,
                                                          int* __restrict__ cinfo, int cap_border) {
    extern __shared__ uint32_t smem[];
    __shared__ uint16_t lut[kTraceLutEntries];   // only the sequential fallback uses it
    const int tid = threadIdx.x, b = blockIdx.x;
    const int P = wpitch + 2, nwords = H * wpitch;
    const int c_lo = slice_start[b], c_hi = min(slice_start[b + 1], cap_contours);
    if (c_lo >= c_hi) return;
    uint32_t* sbits = smem;
    uint16_t* pix_off = reinterpret_cast<uint16_t*>(sbits + (H + 2) * P);
    uint32_t* slots = reinterpret_cast<uint32_t*>(pix_off + ((nwords + 1) & ~1));   // [4 * cap_border + 1]
    uint32_t* pos = slots + 4 * cap_border + 1;                                     // [4 * cap_border]
    uint32_t* slot_px = pos + 4 * cap_border;                                       // [cap_border]
    // per-contour scratch (global, indexed by contour): position base, rotation, start slot, border length, first chunk, vertex base
    int* c_base = cinfo;
    int* c_rot = cinfo + (size_t)(cap_contours + 1);
    int* c_slot = cinfo + 2 * (size_t)(cap_contours + 1);
    int* c_len = cinfo + 3 * (size_t)(cap_contours + 1);
    int* c_chunk = cinfo + 4 * (size_t)(cap_contours + 1);
    int* c_vbase = cinfo + 5 * (size_t)(cap_contours + 1);

    for (int i = tid; i < (H + 2) * P; i += 1024) {
        const int r = i / P, c = i % P;
        sbits[i] = (r >= 1 && r <= H && c >= 1 && c <= wpitch) ? fgbits[((size_t)b * H + (r - 1)) * wpitch + (c - 1)] : 0u;
    }
    __syncthreads();

    // ---- border pixels per word -> exclusive offsets
    int n_border = 0;
    for (int base = 0; base < nwords; base += 1024) {
        const int w = base + tid;
        int cnt = 0;
        if (w < nwords) cnt = __popc(nbhd(sbits, P, w / wpitch, w % wpitch).border);
        int tot;
        const int ex = block_exscan_1024(cnt, &tot);
        if (w < nwords) pix_off[w] = (uint16_t)min(n_border + ex, 0xFFFF);
        n_border += tot;
    }
    if (n_border > cap_border || c_hi - c_lo >= (int)(kInvalidNext - kMark)) {
        // ---- does not fit the table: one thread per contour walks its border (contour_trace.cuh)
        build_trace_lut(lut);
        __syncthreads();
        const PaddedBitsWindow bits{sbits, P};
        for (int c = c_lo + tid; c < c_hi; c += 1024) {
            TraceState ts;
            trace_begin(ts, W, starts[c]);
            ChunkEmit emit{chunk_tmp, chunk_meta, (unsigned long long*)&header[4], cap_chunks, c, 0, nullptr};
            const int cnt = trace_run(bits, lut, W, ts, 8 * H * W + 8, emit, TraceAlwaysInside{}) == 1 ? ts.n : -1;
            if (cnt < 0) atomicAdd((unsigned long long*)&header[3], 1ull);
            npts[c] = cnt < 0 ? 0 : cnt;
        }
        return;
    }
    __syncthreads();
    const Table T{sbits, pix_off, P, wpitch};
    const int n_slots = 4 * n_border;

    // ---- init: every open side points at its successor
    for (int w = tid; w < nwords; w += 1024) {
        const int y = w / wpitch, wx = w - y * wpitch;
        const crack::Nbhd n = nbhd(sbits, P, y, wx);
        uint32_t m = n.border;
        uint32_t idx = pix_off[w];
        while (m) {
            const int j = __ffs((int)m) - 1;
            m &= m - 1;
            const unsigned code = crack::code_at(n, j);
            const int x = wx * 32 + j;
            slot_px[idx] = (uint32_t)(y * W + x);
#pragma unroll
            for (int sd = 0; sd < 4; ++sd) {
                uint32_t v = pack(kInvalidNext, 0);
                if (crack::side_open(code, sd)) {
                    uint32_t nx;   // crack::succ, resolved to a slot
                    if (crack::has(code, 5 + 2 * sd)) nx = T.slot(x + trace_dx((5 + 2 * sd) & 7), y + trace_dy((5 + 2 * sd) & 7), (sd + 3) & 3);
                    else if (crack::has(code, 6 + 2 * sd)) nx = T.slot(x + trace_dx((6 + 2 * sd) & 7), y + trace_dy((6 + 2 * sd) & 7), sd);
                    else nx = idx * 4u + (uint32_t)((sd + 1) & 3);
                    v = pack(nx, 1);
                }
                slots[idx * 4 + sd] = v;
            }
            ++idx;
        }
    }
    __syncthreads();

    // ---- cut: the predecessor of each contour's start crack ends the list and names the contour
    for (int c = c_lo + tid; c < c_hi; c += 1024) {
        const int p = starts[c], x = p % W, y = p / W;
        const unsigned code = T.code(x, y);
        int back = 0, sd = 0;
        while (back < 3 && crack::pred_dir(code, sd) < 0) {
            sd = (sd + 3) & 3;
            ++back;
        }
        c_rot[c] = back;
        const uint32_t s0 = T.slot(x, y, 0);
        c_slot[c] = (int)s0;
        uint32_t pr;       // crack::pred of (p, side 0)
        if (crack::has(code, 3)) pr = T.slot(x - 1, y - 1, 1);
        else if (crack::has(code, 2)) pr = T.slot(x, y - 1, 0);
        else pr = (s0 & ~3u) | 3u;
        slots[pr] = pack(kMark + (uint32_t)(c - c_lo), 1);
    }
    __syncthreads();

    // ---- pointer jumping (in place: every observable {next, rank} is a true statement); a border of length L is done once
    //      its start crack is ranked and 2^rounds >= L
    for (int round = 0; round < 20; ++round) {
        for (int i = tid; i < n_slots; i += 1024) {
            const uint32_t v = slots[i];
            const uint32_t nx = next_of(v);
            if (nx < kMark) {
                const uint32_t t = slots[nx];
                slots[i] = pack(next_of(t), rank_of(v) + rank_of(t));
            }
        }
        __syncthreads();
        int busy = 0;
        for (int c = c_lo + tid; c < c_hi; c += 1024) {
            const uint32_t t = slots[c_slot[c]];
            if (next_of(t) < kMark || rank_of(t) > (1u << (round + 1))) busy = 1;
        }
        if (!__syncthreads_or(busy)) break;
    }

    // ---- border lengths -> position bases
    int n_pos = 0;
    for (int base = 0; base < c_hi - c_lo; base += 1024) {
        const int c = c_lo + base + tid;
        int len = 0;
        if (c < c_hi) {
            const uint32_t t = slots[c_slot[c]];
            if (next_of(t) == kMark + (uint32_t)(c - c_lo)) len = (int)rank_of(t);
            else atomicAdd((unsigned long long*)&header[3], 1ull);   // unranked border: cannot happen
            c_len[c] = len;
        }
        int tot;
        const int ex = block_exscan_1024(len, &tot);
        if (c < c_hi) c_base[c] = n_pos + ex;
        n_pos += tot;
    }
    __syncthreads();

    // ---- flags: position of every crack of an external border; the first crack of a pixel visit decides CHAIN_APPROX_SIMPLE
    for (int i = tid; i < n_slots; i += 1024) {
        const uint32_t v = slots[i];
        const uint32_t nx = next_of(v);
        if (nx < kMark || nx == kInvalidNext) continue;       // hole border / border of a non-external component / closed side
        const int c = c_lo + (int)(nx - kMark);
        const int len = c_len[c];
        int ps = len - (int)rank_of(v) + c_rot[c];
        if (ps >= len) ps -= len;
        const int p = (int)slot_px[i >> 2], sd = i & 3;
        const unsigned code = T.code(p % W, p / W);
        bool kept = false;
        const int pd = crack::pred_dir(code, sd);
        if (pd >= 0) {
            const int d_prev = (pd + 4) & 7;
            int d_out = -1;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int ss = (sd + t) & 3;
                if (d_out < 0 && crack::has(code, 5 + 2 * ss)) d_out = (5 + 2 * ss) & 7;
                if (d_out < 0 && crack::has(code, 6 + 2 * ss)) d_out = (6 + 2 * ss) & 7;
            }
            kept = d_out != d_prev;
        } else if (code == 0 && sd == 0) {
            kept = true;
        }
        pos[c_base[c] + ps] = (uint32_t)p | (kept ? crack::kKept : 0u);
    }
    __syncthreads();

    // ---- kept-vertex prefix over positions (reuses the slot table's memory)
    uint32_t* prefix = slots;
    int n_kept = 0;
    for (int base = 0; base < n_pos; base += 1024) {
        const int i = base + tid;
        const int k = (i < n_pos && (pos[i] & crack::kKept)) ? 1 : 0;
        int tot;
        const int ex = block_exscan_1024(k, &tot);
        if (i < n_pos) prefix[i] = (uint32_t)(n_kept + ex);
        n_kept += tot;
    }
    if (tid == 0) prefix[n_pos] = (uint32_t)n_kept;
    __syncthreads();

    // ---- per contour: vertex count, chunks
    for (int c = c_lo + tid; c < c_hi; c += 1024) {
        const int k0 = (int)prefix[c_base[c]], n = (int)prefix[c_base[c] + c_len[c]] - k0;
        npts[c] = n;
        const int m = (n + kChunk - 1) / kChunk;
        const unsigned long long j0 = atomicAdd((unsigned long long*)&header[4], (unsigned long long)m);
        for (int q = 0; q < m; ++q)
            if ((long long)(j0 + q) < cap_chunks) chunk_meta[j0 + q] = make_int2(c, q);
        c_chunk[c] = (int)min(j0, (unsigned long long)0x7FFFFFFF);
        c_vbase[c] = k0;
    }
    __syncthreads();

    // ---- emit kept vertices into their chunks, in position order
    for (int i = tid; i < n_pos; i += 1024) {
        const uint32_t e = pos[i];
        if (!(e & crack::kKept)) continue;
        int lo = c_lo, hi = c_hi - 1;                          // last contour whose base is <= i
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (c_base[mid] <= i) lo = mid; else hi = mid - 1;
        }
        const int k = (int)prefix[i] - c_vbase[lo];
        const long long j = (long long)c_chunk[lo] + k / kChunk;
        const int p = (int)(e & ~crack::kKept);
        if (j < cap_chunks) chunk_tmp[j * kChunk + (k % kChunk)] = make_int2(p % W, p / W);
    }
}

}  // namespace srank

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

#include "slice_fused.cuh"
#include "dp_simplify.cuh"

enum TraceMode { kTraceSmem = 0, kTraceWindow = 1, kTraceCrack = 2, kTraceRank = 3 };

constexpr size_t kRankSmemMax = 218 * 1024;   // rank_smem_kernel: bits + offsets + tables (8 KiB static step table on top)
// border pixels the in-shared-memory crack table can hold for an h x w slice (0: does not fit at all)
int rank_border_capacity(int h, int w) {
    const int wpitch = cdiv(w, 32);
    const int64_t fixed = (int64_t)(h + 2) * (wpitch + 2) * 4 + (((int64_t)h * wpitch + 1) & ~1) * 2 + 16;
    const int64_t cap = ((int64_t)kRankSmemMax - fixed) / srank::kBytesPerBorder;
    return (int)std::max<int64_t>(0, std::min<int64_t>(cap, srank::kMaxBorder));
}

// MEDSEG_TRACE = rank | smem | window | crack forces a variant (tests exercise all of them on the same masks)
TraceMode pick_trace_mode(int h, int w, int batch) {
    const int wpitch = cdiv(w, 32);
    const size_t trace_smem = (size_t)(h + 2) * (wpitch + 2) * 4;
    const bool smem_ok = trace_smem <= kTraceSmemMax;
    const bool crack_ok = (int64_t)h * w * batch * 4 < ((int64_t)1 << 31);   // positions are int32
    if (const char* e = std::getenv("MEDSEG_TRACE")) {
        if (!std::strcmp(e, "window")) return kTraceWindow;
        if (!std::strcmp(e, "crack") && crack_ok) return kTraceCrack;
        if (!std::strcmp(e, "smem") && smem_ok) return kTraceSmem;
        if (!std::strcmp(e, "rank") && smem_ok && rank_border_capacity(h, w) > 0) return kTraceRank;
    }
    if (smem_ok) return rank_border_capacity(h, w) >= 1024 ? kTraceRank : kTraceSmem;
    return crack_ok ? kTraceCrack : kTraceWindow;
}

// the single walk: per-contour counts into P.npts, kept vertices into P.chunks
void launch_trace(M2pWs& ws, PolyDev& P, TraceMode mode, int h, int w, int batch, cudaStream_t st) {
    const int wpitch = cdiv(w, 32);
    const size_t trace_smem = (size_t)(h + 2) * (wpitch + 2) * 4;
    long long* header = P.header.as<long long>();
    const long long cap_chunks = chunk_capacity(P);
    if (mode == kTraceRank) {
        set_max_dynamic_smem(srank::rank_smem_kernel, (int)kRankSmemMax);
        const int cap_border = rank_border_capacity(h, w);
        const size_t smem = (size_t)(h + 2) * (wpitch + 2) * 4 + (((size_t)h * wpitch + 1) & ~(size_t)1) * 2 + 16 +
                            (size_t)cap_border * srank::kBytesPerBorder;
        ws.crack_contour.reserve(((size_t)P.cap_contours + 1) * srank::kInfo * sizeof(int));
        srank::rank_smem_kernel<<<batch, 1024, smem, st>>>(ws.fgbits.as<uint32_t>(), h, w, wpitch, P.starts.as<int>(), P.slice_start.as<int>(),
                                                          header, (int)P.cap_contours, P.npts.as<int>(), P.chunks.as<int2>(),
                                                          P.chunk_meta.as<int2>(), cap_chunks, ws.crack_contour.as<int>(), cap_border);
    } else if (mode == kTraceSmem) {
        set_max_dynamic_smem(trace_smem_kernel, (int)kTraceSmemMax);
        trace_smem_kernel<<<batch, 128, trace_smem, st>>>(ws.fgbits.as<uint32_t>(), h, w, wpitch, P.starts.as<int>(), P.slice_start.as<int>(),
                                                         header, (int)P.cap_contours, P.npts.as<int>(), P.chunks.as<int2>(),
                                                         P.chunk_meta.as<int2>(), cap_chunks);
    } else {
        trace_window_kernel<<<(unsigned)P.cap_contours, 64, 0, st>>>(ws.fgbits.as<uint32_t>(), h, w, wpitch, P.starts.as<int>(),
                                                                    P.start_slice.as<int>(), header, (int)P.cap_contours, P.npts.as<int>(),
                                                                    P.chunks.as<int2>(), P.chunk_meta.as<int2>(), cap_chunks);
    }
    MS_LAUNCH_CHECK();
}
void launch_gather(PolyDev& P, double sx, double sy, cudaStream_t st) {
    const long long cap_chunks = chunk_capacity(P);
    gather_kernel<<<(unsigned)cdiv64(cap_chunks * kChunk, 256), 256, 0, st>>>(P.chunks.as<int2>(), P.chunk_meta.as<int2>(), P.header.as<long long>(),
                                                                             cap_chunks, P.npts.as<int>(), (int)P.cap_contours,
                                                                             (long long)P.cap_points, sx, sy, P.xy.as<int2>());
    MS_LAUNCH_CHECK();
}

// list-ranking variant of the counting pass: leaves per-contour vertex counts in P.npts and {pixel, kept} per border
// position in ws.crack_pos (consumed by crack_emit)
void crack_count(M2pWs& ws, PolyDev& P, int h, int w, int batch, cudaStream_t st) {
    const int wpitch = cdiv(w, 32);
    const size_t ncracks = (size_t)h * w * batch * 4;
    const int nblocks = (int)cdiv64((int64_t)ncracks, 1024);
    ws.crack_pair.reserve(ncracks * 8);
    ws.crack_pos.reserve(ncracks * 4);
    ws.crack_blocks.reserve((size_t)nblocks * 4);
    ws.crack_contour.reserve(((size_t)P.cap_contours + 1) * 8);
    ws.crack_meta.reserve(8 + 32 * 4);
    crack::u64* pair = ws.crack_pair.as<crack::u64>();
    long long* meta = ws.crack_meta.as<long long>();
    int* busy = reinterpret_cast<int*>(meta + 1);
    int* cbase = ws.crack_contour.as<int>();
    int* rot = cbase + P.cap_contours + 1;
    const uint32_t* B = ws.fgbits.as<uint32_t>();
    long long* header = P.header.as<long long>();
    const dim3 gw = ccl::grid_for(h, wpitch, batch);
    const int gc = (int)cdiv64(P.cap_contours, 256);
    int rounds = 0;
    while (((int64_t)1 << rounds) < (int64_t)h * w * 4) ++rounds;   // a border has at most 4 * h * w cracks

    MS_CUDA(cudaMemsetAsync(meta, 0, 8 + 32 * 4, st));
    crack::init_kernel<<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, pair);
    MS_LAUNCH_CHECK();
    crack::cut_kernel<<<gc, 256, 0, st>>>(B, h, w, wpitch, P.starts.as<int>(), P.start_slice.as<int>(), header, (int)P.cap_contours, pair,
                                          rot, P.npts.as<int>());
    MS_LAUNCH_CHECK();
    for (int r = 0; r < rounds; ++r) {
        crack::jump_kernel<<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, pair, busy, r);
        MS_LAUNCH_CHECK();
        crack::check_kernel<<<gc, 256, 0, st>>>(h, w, P.starts.as<int>(), P.start_slice.as<int>(), header, (int)P.cap_contours, pair, busy, r);
        MS_LAUNCH_CHECK();
    }
    crack::length_kernel<<<1, 1024, 0, st>>>(h, w, P.starts.as<int>(), P.start_slice.as<int>(), header, (int)P.cap_contours, pair, cbase, meta);
    MS_LAUNCH_CHECK();
    crack::flags_kernel<<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, pair, cbase, rot, ws.crack_pos.as<uint32_t>(), P.npts.as<int>());
    MS_LAUNCH_CHECK();
    crack::pos_count_kernel<<<nblocks, 1024, 0, st>>>(ws.crack_pos.as<uint32_t>(), meta, ws.crack_blocks.as<int>());
    MS_LAUNCH_CHECK();
    crack::pos_scan_kernel<<<1, 1024, 0, st>>>(ws.crack_blocks.as<int>(), meta);
    MS_LAUNCH_CHECK();
}

void crack_emit(M2pWs& ws, PolyDev& P, int h, int w, int batch, double sx, double sy, cudaStream_t st) {
    const int nblocks = (int)cdiv64((int64_t)h * w * batch * 4, 1024);
    crack::pos_emit_kernel<<<nblocks, 1024, 0, st>>>(ws.crack_pos.as<uint32_t>(), ws.crack_meta.as<long long>(), ws.crack_blocks.as<int>(),
                                                     P.header.as<long long>(), (long long)P.cap_points, w, sx, sy, P.xy.as<int2>());
    MS_LAUNCH_CHECK();
}

void default_capacities(PolyDev& P, int batch) {
    if (P.cap_contours == 0) P.cap_contours = std::max<int64_t>(1024, 64 * (int64_t)batch);
    if (P.cap_points == 0) P.cap_points = std::max<int64_t>(65536, 4096 * (int64_t)batch);
}

}  // namespace

// mapped host buffer for the phase timestamps of slice 0 (MEDSEG_FUSED_DBG=1, tools/fused_phases.py); null when off
long long* fused_debug_buffer(bool create) {
    static long long* buf = nullptr;
    if (!buf && (create || std::getenv("MEDSEG_FUSED_DBG"))) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 32 * sizeof(long long), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
            buf = static_cast<long long*>(p);
            std::memset(buf, 0, 32 * sizeof(long long));
        } else {
            cudaGetLastError();
        }
    }
    return buf;
}

bool slice_fused_supported(int h, int w) {
    const char* e = std::getenv("MEDSEG_FUSED");      // read per call: the tests flip it to run both paths in one process
    if (e && e[0] == '0') return false;
    return h > 0 && w > 0 && fused::plan(h, w).ok;
}

void slice_fused_launch(FusedWs& fws, M2pWs* ws, PolyDev* P, const uint8_t* d_in, uint8_t* d_out, int h, int w, int batch, bool do_post,
                        bool do_poly, int fg_value, float min_area_ratio, int thr, cudaStream_t st, const FgSpec* multi) {
    MS_REQUIRE(h > 0 && w > 0 && batch > 0 && batch <= 65535, MS_ERR_ARG, "slice kernel: bad shape");
    const fused::Plan pl = fused::plan(h, w);
    MS_REQUIRE(pl.ok, MS_ERR_INTERNAL, "slice kernel: slice does not fit in shared memory");
    MS_REQUIRE(!do_poly || (P && ws), MS_ERR_INTERNAL, "slice kernel: polygon store missing");
    fused::Params a{};
    a.in = d_in; a.out = d_out;
    a.H = h; a.W = w; a.wpitch = cdiv(w, 32); a.batch = batch;
    a.do_post = do_post; a.do_poly = do_poly;
    a.fg = multi ? *multi : FgSpec::single(fg_value, batch);
    a.thr = thr;
    // src/postprocess.cpp:30,66: static_cast<int>(w * h * MIN_AREA_RATIO) -- int product, float multiply, truncate
    a.min_area = static_cast<int>(static_cast<float>(w * h) * min_area_ratio);
    a.off_z = pl.off_z; a.off_y = pl.off_y; a.off_roff = pl.off_roff; a.off_tab = pl.off_tab;
    a.rcap = pl.rcap; a.cap_border = pl.cap_border;
    a.g_runs = h * a.wpitch * 16;
    fws.tab.reserve((size_t)batch * 3 * a.g_runs * 4);
    a.g_tab = fws.tab.as<int>();
    if (do_poly) {
        default_capacities(*P, batch);
        P->slice_info.reserve((size_t)batch * sizeof(int4));
        P->rec.reserve((size_t)P->cap_contours * sizeof(int2));
        P->vstore.reserve((size_t)P->cap_points * 4);
        P->starts.reserve((size_t)P->cap_contours * 4);
        P->npts.reserve(((size_t)P->cap_contours + 1) * 4);
        P->slice_start.reserve(((size_t)batch + 1) * 4);
        P->xy.reserve((size_t)P->cap_points * 8);
        P->header.reserve(8 * sizeof(long long));
        ws->crack_contour.reserve(((size_t)P->cap_contours + 1) * srank::kInfo * sizeof(int));
        a.slice_info = P->slice_info.as<int4>(); a.rec = P->rec.as<int2>(); a.vstore = P->vstore.as<uint32_t>();
        a.starts = P->starts.as<int>(); a.cinfo = ws->crack_contour.as<int>();
        a.header = P->header.as<unsigned long long>();
        a.cap_contours = (int)std::min<int64_t>(P->cap_contours, 0x7FFFFFF0);
        a.cap_points = P->cap_points;
        MS_CUDA(cudaMemsetAsync(P->header.p, 0, 8 * sizeof(long long), st));
        P->fused = true;
    }
    a.dbg = fused_debug_buffer(false);
    set_max_dynamic_smem(fused::slice_kernel, (int)fused::kSmemTotal);
    fused::slice_kernel<<<batch, fused::kT, fused::kSmemTotal, st>>>(a);
    MS_LAUNCH_CHECK();
}

void post_poly_phase_a(PostprocessWs& pws, M2pWs& ws, PolyDev& P, const uint8_t* d_raw, uint8_t* d_clean, int h, int w, int batch,
                       int fg_value, float min_area_ratio, cudaStream_t st, const FgSpec* multi) {
    if (slice_fused_supported(h, w)) {
        slice_fused_launch(ws.fused, &ws, &P, d_raw, d_clean, h, w, batch, true, true, fg_value, min_area_ratio, 0, st, multi);
        return;
    }
    postprocess_launch(pws, d_raw, d_clean, h, w, batch, fg_value, min_area_ratio, st, multi);
    // mask_to_image + threshold(127) (src/process.cpp:234, src/mask2polygon.cpp:31): after postprocess the mask is {0, fg};
    // LUT(fg) > 127 <=> value == fg <=> value > fg - 1
    // (several labels: every clean mask is {0, its label}, so "value > 0" is the same test for all of them)
    m2p_phase_a(ws, P, d_clean, h, w, batch, multi ? 0 : fg_value - 1, st);
}

void m2p_phase_a(M2pWs& ws, PolyDev& P, const uint8_t* d_mask, int h, int w, int batch, int threshold, cudaStream_t st) {
    MS_REQUIRE(h > 0 && w > 0 && batch > 0 && h <= 65535 && batch <= 32767, MS_ERR_ARG, "mask2polygon: bad shape");
    MS_REQUIRE((int64_t)h * w <= (int64_t)(0x7FFFFFFF - 8) / 8, MS_ERR_ARG, "mask2polygon: slice too large");
    if (slice_fused_supported(h, w) && !std::getenv("MEDSEG_TRACE")) {
        slice_fused_launch(ws.fused, &ws, &P, d_mask, nullptr, h, w, batch, false, true, 0, 0.0f, threshold, st);
        return;
    }
    P.fused = false;
    const int n = h * w;
    const size_t nb = (size_t)n * batch;
    default_capacities(P, batch);
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
    P.header.reserve(8 * sizeof(long long));
    P.chunks.reserve((size_t)chunk_capacity(P) * kChunk * sizeof(int2));
    P.chunk_meta.reserve((size_t)chunk_capacity(P) * sizeof(int2));
    int* Lfg = ws.fg.labels.as<int>();
    int* Lbg = ws.bg.labels.as<int>();
    uint8_t* flag = ws.bg.flag.as<uint8_t>();
    uint32_t* B = ws.fgbits.as<uint32_t>();
    int* bc = P.block_counts.as<int>();
    int* slice_total = bc + (size_t)bps * batch;
    long long* header = P.header.as<long long>();
    const dim3 gw = ccl::grid_for(h, wpitch, batch);

    thr_bits_kernel<<<dim3(cdiv(w, 256), h, batch), 256, 0, st>>>(d_mask, h, w, wpitch, threshold, B, Lfg, Lbg, flag);
    MS_LAUNCH_CHECK();
    ccl::merge_fg8_bg4_kernel<<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lfg, Lbg);      // 8-connected foreground, 4-connected background
    MS_LAUNCH_CHECK();
    ccl::resolve_fg_bg_kernel<<<gw, ccl::kThreads, 0, st>>>(B, h, w, wpitch, Lfg, Lbg, flag);
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
    const TraceMode mode = pick_trace_mode(h, w, batch);
    if (mode == kTraceCrack) crack_count(ws, P, h, w, batch, st);
    else launch_trace(ws, P, mode, h, w, batch, st);
    scan_points_kernel<<<1, 1024, 0, st>>>(P.npts.as<int>(), (int)P.cap_contours, header);
    MS_LAUNCH_CHECK();
}

// Opt-in simplification of the polygon set phase B just wrote in NETWORK space: approxPolyDP per contour, new contour
// offsets, then the coordinate mapping.  Three launches, all sizes on the device.
static void dp_launch(PolyDev& P, int h, int w, double eps, double sx, double sy, cudaStream_t st) {
    // exact integer distances: |cross|^2 and d^2 * L stay below 2^62
    MS_REQUIRE((int64_t)w * w + (int64_t)h * h < ((int64_t)1 << 31), MS_ERR_ARG, "dp_epsilon: slice too large for exact simplification");
    const int cap_c = (int)std::min<int64_t>(P.cap_contours, 0x7FFFFFF0);
    P.dp_tmp.reserve((size_t)P.cap_points * 8);
    P.dp_list.reserve((size_t)P.cap_points * 8);
    P.dp_keep.reserve(((size_t)P.cap_points / 32 + (size_t)P.cap_contours + 2) * 4);
    P.dp_cnt.reserve((size_t)P.cap_contours * 4);
    P.dp_old.reserve(((size_t)P.cap_contours + 1) * 4);
    dp::Args a{};
    a.xy = P.xy.as<int2>(); a.cstart = P.npts.as<int>(); a.header = P.header.as<long long>();
    a.cap_contours = cap_c; a.cap_points = (long long)P.cap_points;
    a.eps2 = eps * eps;                                 // the library squares eps once, in double
    a.g_list = P.dp_list.as<uint2>(); a.g_keep = P.dp_keep.as<uint32_t>(); a.tmp = P.dp_tmp.as<int2>(); a.cnt = P.dp_cnt.as<int>();
    set_max_dynamic_smem(dp::simplify_kernel, (int)dp::kSmemBytes);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(P.cap_contours, 148 * 3));
    dp::simplify_kernel<<<grid, dp::kT, dp::kSmemBytes, st>>>(a);
    MS_LAUNCH_CHECK();
    dp::rescan_kernel<<<1, 1024, 0, st>>>(P.npts.as<int>(), P.dp_old.as<int>(), P.dp_cnt.as<int>(), P.header.as<long long>(), cap_c,
                                          (long long)P.cap_points);
    MS_LAUNCH_CHECK();
    dp::compact_kernel<<<grid, 256, 0, st>>>(P.dp_tmp.as<int2>(), P.dp_old.as<int>(), P.npts.as<int>(), P.dp_cnt.as<int>(),
                                             P.header.as<long long>(), sx, sy, P.xy.as<int2>());
    MS_LAUNCH_CHECK();
}

void m2p_phase_b(M2pWs& ws, PolyDev& P, int h, int w, int batch, int orig_w, int orig_h, cudaStream_t st) {
    // src/mask2polygon.cpp:199-200
    const double sx = static_cast<double>(orig_w) / w;
    const double sy = static_cast<double>(orig_h) / h;
    // with simplification on, the vertices are first emitted unmapped ((int)(x * 1.0) == x) and mapped after it
    const bool simplify = ws.dp_eps > 0.0;
    const double ex = simplify ? 1.0 : sx, ey = simplify ? 1.0 : sy;
    if (P.fused) {
        fused::finalize_kernel<<<batch, 256, 0, st>>>(P.slice_info.as<int4>(), P.rec.as<int2>(), P.vstore.as<uint32_t>(), batch,
                                                     (int)std::min<int64_t>(P.cap_contours, 0x7FFFFFF0), (long long)P.cap_points, ex, ey,
                                                     P.slice_start.as<int>(), P.npts.as<int>(), P.xy.as<int2>(), P.header.as<long long>());
        MS_LAUNCH_CHECK();
    } else {
        const TraceMode mode = pick_trace_mode(h, w, batch);
        if (mode == kTraceCrack) crack_emit(ws, P, h, w, batch, ex, ey, st);
        else launch_gather(P, ex, ey, st);
    }
    if (simplify) dp_launch(P, h, w, ws.dp_eps, sx, sy, st);
}

}  // namespace ms
