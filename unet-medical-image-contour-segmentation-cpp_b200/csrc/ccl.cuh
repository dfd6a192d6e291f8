// ccl.cuh -- run-based union-find connected-component labelling on bit-packed rows (header-only).
//
// One labeller, parametrised by connectivity, used four times on the hot path (SURVEY.md hard part H3):
//   postprocess  : 8-connected components of the *inverse* foreground   (src/postprocess.cpp:26)
//                  8-connected components of the opened foreground       (src/postprocess.cpp:64)
//   mask2polygon : 8-connected foreground components  (what cv::findContours traces)
//                  4-connected background components  (the RETR_EXTERNAL test)
// It replaces cv::connectedComponentsWithStats; label numbering differs (the reference's results do not depend on
// numbering): a component's label is the slice-local linear index of its raster-first pixel, which is exactly the
// start pixel cv::findContours uses.
//
// Representation: the predicate image is one bit per pixel, 32 pixels per word (`wpitch` words per row).  A *run* is a
// maximal sequence of set bits inside one word; its *head* is its first pixel.  Only heads own an entry of the int32
// label plane, so the label traffic scales with the number of runs, not pixels (CT-like masks: ~2 runs per row), and
// everything else -- neighbourhood tests, hole filling, 3x3 morphology -- is word-wide bit arithmetic.  One thread owns
// one word:
//   heads   : L[head] = head (area / flag cleared)                       -- fused into the kernels that produce the bits
//   merge   : seam with the previous word, then for the row above only the leftmost pixel of every "both rows set"
//             stretch (plus the two diagonal cases for 8-connectivity) issues a lock-free atomicMin union
//   resolve : every head is compressed to its root; per-root area (run lengths) and "touches the border" flag
// After resolve the root of any pixel is L[head of its run], two dependent loads.
#pragma once
#include "common.cuh"

namespace ms {
namespace ccl {

constexpr int kThreads = 256;

struct BitImage {           // one slice
    const uint32_t* bits;   // H rows x wpitch words
    int H, W, wpitch;
};

__device__ __forceinline__ int ld_label(const int* L, int i) { return __ldcg(L + i); }

__device__ __forceinline__ int find_root(const int* L, int a) {
    int p;
    while ((p = ld_label(L, a)) != a) a = p;
    return a;
}

// Lock-free union keeping the smaller index as root (stale reads are safe: a lost race shows up
// as old != expected and the loop continues with the value actually stored).
__device__ __forceinline__ void unite(int* L, int a, int b) {
    bool done;
    do {
        a = find_root(L, a);
        b = find_root(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

__device__ __forceinline__ uint32_t valid_mask(int W, int wx) {   // bits of word wx that lie inside the image
    const int rem = W - wx * 32;
    return rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}
__device__ __forceinline__ uint32_t load_word(const uint32_t* bits, int H, int wpitch, int y, int wx) {
    return (y >= 0 && y < H && wx >= 0 && wx < wpitch) ? __ldg(bits + (size_t)y * wpitch + wx) : 0u;
}
// first bit of the run (inside the word) that contains set bit x
__device__ __forceinline__ int run_head_bit(uint32_t bits, int x) {
    const uint32_t zeros_below = ~bits & ((1u << x) - 1u);
    return zeros_below ? 32 - __clz(zeros_below) : 0;
}
__device__ __forceinline__ uint32_t head_mask(uint32_t bits) { return bits & ~(bits << 1); }
// mask of the run starting at head bit x
__device__ __forceinline__ uint32_t run_mask(uint32_t bits, int x) {
    const uint32_t above = ~(bits >> x);                 // first zero at or above x
    const int len = above ? __ffs((int)above) - 1 : 32 - x;
    return (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << x;
}
// head pixel (linear index) of the run containing pixel (X, y); the bit must be set
__device__ __forceinline__ int head_pixel(const uint32_t* bits, int wpitch, int W, int y, int X) {
    const int wx = X >> 5;
    const uint32_t b = __ldg(bits + (size_t)y * wpitch + wx);
    return y * W + (wx << 5) + run_head_bit(b, X & 31);
}
// root of the component containing pixel (X, y) -- valid after resolve_kernel
__device__ __forceinline__ int root_of(const uint32_t* bits, const int* L, int wpitch, int W, int y, int X) {
    return L[head_pixel(bits, wpitch, W, y, X)];
}

// word-parallel kernels: thread t of slice blockIdx.y owns word t of the slice (row t / wpitch, word t % wpitch);
// grid = (ceil(H * wpitch / 256), batch)
#define MS_CCL_WORD_COORDS()                                            \
    const int widx = blockIdx.x * ::ms::ccl::kThreads + threadIdx.x;      \
    if (widx >= H * wpitch) return;                                      \
    const int y = widx / wpitch, wx = widx - y * wpitch;                 \
    const int sl = blockIdx.y

// ---- heads: L[head] = head, statistics cleared, for the runs of one word.  INVERT labels the complement of `bits`.
template <bool INVERT>
__device__ __forceinline__ void init_heads(uint32_t b, int W, int y, int wx, int* __restrict__ L, int* __restrict__ area,
                                           uint8_t* __restrict__ flag) {   // L / area / flag: this slice's planes
    if (INVERT) b = ~b & valid_mask(W, wx);
    uint32_t h = head_mask(b);
    while (h) {
        const int x = __ffs((int)h) - 1;
        h &= h - 1;
        const int p = y * W + wx * 32 + x;
        L[p] = p;
        if (area) area[p] = 0;
        if (flag) flag[p] = 0;
    }
}
// ---- merge.  One thread per word.  CONN = 4 or 8.
template <int CONN, bool INVERT>
__device__ __forceinline__ void merge_word(const uint32_t* __restrict__ B, int W, int wpitch, int y, int wx, int* __restrict__ L) {
    auto word = [&](int yy, int ww) -> uint32_t {
        if (yy < 0 || ww < 0 || ww >= wpitch) return 0u;
        uint32_t v = __ldg(B + (size_t)yy * wpitch + ww);
        return INVERT ? (~v & valid_mask(W, ww)) : v;
    };
    const uint32_t cur = word(y, wx);
    if (cur == 0) return;
    const uint32_t cl = word(y, wx - 1), cr = word(y, wx + 1);
    const uint32_t up = word(y - 1, wx), ul = word(y - 1, wx - 1), ur = word(y - 1, wx + 1);
    const int row0 = y * W + wx * 32;
    auto head_cur = [&](int x) { return row0 + run_head_bit(cur, x); };
    auto head_up = [&](int X) {   // X in -1 .. 32 relative to this word
        if (X < 0) return row0 - W - 32 + run_head_bit(ul, 31);
        if (X > 31) return row0 - W + 32;                       // bit 0 of the right word is its own run head
        return row0 - W + run_head_bit(up, X);
    };
    // horizontal: runs inside a word are one run by construction; only the seam with the previous word remains
    if ((cur & 1u) && (cl >> 31)) unite(L, row0, row0 - 32 + run_head_bit(cl, 31));
    if (y == 0) return;
    const uint32_t curW = (cur << 1) | (cl >> 31), curE = (cur >> 1) | (cr << 31);
    const uint32_t upNW = (up << 1) | (ul >> 31), upNE = (up >> 1) | (ur << 31);
    // the leftmost pixel of a stretch where both rows are set links the two runs
    uint32_t m = cur & up & ~(curW & upNW);
    while (m) {
        const int x = __ffs((int)m) - 1;
        m &= m - 1;
        unite(L, head_cur(x), head_up(x));
    }
    if (CONN == 8) {
        m = cur & ~up & upNW & ~curW;      // if W is set, W links to NW (its N) itself
        while (m) {
            const int x = __ffs((int)m) - 1;
            m &= m - 1;
            unite(L, head_cur(x), head_up(x - 1));
        }
        m = cur & ~up & upNE & ~curE;      // if E is set, E links to NE (its N) itself
        while (m) {
            const int x = __ffs((int)m) - 1;
            m &= m - 1;
            unite(L, head_cur(x), head_up(x + 1));
        }
    }
}
template <int CONN, bool INVERT>
__global__ void __launch_bounds__(kThreads) merge_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                          int* __restrict__ L_all) {
    MS_CCL_WORD_COORDS();
    merge_word<CONN, INVERT>(bits_all + (size_t)sl * H * wpitch, W, wpitch, y, wx, L_all + (size_t)sl * H * W);
}
// the two labellings mask2polygon needs, in one launch: 8-connected foreground into L_fg, 4-connected background into L_bg
static __global__ void __launch_bounds__(kThreads) merge_fg8_bg4_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                                  int* __restrict__ L_fg, int* __restrict__ L_bg) {
    MS_CCL_WORD_COORDS();
    const uint32_t* B = bits_all + (size_t)sl * H * wpitch;
    merge_word<8, false>(B, W, wpitch, y, wx, L_fg + (size_t)sl * H * W);
    merge_word<4, true>(B, W, wpitch, y, wx, L_bg + (size_t)sl * H * W);
}

// ---- resolve.  One thread per word: heads -> roots, per-root area and border flag.
template <bool INVERT>
__device__ __forceinline__ void resolve_word(uint32_t b, int H, int W, int y, int wx, int* __restrict__ L, int* __restrict__ area,
                                             uint8_t* __restrict__ flag) {   // L / area / flag: this slice's planes
    if (INVERT) b = ~b & valid_mask(W, wx);
    uint32_t h = head_mask(b);
    while (h) {
        const int x = __ffs((int)h) - 1;
        h &= h - 1;
        const int p = y * W + wx * 32 + x;
        const int r = find_root(L, p);
        L[p] = r;
        const uint32_t rm = run_mask(b, x);
        if (area) atomicAdd(&area[r], __popc(rm));
        if (flag) {
            const int x_first = wx * 32 + x, x_last = wx * 32 + 31 - __clz(rm);
            if (y == 0 || y == H - 1 || x_first == 0 || x_last == W - 1) flag[r] = 1;
        }
    }
}
template <bool INVERT>
__global__ void __launch_bounds__(kThreads) resolve_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                            int* __restrict__ L_all, int* __restrict__ area_all,
                                                            uint8_t* __restrict__ flag_all) {
    MS_CCL_WORD_COORDS();
    const size_t slice = (size_t)sl * H * W;
    resolve_word<INVERT>(bits_all[((size_t)sl * H + y) * wpitch + wx], H, W, y, wx, L_all + slice, area_all ? area_all + slice : nullptr,
                         flag_all ? flag_all + slice : nullptr);
}
static __global__ void __launch_bounds__(kThreads) resolve_fg_bg_kernel(const uint32_t* __restrict__ bits_all, int H, int W, int wpitch,
                                                                  int* __restrict__ L_fg, int* __restrict__ L_bg,
                                                                  uint8_t* __restrict__ bg_flag) {
    MS_CCL_WORD_COORDS();
    const size_t slice = (size_t)sl * H * W;
    const uint32_t b = bits_all[((size_t)sl * H + y) * wpitch + wx];
    resolve_word<false>(b, H, W, y, wx, L_fg + slice, nullptr, nullptr);
    resolve_word<true>(b, H, W, y, wx, L_bg + slice, nullptr, bg_flag + slice);
}

inline dim3 grid_for(int H, int wpitch, int batch) { return dim3(cdiv(H * wpitch, kThreads), batch); }

}  // namespace ccl
}  // namespace ms
