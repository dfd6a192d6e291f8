// ccl.cuh -- warp-level union-find connected-component labelling (header-only templates).
//
// One labeller, parametrised by pixel predicate and connectivity, used four times on the hot path
// (SURVEY.md hard part H3):
//   postprocess  : 8-connected components of the *inverse* foreground   (src/postprocess.cpp:26)
//                  8-connected components of the opened foreground       (src/postprocess.cpp:64)
//   mask2polygon : 8-connected foreground components  (what cv::findContours traces)
//                  4-connected background components  (the RETR_EXTERNAL test)
// It replaces cv::connectedComponentsWithStats; label numbering differs (the reference's results do
// not depend on numbering): here a component's label is the slice-local linear index of its
// raster-first pixel, which is exactly the start pixel cv::findContours uses.
//
// Three passes over int32 labels (4 B/px scratch, not algorithmic traffic):
//   init    : each warp owns a 32-pixel row segment; one ballot gives every pixel the index of the
//             first pixel of its horizontal run inside the segment (runs are pre-merged for free).
//   merge   : unions across segment boundaries and with the row above, pruned so that only the
//             leftmost pixel of every "both rows set" stretch issues a union (lock-free atomicMin).
//   resolve : path-compress to the root, then per-component area (warp-aggregated atomics) and a
//             "touches the image border" flag.
#pragma once
#include "common.cuh"

namespace ms {
namespace ccl {

constexpr int kThreads = 256;  // 8 warps = 8 consecutive 32-pixel segments of one row

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

// grid = (ceil(W / 256), H, batch)
template <class Pred>
__global__ void __launch_bounds__(kThreads) init_kernel(const uint8_t* __restrict__ mask, int H, int W, Pred pred,
                                                         int* __restrict__ labels, int* __restrict__ area,
                                                         uint8_t* __restrict__ flag) {
    const int x = blockIdx.x * kThreads + threadIdx.x, y = blockIdx.y;
    const size_t slice = (size_t)blockIdx.z * H * W;
    const int lane = threadIdx.x & 31;
    const bool in = x < W;
    const int p = y * W + x;
    const bool fg = in && pred(mask[slice + p]);
    const unsigned bits = __ballot_sync(0xFFFFFFFFu, fg);
    if (!in) return;
    int lab = -1;
    if (fg) {
        const unsigned zeros_below = ~bits & ((1u << lane) - 1u);
        const int start = zeros_below ? 32 - __clz(zeros_below) : 0;
        lab = p - lane + start;
    }
    labels[slice + p] = lab;
    if (area) area[slice + p] = 0;
    if (flag) flag[slice + p] = 0;
}

// grid = (ceil(W / 256), H, batch).  CONN = 4 or 8.
template <int CONN>
__global__ void __launch_bounds__(kThreads) merge_kernel(int* __restrict__ labels_all, int H, int W) {
    const int x = blockIdx.x * kThreads + threadIdx.x, y = blockIdx.y;
    int* L = labels_all + (size_t)blockIdx.z * H * W;
    const int lane = threadIdx.x & 31;
    const int seg0 = x - lane;  // first column of this warp's segment
    const bool in = x < W;
    const int p = y * W + x;
    const bool fg = in && (ld_label(L, p) >= 0);
    const bool up = in && y > 0 && (ld_label(L, p - W) >= 0);
    const unsigned cur = __ballot_sync(0xFFFFFFFFu, fg);
    const unsigned upb = __ballot_sync(0xFFFFFFFFu, up);
    if (cur == 0) return;  // warp-uniform
    // columns just outside the segment (loaded by every lane from the same address: one broadcast)
    const bool w_out = seg0 > 0 && (ld_label(L, y * W + seg0 - 1) >= 0);
    const bool e_out = seg0 + 32 < W && (ld_label(L, y * W + seg0 + 32) >= 0);
    const bool nw_out = y > 0 && seg0 > 0 && (ld_label(L, (y - 1) * W + seg0 - 1) >= 0);
    const bool ne_out = y > 0 && seg0 + 32 < W && (ld_label(L, (y - 1) * W + seg0 + 32) >= 0);
    if (!fg) return;
    const bool Wn = lane > 0 ? ((cur >> (lane - 1)) & 1u) : w_out;
    const bool En = lane < 31 ? ((cur >> (lane + 1)) & 1u) : e_out;
    const bool NWn = lane > 0 ? ((upb >> (lane - 1)) & 1u) : nw_out;
    const bool NEn = lane < 31 ? ((upb >> (lane + 1)) & 1u) : ne_out;
    const bool Nn = up;
    // horizontal: runs inside a segment were merged by init; only the segment seam remains
    if (lane == 0 && Wn) unite(L, p, p - 1);
    if (Nn) {
        // the leftmost pixel of a stretch where both rows are set links the two runs
        if (!(Wn && NWn)) unite(L, p, p - W);
    } else if (CONN == 8) {
        if (NWn && !Wn) unite(L, p, p - W - 1);  // if W is set, W links to NW (its N) itself
        if (NEn && !En) unite(L, p, p - W + 1);  // if E is set, E links to NE (its N) itself
    }
}

// grid = (ceil(W / 256), H, batch).  After this pass labels[p] is the component root (raster-first
// pixel), area[root] the pixel count and flag[root] != 0 iff the component touches the image border.
static __global__ void __launch_bounds__(kThreads) resolve_kernel(int* __restrict__ labels_all, int H, int W,
                                                            int* __restrict__ area_all, uint8_t* __restrict__ flag_all) {
    const int x = blockIdx.x * kThreads + threadIdx.x, y = blockIdx.y;
    const size_t slice = (size_t)blockIdx.z * H * W;
    int* L = labels_all + slice;
    const bool in = x < W;
    const int p = y * W + x;
    int r = -1;
    if (in && ld_label(L, p) >= 0) {
        r = find_root(L, p);
        L[p] = r;
    }
    if (area_all) {
        // one atomic per distinct root per warp
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, r);
        if (r >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&area_all[slice + r], __popc(peers));
    }
    if (flag_all && r >= 0 && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) flag_all[slice + r] = 1;
}

inline dim3 grid_for(int H, int W, int batch) { return dim3(cdiv(W, kThreads), H, batch); }

}  // namespace ccl
}  // namespace ms
