// dp_simplify.cuh -- opt-in closed-curve Douglas-Peucker on the device polygon set (included by mask2polygon.cu).
//
// NOT on the reference's path: src/mask2polygon.cpp:34 stops at CHAIN_APPROX_SIMPLE.  BASELINE.json's north_star (3) names
// "Douglas-Peucker simplification in shared memory" as part of the B200 design, so it exists behind "dp_epsilon" (0 = off,
// the default) and, when on, reproduces cv2.approxPolyDP(contour, eps, closed = true) of OpenCV 4.13 vertex for vertex
// (oracle/pipeline.py: approx_poly_dp is the restatement, pinned by fuzzing against cv2; tests/test_gpu_dp.py).
//
// One CTA per contour, vertices staged in shared memory (x | y << 16), contours in flight = grid size:
//   1. three block-wide "farthest vertex from the current start" sweeps pick the two seed split points (the library's
//      approximation of the diameter; the output starts at the second);
//   2. the library's explicit stack becomes LEVEL-SYNCHRONOUS work lists: a slice (s0, s1) is independent of every other
//      slice, so all slices of a level are examined at once -- one warp per slice (the whole CTA per slice while a level has
//      fewer slices than warps): arg-max of the distance to the SEGMENT s0-s1 over the interior vertices, first maximum in
//      walk order, compared exactly (integers on the common scale |s1 - s0|^2).  A slice within eps marks its start vertex
//      in a bitmap, otherwise its two halves go to the next level's list.  The output of the recursion is the marked
//      vertices in walk order from the first seed, whatever order the slices were examined in;
//   3. bitmap -> vertex list (warp prefix sums), then the library's sequential clean-up pass over the few kept vertices.
// Contours longer than kSmemPts vertices run the same code on global scratch (generic pointers).
// Then `rescan_kernel` turns the new counts into contour offsets and `compact_kernel` copies the kept vertices to their
// final place with the (int)(x * scale) mapping of src/mask2polygon.cpp:54-55.
#pragma once

namespace dp {

constexpr int kT = 256;
constexpr int kW = kT / 32;
constexpr int kSmemPts = 6144;
constexpr int kLaneSlice = 48;          // slices with at most this many interior vertices are scanned by a single lane
// vertices u32[kSmemPts] | work lists uint2[kSmemPts] (two lists of n / 2 slices; later the kept vertices int2[n]) | bitmap
constexpr size_t kSmemBytes = (size_t)kSmemPts * 4 + (size_t)kSmemPts * 8 + (size_t)(kSmemPts / 32) * 4;

struct Args {
    const int2* xy;            // network-space vertices (phase B run with unit scale)
    const int* cstart;         // [n_contours + 1]
    const long long* header;   // {n_contours, n_points, overflow flags, ...}
    int cap_contours;
    long long cap_points;
    double eps2;
    uint2* g_list;             // [cap_points]: work lists of contours longer than kSmemPts
    uint32_t* g_keep;          // [cap_points / 32 + cap_contours + 1]: their bitmaps
    int2* tmp;                 // [cap_points]: simplified vertices of contour c at cstart[c]
    int* cnt;                  // [cap_contours]: simplified vertex counts
};

struct Best {
    unsigned long long v;
    int k;
};
__device__ __forceinline__ Best better(Best a, Best b) { return (b.v > a.v || (b.v == a.v && b.k < a.k)) ? b : a; }
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best t;
        t.v = __shfl_xor_sync(0xFFFFFFFFu, b.v, o);
        t.k = __shfl_xor_sync(0xFFFFFFFFu, b.k, o);
        b = better(b, t);
    }
    return b;
}
__device__ __forceinline__ Best block_best(Best b, Best* s_best) {
    b = warp_best(b);
    __syncthreads();                       // s_best may still be read from the previous call
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = b;
    __syncthreads();
    Best r = s_best[0];
#pragma unroll
    for (int i = 1; i < kW; ++i) r = better(r, s_best[i]);
    return r;
}

__device__ __forceinline__ bool results_valid(const long long* header, int cap_contours, long long cap_points) {
    const unsigned long long* hu = reinterpret_cast<const unsigned long long*>(header);
    return (hu[2] & 7ull) == 0 && header[0] <= cap_contours && header[1] <= cap_points;
}

__global__ void __launch_bounds__(kT) simplify_kernel(const Args a) {
    extern __shared__ __align__(16) uint8_t dp_smem[];
    uint32_t* s_pts = reinterpret_cast<uint32_t*>(dp_smem);
    uint2* s_list = reinterpret_cast<uint2*>(dp_smem + (size_t)kSmemPts * 4);
    uint32_t* s_keep = reinterpret_cast<uint32_t*>(dp_smem + (size_t)kSmemPts * 12);
    __shared__ Best s_best[kW];
    __shared__ int s_n[2];                 // slices in the current / the next level's list
    __shared__ int s_m;
    if (!results_valid(a.header, a.cap_contours, a.cap_points)) return;   // the caller grows the buffers and runs again
    const int n_contours = (int)a.header[0];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int c = blockIdx.x; c < n_contours; c += gridDim.x) {
        const int off = a.cstart[c], cnt = a.cstart[c + 1] - off;
        if (cnt <= 0) {
            if (tid == 0) a.cnt[c] = 0;
            continue;
        }
        const bool fits = cnt <= kSmemPts;
        const int2* gp = a.xy + off;
        uint2* list = fits ? s_list : a.g_list + off;
        uint32_t* keep = fits ? s_keep : a.g_keep + (off >> 5) + c;
        __syncthreads();                   // the previous contour's shared memory is no longer read
        if (fits)
            for (int i = tid; i < cnt; i += kT) {
                const int2 p = gp[i];
                s_pts[i] = (uint32_t)p.x | ((uint32_t)p.y << 16);
            }
        for (int i = tid; i < (cnt + 31) >> 5; i += kT) keep[i] = 0u;
        if (tid == 0) { s_n[0] = 0; s_n[1] = 0; }
        __syncthreads();
        auto P = [&](int i) -> int2 {
            if (fits) {
                const uint32_t u = s_pts[i];
                return make_int2((int)(u & 0xFFFFu), (int)(u >> 16));
            }
            return gp[i];
        };
        auto interior = [&](int s0, int s1) { const int d = s1 - s0; return (d < 0 ? d + cnt : d) - 1; };

        // ---- 1. seeds: start at vertex 0, jump to the farthest vertex, three times
        int pos = 0, right_start = 0;
        bool le_eps = false;
        for (int it = 0; it < 3; ++it) {
            pos += right_start;
            if (pos >= cnt) pos -= cnt;
            const int2 s = P(pos);
            Best b{0ull, 0x7FFFFFFF};
            for (int j = 1 + tid; j < cnt; j += kT) {
                int idx = pos + j;
                if (idx >= cnt) idx -= cnt;
                const int2 p = P(idx);
                const long long dx = p.x - s.x, dy = p.y - s.y;
                const unsigned long long d = (unsigned long long)(dx * dx + dy * dy);
                if (d > b.v) { b.v = d; b.k = j; }
            }
            b = block_best(b, s_best);
            if (b.v > 0ull) right_start = b.k;
            le_eps = __ull2double_rn(b.v) <= a.eps2;
        }
        if (le_eps) {                      // the whole contour lies within eps of one vertex
            if (tid == 0) {
                a.tmp[off] = P(pos);
                a.cnt[c] = 1;
            }
            continue;
        }
        const int seed0 = pos;
        int seed1 = right_start + seed0;
        if (seed1 >= cnt) seed1 -= cnt;

        // ---- 2. level-synchronous splitting
        const int half = cnt >> 1;         // a level holds at most cnt / 2 slices with an interior (2 vertices each, disjoint)
        uint2* cur = list;
        uint2* nxt = list + half;
        auto keep_vertex = [&](int i) { atomicOr(&keep[i >> 5], 1u << (i & 31)); };
        auto child = [&](uint2* dst, int s0, int s1) {
            if (interior(s0, s1) <= 0) keep_vertex(s0);
            else dst[atomicAdd(&s_n[1], 1)] = make_uint2((uint32_t)s0, (uint32_t)s1);
        };
        // lane-local first maximum of the distance to segment s0-s1 over interior vertices first, first + stride, ...
        auto scan_slice = [&](int s0, int s1, int first, int stride, long long& L) -> Best {
            const int2 A = P(s0), B = P(s1);
            const long long dx = B.x - A.x, dy = B.y - A.y;
            L = dx * dx + dy * dy;
            const unsigned long long Ls = (unsigned long long)(L > 0 ? L : 1);
            const int n_in = interior(s0, s1);
            Best b{0ull, 0x7FFFFFFF};
            auto dist = [&](const int2 p) -> unsigned long long {
                const long long px = p.x - A.x, py = p.y - A.y;
                const long long dot = px * dx + py * dy;
                if (L == 0 || dot < 0) return (unsigned long long)(px * px + py * py) * Ls;
                if (dot > L) {
                    const long long qx = p.x - B.x, qy = p.y - B.y;
                    return (unsigned long long)(qx * qx + qy * qy) * Ls;
                }
                const long long cr = py * dx - px * dy;
                return (unsigned long long)(cr * cr);
            };
            // four vertices per step, all four loads in flight together: a contour too long for shared memory is read from
            // L2, and one dependent load per iteration made the 51 k-vertex cfg5 contour take milliseconds
            int k = 1 + first;
            for (; k + 3 * stride <= n_in; k += 4 * stride) {
                int2 p[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int idx = s0 + k + u * stride;
                    if (idx >= cnt) idx -= cnt;
                    p[u] = P(idx);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long v = dist(p[u]);
                    if (v > b.v) { b.v = v; b.k = k + u * stride; }
                }
            }
            for (; k <= n_in; k += stride) {
                int idx = s0 + k;
                if (idx >= cnt) idx -= cnt;
                const unsigned long long v = dist(P(idx));
                if (v > b.v) { b.v = v; b.k = k; }
            }
            return b;
        };
        auto decide = [&](uint2* dst, int s0, int s1, Best b, long long L) {
            const double den = (double)(L > 0 ? L : 1);
            if (__ull2double_rn(b.v) <= __dmul_rn(a.eps2, den)) {
                keep_vertex(s0);
            } else {
                int m = s0 + b.k;
                if (m >= cnt) m -= cnt;
                child(dst, s0, m);
                child(dst, m, s1);
            }
        };
        if (tid == 0) {
            child(cur, seed0, seed1);
            child(cur, seed1, seed0);
            s_n[0] = s_n[1];
            s_n[1] = 0;
        }
        __syncthreads();
        while (true) {
            const int nc = s_n[0];
            if (nc == 0) break;
            if (nc >= kW) {
                // long slices: one warp each.  On the global-scratch path (a contour too long for shared memory) the short ones
                // -- the bulk of the deep levels -- take one LANE each, 256 slices at a time: a warp per 5-vertex slice spent its
                // time on the slice's fixed L2 latencies (list entry, end points, atomics).  From shared memory the warp form is
                // the faster one.
                const int lane_max = fits ? 0 : kLaneSlice;
                for (int it = warp; it < nc; it += kW) {
                    const uint2 s = cur[it];
                    if (interior((int)s.x, (int)s.y) <= lane_max) continue;
                    long long L;
                    Best b = warp_best(scan_slice((int)s.x, (int)s.y, lane, 32, L));
                    if (lane == 0) decide(nxt, (int)s.x, (int)s.y, b, L);
                }
                for (int it = tid; it < nc; it += kT) {
                    const uint2 s = cur[it];
                    if (interior((int)s.x, (int)s.y) > lane_max) continue;
                    long long L;
                    const Best b = scan_slice((int)s.x, (int)s.y, 0, 1, L);
                    decide(nxt, (int)s.x, (int)s.y, b, L);
                }
            } else {
                for (int it = 0; it < nc; ++it) {
                    const uint2 s = cur[it];
                    long long L;
                    Best b = block_best(scan_slice((int)s.x, (int)s.y, tid, kT, L), s_best);
                    if (tid == 0) decide(nxt, (int)s.x, (int)s.y, b, L);
                }
            }
            __syncthreads();
            if (tid == 0) {
                s_n[0] = s_n[1];
                s_n[1] = 0;
            }
            uint2* t = cur; cur = nxt; nxt = t;
            __syncthreads();
        }

        // ---- 3. marked vertices in walk order from seed0 -> dst, then the clean-up pass
        int2* dst = fits ? reinterpret_cast<int2*>(s_list) : a.tmp + off;
        if (warp == 0) {
            int base = 0;
            for (int pass = 0; pass < 2; ++pass) {
                const int lo = pass ? 0 : seed0, hi = pass ? seed0 : cnt;
                if (hi <= lo) continue;
                const int w_lo = lo >> 5, w_hi = (hi - 1) >> 5;
                for (int w0 = w_lo; w0 <= w_hi; w0 += 32) {
                    const int w = w0 + lane;
                    uint32_t word = w <= w_hi ? keep[w] : 0u;
                    if (w == w_lo) word &= 0xFFFFFFFFu << (lo & 31);
                    if (w == w_hi && (hi & 31)) word &= (1u << (hi & 31)) - 1u;
                    const int n = __popc(word);
                    int incl = n;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    int at = base + incl - n;
                    while (word) {
                        const int bit = __ffs(word) - 1;
                        word &= word - 1;
                        dst[at++] = P(w * 32 + bit);
                    }
                    base += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
            }
            if (lane == 0) s_m = base;
        }
        __syncthreads();
        if (tid == 0) {
            // the library's last stage, in place with wrap-around: drop a vertex that lies within eps / sqrt(2) of the chord of
            // its neighbours when the chord is not axis-parallel and the path does not turn back
            const int count = s_m;
            int new_count = count;
            auto rd = [&](int& p) { const int2 v = dst[p]; if (++p >= count) p = 0; return v; };
            int p = count - 1;
            int2 start_pt = rd(p);
            int wpos = p;
            int2 pt = rd(p);
            for (int i = 0; i < count && new_count > 2; ++i) {
                const int2 end_pt = rd(p);
                const double dx = (double)(end_pt.x - start_pt.x), dy = (double)(end_pt.y - start_pt.y);
                const double dist = fabs(__dsub_rn(__dmul_rn((double)(pt.x - start_pt.x), dy), __dmul_rn((double)(pt.y - start_pt.y), dx)));
                const long long sip = (long long)(pt.x - start_pt.x) * (end_pt.x - pt.x) + (long long)(pt.y - start_pt.y) * (end_pt.y - pt.y);
                const double lim = __dmul_rn(__dmul_rn(0.5, a.eps2), __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                if (__dmul_rn(dist, dist) <= lim && dx != 0.0 && dy != 0.0 && sip >= 0) {
                    --new_count;
                    dst[wpos] = start_pt = end_pt;
                    if (++wpos >= count) wpos = 0;
                    pt = rd(p);
                    ++i;
                    continue;
                }
                dst[wpos] = start_pt = pt;
                if (++wpos >= count) wpos = 0;
                pt = end_pt;
            }
            s_m = new_count;
            a.cnt[c] = new_count;
        }
        __syncthreads();
        if (fits) {
            const int m = s_m;
            for (int i = tid; i < m; i += kT) a.tmp[off + i] = dst[i];
        }
    }
}

// one CTA of 1024: old offsets -> old_start, exclusive scan of the simplified counts -> cstart, header[1] = new total
__global__ void __launch_bounds__(1024) rescan_kernel(int* __restrict__ cstart, int* __restrict__ old_start, const int* __restrict__ cnt,
                                                       long long* __restrict__ header, int cap_contours, long long cap_points) {
    const bool ok = results_valid(header, cap_contours, cap_points);
    if (threadIdx.x == 0) header[7] = ok ? 1 : 0;      // tells compact_kernel whether tmp / cnt / old_start are this batch's
    if (!ok) return;
    const int n = (int)header[0];
    long long carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? cnt[i] : 0;
        if (i < n) old_start[i] = cstart[i];
        int tot;
        const int ex = block_exscan_1024(v, &tot);
        if (i < n) cstart[i] = (int)(carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0) {
        cstart[n] = (int)carry;
        header[6] = header[1];             // vertices before simplification (diagnostics)
        header[1] = carry;
    }
}

__global__ void __launch_bounds__(256) compact_kernel(const int2* __restrict__ tmp, const int* __restrict__ old_start, const int* __restrict__ cstart,
                                                       const int* __restrict__ cnt, const long long* __restrict__ header, double sx, double sy,
                                                       int2* __restrict__ xy) {
    if (header[7] == 0) return;
    const int n = (int)header[0];
    for (int c = blockIdx.x; c < n; c += gridDim.x) {
        const int2* src = tmp + old_start[c];
        int2* dst = xy + cstart[c];
        const int m = cnt[c];
        for (int i = threadIdx.x; i < m; i += 256) dst[i] = map_point(src[i], sx, sy);
    }
}

}  // namespace dp
