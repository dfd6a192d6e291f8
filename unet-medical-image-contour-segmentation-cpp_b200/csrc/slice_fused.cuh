// slice_fused.cuh -- K5 + K6 for one slice in ONE CTA, entirely in shared memory (fragment of mask2polygon.cu: it reuses
// the crack / srank helpers of that translation unit and is included inside `namespace ms { namespace {`).
//
// Replaces, for slices whose bit image fits on chip (<= ~13 k words: 512 x 512, 512 x 768, ...), the 19 launches of
// postprocess.cu + mask2polygon.cu -- /root/reference/src/postprocess.cpp:13-79 and src/mask2polygon.cpp:29-36 -- by
//     slice_kernel     one CTA of 1,024 threads per slice:
//        u8 class mask -> fg bits -> 8-conn CCL of the inverse + area / border flag -> hole fill -> 3 x 3 open ->
//        8-conn CCL + area filter -> clean u8 mask out; the kept components' roots ARE mask2polygon's contour
//        starts (a component's root is its raster-first pixel) -> 4-conn background CCL + frame flag (RETR_EXTERNAL
//        test) -> starts in descending raster order -> crack list ranking (srank) -> CHAIN_APPROX_SIMPLE ->
//        per-contour vertex counts and the kept vertices, packed x | y << 16, into a per-slice contiguous range of the
//        vertex store
//     finalize_kernel  one CTA per slice: prefix over the slices' totals (slice order), contour offsets, coordinate
//        mapping (int)(x * scale) -> the CSR polygon set the C ABI returns.  Re-runnable with another mapping.
// The HBM traffic is the algorithmic minimum: the class mask is read once, the clean mask written once, the vertices
// written twice (packed, then mapped).  All label / run / crack tables live in shared memory:
//   * runs (maximal horizontal set-bit sequences of a row, across words) are numbered in raster order by a prefix sum
//     over the words' run-head counts, so a pixel's run is roff[word] + popc(heads at or below it) - 1; labels and areas
//     are indexed by run number (int32, up to `rcap` runs on chip; a noisier slice uses a global scratch table through
//     the same pointers).  A CT-like slice has 2-3 runs per row: ~1.5 k table entries, ~1.5 k unions;
//   * union-find: atomicMin on the label table, smaller run number = root = raster-first run;
//   * area and the "touches the image border" flag share one word (bit 31 = flag).
// Modes: do_post && do_poly (the pipeline), do_post only (ms_postprocess_*), do_poly only (ms_mask2polygon_*: bits =
// mask > thr, every component is kept).
namespace fused {

constexpr int kT = 1024;
constexpr uint32_t kFlagBit = 0x80000000u;
constexpr size_t kSmemTotal = 216 * 1024;     // dynamic shared memory of slice_kernel (+ ~8.5 KiB static)

struct Params {
    const uint8_t* in;       // [batch][H][W]
    uint8_t* out;            // [batch][H][W] clean mask {0, fg} (do_post; may be null)
    int H, W, wpitch, batch;
    int do_post, do_poly;
    FgSpec fg;               // do_post: label kept per (virtual) slice and the class mask it reads
    int min_area, thr;
    // shared-memory layout (word offsets into the dynamic array), computed on the host by plan()
    int off_z, off_y, off_roff, off_tab, rcap, cap_border;
    // global fallback run tables: [batch][3][g_runs]  (labels, areas, run head positions)
    int* g_tab;
    int g_runs;
    // polygon staging
    int4* slice_info;        // [batch] {n_contours, n_points, record base, vertex base}  (base < 0: did not fit)
    int2* rec;               // [cap_contours] {slice-local vertex offset, vertex count}
    uint32_t* vstore;        // [cap_points] x | y << 16, slice-local CSR order
    int* starts;             // [cap_contours]
    int* cinfo;              // [srank::kInfo * (cap_contours + 1)]
    unsigned long long* header;   // [8]: 2 overflow flags, 3 errors, 4 contour records requested, 5 vertices requested
    int cap_contours;
    long long cap_points;
    long long* dbg;          // optional [32]: clock64 of slice 0 at the phase boundaries (tools/fused_phases.py)
};
#define MS_FUSED_MARK(k)                                             \
    do {                                                             \
        if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[k] = clock64(); \
    } while (0)

struct Plan {
    bool ok;
    int off_z, off_y, off_roff, off_tab, rcap, cap_border;
};
inline Plan plan(int H, int W) {
    Plan p{};
    const int64_t wpitch = cdiv(W, 32), nw = (int64_t)H * wpitch, PB = (int64_t)(H + 2) * (wpitch + 2);
    const int64_t total = (int64_t)kSmemTotal / 4;
    const int64_t off_y = std::max<int64_t>(2 * nw, PB);
    const int64_t off_roff = off_y + nw, off_tab = off_roff + nw + 1;
    const int64_t rcap = (total - off_tab) / 3;       // labels, areas, head positions
    const int64_t k6_fixed = PB * 4 + ((nw + 1) & ~(int64_t)1) * 2 + 16;
    const int64_t cb = std::min<int64_t>(((int64_t)kSmemTotal - k6_fixed) / srank::kBytesPerBorder, srank::kMaxBorder);
    p.ok = rcap >= 512 && cb >= 256 && H <= 65535 && W <= 65535 && nw * 16 < (1 << 30);
    p.off_z = (int)nw; p.off_y = (int)off_y; p.off_roff = (int)off_roff; p.off_tab = (int)off_tab;
    p.rcap = (int)std::max<int64_t>(rcap, 0);
    p.cap_border = (int)std::max<int64_t>(cb, 0);
    return p;
}

__device__ __forceinline__ uint32_t le_mask(int x) { return (2u << x) - 1u; }      // bits 0 .. x  (x = 31: all ones)
__device__ __forceinline__ uint32_t heads(uint32_t b) { return b & ~(b << 1); }    // run heads inside one word

// A run table (labels or areas).  SMEM: explicit shared-state-space accesses -- a `volatile` access through a generic pointer
// compiles to LD.E.STRONG.SYS, ~50x the latency of an LDS, and pointer chasing is nothing but dependent loads.
template <bool SMEM>
struct Tab {
    int* p;
    __device__ __forceinline__ uint32_t sa(int i) const { return (uint32_t)__cvta_generic_to_shared(p) + 4u * (uint32_t)i; }
    __device__ __forceinline__ int ld(int i) const {
        if (SMEM) {
            int v;
            asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(sa(i)) : "memory");
            return v;
        }
        return __ldcg(p + i);
    }
    __device__ __forceinline__ void st(int i, int v) const {
        if (SMEM) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(sa(i)), "r"(v) : "memory");
        else __stcg(p + i, v);
    }
    __device__ __forceinline__ int amin(int i, int v) const {
        if (SMEM) {
            int o;
            asm volatile("atom.shared.min.s32 %0, [%1], %2;" : "=r"(o) : "r"(sa(i)), "r"(v) : "memory");
            return o;
        }
        return atomicMin(p + i, v);
    }
    __device__ __forceinline__ void add(int i, int v) const {
        if (SMEM) asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(sa(i)), "r"(v) : "memory");
        else atomicAdd(p + i, v);
    }
    __device__ __forceinline__ void bor(int i, uint32_t v) const {
        if (SMEM) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(sa(i)), "r"(v) : "memory");
        else atomicOr(reinterpret_cast<unsigned*>(p) + i, v);
    }
};

// ---- everything below works on bit planes in shared memory: nw = H * wpitch words, row-major, bits beyond W are zero
struct Geo {
    int H, W, wpitch, nw, wp_shift;      // wp_shift >= 0: wpitch is that power of two (512-wide slices: 4)
    __device__ __forceinline__ void split(int w, int& y, int& wx) const {
        if (wp_shift >= 0) {
            y = w >> wp_shift;
            wx = w & (wpitch - 1);
        } else {
            y = w / wpitch;
            wx = w - y * wpitch;
        }
    }
    __device__ __forceinline__ int col(int w) const { return wp_shift >= 0 ? (w & (wpitch - 1)) : (w % wpitch); }
};
// heads of the ROW runs that start inside word w: a set bit whose left neighbour (possibly in the previous word) is clear
__device__ __forceinline__ uint32_t row_heads(const uint32_t* bits, const Geo& g, int w) {
    const uint32_t b = bits[w];
    if (b == 0) return 0u;
    const uint32_t prev = g.col(w) > 0 ? bits[w - 1] : 0u;
    return b & ~((b << 1) | (prev >> 31));
}
// run number of set bit x of a word with row-run heads `rh` and run offset `off` (a run entering from the previous word
// is the last one started before this word: off - 1)
__device__ __forceinline__ int run_id(uint32_t rh, int off, int x) { return off + __popc(rh & le_mask(x)) - 1; }

// find with path halving (every value written is an ancestor, and atomicMin keeps parents decreasing)
template <bool SMEM>
__device__ __forceinline__ int find_root_h(const Tab<SMEM>& L, int a) {
    for (;;) {
        const int p = L.ld(a);
        if (p == a) return a;
        const int g = L.ld(p);
        if (g == p) return p;
        L.amin(a, g);
        a = g;
    }
}
template <bool SMEM>
__device__ __forceinline__ void unite_h(const Tab<SMEM>& L, int a, int b) {
    bool done;
    do {
        a = find_root_h(L, a);
        b = find_root_h(L, b);
        if (a < b) {
            const int old = L.amin(b, a);
            done = old == b;
            b = old;
        } else if (b < a) {
            const int old = L.amin(a, b);
            done = old == a;
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// exclusive prefix of count(w) over the words: off[w], off[nw] = total.  Warp i owns a contiguous block of words and walks
// it 32 words (one per lane) at a time, so count() runs once per word and shared memory is read without bank conflicts.
template <class OffT, class F>
__device__ __forceinline__ int scan_words(int nw, OffT* off, F count, bool write_total = true) {
    __shared__ int wtot[32];
    __shared__ int total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = (((nw + 31) >> 5) + 31) & ~31;
    const int base = warp * per_warp;
    int carry = 0;
    for (int j = 0; j < per_warp; j += 32) {
        const int w = base + j + lane;
        const int c = w < nw ? count(w) : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        if (w < nw) off[w] = (OffT)(carry + inc - c);
        carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) wtot[warp] = carry;
    __syncthreads();
    if (warp == 0) {
        const int v = wtot[lane];
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        wtot[lane] = inc - v;
        if (lane == 31) total = inc;
    }
    __syncthreads();
    const int wb = wtot[warp];
    if (wb)
        for (int j = lane; j < per_warp && base + j < nw; j += 32) off[base + j] = (OffT)(off[base + j] + wb);
    const int tot = total;
    if (threadIdx.x == 0 && write_total) off[nw] = (OffT)tot;
    __syncthreads();
    return tot;
}

// adds {pixels, border flag} to area[root]; lanes of the warp that hold the same root combine first (a CT-like slice is a
// handful of components: without this every run would hit the same shared-memory word).  Called by all 32 lanes.
template <bool SMEM>
__device__ __forceinline__ void flush_area_warp(const Tab<SMEM>& area, int root, int acc, bool edge) {
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, root);
    const int lane = threadIdx.x & 31, leader = __ffs((int)peers) - 1;
    const int sum = __reduce_add_sync(peers, acc);
    const unsigned e = __reduce_or_sync(peers, edge ? 1u : 0u);
    if (lane != leader || root < 0) return;
    if (sum) area.add(root, sum);
    if (e) area.bor(root, kFlagBit);
}

// Connected components of the bit plane (CONN = 4 or 8) over row-run numbers.  On return roff = run offsets per word,
// lab[run] = root run (the component's raster-first run), area[root] = pixel count | kFlagBit if the component touches
// the image border.  Returns the table pointers actually used (shared memory, or the global fallback for > rcap runs).
struct Labels {
    int* lab;
    int* area;
};
// visits every vertical link of the plane: f(run in this row, run in the row above).  A link = the leftmost pixel of a
// stretch where this row and the row above are both set, plus the two diagonal cases for 8-connectivity; horizontal
// adjacency is inside a run by construction.
template <int CONN, class F>
__device__ __forceinline__ void for_each_link(const uint32_t* bits, const Geo& g, const uint32_t* roff, F&& f) {
    const int nw = g.nw, wpitch = g.wpitch;
    for (int w = wpitch + threadIdx.x; w < nw; w += kT) {
        const uint32_t cur = bits[w];
        if (cur == 0) continue;
        const int wx = g.col(w);
        const uint32_t up = bits[w - wpitch];
        const uint32_t cl = wx > 0 ? bits[w - 1] : 0u, ul = wx > 0 ? bits[w - wpitch - 1] : 0u;
        if ((cur & up) == 0xFFFFFFFFu && (cl & ul) >> 31) continue;          // interior of a solid region: no stretch starts here
        const uint32_t cr = wx + 1 < wpitch ? bits[w + 1] : 0u, ur = wx + 1 < wpitch ? bits[w - wpitch + 1] : 0u;
        const uint32_t curW = (cur << 1) | (cl >> 31), upNW = (up << 1) | (ul >> 31);
        uint32_t m = cur & up & ~(curW & upNW);
        uint32_t mw = 0, me = 0;
        if (CONN == 8) {
            const uint32_t curE = (cur >> 1) | (cr << 31), upNE = (up >> 1) | (ur << 31);
            mw = cur & ~up & upNW & ~curW;       // if W is set, W links to NW (its N) itself
            me = cur & ~up & upNE & ~curE;       // if E is set, E links to NE (its N) itself
        }
        if ((m | mw | me) == 0) continue;
        const uint32_t rh_c = cur & ~((cur << 1) | (cl >> 31)), rh_u = up & ~((up << 1) | (ul >> 31));
        const int oc = (int)roff[w], ou = (int)roff[w - wpitch];
        auto run_up = [&](int X) {               // X in -1 .. 32 relative to this word
            if (X < 0) return ou - 1;            // bit 31 of the word left of `up`: the run entering (or last before) `up`
            if (X > 31) return (int)roff[w - wpitch + 1] + (int)((ur & 1u) & ~(up >> 31)) - 1;
            return run_id(rh_u, ou, X);
        };
        while (m) {
            const int x = __ffs((int)m) - 1;
            m &= m - 1;
            f(run_id(rh_c, oc, x), run_up(x));
        }
        while (mw) {
            const int x = __ffs((int)mw) - 1;
            mw &= mw - 1;
            f(run_id(rh_c, oc, x), run_up(x - 1));
        }
        while (me) {
            const int x = __ffs((int)me) - 1;
            me &= me - 1;
            f(run_id(rh_c, oc, x), run_up(x + 1));
        }
    }
}
// pointer jumping until every run points at its root
template <bool SMEM>
__device__ __forceinline__ int compress(const Tab<SMEM>& lab, int n_runs) {
    int round = 0;
    for (; round < 40; ++round) {
        int changed = 0;
        for (int i = threadIdx.x; i < n_runs; i += kT) {
            const int p = lab.ld(i);
            if (p != i) {
                const int gp = lab.ld(p);
                if (gp != p) {
                    lab.st(i, gp);
                    changed = 1;
                }
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    return round + 1;
}

template <int CONN, bool SMEM>
__device__ __forceinline__ void label_body(const uint32_t* bits, const Geo& g, const uint32_t* roff, int n_runs, int* lab_p, int* area_p,
                                           int* hpos_p, long long* dbg) {
#define MS_LABEL_MARK(k)                                              \
    do {                                                              \
        if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[k] = clock64(); \
    } while (0)
    const int nw = g.nw, wpitch = g.wpitch;
    const Tab<SMEM> lab{lab_p}, area{area_p}, hpos{hpos_p};
    for (int i = threadIdx.x; i < n_runs; i += kT) {
        lab.st(i, i);
        area.st(i, 0);
    }
    __syncthreads();
    MS_LABEL_MARK(0);
    // 1. hook: every run takes the smallest run above it that it touches -- one atomicMin per link, nothing is chased, so
    //    all rows hook at once (uniting as we go would make sweep k walk the chain sweep k - 1 left: a 512-row background
    //    is a 512-deep list)
    for_each_link<CONN>(bits, g, roff, [&](int rc, int ru) { lab.amin(rc, ru); });
    __syncthreads();
    MS_LABEL_MARK(6);
    // 2. flatten the forest
    const int rounds1 = compress(lab, n_runs);
    MS_LABEL_MARK(1);
    // 3. links the hook did not use join different trees (U shapes): unite their roots -- paths are one step long now
    int joined = 0;
    for_each_link<CONN>(bits, g, roff, [&](int rc, int ru) {
        const int a = lab.ld(rc), b = lab.ld(ru);
        if (a != b) {
            unite_h(lab, a, b);
            joined = 1;
        }
    });
    // 4. and flatten again where anything was joined
    const int any_joined = __syncthreads_or(joined);
    MS_LABEL_MARK(7);
    const int rounds2 = any_joined ? compress(lab, n_runs) : 0;
    MS_LABEL_MARK(2);
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
        dbg[4] = n_runs;
        dbg[5] = rounds1 * 100 + rounds2;
    }
    // 5. pixel counts and border flags per root.  A row run is contiguous, so its size is tail - head + 1: the thread that
    //    owns a head records its x, the thread that owns the tail adds the length -- only words with a run end do anything.
    //    Every thread accumulates for the root it saw last (a column of words mostly stays inside one component) and the
    //    warp combines equal roots once, at the end.
    for (int w = threadIdx.x; w < nw; w += kT) {
        const uint32_t rh = row_heads(bits, g, w);
        if (rh == 0) continue;
        const int wx = g.col(w), off = (int)roff[w];
        uint32_t h = rh;
        int id = off;
        while (h) {
            hpos.st(id++, wx * 32 + __ffs((int)h) - 1);
            h &= h - 1;
        }
    }
    __syncthreads();
    MS_LABEL_MARK(8);
    {
        int root = -1, acc = 0;
        bool edge = false;
        for (int w = threadIdx.x; w < nw; w += kT) {
            const uint32_t cur = bits[w];
            if (cur == 0) continue;
            int y, wx;
            g.split(w, y, wx);
            const uint32_t nx = wx + 1 < wpitch ? bits[w + 1] : 0u;
            uint32_t tails = cur & ~((cur >> 1) | (nx << 31));
            if (tails == 0) continue;
            const uint32_t cl = wx > 0 ? bits[w - 1] : 0u;
            const uint32_t rh = cur & ~((cur << 1) | (cl >> 31));
            const int off = (int)roff[w];
            const bool row_edge = y == 0 || y == g.H - 1;
            while (tails) {
                const int x = __ffs((int)tails) - 1;
                tails &= tails - 1;
                const int id = run_id(rh, off, x);
                const int r = lab.ld(id), x0 = hpos.ld(id), x1 = wx * 32 + x;
                const bool e = row_edge || x0 == 0 || x1 == g.W - 1;
                if (r != root) {
                    if (root >= 0) {
                        area.add(root, acc);
                        if (edge) area.bor(root, kFlagBit);
                    }
                    root = r;
                    acc = 0;
                    edge = false;
                }
                acc += x1 - x0 + 1;
                edge |= e;
            }
        }
        flush_area_warp(area, root, acc, edge);
    }
    __syncthreads();
    MS_LABEL_MARK(3);
#undef MS_LABEL_MARK
}
template <int CONN>
__device__ Labels label_runs(const uint32_t* bits, const Geo& g, uint32_t* roff, int* s_tab, int rcap, int* g_tab, int g_runs,
                             long long* dbg = nullptr) {
    const int n_runs = scan_words(g.nw, roff, [&](int w) { return __popc(row_heads(bits, g, w)); });
    if (n_runs <= rcap) {
        label_body<CONN, true>(bits, g, roff, n_runs, s_tab, s_tab + rcap, s_tab + 2 * rcap, dbg);
        return Labels{s_tab, s_tab + rcap};
    }
    label_body<CONN, false>(bits, g, roff, n_runs, g_tab, g_tab + g_runs, g_tab + 2 * (size_t)g_runs, dbg);   // noisy slice: tables in L2
    return Labels{g_tab, g_tab + g_runs};
}

// 3 x 3 erode (ERODE) or dilate of plane S into D with OpenCV's default border (outside pixels never win the min / max).
// One thread = one word column of an 8-row strip: every new row costs three loads, the previous two stay in registers.
template <bool ERODE>
__device__ void morph3(const uint32_t* S, uint32_t* D, const Geo& g) {
    constexpr int R = 8;
    const int wpitch = g.wpitch, H = g.H;
    const int n_strips = wpitch * ((H + R - 1) / R);
    for (int s = threadIdx.x; s < n_strips; s += kT) {
        int rb, wx;
        g.split(s, rb, wx);
        const uint32_t valid = ccl::valid_mask(g.W, wx);
        const bool has_l = wx > 0, has_r = wx + 1 < wpitch;
        auto hrow = [&](int y) -> uint32_t {
            if (y < 0 || y >= H) return ERODE ? 0xFFFFFFFFu : 0u;
            const uint32_t* r = S + y * wpitch + wx;
            if (ERODE) {
                const uint32_t c = r[0] | ~valid, l = has_l ? r[-1] : 0xFFFFFFFFu, rr = has_r ? r[1] : 0xFFFFFFFFu;
                return c & ((c << 1) | (l >> 31)) & ((c >> 1) | (rr << 31));
            }
            const uint32_t c = r[0], l = has_l ? r[-1] : 0u, rr = has_r ? r[1] : 0u;
            return c | (c << 1) | (l >> 31) | (c >> 1) | (rr << 31);
        };
        const int y0 = rb * R, y1 = min(H, y0 + R);
        uint32_t a = hrow(y0 - 1), b = hrow(y0);
        for (int y = y0; y < y1; ++y) {
            const uint32_t c = hrow(y + 1);
            D[y * wpitch + wx] = (ERODE ? (a & b & c) : (a | b | c)) & valid;
            a = b;
            b = c;
        }
    }
    __syncthreads();
}

// 4 mask bytes -> 4 bits (bit i = byte i matches)
__device__ __forceinline__ uint32_t match4(uint32_t q, uint32_t key4, bool eq) {
    const uint32_t r = eq ? __vcmpeq4(q, key4) : __vcmpgtu4(q, key4);
    return ((r & 0x01010101u) * 0x01020408u) >> 24;
}

__global__ void __launch_bounds__(kT, 1) slice_kernel(const Params P) {
    extern __shared__ uint32_t smem[];
    __shared__ uint16_t lut[kTraceLutEntries];      // only the sequential fallback uses it
    __shared__ int s_base[2];
    constexpr int kSC = 64;                         // contours whose scratch stays on chip (postprocess keeps <= 1 / 0.06 = 16)
    __shared__ int s_contour[5 * kSC];
    const int tid = threadIdx.x, b = blockIdx.x;
    const int H = P.H, W = P.W, wpitch = P.wpitch, nw = H * wpitch;
    const Geo g{H, W, wpitch, nw, (wpitch & (wpitch - 1)) == 0 ? 31 - __clz(wpitch) : -1};
    uint32_t* X = smem;
    uint32_t* Z = smem + P.off_z;
    uint32_t* Y = smem + P.off_y;
    uint32_t* roff = smem + P.off_roff;
    int* s_tab = reinterpret_cast<int*>(smem + P.off_tab);
    int* g_tab = P.g_tab + (size_t)b * 3 * P.g_runs;
    const uint8_t* in = P.in + (size_t)(P.do_post ? P.fg.src(b) : b) * H * W;
    const int fg_value = P.fg.fg(b);
    MS_FUSED_MARK(0);

    // ---------------------------------------------------------------- mask -> bits (X)
    {
        const uint32_t key = (uint32_t)(P.do_post ? fg_value : P.thr) & 0xFFu, key4 = key * 0x01010101u;
        const bool eq = P.do_post != 0;       // postprocess: mask == FG (postprocess.cpp:18);  mask2polygon: mask > thr (mask2polygon.cpp:31)
        if ((W & 31) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
            const int n16 = H * W / 16;     // 16 pixels per thread and step; lane pairs build one word
            for (int base = 0; base < n16; base += 4 * kT) {
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * kT + tid;
                    q[u] = i < n16 ? __ldg(reinterpret_cast<const uint4*>(in) + i) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * kT + tid;
                    const uint32_t m = match4(q[u].x, key4, eq) | (match4(q[u].y, key4, eq) << 4) | (match4(q[u].z, key4, eq) << 8) |
                                       (match4(q[u].w, key4, eq) << 12);
                    const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, m, 1);
                    if (!(tid & 1) && i < n16) X[i >> 1] = m | (other << 16);
                }
            }
        } else {
            for (int w = tid; w < nw; w += kT) {
                int y, wx;
                g.split(w, y, wx);
                uint32_t m = 0;
                for (int j = 0; j < 32 && wx * 32 + j < W; ++j) {
                    const uint32_t v = in[(size_t)y * W + wx * 32 + j];
                    m |= (uint32_t)(eq ? v == key : v > key) << j;
                }
                X[w] = m;
            }
        }
    }
    __syncthreads();
    MS_FUSED_MARK(1);

    if (P.do_post) {
        // ------------------------------------------------------------ holes: 8-conn CCL of the inverse  (postprocess.cpp:18-43)
        for (int w = tid; w < nw; w += kT) Z[w] = ~X[w] & ccl::valid_mask(W, g.col(w));      // inv = ~bin  (:22)
        __syncthreads();
        {
            const Labels L = label_runs<8>(Z, g, roff, s_tab, P.rcap, g_tab, P.g_runs, P.dbg ? P.dbg + 20 : nullptr);
            MS_FUSED_MARK(2);
            for (int w = tid; w < nw; w += kT) {
                const uint32_t inv = Z[w];
                uint32_t out = X[w];
                if (inv) {
                    const uint32_t rh = row_heads(Z, g, w);
                    const int off = (int)roff[w];
                    uint32_t h = heads(inv);
                    while (h) {
                        const int x = __ffs((int)h) - 1;
                        h &= h - 1;
                        const uint32_t a = (uint32_t)L.area[L.lab[run_id(rh, off, x)]];
                        if (!(a & kFlagBit) && (int)a < P.min_area) out |= ccl::run_mask(inv, x);     // :40-41
                    }
                }
                Y[w] = out;
            }
            __syncthreads();
            MS_FUSED_MARK(3);
        }
        // ------------------------------------------------------------ 3 x 3 open  (postprocess.cpp:57-60): Y -> Z -> X
        morph3<true>(Y, Z, g);
        morph3<false>(Z, X, g);
        MS_FUSED_MARK(4);
    }

    // ---------------------------------------------------------------- 8-conn components of X: area filter, roots
    //   Y = kept components (= mask2polygon's foreground), Z = one bit per kept component at its raster-first pixel
    {
        const Labels L = label_runs<8>(X, g, roff, s_tab, P.rcap, g_tab, P.g_runs);
        MS_FUSED_MARK(5);
        const int min_area = P.do_post ? P.min_area : 0;
        uint8_t* out = (P.do_post && P.out) ? P.out + (size_t)b * H * W : nullptr;
        const uint32_t v = (uint32_t)fg_value;
        for (int w = tid; w < nw; w += kT) {
            const uint32_t o = X[w];
            uint32_t keep = 0, roots = 0;
            if (o) {
                const uint32_t rh = row_heads(X, g, w);
                const int off = (int)roff[w];
                uint32_t h = heads(o);
                while (h) {
                    const int x = __ffs((int)h) - 1;
                    h &= h - 1;
                    const int run = run_id(rh, off, x);
                    const int r = L.lab[run];
                    if ((int)((uint32_t)L.area[r] & ~kFlagBit) >= min_area) {          // postprocess.cpp:66-72
                        keep |= ccl::run_mask(o, x);
                        if (r == run && ((rh >> x) & 1u)) roots |= 1u << x;          // the component's raster-first pixel
                    }
                }
            }
            Y[w] = keep;
            Z[w] = roots;
            if (out) {
                int y, wx;
                g.split(w, y, wx);
                uint8_t* dst = out + (size_t)y * W + wx * 32;
                if (wx * 32 + 32 <= W && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    uint32_t wds[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) wds[j] = ((((keep >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u) * v;   // 4 bits -> 4 bytes
                    reinterpret_cast<uint4*>(dst)[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                    reinterpret_cast<uint4*>(dst)[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
                } else {
                    for (int j = 0; j < 32 && wx * 32 + j < W; ++j) dst[j] = (keep >> j) & 1u ? (uint8_t)fg_value : (uint8_t)0;
                }
            }
        }
        __syncthreads();
        MS_FUSED_MARK(6);
    }
    if (!P.do_poly) return;

    // ---------------------------------------------------------------- 4-conn background + frame flag; external starts
    int n_c = 0;
    {
        for (int w = tid; w < nw; w += kT) X[w] = ~Y[w] & ccl::valid_mask(W, g.col(w));
        __syncthreads();
        const Labels L = label_runs<4>(X, g, roff, s_tab, P.rcap, g_tab, P.g_runs);
        MS_FUSED_MARK(7);
        // a root is external <=> the pixel left of it belongs to a background component reaching the frame (x == 0: the
        // frame itself).  SURVEY.md section 8(c) clause (2).  Z keeps only the external roots.
        for (int w = tid; w < nw; w += kT) {
            uint32_t roots = Z[w];
            if (roots == 0) continue;
            uint32_t ext = 0;
            int y, wx;
            g.split(w, y, wx);
            while (roots) {
                const int x = __ffs((int)roots) - 1;
                roots &= roots - 1;
                const int Xp = wx * 32 + x;
                bool e = Xp == 0;
                if (!e) {
                    const int wi = y * wpitch + ((Xp - 1) >> 5);
                    const int run = run_id(row_heads(X, g, wi), (int)roff[wi], (Xp - 1) & 31);
                    e = ((uint32_t)L.area[L.lab[run]] & kFlagBit) != 0;
                }
                if (e) ext |= 1u << x;
            }
            Z[w] = ext;
        }
        __syncthreads();
        n_c = scan_words(nw, roff, [&](int w) { return __popc(Z[w]); });
        MS_FUSED_MARK(8);
    }
    // reserve the slice's contour records (starts / per-contour scratch / {offset, count} records)
    if (tid == 0) s_base[0] = (int)min(atomicAdd(&P.header[4], (unsigned long long)n_c), (unsigned long long)0x7FFFFFFF);
    __syncthreads();
    const int r0 = s_base[0];
    if (n_c == 0) {
        if (tid == 0) P.slice_info[b] = make_int4(0, 0, r0, 0);
        return;
    }
    if ((long long)r0 + n_c > P.cap_contours) {       // does not fit: the host grows the capacities and runs the batch again
        if (tid == 0) {
            P.slice_info[b] = make_int4(n_c, 0, -1, -1);
            atomicOr(&P.header[2], 1ull);
        }
        return;
    }
    // per-contour scratch (start pixel, position base, rotation, start slot, border length): shared memory for a normal
    // slice, the global arrays for one with more than kSC contours (mask2polygon-only stress masks)
    const bool cs = n_c <= kSC;
    int* starts = cs ? s_contour : P.starts + r0;
    for (int w = tid; w < nw; w += kT) {              // descending raster order (clause (3))
        uint32_t m = Z[w];
        if (m == 0) continue;
        int k = (int)roff[w];
        int y, wx;
        g.split(w, y, wx);
        while (m) {
            const int x = __ffs((int)m) - 1;
            m &= m - 1;
            starts[n_c - 1 - k] = y * W + wx * 32 + x;
            ++k;
        }
    }
    __syncthreads();      // Z and roff are overwritten below
    // ---------------------------------------------------------------- zero-framed bit image for the crack tables
    const int Pp = wpitch + 2;
    uint32_t* sbits = smem;
    for (int i = tid; i < (H + 2) * Pp; i += kT) {
        const int r = i / Pp, c = i - r * Pp;
        sbits[i] = (r >= 1 && r <= H && c >= 1 && c <= wpitch) ? Y[(r - 1) * wpitch + (c - 1)] : 0u;
    }
    __syncthreads();      // (also orders the `starts` writes before the reads below)
    uint16_t* pix_off = reinterpret_cast<uint16_t*>(sbits + (H + 2) * Pp);
    uint32_t* slots = reinterpret_cast<uint32_t*>(pix_off + ((nw + 1) & ~1));      // [4 * cap_border + 1]
    const int cap_border = P.cap_border;
    uint32_t* pos = slots + 4 * cap_border + 1;                                    // [4 * cap_border]
    uint32_t* slot_px = pos + 4 * cap_border;                                      // [cap_border]
    const size_t cstride = cs ? (size_t)kSC : (size_t)P.cap_contours + 1;
    int* c_base = cs ? s_contour + kSC : P.cinfo + r0;
    int* c_rot = c_base + cstride;
    int* c_slot = c_base + 2 * cstride;
    int* c_len = c_base + 3 * cstride;
    int2* rec = P.rec + r0;

    // border pixels per word -> exclusive offsets (u16: a slice with more than 65,535 border pixels takes the walk below,
    // which never reads them)
    const int n_border = scan_words(nw, pix_off, [&](int w) {
        int y, wx;
        g.split(w, y, wx);
        return __popc(srank::nbhd(sbits, Pp, y, wx).border);
    }, false);
    MS_FUSED_MARK(9);

    if (n_border > cap_border || n_c >= (int)(srank::kInvalidNext - srank::kMark)) {
        // ------------------------------------------------------------ does not fit the crack table: walk the borders, twice
        //   (count, reserve, emit) -- one thread per contour, contour_trace.cuh
        build_trace_lut(lut);
        __syncthreads();
        const PaddedBitsWindow bits{sbits, Pp};
        struct CountEmit { __device__ void operator()(int, int) {} };
        for (int c = tid; c < n_c; c += kT) {
            TraceState ts;
            trace_begin(ts, W, starts[c]);
            CountEmit ce;
            const int cnt = trace_run(bits, lut, W, ts, 8 * H * W + 8, ce, TraceAlwaysInside{}) == 1 ? ts.n : -1;
            if (cnt < 0) atomicAdd(&P.header[3], 1ull);
            c_len[c] = cnt < 0 ? 0 : cnt;
        }
        __syncthreads();
        int n_p = 0;
        for (int base = 0; base < n_c; base += kT) {
            const int c = base + tid;
            const int v = c < n_c ? c_len[c] : 0;
            int tot;
            const int ex = block_exscan_1024(v, &tot);
            if (c < n_c) rec[c] = make_int2(n_p + ex, v);
            n_p += tot;
        }
        if (tid == 0) s_base[1] = (int)min(atomicAdd(&P.header[5], (unsigned long long)n_p), (unsigned long long)0x7FFFFFFF);
        __syncthreads();
        const int v0 = s_base[1];
        const bool fits = (long long)v0 + n_p <= P.cap_points;
        if (tid == 0) {
            P.slice_info[b] = make_int4(n_c, n_p, r0, fits ? v0 : -1);
            if (!fits) atomicOr(&P.header[2], 2ull);
        }
        if (!fits) return;
        struct StoreEmit {
            uint32_t* dst;
            __device__ void operator()(int x, int y) { *dst++ = (uint32_t)x | ((uint32_t)y << 16); }
        };
        for (int c = tid; c < n_c; c += kT) {
            TraceState ts;
            trace_begin(ts, W, starts[c]);
            StoreEmit se{P.vstore + v0 + rec[c].x};
            trace_run(bits, lut, W, ts, 8 * H * W + 8, se, TraceAlwaysInside{});
        }
        return;
    }

    // ---------------------------------------------------------------- crack list ranking (see srank::rank_smem_kernel)
    const srank::Table T{sbits, pix_off, Pp, wpitch};
    const int n_slots = 4 * n_border;
    for (int w = tid; w < nw; w += kT) {              // init: every open side points at its successor
        int y, wx;
        g.split(w, y, wx);
        const crack::Nbhd n = srank::nbhd(sbits, Pp, y, wx);
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
                uint32_t v = srank::pack(srank::kInvalidNext, 0);
                if (crack::side_open(code, sd)) {
                    uint32_t nx;
                    if (crack::has(code, 5 + 2 * sd)) nx = T.slot(x + trace_dx((5 + 2 * sd) & 7), y + trace_dy((5 + 2 * sd) & 7), (sd + 3) & 3);
                    else if (crack::has(code, 6 + 2 * sd)) nx = T.slot(x + trace_dx((6 + 2 * sd) & 7), y + trace_dy((6 + 2 * sd) & 7), sd);
                    else nx = idx * 4u + (uint32_t)((sd + 1) & 3);
                    v = srank::pack(nx, 1);
                }
                slots[idx * 4 + sd] = v;
            }
            ++idx;
        }
    }
    __syncthreads();
    MS_FUSED_MARK(10);
    for (int c = tid; c < n_c; c += kT) {             // cut: the predecessor of each start crack ends its list
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
        uint32_t pr;
        if (crack::has(code, 3)) pr = T.slot(x - 1, y - 1, 1);
        else if (crack::has(code, 2)) pr = T.slot(x, y - 1, 0);
        else pr = (s0 & ~3u) | 3u;
        slots[pr] = srank::pack(srank::kMark + (uint32_t)c, 1);
    }
    __syncthreads();
    MS_FUSED_MARK(11);
    for (int round = 0; round < 20; ++round) {        // pointer jumping, in place
        for (int i = tid; i < n_slots; i += kT) {
            const uint32_t v = slots[i];
            const uint32_t nx = srank::next_of(v);
            if (nx < srank::kMark) {
                const uint32_t t = slots[nx];
                slots[i] = srank::pack(srank::next_of(t), srank::rank_of(v) + srank::rank_of(t));
            }
        }
        __syncthreads();
        int busy = 0;
        for (int c = tid; c < n_c; c += kT) {
            const uint32_t t = slots[c_slot[c]];
            if (srank::next_of(t) < srank::kMark || srank::rank_of(t) > (1u << (round + 1))) busy = 1;
        }
        if (!__syncthreads_or(busy)) break;
    }
    MS_FUSED_MARK(12);
    int n_pos = 0;                                    // border lengths -> position bases
    for (int base = 0; base < n_c; base += kT) {
        const int c = base + tid;
        int len = 0;
        if (c < n_c) {
            const uint32_t t = slots[c_slot[c]];
            if (srank::next_of(t) == srank::kMark + (uint32_t)c) len = (int)srank::rank_of(t);
            else atomicAdd(&P.header[3], 1ull);      // unranked border: cannot happen
            c_len[c] = len;
        }
        int tot;
        const int ex = block_exscan_1024(len, &tot);
        if (c < n_c) c_base[c] = n_pos + ex;
        n_pos += tot;
    }
    __syncthreads();
    MS_FUSED_MARK(13);
    for (int i = tid; i < n_slots; i += kT) {         // positions + the CHAIN_APPROX_SIMPLE decision per pixel visit
        const uint32_t v = slots[i];
        const uint32_t nx = srank::next_of(v);
        if (nx < srank::kMark || nx == srank::kInvalidNext) continue;
        const int c = (int)(nx - srank::kMark);
        const int len = c_len[c];
        int ps = len - (int)srank::rank_of(v) + c_rot[c];
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
    MS_FUSED_MARK(14);
    uint32_t* prefix = slots;                         // kept-vertex prefix over positions (the slot table is dead)
    int n_p;
    {
        const int ppt = (n_pos + kT - 1) / kT;
        const int i0 = min(n_pos, tid * ppt), i1 = min(n_pos, i0 + ppt);
        int s = 0;
        for (int i = i0; i < i1; ++i) s += (pos[i] & crack::kKept) ? 1 : 0;
        int tot;
        int ex = block_exscan_1024(s, &tot);
        for (int i = i0; i < i1; ++i) {
            prefix[i] = (uint32_t)ex;
            ex += (pos[i] & crack::kKept) ? 1 : 0;
        }
        if (tid == 0) prefix[n_pos] = (uint32_t)tot;
        n_p = tot;
    }
    if (tid == 0) s_base[1] = (int)min(atomicAdd(&P.header[5], (unsigned long long)n_p), (unsigned long long)0x7FFFFFFF);
    __syncthreads();
    const int v0 = s_base[1];
    const bool fits = (long long)v0 + n_p <= P.cap_points;
    if (tid == 0) {
        P.slice_info[b] = make_int4(n_c, n_p, r0, fits ? v0 : -1);
        if (!fits) atomicOr(&P.header[2], 2ull);
    }
    for (int c = tid; c < n_c; c += kT) {
        const int k0 = (int)prefix[c_base[c]];
        rec[c] = make_int2(k0, (int)prefix[c_base[c] + c_len[c]] - k0);
    }
    MS_FUSED_MARK(15);
    if (!fits) return;
    uint32_t* vs = P.vstore + v0;
    for (int i = tid; i < n_pos; i += kT) {
        const uint32_t e = pos[i];
        if (!(e & crack::kKept)) continue;
        const int p = (int)(e & ~crack::kKept);
        vs[prefix[i]] = (uint32_t)(p % W) | ((uint32_t)(p / W) << 16);
    }
    MS_FUSED_MARK(16);
}

// One CTA per slice: global offsets (slice order), contour offsets, mapped coordinates.  Writes the CSR set the C ABI
// returns: slice_start[batch + 1], cstart (P.npts) [n_contours + 1], xy [n_points], header {n_contours, n_points, flags}.
__global__ void __launch_bounds__(256) finalize_kernel(const int4* __restrict__ slice_info, const int2* __restrict__ rec,
                                                        const uint32_t* __restrict__ vstore, int batch, int cap_contours,
                                                        long long cap_points, double sx, double sy, int* __restrict__ slice_start,
                                                        int* __restrict__ cstart, int2* __restrict__ xy, long long* __restrict__ header) {
    __shared__ long long s_c[8], s_p[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    long long c_acc = 0, p_acc = 0;
    for (int i = tid; i < b; i += 256) {
        const int4 s = slice_info[i];
        c_acc += s.x;
        p_acc += s.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_acc += __shfl_xor_sync(0xFFFFFFFFu, c_acc, o);
        p_acc += __shfl_xor_sync(0xFFFFFFFFu, p_acc, o);
    }
    if ((tid & 31) == 0) { s_c[tid >> 5] = c_acc; s_p[tid >> 5] = p_acc; }
    __syncthreads();
    long long c_base = 0, p_base = 0;
    for (int i = 0; i < 8; ++i) { c_base += s_c[i]; p_base += s_p[i]; }
    const int4 me = slice_info[b];
    const unsigned long long* hu = reinterpret_cast<const unsigned long long*>(header);
    const long long want_c = (long long)hu[4], want_p = (long long)hu[5];
    const bool overflow = (hu[2] & 3ull) != 0 || want_c > cap_contours || want_p > cap_points;
    if (tid == 0) {
        slice_start[b] = (int)c_base;
        if (b == batch - 1) {
            slice_start[batch] = (int)(c_base + me.x);
            // totals: what was requested (exact for contours; a lower bound for vertices while contours overflow)
            header[0] = want_c;
            header[1] = want_p > p_base + me.y ? want_p : p_base + me.y;
            if (want_p > 0x7FFFFFFFll) header[2] |= 4;
            if (!overflow) cstart[c_base + me.x] = (int)(p_base + me.y);
        }
    }
    if (overflow || me.z < 0 || me.w < 0) return;     // the caller grows the buffers and runs the batch again
    for (int c = tid; c < me.x; c += 256) cstart[c_base + c] = (int)(p_base + rec[me.z + c].x);
    const uint32_t* vs = vstore + me.w;
    int2* dst = xy + p_base;
    for (int k = tid; k < me.y; k += 256) {
        const uint32_t v = vs[k];
        dst[k] = map_point(make_int2((int)(v & 0xFFFFu), (int)(v >> 16)), sx, sy);
    }
}

}  // namespace fused
