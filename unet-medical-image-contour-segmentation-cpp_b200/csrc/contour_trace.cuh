// contour_trace.cuh -- border following + CHAIN_APPROX_SIMPLE on an 8-neighbour code image.
//
// Replaces the tracing half of cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
// (/root/reference/src/mask2polygon.cpp:34).  The closed form followed here is SURVEY.md section
// 8(c) clauses (4)-(7) (Suzuki-Abe border following as OpenCV performs it):
//   * the contour starts at the component's raster-first pixel;
//   * from the start, neighbours are probed clockwise on screen beginning after W
//     (NW, N, NE, E, SE, S, SW, W); the first foreground hit is the contour's LAST pixel L;
//   * then, from pixel p entered from direction d_prev (pointing at the previous pixel), directions
//     d_prev+1, d_prev+2, ... (mod 8, counter-clockwise on screen) are probed; the first hit q gives
//     d_out; the walk ends when q == start and p == L;
//   * CHAIN_APPROX_SIMPLE keeps p iff d_out(p) differs from d_out of the previous vertex (cyclic;
//     the vertex before the start is L, whose d_out points at the start).
// Direction codes: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (y grows downwards).
//
// One step is a dependent chain (position -> 3x3 window -> next direction -> position), so its depth is the walking
// speed.  "Assemble the direction-indexed code, rotate by d_prev, find first set, look up dx / dy" is therefore ONE
// table lookup indexed by {d_prev, raw 3x3 window bits} (4096 entries, built per CTA in shared memory).
//
// Everything here is __host__ __device__ so tests/hostsim can exercise this exact code on the CPU against cv2; the
// library itself only ever calls it from kernels.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MS_HD __host__ __device__ __forceinline__
#else
#define MS_HD inline
#endif

namespace ms {

MS_HD int trace_dx(int d) { return (int)((0x10FFF011u >> (4 * d)) & 0xF) - ((0x10FFF011u >> (4 * d)) & 0x8) * 2; }  // 1,1,0,-1,-1,-1,0,1
MS_HD int trace_dy(int d) { return (int)((0x1110FFF0u >> (4 * d)) & 0xF) - ((0x1110FFF0u >> (4 * d)) & 0x8) * 2; }  // 0,-1,-1,-1,0,1,1,1

MS_HD int trace_first_set(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// 8-neighbour foreground code from three 3-bit row windows (bit 0 = x-1, bit 1 = x, bit 2 = x+1): bit d <-> direction d
MS_HD unsigned code_from_rows(unsigned up, unsigned cu, unsigned dn) {
    return ((cu >> 2) & 1u) | (((up >> 2) & 1u) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) | ((cu & 1u) << 4) |
           ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | (((dn >> 2) & 1u) << 7);
}
// raw window: up3 | cu3 << 3 | dn3 << 6
MS_HD unsigned code_from_window(unsigned w9) { return code_from_rows(w9 & 7u, (w9 >> 3) & 7u, (w9 >> 6) & 7u); }

// Step table entry i = d_prev << 9 | window:  bits 0-2 d_out | bits 3-4 dx + 1 | bits 5-6 dy + 1 | bits 7-9 next d_prev.
// From pixel p entered from direction d_prev, directions d_prev+1, d_prev+2, ... are probed; the first hit is d_out.
constexpr int kTraceLutEntries = 8 * 512;
MS_HD uint16_t trace_lut_entry(int i) {
    const unsigned cc = code_from_window((unsigned)i & 511u), dp = (unsigned)i >> 9;
    if (cc == 0) return 0;
    const unsigned rot = ((cc | (cc << 8)) >> ((dp + 1) & 7)) & 0xFFu;   // bit k <-> direction d_prev + 1 + k
    const int d = (int)(dp + 1 + (unsigned)trace_first_set(rot)) & 7;
    return (uint16_t)((unsigned)d | ((unsigned)(trace_dx(d) + 1) << 3) | ((unsigned)(trace_dy(d) + 1) << 5) | ((unsigned)((d + 4) & 7) << 7));
}

// Resumable walk.  `win9(x, y)` returns the raw 3x3 window of pixel (x, y); `emit(x, y)` is called for every kept vertex,
// in order; `inside(x, y)` says whether `win9` can currently be evaluated there (always true for a whole-image source; a
// shared-memory window returns false when the walk leaves it, the caller then re-centres it and calls again with the
// same state).
struct TraceState {
    int start, last, p, x, y, d_prev, prev_out, n;
    int phase;   // 0 = not started, 1 = walking, 2 = finished
};
MS_HD void trace_begin(TraceState& s, int W, int start) {
    s.start = start; s.p = start; s.x = start % W; s.y = start / W; s.n = 0; s.phase = 0;
    s.last = 0; s.d_prev = 0; s.prev_out = 0;
}
struct TraceAlwaysInside {
    MS_HD bool operator()(int, int) const { return true; }
};
// returns 1 = finished (s.n kept vertices), 0 = paused because the current pixel is not `inside`, -1 = `max_steps` exhausted
template <class Window, class Emit, class Inside>
MS_HD int trace_run(Window win9, const uint16_t* lut, int W, TraceState& s, int max_steps, Emit& emit, Inside inside) {
    if (s.phase == 2) return 1;
    if (s.phase == 0) {
        if (!inside(s.x, s.y)) return 0;
        const unsigned c0 = code_from_window(win9(s.x, s.y));
        if (c0 == 0) {  // isolated pixel
            emit(s.x, s.y);
            s.n = 1;
            s.phase = 2;
            return 1;
        }
        unsigned rev = 0;   // probe NW, N, NE, E, SE, S, SW, W: bit k <-> direction (3 - k) & 7
        for (int k = 0; k < 8; ++k) rev |= ((c0 >> ((3 - k) & 7)) & 1u) << k;
        const int dL = (3 - trace_first_set(rev)) & 7;
        s.last = s.start + trace_dy(dL) * W + trace_dx(dL);
        s.d_prev = dL;
        s.prev_out = (dL + 4) & 7;
        s.phase = 1;
    }
    int x = s.x, y = s.y, p = s.p, prev_out = s.prev_out, n = s.n;
    unsigned dp9 = (unsigned)s.d_prev << 9;
    int status = -1;
    for (int step = 0; step < max_steps; ++step) {
        if (!inside(x, y)) { status = 0; break; }
        const unsigned e = lut[dp9 | win9(x, y)];
        const int d = (int)(e & 7u);
        if (d != prev_out) {   // CHAIN_APPROX_SIMPLE: a vertex is kept iff the outgoing direction changes
            emit(x, y);
            ++n;
            prev_out = d;
        }
        const int ddx = (int)((e >> 3) & 3u) - 1, ddy = (int)((e >> 5) & 3u) - 1;
        const int q = p + ddy * W + ddx;
        if (q == s.start && p == s.last) { status = 1; break; }
        p = q;
        x += ddx;
        y += ddy;
        dp9 = (e << 2) & (7u << 9);
    }
    s.x = x; s.y = y; s.p = p; s.prev_out = prev_out; s.n = n; s.d_prev = (int)(dp9 >> 9);
    if (status == 1) s.phase = 2;
    return status;
}

}  // namespace ms
