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
// `nb[p]` holds bit d set iff the neighbour of p in direction d is foreground (pixels outside the
// image count as background), so one byte load per step replaces eight probes.
//
// The function is __host__ __device__ so tests/hostsim can exercise this exact code on the CPU
// against cv2; the library itself only ever calls it from kernels.
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

MS_HD uint8_t trace_load(const uint8_t* nb, int i) {
#if defined(__CUDA_ARCH__)
    return __ldg(nb + i);
#else
    return nb[i];
#endif
}

// Resumable walk.  `code(p, x, y)` returns the 8-neighbour foreground code of pixel p = y * W + x; `emit(x, y)` is called
// for every kept vertex, in order; `inside(x, y)` says whether `code` can currently be evaluated at that pixel (always
// true for a whole-image code source; a shared-memory window returns false when the walk leaves it, the caller then
// re-centres the window and calls trace_run again with the same state).
struct TraceState {
    int start, last, p, x, y, d_prev, prev_out, n;
    int phase;   // 0 = not started, 1 = walking, 2 = finished
};
MS_HD void trace_begin(TraceState& s, int W, int start) {
    s.start = start; s.p = start; s.x = start % W; s.y = start / W; s.n = 0; s.phase = 0;
    s.last = 0; s.d_prev = 0; s.prev_out = 0;
}
// returns 1 = finished (s.n kept vertices), 0 = paused because the current pixel is not `inside`, -1 = `max_steps` exhausted
template <class Code, class Emit, class Inside>
MS_HD int trace_run(Code code, int W, TraceState& s, int max_steps, Emit& emit, Inside inside) {
    if (s.phase == 2) return 1;
    if (s.phase == 0) {
        if (!inside(s.x, s.y)) return 0;
        const unsigned c0 = code(s.p, s.x, s.y);
        if (c0 == 0) {  // isolated pixel
            emit(s.x, s.y);
            s.n = 1;
            s.phase = 2;
            return 1;
        }
        // probe 3,2,1,0,7,6,5,4: reverse the byte so that bit k <-> direction (3 - k) & 7
        unsigned rev = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) rev |= ((c0 >> ((3 - k) & 7)) & 1u) << k;
        const int dL = (3 - trace_first_set(rev)) & 7;
        s.last = s.start + trace_dy(dL) * W + trace_dx(dL);
        s.d_prev = dL;
        s.prev_out = (dL + 4) & 7;
        s.phase = 1;
    }
    for (int step = 0; step < max_steps; ++step) {
        if (!inside(s.x, s.y)) return 0;
        const unsigned cc = code(s.p, s.x, s.y);
        const unsigned rot = ((cc | (cc << 8)) >> ((s.d_prev + 1) & 7)) & 0xFFu;  // bit k <-> direction d_prev+1+k
        const int d = (s.d_prev + 1 + trace_first_set(rot)) & 7;                  // rot != 0: the way back is always set
        if (d != s.prev_out) {
            emit(s.x, s.y);
            ++s.n;
            s.prev_out = d;
        }
        const int ddx = trace_dx(d), ddy = trace_dy(d);
        const int q = s.p + ddy * W + ddx;
        if (q == s.start && s.p == s.last) {
            s.phase = 2;
            return 1;
        }
        s.p = q;
        s.x += ddx;
        s.y += ddy;
        s.d_prev = (d + 4) & 7;
    }
    return -1;
}

struct TraceAlwaysInside {
    MS_HD bool operator()(int, int) const { return true; }
};

// Walks one outer border in one go.  Returns the number of kept vertices, or -1 if `max_steps` was exhausted.
template <class Code, class Emit>
MS_HD int trace_contour_fn(Code code, int W, int start, int max_steps, Emit emit) {
    TraceState s;
    trace_begin(s, W, start);
    const int r = trace_run(code, W, s, max_steps, emit, TraceAlwaysInside{});
    return r == 1 ? s.n : -1;
}

struct NbImageCode {   // neighbour codes precomputed per pixel in global memory
    const uint8_t* nb;
    MS_HD unsigned operator()(int p, int, int) const { return trace_load(nb, p); }
};
template <class Emit>
MS_HD int trace_contour(const uint8_t* nb, int W, int start, int max_steps, Emit emit) {
    return trace_contour_fn(NbImageCode{nb}, W, start, max_steps, emit);
}

}  // namespace ms
