// trace_sim.cpp -- TEST-ONLY host build of csrc/contour_trace.cuh.
//
// Compiles the exact __host__ __device__ border-following code the CUDA kernel runs, drives it with
// a sequential stand-in for the labelling kernels, and exposes it over a C ABI so
// tests/test_hostsim_trace.py can compare it with cv2 on the CPU.  This is NOT a product path: the
// library never executes contour_trace.cuh on the host.
#include <cstdint>
#include <cstring>
#include <vector>
#include <numeric>
#include "../../unet-medical-image-contour-segmentation-cpp_b200/csrc/contour_trace.cuh"

namespace {
struct UF {
    std::vector<int> p;
    explicit UF(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int a) { while (p[a] != a) { p[a] = p[p[a]]; a = p[a]; } return a; }
    void unite(int a, int b) { a = find(a); b = find(b); if (a < b) p[b] = a; else p[a] = b; }
};
}

extern "C" int sim_find_contours(const uint8_t* mask, int H, int W, int thr, int32_t* xy, int64_t cap_pts,
                                 int32_t* cstart, int cap_c, int64_t* n_pts) {
    const size_t n = (size_t)H * W;
    std::vector<uint8_t> fg(n);
    for (size_t i = 0; i < n; ++i) fg[i] = mask[i] > thr;
    auto F = [&](int x, int y) { return x >= 0 && y >= 0 && x < W && y < H && fg[(size_t)y * W + x]; };
    // the step table exactly as a CTA builds it, and a raw 3x3 window source over the byte image
    std::vector<uint16_t> lut(ms::kTraceLutEntries);
    for (int i = 0; i < ms::kTraceLutEntries; ++i) lut[i] = ms::trace_lut_entry(i);
    auto win9 = [&](int x, int y) {
        unsigned w = 0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) w |= (unsigned)F(x - 1 + c, y - 1 + r) << (3 * r + c);
        return w;
    };
    UF f8(n), b4(n);
    std::vector<uint8_t> outer(n, 0);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int p = y * W + x;
            if (fg[p]) {
                if (F(x - 1, y)) f8.unite(p, p - 1);
                if (F(x - 1, y - 1)) f8.unite(p, p - W - 1);
                if (F(x, y - 1)) f8.unite(p, p - W);
                if (F(x + 1, y - 1)) f8.unite(p, p - W + 1);
            } else {
                if (x > 0 && !fg[p - 1]) b4.unite(p, p - 1);
                if (y > 0 && !fg[p - W]) b4.unite(p, p - W);
            }
        }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int p = y * W + x;
            if (!fg[p] && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) outer[b4.find(p)] = 1;
        }
    int nout = 0;
    int64_t np = 0;
    for (int64_t p = (int64_t)n - 1; p >= 0; --p) {
        if (!fg[p] || f8.find((int)p) != p) continue;
        int x = (int)(p % W);
        if (!(x == 0 || outer[b4.find((int)p - 1)])) continue;
        if (nout < cap_c) cstart[nout] = (int32_t)np;
        auto emit = [&](int px, int py) {
            if (np < cap_pts) { xy[2 * np] = px; xy[2 * np + 1] = py; }
            ++np;
        };
        ms::TraceState st;
        ms::trace_begin(st, W, (int)p);
        if (ms::trace_run(win9, lut.data(), W, st, (int)(8 * n + 8), emit, ms::TraceAlwaysInside{}) != 1) return -1;
        ++nout;
    }
    if (nout <= cap_c) cstart[nout] = (int32_t)np;
    *n_pts = np;
    return nout;
}
