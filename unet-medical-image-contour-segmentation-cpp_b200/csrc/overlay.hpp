// overlay.hpp -- 1-px closed polylines on an RGB image, as cv::drawContours(img, contours, -1,
// Scalar(0,0,255) /*BGR red*/, 1) does at /root/reference/src/mask2polygon.cpp:123.  Consecutive
// CHAIN_APPROX_SIMPLE vertices are always joined by a horizontal, vertical or exact-diagonal run, for
// which 8-connected Bresenham (cv::LINE_8) is unambiguous.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace ms {

// contour c = xy[2*cstart[c] .. 2*cstart[c+1]) ; rgb is w*h*3 (R,G,B)
inline void draw_contours_red(std::vector<uint8_t>& rgb, int w, int h, const int32_t* xy, const int32_t* cstart, int n_contours) {
    auto plot = [&](int x, int y) {
        if (x < 0 || y < 0 || x >= w || y >= h) return;
        uint8_t* p = &rgb[((size_t)y * w + x) * 3];
        p[0] = 255; p[1] = 0; p[2] = 0;
    };
    for (int c = 0; c < n_contours; ++c) {
        const int a = cstart[c], b = cstart[c + 1];
        for (int i = a; i < b; ++i) {
            const int j = i + 1 < b ? i + 1 : a;  // closed
            int x0 = xy[2 * i], y0 = xy[2 * i + 1];
            const int x1 = xy[2 * j], y1 = xy[2 * j + 1];
            const int sx = (x1 > x0) - (x1 < x0), sy = (y1 > y0) - (y1 < y0);
            const int dx = std::abs(x1 - x0), dy = std::abs(y1 - y0);
            int err = dx - dy;
            for (;;) {
                plot(x0, y0);
                if (x0 == x1 && y0 == y1) break;
                const int e2 = 2 * err;
                if (e2 > -dy) { err -= dy; x0 += sx; }
                if (e2 < dx) { err += dx; y0 += sy; }
            }
        }
    }
}

// The same polylines as a bitmap (1 bit per pixel, row pitch (w + 31) / 32 words): the overlay PNG is then produced row
// by row from the grey image plus this 32 KiB map, without ever materialising the 768 KiB RGB image.
inline std::vector<uint32_t> contour_bitmap(int w, int h, const int32_t* xy, const int32_t* cstart, int n_contours) {
    const int pitch = (w + 31) / 32;
    std::vector<uint32_t> bits((size_t)pitch * h, 0u);
    auto plot = [&](int x, int y) {
        if (x < 0 || y < 0 || x >= w || y >= h) return;
        bits[(size_t)y * pitch + (x >> 5)] |= 1u << (x & 31);
    };
    for (int c = 0; c < n_contours; ++c) {
        const int a = cstart[c], b = cstart[c + 1];
        for (int i = a; i < b; ++i) {
            const int j = i + 1 < b ? i + 1 : a;  // closed
            int x0 = xy[2 * i], y0 = xy[2 * i + 1];
            const int x1 = xy[2 * j], y1 = xy[2 * j + 1];
            const int sx = (x1 > x0) - (x1 < x0), sy = (y1 > y0) - (y1 < y0);
            const int dx = std::abs(x1 - x0), dy = std::abs(y1 - y0);
            int err = dx - dy;
            for (;;) {
                plot(x0, y0);
                if (x0 == x1 && y0 == y1) break;
                const int e2 = 2 * err;
                if (e2 > -dy) { err -= dy; x0 += sx; }
                if (e2 < dx) { err += dx; y0 += sy; }
            }
        }
    }
    return bits;
}

}  // namespace ms
