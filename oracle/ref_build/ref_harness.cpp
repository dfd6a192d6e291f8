// ref_harness.cpp -- ORACLE (test infrastructure only).  A C ABI over the REFERENCE'S OWN object code:
// /root/reference/src/{preprocess,postprocess,mask2polygon}.cpp compiled unmodified, from where they lie, against the
// OpenCV stub in oracle/ref_build/stubs (see its header for what is and is not the reference's) and the reference's
// vendored nlohmann/json.hpp.  Output: oracle/_ref/libref_pipeline.so (git-ignored, travels to the GPU box).
// Used by tests/test_ref_pin.py (live) and tests/golden/make_ref_golden.py (committed fixtures).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "preprocess.h"      // /root/reference/include
#include "mask2polygon.h"    // /root/reference/include (pulls the stub opencv2/opencv.hpp + nlohmann/json.hpp)

cv::Mat postprocess_mask(const cv::Mat& src);   // /root/reference/src/postprocess.cpp:47 (no header in the reference)

namespace {
std::vector<std::vector<cv::Point>> from_csr(const int32_t* xy, const int32_t* cstart, int nc) {
    std::vector<std::vector<cv::Point>> out((size_t)nc);
    for (int c = 0; c < nc; ++c)
        for (int i = cstart[c]; i < cstart[c + 1]; ++i) out[c].emplace_back(xy[2 * i], xy[2 * i + 1]);
    return out;
}
// returns the contour count; *n_pts the total; writes what fits
int to_csr(const std::vector<std::vector<cv::Point>>& cs, int32_t* xy, int64_t cap_pts, int32_t* cstart, int cap_c, int64_t* n_pts) {
    int64_t np = 0;
    int c = 0;
    for (const auto& v : cs) {
        if (c < cap_c) cstart[c] = (int32_t)np;
        for (const auto& p : v) {
            if (np < cap_pts) { xy[2 * np] = p.x; xy[2 * np + 1] = p.y; }
            ++np;
        }
        ++c;
    }
    if (c <= cap_c) cstart[c] = (int32_t)np;
    *n_pts = np;
    return c;
}
cv::Mat wrap_u8(const uint8_t* p, int h, int w) {
    cv::Mat m(h, w, CV_8UC1);
    std::memcpy(m.data(), p, (size_t)h * w);
    return m;
}
}  // namespace

extern "C" {

// Preprocess::preprocess_raw (src/preprocess.cpp:76-141): RAW file -> `_normalized.png` + `_original_sizes.json`
int ref_preprocess_raw(const char* raw_path, const char* png_path, const char* json_path, int w, int h) {
    return Preprocess::preprocess_raw(raw_path, png_path, json_path, w, h) ? 1 : 0;
}

// stub cv::imread(IMREAD_GRAYSCALE) of a file the stub cv::imwrite wrote
int ref_read_png_gray(const char* path, uint8_t* dst, int64_t cap, int* w, int* h) {
    cv::Mat m = cv::imread(path, cv::IMREAD_GRAYSCALE);
    if (m.empty()) return 0;
    *w = m.cols; *h = m.rows;
    if ((int64_t)m.rows * m.cols > cap) return 0;
    std::memcpy(dst, m.data(), (size_t)m.rows * m.cols);
    return 1;
}
int ref_read_png_bgr(const char* path, uint8_t* dst, int64_t cap, int* w, int* h) {
    cv::Mat m = cv::imread(path);
    if (m.empty()) return 0;
    *w = m.cols; *h = m.rows;
    if ((int64_t)m.rows * m.cols * 3 > cap) return 0;
    std::memcpy(dst, m.data(), (size_t)m.rows * m.cols * 3);
    return 1;
}
int ref_write_png_gray(const char* path, const uint8_t* src, int w, int h) { return cv::imwrite(path, wrap_u8(src, h, w)) ? 1 : 0; }

// ::postprocess_mask (src/postprocess.cpp:47-79, fill_holes_inside_foreground :13-44)
int ref_postprocess_mask(const uint8_t* in, uint8_t* out, int h, int w) {
    try {
        cv::Mat r = postprocess_mask(wrap_u8(in, h, w));
        std::memcpy(out, r.data(), (size_t)h * w);
        return 1;
    } catch (const std::exception&) {
        return 0;
    }
}

// Mask2Polygon::extract_contours (src/mask2polygon.cpp:29-36)
int ref_extract_contours(const uint8_t* mask, int h, int w, int32_t* xy, int64_t cap_pts, int32_t* cstart, int cap_c, int64_t* n_pts) {
    return to_csr(Mask2Polygon::extract_contours(wrap_u8(mask, h, w)), xy, cap_pts, cstart, cap_c, n_pts);
}

// Mask2Polygon::map_contour_points (src/mask2polygon.cpp:41-63); in place on the CSR points.
// (The function is defined in the reference's TU with external linkage but not declared in its header.)
}  // extern "C"
namespace Mask2Polygon {
std::vector<std::vector<cv::Point>> map_contour_points(const std::vector<std::vector<cv::Point>>& contours, double scale_x, double scale_y);
}
extern "C" {
void ref_map_contour_points(int32_t* xy, const int32_t* cstart, int nc, double scale_x, double scale_y) {
    const auto mapped = Mask2Polygon::map_contour_points(from_csr(xy, cstart, nc), scale_x, scale_y);
    int64_t i = 0;
    for (const auto& v : mapped)
        for (const auto& p : v) { xy[2 * i] = p.x; xy[2 * i + 1] = p.y; ++i; }
}

// Mask2Polygon::generate_json (src/mask2polygon.cpp:68-109): writes json_path
int ref_generate_json(const int32_t* xy, const int32_t* cstart, int nc, const char* json_path, const char* base_name, int ow, int oh) {
    try {
        Mask2Polygon::generate_json(from_csr(xy, cstart, nc), json_path, base_name, ow, oh);
        return 1;
    } catch (const std::exception&) {
        return 0;
    }
}

// Mask2Polygon::process_single_mask (src/mask2polygon.cpp:134-222): files -> files (swallows its own exceptions)
void ref_process_single_mask(const char* mask_path, const char* output_dir, const char* json_path, const char* original_png,
                             const char* base_name) {
    Mask2Polygon::process_single_mask(mask_path, output_dir, json_path, original_png, base_name);
}

}  // extern "C"
