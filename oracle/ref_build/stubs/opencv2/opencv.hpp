// opencv2/opencv.hpp -- ORACLE STUB (test infrastructure only; never part of the product).
//
// A minimal stand-in for the OpenCV C++ SDK, which is not installed in this image, so that the reference's own
// translation units compile UNMODIFIED from where they lie:
//     /root/reference/src/preprocess.cpp, src/postprocess.cpp, src/mask2polygon.cpp
// (recipe: oracle/ref_build/Makefile -> oracle/_ref/libref_pipeline.so).  Everything the reference authored --
// the resample / normalise arithmetic, the hole / area thresholds and bbox test, the label loops, the coordinate
// mapping, the JSON document, the file protocol of process_single_mask -- is then the reference's object code.
//
// What is NOT the reference's: the OpenCV primitives below.  They are textbook implementations of the published
// definitions (containers, comparisons, 3x3 morphology with OpenCV's "border never wins" rule, flood-fill connected
// components with LEFT/TOP/WIDTH/HEIGHT/AREA stats, Suzuki-Abe border following for RETR_EXTERNAL +
// CHAIN_APPROX_SIMPLE, 8-connected polylines, a PNG codec restricted to stored deflate blocks), each checked against
// cv2 4.13.0 in tests/test_ref_pin.py.  The labelling / border-following loops are the plain-C oracle's
// (oracle/c/medseg_oracle.c: orc_ccl, orc_find_contours), linked into the same library.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32S 4
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(2, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_Assert(expr) \
    do { if (!(expr)) throw std::runtime_error(std::string("CV_Assert failed: ") + #expr); } while (0)

extern "C" {
int orc_ccl(const uint8_t* bin, int H, int W, int conn, int32_t* labels, int32_t* stats, int stats_cap);
int orc_find_contours(const uint8_t* mask, int H, int W, int thr, int32_t* xy, int64_t cap_pts, int32_t* cstart, int cap_c,
                      int64_t* n_pts);
}

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    double operator[](int i) const { return val[i]; }
};

template <class T>
struct Point_ {
    T x = 0, y = 0;
    Point_() {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;

enum { CC_STAT_LEFT = 0, CC_STAT_TOP = 1, CC_STAT_WIDTH = 2, CC_STAT_HEIGHT = 3, CC_STAT_AREA = 4 };
enum { THRESH_BINARY = 0 };
enum { RETR_EXTERNAL = 0 };
enum { CHAIN_APPROX_SIMPLE = 2 };
enum { MORPH_RECT = 0 };
enum { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2 };
enum { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { IMWRITE_PNG_COMPRESSION = 16 };

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type, const Scalar& v) {
        create(s.height, s.width, type);
        setTo(v);
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * elemSize(), 0);
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : depth() == 2 ? 2 : 4); }
    bool empty() const { return !buf_ || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    uchar* data() const { return buf_ ? buf_->data() : nullptr; }
    Mat clone() const {
        Mat m;
        m.rows = rows; m.cols = cols; m.type_ = type_;
        if (buf_) m.buf_ = std::make_shared<std::vector<uchar>>(*buf_);
        return m;
    }
    template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data() + (size_t)y * cols * elemSize()); }
    template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data() + (size_t)y * cols * elemSize()); }
    template <class T> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <class T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }
    // setTo(value[, mask]): every element (where mask != 0) becomes the saturate-cast scalar
    Mat& setTo(const Scalar& v, const Mat& mask = Mat()) {
        const bool use_mask = !mask.empty();
        if (use_mask) CV_Assert(mask.type() == CV_8UC1 && mask.rows == rows && mask.cols == cols);
        const int cn = channels();
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                if (use_mask && !mask.at<uchar>(y, x)) continue;
                for (int c = 0; c < cn; ++c) {
                    if (depth() == CV_8U) ptr<uchar>(y)[x * cn + c] = (uchar)std::min(255.0, std::max(0.0, std::round(v[c])));
                    else if (depth() == CV_32S) ptr<int>(y)[x * cn + c] = (int)v[c];
                    else throw std::runtime_error("stub Mat::setTo: unsupported depth");
                }
            }
        return *this;
    }

private:
    int type_ = 0;
    std::shared_ptr<std::vector<uchar>> buf_;
};

// cv::compare(src, value, dst, CMP_EQ): 255 where equal, else 0 (single channel)
inline Mat operator==(const Mat& a, double v) {
    CV_Assert(a.channels() == 1);
    Mat m(a.rows, a.cols, CV_8UC1);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) {
            const double e = a.depth() == CV_8U ? (double)a.at<uchar>(y, x) : (double)a.at<int>(y, x);
            m.at<uchar>(y, x) = e == v ? 255 : 0;
        }
    return m;
}

inline void bitwise_not(const Mat& src, Mat& dst) {
    CV_Assert(src.depth() == CV_8U);
    Mat out(src.rows, src.cols, src.type());
    const size_t n = (size_t)src.rows * src.cols * src.channels();
    for (size_t i = 0; i < n; ++i) out.data()[i] = (uchar)~src.data()[i];
    dst = out;
}

inline double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type) {
    CV_Assert(src.type() == CV_8UC1 && type == THRESH_BINARY);
    Mat out(src.rows, src.cols, CV_8UC1);
    const int t = (int)std::floor(thresh);          // 8-bit THRESH_BINARY: v > floor(thresh)
    const size_t n = (size_t)src.rows * src.cols;
    for (size_t i = 0; i < n; ++i) out.data()[i] = src.data()[i] > t ? (uchar)maxval : 0;
    dst = out;
    return thresh;
}

// labels CV_32S (0 = background), stats CV_32S [n][5] = LEFT, TOP, WIDTH, HEIGHT, AREA (row 0 = background),
// centroids left empty (the reference never reads them).  Returns the label count including background.
inline int connectedComponentsWithStats(const Mat& image, Mat& labels, Mat& stats, Mat& centroids, int connectivity = 8) {
    CV_Assert(image.type() == CV_8UC1);
    const int H = image.rows, W = image.cols;
    labels.create(H, W, CV_32SC1);
    std::vector<int32_t> st(((size_t)H * W + 2) * 5, 0);
    const int nc = orc_ccl(image.data(), H, W, connectivity, labels.ptr<int>(), st.data(), H * W + 2);
    stats.create(nc + 1, 5, CV_32SC1);
    int bl = W, bt = H, br = -1, bb = -1, ba = 0;      // background row, as OpenCV reports it
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (!labels.at<int>(y, x)) {
                bl = std::min(bl, x); bt = std::min(bt, y); br = std::max(br, x); bb = std::max(bb, y); ++ba;
            }
    int* s0 = stats.ptr<int>(0);
    s0[0] = ba ? bl : 0; s0[1] = ba ? bt : 0; s0[2] = ba ? br - bl + 1 : 0; s0[3] = ba ? bb - bt + 1 : 0; s0[4] = ba;
    for (int i = 1; i <= nc; ++i) {
        const int32_t* o = st.data() + (size_t)i * 5;      // {left, top, right, bottom, area}
        int* s = stats.ptr<int>(i);
        s[0] = o[0]; s[1] = o[1]; s[2] = o[2] - o[0] + 1; s[3] = o[3] - o[1] + 1; s[4] = o[4];
    }
    centroids = Mat();
    return nc + 1;
}

inline Mat getStructuringElement(int shape, Size ksize) {
    CV_Assert(shape == MORPH_RECT);
    return Mat(ksize, CV_8UC1, Scalar(1));
}

// erode / dilate with a rectangular kernel anchored at its centre; OpenCV's default border for morphology is a
// constant that never wins (+inf for erode, -inf for dilate), i.e. out-of-image taps are skipped
inline Mat morph_rect_(const Mat& src, const Mat& kernel, bool erode) {
    const int H = src.rows, W = src.cols, ry = kernel.rows / 2, rx = kernel.cols / 2;
    Mat out(H, W, CV_8UC1);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int v = erode ? 255 : 0;
            for (int dy = -ry; dy <= ry; ++dy)
                for (int dx = -rx; dx <= rx; ++dx) {
                    const int yy = y + dy, xx = x + dx;
                    if (yy < 0 || xx < 0 || yy >= H || xx >= W) continue;
                    const int s = src.at<uchar>(yy, xx);
                    v = erode ? std::min(v, s) : std::max(v, s);
                }
            out.at<uchar>(y, x) = (uchar)v;
        }
    return out;
}
inline void morphologyEx(const Mat& src, Mat& dst, int op, const Mat& kernel) {
    CV_Assert(src.type() == CV_8UC1 && op == MORPH_OPEN);
    dst = morph_rect_(morph_rect_(src, kernel, true), kernel, false);
}

inline void findContours(const Mat& image, std::vector<std::vector<Point>>& contours, int mode, int method) {
    CV_Assert(image.type() == CV_8UC1 && mode == RETR_EXTERNAL && method == CHAIN_APPROX_SIMPLE);
    const int H = image.rows, W = image.cols;
    const int64_t cap_pts = 4 * (int64_t)H * W + 16;
    const int cap_c = H * W + 1;
    std::vector<int32_t> xy((size_t)cap_pts * 2), cs((size_t)cap_c + 1);
    int64_t np = 0;
    const int nc = orc_find_contours(image.data(), H, W, 0 /* findContours: non-zero = foreground */, xy.data(), cap_pts, cs.data(), cap_c, &np);
    contours.clear();
    for (int c = 0; c < nc; ++c) {
        std::vector<Point> v;
        for (int i = cs[c]; i < cs[c + 1]; ++i) v.emplace_back(xy[2 * i], xy[2 * i + 1]);
        contours.push_back(std::move(v));
    }
}

// closed 1-px polylines, 8-connected (LINE_8); colour is BGR for 3-channel images
inline void drawContours(Mat& img, const std::vector<std::vector<Point>>& contours, int idx, const Scalar& color, int thickness = 1) {
    CV_Assert(img.depth() == CV_8U && thickness == 1);
    const int cn = img.channels();
    auto plot = [&](int x, int y) {
        if (x < 0 || y < 0 || x >= img.cols || y >= img.rows) return;
        for (int c = 0; c < cn; ++c) img.ptr<uchar>(y)[x * cn + c] = (uchar)color[c];
    };
    for (size_t k = 0; k < contours.size(); ++k) {
        if (idx >= 0 && (size_t)idx != k) continue;
        const auto& C = contours[k];
        for (size_t i = 0; i < C.size(); ++i) {
            const Point a = C[i], b = C[(i + 1) % C.size()];
            const int dx = std::abs(b.x - a.x), dy = std::abs(b.y - a.y), n = std::max(dx, dy);
            for (int t = 0; t <= n; ++t)      // axis-aligned / exact-diagonal runs are exact; general lines are rounded DDA
                plot(n ? a.x + (int)std::lround((double)(b.x - a.x) * t / n) : a.x, n ? a.y + (int)std::lround((double)(b.y - a.y) * t / n) : a.y);
        }
    }
}

// ---------------------------------------------------------------- PNG (8-bit gray / RGB, stored deflate blocks only)
namespace stub_png {
inline uint32_t crc32(const uchar* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}
inline void put32(std::vector<uchar>& v, uint32_t x) {
    for (int s = 24; s >= 0; s -= 8) v.push_back((uchar)(x >> s));
}
inline void chunk(std::vector<uchar>& out, const char* tag, const std::vector<uchar>& body) {
    put32(out, (uint32_t)body.size());
    std::vector<uchar> t(tag, tag + 4);
    t.insert(t.end(), body.begin(), body.end());
    out.insert(out.end(), t.begin(), t.end());
    put32(out, crc32(t.data(), t.size()));
}
}  // namespace stub_png

inline bool imwrite(const std::string& path, const Mat& img, const std::vector<int>& params = std::vector<int>()) {
    (void)params;
    if (img.empty() || img.depth() != CV_8U || (img.channels() != 1 && img.channels() != 3)) return false;
    const int cn = img.channels(), W = img.cols, H = img.rows;
    std::vector<uchar> raw;
    raw.reserve((size_t)H * (W * cn + 1));
    for (int y = 0; y < H; ++y) {
        raw.push_back(0);                                 // filter: None
        const uchar* r = img.ptr<uchar>(y);
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < cn; ++c) raw.push_back(r[x * cn + (cn == 3 ? 2 - c : c)]);   // BGR in memory -> RGB on disk
    }
    std::vector<uchar> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uchar v : raw) { a = (a + v) % 65521; b = (b + a) % 65521; }
    size_t pos = 0;
    do {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uchar)(n & 0xFF)); z.push_back((uchar)(n >> 8));
        z.push_back((uchar)(~n & 0xFF)); z.push_back((uchar)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        pos += n;
    } while (pos < raw.size());
    stub_png::put32(z, (b << 16) | a);
    std::vector<uchar> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    stub_png::put32(ihdr, (uint32_t)W); stub_png::put32(ihdr, (uint32_t)H);
    ihdr.push_back(8); ihdr.push_back(cn == 3 ? 2 : 0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    stub_png::chunk(out, "IHDR", ihdr);
    stub_png::chunk(out, "IDAT", z);
    stub_png::chunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    if (!f.good()) return false;
    f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
    return f.good();
}

// reads the PNGs imwrite above produces (stored blocks, filter None); anything else -> empty Mat, as cv::imread
// reports an unreadable file
inline Mat imread(const std::string& path, int flags = IMREAD_COLOR) {
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) return Mat();
    std::vector<uchar> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uchar sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8) != 0) return Mat();
    auto rd32 = [&](size_t p) { return ((uint32_t)d[p] << 24) | ((uint32_t)d[p + 1] << 16) | ((uint32_t)d[p + 2] << 8) | d[p + 3]; };
    int W = 0, H = 0, cn = 0;
    std::vector<uchar> z;
    for (size_t p = 8; p + 12 <= d.size();) {
        const uint32_t n = rd32(p);
        if (p + 12 + n > d.size()) return Mat();
        const std::string tag(d.begin() + p + 4, d.begin() + p + 8);
        if (tag == "IHDR") {
            W = (int)rd32(p + 8); H = (int)rd32(p + 12);
            if (d[p + 16] != 8 || (d[p + 17] != 0 && d[p + 17] != 2) || d[p + 20] != 0) return Mat();
            cn = d[p + 17] == 2 ? 3 : 1;
        } else if (tag == "IDAT") {
            z.insert(z.end(), d.begin() + p + 8, d.begin() + p + 8 + n);
        }
        p += 12 + n;
    }
    if (!W || !H || z.size() < 6) return Mat();
    std::vector<uchar> raw;
    size_t p = 2;
    for (bool last = false; !last;) {
        if (p + 5 > z.size() || (z[p] & 6) != 0) return Mat();      // BTYPE must be 00 (stored)
        last = z[p] & 1;
        const size_t n = z[p + 1] | ((size_t)z[p + 2] << 8);
        p += 5;
        if (p + n > z.size()) return Mat();
        raw.insert(raw.end(), z.begin() + p, z.begin() + p + n);
        p += n;
    }
    if (raw.size() != (size_t)H * (W * cn + 1)) return Mat();
    const bool want_gray = flags == IMREAD_GRAYSCALE || (flags == IMREAD_UNCHANGED && cn == 1);
    Mat out(H, W, want_gray ? CV_8UC1 : CV_8UC3);
    for (int y = 0; y < H; ++y) {
        const uchar* r = raw.data() + (size_t)y * (W * cn + 1);
        if (r[0] != 0) return Mat();
        for (int x = 0; x < W; ++x) {
            if (want_gray) {
                // RGB -> gray only ever needed for equal channels here; use OpenCV's fixed-point weights anyway
                out.at<uchar>(y, x) = cn == 1 ? r[1 + x] : (uchar)((r[1 + 3 * x] * 4899 + r[2 + 3 * x] * 9617 + r[3 + 3 * x] * 1868 + 8192) >> 14);
            } else {
                uchar* o = out.ptr<uchar>(y) + 3 * x;
                if (cn == 1) o[0] = o[1] = o[2] = r[1 + x];
                else { o[0] = r[3 + 3 * x]; o[1] = r[2 + 3 * x]; o[2] = r[1 + 3 * x]; }
            }
        }
    }
    return out;
}

}  // namespace cv
