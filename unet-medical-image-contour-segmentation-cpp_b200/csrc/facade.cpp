// facade.cpp -- the reference's C++ stage API (namespaces MedicalSeg / Preprocess / Mask2Polygon)
// as a thin layer over the C ABI in include/medseg_b200.h, so /root/reference/src/main.cpp builds
// and links against this library unchanged.  Error convention of the reference is kept: bool
// results, messages on std::cerr and in the log file, nothing throws across the API
// (src/process.cpp:256-261, src/initialize.cpp:69-75).
#include "../../include/medseg_b200.h"
#include "../../include/initialize.h"
#include "../../include/process.h"
#include "../../include/cleanup.h"
#include "../../include/preprocess.h"
#include "../../include/postprocess.h"
#include "../../include/mask2polygon.h"
#include "json_min.hpp"
#include "png_min.hpp"
#include "overlay.hpp"

#include <cstdio>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>

namespace {
ms_handle* g_handle = nullptr;       // replaces g_runtime / g_engine + the thread_local context
std::ofstream g_log_file;            // src/initialize.cpp:22
std::string g_log_path;              // src/initialize.cpp:23

ms_handle* stage_handle() {          // stage-only handle for callers that never ran initialize_engine
    if (!g_handle) {
        if (ms_init(nullptr, nullptr, &g_handle) != MS_OK) {
            std::cerr << "medseg_b200: " << ms_last_error(nullptr) << std::endl;
            g_handle = nullptr;
        }
    }
    return g_handle;
}
std::string file_name(const std::string& p) {
    const size_t s = p.find_last_of("/\\");
    return s == std::string::npos ? p : p.substr(s + 1);
}
struct Csr {
    std::vector<int32_t> xy, cstart;
};
Csr to_csr(const std::vector<std::vector<Mask2Polygon::Point>>& contours) {
    Csr c;
    c.cstart.push_back(0);
    for (const auto& ct : contours) {
        for (const auto& p : ct) { c.xy.push_back(p.x); c.xy.push_back(p.y); }
        c.cstart.push_back((int32_t)(c.xy.size() / 2));
    }
    return c;
}
}  // namespace

namespace MedicalSeg {

bool initialize_engine(const std::string& engine_path, const std::string& log_dir) {
    if (g_handle) {
        ms_destroy(g_handle);
        g_handle = nullptr;
    }
    if (g_log_file.is_open()) g_log_file.close();
    const int rc = ms_init(engine_path.c_str(), log_dir.c_str(), &g_handle);
    if (rc != MS_OK) {
        std::cerr << "Initialization error: " << ms_last_error(nullptr) << std::endl;   // src/initialize.cpp:70
        g_handle = nullptr;
        return false;
    }
    g_log_path = log_dir + "/segmentation_log.txt";                                    // src/initialize.cpp:30
    g_log_file.open(g_log_path, std::ios::out | std::ios::app);                         // the handle created / truncated it
    if (!g_log_file.is_open()) {
        std::cerr << "Failed to create log file: " << g_log_path << std::endl;         // :34
        return false;
    }
    g_log_file << "Engine initialized successfully" << std::endl;                      // :67
    return true;
}

std::ofstream& get_log_file() { return g_log_file; }
std::string get_log_path() { return g_log_path; }
ms_handle* get_handle() { return g_handle; }

bool process_single_image(const std::string& raw_path, int width, int height, const std::string& output_dir) {
    if (!g_handle) {
        std::cerr << "Processing error: Engine not initialized" << std::endl;          // src/process.cpp:195,257
        return false;
    }
    const int rc = ms_process_raw_file(g_handle, raw_path.c_str(), width, height, output_dir.c_str());
    if (rc != MS_OK) {
        std::cerr << "Processing error: " << ms_last_error(g_handle) << std::endl;     // :257 (the handle logs it too, :259)
        return false;
    }
    return true;
}

int process_image_batch(const std::vector<std::string>& raw_paths, int width, int height, const std::vector<std::string>& output_dirs,
                        std::vector<bool>* ok) {
    if (ok) ok->assign(raw_paths.size(), false);
    if (!g_handle) {
        std::cerr << "Processing error: Engine not initialized" << std::endl;
        return 0;
    }
    if (output_dirs.size() != raw_paths.size()) {
        std::cerr << "Processing error: one output directory per file is required" << std::endl;
        return 0;
    }
    std::vector<const char*> a, b;
    for (size_t i = 0; i < raw_paths.size(); ++i) {
        a.push_back(raw_paths[i].c_str());
        b.push_back(output_dirs[i].c_str());
    }
    std::vector<uint8_t> flags(raw_paths.size() + 1, 0);
    int64_t good = 0, bad = 0;
    const int rc = ms_process_raw_files(g_handle, a.data(), b.data(), (int64_t)raw_paths.size(), width, height, flags.data(), &good, &bad);
    if (rc != MS_OK) {
        std::cerr << "Processing error: " << ms_last_error(g_handle) << std::endl;
        return 0;
    }
    if (ok)
        for (size_t i = 0; i < raw_paths.size(); ++i) (*ok)[i] = flags[i] != 0;
    return (int)good;
}

bool process_directory(const std::string& input_dir, int width, int height, const std::string& output_dir, bool recursive,
                       int* success_count, int* fail_count) {
    if (success_count) *success_count = 0;
    if (fail_count) *fail_count = 0;
    if (!g_handle) {
        std::cerr << "Processing error: Engine not initialized" << std::endl;
        return false;
    }
    std::cout << "Processing directory: " << input_dir << std::endl;                   // src/main.cpp:135
    std::cout << "Recursive: " << (recursive ? "Yes" : "No") << std::endl;             // :136
    int64_t found = 0, good = 0, bad = 0;
    const int rc = ms_process_directory(g_handle, input_dir.c_str(), width, height, output_dir.c_str(), recursive ? 1 : 0, 0, 1, &found,
                                        &good, &bad);
    if (rc != MS_OK) {
        std::cerr << "Directory error: " << ms_last_error(g_handle) << std::endl;      // :44
        return false;
    }
    if (found == 0) {
        std::cerr << "No 16-bit images found in directory" << std::endl;               // :140
        return false;
    }
    std::cout << "\nDirectory processing completed:" << std::endl;                    // :166-168
    std::cout << "  Success: " << good << " files" << std::endl;
    std::cout << "  Failed: " << bad << " files" << std::endl;
    if (success_count) *success_count = (int)good;
    if (fail_count) *fail_count = (int)bad;
    return true;
}

void cleanup_resources() {
    if (g_handle) {
        ms_destroy(g_handle);                                                          // src/cleanup.cpp:16-45
        g_handle = nullptr;
    }
    if (g_log_file.is_open()) g_log_file.close();                                      // :53
    std::cout << "Resources cleaned up successfully" << std::endl;                     // :56
}

std::vector<uint8_t> postprocess_mask(const MaskView& src) {
    std::vector<uint8_t> out((size_t)src.rows * src.cols);
    ms_handle* h = stage_handle();
    if (!h || ms_postprocess_host(h, src.data, out.data(), src.rows, src.cols, 1, 0) != MS_OK) {
        std::cerr << "postprocess_mask error: " << ms_last_error(h) << std::endl;
        return {};
    }
    return out;
}

std::vector<uint8_t> mask_to_image(const MaskView& mask) {
    std::vector<uint8_t> out((size_t)mask.rows * mask.cols);
    for (size_t i = 0; i < out.size(); ++i) out[i] = mask.data[i] == 1 ? 128 : (mask.data[i] == 2 ? 255 : 0);  // src/process.cpp:180-183
    return out;
}

}  // namespace MedicalSeg

namespace Preprocess {

bool preprocess_raw(const std::string& raw_path, const std::string& png_path, const std::string& json_path, int w, int h) {
    ms_handle* hd = stage_handle();
    if (!hd || w <= 0 || h <= 0) {
        std::cerr << "preprocess_raw error: " << (hd ? "bad size" : ms_last_error(nullptr)) << '\n';
        return false;
    }
    std::vector<uint16_t> src((size_t)w * h);
    FILE* f = std::fopen(raw_path.c_str(), "rb");
    if (!f) {
        std::cerr << "preprocess_raw error: open failed" << '\n';                       // src/preprocess.cpp:39,138
        return false;
    }
    const size_t got = std::fread(src.data(), 2, src.size(), f);
    std::fclose(f);
    ms_info info;
    ms_get_info(hd, &info);
    std::vector<uint8_t> dst((size_t)info.net_w * info.net_h);
    if (got != src.size() || ms_preprocess_host(hd, src.data(), w, h, 1, dst.data()) != MS_OK) {
        std::cerr << "preprocess_raw error: " << (got != src.size() ? "short read" : ms_last_error(hd)) << '\n';
        return false;
    }
    {
        std::error_code ec;
        std::filesystem::create_directories(std::filesystem::path(png_path).parent_path(), ec);   // :121
    }
    if (!ms::png::write_file(png_path, dst.data(), info.net_w, info.net_h, 1)) {       // :122
        std::cerr << "preprocess_raw error: imwrite failed" << '\n';
        return false;
    }
    std::ofstream jf(json_path, std::ios::binary);                                      // :133
    jf << ms::json::sidecar_text(file_name(raw_path), w, h, info.net_w, info.net_h);    // :126-134
    return jf.good();
}

}  // namespace Preprocess

namespace Mask2Polygon {

SizeInfo load_size_json(const std::string& json_path, const std::string& base_name) {
    std::ifstream f(json_path, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Fail to Open JSON File: " + json_path);  // src/mask2polygon.cpp:18-20
    std::stringstream ss;
    ss << f.rdbuf();
    const ms::json::Value j = ms::json::parse(ss.str());
    SizeInfo s;
    if (j.has(base_name + ".raw")) s.filename = base_name + ".raw";                      // :147
    else if (j.has(base_name + ".tif")) s.filename = base_name + ".tif";                 // :150
    else throw std::runtime_error("Cannot Find Size Info in JSON: " + base_name + ".raw/.tif");  // :154
    const ms::json::Value& e = j.at(s.filename);
    s.original_width = (int)e.at("original_width").num;                                  // :157-160
    s.original_height = (int)e.at("original_height").num;
    s.scaled_width = (int)e.at("scaled_width").num;
    s.scaled_height = (int)e.at("scaled_height").num;
    return s;
}

std::vector<std::vector<Point>> extract_contours(const MaskView& mask) {
    std::vector<std::vector<Point>> out;
    ms_handle* h = stage_handle();
    if (!h) return out;
    std::vector<int32_t> xy(2 * 4096), cstart(1025), sl(2);
    for (int attempt = 0; attempt < 2; ++attempt) {
        ms_polygons pg{xy.data(), (int64_t)xy.size() / 2, cstart.data(), (int64_t)cstart.size() - 1, sl.data(), 0, 0};
        const int rc = ms_mask2polygon_host(h, mask.data, mask.rows, mask.cols, 1, 127, mask.cols, mask.rows, &pg);
        if (rc == MS_ERR_CAPACITY && attempt == 0) {
            xy.resize((size_t)pg.n_points * 2 + 2);
            cstart.resize((size_t)pg.n_contours + 1);
            continue;
        }
        if (rc != MS_OK) {
            std::cerr << "extract_contours error: " << ms_last_error(h) << std::endl;
            return out;
        }
        out.resize((size_t)pg.n_contours);
        for (int64_t c = 0; c < pg.n_contours; ++c)
            for (int i = cstart[c]; i < cstart[c + 1]; ++i) out[c].push_back(Point{xy[2 * i], xy[2 * i + 1]});
        return out;
    }
    return out;
}

std::vector<std::vector<Point>> map_contour_points(const std::vector<std::vector<Point>>& contours, double scale_x, double scale_y) {
    std::vector<std::vector<Point>> mapped;
    mapped.reserve(contours.size());
    for (const auto& c : contours) {
        std::vector<Point> m;
        m.reserve(c.size());
        for (const auto& p : c) m.push_back(Point{static_cast<int>(p.x * scale_x), static_cast<int>(p.y * scale_y)});  // :54-55
        mapped.push_back(m);
    }
    return mapped;
}

void generate_json(const std::vector<std::vector<Point>>& contours, const std::string& json_path, const std::string& base_name,
                   int original_width, int original_height) {
    const Csr c = to_csr(contours);
    std::ofstream f(json_path, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Fail to Create JSON File: " + json_path);  // src/mask2polygon.cpp:105-107
    f << ms::json::labelme_text(c.xy.data(), c.cstart.data(), (int)contours.size(), base_name, original_width, original_height);
}

void create_overlay_image(const std::vector<std::vector<Point>>& contours, const MaskView& gray, const std::string& overlay_path) {
    const Csr c = to_csr(contours);
    std::vector<uint8_t> rgb((size_t)gray.rows * gray.cols * 3);
    for (size_t i = 0; i < (size_t)gray.rows * gray.cols; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = gray.data[i];
    ms::draw_contours_red(rgb, gray.cols, gray.rows, c.xy.data(), c.cstart.data(), (int)contours.size());
    if (!ms::png::write_file(overlay_path, rgb.data(), gray.cols, gray.rows, 3))
        throw std::runtime_error("Fail to Save Overlay PNG: " + overlay_path);             // src/mask2polygon.cpp:126-128
}

void process_single_mask(const std::string& mask_path, const std::string& output_dir, const std::string& json_path,
                         const std::string& original_png, const std::string& base_name) {
    try {
        std::cout << "Processing Mask: " << base_name + ".png" << std::endl;                       // :141
        const SizeInfo sz = load_size_json(json_path, base_name);                                   // :144-160
        std::cout << "Original Size: " << sz.original_width << "x" << sz.original_height << std::endl;
        std::cout << "Scaled Size: " << sz.scaled_width << "x" << sz.scaled_height << std::endl;
        ms::png::Image mimg;
        if (!ms::png::read_file(mask_path, mimg)) throw std::runtime_error("Fail to Read Mask File: " + mask_path);   // :166-169
        if (mimg.w != sz.scaled_width || mimg.h != sz.scaled_height)                                 // :172-179
            throw std::runtime_error("Mask size mismatch: " + std::to_string(mimg.w) + "x" + std::to_string(mimg.h) + " (actual) vs " +
                                     std::to_string(sz.scaled_width) + "x" + std::to_string(sz.scaled_height) + " (JSON)");
        const std::vector<uint8_t> grey = ms::png::to_grey(mimg);
        const auto contours = extract_contours(MaskView{grey.data(), mimg.h, mimg.w});              // :182
        if (contours.empty()) {
            std::cout << "Warning: No Contours Detected" << std::endl;                              // :184-185
            return;
        }
        std::cout << "Extracted " << contours.size() << " Contours" << std::endl;
        if (!original_png.empty()) {                                                                // :189-196
            ms::png::Image oimg;
            if (!ms::png::read_file(original_png, oimg)) throw std::runtime_error("Fail to Read Original Image: " + original_png);
            const std::vector<uint8_t> og = ms::png::to_grey(oimg);
            const std::string overlay_path = output_dir + "/" + base_name + "_contour_overlay.png";
            create_overlay_image(contours, MaskView{og.data(), oimg.h, oimg.w}, overlay_path);
            std::cout << "Overlay Image Saved to: " << overlay_path << std::endl;
        } else {
            std::cout << "Warning: Original PNG not provided, skipping overlay generation" << std::endl;
        }
        const double scale_x = static_cast<double>(sz.original_width) / sz.scaled_width;            // :199-200
        const double scale_y = static_cast<double>(sz.original_height) / sz.scaled_height;
        const auto mapped = map_contour_points(contours, scale_x, scale_y);
        const std::string out_json = output_dir + "/" + base_name + ".json";                        // :206
        generate_json(mapped, out_json, base_name, sz.original_width, sz.original_height);
        std::cout << "JSON Saved to: " << out_json << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "Processing Failure: " << e.what() << std::endl;                               // :219-221
    }
}

}  // namespace Mask2Polygon
