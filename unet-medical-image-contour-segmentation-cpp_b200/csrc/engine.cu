// engine.cu -- the C ABI of libmedseg_b200.so (include/medseg_b200.h): handle, config, stage entry
// points, the in-memory per-batch pipeline and the per-file artefact writer.
//
// Host-side structure replaced: MedicalSeg::initialize_engine / TensorRTContext / execute_inference /
// process_single_image / cleanup_resources (/root/reference/src/initialize.cpp, src/process.cpp,
// src/cleanup.cpp).  Differences by design: state lives in a
// handle instead of process globals + a thread_local context, every CUDA return code is checked,
// intermediates stay in HBM instead of round-tripping through PNG/JSON files, and the batch is a
// launch parameter instead of a hard-wired 1.
#include "common.cuh"
#include "unet.hpp"
#include "json_min.hpp"
#include "png_min.hpp"
#include "overlay.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <filesystem>
#include <thread>
#include <fstream>
#include <iostream>
#include <sstream>
#include <sys/stat.h>

namespace ms {

thread_local LaunchCounter* g_counter = nullptr;
std::atomic<uint64_t> g_alloc_epoch{1};
static thread_local std::string t_last_error = "";

void fail(int code, const std::string& what) { throw Error{code, what}; }

unsigned* watchdog_host_word() {
    static std::mutex mu;
    static unsigned* word = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!word) {
        void* p = nullptr;
        MS_CUDA(cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable));
        word = static_cast<unsigned*>(p);
        *word = 0u;
    }
    return word;
}
static unsigned* g_watchdog_seen = nullptr;   // non-null once any handle loaded a UNet

}  // namespace ms

using namespace ms;

// host copies of one batch's results on the file path (double-buffered: writers of batch k run under batch k + 1)
struct BatchHost {
    PinBuf in, norm, mask;
    std::vector<int32_t> slice_start, cstart, xy, uxy;
    void release() {
        in.release();
        norm.release();
        mask.release();
    }
};

struct ms_handle {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    // config (defaults = the reference's literals)
    int net_h = 512, net_w = 512, n_classes_cfg = 0, max_batch = 32;
    int fg_value = 2, threshold = 127;
    float min_area_ratio = 0.06f;
    std::string weights_path;
    // stages
    UNet unet;
    PreprocessWs pre;
    PostprocessWs post;
    M2pWs m2p;
    // pipeline buffers
    DevBuf d_src, d_norm, d_mask_raw, d_mask, d_logits, d_scratch_in, d_scratch_out;
    PinBuf staging, h_header;
    // double-buffered asynchronous pipeline (ms_submit_batch_host / ms_wait_batch)
    struct Slot {
        DevBuf d_src;
        PolyDev poly;
        PinBuf h_src, h_header, h_slice_start, h_cstart, h_xy;
        cudaEvent_t ev_h2d = nullptr, ev_m2p = nullptr, ev_done = nullptr;
        int batch = 0, w = 0, hgt = 0;
        bool busy = false;
        // the D2H issued with the header is sized from the previous batch on this slot (+25 %); ms_wait_batch fetches the
        // remainder only when a batch turned out larger, so the steady state copies ~what it uses, in one round trip
        int64_t spec_contours = 0, spec_points = 0, last_contours = 0, last_points = 0;
        // CUDA graph of the slot's kernel chain (K1 .. K6), replayed while shape and buffers are unchanged
        cudaGraphExec_t gexec = nullptr;
        int g_w = 0, g_h = 0, g_batch = 0;      // shape the graph was captured for
        int seen_w = 0, seen_h = 0, seen_batch = 0;   // shape of the last eager run (buffers are sized for it)
        uint64_t g_epoch = 0, seen_epoch = 0;
        int64_t g_launches = 0;
    } slots[2];
    bool graphs_on = true;     // MEDSEG_GRAPH=0 disables graph replay
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    BatchHost file_host[2];   // ms_process_raw_file(s) / ms_process_directory
    // log: every line is appended with open/append/close, so the C++ facade's own std::ofstream
    // (opened with ios::app on the same file, see facade.cpp) interleaves correctly with it
    std::string log_path;
    bool log_on() const { return !log_path.empty(); }
    void log(const std::string& line) const {
        if (log_path.empty()) return;
        FILE* f = std::fopen(log_path.c_str(), "a");
        if (!f) return;
        std::fputs(line.c_str(), f);
        std::fputc('\n', f);
        std::fclose(f);
    }
    std::string last_error;
    LaunchCounter counter;
    // single-process multi-GPU ("devices" in the config / MEDSEG_DEVICES): this handle drives devices[0]; `peers` are full
    // handles (own stream, weights, workspaces) for devices[1..], owned by this one.  Slices are independent, so the
    // volume / file-list entry points give every GPU a contiguous block and one host thread (src/main.cpp:148-164).
    std::vector<int> devices;
    std::vector<ms_handle*> peers;
    std::vector<int32_t> mc_slice, mc_cstart, mc_xy;   // ms_process_batch_multiclass_host: the label-major polygon set before the split
    std::string cfg_text;            // the JSON config this handle was created from (peers are created from it)
    bool quiet_console = false;      // peers: console lines are collected per job and printed in job order by the owner
    // bytes this handle moved over PCIe with its own copies (bench.py: e2e.h2d/d2h_bytes_per_step are read from here)
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    // ms_process_batch_dev without result pointers: header lands in h_header asynchronously, ms_last_counts collects it
    cudaEvent_t ev_header = nullptr;
    bool header_pending = false;
    // in-step stage timing (ms_profile_layers_begin switches it on together with the per-layer events): five events per
    // pass of pipeline_dev -> K1 | UNet | K5 | K6
    std::vector<cudaEvent_t> stage_events;
    int stage_cap = 0, stage_n = 0;
};

namespace {

std::string dirname_of(const std::string& p) {
    const size_t s = p.find_last_of('/');
    return s == std::string::npos ? std::string(".") : p.substr(0, s);
}
std::string basename_of(const std::string& p) {
    const size_t s = p.find_last_of('/');
    return s == std::string::npos ? p : p.substr(s + 1);
}
std::string stem_of(const std::string& p) {
    std::string b = basename_of(p);
    const size_t d = b.find_last_of('.');
    return (d == std::string::npos || d == 0) ? b : b.substr(0, d);
}
bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty()) ::mkdir(cur.c_str(), 0777);
        }
        if (i < path.size()) cur += path[i];
    }
}
std::string read_text(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    MS_REQUIRE(f.good(), MS_ERR_IO, "cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// An entry point that fails may have copies into CALLER memory (or kernels reading caller buffers) still in flight on the
// handle's streams: wait for them before the error is reported, so the caller may free or reuse its buffers at once.
void drain(ms_handle* h) {
    if (!h) return;
    for (cudaStream_t st : {h->stream, h->copy_stream, h->d2h_stream})
        if (st) cudaStreamSynchronize(st);
    cudaGetLastError();
}

template <class F>
int guarded(ms_handle* h, F&& f) {
    try {
        if (h) {
            MS_CUDA(cudaSetDevice(h->device));
            g_counter = &h->counter;
        }
        f();
        g_counter = nullptr;
        if (g_watchdog_seen && *reinterpret_cast<volatile unsigned*>(g_watchdog_seen))
            fail(MS_ERR_INTERNAL, "device pipeline watchdog expired (an mbarrier wait exceeded 20 s): results of this process are not trustworthy");
        return MS_OK;
    } catch (const Error& e) {
        g_counter = nullptr;
        drain(h);
        (h ? h->last_error : t_last_error) = e.what;
        if (h) h->log("error: " + e.what);
        return e.code;
    } catch (const std::exception& e) {
        g_counter = nullptr;
        drain(h);
        (h ? h->last_error : t_last_error) = e.what();
        return MS_ERR_INTERNAL;
    } catch (...) {
        g_counter = nullptr;
        drain(h);
        (h ? h->last_error : t_last_error) = "unknown exception";
        return MS_ERR_INTERNAL;
    }
}

cudaStream_t pick_stream(ms_handle* h, void* s) { return s ? reinterpret_cast<cudaStream_t>(s) : h->stream; }

bool is_cuda_host_ptr(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// H2D of caller memory: pinned -> direct async copy, pageable -> through the handle's pinned staging buffer.
void upload(ms_handle* h, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    h->h2d_bytes += (int64_t)bytes;
    if (is_cuda_host_ptr(h_src)) {
        MS_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
    } else {
        h->staging.reserve(bytes);
        std::memcpy(h->staging.p, h_src, bytes);
        MS_CUDA(cudaMemcpyAsync(d_dst, h->staging.p, bytes, cudaMemcpyHostToDevice, st));
    }
}
void download(ms_handle* h, void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
    h->d2h_bytes += (int64_t)bytes;
    MS_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
}
void download_sync(ms_handle* h, void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
    download(h, h_dst, d_src, bytes, st);
    MS_CUDA(cudaStreamSynchronize(st));
}

void parse_devices(ms_handle* h, const std::string& devs) {
    h->devices.clear();
    if (devs == "all") {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        for (int i = 0; i < n; ++i) h->devices.push_back(i);
    } else if (!devs.empty()) {
        std::stringstream ss(devs);
        std::string tok;
        while (std::getline(ss, tok, ',')) {
            MS_REQUIRE(!tok.empty() && tok.find_first_not_of("0123456789 ") == std::string::npos, MS_ERR_FORMAT,
                       "devices: expected a list of ordinals or \"all\"");
            const int d = std::atoi(tok.c_str());
            MS_REQUIRE(std::find(h->devices.begin(), h->devices.end(), d) == h->devices.end(), MS_ERR_FORMAT, "devices: duplicate ordinal");
            h->devices.push_back(d);
        }
    }
    if (!h->devices.empty()) h->device = h->devices[0];
}

void apply_config(ms_handle* h, const json::Value& cfg, const std::string& base_dir) {
    MS_REQUIRE(cfg.type == json::Value::Obj, MS_ERR_FORMAT, "config must be a JSON object");
    h->net_h = (int)cfg.integer("net_h", cfg.integer("net_size", 512));
    h->net_w = (int)cfg.integer("net_w", cfg.integer("net_size", 512));
    h->n_classes_cfg = (int)cfg.integer("n_classes", 0);
    h->max_batch = (int)cfg.integer("max_batch", 32);
    h->fg_value = (int)cfg.integer("foreground_value", 2);
    h->threshold = (int)cfg.integer("threshold", 127);
    h->min_area_ratio = (float)cfg.number("min_area_ratio", 0.06f);
    // opt-in extra, never on the reference's path (src/mask2polygon.cpp:34 keeps CHAIN_APPROX_SIMPLE): > 0 = every contour goes
    // through approxPolyDP(eps, closed) in network space before map_contour_points
    h->m2p.dp_eps = cfg.number("dp_epsilon", 0.0);
    MS_REQUIRE(h->m2p.dp_eps >= 0.0 && h->m2p.dp_eps < 1e9, MS_ERR_ARG, "dp_epsilon must be a finite number >= 0");
    h->device = (int)cfg.integer("device", 0);
    // "devices": [0, 1, ...] or "all" (MEDSEG_DEVICES=all|0,1,.. supplies it when the config names neither, so the
    // reference's `init <path>` needs no new argument): one process, one handle per listed GPU
    std::string devs;
    if (cfg.has("devices")) {
        const json::Value& dv = cfg.at("devices");
        if (dv.type == json::Value::Arr) {
            for (const auto& e : dv.arr) devs += (devs.empty() ? "" : ",") + std::to_string((int)e.num);
        } else if (dv.type == json::Value::Str) {
            devs = dv.str;
        }
    } else if (!cfg.has("device")) {
        if (const char* e = std::getenv("MEDSEG_DEVICES")) devs = e;
    }
    parse_devices(h, devs);
    std::string w = cfg.string("weights", "");
    if (!w.empty() && w[0] != '/') w = base_dir + "/" + w;
    h->weights_path = w;
    const std::string head = cfg.string("head", "");
    if (head == "binary") MS_REQUIRE(h->n_classes_cfg == 0 || h->n_classes_cfg == 1, MS_ERR_FORMAT, "head=binary needs n_classes=1");
    MS_REQUIRE(h->fg_value >= 1 && h->fg_value <= 255, MS_ERR_ARG, "foreground_value out of range");
    MS_REQUIRE(h->max_batch >= 1, MS_ERR_ARG, "max_batch must be >= 1");
}

void finish_init(ms_handle* h, const char* log_dir) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail(MS_ERR_CUDA, "no CUDA device available (libmedseg_b200 has no CPU fallback)");
    }
    MS_REQUIRE(h->device >= 0 && h->device < ndev, MS_ERR_ARG, "device ordinal out of range");
    MS_CUDA(cudaSetDevice(h->device));
    if (const char* g = std::getenv("MEDSEG_GRAPH")) h->graphs_on = g[0] != '0';
    cudaDeviceProp prop{};
    MS_CUDA(cudaGetDeviceProperties(&prop, h->device));
    MS_REQUIRE(prop.major == 10, MS_ERR_CUDA,
               std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                   "; libmedseg_b200 carries sm_100a code only");
    h->sm_count = prop.multiProcessorCount;
    MS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (log_dir && log_dir[0]) {
        mkdirs(log_dir);                                                   // src/initialize.cpp:29
        h->log_path = std::string(log_dir) + "/segmentation_log.txt";      // :30
        FILE* lf = std::fopen(h->log_path.c_str(), "w");                   // :31 (truncate)
        if (!lf) {
            const std::string p = h->log_path;
            h->log_path.clear();
            fail(MS_ERR_IO, "Failed to create log file: " + p);
        }
        std::fclose(lf);
        h->log("=== Initializing Medical Image Segmentation Engine ===");
        h->log(std::string("Engine: libmedseg_b200 (sm_100a) on ") + prop.name + ", weights: " + h->weights_path);
    }
    if (!h->weights_path.empty()) {
        g_counter = &h->counter;
        h->unet.load(h->weights_path, h->net_h, h->net_w, h->n_classes_cfg, h->max_batch, h->fg_value, h->sm_count);
        g_watchdog_seen = watchdog_host_word();
        h->log("UNet loaded: " + std::to_string(h->unet.n_params()) + " parameters, " +
               std::to_string(h->unet.flops_per_slice() / 1e9) + " GFLOP per slice, n_classes=" + std::to_string(h->unet.n_classes()));
    }
    const size_t npx = (size_t)h->max_batch * h->net_h * h->net_w;
    h->d_norm.reserve(npx);
    h->d_mask_raw.reserve(npx);
    h->d_mask.reserve(npx);
    h->h_header.reserve(4 * sizeof(long long));
    MS_CUDA(cudaStreamSynchronize(h->stream));
}

// Runs phase A (+ B) of mask2polygon with automatic growth of the device-side capacities.  On return
// the pinned header holds {n_contours, n_points, overflow, trace_errors} and the stream is idle.
// `d_raw` != null: the class mask straight from the head; K5 (clean mask into `d_mask`) then runs as part of phase A.
void run_m2p(ms_handle* h, const uint8_t* d_mask, int hgt, int w, int batch, int threshold, int orig_w, int orig_h, cudaStream_t st,
             const uint8_t* d_raw = nullptr, const FgSpec* multi = nullptr) {
    long long* hh = h->h_header.as<long long>();
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (d_raw) post_poly_phase_a(h->post, h->m2p, h->m2p.poly, d_raw, const_cast<uint8_t*>(d_mask), hgt, w, batch, h->fg_value, h->min_area_ratio, st, multi);
        else m2p_phase_a(h->m2p, h->m2p.poly, d_mask, hgt, w, batch, threshold, st);
        m2p_phase_b(h->m2p, h->m2p.poly, hgt, w, batch, orig_w, orig_h, st);
        download_sync(h, hh, h->m2p.poly.header.p, 4 * sizeof(long long), st);
        MS_REQUIRE(hh[3] == 0, MS_ERR_INTERNAL, "mask2polygon: border following did not terminate");
        MS_REQUIRE((hh[2] & 4) == 0, MS_ERR_CAPACITY, "mask2polygon: more than 2^31 points");
        bool grown = false;
        if (hh[0] > h->m2p.poly.cap_contours) { h->m2p.poly.cap_contours = hh[0] + hh[0] / 4 + 64; grown = true; }
        if (hh[1] > h->m2p.poly.cap_points) { h->m2p.poly.cap_points = hh[1] + hh[1] / 4 + 1024; grown = true; }
        if (!grown) return;
    }
    fail(MS_ERR_INTERNAL, "mask2polygon: capacity growth did not converge");
}

void copy_polygons_out(ms_handle* h, int batch, ms_polygons* out, cudaStream_t st) {
    const long long* hh = h->h_header.as<long long>();
    out->n_contours = hh[0];
    out->n_points = hh[1];
    MS_REQUIRE(out->cap_contours >= hh[0] && out->cap_points >= hh[1], MS_ERR_CAPACITY,
               "polygon buffers too small: need " + std::to_string(hh[0]) + " contours, " + std::to_string(hh[1]) + " points");
    download(h, out->slice_start, h->m2p.poly.slice_start.p, ((size_t)batch + 1) * 4, st);
    download(h, out->contour_start, h->m2p.poly.npts.p, ((size_t)hh[0] + 1) * 4, st);
    if (hh[1] > 0) download(h, out->xy, h->m2p.poly.xy.p, (size_t)hh[1] * 8, st);
    MS_CUDA(cudaStreamSynchronize(st));
}

void check_polys(const ms_polygons* out, int batch) {
    MS_REQUIRE(out && out->slice_start && out->contour_start && (out->xy || out->cap_points == 0) && out->cap_contours >= 0 &&
                   out->cap_points >= 0 && batch > 0,
               MS_ERR_ARG, "ms_polygons: null buffer or negative capacity");
}

// device pipeline shared by the batch entry points; input u16 already at d_src.  `sync` = false: one pass with the current
// polygon capacities, the header is copied to pinned memory asynchronously and collected by ms_last_counts (no host
// round trip inside the call); true: wait, and grow the capacities / re-run mask2polygon when they were too small.
void pipeline_dev(ms_handle* h, const uint16_t* d_src, int w, int hgt, int batch, cudaStream_t st, bool sync = true) {
    MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded (handle was created without a weight blob)");
    MS_REQUIRE(batch >= 1 && batch <= h->max_batch, MS_ERR_ARG, "batch exceeds max_batch of this handle");
    MS_REQUIRE(w > 0 && hgt > 0, MS_ERR_ARG, "bad slice size");
    uint8_t* norm = h->d_norm.as<uint8_t>();
    uint8_t* raw = h->d_mask_raw.as<uint8_t>();
    uint8_t* mask = h->d_mask.as<uint8_t>();
    cudaEvent_t* sev = h->stage_n < h->stage_cap ? h->stage_events.data() + (size_t)h->stage_n * 5 : nullptr;
    if (sev) MS_CUDA(cudaEventRecord(sev[0], st));
    preprocess_launch(h->pre, d_src, w, hgt, batch, h->net_w, h->net_h, norm, nullptr, st);         // src/process.cpp:211
    if (sev) MS_CUDA(cudaEventRecord(sev[1], st));
    h->unet.forward(norm, batch, raw, nullptr, st);                                                  // :224
    if (sev) MS_CUDA(cudaEventRecord(sev[2], st));
    // K5 postprocess (:231) + mask_to_image / threshold(127) (:234, src/mask2polygon.cpp:31) + K6: after postprocess the
    // mask is {0, fg} and LUT(fg) > 127 <=> value == fg, so the kept components ARE mask2polygon's foreground.  Slices that
    // fit in shared memory run K5 and phase A of K6 as ONE kernel (the K5 | K6 stage boundary then sits before the
    // finalize kernel: K5's share is not separable, stage timing reports the sum under K6 and 0 under K5).
    if (sync) {
        if (sev) MS_CUDA(cudaEventRecord(sev[3], st));
        run_m2p(h, mask, h->net_h, h->net_w, batch, h->fg_value - 1, w, hgt, st, raw);               // :231-242
        h->header_pending = false;
        if (sev) {
            MS_CUDA(cudaEventRecord(sev[4], st));
            ++h->stage_n;
        }
        return;
    }
    const bool one_kernel = slice_fused_supported(h->net_h, h->net_w);
    if (!one_kernel) postprocess_launch(h->post, raw, mask, h->net_h, h->net_w, batch, h->fg_value, h->min_area_ratio, st);
    if (sev) MS_CUDA(cudaEventRecord(sev[3], st));
    if (one_kernel) post_poly_phase_a(h->post, h->m2p, h->m2p.poly, raw, mask, h->net_h, h->net_w, batch, h->fg_value, h->min_area_ratio, st);
    else m2p_phase_a(h->m2p, h->m2p.poly, mask, h->net_h, h->net_w, batch, h->fg_value - 1, st);
    m2p_phase_b(h->m2p, h->m2p.poly, h->net_h, h->net_w, batch, w, hgt, st);
    if (sev) {
        MS_CUDA(cudaEventRecord(sev[4], st));
        ++h->stage_n;
    }
    if (!h->ev_header) MS_CUDA(cudaEventCreateWithFlags(&h->ev_header, cudaEventDisableTiming));
    download(h, h->h_header.p, h->m2p.poly.header.p, 4 * sizeof(long long), st);
    MS_CUDA(cudaEventRecord(h->ev_header, st));
    h->header_pending = true;
}

// waits for the header of the last asynchronous pipeline_dev and validates it; grows the capacities on overflow so that
// the next call fits
void collect_header(ms_handle* h) {
    if (h->header_pending) {
        MS_CUDA(cudaEventSynchronize(h->ev_header));
        h->header_pending = false;
    }
    const long long* hh = h->h_header.as<long long>();
    MS_REQUIRE(hh[3] == 0, MS_ERR_INTERNAL, "mask2polygon: border following did not terminate");
    MS_REQUIRE((hh[2] & 4) == 0, MS_ERR_CAPACITY, "mask2polygon: more than 2^31 points");
    bool grown = false;
    if (hh[0] > h->m2p.poly.cap_contours) { h->m2p.poly.cap_contours = hh[0] + hh[0] / 4 + 64; grown = true; }
    if (hh[1] > h->m2p.poly.cap_points) { h->m2p.poly.cap_points = hh[1] + hh[1] / 4 + 1024; grown = true; }
    MS_REQUIRE(!grown && (hh[2] & 3) == 0, MS_ERR_CAPACITY,
               "device polygon workspace was too small for the last batch (it has been grown: run the batch again)");
}

}  // namespace

// ================================================================================== C ABI
extern "C" {

static int init_common(ms_handle* h, const char* log_dir, ms_handle** out) {
    int rc = guarded(nullptr, [&] { finish_init(h, log_dir); });
    if (rc != MS_OK) {
        t_last_error = t_last_error.empty() ? h->last_error : t_last_error;
        if (h->stream) cudaStreamDestroy(h->stream);
        h->stream = nullptr;
        delete h;
        return rc;
    }
    // one more full handle per additional GPU (same configuration, quiet console, same log file)
    for (size_t i = 1; i < h->devices.size(); ++i) {
        ms_handle* c = new ms_handle();
        c->device = h->devices[i];
        c->net_h = h->net_h; c->net_w = h->net_w; c->n_classes_cfg = h->n_classes_cfg; c->max_batch = h->max_batch;
        c->fg_value = h->fg_value; c->threshold = h->threshold; c->min_area_ratio = h->min_area_ratio;
        c->m2p.dp_eps = h->m2p.dp_eps;
        c->weights_path = h->weights_path;
        c->quiet_console = true;
        rc = guarded(nullptr, [&] { finish_init(c, nullptr); });
        if (rc != MS_OK) {
            t_last_error = "device " + std::to_string(c->device) + ": " + t_last_error;
            if (c->stream) cudaStreamDestroy(c->stream);
            c->stream = nullptr;
            delete c;
            ms_destroy(h);
            return rc;
        }
        c->log_path = h->log_path;
        h->peers.push_back(c);
    }
    if (!h->peers.empty()) h->log("Devices: " + std::to_string(h->devices.size()) + " GPUs in this process, one worker thread each");
    cudaSetDevice(h->device);
    *out = h;
    return MS_OK;
}

int ms_init(const char* path, const char* log_dir, ms_handle** out) {
    if (!out) return MS_ERR_ARG;
    *out = nullptr;
    ms_handle* h = new ms_handle();
    int rc = guarded(nullptr, [&] {
        const std::string p = path ? path : "";
        if (ends_with(p, ".json")) {
            json::Value cfg;
            try {
                cfg = json::parse(read_text(p));
            } catch (const std::exception& e) {
                fail(MS_ERR_FORMAT, std::string("config ") + p + ": " + e.what());
            }
            apply_config(h, cfg, dirname_of(p));
        } else if (!p.empty()) {
            struct stat sb;
            MS_REQUIRE(::stat(p.c_str(), &sb) == 0, MS_ERR_IO, "Error: engine file not found - " + p);  // src/initialize.cpp:42-45
            h->weights_path = p;
            if (const char* e = std::getenv("MEDSEG_DEVICES")) parse_devices(h, e);
        }
    });
    if (rc != MS_OK) { delete h; return rc; }
    return init_common(h, log_dir, out);
}

int ms_init_json(const char* cfg_json_text, const char* log_dir, ms_handle** out) {
    if (!out || !cfg_json_text) return MS_ERR_ARG;
    *out = nullptr;
    ms_handle* h = new ms_handle();
    int rc = guarded(nullptr, [&] {
        json::Value cfg;
        try {
            cfg = json::parse(cfg_json_text);
        } catch (const std::exception& e) {
            fail(MS_ERR_FORMAT, std::string("config: ") + e.what());
        }
        apply_config(h, cfg, ".");
    });
    if (rc != MS_OK) { delete h; return rc; }
    return init_common(h, log_dir, out);
}

void ms_destroy(ms_handle* h) {
    if (!h) return;
    for (ms_handle* c : h->peers) {
        c->log_path.clear();          // the owner writes the clean-up lines once
        ms_destroy(c);
    }
    h->peers.clear();
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->log("\n=== Cleaning Up Resources ===");          // src/cleanup.cpp:13
    h->log("All resources cleaned up successfully");    // :52
    for (auto& S : h->slots) {
        if (S.ev_done) cudaEventSynchronize(S.ev_done);
        if (S.gexec) cudaGraphExecDestroy(S.gexec);
        for (cudaEvent_t e : {S.ev_h2d, S.ev_m2p, S.ev_done})
            if (e) cudaEventDestroy(e);
    }
    if (h->ev_header) cudaEventDestroy(h->ev_header);
    for (cudaEvent_t e : h->stage_events) cudaEventDestroy(e);
    for (cudaStream_t st : {h->copy_stream, h->d2h_stream, h->stream})
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    // every DevBuf / PinBuf member (stage workspaces, polygon stores incl. the chunk store, slots, file staging) frees
    // itself in its destructor, on the success path here and on a failed ms_init alike
    delete h;  // ~UNet frees weights and activations
}

const char* ms_last_error(ms_handle* h) { return h ? h->last_error.c_str() : t_last_error.c_str(); }

int ms_get_info(ms_handle* h, ms_info* out) {
    if (!h || !out) return MS_ERR_ARG;
    out->device = h->device;
    out->sm_count = h->sm_count;
    out->net_h = h->net_h;
    out->net_w = h->net_w;
    out->n_classes = h->unet.loaded() ? h->unet.n_classes() : h->n_classes_cfg;
    out->max_batch = h->max_batch;
    out->foreground_value = h->fg_value;
    out->min_area_ratio = h->min_area_ratio;
    out->has_weights = h->unet.loaded() ? 1 : 0;
    out->n_params = h->unet.loaded() ? h->unet.n_params() : 0;
    out->flops_per_slice = h->unet.loaded() ? (int64_t)h->unet.flops_per_slice() : 0;
    return MS_OK;
}

void* ms_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void ms_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------- preprocess
int ms_preprocess_dev(ms_handle* h, const uint16_t* d_src, int w, int hgt, int batch, uint8_t* d_out_u8, void* d_out_bf16, void* stream) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(d_src && d_out_u8, MS_ERR_ARG, "preprocess: null pointer");
        preprocess_launch(h->pre, d_src, w, hgt, batch, h->net_w, h->net_h, d_out_u8, reinterpret_cast<__nv_bfloat16*>(d_out_bf16),
                          pick_stream(h, stream));
    });
}
int ms_preprocess_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int batch, uint8_t* h_out_u8) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h_src && h_out_u8 && w > 0 && hgt > 0 && batch > 0, MS_ERR_ARG, "preprocess: bad argument");
        const size_t in_bytes = (size_t)w * hgt * 2 * batch, out_bytes = (size_t)h->net_w * h->net_h * batch;
        h->d_scratch_in.reserve(in_bytes);
        h->d_scratch_out.reserve(out_bytes);
        upload(h, h->d_scratch_in.p, h_src, in_bytes, h->stream);
        preprocess_launch(h->pre, h->d_scratch_in.as<uint16_t>(), w, hgt, batch, h->net_w, h->net_h, h->d_scratch_out.as<uint8_t>(), nullptr, h->stream);
        download_sync(h, h_out_u8, h->d_scratch_out.p, out_bytes, h->stream);
    });
}

// ---------------------------------------------------------------- UNet
int ms_unet_forward_dev(ms_handle* h, const uint8_t* d_in_u8, int batch, uint8_t* d_mask, float* d_logits, void* stream) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(d_in_u8 && d_mask, MS_ERR_ARG, "unet_forward: null pointer");
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        h->unet.forward(d_in_u8, batch, d_mask, d_logits, pick_stream(h, stream));
    });
}
int ms_unet_forward_host(ms_handle* h, const uint8_t* h_in_u8, int batch, uint8_t* h_mask, float* h_logits) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h_in_u8 && h_mask, MS_ERR_ARG, "unet_forward: null pointer");
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        MS_REQUIRE(batch >= 1 && batch <= h->max_batch, MS_ERR_ARG, "batch exceeds max_batch");
        const size_t npx = (size_t)h->net_w * h->net_h * batch;
        upload(h, h->d_norm.p, h_in_u8, npx, h->stream);
        float* dl = nullptr;
        if (h_logits) {
            h->d_logits.reserve(npx * h->unet.n_classes() * 4);
            dl = h->d_logits.as<float>();
        }
        h->unet.forward(h->d_norm.as<uint8_t>(), batch, h->d_mask_raw.as<uint8_t>(), dl, h->stream);
        if (h_logits) MS_CUDA(cudaMemcpyAsync(h_logits, dl, npx * h->unet.n_classes() * 4, cudaMemcpyDeviceToHost, h->stream));
        download_sync(h, h_mask, h->d_mask_raw.p, npx, h->stream);
    });
}

// ---------------------------------------------------------------- postprocess
int ms_postprocess_dev(ms_handle* h, const uint8_t* d_in, uint8_t* d_out, int hgt, int w, int batch, int fg_value, void* stream) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(d_in && d_out, MS_ERR_ARG, "postprocess: null pointer");
        postprocess_launch(h->post, d_in, d_out, hgt, w, batch, fg_value > 0 ? fg_value : h->fg_value, h->min_area_ratio, pick_stream(h, stream));
    });
}
int ms_postprocess_host(ms_handle* h, const uint8_t* h_in, uint8_t* h_out, int hgt, int w, int batch, int fg_value) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h_in && h_out && hgt > 0 && w > 0 && batch > 0, MS_ERR_ARG, "postprocess: bad argument");
        const size_t n = (size_t)hgt * w * batch;
        h->d_scratch_in.reserve(n);
        h->d_scratch_out.reserve(n);
        upload(h, h->d_scratch_in.p, h_in, n, h->stream);
        postprocess_launch(h->post, h->d_scratch_in.as<uint8_t>(), h->d_scratch_out.as<uint8_t>(), hgt, w, batch,
                           fg_value > 0 ? fg_value : h->fg_value, h->min_area_ratio, h->stream);
        download_sync(h, h_out, h->d_scratch_out.p, n, h->stream);
    });
}

// ---------------------------------------------------------------- mask2polygon
int ms_mask2polygon_dev(ms_handle* h, const uint8_t* d_mask, int hgt, int w, int batch, int threshold, int orig_w, int orig_h,
                        ms_polygons* out, void* stream) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        check_polys(out, batch);
        MS_REQUIRE(d_mask && orig_w > 0 && orig_h > 0, MS_ERR_ARG, "mask2polygon: bad argument");
        cudaStream_t st = pick_stream(h, stream);
        run_m2p(h, d_mask, hgt, w, batch, threshold, orig_w, orig_h, st);
        copy_polygons_out(h, batch, out, st);
    });
}
int ms_mask2polygon_host(ms_handle* h, const uint8_t* h_mask, int hgt, int w, int batch, int threshold, int orig_w, int orig_h,
                         ms_polygons* out) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        check_polys(out, batch);
        MS_REQUIRE(h_mask && hgt > 0 && w > 0 && orig_w > 0 && orig_h > 0, MS_ERR_ARG, "mask2polygon: bad argument");
        const size_t n = (size_t)hgt * w * batch;
        h->d_scratch_in.reserve(n);
        upload(h, h->d_scratch_in.p, h_mask, n, h->stream);
        run_m2p(h, h->d_scratch_in.as<uint8_t>(), hgt, w, batch, threshold, orig_w, orig_h, h->stream);
        copy_polygons_out(h, batch, out, h->stream);
    });
}

// ---------------------------------------------------------------- whole path
int ms_process_batch_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int batch, ms_polygons* out, uint8_t* h_norm_u8,
                          uint8_t* h_mask_u8) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        check_polys(out, batch);
        MS_REQUIRE(h_src && w > 0 && hgt > 0, MS_ERR_ARG, "process_batch: bad argument");
        const size_t in_bytes = (size_t)w * hgt * 2 * batch;
        h->d_src.reserve(in_bytes);
        upload(h, h->d_src.p, h_src, in_bytes, h->stream);
        pipeline_dev(h, h->d_src.as<uint16_t>(), w, hgt, batch, h->stream);
        const size_t npx = (size_t)h->net_w * h->net_h * batch;
        if (h_norm_u8) MS_CUDA(cudaMemcpyAsync(h_norm_u8, h->d_norm.p, npx, cudaMemcpyDeviceToHost, h->stream));
        if (h_mask_u8) MS_CUDA(cudaMemcpyAsync(h_mask_u8, h->d_mask.p, npx, cudaMemcpyDeviceToHost, h->stream));
        copy_polygons_out(h, batch, out, h->stream);
    });
}

// cfg4 (BASELINE.json configs[3]): per-class contours of a multi-class head in ONE call.  The reference runs the path for one
// label (FOREGROUND_VALUE = 2, src/postprocess.cpp:5); here K1 and the UNet run once and K5 + K6 run ONCE over
// n_labels x batch virtual slices (FgSpec: virtual slice s cleans label classes[s / batch] of class mask s % batch), so the
// labels share every launch; the polygon set comes back in one piece and is split per label on the host.
int ms_process_batch_multiclass_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int batch, const int32_t* classes, int n_classes,
                                     ms_polygons* outs, uint8_t* h_raw_mask_u8, uint8_t* h_clean_masks_u8) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h_src && w > 0 && hgt > 0 && classes && outs && n_classes >= 1 && n_classes <= 16, MS_ERR_ARG,
                   "process_batch_multiclass: bad argument (1 .. 16 labels)");
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded (handle was created without a weight blob)");
        MS_REQUIRE(batch >= 1 && batch <= h->max_batch, MS_ERR_ARG, "batch exceeds max_batch of this handle");
        FgSpec fgs{};
        fgs.in_mod = batch;
        for (int i = 0; i < n_classes; ++i) {
            check_polys(&outs[i], batch);
            MS_REQUIRE(classes[i] >= 1 && classes[i] <= 255, MS_ERR_ARG, "process_batch_multiclass: labels must be 1 .. 255");
            fgs.v[i] = (unsigned char)classes[i];
        }
        cudaStream_t st = h->stream;
        const int vbatch = batch * n_classes;
        const size_t in_bytes = (size_t)w * hgt * 2 * batch;
        const size_t npx = (size_t)h->net_w * h->net_h * batch;
        h->d_src.reserve(in_bytes);
        h->d_mask.reserve(npx * n_classes);
        upload(h, h->d_src.p, h_src, in_bytes, st);
        uint8_t* norm = h->d_norm.as<uint8_t>();
        uint8_t* raw = h->d_mask_raw.as<uint8_t>();
        uint8_t* mask = h->d_mask.as<uint8_t>();
        preprocess_launch(h->pre, h->d_src.as<uint16_t>(), w, hgt, batch, h->net_w, h->net_h, norm, nullptr, st);
        h->unet.forward(norm, batch, raw, nullptr, st);
        if (h_raw_mask_u8) MS_CUDA(cudaMemcpyAsync(h_raw_mask_u8, raw, npx, cudaMemcpyDeviceToHost, st));
        run_m2p(h, mask, h->net_h, h->net_w, vbatch, 0, w, hgt, st, raw, &fgs);
        h->header_pending = false;
        if (h_clean_masks_u8) MS_CUDA(cudaMemcpyAsync(h_clean_masks_u8, mask, npx * n_classes, cudaMemcpyDeviceToHost, st));
        // the whole set (virtual slices in label-major order), then one slice-range per label
        const long long* hh = h->h_header.as<long long>();
        const int64_t nc = hh[0], np = hh[1];
        h->mc_slice.resize((size_t)vbatch + 1);
        h->mc_cstart.resize((size_t)nc + 1);
        h->mc_xy.resize((size_t)np * 2);
        ms_polygons all{h->mc_xy.data(), np, h->mc_cstart.data(), nc, h->mc_slice.data(), 0, 0};
        copy_polygons_out(h, vbatch, &all, st);
        int rc_capacity = MS_OK;
        std::string cap_msg;
        for (int i = 0; i < n_classes; ++i) {
            const int32_t c0 = h->mc_slice[(size_t)i * batch], c1 = h->mc_slice[(size_t)(i + 1) * batch];
            const int32_t p0 = h->mc_cstart[c0], p1 = h->mc_cstart[c1];
            outs[i].n_contours = c1 - c0;
            outs[i].n_points = p1 - p0;
            if (outs[i].cap_contours < c1 - c0 || outs[i].cap_points < p1 - p0) {     // report every label's sizes before failing
                rc_capacity = MS_ERR_CAPACITY;
                cap_msg = "polygon buffers too small for label " + std::to_string(classes[i]) + ": need " + std::to_string(c1 - c0) +
                          " contours, " + std::to_string(p1 - p0) + " points";
                continue;
            }
            for (int b2 = 0; b2 <= batch; ++b2) outs[i].slice_start[b2] = h->mc_slice[(size_t)i * batch + b2] - c0;
            for (int32_t c = c0; c <= c1; ++c) outs[i].contour_start[c - c0] = h->mc_cstart[c] - p0;
            if (p1 > p0) std::memcpy(outs[i].xy, h->mc_xy.data() + 2 * (size_t)p0, (size_t)(p1 - p0) * 8);
        }
        MS_REQUIRE(rc_capacity == MS_OK, MS_ERR_CAPACITY, cap_msg);
    });
}

int ms_process_batch_dev(ms_handle* h, const uint16_t* d_src, int w, int hgt, int batch, int64_t* n_points, int64_t* n_contours, void* stream) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(d_src, MS_ERR_ARG, "process_batch_dev: null pointer");
        const bool sync = n_points || n_contours;
        pipeline_dev(h, d_src, w, hgt, batch, pick_stream(h, stream), sync);
        const long long* hh = h->h_header.as<long long>();
        if (n_contours) *n_contours = hh[0];
        if (n_points) *n_points = hh[1];
    });
}

int ms_last_counts(ms_handle* h, int64_t* n_points, int64_t* n_contours) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        collect_header(h);
        const long long* hh = h->h_header.as<long long>();
        if (n_contours) *n_contours = hh[0];
        if (n_points) *n_points = hh[1];
    });
}

int ms_get_transfer_bytes(ms_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!h) return MS_ERR_ARG;
    int64_t a = h->h2d_bytes, b = h->d2h_bytes;
    for (const ms_handle* c : h->peers) { a += c->h2d_bytes; b += c->d2h_bytes; }
    if (h2d_bytes) *h2d_bytes = a;
    if (d2h_bytes) *d2h_bytes = b;
    return MS_OK;
}

int64_t ms_polygons_to_json(const int32_t* xy, const int32_t* contour_start, int n_contours, const char* base_name, int orig_w,
                            int orig_h, char* dst, int64_t cap) {
    if (n_contours < 0 || (n_contours > 0 && (!xy || !contour_start)) || !base_name) return MS_ERR_ARG;
    try {
        const std::string s = json::labelme_text(xy, contour_start, n_contours, base_name, orig_w, orig_h);
        if (dst && cap > 0) std::memcpy(dst, s.data(), (size_t)std::min<int64_t>(cap, (int64_t)s.size()));
        return (int64_t)s.size();
    } catch (...) {
        return MS_ERR_INTERNAL;
    }
}

int64_t ms_polygons_to_json_batch(const int32_t* xy, const int32_t* contour_start, const int32_t* slice_start, int n_slices,
                                  const char* const* base_names, int orig_w, int orig_h, int n_threads, char* dst, int64_t cap,
                                  int64_t* offsets) {
    if (n_slices < 0 || !contour_start || !slice_start || !base_names || (n_slices > 0 && slice_start[n_slices] > 0 && !xy)) return MS_ERR_ARG;
    try {
        const int T = std::max(1, std::min(n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency(), std::max(1, n_slices)));
        std::vector<std::string> text((size_t)n_slices);
        std::atomic<int> next{0};
        auto work = [&] {
            for (int s = next.fetch_add(1); s < n_slices; s = next.fetch_add(1)) {
                const int c0 = slice_start[s], nc = slice_start[s + 1] - c0;
                // the reference writes no document for a slice without contours (src/mask2polygon.cpp:183-186)
                if (nc > 0) json::labelme_append(text[s], xy, contour_start + c0, nc, base_names[s] ? base_names[s] : "", orig_w, orig_h);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        int64_t total = 0;
        for (const auto& t : text) total += (int64_t)t.size();
        if (offsets) {
            int64_t o = 0;
            for (int s = 0; s < n_slices; ++s) {
                offsets[s] = o;
                o += (int64_t)text[s].size();
            }
            offsets[n_slices] = o;
        }
        if (dst && cap >= total) {
            int64_t o = 0;
            for (const auto& t : text) {
                std::memcpy(dst + o, t.data(), t.size());
                o += (int64_t)t.size();
            }
        }
        return total;
    } catch (...) {
        return MS_ERR_INTERNAL;
    }
}

// ---------------------------------------------------------------- file path (P0 + N1 + N3)
namespace {

// contiguous block of `total` items for part `i` of `parts` (sharding.py: shard_range; earlier parts take the remainder)
std::pair<int64_t, int64_t> block_of(int64_t total, int parts, int i) {
    const int64_t base = total / parts, rem = total % parts;
    const int64_t lo = i * base + std::min<int64_t>(i, rem);
    return {lo, lo + base + (i < rem ? 1 : 0)};
}

// Everything the artefact writer of one slice needs; no CUDA, no handle state, so slices are written concurrently.
struct SliceJob {
    std::string raw, dir, base;
    int status = MS_OK;          // first failure (read or write) of this file
    std::string error;
    std::string console;         // the reference's std::cout lines for this file, printed in file order by the caller
    int slot = -1;               // index inside its batch (-1: never reached the GPU)
};
// the five artefacts of src/process.cpp:207-242 / src/mask2polygon.cpp:134-222 for one slice
// `spare_threads` > 0 (fewer slices than writer threads, e.g. the single-file call): the two grey PNGs are encoded by helper
// threads while this one draws and encodes the overlay.
void write_artefacts(SliceJob& J, const BatchHost& B, int w, int hgt, int net_w, int net_h, int fg_value, int spare_threads = 0) {
    std::thread helper_norm, helper_mask;
    auto join_helpers = [&] {
        if (helper_norm.joinable()) helper_norm.join();
        if (helper_mask.joinable()) helper_mask.join();
    };
    try {
        const size_t npx = (size_t)net_w * net_h;
        const uint8_t* norm = B.norm.as<uint8_t>() + (size_t)J.slot * npx;
        const uint8_t* mask = B.mask.as<uint8_t>() + (size_t)J.slot * npx;
        const int c0 = B.slice_start[J.slot], nc = B.slice_start[J.slot + 1] - c0;
        const int32_t* cstart = B.cstart.data() + c0;
        std::ostringstream out;
        mkdirs(J.dir);
        const std::string png_path = J.dir + "/" + J.base + "_normalized.png";                 // src/process.cpp:207
        const std::string sizes_path = J.dir + "/" + J.base + "_original_sizes.json";          // :208
        const std::string mask_path = J.dir + "/" + J.base + "_mask.png";                      // :209
        bool ok_norm = true, ok_mask = true;
        auto write_norm = [&] { ok_norm = png::write_file(png_path, norm, net_w, net_h, 1); };                  // src/preprocess.cpp:121-122
        auto write_mask = [&] {
            uint8_t lut[256] = {0};                                                                              // src/process.cpp:178-185
            lut[fg_value & 255] = 255;
            lut[1] = 128;
            lut[2] = 255;
            ok_mask = png::write_bytes(mask_path, png::encode_rows(net_w, net_h, 1, [&](int y, uint8_t* dst) {   // :236-239
                const uint8_t* m = mask + (size_t)y * net_w;
                for (int x = 0; x < net_w; ++x) dst[x] = lut[m[x]];
            }));
        };
        if (spare_threads >= 2) {
            helper_norm = std::thread(write_norm);
            helper_mask = std::thread(write_mask);
        } else {
            write_norm();
            write_mask();
        }
        {   // src/preprocess.cpp:126-134
            std::ofstream jf(sizes_path, std::ios::binary);
            MS_REQUIRE(jf.good(), MS_ERR_IO, "cannot write " + sizes_path);
            jf << json::sidecar_text(basename_of(J.raw), w, hgt, net_w, net_h);
        }
        out << "Processing Mask: " << J.base + ".png" << "\n";                                // src/mask2polygon.cpp:141
        out << "Original Size: " << w << "x" << hgt << "\n";                                  // :162
        out << "Scaled Size: " << net_w << "x" << net_h << "\n";                              // :163
        if (nc == 0) {
            out << "Warning: No Contours Detected\n";                                         // :184 (no JSON is written)
        } else {
            out << "Extracted " << nc << " Contours\n";                                       // :187
            // overlay with unmapped (network-space) contours, red, 1 px (src/mask2polygon.cpp:114-129, 189-193): rows are
            // expanded grey -> RGB and painted from a 1-bit contour map while the PNG is written
            const std::vector<uint32_t> red = contour_bitmap(net_w, net_h, B.uxy.data(), cstart, nc);
            const int rp = (net_w + 31) / 32;
            const std::string overlay_path = J.dir + "/" + J.base + "_contour_overlay.png";  // :190
            const bool ok_overlay = png::write_bytes(overlay_path, png::encode_rows(net_w, net_h, 3, [&](int y, uint8_t* dst) {
                const uint8_t* g = norm + (size_t)y * net_w;
                const uint32_t* rb = red.data() + (size_t)y * rp;
                for (int x = 0; x < net_w; ++x) {
                    const uint8_t v = g[x];
                    dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = v;
                }
                for (int wq = 0; wq < rp; ++wq)
                    for (uint32_t m = rb[wq]; m; m &= m - 1) {
                        const int x = wq * 32 + __builtin_ctz(m);
                        dst[3 * x] = 255; dst[3 * x + 1] = 0; dst[3 * x + 2] = 0;
                    }
            }));
            MS_REQUIRE(ok_overlay, MS_ERR_IO, "Fail to Save Overlay PNG: " + overlay_path);
            out << "Overlay Image Saved to: " << overlay_path << "\n";                        // :193
            const std::string out_json = J.dir + "/" + J.base + ".json";                      // :206
            std::ofstream f(out_json, std::ios::binary);
            MS_REQUIRE(f.good(), MS_ERR_IO, "Fail to Create JSON File: " + out_json);
            f << json::labelme_text(B.xy.data(), cstart, nc, J.base, w, hgt);                  // :207
            out << "JSON Saved to: " << out_json << "\n";                                     // :208
        }
        join_helpers();
        MS_REQUIRE(ok_norm, MS_ERR_IO, "imwrite failed: " + png_path);
        MS_REQUIRE(ok_mask, MS_ERR_IO, "Failed to save mask");
        J.console = out.str();
    } catch (const Error& e) {
        join_helpers();
        J.status = e.code;
        J.error = e.what;
    } catch (const std::exception& e) {
        join_helpers();
        J.status = MS_ERR_INTERNAL;
        J.error = e.what();
    }
}

// writer threads of one pipeline: MEDSEG_WRITERS, else every host core, shared evenly when several GPUs of one handle
// run their pipelines side by side
int writer_threads(int share = 1) {
    if (const char* e = std::getenv("MEDSEG_WRITERS")) return std::max(1, std::atoi(e));
    const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
    return std::max(2, (int)hc / std::max(1, share));
}

// Reads batch `k` of the job list into B.in (good slices packed), assigning slots.  Runs on the prefetch thread.
void read_batch(SliceJob* jobs, size_t first, size_t last, BatchHost& B, int w, int hgt) {
    const size_t n_in = (size_t)w * hgt;
    int slot = 0;
    for (size_t i = first; i < last; ++i) {
        SliceJob& J = jobs[i];
        // the reference maps w*h*2 bytes of the headerless file from offset 0 (src/preprocess.cpp:28-61, :86)
        FILE* f = std::fopen(J.raw.c_str(), "rb");
        if (!f) {
            J.status = MS_ERR_IO;
            J.error = "open failed: " + J.raw;
            continue;
        }
        const size_t got = std::fread(B.in.as<uint16_t>() + (size_t)slot * n_in, 2, n_in, f);
        std::fclose(f);
        if (got != n_in) {
            J.status = MS_ERR_IO;
            J.error = "file shorter than width*height*2 bytes: " + J.raw;
            continue;
        }
        J.slot = slot++;
    }
}

// The batched file pipeline: prefetch thread (disk -> pinned) | GPU (this thread) | writer threads (artefacts), each
// working on a different batch.  Per-file failures are recorded in the jobs; only CUDA / argument errors throw.
// `console`: where the reference's std::cout / std::cerr lines go -- straight out (null), or into a string the caller prints
// in job order (a GPU of a multi-device handle must not interleave its lines with the others')
void process_jobs_on(ms_handle* h, SliceJob* jobs, size_t n, int w, int hgt, bool report_errors, std::string* console, int share = 1) {
    MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded (handle was created without a weight blob)");
    MS_REQUIRE(w > 0 && hgt > 0, MS_ERR_ARG, "bad slice size");
    const size_t mb = (size_t)h->max_batch;
    const size_t n_batches = (n + mb - 1) / mb;
    const size_t n_in = (size_t)w * hgt, npx = (size_t)h->net_w * h->net_h;
    BatchHost* host = h->file_host;   // kept across calls: pinned allocations cost milliseconds
    for (int b = 0; b < 2; ++b) {
        BatchHost& B = host[b];
        const size_t cap = std::min(n, mb);
        B.in.reserve(cap * n_in * 2);
        B.norm.reserve(cap * npx);
        B.mask.reserve(cap * npx);
    }
    std::thread reader;
    std::vector<std::thread> writers;
    std::atomic<size_t> next_job{0};
    auto range = [&](size_t k) { return std::make_pair(k * mb, std::min(n, (k + 1) * mb)); };
    auto join_writers = [&] {
        for (auto& t : writers) t.join();
        writers.clear();
    };
    auto report = [&](size_t k, long long infer_ms) {     // console + log lines of batch k, in file order
        const auto r = range(k);
        for (size_t i = r.first; i < r.second; ++i) {
            const SliceJob& J = jobs[i];
            h->log("\n=== Processing Image: " + basename_of(J.raw) + " ===");                 // src/process.cpp:198
            if (J.status != MS_OK) {
                if (report_errors) {
                    if (console) *console += "\x01Processing error: " + J.error + "\n";        // \x01: a std::cerr line
                    else std::cerr << "Processing error: " << J.error << std::endl;           // :257
                    h->log("error: " + J.error);                                              // :259
                }
                continue;
            }
            h->log("Inference time: " + std::to_string(infer_ms) + " ms");                    // :228 (the whole batch)
            if (console) *console += J.console;
            else std::cout << J.console;
            h->log("Processing completed for: " + J.base);                                    // :250
        }
    };
    long long infer_ms[2] = {0, 0};
    try {
        read_batch(jobs, range(0).first, range(0).second, host[0], w, hgt);
        for (size_t k = 0; k < n_batches; ++k) {
            BatchHost& B = host[k & 1];
            if (k + 1 < n_batches) {
                const auto r = range(k + 1);
                reader = std::thread(read_batch, jobs, r.first, r.second, std::ref(host[(k + 1) & 1]), w, hgt);
            }
            const auto r = range(k);
            int nb = 0;
            for (size_t i = r.first; i < r.second; ++i) nb += jobs[i].slot >= 0;
            if (nb > 0) {
                const auto t0 = std::chrono::high_resolution_clock::now();
                h->d_src.reserve((size_t)nb * n_in * 2);
                MS_CUDA(cudaMemcpyAsync(h->d_src.p, B.in.p, (size_t)nb * n_in * 2, cudaMemcpyHostToDevice, h->stream));
                pipeline_dev(h, h->d_src.as<uint16_t>(), w, hgt, nb, h->stream);
                const long long* hh = h->h_header.as<long long>();
                B.slice_start.resize((size_t)nb + 1);
                B.cstart.resize((size_t)hh[0] + 1);
                B.xy.resize((size_t)std::max<long long>(hh[1], 1) * 2);
                B.uxy.resize(B.xy.size());
                ms_polygons pg{B.xy.data(), (int64_t)B.xy.size() / 2, B.cstart.data(), hh[0], B.slice_start.data(), 0, 0};
                MS_CUDA(cudaMemcpyAsync(B.norm.p, h->d_norm.p, (size_t)nb * npx, cudaMemcpyDeviceToHost, h->stream));
                MS_CUDA(cudaMemcpyAsync(B.mask.p, h->d_mask.p, (size_t)nb * npx, cudaMemcpyDeviceToHost, h->stream));
                copy_polygons_out(h, nb, &pg, h->stream);
                if (hh[1] > 0) {   // the overlay is drawn in network space: emit once more with the identity mapping
                    m2p_phase_b(h->m2p, h->m2p.poly, h->net_h, h->net_w, nb, h->net_w, h->net_h, h->stream);
                    download_sync(h, B.uxy.data(), h->m2p.poly.xy.p, (size_t)hh[1] * 8, h->stream);
                }
                infer_ms[k & 1] = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t0).count();
            }
            join_writers();                                   // batch k - 1 is on disk
            if (k > 0) report(k - 1, infer_ms[(k - 1) & 1]);
            next_job = r.first;
            const int nt = std::min<int>(writer_threads(share), std::max(nb, 1));
            const int spare = writer_threads(share) / std::max(nb, 1) - 1;   // helper threads each slice may use
            for (int t = 0; t < nt; ++t)
                writers.emplace_back([&, r, k, spare] {
                    for (size_t i = next_job.fetch_add(1); i < r.second; i = next_job.fetch_add(1))
                        if (jobs[i].slot >= 0) write_artefacts(jobs[i], host[k & 1], w, hgt, h->net_w, h->net_h, h->fg_value, spare);
                });
            if (reader.joinable()) reader.join();
        }
        join_writers();
        if (n_batches > 0) report(n_batches - 1, infer_ms[(n_batches - 1) & 1]);
    } catch (...) {
        if (reader.joinable()) reader.join();
        join_writers();
        throw;
    }
}

// The per-file loop of src/main.cpp:148-164 over every GPU of the handle: contiguous blocks of the (sorted) job list, one
// host thread per GPU, each running the prefetch | GPU | writers pipeline on its own handle.  Console lines come out in
// job order, block after block.
void process_jobs(ms_handle* h, std::vector<SliceJob>& jobs, int w, int hgt, bool report_errors = true) {
    std::vector<ms_handle*> devs{h};
    devs.insert(devs.end(), h->peers.begin(), h->peers.end());
    const int G = (int)std::min<size_t>(devs.size(), std::max<size_t>(1, (jobs.size() + (size_t)h->max_batch - 1) / (size_t)h->max_batch));
    if (G <= 1) {
        process_jobs_on(h, jobs.data(), jobs.size(), w, hgt, report_errors, nullptr);
        return;
    }
    struct Part {
        std::string console, error;
        int status = MS_OK;
    };
    std::vector<Part> parts((size_t)G);
    auto run = [&](int g) {
        const auto r = block_of((int64_t)jobs.size(), G, g);
        ms_handle* d = devs[g];
        try {
            MS_CUDA(cudaSetDevice(d->device));
            g_counter = &d->counter;
            process_jobs_on(d, jobs.data() + r.first, (size_t)(r.second - r.first), w, hgt, report_errors, &parts[g].console, G);
        } catch (const Error& e) {
            parts[g].status = e.code;
            parts[g].error = e.what;
        } catch (const std::exception& e) {
            parts[g].status = MS_ERR_INTERNAL;
            parts[g].error = e.what();
        }
        g_counter = nullptr;
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; ++g) th.emplace_back(run, g);
    run(0);
    for (auto& t : th) t.join();
    MS_CUDA(cudaSetDevice(h->device));
    g_counter = &h->counter;
    for (const auto& P : parts) {             // lines in job order; \x01 marks a std::cerr line
        size_t i = 0;
        while (i < P.console.size()) {
            const size_t e = P.console.find('\n', i);
            const size_t end = e == std::string::npos ? P.console.size() : e + 1;
            if (P.console[i] == '\x01') std::cerr << P.console.substr(i + 1, end - i - 1) << std::flush;
            else std::cout << P.console.substr(i, end - i);
            i = end;
        }
    }
    for (int g = 0; g < G; ++g)
        if (parts[g].status != MS_OK) fail(parts[g].status, "device " + std::to_string(devs[g]->device) + ": " + parts[g].error);
}

bool is_16bit_image(const std::string& path) {               // src/main.cpp:18-25
    std::string ext = std::filesystem::path(path).extension().string();
    for (auto& c : ext) c = (char)std::tolower((unsigned char)c);
    return ext == ".raw" || ext == ".dcm" || ext == ".tif" || ext == ".tiff";
}

}  // namespace

int ms_process_raw_file(ms_handle* h, const char* raw_path, int w, int hgt, const char* out_dir) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(raw_path && out_dir && w > 0 && hgt > 0, MS_ERR_ARG, "process_raw_file: bad argument");
        const auto t0 = std::chrono::high_resolution_clock::now();
        std::vector<SliceJob> jobs(1);
        jobs[0].raw = raw_path;
        jobs[0].dir = out_dir;
        jobs[0].base = stem_of(jobs[0].raw);                                              // src/process.cpp:201
        // process_jobs reports per-file failures on std::cerr itself; here the failure is also the call's status
        process_jobs(h, jobs, w, hgt, false);
        if (jobs[0].status != MS_OK) fail(jobs[0].status, jobs[0].error);
        const auto total_ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t0).count();
        h->log("Total processing time: " + std::to_string(total_ms) + " ms");            // src/process.cpp:249
        std::cout << "Total processing time: " << total_ms << " ms" << std::endl;         // :253
    });
}

int ms_process_raw_files(ms_handle* h, const char* const* raw_paths, const char* const* out_dirs, int64_t n, int w, int hgt, uint8_t* ok,
                         int64_t* n_ok, int64_t* n_failed) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(n >= 0 && (n == 0 || (raw_paths && out_dirs)) && w > 0 && hgt > 0, MS_ERR_ARG, "process_raw_files: bad argument");
        std::vector<SliceJob> jobs((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            MS_REQUIRE(raw_paths[i] && out_dirs[i], MS_ERR_ARG, "process_raw_files: null path");
            jobs[i].raw = raw_paths[i];
            jobs[i].dir = out_dirs[i];
            jobs[i].base = stem_of(jobs[i].raw);
        }
        const auto t0 = std::chrono::high_resolution_clock::now();
        process_jobs(h, jobs, w, hgt);
        int64_t good = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (ok) ok[i] = jobs[i].status == MS_OK;
            good += jobs[i].status == MS_OK;
        }
        if (n_ok) *n_ok = good;
        if (n_failed) *n_failed = n - good;
        const auto total_ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t0).count();
        h->log("Total processing time: " + std::to_string(total_ms) + " ms (" + std::to_string(n) + " files)");
    });
}

int ms_process_directory(ms_handle* h, const char* input_dir, int w, int hgt, const char* out_dir, int recursive, int shard_index,
                         int shard_count, int64_t* n_found, int64_t* n_ok, int64_t* n_failed) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        namespace fs = std::filesystem;
        MS_REQUIRE(input_dir && out_dir && w > 0 && hgt > 0 && shard_count >= 1 && shard_index >= 0 && shard_index < shard_count, MS_ERR_ARG,
                   "process_directory: bad argument");
        std::error_code ec;
        MS_REQUIRE(fs::is_directory(input_dir, ec), MS_ERR_IO, std::string("not a directory: ") + input_dir);
        // src/main.cpp:28-48; sorted so that every rank sees the same list
        std::vector<std::string> files;
        if (recursive) {
            for (fs::recursive_directory_iterator it(input_dir, fs::directory_options::skip_permission_denied, ec), end; it != end && !ec; it.increment(ec))
                if (it->is_regular_file(ec) && is_16bit_image(it->path().string())) files.push_back(it->path().string());
        } else {
            for (fs::directory_iterator it(input_dir, fs::directory_options::skip_permission_denied, ec), end; it != end && !ec; it.increment(ec))
                if (it->is_regular_file(ec) && is_16bit_image(it->path().string())) files.push_back(it->path().string());
        }
        std::sort(files.begin(), files.end());
        if (n_found) *n_found = (int64_t)files.size();
        // contiguous block of the list per shard; earlier shards take the remainder
        const size_t total = files.size();
        const size_t base = total / (size_t)shard_count, rem = total % (size_t)shard_count;   // sharding.py: shard_range
        const size_t lo = (size_t)shard_index * base + std::min<size_t>((size_t)shard_index, rem), hi = lo + base + ((size_t)shard_index < rem);
        std::vector<SliceJob> jobs(hi - lo);
        for (size_t i = lo; i < hi; ++i) {
            SliceJob& J = jobs[i - lo];
            J.raw = files[i];
            J.base = stem_of(J.raw);
            J.dir = out_dir;
            if (recursive) {                                                               // src/main.cpp:151-156: keep the tree
                const std::string rel = fs::relative(fs::path(J.raw), fs::path(input_dir), ec).parent_path().string();
                if (!rel.empty()) J.dir = (fs::path(out_dir) / rel).string();
            }
        }
        mkdirs(out_dir);                                                                   // src/main.cpp:130
        process_jobs(h, jobs, w, hgt);
        int64_t good = 0;
        for (const auto& J : jobs) good += J.status == MS_OK;
        if (n_ok) *n_ok = good;
        if (n_failed) *n_failed = (int64_t)jobs.size() - good;
    });
}

// ---------------------------------------------------------------- asynchronous double-buffered pipeline
namespace {

// the slot's kernel chain K1 .. K6 (no host round trip: capturable)
void enqueue_chain(ms_handle* h, ms_handle::Slot& S, int w, int hgt, int batch, cudaStream_t st) {
    uint8_t* norm = h->d_norm.as<uint8_t>();
    uint8_t* raw = h->d_mask_raw.as<uint8_t>();
    uint8_t* mask = h->d_mask.as<uint8_t>();
    preprocess_launch(h->pre, S.d_src.as<uint16_t>(), w, hgt, batch, h->net_w, h->net_h, norm, nullptr, st);
    h->unet.forward(norm, batch, raw, nullptr, st);
    post_poly_phase_a(h->post, h->m2p, S.poly, raw, mask, h->net_h, h->net_w, batch, h->fg_value, h->min_area_ratio, st);
    m2p_phase_b(h->m2p, S.poly, h->net_h, h->net_w, batch, w, hgt, st);
}

// pinned result buffers follow the device capacities
void size_slot_host(ms_handle* h, ms_handle::Slot& S) {
    S.h_header.reserve(4 * sizeof(long long));
    S.h_slice_start.reserve(((size_t)h->max_batch + 1) * 4);
    S.h_cstart.reserve(((size_t)S.poly.cap_contours + 1) * 4);
    S.h_xy.reserve((size_t)S.poly.cap_points * 8);
}

// header + slice offsets + the first spec_contours + 1 offsets and spec_points vertices, on the D2H stream
void enqueue_slot_d2h(ms_handle* h, ms_handle::Slot& S, cudaStream_t ds) {
    download(h, S.h_header.p, S.poly.header.p, 4 * sizeof(long long), ds);
    download(h, S.h_slice_start.p, S.poly.slice_start.p, ((size_t)S.batch + 1) * 4, ds);
    download(h, S.h_cstart.p, S.poly.npts.p, ((size_t)S.spec_contours + 1) * 4, ds);
    if (S.spec_points > 0) download(h, S.h_xy.p, S.poly.xy.p, (size_t)S.spec_points * 8, ds);
}

}  // namespace

int ms_submit_batch_host(ms_handle* h, int slot, const uint16_t* h_src, int w, int hgt, int batch) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(slot >= 0 && slot < 2 && h_src && w > 0 && hgt > 0, MS_ERR_ARG, "submit_batch: bad argument");
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        MS_REQUIRE(batch >= 1 && batch <= h->max_batch, MS_ERR_ARG, "batch exceeds max_batch of this handle");
        ms_handle::Slot& S = h->slots[slot];
        MS_REQUIRE(!S.busy, MS_ERR_STATE, "submit_batch: slot still holds an uncollected batch (call ms_wait_batch)");
        if (!h->copy_stream) {
            MS_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            MS_CUDA(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        }
        if (!S.ev_h2d) {
            MS_CUDA(cudaEventCreateWithFlags(&S.ev_h2d, cudaEventDisableTiming));
            MS_CUDA(cudaEventCreateWithFlags(&S.ev_m2p, cudaEventDisableTiming));
            MS_CUDA(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
        }
        // initial capacities: postprocess keeps at most floor(1 / min_area_ratio) components per slice (the
        // ratio is configurable, hence derived, not a literal) and CT-like borders have ~10^3 vertices; a batch that needs
        // more grows the slot inside ms_wait_batch
        if (S.poly.cap_contours == 0) {
            const double r = (double)h->min_area_ratio;
            const int64_t per_slice = r > 0.0 ? std::min<int64_t>(4096, (int64_t)(1.0 / r) + 1) : 4096;
            S.poly.cap_contours = (int64_t)h->max_batch * std::max<int64_t>(64, per_slice);
            S.poly.cap_points = (int64_t)h->max_batch * 8192;
            size_slot_host(h, S);
        }
        const size_t in_bytes = (size_t)w * hgt * 2 * batch;
        S.d_src.reserve(in_bytes);
        const void* src = h_src;
        if (!is_cuda_host_ptr(h_src)) {          // pageable memory: stage through this slot's pinned buffer
            S.h_src.reserve(in_bytes);
            std::memcpy(S.h_src.p, h_src, in_bytes);
            src = S.h_src.p;
        }
        h->h2d_bytes += (int64_t)in_bytes;
        MS_CUDA(cudaMemcpyAsync(S.d_src.p, src, in_bytes, cudaMemcpyHostToDevice, h->copy_stream));
        MS_CUDA(cudaEventRecord(S.ev_h2d, h->copy_stream));
        cudaStream_t st = h->stream;
        MS_CUDA(cudaStreamWaitEvent(st, S.ev_h2d, 0));
        // The chain has no host round trip, so it is captured once per (slot, shape) and replayed: the reference does the
        // same for its inference (cudaGraphLaunch, src/process.cpp:147).  First call with a shape: eager (sizes every
        // buffer).  Second call: capture + launch.  Later calls: launch.  Any reallocation bumps the epoch and drops it.
        const bool same_graph = S.gexec && S.g_w == w && S.g_h == hgt && S.g_batch == batch && S.g_epoch == g_alloc_epoch;
        const bool same_eager = S.seen_w == w && S.seen_h == hgt && S.seen_batch == batch && S.seen_epoch == g_alloc_epoch;
        if (h->graphs_on && same_graph) {
            MS_CUDA(cudaGraphLaunch(S.gexec, st));
            h->counter.n += S.g_launches;
        } else if (h->graphs_on && same_eager) {
            if (S.gexec) {
                cudaGraphExecDestroy(S.gexec);
                S.gexec = nullptr;
            }
            const int64_t l0 = h->counter.n;
            const uint64_t e0 = g_alloc_epoch;
            cudaGraph_t graph = nullptr;
            MS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            try {
                enqueue_chain(h, S, w, hgt, batch, st);
            } catch (...) {
                cudaStreamEndCapture(st, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            MS_CUDA(cudaStreamEndCapture(st, &graph));
            const int64_t captured = h->counter.n - l0;
            h->counter.n = l0;
            if (g_alloc_epoch == e0 && cudaGraphInstantiate(&S.gexec, graph, 0) == cudaSuccess) {
                S.g_w = w; S.g_h = hgt; S.g_batch = batch; S.g_epoch = e0; S.g_launches = captured;
                cudaGraphDestroy(graph);
                MS_CUDA(cudaGraphLaunch(S.gexec, st));
                h->counter.n += captured;
            } else {            // a buffer moved while capturing (or instantiation failed): run this batch eagerly
                cudaGetLastError();
                S.gexec = nullptr;
                if (graph) cudaGraphDestroy(graph);
                enqueue_chain(h, S, w, hgt, batch, st);
                S.seen_epoch = g_alloc_epoch;
            }
        } else {
            enqueue_chain(h, S, w, hgt, batch, st);
            S.seen_w = w; S.seen_h = hgt; S.seen_batch = batch; S.seen_epoch = g_alloc_epoch;
        }
        // results leave on their own stream so the next batch's kernels are not held up.  Their sizes are not known on the
        // host yet: copy what the previous batch of this slot needed plus a quarter, fetch any remainder in ms_wait_batch.
        S.batch = batch; S.w = w; S.hgt = hgt;
        S.spec_contours = std::min<int64_t>(S.poly.cap_contours, std::max<int64_t>(S.last_contours + S.last_contours / 4 + 8, 2 * (int64_t)batch));
        S.spec_points = std::min<int64_t>(S.poly.cap_points, std::max<int64_t>(S.last_points + S.last_points / 4 + 256, 256 * (int64_t)batch));
        MS_CUDA(cudaEventRecord(S.ev_m2p, st));
        cudaStream_t ds = h->d2h_stream;
        MS_CUDA(cudaStreamWaitEvent(ds, S.ev_m2p, 0));
        enqueue_slot_d2h(h, S, ds);
        MS_CUDA(cudaEventRecord(S.ev_done, ds));
        S.busy = true;
    });
}

int ms_wait_batch(ms_handle* h, int slot, ms_polygons* out) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(slot >= 0 && slot < 2, MS_ERR_ARG, "wait_batch: bad slot");
        ms_handle::Slot& S = h->slots[slot];
        MS_REQUIRE(S.busy, MS_ERR_STATE, "wait_batch: nothing was submitted on this slot");
        check_polys(out, S.batch);
        MS_CUDA(cudaEventSynchronize(S.ev_done));
        const long long* hh = S.h_header.as<long long>();
        cudaStream_t ds = h->d2h_stream;
        // the batch did not fit the slot's device capacities: grow them (and the pinned buffers) and run the slot's chain
        // again on its still-resident input -- stream order keeps the other slot's in-flight batch intact
        for (int attempt = 0; hh[0] > S.poly.cap_contours || hh[1] > S.poly.cap_points || (hh[2] & 3); ++attempt) {
            if (attempt >= 3 || hh[3] != 0 || (hh[2] & 4)) S.busy = false;
            MS_REQUIRE(attempt < 3, MS_ERR_INTERNAL, "async pipeline: capacity growth did not converge");
            MS_REQUIRE(hh[3] == 0 && (hh[2] & 4) == 0, MS_ERR_INTERNAL, "mask2polygon failed on an oversized batch");
            if (hh[0] > S.poly.cap_contours) S.poly.cap_contours = hh[0] + hh[0] / 4 + 64;
            if (hh[1] > S.poly.cap_points) S.poly.cap_points = hh[1] + hh[1] / 4 + 1024;
            size_slot_host(h, S);
            enqueue_chain(h, S, S.w, S.hgt, S.batch, h->stream);
            MS_CUDA(cudaEventRecord(S.ev_m2p, h->stream));
            MS_CUDA(cudaStreamWaitEvent(ds, S.ev_m2p, 0));
            S.spec_contours = S.spec_points = 0;
            enqueue_slot_d2h(h, S, ds);
            MS_CUDA(cudaStreamSynchronize(ds));
            hh = S.h_header.as<long long>();
        }
        if (hh[3] != 0) S.busy = false;
        MS_REQUIRE(hh[3] == 0, MS_ERR_INTERNAL, "mask2polygon: border following did not terminate");
        out->n_contours = hh[0];
        out->n_points = hh[1];
        // (the slot stays collectable: call again with the sizes just reported)
        MS_REQUIRE(out->cap_contours >= hh[0] && out->cap_points >= hh[1], MS_ERR_CAPACITY,
                   "polygon buffers too small: need " + std::to_string(hh[0]) + " contours, " + std::to_string(hh[1]) + " points");
        bool more = false;
        if (hh[0] > S.spec_contours) {
            download(h, S.h_cstart.as<int32_t>() + S.spec_contours + 1, S.poly.npts.as<int32_t>() + S.spec_contours + 1,
                     (size_t)(hh[0] - S.spec_contours) * 4, ds);
            more = true;
        }
        if (hh[1] > S.spec_points) {
            download(h, S.h_xy.as<int32_t>() + 2 * S.spec_points, S.poly.xy.as<int32_t>() + 2 * S.spec_points,
                     (size_t)(hh[1] - S.spec_points) * 8, ds);
            more = true;
        }
        if (more) MS_CUDA(cudaStreamSynchronize(ds));
        S.busy = false;
        S.last_contours = hh[0];
        S.last_points = hh[1];
        std::memcpy(out->slice_start, S.h_slice_start.p, ((size_t)S.batch + 1) * 4);
        std::memcpy(out->contour_start, S.h_cstart.p, ((size_t)hh[0] + 1) * 4);
        if (hh[1] > 0) std::memcpy(out->xy, S.h_xy.p, (size_t)hh[1] * 8);
    });
}

// ---------------------------------------------------------------- one volume, all GPUs of the handle (cfg3)
namespace {

struct VolumePart {                     // one GPU's contiguous block of slices
    ms_handle* dev = nullptr;
    int64_t first = 0, count = 0;
    std::vector<int32_t> per_slice;     // contours per slice
    std::vector<int32_t> per_contour;   // vertices per contour
    std::vector<int32_t> xy;
    int status = MS_OK;
    std::string error;
};

// the double-buffered batcher over one block: submit sub-batch k + 1, collect sub-batch k
void volume_worker(VolumePart& V, const uint16_t* h_src, int w, int hgt, uint8_t* h_norm, uint8_t* h_mask) {
    ms_handle* d = V.dev;
    const int64_t mb = d->max_batch, n_sub = (V.count + mb - 1) / mb;
    const size_t n_in = (size_t)w * hgt, npx = (size_t)d->net_w * d->net_h;
    std::vector<int32_t> ss((size_t)mb + 1), cs, xy;
    int64_t cap_c = mb * 64, cap_p = mb * 8192;
    cs.resize((size_t)cap_c + 1);
    xy.resize((size_t)cap_p * 2);
    auto fail_with = [&](int rc) {
        V.status = rc;
        V.error = ms_last_error(d);
    };
    auto sub = [&](int64_t k) { return std::make_pair(k * mb, std::min(V.count, (k + 1) * mb)); };
    auto append = [&](const ms_polygons& pg, int nb) {
        for (int i = 0; i < nb; ++i) V.per_slice.push_back(ss[i + 1] - ss[i]);
        for (int64_t c = 0; c < pg.n_contours; ++c) V.per_contour.push_back(cs[c + 1] - cs[c]);
        V.xy.insert(V.xy.end(), xy.begin(), xy.begin() + 2 * pg.n_points);
    };
    if (h_norm || h_mask) {               // side outputs wanted: the synchronous call per sub-batch returns them
        for (int64_t k = 0; k < n_sub; ++k) {
            const auto r = sub(k);
            const int nb = (int)(r.second - r.first);
            const size_t g0 = (size_t)(V.first + r.first);
            for (;;) {
                ms_polygons pg{xy.data(), cap_p, cs.data(), cap_c, ss.data(), 0, 0};
                const int rc = ms_process_batch_host(d, h_src + g0 * n_in, w, hgt, nb, &pg, h_norm ? h_norm + g0 * npx : nullptr,
                                                     h_mask ? h_mask + g0 * npx : nullptr);
                if (rc == MS_ERR_CAPACITY && (pg.n_points > cap_p || pg.n_contours > cap_c)) {
                    cap_p = std::max(cap_p, pg.n_points + pg.n_points / 4 + 16);
                    cap_c = std::max(cap_c, pg.n_contours + pg.n_contours / 4 + 16);
                    cs.resize((size_t)cap_c + 1);
                    xy.resize((size_t)cap_p * 2);
                    continue;
                }
                if (rc != MS_OK) return fail_with(rc);
                append(pg, nb);
                break;
            }
        }
        return;
    }
    auto submit = [&](int64_t k) {
        const auto r = sub(k);
        return ms_submit_batch_host(d, (int)(k & 1), h_src + (size_t)(V.first + r.first) * n_in, w, hgt, (int)(r.second - r.first));
    };
    int rc = n_sub > 0 ? submit(0) : MS_OK;
    if (rc != MS_OK) return fail_with(rc);
    for (int64_t k = 0; k < n_sub; ++k) {
        if (k + 1 < n_sub && (rc = submit(k + 1)) != MS_OK) {
            ms_polygons drop{xy.data(), cap_p, cs.data(), cap_c, ss.data(), 0, 0};
            ms_wait_batch(d, (int)(k & 1), &drop);
            return fail_with(rc);
        }
        const auto r = sub(k);
        for (;;) {
            ms_polygons pg{xy.data(), cap_p, cs.data(), cap_c, ss.data(), 0, 0};
            rc = ms_wait_batch(d, (int)(k & 1), &pg);
            if (rc == MS_ERR_CAPACITY && (pg.n_points > cap_p || pg.n_contours > cap_c)) {     // the slot stays collectable
                cap_p = std::max(cap_p, pg.n_points + pg.n_points / 4 + 16);
                cap_c = std::max(cap_c, pg.n_contours + pg.n_contours / 4 + 16);
                cs.resize((size_t)cap_c + 1);
                xy.resize((size_t)cap_p * 2);
                continue;
            }
            if (rc != MS_OK) {
                if (k + 1 < n_sub) {      // drain the sub-batch already in flight
                    ms_polygons drop{xy.data(), cap_p, cs.data(), cap_c, ss.data(), 0, 0};
                    ms_wait_batch(d, (int)((k + 1) & 1), &drop);
                }
                return fail_with(rc);
            }
            append(pg, (int)(r.second - r.first));
            break;
        }
    }
}

}  // namespace

int ms_process_volume_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int64_t n_slices, ms_polygons* out, uint8_t* h_norm_u8,
                           uint8_t* h_mask_u8) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h_src && w > 0 && hgt > 0 && n_slices > 0 && n_slices < ((int64_t)1 << 31), MS_ERR_ARG, "process_volume: bad argument");
        MS_REQUIRE(out && out->slice_start && out->contour_start && (out->xy || out->cap_points == 0), MS_ERR_ARG,
                   "ms_polygons: null buffer (slice_start needs n_slices + 1 entries)");
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        std::vector<ms_handle*> devs{h};
        devs.insert(devs.end(), h->peers.begin(), h->peers.end());
        const int G = (int)std::min<int64_t>((int64_t)devs.size(), n_slices);
        std::vector<VolumePart> parts((size_t)G);
        for (int g = 0; g < G; ++g) {
            const auto r = block_of(n_slices, G, g);
            parts[g].dev = devs[g];
            parts[g].first = r.first;
            parts[g].count = r.second - r.first;
        }
        std::vector<std::thread> th;
        for (int g = 1; g < G; ++g) th.emplace_back(volume_worker, std::ref(parts[g]), h_src, w, hgt, h_norm_u8, h_mask_u8);
        volume_worker(parts[0], h_src, w, hgt, h_norm_u8, h_mask_u8);
        for (auto& t : th) t.join();
        MS_CUDA(cudaSetDevice(h->device));
        g_counter = &h->counter;
        for (const auto& V : parts)
            if (V.status != MS_OK) fail(V.status, "device " + std::to_string(V.dev->device) + ": " + V.error);
        // concatenate in slice order
        int64_t nc = 0, np = 0;
        for (const auto& V : parts) {
            nc += (int64_t)V.per_contour.size();
            np += (int64_t)V.xy.size() / 2;
        }
        out->n_contours = nc;
        out->n_points = np;
        MS_REQUIRE(out->cap_contours >= nc && out->cap_points >= np, MS_ERR_CAPACITY,
                   "polygon buffers too small: need " + std::to_string(nc) + " contours, " + std::to_string(np) + " points");
        int64_t s = 0, c = 0, p = 0;
        out->slice_start[0] = 0;
        out->contour_start[0] = 0;
        for (const auto& V : parts) {
            for (int32_t k : V.per_slice) {
                out->slice_start[s + 1] = out->slice_start[s] + k;
                ++s;
            }
            for (int32_t k : V.per_contour) {
                out->contour_start[c + 1] = out->contour_start[c] + k;
                ++c;
            }
            if (!V.xy.empty()) std::memcpy(out->xy + 2 * p, V.xy.data(), V.xy.size() * 4);
            p += (int64_t)V.xy.size() / 2;
        }
    });
}

int ms_set_dp_epsilon(ms_handle* h, double eps) {
    if (!h || !(eps >= 0.0 && eps < 1e9)) return MS_ERR_ARG;
    h->m2p.dp_eps = eps;
    for (ms_handle* c : h->peers) c->m2p.dp_eps = eps;
    ++g_alloc_epoch;        // captured per-slot graphs hold the old launch sequence: capture again
    return MS_OK;
}

double ms_dp_epsilon(ms_handle* h) { return h ? h->m2p.dp_eps : 0.0; }

// ---------------------------------------------------------------- instrumentation
int64_t ms_launch_count(ms_handle* h) {
    if (!h) return 0;
    int64_t n = h->counter.n;
    for (const ms_handle* c : h->peers) n += c->counter.n;
    return n;
}
int ms_device_count(ms_handle* h) { return h ? 1 + (int)h->peers.size() : 0; }
int ms_layer_count(ms_handle* h) { return h && h->unet.loaded() ? (int)h->unet.layers().size() : 0; }
const char* ms_layer_name(ms_handle* h, int layer) {
    if (!h || !h->unet.loaded() || layer < 0 || layer >= (int)h->unet.layers().size()) return "";
    return h->unet.layers()[layer].name.c_str();
}
const char* ms_layer_kernel(ms_handle* h, int layer) {
    static thread_local std::string name;
    if (!h || !h->unet.loaded() || layer < 0 || layer >= (int)h->unet.layers().size()) return "";
    name = h->unet.layers()[layer].kernel_name();
    return name.c_str();
}
int ms_time_layer(ms_handle* h, int layer, int batch, int iters, float* ms_per_launch, double* flops) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        MS_REQUIRE(iters >= 1 && ms_per_launch, MS_ERR_ARG, "time_layer: bad argument");
        cudaEvent_t e0, e1;
        MS_CUDA(cudaEventCreate(&e0));
        MS_CUDA(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) h->unet.run_layer(layer, h->d_norm.as<uint8_t>(), batch, h->d_mask_raw.as<uint8_t>(), nullptr, h->stream);
        MS_CUDA(cudaEventRecord(e0, h->stream));
        for (int i = 0; i < iters; ++i) h->unet.run_layer(layer, h->d_norm.as<uint8_t>(), batch, h->d_mask_raw.as<uint8_t>(), nullptr, h->stream);
        MS_CUDA(cudaEventRecord(e1, h->stream));
        MS_CUDA(cudaStreamSynchronize(h->stream));
        float ms = 0;
        MS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms_per_launch = ms / iters;
        if (flops) *flops = h->unet.layers()[layer].flops_per_slice * batch;
    });
}

int ms_profile_layers_begin(ms_handle* h, int max_forwards) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        MS_REQUIRE(max_forwards >= 0 && max_forwards <= 4096, MS_ERR_ARG, "profile_layers_begin: bad argument");
        h->unet.profile_begin(max_forwards);
        while (h->stage_events.size() < (size_t)max_forwards * 5) {
            cudaEvent_t e;
            MS_CUDA(cudaEventCreate(&e));
            h->stage_events.push_back(e);
        }
        h->stage_cap = max_forwards;
        h->stage_n = 0;
    });
}
int ms_profile_stages_read(ms_handle* h, float* ms4, int* n_passes) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(ms4, MS_ERR_ARG, "profile_stages_read: bad argument");
        const int n = h->stage_n;
        for (int k = 0; k < 4; ++k) ms4[k] = 0.0f;
        for (int f = 0; f < n; ++f) {
            cudaEvent_t* ev = h->stage_events.data() + (size_t)f * 5;
            MS_CUDA(cudaEventSynchronize(ev[4]));
            for (int k = 0; k < 4; ++k) {
                float ms = 0;
                MS_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
                ms4[k] += ms / n;
            }
        }
        if (n_passes) *n_passes = n;
        h->stage_cap = h->stage_n = 0;
    });
}
int ms_profile_layers_read(ms_handle* h, float* ms_per_layer, int n_layers, int* n_forwards) {
    if (!h) return MS_ERR_ARG;
    return guarded(h, [&] {
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        MS_REQUIRE(ms_per_layer && n_layers == (int)h->unet.layers().size(), MS_ERR_ARG, "profile_layers_read: bad argument");
        std::vector<float> v;
        const int passes = h->unet.profile_read(v);
        std::copy(v.begin(), v.end(), ms_per_layer);
        if (n_forwards) *n_forwards = passes;
    });
}

int ms_debug_fused_phases(long long* out32) {
    long long* b = fused_debug_buffer(false);
    if (!b || !out32) return MS_ERR_STATE;
    std::memcpy(out32, b, 32 * sizeof(long long));
    return MS_OK;
}

int64_t ms_debug_read_activation(ms_handle* h, const char* name, int batch, float* h_dst, int64_t cap) {
    if (!h || !name) return MS_ERR_ARG;
    int64_t result = 0;
    int rc = guarded(h, [&] {
        MS_REQUIRE(h->unet.loaded(), MS_ERR_STATE, "no weights loaded");
        const ActBuf* buf = nullptr;
        for (const auto& b : h->unet.buffers())
            if (b.name == name) buf = &b;
        MS_REQUIRE(buf, MS_ERR_ARG, std::string("unknown activation buffer: ") + name);
        const int hh = h->net_h >> buf->level, ww = h->net_w >> buf->level, C = buf->C;
        const size_t n = (size_t)batch * hh * ww * C;
        result = (int64_t)n;
        if (!h_dst || cap < (int64_t)n) return;
        std::vector<__nv_bfloat16> tmp(n);
        MS_CUDA(cudaStreamSynchronize(h->stream));
        MS_CUDA(cudaMemcpy(tmp.data(), buf->p, n * 2, cudaMemcpyDeviceToHost));
        for (int b = 0; b < batch; ++b)          // NHWC bf16 -> NCHW fp32
            for (int y = 0; y < hh; ++y)
                for (int x = 0; x < ww; ++x)
                    for (int c = 0; c < C; ++c)
                        h_dst[(((size_t)b * C + c) * hh + y) * ww + x] = __bfloat162float(tmp[(((size_t)b * hh + y) * ww + x) * C + c]);
    });
    return rc == MS_OK ? result : rc;
}

}  // extern "C"
