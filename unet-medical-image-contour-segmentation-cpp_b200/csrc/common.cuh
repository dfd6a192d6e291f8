// common.cuh -- shared declarations for libmedseg_b200 (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/medseg_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "libmedseg_b200 device code is written for sm_100a only (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif

namespace ms {

struct Error {
    int code;
    std::string what;
};

// Thrown inside the library, caught at every C-ABI entry point (never crosses the boundary).
[[noreturn]] void fail(int code, const std::string& what);

#define MS_CUDA(expr)                                                                               \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            ::ms::fail(MS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                                        __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
    } while (0)

#define MS_REQUIRE(cond, code, msg)                                                                 \
    do {                                                                                            \
        if (!(cond)) ::ms::fail((code), std::string(msg));                                          \
    } while (0)

// Launch bookkeeping: every kernel launch goes through LAUNCH so ms_launch_count is honest.
struct LaunchCounter {
    int64_t n = 0;
};
extern thread_local LaunchCounter* g_counter;

#define MS_LAUNCH_CHECK()                                                                           \
    do {                                                                                            \
        if (::ms::g_counter) ::ms::g_counter->n++;                                                  \
        MS_CUDA(cudaGetLastError());                                                                \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: set it once per (kernel, device), so
// several handles on different GPUs of one process all get it (thread-safe).
template <class Kernel>
inline void set_max_dynamic_smem(Kernel kernel, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    MS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(reinterpret_cast<const void*>(kernel), dev);
    if (done.count(key)) return;
    MS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert(key);
}

// Programmatic dependent launch (PDL).  Every layer kernel announces, right after its set-up, that its successor may be
// scheduled (`launch_dependents`); the successor is launched with programmatic stream serialisation and runs ITS set-up
// (tensor-map prefetch, barrier init, TMEM allocation, resident-weight TMA loads) on whatever SMs the predecessor's tail
// leaves idle, then blocks in `wait` -- which returns once the predecessor grid has completed and its writes are visible --
// before it touches an activation.  All of a kernel's global writes depend on loads issued after its wait, so no write can
// overtake the predecessor either.  MEDSEG_PDL=0 falls back to plain stream order.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("MEDSEG_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
// kernel<<<grid, block, smem, st>>>(args...) with the PDL attribute when `dependent` (the kernel calls pdl_wait())
template <class... KArgs, class... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool dependent, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (dependent && pdl_enabled()) ? 1 : 0;
    MS_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bumped whenever a device buffer is (re)allocated or freed: a captured CUDA graph holds raw pointers, so it is only
// replayed while the epoch it was captured under is still current.
extern std::atomic<uint64_t> g_alloc_epoch;   // process-wide and monotonic, so a stale graph can never look current

// Device-side pipeline watchdog (unet_conv_tc.cuh: mbar_wait): one word of mapped pinned host memory, process-wide.
// A kernel whose mbarrier wait exceeded 20 s of wall time sets it; every C-ABI entry point checks it on return.
unsigned* watchdog_host_word();        // allocates on first use (cudaHostAllocMapped | Portable); never freed

// Simple growable device buffer (never shrinks; growth is outside any timed / captured region).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }     // RAII: a handle (or a half-initialised one) frees everything it reserved
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        ++g_alloc_epoch;
        if (p) MS_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        MS_CUDA(cudaMalloc(&p, bytes));
        cap = bytes;
    }
    void release() {
        if (p) {
            ++g_alloc_epoch;
            cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { release(); }
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) MS_CUDA(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        MS_CUDA(cudaMallocHost(&p, bytes));
        cap = bytes;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------- stage launchers (device ptrs)

// K1 preprocess.cu
struct PreprocessWs {
    DevBuf minmax;  // batch x 2 u32
};
void preprocess_launch(PreprocessWs& ws, const uint16_t* d_src, int w, int h, int batch, int out_w, int out_h,
                       uint8_t* d_out_u8, __nv_bfloat16* d_out_bf16, cudaStream_t st);

// ccl.cu -- union-find labelling shared by K5 and K6
struct CclWs {
    DevBuf labels;  // int32 per pixel
    DevBuf area;    // int32 per pixel (valid at roots)
    DevBuf flag;    // u8 per pixel (valid at roots): component touches the image border
};

// slice_fused.cuh (built into mask2polygon.cu): global fallback run tables of the one-CTA-per-slice kernel, used only by
// slices with more runs than its shared-memory tables hold
struct FusedWs {
    DevBuf tab;           // int32 [batch][3][16 * words per slice]: labels, areas, run head positions
};

// Foreground label per slice of a (possibly virtual) batch.  The reference cleans ONE label (FOREGROUND_VALUE, src/postprocess.cpp:5);
// the multi-label call (cfg4) runs K5 / K6 once over n_labels x batch VIRTUAL slices: virtual slice s reads class mask s % in_mod
// and keeps label v[s / in_mod].  single(): the plain batch.
struct FgSpec {
    int in_mod;
    unsigned char v[16];
    __host__ __device__ int fg(int slice) const { return v[slice / in_mod]; }
    __host__ __device__ int src(int slice) const { return slice % in_mod; }
    static FgSpec single(int fg_value, int batch) {
        FgSpec f{};
        f.in_mod = batch > 0 ? batch : 1;
        f.v[0] = (unsigned char)fg_value;
        return f;
    }
};

// K5 postprocess.cu
struct PostprocessWs {
    CclWs ccl;
    DevBuf bin_a, bin_b;  // u8 per pixel
    FusedWs fused;
};
void postprocess_launch(PostprocessWs& ws, const uint8_t* d_in, uint8_t* d_out, int h, int w, int batch, int fg_value,
                        float min_area_ratio, cudaStream_t st, const FgSpec* multi = nullptr);

// K6 mask2polygon.cu
struct PolyDev {              // device-resident polygon set + counters
    DevBuf starts;            // int32 [cap_contours]  start pixel (slice-local linear index)
    DevBuf start_slice;       // int32 [cap_contours]
    DevBuf npts;              // int32 [cap_contours + 1]  counts, then exclusive offsets
    DevBuf slice_start;       // int32 [batch + 1]
    DevBuf block_counts;      // int32 scratch
    DevBuf xy;                // int32 [cap_points * 2]
    DevBuf header;            // int64 [8]: n_contours, n_points, overflow, trace_error, n_chunks
    DevBuf chunks;            // int2 [cap_chunks * 64]: kept vertices (unmapped) in walk order, 64 per chunk
    DevBuf chunk_meta;        // int2 [cap_chunks]: {contour, sequence number} of each chunk
    // one-CTA-per-slice path (slice_fused.cuh): per-slice totals and bases, per-contour {offset, count}, packed vertices
    DevBuf slice_info;        // int4 [batch]
    DevBuf rec;               // int2 [cap_contours]
    DevBuf vstore;            // u32  [cap_points]  x | y << 16
    bool fused = false;       // phase A ran on the fused path (phase B must finalize accordingly)
    // opt-in Douglas-Peucker (dp_simplify.cuh), reserved only when "dp_epsilon" > 0
    DevBuf dp_tmp;            // int2 [cap_points]: simplified vertices of contour c at its old offset
    DevBuf dp_list;           // uint2 [cap_points]: work lists of contours too long for shared memory
    DevBuf dp_keep;           // u32 bitmaps of those contours
    DevBuf dp_cnt, dp_old;    // int32 [cap_contours (+ 1)]: simplified counts, offsets before simplification
    int64_t cap_contours = 0, cap_points = 0;
    void release() {
        for (DevBuf* b : {&starts, &start_slice, &npts, &slice_start, &block_counts, &xy, &header, &chunks, &chunk_meta, &slice_info, &rec, &vstore,
                          &dp_tmp, &dp_list, &dp_keep, &dp_cnt, &dp_old})
            b->release();
    }
};
struct M2pWs {
    CclWs fg, bg;
    DevBuf fgbits;            // u32 per 32 pixels: bit-packed foreground
    PolyDev poly;
    PinBuf h_header;          // pinned int64[4]
    // parallel contour ordering (large slices, mask2polygon.cu "crack" kernels)
    DevBuf crack_pair;        // u64 per directed crack (4 per pixel): {next crack | contour marker, rank}
    DevBuf crack_pos;         // u32 per crack position of the external contours: pixel index | kept bit
    DevBuf crack_blocks;      // int32 per 1024 positions: kept-vertex counts, then exclusive offsets
    DevBuf crack_contour;     // int32 [2 * (cap_contours + 1)]: position base per contour, rotation
    DevBuf crack_meta;        // int64 total positions, int32 round flags
    FusedWs fused;
    double dp_eps = 0.0;      // > 0: phase B simplifies every contour (cv2.approxPolyDP, closed) before the coordinate mapping
    void release_crack() {
        for (DevBuf* b : {&crack_pair, &crack_pos, &crack_blocks, &crack_contour, &crack_meta}) b->release();
    }
};
// Phase A: labels, external starts (descending raster order per slice), per-contour vertex counts, offsets.
void m2p_phase_a(M2pWs& ws, PolyDev& poly, const uint8_t* d_mask, int h, int w, int batch, int threshold, cudaStream_t st);
// Phase B: emit mapped vertices into ws.poly.xy (requires cap_points >= n_points).
void m2p_phase_b(M2pWs& ws, PolyDev& poly, int h, int w, int batch, int orig_w, int orig_h, cudaStream_t st);

// One CTA per slice, everything in shared memory (slice_fused.cuh): K5 and / or phase A of K6 in one launch, for slices
// whose bit image fits on chip.  MEDSEG_FUSED=0 switches it off (A/B measurements, and the tests run both paths).
bool slice_fused_supported(int h, int w);
long long* fused_debug_buffer(bool create);   // [32] clock64 stamps of slice 0's phases (debug; null unless MEDSEG_FUSED_DBG)
// do_post: d_in = class mask, d_out = clean mask {0, fg} (postprocess_mask);  do_poly: contours of the result (do_post) or
// of d_in > thr (otherwise) staged in `poly` for m2p_phase_b.  `poly` may be null when !do_poly.
void slice_fused_launch(FusedWs& fws, M2pWs* ws, PolyDev* poly, const uint8_t* d_in, uint8_t* d_out, int h, int w, int batch, bool do_post,
                        bool do_poly, int fg_value, float min_area_ratio, int thr, cudaStream_t st, const FgSpec* multi = nullptr);
// K5 + phase A of K6 on the class mask: the fused kernel when the slice fits, else postprocess_launch + m2p_phase_a
void post_poly_phase_a(PostprocessWs& pws, M2pWs& ws, PolyDev& poly, const uint8_t* d_raw, uint8_t* d_clean, int h, int w, int batch,
                       int fg_value, float min_area_ratio, cudaStream_t st, const FgSpec* multi = nullptr);

}  // namespace ms
