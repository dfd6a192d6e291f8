/* medseg_b200.h -- C ABI of libmedseg_b200.so (sm_100a).
 *
 * Drop-in boundary for the per-slice segmentation hot path of
 * Florescence/UNet-Medical-Image-Contour-Segmentation-cpp.  The reference has no C ABI; every entry
 * point below names the reference function (file:line under /root/reference) it replaces.  The C++
 * facade in include/{initialize,process,cleanup,preprocess,postprocess,mask2polygon}.h is a thin
 * layer over these calls and keeps the reference's namespaces and signatures, so the reference's
 * src/main.cpp compiles and links unchanged.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary;
 *   - return 0 (MS_OK) on success, a negative ms_status otherwise; nothing throws;
 *   - `d_` pointers are device memory on the handle's GPU, `h_` pointers are host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the handle's own stream); `_dev` calls are
 *     asynchronous on that stream, `_host` calls are synchronous and do their own H2D / D2H;
 *   - there is NO CPU fallback: every compute call fails with MS_ERR_CUDA when no sm_100 GPU is usable.
 */
#ifndef MEDSEG_B200_H
#define MEDSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MS_API __attribute__((visibility("default")))
#else
#define MS_API
#endif

typedef struct ms_handle ms_handle;

typedef enum ms_status {
    MS_OK = 0,
    MS_ERR_ARG = -1,       /* bad argument / unsupported shape                       */
    MS_ERR_IO = -2,        /* file missing / unreadable / unwritable                 */
    MS_ERR_FORMAT = -3,    /* malformed config JSON or weight blob                   */
    MS_ERR_CUDA = -4,      /* CUDA runtime / driver error, or no sm_100 device       */
    MS_ERR_CAPACITY = -5,  /* caller-provided output buffer too small (see counts)   */
    MS_ERR_STATE = -6,     /* call order (e.g. forward before weights are loaded)    */
    MS_ERR_INTERNAL = -7
} ms_status;

/* Reference literals (defaults of the JSON config): 512 (src/preprocess.cpp:81, src/process.cpp:70),
 * 3 classes (src/process.cpp:162), FOREGROUND_VALUE 2 / 3x3 kernel / MIN_AREA_RATIO 0.06f
 * (src/postprocess.cpp:5-9), threshold 127 (src/mask2polygon.cpp:31). */
typedef struct ms_info {
    int32_t device;          /* CUDA device ordinal                                  */
    int32_t sm_count;
    int32_t net_h, net_w;    /* network input size (512 x 512)                       */
    int32_t n_classes;       /* 3 = reference argmax head, 1 = binary (logit > 0)    */
    int32_t max_batch;       /* slices per launch the workspace was sized for        */
    int32_t foreground_value;
    float   min_area_ratio;
    int32_t has_weights;     /* 0 when created with weights=null (stage-only handle) */
    int64_t n_params;        /* 31,036,611 for the canonical 3-class UNet            */
    int64_t flops_per_slice; /* 2*MAC of conv3x3 + convT + head */
} ms_info;

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Replaces MedicalSeg::initialize_engine(trt_cache_path, log_dir)  (src/initialize.cpp:26-76).
 * `path` is a JSON config ({"weights": "...", "max_batch": 32, ...}) or directly a MSEGW001 weight
 * blob; NULL / "" creates a stage-only handle (preprocess / postprocess / mask2polygon usable,
 * UNet calls return MS_ERR_STATE).  `log_dir` may be NULL (no log file). */
MS_API int ms_init(const char* path, const char* log_dir, ms_handle** out);

/* Same, with the JSON config given as text (keys as in the file form). */
MS_API int ms_init_json(const char* cfg_json_text, const char* log_dir, ms_handle** out);

/* Replaces MedicalSeg::cleanup_resources()  (src/cleanup.cpp:10-64). */
MS_API void ms_destroy(ms_handle* h);

/* Last error text of this handle (or of the calling thread when h == NULL).  Never NULL. */
MS_API const char* ms_last_error(ms_handle* h);

MS_API int ms_get_info(ms_handle* h, ms_info* out);

/* Pinned host memory for the `_host` calls (pageable memory works too, but is staged). */
MS_API void* ms_alloc_pinned(size_t bytes);
MS_API void  ms_free_pinned(void* p);

/* ---- stage 1: preprocess -------------------------------------------------------------------- */

/* Replaces Preprocess::compute_minmax + the resample loop of Preprocess::preprocess_raw
 * (src/preprocess.cpp:65-74, 81-118) and MedicalSeg::preprocess_image (src/process.cpp:22-42).
 * src: `batch` slices of h x w little-endian u16, row-major, contiguous.
 * d_out_u8  : batch x net_h x net_w   u8  (the `_normalized.png` pixels), required
 * d_out_bf16: batch x net_h x net_w   bf16 = float(u8)/255.0f, may be NULL            */
MS_API int ms_preprocess_dev(ms_handle* h, const uint16_t* d_src, int w, int hgt, int batch,
                             uint8_t* d_out_u8, void* d_out_bf16, void* stream);
MS_API int ms_preprocess_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int batch,
                              uint8_t* h_out_u8);

/* ---- stage 2: UNet forward + head ------------------------------------------------------------ */

/* Replaces MedicalSeg::execute_inference (src/process.cpp:123-175): the TensorRT engine launch
 * (:147) and the 3-class first-max argmax (:158-170).
 * d_in_u8 : batch x net_h x net_w u8 (output of ms_preprocess_dev)
 * d_mask  : batch x net_h x net_w u8 class index (binary head: foreground_value / 0), required
 * d_logits: batch x n_classes x net_h x net_w fp32 (the reference's "output" tensor), may be NULL */
MS_API int ms_unet_forward_dev(ms_handle* h, const uint8_t* d_in_u8, int batch, uint8_t* d_mask,
                               float* d_logits, void* stream);
MS_API int ms_unet_forward_host(ms_handle* h, const uint8_t* h_in_u8, int batch, uint8_t* h_mask,
                                float* h_logits);

/* ---- stage 3: postprocess -------------------------------------------------------------------- */

/* Replaces postprocess_mask (src/postprocess.cpp:47-79) incl. fill_holes_inside_foreground
 * (:13-44): hole fill, 3x3 open, 8-connected area filter.  Output values {0, fg_value}.
 * fg_value <= 0 selects the handle's foreground_value (2).  In-place (d_out == d_in) is allowed. */
MS_API int ms_postprocess_dev(ms_handle* h, const uint8_t* d_in, uint8_t* d_out, int hgt, int w,
                              int batch, int fg_value, void* stream);
MS_API int ms_postprocess_host(ms_handle* h, const uint8_t* h_in, uint8_t* h_out, int hgt, int w,
                               int batch, int fg_value);

/* ---- stage 4: mask2polygon ------------------------------------------------------------------- */

/* Polygon set in CSR form.  Contours of slice s are [slice_start[s], slice_start[s+1]); the points
 * of contour c are xy[2*contour_start[c]] .. xy[2*contour_start[c+1]-1] as (x, y) int32 pairs. */
typedef struct ms_polygons {
    int32_t* xy;             /* capacity cap_points * 2                                */
    int64_t  cap_points;
    int32_t* contour_start;  /* capacity cap_contours + 1                              */
    int64_t  cap_contours;
    int32_t* slice_start;    /* capacity batch + 1                                     */
    int64_t  n_points;       /* out: totals (also set on MS_ERR_CAPACITY = what is needed) */
    int64_t  n_contours;     /* out */
} ms_polygons;

/* Replaces Mask2Polygon::extract_contours (src/mask2polygon.cpp:29-36: threshold(127) +
 * findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)) and map_contour_points (:41-63).
 * mask: batch x hgt x w u8; foreground = value > threshold (127 in the reference).
 * Points are scaled by (orig_w / w, orig_h / hgt) in double and truncated (:54-55); pass
 * orig_w = w, orig_h = hgt for unmapped mask-space coordinates.
 * Host buffers in `out` are caller-allocated; on MS_ERR_CAPACITY n_points / n_contours hold the
 * required sizes and nothing else is written. */
MS_API int ms_mask2polygon_host(ms_handle* h, const uint8_t* h_mask, int hgt, int w, int batch,
                                int threshold, int orig_w, int orig_h, ms_polygons* out);
/* Same with the mask already on the device (polygons still land in host buffers). */
MS_API int ms_mask2polygon_dev(ms_handle* h, const uint8_t* d_mask, int hgt, int w, int batch,
                               int threshold, int orig_w, int orig_h, ms_polygons* out, void* stream);

/* ---- whole path ------------------------------------------------------------------------------ */

/* Replaces the compute of MedicalSeg::process_single_image (src/process.cpp:188-262) for a batch of
 * slices held in memory: preprocess -> UNet -> head -> postprocess -> LUT -> contours -> mapping,
 * with no PNG / JSON round trips.  h_src: batch slices of hgt x w u16.  Optional side outputs
 * (host, may be NULL): h_norm_u8 batch x net_h x net_w, h_mask_u8 (values {0,fg}) same size. */
MS_API int ms_process_batch_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int batch,
                                 ms_polygons* out, uint8_t* h_norm_u8, uint8_t* h_mask_u8);

/* BASELINE.json configs[3] ("multi-class UNet with argmax head and per-class contours"): K1 and the UNet once, then for every
 * requested label k = classes[i] (1 .. 255) postprocess with FOREGROUND_VALUE = k (src/postprocess.cpp:5 made a parameter) and
 * mask2polygon of (mask == k), all on the device.  outs[i] receives label classes[i]'s polygons in original coordinates;
 * h_raw_mask_u8 (optional, [batch][net_h][net_w]) the argmax mask, h_clean_masks_u8 (optional, [n_classes][batch][net_h][net_w])
 * the cleaned masks.  MS_ERR_CAPACITY: every outs[i].n_contours / n_points says what that label needs. */
MS_API int ms_process_batch_multiclass_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int batch, const int32_t* classes,
                                            int n_classes, ms_polygons* outs, uint8_t* h_raw_mask_u8, uint8_t* h_clean_masks_u8);

/* Device-resident variant used for kernel-only timing: input already in HBM, polygons stay in the
 * handle's device workspace.  With n_points / n_contours given the call waits for the two totals (and grows the
 * device polygon capacities when the batch needed more); with both NULL it is fully asynchronous on `stream` -- no
 * host round trip, the 32-byte result header lands in pinned memory behind the kernels -- and ms_last_counts
 * collects it later (MS_ERR_CAPACITY there = the workspace was too small for that batch and has been grown). */
MS_API int ms_process_batch_dev(ms_handle* h, const uint16_t* d_src, int w, int hgt, int batch,
                                int64_t* n_points, int64_t* n_contours, void* stream);
MS_API int ms_last_counts(ms_handle* h, int64_t* n_points, int64_t* n_contours);

/* Asynchronous, double-buffered form of ms_process_batch_host for streaming a volume (the serial per-file loop at
 * src/main.cpp:148-164): submit batch i+1 on the other slot before collecting batch i, and the H2D copy of i+1 and
 * the host-side collection of i overlap the GPU work.  slot is 0 or 1.  h_src should be pinned (ms_alloc_pinned);
 * it must stay valid until the matching ms_wait_batch returns.  Device-side polygon capacities start at
 * max(64, 1 / min_area_ratio + 1) contours and 8,192 points per slice; a batch that exceeds them is re-run inside
 * ms_wait_batch on the slot's still-resident input with grown buffers (the call still succeeds).  The D2H issued behind
 * the kernels is sized from the slot's previous batch (+25 %); ms_wait_batch fetches a remainder only when needed, so the
 * bytes copied track the bytes used (ms_get_transfer_bytes reports them).
 * The slot's kernel chain (K1 .. K6, 42 launches, no host round trip) is captured into a CUDA graph on the second call with
 * the same (w, hgt, batch) and replayed afterwards -- what the reference does for its inference (cudaGraphLaunch,
 * src/process.cpp:147); a change of shape or any buffer reallocation drops the graph.  MEDSEG_GRAPH=0 disables it. */
MS_API int ms_submit_batch_host(ms_handle* h, int slot, const uint16_t* h_src, int w, int hgt, int batch);
MS_API int ms_wait_batch(ms_handle* h, int slot, ms_polygons* out);

/* One volume of n_slices slices (h_src: n_slices x hgt x w u16, pinned preferred) through ONE call on every GPU of the
 * handle -- BASELINE cfg3, the per-file loop of src/main.cpp:148-164 turned into a sharded batcher.  A handle created with
 * "devices": [0, 1, ...] (or "all"; MEDSEG_DEVICES supplies the list when the config names no device) owns one full engine
 * per GPU; GPU g takes the contiguous block [g * n / G, (g + 1) * n / G) and one host thread, which streams it in
 * max_batch-sized sub-batches through the double-buffered ms_submit_batch_host / ms_wait_batch pair.  No collective: the
 * blocks' polygon sets are concatenated on the host in slice order (out->slice_start needs n_slices + 1 entries).
 * h_norm_u8 / h_mask_u8 (n_slices x net_h x net_w, may be NULL) switch the workers to the synchronous per-batch call.
 * With one device it is simply the streaming batcher. */
MS_API int ms_process_volume_host(ms_handle* h, const uint16_t* h_src, int w, int hgt, int64_t n_slices, ms_polygons* out,
                                  uint8_t* h_norm_u8, uint8_t* h_mask_u8);
/* GPUs behind this handle (1 unless it was created with "devices"). */
MS_API int ms_device_count(ms_handle* h);

/* Replaces MedicalSeg::process_single_image(raw_path, width, height, output_dir) including its
 * artefacts: <stem>_normalized.png, <stem>_original_sizes.json, <stem>_mask.png,
 * <stem>_contour_overlay.png, <stem>.json (src/process.cpp:207-242, src/mask2polygon.cpp:134-222). */
MS_API int ms_process_raw_file(ms_handle* h, const char* raw_path, int w, int hgt, const char* out_dir);

/* Batched form of the per-file loop of src/main.cpp:148-164: n headerless u16 files of the same w x hgt, artefacts of
 * file i (the same five as ms_process_raw_file, byte-identical to calling it per file) into out_dirs[i].  A prefetch
 * (a multi-GPU handle splits the list into one contiguous block per GPU, each with its own pipeline and host thread) thread
 * reads batch k+1 into pinned memory and a pool of writer threads (MEDSEG_WRITERS, default min(cores, 16))
 * encodes batch k-1 while the GPU works on batch k (max_batch files per launch sequence).  A file that cannot be read
 * or written is counted in *n_failed (and ok[i] = 0) without stopping the others -- the reference's success_count /
 * fail_count (src/main.cpp:160-164); the return value is MS_OK unless the call itself failed.  ok, n_ok, n_failed may
 * be NULL. */
MS_API int ms_process_raw_files(ms_handle* h, const char* const* raw_paths, const char* const* out_dirs, int64_t n, int w, int hgt,
                                uint8_t* ok, int64_t* n_ok, int64_t* n_failed);

/* Replaces find_16bit_images + the directory branch of src/main.cpp:28-48,134-168: regular files with extension .raw
 * .dcm .tif .tiff (case-insensitive) in input_dir, optionally recursive with the sub-directory structure reproduced
 * under out_dir (:151-156).  The list is sorted by path; shard (shard_index of shard_count)
 * takes a contiguous block of it, so N processes -- one per GPU -- cover a directory without talking to each other.
 * *n_found = files in the whole list, *n_ok / *n_failed = this shard's outcome. */
MS_API int ms_process_directory(ms_handle* h, const char* input_dir, int w, int hgt, const char* out_dir, int recursive,
                                int shard_index, int shard_count, int64_t* n_found, int64_t* n_ok, int64_t* n_failed);

/* Byte-exact LabelMe-style document of Mask2Polygon::generate_json (src/mask2polygon.cpp:68-109,
 * nlohmann::json dump with std::setw(4)).  Writes at most `cap` bytes (no NUL) to `dst`, returns the
 * full length (call with cap = 0 to size) or a negative ms_status. */
MS_API int64_t ms_polygons_to_json(const int32_t* xy, const int32_t* contour_start, int n_contours,
                                   const char* base_name, int orig_w, int orig_h, char* dst, int64_t cap);

/* The same document for every slice of a polygon set at once, formatted by `n_threads` host threads (0 = all cores):
 * the host stage that follows the GPU at > 10 k slices/s.  Texts are concatenated into `dst`
 * (no NUL) in slice order, offsets[s] .. offsets[s + 1] delimiting slice s (offsets has n_slices + 1 entries, may be
 * NULL); a slice without contours contributes nothing, as the reference writes no file for it
 * (src/mask2polygon.cpp:183-186).  base_names[s] is the stem written into "imagePath".  Returns the total length; when
 * it exceeds `cap` nothing is copied (call again with a larger buffer), negative ms_status on error. */
MS_API int64_t ms_polygons_to_json_batch(const int32_t* xy, const int32_t* contour_start, const int32_t* slice_start, int n_slices,
                                         const char* const* base_names, int orig_w, int orig_h, int n_threads, char* dst,
                                         int64_t cap, int64_t* offsets);

/* ---- opt-in polygon simplification -------------------------------------------------------------
 * NOT part of the reference (src/mask2polygon.cpp:34 stops at CHAIN_APPROX_SIMPLE; SURVEY.md finding 4).  With
 * eps > 0 every polygon-producing call runs each contour through closed-curve Douglas-Peucker in network space,
 * vertex for vertex what cv2.approxPolyDP(contour, eps, true) of OpenCV 4.13 returns, before the coordinate mapping
 * of src/mask2polygon.cpp:41-63.  Config key "dp_epsilon" sets the initial value; 0 (the default) = off = the
 * reference's output.  Applies to the handle and its per-GPU peers; not to be called while a batch is in flight. */
MS_API int ms_set_dp_epsilon(ms_handle* h, double eps);
MS_API double ms_dp_epsilon(ms_handle* h);

/* ---- instrumentation ------------------------------------------------------------------------- */

/* Number of kernels this library launched on the handle since creation (bench.py `gpu_launches`). */
MS_API int64_t ms_launch_count(ms_handle* h);

/* Bytes this handle has copied host->device / device->host with its own transfers since creation (bench.py reports
 * the per-step difference as e2e.h2d_bytes_per_step / d2h_bytes_per_step: what was copied, not what was useful). */
MS_API int ms_get_transfer_bytes(ms_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* Run only UNet conv layer `layer` (0-based index into the layer table, see ms_layer_name) `iters`
 * times on the handle's stream and return the average milliseconds per launch measured with CUDA
 * events; `flops` receives the layer's 2*MAC count for `batch` slices.  For roofline reporting. */
MS_API int ms_time_layer(ms_handle* h, int layer, int batch, int iters, float* ms_per_launch, double* flops);
MS_API int ms_layer_count(ms_handle* h);
/* In-step layer timing: after ms_profile_layers_begin(h, n) the next n UNet forward passes of the eager entry points
 * (ms_unet_forward_*, ms_process_batch_host/_dev; not the graph-replayed ms_submit_batch_host) bracket every layer launch
 * with CUDA events on the launching stream.  ms_profile_layers_read waits for them, returns the average milliseconds per
 * layer (ms_layer_count entries) and the number of passes recorded, and switches the timing off again. */
MS_API int ms_profile_layers_begin(ms_handle* h, int max_forwards);
MS_API int ms_profile_layers_read(ms_handle* h, float* ms_per_layer, int n_layers, int* n_forwards);
/* Same switch, stage granularity: average milliseconds of K1 preprocess | UNet | K5 postprocess | K6 mask2polygon
 * (ms4[0..3]) over the passes of ms_process_batch_host/_dev recorded since ms_profile_layers_begin. */
MS_API int ms_profile_stages_read(ms_handle* h, float* ms4, int* n_passes);
MS_API const char* ms_layer_name(ms_handle* h, int layer);
/* Kernel instantiation that layer runs on (as ncu prints it), valid until the next call on this thread. */
MS_API const char* ms_layer_kernel(ms_handle* h, int layer);

/* Debug (MEDSEG_FUSED_DBG=1): clock64 stamps of slice 0 at the phase boundaries of the last one-CTA-per-slice
 * postprocess / mask2polygon kernel (csrc/slice_fused.cuh); 32 entries.  MS_ERR_STATE when the switch is off. */
MS_API int ms_debug_fused_phases(long long* out32);

/* Debug: copy an internal activation (bf16 NHWC) of the last forward to host as fp32 NCHW.
 * Returns element count or negative status.  `name` as in ms_layer_name. */
MS_API int64_t ms_debug_read_activation(ms_handle* h, const char* name, int batch, float* h_dst, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* MEDSEG_B200_H */
