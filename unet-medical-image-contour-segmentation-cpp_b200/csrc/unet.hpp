// unet.hpp -- the UNet forward ("process" stage) as a table of tcgen05 implicit-GEMM launches.
#pragma once
#include <cuda.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace ms {

struct UNetLayer {
    std::string name;        // enc1a, enc1b, ..., up4, dec4a, ..., dec1b_head
    int kind = 0;            // 0 = direct first conv (Cin = 1), 1 = conv3x3, 2 = ConvT 2x2 s2, 3 = conv3x3 + head
    int level = 0;           // pixel grid of the GEMM rows: (H >> level) x (W >> level)
    int Cin = 0, Cout = 0;
    int taps = 9;
    int n_total = 0;         // GEMM N
    int block_n = 0;
    int src = -1, dst = -1, pool_dst = -1;   // activation buffer ids
    int dst_coff = 0;
    __nv_bfloat16* w = nullptr;  // device [n_total][taps * Cin] bf16 (kinds 1..3)
    float* w_f32 = nullptr;      // device [64][9] fp32 (kind 0)
    float* bias = nullptr;       // device [Cout] fp32
    CUtensorMap map_a, map_b;
    CUtensorMap map_b_half;      // cta_group::2 kernel: box {64 k, block_n / 2 rows}
    CUtensorMap map_b_half_alt;  // same for the small-batch tiling (block_n_alt / 2 rows)
    int block_n_alt = 0;         // 0 = none; 128 = deep layers switch to 128-wide N tiles when 256-wide ones cannot fill a wave
    CUtensorMap map_out;         // TMA store of the epilogue (one epilogue warp's 32-pixel x 64-channel slab)
    CUtensorMap map_pool;        // row-pair kernels: TMA store of the fused 2x2 max-pool
    CUtensorMap map_a_row;       // halo kernel: box {64 ch, 10 px, 18 rows}
    bool convt_pair = false;     // ConvT: cta_group::2 GEMM (convt_pair_kernel) with map_a_row = {64 ch, 8 px, 16 rows} tiles
    int halo = 0;                // 0 = per-tap streaming kernel, 1 = halo-stationary kernel, 2 = its cta_group::2 version, 3 = row-pair kernel, 4 = row-pair cta_group::2
    int resident_kc = 0;         // halo kernel: > 0 when all weights stay in shared memory
    int halo_pitch = 16;         // halo kernel: shared-memory rows per halo image row (10 dense / 16 aligned)
    double flops_per_slice = 0;  // 2 * MAC
    // the kernel instantiation run_layer launches for this layer, as ncu prints it
    std::string kernel_name() const {
        const char* epi = kind == 3 ? "EPI_HEAD" : (kind == 2 ? "EPI_CONVT" : "EPI_STORE");
        if (kind == 0) return "first_conv_kernel";
        if (convt_pair) return "tc::convt_pair_kernel<" + std::to_string(block_n) + ">";
        if (halo == 4) return std::string("tc::conv_rowpair2_kernel<") + epi + ", " + std::to_string(resident_kc) + ">";
        if (halo == 3) return std::string("tc::conv_rowpair_kernel<") + epi + ", " + std::to_string(resident_kc) + ">";
        if (halo == 2) return "tc::conv_halo2_kernel<" + std::to_string(block_n) + ", " + epi + ", " + std::to_string(resident_kc) + ">";
        if (halo == 1) return "tc::conv_halo_kernel<" + std::to_string(block_n) + ", " + epi + ", " + std::to_string(resident_kc) + ", 10>";
        return "tc::conv_gemm_kernel<" + std::to_string(block_n) + ", " + epi + ">";
    }
};

struct ActBuf {
    std::string name;
    int level = 0, C = 0;
    __nv_bfloat16* p = nullptr;
};

class UNet {
  public:
    ~UNet();
    // Loads a MSEGW001 blob, folds BN, converts to bf16 K-major, builds buffers and tensor maps.
    void load(const std::string& blob_path, int net_h, int net_w, int n_classes_cfg, int max_batch, int fg_value, int sm_count);
    bool loaded() const { return loaded_; }
    void forward(const uint8_t* d_in_u8, int batch, uint8_t* d_mask, float* d_logits, cudaStream_t st);
    void run_layer(int li, const uint8_t* d_in_u8, int batch, uint8_t* d_mask, float* d_logits, cudaStream_t st);
    const std::vector<UNetLayer>& layers() const { return layers_; }
    const std::vector<ActBuf>& buffers() const { return bufs_; }
    int n_classes() const { return n_classes_; }
    int64_t n_params() const { return n_params_; }
    double flops_per_slice() const { return flops_; }
    int net_h() const { return H_; }
    int net_w() const { return W_; }
    int max_batch() const { return max_batch_; }
    uint8_t* scratch_mask() { return scratch_mask_.as<uint8_t>(); }
    // In-step layer timing: while enabled, forward() brackets every layer launch with CUDA events (up to `max_forwards`
    // passes); profile_read() waits for them and returns the average milliseconds per layer and the number of passes.
    void profile_begin(int max_forwards);
    int profile_read(std::vector<float>& ms_per_layer);

  private:
    std::vector<cudaEvent_t> prof_events_;   // [pass][layer + 1]
    int prof_cap_ = 0, prof_n_ = 0;          // passes: capacity / recorded
    bool loaded_ = false;
    int H_ = 0, W_ = 0, n_classes_ = 0, max_batch_ = 0, fg_value_ = 2, sm_count_ = 148;
    bool naive_ = false;  // MEDSEG_NAIVE_CONV=1: CUDA-core reference kernels (debug / validation only)
    bool halo_enabled_ = true;  // MEDSEG_HALO=0: force the per-tap kernel everywhere (A/B measurements)
    bool cta2_enabled_ = true;  // MEDSEG_CTA2=0: single-CTA halo kernel only
    bool cta2_force_ = false;
    bool cta2_single_pref_ = false;   // MEDSEG_CTA2=1: single-CTA halo kernel where both forms fit (A/B measurements)
    bool deep2_enabled_ = true;    // MEDSEG_DEEP2=0: per-tap kernel for the N = 256 layers
    bool res_big_ = true;          // MEDSEG_RES_BIG=0: 144 KiB half-weight sets stream instead of staying resident
    bool rowpair_enabled_ = true;  // MEDSEG_ROWPAIR=0: kernels 2 / 3 for the Cout = 64 layers
    bool rowpair2_enabled_ = true; // MEDSEG_ROWPAIR2=0: single-CTA row-pair kernel instead of its cta_group::2 form
    bool rowpair_stream_ = true;   // MEDSEG_ROWPAIR=1: row-pair kernel only where its weights stay resident (not dec1a)
    bool stream2_enabled_ = true;  // MEDSEG_STREAM2=0: per-tap kernel instead of the streaming pair kernel (dec2a)
    int halo_pitch_ = 16;       // MEDSEG_HALO_PITCH
    int64_t n_params_ = 0;
    double flops_ = 0;
    std::vector<UNetLayer> layers_;
    std::vector<ActBuf> bufs_;
    std::vector<void*> allocs_;
    float* head_w_ = nullptr;
    float* head_b_ = nullptr;
    DevBuf scratch_mask_;
};

}  // namespace ms
