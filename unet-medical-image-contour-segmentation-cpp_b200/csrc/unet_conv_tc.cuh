// unet_conv_tc.cuh -- implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA (sm_100a).
//
// This is the dense contraction of the UNet forward that the reference runs inside an opaque
// TensorRT engine (/root/reference/src/process.cpp:94,147).  Six persistent, warp-specialised
// kernels share the PTX wrappers, the epilogues and the pipeline scheme described here (kernel 1);
// unet.cu picks one per layer from measurements:
//   1 conv_gemm_kernel      per-tap operand streaming           the transposed convs up1 - up3 (8 epilogue warps)
//   2 conv_halo_kernel      halo-stationary A, resident weights  A/B fallback of the narrow layers
//   3 conv_halo2_kernel     kernel 2 as a cta_group::2 pair      enc2*, dec2*, and every N = 256 layer (streamed weights)
//   4 convt_pair_kernel     transposed conv as a pair GEMM       up4
//   5 conv_rowpair_kernel   two output rows per GEMM row         A/B fallback of kernel 6
//   6 conv_rowpair2_kernel  kernel 5 as a cta_group::2 pair      the Cout = 64 layers: enc1b, dec1a, dec1b + head
//
// Every GEMM-shaped layer of the network is one of:
//
//   conv3x3 (pad 1) + folded BN + ReLU   M = pixels, N = Cout, K = 9 * Cin     (18 layers)
//   ConvTranspose2d(k=2, s=2) + bias     M = input pixels, N = 4 * Cout, K = Cin  (4 layers)
//
// Data layout: activations NHWC bf16; weights [N][K] bf16, K-major, K = tap * Cin + ci.
//
// Tiling: one CTA tile = 16 (x) x 8 (y) output pixels = 128 GEMM rows, BLOCK_N output channels.
//   A operand : for every (tap, 64-channel chunk) ONE 4-D TMA box {64 ch, 16 px, 8 rows, 1 image}
//               at the tap's (dx, dy) shift.  TMA zero-fills out-of-bounds pixels, which *is* the
//               conv's zero padding -- im2col never exists in memory.  The box lands as 128 rows of
//               128 B with the 128-byte swizzle = the canonical K-major SW128 UMMA operand.
//   B operand : 2-D TMA box {64 k, BLOCK_N rows}, same swizzle.
//   D         : fp32 accumulators in TMEM, 128 lanes x BLOCK_N columns, double buffered
//               (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer
// (one lane issues tcgen05.mma / tcgen05.commit), warps 2..5 = epilogue (TMEM lane quarter =
// warp_id % 4).  Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue).
//
// Epilogues (template EPI):
//   EPI_STORE : + bias, ReLU, bf16, NHWC store at a channel offset / stride (so encoder outputs are
//               written straight into the decoder's concat buffer -- torch.cat never runs), and an
//               optional fused 2x2 max-pool: the 16x8 tile puts every 2x2 window inside one warp
//               (lanes l, l^1, l^16, l^17), so the pool is two shuffles.
//   EPI_CONVT : + bias, bf16, scatter to (2y+dy, 2x+dx) of the up-sampled image at the concat
//               buffer's channel offset (transposed conv fused with the skip concat).
//   EPI_HEAD  : + bias, ReLU, then the 1x1 head (64 -> n_classes) on the fp32 accumulators in
//               registers, first-max argmax (src/process.cpp:158-170) or `logit > 0`; the last
//               feature map never touches HBM.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ms {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int TILE_W = 16, TILE_H = 8;
constexpr int BLOCK_K = 64;   // bf16 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB

enum Epi { EPI_STORE = 0, EPI_CONVT = 1, EPI_HEAD = 2 };

struct ConvArgs {
    int H, W, batch;        // pixel grid of the GEMM rows (input grid for ConvT)
    int Cin, taps;          // taps = 9 (conv3x3) or 1 (ConvT)
    int n_total;            // GEMM N: Cout, or 4 * Cout for ConvT
    int Cout;               // bias length / channel modulus
    const float* bias;      // [Cout] fp32
    __nv_bfloat16* out;     // NHWC destination (may be null for EPI_HEAD)
    int out_cstride, out_coff;
    __nv_bfloat16* pool;    // optional pooled destination (EPI_STORE), channel stride pool_cstride
    int pool_cstride;
    const float* head_w;    // [n_classes][64]
    const float* head_b;    // [n_classes]
    int n_classes;          // 1 = binary head (logit > 0), else argmax over n_classes
    int fg_value;           // value written by the binary head
    uint8_t* mask;          // [batch][H][W]
    float* logits;          // optional [batch][n_classes][H][W]
};

template <int BLOCK_N, int EPI = 0>
struct Cfg {
    // The transposed convs (K = Cin only: 2 - 16 chunks per tile, N = 256) are bound by their EPILOGUE -- four 64-column
    // blocks of TMEM load / pack / stage / TMA scatter per warp against 8 - 64 MMAs -- so they get two epilogue warps per
    // TMEM lane quarter, each taking half of the columns.
    static constexpr int EPI_GROUPS = (EPI == 1 /* EPI_CONVT */ && BLOCK_N == 256) ? 2 : 1;
    static constexpr int THREADS = 64 + 128 * EPI_GROUPS;
    static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BLOCK_N == 256) ? (EPI_GROUPS == 2 ? 3 : 4) : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;  // 128 / 256 / 512: powers of two
    static constexpr int AUX_BYTES = 4096;  // barriers, tmem pointer, head weights
    static constexpr int STG_BYTES = EPI_GROUPS * 4 * 4096;  // one 4 KiB output staging slab per epilogue warp
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + AUX_BYTES + 1024;  // + alignment slack
    static_assert(SMEM_BYTES <= 227 * 1024, "conv_gemm_kernel: shared memory");
};

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// "Accumulator drained" arrive of an epilogue thread.  Relaxed: what it orders (this thread's tcgen05.ld before the next MMA
// into the accumulator) is carried by tcgen05.wait::ld + tcgen05.fence::before_thread_sync here and fence::after_thread_sync
// behind the issuer's wait.  The default .release compiled to a MEMBAR.ALL.CTA in front of every arrive, which waits for the
// thread's outstanding GLOBAL stores (previous tile's mask / pooled pixels) -- the top stall of the epilogue warps in ncu.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait.  Real waits are microseconds; a pipeline bug would otherwise hang the GPU.  The bound is 20 s of WALL
// time (%globaltimer, polled only every ~2^28 cycles so the spin stays cheap), far beyond anything time-slicing (MPS),
// preemption or a profiler replay can add to a legitimate wait.  On expiry the kernel raises a flag in mapped host
// memory (checked by every C-ABI entry point -> MS_ERR_INTERNAL) and stops waiting -- its output is garbage, but the
// context and every other handle in the process survive.  -DMEDSEG_DEBUG_WATCHDOG traps instead (sticky context error).
__device__ unsigned* g_watchdog_dev = nullptr;    // set per device by UNet::load (cudaMemcpyToSymbol)
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __noinline__ void watchdog_expired() {
    if (g_watchdog_dev) {
        *reinterpret_cast<volatile unsigned*>(g_watchdog_dev) = 1u;
        __threadfence_system();
    }
#ifdef MEDSEG_DEBUG_WATCHDOG
    __trap();
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long c0 = clock64();
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - c0 > (1ll << 28)) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) {
                watchdog_expired();
                return;
            }
            c0 = clock64();
        }
    }
}
// One lane of a fully converged warp (the same lane every time).  Issuing tcgen05 / TMA instructions under this
// predicate inside warp-convergent control flow lets ptxas emit them straight-line; issuing them from a
// divergent `if (lane == 0)` region makes it wrap every UTCHMMA in an ELECT / BRA.U.ANY loop (measured:
// ~17 issue-slot instructions per MMA, the issuing thread became the bottleneck of the N <= 128 kernels).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// K-major, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart (SBO = 64 x 16 B), LBO unused (1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t make_smem_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = BLOCK_N.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}


// ------------------------------------------------------------------------------------ shared epilogue
// One epilogue thread owns one GEMM row (= one pixel of image b) and walks its BLOCK_N columns 64 at a
// time.  TW is the tile width in pixels: lane l of a warp holds pixel (l / TW, l % TW) of the warp's
// 32-row slab (32 / TW tile rows), so the 2x2 pool partners are lanes l^1 and l^TW.
//
// Stores: the 32 pixels x 64 channels of a warp (4 KiB) are staged in the warp's own shared-memory slab
// in the 128-byte-swizzled layout and written with ONE TMA tensor store (full 128-byte lines, channel
// offset / stride of the concat buffers and the (2y+dy, 2x+dx) scatter of the transposed conv are all
// in the tensor map), instead of 16-byte stores that each touch 32 different lines.
struct EpiCtx {
    const CUtensorMap* map_out;
    uint32_t slab;       // shared-memory address of this warp's 4 KiB staging slab (1024-aligned)
    int b, y0, x0, n0;   // tile origin
    int slab_y;          // first tile row of this warp's slab, relative to y0
    int y, x;            // this thread's pixel
    int lane;
};

__device__ __forceinline__ void mbar_arrive_cluster_fwd(uint32_t cluster_addr);

template <int BLOCK_N, int EPI, int TW, bool REMOTE = false>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& args, const float* s_head, uint32_t taddr, const EpiCtx& e,
                                              uint64_t* tmem_empty_bar, uint32_t remote_empty = 0, int c_begin = 0, int c_end = BLOCK_N) {
    if (EPI == EPI_HEAD) {
        // BLOCK_N == 64: the whole feature vector of this pixel
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        if (REMOTE) mbar_arrive_cluster_fwd(remote_empty); else mbar_arrive(tmem_empty_bar);
        float f[64];
        const float4* b4 = reinterpret_cast<const float4*>(args.bias);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 lo = __ldg(b4 + j), hi = __ldg(b4 + 8 + j);
            f[4 * j + 0] = fmaxf(__uint_as_float(r0[4 * j + 0]) + lo.x, 0.0f);
            f[4 * j + 1] = fmaxf(__uint_as_float(r0[4 * j + 1]) + lo.y, 0.0f);
            f[4 * j + 2] = fmaxf(__uint_as_float(r0[4 * j + 2]) + lo.z, 0.0f);
            f[4 * j + 3] = fmaxf(__uint_as_float(r0[4 * j + 3]) + lo.w, 0.0f);
            f[32 + 4 * j + 0] = fmaxf(__uint_as_float(r1[4 * j + 0]) + hi.x, 0.0f);
            f[32 + 4 * j + 1] = fmaxf(__uint_as_float(r1[4 * j + 1]) + hi.y, 0.0f);
            f[32 + 4 * j + 2] = fmaxf(__uint_as_float(r1[4 * j + 2]) + hi.z, 0.0f);
            f[32 + 4 * j + 3] = fmaxf(__uint_as_float(r1[4 * j + 3]) + hi.w, 0.0f);
        }
        const size_t pix = ((size_t)e.b * args.H + e.y) * args.W + e.x;
        const size_t plane = (size_t)args.H * args.W;
        float best = -3.402823466e+38f;  // -FLT_MAX, src/process.cpp:159
        int best_c = 0;
        for (int c = 0; c < args.n_classes; ++c) {
            float s = s_head[args.n_classes * 64 + c];
            const float4* w4 = reinterpret_cast<const float4*>(s_head + c * 64);   // broadcast 16-byte loads
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 w = w4[j];
                s = fmaf(f[4 * j + 0], w.x, s);
                s = fmaf(f[4 * j + 1], w.y, s);
                s = fmaf(f[4 * j + 2], w.z, s);
                s = fmaf(f[4 * j + 3], w.w, s);
            }
            if (args.logits) args.logits[((size_t)e.b * args.n_classes + c) * plane + (size_t)e.y * args.W + e.x] = s;
            if (s > best) { best = s; best_c = c; }   // strict >: first max wins, NaN never wins
        }
        args.mask[pix] = args.n_classes == 1 ? (uint8_t)(best > 0.0f ? args.fg_value : 0) : (uint8_t)best_c;
    } else {
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 64) {
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + c0, r0);
            tmem_ld32(taddr + c0 + 32, r1);
            tmem_ld_wait();
            if (c0 + 64 >= c_end) {  // last block read: hand the accumulator back to the MMA warp
                tc_fence_before();
                if (REMOTE) mbar_arrive_cluster_fwd(remote_empty); else mbar_arrive(tmem_empty_bar);
            }
            const int col = e.n0 + c0;             // first GEMM column of this 64-block
            const int co = col % args.Cout;        // 64-aligned, never straddles Cout
            const float4* b4 = reinterpret_cast<const float4*>(args.bias + co);
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 lo = __ldg(b4 + j), hi = __ldg(b4 + 8 + j);
                float v0 = __uint_as_float(r0[4 * j + 0]) + lo.x, v1 = __uint_as_float(r0[4 * j + 1]) + lo.y;
                float v2 = __uint_as_float(r0[4 * j + 2]) + lo.z, v3 = __uint_as_float(r0[4 * j + 3]) + lo.w;
                float w0 = __uint_as_float(r1[4 * j + 0]) + hi.x, w1 = __uint_as_float(r1[4 * j + 1]) + hi.y;
                float w2 = __uint_as_float(r1[4 * j + 2]) + hi.z, w3 = __uint_as_float(r1[4 * j + 3]) + hi.w;
                if (EPI == EPI_STORE) {
                    v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); v2 = fmaxf(v2, 0.0f); v3 = fmaxf(v3, 0.0f);
                    w0 = fmaxf(w0, 0.0f); w1 = fmaxf(w1, 0.0f); w2 = fmaxf(w2, 0.0f); w3 = fmaxf(w3, 0.0f);
                }
                pk[2 * j] = pack_bf16(v0, v1);
                pk[2 * j + 1] = pack_bf16(v2, v3);
                pk[16 + 2 * j] = pack_bf16(w0, w1);
                pk[16 + 2 * j + 1] = pack_bf16(w2, w3);
            }
            // the previous TMA store must have finished reading the slab before it is overwritten
            if (e.lane == 0) tma_store_wait_read();
            __syncwarp();
            const uint32_t row_addr = e.slab + (uint32_t)e.lane * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j)   // 16-byte chunk j of row `lane` lives at chunk j ^ (lane & 7)
                st_shared_v4(row_addr + (uint32_t)((j ^ (e.lane & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (e.lane == 0) {
                if (EPI == EPI_STORE) {
                    // box {64 ch, TW px, 32/TW rows, 1 image}
                    tma_store_4d(e.map_out, e.slab, args.out_coff + col, e.x0, e.y0 + e.slab_y, e.b);
                } else {
                    // EPI_CONVT: destination viewed as (C, dx, W, dy, B*H); box {64, 1, TW, 1, 32/TW}
                    const int q = col / args.Cout;
                    tma_store_5d(e.map_out, e.slab, args.out_coff + co, q & 1, e.x0, q >> 1, e.b * args.H + e.y0 + e.slab_y);
                }
                tma_store_commit();
            }
            if (EPI == EPI_STORE && args.pool) {
                // 2x2 max-pool out of the staged slab (the TMA store only reads it): the slab holds 32 px x 64 ch, i.e.
                // 8 pooled pixels x 8 sixteen-byte channel chunks = 64 work items, two per lane.  Item (q, j): max over
                // the rows (= pixels) l0, l0+1, l0+TW, l0+TW+1 of chunk j, which the 128-byte swizzle put at chunk
                // position j ^ (row & 7).  Lanes 4q .. 4q+3 write pooled pixel q's 128 bytes contiguously.
                // (The earlier form exchanged the 32 packed registers with two shuffles each: 4 warps x 128 SHFL per tile
                // made the epilogue, not the MMA, the bound of enc2b.)
                const int q = e.lane >> 2;
                const int l0 = (q % (TW / 2)) * 2 + (q / (TW / 2)) * 2 * TW;
                const int px = e.x0 + (l0 % TW), py = e.y0 + e.slab_y + (l0 / TW);
                uint4* dst = reinterpret_cast<uint4*>(args.pool + (((size_t)e.b * (args.H / 2) + (py >> 1)) * (args.W / 2) + (px >> 1)) * args.pool_cstride + col);
                uint4 v[2][4];     // all eight loads first: one exposed shared-memory latency, not two
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int j = (e.lane & 3) * 2 + it;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int row = l0 + (k & 1) + (k >> 1) * TW;
                        v[it][k] = ld_shared_v4(e.slab + (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4));
                    }
                }
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    uint4 m;
                    m.x = max_bf16x2(max_bf16x2(v[it][0].x, v[it][1].x), max_bf16x2(v[it][2].x, v[it][3].x));
                    m.y = max_bf16x2(max_bf16x2(v[it][0].y, v[it][1].y), max_bf16x2(v[it][2].y, v[it][3].y));
                    m.z = max_bf16x2(max_bf16x2(v[it][0].z, v[it][1].z), max_bf16x2(v[it][2].z, v[it][3].z));
                    m.w = max_bf16x2(max_bf16x2(v[it][0].w, v[it][1].w), max_bf16x2(v[it][2].w, v[it][3].w));
                    dst[(e.lane & 3) * 2 + it] = m;
                }
            }
        }
    }
}

struct TileCoord {
    int b, y0, x0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(int t, int n_tiles, int tiles_x, int tiles_y, int block_n, int tw, int th) {
    TileCoord c;
    const int n_idx = t % n_tiles;
    int m = t / n_tiles;
    c.x0 = (m % tiles_x) * tw;
    m /= tiles_x;
    c.y0 = (m % tiles_y) * th;
    c.b = m / tiles_y;
    c.n0 = n_idx * block_n;
    return c;
}

__device__ __forceinline__ void tmem_alloc_warp(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_warp(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// ==================================================================================== kernel 1
// Per-tap operand streaming: every k-step loads one shifted 16x8-pixel A box and one B box.  Used where
// N >= 256 keeps the tensor pipe busy per byte fetched (deep layers) and for the ConvT GEMMs (one tap).
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(Cfg<BLOCK_N, EPI>::THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_out, const ConvArgs args) {
    using C = Cfg<BLOCK_N, EPI>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_stg = smem + C::STAGES * C::STAGE_BYTES;
    uint8_t* aux = s_stg + C::STG_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_head = reinterpret_cast<float*>(aux + 512);  // [n_classes][64] weights then [n_classes] bias

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = args.W / TILE_W, tiles_y = args.H / TILE_H;
    const int n_tiles = args.n_total / BLOCK_N;
    const int total = args.batch * tiles_y * tiles_x * n_tiles;
    const int kchunks = args.Cin / BLOCK_K;
    const int ksteps = args.taps * kchunks;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        if (EPI != EPI_HEAD) prefetch_tmap(&map_out);
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128 * C::EPI_GROUPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_warp(tmem_ptr, C::TMEM_COLS);
    if (EPI == EPI_HEAD) {
        for (int i = threadIdx.x; i < args.n_classes * 64 + args.n_classes; i += C::THREADS)
            s_head[i] = i < args.n_classes * 64 ? args.head_w[i] : args.head_b[i - args.n_classes * 64];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            pdl_wait();            // the activations are the previous layer's output
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, n_tiles, tiles_x, tiles_y, BLOCK_N, TILE_W, TILE_H);
                for (int tap = 0; tap < args.taps; ++tap) {
                    const int dy = args.taps == 9 ? tap / 3 - 1 : 0;
                    const int dx = args.taps == 9 ? tap % 3 - 1 : 0;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * C::STAGE_BYTES;
                        uint8_t* sb = sa + A_STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                        tma_load_4d(sa, &map_a, &full_bar[stage], kc * BLOCK_K, tc.x0 + dx, tc.y0 + dy, tc.b);
                        tma_load_2d(sb, &map_b, &full_bar[stage], tap * args.Cin + kc * BLOCK_K, tc.n0);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (whole warp walks the
        // pipeline, one elected lane issues)
        {
            constexpr uint32_t idesc = make_idesc(BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa);
                        const uint64_t bdesc = make_smem_desc(sa + A_STAGE_BYTES);
                        // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                        umma_f16(d_tmem, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 1; k < BLOCK_K / UMMA_K; ++k) umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                        if (ks == ksteps - 1) umma_commit(&tmem_full[acc]);   // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue (warps 2..5)
        const int quarter = warp & 3;           // TMEM lane quarter this warp may access
        const int egroup = (warp - 2) >> 2;     // which share of the columns (two warps per quarter for the transposed convs)
        constexpr int kColsPerGroup = BLOCK_N / C::EPI_GROUPS;
        const int row = quarter * 32 + lane;    // GEMM row inside the tile = pixel
        const int ly = row / TILE_W, lx = row % TILE_W;
        EpiCtx e;
        e.map_out = &map_out;
        e.slab = smem_u32(s_stg + (egroup * 4 + quarter) * 4096);
        e.slab_y = quarter * (32 / TILE_W);
        e.lane = lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tcd = decode_tile(t, n_tiles, tiles_x, tiles_y, BLOCK_N, TILE_W, TILE_H);
            e.b = tcd.b; e.y0 = tcd.y0; e.x0 = tcd.x0; e.n0 = tcd.n0; e.y = tcd.y0 + ly; e.x = tcd.x0 + lx;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            epilogue_tile<BLOCK_N, EPI, TILE_W>(args, s_head, taddr, e, &tmem_empty[acc], 0u, egroup * kColsPerGroup, (egroup + 1) * kColsPerGroup);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI != EPI_HEAD && lane == 0) tma_store_wait_read();   // the slab must outlive the last store's read
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_warp(tmem_base, C::TMEM_COLS);
    }
}

// ==================================================================================== kernel 2
// Halo-stationary conv3x3 for the large-grid, narrow-N layers (levels 0-1: N = 64 / 128), which kernel 1
// leaves L2-bandwidth bound: it re-fetches the same input pixels for each of the 9 taps.
// Here one CTA tile is 8 (x) x 16 (y) pixels and, per 64-channel chunk, the (8+2) x (16+2) input halo is
// loaded ONCE (one TMA box {64 ch, 10 px, 18 rows} = 180 rows of 128 B).  The nine taps are nine UMMA A
// descriptors into that one buffer, start address shifted by ((dy+1) * 10 + (dx+1)) rows of 128 B: an
// 8-pixel tile row is exactly one 8-row group and consecutive tile rows are one halo row (1280 B) apart,
// so SBO = 1280.  The 128-byte swizzle is a function of the absolute shared-memory address bits, which TMA
// (writer) and UMMA (reader) share, so 128-byte-aligned start addresses need no descriptor base offset
// (measured on B200: base_offset = 0 is bit-correct, base_offset = (addr >> 7) & 7 is not).  Weights are either resident in shared memory for the whole
// persistent CTA (RESIDENT_KC = Cin / 64 > 0: all 9 * RESIDENT_KC B tiles, loaded once) or streamed
// through their own ring (RESIDENT_KC == 0).
constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_ROWS = HALO_TH + 2;
constexpr int HALO_BOX_BYTES = HALO_ROWS * (HALO_TW + 2) * 128;  // 23,040 B of pixels per halo tile
// PITCH = rows of 128 B per halo image row in shared memory:
//   10 : dense, one TMA box {64, 10, 18} per chunk (SBO = 1280 B)
//   16 : every halo row starts a fresh 2048 B line group (SBO = 2048 B), 18 row boxes {64, 10, 1} per chunk
template <int PITCH>
struct HaloGeom {
    static constexpr int STAGE_BYTES = (HALO_ROWS * PITCH * 128 + 1023) / 1024 * 1024;
};

template <int BLOCK_N, int RESIDENT_KC, int PITCH>
struct HaloCfg {
    static constexpr int HALO_STAGE_BYTES = HaloGeom<PITCH>::STAGE_BYTES;
    static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int RES_BYTES = 9 * RESIDENT_KC * B_TILE_BYTES;
    static constexpr int B_STAGES = RESIDENT_KC > 0 ? 0 : (BLOCK_N == 64 ? 8 : 5);
    static constexpr int STG_BYTES = 4 * 4096;  // one 4 KiB output staging slab per epilogue warp
    static constexpr int BUDGET = 227 * 1024 - 4096 - 1024 - STG_BYTES;
    static constexpr int A_STAGES_RAW = (BUDGET - RES_BYTES - B_STAGES * B_TILE_BYTES) / HALO_STAGE_BYTES;
    static constexpr int A_STAGES = A_STAGES_RAW > 4 ? 4 : A_STAGES_RAW;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = RES_BYTES + B_STAGES * B_TILE_BYTES + A_STAGES * HALO_STAGE_BYTES + STG_BYTES + 4096 + 1024;
    static_assert(A_STAGES >= 2, "halo kernel needs at least two A stages");
};

template <int BLOCK_N, int EPI, int RESIDENT_KC, int PITCH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a_row, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_out, const ConvArgs args) {
    using C = HaloCfg<BLOCK_N, RESIDENT_KC, PITCH>;
    static_assert(RESIDENT_KC > 0, "the single-CTA halo kernel keeps the layer's weights resident in shared memory");
    constexpr int HALO_STAGE_BYTES = C::HALO_STAGE_BYTES;
    constexpr int HALO_PITCH = PITCH;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_res = smem;                                            // resident weights (or nothing)
    uint8_t* s_b = smem + C::RES_BYTES;                               // streamed B ring
    uint8_t* s_a = s_b + C::B_STAGES * C::B_TILE_BYTES;               // halo ring
    uint8_t* s_stg = s_a + C::A_STAGES * HALO_STAGE_BYTES;            // output staging slabs
    uint8_t* aux = s_stg + C::STG_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* b_empty = b_full + 8;
    uint64_t* res_full = b_empty + 8;
    uint64_t* tmem_full = res_full + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_head = reinterpret_cast<float*>(aux + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = args.W / HALO_TW, tiles_y = args.H / HALO_TH;
    const int total = args.batch * tiles_y * tiles_x;                 // n_total == BLOCK_N
    const int kchunks = args.Cin / BLOCK_K;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_row);
        prefetch_tmap(&map_b);
        if (EPI != EPI_HEAD) prefetch_tmap(&map_out);
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(res_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_warp(tmem_ptr, C::TMEM_COLS);
    if (EPI == EPI_HEAD) {
        for (int i = threadIdx.x; i < args.n_classes * 64 + args.n_classes; i += NUM_THREADS)
            s_head[i] = i < args.n_classes * 64 ? args.head_w[i] : args.head_b[i - args.n_classes * 64];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0 && blockIdx.x < total) {
            if (RESIDENT_KC > 0) {   // all weights of the layer, once per CTA: tile index = tap * kchunks + kc
                mbar_expect_tx(res_full, C::RES_BYTES);
                for (int i = 0; i < 9 * RESIDENT_KC; ++i)
                    tma_load_2d(s_res + i * C::B_TILE_BYTES, &map_b, res_full, (i / RESIDENT_KC) * args.Cin + (i % RESIDENT_KC) * BLOCK_K, 0);
            }
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            pdl_wait();            // weights are in flight; the activations are the previous layer's output
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, 1, tiles_x, tiles_y, BLOCK_N, HALO_TW, HALO_TH);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    uint8_t* dst = s_a + sa * HALO_STAGE_BYTES;
                    mbar_expect_tx(&a_full[sa], HALO_BOX_BYTES);
                    if (PITCH == HALO_TW + 2) {
                        tma_load_4d(dst, &map_a_row, &a_full[sa], kc * BLOCK_K, tc.x0 - 1, tc.y0 - 1, tc.b);
                    } else {
                        for (int r = 0; r < HALO_ROWS; ++r)
                            tma_load_4d(dst + r * HALO_PITCH * 128, &map_a_row, &a_full[sa], kc * BLOCK_K, tc.x0 - 1, tc.y0 - 1 + r, tc.b);
                    }
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                    if (RESIDENT_KC == 0) {
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&b_empty[sb], pb ^ 1);
                            mbar_expect_tx(&b_full[sb], C::B_TILE_BYTES);
                            tma_load_2d(s_b + sb * C::B_TILE_BYTES, &map_b, &b_full[sb], tap * args.Cin + kc * BLOCK_K, 0);
                            if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (whole warp walks the
        // pipeline, one elected lane issues; the nine taps are straight-line code with immediate offsets)
        if (blockIdx.x < total) {
            constexpr uint32_t idesc = make_idesc(BLOCK_N);
            int sa = 0, acc = 0;
            uint32_t pa = 0, acc_phase = 0;
            mbar_wait(res_full, 0);
            tc_fence_after();
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    if (elect_one()) {
                        // shifted views of the halo buffer: tile row ty lives at halo row ty + (dy+1), column dx+1
                        const uint64_t a0 = make_smem_desc_sbo(smem_u32(s_a + sa * HALO_STAGE_BYTES), HALO_PITCH * 128);
                        const uint64_t b0 = make_smem_desc(smem_u32(s_res + kc * C::B_TILE_BYTES));
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            constexpr int kAUnit = 128 / 16;                        // one 128-byte row in descriptor units
                            const uint64_t adesc = a0 + (uint64_t)(((tap / 3) * HALO_PITCH + (tap % 3)) * kAUnit);
                            const uint64_t bdesc = b0 + (uint64_t)(tap * RESIDENT_KC * (C::B_TILE_BYTES / 16));
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                         (tap | k) != 0 ? 1u : (kc != 0 ? 1u : 0u));
                        }
                        umma_commit(&a_empty[sa]);
                        if (kc == kchunks - 1) umma_commit(&tmem_full[acc]);
                    }
                    __syncwarp();
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue (warps 2..5)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int ly = row / HALO_TW, lx = row % HALO_TW;
        EpiCtx e;
        e.map_out = &map_out;
        e.slab = smem_u32(s_stg + quarter * 4096);
        e.slab_y = quarter * (32 / HALO_TW);
        e.lane = lane;
        e.n0 = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tcd = decode_tile(t, 1, tiles_x, tiles_y, BLOCK_N, HALO_TW, HALO_TH);
            e.b = tcd.b; e.y0 = tcd.y0; e.x0 = tcd.x0; e.y = tcd.y0 + ly; e.x = tcd.x0 + lx;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            epilogue_tile<BLOCK_N, EPI, HALO_TW>(args, s_head, taddr, e, &tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI != EPI_HEAD && lane == 0) tma_store_wait_read();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_warp(tmem_base, C::TMEM_COLS);
    }
}


// ==================================================================================== kernel 3
// cta_group::2 version of the halo-stationary kernel: a CTA pair (cluster of 2, one TPC) works on two
// adjacent 8x16-pixel tiles as ONE UMMA of M = 256.  Each CTA stages its own halo tile (its 128 rows of A)
// and only HALF of the weight tile (N/2 rows of B), so per CTA the weights take half the shared memory
// (144 KiB layers become resident, resident layers get a deeper A ring) and half the operand bandwidth.
// The leader CTA (rank 0) issues tcgen05.mma.cta_group::2; both CTAs' TMA loads signal the leader's "full"
// barriers, tcgen05.commit multicasts "empty" / "accumulator ready" to both CTAs, and both epilogues report
// "accumulator drained" to the leader.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// Remote "accumulator drained" arrive.  The ordering that matters (this thread's tcgen05.ld before the leader's next MMA) is
// carried by tcgen05.wait::ld + tcgen05.fence::before_thread_sync on this side and the fence::after_thread_sync behind the
// leader's wait, so the arrive itself needs no cluster-scope release: `.release.cluster` compiled to MEMBAR.ALL + ERRBAR in
// front of every arrive and was ~45 % of the epilogue warps' stall samples (profiles/r2_ncu_rowpair2_before.txt).
__device__ __forceinline__ void mbar_arrive_cluster_fwd(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
constexpr int HALO2_THREADS = 224;   // warps 0/1 producer + MMA, 2..5 epilogue, 6 weight-tile producer (streaming mode)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {   // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BLOCK_N, int RESIDENT_KC>
struct Halo2Cfg {
    static constexpr int HALO_STAGE_BYTES = HaloGeom<10>::STAGE_BYTES;
    static constexpr int B_TILE_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;      // this CTA's half of one weight tile
    static constexpr int RES_BYTES = 9 * RESIDENT_KC * B_TILE_BYTES;
    // streaming mode: one B stage = the three weight tiles of one filter row (3 taps), so a stage is worth
    // 12 MMAs (~900 cycles) and five stages cover the TMA latency comfortably
    static constexpr int B_STAGE_BYTES = 3 * B_TILE_BYTES;
    static constexpr int B_STAGES = RESIDENT_KC > 0 ? 0 : (BLOCK_N == 256 ? 3 : 5);
    static constexpr int STG_BYTES = 4 * 4096;
    static constexpr int BUDGET = 227 * 1024 - 4096 - 1024 - STG_BYTES;
    static constexpr int A_STAGES_RAW = (BUDGET - RES_BYTES - B_STAGES * B_STAGE_BYTES) / HALO_STAGE_BYTES;
    static constexpr int A_STAGES = A_STAGES_RAW > 4 ? 4 : A_STAGES_RAW;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = RES_BYTES + B_STAGES * B_STAGE_BYTES + A_STAGES * HALO_STAGE_BYTES + STG_BYTES + 4096 + 1024;
    static_assert(A_STAGES >= 2, "halo2 kernel needs at least two A stages");
};

template <int BLOCK_N, int EPI, int RESIDENT_KC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO2_THREADS, 1)
conv_halo2_kernel(const __grid_constant__ CUtensorMap map_a_halo, const __grid_constant__ CUtensorMap map_b_half,
                  const __grid_constant__ CUtensorMap map_out, const ConvArgs args) {
    using C = Halo2Cfg<BLOCK_N, RESIDENT_KC>;
    constexpr int HALO_STAGE_BYTES = C::HALO_STAGE_BYTES;
    constexpr int HALO_PITCH = 10;
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs: the UMMA descriptors and multicast barrier addresses are CTA-relative offsets
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_res = smem;
    uint8_t* s_b = smem + C::RES_BYTES;
    uint8_t* s_a = s_b + C::B_STAGES * C::B_STAGE_BYTES;
    uint8_t* s_stg = s_a + C::A_STAGES * HALO_STAGE_BYTES;
    uint8_t* aux = s_stg + C::STG_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* b_empty = b_full + 8;
    uint64_t* res_full = b_empty + 8;
    uint64_t* tmem_full = res_full + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_head = reinterpret_cast<float*>(aux + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_x = args.W / HALO_TW, tiles_y = args.H / HALO_TH;
    const int n_tiles = args.n_total / BLOCK_N;                        // > 1 only in streaming mode (Cout = 512 / 1024)
    const int total_pairs = ((args.batch * tiles_y * tiles_x) >> 1) * n_tiles;   // work items: (tile pair, n tile); even tile count
    const int kchunks = args.Cin / BLOCK_K;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_halo);
        prefetch_tmap(&map_b_half);
        if (EPI != EPI_HEAD) prefetch_tmap(&map_out);
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(res_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (EPI == EPI_HEAD) {
        for (int i = threadIdx.x; i < args.n_classes * 64 + args.n_classes; i += HALO2_THREADS)
            s_head[i] = i < args.n_classes * 64 ? args.head_w[i] : args.head_b[i - args.n_classes * 64];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // barrier inits of both CTAs are visible before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs)
        if (lane == 0 && pair_id < total_pairs) {
            const int n_half0 = (int)rank * (BLOCK_N / 2);
            if (RESIDENT_KC > 0) {
                if (leader) mbar_expect_tx(res_full, 2 * C::RES_BYTES);
                for (int i = 0; i < 9 * RESIDENT_KC; ++i)
                    tma2_load_2d(s_res + i * C::B_TILE_BYTES, &map_b_half, res_full, (i / RESIDENT_KC) * args.Cin + (i % RESIDENT_KC) * BLOCK_K, n_half0);
            }
            int sa = 0;
            uint32_t pa = 0;
            pdl_wait();            // weights are in flight; the activations are the previous layer's output
            for (int p = pair_id; p < total_pairs; p += n_pairs) {
                const TileCoord tc = decode_tile(2 * (p / n_tiles) + (int)rank, 1, tiles_x, tiles_y, BLOCK_N, HALO_TW, HALO_TH);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    if (leader) mbar_expect_tx(&a_full[sa], 2 * HALO_BOX_BYTES);
                    tma2_load_4d(s_a + sa * HALO_STAGE_BYTES, &map_a_halo, &a_full[sa], kc * BLOCK_K, tc.x0 - 1, tc.y0 - 1, tc.b);
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == 6) {
        // ===================================================================== weight-tile producer (streaming mode)
        // Its own warp, so the A ring (one halo per 64-channel chunk) and the B ring (one half tile per tap) run
        // ahead independently instead of the halo of chunk c+1 queueing behind the nine weight tiles of chunk c.
        if (RESIDENT_KC == 0 && lane == 0 && pair_id < total_pairs) {
            const int n_half0 = (int)rank * (BLOCK_N / 2);
            int sb = 0;
            uint32_t pb = 0;
            for (int p = pair_id; p < total_pairs; p += n_pairs) {
                const int n_row0 = (p % n_tiles) * BLOCK_N + n_half0;
                for (int kc = 0; kc < kchunks; ++kc) {
                    for (int row = 0; row < 3; ++row) {
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        if (leader) mbar_expect_tx(&b_full[sb], 2 * C::B_STAGE_BYTES);
                        for (int j = 0; j < 3; ++j)
                            tma2_load_2d(s_b + sb * C::B_STAGE_BYTES + j * C::B_TILE_BYTES, &map_b_half, &b_full[sb],
                                         (row * 3 + j) * args.Cin + kc * BLOCK_K, n_row0);
                        if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only; the whole
        // warp walks the pipeline, one elected lane issues, taps are straight-line code with immediate offsets)
        if (leader && pair_id < total_pairs) {
            constexpr uint32_t idesc = make_idesc_m256(BLOCK_N);
            int sa = 0, sb = 0, acc = 0;
            uint32_t pa = 0, pb = 0, acc_phase = 0;
            if (RESIDENT_KC > 0) {
                mbar_wait(res_full, 0);
                tc_fence_after();
            }
            for (int p = pair_id; p < total_pairs; p += n_pairs) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    const uint64_t a0 = make_smem_desc_sbo(smem_u32(s_a + sa * HALO_STAGE_BYTES), HALO_PITCH * 128);
                    constexpr int kAUnit = 128 / 16;
                    if (RESIDENT_KC > 0) {
                        if (elect_one()) {
                            const uint64_t b0 = make_smem_desc(smem_u32(s_res + kc * C::B_TILE_BYTES));
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint64_t adesc = a0 + (uint64_t)(((tap / 3) * HALO_PITCH + (tap % 3)) * kAUnit);
                                const uint64_t bdesc = b0 + (uint64_t)(tap * RESIDENT_KC * (C::B_TILE_BYTES / 16));
#pragma unroll
                                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                    umma2_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                              (tap | k) != 0 ? 1u : (kc != 0 ? 1u : 0u));
                            }
                            umma2_commit_mc(&a_empty[sa]);
                            if (kc == kchunks - 1) umma2_commit_mc(&tmem_full[acc]);
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int row = 0; row < 3; ++row) {                 // one B stage = the three taps of a filter row
                            mbar_wait(&b_full[sb], pb);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t b0 = make_smem_desc(smem_u32(s_b + sb * C::B_STAGE_BYTES));
#pragma unroll
                                for (int j = 0; j < 3; ++j) {
                                    const uint64_t adesc = a0 + (uint64_t)((row * HALO_PITCH + j) * kAUnit);
                                    const uint64_t bdesc = b0 + (uint64_t)(j * (C::B_TILE_BYTES / 16));
#pragma unroll
                                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                        umma2_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                                  (row | j | k) != 0 ? 1u : (kc != 0 ? 1u : 0u));
                                }
                                umma2_commit_mc(&b_empty[sb]);
                                if (row == 2) {
                                    umma2_commit_mc(&a_empty[sa]);
                                    if (kc == kchunks - 1) umma2_commit_mc(&tmem_full[acc]);
                                }
                            }
                            __syncwarp();
                            if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
                        }
                    }
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue (both CTAs, own 128 rows)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int ly = row / HALO_TW, lx = row % HALO_TW;
        EpiCtx e;
        e.map_out = &map_out;
        e.slab = smem_u32(s_stg + quarter * 4096);
        e.slab_y = quarter * (32 / HALO_TW);
        e.lane = lane;
        e.n0 = 0;
        const uint32_t empty0 = mapa_rank(smem_u32(&tmem_empty[0]), 0), empty1 = mapa_rank(smem_u32(&tmem_empty[1]), 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int p = pair_id; p < total_pairs; p += n_pairs) {
            const TileCoord tcd = decode_tile(2 * (p / n_tiles) + (int)rank, 1, tiles_x, tiles_y, BLOCK_N, HALO_TW, HALO_TH);
            e.b = tcd.b; e.y0 = tcd.y0; e.x0 = tcd.x0; e.y = tcd.y0 + ly; e.x = tcd.x0 + lx;
            e.n0 = (p % n_tiles) * BLOCK_N;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            epilogue_tile<BLOCK_N, EPI, HALO_TW, true>(args, s_head, taddr, e, nullptr, acc ? empty1 : empty0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI != EPI_HEAD && lane == 0) tma_store_wait_read();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // neither CTA may free TMEM / exit while its partner can still touch it
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// ==================================================================================== kernel 4
// Transposed conv 2x2 s2 as a cta_group::2 GEMM (M = pixels, N = 4 * Cout, K = Cin, one tap).  The per-tap kernel 1 runs
// the up-sampling layers L2 -> SM bound: every CTA re-fetches the whole 256-row weight tile for each of its pixel tiles
// (48 KiB of operands per k-step for 128 x 256 x 64 MACs).  Here a CTA pair works on two adjacent 8 x 16-pixel tiles as ONE
// UMMA of M = 256: each CTA stages its own 128 pixel rows (16 KiB) and only HALF of the weight tile (BLOCK_N / 2 rows,
// 16 KiB), so a k-step moves 32 KiB per CTA for twice the MACs per byte of weights.  Pipeline, barriers and epilogue are
// those of kernel 3 (both CTAs' TMA loads signal the leader's barriers, tcgen05.commit multicasts to both CTAs, epilogues
// report back over DSMEM); the epilogue scatters through the 5-D tensor map of the (2H x 2W) destination.
template <int BLOCK_N>
struct ConvT2Cfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;                  // this CTA's 128 pixel rows of one K chunk
    static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;            // this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STG_BYTES = 4 * 4096;
    static constexpr int STAGES_RAW = (227 * 1024 - 4096 - 1024 - STG_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 4096 + 1024;
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
convt_pair_kernel(const __grid_constant__ CUtensorMap map_a_tile, const __grid_constant__ CUtensorMap map_b_half,
                  const __grid_constant__ CUtensorMap map_out, const ConvArgs args) {
    using C = ConvT2Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_stg = smem + C::STAGES * C::STAGE_BYTES;
    uint8_t* aux = s_stg + C::STG_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full = empty_bar + 8;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_x = args.W / HALO_TW, tiles_y = args.H / HALO_TH;
    const int n_tiles = args.n_total / BLOCK_N;
    const int total_pairs = ((args.batch * tiles_y * tiles_x) >> 1) * n_tiles;   // work items: (tile pair, n tile)
    const int kchunks = args.Cin / BLOCK_K;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_tile);
        prefetch_tmap(&map_b_half);
        prefetch_tmap(&map_out);
        for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs)
        if (lane == 0 && pair_id < total_pairs) {
            int stage = 0;
            uint32_t phase = 0;
            pdl_wait();            // the activations are the previous layer's output
            for (int p = pair_id; p < total_pairs; p += n_pairs) {
                const TileCoord tc = decode_tile(2 * (p / n_tiles) + (int)rank, 1, tiles_x, tiles_y, BLOCK_N, HALO_TW, HALO_TH);
                const int n_row0 = (p % n_tiles) * BLOCK_N + (int)rank * (BLOCK_N / 2);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                    tma2_load_4d(sa, &map_a_tile, &full_bar[stage], kc * BLOCK_K, tc.x0, tc.y0, tc.b);
                    tma2_load_2d(sa + C::A_BYTES, &map_b_half, &full_bar[stage], kc * BLOCK_K, n_row0);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only)
        if (leader && pair_id < total_pairs) {
            constexpr uint32_t idesc = make_idesc_m256(BLOCK_N);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int p = pair_id; p < total_pairs; p += n_pairs) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa);
                        const uint64_t bdesc = make_smem_desc(sa + C::A_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma2_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
                        umma2_commit_mc(&empty_bar[stage]);
                        if (kc == kchunks - 1) umma2_commit_mc(&tmem_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue (both CTAs, own 128 rows)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int ly = row / HALO_TW, lx = row % HALO_TW;
        EpiCtx e;
        e.map_out = &map_out;
        e.slab = smem_u32(s_stg + quarter * 4096);
        e.slab_y = quarter * (32 / HALO_TW);
        e.lane = lane;
        const uint32_t empty0 = mapa_rank(smem_u32(&tmem_empty[0]), 0), empty1 = mapa_rank(smem_u32(&tmem_empty[1]), 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int p = pair_id; p < total_pairs; p += n_pairs) {
            const TileCoord tcd = decode_tile(2 * (p / n_tiles) + (int)rank, 1, tiles_x, tiles_y, BLOCK_N, HALO_TW, HALO_TH);
            e.b = tcd.b; e.y0 = tcd.y0; e.x0 = tcd.x0; e.y = tcd.y0 + ly; e.x = tcd.x0 + lx;
            e.n0 = (p % n_tiles) * BLOCK_N;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            epilogue_tile<BLOCK_N, EPI_CONVT, HALO_TW, true>(args, nullptr, taddr, e, nullptr, acc ? empty1 : empty0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_read();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}


// ==================================================================================== kernel 5
// Row-pair halo kernel for the Cout = 64 layers (enc1b, dec1a, dec1b + head), which kernels 2 / 3 leave on the per-instruction
// floor: an M = 128 UMMA costs ~64 - 72 cycles whether N is 64 or 128 (profiles/r1_mma_rate_microbench.log; what it waits for is
// its operand read from shared memory, see kernel 6), so at N = 64 half of every instruction slot is empty and those layers
// cannot exceed ~45 % of the tensor peak with one output pixel per GEMM row.  Here ONE GEMM row carries TWO output
// pixels, (y, x) and (y + 1, x): accumulator columns [0, 64) are the 64 channels of the even output row, [64, 128) those of
// the odd row below it.  An A operand read at input-row shift a (relative to the even row) feeds tap dy = a of the even row
// and tap dy = a - 1 of the odd row, so with the weight tiles of one filter column laid out as [W(dy=+1) | W(0) | W(-1)]
// (64 rows each, contiguous) the twelve (dy, row) products of a filter column become FOUR instructions:
//     a =  0 : N = 128, B = [W(0) ; W(-1)]     a = +1 : N = 128, B = [W(+1) ; W(0)]
//     a = -1 : N =  64, B = W(-1) -> cols [0, 64)      a = +2 : N = 64, B = W(+1) -> cols [64, 128)
// i.e. 4 issue slots per 256 pixels instead of 6: 1.5x the MMA-bound rate of kernel 2.  The tile is 8 px x 32 rows; GEMM row
// group g (8 px) is output rows y0 + 2g, y0 + 2g + 1, so consecutive row groups are TWO halo rows apart: the A descriptors
// are those of kernel 2 with SBO = 2 * 1280 B.  One TMA box {64 ch, 10 px, 34 rows} per 64-channel chunk; weights resident.
// Epilogue: lane = (row group, px) holds both pixels of a column pair, so the 2 x 2 max-pool is one in-lane max plus one
// exchange with lane ^ 1; an epilogue warp stages its 8 rows x 8 px x 64 ch (8 KiB) and writes them with ONE TMA store.
constexpr int RP_TW = 8, RP_TH = 32;
constexpr int RP_HALO_ROWS = RP_TH + 2;
constexpr int RP_HALO_BOX_BYTES = RP_HALO_ROWS * (RP_TW + 2) * 128;     // 43,520 B per 64-channel chunk
template <int EPI, int RESIDENT_KC>
struct RowPairCfg {
    static constexpr int HALO_STAGE_BYTES = (RP_HALO_BOX_BYTES + 1023) / 1024 * 1024;
    static constexpr int W_TILE_BYTES = 64 * BLOCK_K * 2;                // 64 output channels x 64 k
    static constexpr int RES_BYTES = 9 * RESIDENT_KC * W_TILE_BYTES;
    // RESIDENT_KC == 0 (Cin > 64: the weight set does not fit beside two halo stages): the three tiles of one filter column
    // and chunk ([W(+1) | W(0) | W(-1)], 24 KiB = 16 MMAs) stream through their own ring, fed by their own producer warp
    static constexpr int B_STAGE_BYTES = 3 * W_TILE_BYTES;
    static constexpr int B_STAGES = RESIDENT_KC > 0 ? 0 : 3;
    // one 4 KiB staging slab per epilogue warp (the four rows of one parity; none for the head, which stores no feature map)
    // two 4 KiB staging slabs per epilogue warp (the four rows of each parity; none for the head, which stores no feature map)
    static constexpr int STG_BYTES = EPI == EPI_HEAD ? 0 : 4 * 10240;   // per warp: two 4 KiB row slabs + a 2 KiB pooled slab
    // halo stages: two where the slabs take 32 KiB (a third stage measured no faster: the loads are not the bound)
    static constexpr int A_STAGES = EPI == EPI_HEAD ? 3 : 2;
    static constexpr int TMEM_COLS = 256;                                // two accumulators of 128 columns
    static constexpr int THREADS = RESIDENT_KC > 0 ? NUM_THREADS : NUM_THREADS + 32;
    static constexpr int SMEM_BYTES = RES_BYTES + B_STAGES * B_STAGE_BYTES + A_STAGES * HALO_STAGE_BYTES + STG_BYTES + 4096 + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "row-pair kernel: weights do not fit beside two halo stages");
};

// 64 -> n_classes head (src/process.cpp:158-170) of the TWO pixels of a row-pair lane, (y, x) and (y + 1, x), on their fp32
// features (bias + ReLU already applied).  Both pixels share every broadcast weight load and run eight independent FMA chains
// (four partial sums per pixel): with one pixel at a time the multi-class head waited on its shared-memory weight loads
// (ncu: long-scoreboard stalls on the FFMAs) and ran the layer at half the rate of the binary head.
__device__ __forceinline__ void head_pixel_pair(const ConvArgs& args, const float* s_head, const float (&f0)[64], const float (&f1)[64],
                                                int b, int y, int x) {
    const size_t plane = (size_t)args.H * args.W;
    const size_t pix = (size_t)b * plane + (size_t)y * args.W + x;
    float best0 = -3.402823466e+38f, best1 = -3.402823466e+38f;  // -FLT_MAX, src/process.cpp:159
    int c0 = 0, c1 = 0;
    for (int c = 0; c < args.n_classes; ++c) {
        const float hb = s_head[args.n_classes * 64 + c];
        float p0 = hb, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f, q0 = hb, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
        const float4* w4 = reinterpret_cast<const float4*>(s_head + c * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 w = w4[j];
            p0 = fmaf(f0[4 * j + 0], w.x, p0); q0 = fmaf(f1[4 * j + 0], w.x, q0);
            p1 = fmaf(f0[4 * j + 1], w.y, p1); q1 = fmaf(f1[4 * j + 1], w.y, q1);
            p2 = fmaf(f0[4 * j + 2], w.z, p2); q2 = fmaf(f1[4 * j + 2], w.z, q2);
            p3 = fmaf(f0[4 * j + 3], w.w, p3); q3 = fmaf(f1[4 * j + 3], w.w, q3);
        }
        const float s0 = (p0 + p1) + (p2 + p3), s1 = (q0 + q1) + (q2 + q3);
        if (args.logits) {
            float* lg = args.logits + ((size_t)b * args.n_classes + c) * plane + (size_t)y * args.W + x;
            lg[0] = s0;
            lg[args.W] = s1;
        }
        if (s0 > best0) { best0 = s0; c0 = c; }   // strict >: first max wins, NaN never wins
        if (s1 > best1) { best1 = s1; c1 = c; }
    }
    const bool binary = args.n_classes == 1;
    args.mask[pix] = binary ? (uint8_t)(best0 > 0.0f ? args.fg_value : 0) : (uint8_t)c0;
    args.mask[pix + args.W] = binary ? (uint8_t)(best1 > 0.0f ? args.fg_value : 0) : (uint8_t)c1;
}

// Epilogue of one row-pair tile for one epilogue warp: lane = (row group gl, px) owns output pixels (yw + 2 gl, x0 + px) in
// accumulator columns [0, 64) and (yw + 2 gl + 1, x0 + px) in [64, 128).  The epilogue, not the MMA pipe, bounds these layers
// (stripped of its math the kernel runs at ~1.6 PFLOP/s), and it is LATENCY bound -- one warp per scheduler, every TMEM load,
// store-read wait and proxy fence exposed -- so the tile is handled in one pass: all four TMEM loads in flight together and
// the accumulator handed back before any math, the bias from shared memory, one slab per row parity (the wait for the
// previous tile's TMA stores is long satisfied), one proxy fence and both stores issued together.
template <int EPI, bool REMOTE>
__device__ __forceinline__ void rowpair_epilogue_tile(const ConvArgs& args, const CUtensorMap* map_out, const CUtensorMap* map_pool,
                                                      const float* s_head, const float* s_bias,
                                                      uint32_t taddr, uint32_t slab, const TileCoord& tcd, int yw, int lane,
                                                      uint64_t* tmem_empty_bar, uint32_t remote_empty) {
    const int gl = lane >> 3, px = lane & 7;
    uint32_t r[4][32];
    tmem_ld32(taddr, r[0]);
    tmem_ld32(taddr + 32, r[1]);
    tmem_ld32(taddr + 64, r[2]);
    tmem_ld32(taddr + 96, r[3]);
    tmem_ld_wait();
    tc_fence_before();
    if (REMOTE) mbar_arrive_cluster_fwd(remote_empty); else mbar_arrive(tmem_empty_bar);
    const float4* b4 = reinterpret_cast<const float4*>(s_bias);     // broadcast 16-byte loads
    if (EPI == EPI_HEAD) {
        float f[2][64];
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 bb = b4[j];
                const uint32_t* q = &r[2 * half + (j >> 3)][4 * (j & 7)];
                f[half][4 * j + 0] = fmaxf(__uint_as_float(q[0]) + bb.x, 0.0f);
                f[half][4 * j + 1] = fmaxf(__uint_as_float(q[1]) + bb.y, 0.0f);
                f[half][4 * j + 2] = fmaxf(__uint_as_float(q[2]) + bb.z, 0.0f);
                f[half][4 * j + 3] = fmaxf(__uint_as_float(q[3]) + bb.w, 0.0f);
            }
        head_pixel_pair(args, s_head, f[0], f[1], tcd.b, yw + 2 * gl, tcd.x0 + px);
        return;
    }
    uint32_t pk[2][32];                                              // pk[half][i] = channels 2 i, 2 i + 1
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 bb = b4[j];
            const uint32_t* q = &r[2 * half + (j >> 3)][4 * (j & 7)];
            pk[half][2 * j] = pack_bf16(fmaxf(__uint_as_float(q[0]) + bb.x, 0.0f), fmaxf(__uint_as_float(q[1]) + bb.y, 0.0f));
            pk[half][2 * j + 1] = pack_bf16(fmaxf(__uint_as_float(q[2]) + bb.z, 0.0f), fmaxf(__uint_as_float(q[3]) + bb.w, 0.0f));
        }
    // 2 x 2 max-pool in registers: rows (2 gl, 2 gl + 1) are this lane's two halves, the column partner is lane ^ 1.  The even
    // lane finishes channels [0, 32), the odd lane [32, 64): each sends the half the other one needs.
    const bool odd = lane & 1;
    uint32_t res[16];
    if (args.pool) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t lo = max_bf16x2(pk[0][i], pk[1][i]), hi = max_bf16x2(pk[0][16 + i], pk[1][16 + i]);
            const uint32_t got = __shfl_xor_sync(0xFFFFFFFFu, odd ? lo : hi, 1);
            res[i] = max_bf16x2(odd ? hi : lo, got);
        }
    }
    // the previous tile's TMA stores must have finished reading the slabs
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        // slab row = (row group) * 8 + px; 16-byte chunk j of a row lives at j ^ (row & 7)
        const uint32_t row_addr = slab + (uint32_t)(half * 4096 + lane * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            st_shared_v4(row_addr + (uint32_t)((j ^ px) << 4), pk[half][4 * j], pk[half][4 * j + 1], pk[half][4 * j + 2], pk[half][4 * j + 3]);
    }
    if (args.pool) {
        // pooled slab (2 KiB behind the two row slabs): row = gl * 4 + px / 2, this lane's four 16-byte chunks.  The pooled map
        // leaves through TMA like the rest: a generic st.global here would still be in flight at the NEXT tile's proxy fence
        // (MEMBAR.ALL.CTA), which then waits out an HBM write latency per tile.
        const int prow = gl * 4 + (px >> 1);
        const uint32_t paddr = slab + 8192u + (uint32_t)(prow * 128);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st_shared_v4(paddr + (uint32_t)((((odd ? 4 : 0) + i) ^ (prow & 7)) << 4), res[4 * i], res[4 * i + 1], res[4 * i + 2], res[4 * i + 3]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        // destination viewed as (C, W, row parity, H / 2, B): box {64 ch, 8 px, 1, 4 row pairs, 1} = rows yw + half + 2 g
        tma_store_5d(map_out, slab, args.out_coff, tcd.x0, 0, yw >> 1, tcd.b);
        tma_store_5d(map_out, slab + 4096u, args.out_coff, tcd.x0, 1, yw >> 1, tcd.b);
        if (args.pool) tma_store_4d(map_pool, slab + 8192u, 0, tcd.x0 >> 1, yw >> 1, tcd.b);   // box {64 ch, 4 px, 4 rows, 1}
        tma_store_commit();
    }
}

template <int EPI, int RESIDENT_KC>
__global__ void __launch_bounds__(RowPairCfg<EPI, RESIDENT_KC>::THREADS, 1)
conv_rowpair_kernel(const __grid_constant__ CUtensorMap map_a_halo, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_pool, const ConvArgs args) {
    using C = RowPairCfg<EPI, RESIDENT_KC>;
    static_assert(EPI == EPI_STORE || EPI == EPI_HEAD, "row-pair kernel: conv3x3 layers only");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_res = smem;                                            // [filter column][chunk][dy = +1, 0, -1] weight tiles
    uint8_t* s_b = smem + C::RES_BYTES;                               // streamed weight ring (RESIDENT_KC == 0)
    uint8_t* s_a = s_b + C::B_STAGES * C::B_STAGE_BYTES;              // halo ring
    uint8_t* s_stg = s_a + C::A_STAGES * C::HALO_STAGE_BYTES;         // output staging slabs
    uint8_t* aux = s_stg + C::STG_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* b_empty = b_full + 4;
    uint64_t* res_full = b_empty + 4;
    uint64_t* tmem_full = res_full + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_head = reinterpret_cast<float*>(aux + 512);
    float* s_bias = reinterpret_cast<float*>(aux + 3072);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = args.W / RP_TW, tiles_y = args.H / RP_TH;
    const int total = args.batch * tiles_y * tiles_x;
    const int kchunks = args.Cin / BLOCK_K;                           // == RESIDENT_KC when the weights are resident

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_halo);
        prefetch_tmap(&map_b);
        if (EPI != EPI_HEAD) { prefetch_tmap(&map_out); prefetch_tmap(&map_pool); }
        for (int i = 0; i < C::A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(res_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_warp(tmem_ptr, C::TMEM_COLS);
    if (EPI == EPI_HEAD) {
        for (int i = threadIdx.x; i < args.n_classes * 64 + args.n_classes; i += C::THREADS)
            s_head[i] = i < args.n_classes * 64 ? args.head_w[i] : args.head_b[i - args.n_classes * 64];
    }
    if (threadIdx.x < 64) s_bias[threadIdx.x] = args.bias[threadIdx.x];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();

    if (warp == 6) {
        // ===================================================================== weight producer (streaming mode only)
        if (RESIDENT_KC == 0 && lane == 0 && blockIdx.x < total) {
            int sb = 0;
            uint32_t pb = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x)
                for (int kc = 0; kc < kchunks; ++kc)
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        mbar_expect_tx(&b_full[sb], C::B_STAGE_BYTES);
                        for (int j = 0; j < 3; ++j)              // j = 0, 1, 2 <-> dy = +1, 0, -1
                            tma_load_2d(s_b + sb * C::B_STAGE_BYTES + j * C::W_TILE_BYTES, &map_b, &b_full[sb],
                                        ((2 - j) * 3 + dxi) * args.Cin + kc * BLOCK_K, 0);
                        if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
                    }
        }
    } else if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0 && blockIdx.x < total) {
            if (RESIDENT_KC > 0) {
                mbar_expect_tx(res_full, C::RES_BYTES);
                for (int i = 0; i < 9 * RESIDENT_KC; ++i) {           // slot i = (dx, chunk, j): j = 0, 1, 2 <-> dy = +1, 0, -1
                    const int dxi = i / (3 * RESIDENT_KC), kc = (i / 3) % RESIDENT_KC, j = i % 3;
                    const int tap = (2 - j) * 3 + dxi;
                    tma_load_2d(s_res + i * C::W_TILE_BYTES, &map_b, res_full, tap * args.Cin + kc * BLOCK_K, 0);
                }
            }
            int sa = 0;
            uint32_t pa = 0;
            pdl_wait();            // weights are in flight; the activations are the previous layer's output
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, 1, tiles_x, tiles_y, 64, RP_TW, RP_TH);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    mbar_expect_tx(&a_full[sa], RP_HALO_BOX_BYTES);
                    tma_load_4d(s_a + sa * C::HALO_STAGE_BYTES, &map_a_halo, &a_full[sa], kc * BLOCK_K, tc.x0 - 1, tc.y0 - 1, tc.b);
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (blockIdx.x < total) {
            constexpr uint32_t idesc128 = make_idesc(128), idesc64 = make_idesc(64);
            constexpr int kPitch = RP_TW + 2;                           // halo row = 10 rows of 128 B
            constexpr int kAUnit = 128 / 16;                            // one 128-byte row in descriptor units
            constexpr int kWUnit = C::W_TILE_BYTES / 16;
            int sa = 0, sb = 0, acc = 0;
            uint32_t pa = 0, pb = 0, acc_phase = 0;
            if (RESIDENT_KC > 0) {
                mbar_wait(res_full, 0);
                tc_fence_after();
            }
            // the 16 instructions of one filter column and chunk; a = 0 first: its N = 128 form initialises both halves of a
            // fresh accumulator
            auto issue_column = [&](uint32_t d_tmem, uint64_t a_col, uint64_t w0, bool fresh) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t ar = a_col + (uint64_t)(2 * k);      // + (a + 1) halo rows below
                    const uint64_t wk = w0 + (uint64_t)(2 * k);
                    umma_f16(d_tmem, ar + (uint64_t)(1 * kPitch * kAUnit), wk + (uint64_t)kWUnit, idesc128, (k != 0 || !fresh) ? 1u : 0u);
                    umma_f16(d_tmem, ar + (uint64_t)(2 * kPitch * kAUnit), wk, idesc128, 1u);
                    umma_f16(d_tmem, ar, wk + (uint64_t)(2 * kWUnit), idesc64, 1u);
                    umma_f16(d_tmem + 64u, ar + (uint64_t)(3 * kPitch * kAUnit), wk, idesc64, 1u);
                }
            };
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    // row group g of the A operand = halo rows 2g + (a + 1): SBO = two halo rows
                    const uint64_t a0 = make_smem_desc_sbo(smem_u32(s_a + sa * C::HALO_STAGE_BYTES), 2 * kPitch * 128);
                    if (RESIDENT_KC > 0) {
                        if (elect_one()) {
#pragma unroll
                            for (int dxi = 0; dxi < 3; ++dxi)
                                issue_column(d_tmem, a0 + (uint64_t)(dxi * kAUnit),
                                             make_smem_desc(smem_u32(s_res + (dxi * RESIDENT_KC + kc) * 3 * C::W_TILE_BYTES)), dxi == 0 && kc == 0);
                            umma_commit(&a_empty[sa]);
                            if (kc == kchunks - 1) umma_commit(&tmem_full[acc]);
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int dxi = 0; dxi < 3; ++dxi) {
                            mbar_wait(&b_full[sb], pb);
                            tc_fence_after();
                            if (elect_one()) {
                                issue_column(d_tmem, a0 + (uint64_t)(dxi * kAUnit), make_smem_desc(smem_u32(s_b + sb * C::B_STAGE_BYTES)),
                                             dxi == 0 && kc == 0);
                                umma_commit(&b_empty[sb]);
                                if (dxi == 2) {
                                    umma_commit(&a_empty[sa]);
                                    if (kc == kchunks - 1) umma_commit(&tmem_full[acc]);
                                }
                            }
                            __syncwarp();
                            if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
                        }
                    }
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue (warps 2..5): lane = (row group, px)
        const int quarter = warp & 3;
        const uint32_t slab = smem_u32(s_stg + quarter * 10240);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const TileCoord tcd = decode_tile(t, 1, tiles_x, tiles_y, 64, RP_TW, RP_TH);
            const int yw = tcd.y0 + quarter * 8;                       // first of this warp's 8 output rows
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 128);
            rowpair_epilogue_tile<EPI, false>(args, &map_out, &map_pool, s_head, s_bias, taddr, slab, tcd, yw, lane, &tmem_empty[acc], 0u);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI != EPI_HEAD && lane == 0) tma_store_wait_read();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_warp(tmem_base, C::TMEM_COLS);
    }
}


// ==================================================================================== kernel 6
// cta_group::2 form of the row-pair kernel.  Kernel 5 is not bound by the MMA rate but by SHARED-MEMORY BANDWIDTH: an M = 128
// instruction reads 4 KiB of A and N x 32 B of B, so at N <= 128 the operand fetch (6 - 8 KiB per instruction against
// ~128 B / cycle) takes as long as the math, and the halo writes and the epilogue's staging traffic come on top
// (tools/microbench/mma_rowpair*.cu: 63 cycles per instruction of the mix with idle shared memory, ~95 inside the kernel).
// As a CTA pair (M = 256: each CTA its own 8 x 32-pixel tile) every CTA supplies only HALF of each B operand: per filter
// column and K-slice a CTA reads 4 x 4 KiB of A + (2 + 2 + 1 + 1) KiB of B = 22 KiB instead of 28.  The B halves fall out of
// the column layout for free: for the N = 128 instructions CTA 0 holds the tile of the even output row and CTA 1 that of the
// odd row ([W(0) ; W(-1)] and [W(+1) ; W(0)] are split exactly there); for the N = 64 instructions each CTA holds 32 of the
// 64 output channels.  Per (filter column, chunk) a CTA stores S0 | S1 | S2 | S3 = 8 + 8 + 4 + 4 KiB:
//     CTA 0: W(+1) | W(0)  | W(-1)[0:32]  | W(+1)[0:32]        CTA 1: W(0) | W(-1) | W(-1)[32:64] | W(+1)[32:64]
//     a = +1 reads S0 (N = 128), a = 0 reads S1 (N = 128), a = -1 reads S2 (N = 64), a = +2 reads S3 (N = 64, columns [64, 128)).
// Pipeline, barriers and the remote "accumulator drained" arrive are those of kernel 3.
template <int EPI, int RESIDENT_KC>
struct RowPair2Cfg {
    static constexpr int HALO_STAGE_BYTES = (RP_HALO_BOX_BYTES + 1023) / 1024 * 1024;
    static constexpr int COL_BYTES = 3 * 64 * BLOCK_K * 2;               // S0 | S1 | S2 | S3 of one filter column and chunk: 24 KiB
    static constexpr int RES_BYTES = 3 * RESIDENT_KC * COL_BYTES;
    static constexpr int B_STAGES = RESIDENT_KC > 0 ? 0 : 3;
    static constexpr int STG_BYTES = EPI == EPI_HEAD ? 0 : 4 * 10240;   // per warp: two 4 KiB row slabs + a 2 KiB pooled slab
    static constexpr int A_STAGES = EPI == EPI_HEAD ? 3 : 2;
    static constexpr int TMEM_COLS = 256;
    static constexpr int SMEM_BYTES = RES_BYTES + B_STAGES * COL_BYTES + A_STAGES * HALO_STAGE_BYTES + STG_BYTES + 4096 + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "row-pair pair kernel: shared memory");
};

template <int EPI, int RESIDENT_KC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO2_THREADS, 1)
conv_rowpair2_kernel(const __grid_constant__ CUtensorMap map_a_halo, const __grid_constant__ CUtensorMap map_b64,
                     const __grid_constant__ CUtensorMap map_b32, const __grid_constant__ CUtensorMap map_out,
                     const __grid_constant__ CUtensorMap map_pool, const ConvArgs args) {
    using C = RowPair2Cfg<EPI, RESIDENT_KC>;
    static_assert(EPI == EPI_STORE || EPI == EPI_HEAD, "row-pair kernel: conv3x3 layers only");
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs: the UMMA descriptors and multicast barrier addresses are CTA-relative offsets
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_res = smem;
    uint8_t* s_b = smem + C::RES_BYTES;
    uint8_t* s_a = s_b + C::B_STAGES * C::COL_BYTES;
    uint8_t* s_stg = s_a + C::A_STAGES * C::HALO_STAGE_BYTES;
    uint8_t* aux = s_stg + C::STG_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* b_empty = b_full + 4;
    uint64_t* res_full = b_empty + 4;
    uint64_t* tmem_full = res_full + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_head = reinterpret_cast<float*>(aux + 512);
    float* s_bias = reinterpret_cast<float*>(aux + 3072);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_x = args.W / RP_TW, tiles_y = args.H / RP_TH;
    const int total_pairs = (args.batch * tiles_y * tiles_x) >> 1;    // tiles 2p, 2p + 1 are neighbours in x (even tile count)
    const int kchunks = args.Cin / BLOCK_K;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a_halo);
        prefetch_tmap(&map_b64);
        prefetch_tmap(&map_b32);
        if (EPI != EPI_HEAD) { prefetch_tmap(&map_out); prefetch_tmap(&map_pool); }
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(res_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (EPI == EPI_HEAD) {
        for (int i = threadIdx.x; i < args.n_classes * 64 + args.n_classes; i += HALO2_THREADS)
            s_head[i] = i < args.n_classes * 64 ? args.head_w[i] : args.head_b[i - args.n_classes * 64];
    }
    if (threadIdx.x < 64) s_bias[threadIdx.x] = args.bias[threadIdx.x];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // barrier inits of both CTAs are visible before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();

    // this CTA's share of one filter column's weights: S0 | S1 | S2 | S3 (dy index j: 0, 1, 2 <-> dy = +1, 0, -1; tap = (2 - j) * 3 + dxi)
    auto load_column = [&](uint8_t* dst, uint64_t* bar, int dxi, int kc) {
        const int r = (int)rank, k0 = kc * BLOCK_K;
        tma2_load_2d(dst, &map_b64, bar, ((2 - r) * 3 + dxi) * args.Cin + k0, 0);                    // S0: j = r
        tma2_load_2d(dst + 8192, &map_b64, bar, ((1 - r) * 3 + dxi) * args.Cin + k0, 0);             // S1: j = 1 + r
        tma2_load_2d(dst + 16384, &map_b32, bar, (0 * 3 + dxi) * args.Cin + k0, 32 * r);             // S2: W(-1), channels 32 r ..
        tma2_load_2d(dst + 20480, &map_b32, bar, (2 * 3 + dxi) * args.Cin + k0, 32 * r);             // S3: W(+1), channels 32 r ..
    };

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs)
        if (lane == 0 && pair_id < total_pairs) {
            if (RESIDENT_KC > 0) {
                if (leader) mbar_expect_tx(res_full, 2 * C::RES_BYTES);
                for (int i = 0; i < 3 * RESIDENT_KC; ++i) load_column(s_res + i * C::COL_BYTES, res_full, i / RESIDENT_KC, i % RESIDENT_KC);
            }
            int sa = 0;
            uint32_t pa = 0;
            pdl_wait();            // weights are in flight; the activations are the previous layer's output
            for (int p = pair_id; p < total_pairs; p += n_pairs) {
                const TileCoord tc = decode_tile(2 * p + (int)rank, 1, tiles_x, tiles_y, 64, RP_TW, RP_TH);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    if (leader) mbar_expect_tx(&a_full[sa], 2 * RP_HALO_BOX_BYTES);
                    tma2_load_4d(s_a + sa * C::HALO_STAGE_BYTES, &map_a_halo, &a_full[sa], kc * BLOCK_K, tc.x0 - 1, tc.y0 - 1, tc.b);
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == 6) {
        // ===================================================================== weight producer (streaming mode, both CTAs)
        if (RESIDENT_KC == 0 && lane == 0 && pair_id < total_pairs) {
            int sb = 0;
            uint32_t pb = 0;
            for (int p = pair_id; p < total_pairs; p += n_pairs)
                for (int kc = 0; kc < kchunks; ++kc)
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        if (leader) mbar_expect_tx(&b_full[sb], 2 * C::COL_BYTES);
                        load_column(s_b + sb * C::COL_BYTES, &b_full[sb], dxi, kc);
                        if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only)
        if (leader && pair_id < total_pairs) {
            constexpr uint32_t idesc128 = make_idesc_m256(128), idesc64 = make_idesc_m256(64);
            constexpr int kPitch = RP_TW + 2;
            constexpr int kAUnit = 128 / 16;
            int sa = 0, sb = 0, acc = 0;
            uint32_t pa = 0, pb = 0, acc_phase = 0;
            if (RESIDENT_KC > 0) {
                mbar_wait(res_full, 0);
                tc_fence_after();
            }
            // the 16 instructions of one filter column and chunk; a = 0 first: its N = 128 form initialises both halves of a
            // fresh accumulator
            auto issue_column = [&](uint32_t d_tmem, uint64_t a_col, uint64_t w0, bool fresh) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t ar = a_col + (uint64_t)(2 * k);      // + (a + 1) halo rows below
                    const uint64_t wk = w0 + (uint64_t)(2 * k);
                    umma2_f16(d_tmem, ar + (uint64_t)(1 * kPitch * kAUnit), wk + (uint64_t)(8192 / 16), idesc128, (k != 0 || !fresh) ? 1u : 0u);
                    umma2_f16(d_tmem, ar + (uint64_t)(2 * kPitch * kAUnit), wk, idesc128, 1u);
                    umma2_f16(d_tmem, ar, wk + (uint64_t)(16384 / 16), idesc64, 1u);
                    umma2_f16(d_tmem + 64u, ar + (uint64_t)(3 * kPitch * kAUnit), wk + (uint64_t)(20480 / 16), idesc64, 1u);
                }
            };
            for (int p = pair_id; p < total_pairs; p += n_pairs) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    const uint64_t a0 = make_smem_desc_sbo(smem_u32(s_a + sa * C::HALO_STAGE_BYTES), 2 * kPitch * 128);
                    if (RESIDENT_KC > 0) {
                        if (elect_one()) {
#pragma unroll
                            for (int dxi = 0; dxi < 3; ++dxi)
                                issue_column(d_tmem, a0 + (uint64_t)(dxi * kAUnit),
                                             make_smem_desc(smem_u32(s_res + (dxi * RESIDENT_KC + kc) * C::COL_BYTES)), dxi == 0 && kc == 0);
                            umma2_commit_mc(&a_empty[sa]);
                            if (kc == kchunks - 1) umma2_commit_mc(&tmem_full[acc]);
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int dxi = 0; dxi < 3; ++dxi) {
                            mbar_wait(&b_full[sb], pb);
                            tc_fence_after();
                            if (elect_one()) {
                                issue_column(d_tmem, a0 + (uint64_t)(dxi * kAUnit), make_smem_desc(smem_u32(s_b + sb * C::COL_BYTES)),
                                             dxi == 0 && kc == 0);
                                umma2_commit_mc(&b_empty[sb]);
                                if (dxi == 2) {
                                    umma2_commit_mc(&a_empty[sa]);
                                    if (kc == kchunks - 1) umma2_commit_mc(&tmem_full[acc]);
                                }
                            }
                            __syncwarp();
                            if (++sb == C::B_STAGES) { sb = 0; pb ^= 1; }
                        }
                    }
                    if (++sa == C::A_STAGES) { sa = 0; pa ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue (both CTAs, own tile)
        const int quarter = warp & 3;
        const uint32_t slab = smem_u32(s_stg + quarter * 10240);
        const uint32_t empty0 = mapa_rank(smem_u32(&tmem_empty[0]), 0), empty1 = mapa_rank(smem_u32(&tmem_empty[1]), 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int p = pair_id; p < total_pairs; p += n_pairs) {
            const TileCoord tcd = decode_tile(2 * p + (int)rank, 1, tiles_x, tiles_y, 64, RP_TW, RP_TH);
            const int yw = tcd.y0 + quarter * 8;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 128);
            rowpair_epilogue_tile<EPI, true>(args, &map_out, &map_pool, s_head, s_bias, taddr, slab, tcd, yw, lane, nullptr, acc ? empty1 : empty0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI != EPI_HEAD && lane == 0) tma_store_wait_read();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // neither CTA may free TMEM / exit while its partner can still touch it
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace ms
