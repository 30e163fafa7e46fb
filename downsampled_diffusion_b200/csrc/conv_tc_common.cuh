// Shared pieces of the tcgen05 / TMEM / TMA convolution kernels (conv_tc.cu: bf16 sampling path, conv_tc32.cu: TF32 forward /
// input gradient, conv_wgrad_tc32.cu: TF32 weight gradient): launch parameters, PTX wrappers (mbarrier, TMA, UMMA, TMEM loads),
// the halo-tile geometry and the host-side tensor-map builders.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace dd {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;        // 16 KB
constexpr int TC_THREADS = 224;      // warps: A producer, MMA, 4 x epilogue, B producer
// Pipeline variants <STAGES, BROWS>: BROWS = rows of the weight slot (>= bn).
//   <3,128>: 96 KB  -> 2 CTAs/SM, for grids that fill the GPU (epilogue of one CTA overlaps the other's loop)
//   <6,128>: 192 KB -> 1 CTA/SM, grids of at most one wave: twice the loads in flight per CTA
//   <8, 64>: 192 KB -> 1 CTA/SM, low-resolution layers run with bn = 64 (twice the CTAs) and 8 stages
constexpr int tc_smem_bytes(int stages, int brows, int kch) { return stages * kch * (TC_A_BYTES + brows * TC_BK * 2) + 1024 /*align*/ + 1536 /*barriers + bias + weight row sums*/; }
constexpr int TC_TMEM_COLS = 128;

struct TcParams {
    CUtensorMap tmA0, tmA1, tmB;
    CUtensorMap tmB3;            // fp32 weights as (c, row, tap): one box = all nine taps of a 32-channel chunk (conv_tc32_halo_kernel)
    CUtensorMap tmH0, tmH1;      // halo boxes (64 ch, tw+2, th+2, 1, 1) of source 0 / 1 (halo kernel only)
    int8_t tap_dw[16], tap_dh[16], tap_plane[16];
    int ntaps;              // taps per phase
    int chunks0, chunks1;   // 64-channel chunks of source 0 / 1
    int tw, th, tn, tiles_w, tiles_h;
    int B, H, W;            // GEMM pixel grid
    int Cout, cout_valid, bn, rows_per_phase;
    int out_mul;            // 1, or 2 for the sub-pixel phases of the transposed conv
    int in_mul;             // 1, or 2 when a stride-2 conv reads its input through a stride-2 tensor map (DD_TC_STRIDED_IN)
    int out_nchw_f32;
    int G, cpg_mask, cpg_shift;
    int tw_sh, th_sh;           // log2(tw), log2(th): tile geometry is all powers of two
    int rows_valid;             // tw*th*tn (< 128 when one image has fewer than 128 pixels and tn is forced to 1)
    int pair_nt;                // CTA-pair halo kernel: 128-column accumulator blocks per CTA (N of the pair's MMA = bn * pair_nt)
    int mma_parts;              // partial accumulators left by the MMA issuers (1, or 2: the epilogue folds them, tc_fold_partials)
    int w_per_sample;           // weights are (B, rows, K): every image multiplies its own matrix (fused attention output)
    int splits, kb_per_split;   // split-K over the (tap, chunk) loop; partial sums meet in splitk_ws
    void* out;
    const float* bias;
    const __nv_bfloat16* residual;
    float* gn_stats;
    float* splitk_ws;           // (tiles, bn/4, 128, 4) fp32, all zero between launches (self-cleaning)
    int32_t* splitk_cnt;        // per-tile arrival counters, all zero between launches
    float* out2;                // dd_conv_tc32: optional second output mish(y) (the next conv's activated input)
    const float* mgrad;         // dd_conv_tc32: optional z, the result is multiplied by mish'(z) (input gradient through a pre-activation)
    long long* dbg;             // optional per-CTA timeline (8 clock64 stamps per CTA), NULL in production
    // ---- fused GroupNorm + Mish epilogue (dd_conv_tc_gn): y = mish(gn(acc + bias)) [+ tbias[row(n), c]] [+ residual] ----
    int gn_fuse;                // the epilogue normalises its own tile; statistics of one image meet inside the thread-block cluster
    int gn_cluster;             // CTAs (consecutive pixel tiles = one image) per cluster: 1 (a tile holds whole images), 2, 4 or 8
    float gn_eps, gn_inv_n;     // inv_n = 1 / (H * W * channels per group)
    const float* gn_gamma;
    const float* gn_beta;
    const float* tbias;         // time-embedding bias rows (rows, tb_stride) fp32, already offset to this layer's first column
    int tb_stride;
    const int32_t* trow;        // row index per sample (stride trow_stride; 0 = one shared step counter); NULL: row = n
    int trow_stride;
    float* ln_part;             // optional (pixels, Cout / bn, 2) fp32: per-pixel {sum, sum of squares} of the written tile row
    // ---- channel LayerNorm folded into a 1x1 convolution (dd_conv_tc_ln): y = inv_p * (acc - mean_p * wsum[c]) + bias[c] ----
    const float* ln_in;         // (pixels, ln_in_parts, 2) fp32 {sum, sum of squares} over the INPUT channels of every pixel
    int ln_in_parts;
    const float* ln_wsum;       // (Cout) row sums of the packed bf16 weights (the gain already folded into them)
    float ln_eps, ln_inv_c;     // eps is added to the standard deviation (blocks.py:57-60); inv_c = 1 / input channels
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread up to a system-dependent time limit: useless for polling two barriers)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();      // ~2 s: surfaces as a launch failure on the host
    }
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// One elected lane of a converged warp.  The producer / MMA warps run their loops with all 32 lanes so every
// loop variable stays warp-uniform (uniform registers feed UTMALDG / UTCHMMA directly); wrapping the loops in
// `if (lane == 0)` instead costs ~25 R2UR/ELECT/vote instructions per TMA issue (profiles/README.md).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: 8-row atoms of 1024 B (SBO), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pair (cta_group::2): the exact instruction forms are those of CUTLASS' cute/arch/{copy_sm100_tma,mma_sm100_umma}.hpp
//      and cutlass/arch/barrier.h (the pair's shared-memory windows differ in bit 24 of the shared::cluster address)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {          // same offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t caddr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(caddr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's mbarrier (same offset, peer bit cleared)
__device__ __forceinline__ void tma_load_5d_2sm(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
// arrive on the same barrier offset in both CTAs of the pair when the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// two 16-column loads in flight, one wait -- a single asm statement, so no use of the results can be scheduled before the wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t t0, uint32_t t1, uint32_t (&r)[16], uint32_t (&s)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]),
          "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15])
        : "r"(t0), "r"(t1) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#ifndef DD_TC_TIMELINE
#define DD_TC_TIMELINE 0          // build with -DDD_TC_TIMELINE=1 to record per-CTA clock64 stamps (scripts/timeline.py)
#endif
__device__ __forceinline__ void tstamp(const TcParams& p, int slot) {
    if (DD_TC_TIMELINE && p.dbg) p.dbg[((int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z))) * 32 + slot] = clock64();
}
__device__ __forceinline__ void tstore(const TcParams& p, int slot, long long v) {
    if (DD_TC_TIMELINE && p.dbg) p.dbg[((int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z))) * 32 + slot] = v;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

// halo tile of the 3x3 stride-1 kernels: 16 rows x 8 columns of one image, (18 x 10)-pixel input box per channel chunk
constexpr int HALO_TH = 16, HALO_TW = 8;
constexpr int HALO_ROWS = (HALO_TH + 2) * (HALO_TW + 2);            // 180 pixels
constexpr int HALO_TX = HALO_ROWS * TC_BK * 2;                      // 23040 bytes per TMA box
constexpr int HALO_SLOT = (HALO_TX + 1023) / 1024 * 1024;           // 23552
constexpr int HALO_NH = 2, HALO_NB = 4;
constexpr int HALO_B_BYTES = 128 * TC_BK * 2;
constexpr int HALO_SMEM = HALO_NH * HALO_SLOT + HALO_NB * HALO_B_BYTES + 1024 + 1024;

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// stride 2: the map traverses every other pixel of a (2W x 2H) image (TMA elementStrides), so a box still lands as tw x th rows
static int make_act_map(CUtensorMap* tm, const void* ptr, int C, int pitch, int W, int H, int N, int P, int tw, int th, int tn,
                        CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B, bool f32 = false, int stride = 1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DD_ERR_CUDA; }
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)P};
    const cuuint64_t es = f32 ? 4 : 2;
    cuuint64_t strides[4] = {(cuuint64_t)pitch * es, (cuuint64_t)W * pitch * es, (cuuint64_t)H * W * pitch * es,
                             (cuuint64_t)N * H * W * pitch * es};
    cuuint32_t box[5] = {f32 ? 32u : 64u, (cuuint32_t)(tw * stride), (cuuint32_t)(th * stride), (cuuint32_t)tn, 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation C=%d W=%d H=%d N=%d P=%d box %d,%d,%d) failed: %d", C, W, H, N, P, tw, th, tn, (int)r); return DD_ERR_CUDA; }
    return DD_OK;
}

static int make_w_map(CUtensorMap* tm, const void* ptr, int K, int rows, int bn, int batch, bool f32 = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DD_ERR_CUDA; }
    const cuuint64_t es = f32 ? 4 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)K * es, (cuuint64_t)K * es * rows};
    cuuint32_t box[3] = {f32 ? 32u : 64u, (cuuint32_t)bn, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, batch > 0 ? 3 : 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights K=%d rows=%d bn=%d) failed: %d", K, rows, bn, (int)r); return DD_ERR_CUDA; }
    return DD_OK;
}

// (c, row, tap) view of the packed fp32 [row][tap*Cin + c] 3x3 weights: box (32, bn, 9) = all nine taps of a 32-channel chunk
static int make_w_map_taps(CUtensorMap* tm, const void* ptr, int Cin, int rows, int bn, bool f32 = true) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DD_ERR_CUDA; }
    const cuuint64_t es = f32 ? 4 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)rows, 9};
    cuuint64_t strides[2] = {(cuuint64_t)9 * Cin * es, (cuuint64_t)Cin * es};
    cuuint32_t box[3] = {f32 ? 32u : 64u, (cuuint32_t)bn, f32 ? 9u : 3u};          // fp32: all nine taps of a 32-channel chunk
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weight taps Cin=%d rows=%d bn=%d) failed: %d", Cin, rows, bn, (int)r); return DD_ERR_CUDA; }
    return DD_OK;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// launch with a (2,1,1) thread-block cluster + programmatic dependent launch
template <typename... KArgs, typename... Args>
static void launch_pair_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


// launch with an (n,1,1) thread-block cluster (n = 1: plain launch) + programmatic dependent launch
template <typename... KArgs, typename... Args>
static void launch_cluster_pdl(void (*kernel)(KArgs...), int cluster_x, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// persistent halo convolution with the fused GroupNorm epilogue (conv_tc_persist.cu)
bool halo_persist_ok(int kind, int H, int W, int Cout, int G);
int launch_halo_persist(const TcParams& p, cudaStream_t st);
int halo_persist_col_split();          // ln_part blocks per 128-channel tile written by the persistent halo kernel
bool gemm_persist_ok(int kind, int B, int H, int W, int C1, int C2, int Cout, int G, bool stats, bool wps);
int launch_gemm_persist(TcParams& p, const void* x, int x_pitch, int C1, cudaStream_t st);

extern long long* g_tc_dbg;          // optional in-kernel timeline buffer (dd_debug_set_timeline), defined in conv_tc.cu

}  // namespace dd
