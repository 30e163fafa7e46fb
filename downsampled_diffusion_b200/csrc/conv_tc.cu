// tcgen05 / TMEM / TMA implicit-GEMM convolution for the U-Net (sm_100a only).
//
//   D[pixel, co] = sum_{tap} sum_{c} X[pixel shifted by tap, c] * Wp[co, tap*Cin + c]
//
// GEMM view: M = 128 output pixels per CTA (a th x tw rectangle of tn images), N = bn output
// channels, K = 64 input channels of one filter tap per pipeline stage.
//   * A operand: one TMA box load per (tap, 64-channel chunk) straight from the NHWC activation,
//     coordinates shifted by the tap offset; out-of-bounds rows/columns are zero-filled by TMA,
//     which *is* the conv's zero padding.  The box lands in shared memory as 128 rows x 128 B with
//     the 128B swizzle = the canonical K-major UMMA layout, so no im2col buffer ever exists.
//   * B operand: TMA box (64 K x bn rows) of the packed bf16 weights, same swizzle.
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=bn, K=16) accumulates fp32 in TMEM.
//   * warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue
//     (tcgen05.ld -> +bias -> GroupNorm sum/sumsq atomics -> +residual -> bf16 NHWC / fp32 NCHW store).
// Two CTAs are resident per SM (3 stages x 32 KB each) so one CTA's epilogue overlaps the
// other's main loop.
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
constexpr int tc_smem_bytes(int stages, int brows, int kch) { return stages * kch * (TC_A_BYTES + brows * TC_BK * 2) + 1024 /*align*/ + 1024 /*barriers + bias*/; }
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#ifndef DD_TC_TIMELINE
#define DD_TC_TIMELINE 0          // build with -DDD_TC_TIMELINE=1 to record per-CTA clock64 stamps (scripts/timeline.py)
#endif
__device__ __forceinline__ void tstamp(const TcParams& p, int slot) {
    if (DD_TC_TIMELINE && p.dbg) p.dbg[((int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z))) * 16 + slot] = clock64();
}
__device__ __forceinline__ void tstore(const TcParams& p, int slot, long long v) {
    if (DD_TC_TIMELINE && p.dbg) p.dbg[((int64_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z))) * 16 + slot] = v;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// ---- epilogue shared by every pipeline variant: 4 warps, TMEM lane quadrant = warp % 4 ----
__device__ __forceinline__ void tc_epilogue(const TcParams& p, uint32_t tmem_base, uint32_t tmem_full_bar, float* s_bias,
                                            int m_tile, int n_tile, int phase, int w0, int h0, int n0, int warp, int lane) {
    // ===== epilogue: 4 warps, TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ww = r % p.tw, hh = (r / p.tw) % p.th, nl = r / (p.tw * p.th);
    const int n = n0 + nl;
    const bool valid = n < p.B && r < p.rows_valid;
    const int mul = p.out_mul;
    const int Ho = p.H * mul, Wo = p.W * mul;
    const int oh = (h0 + hh) * mul + (phase >> 1), ow = (w0 + ww) * mul + (phase & 1);
    const int64_t pix = ((int64_t)n * Ho + oh) * Wo + ow;
    const int cbase = n_tile * p.bn;
    const int seg = min(32, p.tw * p.th);       // lanes of this warp that share a sample
    const int et = threadIdx.x - 64;            // 0..127
    if (et < p.bn) s_bias[et] = (p.bias && cbase + et < p.Cout) ? p.bias[cbase + et] : 0.f;
    const bool use_res = p.residual != nullptr && valid && !p.out_nchw_f32;
    const uint4* res_ptr = reinterpret_cast<const uint4*>(p.residual + (use_res ? pix * p.Cout + cbase : 0));
    uint4 res_cur[2], res_nxt[2];
    if (use_res) { res_cur[0] = res_ptr[0]; res_cur[1] = res_ptr[1]; }
    epi_bar();
    mbar_wait(tmem_full_bar, 0);
    if (et == 0) tstamp(p, 5);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);

    const bool finisher = true;
    if (finisher) {
        float gs = 0.f, gq = 0.f;
        uint32_t acc[16], nxt[16];
        tmem_ld16_issue(trow, acc);
        tmem_ld_wait();
        for (int ch = 0; ch < p.bn; ch += 16) {
            const bool more = ch + 16 < p.bn;
            float v[16];
            if (more) tmem_ld16_issue(trow + (uint32_t)(ch + 16), nxt);      // overlaps the math below
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
            if (use_res && more) { res_nxt[0] = res_ptr[(ch >> 3) + 2]; res_nxt[1] = res_ptr[(ch >> 3) + 3]; }
            const int c0 = cbase + ch;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += s_bias[ch + j];
            if (p.gn_stats) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float s8 = 0.f, q8 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float t = v[hf * 8 + j]; s8 += t; q8 += t * t; }
                    if (valid) { gs += s8; gq += q8; }
                    const int c_end = c0 + hf * 8 + 8;
                    if ((c_end & p.cpg_mask) == 0) {           // warp-uniform
                        float a = gs, b = gq;
                        for (int o = 1; o < seg; o <<= 1) {
                            a += __shfl_xor_sync(0xffffffffu, a, o);
                            b += __shfl_xor_sync(0xffffffffu, b, o);
                        }
                        if (valid && (lane & (seg - 1)) == 0) {
                            const int g = (c_end >> p.cpg_shift) - 1;
                            float* st = p.gn_stats + ((int64_t)n * p.G + g) * 2;
                            atomicAdd(st, a);
                            atomicAdd(st + 1, b);
                        }
                        gs = 0.f; gq = 0.f;
                    }
                }
            }
            if (valid) {
                if (p.out_nchw_f32) {
                    float* o = reinterpret_cast<float*>(p.out);
                    const int64_t hw = (int64_t)Ho * Wo;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.cout_valid) o[((int64_t)n * p.cout_valid + c0 + j) * hw + (int64_t)oh * Wo + ow] = v[j];
                } else if (c0 < p.Cout) {
                    if (use_res) {
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&res_cur[hf]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = __bfloat1622float2(rh[j]);
                                v[hf * 8 + 2 * j] += f.x; v[hf * 8 + 2 * j + 1] += f.y;
                            }
                        }
                    }
                    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + c0);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint4 ov;
                        __nv_bfloat162* oh2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                        for (int j = 0; j < 4; ++j) oh2[j] = __floats2bfloat162_rn(v[hf * 8 + 2 * j], v[hf * 8 + 2 * j + 1]);
                        op[hf] = ov;
                    }
                }
            }
            if (more) {
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = nxt[j];
            }
            if (use_res && more) { res_cur[0] = res_nxt[0]; res_cur[1] = res_nxt[1]; }
        }
    }
    if (et == 0) tstamp(p, 6);
}


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

// ---- staged epilogue (bf16 NHWC output, bn >= 32) -------------------------------------------------
// Phase A: each of the 128 epilogue threads drains its accumulator row from TMEM (32-column loads, two in
//          flight), adds the bias, keeps per-8-channel {sum, sumsq} partials in registers and writes the bf16
//          row into a padded shared-memory tile (the pipeline stages are free once the accumulator is complete).
// Phase B: GroupNorm partials are reduced through shared memory (16 lanes per (sample, group), one atomic
//          pair each), and the tile is written out with fully coalesced 16-byte stores (+ coalesced residual
//          reads).  The first version did both per row with warp shuffles and 16-byte scattered stores and took
//          as long as the main loop (profiles/README.md).
__device__ __forceinline__ void tc_epilogue_staged(const TcParams& p, uint32_t tmem_base, uint32_t tmem_full_bar, float* s_bias,
                                                   uint8_t* stage, int n_tile, int phase, int w0, int h0, int n0, int warp,
                                                   int lane) {
    const int q = warp & 3, r = q * 32 + lane, et = threadIdx.x - 64;
    const int bn = p.bn, cbase = n_tile * bn;
    const int rps = p.tw * p.th;                       // rows of this tile that belong to one image
    const int pitch = bn * 2 + 16;                     // bytes; +16 keeps 16-byte row accesses conflict free
    if (et < bn) s_bias[et] = (p.bias && cbase + et < p.Cout) ? p.bias[cbase + et] : 0.f;
    const bool valid = (n0 + (r >> (p.tw_sh + p.th_sh))) < p.B && r < p.rows_valid;
    const bool do_stats = p.gn_stats != nullptr;
    epi_bar();
    mbar_wait(tmem_full_bar, 0);
    if (et == 0) tstamp(p, 5);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float s8[16], q8[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { s8[k] = 0.f; q8[k] = 0.f; }
    uint32_t a0[32], a1[32];
    auto process = [&](const int c32, uint32_t (&acc)[32]) {
        uint8_t* dst = stage + r * pitch + c32 * 64;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
            float v[8];
            float sa = 0.f, qa = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = __uint_as_float(acc[g8 * 8 + j]) + s_bias[c32 * 32 + g8 * 8 + j];
                sa += v[j]; qa = fmaf(v[j], v[j], qa);
            }
            if (valid) { s8[c32 * 4 + g8] = sa; q8[c32 * 4 + g8] = qa; }
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            *reinterpret_cast<uint4*>(dst + g8 * 16) = pk;
        }
    };
    tmem_ld32_issue(trow, a0);
    tmem_ld_wait();
    if (bn > 32) tmem_ld32_issue(trow + 32u, a1);
    process(0, a0);
    if (bn > 32) {
        tmem_ld_wait();
        if (bn > 64) tmem_ld32_issue(trow + 64u, a0);
        process(1, a1);
        if (bn > 64) {
            tmem_ld_wait();
            tmem_ld32_issue(trow + 96u, a1);
            process(2, a0);
            tmem_ld_wait();
            process(3, a1);
        }
    }
    if (et == 0) tstamp(p, 10);
    float* s_part = reinterpret_cast<float*>(stage + TC_BM * pitch);      // [(sub*2 + {sum,sq})][128 rows]
    if (do_stats) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k * 8 < bn) { s_part[(2 * k) * TC_BM + r] = s8[k]; s_part[(2 * k + 1) * TC_BM + r] = q8[k]; }
    }
    epi_bar();
    if (do_stats) {
        // (sample, group) outputs: 16 lanes each, rows of the sample split across the lanes; all sizes are
        // powers of two, so only shifts and masks appear below
        const int cpg_sh = p.cpg_shift, spg = 1 << (cpg_sh - 3);           // 8-channel partials per group
        const int ng_sh = (31 - __clz(bn)) - cpg_sh;                        // log2(groups in this tile)
        const int rps_sh2 = p.tw_sh + p.th_sh;
        const int nout = (1 << ng_sh) * (TC_BM >> rps_sh2);
        const int hw = et >> 4, l16 = et & 15;
        const int gbase = cbase >> cpg_sh;
        for (int o0 = 0; o0 < nout; o0 += 8) {
            const int o = o0 + hw;
            const bool act = o < nout;
            const int gl = o & ((1 << ng_sh) - 1), sl = o >> ng_sh;
            float sa0 = 0.f, qa0 = 0.f, sa1 = 0.f, qa1 = 0.f;
            if (act) {
                constexpr int rs = 16;                                      // row stride between a lane's rows
                const float* ps = s_part + (2 * (gl * spg)) * TC_BM + (sl << rps_sh2) + l16;
                const int cnt = rps >> 4;                                   // rows per lane (0 when rps < 16)
                for (int k = 0; k < spg; ++k, ps += 2 * TC_BM) {
                    if (cnt == 0) { if (l16 < rps) { sa0 += ps[0]; qa0 += ps[TC_BM]; } continue; }
                    int i = 0;
                    for (; i + 1 < cnt; i += 2) {
                        sa0 += ps[rs * i]; qa0 += ps[TC_BM + rs * i];
                        sa1 += ps[rs * i + rs]; qa1 += ps[TC_BM + rs * i + rs];
                    }
                    if (i < cnt) { sa0 += ps[rs * i]; qa0 += ps[TC_BM + rs * i]; }
                }
            }
            float sa = sa0 + sa1, qa = qa0 + qa1;
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, d);
                qa += __shfl_xor_sync(0xffffffffu, qa, d);
            }
            if (act && l16 == 0 && n0 + sl < p.B) {
                float* st = p.gn_stats + ((int64_t)(n0 + sl) * p.G + (gbase + gl)) * 2;
                atomicAdd(st, sa);
                atomicAdd(st + 1, qa);
            }
        }
    }
    if (et == 0) tstamp(p, 11);
    // coalesced write-out: 16 bytes per thread, consecutive threads walk along a row (shifts only: every
    // tile dimension is a power of two)
    const int ppr_sh = 31 - __clz(bn >> 3);        // log2(16-byte pieces per row)
    const int rps_sh = p.tw_sh + p.th_sh;
    const int pc = et & ((1 << ppr_sh) - 1);
    const int row_step = TC_BM >> ppr_sh;
    const int mul = p.out_mul, Ho = p.H * mul, Wo = p.W * mul;
    const int py = phase >> 1, px = phase & 1;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    const __nv_bfloat16* resp = p.residual;
    const int rows_valid = p.rows_valid, Bn = p.B, Cout = p.Cout;
    // four rows per trip: all shared/global loads first, then the stores (independent -> latencies overlap)
    for (int row0 = et >> ppr_sh; row0 < TC_BM; row0 += 4 * row_step) {
        uint4 v[4], rv[4];
        int64_t off[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = row0 + u * row_step;
            const int n = n0 + (row >> rps_sh);
            ok[u] = row < TC_BM && row < rows_valid && n < Bn;
            const int ww = row & (p.tw - 1), hh = (row >> p.tw_sh) & (p.th - 1);
            const int oh = (h0 + hh) * mul + py, ow = (w0 + ww) * mul + px;
            off[u] = (((int64_t)n * Ho + oh) * Wo + ow) * Cout + cbase + pc * 8;
            if (ok[u]) {
                v[u] = *reinterpret_cast<const uint4*>(stage + row * pitch + pc * 16);
                if (resp) rv[u] = *reinterpret_cast<const uint4*>(resp + off[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            if (resp) {
                __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&v[u]);
                const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&rv[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 x = __bfloat1622float2(a[j]), y = __bfloat1622float2(b[j]);
                    a[j] = __floats2bfloat162_rn(x.x + y.x, x.y + y.y);
                }
            }
            *reinterpret_cast<uint4*>(outp + off[u]) = v[u];
        }
    }
    if (et == 0) tstamp(p, 6);
}

// ---- split-K partial epilogue ------------------------------------------------------------------------
// The CTA's fp32 accumulator goes, unreduced and without bias, to ws[split][pixel][Cout].  dd_gn_mish_sum adds the
// splits (+ bias) while it computes the GroupNorm statistics: no atomics, no counters, deterministic, and the
// low-resolution layers (8 - 32 tiles per launch) spread their operand streams over 4x as many SMs.
__device__ __forceinline__ void tc_epilogue_partial(const TcParams& p, uint32_t tmem_base, uint32_t tmem_full_bar, int n_tile, int split,
                                                    int w0, int h0, int n0, int warp, int lane) {
    const int q = warp & 3, r = q * 32 + lane;
    const int ww = r & (p.tw - 1), hh = (r >> p.tw_sh) & (p.th - 1), n = n0 + (r >> (p.tw_sh + p.th_sh));
    const bool valid = n < p.B && r < p.rows_valid;
    const int64_t pix = ((int64_t)n * p.H + (h0 + hh)) * p.W + (w0 + ww);
    float* dst = p.splitk_ws + (((int64_t)split * p.B * p.H * p.W + pix) * p.Cout + n_tile * p.bn);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t a0[32], a1[32];
    tmem_ld32_issue(trow, a0);
    for (int c = 0; c < p.bn; c += 64) {
        tmem_ld_wait();
        if (c + 32 < p.bn) tmem_ld32_issue(trow + (uint32_t)(c + 32), a1);
        if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                reinterpret_cast<uint4*>(dst + c)[j] = make_uint4(a0[4 * j], a0[4 * j + 1], a0[4 * j + 2], a0[4 * j + 3]);
        }
        if (c + 32 < p.bn) {
            tmem_ld_wait();
            if (c + 64 < p.bn) tmem_ld32_issue(trow + (uint32_t)(c + 64), a0);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    reinterpret_cast<uint4*>(dst + c + 32)[j] = make_uint4(a1[4 * j], a1[4 * j + 1], a1[4 * j + 2], a1[4 * j + 3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Generic pipeline.  <STAGES, BROWS, KCH>: KCH 64-channel chunks per stage (KCH = 2 halves the per-k-block
// barrier / issue overhead, which -- not bandwidth -- bounds small tiles: one warp needs ~400 clk to issue a
// stage, see profiles/README.md).  Warp roles: 0 = A-operand TMA producer, 6 = B-operand TMA producer,
// 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
// ---------------------------------------------------------------------------------------------
template <int TC_STAGES, int BROWS, int KCH, int EPI>      // EPI: 0 staged bf16, 1 legacy (fp32 NCHW / narrow), 2 split-K partial
__global__ void __launch_bounds__(TC_THREADS, (TC_STAGES * KCH <= 3 ? 2 : 1)) conv_tc_kernel(const __grid_constant__ TcParams p) {
    constexpr int CH = 64;                               // channels per 128-byte operand row (bf16)
    constexpr int B_SLOT = BROWS * TC_BK * 2;
    constexpr int TC_STAGE_BYTES = KCH * (TC_A_BYTES + B_SLOT);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;     // full[S], empty[S], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * TC_STAGES);
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * TC_STAGES + 1);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) tstamp(p, 0);
    const int m_tile = blockIdx.x, n_tile = blockIdx.y;
    const int phase = EPI == 2 ? 0 : blockIdx.z, split = EPI == 2 ? blockIdx.z : 0;     // split-K: blockIdx.z walks the K ranges
    const int w0 = (m_tile % p.tiles_w) * p.tw;
    const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
    const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.tn;
    const int cpt = p.chunks0 + p.chunks1;                  // 64-channel chunks per tap (multiple of KCH)
    const int num_st = p.kb_per_split / KCH;                // pipeline stages this CTA consumes
    const int kb0 = split * p.kb_per_split;                 // first (tap, chunk) block of this CTA's K range
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));   // 128 floats (barriers use < 160 B)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (threadIdx.x == 0) tstamp(p, 1);
    pdl_sync();      // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
    if (threadIdx.x == 0) tstamp(p, 2);

    if (warp == 0) {
        // ===== A-operand producer: whole warp runs the (uniform) loop, one elected lane issues =====
        const uint32_t tx = (uint32_t)KCH * p.rows_valid * TC_BK * 2;
        const int chunks0 = p.chunks0;
        int rem = kb0 % cpt, ti = phase * p.ntaps + kb0 / cpt;
        const int wi0 = w0 * p.in_mul, hi0 = h0 * p.in_mul;      // tile origin in input pixels
        int cx = wi0 + p.tap_dw[ti], cy = hi0 + p.tap_dh[ti], cp = p.tap_plane[ti];
        uint32_t sA = base;
        int st = 0, round = 0;
        for (int i = 0; i < num_st; ++i) {
            const uint32_t fb = full_bar(st);
            if (round > 0) mbar_wait(empty_bar(st), (round - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(fb, tx);
#pragma unroll
                for (int j = 0; j < KCH; ++j) {
                    const int c = rem + j;
                    if (c < chunks0) tma_load_5d(&p.tmA0, fb, sA + j * TC_A_BYTES, c * CH, cx, cy, n0, cp);
                    else tma_load_5d(&p.tmA1, fb, sA + j * TC_A_BYTES, (c - chunks0) * CH, cx, cy, n0, cp);
                }
            }
            __syncwarp();
            sA += TC_STAGE_BYTES;
            if (++st == TC_STAGES) { st = 0; ++round; sA = base; }
            rem += KCH;
            if (rem == cpt) {
                rem = 0; ++ti;
                if (i + 1 < num_st) { cx = wi0 + p.tap_dw[ti]; cy = hi0 + p.tap_dh[ti]; cp = p.tap_plane[ti]; }
            }
        }
    } else if (warp == 6) {
        // ===== B-operand (weights) producer =====
        const uint32_t tx = (uint32_t)KCH * p.bn * TC_BK * 2;
        const int wps = p.w_per_sample;
        const int brow = wps ? n_tile * p.bn : phase * p.rows_per_phase + n_tile * p.bn;
        int kcoord = kb0 * CH;
        uint32_t sB = base + KCH * TC_A_BYTES;
        int st = 0, round = 0;
        for (int i = 0; i < num_st; ++i) {
            const uint32_t fb = full_bar(st);
            if (round > 0) mbar_wait(empty_bar(st), (round - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(fb, tx);
#pragma unroll
                for (int j = 0; j < KCH; ++j) {
                    if (wps) tma_load_3d(&p.tmB, fb, sB + j * B_SLOT, kcoord + CH * j, brow, n0);
                    else tma_load_2d(&p.tmB, fb, sB + j * B_SLOT, kcoord + CH * j, brow);
                }
            }
            __syncwarp();
            kcoord += CH * KCH;
            sB += TC_STAGE_BYTES;
            if (++st == TC_STAGES) { st = 0; ++round; sB = base + KCH * TC_A_BYTES; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: whole warp loops, one elected lane issues tcgen05.mma / commit =====
        // instruction descriptor: D=f32, A=B=bf16, both K-major, N=bn, M=128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        uint32_t sA = base;
        int st = 0;
        uint32_t par = 0;
        for (int i = 0; i < num_st; ++i) {
            mbar_wait(full_bar(st), par);
            if (i == 0 && lane == 0) tstamp(p, 3);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < KCH; ++j) {
                    const uint64_t ad = umma_desc(sA + j * TC_A_BYTES);
                    const uint64_t bd = umma_desc(sA + KCH * TC_A_BYTES + j * B_SLOT);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)
                        umma_f16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i | j | k) ? 1u : 0u);
                }
                umma_commit(empty_bar(st));          // frees this smem stage when the MMAs retire
            }
            __syncwarp();
            sA += TC_STAGE_BYTES;
            if (++st == TC_STAGES) { st = 0; par ^= 1u; sA = base; }
        }
        if (lane == 0) tstamp(p, 4);
        if (elect_one()) umma_commit(tmem_full_bar);             // accumulator complete
        __syncwarp();
    } else {
        if constexpr (EPI == 1)         // fp32 NCHW output / tiles narrower than 32 channels (the final 1x1 conv)
            tc_epilogue(p, tmem_base, tmem_full_bar, s_bias, m_tile, n_tile, phase, w0, h0, n0, warp, lane);
        else if constexpr (EPI == 2)
            tc_epilogue_partial(p, tmem_base, tmem_full_bar, n_tile, split, w0, h0, n0, warp, lane);
        else
            tc_epilogue_staged(p, tmem_base, tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)), n_tile, phase, w0, h0, n0,
                               warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
    }
}

// =============================================================================================
// Halo variant for 3x3 stride-1 convolutions on maps of at least 16x8 pixels.
// The CTA tile is 16 rows x 8 columns of one image.  For every 64-channel chunk the (18 x 10)-pixel
// input halo is loaded ONCE (one TMA box, 23 KB) and all nine filter taps read it in place: the tap
// (r, s) operand is the same shared-memory tile addressed (r*10 + s) rows further on, with an 8-row
// group stride of 10 rows (SBO = 1280 B).  The 128-byte swizzle is a function of the absolute
// shared-memory address on both the TMA and the UMMA side (measured: profiles/README.md), so the
// shifted descriptors see exactly the bytes TMA wrote.  A-operand traffic drops from 9 x 16 KB to
// 23 KB per chunk -- these layers are bound by the L2 -> SM path, not by the tensor pipe.
// =============================================================================================
constexpr int HALO_TH = 16, HALO_TW = 8;
constexpr int HALO_ROWS = (HALO_TH + 2) * (HALO_TW + 2);            // 180 pixels
constexpr int HALO_TX = HALO_ROWS * TC_BK * 2;                      // 23040 bytes per TMA box
constexpr int HALO_SLOT = (HALO_TX + 1023) / 1024 * 1024;           // 23552
constexpr int HALO_NH = 2, HALO_NB = 4;
constexpr int HALO_B_BYTES = 128 * TC_BK * 2;
constexpr int HALO_SMEM = HALO_NH * HALO_SLOT + HALO_NB * HALO_B_BYTES + 1024 + 1024;
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_halo_kernel(const __grid_constant__ TcParams p) {
    constexpr int NT = 1;                               // output tiles per CTA (a two-tile form was measured and dropped, profiles/README.md)
    constexpr int NSTG = HALO_NH;                       // halo pipeline depth
    constexpr uint32_t DY_BYTES = (HALO_TW + 2) * 128u;                       // smem bytes between filter rows
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bbase = base + HALO_NH * HALO_SLOT;
    const uint32_t bars = bbase + HALO_NB * HALO_B_BYTES;
    auto hfull = [&](int s) { return bars + 8u * s; };
    auto hempty = [&](int s) { return bars + 8u * (HALO_NH + s); };
    auto bfull = [&](int s) { return bars + 8u * (2 * HALO_NH + s); };
    auto bempty = [&](int s) { return bars + 8u * (2 * HALO_NH + HALO_NB + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * HALO_NH + 2 * HALO_NB);
    const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 512u - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) tstamp(p, 0);
    const int n_tile = blockIdx.y;
    int w0[NT], h0[NT], n0[NT];
#pragma unroll
    for (int s = 0; s < NT; ++s) {
        const int m_tile = blockIdx.x * NT + s;
        w0[s] = (m_tile % p.tiles_w) * p.tw;
        h0[s] = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
        n0[s] = m_tile / (p.tiles_w * p.tiles_h);
    }
    const int nchunks = p.chunks0 + p.chunks1;
    const int cin = nchunks * 64;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmH0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        for (int s = 0; s < HALO_NH; ++s) { mbar_init(hfull(s), 1); mbar_init(hempty(s), 1); }
        for (int s = 0; s < HALO_NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(TC_TMEM_COLS * NT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (threadIdx.x == 0) tstamp(p, 1);
    pdl_sync();
    if (threadIdx.x == 0) tstamp(p, 2);

    if (warp == 0) {
        const uint32_t b_tx = (uint32_t)p.bn * TC_BK * 2;
        const int brow = n_tile * p.bn, chunks0 = p.chunks0;
        int bs = 0, bround = 0, hs = 0, hround = 0;
        uint32_t sB = bbase;
        for (int c = 0; c < nchunks; ++c) {
            if (hround > 0) mbar_wait(hempty(hs), (hround - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(hfull(hs), NT * HALO_TX);
#pragma unroll
                for (int s = 0; s < NT; ++s) {
                    const uint32_t dst = base + (hs * NT + s) * HALO_SLOT;
                    const CUtensorMap* tm = c < chunks0 ? &p.tmH0 : &p.tmH1;
                    const int cc = (c < chunks0 ? c : c - chunks0) * 64;
                    tma_load_5d(tm, hfull(hs), dst, cc, w0[s] - 1, h0[s] - 1, n0[s], 0);
                }
            }
            __syncwarp();
            if (++hs == NSTG) { hs = 0; ++hround; }
            int kcoord = c * 64;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const uint32_t fb = bfull(bs);
                if (bround > 0) mbar_wait(bempty(bs), (bround - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(fb, b_tx);
                    tma_load_2d(&p.tmB, fb, sB, kcoord, brow);
                }
                __syncwarp();
                kcoord += cin;
                sB += HALO_B_BYTES;
                if (++bs == HALO_NB) { bs = 0; ++bround; sB = bbase; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        // A: 8-row groups (one tile row of 8 pixels) are (HALO_TW + 2) halo rows apart
        const uint64_t a_hi = ((uint64_t)(((HALO_TW + 2) * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        int bs = 0, hs = 0;
        uint32_t bpar = 0, hpar = 0, acc = 0;
        uint32_t sB = bbase;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(hfull(hs), hpar);
            uint32_t rowA = base + hs * NT * HALO_SLOT;             // tap (0,0) of the first tile
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
#pragma unroll 1
                for (int sx = 0; sx < 3; ++sx) {
                    mbar_wait(bfull(bs), bpar);
                    if (acc == 0 && lane == 0) tstamp(p, 3);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bd = umma_desc(sB);
#pragma unroll
                        for (int s = 0; s < NT; ++s) {
                            const uint64_t ad = (uint64_t)(((rowA + s * HALO_SLOT + 128u * sx) & 0x3FFFFu) >> 4) | a_hi;
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; ++k)
                                umma_f16(tmem_base + (uint32_t)(s * TC_TMEM_COLS), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc,
                                         (acc | k) ? 1u : 0u);
                        }
                        umma_commit(bempty(bs));
                    }
                    __syncwarp();
                    acc = 1u;
                    sB += HALO_B_BYTES;
                    if (++bs == HALO_NB) { bs = 0; bpar ^= 1u; sB = bbase; }
                }
                rowA += DY_BYTES;
            }
            if (elect_one()) umma_commit(hempty(hs));           // halo slot(s) free once their nine taps have retired
            __syncwarp();
            if (++hs == NSTG) { hs = 0; hpar ^= 1u; }
        }
        if (lane == 0) tstamp(p, 4);
        if (elect_one()) umma_commit(tmem_full_bar);
        __syncwarp();
    } else if (warp < 6) {
#pragma unroll
        for (int s = 0; s < NT; ++s)
            tc_epilogue_staged(p, tmem_base + (uint32_t)(s * TC_TMEM_COLS), tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)), n_tile, 0,
                               w0[s], h0[s], n0[s], warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS * NT) : "memory");
    }
}

// =============================================================================================
// CTA-pair form of the halo kernel (tcgen05 cta_group::2).
// With both operands in shared memory a cta_group::1 MMA is paced by its operand reads (~64 B/clk: M128 x N128 x K16 takes
// ~128 clk, half the tensor-pipe rate -- profiles/README.md).  Two CTAs of a cluster (one SM pair) compute a 256-pixel x N tile
// together: each loads the halo of ITS 128 pixels and HALF of the weight rows, the leader issues one M = 256 MMA per k-step,
// each SM reads 128 A rows + N/2 B rows and the accumulator rows land in each CTA's own TMEM.  For N = 256 an SM reads the
// same bytes as before for twice the math.
//   * full barriers live in the leader; both CTAs' TMA loads signal them (cta_group::2 loads, peer bit of the barrier address
//     cleared), the leader's producer arms them with the pair's byte count;
//   * empty / accumulator-ready barriers exist in both CTAs; the leader's tcgen05.commit multicasts to both;
//   * TMEM is allocated with cta_group::2 by warp 1 of both CTAs; a cluster barrier brackets setup and teardown.
// =============================================================================================
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_halo2_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bbase = base + HALO_NH * HALO_SLOT;
    const uint32_t bars = bbase + HALO_NB * HALO_B_BYTES;
    auto hfull = [&](int s) { return bars + 8u * s; };
    auto hempty = [&](int s) { return bars + 8u * (HALO_NH + s); };
    auto bfull = [&](int s) { return bars + 8u * (2 * HALO_NH + s); };
    auto bempty = [&](int s) { return bars + 8u * (2 * HALO_NH + HALO_NB + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * HALO_NH + 2 * HALO_NB);
    const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 512u - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool leader = cluster_ctarank() == 0;
    const int m_tile = blockIdx.x, n_pair = blockIdx.y;           // blockIdx.x pairs (2k, 2k+1) form a cluster
    const int w0 = (m_tile % p.tiles_w) * p.tw, h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th, n0 = m_tile / (p.tiles_w * p.tiles_h);
    const int nchunks = p.chunks0 + p.chunks1;
    const int cin = nchunks * 64;
    const int N2 = p.bn * p.pair_nt;                                // N of the pair's MMA: 128 or 256
    const int half_rows = N2 >> 1;                                  // weight rows each CTA loads per tap

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmH0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        for (int s = 0; s < HALO_NH; ++s) { mbar_init(hfull(s), 1); mbar_init(hempty(s), 1); }
        for (int s = 0; s < HALO_NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if (N2 == 256) asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(256) : "memory");
        else asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                     // barriers of both CTAs are initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    pdl_sync();

    if (warp == 0) {
        // ===== producer (both CTAs): own halo, own half of the weight rows; completion goes to the leader's barriers =====
        const uint32_t b_tx = (uint32_t)half_rows * TC_BK * 2;
        const int brow = n_pair * N2 + (leader ? 0 : half_rows), chunks0 = p.chunks0;
        int bs = 0, bround = 0, hs = 0, hround = 0;
        uint32_t sB = bbase;
        for (int c = 0; c < nchunks; ++c) {
            if (hround > 0) mbar_wait(hempty(hs), (hround - 1) & 1);
            if (elect_one()) {
                if (leader) mbar_expect_tx(hfull(hs), 2 * HALO_TX);
                const CUtensorMap* tm = c < chunks0 ? &p.tmH0 : &p.tmH1;
                tma_load_5d_2sm(tm, hfull(hs), base + hs * HALO_SLOT, (c < chunks0 ? c : c - chunks0) * 64, w0 - 1, h0 - 1, n0, 0);
            }
            __syncwarp();
            if (++hs == HALO_NH) { hs = 0; ++hround; }
            int kcoord = c * 64;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                if (bround > 0) mbar_wait(bempty(bs), (bround - 1) & 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(bfull(bs), 2 * b_tx);
                    tma_load_2d_2sm(&p.tmB, bfull(bs), sB, kcoord, brow);
                }
                __syncwarp();
                kcoord += cin;
                sB += HALO_B_BYTES;
                if (++bs == HALO_NB) { bs = 0; ++bround; sB = bbase; }
            }
        }
    } else if (warp == 1 && leader) {
        // ===== MMA issuer (leader only): M = 256 over the pair, N = N2 =====
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint64_t a_hi = ((uint64_t)(((HALO_TW + 2) * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        int bs = 0, hs = 0;
        uint32_t bpar = 0, hpar = 0, acc = 0;
        uint32_t sB = bbase;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(hfull(hs), hpar);
            uint32_t rowA = base + hs * HALO_SLOT;
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
#pragma unroll 1
                for (int sx = 0; sx < 3; ++sx) {
                    mbar_wait(bfull(bs), bpar);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = (uint64_t)(((rowA + 128u * sx) & 0x3FFFFu) >> 4) | a_hi;
                        const uint64_t bd = umma_desc(sB);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                            umma_f16_2sm(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (acc | k) ? 1u : 0u);
                        umma_commit_2sm(bempty(bs));
                    }
                    __syncwarp();
                    acc = 1u;
                    sB += HALO_B_BYTES;
                    if (++bs == HALO_NB) { bs = 0; bpar ^= 1u; sB = bbase; }
                }
                rowA += (HALO_TW + 2) * 128u;
            }
            if (elect_one()) umma_commit_2sm(hempty(hs));
            __syncwarp();
            if (++hs == HALO_NH) { hs = 0; hpar ^= 1u; }
        }
        if (elect_one()) umma_commit_2sm(tmem_full_bar);
        __syncwarp();
    } else if (warp >= 2 && warp < 6) {
        // ===== epilogue (both CTAs): own 128 rows, N2 columns in 128-wide passes =====
        for (int s = 0; s < p.pair_nt; ++s)
            tc_epilogue_staged(p, tmem_base + (uint32_t)(s * TC_TMEM_COLS), tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)),
                               n_pair * p.pair_nt + s, 0, w0, h0, n0, warp, lane);
    }
    tc_fence_before();
    cluster_sync_all();                     // the peer's shared memory and TMEM stay alive until both CTAs are done
    if (warp == 1) {
        tc_fence_after();
        if (N2 == 256) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
    }
}

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

}  // namespace dd

using namespace dd;

static long long* g_tc_dbg = nullptr;
extern "C" int dd_debug_set_timeline(long long* buf) { g_tc_dbg = buf; return DD_OK; }

extern "C" int dd_zero(void* ptr, int64_t bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dd_zero: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
    return DD_OK;
}

// Split-K policy: 3x3 stride-1 convs on maps of at most 4x4 pixels whose 64-wide tiles would occupy less than a
// third of the SMs.  Each CTA streams its whole K range through one SM's L2 port (~64 B/clk), so 32 CTAs take ~7 us
// for 27 MB of operands while 116 SMs idle; S splits cut that stream S-fold.  Returns S (>= 2) or 1.
extern "C" int dd_conv_tc_splits(int kind, int B, int H, int W, int Cin, int Cout) {
    if (kind != DD_TC_CONV3x3 || H * W > 16 || Cout < 128 || Cout % 64 || Cin % 64 || getenv("DD_NO_SPLITK")) return 1;
    const int rows = B * H * W, tiles = (rows + 127) / 128 * (Cout / 64), num_kb = 9 * (Cin / 64);
    if (tiles * 3 > dd::num_sms()) return 1;
    int S = dd::num_sms() / tiles;
    if (S > 8) S = 8;
    while (S > 1 && (num_kb % S != 0 || num_kb / S < 4)) --S;
    return S;
}

extern "C" int dd_conv_tc(int kind, const void* x, int x_pitch, const void* x2, int C1, int C2, const void* wp, int w_rows,
                          const float* bias, const void* residual, void* y, int out_nchw_f32, int cout_valid,
                          float* gn_stats, int G, int B, int H, int W, int Cout, int flags,
                          float* splitk_ws, int64_t splitk_ws_floats, int32_t* splitk_cnt, int splitk_cnt_n, void* stream) {
    DD_REQUIRE(kind >= 0 && kind <= 3, "conv_tc: bad kind %d", kind);
    DD_REQUIRE(C1 > 0 && C1 % 64 == 0 && C2 >= 0 && C2 % 64 == 0, "conv_tc: channel counts (%d,%d) must be multiples of 64", C1, C2);
    DD_REQUIRE((C2 == 0) == (x2 == nullptr), "conv_tc: x2/C2 mismatch");
    DD_REQUIRE(is_pow2(H) && is_pow2(W) && B > 0, "conv_tc: H=%d, W=%d must be powers of two (use the direct kernel otherwise)", H, W);
    DD_REQUIRE(Cout > 0 && Cout % 16 == 0 && w_rows >= Cout, "conv_tc: Cout=%d must be a multiple of 16 (pad the weights)", Cout);
    DD_REQUIRE(!out_nchw_f32 || (residual == nullptr && gn_stats == nullptr && cout_valid > 0 && cout_valid <= Cout),
               "conv_tc: fp32 NCHW output takes no residual / GroupNorm statistics");
    DD_REQUIRE(!(kind == DD_TC_UPT) || gn_stats == nullptr, "conv_tc: transposed conv has no GroupNorm epilogue");

    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<3, 128, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(3, 128, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<6, 128, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(6, 128, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<8, 64, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(8, 64, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<3, 128, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(3, 128, 2));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<4, 64, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(4, 64, 2));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<3, 128, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(3, 128, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<8, 64, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(8, 64, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
        attr_done = true;
    }
    const bool wps = (flags & DD_TC_W_PER_SAMPLE) != 0;
    DD_REQUIRE(!wps || kind == DD_TC_CONV1x1, "conv_tc: per-sample weights are for 1x1 convs only");

    TcParams p;
    memset(&p, 0, sizeof(p));
    // tile geometry over the GEMM pixel grid
    static const bool halo_off = getenv("DD_NO_HALO") != nullptr;
    const bool halo = !halo_off && kind == DD_TC_CONV3x3 && H >= HALO_TH && W >= HALO_TW && Cout >= 64 && !out_nchw_f32;
    p.tw = W < 128 ? W : 128;
    p.th = (128 / p.tw) < H ? (128 / p.tw) : H;
    if (halo) { p.tw = HALO_TW; p.th = HALO_TH; }
    p.tn = wps ? 1 : 128 / (p.tw * p.th);      // per-sample weights: one image per tile (rows beyond it are ignored)
    p.rows_valid = p.tw * p.th * p.tn;
    p.tw_sh = 0; while ((1 << p.tw_sh) < p.tw) ++p.tw_sh;
    p.th_sh = 0; while ((1 << p.th_sh) < p.th) ++p.th_sh;
    p.w_per_sample = wps ? 1 : 0;
    p.tiles_w = W / p.tw; p.tiles_h = H / p.th;
    const int tiles_n = (B + p.tn - 1) / p.tn;
    p.B = B; p.H = H; p.W = W;
    p.chunks0 = C1 / 64; p.chunks1 = C2 / 64;
    p.Cout = Cout; p.cout_valid = out_nchw_f32 ? cout_valid : Cout;
    p.bn = Cout >= 128 ? 128 : Cout;
    // low-resolution layers (at most half a wave of 128-wide tiles): halve the N tile to double the CTA count
    const int tiles128 = p.tiles_w * p.tiles_h * tiles_n * ((Cout + 127) / 128);
    if (tiles128 * 2 <= num_sms() && Cout % 64 == 0 && Cout >= 128) p.bn = 64;
    if (flags & DD_TC_SPLITK) p.bn = 64;
    DD_REQUIRE(Cout % p.bn == 0 && (p.bn == 16 || p.bn == 32 || p.bn == 64 || p.bn == 128), "conv_tc: unsupported Cout=%d", Cout);
    p.out = y; p.bias = bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.gn_stats = gn_stats; p.G = G; p.out_nchw_f32 = out_nchw_f32; p.out_mul = 1;
    if (gn_stats) {
        DD_REQUIRE(G > 0 && Cout % G == 0 && is_pow2(Cout / G) && Cout / G >= 8, "conv_tc: GroupNorm needs power-of-two channels per group >= 8");
        const int cpg = Cout / G;
        p.cpg_mask = cpg - 1;
        p.cpg_shift = 0;
        while ((1 << p.cpg_shift) < cpg) ++p.cpg_shift;
    }
    const int Cin = C1 + C2;
    int planes = 1, phases = 1, K, in_stride = 1;
    if (kind == DD_TC_CONV3x3) {
        p.ntaps = 9;
        for (int t = 0; t < 9; ++t) { p.tap_dh[t] = (int8_t)(t / 3 - 1); p.tap_dw[t] = (int8_t)(t % 3 - 1); p.tap_plane[t] = 0; }
    } else if (kind == DD_TC_CONV1x1) {
        p.ntaps = 1;
    } else if (kind == DD_TC_DOWN && (flags & DD_TC_STRIDED_IN)) {
        // x is the plain (B, 2H, 2W, C) input: input pixel (2*ho + ky - 1, 2*wo + kx - 1) through a stride-2 tensor map
        p.ntaps = 9; in_stride = 2;
        for (int t = 0; t < 9; ++t) { p.tap_dh[t] = (int8_t)(t / 3 - 1); p.tap_dw[t] = (int8_t)(t % 3 - 1); p.tap_plane[t] = 0; }
    } else if (kind == DD_TC_DOWN) {
        // input row 2*ho + ky - 1: ky=0 -> odd plane, ho-1; ky=1 -> even plane, ho; ky=2 -> odd plane, ho
        p.ntaps = 9; planes = 4;
        for (int t = 0; t < 9; ++t) {
            const int ky = t / 3, kx = t % 3;
            p.tap_dh[t] = (int8_t)(ky == 0 ? -1 : 0); p.tap_dw[t] = (int8_t)(kx == 0 ? -1 : 0);
            p.tap_plane[t] = (int8_t)(((ky != 1) ? 2 : 0) + ((kx != 1) ? 1 : 0));
        }
    } else {
        // sub-pixel phase (py,px), tap (a,b): input row h + (py - a), kernel row 2a + 1 - py
        p.ntaps = 4; phases = 4; p.out_mul = 2;
        for (int ph = 0; ph < 4; ++ph)
            for (int t = 0; t < 4; ++t) {
                const int py = ph >> 1, px = ph & 1, a = t >> 1, b = t & 1;
                p.tap_dh[ph * 4 + t] = (int8_t)(py - a); p.tap_dw[ph * 4 + t] = (int8_t)(px - b); p.tap_plane[ph * 4 + t] = 0;
            }
    }
    K = p.ntaps * Cin;
    p.rows_per_phase = w_rows / phases;
    DD_REQUIRE(w_rows % phases == 0 && p.rows_per_phase >= Cout && p.rows_per_phase % p.bn == 0, "conv_tc: packed weight rows %d do not match", w_rows);

    if (x_pitch <= 0) x_pitch = C1;
    DD_REQUIRE(x_pitch >= C1 && x_pitch % 8 == 0, "conv_tc: bad channel pitch %d", x_pitch);
    p.in_mul = in_stride;
    DD_REQUIRE(in_stride == 1 || (x2 == nullptr && p.tw * in_stride <= 256 && p.th * in_stride <= 256), "conv_tc: strided input needs one source and tiles <= 128 wide");
    int rc = make_act_map(&p.tmA0, x, C1, x_pitch, W * in_stride, H * in_stride, B, planes, p.tw, p.th, p.tn,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, false, in_stride);
    if (rc) return rc;
    rc = make_act_map(&p.tmA1, x2 ? x2 : x, x2 ? C2 : C1, x2 ? C2 : x_pitch, W * in_stride, H * in_stride, B, planes, p.tw, p.th, p.tn,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, false, in_stride);
    if (rc) return rc;
    rc = make_w_map(&p.tmB, wp, K, w_rows, p.bn, wps ? B : 0);
    if (rc) return rc;
    if (halo) {
        rc = make_act_map(&p.tmH0, x, C1, x_pitch, W, H, B, 1, HALO_TW + 2, HALO_TH + 2, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
        rc = make_act_map(&p.tmH1, x2 ? x2 : x, x2 ? C2 : C1, x2 ? C2 : x_pitch, W, H, B, 1, HALO_TW + 2, HALO_TH + 2, 1,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }

    p.splits = 1; p.kb_per_split = p.ntaps * (p.chunks0 + p.chunks1);
    p.splitk_ws = nullptr; p.splitk_cnt = nullptr;
    p.dbg = g_tc_dbg;
    (void)splitk_cnt; (void)splitk_cnt_n;
    if (flags & DD_TC_SPLITK) {
        const int S = dd_conv_tc_splits(kind, B, H, W, Cin, Cout);
        DD_REQUIRE(S > 1, "conv_tc: DD_TC_SPLITK on a shape dd_conv_tc_splits() does not split");
        DD_REQUIRE(splitk_ws != nullptr && splitk_ws_floats >= (int64_t)S * B * H * W * Cout, "conv_tc: split-K workspace too small");
        DD_REQUIRE(residual == nullptr && !out_nchw_f32, "conv_tc: split-K writes raw partials only");
        p.splits = S; p.kb_per_split /= S; p.splitk_ws = splitk_ws; p.gn_stats = nullptr;
    }
    dim3 grid(p.tiles_w * p.tiles_h * tiles_n, Cout / p.bn, phases * p.splits);
    static const bool verbose = getenv("DD_TC_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "conv_tc kind=%d B=%d H=%d W=%d C=%d+%d Cout=%d grid=(%u,%u,%u) bn=%d num_kb=%d splits=%d kb_per=%d\n", kind, B, H, W,
                C1, C2, Cout, grid.x, grid.y, grid.z, p.bn, p.kb_per_split, p.splits, p.kb_per_split);
    const int ctas = (int)(grid.x * grid.y * grid.z);
    const bool pair = ((p.chunks0 + p.chunks1) % 2 == 0) && (p.chunks0 % 2 == 0);     // two chunks per stage never straddle the sources
    // CTA-pair (cta_group::2) form of the halo kernel: two consecutive pixel tiles per cluster, N = 128 or 256 per MMA.
    // Opt-in (DD_TC_PAIR): parity-green, but as a one-tile-per-CTA kernel it loses to the single-CTA form on this network
    // (3x3 128->128 @32x32 22.2 -> 25.0 us, 256->256 @16x16 17.5 -> 20.5 us, 512->128 @16x16 21.5 -> 18.8 us; profiles/README.md):
    // the 256-column epilogue is exposed and clusters schedule in pairs.  It is the base for a persistent pair kernel.
    const bool cta_pair = halo && (flags & DD_TC_PAIR) && p.bn == 128 && grid.x % 2 == 0 && grid.z == 1;
    p.pair_nt = (cta_pair && Cout % 256 == 0) ? 2 : 1;
    if (cta_pair) {
        static bool pattr = false;
        if (!pattr) {
            cudaError_t e = cudaFuncSetAttribute(conv_tc_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_halo2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
            if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute(pair): %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
            pattr = true;
        }
        rc = make_w_map(&p.tmB, wp, K, w_rows, 64 * p.pair_nt, 0);         // each CTA loads half of the pair's weight rows
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (p.splits > 1)
        launch_pdl(conv_tc_kernel<8, 64, 1, 2>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(8, 64, 1), st, p);
    else if (out_nchw_f32 || p.bn < 32)
        launch_pdl(conv_tc_kernel<3, 128, 1, 1>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(3, 128, 1), st, p);
    else if (cta_pair)
        launch_pair_pdl(conv_tc_halo2_kernel, dim3(grid.x, Cout / (p.bn * p.pair_nt), 1), dim3(TC_THREADS), HALO_SMEM, st, p);
    else if (halo)
        launch_pdl(conv_tc_halo_kernel, dim3(grid), dim3(TC_THREADS), HALO_SMEM, st, p);
    else if (ctas > num_sms())      // more than one wave: two CTAs per SM so epilogues overlap main loops
        launch_pdl(conv_tc_kernel<3, 128, 1, 0>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(3, 128, 1), st, p);
    else if (p.bn <= 64 && pair)
        launch_pdl(conv_tc_kernel<4, 64, 2, 0>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(4, 64, 2), st, p);
    else if (p.bn <= 64)
        launch_pdl(conv_tc_kernel<8, 64, 1, 0>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(8, 64, 1), st, p);
    else if (pair)
        launch_pdl(conv_tc_kernel<3, 128, 2, 0>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(3, 128, 2), st, p);
    else
        launch_pdl(conv_tc_kernel<6, 128, 1, 0>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(6, 128, 1), st, p);
    return check_launch("conv_tc");
}


// =============================================================================================
// Persistent TF32 convolution for the fp32 training programs (dd_conv_tc32).
// The resampling nets run 3x3 32->32 / 1x1 32<->64 convolutions over 2 M pixels: 16 384 tiles of 128 pixels with nine
// (or one, two) k-blocks each.  One CTA per tile spends ~4.6 of its 6 us in launch, TMEM allocation, barrier setup and
// the first load's latency (launch list: 340 us per conv against an 84 us HBM floor), so here a CTA walks tiles
// t = blockIdx.x, += gridDim.x with the operand ring running across tile boundaries and TWO accumulator buffers in TMEM:
// the epilogue of tile i (TMEM -> registers -> fp32 NHWC, bias / addend fused) overlaps the loads and MMAs of tile i+1.
// Warps: 0 = A-operand TMA, 6 = weight TMA, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
// =============================================================================================
namespace dd {

constexpr int P32_STAGES = 3;
constexpr int P32_STAGE_BYTES = TC_A_BYTES + 128 * 128;        // 128 pixels + up to 128 weight rows, 32 fp32 channels each
constexpr int P32_SMEM = P32_STAGES * P32_STAGE_BYTES + 1024 + 2048;
constexpr int P32_TMEM_COLS = 256;                              // two 128-column accumulators

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc32_persist_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + P32_STAGES * P32_STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (P32_STAGES + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * P32_STAGES + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * P32_STAGES + 2 + b); };
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * P32_STAGES + 4);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 1024u - smem_u32(smem_raw)));      // [2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cpt = p.chunks0 + p.chunks1;                  // 32-channel chunks per tap
    const int num_kb = p.ntaps * cpt;
    const int n_tiles = p.Cout / p.bn;
    const int tiles_mn = p.tiles_w * p.tiles_h * ((p.B + p.tn - 1) / p.tn) * n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        for (int s = 0; s < P32_STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(P32_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    pdl_sync();

    if (warp == 0) {
        // ===== A-operand producer =====
        const uint32_t tx = (uint32_t)p.rows_valid * 128u;
        const int chunks0 = p.chunks0;
        int st = 0, round = 0;
        for (int t = blockIdx.x; t < tiles_mn; t += gridDim.x) {
            const int m_tile = t / n_tiles;
            const int w0 = (m_tile % p.tiles_w) * p.tw, h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
            const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.tn;
            int rem = 0, ti = 0;
            for (int i = 0; i < num_kb; ++i) {
                const uint32_t fb = full_bar(st);
                if (round > 0) mbar_wait(empty_bar(st), (round - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(fb, tx);
                    const int cx = w0 + p.tap_dw[ti], cy = h0 + p.tap_dh[ti];
                    if (rem < chunks0) tma_load_5d(&p.tmA0, fb, base + st * P32_STAGE_BYTES, rem * 32, cx, cy, n0, 0);
                    else tma_load_5d(&p.tmA1, fb, base + st * P32_STAGE_BYTES, (rem - chunks0) * 32, cx, cy, n0, 0);
                }
                __syncwarp();
                if (++st == P32_STAGES) { st = 0; ++round; }
                if (++rem == cpt) { rem = 0; ++ti; }
            }
        }
    } else if (warp == 6) {
        // ===== weight producer =====
        const uint32_t tx = (uint32_t)p.bn * 128u;
        int st = 0, round = 0;
        for (int t = blockIdx.x; t < tiles_mn; t += gridDim.x) {
            const int brow = (t % n_tiles) * p.bn;
            for (int i = 0; i < num_kb; ++i) {
                const uint32_t fb = full_bar(st);
                if (round > 0) mbar_wait(empty_bar(st), (round - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(fb, tx);
                    tma_load_2d(&p.tmB, fb, base + st * P32_STAGE_BYTES + TC_A_BYTES, i * 32, brow);
                }
                __syncwarp();
                if (++st == P32_STAGES) { st = 0; ++round; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: accumulator buffer (it & 1), released by the epilogue through tempty =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        int st = 0, it = 0;
        uint32_t par = 0;
        for (int t = blockIdx.x; t < tiles_mn; t += gridDim.x, ++it) {
            const int ab = it & 1;
            if (it >= 2) mbar_wait(tempty_bar(ab), ((it >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)(ab * 128);
            for (int i = 0; i < num_kb; ++i) {
                mbar_wait(full_bar(st), par);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = umma_desc(base + st * P32_STAGE_BYTES), bd = umma_desc(base + st * P32_STAGE_BYTES + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32(dcol, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i | k) ? 1u : 0u);
                    umma_commit(empty_bar(st));
                }
                __syncwarp();
                if (++st == P32_STAGES) { st = 0; par ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(ab));
            __syncwarp();
        }
    } else {
        // ===== epilogue: y[pixel][c] = (acc + bias) [* mish'(z)] [+ addend];  optionally y2 = mish(y) =====
        const int q = warp & 3, r = q * 32 + lane, et = threadIdx.x - 64;
        const int bn = p.bn;
        float* yout = reinterpret_cast<float*>(p.out);
        float* yout2 = p.out2;
        const float* addp = reinterpret_cast<const float*>(p.residual);
        const float* mgp = p.mgrad;
        int it = 0;
        for (int t = blockIdx.x; t < tiles_mn; t += gridDim.x, ++it) {
            const int ab = it & 1;
            const int m_tile = t / n_tiles, cbase = (t % n_tiles) * bn;
            const int w0 = (m_tile % p.tiles_w) * p.tw, h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
            const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.tn;
            float* sb = s_bias + ab * 128;
            if (et < bn) sb[et] = p.bias ? p.bias[cbase + et] : 0.f;
            const int ww = r & (p.tw - 1), hh = (r >> p.tw_sh) & (p.th - 1), n = n0 + (r >> (p.tw_sh + p.th_sh));
            const bool valid = n < p.B && r < p.rows_valid;
            const int64_t off = (((int64_t)n * p.H + (h0 + hh)) * p.W + (w0 + ww)) * p.Cout + cbase;
            epi_bar();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 128);
            // operands of the fused epilogue do not depend on the accumulator: the first 32-column group is requested
            // BEFORE the wait on the MMAs (each lane reads its own pixel row: a latency-bound gather), later groups one ahead
            float4 ad[8], zg[8];
            auto fetch = [&](int c) {
                if (!valid) return;
                if (addp) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) ad[j] = reinterpret_cast<const float4*>(addp + off + c)[j];
                }
                if (mgp) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) zg[j] = reinterpret_cast<const float4*>(mgp + off + c)[j];
                }
            };
            fetch(0);
            mbar_wait(tfull_bar(ab), (it >> 1) & 1);
            tc_fence_after();
            for (int c = 0; c < bn; c += 32) {
                uint32_t acc[32];
                tmem_ld32_issue(trow + (uint32_t)c, acc);
                tmem_ld_wait();
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = make_float4(__uint_as_float(acc[4 * j]) + sb[c + 4 * j], __uint_as_float(acc[4 * j + 1]) + sb[c + 4 * j + 1],
                                       __uint_as_float(acc[4 * j + 2]) + sb[c + 4 * j + 2], __uint_as_float(acc[4 * j + 3]) + sb[c + 4 * j + 3]);
                    if (mgp) {
                        v[j].x *= mish_grad_fast(zg[j].x); v[j].y *= mish_grad_fast(zg[j].y);
                        v[j].z *= mish_grad_fast(zg[j].z); v[j].w *= mish_grad_fast(zg[j].w);
                    }
                    if (addp) { v[j].x += ad[j].x; v[j].y += ad[j].y; v[j].z += ad[j].z; v[j].w += ad[j].w; }
                }
                if (c + 32 < bn) fetch(c + 32);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        reinterpret_cast<float4*>(yout + off + c)[j] = v[j];
                        if (yout2)
                            reinterpret_cast<float4*>(yout2 + off + c)[j] =
                                make_float4(mish_fast(v[j].x), mish_fast(v[j].y), mish_fast(v[j].z), mish_fast(v[j].w));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(ab));            // 128 arrivals release the accumulator buffer
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(P32_TMEM_COLS) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------
// Halo form of the persistent TF32 convolution for the narrow 3x3 layers of the resampling nets (32 -> 32 channels on up to
// 256x256 maps, convblocks.py:103-104).  Loading nine tap-shifted operand tiles per 128 pixels makes these layers
// L2 -> SM bound (180 KB per tile, 2.95 GB per launch against ~12 TB/s: 350 us).  Here the CTA keeps ALL filter taps
// resident in shared memory (9 * Cin * 32 * 4 bytes <= 72 KB, one TMA box at start) and loads one (18 x 10)-pixel halo per
// 16 x 8 tile and 32-channel chunk (23 KB); the nine taps are nine shifted UMMA descriptors into it (row-group stride of
// 10 halo rows), as in the bf16 halo kernel.  Per tile 23 KB arrive instead of 180 KB.
// ---------------------------------------------------------------------------------------------
constexpr int H32_RING = 4;
constexpr int H32_W_MAX = 9 * 64 * 32 * 4;                                  // resident weights: Cin <= 64, 32 output channels
constexpr int H32_SMEM = H32_W_MAX + H32_RING * HALO_SLOT + 1024 + 2048 + 4 * 32 * 36 * 4;
constexpr int H32_BN = 32;

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc32_halo_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t hbase = base + H32_W_MAX;
    const uint32_t bars = hbase + H32_RING * HALO_SLOT;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (H32_RING + s); };
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * H32_RING + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * H32_RING + 2 + b); };
    const uint32_t wfull_bar = bars + 8u * (2 * H32_RING + 4);
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * H32_RING + 5);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 1024u - smem_u32(smem_raw)));      // [2][32]
    float* s_stage = reinterpret_cast<float*>(smem_raw + (bars + 2048u - smem_u32(smem_raw)));     // [4 warps][32][36] epilogue transpose

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = p.chunks0;                           // 32-channel chunks (single source)
    const int n_tiles = p.Cout / H32_BN;
    const int n_tile = blockIdx.x % n_tiles;                // fixed per CTA: its weights stay resident
    const int m_tiles = p.tiles_w * p.tiles_h * p.B;
    const int m_first = blockIdx.x / n_tiles, m_step = gridDim.x / n_tiles;
    const uint32_t w_tap_bytes = H32_BN * 128u, w_chunk_bytes = 9u * w_tap_bytes;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmH0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB3)) : "memory");
        for (int s = 0; s < H32_RING; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 128); }
        mbar_init(wfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    pdl_sync();

    if (warp == 0) {
        // ===== producer: the resident weights once, then one halo per (tile, chunk) =====
        if (elect_one()) {
            mbar_expect_tx(wfull_bar, (uint32_t)chunks * w_chunk_bytes);
            for (int c = 0; c < chunks; ++c) tma_load_3d(&p.tmB3, wfull_bar, base + c * w_chunk_bytes, c * 32, n_tile * H32_BN, 0);
        }
        __syncwarp();
        int st = 0, round = 0;
        for (int m = m_first; m < m_tiles; m += m_step) {
            const int w0 = (m % p.tiles_w) * HALO_TW, h0 = ((m / p.tiles_w) % p.tiles_h) * HALO_TH, n0 = m / (p.tiles_w * p.tiles_h);
            for (int c = 0; c < chunks; ++c) {
                if (round > 0) mbar_wait(empty_bar(st), (round - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(full_bar(st), HALO_TX);
                    tma_load_5d(&p.tmH0, full_bar(st), hbase + st * HALO_SLOT, c * 32, w0 - 1, h0 - 1, n0, 0);
                }
                __syncwarp();
                if (++st == H32_RING) { st = 0; ++round; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H32_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const uint64_t a_hi = ((uint64_t)(((HALO_TW + 2) * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        mbar_wait(wfull_bar, 0);
        int st = 0, it = 0;
        uint32_t par = 0;
        for (int m = m_first; m < m_tiles; m += m_step, ++it) {
            const int ab = it & 1;
            if (it >= 2) mbar_wait(tempty_bar(ab), ((it >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)(ab * H32_BN);
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(full_bar(st), par);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t halo = hbase + st * HALO_SLOT, wch = base + c * w_chunk_bytes;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t rowA = halo + (uint32_t)((tap / 3) * (HALO_TW + 2) + tap % 3) * 128u;
                        const uint64_t ad = (uint64_t)((rowA & 0x3FFFFu) >> 4) | a_hi;
                        const uint64_t bd = umma_desc(wch + tap * w_tap_bytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_tf32(dcol, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (c | tap | k) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(st));
                }
                __syncwarp();
                if (++st == H32_RING) { st = 0; par ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(ab));
            __syncwarp();
        }
    } else if (warp < 6) {
        // ===== epilogue: TMEM row per lane -> warp-private shared-memory transpose -> coalesced fp32 NHWC I/O =====
        // A lane owns one pixel's 32 channels after the TMEM load; written that way every 16-byte store of a warp lands in a
        // different 128-byte row (and the fused operands are gathered the same way).  Through a [32][36]-float staging tile the
        // warp instead moves (4 pixels x 128 bytes) per instruction: lanes 8k..8k+7 cover one pixel's row.
        const int q = warp & 3, et = threadIdx.x - 64;
        float* yout = reinterpret_cast<float*>(p.out);
        float* yout2 = p.out2;
        const float* addp = reinterpret_cast<const float*>(p.residual);
        const float* mgp = p.mgrad;
        const int cbase = n_tile * H32_BN;
        float* stg = s_stage + q * (32 * 36);
        if (et < H32_BN) { s_bias[et] = p.bias ? p.bias[cbase + et] : 0.f; }
        epi_bar();
        const int c4 = lane & 7, prow = lane >> 3;          // transposed view: this lane's 4 channels, its pixel within a group of 4
        const float4 bv = *reinterpret_cast<const float4*>(s_bias + 4 * c4);
        int it = 0;
        for (int m = m_first; m < m_tiles; m += m_step, ++it) {
            const int ab = it & 1;
            const int w0 = (m % p.tiles_w) * HALO_TW, h0 = ((m / p.tiles_w) % p.tiles_h) * HALO_TH, n0 = m / (p.tiles_w * p.tiles_h);
            // tile rows q*32 + 4i + prow, i = 0..7: image row h0 + 4q + (4i + prow) / 8, column w0 + (4i + prow) % 8
            int64_t off[8];
            float4 ad[8], zg[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + prow;
                off[i] = (((int64_t)n0 * p.H + (h0 + 4 * q + (rr >> 3))) * p.W + (w0 + (rr & 7))) * p.Cout + cbase + 4 * c4;
                if (addp) ad[i] = *reinterpret_cast<const float4*>(addp + off[i]);
                if (mgp) zg[i] = *reinterpret_cast<const float4*>(mgp + off[i]);
            }
            mbar_wait(tfull_bar(ab), (it >> 1) & 1);
            tc_fence_after();
            uint32_t acc[32];
            tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * H32_BN), acc);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(tempty_bar(ab));            // the accumulator is in registers: release the buffer before the stores
            __syncwarp();                           // the previous tile's transposed reads of the staging tile are done
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * j) = make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + prow) * 36 + 4 * c4);
                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                if (mgp) {
                    v.x *= mish_grad_fast(zg[i].x); v.y *= mish_grad_fast(zg[i].y); v.z *= mish_grad_fast(zg[i].z); v.w *= mish_grad_fast(zg[i].w);
                }
                if (addp) { v.x += ad[i].x; v.y += ad[i].y; v.z += ad[i].z; v.w += ad[i].w; }
                *reinterpret_cast<float4*>(yout + off[i]) = v;
                if (yout2) *reinterpret_cast<float4*>(yout2 + off[i]) = make_float4(mish_fast(v.x), mish_fast(v.y), mish_fast(v.z), mish_fast(v.w));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64) : "memory");
    }
}

}  // namespace dd

// fp32 training form of dd_conv_tc: fp32 NHWC activations, fp32 packed weights [rows][tap*Cin + c], TF32 tensor-core math
// (10-bit mantissa operands, fp32 accumulate -- what torch's cudnn.allow_tf32 default gives the reference on a GPU).
extern "C" int dd_conv_tc32(int kind, const float* x, const float* x2, int C1, int C2, const float* wp, int w_rows, const float* bias,
                            const float* addend, float* y, float* y_mish, const float* mish_grad_of, int B, int H, int W, int Cout,
                            void* stream) {
    DD_REQUIRE(kind == DD_TC_CONV3x3 || kind == DD_TC_CONV1x1, "conv_tc32: 3x3 stride-1 and 1x1 only (kind %d)", kind);
    DD_REQUIRE(C1 > 0 && C1 % 32 == 0 && C2 >= 0 && C2 % 32 == 0, "conv_tc32: channel counts (%d,%d) must be multiples of 32", C1, C2);
    DD_REQUIRE((C2 == 0) == (x2 == nullptr), "conv_tc32: x2/C2 mismatch");
    DD_REQUIRE(is_pow2(H) && is_pow2(W) && B > 0, "conv_tc32: H=%d, W=%d must be powers of two", H, W);
    DD_REQUIRE(Cout >= 32 && Cout % 32 == 0 && w_rows >= Cout, "conv_tc32: Cout=%d must be a multiple of 32", Cout);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc32_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P32_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc32_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H32_SMEM);
        if (e != cudaSuccess) { set_error("conv_tc32: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
        attr_done = true;
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.tw = W < 128 ? W : 128;
    p.th = (128 / p.tw) < H ? (128 / p.tw) : H;
    p.tn = 128 / (p.tw * p.th);
    p.rows_valid = p.tw * p.th * p.tn;
    while ((1 << p.tw_sh) < p.tw) ++p.tw_sh;
    while ((1 << p.th_sh) < p.th) ++p.th_sh;
    p.tiles_w = W / p.tw; p.tiles_h = H / p.th;
    const int tiles_n = (B + p.tn - 1) / p.tn;
    p.B = B; p.H = H; p.W = W;
    p.chunks0 = C1 / 32; p.chunks1 = C2 / 32;
    p.Cout = Cout; p.cout_valid = Cout;
    p.bn = Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32);
    p.out = y; p.bias = bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(addend);
    p.out2 = y_mish; p.mgrad = mish_grad_of;
    p.out_mul = 1;
    if (kind == DD_TC_CONV3x3) {
        p.ntaps = 9;
        for (int t = 0; t < 9; ++t) { p.tap_dh[t] = (int8_t)(t / 3 - 1); p.tap_dw[t] = (int8_t)(t % 3 - 1); p.tap_plane[t] = 0; }
    } else {
        p.ntaps = 1;
    }
    const int Cin = C1 + C2, K = p.ntaps * Cin;
    p.rows_per_phase = w_rows;
    DD_REQUIRE(w_rows % p.bn == 0, "conv_tc32: packed weight rows %d must be a multiple of the %d-wide tile", w_rows, p.bn);
    int rc = make_act_map(&p.tmA0, x, C1, C1, W, H, B, 1, p.tw, p.th, p.tn, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, true);
    if (rc) return rc;
    rc = make_act_map(&p.tmA1, x2 ? x2 : x, x2 ? C2 : C1, x2 ? C2 : C1, W, H, B, 1, p.tw, p.th, p.tn, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, true);
    if (rc) return rc;
    rc = make_w_map(&p.tmB, wp, K, w_rows, p.bn, 0, true);
    if (rc) return rc;
    p.splits = 1; p.kb_per_split = p.ntaps * (p.chunks0 + p.chunks1);
    p.dbg = g_tc_dbg;
    static const bool halo32_off = getenv("DD_NO_HALO32") != nullptr;
    if (!halo32_off && kind == DD_TC_CONV3x3 && C2 == 0 && C1 <= 64 && H >= HALO_TH && W >= HALO_TW && Cout <= 64) {
        // narrow 3x3 layers: resident weights + one halo per tile
        p.tw = HALO_TW; p.th = HALO_TH; p.tn = 1; p.rows_valid = 128; p.tw_sh = 3; p.th_sh = 4;
        p.tiles_w = W / HALO_TW; p.tiles_h = H / HALO_TH; p.bn = H32_BN;
        rc = make_act_map(&p.tmH0, x, C1, C1, W, H, B, 1, HALO_TW + 2, HALO_TH + 2, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, true);
        if (rc) return rc;
        rc = make_w_map_taps(&p.tmB3, wp, C1, w_rows, H32_BN, true);
        if (rc) return rc;
        const int n_t = Cout / H32_BN, m_t = p.tiles_w * p.tiles_h * B;
        int per = num_sms() / n_t;                      // CTAs per output-channel tile: one CTA per SM
        if (per > m_t) per = m_t;
        launch_pdl(conv_tc32_halo_kernel, dim3(per * n_t), dim3(TC_THREADS), H32_SMEM, (cudaStream_t)stream, p);
        return check_launch("conv_tc32");
    }
    const int tiles = p.tiles_w * p.tiles_h * tiles_n * (Cout / p.bn);
    const int ctas = tiles < 2 * num_sms() ? tiles : 2 * num_sms();            // persistent: two CTAs per SM walk the tiles
    launch_pdl(conv_tc32_persist_kernel, dim3(ctas), dim3(TC_THREADS), P32_SMEM, (cudaStream_t)stream, p);
    return check_launch("conv_tc32");
}

// =============================================================================================
// TF32 weight gradient on tcgen05:  dW[tap][ci][co] += sum_pixels Xa[pixel + tap][ci] * dY[pixel][co]
// (autograd of blocks.py:78,103,123-124 / convblocks.py:29-67; Xa is the conv's input, already activated when the
// block applies Mish first).  GEMM view with the PIXELS as the K dimension:
//   M = 128 rows = four "units" (tap, 32-channel chunk) of the input, N = a tile of output channels, K = pixels.
// kind::tf32 has no MN-major operand form (measured: any transpose bit in the instruction descriptor yields an all-zero
// accumulator, profiles/README.md), so both operands are read K-major from CHANNEL-MAJOR (NCHW) fp32 copies of Xa and dY:
// one operand row = one channel's 32 consecutive pixels (128 bytes), boxes (kw x kh pixels, 32 or N channels) land as the
// canonical 128B-swizzled tiles.  A unit's tap shift is a shift of its TMA coordinates (zero fill outside the map = the
// conv padding).  The pixel range is split over the grid; partial sums meet in dW through red.global.add.f32.
// =============================================================================================
namespace dd {

struct WgParams {
    CUtensorMap tmX0, tmX1, tmG;      // (Wp, H + 2, C, B) / (Wp, H, Cout, B) views of the padded channel-major copies
    int8_t tap_dw[9], tap_dh[9];
    int ntaps, chunks0, chunks1, units;
    int cw, chn;                      // a K chunk = 32 consecutive pixels of one (padded) row; cw chunks per row, chn per image
    int chunks_total, stages_per_cta; // 32-pixel chunks over the batch; pipeline stages (WG_KC chunks each) per CTA
    int N, Cout, rows_total;
    float* dw;
};
constexpr int WG_KC = 2;                          // 32-pixel chunks per pipeline stage
constexpr int WG_A_BYTES = 128 * 128;             // 128 rows (4 units x 32 channels) x 32 pixels fp32
constexpr int WG_STAGES = 3;
constexpr int WG_STAGE_BYTES = WG_KC * (WG_A_BYTES + 128 * 128);      // + up to 128 dY channels
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 1024;
constexpr int WG_THREADS = 192;                   // warps: TMA producer, MMA issuer + TMEM owner, 4 x epilogue

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int TMEM_COLS>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc32_kernel(const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + WG_STAGES * WG_STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (WG_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * WG_STAGES);
    const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
    const int chunk0 = split * p.stages_per_cta * WG_KC;
    const int nch = min(p.stages_per_cta * WG_KC, p.chunks_total - chunk0);      // chunks of this CTA (>= 1)
    const int nst = (nch + WG_KC - 1) / WG_KC;
    const uint32_t b_bytes = (uint32_t)p.N * 128u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmX0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmG)) : "memory");
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    pdl_sync();

    if (warp == 0) {
        // ===== producer: per 32-pixel chunk, the four input units of this group (each with its tap shift) + N dY channels =====
        const int cpt = p.chunks0 + p.chunks1;
        int u_tap[4], u_ch[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int u = g * 4 + j;
            if (u >= p.units) u = 0;                        // padding unit of the last group: loaded, never written back
            u_tap[j] = u / cpt; u_ch[j] = u % cpt;
        }
        int st = 0, round = 0;
        for (int s = 0; s < nst; ++s) {
            const int kcs = min(WG_KC, nch - s * WG_KC);
            const uint32_t fb = full_bar(st), sS = base + st * WG_STAGE_BYTES;
            if (round > 0) mbar_wait(empty_bar(st), (round - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(fb, (uint32_t)kcs * (WG_A_BYTES + b_bytes));
                for (int kc = 0; kc < kcs; ++kc) {
                    const int chunk = chunk0 + s * WG_KC + kc;
                    const int n = chunk / p.chn, rc = chunk % p.chn;
                    const int w0 = (rc % p.cw) * 32, h0 = rc / p.cw;
                    const uint32_t sA = sS + kc * (WG_A_BYTES + b_bytes), sB = sA + WG_A_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = u_ch[j], tp = u_tap[j];
                        // row shift = coordinate (+1: one zero row above and below); column shift = which pre-shifted copy:
                        // a TMA box origin must be 16-byte aligned, w0 - 1 in the innermost dimension is not (illegal instruction)
                        const int sc = p.ntaps == 9 ? p.tap_dw[tp] + 1 : 0;
                        if (c < p.chunks0) tma_load_5d(&p.tmX0, fb, sA + j * 4096, w0, h0 + p.tap_dh[tp] + 1, c * 32, n, sc);
                        else tma_load_5d(&p.tmX1, fb, sA + j * 4096, w0, h0 + p.tap_dh[tp] + 1, (c - p.chunks0) * 32, n, sc);
                    }
                    tma_load_4d(&p.tmG, fb, sB, w0, h0, n_tile * p.N, n);
                }
            }
            __syncwarp();
            if (++st == WG_STAGES) { st = 0; ++round; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[128 x N] += A[128 x 8 pixels] . B[N x 8 pixels]^T, K-major TF32, four K = 8 steps per chunk =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int st = 0;
        uint32_t par = 0;
        for (int s = 0; s < nst; ++s) {
            const int kcs = min(WG_KC, nch - s * WG_KC);
            mbar_wait(full_bar(st), par);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sS = base + st * WG_STAGE_BYTES;
                for (int kc = 0; kc < kcs; ++kc) {
                    const uint64_t ad = umma_desc(sS + kc * (WG_A_BYTES + b_bytes)), bd = umma_desc(sS + kc * (WG_A_BYTES + b_bytes) + WG_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (s | kc | k) ? 1u : 0u);
                }
                umma_commit(empty_bar(st));
            }
            __syncwarp();
            if (++st == WG_STAGES) { st = 0; par ^= 1u; }
        }
        if (elect_one()) umma_commit(tmem_full_bar);
        __syncwarp();
    } else {
        // ===== epilogue: accumulate the partial tile into dW (rows = tap*Cin + ci, columns = co) =====
        const int q = warp & 3, r = q * 32 + lane;
        const int row = g * 128 + r;
        const bool valid = row < p.rows_total;
        float* dst = p.dw + (int64_t)row * p.Cout + n_tile * p.N;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c = 0; c < p.N; c += 32) {
            uint32_t a[32];
            tmem_ld32_issue(trow + (uint32_t)c, a);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    red_add_v4(dst + c + 4 * j, __uint_as_float(a[4 * j]), __uint_as_float(a[4 * j + 1]), __uint_as_float(a[4 * j + 2]),
                               __uint_as_float(a[4 * j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// (W, H, C, B) fp32 view of a channel-major tensor; box = (32, 1, rows channels, 1): `rows` operand rows of 32 pixels (128 bytes)
static int make_nchw_map(CUtensorMap* tm, const void* ptr, int C, int W, int H, int B, int kw, int kh, int rows, int copies = 0) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DD_ERR_CUDA; }
    cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B, (cuuint64_t)(copies > 0 ? copies : 1)};
    cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4, (cuuint64_t)B * C * H * W * 4};
    cuuint32_t box[5] = {(cuuint32_t)kw, (cuuint32_t)kh, (cuuint32_t)rows, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, copies > 0 ? 5 : 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(NCHW C=%d W=%d H=%d B=%d box %d,%d,%d) failed: %d", C, W, H, B, kw, kh, rows, (int)r); return DD_ERR_CUDA; }
    return DD_OK;
}

}  // namespace dd

extern "C" int dd_conv_wgrad_tc32(int kind, const float* x_nchw, const float* x2_nchw, int C1, int C2, const float* dy_nchw, float* dw,
                                  int B, int H, int W, int Wp, int Cout, void* stream) {
    DD_REQUIRE(kind == DD_TC_CONV3x3 || kind == DD_TC_CONV1x1, "conv_wgrad_tc32: 3x3 stride-1 and 1x1 only (kind %d)", kind);
    DD_REQUIRE(C1 > 0 && C1 % 32 == 0 && C2 >= 0 && C2 % 32 == 0 && Cout > 0 && Cout % 32 == 0,
               "conv_wgrad_tc32: channel counts (%d,%d -> %d) must be multiples of 32", C1, C2, Cout);
    DD_REQUIRE((C2 == 0) == (x2_nchw == nullptr), "conv_wgrad_tc32: x2/C2 mismatch");
    DD_REQUIRE(H > 0 && W > 0 && B > 0 && Wp >= W && Wp % 32 == 0, "conv_wgrad_tc32: padded row width Wp=%d must be a multiple of 32 >= W=%d", Wp, W);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        if (e != cudaSuccess) { set_error("conv_wgrad_tc32: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
        attr_done = true;
    }
    WgParams p;
    memset(&p, 0, sizeof(p));
    p.cw = Wp / 32;
    p.chn = p.cw * H;
    p.chunks_total = p.chn * B;
    p.ntaps = kind == DD_TC_CONV3x3 ? 9 : 1;
    for (int t = 0; t < p.ntaps; ++t) {
        p.tap_dh[t] = (int8_t)(p.ntaps == 9 ? t / 3 - 1 : 0);
        p.tap_dw[t] = (int8_t)(p.ntaps == 9 ? t % 3 - 1 : 0);
    }
    p.chunks0 = C1 / 32; p.chunks1 = C2 / 32;
    p.units = p.ntaps * (p.chunks0 + p.chunks1);
    p.N = Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32);
    p.Cout = Cout; p.rows_total = p.ntaps * (C1 + C2); p.dw = dw;
    const int groups = (p.units + 3) / 4, n_tiles = Cout / p.N;
    const int stages_total = (p.chunks_total + WG_KC - 1) / WG_KC;
    int S = (2 * num_sms()) / (groups * n_tiles);
    if (S < 1) S = 1;
    if (S > stages_total) S = stages_total;
    p.stages_per_cta = (stages_total + S - 1) / S;
    S = (stages_total + p.stages_per_cta - 1) / p.stages_per_cta;
    const int copies = p.ntaps == 9 ? 3 : 1;
    int rc = make_nchw_map(&p.tmX0, x_nchw, C1, Wp, H + 2, B, 32, 1, 32, copies);
    if (rc) return rc;
    rc = make_nchw_map(&p.tmX1, x2_nchw ? x2_nchw : x_nchw, x2_nchw ? C2 : C1, Wp, H + 2, B, 32, 1, 32, copies);
    if (rc) return rc;
    rc = make_nchw_map(&p.tmG, dy_nchw, Cout, Wp, H, B, 32, 1, p.N);
    if (rc) return rc;
    dim3 grid(groups, n_tiles, S);
    cudaStream_t st = (cudaStream_t)stream;
    if (p.N == 32) launch_pdl(wgrad_tc32_kernel<32>, dim3(grid), dim3(WG_THREADS), WG_SMEM, st, p);
    else if (p.N == 64) launch_pdl(wgrad_tc32_kernel<64>, dim3(grid), dim3(WG_THREADS), WG_SMEM, st, p);
    else launch_pdl(wgrad_tc32_kernel<128>, dim3(grid), dim3(WG_THREADS), WG_SMEM, st, p);
    return check_launch("conv_wgrad_tc32");
}
