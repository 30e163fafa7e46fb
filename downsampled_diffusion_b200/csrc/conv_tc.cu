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
constexpr int TC_THREADS = 192;
// Pipeline variants <STAGES, BROWS>: BROWS = rows of the weight slot (>= bn).
//   <3,128>: 96 KB  -> 2 CTAs/SM, for grids that fill the GPU (epilogue of one CTA overlaps the other's loop)
//   <6,128>: 192 KB -> 1 CTA/SM, grids of at most one wave: twice the loads in flight per CTA
//   <8, 64>: 192 KB -> 1 CTA/SM, low-resolution layers run with bn = 64 (twice the CTAs) and 8 stages
constexpr int tc_stage_bytes(int brows) { return TC_A_BYTES + brows * TC_BK * 2; }
constexpr int tc_smem_bytes(int stages, int brows) { return stages * tc_stage_bytes(brows) + 1024 /*align*/ + 1024 /*barriers + bias*/; }
constexpr int TC_TMEM_COLS = 128;

struct TcParams {
    CUtensorMap tmA0, tmA1, tmB;
    int8_t tap_dw[16], tap_dh[16], tap_plane[16];
    int ntaps;              // taps per phase
    int chunks0, chunks1;   // 64-channel chunks of source 0 / 1
    int tw, th, tn, tiles_w, tiles_h;
    int B, H, W;            // GEMM pixel grid
    int Cout, cout_valid, bn, rows_per_phase;
    int out_mul;            // 1, or 2 for the sub-pixel phases of the transposed conv
    int out_nchw_f32;
    int G, cpg_mask, cpg_shift;
    int rows_valid;             // tw*th*tn (< 128 when one image has fewer than 128 pixels and tn is forced to 1)
    int w_per_sample;           // weights are (B, rows, K): every image multiplies its own matrix (fused attention output)
    int splits, kb_per_split;   // split-K over the (tap, chunk) loop; partial sums meet in splitk_ws
    void* out;
    const float* bias;
    const __nv_bfloat16* residual;
    float* gn_stats;
    float* splitk_ws;           // (tiles, bn/4, 128, 4) fp32, all zero between launches (self-cleaning)
    int32_t* splitk_cnt;        // per-tile arrival counters, all zero between launches
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
        if (clock64() - t0 > 4000000000LL) { printf("conv_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// ---------------------------------------------------------------------------------------------
template <int TC_STAGES, int BROWS>
__global__ void __launch_bounds__(TC_THREADS, (TC_STAGES <= 3 ? 2 : 1)) conv_tc_kernel(const __grid_constant__ TcParams p) {
    constexpr int TC_STAGE_BYTES = tc_stage_bytes(BROWS);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;     // full[S], empty[S], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * TC_STAGES);
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * TC_STAGES + 1);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y;
    const int phase = p.splits > 1 ? 0 : blockIdx.z;
    const int split = p.splits > 1 ? blockIdx.z : 0;
    const int w0 = (m_tile % p.tiles_w) * p.tw;
    const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
    const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.tn;
    const int cpt = p.chunks0 + p.chunks1;
    const int kb_lo = split * p.kb_per_split;
    const int num_kb = min(p.ntaps * cpt, kb_lo + p.kb_per_split) - kb_lo;     // k-blocks of this CTA
    float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));   // 128 floats (barriers use < 160 B)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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
    pdl_sync();      // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t stage_tx = (uint32_t)(p.rows_valid + p.bn) * TC_BK * 2;   // bytes the two TMA boxes deliver
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                if (kb >= TC_STAGES) mbar_wait(empty_bar(s), ((kb / TC_STAGES) - 1) & 1);
                const int kg = kb_lo + kb;                       // global k-block index
                const int tap = kg / cpt, rem = kg - tap * cpt;
                const int ti = phase * p.ntaps + tap;
                const uint32_t sA = base + s * TC_STAGE_BYTES, sB = sA + TC_A_BYTES;
                mbar_expect_tx(full_bar(s), stage_tx);
                if (rem < p.chunks0)
                    tma_load_5d(&p.tmA0, full_bar(s), sA, rem * 64, w0 + p.tap_dw[ti], h0 + p.tap_dh[ti], n0, p.tap_plane[ti]);
                else
                    tma_load_5d(&p.tmA1, full_bar(s), sA, (rem - p.chunks0) * 64, w0 + p.tap_dw[ti], h0 + p.tap_dh[ti], n0, p.tap_plane[ti]);
                if (p.w_per_sample) tma_load_3d(&p.tmB, full_bar(s), sB, kg * 64, n_tile * p.bn, n0);
                else tma_load_2d(&p.tmB, full_bar(s), sB, kg * 64, phase * p.rows_per_phase + n_tile * p.bn);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=bn, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                mbar_wait(full_bar(s), (kb / TC_STAGES) & 1);
                tc_fence_after();
                const uint32_t sA = base + s * TC_STAGE_BYTES, sB = sA + TC_A_BYTES;
                const uint64_t ad = umma_desc(sA), bd = umma_desc(sB);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                    umma_f16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                umma_commit(empty_bar(s));          // frees this smem stage when the MMAs retire
            }
            umma_commit(tmem_full_bar);             // accumulator complete
        }
    } else {
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
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);

        bool finisher = true;
        float* wst = nullptr;
        if (p.splits > 1) {
            // ---- split-K: add this CTA's partial tile into the fp32 workspace, last arrival finishes ----
            const int tile_id = m_tile * gridDim.y + n_tile;
            wst = p.splitk_ws + (int64_t)tile_id * TC_BM * p.bn;
            uint32_t acc[16];
            for (int ch = 0; ch < p.bn; ch += 16) {
                tmem_ld16_issue(trow + (uint32_t)ch, acc);
                tmem_ld_wait();
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    red_add_v4(wst + ((int64_t)((ch >> 2) + k4) * TC_BM + r) * 4, __uint_as_float(acc[4 * k4]),
                               __uint_as_float(acc[4 * k4 + 1]), __uint_as_float(acc[4 * k4 + 2]), __uint_as_float(acc[4 * k4 + 3]));
            }
            __threadfence();
            epi_bar();
            int* s_flag = reinterpret_cast<int*>(s_bias + 128);
            if (et == 0) {
                const int ticket = atomicAdd(p.splitk_cnt + tile_id, 1);
                const int last = (ticket == p.splits - 1);
                if (last) p.splitk_cnt[tile_id] = 0;                 // self-cleaning for the next launch
                *s_flag = last;
            }
            epi_bar();
            finisher = (*s_flag != 0);
            if (finisher) __threadfence();
        }

        if (finisher) {
            float gs = 0.f, gq = 0.f;
            uint32_t acc[16], nxt[16];
            if (p.splits == 1) { tmem_ld16_issue(trow, acc); tmem_ld_wait(); }
            for (int ch = 0; ch < p.bn; ch += 16) {
                const bool more = ch + 16 < p.bn;
                float v[16];
                if (p.splits > 1) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        float4* wp4 = reinterpret_cast<float4*>(wst + ((int64_t)((ch >> 2) + k4) * TC_BM + r) * 4);
                        const float4 t = __ldcg(wp4);
                        *wp4 = make_float4(0.f, 0.f, 0.f, 0.f);      // leave the workspace zeroed
                        v[4 * k4] = t.x; v[4 * k4 + 1] = t.y; v[4 * k4 + 2] = t.z; v[4 * k4 + 3] = t.w;
                    }
                } else {
                    if (more) tmem_ld16_issue(trow + (uint32_t)(ch + 16), nxt);      // overlaps the math below
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
                }
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
                if (p.splits == 1 && more) {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = nxt[j];
                }
                if (use_res && more) { res_cur[0] = res_nxt[0]; res_cur[1] = res_nxt[1]; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
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

static int make_act_map(CUtensorMap* tm, const void* ptr, int C, int pitch, int W, int H, int N, int P, int tw, int th, int tn) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DD_ERR_CUDA; }
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)P};
    cuuint64_t strides[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2,
                             (cuuint64_t)N * H * W * pitch * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation C=%d W=%d H=%d N=%d P=%d box %d,%d,%d) failed: %d", C, W, H, N, P, tw, th, tn, (int)r); return DD_ERR_CUDA; }
    return DD_OK;
}

static int make_w_map(CUtensorMap* tm, const void* ptr, int K, int rows, int bn, int batch) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DD_ERR_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * rows};
    cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, batch > 0 ? 3 : 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights K=%d rows=%d bn=%d) failed: %d", K, rows, bn, (int)r); return DD_ERR_CUDA; }
    return DD_OK;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace dd

using namespace dd;

extern "C" int dd_zero(void* ptr, int64_t bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dd_zero: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
    return DD_OK;
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
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(3, 128));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<6, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(6, 128));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<8, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(8, 64));
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
        attr_done = true;
    }
    const bool wps = (flags & DD_TC_W_PER_SAMPLE) != 0;
    DD_REQUIRE(!wps || kind == DD_TC_CONV1x1, "conv_tc: per-sample weights are for 1x1 convs only");

    TcParams p;
    memset(&p, 0, sizeof(p));
    // tile geometry over the GEMM pixel grid
    p.tw = W < 128 ? W : 128;
    p.th = (128 / p.tw) < H ? (128 / p.tw) : H;
    p.tn = wps ? 1 : 128 / (p.tw * p.th);      // per-sample weights: one image per tile (rows beyond it are ignored)
    p.rows_valid = p.tw * p.th * p.tn;
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
    int planes = 1, phases = 1, K;
    if (kind == DD_TC_CONV3x3) {
        p.ntaps = 9;
        for (int t = 0; t < 9; ++t) { p.tap_dh[t] = (int8_t)(t / 3 - 1); p.tap_dw[t] = (int8_t)(t % 3 - 1); p.tap_plane[t] = 0; }
    } else if (kind == DD_TC_CONV1x1) {
        p.ntaps = 1;
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
    int rc = make_act_map(&p.tmA0, x, C1, x_pitch, W, H, B, planes, p.tw, p.th, p.tn);
    if (rc) return rc;
    rc = make_act_map(&p.tmA1, x2 ? x2 : x, x2 ? C2 : C1, x2 ? C2 : x_pitch, W, H, B, planes, p.tw, p.th, p.tn);
    if (rc) return rc;
    rc = make_w_map(&p.tmB, wp, K, w_rows, p.bn, wps ? B : 0);
    if (rc) return rc;

    // split-K for layers whose output tiles cannot fill the GPU (low-resolution levels): spread the
    // (tap, chunk) loop over up to 12 CTAs per tile, aiming at ~2 CTAs per SM.
    const int tiles = p.tiles_w * p.tiles_h * tiles_n * (Cout / p.bn);
    const int num_kb = p.ntaps * (p.chunks0 + p.chunks1);
    p.splits = 1; p.kb_per_split = num_kb;
    p.splitk_ws = splitk_ws; p.splitk_cnt = splitk_cnt;
    static const bool splitk_on = getenv("DD_SPLITK") != nullptr;   // red.add reduction measured slower than deep pipelines (profiles/README.md)
    if (splitk_on && splitk_ws && splitk_cnt && phases == 1 && !wps && tiles < num_sms() && num_kb >= 6 &&
        (int64_t)tiles * TC_BM * p.bn <= splitk_ws_floats && tiles <= splitk_cnt_n) {
        int sp = (2 * num_sms()) / tiles;
        if (sp > num_kb / 3) sp = num_kb / 3;
        if (sp > 12) sp = 12;
        if (sp >= 2) {
            p.kb_per_split = (num_kb + sp - 1) / sp;
            p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
        }
    }
    dim3 grid(p.tiles_w * p.tiles_h * tiles_n, Cout / p.bn, phases * p.splits);
    static const bool verbose = getenv("DD_TC_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "conv_tc kind=%d B=%d H=%d W=%d C=%d+%d Cout=%d grid=(%u,%u,%u) bn=%d num_kb=%d splits=%d kb_per=%d\n", kind, B, H, W,
                C1, C2, Cout, grid.x, grid.y, grid.z, p.bn, num_kb, p.splits, p.kb_per_split);
    const int ctas = (int)(grid.x * grid.y * grid.z);
    if (p.bn <= 64 && ctas <= num_sms())
        launch_pdl(conv_tc_kernel<8, 64>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(8, 64), (cudaStream_t)stream, p);
    else if (ctas <= num_sms())
        launch_pdl(conv_tc_kernel<6, 128>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(6, 128), (cudaStream_t)stream, p);
    else
        launch_pdl(conv_tc_kernel<3, 128>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(3, 128), (cudaStream_t)stream, p);
    return check_launch("conv_tc");
}
