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
#include "conv_tc_common.cuh"

namespace dd {

// Two MMA issuers (conv_tc_kernel<..., MW = 2>) leave two partial accumulators, 128 TMEM columns apart: every epilogue warp folds
// the second into the first for its own 32 lanes before the epilogue proper reads it (tcgen05.ld x2, add, tcgen05.st).
__device__ __forceinline__ void tc_fold_partials(const TcParams& p, uint32_t tmem_base, int warp) {
    if (p.mma_parts < 2) return;
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = 0; c < p.bn; c += 16) {
        uint32_t a[16], b[16];
        tmem_ld16x2(trow + (uint32_t)c, trow + (uint32_t)(TC_TMEM_COLS + c), a, b);
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(a[j]) + __uint_as_float(b[j]));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(trow + (uint32_t)c), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
                       "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- epilogue shared by every pipeline variant: 4 warps, TMEM lane quadrant = warp % 4 ----
__device__ __forceinline__ void tc_epilogue(const TcParams& p, uint32_t tmem_base, uint32_t tmem_full_bar, float* s_bias,
                                            int m_tile, int n_tile, int phase, int w0, int h0, int n0, int warp, int lane) {
    // ===== epilogue: 4 warps, TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ww = r % p.tw, hh = (r / p.tw) % p.th, nl = r / (p.tw * p.th);
    const int n = n0 + nl;
    const bool valid = n < p.B && r < p.rows_valid && h0 + hh < p.H && w0 + ww < p.W;     // tiles may hang over the edge of a ragged map
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
    tc_fold_partials(p, tmem_base, warp);
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



// ---- staged epilogue (bf16 NHWC output, bn >= 32) -------------------------------------------------
// Phase A: each of the 128 epilogue threads drains its accumulator row from TMEM (32-column loads, two in
//          flight), adds the bias, keeps per-8-channel {sum, sumsq} partials in registers and writes the bf16
//          row into a padded shared-memory tile (the pipeline stages are free once the accumulator is complete).
// Phase B: GroupNorm partials are reduced through shared memory (16 lanes per (sample, group), one atomic
//          pair each), and the tile is written out with fully coalesced 16-byte stores (+ coalesced residual
//          reads).  The first version did both per row with warp shuffles and 16-byte scattered stores and took
//          as long as the main loop (profiles/README.md).
template <bool LN_FOLD>
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
    // channel LayerNorm of the input folded into this GEMM (dd_conv_tc_ln): per-row {1 / (std + eps), -mean / (std + eps)}
    float* s_wsum = s_bias + 128;
    float ln_a = 1.f, ln_b = 0.f;
    if (LN_FOLD) {
        if (et < bn) s_wsum[et] = p.ln_wsum[cbase + et];
        if (valid) {
            const int ww = r & (p.tw - 1), hh = (r >> p.tw_sh) & (p.th - 1), n = n0 + (r >> (p.tw_sh + p.th_sh));
            const float2* lp = reinterpret_cast<const float2*>(p.ln_in) + (((int64_t)n * p.H + (h0 + hh)) * p.W + (w0 + ww)) * p.ln_in_parts;
            float su = 0.f, sq = 0.f;
            for (int i = 0; i < p.ln_in_parts; ++i) { const float2 v = __ldg(lp + i); su += v.x; sq += v.y; }
            const float mean = su * p.ln_inv_c;
            ln_a = 1.f / (sqrtf(fmaxf(sq * p.ln_inv_c - mean * mean, 0.f)) + p.ln_eps);
            ln_b = -mean * ln_a;
        }
    }
    epi_bar();
    mbar_wait(tmem_full_bar, 0);
    if (et == 0) tstamp(p, 5);
    tc_fence_after();
    tc_fold_partials(p, tmem_base, warp);
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
                const int c = c32 * 32 + g8 * 8 + j;
                v[j] = LN_FOLD ? fmaf(__uint_as_float(acc[g8 * 8 + j]), ln_a, fmaf(ln_b, s_wsum[c], s_bias[c]))
                               : __uint_as_float(acc[g8 * 8 + j]) + s_bias[c];
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

// ---- fused GroupNorm + Mish epilogue (dd_conv_tc_gn; bf16 NHWC output, bn = 64 or 128) ----------------------------------
// blocks.py:79-84 and :108-115 behind the convolution that feeds them: y = mish(GN(acc + bias)) [+ time bias] [+ residual].
// The statistics of a (sample, group) span every pixel of the image, so the tiles of one image run as ONE thread-block
// cluster (1, 2, 4 or 8 CTAs; a tile of a small map holds whole images and needs no peer):
//   part 1  drain the accumulator from TMEM (+ bias) into an fp32 staging tile (the pipeline stages are free by then) and
//           reduce this tile's {sum, sum of squares} per (sample of the tile, group of the N tile) into s_stat;
//   --      cluster barrier: every tile of the image has published its partial sums;
//   part 2  add the peers' partials through distributed shared memory (same order in every CTA: bit-identical mean / rstd),
//           then normalise, activate, add the time bias / residual and write bf16 NHWC with coalesced 16-byte stores.
// The normalisation reads the fp32 accumulator, not a bf16 round trip through memory: one rounding less per convolution than
// the separate dd_gn_mish launch this replaces, no statistics atomics, no arena memset.
// Staging layout: two fp32 planes [128 rows][bn/2 floats + 16 B]: float4 j of a row lives in plane (j & 1) at position j >> 1,
// so both the row-per-thread writes of part 1 and the 8-channels-per-thread reads of part 2 are bank-conflict free;
// then s_part [2 * bn/8][128], s_stat [nout][2], s_mr [nout][2] ({mean, rstd}).
struct GnLayout { int pp, plane, off_part, off_stat, off_mr; };
__device__ __forceinline__ GnLayout gn_layout(int bn) {
    GnLayout L;
    L.pp = bn * 2 + 16; L.plane = TC_BM * L.pp; L.off_part = 2 * L.plane; L.off_stat = L.off_part + bn * TC_BM; L.off_mr = L.off_stat + 1024;
    return L;
}
struct GnRegs { float ga[8], be[8]; };

// gain / offset of the eight channels thread `t` writes in part 2 (t = 0 .. nthr-1 over ALL warps of the CTA: the producer and
// MMA warps have nothing left to do once the accumulator is complete, and part 2 reads shared memory, not TMEM)
__device__ __forceinline__ GnRegs tc_gn_load_regs(const TcParams& p, int n_tile, int t) {
    GnRegs g;
    const int bn = p.bn, cbase = n_tile * bn, c8 = (t & ((bn >> 3) - 1)) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + cbase + c8)), g1 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + cbase + c8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + cbase + c8)), b1 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + cbase + c8) + 1);
    g.ga[0] = g0.x; g.ga[1] = g0.y; g.ga[2] = g0.z; g.ga[3] = g0.w; g.ga[4] = g1.x; g.ga[5] = g1.y; g.ga[6] = g1.z; g.ga[7] = g1.w;
    g.be[0] = b0.x; g.be[1] = b0.y; g.be[2] = b0.z; g.be[3] = b0.w; g.be[4] = b1.x; g.be[5] = b1.y; g.be[6] = b1.z; g.be[7] = b1.w;
    return g;
}

__device__ __forceinline__ GnRegs tc_epi_gn_part1(const TcParams& p, uint32_t tmem_base, uint32_t tmem_full_bar, float* s_bias,
                                                  uint8_t* stage, int n_tile, int n0, int warp, int lane) {
    const int q = warp & 3, r = q * 32 + lane, et = threadIdx.x - 64;
    const int bn = p.bn, cbase = n_tile * bn;
    const GnLayout L = gn_layout(bn);
    const int rps_sh = p.tw_sh + p.th_sh, rps = 1 << rps_sh;
    if (et < bn) s_bias[et] = p.bias ? p.bias[cbase + et] : 0.f;
    // the eight channels this thread writes in part 2: their gain / offset travel in registers (fetched under the main loop)
    const GnRegs g = tc_gn_load_regs(p, n_tile, threadIdx.x);
    const bool valid = (n0 + (r >> rps_sh)) < p.B && r < p.rows_valid;
    epi_bar();
    mbar_wait(tmem_full_bar, 0);
    if (et == 0) tstamp(p, 5);
    tc_fence_after();
    tc_fold_partials(p, tmem_base, warp);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float s8[16], q8[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { s8[k] = 0.f; q8[k] = 0.f; }
    uint32_t a0[32], a1[32];
    auto process = [&](const int c32, uint32_t (&acc)[32]) {
        uint8_t* dst = stage + r * L.pp + c32 * 64;
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
            *reinterpret_cast<float4*>(dst + g8 * 16) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + L.plane + g8 * 16) = make_float4(v[4], v[5], v[6], v[7]);
        }
    };
    tmem_ld32_issue(trow, a0);
    tmem_ld_wait();
    tmem_ld32_issue(trow + 32u, a1);
    process(0, a0);
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
    if (et == 0) tstamp(p, 10);
    // {sum, sum of squares} per (sample of the tile, group of the N tile): 8-channel slices -> groups inside the thread, rows of
    // a sample across lanes with shuffles (a sample's rows are consecutive threads), segments of 32 rows through s_seg
    {
        const int cpg_sh = p.cpg_shift, spg_sh = cpg_sh - 3;                // log2(8-channel slices per group)
        const int ng = bn >> cpg_sh;                                        // groups in this tile (<= 16)
        const int seg_sh = rps_sh < 5 ? rps_sh : 5, seg = 1 << seg_sh;      // lanes that share a sample
        float* s_seg = reinterpret_cast<float*>(stage + L.off_part);        // [128 / seg segments][ng][2]
        // fold the 8-channel slices of a group (1, 2 or 4 of them) with static register indices
        if (spg_sh >= 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s8[i] = s8[2 * i] + s8[2 * i + 1]; q8[i] = q8[2 * i] + q8[2 * i + 1]; }
        }
        if (spg_sh >= 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { s8[i] = s8[2 * i] + s8[2 * i + 1]; q8[i] = q8[2 * i] + q8[2 * i + 1]; }
        }
#pragma unroll
        for (int gi = 0; gi < 16; ++gi) {
            if (gi < ng) {
                float sa = s8[gi], qa = q8[gi];
                for (int d = seg >> 1; d > 0; d >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, d);
                    qa += __shfl_xor_sync(0xffffffffu, qa, d);
                }
                if ((lane & (seg - 1)) == 0) {
                    float* dst = s_seg + (((r >> seg_sh) * ng + gi) << 1);
                    dst[0] = sa; dst[1] = qa;
                }
            }
        }
        epi_bar();
        float* s_stat = reinterpret_cast<float*>(stage + L.off_stat);
        const int ng_sh = (31 - __clz(bn)) - cpg_sh;
        const int nout = ng << (7 - rps_sh);                                // (samples of the tile) x groups
        const int sps = 1 << (rps_sh - seg_sh);                             // segments per sample
        for (int o = et; o < nout; o += 128) {
            const int gl = o & (ng - 1), sl = o >> ng_sh;
            float sa = 0.f, qa = 0.f;
            for (int i = 0; i < sps; ++i) {
                const float* src = s_seg + (((sl * sps + i) * ng + gl) << 1);
                sa += src[0]; qa += src[1];
            }
            s_stat[2 * o] = sa; s_stat[2 * o + 1] = qa;
        }
    }
    if (et == 0) tstamp(p, 11);
    return g;
}

// part 2 runs on every thread of the CTA (t = threadIdx.x, nthr = blockDim.x)
__device__ __forceinline__ void tc_epi_gn_part2(const TcParams& p, const GnRegs& g, uint8_t* stage, int n_tile, int w0, int h0, int n0) {
    const int t = threadIdx.x, nthr = blockDim.x, lane = threadIdx.x & 31;
    const int bn = p.bn, cbase = n_tile * bn;
    const GnLayout L = gn_layout(bn);
    const int rps_sh = p.tw_sh + p.th_sh;
    const int cpg_sh = p.cpg_shift, ng_sh = (31 - __clz(bn)) - cpg_sh;
    const int nout = (1 << ng_sh) * (TC_BM >> rps_sh);
    float* s_stat = reinterpret_cast<float*>(stage + L.off_stat);
    float* s_mr = reinterpret_cast<float*>(stage + L.off_mr);
    for (int o = t; o < nout; o += nthr) {
        float su = 0.f, sq = 0.f;
        if (p.gn_cluster > 1) {
            // all loads first (each is a ~200-cycle round trip to a peer SM), then the sums, in rank order in every CTA
            float vs[8], vq[8];
            const uint32_t a = smem_u32(s_stat + 2 * o);
#pragma unroll
            for (int rk = 0; rk < 8; ++rk) {
                vs[rk] = 0.f; vq[rk] = 0.f;
                if (rk < p.gn_cluster) {
                    const uint32_t ra = mapa_u32(a, (uint32_t)rk);
                    vs[rk] = ld_dsmem_f32(ra); vq[rk] = ld_dsmem_f32(ra + 4u);
                }
            }
#pragma unroll
            for (int rk = 0; rk < 8; ++rk) { su += vs[rk]; sq += vq[rk]; }
        } else { su = s_stat[2 * o]; sq = s_stat[2 * o + 1]; }
        const float mean = su * p.gn_inv_n;
        s_mr[2 * o] = mean;
        s_mr[2 * o + 1] = rsqrtf(fmaxf(sq * p.gn_inv_n - mean * mean, 0.f) + p.gn_eps);
    }
    __syncthreads();
    if (t == 64) tstamp(p, 12);
    // write-out: 16 bytes (8 channels) per thread, consecutive threads walk along a row
    const int ppr_sh = 31 - __clz(bn >> 3);
    const int pc = t & ((1 << ppr_sh) - 1), c8 = pc * 8, gl = c8 >> cpg_sh;
    const int row_step = nthr >> ppr_sh;
    const int Hh = p.H, Ww = p.W, Cout = p.Cout, rows_valid = p.rows_valid, Bn = p.B;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    const __nv_bfloat16* resp = p.residual;
    const unsigned gmask = (bn == 128) ? (0xFFFFu << (lane & 16)) : (0xFFu << (lane & 24));      // the lanes that share this thread's row
    const int ntiles = Cout / bn;
    float a8[8], b8[8], t8[8];
    int cur_sl = -1;
    for (int row0 = t >> ppr_sh; row0 < TC_BM; row0 += 2 * row_step) {
        float4 lo[2], hi[2];
        uint4 rv[2];
        int64_t pix[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int row = row0 + u * row_step;
            const int n = n0 + (row >> rps_sh);
            ok[u] = row < TC_BM && row < rows_valid && n < Bn;
            const int ww = row & (p.tw - 1), hh = (row >> p.tw_sh) & (p.th - 1);
            pix[u] = ((int64_t)n * Hh + (h0 + hh)) * Ww + (w0 + ww);
            if (ok[u]) {
                lo[u] = *reinterpret_cast<const float4*>(stage + row * L.pp + pc * 16);
                hi[u] = *reinterpret_cast<const float4*>(stage + L.plane + row * L.pp + pc * 16);
                if (resp) rv[u] = *reinterpret_cast<const uint4*>(resp + pix[u] * Cout + cbase + c8);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!ok[u]) continue;
            const int row = row0 + u * row_step, sl = row >> rps_sh;
            if (sl != cur_sl) {
                cur_sl = sl;
                const float mean = s_mr[2 * ((sl << ng_sh) + gl)], rstd = s_mr[2 * ((sl << ng_sh) + gl) + 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) { a8[j] = rstd * g.ga[j]; b8[j] = fmaf(-mean, a8[j], g.be[j]); t8[j] = 0.f; }
                if (p.tbias) {
                    const int n = n0 + sl;
                    const int trow_i = p.trow ? p.trow[(int64_t)n * p.trow_stride] : n;
                    const float4* tp = reinterpret_cast<const float4*>(p.tbias + (int64_t)trow_i * p.tb_stride + cbase + c8);
                    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
                    t8[0] = t0.x; t8[1] = t0.y; t8[2] = t0.z; t8[3] = t0.w; t8[4] = t1.x; t8[5] = t1.y; t8[6] = t1.z; t8[7] = t1.w;
                }
            }
            float v[8] = {lo[u].x, lo[u].y, lo[u].z, lo[u].w, hi[u].x, hi[u].y, hi[u].z, hi[u].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = mish_fast(fmaf(v[j], a8[j], b8[j])) + t8[j];
            if (resp) {
                const uint32_t w4[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { v[2 * j] += __uint_as_float(w4[j] << 16); v[2 * j + 1] += __uint_as_float(w4[j] & 0xffff0000u); }
            }
            uint4 ov;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            *reinterpret_cast<uint4*>(outp + pix[u] * Cout + cbase + c8) = ov;
            if (p.ln_part) {
                // channel-LayerNorm statistics of the row as the next kernel will read it (bf16-rounded): blocks.py:57-60
                const uint32_t w4[4] = {ov.x, ov.y, ov.z, ov.w};
                float ls = 0.f, lq = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x0 = __uint_as_float(w4[j] << 16), x1 = __uint_as_float(w4[j] & 0xffff0000u);
                    ls += x0 + x1; lq = fmaf(x0, x0, fmaf(x1, x1, lq));
                }
                for (int d = (1 << ppr_sh) >> 1; d > 0; d >>= 1) {
                    ls += __shfl_xor_sync(gmask, ls, d);
                    lq += __shfl_xor_sync(gmask, lq, d);
                }
                if (pc == 0) *reinterpret_cast<float2*>(p.ln_part + (pix[u] * ntiles + n_tile) * 2) = make_float2(ls, lq);
            }
        }
    }
    if (t == 64) tstamp(p, 6);
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
    tc_fold_partials(p, tmem_base, warp);
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
// MW = 2 (the one-CTA-per-SM forms): a second MMA issuer (warp 7) takes half of the K = 16 slices of every stage into its own
// accumulator -- ONE issuing warp needs ~430 clk of instruction latency per barrier wait + four MMAs, whatever the operand rings do
// (433 clk per k-block measured on the 8x8 layers against 192 clk of tensor-pipe work; profiles/README.md, round 2 passes p - z).
template <int TC_STAGES, int BROWS, int KCH, int EPI, int MW = 1>      // EPI: 0 staged bf16, 1 legacy (fp32 NCHW / narrow), 2 split-K partial
__global__ void __launch_bounds__(TC_THREADS + 32 * (MW - 1), (TC_STAGES * KCH <= 3 ? 2 : 1)) conv_tc_kernel(const __grid_constant__ TcParams p) {
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
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), MW); }
        mbar_init(tmem_full_bar, MW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(TC_TMEM_COLS * MW) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (threadIdx.x == 0) tstamp(p, 1);
    // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail; the weight producer (warp 6)
    // runs on: packed weights do not depend on the preceding launch (per-sample matrices do: it waits below)
    if (warp != 6) pdl_sync();
    if (threadIdx.x == 0) tstamp(p, 2);
    GnRegs gnr;

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
        if (wps) pdl_sync();
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
    } else if (warp == 1 || (MW == 2 && warp == 7)) {
        // ===== MMA issuer(s): whole warp loops, one elected lane issues tcgen05.mma / commit =====
        // instruction descriptor: D=f32, A=B=bf16, both K-major, N=bn, M=128.  Issuer mw takes the K = 16 slices
        // [mw * KK_N, (mw + 1) * KK_N) of every stage into accumulator mw (every issuer waits on every stage: all barrier phases seen).
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        constexpr int KK_N = (TC_BK / 16) / MW;
        const int mw = warp == 1 ? 0 : 1;
        const uint32_t dcol = tmem_base + (uint32_t)(mw * TC_TMEM_COLS);
        uint32_t sA = base;
        int st = 0;
        uint32_t par = 0;
        for (int i = 0; i < num_st; ++i) {
            mbar_wait(full_bar(st), par);
            if (mw == 0 && i == 0 && lane == 0) tstamp(p, 3);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < KCH; ++j) {
                    const uint64_t ad = umma_desc(sA + j * TC_A_BYTES);
                    const uint64_t bd = umma_desc(sA + KCH * TC_A_BYTES + j * B_SLOT);
#pragma unroll
                    for (int k = 0; k < KK_N; ++k)
                        umma_f16(dcol, ad + (uint64_t)(2 * (mw * KK_N + k)), bd + (uint64_t)(2 * (mw * KK_N + k)), idesc, (i | j | k) ? 1u : 0u);
                }
                umma_commit(empty_bar(st));          // frees this smem stage when the MMAs (of both issuers) retire
            }
            __syncwarp();
            sA += TC_STAGE_BYTES;
            if (++st == TC_STAGES) { st = 0; par ^= 1u; sA = base; }
        }
        if (mw == 0 && lane == 0) tstamp(p, 4);
        if (elect_one()) umma_commit(tmem_full_bar);             // this issuer's accumulator complete
        __syncwarp();
    } else if (warp >= 2 && warp < 6) {
        if constexpr (EPI == 1)         // fp32 NCHW output / tiles narrower than 32 channels (the final 1x1 conv)
            tc_epilogue(p, tmem_base, tmem_full_bar, s_bias, m_tile, n_tile, phase, w0, h0, n0, warp, lane);
        else if constexpr (EPI == 2)
            tc_epilogue_partial(p, tmem_base, tmem_full_bar, n_tile, split, w0, h0, n0, warp, lane);
        else if (p.gn_fuse)
            gnr = tc_epi_gn_part1(p, tmem_base, tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)), n_tile, n0, warp, lane);
        else if (p.ln_in)
            tc_epilogue_staged<true>(p, tmem_base, tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)), n_tile, phase, w0, h0, n0,
                                     warp, lane);
        else
            tc_epilogue_staged<false>(p, tmem_base, tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)), n_tile, phase, w0, h0, n0,
                                      warp, lane);
    }
    const bool gn_cl = EPI == 0 && p.gn_fuse && p.gn_cluster > 1;
    if (EPI == 0 && p.gn_fuse) {
        if (warp < 2 || warp >= 6) gnr = tc_gn_load_regs(p, n_tile, threadIdx.x);      // producer / MMA warps join the write-out
        // every tile of the image has published its partial sums (s_stat)
        if (gn_cl) cluster_sync_all();
        else __syncthreads();
        tc_epi_gn_part2(p, gnr, smem_raw + (base - smem_u32(smem_raw)), n_tile, w0, h0, n0);
    }
    tc_fence_before();
    if (gn_cl) cluster_sync_all();          // peers may still be reading this CTA's partial sums
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS * MW) : "memory");
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
    GnRegs gnr;

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
        if (p.gn_fuse)
            gnr = tc_epi_gn_part1(p, tmem_base, tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)), n_tile, n0[0], warp, lane);
        else
#pragma unroll
        for (int s = 0; s < NT; ++s)
            tc_epilogue_staged<false>(p, tmem_base + (uint32_t)(s * TC_TMEM_COLS), tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)), n_tile, 0,
                                      w0[s], h0[s], n0[s], warp, lane);
    }
    const bool gn_cl = p.gn_fuse && p.gn_cluster > 1;
    if (p.gn_fuse) {
        if (warp < 2 || warp >= 6) gnr = tc_gn_load_regs(p, n_tile, threadIdx.x);
        if (gn_cl) cluster_sync_all();
        else __syncthreads();
        tc_epi_gn_part2(p, gnr, smem_raw + (base - smem_u32(smem_raw)), n_tile, w0[0], h0[0], n0[0]);
    }
    tc_fence_before();
    if (gn_cl) cluster_sync_all();
    else __syncthreads();
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
            tc_epilogue_staged<false>(p, tmem_base + (uint32_t)(s * TC_TMEM_COLS), tmem_full_bar, s_bias, smem_raw + (base - smem_u32(smem_raw)),
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

}  // namespace dd

using namespace dd;

long long* dd::g_tc_dbg = nullptr;
extern "C" int dd_debug_set_timeline(long long* buf) { g_tc_dbg = buf; return DD_OK; }

// Debug / tuning aid: how many thread-block clusters of `cluster` halo-kernel CTAs the device can keep resident
// (cudaOccupancyMaxActiveClusters); -1 on error.
extern "C" int dd_debug_max_clusters(int cluster) {
    cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster * 64); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = HALO_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, conv_tc_halo_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

extern "C" int dd_zero(void* ptr, int64_t bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dd_zero: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
    return DD_OK;
}

// Split-K policy: 3x3 stride-1 convs on maps of at most 4x4 pixels whose 64-wide tiles would occupy less than a
// third of the SMs.  Each CTA streams its whole K range through one SM's L2 port (~64 B/clk), so 32 CTAs take ~7 us
// for 27 MB of operands while 116 SMs idle; S splits cut that stream S-fold.  Returns S (>= 2) or 1.
extern "C" int dd_conv_tc_splits(int kind, int B, int H, int W, int Cin, int Cout) {
    if (kind != DD_TC_CONV3x3 || H * W > 16 || !is_pow2(H) || !is_pow2(W) || Cout < 128 || Cout % 64 || Cin % 64 || getenv("DD_NO_SPLITK")) return 1;
    const int rows = B * H * W, tiles = (rows + 127) / 128 * (Cout / 64), num_kb = 9 * (Cin / 64);
    if (tiles * 3 > dd::num_sms()) return 1;
    int S = dd::num_sms() / tiles;
    if (S > 8) S = 8;
    while (S > 1 && (num_kb % S != 0 || num_kb / S < 4)) --S;
    return S;
}

namespace {
struct GnFuse {                 // arguments of the fused GroupNorm + Mish epilogue (dd_conv_tc_gn); gamma == nullptr: not fused
    const float* gamma = nullptr;
    const float* beta = nullptr;
    float eps = 1e-5f;
    const float* tbias = nullptr;
    int tb_stride = 0;
    const int32_t* trow = nullptr;
    int trow_stride = 0;
    float* ln_part = nullptr;
    float* ws = nullptr;        // zeroed workspace of dd_conv_tc_gn_ws_floats() floats: enables the persistent kernel (statistics + arrival counters)
    // channel LayerNorm folded into a 1x1 convolution (dd_conv_tc_ln); ln_in == nullptr: none
    const float* ln_in = nullptr;
    int ln_in_parts = 0;
    const float* ln_wsum = nullptr;
    float ln_eps = 1e-5f;
};

// Tile geometry of a launch (shared by dd_conv_tc and the dd_conv_tc_gn_cluster query).
// Maps whose sides are not powers of two (28x28 -> 14x14 -> 7x7 of the MNIST-shaped DDPM) are tiled as
// if padded to the next power of two: the TMA boxes hang over the right / bottom edge (out-of-bounds elements arrive as zeros,
// which is also the convolution's padding) and the epilogue skips the rows of a tile that lie outside the image.
struct TcGeom { int tw, th, tn, bn, tiles_w, tiles_h, tiles_n; bool halo, ragged; };
static int next_pow2(int v) { int r = 1; while (r < v) r <<= 1; return r; }
TcGeom tc_geometry(int kind, int B, int H, int W, int Cout, int flags, int out_nchw_f32) {
    TcGeom g;
    static const bool halo_off = getenv("DD_NO_HALO") != nullptr;
    const bool wps = (flags & DD_TC_W_PER_SAMPLE) != 0;
    g.ragged = !is_pow2(H) || !is_pow2(W);
    const int Wp = next_pow2(W), Hp = next_pow2(H);
    g.halo = !g.ragged && !halo_off && kind == DD_TC_CONV3x3 && H >= HALO_TH && W >= HALO_TW && Cout >= 64 && !out_nchw_f32;
    g.tw = Wp < 128 ? Wp : 128;
    g.th = (128 / g.tw) < Hp ? (128 / g.tw) : Hp;
    if (g.halo) { g.tw = HALO_TW; g.th = HALO_TH; }
    g.tn = wps ? 1 : 128 / (g.tw * g.th);      // per-sample weights: one image per tile (rows beyond it are ignored)
    g.tiles_w = (W + g.tw - 1) / g.tw; g.tiles_h = (H + g.th - 1) / g.th;
    g.tiles_n = (B + g.tn - 1) / g.tn;
    g.bn = Cout >= 128 ? 128 : Cout;
    // low-resolution layers (at most half a wave of 128-wide tiles): halve the N tile to double the CTA count
    const int tiles128 = g.tiles_w * g.tiles_h * g.tiles_n * ((Cout + 127) / 128);
    if (tiles128 * 2 <= num_sms() && Cout % 64 == 0 && Cout >= 128) g.bn = 64;
    if (flags & DD_TC_SPLITK) g.bn = 64;
    return g;
}
}  // namespace

// Cluster size the fused GroupNorm epilogue would use for this layer: 1 (a tile holds whole images), 2, 4 or 8 (the tiles of
// one image form a thread-block cluster), 0 when the layer cannot take the fused epilogue (dd_conv_tc + dd_gn_mish then).
extern "C" int dd_conv_tc_gn_cluster(int kind, int B, int H, int W, int Cout, int G) {
    if (kind == DD_TC_UPT || kind < 0 || kind > 3 || !is_pow2(H) || !is_pow2(W) || B <= 0 || G <= 0 || Cout % G) return 0;
    if (getenv("DD_NO_GN_FUSE")) return 0;
    const TcGeom g = tc_geometry(kind, B, H, W, Cout, 0, 0);
    const int cpg = Cout / G;
    if ((g.bn != 64 && g.bn != 128) || Cout % g.bn || !is_pow2(cpg) || cpg < 8 || g.bn % cpg) return 0;
    const int tpi = g.tiles_w * g.tiles_h;                          // tiles per image
    if (tpi == 1) {
        const int nsamp = g.tn, nout = nsamp * (g.bn / cpg);
        return (g.tw * g.th >= 16 && nout <= 128) ? 1 : 0;          // s_stat / s_mr hold 128 (sample, group) pairs
    }
    return (tpi == 2 || tpi == 4 || tpi == 8) ? tpi : 0;
}

// N tile (output channels per CTA) dd_conv_tc / dd_conv_tc_gn choose for a layer: the `parts` dimension of ln_part is Cout / it.
extern "C" int dd_conv_tc_tile_n(int kind, int B, int H, int W, int Cout, int flags) {
    if (kind < 0 || kind > 3 || H <= 0 || W <= 0 || B <= 0 || Cout <= 0) return 0;
    return tc_geometry(kind, B, H, W, Cout, flags, 0).bn;
}

// Workspace of dd_conv_tc_gn for this layer in floats (0: none needed): one 16-byte packet {sum, flag, sum of squares, flag} per
// (image, 128-channel tile, pixel tile, group of the channel tile); must be ALL ZERO when the launch starts.
extern "C" int64_t dd_conv_tc_gn_ws_floats(int kind, int B, int H, int W, int Cout, int G) {
    if (kind != DD_TC_CONV3x3 || !is_pow2(H) || !is_pow2(W) || B <= 0 || G <= 0 || Cout % G || getenv("DD_NO_GN_FUSE")) return 0;
    const TcGeom g = tc_geometry(kind, B, H, W, Cout, 0, 0);
    if (!(g.halo && halo_persist_ok(kind, H, W, Cout, G))) return 0;
    const int64_t tpi = (int64_t)(H / HALO_TH) * (W / HALO_TW), groups_per_tile = 128 / (Cout / G);
    return (int64_t)B * (Cout / 128) * tpi * groups_per_tile * 4;
}

// `parts` dimension of dd_conv_tc_gn's ln_part output for this layer
extern "C" int dd_conv_tc_gn_ln_parts(int kind, int B, int H, int W, int Cout, int G, int persistent) {
    if (persistent) return dd_conv_tc_gn_ws_floats(kind, B, H, W, Cout, G) > 0 ? halo_persist_col_split() * (Cout / 128) : 0;
    if (dd_conv_tc_gn_cluster(kind, B, H, W, Cout, G) <= 0) return 0;
    return Cout / tc_geometry(kind, B, H, W, Cout, 0, 0).bn;
}

static int conv_tc_impl(int kind, const void* x, int x_pitch, const void* x2, int C1, int C2, const void* wp, int w_rows,
                        const float* bias, const void* residual, void* y, int out_nchw_f32, int cout_valid,
                        float* gn_stats, int G, int B, int H, int W, int Cout, int flags,
                        float* splitk_ws, int64_t splitk_ws_floats, int32_t* splitk_cnt, int splitk_cnt_n, const GnFuse& gf, void* stream) {
    DD_REQUIRE(kind >= 0 && kind <= 3, "conv_tc: bad kind %d", kind);
    DD_REQUIRE(C1 > 0 && C1 % 64 == 0 && C2 >= 0 && C2 % 64 == 0, "conv_tc: channel counts (%d,%d) must be multiples of 64", C1, C2);
    DD_REQUIRE((C2 == 0) == (x2 == nullptr), "conv_tc: x2/C2 mismatch");
    DD_REQUIRE(H > 0 && W > 0 && B > 0, "conv_tc: bad sizes B=%d H=%d W=%d", B, H, W);
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
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<6, 128, 1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(6, 128, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<8, 64, 1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(8, 64, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<3, 128, 2, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(3, 128, 2));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<4, 64, 2, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(4, 64, 2));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<8, 64, 1, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(8, 64, 1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
        attr_done = true;
    }
    const bool wps = (flags & DD_TC_W_PER_SAMPLE) != 0;
    DD_REQUIRE(!wps || kind == DD_TC_CONV1x1, "conv_tc: per-sample weights are for 1x1 convs only");

    TcParams p;
    memset(&p, 0, sizeof(p));
    // tile geometry over the GEMM pixel grid
    const TcGeom geo = tc_geometry(kind, B, H, W, Cout, flags, out_nchw_f32);
    const bool halo = geo.halo, ragged = geo.ragged;
    DD_REQUIRE(!ragged || (gf.gamma == nullptr && gf.ln_in == nullptr && !(flags & (DD_TC_SPLITK | DD_TC_PAIR))),
               "conv_tc: maps that are not powers of two (%dx%d) take the plain epilogue only (no fused GroupNorm, LayerNorm fold, split-K)", H, W);
    p.tw = geo.tw; p.th = geo.th; p.tn = geo.tn;
    p.rows_valid = p.tw * p.th * p.tn;
    p.tw_sh = 0; while ((1 << p.tw_sh) < p.tw) ++p.tw_sh;
    p.th_sh = 0; while ((1 << p.th_sh) < p.th) ++p.th_sh;
    p.w_per_sample = wps ? 1 : 0;
    p.tiles_w = geo.tiles_w; p.tiles_h = geo.tiles_h;
    const int tiles_n = geo.tiles_n;
    p.B = B; p.H = H; p.W = W;
    p.chunks0 = C1 / 64; p.chunks1 = C2 / 64;
    p.Cout = Cout; p.cout_valid = out_nchw_f32 ? cout_valid : Cout;
    p.bn = geo.bn;
    DD_REQUIRE(Cout % p.bn == 0 && (p.bn == 16 || p.bn == 32 || p.bn == 64 || p.bn == 128), "conv_tc: unsupported Cout=%d", Cout);
    p.out = y; p.bias = bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.gn_stats = gn_stats; p.G = G; p.out_nchw_f32 = out_nchw_f32; p.out_mul = 1;
    const bool fuse = gf.gamma != nullptr;
    bool persist = false;
    if (fuse) {
        persist = gf.ws != nullptr && dd_conv_tc_gn_ws_floats(kind, B, H, W, Cout, G) > 0;
        const int cl = persist ? 1 : dd_conv_tc_gn_cluster(kind, B, H, W, Cout, G);
        DD_REQUIRE(cl > 0, "conv_tc_gn: this layer cannot take the fused GroupNorm epilogue (ask dd_conv_tc_gn_cluster / dd_conv_tc_gn_ws_floats first)");
        DD_REQUIRE(gf.beta != nullptr && gn_stats == nullptr && !out_nchw_f32 && !(flags & (DD_TC_SPLITK | DD_TC_W_PER_SAMPLE | DD_TC_PAIR)),
                   "conv_tc_gn: bad argument combination");
        DD_REQUIRE(gf.tbias == nullptr || gf.tb_stride % 4 == 0, "conv_tc_gn: time-bias stride must be a multiple of 4 floats");
        p.gn_fuse = 1; p.gn_cluster = cl; p.gn_eps = gf.eps; p.gn_inv_n = 1.f / ((float)H * (float)W * (float)(Cout / G));
        p.gn_gamma = gf.gamma; p.gn_beta = gf.beta; p.tbias = gf.tbias; p.tb_stride = gf.tb_stride; p.trow = gf.trow;
        p.trow_stride = gf.trow_stride; p.ln_part = gf.ln_part;
        if (persist) p.gn_stats = gf.ws;
    }
    if (gf.ln_in) {
        DD_REQUIRE(kind == DD_TC_CONV1x1 && C2 == 0 && !out_nchw_f32 && !fuse && gn_stats == nullptr && !(flags & (DD_TC_SPLITK | DD_TC_W_PER_SAMPLE)),
                   "conv_tc_ln: the LayerNorm fold is for plain 1x1 convolutions");
        DD_REQUIRE(gf.ln_wsum != nullptr && gf.ln_in_parts >= 1 && p.bn >= 32, "conv_tc_ln: bad arguments");
        p.ln_in = gf.ln_in; p.ln_in_parts = gf.ln_in_parts; p.ln_wsum = gf.ln_wsum; p.ln_eps = gf.ln_eps; p.ln_inv_c = 1.f / (float)C1;
    }
    if (gn_stats || fuse) {
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
    const int cl_x = fuse ? p.gn_cluster : 1;
    if (persist) {
        rc = make_w_map(&p.tmB, wp, K, w_rows, 128, 0);
        if (rc) return rc;
        return launch_halo_persist(p, st);
    }
    if (!ragged && !fuse && !out_nchw_f32 && p.splits == 1 && gemm_persist_ok(kind, B, H, W, C1, C2, Cout, G, gn_stats != nullptr, wps)) {
        rc = make_w_map(&p.tmB, wp, K, w_rows, 128, wps ? B : 0);
        if (rc) return rc;
        return launch_gemm_persist(p, x, x_pitch, C1, st);
    }
    static const bool one_issuer = getenv("DD_TC_ONE_ISSUER") != nullptr;
    const dim3 blk2(TC_THREADS + 32);
    if (!one_issuer && !(out_nchw_f32 || p.bn < 32) && !cta_pair && !halo && !ragged && (p.splits > 1 || ctas <= num_sms())) p.mma_parts = 2;
    if (p.splits > 1 && p.mma_parts == 2)
        launch_pdl(conv_tc_kernel<8, 64, 1, 2, 2>, dim3(grid), blk2, tc_smem_bytes(8, 64, 1), st, p);
    else if (p.splits > 1)
        launch_pdl(conv_tc_kernel<8, 64, 1, 2>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(8, 64, 1), st, p);
    else if (out_nchw_f32 || p.bn < 32 || ragged)        // per-row epilogue: the only one that masks rows outside the image
        launch_pdl(conv_tc_kernel<3, 128, 1, 1>, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(3, 128, 1), st, p);
    else if (cta_pair)
        launch_pair_pdl(conv_tc_halo2_kernel, dim3(grid.x, Cout / (p.bn * p.pair_nt), 1), dim3(TC_THREADS), HALO_SMEM, st, p);
    else if (halo)
        launch_cluster_pdl(conv_tc_halo_kernel, cl_x, dim3(grid), dim3(TC_THREADS), HALO_SMEM, st, p);
    else if (ctas > num_sms())      // more than one wave: two CTAs per SM so epilogues overlap main loops
        launch_cluster_pdl(conv_tc_kernel<3, 128, 1, 0>, cl_x, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(3, 128, 1), st, p);
    else if (p.mma_parts == 2) {           // at most one wave, one CTA per SM: two MMA issuers
        if (p.bn <= 64 && pair) launch_cluster_pdl(conv_tc_kernel<4, 64, 2, 0, 2>, cl_x, dim3(grid), blk2, tc_smem_bytes(4, 64, 2), st, p);
        else if (p.bn <= 64) launch_cluster_pdl(conv_tc_kernel<8, 64, 1, 0, 2>, cl_x, dim3(grid), blk2, tc_smem_bytes(8, 64, 1), st, p);
        else if (pair) launch_cluster_pdl(conv_tc_kernel<3, 128, 2, 0, 2>, cl_x, dim3(grid), blk2, tc_smem_bytes(3, 128, 2), st, p);
        else launch_cluster_pdl(conv_tc_kernel<6, 128, 1, 0, 2>, cl_x, dim3(grid), blk2, tc_smem_bytes(6, 128, 1), st, p);
    }
    else if (p.bn <= 64 && pair)
        launch_cluster_pdl(conv_tc_kernel<4, 64, 2, 0>, cl_x, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(4, 64, 2), st, p);
    else if (p.bn <= 64)
        launch_cluster_pdl(conv_tc_kernel<8, 64, 1, 0>, cl_x, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(8, 64, 1), st, p);
    else if (pair)
        launch_cluster_pdl(conv_tc_kernel<3, 128, 2, 0>, cl_x, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(3, 128, 2), st, p);
    else
        launch_cluster_pdl(conv_tc_kernel<6, 128, 1, 0>, cl_x, dim3(grid), dim3(TC_THREADS), tc_smem_bytes(6, 128, 1), st, p);
    return check_launch("conv_tc");
}

extern "C" int dd_conv_tc(int kind, const void* x, int x_pitch, const void* x2, int C1, int C2, const void* wp, int w_rows,
                          const float* bias, const void* residual, void* y, int out_nchw_f32, int cout_valid,
                          float* gn_stats, int G, int B, int H, int W, int Cout, int flags,
                          float* splitk_ws, int64_t splitk_ws_floats, int32_t* splitk_cnt, int splitk_cnt_n, void* stream) {
    return conv_tc_impl(kind, x, x_pitch, x2, C1, C2, wp, w_rows, bias, residual, y, out_nchw_f32, cout_valid, gn_stats, G, B, H, W, Cout,
                        flags, splitk_ws, splitk_ws_floats, splitk_cnt, splitk_cnt_n, GnFuse(), stream);
}

extern "C" int dd_conv_tc_ln(const void* x, int C, const void* wp, int w_rows, const float* bias, const float* wsum,
                             const float* ln_in, int ln_in_parts, float ln_eps, void* y, int B, int H, int W, int Cout, void* stream) {
    DD_REQUIRE(ln_in != nullptr && wsum != nullptr, "conv_tc_ln: statistics and weight row sums required");
    GnFuse gf;
    gf.ln_in = ln_in; gf.ln_in_parts = ln_in_parts; gf.ln_wsum = wsum; gf.ln_eps = ln_eps;
    return conv_tc_impl(DD_TC_CONV1x1, x, 0, nullptr, C, 0, wp, w_rows, bias, nullptr, y, 0, 0, nullptr, 0, B, H, W, Cout, 0, nullptr, 0,
                        nullptr, 0, gf, stream);
}

extern "C" int dd_conv_tc_gn(int kind, const void* x, int x_pitch, const void* x2, int C1, int C2, const void* wp, int w_rows,
                             const float* bias, void* y, int B, int H, int W, int Cout, int flags,
                             int G, float eps, const float* gamma, const float* beta,
                             const float* tbias, int tb_stride, const int32_t* trow, int trow_stride,
                             const void* residual, float* ln_part, float* ws, void* stream) {
    DD_REQUIRE(gamma != nullptr && beta != nullptr, "conv_tc_gn: gamma / beta required");
    GnFuse gf;
    gf.ws = ws;
    gf.gamma = gamma; gf.beta = beta; gf.eps = eps; gf.tbias = tbias; gf.tb_stride = tb_stride; gf.trow = trow; gf.trow_stride = trow_stride;
    gf.ln_part = ln_part;
    return conv_tc_impl(kind, x, x_pitch, x2, C1, C2, wp, w_rows, bias, residual, y, 0, 0, nullptr, G, B, H, W, Cout, flags, nullptr, 0,
                        nullptr, 0, gf, stream);
}
