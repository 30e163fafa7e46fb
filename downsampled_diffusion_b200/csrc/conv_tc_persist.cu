// Persistent form of the 3x3 stride-1 halo convolution with the GroupNorm + Mish (+ time bias, + residual) of
// models/unet/blocks.py:73-84, 105-115 fused behind it (dd_conv_tc_gn on maps of at least 16x8 pixels, Cout % 128 == 0).
//
// Why: with one tile per CTA (conv_tc_halo_kernel) the two CTAs of an SM run in lock step -- both in their main loop, then both
// in their epilogue -- so the tensor pipe idles during every epilogue and the activation math (2 MUFU per element) is exposed
// (profiles/README.md, round 2: 9.0 k clk main loop + 6.1 k clk epilogue per wave, 16 - 21 k clk with the fused normalisation).
// Here ONE CTA per SM walks the work items (pixel tile x 128-channel tile) of the layer:
//   warp 0      TMA producer: one (18 x 10)-pixel halo box per 64-channel chunk, one 128 x 64 weight tile per (chunk, tap);
//               the rings (3 halos, 8 weight tiles = 198 KB in flight) run across item boundaries;
//   warp 1      MMA issuer: tcgen05.mma M128 x N128 x K16 into one of TWO accumulators in TMEM (2 x 128 columns);
//   warps 2..9  epilogue (8 warps: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4): drain the accumulator of item i
//               into registers (64 channels of one pixel per thread, bf16) and hand the TMEM buffer back at once, so the MMAs of
//               item i+1 run under the statistics exchange, the normalisation / Mish / residual math and the stores of item i.
// GroupNorm statistics of an image span several items (8 at 32x32): every CTA adds its {sum, sum of squares} per (image, group)
// to a zeroed fp32 workspace with red.global.add, then bumps a per-(image, N tile) arrival counter (release) and waits until all
// tiles of the image have arrived (acquire).  Items of one image are consecutive and the grid is a multiple of the tiles per
// image, so they are in flight on co-resident CTAs in the same round: the wait cannot deadlock (and is bounded: a protocol bug
// traps instead of hanging the GPU).  Output: bf16 NHWC, 128 contiguous bytes per thread as four 32-byte stores; no staging in
// shared memory at all.
#include "conv_tc_common.cuh"

namespace dd {

constexpr int PS_EPI_WARPS = 8;
constexpr int PS_THREADS = 64 + 32 * PS_EPI_WARPS;         // 320
constexpr int PS_MAX_COUT = 512;                           // bias, gamma, beta of the WHOLE layer are staged once per CTA
constexpr int PS_PAR_BYTES = 3 * PS_MAX_COUT * 4;
constexpr int ps_smem(int nh, int nb) { return nh * HALO_SLOT + nb * HALO_B_BYTES + 1024 /*align*/ + 256 /*barriers*/ + PS_PAR_BYTES; }

__device__ __forceinline__ void ps_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void ldg_v8(const void* p, uint32_t (&a)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t (&a)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]) : "memory");
}

struct PsItem { int w0, h0, img, n_tile; };
__device__ __forceinline__ PsItem ps_decode(const TcParams& p, int it) {
    const int tpi = p.tiles_w * p.tiles_h, ntn = p.Cout >> 7;
    PsItem i;
    const int m = it % tpi, rest = it / tpi;
    i.n_tile = rest % ntn; i.img = rest / ntn;
    i.w0 = (m % p.tiles_w) * HALO_TW; i.h0 = (m / p.tiles_w) * HALO_TH;
    return i;
}

// <CPG_SH: log2(channels per GroupNorm group), NH / NB: halo / weight ring depth>.  One CTA per SM.  (A two-CTAs-per-SM form
// with 2 + 4 ring slots and 96 registers was built and measured: the occupancy calculator still grants one CTA per SM and the
// shallower rings lose, 1.02 - 1.10 ms per step against 0.94; removed again -- profiles/README.md, round 2 passes g - k.)
template <int CPG_SH, int PS_NH, int PS_NB>
__global__ void __launch_bounds__(PS_THREADS, 1) conv_tc_halo_persist_kernel(const __grid_constant__ TcParams p) {
    constexpr uint32_t DY_BYTES = (HALO_TW + 2) * 128u;                 // shared-memory bytes between filter rows of the halo
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bbase = base + PS_NH * HALO_SLOT;
    const uint32_t bars = bbase + PS_NB * HALO_B_BYTES;
    auto hfull = [&](int s) { return bars + 8u * s; };
    auto hempty = [&](int s) { return bars + 8u * (PS_NH + s); };
    auto bfull = [&](int s) { return bars + 8u * (2 * PS_NH + s); };
    auto bempty = [&](int s) { return bars + 8u * (2 * PS_NH + PS_NB + s); };
    auto tfull = [&](int s) { return bars + 8u * (2 * PS_NH + 2 * PS_NB + s); };
    auto tempty = [&](int s) { return bars + 8u * (2 * PS_NH + 2 * PS_NB + 2 + s); };
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * PS_NH + 2 * PS_NB + 4);
    static_assert(8 * (2 * PS_NH + 2 * PS_NB + 4) + 8 <= 256, "barrier block");
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    float* s_par = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));      // [bias | gamma | beta][PS_MAX_COUT]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.chunks0 + p.chunks1, cin = nchunks * 64;
    const int tpi = p.tiles_w * p.tiles_h, ntn = p.Cout >> 7;
    const int n_items = tpi * ntn * p.B;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmH0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        for (int s = 0; s < PS_NH; ++s) { mbar_init(hfull(s), 1); mbar_init(hempty(s), 1); }
        for (int s = 0; s < PS_NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), PS_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the layer's parameters (packed before the step, not written by the preceding launch): staged under its tail
    for (int c = threadIdx.x; c < p.Cout; c += PS_THREADS) {
        s_par[c] = p.bias ? p.bias[c] : 0.f;
        s_par[PS_MAX_COUT + c] = p.gn_gamma[c];
        s_par[2 * PS_MAX_COUT + c] = p.gn_beta[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (threadIdx.x == 0) tstamp(p, 0);
    pdl_sync();
    if (threadIdx.x == 0) tstamp(p, 2);

    if (warp == 0) {
        // ===== TMA producer: the rings keep running across item boundaries =====
        const uint32_t b_tx = 128u * TC_BK * 2;
        const int chunks0 = p.chunks0;
        int bs = 0, bround = 0, hs = 0, hround = 0;
        uint32_t sB = bbase;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const PsItem im = ps_decode(p, it);
            const int brow = im.n_tile * 128;
            for (int c = 0; c < nchunks; ++c) {
                if (hround > 0) mbar_wait(hempty(hs), (hround - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(hfull(hs), HALO_TX);
                    const CUtensorMap* tm = c < chunks0 ? &p.tmH0 : &p.tmH1;
                    tma_load_5d(tm, hfull(hs), base + hs * HALO_SLOT, (c < chunks0 ? c : c - chunks0) * 64, im.w0 - 1, im.h0 - 1, im.img, 0);
                }
                __syncwarp();
                if (++hs == PS_NH) { hs = 0; ++hround; }
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
                    if (++bs == PS_NB) { bs = 0; ++bround; sB = bbase; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: item k accumulates into TMEM buffer k & 1 =====
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const uint64_t a_hi = ((uint64_t)(((HALO_TW + 2) * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        int bs = 0, hs = 0, k = 0;
        uint32_t bpar = 0, hpar = 0;
        uint32_t sB = bbase;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            const int buf = k & 1;
            if (k >= 2) mbar_wait(tempty(buf), ((k >> 1) - 1) & 1);          // the epilogue of item k - 2 has drained this buffer
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)(buf * 128);
            uint32_t acc = 0;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(hfull(hs), hpar);
                uint32_t rowA = base + hs * HALO_SLOT;
#pragma unroll 1
                for (int r = 0; r < 3; ++r) {
#pragma unroll 1
                    for (int sx = 0; sx < 3; ++sx) {
                        mbar_wait(bfull(bs), bpar);
                        if (k == 0 && acc == 0 && lane == 0) tstamp(p, 3);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t bd = umma_desc(sB);
                            const uint64_t ad = (uint64_t)(((rowA + 128u * sx) & 0x3FFFFu) >> 4) | a_hi;
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 16; ++kk)
                                umma_f16(dcol, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, (acc | kk) ? 1u : 0u);
                            umma_commit(bempty(bs));
                        }
                        __syncwarp();
                        acc = 1u;
                        sB += HALO_B_BYTES;
                        if (++bs == PS_NB) { bs = 0; bpar ^= 1u; sB = bbase; }
                    }
                    rowA += DY_BYTES;
                }
                if (elect_one()) umma_commit(hempty(hs));
                __syncwarp();
                if (++hs == PS_NH) { hs = 0; hpar ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull(buf));
            __syncwarp();
            if (lane == 0) tstamp(p, k == 0 ? 4 : 7);                      // MMAs of the first / of the latest item issued
        }
    } else {
        // ===== epilogue: 8 warps; thread = (pixel row r of the tile, 64-channel half) =====
        // Software-pipelined over the CTA's items: phase 1 of item k (drain TMEM, publish the statistics) runs BEFORE phase 2 of
        // item k - 1 (normalise, activate, store), so the other tiles of image k - 1 have had a whole main loop to arrive and the
        // wait costs nothing; a thread carries two packed rows (2 x 32 registers).
        constexpr int NGH = 64 >> CPG_SH;                                 // GroupNorm groups inside a thread's 64 channels (1, 2, 4, 8)
        const int et = threadIdx.x - 64;                                  // 0 .. 255
        const int q = warp & 3, hsel = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const int ww = r & (HALO_TW - 1), hh = r >> 3;
        const int G = p.G;
        uint32_t* cnt_base = reinterpret_cast<uint32_t*>(p.gn_stats + (int64_t)p.B * G * 2);
        const int c0 = hsel * 64;

        auto phase1 = [&](const int k, const PsItem& im, uint32_t (&row)[32]) {
            const int buf = k & 1, cbase = im.n_tile * 128;
            const float* par = s_par + cbase;                             // bias of this item's 128 channels
            mbar_wait(tfull(buf), (k >> 1) & 1);
            if (et == 0) tstamp(p, k == 0 ? 5 : 13);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(buf * 128 + c0) + ((uint32_t)(q * 32) << 16);
            // drain 64 columns, 16 at a time: + bias, group statistics of the fp32 values, pack to bf16
            float gs[NGH], gq[NGH];
#pragma unroll
            for (int g = 0; g < NGH; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                uint32_t a[16];
                tmem_ld16(taddr + (uint32_t)(16 * qd), a);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = 16 * qd + 2 * j;                         // channel within the half (static)
                    const float2 bi = *reinterpret_cast<const float2*>(par + c0 + c);
                    const float x0 = __uint_as_float(a[2 * j]) + bi.x, x1 = __uint_as_float(a[2 * j + 1]) + bi.y;
                    gs[c >> CPG_SH] += x0 + x1;
                    gq[c >> CPG_SH] = fmaf(x0, x0, fmaf(x1, x1, gq[c >> CPG_SH]));
                    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                    row[8 * qd + j] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            // the accumulator is in registers: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(buf));
            if (et == 0 && k == 0) tstamp(p, 10);
            float* st = p.gn_stats + ((int64_t)im.img * G + ((cbase + c0) >> CPG_SH)) * 2;
#pragma unroll
            for (int g = 0; g < NGH; ++g) {
                float sa = gs[g], qa = gq[g];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, d);
                    qa += __shfl_xor_sync(0xffffffffu, qa, d);
                }
                if (lane == 0) { red_add_f32(st + 2 * g, sa); red_add_f32(st + 2 * g + 1, qa); }
            }
            ps_epi_bar();                                                  // every warp of the CTA has issued its partial sums
            if (et == 0) red_release_add_u32(cnt_base + (im.img * ntn + im.n_tile), 1u);
            if (et == 0 && k == 0) tstamp(p, 11);
        };

        auto phase2 = [&](const int k, const PsItem& im, uint32_t (&row)[32]) {
            const int cbase = im.n_tile * 128;
            const float* s_gamma = s_par + PS_MAX_COUT + cbase, *s_beta = s_par + 2 * PS_MAX_COUT + cbase;
            const int64_t pix = ((int64_t)im.img * p.H + (im.h0 + hh)) * p.W + (im.w0 + ww);
            const float* tbp = nullptr;
            if (p.tbias) {
                const int trow_i = p.trow ? p.trow[(int64_t)im.img * p.trow_stride] : im.img;
                tbp = p.tbias + (int64_t)trow_i * p.tb_stride + cbase + c0;
            }
            const __nv_bfloat16* resp = p.residual ? p.residual + pix * p.Cout + cbase + c0 : nullptr;
            uint32_t res[2][8];
            if (resp) ldg_v8(resp, res[0]);                                // first quarter of the residual row, under the wait
            if (et == 0) {
                const uint32_t* cnt = cnt_base + (im.img * ntn + im.n_tile);
                if (ld_acquire_u32(cnt) < (uint32_t)tpi) {
                    const long long t0 = clock64();
                    while (ld_acquire_u32(cnt) < (uint32_t)tpi) {
                        if (clock64() - t0 > 4000000000LL) __trap();
                    }
                }
            }
            ps_epi_bar();                                                  // all tiles of the image have arrived
            if (et == 0 && k == 0) tstamp(p, 12);
            const float* st = p.gn_stats + ((int64_t)im.img * G + ((cbase + c0) >> CPG_SH)) * 2;
            float mean[NGH], rstd[NGH];
#pragma unroll
            for (int g = 0; g < NGH; ++g) {
                const float2 sv = __ldcg(reinterpret_cast<const float2*>(st) + g);
                mean[g] = sv.x * p.gn_inv_n;
                rstd[g] = rsqrtf(fmaxf(sv.y * p.gn_inv_n - mean[g] * mean[g], 0.f) + p.gn_eps);
            }
            // normalise, activate, + time bias, + residual; LayerNorm partial sums of the rounded result; 32-byte stores
            float ls = 0.f, lq = 0.f;
            const bool want_ln = p.ln_part != nullptr;
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + cbase + c0;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                if (resp && qd < 3) ldg_v8(resp + 16 * (qd + 1), res[(qd + 1) & 1]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cl = 16 * qd + 2 * j, c = c0 + cl;           // cl static
                    const float m_ = mean[cl >> CPG_SH], r_ = rstd[cl >> CPG_SH];
                    const float2 gm = *reinterpret_cast<const float2*>(s_gamma + c), be = *reinterpret_cast<const float2*>(s_beta + c);
                    const uint32_t rw = row[8 * qd + j];
                    const float x0 = __uint_as_float(rw << 16), x1 = __uint_as_float(rw & 0xffff0000u);
                    const float ga0 = r_ * gm.x, ga1 = r_ * gm.y;
                    float y0 = mish_fast(fmaf(x0, ga0, fmaf(-m_, ga0, be.x)));
                    float y1 = mish_fast(fmaf(x1, ga1, fmaf(-m_, ga1, be.y)));
                    if (tbp) { const float2 tb = __ldg(reinterpret_cast<const float2*>(tbp + cl)); y0 += tb.x; y1 += tb.y; }
                    if (resp) {
                        const uint32_t rr = res[qd & 1][j];
                        y0 += __uint_as_float(rr << 16); y1 += __uint_as_float(rr & 0xffff0000u);
                    }
                    __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                    const uint32_t hv = *reinterpret_cast<uint32_t*>(&h);
                    row[8 * qd + j] = hv;
                    if (want_ln) {
                        const float z0 = __uint_as_float(hv << 16), z1 = __uint_as_float(hv & 0xffff0000u);
                        ls += z0 + z1; lq = fmaf(z0, z0, fmaf(z1, z1, lq));
                    }
                }
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             ::"l"(op + 16 * qd), "r"(row[8 * qd]), "r"(row[8 * qd + 1]), "r"(row[8 * qd + 2]), "r"(row[8 * qd + 3]),
                               "r"(row[8 * qd + 4]), "r"(row[8 * qd + 5]), "r"(row[8 * qd + 6]), "r"(row[8 * qd + 7]) : "memory");
            }
            if (want_ln) *reinterpret_cast<float2*>(p.ln_part + (pix * (2 * ntn) + (2 * im.n_tile + hsel)) * 2) = make_float2(ls, lq);
            if (et == 0) tstamp(p, k == 0 ? 6 : 14);
        };

        uint32_t cur[32], prv[32];
        PsItem pim = {0, 0, 0, 0};
        int k = 0;
        for (int it = blockIdx.x; ; it += gridDim.x, ++k) {
            const bool has = it < n_items;
            PsItem im = pim;
            if (has) { im = ps_decode(p, it); phase1(k, im, cur); }
            if (k > 0) phase2(k - 1, pim, prv);
            if (!has) break;
#pragma unroll
            for (int j = 0; j < 32; ++j) prv[j] = cur[j];
            pim = im;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

bool halo_persist_ok(int kind, int H, int W, int Cout, int G) {
    const bool off = getenv("DD_NO_PERSIST") != nullptr;       // read per call: tests switch it inside one process
    if (off || kind != DD_TC_CONV3x3 || H < HALO_TH || W < HALO_TW || Cout % 128 || Cout > PS_MAX_COUT || G <= 0 || Cout % G) return false;
    const int cpg = Cout / G, tpi = (H / HALO_TH) * (W / HALO_TW);
    return (cpg == 8 || cpg == 16 || cpg == 32 || cpg == 64) && tpi <= num_sms();
}

template <int CPG_SH, int NH, int NB>
static int launch_ps(const TcParams& p, cudaStream_t st) {
    auto kern = conv_tc_halo_persist_kernel<CPG_SH, NH, NB>;
    constexpr int smem = ps_smem(NH, NB);
    static int ctas_per_sm = -1;            // per template instance
    if (ctas_per_sm < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int n = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, PS_THREADS, smem);
        if (e != cudaSuccess || n < 1) { set_error("conv_tc_gn(persistent): %s (occupancy %d)", cudaGetErrorString(e), n); return DD_ERR_CUDA; }
        ctas_per_sm = 1;
    }
    // every CTA of the grid must be resident at once (the tiles of an image wait for each other) and the grid is a multiple of
    // the tiles per image (they then sit in the same round)
    const int tpi = p.tiles_w * p.tiles_h, n_items = tpi * (p.Cout / 128) * p.B;
    int g = num_sms() * ctas_per_sm / tpi * tpi;
    if (g > n_items) g = n_items;
    if (g < tpi) { set_error("conv_tc_gn(persistent): an image of %d tiles does not fit the %d resident CTAs", tpi, num_sms() * ctas_per_sm); return DD_ERR_ARG; }
    launch_pdl(kern, dim3(g), dim3(PS_THREADS), smem, st, p);
    return check_launch("conv_tc_gn(persistent)");
}

int launch_halo_persist(const TcParams& p, cudaStream_t st) {
    switch (p.cpg_shift) {          // rings: 3 halos (70.7 KB) + 8 weight tiles (128 KB)
        case 3: return launch_ps<3, 3, 8>(p, st);
        case 4: return launch_ps<4, 3, 8>(p, st);
        case 5: return launch_ps<5, 3, 8>(p, st);
        case 6: return launch_ps<6, 3, 8>(p, st);
    }
    set_error("conv_tc_gn(persistent): unsupported channels per group (shift %d)", p.cpg_shift);
    return DD_ERR_ARG;
}


// =============================================================================================================================
// Persistent GEMM for the 1x1 convolutions (to_qkv with the LayerNorm fold, the attention output with per-sample weights, the
// ResnetBlock res_conv, the im2col'd first convolution): blocks.py:103, 123-124, unet.py:71.
// These layers have K = 128 .. 512, i.e. 2 .. 8 k-blocks per 128 x 128 tile: as one tile per CTA they are all prologue and
// epilogue (qkv @32x32: 1536 CTAs x (2.3 k clk set-up + 0.8 k clk of MMAs + 4.5 - 6 k clk staged epilogue) = 30 us for a 50 MB
// output).  Here one CTA per SM walks the (row tile, 128-channel tile) items with a 6-stage operand ring running across items,
// two accumulators in TMEM and the same 8-warp register epilogue as the halo kernel above (64 channels of one row per thread,
// 32-byte stores), so an item costs its epilogue (~2 k clk) and nothing else.
//   rows are the M = B*H*W pixels of the NHWC tensor (a 2-D tensor map; TMA zero-fills past M), optional per-row LayerNorm
//   statistics (dd_conv_tc_ln), optional residual, optional GroupNorm {sum, sum of squares} atomics for a following dd_gn_mish
//   (template CPG_SH > 0; needs H*W % 128 == 0 so that a tile lies inside one image), optional per-image weights.
// =============================================================================================================================
constexpr int GS_STAGES = 6;
constexpr int GS_STAGE_BYTES = 2 * TC_A_BYTES;                     // 128 x 64 bf16 of A + 128 x 64 of B
constexpr int GS_MAX_COUT = 1024;
constexpr int GS_SMEM = GS_STAGES * GS_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * GS_MAX_COUT * 4 /*bias, weight row sums of the layer*/;

template <int CPG_SH>
__global__ void __launch_bounds__(PS_THREADS, 1) conv_tc_gemm_persist_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + GS_STAGES * GS_STAGE_BYTES;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (GS_STAGES + s); };
    auto tfull = [&](int s) { return bars + 8u * (2 * GS_STAGES + s); };
    auto tempty = [&](int s) { return bars + 8u * (2 * GS_STAGES + 2 + s); };
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * GS_STAGES + 4);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    float* s_par = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));      // [bias | wsum][GS_MAX_COUT]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.chunks0;
    const int HW = p.H * p.W;
    const int64_t M = (int64_t)p.B * HW;
    const int n_m = (int)((M + 127) >> 7), ntn = p.Cout >> 7;
    const int n_items = n_m * ntn;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        for (int s = 0; s < GS_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), PS_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int c = threadIdx.x; c < p.Cout; c += PS_THREADS) {       // packed before the step: safe to read before the grid dependency resolves
        s_par[c] = p.bias ? p.bias[c] : 0.f;
        s_par[GS_MAX_COUT + c] = p.ln_in ? p.ln_wsum[c] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (threadIdx.x == 0) tstamp(p, 0);
    pdl_sync();
    if (threadIdx.x == 0) tstamp(p, 2);

    if (warp == 0) {
        // ===== TMA producer =====
        int st = 0, round = 0;
        uint32_t sA = base;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int n_tile = it % ntn, m_tile = it / ntn;
            const int img = p.w_per_sample ? (int)(((int64_t)m_tile << 7) / HW) : 0;
            for (int kb = 0; kb < nkb; ++kb) {
                if (round > 0) mbar_wait(empty(st), (round - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(full(st), GS_STAGE_BYTES);
                    tma_load_2d(&p.tmA0, full(st), sA, kb * 64, m_tile * 128);
                    if (p.w_per_sample) tma_load_3d(&p.tmB, full(st), sA + TC_A_BYTES, kb * 64, n_tile * 128, img);
                    else tma_load_2d(&p.tmB, full(st), sA + TC_A_BYTES, kb * 64, n_tile * 128);
                }
                __syncwarp();
                sA += GS_STAGE_BYTES;
                if (++st == GS_STAGES) { st = 0; ++round; sA = base; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        int st = 0, k = 0;
        uint32_t par = 0;
        uint32_t sA = base;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            const int buf = k & 1;
            if (k >= 2) mbar_wait(tempty(buf), ((k >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)(buf * 128);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full(st), par);
                if (k == 0 && kb == 0 && lane == 0) tstamp(p, 3);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = umma_desc(sA), bd = umma_desc(sA + TC_A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 16; ++kk)
                        umma_f16(dcol, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, (kb | kk) ? 1u : 0u);
                    umma_commit(empty(st));
                }
                __syncwarp();
                sA += GS_STAGE_BYTES;
                if (++st == GS_STAGES) { st = 0; par ^= 1u; sA = base; }
            }
            if (elect_one()) umma_commit(tfull(buf));
            __syncwarp();
            if (lane == 0) tstamp(p, k == 0 ? 4 : 7);
        }
    } else {
        // ===== epilogue: 8 warps; thread = (row of the tile, 64-channel half) =====
        constexpr int NGH = CPG_SH > 0 ? (64 >> CPG_SH) : 1;
        const int et = threadIdx.x - 64;
        const int q = warp & 3, hsel = (warp - 2) >> 2;
        const int r = q * 32 + lane, c0 = hsel * 64;
        const bool ln_fold = p.ln_in != nullptr;
        // channel sums of row m for the LayerNorm fold: {sum, sum of squares} over the parts, fetched one item ahead
        auto ln_sums = [&](const int64_t m, float& su, float& sq) {
            su = 0.f; sq = 0.f;
            if (ln_fold && m < M) {
                const float2* lp = reinterpret_cast<const float2*>(p.ln_in) + m * p.ln_in_parts;
                for (int i = 0; i < p.ln_in_parts; ++i) { const float2 v = __ldg(lp + i); su += v.x; sq += v.y; }
            }
        };
        float nsu, nsq;
        ln_sums((((int64_t)(blockIdx.x / ntn)) << 7) + r, nsu, nsq);
        int k = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            const int n_tile = it % ntn, m_tile = it / ntn;
            const int buf = k & 1, cbase = n_tile * 128;
            const int64_t m = ((int64_t)m_tile << 7) + r;
            const bool valid = m < M;
            const float* par = s_par + cbase;
            float ln_a = 1.f, ln_b = 0.f;
            if (ln_fold && valid) {
                const float mean = nsu * p.ln_inv_c;
                ln_a = 1.f / (sqrtf(fmaxf(nsq * p.ln_inv_c - mean * mean, 0.f)) + p.ln_eps);
                ln_b = -mean * ln_a;
            }
            if (it + (int)gridDim.x < n_items) ln_sums((((int64_t)((it + (int)gridDim.x) / ntn)) << 7) + r, nsu, nsq);      // next item's rows
            const __nv_bfloat16* resp = (p.residual && valid) ? p.residual + m * p.Cout + cbase + c0 : nullptr;
            uint32_t res[2][8];
            if (resp) ldg_v8(resp, res[0]);
            if (et == 0 && k == 1) tstamp(p, 11);
            mbar_wait(tfull(buf), (k >> 1) & 1);
            if (et == 0) tstamp(p, k == 0 ? 5 : (k == 1 ? 13 : 12));
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(buf * 128 + c0) + ((uint32_t)(q * 32) << 16);
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.Cout + cbase + c0;
            float gs[NGH], gq[NGH];
#pragma unroll
            for (int g = 0; g < NGH; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                uint32_t a[16];
                tmem_ld16(taddr + (uint32_t)(16 * qd), a);
                if (qd == 3) {                                             // every column of this thread is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty(buf));
                }
                if (resp && qd < 3) ldg_v8(resp + 16 * (qd + 1), res[(qd + 1) & 1]);
                uint32_t ov[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + 16 * qd + 2 * j;
                    const float2 bi = *reinterpret_cast<const float2*>(par + c);
                    float x0 = __uint_as_float(a[2 * j]), x1 = __uint_as_float(a[2 * j + 1]);
                    if (ln_fold) {
                        const float2 ws = *reinterpret_cast<const float2*>(par + GS_MAX_COUT + c);
                        x0 = fmaf(x0, ln_a, fmaf(ln_b, ws.x, bi.x)); x1 = fmaf(x1, ln_a, fmaf(ln_b, ws.y, bi.y));
                    } else { x0 += bi.x; x1 += bi.y; }
                    if (CPG_SH > 0 && valid) {
                        constexpr int SH = CPG_SH > 0 ? CPG_SH : 6;
                        gs[(16 * qd + 2 * j) >> SH] += x0 + x1;
                        gq[(16 * qd + 2 * j) >> SH] = fmaf(x0, x0, fmaf(x1, x1, gq[(16 * qd + 2 * j) >> SH]));
                    }
                    if (resp) {
                        const uint32_t rr = res[qd & 1][j];
                        x0 += __uint_as_float(rr << 16); x1 += __uint_as_float(rr & 0xffff0000u);
                    }
                    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                    ov[j] = *reinterpret_cast<uint32_t*>(&h);
                }
                if (valid) stg_v8(op + 16 * qd, ov);
            }
            if (et == 0) tstamp(p, k == 0 ? 6 : (k == 1 ? 10 : 14));
            if (CPG_SH > 0 && p.gn_stats) {
                // a tile lies inside one image (H*W % 128 == 0): one (image, group) target per warp and group
                constexpr int SH = CPG_SH > 0 ? CPG_SH : 6;
                const int img = (int)(((int64_t)m_tile << 7) / HW);
                float* st = p.gn_stats + ((int64_t)img * p.G + ((cbase + c0) >> SH)) * 2;
#pragma unroll
                for (int g = 0; g < NGH; ++g) {
                    float sa = gs[g], qa = gq[g];
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        sa += __shfl_xor_sync(0xffffffffu, sa, d);
                        qa += __shfl_xor_sync(0xffffffffu, qa, d);
                    }
                    if (lane == 0) { red_add_f32(st + 2 * g, sa); red_add_f32(st + 2 * g + 1, qa); }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// rows x channels view of an NHWC tensor: box (64 channels, 128 rows)
static int make_rows_map(CUtensorMap* tm, const void* ptr, int C, int pitch, int64_t M) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DD_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {64u, 128u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rows C=%d M=%lld) failed: %d", C, (long long)M, (int)r); return DD_ERR_CUDA; }
    return DD_OK;
}

// Layers worth the persistent GEMM: plain 1x1 convolutions with bf16 NHWC output, Cout % 128 == 0, at least two items per SM
// (smaller layers are a single wave either way and keep the one-tile-per-CTA kernel, whose 96 KB let neighbours overlap).
bool gemm_persist_ok(int kind, int B, int H, int W, int C1, int C2, int Cout, int G, bool stats, bool wps) {
    if (getenv("DD_NO_PERSIST_GEMM") || kind != DD_TC_CONV1x1 || C2 != 0 || C1 % 64 || Cout % 128 || Cout > GS_MAX_COUT) return false;
    const int64_t M = (int64_t)B * H * W;
    const int64_t items = ((M + 127) / 128) * (Cout / 128);
    if (items < 2 * num_sms()) return false;
    if ((stats || wps) && (H * W) % 128) return false;
    if (stats) { const int cpg = (G > 0 && Cout % G == 0) ? Cout / G : 0; if (cpg != 8 && cpg != 16 && cpg != 32) return false; }
    return true;
}

int launch_gemm_persist(TcParams& p, const void* x, int x_pitch, int C1, cudaStream_t st) {
    const int64_t M = (int64_t)p.B * p.H * p.W;
    int rc = make_rows_map(&p.tmA0, x, C1, x_pitch, M);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_gemm_persist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_gemm_persist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_gemm_persist_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_gemm_persist_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM);
        if (e != cudaSuccess) { set_error("conv_tc(persistent GEMM): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
        attr_done = true;
    }
    const int64_t items = ((M + 127) / 128) * (p.Cout / 128);
    const int grid = (int)(items < num_sms() ? items : num_sms());
    const int sh = p.gn_stats ? p.cpg_shift : 0;
    switch (sh) {
        case 0: launch_pdl(conv_tc_gemm_persist_kernel<0>, dim3(grid), dim3(PS_THREADS), GS_SMEM, st, p); break;
        case 3: launch_pdl(conv_tc_gemm_persist_kernel<3>, dim3(grid), dim3(PS_THREADS), GS_SMEM, st, p); break;
        case 4: launch_pdl(conv_tc_gemm_persist_kernel<4>, dim3(grid), dim3(PS_THREADS), GS_SMEM, st, p); break;
        case 5: launch_pdl(conv_tc_gemm_persist_kernel<5>, dim3(grid), dim3(PS_THREADS), GS_SMEM, st, p); break;
        default: set_error("conv_tc(persistent GEMM): unsupported channels per group"); return DD_ERR_ARG;
    }
    return check_launch("conv_tc(persistent GEMM)");
}

}  // namespace dd
