// Persistent form of the 3x3 stride-1 halo convolution with the GroupNorm + Mish (+ time bias, + residual) of
// models/unet/blocks.py:73-84, 105-115 fused behind it (dd_conv_tc_gn on maps of at least 16x8 pixels, Cout % 128 == 0).
//
// Why: with one tile per CTA (conv_tc_halo_kernel) the two CTAs of an SM run in lock step -- both in their main loop, then both
// in their epilogue -- so the tensor pipe idles during every epilogue and the activation math (2 MUFU per element) is exposed
// (profiles/README.md, round 2).  Here ONE CTA per SM (20 warps) walks the work items (pixel tile x 128-channel tile) of the layer:
//   warp 0       TMA producer: one (18 x 10)-pixel halo box per 64-channel chunk (requested a chunk ahead), one weight STAGE per
//                (chunk, filter row) = three 128 x 64 tiles, 48 KB; rings of 3 halos + 3 stages (215 KB) run across item boundaries;
//                the first three stages are requested BEFORE griddepcontrol.wait (weights never depend on the preceding launch);
//   warps 1, 2   MMA issuers: twelve tcgen05.mma M128 x N128 x K16 per barrier wait into one of TWO accumulator buffers in TMEM.
//                One issuer on short-K layers; from 256 input channels on two, each taking half of the K = 16 slices of every
//                stage into its own accumulator (one warp needs ~430 clk of instruction latency per wait + elect + descriptor
//                arithmetic + commit, whatever the rings do: four MMAs per wait ran the pipe at 111 clk per MMA instead of 64);
//   warp 3       statistics warp (below);
//   warps 4..19  epilogue (16 warps: TMEM lane quadrant = warp % 4, 32-channel block = (warp - 4) / 4): drain the accumulator(s) of
//                item i into registers (32 channels of one pixel per thread, bf16) and hand the TMEM buffer back at once, so the MMAs
//                of item i+1 run under the statistics exchange, the normalisation / Mish / residual math and the stores of item i.
// GroupNorm statistics of an image span several items (8 at 32x32).  Every epilogue warp leaves its {sum, sum of squares} per
// group in shared memory and arrives on a named barrier; the statistics warp adds them, publishes one 16-byte packet
// {sum, 1, sum of squares, 1} per (image, channel tile, pixel tile, group) in a zeroed workspace (st.relaxed.gpu.v4: value and flag
// travel in the same 8 bytes, so neither atomics nor fences are needed -- a gpu-scope release cost 2 - 9 k clk per item here),
// polls the packets of the image's other tiles, computes {rstd * gamma, beta - mean * rstd * gamma, time bias} for the item's 128
// channels into shared memory and releases the epilogue through an mbarrier.  Items of one image are consecutive and the grid is
// a multiple of the tiles per image, so they are in flight on co-resident CTAs in the same round: the wait cannot deadlock (and
// is bounded: a protocol bug traps instead of hanging the GPU).  The epilogue probes (test_wait) "next accumulator ready" against
// "statistics of the previous item complete" and finishes the previous item first when it can, so nothing but the last item's own
// epilogue is left behind the last MMA.  Output: bf16 NHWC, 64 contiguous bytes per thread as two 32-byte stores; no staging.
#include "conv_tc_common.cuh"

namespace dd {

constexpr int PS_EPI_WARPS = 8;
constexpr int PS_THREADS = 64 + 32 * PS_EPI_WARPS;         // 320: persistent GEMM
#ifndef DD_PS_EPI_WARPS
#define DD_PS_EPI_WARPS 16         // epilogue warps of the halo kernel: 8 (64 channels of a pixel per thread) or 16 (32 channels)
#endif
// warps of the halo kernel: 0 TMA producer, 1 (2) MMA issuers, 3 statistics warp, 4 .. 4 + PH_EW - 1 epilogue
// (20 warps: with 21 the register file is split for 24 and a thread gets 80 registers instead of 96)
constexpr int PH_EW = DD_PS_EPI_WARPS;
constexpr int PH_CS = PH_EW / 4;                           // channel blocks per 128-channel tile (one per group of four epilogue warps)
constexpr int PH_CW = 128 / PH_CS;                         // channels of one pixel per epilogue thread
constexpr int PH_WARP_B = 0, PH_WARP_MMA = 1, PH_WARP_STATS = 3, PH_WARP_EPI = 4;
constexpr int PH_THREADS = 32 * (4 + PH_EW);
static_assert(PH_EW == 8 || PH_EW == 16, "8 or 16 epilogue warps");
// named barriers 2 / 3 (item parity): the epilogue threads arrive, the statistics warp waits
__device__ __forceinline__ void ph_stats_bar_arrive(int par) {
    if (par) asm volatile("bar.arrive 3, %0;" ::"n"(32 * PH_EW + 32) : "memory");
    else asm volatile("bar.arrive 2, %0;" ::"n"(32 * PH_EW + 32) : "memory");
}
__device__ __forceinline__ void ph_stats_bar_sync(int par) {
    if (par) asm volatile("bar.sync 3, %0;" ::"n"(32 * PH_EW + 32) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"n"(32 * PH_EW + 32) : "memory");
}
#if DD_TC_TIMELINE
#define TL_WAIT(acc, stmt) do { const long long t_ = clock64(); stmt; acc += clock64() - t_; } while (0)
#else
#define TL_WAIT(acc, stmt) do { stmt; } while (0)
#endif
constexpr int PS_MAX_COUT = 512;                           // bias, gamma, beta of the WHOLE layer are staged once per CTA
constexpr int PS_PAR_BYTES = 3 * PS_MAX_COUT * 4;
constexpr int PS_ROW_BYTES = 3 * HALO_B_BYTES;            // one weight stage: the three taps of a filter row for one 64-channel chunk (48 KB)
constexpr int PS_XCH_BYTES = 2 * PH_EW * 16 * 4 /*per-warp partial sums, double buffered*/ + 2 * 128 * 3 * 4 /*scale, shift, time bias per channel of the item, double buffered*/;
constexpr int ps_smem(int nh, int nb) { return nh * HALO_SLOT + nb * PS_ROW_BYTES + 1024 /*align*/ + 256 /*barriers*/ + PS_PAR_BYTES + PS_XCH_BYTES; }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// PS_MMA_WARPS: MMA-issuing warps (1 or 2); with two, each takes half of the K = 16 slices of every step into its own accumulator
// and the epilogue adds the partials (pays on the long-K layers: 256 input channels and more).
template <int CPG_SH, int PS_NH, int PS_NB, int PS_MMA_WARPS>
__global__ void __launch_bounds__(PH_THREADS, 1) conv_tc_halo_persist_kernel(const __grid_constant__ TcParams p) {
    constexpr int PS_ACC_COLS = 128 * PS_MMA_WARPS;                     // TMEM columns of one accumulator buffer
    constexpr uint32_t DY_BYTES = (HALO_TW + 2) * 128u;                 // shared-memory bytes between filter rows of the halo
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bbase = base + PS_NH * HALO_SLOT;
    const uint32_t bars = bbase + PS_NB * PS_ROW_BYTES;
    auto hfull = [&](int s) { return bars + 8u * s; };
    auto hempty = [&](int s) { return bars + 8u * (PS_NH + s); };
    auto bfull = [&](int s) { return bars + 8u * (2 * PS_NH + s); };
    auto bempty = [&](int s) { return bars + 8u * (2 * PS_NH + PS_NB + s); };
    auto tfull = [&](int s) { return bars + 8u * (2 * PS_NH + 2 * PS_NB + s); };
    auto tempty = [&](int s) { return bars + 8u * (2 * PS_NH + 2 * PS_NB + 2 + s); };
    auto sready = [&](int s) { return bars + 8u * (2 * PS_NH + 2 * PS_NB + 4 + s); };   // statistics of the image of item k (k & 1) complete
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * PS_NH + 2 * PS_NB + 6);
    static_assert(8 * (2 * PS_NH + 2 * PS_NB + 6) + 8 <= 256, "barrier block");
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
    float* s_par = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));      // [bias | gamma | beta][PS_MAX_COUT]
    float* s_part = s_par + 3 * PS_MAX_COUT;                 // [item parity][epilogue warp][group slot][sum, sumsq]: partial sums of 32 pixel rows
    float2* s_ss = reinterpret_cast<float2*>(s_part + 2 * PH_EW * 16);       // [item parity][channel of the tile] {rstd * gamma, beta - mean * rstd * gamma}
    float* s_tb = reinterpret_cast<float*>(s_ss + 2 * 128);                   // [item parity][channel of the tile] time bias of the item's image
    constexpr int NGT = 128 >> CPG_SH;                       // GroupNorm groups per 128-channel tile (2 .. 16)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.chunks0 + p.chunks1, cin = nchunks * 64;
    const int tpi = p.tiles_w * p.tiles_h, ntn = p.Cout >> 7;
    const int n_items = tpi * ntn * p.B;
    // exchange of the GroupNorm partial sums between the CTAs of an image: one 16-byte packet {sum, 1, sumsq, 1} per (image,
    // 128-channel tile, pixel tile, group) in the zeroed workspace; the flag travels in the same 8 bytes as the value, so neither
    // atomics nor fences are needed (a gpu-scope release cost every item 2 - 9 k clk: it waits for the CTA's output stores)
    uint4* pk_base = reinterpret_cast<uint4*>(p.gn_stats);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmH0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
        // halo slots and accumulators are released / published by EVERY MMA-issuing warp (each commits its own MMAs)
        for (int s = 0; s < PS_NH; ++s) { mbar_init(hfull(s), 1); mbar_init(hempty(s), PS_MMA_WARPS); }
        for (int s = 0; s < PS_NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), PS_MMA_WARPS); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), PS_MMA_WARPS); mbar_init(tempty(s), PH_EW); mbar_init(sready(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == PH_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(2 * PS_ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the layer's parameters (packed before the step, not written by the preceding launch): staged under its tail
    for (int c = threadIdx.x; c < p.Cout; c += PH_THREADS) {
        s_par[c] = p.bias ? p.bias[c] : 0.f;
        s_par[PS_MAX_COUT + c] = p.gn_gamma[c];
        s_par[2 * PS_MAX_COUT + c] = p.gn_beta[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (threadIdx.x == 0) tstamp(p, 0);
    // the weight-tile producer runs ahead: the packed weights do not depend on the preceding launch, so its first stages are
    // requested while the grid dependency is still pending (144 KB of the ~170 KB a CTA loads before its first MMA)
    if (warp != PH_WARP_B) pdl_sync();
    if (threadIdx.x == 32) tstamp(p, 2);

    if (warp == PH_WARP_B) {
        // ===== TMA producer: one (18 x 10)-pixel halo box per 64-channel chunk, three 128 x 64 weight tiles per (chunk, filter row);
        //       the rings keep running across item boundaries.  (A separate warp for the halos, so that they are requested up to
        //       two chunks ahead instead of behind the previous chunk's weight tiles, changed nothing: the MMA warp waits ~1.5 k
        //       clk per CTA for halos either way -- profiles/README.md, round 2 pass p.) =====
        const uint32_t b_tx = 128u * TC_BK * 2;
        const int chunks0 = p.chunks0;
        int bs = 0, bround = 0, hs = 0, hround = 0;
        uint32_t sB = bbase;
        const int my_items = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        const int total_chunks = my_items * nchunks;
        // the halo of chunk g + 1 is requested BEFORE the weight stages of chunk g (its slot is free by then): it has a whole
        // chunk of MMAs to arrive instead of the last stage's
        auto request_halo = [&](int g) {
            const int c = g % nchunks;
            const PsItem im = ps_decode(p, blockIdx.x + (g / nchunks) * gridDim.x);
            if (hround > 0) mbar_wait(hempty(hs), (hround - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(hfull(hs), HALO_TX);
                const CUtensorMap* tm = c < chunks0 ? &p.tmH0 : &p.tmH1;
                tma_load_5d(tm, hfull(hs), base + hs * HALO_SLOT, (c < chunks0 ? c : c - chunks0) * 64, im.w0 - 1, im.h0 - 1, im.img, 0);
            }
            __syncwarp();
            if (++hs == PS_NH) { hs = 0; ++hround; }
        };
        auto weight_stages = [&](int g) {
            const int c = g % nchunks;
            const int brow = ps_decode(p, blockIdx.x + (g / nchunks) * gridDim.x).n_tile * 128;
            int kcoord = c * 64;
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {                                // one stage = the three taps of filter row r
                const uint32_t fb = bfull(bs);
                if (bround > 0) mbar_wait(bempty(bs), (bround - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(fb, 3u * b_tx);
                    tma_load_2d(&p.tmB, fb, sB, kcoord, brow);
                    tma_load_2d(&p.tmB, fb, sB + HALO_B_BYTES, kcoord + cin, brow);
                    tma_load_2d(&p.tmB, fb, sB + 2 * HALO_B_BYTES, kcoord + 2 * cin, brow);
                }
                __syncwarp();
                kcoord += 3 * cin;
                sB += PS_ROW_BYTES;
                if (++bs == PS_NB) { bs = 0; ++bround; sB = bbase; }
            }
        };
        static_assert(PS_NB >= 3, "the first chunk's three weight stages are requested before the grid dependency resolves");
        if (total_chunks > 0) weight_stages(0);
        pdl_sync();
        if (total_chunks > 0) request_halo(0);
        for (int g = 0; g < total_chunks; ++g) {
            if (g + 1 < total_chunks) request_halo(g + 1);
            if (g > 0) weight_stages(g);
        }
    } else if (warp >= PH_WARP_MMA && warp < PH_WARP_MMA + PS_MMA_WARPS) {
        // ===== MMA issuers: item k accumulates into TMEM buffer k & 1; issuer w takes its share of the K = 16 slices of every
        //       step and accumulates them into its OWN 128 columns of the buffer (the epilogue adds the partials).
        //       A step is a filter row of one chunk: twelve MMAs.  Why: ONE warp needs ~430 clk of instruction latency per barrier
        //       wait + elect + R2UR / uniform descriptor arithmetic + commit, whatever the operand rings do; with four MMAs (256 clk
        //       of tensor-pipe work) behind each wait the pipe ran at 111 clk per M128 x N128 x K16 MMA (profiles/README.md, round 2
        //       passes p - t). =====
        const int mw = warp - PH_WARP_MMA;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const uint64_t a_hi = ((uint64_t)(((HALO_TW + 2) * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        int bs = 0, hs = 0, k = 0;
        uint32_t bpar = 0, hpar = 0;
        uint32_t sB = bbase;
        long long wt_h = 0, wt_b = 0, wt_t = 0;                              // timeline builds: clocks spent waiting for halos / weight tiles / a TMEM buffer
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            const int buf = k & 1;
            if (k >= 2) TL_WAIT(wt_t, mbar_wait(tempty(buf), ((k >> 1) - 1) & 1));          // the epilogue of item k - 2 has drained this buffer
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)(buf * PS_ACC_COLS + mw * 128);
            uint32_t acc = 0;
            for (int c = 0; c < nchunks; ++c) {
                TL_WAIT(wt_h, mbar_wait(hfull(hs), hpar));
                uint32_t rowA = base + hs * HALO_SLOT;
#pragma unroll 1
                for (int r = 0; r < 3; ++r) {
                    TL_WAIT(wt_b, mbar_wait(bfull(bs), bpar));
                    if (mw == 0 && k == 0 && acc == 0 && lane == 0) tstamp(p, 3);
                    tc_fence_after();
                    if (elect_one()) {
                        // 12 / PS_MMA_WARPS MMAs behind one barrier wait: the per-wait instruction latency (~430 clk) is what bounds an
                        // issuing warp, not the tensor pipe (64 clk per MMA).  Every issuer waits on EVERY stage (a warp that skipped
                        // stages could see a barrier two phases on and mistake it for ready) and takes its share of the K = 16 slices.
                        const uint64_t bd = umma_desc(sB);
                        const uint64_t ad = (uint64_t)((rowA & 0x3FFFFu) >> 4) | a_hi;
                        constexpr int KK_N = (TC_BK / 16) / PS_MMA_WARPS;
                        const int kk0 = mw * KK_N;
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx)
#pragma unroll
                            for (int kq = 0; kq < KK_N; ++kq)
                                umma_f16(dcol, ad + (uint64_t)(8 * sx + 2 * (kk0 + kq)), bd + (uint64_t)((HALO_B_BYTES >> 4) * sx + 2 * (kk0 + kq)), idesc,
                                         (acc | sx | kq) ? 1u : 0u);
                        umma_commit(bempty(bs));
                    }
                    __syncwarp();
                    acc = 1u;
                    sB += PS_ROW_BYTES;
                    if (++bs == PS_NB) { bs = 0; bpar ^= 1u; sB = bbase; }
                    rowA += DY_BYTES;
                }
                if (elect_one()) umma_commit(hempty(hs));                  // this issuer's MMAs on the halo have retired
                __syncwarp();
                if (++hs == PS_NH) { hs = 0; hpar ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull(buf));
            __syncwarp();
            if (mw == 0 && lane == 0) tstamp(p, k == 0 ? 4 : 7);           // MMAs of the first / of the latest item issued
        }
        if (mw == 0 && lane == 0) { tstore(p, 1, wt_h); tstore(p, 8, wt_b); tstore(p, 9, wt_t); }
    } else if (warp == PH_WARP_STATS) {
        // ===== statistics warp: publishes the arrival of every item's partial sums and waits for the rest of its image, so the
        //       gpu-scope fence and the polling (2 - 5 k clk per item when the epilogue threads did them) cost the epilogue nothing =====
        int k = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            const PsItem im = ps_decode(p, it);
            const int m_in_img = (it % tpi);
            float4 tbv = make_float4(0.f, 0.f, 0.f, 0.f);                  // time bias of the image, channels 4 * lane .. + 3 of the tile
            if (p.tbias) {
                const int trow_i = p.trow ? p.trow[(int64_t)im.img * p.trow_stride] : im.img;
                tbv = __ldg(reinterpret_cast<const float4*>(p.tbias + (int64_t)trow_i * p.tb_stride + im.n_tile * 128 + 4 * lane));
            }
            ph_stats_bar_sync(k & 1);                                      // every epilogue warp has left its partial sums of item k in s_part
            if (k == 0 && lane == 0) tstamp(p, 16);
            uint4* pk_img = pk_base + ((int64_t)(im.img * ntn + im.n_tile) * tpi) * NGT;
            if (lane < NGT) {
                // group `lane` of the tile: columns [lane << CPG_SH, (lane + 1) << CPG_SH) = column blocks cb0 .. cb1 of PH_CW channels
                const int cb0 = (lane << CPG_SH) / PH_CW, cb1 = (((lane + 1) << CPG_SH) - 1) / PH_CW;
                const int slot = (PH_CW >> CPG_SH) > 0 ? lane - cb0 * (PH_CW >> CPG_SH) : 0;
                float su = 0.f, sq = 0.f;
                for (int cb = cb0; cb <= cb1; ++cb)
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const float2 v = *reinterpret_cast<const float2*>(s_part + (((k & 1) * PH_EW + cb * 4 + qq) * 8 + slot) * 2);
                        su += v.x; sq += v.y;
                    }
                asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};"
                             ::"l"(pk_img + (int64_t)m_in_img * NGT + lane), "r"(__float_as_uint(su)), "r"(1u), "r"(__float_as_uint(sq)), "r"(1u) : "memory");
            }
            if (k == 0 && lane == 0) tstamp(p, 17);
            // gather the packets of all tiles of the image: lane handles packets lane, lane + 32, ... (packet % NGT == lane % NGT)
            float su = 0.f, sq = 0.f;
            const long long t0 = clock64();
            for (int i = lane; i < tpi * NGT; i += 32) {
                uint32_t a, fa, b, fb;
                while (true) {
                    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(pk_img + i) : "memory");
                    if (fa == 1u && fb == 1u) break;
                    if (clock64() - t0 > 4000000000LL) __trap();
                }
                su += __uint_as_float(a); sq += __uint_as_float(b);
            }
            if (k == 0 && lane == 0) tstamp(p, 18);
#pragma unroll
            for (int d = 16; d >= NGT; d >>= 1) {
                su += __shfl_xor_sync(0xffffffffu, su, d);
                sq += __shfl_xor_sync(0xffffffffu, sq, d);
            }
            // per-channel scale / shift (and the image's time bias) of the item's 128 channels, four channels per lane
            float mean = su * p.gn_inv_n;
            float rstd = rsqrtf(fmaxf(sq * p.gn_inv_n - mean * mean, 0.f) + p.gn_eps);
            const int gsrc = (4 * lane) >> CPG_SH;                         // lanes 0 .. NGT - 1 hold the groups' totals
            mean = __shfl_sync(0xffffffffu, mean, gsrc);
            rstd = __shfl_sync(0xffffffffu, rstd, gsrc);
            {
                const int cb4 = im.n_tile * 128 + 4 * lane;
                const float4 gm = *reinterpret_cast<const float4*>(s_par + PS_MAX_COUT + cb4);
                const float4 be = *reinterpret_cast<const float4*>(s_par + 2 * PS_MAX_COUT + cb4);
                float2* ss = s_ss + (k & 1) * 128 + 4 * lane;
                float sc;
                sc = rstd * gm.x; ss[0] = make_float2(sc, fmaf(-mean, sc, be.x));
                sc = rstd * gm.y; ss[1] = make_float2(sc, fmaf(-mean, sc, be.y));
                sc = rstd * gm.z; ss[2] = make_float2(sc, fmaf(-mean, sc, be.z));
                sc = rstd * gm.w; ss[3] = make_float2(sc, fmaf(-mean, sc, be.w));
                *reinterpret_cast<float4*>(s_tb + (k & 1) * 128 + 4 * lane) = tbv;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sready(k & 1));                     // statistics of the image of item k complete
            if (k == 0 && lane == 0) tstamp(p, 19);
            __syncwarp();
        }
    } else if (warp >= PH_WARP_EPI && warp < PH_WARP_EPI + PH_EW) {
        // ===== epilogue: PH_EW warps; thread = (pixel row r of the tile, block of PH_CW channels) =====
        // Software-pipelined over the CTA's items: phase 1 of item k (drain TMEM, publish the statistics) runs BEFORE phase 2 of
        // item k - 1 (normalise, activate, store), so the other tiles of image k - 1 have had a whole main loop to arrive;
        // a thread carries two packed rows (2 x PH_CW / 2 registers).
        constexpr int NGH = (PH_CW >> CPG_SH) > 0 ? (PH_CW >> CPG_SH) : 1;   // GroupNorm groups a thread's channels touch
        constexpr int GSH = (PH_CW >> CPG_SH) > 0 ? CPG_SH : 31;             // channel -> group slot of the thread
        constexpr int NQ = PH_CW / 16;                                       // 16-column TMEM loads per thread
        const int et = threadIdx.x - 32 * PH_WARP_EPI;
        const int q = warp & 3, csel = (warp - PH_WARP_EPI) >> 2;
        const int r = q * 32 + lane;
        const int ww = r & (HALO_TW - 1), hh = r >> 3;
        const int c0 = csel * PH_CW;
        long long wt_f = 0;                                               // timeline builds: clocks spent waiting for an accumulator

        auto phase1 = [&](const int k, const PsItem& im, uint32_t (&row)[PH_CW / 2]) {
            const int buf = k & 1, cbase = im.n_tile * 128;
            const float* par = s_par + cbase;                             // bias of this item's 128 channels
            TL_WAIT(wt_f, mbar_wait(tfull(buf), (k >> 1) & 1));
            if (et == 0) tstamp(p, k == 0 ? 5 : 13);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(buf * PS_ACC_COLS + c0) + ((uint32_t)(q * 32) << 16);
            // drain PH_CW columns, 16 at a time: sum of the issuers' partials + bias, group statistics of the fp32 values, pack to bf16
            float gs[NGH], gq[NGH];
#pragma unroll
            for (int g = 0; g < NGH; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
#pragma unroll
            for (int qd = 0; qd < NQ; ++qd) {
                uint32_t a[16];
                static_assert(PS_MMA_WARPS == 1 || PS_MMA_WARPS == 2, "one or two MMA issuers");
                if (PS_MMA_WARPS == 1) {
                    tmem_ld16(taddr + (uint32_t)(16 * qd), a);
                } else {
                    uint32_t a2[16];
                    tmem_ld16x2(taddr + (uint32_t)(16 * qd), taddr + (uint32_t)(128 + 16 * qd), a, a2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(a[j]) + __uint_as_float(a2[j]));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = 16 * qd + 2 * j;                         // channel within the thread's block (static)
                    const float2 bi = *reinterpret_cast<const float2*>(par + c0 + c);
                    const float x0 = __uint_as_float(a[2 * j]) + bi.x, x1 = __uint_as_float(a[2 * j + 1]) + bi.y;
                    gs[c >> GSH] += x0 + x1;
                    gq[c >> GSH] = fmaf(x0, x0, fmaf(x1, x1, gq[c >> GSH]));
                    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                    row[8 * qd + j] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            // the accumulator is in registers: hand the TMEM buffer back to the MMA warps
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(buf));
            if (et == 0 && k == 0) tstamp(p, 10);
            float* sp = s_part + (((k & 1) * PH_EW + (warp - PH_WARP_EPI)) * 8) * 2;
#pragma unroll
            for (int g = 0; g < NGH; ++g) {
                float sa = gs[g], qa = gq[g];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, d);
                    qa += __shfl_xor_sync(0xffffffffu, qa, d);
                }
                if (lane == 0) *reinterpret_cast<float2*>(sp + 2 * g) = make_float2(sa, qa);
            }
            ph_stats_bar_arrive(k & 1);                                    // the statistics warp publishes them
            if (et == 0 && k == 0) tstamp(p, 11);
        };

        auto phase2 = [&](const int k, const PsItem& im, uint32_t (&row)[PH_CW / 2]) {
            const int cbase = im.n_tile * 128;
            const float2* ssp = s_ss + (k & 1) * 128 + c0;
            const float* tbs = s_tb + (k & 1) * 128 + c0;
            const int64_t pix = ((int64_t)im.img * p.H + (im.h0 + hh)) * p.W + (im.w0 + ww);
            const __nv_bfloat16* resp = p.residual ? p.residual + pix * p.Cout + cbase + c0 : nullptr;
            uint32_t res[2][8];
            if (resp) ldg_v8(resp, res[0]);                                // first 16 channels of the residual row, under the wait
            mbar_wait(sready(k & 1), (k >> 1) & 1);                        // all tiles of the image have arrived
            if (et == 0 && k == 0) tstamp(p, 12);
            // normalise, activate, + time bias, + residual; LayerNorm partial sums of the rounded result; 32-byte stores
            float ls = 0.f, lq = 0.f;
            const bool want_ln = p.ln_part != nullptr;
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + cbase + c0;
            // The per-channel parameters come from shared memory while the tensor pipe is reading its operands from it at full rate:
            // an LDS then takes hundreds of clocks.  They are requested one sub-block of four channels AHEAD of their use
            // (volatile loads keep their place in the instruction stream); without this the parameter reads were 2.8 k of the
            // 6.3 k clk this phase takes per tile (profiles/README.md, round 2 passes x - ae).
            const uint32_t ss_a = smem_u32(ssp), tb_a = smem_u32(tbs);
            constexpr int PSUB = 2;                                       // channel pairs per sub-block (registers: 2 x PSUB x 6)
            float4 ssv[2][PSUB];
            float2 tbv[2][PSUB];
            auto ldp = [&](const int buf, const int cl0) {
#pragma unroll
                for (int j = 0; j < PSUB; ++j) {
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(ssv[buf][j].x), "=f"(ssv[buf][j].y), "=f"(ssv[buf][j].z), "=f"(ssv[buf][j].w) : "r"(ss_a + (uint32_t)((cl0 + 2 * j) * 8)));
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(tbv[buf][j].x), "=f"(tbv[buf][j].y) : "r"(tb_a + (uint32_t)((cl0 + 2 * j) * 4)));
                }
            };
            ldp(0, 0);
#pragma unroll
            for (int sub = 0; sub < PH_CW / (2 * PSUB); ++sub) {
                constexpr int SPQ = 8 / PSUB;                             // sub-blocks per 16-channel store
                const int qd = sub / SPQ;
                if (sub % SPQ == 0 && resp && qd < NQ - 1) ldg_v8(resp + 16 * (qd + 1), res[(qd + 1) & 1]);
                if (sub + 1 < PH_CW / (2 * PSUB)) ldp((sub + 1) & 1, 2 * PSUB * (sub + 1));
#pragma unroll
                for (int j = 0; j < PSUB; ++j) {
                    const int pj = PSUB * sub + j;                        // pair index within the thread's channels (static)
                    const float4 ss = ssv[sub & 1][j];                    // {scale, shift} of channels 2 pj, 2 pj + 1
                    const float2 tb = tbv[sub & 1][j];
                    const uint32_t rw = row[pj];
                    const float x0 = __uint_as_float(rw << 16), x1 = __uint_as_float(rw & 0xffff0000u);
                    float y0 = mish_fast(fmaf(x0, ss.x, ss.y)) + tb.x;
                    float y1 = mish_fast(fmaf(x1, ss.z, ss.w)) + tb.y;
                    if (resp) {
                        const uint32_t rr = res[qd & 1][pj & 7];
                        y0 += __uint_as_float(rr << 16); y1 += __uint_as_float(rr & 0xffff0000u);
                    }
                    __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                    const uint32_t hv = *reinterpret_cast<uint32_t*>(&h);
                    row[pj] = hv;
                    if (want_ln) {
                        const float z0 = __uint_as_float(hv << 16), z1 = __uint_as_float(hv & 0xffff0000u);
                        ls += z0 + z1; lq = fmaf(z0, z0, fmaf(z1, z1, lq));
                    }
                }
                if (sub % SPQ == SPQ - 1)
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                                 ::"l"(op + 16 * qd), "r"(row[8 * qd]), "r"(row[8 * qd + 1]), "r"(row[8 * qd + 2]), "r"(row[8 * qd + 3]),
                                   "r"(row[8 * qd + 4]), "r"(row[8 * qd + 5]), "r"(row[8 * qd + 6]), "r"(row[8 * qd + 7]) : "memory");
            }
            if (want_ln) *reinterpret_cast<float2*>(p.ln_part + (pix * (PH_CS * ntn) + (PH_CS * im.n_tile + csel)) * 2) = make_float2(ls, lq);
            if (et == 0) tstamp(p, k == 0 ? 6 : 14);
        };

        uint32_t cur[PH_CW / 2], prv[PH_CW / 2];
        PsItem pim = {0, 0, 0, 0};
        int k = 0;
        for (int it = blockIdx.x; ; it += gridDim.x, ++k) {
            const bool has = it < n_items;
            PsItem im = pim;
            bool early = false;                                            // finish item k - 1 before draining item k
            if (has) {
                im = ps_decode(p, it);
                if (k > 0) {
                    // whichever comes first: the accumulator of item k (drain it, then finish item k - 1) or the statistics of the
                    // image of item k - 1 (finish it now: at the end of the CTA's work nothing is left behind the last drain)
                    const long long t0 = clock64();
                    while (true) {
                        if (__all_sync(0xffffffffu, mbar_test_wait(tfull(k & 1), (k >> 1) & 1))) break;
                        if (__all_sync(0xffffffffu, mbar_test_wait(sready((k - 1) & 1), ((k - 1) >> 1) & 1))) { early = true; break; }
                        if (clock64() - t0 > 4000000000LL) __trap();
                    }
                }
            }
#pragma unroll 1
            for (int step = 0; step < 2; ++step) {                          // one call site per phase (the bodies are large)
                if ((step == 0) == early) { if (k > 0) phase2(k - 1, pim, prv); }
                else if (has) phase1(k, im, cur);
            }
            if (!has) break;
#pragma unroll
            for (int j = 0; j < PH_CW / 2; ++j) prv[j] = cur[j];
            pim = im;
        }
        if (et == 0) tstore(p, 15, wt_f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PH_WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * PS_ACC_COLS) : "memory");
    }
}

int halo_persist_col_split() { return PH_CS; }

bool halo_persist_ok(int kind, int H, int W, int Cout, int G) {
    const bool off = getenv("DD_NO_PERSIST") != nullptr;       // read per call: tests switch it inside one process
    if (off || kind != DD_TC_CONV3x3 || H < HALO_TH || W < HALO_TW || Cout % 128 || Cout > PS_MAX_COUT || G <= 0 || Cout % G) return false;
    const int cpg = Cout / G, tpi = (H / HALO_TH) * (W / HALO_TW);
    return (cpg == 8 || cpg == 16 || cpg == 32 || cpg == 64) && tpi <= num_sms();
}

template <int CPG_SH, int NH, int NB, int MW>
static int launch_ps(const TcParams& p, cudaStream_t st) {
    auto kern = conv_tc_halo_persist_kernel<CPG_SH, NH, NB, MW>;
    constexpr int smem = ps_smem(NH, NB);
    static int ctas_per_sm = -1;            // per template instance
    if (ctas_per_sm < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int n = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, PH_THREADS, smem);
        if (e != cudaSuccess || n < 1) { set_error("conv_tc_gn(persistent): %s (occupancy %d)", cudaGetErrorString(e), n); return DD_ERR_CUDA; }
        ctas_per_sm = 1;
    }
    // every CTA of the grid must be resident at once (the tiles of an image wait for each other) and the grid is a multiple of
    // the tiles per image (they then sit in the same round)
    const int tpi = p.tiles_w * p.tiles_h, n_items = tpi * (p.Cout / 128) * p.B;
    int g = num_sms() * ctas_per_sm / tpi * tpi;
    if (g > n_items) g = n_items;
    if (g < tpi) { set_error("conv_tc_gn(persistent): an image of %d tiles does not fit the %d resident CTAs", tpi, num_sms() * ctas_per_sm); return DD_ERR_ARG; }
    launch_pdl(kern, dim3(g), dim3(PH_THREADS), smem, st, p);
    return check_launch("conv_tc_gn(persistent)");
}

int launch_halo_persist(const TcParams& p, cudaStream_t st) {
    // rings: 3 halos (70.7 KB) + 3 filter rows of weight tiles (144 KB); two MMA issuers from 256 input channels on
    const bool two = p.chunks0 + p.chunks1 >= 4 && !getenv("DD_PS_ONE_ISSUER");
    switch (p.cpg_shift) {
        case 3: return two ? launch_ps<3, 3, 3, 2>(p, st) : launch_ps<3, 3, 3, 1>(p, st);
        case 4: return two ? launch_ps<4, 3, 3, 2>(p, st) : launch_ps<4, 3, 3, 1>(p, st);
        case 5: return two ? launch_ps<5, 3, 3, 2>(p, st) : launch_ps<5, 3, 3, 1>(p, st);
        case 6: return two ? launch_ps<6, 3, 3, 2>(p, st) : launch_ps<6, 3, 3, 1>(p, st);
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
    if (warp != 0) pdl_sync();            // the producer requests the weight halves of its first stages before the dependency resolves
    if (threadIdx.x == 32) tstamp(p, 2);

    if (warp == 0) {
        // ===== TMA producer: stage f = (item f / nkb of this CTA, k-block f % nkb) =====
        const int my_items = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        const int total = my_items * nkb;
        const int early = total < GS_STAGES ? total : GS_STAGES;
        auto coords = [&](int f, int& kb, int& n_tile, int& m_tile, int& img) {
            const int it = blockIdx.x + (f / nkb) * gridDim.x;
            kb = f % nkb; n_tile = it % ntn; m_tile = it / ntn;
            img = p.w_per_sample ? (int)(((int64_t)m_tile << 7) / HW) : 0;
        };
        auto load_b = [&](int f, uint32_t bar, uint32_t dst) {
            int kb, n_tile, m_tile, img; coords(f, kb, n_tile, m_tile, img);
            if (p.w_per_sample) tma_load_3d(&p.tmB, bar, dst + TC_A_BYTES, kb * 64, n_tile * 128, img);
            else tma_load_2d(&p.tmB, bar, dst + TC_A_BYTES, kb * 64, n_tile * 128);
        };
        auto load_a = [&](int f, uint32_t bar, uint32_t dst) {
            int kb, n_tile, m_tile, img; coords(f, kb, n_tile, m_tile, img);
            tma_load_2d(&p.tmA0, bar, dst, kb * 64, m_tile * 128);
        };
        // the first ring of stages: weight halves before the grid dependency resolves (unless the weights are per-sample matrices
        // the preceding launch wrote), activation halves after
        const bool b_early = !p.w_per_sample;
        for (int f = 0; f < early; ++f) {
            if (elect_one()) {
                mbar_expect_tx(full(f), GS_STAGE_BYTES);
                if (b_early) load_b(f, full(f), base + f * GS_STAGE_BYTES);
            }
            __syncwarp();
        }
        pdl_sync();
        for (int f = 0; f < early; ++f) {
            if (elect_one()) {
                if (!b_early) load_b(f, full(f), base + f * GS_STAGE_BYTES);
                load_a(f, full(f), base + f * GS_STAGE_BYTES);
            }
            __syncwarp();
        }
        int st = 0, round = 1;
        uint32_t sA = base;
        for (int f = early; f < total; ++f) {
            mbar_wait(empty(st), (round - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(full(st), GS_STAGE_BYTES);
                load_a(f, full(st), sA);
                load_b(f, full(st), sA);
            }
            __syncwarp();
            sA += GS_STAGE_BYTES;
            if (++st == GS_STAGES) { st = 0; ++round; sA = base; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        int st = 0, k = 0;
        uint32_t par = 0;
        uint32_t sA = base;
        long long w_full = 0, w_tempty = 0;               // timeline builds: clocks this warp waited for operand stages / a free accumulator
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            const int buf = k & 1;
            if (k >= 2) TL_WAIT(w_tempty, mbar_wait(tempty(buf), ((k >> 1) - 1) & 1));
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)(buf * 128);
            for (int kb = 0; kb < nkb; ++kb) {
                TL_WAIT(w_full, mbar_wait(full(st), par));
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
        if (lane == 0) { tstore(p, 8, w_full); tstore(p, 9, w_tempty); }
    } else {
        // ===== epilogue: 8 warps; thread = (row of the tile, 64-channel half) =====
        constexpr int NGH = CPG_SH > 0 ? (64 >> CPG_SH) : 1;
        const int et = threadIdx.x - 64;
        const int q = warp & 3, hsel = (warp - 2) >> 2;
        const int r = q * 32 + lane, c0 = hsel * 64;
        const bool ln_fold = p.ln_in != nullptr;
        // channel sums of row m for the LayerNorm fold: {sum, sum of squares} over the parts, fetched one item ahead.  The loads are
        // only ISSUED here; the raw values stay in registers and are summed when the next item starts (the empty asm below keeps the
        // compiler from hoisting the adds back up to the loads): summing in the fetch loop made every add wait for its load and the
        // next load for the add -- 4 (8) dependent L2 round trips = 0.9 k (1.3 k) clk per item on the warp that bounds the item
        // rate (scripts/timeline.py: the MMA warp waits for a free accumulator, profiles/timeline_gemm_r02_aw2.txt)
        constexpr int LN_MAXP = 8;
        float2 lnv[LN_MAXP];
        auto ln_fetch = [&](const int64_t m) {
#pragma unroll
            for (int i = 0; i < LN_MAXP; ++i) lnv[i] = make_float2(0.f, 0.f);
            if (ln_fold && m < M) {
                const float2* lp = reinterpret_cast<const float2*>(p.ln_in) + m * p.ln_in_parts;
#pragma unroll
                for (int i = 0; i < LN_MAXP; ++i)
                    if (i < p.ln_in_parts) lnv[i] = __ldg(lp + i);
                for (int i = LN_MAXP; i < p.ln_in_parts; ++i) { const float2 v = __ldg(lp + i); lnv[0].x += v.x; lnv[0].y += v.y; }
            }
        };
        ln_fetch((((int64_t)(blockIdx.x / ntn)) << 7) + r);
        int k = 0;
        long long w_tfull = 0;                            // timeline builds: clocks this warp waited for accumulators
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
            const int n_tile = it % ntn, m_tile = it / ntn;
            const int buf = k & 1, cbase = n_tile * 128;
            const int64_t m = ((int64_t)m_tile << 7) + r;
            const bool valid = m < M;
            const float* par = s_par + cbase;
            float ln_a = 1.f, ln_b = 0.f;
            if (ln_fold) {
#pragma unroll
                for (int i = 0; i < LN_MAXP; ++i) asm volatile("" : "+f"(lnv[i].x), "+f"(lnv[i].y));      // consumed here, not earlier
            }
            if (ln_fold && valid) {
                float nsu = 0.f, nsq = 0.f;
#pragma unroll
                for (int i = 0; i < LN_MAXP; ++i) { nsu += lnv[i].x; nsq += lnv[i].y; }
                const float mean = nsu * p.ln_inv_c;
                ln_a = 1.f / (sqrtf(fmaxf(nsq * p.ln_inv_c - mean * mean, 0.f)) + p.ln_eps);
                ln_b = -mean * ln_a;
            }
            if (it + (int)gridDim.x < n_items) ln_fetch((((int64_t)((it + (int)gridDim.x) / ntn)) << 7) + r);      // next item's rows
            const __nv_bfloat16* resp = (p.residual && valid) ? p.residual + m * p.Cout + cbase + c0 : nullptr;
            uint32_t res[2][8];
            if (resp) ldg_v8(resp, res[0]);
            if (et == 0 && k == 1) tstamp(p, 11);
            TL_WAIT(w_tfull, mbar_wait(tfull(buf), (k >> 1) & 1));
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
                if (valid) stg_v8(op + 16 * qd, ov);       // (holding the four pieces back to store the 128-byte line at once was slower)
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
        if (et == 0) { tstore(p, 15, w_tfull); tstore(p, 20, (long long)k); }
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
