// =============================================================================================
// Persistent TF32 convolution for the fp32 training programs (dd_conv_tc32).
// The resampling nets run 3x3 32->32 / 1x1 32<->64 convolutions over 2 M pixels: 16 384 tiles of 128 pixels with nine
// (or one, two) k-blocks each.  One CTA per tile spends ~4.6 of its 6 us in launch, TMEM allocation, barrier setup and
// the first load's latency (launch list: 340 us per conv against an 84 us HBM floor), so here a CTA walks tiles
// t = blockIdx.x, += gridDim.x with the operand ring running across tile boundaries and TWO accumulator buffers in TMEM:
// the epilogue of tile i (TMEM -> registers -> fp32 NHWC, bias / addend fused) overlaps the loads and MMAs of tile i+1.
// Warps: 0 = A-operand TMA, 6 = weight TMA, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
// =============================================================================================
#include "conv_tc_common.cuh"

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

using namespace dd;

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
