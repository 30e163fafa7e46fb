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
#include "conv_tc_common.cuh"

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

using namespace dd;

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
