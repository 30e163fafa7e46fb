// LinearAttention context on warp-level tensor-core MMA (bf16 in, fp32 accumulate), fused with the fold of the
// attention's output projection into per-sample matrices (reference: models/unet/blocks.py:118-134).
//
//   k, v : (n, 32) slices of the NHWC qkv tensor for one (b, head)
//   ctx[d][e] = sum_n softmax_n(k)[n][d] * v[n][e]                           (blocks.py:129-130)
//   Mb[b][c][h*32+d] = sum_e Wout[c][h*32+e] * ctx[d][e]                     (blocks.py:131-133 folded)
// so that to_out(attention)[n][c] = sum_k q[n][k] Mb[b][c][k] + bias[c] is one per-sample tcgen05 GEMM (dd_conv_tc
// with DD_TC_W_PER_SAMPLE).  The op is bound by reading k and v once (128 B per pixel and head); the 32x32x n
// contraction runs on mma.sync.m16n8k16 because a 32x32 output per (b, head) is far too small for a tcgen05 tile.
//
// CTA (b, head, split) = 8 warps; a warp owns 32-row slabs of the split's pixel range, keeps an online softmax
// (running max per d, rescaled fp32 accumulators) in registers, operands via ldmatrix.trans from warp-private
// shared-memory slabs, double-buffered with cp.async (the loads of slab i + 1 are in flight under slab i).  Warps merge through shared memory, splits through a workspace + self-resetting arrival
// ticket: the last CTA of a (b, head) normalises the context and does the projection fold, also on mma.sync.
#include "common.cuh"

namespace dd {

constexpr int LA_WS = 64 + 32 * 32;      // floats per (b, head, split): max_d[32], sum_d[32], ctx[32][32]
constexpr int LM_SLAB = 32;              // rows per warp slab
constexpr int LM_PITCH = 40;             // bf16 elements per smem row (80 B: conflict-free ldmatrix)
constexpr int LM_ROWS = 8 * LM_SLAB;     // rows per CTA iteration

// rows per split: a multiple of 256, at least 256, at most 16 splits; the number of splits aims at `target` CTAs in all
// (every split pays an eight-warp merge, a workspace round trip and a ticket: with 256 (b, head) pairs four splits of a 32x32 map
// were 1024 CTAs of ONE 32-row slab per warp -- profiles/README.md, round 2)
inline int lm_target_ctas() {
    static int t = -1;
    if (t < 0) { const char* e = getenv("DD_LM_TARGET_CTAS"); t = e ? atoi(e) : 148; if (t < 1) t = 1; }
    return t;
}
inline int lm_chunk(int n, int bh) {
    // one CTA per (b, head) once those alone fill the SMs; below that three CTAs per SM's worth of splits (sweep: profiles/README.md)
    const int target = bh >= lm_target_ctas() ? lm_target_ctas() : 3 * lm_target_ctas();
    int S = (target + bh - 1) / bh;
    if (S > 16) S = 16;
    if (S < 1) S = 1;
    int c = (n + S - 1) / S;
    c = (c + LM_ROWS - 1) / LM_ROWS * LM_ROWS;
    return c < LM_ROWS ? LM_ROWS : c;
}

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

#ifndef DD_ATTN_TIMELINE
#define DD_ATTN_TIMELINE 0          // -DDD_ATTN_TIMELINE=1: per-CTA clock64 stamps (scripts/timeline_attn.py)
#endif
__device__ long long* g_attn_dbg = nullptr;
__device__ __forceinline__ void astamp(int slot) {
    if (DD_ATTN_TIMELINE && g_attn_dbg && threadIdx.x == 0) g_attn_dbg[(blockIdx.x + gridDim.x * blockIdx.y) * 8 + slot] = clock64();
}

__global__ void __launch_bounds__(256, 3) linattn_ctxmix_kernel(const __nv_bfloat16* __restrict__ qkv, float* __restrict__ ws,
                                                             int* __restrict__ tickets, int n, int heads, int chunk,
                                                             const __nv_bfloat16* __restrict__ Wout, int C,
                                                             __nv_bfloat16* __restrict__ Mb) {
    astamp(0);
    constexpr int DH = 32;
    const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
    const int S = gridDim.y, sp = blockIdx.y;
    const int HD = heads * DH, C3 = 3 * HD;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // the projection weights do not depend on the previous kernel: this head's (C x 32) slice is copied to shared memory
    // (cp.async, 80-byte rows) while the grid dependency is still pending -- in the fold at the end these loads were two
    // exposed L2 round trips (4600 of the 12 400 clk a CTA lives on the 4x4 maps, scripts/timeline_attn.py)
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __nv_bfloat16* s_w = reinterpret_cast<__nv_bfloat16*>(s_dyn);                  // [C][LM_PITCH]
    for (int i = threadIdx.x; i < C * 4; i += 256) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_w + (i >> 2) * LM_PITCH + (i & 3) * 8);
        const __nv_bfloat16* src = Wout + (int64_t)(i >> 2) * HD + hd * DH + (i & 3) * 8;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pdl_sync();
    astamp(1);

    __shared__ __align__(16) __nv_bfloat16 s_kv[8][2][LM_SLAB][LM_PITCH];     // 40 KB, later reused (fold stage)
    __shared__ float s_ctx[DH][DH + 1];
    __shared__ float s_mw[8][DH];
    __shared__ float s_s[DH], s_M[DH];
    __shared__ float s_sw[8][DH];                                            // per-warp softmax denominators for the merge
    __shared__ int s_last;

    const __nv_bfloat16* kb = qkv + (int64_t)b * n * C3 + HD + hd * DH;      // v = k + HD
    const int n_lo = sp * chunk, n_hi = min(n, n_lo + chunk);

    // per-thread state: d rows {g, g+8, g+16, g+24} (index mt*2+half), e columns nt*8 + 2t, +1
    float m_run[4], s_run[4], acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; s_run[i] = 0.f; }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mt][nt][j] = 0.f;

    // The slabs are double-buffered with cp.async: the loads of slab i + 1 are in flight while slab i is processed (one L2 round
    // trip per slab was exposed: 26 k clk for the four slabs per warp of a 32x32 map, scripts/timeline_attn.py).  Buffer 0 is the
    // warp's static slab, buffer 1 sits behind the weight slice in dynamic shared memory.
    __nv_bfloat16* kvb[2] = {&s_kv[warp][0][0][0],
                             reinterpret_cast<__nv_bfloat16*>(s_dyn) + (size_t)C * LM_PITCH + (size_t)warp * 2 * LM_SLAB * LM_PITCH};
    auto issue_slab = [&](const int r0, __nv_bfloat16* dstb) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = lane + 32 * u, isv = idx >> 7, r = (idx & 127) >> 2, c16 = idx & 3;
            __nv_bfloat16* d = dstb + (isv * LM_SLAB + r) * LM_PITCH + c16 * 8;
            if (r0 + r < n_hi) {
                const uint32_t da = (uint32_t)__cvta_generic_to_shared(d);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(kb + (int64_t)(r0 + r) * C3 + isv * HD + c16 * 8) : "memory");
            } else {
                *reinterpret_cast<uint4*>(d) = isv ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // k = -inf
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cur = 0;
    if (n_lo + warp * LM_SLAB < n_hi) issue_slab(n_lo + warp * LM_SLAB, kvb[0]);
    for (int r0 = n_lo + warp * LM_SLAB; r0 < n_hi; r0 += LM_ROWS, cur ^= 1) {
        if (r0 + LM_ROWS < n_hi) {
            issue_slab(r0 + LM_ROWS, kvb[cur ^ 1]);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        __nv_bfloat16 (*sk)[LM_PITCH] = reinterpret_cast<__nv_bfloat16 (*)[LM_PITCH]>(kvb[cur]);
        __nv_bfloat16 (*sv)[LM_PITCH] = reinterpret_cast<__nv_bfloat16 (*)[LM_PITCH]>(kvb[cur] + LM_SLAB * LM_PITCH);
        // ---- A = exp(k - m)^T fragments (d x n): ldmatrix.trans of the [n][d] slab
        uint32_t a[2][2][4];
        const int mi = lane >> 3, li = lane & 7;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
                ldsm_x4_t(a[mt][ks], &sk[ks * 16 + li + (mi >> 1) * 8][mt * 16 + (mi & 1) * 8]);
        float fac[4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float mx = -INFINITY;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float2 f = unpack_bf2(a[mt][ks][h + 2 * q]);
                        mx = fmaxf(mx, fmaxf(f.x, f.y));
                    }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const int i = mt * 2 + h;
                const float mn = fmaxf(m_run[i], mx);            // finite: the slab has at least one valid row
                fac[i] = __expf(m_run[i] - mn);                  // 0 on the first slab
                m_run[i] = mn;
                float ssum = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float2 f = unpack_bf2(a[mt][ks][h + 2 * q]);
                        const uint32_t p = pack_bf2(__expf(f.x - mn), __expf(f.y - mn));
                        const float2 pr = unpack_bf2(p);         // the sum uses the rounded weights the MMA sees
                        ssum += pr.x + pr.y;
                        a[mt][ks][h + 2 * q] = p;
                    }
                s_run[i] = s_run[i] * fac[i] + ssum;
            }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                acc[mt][nt][0] *= fac[mt * 2]; acc[mt][nt][1] *= fac[mt * 2];
                acc[mt][nt][2] *= fac[mt * 2 + 1]; acc[mt][nt][3] *= fac[mt * 2 + 1];
            }
        // ---- B = v fragments (n x e) and the MMAs
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t bb[4];
                ldsm_x4_t(bb, &sv[ks * 16 + li + (mi & 1) * 8][np * 16 + (mi >> 1) * 8]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16(acc[mt][np * 2], a[mt][ks], bb[0], bb[1]);
                    mma_bf16(acc[mt][np * 2 + 1], a[mt][ks], bb[2], bb[3]);
                }
            }
        __syncwarp();                     // this buffer is refilled by the next iteration's cp.async
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s_run[i] += __shfl_xor_sync(0xffffffffu, s_run[i], 1);
        s_run[i] += __shfl_xor_sync(0xffffffffu, s_run[i], 2);
    }
    astamp(2);
    __nv_bfloat16 (*s_cb)[LM_PITCH] = reinterpret_cast<__nv_bfloat16 (*)[LM_PITCH]>(&s_kv[0][0][0][0]);      // bf16 [d][e] for the fold
    const bool single = (S == 1) && (n_hi - n_lo <= LM_SLAB);         // only warp 0 had rows (4x4 maps): nothing to merge
    if (single) {
        if (warp == 0) {
            __syncwarp();                                               // the slab's ldmatrix reads are done: s_cb aliases it
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float i0 = 1.f / s_run[mt * 2], i1 = 1.f / s_run[mt * 2 + 1];
                    *reinterpret_cast<uint32_t*>(&s_cb[mt * 16 + g][nt * 8 + 2 * t]) = pack_bf2(acc[mt][nt][0] * i0, acc[mt][nt][1] * i0);
                    *reinterpret_cast<uint32_t*>(&s_cb[mt * 16 + g + 8][nt * 8 + 2 * t]) = pack_bf2(acc[mt][nt][2] * i1, acc[mt][nt][3] * i1);
                }
        }
    } else {
    // ---- merge the 8 warps: CTA max, rescale, per-warp partials parked in shared memory, summed by all threads
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_mw[warp][g + 8 * i] = m_run[i];
    }
    __syncthreads();
    float f_own[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, s_mw[w][g + 8 * i]);
        f_own[i] = (m_run[i] == -INFINITY) ? 0.f : __expf(m_run[i] - M);       // a warp without rows contributes nothing
        if (warp == 0 && t == 0) s_M[g + 8 * i] = M;
    }
    // each warp parks its rescaled partial in its own (now idle) slab as [32][40] floats, 8-byte stores (conflict free per
    // half warp; the first version's scalar stores at pitch 33 were 4-way conflicted), then a tree-free sum over the 8 warps
    // (shared-memory float atomics would be CAS loops under 8-way contention)
    float* my = reinterpret_cast<float*>(&s_kv[warp][0][0][0]);          // 5120 B per warp = 32 * 40 * 4
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<float2*>(&my[(mt * 16 + g) * 40 + nt * 8 + 2 * t]) =
                make_float2(acc[mt][nt][0] * f_own[mt * 2], acc[mt][nt][1] * f_own[mt * 2]);
            *reinterpret_cast<float2*>(&my[(mt * 16 + g + 8) * 40 + nt * 8 + 2 * t]) =
                make_float2(acc[mt][nt][2] * f_own[mt * 2 + 1], acc[mt][nt][3] * f_own[mt * 2 + 1]);
        }
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_sw[warp][g + 8 * i] = s_run[i] * f_own[i];
    }
    __syncthreads();
    {
        constexpr int WSTRIDE = 2 * LM_SLAB * LM_PITCH / 2;                 // floats between warp regions
        const float* w0 = reinterpret_cast<const float*>(&s_kv[0][0][0][0]);
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < 8; ++w)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = threadIdx.x + 256 * k;
                part[k] += w0[w * WSTRIDE + (i >> 5) * 40 + (i & 31)];
            }
        float ssum = 0.f;
        if (threadIdx.x < DH) {
#pragma unroll
            for (int w = 0; w < 8; ++w) ssum += s_sw[w][threadIdx.x];
        }
        __syncthreads();                                                 // slab memory is reused below
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = threadIdx.x + 256 * k;
            s_ctx[i >> 5][i & 31] = part[k];
        }
        if (threadIdx.x < DH) s_s[threadIdx.x] = ssum;
    }
    __syncthreads();

    astamp(3);
    // normalised context as bf16 [d][e] (pitch 40) for the fold; reuses the slab memory
    float* s_fac = reinterpret_cast<float*>(&s_kv[1][0][0][0]);              // [S][32] split factors, then 1/total at [16][32]
    if (S == 1) {
        for (int i = threadIdx.x; i < DH * DH; i += 256) {
            const int d = i >> 5, e = i & 31;
            s_cb[d][e] = __float2bfloat16_rn(s_ctx[d][e] / s_s[d]);
        }
    } else {
        float* w = ws + ((int64_t)bh * S + sp) * LA_WS;
        if (threadIdx.x < DH) { w[threadIdx.x] = s_M[threadIdx.x]; w[32 + threadIdx.x] = s_s[threadIdx.x]; }
        for (int i = threadIdx.x; i < DH * DH; i += 256) w[64 + i] = s_ctx[i >> 5][i & 31];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();                              // release: cumulative over the CTA's writes ordered by the barrier
            const int tk = atomicAdd(tickets + bh, 1);
            s_last = (tk == S - 1);
            if (s_last) tickets[bh] = 0;                  // self-resetting: zero again for the next launch
            __threadfence();                              // acquire side for the reads below
        }
        __syncthreads();
        if (!s_last) {
            asm volatile("cp.async.wait_all;" ::: "memory");        // no copy may be in flight when the CTA retires
            return;
        }
        const float* w0 = ws + (int64_t)bh * S * LA_WS;
        if (threadIdx.x < DH) {
            const int d = threadIdx.x;
            float M = -INFINITY;
            for (int s = 0; s < S; ++s) M = fmaxf(M, __ldcg(w0 + s * LA_WS + d));
            float tot = 0.f;
            for (int s = 0; s < S; ++s) {
                const float f = __expf(__ldcg(w0 + s * LA_WS + d) - M);
                s_fac[s * DH + d] = f;
                tot += __ldcg(w0 + s * LA_WS + 32 + d) * f;
            }
            s_fac[16 * DH + d] = 1.f / tot;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < DH * DH; i += 256) {
            const int d = i >> 5, e = i & 31;
            float o = 0.f;
            for (int s = 0; s < S; ++s) o += __ldcg(w0 + s * LA_WS + 64 + i) * s_fac[s * DH + d];
            s_cb[d][e] = __float2bfloat16_rn(o * s_fac[16 * DH + d]);
        }
    }
    }   // !single
    asm volatile("cp.async.wait_group 0;" ::: "memory");       // this thread's share of the weight slice has landed
    __syncthreads();

    astamp(4);
    // ---- projection fold: Mb[c][hd*32 + d] = sum_e Wout[c][hd*32 + e] * ctx[d][e]   (M = c, N = d, K = e)
    uint32_t bf[2][4][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            bf[ks][nt][0] = *reinterpret_cast<const uint32_t*>(&s_cb[nt * 8 + g][ks * 16 + 2 * t]);
            bf[ks][nt][1] = *reinterpret_cast<const uint32_t*>(&s_cb[nt * 8 + g][ks * 16 + 2 * t + 8]);
        }
    __nv_bfloat16* mb = Mb + (int64_t)b * C * HD + hd * DH;
    for (int c0 = warp * 16; c0 < C; c0 += 128) {
        uint32_t aw[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const __nv_bfloat16* p0 = s_w + (c0 + g) * LM_PITCH + ks * 16 + 2 * t;
            const __nv_bfloat16* p1 = p0 + 8 * LM_PITCH;
            aw[ks][0] = *reinterpret_cast<const uint32_t*>(p0);
            aw[ks][1] = *reinterpret_cast<const uint32_t*>(p1);
            aw[ks][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
            aw[ks][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16(o, aw[0], bf[0][nt][0], bf[0][nt][1]);
            mma_bf16(o, aw[1], bf[1][nt][0], bf[1][nt][1]);
            *reinterpret_cast<uint32_t*>(mb + (int64_t)(c0 + g) * HD + nt * 8 + 2 * t) = pack_bf2(o[0], o[1]);
            *reinterpret_cast<uint32_t*>(mb + (int64_t)(c0 + g + 8) * HD + nt * 8 + 2 * t) = pack_bf2(o[2], o[3]);
        }
    }
    astamp(5);
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_debug_set_attn_timeline(long long* buf) {
    cudaError_t e = cudaMemcpyToSymbol(g_attn_dbg, &buf, sizeof(buf));
    return e == cudaSuccess ? DD_OK : DD_ERR_CUDA;
}

int64_t dd_linattn_mix_ws_floats(int B, int n, int heads) {
    const int chunk = lm_chunk(n, B * heads);
    return (int64_t)B * heads * ((n + chunk - 1) / chunk) * LA_WS + (int64_t)B * heads;
}

int dd_linattn_mix(const void* qkv, int dtype, int B, int n, int heads, int dh, float* ws, int64_t ws_floats,
                   const void* Wout_bf16, int C, void* Mb_bf16, void* stream) {
    DD_REQUIRE(dtype == DD_BF16, "linattn_mix: bf16 tensor-core path only (dtype %d)", dtype);
    DD_REQUIRE(dh == 32 && heads > 0 && n > 0, "linattn_mix: dim_head must be 32 (got %d)", dh);
    DD_REQUIRE(C > 0 && C % 16 == 0, "linattn_mix: C=%d must be a multiple of 16", C);
    const int chunk = lm_chunk(n, B * heads);
    const int S = (n + chunk - 1) / chunk;
    DD_REQUIRE(ws != nullptr && ws_floats >= dd_linattn_mix_ws_floats(B, n, heads), "linattn_mix: workspace too small");
    int* tickets = reinterpret_cast<int*>(ws + (int64_t)B * heads * S * LA_WS);
    // this head's (C x 32) projection slice, 80-byte rows + the second k / v slab of every warp
    const size_t dyn = (size_t)C * LM_PITCH * sizeof(__nv_bfloat16) + (size_t)8 * 2 * LM_SLAB * LM_PITCH * sizeof(__nv_bfloat16);
    DD_REQUIRE(dyn <= 136 * 1024, "linattn_mix: C=%d too large for the shared-memory weight slice", C);
    static size_t dyn_set = 0;
    if (dyn > dyn_set) {
        cudaError_t e = cudaFuncSetAttribute(linattn_ctxmix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) { set_error("linattn_mix: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
        dyn_set = dyn;
    }
    launch_pdl(linattn_ctxmix_kernel, dim3(B * heads, S), dim3(256), dyn, (cudaStream_t)stream, (const __nv_bfloat16*)qkv, ws,
               tickets, n, heads, chunk, (const __nv_bfloat16*)Wout_bf16, C, (__nv_bfloat16*)Mb_bf16);
    return check_launch("linattn_mix");
}

}  // extern "C"
