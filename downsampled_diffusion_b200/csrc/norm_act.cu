// GroupNorm(+Mish,+time bias,+residual), channel LayerNorm, time-embedding MLPs and the
// LinearAttention core of the U-Net (models/unet/blocks.py).  All HBM/L2-bound or tiny; fp32 math.
#include "common.cuh"

namespace dd {

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: one CTA per (b, group).  Two-pass (mean, then centred sum of squares) in
// fp32 with a double block reduction: the validation mode needs 1e-4 on eps_hat.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x, int HW, int C, int G, float eps, float* __restrict__ stats) {
    pdl_sync();
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G;
    const T* xb = x + (int64_t)b * HW * C + g * cpg;
    const int n = HW * cpg;
    __shared__ double red[32];
    __shared__ float s_mean;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += to_f(xb[(int64_t)(i / cpg) * C + (i % cpg)]);
    double v = (double)warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        s_mean = (float)(s / n);
    }
    __syncthreads();
    const float mean = s_mean;
    acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float d = to_f(xb[(int64_t)(i / cpg) * C + (i % cpg)]) - mean;
        acc += d * d;
    }
    v = (double)warp_sum(acc);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        stats[2 * blockIdx.x] = mean;
        stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(s / n + (double)eps));
    }
}

// ---------------------------------------------------------------------------------------------
// y = mish(gn(x)*gamma+beta) [+ tbias[row(b), c]] [+ residual].  One 16-byte vector (8 bf16 / 4 fp32) per
// thread, every operand fetched with 16-byte loads.  Algorithmic bytes: read x + write y (+ residual read)
// = 4 (+2) B/element in bf16.
// ---------------------------------------------------------------------------------------------
template <int N> struct FVec { float v[N]; };
template <int N> __device__ __forceinline__ FVec<N> ldg_f(const float* p) {
    FVec<N> r;
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}

constexpr int GN_UNR = 4;
// grid = (vector blocks within one image, B): no division to find the image, 32-bit index math only
// (the first version spent ~25 of its 36 instructions per element on 64-bit index divisions -- profiles/README.md).
template <typename T>
__global__ void __launch_bounds__(256, 4) gn_mish_kernel(const T* __restrict__ x, T* __restrict__ y, int HW, int C, int G,
                               const float* __restrict__ stats, int stats_mode, float eps,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ tbias, int tb_stride, const int32_t* __restrict__ trow,
                               int trow_stride, const T* __restrict__ residual, int vec_per_img) {
    pdl_sync();
    constexpr int VN = Vec<T>::N;
    constexpr int UNR = GN_UNR;                      // independent 16-byte vectors per thread: loads first, math after
    const int b = blockIdx.y;
    const unsigned cv = (unsigned)C / VN, cpg = (unsigned)C / G;   // cv divides 256: a thread's vectors share their channels
    const unsigned i0 = blockIdx.x * (256u * UNR) + threadIdx.x;
    if (i0 >= (unsigned)vec_per_img) return;
    const int c = (int)(i0 % cv) * VN;
    const int g = c / (int)cpg;
    const int64_t base = (int64_t)b * vec_per_img;
    uint4 xv[UNR], rv[UNR];                          // raw 16-byte vectors: 64 (+64) bytes in flight per thread
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const unsigned i = i0 + u * 256u;
        if (i < (unsigned)vec_per_img) {
            xv[u] = *reinterpret_cast<const uint4*>(x + (base + i) * VN);
            if (residual) rv[u] = *reinterpret_cast<const uint4*>(residual + (base + i) * VN);
        }
    }
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + (int64_t)b * G + g);
    float mean, rstd;
    if (stats_mode == 0) { mean = st.x; rstd = st.y; }
    else {
        const float inv_n = 1.f / ((float)HW * (float)cpg);
        mean = st.x * inv_n;
        rstd = rsqrtf(fmaxf(st.y * inv_n - mean * mean, 0.f) + eps);
    }
    FVec<VN> sc = ldg_f<VN>(gamma + c), sh = ldg_f<VN>(beta + c), tb;
#pragma unroll
    for (int j = 0; j < VN; ++j) { sc.v[j] *= rstd; sh.v[j] -= mean * sc.v[j]; }     // h = x * sc + sh
    if (tbias) {
        const int row = trow ? trow[(int64_t)b * trow_stride] : b;
        tb = ldg_f<VN>(tbias + (int64_t)row * tb_stride + c);
    } else {
#pragma unroll
        for (int j = 0; j < VN; ++j) tb.v[j] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const unsigned i = i0 + u * 256u;
        if (i >= (unsigned)vec_per_img) break;
        Vec<T> v, r;
        v.from_raw(xv[u]);
        if (residual) r.from_raw(rv[u]);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
            float h = mish_t<T>(fmaf(v.v[j], sc.v[j], sh.v[j])) + tb.v[j];
            if (residual) h += r.v[j];
            v.v[j] = h;
        }
        *reinterpret_cast<uint4*>(y + (base + i) * VN) = v.to_raw();
    }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm + Mish over the partial sums of a split-K convolution (dd_conv_tc with DD_TC_SPLITK), for the
// low-resolution maps (one CTA per image, HW*C <= 16384 values staged in shared memory):
//   x = sum_s part[s] + bias -> {mean, rstd} per group (block reduction) -> mish(x_hat*gamma+beta) + tbias + residual.
// Thread t owns channels 4*(t % (C/4)).. of the pixels t/(C/4), +256/(C/4), ...: one group per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_mish_sum_kernel(const float* __restrict__ part, int S, int64_t split_stride,
                                                          const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                          int HW, int C, int G, float eps,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ tbias, int tb_stride,
                                                          const int32_t* __restrict__ trow, int trow_stride,
                                                          const __nv_bfloat16* __restrict__ residual, float* __restrict__ ln_part) {
    extern __shared__ float4 s_x[];                        // this CTA's summed vectors: HW * (C/4) / gridDim.y
    __shared__ float s_stat[64][2];
    // grid (B, parts): a CTA owns C/parts channels (whole groups) of one image -- one CTA per image left the 4x4 layers with
    // 64 CTAs on 148 SMs and every load latency exposed
    const int b = blockIdx.x, cv = C >> 2, cvp = cv / gridDim.y, nloc = HW * cvp;
    const int c4 = blockIdx.y * cvp + threadIdx.x % cvp, c = c4 * 4, cpg = C / G, g = c / cpg;
    if (threadIdx.x < 2 * G) (&s_stat[0][0])[threadIdx.x] = 0.f;
    // Only the split-K partials come from the preceding launch: the parameters, the time bias (its row index was written at the end
    // of the previous step) and the residual (an earlier launch's output) are fetched while the grid dependency is still pending.
    const float4* src = reinterpret_cast<const float4*>(part) + (int64_t)b * HW * cv;
    const float4 bi = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 tb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tbias) {
        const int row = trow ? trow[(int64_t)b * trow_stride] : b;
        tb = __ldg(reinterpret_cast<const float4*>(tbias + (int64_t)row * tb_stride) + c4);
    }
    uint2 res0 = make_uint2(0u, 0u);                        // residual of the thread's first vector (the only one on the 4x4 maps)
    if (residual && threadIdx.x < nloc)
        res0 = *reinterpret_cast<const uint2*>(residual + ((int64_t)b * HW * cv + (threadIdx.x / cvp) * cv + c4) * 4);
    __syncthreads();
    pdl_sync();
    float sum = 0.f, sq = 0.f;
    for (int vl = threadIdx.x; vl < nloc; vl += 256) {
        const int v = (vl / cvp) * cv + c4;                 // vector index inside the image
        float4 a = bi;
        for (int s = 0; s < S; ++s) {
            const float4 t = __ldcg(src + s * (split_stride >> 2) + v);
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        s_x[vl] = a;
        sum += a.x + a.y + a.z + a.w;
        sq += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    // lanes of one group are cpg/4 consecutive lanes (a power of two <= 32)
    const int gl = cpg >> 2;
    for (int o = 1; o < gl; o <<= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if ((threadIdx.x & (gl - 1)) == 0) { atomicAdd(&s_stat[g][0], sum); atomicAdd(&s_stat[g][1], sq); }
    __syncthreads();
    const float inv_n = 1.f / ((float)HW * (float)cpg);
    const float mean = s_stat[g][0] * inv_n;
    const float rstd = rsqrtf(fmaxf(s_stat[g][1] * inv_n - mean * mean, 0.f) + eps);
    const float sc[4] = {ga.x * rstd, ga.y * rstd, ga.z * rstd, ga.w * rstd};
    const float sh[4] = {be.x - mean * sc[0], be.y - mean * sc[1], be.z - mean * sc[2], be.w - mean * sc[3]};
    const float tbv[4] = {tb.x, tb.y, tb.z, tb.w};
    for (int vl = threadIdx.x; vl < nloc; vl += 256) {
        const int v = (vl / cvp) * cv + c4;
        const float4 a = s_x[vl];
        const float xin[4] = {a.x, a.y, a.z, a.w};
        float h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = mish_fast(fmaf(xin[j], sc[j], sh[j])) + tbv[j];
        const int64_t off = ((int64_t)b * HW * cv + v) * 4;
        if (residual) {
            const uint2 rr = vl == (int)threadIdx.x ? res0 : *reinterpret_cast<const uint2*>(residual + off);
            h[0] += __uint_as_float(rr.x << 16); h[1] += __uint_as_float(rr.x & 0xffff0000u);
            h[2] += __uint_as_float(rr.y << 16); h[3] += __uint_as_float(rr.y & 0xffff0000u);
        }
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(h[0], h[1]);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(h[2], h[3]);
        *reinterpret_cast<uint2*>(y + off) = o;
        if (ln_part) {
            // channel-LayerNorm statistics of the pixel over this CTA's channels, on the bf16 values just written: the cvp lanes
            // of a pixel are consecutive (cvp a power of two <= 32, checked by the host)
            const float x0 = __uint_as_float(o.x << 16), x1 = __uint_as_float(o.x & 0xffff0000u);
            const float x2 = __uint_as_float(o.y << 16), x3 = __uint_as_float(o.y & 0xffff0000u);
            float ls = (x0 + x1) + (x2 + x3), lq = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, x3 * x3)));
            const unsigned lane = threadIdx.x & 31u;
            const unsigned gm = cvp >= 32 ? 0xffffffffu : (((1u << cvp) - 1u) << (lane & ~(unsigned)(cvp - 1)));
            for (int d = cvp >> 1; d > 0; d >>= 1) {
                ls += __shfl_xor_sync(gm, ls, d);
                lq += __shfl_xor_sync(gm, lq, d);
            }
            if ((threadIdx.x & (cvp - 1)) == 0)
                *reinterpret_cast<float2*>(ln_part + (((int64_t)b * HW + vl / cvp) * gridDim.y + blockIdx.y) * 2) = make_float2(ls, lq);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Channel LayerNorm (blocks.py:57-60): one warp per pixel, eps added to the std.
// ---------------------------------------------------------------------------------------------
template <typename T, int PER_LANE>
__global__ void layernorm_c_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t P, int C,
                                   const float* __restrict__ g, const float* __restrict__ bta, float eps) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < P; p += nwarps) {
        const T* xp = x + p * C;
        float v[PER_LANE];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) { v[j] = to_f(xp[lane * PER_LANE + j]); s += v[j]; }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) { float d = v[j] - mean; q += d * d; }
        const float stdv = sqrtf(warp_sum(q) / (float)C);
        const float inv = 1.f / (stdv + eps);
        T* yp = y + p * C;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const int c = lane * PER_LANE + j;
            yp[c] = from_f<T>((v[j] - mean) * inv * g[c] + bta[c]);
        }
    }
}

// bf16 fast path: 16-byte loads, C/8 lanes per pixel (8, 16 or 32), several pixels per warp.
template <int LPP>
__global__ void __launch_bounds__(256) layernorm_c_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                               int64_t P, const float* __restrict__ g,
                                                               const float* __restrict__ bta, float eps) {
    pdl_sync();
    constexpr int C = LPP * 8, PPW = 32 / LPP;       // pixels per warp
    const int lane = threadIdx.x & 31, sub = lane / LPP, l = lane % LPP;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t p = warp * PPW + sub;
    const bool ok = p < P;
    Vec<__nv_bfloat16> v;
    if (ok) v.load(x + p * C + l * 8);
    else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v.v[j];
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { v.v[j] -= mean; q += v.v[j] * v.v[j]; }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float inv = 1.f / (sqrtf(q * (1.f / C)) + eps);
    const FVec<8> gg = ldg_f<8>(g + l * 8), bb = ldg_f<8>(bta + l * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = v.v[j] * inv * gg.v[j] + bb.v[j];
    if (ok) v.store(y + p * C + l * 8);
}

// ---------------------------------------------------------------------------------------------
// time embedding -> per-block channel biases.  grid (ceil(J/128), R); the two small MLP layers are
// recomputed per CTA (131k MAC), then each CTA produces 128 of the J outputs.
// ---------------------------------------------------------------------------------------------
__global__ void time_bias_kernel(const float* __restrict__ t, int dim, const float* __restrict__ freq,
                                 const float* __restrict__ W1,
                                 const float* __restrict__ b1, const float* __restrict__ W2,
                                 const float* __restrict__ b2, const float* __restrict__ Wcat,
                                 const float* __restrict__ bcat, int J, float* __restrict__ out) {
    pdl_sync();
    extern __shared__ float sm[];
    float* emb = sm;              // dim
    float* h1 = emb + dim;        // 4*dim
    float* act = h1 + 4 * dim;    // dim  = mish(temb)
    const int r = blockIdx.y;
    const int half = dim / 2;
    const float tv = t[r];
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float a = __fmul_rn(tv, freq[i]);
        emb[i] = sinf(a);
        emb[half + i] = cosf(a);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = warp; o < 4 * dim; o += nw) {
        const float* w = W1 + (int64_t)o * dim;
        float s = 0.f;
        for (int k = lane; k < dim; k += 32) s += w[k] * emb[k];
        s = warp_sum(s);
        if (lane == 0) h1[o] = mish_f(s + b1[o]);
    }
    __syncthreads();
    for (int o = warp; o < dim; o += nw) {
        const float* w = W2 + (int64_t)o * 4 * dim;
        float s = 0.f;
        for (int k = lane; k < 4 * dim; k += 32) s += w[k] * h1[k];
        s = warp_sum(s);
        if (lane == 0) act[o] = mish_f(s + b2[o]);
    }
    __syncthreads();
    const int j0 = blockIdx.x * 128;
    for (int o = j0 + warp; o < min(j0 + 128, J); o += nw) {
        const float* w = Wcat + (int64_t)o * dim;
        float s = 0.f;
        for (int k = lane; k < dim; k += 32) s += w[k] * act[k];
        s = warp_sum(s);
        if (lane == 0) out[(int64_t)r * J + o] = s + bcat[o];
    }
}

// ---------------------------------------------------------------------------------------------
// LinearAttention core (blocks.py:128-133), dh = 32, split over the spatial axis so that (b, head) pairs
// with many pixels are spread over several CTAs:
//   ctx kernel : CTA (b, head, split) streams its n-range in 128-row tiles with an online softmax
//                (running max per d), accumulating un-normalised ctx[d][e] = sum_n exp(k[d,n]-m_d) v[e,n]
//                and s_d = sum_n exp(k[d,n]-m_d); writes {m, s, ctx} (1088 floats) to the workspace.
//   out kernel : CTA (b, head, split) merges the S partials (max-rescale), normalises, then
//                out[n][e] = sum_d ctx[d][e] q[n][d] for its n-range (ctx column in registers, q via shuffles).
// ---------------------------------------------------------------------------------------------
constexpr int LA_TILE = 128;
constexpr int LA_WS = 64 + 32 * 32;      // floats per (b, head, split)

__host__ __device__ inline int la_chunk(int n) {
    int c = (n + 15) / 16;                                  // at most 16 splits
    c = (c + LA_TILE - 1) / LA_TILE * LA_TILE;
    return c < LA_TILE ? LA_TILE : c;
}

template <typename T>
__global__ void __launch_bounds__(256) linattn_ctx_kernel(const T* __restrict__ qkv, float* __restrict__ ws,
                                                          int n, int heads, int chunk) {
    pdl_sync();
    constexpr int DH = 32;
    constexpr int VN = Vec<T>::N;                 // 8 (bf16) / 4 (fp32) channels per 16-byte load
    constexpr int VPR = DH / VN;                  // vectors per row of one matrix
    const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
    const int S = gridDim.y, sp = blockIdx.y;
    const int C3 = 3 * heads * DH;
    const T* kb = qkv + (int64_t)b * n * C3 + heads * DH + hd * DH;
    const T* vb = kb + heads * DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float sk[LA_TILE][DH + 1];                    // lane-varying d: conflict free
    __shared__ __align__(16) float sv[LA_TILE][DH + 4];      // 16-byte aligned rows: float4 broadcasts
    __shared__ float red[8][DH];
    __shared__ float s_m[DH], s_scale[DH];
    const int n_lo = sp * chunk, n_hi = min(n, n_lo + chunk);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float s_part = 0.f;                 // partial sum over this thread's rows, d = lane
    if (threadIdx.x < DH) s_m[threadIdx.x] = -INFINITY;
    __syncthreads();
    for (int t0 = n_lo; t0 < n_hi; t0 += LA_TILE) {
        const int rows = min(LA_TILE, n_hi - t0);
        {   // all global loads of the tile are issued before the first use (the kernel is load-latency bound)
            constexpr int NV = LA_TILE * VPR * 2 / 256;          // vectors per thread for a full tile
            Vec<T> xv[NV];
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int i = threadIdx.x + u * 256;
                const int r = i / (VPR * 2), w = i % (VPR * 2);
                const int isv = w / VPR, c = (w % VPR) * VN;
                if (r < rows) xv[u].load((isv ? vb : kb) + (int64_t)(t0 + r) * C3 + c);
            }
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int i = threadIdx.x + u * 256;
                const int r = i / (VPR * 2), w = i % (VPR * 2);
                const int isv = w / VPR, c = (w % VPR) * VN;
                if (r < rows) {
#pragma unroll
                    for (int j = 0; j < VN; ++j) (isv ? sv[r][c + j] : sk[r][c + j]) = xv[u].v[j];
                }
            }
        }
        __syncthreads();
        float m = -INFINITY;
        for (int r = warp; r < rows; r += 8) m = fmaxf(m, sk[r][lane]);
        red[warp][lane] = m;
        __syncthreads();
        if (warp == 0) {
            float mm = s_m[lane];
#pragma unroll
            for (int w = 0; w < 8; ++w) mm = fmaxf(mm, red[w][lane]);
            s_scale[lane] = __expf(s_m[lane] - mm);      // 0 on the first tile (exp(-inf))
            s_m[lane] = mm;
        }
        __syncthreads();
        const float mcur = s_m[lane], sc = s_scale[lane];
        s_part *= sc;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] *= sc;
        for (int r = warp; r < rows; r += 8) {
            const float e = __expf(sk[r][lane] - mcur);
            sk[r][lane] = e;
            s_part += e;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < rows; ++r) {
            const float pe = sk[r][lane];
            const float4 vv = *reinterpret_cast<const float4*>(&sv[r][warp * 4]);
            acc[0] += pe * vv.x; acc[1] += pe * vv.y; acc[2] += pe * vv.z; acc[3] += pe * vv.w;
        }
        __syncthreads();
    }
    float* w = ws + ((int64_t)bh * S + sp) * LA_WS;
    red[warp][lane] = s_part;
    __syncthreads();
    if (warp == 0) {
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) ss += red[k][lane];
        w[lane] = s_m[lane];
        w[32 + lane] = ss;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) w[64 + lane * 32 + warp * 4 + j] = acc[j];
}

template <typename T>
__global__ void __launch_bounds__(256) linattn_out_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                          const float* __restrict__ ws, int n, int heads, int chunk) {
    pdl_sync();
    constexpr int DH = 32;
    const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
    const int S = gridDim.y, sp = blockIdx.y;
    const int C3 = 3 * heads * DH, CO = heads * DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float s_ctx[DH][DH + 1];
    // merge the S partials: thread (d = lane, e block = warp)
    const float* w0 = ws + (int64_t)bh * S * LA_WS;
    float M = -INFINITY;
    for (int s = 0; s < S; ++s) M = fmaxf(M, w0[s * LA_WS + lane]);
    float tot = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < S; ++s) {
        const float* w = w0 + s * LA_WS;
        const float f = __expf(w[lane] - M);
        tot += w[32 + lane] * f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += w[64 + lane * 32 + warp * 4 + j] * f;
    }
    const float inv = 1.f / tot;
#pragma unroll
    for (int j = 0; j < 4; ++j) s_ctx[lane][warp * 4 + j] = acc[j] * inv;
    __syncthreads();
    float ctx[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) ctx[d] = s_ctx[d][lane];     // lane = e from here on
    const T* qb = qkv + (int64_t)b * n * C3 + hd * DH;
    T* ob = out + (int64_t)b * n * CO + hd * DH;
    const int n_lo = sp * chunk, n_hi = min(n, n_lo + chunk);
    for (int i = n_lo + warp; i < n_hi; i += 8) {
        const float qv = to_f(qb[(int64_t)i * C3 + lane]);
        float o = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) o += ctx[d] * __shfl_sync(0xffffffffu, qv, d);
        ob[(int64_t)i * CO + lane] = from_f<T>(o);
    }
}

static inline int grid_cap(int64_t n, int threads) {
    int64_t g = (n + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * 8;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_time_bias(const float* t, int R, int dim, const float* freq, const float* W1, const float* b1, const float* W2, const float* b2,
                 const float* Wcat, const float* bcat, int J, float* out, void* stream) {
    DD_REQUIRE(R > 0 && dim >= 4 && dim % 2 == 0 && J > 0 && dim <= 2048, "time_bias: bad sizes R=%d dim=%d J=%d", R, dim, J);
    dim3 grid((J + 127) / 128, R);
    size_t smem = (size_t)6 * dim * sizeof(float);
    launch_pdl(time_bias_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, t, dim, freq, W1, b1, W2, b2, Wcat, bcat, J, out);
    return check_launch("time_bias");
}

int dd_gn_stats(const void* x, int dtype, int B, int HW, int C, int G, float eps, float* stats, void* stream) {
    DD_REQUIRE(C % G == 0 && B > 0 && HW > 0, "gn_stats: C=%d not divisible by G=%d", C, G);
    DD_DISPATCH_DTYPE(dtype, T, (launch_pdl(gn_stats_kernel<T>, dim3(B * G), dim3(256), 0, (cudaStream_t)stream, (const T*)x, HW, C, G, eps, stats)));
    return check_launch("gn_stats");
}

int dd_gn_mish(const void* x, void* y, int dtype, int B, int HW, int C, int G, const float* stats, int stats_mode,
               float eps, const float* gamma, const float* beta, const float* tbias, int tb_stride,
               const int32_t* trow, int trow_stride, const void* residual, void* stream) {
    DD_REQUIRE(C % G == 0, "gn_mish: C=%d not divisible by G=%d", C, G);
    DD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VN = Vec<T>::N;
        DD_REQUIRE(C % VN == 0 && (C / G) % VN == 0, "gn_mish: channels per group (%d) must be a multiple of %d", C / G, VN);
        const int vpi = HW * (C / VN);                       // vectors per image
        DD_REQUIRE(tb_stride % 4 == 0, "gn_mish: time-bias row stride must be a multiple of 4 floats");
        DD_REQUIRE(256 % (C / VN) == 0, "gn_mish: C=%d must divide 256 vectors", C);
        launch_pdl(gn_mish_kernel<T>, dim3((unsigned)((vpi + 256 * GN_UNR - 1) / (256 * GN_UNR)), (unsigned)B), dim3(256), 0, (cudaStream_t)stream,
            (const T*)x, (T*)y, HW, C, G, stats, stats_mode, eps, gamma, beta, tbias, tb_stride, trow, trow_stride,
            (const T*)residual, vpi);
    });
    return check_launch("gn_mish");
}

static int gn_mish_sum_parts(int C, int G) {
    const int cv = C / 4;
    for (int pt = 4; pt > 1; pt >>= 1)
        if (G % pt == 0 && cv % pt == 0 && 256 % (cv / pt) == 0) return pt;
    return 1;
}

int dd_gn_mish_sum_parts(int C, int G) { return (C > 0 && G > 0 && C % 4 == 0) ? gn_mish_sum_parts(C, G) : 0; }

int dd_gn_mish_sum(const float* part, int S, const float* bias, void* y_bf16, int B, int HW, int C, int G, float eps,
                   const float* gamma, const float* beta, const float* tbias, int tb_stride, const int32_t* trow,
                   int trow_stride, const void* residual_bf16, float* ln_part, void* stream) {
    DD_REQUIRE(S >= 1 && B > 0 && HW > 0 && G > 0 && G <= 64 && C % G == 0, "gn_mish_sum: bad sizes S=%d B=%d HW=%d C=%d G=%d", S, B, HW, C, G);
    const int cv = C / 4, cpg = C / G;
    DD_REQUIRE(C % 4 == 0 && 256 % cv == 0 && cpg % 4 == 0 && cpg / 4 <= 32 && ((cpg / 4) & (cpg / 4 - 1)) == 0,
               "gn_mish_sum: unsupported channel layout C=%d G=%d", C, G);
    DD_REQUIRE((int64_t)HW * C <= 16384, "gn_mish_sum: HW*C=%lld exceeds the shared-memory stage (16384)", (long long)HW * C);
    DD_REQUIRE(tb_stride % 4 == 0, "gn_mish_sum: time-bias row stride must be a multiple of 4 floats");
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(gn_mish_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4);
        attr_done = true;
    }
    const int parts = gn_mish_sum_parts(C, G);
    const int cvp = cv / parts;
    DD_REQUIRE(ln_part == nullptr || (cvp <= 32 && (cvp & (cvp - 1)) == 0), "gn_mish_sum: LayerNorm partials need a power-of-two <= 32 vectors per pixel and CTA");
    launch_pdl(gn_mish_sum_kernel, dim3(B, parts), dim3(256), (size_t)HW * C * 4 / parts, (cudaStream_t)stream, part, S, (int64_t)B * HW * C, bias,
               (__nv_bfloat16*)y_bf16, HW, C, G, eps, gamma, beta, tbias, tb_stride, trow, trow_stride, (const __nv_bfloat16*)residual_bf16, ln_part);
    return check_launch("gn_mish_sum");
}

int dd_layernorm_c(const void* x, void* y, int dtype, int64_t P, int C, const float* g, const float* b, float eps,
                   void* stream) {
    DD_REQUIRE(C % 32 == 0 && C <= 512, "layernorm_c: C=%d must be a multiple of 32 and <= 512", C);
    if (dtype == DD_BF16 && (C == 64 || C == 128 || C == 256)) {
        const int lpp = C / 8, ppw = 32 / lpp;
        const int64_t warps = (P + ppw - 1) / ppw;
        const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
        if (lpp == 8) launch_pdl(layernorm_c_bf16_kernel<8>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, P, g, b, eps);
        else if (lpp == 16) launch_pdl(layernorm_c_bf16_kernel<16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, P, g, b, eps);
        else launch_pdl(layernorm_c_bf16_kernel<32>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, P, g, b, eps);
        return check_launch("layernorm_c");
    }
    const int per = C / 32;
    const int grid = grid_cap(P * 32, 256);
#define LN_CASE(N)                                                                                                   \
    case N:                                                                                                          \
        DD_DISPATCH_DTYPE(dtype, T, (launch_pdl(layernorm_c_kernel<T, N>, dim3(grid), dim3(256), 0, (cudaStream_t)stream,                \
                                        (const T*)x, (T*)y, P, C, g, b, eps)));                                      \
        break;
    switch (per) {
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(6) LN_CASE(8) LN_CASE(12) LN_CASE(16)
        default:
            DD_REQUIRE(false, "layernorm_c: unsupported C=%d", C);
    }
#undef LN_CASE
    return check_launch("layernorm_c");
}

int64_t dd_linattn_ws_floats(int B, int n, int heads) {
    const int chunk = la_chunk(n);
    return (int64_t)B * heads * ((n + chunk - 1) / chunk) * LA_WS;
}

int dd_linattn_core(const void* qkv, void* out, int dtype, int B, int n, int heads, int dh, float* ws, int64_t ws_floats,
                    void* stream) {
    DD_REQUIRE(dh == 32 && heads > 0 && n > 0, "linattn_core: dim_head must be 32 (got %d)", dh);
    DD_REQUIRE(ws != nullptr && ws_floats >= dd_linattn_ws_floats(B, n, heads), "linattn_core: workspace too small");
    const int chunk = la_chunk(n);
    dim3 grid(B * heads, (n + chunk - 1) / chunk);
    DD_DISPATCH_DTYPE(dtype, T, {
        launch_pdl(linattn_ctx_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)qkv, ws, n, heads, chunk);
        launch_pdl(linattn_out_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)qkv, (T*)out, ws, n, heads, chunk);
    });
    return check_launch("linattn_core");
}

}  // extern "C"
