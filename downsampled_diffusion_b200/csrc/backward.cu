// Backward kernels of the training denoising step (fp32 programs): weight/bias gradients of every
// convolution, GroupNorm+Mish(+time bias,+dropout) backward, channel-LayerNorm backward, LinearAttention
// backward, and the small elementwise pieces (Mish / tanh derivatives, accumulate, scale).
// Input gradients of convolutions reuse conv_direct_kernel with transposed / flipped weights (engine side).
#include "common.cuh"

namespace dd {

// ---------------------------------------------------------------------------------------------
// wgrad: dW[tap][ci][co] += sum_pixels X[pixel shifted by tap][ci] * dY[pixel][co]
// GEMM with rows r = (tap, ci), columns co, reduction over output pixels; the pixel range is split over
// grid.z and partial tiles are accumulated with atomics (dW must hold the running gradient or zeros).
// ---------------------------------------------------------------------------------------------
struct WgradArgs {
    const void* x; const void* x2; const float* dy; float* dw;
    int C1, C2, B, H, W, Ho, Wo, Cout;
    int ksize, stride, pad, mode, flags;
    int K;              // taps * (C1 + C2)
    int64_t M;          // B * Ho * Wo
    int64_t m_per_split;
};

template <typename TI>
__device__ __forceinline__ float wgrad_fetch(const WgradArgs& a, int b, int oh, int ow, int k) {
    const int Cin = a.C1 + a.C2;
    const int tap = k / Cin, c = k - tap * Cin;
    int ih, iw;
    if (a.mode == 0) {
        const int ky = tap / a.ksize, kx = tap - ky * a.ksize;
        ih = oh * a.stride - a.pad + ky;
        iw = ow * a.stride - a.pad + kx;
    } else {
        const int ky = tap / a.ksize, kx = tap - ky * a.ksize;
        const int th = oh + a.pad - ky, tw = ow + a.pad - kx;
        if ((th | tw) < 0 || (th % a.stride) || (tw % a.stride)) return 0.f;
        ih = th / a.stride; iw = tw / a.stride;
    }
    if (ih < 0 || ih >= a.H || iw < 0 || iw >= a.W) return 0.f;
    float v;
    if (a.flags & DD_CONV_IN_NCHW) v = reinterpret_cast<const float*>(a.x)[(((int64_t)b * Cin + c) * a.H + ih) * a.W + iw];
    else if (c < a.C1) v = to_f(reinterpret_cast<const TI*>(a.x)[(((int64_t)b * a.H + ih) * a.W + iw) * a.C1 + c]);
    else v = to_f(reinterpret_cast<const TI*>(a.x2)[(((int64_t)b * a.H + ih) * a.W + iw) * a.C2 + (c - a.C1)]);
    if (a.flags & DD_CONV_PRE_MISH) v = mish_f(v);
    return v;
}

template <typename TI>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradArgs a) {
    pdl_sync();
    constexpr int TR = 64, TN = 64, TK = 16;
    __shared__ float As[TK][TR + 4];     // [pixel][row]
    __shared__ float Bs[TK][TN + 4];     // [pixel][co]
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * TR, n0 = blockIdx.y * TN;
    const int64_t m_lo = (int64_t)blockIdx.z * a.m_per_split;
    const int64_t m_hi = min(a.M, m_lo + a.m_per_split);
    const int tx = tid & 15, ty = tid >> 4;
    // A loads: pixel kk = tid / 16, rows (tid % 16) * 4 .. +4 ; B loads: pixel kk = tid / 16, co (tid % 16) * 4 .. +4
    const int lk = tid >> 4, l4 = (tid & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t m0 = m_lo; m0 < m_hi; m0 += TK) {
        const int64_t m = m0 + lk;
        const bool mv = m < m_hi;
        int b = 0, oh = 0, ow = 0;
        if (mv) { ow = (int)(m % a.Wo); const int64_t t = m / a.Wo; oh = (int)(t % a.Ho); b = (int)(t / a.Ho); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + l4 + j;
            As[lk][l4 + j] = (mv && r < a.K) ? wgrad_fetch<TI>(a, b, oh, ow, r) : 0.f;
            const int n = n0 + l4 + j;
            float g = 0.f;
            if (mv && n < a.Cout)
                g = (a.flags & DD_CONV_OUT_NCHW) ? a.dy[(((int64_t)b * a.Cout + n) * a.Ho + oh) * a.Wo + ow] : a.dy[m * a.Cout + n];
            Bs[lk][l4 + j] = g;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r >= a.K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.Cout) atomicAdd(a.dw + (int64_t)r * a.Cout + n, acc[i][j]);
        }
    }
}

// out[c] += sum_m x[m][c]   (bias gradients; batch reduction of per-sample channel sums)
__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t M, int C, int64_t m_per, int nchw,
                              int64_t HW) {
    pdl_sync();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int wy = threadIdx.x >> 5;                 // 8 row lanes
    const int64_t lo = (int64_t)blockIdx.y * m_per, hi = min(M, lo + m_per);
    float s = 0.f;
    if (c < C) {
        if (!nchw) for (int64_t m = lo + wy; m < hi; m += 8) s += x[m * C + c];
        else for (int64_t m = lo + wy; m < hi; m += 8) s += x[((m / HW) * C + c) * HW + (m % HW)];
    }
    __shared__ float red[8][33];
    red[wy][threadIdx.x & 31] = s;
    __syncthreads();
    if (wy == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

// ---------------------------------------------------------------------------------------------
// dropout mask: counter-based hash of (seed, element index); the same function regenerates the mask in
// the backward pass, nothing is stored.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dropout_scale(uint32_t seed, uint64_t idx, uint32_t thresh, float keep_inv) {
    uint64_t z = idx + ((uint64_t)seed << 32) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((uint32_t)z >= thresh) ? keep_inv : 0.f;      // P(drop) = thresh / 2^32
}

// The per-call seed lives in device memory (seed_dev, may be null) so that a launch recorded in a CUDA graph draws a fresh
// mask on every replay; `salt` separates the dropout layers of one network.
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, uint32_t salt,
                               const uint32_t* __restrict__ seed_dev, uint32_t thresh, float keep_inv) {
    pdl_sync();
    const uint32_t seed = (salt * 0x9E3779B9u) ^ (seed_dev ? *seed_dev : 0u);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = x[i] * dropout_scale(seed, (uint64_t)i, thresh, keep_inv);
}

// ---------------------------------------------------------------------------------------------
// GroupNorm + Mish backward.  Forward: xh = (x-mean)*rstd, h = xh*gamma+beta, y = mish(h) [+tb] [+res].
//   reduce: per (b, c): sums over pixels of dh*xh, dh, dy   (dh = dy * mish'(h))
//   apply : dx = rstd * (gamma*dh - S1/n - xh*S2/n), S1 = sum_{c in g} gamma_c*sum(dh), S2 = sum gamma_c*sum(dh*xh)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int HW, int C, int G,
                                                            float* __restrict__ s_dhxh, float* __restrict__ s_dh,
                                                            float* __restrict__ s_dy) {
    pdl_sync();
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G;
    const float mean = stats[2 * blockIdx.x], rstd = stats[2 * blockIdx.x + 1];
    const int cl = threadIdx.x % cpg, pl = threadIdx.x / cpg, pstep = blockDim.x / cpg;
    const int c = g * cpg + cl;
    const float ga = gamma[c], be = beta[c];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (pl < pstep)
        for (int p = pl; p < HW; p += pstep) {
            const int64_t off = ((int64_t)b * HW + p) * C + c;
            const float xh = (x[off] - mean) * rstd;
            const float d = dy[off];
            const float dh = d * mish_grad_f(xh * ga + be);
            a0 += dh * xh; a1 += dh; a2 += d;
        }
    __shared__ float red[3][256];
    red[0][threadIdx.x] = a0; red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.x < cpg) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        for (int k = threadIdx.x; k < pstep * cpg; k += cpg) { t0 += red[0][k]; t1 += red[1][k]; t2 += red[2][k]; }
        const int64_t o = (int64_t)b * C + g * cpg + threadIdx.x;
        s_dhxh[o] = t0; s_dh[o] = t1; s_dy[o] = t2;
    }
}

__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int HW, int C, int G,
                                                           const float* __restrict__ s_dhxh, const float* __restrict__ s_dh,
                                                           float* __restrict__ dx, int accumulate, int64_t total) {
    pdl_sync();
    const int cpg = C / G;
    const float inv_n = 1.f / ((float)HW * (float)cpg);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int b = (int)(i / ((int64_t)C * HW));
        const int g = c / cpg;
        const float mean = stats[((int64_t)b * G + g) * 2], rstd = stats[((int64_t)b * G + g) * 2 + 1];
        float S1 = 0.f, S2 = 0.f;
        for (int k = 0; k < cpg; ++k) {
            const int cc = g * cpg + k;
            S1 += gamma[cc] * s_dh[(int64_t)b * C + cc];
            S2 += gamma[cc] * s_dhxh[(int64_t)b * C + cc];
        }
        const float xh = (x[i] - mean) * rstd;
        const float dh = dy[i] * mish_grad_f(xh * gamma[c] + beta[c]);
        const float v = rstd * (gamma[c] * dh - S1 * inv_n - xh * S2 * inv_n);
        dx[i] = accumulate ? dx[i] + v : v;
    }
}

// The same, one image per blockIdx.y: the group sums S1 / S2, mean and rstd of the image's G groups are formed once per block in
// shared memory (the kernel above recomputes them with 2 * cpg loads for every element), four channels per thread.
// Needs C / G % 4 == 0 and G <= 64.
__global__ void __launch_bounds__(256) gn_bwd_apply4_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int HW, int C, int G,
                                                            const float* __restrict__ s_dhxh, const float* __restrict__ s_dh,
                                                            float4* __restrict__ dx, int accumulate) {
    pdl_sync();
    __shared__ float sS1[64], sS2[64], sMean[64], sRstd[64];
    const int b = blockIdx.y, cpg = C / G, C4 = C >> 2;
    const float inv_n = 1.f / ((float)HW * (float)cpg);
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        float S1 = 0.f, S2 = 0.f;
        for (int k = 0; k < cpg; ++k) {
            const int cc = g * cpg + k;
            S1 += gamma[cc] * s_dh[(int64_t)b * C + cc];
            S2 += gamma[cc] * s_dhxh[(int64_t)b * C + cc];
        }
        sS1[g] = S1; sS2[g] = S2;
        sMean[g] = stats[((int64_t)b * G + g) * 2]; sRstd[g] = stats[((int64_t)b * G + g) * 2 + 1];
    }
    __syncthreads();
    const int n4 = HW * C4;
    const int64_t base = (int64_t)b * n4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const int c = (i % C4) * 4, g = c / cpg;
        const float mean = sMean[g], rstd = sRstd[g], S1 = sS1[g], S2 = sS2[g];
        const float4 xv = x[base + i], dv = dy[base + i];
        const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
        float4 v;
#define GNB_ONE(F)                                                                \
        {                                                                         \
            const float xh = (xv.F - mean) * rstd;                                \
            const float dh = dv.F * mish_grad_f(xh * ga.F + be.F);                \
            v.F = rstd * (ga.F * dh - S1 * inv_n - xh * S2 * inv_n);              \
        }
        GNB_ONE(x) GNB_ONE(y) GNB_ONE(z) GNB_ONE(w)
#undef GNB_ONE
        if (accumulate) { const float4 o = dx[base + i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        dx[base + i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// channel LayerNorm backward (blocks.py:57-60; eps added to the std).  One warp per pixel.
//   y = (x-mu)/(sigma+eps)*g + b ; u = dy*g ; s = sigma+eps
//   dx_j = u_j/s - mean(u)/s - (x_j-mu) * [sum_i u_i (x_i-mu)] / (C * sigma * s^2)
// ---------------------------------------------------------------------------------------------
template <int PER_LANE>
__global__ void __launch_bounds__(256) layernorm_c_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ g, float eps, int64_t P, int C,
                                                              float* __restrict__ dx, int accumulate, float* __restrict__ dg,
                                                              float* __restrict__ db) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float pg[PER_LANE], pb[PER_LANE], gv[PER_LANE];
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j) { pg[j] = 0.f; pb[j] = 0.f; gv[j] = g[lane * PER_LANE + j]; }
    for (int64_t p = warp; p < P; p += nwarps) {
        float v[PER_LANE], u[PER_LANE];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) { v[j] = x[p * C + lane * PER_LANE + j]; s += v[j]; }
        const float mu = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) { v[j] -= mu; q += v[j] * v[j]; }
        const float sigma = sqrtf(warp_sum(q) / (float)C);
        const float sden = sigma + eps, inv = 1.f / sden;
        float su = 0.f, suv = 0.f;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const float d = dy[p * C + lane * PER_LANE + j];
            u[j] = d * gv[j];
            su += u[j]; suv += u[j] * v[j];
            pg[j] += d * v[j] * inv; pb[j] += d;
        }
        su = warp_sum(su) / (float)C;
        suv = warp_sum(suv);
        const float k2 = sigma > 0.f ? suv / ((float)C * sigma * sden * sden) : 0.f;
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const float r = (u[j] - su) * inv - v[j] * k2;
            const int64_t o = p * C + lane * PER_LANE + j;
            dx[o] = accumulate ? dx[o] + r : r;
        }
    }
    // gain / bias gradients: the eight warps of a block are summed in shared memory first, so that a block issues one global
    // atomic per channel (eight times fewer contended atomics on 2*C addresses: 225 -> ~30 us at 32x32x128, batch 32)
    __shared__ float s_g[8][PER_LANE * 32], s_b[8][PER_LANE * 32];
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j) { s_g[w][lane * PER_LANE + j] = pg[j]; s_b[w][lane * PER_LANE + j] = pb[j]; }
    __syncthreads();
    for (int c = threadIdx.x; c < PER_LANE * 32; c += blockDim.x) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += s_g[k][c]; b += s_b[k][c]; }
        atomicAdd(dg + c, a);
        atomicAdd(db + c, b);
    }
}

// ---------------------------------------------------------------------------------------------
// LinearAttention backward (fp32).  saved: per (b, head) {M_d, S_d, ctx[d][e]} = the merged forward state.
//   dctx[d][e] = sum_n q[n][d] dout[n][e]                      (kernel 1, atomics over n-splits)
//   dq[n][d]   = sum_e ctx[d][e] dout[n][e]
//   kt[n][d]   = exp(k[n][d]-M_d)/S_d ; dkt = sum_e dctx[d][e] v[n][e] ; dk = kt * (dkt - sum_e dctx[d][e] ctx[d][e])
//   dv[n][e]   = sum_d kt[n][d] dctx[d][e]                      (kernel 2)
// ---------------------------------------------------------------------------------------------
constexpr int LAB_WS = 64 + 32 * 32;

__global__ void __launch_bounds__(256) linattn_bwd_dctx_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                               float* __restrict__ dctx, int n, int heads, int chunk) {
    pdl_sync();
    constexpr int DH = 32;
    const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
    const int C3 = 3 * heads * DH, CO = heads * DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* qb = qkv + (int64_t)b * n * C3 + hd * DH;
    const float* db_ = dout + (int64_t)b * n * CO + hd * DH;
    const int n_lo = blockIdx.y * chunk, n_hi = min(n, n_lo + chunk);
    __shared__ float sq[64][DH + 1], sd[64][DH + 1];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};       // thread (d = lane, e = warp*4..+4)
    for (int t0 = n_lo; t0 < n_hi; t0 += 64) {
        const int rows = min(64, n_hi - t0);
        for (int r = warp; r < rows; r += 8) {
            sq[r][lane] = qb[(int64_t)(t0 + r) * C3 + lane];
            sd[r][lane] = db_[(int64_t)(t0 + r) * CO + lane];
        }
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            const float qv = sq[r][lane];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += qv * sd[r][warp * 4 + j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dctx + (int64_t)bh * 1024 + lane * 32 + warp * 4 + j, acc[j]);
}

__global__ void __launch_bounds__(256) linattn_bwd_apply_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                                const float* __restrict__ saved, const float* __restrict__ dctx,
                                                                float* __restrict__ dqkv, int n, int heads, int chunk) {
    pdl_sync();
    constexpr int DH = 32;
    const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
    const int C3 = 3 * heads * DH, CO = heads * DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float s_ctx[DH][DH + 1], s_dctx[DH][DH + 1];
    const float* sv = saved + (int64_t)bh * LAB_WS;
    for (int i = threadIdx.x; i < 1024; i += 256) {
        s_ctx[i / 32][i % 32] = sv[64 + i];
        s_dctx[i / 32][i % 32] = dctx[(int64_t)bh * 1024 + i];
    }
    __syncthreads();
    const float Md = sv[lane], Sinv = 1.f / sv[32 + lane];
    float rd = 0.f;                                     // r[d], d = lane
#pragma unroll
    for (int e = 0; e < DH; ++e) rd += s_dctx[lane][e] * s_ctx[lane][e];
    float ctx_row[DH], dctx_row[DH], dctx_col[DH];      // [lane][e], [lane][e], [d][lane]
#pragma unroll
    for (int e = 0; e < DH; ++e) { ctx_row[e] = s_ctx[lane][e]; dctx_row[e] = s_dctx[lane][e]; dctx_col[e] = s_dctx[e][lane]; }
    const float* base = qkv + (int64_t)b * n * C3 + hd * DH;
    float* dbase = dqkv + (int64_t)b * n * C3 + hd * DH;
    const float* dob = dout + (int64_t)b * n * CO + hd * DH;
    const int n_lo = blockIdx.y * chunk, n_hi = min(n, n_lo + chunk);
    for (int i = n_lo + warp; i < n_hi; i += 8) {
        const float kv = base[(int64_t)i * C3 + heads * DH + lane];
        const float vv = base[(int64_t)i * C3 + 2 * heads * DH + lane];
        const float dov = dob[(int64_t)i * CO + lane];
        const float kt = __expf(kv - Md) * Sinv;
        float dq = 0.f, dkt = 0.f, dv = 0.f;
#pragma unroll
        for (int e = 0; e < DH; ++e) {
            dq += ctx_row[e] * __shfl_sync(0xffffffffu, dov, e);
            dkt += dctx_row[e] * __shfl_sync(0xffffffffu, vv, e);
            dv += dctx_col[e] * __shfl_sync(0xffffffffu, kt, e);       // lane = e, loop index = d
        }
        dbase[(int64_t)i * C3 + lane] = dq;
        dbase[(int64_t)i * C3 + heads * DH + lane] = kt * (dkt - rd);
        dbase[(int64_t)i * C3 + 2 * heads * DH + lane] = dv;
    }
}

// merged forward state of the attention core for the backward pass: {M_d, S_d, ctx[d][e]} per (b, head)
__global__ void __launch_bounds__(256) linattn_save_kernel(const float* __restrict__ ws, int S, float* __restrict__ saved) {
    pdl_sync();
    const int bh = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* w0 = ws + (int64_t)bh * S * LAB_WS;
    float M = -INFINITY;
    for (int s = 0; s < S; ++s) M = fmaxf(M, w0[s * LAB_WS + lane]);
    float tot = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < S; ++s) {
        const float* w = w0 + s * LAB_WS;
        const float f = __expf(w[lane] - M);
        tot += w[32 + lane] * f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += w[64 + lane * 32 + warp * 4 + j] * f;
    }
    float* o = saved + (int64_t)bh * LAB_WS;
    if (warp == 0) { o[lane] = M; o[32 + lane] = tot; }
    const float inv = 1.f / tot;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[64 + lane * 32 + warp * 4 + j] = acc[j] * inv;
}

// ---------------------------------------------------------------------------------------------
// small elementwise pieces
// ---------------------------------------------------------------------------------------------
// mode 0: y = mish(x); 1: y = g * mish'(x) (+= if accumulate); 2: y = g * (1 - t^2), t = tanh output;
// 3: y (+)= alpha * x ; 4: y = tanh(x)
__device__ __forceinline__ float ew_one(int mode, float x, float g, float alpha) {
    if (mode == 0) return mish_f(x);
    if (mode == 1) return g * mish_grad_f(x);
    if (mode == 2) return g * (1.f - x * x);
    if (mode == 4) return tanhf(x);
    if (mode == 5) { const float d = __fsub_rn(x, g); return __fmul_rn(d, d); }      // F.mse_loss(x, g, reduction='none')
    return alpha * x;
}

// four elements per thread (all operands 16-byte aligned, n % 4 == 0: every activation / gradient buffer of a program)
__global__ void __launch_bounds__(256) ew4_kernel(int mode, const float4* __restrict__ x, const float4* __restrict__ g,
                                                  float4* __restrict__ y, int64_t n4, float alpha, int accumulate) {
    pdl_sync();
    const bool needs_g = (mode == 1 || mode == 2 || mode == 5);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = x[i];
        const float4 b = needs_g ? g[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = make_float4(ew_one(mode, a.x, b.x, alpha), ew_one(mode, a.y, b.y, alpha), ew_one(mode, a.z, b.z, alpha),
                               ew_one(mode, a.w, b.w, alpha));
        if (accumulate) {
            const float4 o = y[i];
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        y[i] = v;
    }
}

__global__ void ew_kernel(int mode, const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ y, int64_t n,
                          float alpha, int accumulate) {
    pdl_sync();
    const bool needs_g = (mode == 1 || mode == 2 || mode == 5);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = ew_one(mode, x[i], needs_g ? g[i] : 0.f, alpha);
        y[i] = accumulate ? y[i] + v : v;
    }
}

__global__ void sincos_emb_kernel(const float* __restrict__ t, const float* __restrict__ freq, float* __restrict__ out, int R,
                                  int dim) {
    pdl_sync();
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * half) return;
    const int r = i / half, k = i % half;
    const float a = __fmul_rn(t[r], freq[k]);
    out[(int64_t)r * dim + k] = sinf(a);
    out[(int64_t)r * dim + half + k] = cosf(a);
}

// nearest x2 backward / avg-pool forward share one kernel shape: y[b,ho,wo,c] = scale * sum of the 2x2 block of x.
// 16-byte vectors along c, 64-bit divisions only once per vector (the scalar form ran at 1.4 TB/s on the 1 GB maps).
__global__ void pool2_sum_kernel(const float4* __restrict__ x, float4* __restrict__ y, int H, int W, int cv, float scale,
                                 int64_t total) {
    pdl_sync();
    const int Ho = H >> 1, Wo = W >> 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        int64_t r = i / cv;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int64_t b = r / Ho;
        const float4* p = x + (((b * H + 2 * ho) * W + 2 * wo) * (int64_t)cv) + c;
        const float4 a = p[0], bq = p[cv], cq = p[(int64_t)W * cv], d = p[(int64_t)W * cv + cv];
        y[i] = make_float4(scale * (a.x + bq.x + cq.x + d.x), scale * (a.y + bq.y + cq.y + d.y), scale * (a.z + bq.z + cq.z + d.z),
                           scale * (a.w + bq.w + cq.w + d.w));
    }
}

// avg-pool backward / nearest x2 forward: y[b,h,w,c] = scale * x[b,h/2,w/2,c]; one thread per INPUT vector writes its 2x2 block
__global__ void unpool2_kernel(const float4* __restrict__ x, float4* __restrict__ y, int H, int W, int cv, float scale, int64_t total) {
    pdl_sync();
    const int Wo = W * 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        int64_t r = i / cv;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int64_t b = r / H;
        float4 v = x[i];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        float4* q = y + (((b * (2 * H) + 2 * h) * Wo + 2 * w) * (int64_t)cv) + c;
        q[0] = v; q[cv] = v; q[(int64_t)Wo * cv] = v; q[(int64_t)Wo * cv + cv] = v;
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

int dd_conv_wgrad(const void* x, const void* x2, int C1, int C2, int in_dtype, const float* dy, float* dw, int B, int H, int W,
                  int Cout, int ksize, int stride, int pad, int mode, int flags, void* stream) {
    DD_REQUIRE(B > 0 && H > 0 && W > 0 && Cout > 0 && C1 > 0 && C2 >= 0, "conv_wgrad: bad sizes");
    DD_REQUIRE((C2 == 0) == (x2 == nullptr), "conv_wgrad: x2/C2 mismatch");
    WgradArgs a;
    a.x = x; a.x2 = x2; a.dy = dy; a.dw = dw;
    a.C1 = C1; a.C2 = C2; a.B = B; a.H = H; a.W = W; a.Cout = Cout;
    a.ksize = ksize; a.stride = stride; a.pad = pad; a.mode = mode; a.flags = flags;
    if (mode == 0) { a.Ho = (H + 2 * pad - ksize) / stride + 1; a.Wo = (W + 2 * pad - ksize) / stride + 1; }
    else { a.Ho = H * stride; a.Wo = W * stride; }
    a.K = ksize * ksize * (C1 + C2);
    a.M = (int64_t)B * a.Ho * a.Wo;
    const int gx = (a.K + 63) / 64, gy = (Cout + 63) / 64;
    int64_t want = ((int64_t)num_sms() * 4 + gx * gy - 1) / (gx * gy);      // ~4 CTAs per SM in total
    int64_t max_split = (a.M + 255) / 256;
    int64_t splits = want < 1 ? 1 : (want > max_split ? max_split : want);
    a.m_per_split = ((a.M + splits - 1) / splits + 15) / 16 * 16;
    splits = (a.M + a.m_per_split - 1) / a.m_per_split;
    dim3 grid(gx, gy, (unsigned)splits);
    if (in_dtype == DD_F32) launch_pdl(conv_wgrad_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, a);
    else if (in_dtype == DD_BF16) launch_pdl(conv_wgrad_kernel<__nv_bfloat16>, grid, dim3(256), 0, (cudaStream_t)stream, a);
    else { set_error("conv_wgrad: bad dtype"); return DD_ERR_ARG; }
    return check_launch("conv_wgrad");
}

int dd_colsum(const float* x, float* out, int64_t M, int C, int nchw, int64_t HW, void* stream) {
    DD_REQUIRE(M > 0 && C > 0, "colsum: bad sizes");
    const int gx = (C + 31) / 32;
    int64_t splits = ((int64_t)num_sms() * 4 + gx - 1) / gx;
    const int64_t max_split = (M + 63) / 64;
    if (splits > max_split) splits = max_split;
    if (splits < 1) splits = 1;
    const int64_t m_per = (M + splits - 1) / splits;
    splits = (M + m_per - 1) / m_per;
    launch_pdl(colsum_kernel, dim3(gx, (unsigned)splits), dim3(256), 0, (cudaStream_t)stream, x, out, M, C, m_per, nchw, HW);
    return check_launch("colsum");
}

int dd_dropout(const float* x, float* y, int64_t n, uint32_t salt, const uint32_t* seed_dev, float p, void* stream) {
    DD_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0,1)");
    const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
    launch_pdl(dropout_kernel, dim3(grid_cap(n, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, n, salt, seed_dev, thresh,
               1.f / (1.f - p));
    return check_launch("dropout");
}

int dd_gn_mish_bwd(const float* x, const float* dy, const float* stats, const float* gamma, const float* beta, int B, int HW,
                   int C, int G, float* s_dhxh, float* s_dh, float* s_dy, float* dx, int accumulate, void* stream) {
    DD_REQUIRE(C % G == 0 && 256 % (C / G) == 0, "gn_mish_bwd: channels per group must divide 256");
    launch_pdl(gn_bwd_reduce_kernel, dim3(B * G), dim3(256), 0, (cudaStream_t)stream, x, dy, stats, gamma, beta, HW, C, G, s_dhxh,
               s_dh, s_dy);
    const int64_t total = (int64_t)B * HW * C;
    const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                            reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta);
    if ((C / G) % 4 == 0 && G <= 64 && B <= 65535 && (align & 15) == 0 && (int64_t)HW * C < (1LL << 31)) {
        int gx = (HW * (C / 4) + 255) / 256;
        const int cap = (8 * num_sms() + B - 1) / B;
        if (gx > cap) gx = cap;
        launch_pdl(gn_bwd_apply4_kernel, dim3(gx < 1 ? 1 : gx, B), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (const float4*)dy,
                   stats, gamma, beta, HW, C, G, (const float*)s_dhxh, (const float*)s_dh, (float4*)dx, accumulate);
    } else {
        launch_pdl(gn_bwd_apply_kernel, dim3(grid_cap(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, dy, stats, gamma, beta, HW, C,
                   G, (const float*)s_dhxh, (const float*)s_dh, dx, accumulate, total);
    }
    return check_launch("gn_mish_bwd");
}

int dd_layernorm_c_bwd(const float* x, const float* dy, const float* g, float eps, int64_t P, int C, float* dx, int accumulate,
                       float* dg, float* db, void* stream) {
    DD_REQUIRE(C % 32 == 0 && C <= 512, "layernorm_c_bwd: C=%d must be a multiple of 32 and <= 512", C);
    const int grid = grid_cap(P * 32, 256) > 4 * num_sms() ? 4 * num_sms() : grid_cap(P * 32, 256);
#define LNB_CASE(N)                                                                                                           \
    case N:                                                                                                                   \
        launch_pdl(layernorm_c_bwd_kernel<N>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, dy, g, eps, P, C, dx, accumulate, \
                   dg, db);                                                                                                   \
        break;
    switch (C / 32) {
        LNB_CASE(1) LNB_CASE(2) LNB_CASE(3) LNB_CASE(4) LNB_CASE(6) LNB_CASE(8) LNB_CASE(12) LNB_CASE(16)
        default: DD_REQUIRE(false, "layernorm_c_bwd: unsupported C=%d", C);
    }
#undef LNB_CASE
    return check_launch("layernorm_c_bwd");
}

int dd_linattn_save(const float* ws, int B, int n, int heads, float* saved, void* stream) {
    const int chunk = (n + 15) / 16 <= 128 ? 128 : ((n + 15) / 16 + 127) / 128 * 128;
    const int S = (n + chunk - 1) / chunk;
    launch_pdl(linattn_save_kernel, dim3(B * heads), dim3(256), 0, (cudaStream_t)stream, ws, S, saved);
    return check_launch("linattn_save");
}

int dd_linattn_bwd(const float* qkv, const float* dout, const float* saved, float* dctx, float* dqkv, int B, int n, int heads,
                   int dh, void* stream) {
    DD_REQUIRE(dh == 32, "linattn_bwd: dim_head must be 32");
    cudaError_t e = cudaMemsetAsync(dctx, 0, (size_t)B * heads * 1024 * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("linattn_bwd: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
    const int chunk = n <= 256 ? 256 : 512;
    dim3 grid(B * heads, (n + chunk - 1) / chunk);
    launch_pdl(linattn_bwd_dctx_kernel, grid, dim3(256), 0, (cudaStream_t)stream, qkv, dout, dctx, n, heads, chunk);
    launch_pdl(linattn_bwd_apply_kernel, grid, dim3(256), 0, (cudaStream_t)stream, qkv, dout, saved, (const float*)dctx, dqkv, n, heads,
               chunk);
    return check_launch("linattn_bwd");
}

int dd_ew(int mode, const float* x, const float* g, float* y, int64_t n, float alpha, int accumulate, void* stream) {
    DD_REQUIRE(mode >= 0 && mode <= 5 && n > 0, "ew: bad mode");
    const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(y);
    if ((align & 15) == 0 && n % 4 == 0) {
        launch_pdl(ew4_kernel, dim3(grid_cap(n / 4, 256)), dim3(256), 0, (cudaStream_t)stream, mode, (const float4*)x, (const float4*)g,
                   (float4*)y, n / 4, alpha, accumulate);
        return check_launch("ew");
    }
    launch_pdl(ew_kernel, dim3(grid_cap(n, 256)), dim3(256), 0, (cudaStream_t)stream, mode, x, g, y, n, alpha, accumulate);
    return check_launch("ew");
}

int dd_sincos_emb(const float* t, const float* freq, float* out, int R, int dim, void* stream) {
    launch_pdl(sincos_emb_kernel, dim3((R * (dim / 2) + 127) / 128), dim3(128), 0, (cudaStream_t)stream, t, freq, out, R, dim);
    return check_launch("sincos_emb");
}

int dd_pool2_sum(const float* x, float* y, int B, int H, int W, int C, float scale, void* stream) {
    DD_REQUIRE(H % 2 == 0 && W % 2 == 0, "pool2_sum: odd size");
    DD_REQUIRE(C % 4 == 0, "pool2_sum: C=%d must be a multiple of 4", C);
    const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 4);
    launch_pdl(pool2_sum_kernel, dim3(grid_cap(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (float4*)y, H, W, C / 4, scale, total);
    return check_launch("pool2_sum");
}

int dd_unpool2(const float* x, float* y, int B, int H, int W, int C, float scale, void* stream) {
    DD_REQUIRE(C % 4 == 0, "unpool2: C=%d must be a multiple of 4", C);
    const int64_t total = (int64_t)B * H * W * (C / 4);
    launch_pdl(unpool2_kernel, dim3(grid_cap(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (float4*)y, H, W, C / 4, scale, total);
    return check_launch("unpool2");
}

}  // extern "C"
