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
// y = mish(gn(x)*gamma+beta) [+ tbias[row(b), c]] [+ residual].  16-byte vectors along C.
// Algorithmic bytes: read x + write y (+ residual read) = 4 (+2) B/element in bf16.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void gn_mish_kernel(const T* __restrict__ x, T* __restrict__ y, int HW, int C, int G,
                               const float* __restrict__ stats, int stats_mode, float eps,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ tbias, int tb_stride, const int32_t* __restrict__ trow,
                               int trow_stride, const T* __restrict__ residual, int64_t total_vec) {
    constexpr int VN = Vec<T>::N;
    const int cv = C / VN, cpg = C / G;
    const float inv_n = 1.f / ((float)HW * (float)cpg);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VN;
        const int b = (int)(i / ((int64_t)cv * HW));
        const int g = c / cpg;
        float mean, rstd;
        const float s0 = stats[((int64_t)b * G + g) * 2], s1 = stats[((int64_t)b * G + g) * 2 + 1];
        if (stats_mode == 0) { mean = s0; rstd = s1; }
        else {
            mean = s0 * inv_n;
            float var = fmaxf(s1 * inv_n - mean * mean, 0.f);
            rstd = rsqrtf(var + eps);
        }
        Vec<T> v, r;
        v.load(x + i * VN);
        const float* tb = nullptr;
        if (tbias) {
            const int row = trow ? trow[(int64_t)b * trow_stride] : b;
            tb = tbias + (int64_t)row * tb_stride + c;
        }
        if (residual) r.load(residual + i * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
            float h = (v.v[j] - mean) * rstd * gamma[c + j] + beta[c + j];
            h = mish_f(h);
            if (tb) h += tb[j];
            if (residual) h += r.v[j];
            v.v[j] = h;
        }
        v.store(y + i * VN);
    }
}

// ---------------------------------------------------------------------------------------------
// Channel LayerNorm (blocks.py:57-60): one warp per pixel, eps added to the std.
// ---------------------------------------------------------------------------------------------
template <typename T, int PER_LANE>
__global__ void layernorm_c_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t P, int C,
                                   const float* __restrict__ g, const float* __restrict__ bta, float eps) {
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

// ---------------------------------------------------------------------------------------------
// time embedding -> per-block channel biases.  grid (ceil(J/128), R); the two small MLP layers are
// recomputed per CTA (131k MAC), then each CTA produces 128 of the J outputs.
// ---------------------------------------------------------------------------------------------
__global__ void time_bias_kernel(const float* __restrict__ t, int dim, const float* __restrict__ freq,
                                 const float* __restrict__ W1,
                                 const float* __restrict__ b1, const float* __restrict__ W2,
                                 const float* __restrict__ b2, const float* __restrict__ Wcat,
                                 const float* __restrict__ bcat, int J, float* __restrict__ out) {
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
// LinearAttention core (blocks.py:128-133).  One CTA (256 threads) per (b, head), dh = 32.
//   pass 1: column max / sum-exp of k over n (online, per d)
//   pass 2: ctx[d][e] = sum_n softmax(k)[d,n] v[e,n]      (tiles of 32 n staged in smem)
//   pass 3: out[n][e] = sum_d ctx[d][e] q[n][d]           (ctx column in registers, q via shuffles)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) linattn_core_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                           int n, int heads) {
    constexpr int DH = 32;
    const int b = blockIdx.x / heads, hd = blockIdx.x % heads;
    const int C3 = 3 * heads * DH, CO = heads * DH;
    const T* qb = qkv + (int64_t)b * n * C3 + hd * DH;
    const T* kb = qb + heads * DH;
    const T* vb = kb + heads * DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // 8 warps

    __shared__ float s_m[8][DH], s_s[8][DH];
    __shared__ float s_max[DH], s_inv[DH];
    __shared__ float s_k[32][DH + 1], s_v[32][DH + 1];
    __shared__ float s_ctx[DH][DH + 1];

    // pass 1: lane = d, each warp strides over n
    float m = -INFINITY, s = 0.f;
    for (int i = warp; i < n; i += 8) {
        const float kv = to_f(kb[(int64_t)i * C3 + lane]);
        const float nm = fmaxf(m, kv);
        s = s * expf(m - nm) + expf(kv - nm);
        m = nm;
    }
    s_m[warp][lane] = m; s_s[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        float mm = -INFINITY;
        for (int w = 0; w < 8; ++w) mm = fmaxf(mm, s_m[w][lane]);
        float ss = 0.f;
        for (int w = 0; w < 8; ++w) if (s_m[w][lane] > -INFINITY) ss += s_s[w][lane] * expf(s_m[w][lane] - mm);
        s_max[lane] = mm; s_inv[lane] = 1.f / ss;
    }
    __syncthreads();

    // pass 2: thread (d = lane, e block = warp*4 .. +4)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n0 = 0; n0 < n; n0 += 32) {
        for (int r = warp; r < 32; r += 8) {
            const int i = n0 + r;
            float kv = 0.f, vv = 0.f;
            if (i < n) {
                kv = expf(to_f(kb[(int64_t)i * C3 + lane]) - s_max[lane]);
                vv = to_f(vb[(int64_t)i * C3 + lane]);
            }
            s_k[r][lane] = kv; s_v[r][lane] = vv;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float p = s_k[r][lane];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += p * s_v[r][warp * 4 + j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s_ctx[lane][warp * 4 + j] = acc[j] * s_inv[lane];
    __syncthreads();

    // pass 3: lane = e
    float ctx[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) ctx[d] = s_ctx[d][lane];
    T* ob = out + (int64_t)b * n * CO + hd * DH;
    for (int i = warp; i < n; i += 8) {
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
    time_bias_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(t, dim, freq, W1, b1, W2, b2, Wcat, bcat, J, out);
    return check_launch("time_bias");
}

int dd_gn_stats(const void* x, int dtype, int B, int HW, int C, int G, float eps, float* stats, void* stream) {
    DD_REQUIRE(C % G == 0 && B > 0 && HW > 0, "gn_stats: C=%d not divisible by G=%d", C, G);
    DD_DISPATCH_DTYPE(dtype, T, (gn_stats_kernel<T><<<B * G, 256, 0, (cudaStream_t)stream>>>((const T*)x, HW, C, G, eps, stats)));
    return check_launch("gn_stats");
}

int dd_gn_mish(const void* x, void* y, int dtype, int B, int HW, int C, int G, const float* stats, int stats_mode,
               float eps, const float* gamma, const float* beta, const float* tbias, int tb_stride,
               const int32_t* trow, int trow_stride, const void* residual, void* stream) {
    DD_REQUIRE(C % G == 0, "gn_mish: C=%d not divisible by G=%d", C, G);
    DD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VN = Vec<T>::N;
        DD_REQUIRE(C % VN == 0 && (C / G) % VN == 0, "gn_mish: channels per group (%d) must be a multiple of %d", C / G, VN);
        int64_t n = (int64_t)B * HW * (C / VN);
        gn_mish_kernel<T><<<grid_cap(n, 256), 256, 0, (cudaStream_t)stream>>>(
            (const T*)x, (T*)y, HW, C, G, stats, stats_mode, eps, gamma, beta, tbias, tb_stride, trow, trow_stride,
            (const T*)residual, n);
    });
    return check_launch("gn_mish");
}

int dd_layernorm_c(const void* x, void* y, int dtype, int64_t P, int C, const float* g, const float* b, float eps,
                   void* stream) {
    DD_REQUIRE(C % 32 == 0 && C <= 512, "layernorm_c: C=%d must be a multiple of 32 and <= 512", C);
    const int per = C / 32;
    const int grid = grid_cap(P * 32, 256);
#define LN_CASE(N)                                                                                                   \
    case N:                                                                                                          \
        DD_DISPATCH_DTYPE(dtype, T, (layernorm_c_kernel<T, N><<<grid, 256, 0, (cudaStream_t)stream>>>(               \
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

int dd_linattn_core(const void* qkv, void* out, int dtype, int B, int n, int heads, int dh, void* stream) {
    DD_REQUIRE(dh == 32 && heads > 0 && n > 0, "linattn_core: dim_head must be 32 (got %d)", dh);
    DD_DISPATCH_DTYPE(dtype, T, (linattn_core_kernel<T><<<B * heads, 256, 0, (cudaStream_t)stream>>>(
                                    (const T*)qkv, (T*)out, n, heads)));
    return check_launch("linattn_core");
}

}  // extern "C"
