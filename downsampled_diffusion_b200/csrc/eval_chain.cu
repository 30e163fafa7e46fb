// Evaluation-side chain: the per-timestep term of the variational bound and the L_simple term of
// DDPM.test_losses_ (models/diffusion/ddpm.py:391-442), the prior term (ddpm.py:367-389), and the sampler caller's output
// formatting (utils/eval_helpers.py:37-41 fix_samples).  All fp32, HBM-trivial: one block per sample, the row means are
// reduced in-kernel so a step of the chain is q_sample -> U-Net -> vlb_terms -> tick with no host round trip.
#include "common.cuh"

namespace dd {

// row of the evaluation table (T, 8): the per-step scalars the reference gathers with extract() (helpers.py:31-34)
struct EvalCoef { float sa, sb, sr, srm1, c1, c2, lv, pad; };

__device__ __forceinline__ int ring_slot(int t, int T, int period) {
    int step = T - 1 - t;                       // draw order of ddpm.py:409-411: t = T-1 first
    return period > 0 ? step % period : step;
}

// x_t = sqrt_ac[t] x + sqrt_1mac[t] eps for ONE step shared by the batch (device-side step counter), eps from the noise ring
__global__ void q_sample_step_kernel(const float4* __restrict__ x, const float4* __restrict__ noise, int64_t step_stride4,
                                     int period, const EvalCoef* __restrict__ tab, const int32_t* __restrict__ t_idx, int T,
                                     float4* __restrict__ out, int64_t n4) {
    pdl_sync();
    const int t = t_idx[0];
    const float a = tab[t].sa, b = tab[t].sb;
    const float4* nz = noise + (int64_t)ring_slot(t, T, period) * step_stride4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = x[i], e = nz[i], o;
        o.x = __fadd_rn(__fmul_rn(a, v.x), __fmul_rn(b, e.x));
        o.y = __fadd_rn(__fmul_rn(a, v.y), __fmul_rn(b, e.y));
        o.z = __fadd_rn(__fmul_rn(a, v.z), __fmul_rn(b, e.z));
        o.w = __fadd_rn(__fmul_rn(a, v.w), __fmul_rn(b, e.w));
        out[i] = o;
    }
}

// approx_standard_normal_cdf, models/utils/losses.py:55-63
__device__ __forceinline__ float approx_cdf(float v) {
    const float cube = __fmul_rn(__fmul_rn(v, v), v);
    const float arg = __fmul_rn(0.7978845608028654f, __fadd_rn(v, __fmul_rn(0.044715f, cube)));
    return __fmul_rn(0.5f, __fadd_rn(1.0f, tanhf(arg)));
}

// one element of vlb_terms (ddpm.py:339-364): KL(q(x_{t-1}|x_t,x) || p(x_{t-1}|x_t)) for t > 0 (equal variances: the
// log-variance terms of normal_kl cancel exactly in fp32), minus the discretised Gaussian log-likelihood for t == 0.
__device__ __forceinline__ float vlb_elem(float x, float xt, float eh, const EvalCoef& c, bool first, float inv_var, float inv_std) {
    float x0 = __fsub_rn(__fmul_rn(c.sr, xt), __fmul_rn(c.srm1, eh));              // predict_x_from_eps, clip=True (ddpm.py:149-158)
    x0 = fminf(fmaxf(x0, -1.f), 1.f);
    const float c2xt = __fmul_rn(c.c2, xt);
    const float pm = __fadd_rn(__fmul_rn(c.c1, x0), c2xt);                          // p_mean_variance (ddpm.py:187-201)
    if (!first) {
        const float tm = __fadd_rn(__fmul_rn(c.c1, x), c2xt);                       // q_posterior (ddpm.py:160-185)
        const float d = __fsub_rn(tm, pm);
        return __fmul_rn(0.5f, __fmul_rn(__fmul_rn(d, d), inv_var));                // normal_kl, losses.py:17-52
    }
    const float cx = __fsub_rn(x, pm);                                              // losses.py:66-109
    const float cdf_plus = approx_cdf(__fmul_rn(inv_std, __fadd_rn(cx, 1.f / 255.f)));
    const float cdf_min = approx_cdf(__fmul_rn(inv_std, __fsub_rn(cx, 1.f / 255.f)));
    float lp;
    if (x < -0.999f) lp = logf(fmaxf(cdf_plus, 1e-12f));
    else if (x > 0.999f) lp = logf(fmaxf(__fsub_rn(1.f, cdf_min), 1e-12f));
    else lp = logf(fmaxf(__fsub_rn(cdf_plus, cdf_min), 1e-12f));
    return -lp;
}

// grid = B blocks.  vlb[b*out_stride + col] = mean_chw(term) / ln 2 (flat_bits, utils/utils.py:43-48);
// sq[b*out_stride + col] = sum_chw (eps - eps_hat)^2 when eps != null.  col = T-1-t when col_from_t (chain order), else 0.
__global__ void __launch_bounds__(256) vlb_terms_kernel(const float4* __restrict__ x, const float4* __restrict__ xt,
                                                        const float4* __restrict__ eh, const float4* __restrict__ eps,
                                                        int64_t eps_step_stride4, int eps_period,
                                                        const EvalCoef* __restrict__ tab, const int32_t* __restrict__ t_idx,
                                                        int t_stride, int T, float* __restrict__ vlb, float* __restrict__ sq,
                                                        int64_t out_stride, int col_from_t, int64_t chw4) {
    pdl_sync();
    const int b = blockIdx.x;
    const int t = t_idx[(int64_t)b * t_stride];
    const EvalCoef c = tab[t];
    const bool first = (t == 0);
    const float inv_var = expf(-c.lv);                      // exp(-logvar2)
    const float inv_std = expf(-__fmul_rn(0.5f, c.lv));     // exp(-log_scales), log_scales = 0.5 * logvar
    const int64_t base = (int64_t)b * chw4;
    const float4* nz = eps ? eps + (eps_step_stride4 ? (int64_t)ring_slot(t, T, eps_period) * eps_step_stride4 : 0) : nullptr;
    float acc = 0.f, acc2 = 0.f;
    for (int64_t i = threadIdx.x; i < chw4; i += blockDim.x) {
        const float4 a = x[base + i], y = xt[base + i], e = eh[base + i];
        acc += vlb_elem(a.x, y.x, e.x, c, first, inv_var, inv_std) + vlb_elem(a.y, y.y, e.y, c, first, inv_var, inv_std) +
               vlb_elem(a.z, y.z, e.z, c, first, inv_var, inv_std) + vlb_elem(a.w, y.w, e.w, c, first, inv_var, inv_std);
        if (nz) {
            const float4 n = nz[base + i];
            const float d0 = n.x - e.x, d1 = n.y - e.y, d2 = n.z - e.z, d3 = n.w - e.w;
            acc2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
    }
    __shared__ double red[2][8];
    const double v = (double)warp_sum(acc), v2 = (double)warp_sum(acc2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = v; red[1][threadIdx.x >> 5] = v2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0, s2 = 0;
        for (int w = 0; w < 8; ++w) { s += red[0][w]; s2 += red[1][w]; }
        const int64_t o = (int64_t)b * out_stride + (col_from_t ? (T - 1 - t) : 0);
        vlb[o] = (float)(s / (double)(chw4 * 4)) / 0.6931471805599453f;
        if (sq) sq[o] = (float)s2;
    }
}

// calc_prior (ddpm.py:367-389): KL(q(x_T | x) || N(0, I)) = 0.5 (-lv - 1 + exp(lv) + mean^2), mean = sqrt_ac[T-1] x; bits/dim per sample
__global__ void __launch_bounds__(256) prior_kl_kernel(const float4* __restrict__ x, float sa, float lv, float* __restrict__ out,
                                                       int64_t chw4) {
    pdl_sync();
    const int b = blockIdx.x;
    const float k0 = __fadd_rn(__fsub_rn(__fsub_rn(0.f, lv), 1.0f), expf(lv));      // logvar2 - logvar1 - 1 + exp(logvar1 - logvar2)
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < chw4; i += blockDim.x) {
        const float4 a = x[(int64_t)b * chw4 + i];
        const float m0 = __fmul_rn(sa, a.x), m1 = __fmul_rn(sa, a.y), m2 = __fmul_rn(sa, a.z), m3 = __fmul_rn(sa, a.w);
        acc += __fmul_rn(0.5f, __fadd_rn(k0, __fmul_rn(m0, m0))) + __fmul_rn(0.5f, __fadd_rn(k0, __fmul_rn(m1, m1))) +
               __fmul_rn(0.5f, __fadd_rn(k0, __fmul_rn(m2, m2))) + __fmul_rn(0.5f, __fadd_rn(k0, __fmul_rn(m3, m3)));
    }
    __shared__ double red[8];
    const double v = (double)warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        out[b] = (float)(s / (double)(chw4 * 4)) / 0.6931471805599453f;
    }
}

// ---------------------------------------------------------------------------------------------
// fix_samples (utils/eval_helpers.py:37-41, utils/utils.py:16-24): per-image min-max normalisation, x255, NCHW -> NHWC.
// One block per image: pass 1 finds min / max (the image, 786 KB at 3x256x256, then sits in L2), pass 2 writes pixels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) fix_samples_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
    pdl_sync();
    const int b = blockIdx.x;
    const float* src = x + (int64_t)b * C * HW;
    const int n = C * HW;
    float lo = INFINITY, hi = -INFINITY;
    if ((n & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
            const float4 v = s4[i];
            lo = fminf(fminf(lo, fminf(v.x, v.y)), fminf(v.z, v.w));
            hi = fmaxf(fmaxf(hi, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) { lo = fminf(lo, src[i]); hi = fmaxf(hi, src[i]); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float s_lo[16], s_hi[16];
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    lo = s_lo[0]; hi = s_hi[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
    const float range = __fsub_rn(hi, lo);
    float* dst = y + (int64_t)b * C * HW;
    for (int p = threadIdx.x; p < HW; p += blockDim.x)          // reads coalesced per channel plane, writes C contiguous floats per pixel
        for (int c = 0; c < C; ++c)
            dst[(int64_t)p * C + c] = __fmul_rn(__fdiv_rn(__fsub_rn(src[(int64_t)c * HW + p], lo), range), 255.f);
}

}  // namespace dd

using namespace dd;

static inline int grid_cap(int64_t n, int threads) {
    const int64_t g = (n + threads - 1) / threads, cap = 8LL * num_sms();
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" {

int dd_q_sample_step(const float* x, const float* noise, int64_t noise_step_stride, int noise_period, const float* tab,
                     const int32_t* t_idx, int T, float* out, int B, int64_t chw, void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0 && noise_step_stride % 4 == 0, "q_sample_step: chw=%lld must be a multiple of 4", (long long)chw);
    const int64_t n4 = (int64_t)B * chw / 4;
    launch_pdl(q_sample_step_kernel, dim3(grid_cap(n4, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)x,
               (const float4*)noise, noise_step_stride / 4, noise_period, (const EvalCoef*)tab, t_idx, T, (float4*)out, n4);
    return check_launch("q_sample_step");
}

int dd_vlb_terms(const float* x, const float* x_t, const float* eps_hat, const float* eps, int64_t eps_step_stride, int eps_period,
                 const float* tab, const int32_t* t_idx, int t_stride, int T, float* vlb, float* sq, int64_t out_stride,
                 int col_from_t, int B, int64_t chw, void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0 && eps_step_stride % 4 == 0, "vlb_terms: chw=%lld must be a multiple of 4", (long long)chw);
    DD_REQUIRE((sq == nullptr) || (eps != nullptr), "vlb_terms: the squared-error output needs eps");
    launch_pdl(vlb_terms_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (const float4*)x_t,
               (const float4*)eps_hat, (const float4*)eps, eps_step_stride / 4, eps_period, (const EvalCoef*)tab, t_idx, t_stride, T,
               vlb, sq, out_stride, col_from_t, chw / 4);
    return check_launch("vlb_terms");
}

int dd_prior_kl(const float* x, float sqrt_ac_last, float log_1mac_last, float* out, int B, int64_t chw, void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0, "prior_kl: chw=%lld must be a multiple of 4", (long long)chw);
    launch_pdl(prior_kl_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, sqrt_ac_last, log_1mac_last, out, chw / 4);
    return check_launch("prior_kl");
}

int dd_fix_samples(const float* x, float* y, int B, int C, int H, int W, void* stream) {
    DD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "fix_samples: empty batch");
    launch_pdl(fix_samples_kernel, dim3(B), dim3(512), 0, (cudaStream_t)stream, x, y, C, H * W);
    return check_launch("fix_samples");
}

}  // extern "C"
