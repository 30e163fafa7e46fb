// Optimizer step of the training loop's caller (trainers/trainer.py:69 Adam(lr), trainer_ddpm.py:128-135 /
// :243-250: clip_grad_norm_(params, 1.0) -> opt.step() -> opt.zero_grad() -> EMA.update): three launches over all ~380
// parameter tensors instead of ~2000 -- a multi-tensor squared-norm pass, a one-block finish that produces the clip
// coefficient on the device (no host round trip), and one pass that applies clip, Adam, the EMA update and the gradient
// reset (HBM-bound: 28 B read + 16 B written per parameter, +8 with the EMA).
// Arithmetic follows torch.optim.Adam's single-tensor path and torch.nn.utils.clip_grad_norm_ operation by operation in fp32.
#include "common.cuh"

namespace dd {

constexpr int OPT_COLS = 6;     // table row: {param, grad, exp_avg, exp_avg_sq, shadow (0 = none), numel}

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const uint64_t* __restrict__ table, const int32_t* __restrict__ chunks,
                                                          int chunk_elems, float* __restrict__ partial) {
    pdl_sync();
    const int ti = chunks[2 * blockIdx.x], ci = chunks[2 * blockIdx.x + 1];
    const float* __restrict__ g = reinterpret_cast<const float*>(table[OPT_COLS * ti + 1]);
    const int64_t n = (int64_t)table[OPT_COLS * ti + 5];
    const int64_t lo = (int64_t)ci * chunk_elems, hi = min(lo + (int64_t)chunk_elems, n);
    float acc = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int64_t hi4 = lo + ((hi - lo) & ~(int64_t)3);
        for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi4; i += 4 * (int64_t)blockDim.x) {
            const float4 v = *reinterpret_cast<const float4*>(g + i);
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (int64_t i = hi4 + threadIdx.x; i < hi; i += blockDim.x) acc += g[i] * g[i];
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += g[i] * g[i];
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

// out[0] = total L2 norm, out[1] = clamp(max_norm / (norm + 1e-6), max = 1)   (torch/nn/utils/clip_grad.py)
__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                                float* __restrict__ out) {
    pdl_sync();
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partial[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 32; ++w) s += red[w];
        const float norm = (float)sqrt(s);
        out[0] = norm;
        out[1] = max_norm > 0.f ? fminf(__fdiv_rn(max_norm, __fadd_rn(norm, 1e-6f)), 1.0f) : 1.0f;
    }
}

struct AdamArgs {
    float w1;            // 1 - beta1      (exp_avg.lerp_(grad, 1 - beta1))
    float beta2, w2;     // beta2, 1 - beta2 (exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2))
    float bc2_sqrt;      // sqrt(1 - beta2^step)
    float eps;
    float neg_step;      // -lr / (1 - beta1^step)
    float decay, omd;    // EMA: shadow*decay + (1-decay)*param
    int ema_mode;        // 0 none, 1 update (trainers/ema.py:36-44), 2 copy (EMA.reset during warm-up, trainer_ddpm.py:107-109)
    int zero_grad;       // write zeros over the gradient once consumed
};

__device__ __forceinline__ void adam_elem(float& p, float& g, float& m, float& v, float* s, float coef, const AdamArgs& a) {
    const float gc = __fmul_rn(g, coef);                                           // clip_grad_norm_: g.mul_(clip_coef_clamped)
    m = fmaf(a.w1, __fsub_rn(gc, m), m);                                           // lerp (weight < 0.5 branch)
    v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.w2, gc), gc));      // mul_, addcmul_
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);    // (exp_avg_sq.sqrt() / bc2_sqrt).add_(eps)
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(a.neg_step, m), denom));                  // addcdiv_(exp_avg, denom, value = -step_size)
    if (a.ema_mode == 1) *s = __fadd_rn(__fmul_rn(*s, a.decay), __fmul_rn(a.omd, p));
    else if (a.ema_mode == 2) *s = p;
    if (a.zero_grad) g = 0.f;
}

__global__ void __launch_bounds__(256) adam_ema_kernel(const uint64_t* __restrict__ table, const int32_t* __restrict__ chunks,
                                                       int chunk_elems, const float* __restrict__ norm_out, AdamArgs a) {
    pdl_sync();
    const int ti = chunks[2 * blockIdx.x], ci = chunks[2 * blockIdx.x + 1];
    const uint64_t* row = table + OPT_COLS * ti;
    float* __restrict__ p = reinterpret_cast<float*>(row[0]);
    float* __restrict__ g = reinterpret_cast<float*>(row[1]);
    float* __restrict__ m = reinterpret_cast<float*>(row[2]);
    float* __restrict__ v = reinterpret_cast<float*>(row[3]);
    float* __restrict__ s = reinterpret_cast<float*>(row[4]);
    if (s == nullptr) a.ema_mode = 0;
    const int64_t n = (int64_t)row[5];
    const float coef = norm_out ? norm_out[1] : 1.0f;
    const int64_t lo = (int64_t)ci * chunk_elems, hi = min(lo + (int64_t)chunk_elems, n);
    const bool vec = ((row[0] | row[1] | row[2] | row[3] | row[4]) & 15) == 0;
    int64_t tail = lo;
    if (vec) {
        const int64_t hi4 = lo + ((hi - lo) & ~(int64_t)3);
        for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi4; i += 4 * (int64_t)blockDim.x) {
            float4 pv = *reinterpret_cast<float4*>(p + i), gv = *reinterpret_cast<float4*>(g + i);
            float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.ema_mode == 1) sv = *reinterpret_cast<float4*>(s + i);
            adam_elem(pv.x, gv.x, mv.x, vv.x, &sv.x, coef, a);
            adam_elem(pv.y, gv.y, mv.y, vv.y, &sv.y, coef, a);
            adam_elem(pv.z, gv.z, mv.z, vv.z, &sv.z, coef, a);
            adam_elem(pv.w, gv.w, mv.w, vv.w, &sv.w, coef, a);
            *reinterpret_cast<float4*>(p + i) = pv;
            *reinterpret_cast<float4*>(m + i) = mv;
            *reinterpret_cast<float4*>(v + i) = vv;
            if (a.ema_mode) *reinterpret_cast<float4*>(s + i) = sv;
            if (a.zero_grad) *reinterpret_cast<float4*>(g + i) = gv;
        }
        tail = hi4;
    }
    for (int64_t i = tail + threadIdx.x; i < hi; i += blockDim.x) {
        float sv = (a.ema_mode == 1) ? s[i] : 0.f;
        float pv = p[i], gv = g[i], mv = m[i], vv = v[i];
        adam_elem(pv, gv, mv, vv, &sv, coef, a);
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (a.ema_mode) s[i] = sv;
        if (a.zero_grad) g[i] = gv;
    }
}

// ---------------------------------------------------------------------------------------------
// Multi-tensor re-layout: dst_seg[i] = code ? src[(code >> 32) - 1][code & 0xffffffff] : 0 for every element of every
// segment, one launch.  The training programs derive ~600 packed weight buffers (K-major matrices, flipped / transposed
// input-gradient forms, space-to-depth forms of the strided convs) from the fp32 master parameters after every optimizer
// step, and map the packed weight gradients back to parameter layout; each re-layout is a fixed permutation, so it is
// tabulated once (engine side, by pushing index codes through the packing functions) and replayed here instead of ~900
// strided-copy launches per step.  16 B of traffic per element (8 code, 4 gathered read, 4 write).
// segs: n_segs * 3 uint64 {dst pointer, first element in `codes`, element count}; blocks: n_blocks * 2 int32 {segment, chunk}.
// ---------------------------------------------------------------------------------------------
constexpr int GATHER_CHUNK = 4096;

__global__ void __launch_bounds__(256) gather_f32_kernel(const uint64_t* __restrict__ segs, const int32_t* __restrict__ blocks,
                                                         const unsigned long long* __restrict__ codes,
                                                         const uint64_t* __restrict__ src_table, float* dst0) {
    pdl_sync();
    const int si = blocks[2 * blockIdx.x], ci = blocks[2 * blockIdx.x + 1];
    float* __restrict__ dst = (si == 0 && dst0) ? dst0 : reinterpret_cast<float*>(segs[3 * si]);
    const unsigned long long* __restrict__ code = codes + segs[3 * si + 1];
    const int64_t n = (int64_t)segs[3 * si + 2];
    const int64_t lo = (int64_t)ci * GATHER_CHUNK, hi = min(lo + (int64_t)GATHER_CHUNK, n);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const unsigned long long c = code[i];
        const uint32_t ord = (uint32_t)(c >> 32);
        dst[i] = ord ? reinterpret_cast<const float*>(src_table[ord - 1])[(uint32_t)c] : 0.f;
    }
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_grad_norm(const uint64_t* table, const int32_t* chunks, int n_chunks, int chunk_elems, float max_norm, float* partial,
                 float* norm_out, void* stream) {
    DD_REQUIRE(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0, "grad_norm: bad chunking");
    launch_pdl(grad_sqnorm_kernel, dim3(n_chunks), dim3(256), 0, (cudaStream_t)stream, table, chunks, chunk_elems, partial);
    launch_pdl(grad_norm_finish_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, (const float*)partial, n_chunks, max_norm, norm_out);
    return check_launch("grad_norm");
}

int dd_adam_ema_step(const uint64_t* table, const int32_t* chunks, int n_chunks, int chunk_elems, const float* norm_out,
                     float one_minus_beta1, float beta2, float one_minus_beta2, float bias_correction2_sqrt, float eps,
                     float neg_step_size, int ema_mode, float decay, float one_minus_decay, int zero_grad, void* stream) {
    DD_REQUIRE(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0, "adam_ema_step: bad chunking");
    DD_REQUIRE(ema_mode >= 0 && ema_mode <= 2, "adam_ema_step: ema_mode must be 0, 1 or 2");
    DD_REQUIRE(one_minus_beta1 < 0.5f, "adam_ema_step: beta1 <= 0.5 takes lerp's other branch (not implemented)");
    AdamArgs a{one_minus_beta1, beta2, one_minus_beta2, bias_correction2_sqrt, eps, neg_step_size, decay, one_minus_decay, ema_mode,
               zero_grad};
    launch_pdl(adam_ema_kernel, dim3(n_chunks), dim3(256), 0, (cudaStream_t)stream, table, chunks, chunk_elems, norm_out, a);
    return check_launch("adam_ema_step");
}

int dd_gather_f32(const uint64_t* segs, const int32_t* blocks, int n_blocks, const uint64_t* codes, const uint64_t* src_table,
                  float* dst0, void* stream) {
    DD_REQUIRE(n_blocks > 0, "gather_f32: nothing to do");
    launch_pdl(gather_f32_kernel, dim3(n_blocks), dim3(256), 0, (cudaStream_t)stream, segs, blocks,
               reinterpret_cast<const unsigned long long*>(codes), src_table, dst0);
    return check_launch("gather_f32");
}

}  // extern "C"
