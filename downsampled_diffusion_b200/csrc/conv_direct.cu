// Generic implicit-GEMM convolution on CUDA cores (fp32 accumulate).
// Used by the fp32 validation mode, by shapes the tcgen05 kernel does not tile (e.g. 28x28 MNIST),
// and by the 32/64-channel down/up-sampling nets.  64 pixels x 64 output channels per CTA,
// K (= taps * C_in) consumed 16 at a time through shared memory, 4x4 outputs per thread.
#include "common.cuh"

namespace dd {

struct ConvDirectArgs {
    const void* x; const void* x2;
    const float* w; const float* bias; const void* residual; void* y;
    int C1, C2, B, H, W, Ho, Wo, Cout;
    int ksize, stride, pad, mode, flags;
    int K;          // taps * (C1 + C2)
    int64_t M;      // B * Ho * Wo
};

template <typename TI>
__device__ __forceinline__ float conv_fetch(const ConvDirectArgs& a, int b, int oh, int ow, int k) {
    const int Cin = a.C1 + a.C2;
    const int tap = k / Cin, c = k - tap * Cin;
    int ih, iw;
    if (a.mode == 0) {
        const int ky = tap / a.ksize, kx = tap - ky * a.ksize;
        ih = oh * a.stride - a.pad + ky;
        iw = ow * a.stride - a.pad + kx;
    } else {   // transposed convolution: oh = ih*stride - pad + ky  (ConvTranspose2d(4,2,1); dgrad of a strided conv)
        const int ky = tap / a.ksize, kx = tap - ky * a.ksize;
        const int th = oh + a.pad - ky, tw = ow + a.pad - kx;
        if ((th | tw) < 0 || (th % a.stride) || (tw % a.stride)) return 0.f;
        ih = th / a.stride; iw = tw / a.stride;
    }
    if (ih < 0 || ih >= a.H || iw < 0 || iw >= a.W) return 0.f;
    float v;
    if (a.flags & DD_CONV_IN_NCHW) {
        v = reinterpret_cast<const float*>(a.x)[(((int64_t)b * Cin + c) * a.H + ih) * a.W + iw];
    } else if (c < a.C1) {
        v = to_f(reinterpret_cast<const TI*>(a.x)[(((int64_t)b * a.H + ih) * a.W + iw) * a.C1 + c]);
    } else {
        v = to_f(reinterpret_cast<const TI*>(a.x2)[(((int64_t)b * a.H + ih) * a.W + iw) * a.C2 + (c - a.C1)]);
    }
    if (a.flags & DD_CONV_PRE_MISH) v = mish_f(v);
    return v;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) conv_direct_kernel(const ConvDirectArgs a) {
    pdl_sync();
    constexpr int TM = 64, TN = 64, TK = 16;
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int tx = tid & 15, ty = tid >> 4;      // tx -> 4 couts, ty -> 4 pixels

    // A-load assignment: pixel lm = tid / 4, k sub-range (tid % 4) * 4 .. +4
    const int lm = tid >> 2, lk = (tid & 3) * 4;
    const int64_t pm = m0 + lm;
    int pb = 0, poh = 0, pow_ = 0;
    const bool pvalid = pm < a.M;
    if (pvalid) {
        pow_ = (int)(pm % a.Wo);
        const int64_t r = pm / a.Wo;
        poh = (int)(r % a.Ho);
        pb = (int)(r / a.Ho);
    }
    // B-load assignment: k row = tid / 16, 4 couts at (tid % 16) * 4
    const int bk = tid >> 4, bn = (tid & 15) * 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.K; k0 += TK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + lk + j;
            As[lk + j][lm] = (pvalid && k < a.K) ? conv_fetch<TI>(a, pb, poh, pow_, k) : 0.f;
        }
        {
            const int k = k0 + bk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + bn + j;
                Bs[bk][bn + j] = (k < a.K && n < a.Cout) ? a.w[(int64_t)k * a.Cout + n] : 0.f;
            }
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
        const int64_t m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
        const int ow = (int)(m % a.Wo);
        const int64_t r = m / a.Wo;
        const int oh = (int)(r % a.Ho);
        const int b = (int)(r / a.Ho);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.Cout) continue;
            float v = acc[i][j];
            if (a.bias) v += a.bias[n];
            if (a.residual) v += to_f(reinterpret_cast<const TO*>(a.residual)[m * a.Cout + n]);
            if (a.flags & DD_CONV_TANH) v = tanhf(v);
            if (a.flags & DD_CONV_OUT_NCHW)
                reinterpret_cast<float*>(a.y)[(((int64_t)b * a.Cout + n) * a.Ho + oh) * a.Wo + ow] = v;
            else
                reinterpret_cast<TO*>(a.y)[m * a.Cout + n] = from_f<TO>(v);
        }
    }
}

}  // namespace dd

using namespace dd;

extern "C" int dd_conv_direct(const void* x, const void* x2, int C1, int C2, int in_dtype, const float* w,
                              const float* bias, const void* residual, void* y, int out_dtype, int B, int H, int W,
                              int Cout, int ksize, int stride, int pad, int mode, int flags, void* stream) {
    DD_REQUIRE(B > 0 && H > 0 && W > 0 && Cout > 0 && C1 > 0 && C2 >= 0, "conv_direct: bad sizes");
    DD_REQUIRE(mode == 0 || mode == 1, "conv_direct: bad mode %d", mode);
    DD_REQUIRE(!(flags & DD_CONV_IN_NCHW) || (x2 == nullptr && in_dtype == DD_F32), "conv_direct: NCHW input needs fp32, single source");
    DD_REQUIRE(!(flags & DD_CONV_OUT_NCHW) || (residual == nullptr), "conv_direct: NCHW output cannot take a residual");
    DD_REQUIRE((C2 == 0) == (x2 == nullptr), "conv_direct: x2/C2 mismatch");
    ConvDirectArgs a;
    a.x = x; a.x2 = x2; a.w = w; a.bias = bias; a.residual = residual; a.y = y;
    a.C1 = C1; a.C2 = C2; a.B = B; a.H = H; a.W = W; a.Cout = Cout;
    a.ksize = ksize; a.stride = stride; a.pad = pad; a.mode = mode; a.flags = flags;
    if (mode == 0) {
        a.Ho = (H + 2 * pad - ksize) / stride + 1;
        a.Wo = (W + 2 * pad - ksize) / stride + 1;
        a.K = ksize * ksize * (C1 + C2);
    } else {
        DD_REQUIRE(stride >= 1 && ksize >= stride, "conv_direct: bad transposed geometry");
        a.Ho = stride * H; a.Wo = stride * W;          // output_padding chosen so the size is exactly stride*H
        a.K = ksize * ksize * (C1 + C2);
    }
    a.M = (int64_t)B * a.Ho * a.Wo;
    dim3 grid((unsigned)((a.M + 63) / 64), (unsigned)((Cout + 63) / 64));
    cudaStream_t st = (cudaStream_t)stream;
    const int odt = (flags & DD_CONV_OUT_NCHW) ? DD_F32 : out_dtype;
    if (in_dtype == DD_F32 && odt == DD_F32) launch_pdl(conv_direct_kernel<float, float>, dim3(grid), dim3(256), 0, st, a);
    else if (in_dtype == DD_BF16 && odt == DD_BF16) launch_pdl(conv_direct_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, a);
    else if (in_dtype == DD_BF16 && odt == DD_F32) launch_pdl(conv_direct_kernel<__nv_bfloat16, float>, dim3(grid), dim3(256), 0, st, a);
    else if (in_dtype == DD_F32 && odt == DD_BF16) launch_pdl(conv_direct_kernel<float, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, a);
    else { dd::set_error("conv_direct: bad dtypes %d/%d", in_dtype, out_dtype); return DD_ERR_ARG; }
    return check_launch("conv_direct");
}
