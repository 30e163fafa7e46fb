// 1x1 convolutions with a narrow side (<= 8 channels): the 3 -> 64 input layer and the 64 -> 3 (+tanh) output layer of the
// resampling nets (models/downsampled/convblocks.py:143-156, dddpm.py:92-112) on 256x256 maps.  With K or N = 3 there is
// nothing for a tensor core to do: these are HBM streaming kernels (one pass over the 64-channel tensor, 16-byte vectors,
// the narrow operand read from / written to the NCHW tensor the Python API exposes).  The generic tiled CUDA-core kernels
// spent 0.4 - 1.3 ms on each of them (launch list train_list_r01_e.txt) against ~0.1 ms of memory time.
#include "common.cuh"

namespace dd {

constexpr int THIN_MAX = 8;

// y[b][p][c] (+)= sum_s x[b][s][p] * w[s][c] + bias[c]          (narrow NCHW in, wide NHWC out)
__global__ void __launch_bounds__(256) thin_in_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                      float* __restrict__ y, int HW, int Cs, int cv, int accumulate, int64_t total) {
    pdl_sync();
    const int c4 = threadIdx.x % cv;                       // blockDim.x % cv == 0: the thread's channel vector is fixed
    float4 wr[THIN_MAX];
#pragma unroll
    for (int s = 0; s < THIN_MAX; ++s) wr[s] = s < Cs ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)s * cv * 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t ppb = blockDim.x / cv;                  // pixels per CTA pass
    const uint32_t step = gridDim.x * ppb, n = (uint32_t)total, hw = (uint32_t)HW;      // total < 2^31 / cv (checked by the host)
    // two pixels in flight per thread: the narrow loads of the second overlap the store of the first
    for (uint32_t pix = blockIdx.x * ppb + threadIdx.x / cv; pix < n; pix += 2 * step) {
        const uint32_t pix2 = pix + step;
        const bool two = pix2 < n;
        const uint32_t b = pix / hw, p = pix - b * hw;
        const uint32_t b2 = two ? pix2 / hw : b, p2 = two ? pix2 - b2 * hw : p;
        const float* xa = x + (int64_t)b * Cs * HW + p;
        const float* xb = x + (int64_t)b2 * Cs * HW + p2;
        float va[THIN_MAX], vb[THIN_MAX];
#pragma unroll
        for (int s = 0; s < THIN_MAX; ++s) {
            va[s] = s < Cs ? __ldg(xa + (int64_t)s * HW) : 0.f;
            vb[s] = s < Cs ? __ldg(xb + (int64_t)s * HW) : 0.f;
        }
        float4 a = bv, c = bv;
#pragma unroll
        for (int s = 0; s < THIN_MAX; ++s) {
            if (s < Cs) {
                a.x = fmaf(va[s], wr[s].x, a.x); a.y = fmaf(va[s], wr[s].y, a.y); a.z = fmaf(va[s], wr[s].z, a.z); a.w = fmaf(va[s], wr[s].w, a.w);
                c.x = fmaf(vb[s], wr[s].x, c.x); c.y = fmaf(vb[s], wr[s].y, c.y); c.z = fmaf(vb[s], wr[s].z, c.z); c.w = fmaf(vb[s], wr[s].w, c.w);
            }
        }
        float4* da = reinterpret_cast<float4*>(y) + (int64_t)pix * cv + c4;
        float4* dc = reinterpret_cast<float4*>(y) + (int64_t)pix2 * cv + c4;
        if (accumulate) {
            const float4 o = *da; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            if (two) { const float4 q = *dc; c.x += q.x; c.y += q.y; c.z += q.z; c.w += q.w; }
        }
        *da = a;
        if (two) *dc = c;
    }
}

// y[b][s][p] = act(sum_c x[b][p][c] * w[s][c] + bias[s])       (wide NHWC in, narrow NCHW out; cv lanes per pixel)
__global__ void __launch_bounds__(256) thin_out_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ y, int HW, int Cs, int cv, int do_tanh, int64_t total) {
    pdl_sync();
    const int c4 = threadIdx.x % cv;
    float4 wr[THIN_MAX];
#pragma unroll
    for (int s = 0; s < THIN_MAX; ++s) wr[s] = s < Cs ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)s * cv * 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int ppb = blockDim.x / cv;
    const int64_t rounds = (total + (int64_t)gridDim.x * ppb - 1) / ((int64_t)gridDim.x * ppb);       // whole warps stay in the shuffles
    for (int64_t k = 0; k < rounds; ++k) {
        const int64_t pix = (k * gridDim.x + blockIdx.x) * ppb + threadIdx.x / cv;
        const bool ok = pix < total;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = *(reinterpret_cast<const float4*>(x) + pix * cv + c4);
        float acc[THIN_MAX];
#pragma unroll
        for (int s = 0; s < THIN_MAX; ++s) acc[s] = v.x * wr[s].x + v.y * wr[s].y + v.z * wr[s].z + v.w * wr[s].w;
        for (int o = 1; o < cv; o <<= 1) {                   // cv is a power of two <= 32
#pragma unroll
            for (int s = 0; s < THIN_MAX; ++s)
                if (s < Cs) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
        }
        if (ok && c4 < Cs) {                                 // lane s of the pixel's group writes channel s
            const int64_t b = (uint32_t)pix / (uint32_t)HW;  // total < 2^31 (checked by the host): 32-bit division
            const int p = (int)(pix - b * HW);
            float r = 0.f;
#pragma unroll
            for (int s = 0; s < THIN_MAX; ++s) if (s == c4) r = acc[s];
            r += bias ? __ldg(bias + c4) : 0.f;
            y[(b * Cs + c4) * HW + p] = do_tanh ? tanhf(r) : r;
        }
    }
}

// dw[s][c] += sum_{b,p} narrow[b][s][p] * wide[b][p][c]   (dw element (s, c) at dw[s*ds + c*dc]);  dbias_wide[c] += sum wide
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const float* __restrict__ nar, const float* __restrict__ wide, float* __restrict__ dw,
                                                         int ds, int dc, float* __restrict__ dbias_wide, int HW, int Cs, int cv, int64_t total) {
    pdl_sync();
    __shared__ float4 s_red[256];
    const int c4 = threadIdx.x % cv, ppb = blockDim.x / cv, pl = threadIdx.x / cv;
    float4 acc[THIN_MAX], bsum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < THIN_MAX; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t pix = (int64_t)blockIdx.x * ppb + pl; pix < total; pix += (int64_t)gridDim.x * ppb) {
        const int64_t b = (uint32_t)pix / (uint32_t)HW;      // total < 2^31 (checked by the host): 32-bit division
        const int p = (int)(pix - b * HW);
        const float4 v = *(reinterpret_cast<const float4*>(wide) + pix * cv + c4);
        bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
        const float* nb = nar + b * Cs * HW + p;
#pragma unroll
        for (int s = 0; s < THIN_MAX; ++s) {
            if (s < Cs) {
                const float n = __ldg(nb + (int64_t)s * HW);
                acc[s].x = fmaf(n, v.x, acc[s].x); acc[s].y = fmaf(n, v.y, acc[s].y); acc[s].z = fmaf(n, v.z, acc[s].z); acc[s].w = fmaf(n, v.w, acc[s].w);
            }
        }
    }
    // reduce over the ppb pixel lanes that share a channel vector, then one atomic per output element and CTA
    for (int s = -1; s < Cs; ++s) {
        float4 val = bsum;
#pragma unroll
        for (int k = 0; k < THIN_MAX; ++k) if (k == s) val = acc[k];
        __syncthreads();
        s_red[threadIdx.x] = val;
        __syncthreads();
        if (pl == 0) {
            float4 t = s_red[c4];
            for (int k = 1; k < ppb; ++k) { const float4 u = s_red[k * cv + c4]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
            if (s < 0) {
                if (dbias_wide) {
                    atomicAdd(dbias_wide + 4 * c4, t.x); atomicAdd(dbias_wide + 4 * c4 + 1, t.y);
                    atomicAdd(dbias_wide + 4 * c4 + 2, t.z); atomicAdd(dbias_wide + 4 * c4 + 3, t.w);
                }
            } else {
                float* d = dw + (int64_t)s * ds + (int64_t)(4 * c4) * dc;
                atomicAdd(d, t.x); atomicAdd(d + dc, t.y); atomicAdd(d + 2 * dc, t.z); atomicAdd(d + 3 * dc, t.w);
            }
        }
    }
}

static inline bool thin_ok(int Cs, int Cw) {
    const int cv = Cw / 4;
    return Cs >= 1 && Cs <= THIN_MAX && Cw % 4 == 0 && cv >= 1 && cv <= 32 && (cv & (cv - 1)) == 0 && cv >= Cs;
}
static inline int thin_grid(int64_t pixels, int ppb) {
    int64_t g = (pixels + ppb - 1) / ppb, cap = (int64_t)num_sms() * 8;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_conv1x1_thin_in(const float* x_nchw, const float* w, const float* bias, float* y_nhwc, int B, int HW, int Cs, int Cout,
                       int accumulate, void* stream) {
    DD_REQUIRE(thin_ok(Cs, Cout) && B > 0 && HW > 0, "conv1x1_thin_in: narrow side %d (1..8), wide side %d (4..128, power of two)", Cs, Cout);
    DD_REQUIRE((int64_t)B * HW < (1LL << 31) / 32, "conv1x1_thin_in: %lld pixels exceed the 32-bit index range of the kernel", (long long)B * HW);
    const int cv = Cout / 4;
    const int64_t total = (int64_t)B * HW;
    launch_pdl(thin_in_kernel, dim3(thin_grid(total, 256 / cv)), dim3(256), 0, (cudaStream_t)stream, x_nchw, w, bias, y_nhwc, HW, Cs, cv, accumulate, total);
    return check_launch("conv1x1_thin_in");
}

int dd_conv1x1_thin_out(const float* x_nhwc, const float* w, const float* bias, float* y_nchw, int B, int HW, int Cin, int Cs, int do_tanh,
                        void* stream) {
    DD_REQUIRE(thin_ok(Cs, Cin) && B > 0 && HW > 0, "conv1x1_thin_out: narrow side %d (1..8), wide side %d (4..128, power of two)", Cs, Cin);
    DD_REQUIRE((int64_t)B * HW < (1LL << 31) / 32, "conv1x1_thin_out: %lld pixels exceed the 32-bit index range of the kernel", (long long)B * HW);
    const int cv = Cin / 4;
    const int64_t total = (int64_t)B * HW;
    launch_pdl(thin_out_kernel, dim3(thin_grid(total, 256 / cv)), dim3(256), 0, (cudaStream_t)stream, x_nhwc, w, bias, y_nchw, HW, Cs, cv, do_tanh, total);
    return check_launch("conv1x1_thin_out");
}

int dd_conv1x1_thin_wgrad(const float* narrow_nchw, const float* wide_nhwc, float* dw, int narrow_major, float* dbias_wide, int B, int HW,
                          int Cs, int Cw, void* stream) {
    DD_REQUIRE(thin_ok(Cs, Cw) && B > 0 && HW > 0, "conv1x1_thin_wgrad: narrow side %d (1..8), wide side %d (4..128, power of two)", Cs, Cw);
    DD_REQUIRE((int64_t)B * HW < (1LL << 31) / 32, "conv1x1_thin_wgrad: %lld pixels exceed the 32-bit index range of the kernel", (long long)B * HW);
    const int cv = Cw / 4;
    const int64_t total = (int64_t)B * HW;
    const int ds = narrow_major ? Cw : 1, dc = narrow_major ? 1 : Cs;         // dw is (Cs, Cw) or (Cw, Cs)
    int grid = thin_grid(total, 256 / cv);
    if (grid > 4 * num_sms()) grid = 4 * num_sms();                            // bounds the atomics per output element
    launch_pdl(thin_wgrad_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, narrow_nchw, wide_nhwc, dw, ds, dc, dbias_wide, HW, Cs, cv, total);
    return check_launch("conv1x1_thin_wgrad");
}

}  // extern "C"
