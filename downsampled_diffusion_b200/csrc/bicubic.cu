// 'deterministic' resampler of the reference (models/downsampled/convblocks.py:8-26, wrapper.py:22-24, 49-53):
// F.interpolate(size, mode='bicubic', align_corners=True) on NCHW fp32, down (256 -> 32) and up (32 -> 256), and its input
// gradient (the non-autoencoder dDDPM loss back-propagates the reconstruction error through the up-sampler,
// models/diffusion/dddpm.py:122-143).  Cubic convolution with A = -0.75, source index = dst * (in-1)/(out-1), the four taps
// clamped to the image: the arithmetic of ATen's upsample_bicubic2d, tap weights per output row / column computed once per
// thread.  HBM-trivial (16 taps per output element, all L1/L2 hits); one thread per output pixel, planes in grid.y.
#include "common.cuh"

namespace dd {

struct CubicTaps { int idx[4]; float w[4]; };

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ CubicTaps cubic_taps(int dst, int in_size, float scale) {
    const float A = -0.75f;
    const float real = scale * (float)dst;                       // align_corners source index
    int base = (int)floorf(real);
    base = base < in_size - 1 ? base : in_size - 1;
    float t = real - (float)base;
    t = fminf(fmaxf(t, 0.f), 1.f);
    CubicTaps c;
    c.w[0] = cubic2(t + 1.f, A);
    c.w[1] = cubic1(t, A);
    const float u = 1.f - t;
    c.w[2] = cubic1(u, A);
    c.w[3] = cubic2(u + 1.f, A);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = base - 1 + j;
        c.idx[j] = i < 0 ? 0 : (i > in_size - 1 ? in_size - 1 : i);
    }
    return c;
}

__global__ void __launch_bounds__(256) bicubic2d_kernel(const float* __restrict__ x, float* __restrict__ y, int Hin, int Win,
                                                        int Hout, int Wout, float sh, float sw) {
    pdl_sync();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= Hout * Wout) return;
    const int oy = o / Wout, ox = o - oy * Wout;
    const CubicTaps ty = cubic_taps(oy, Hin, sh), tx = cubic_taps(ox, Win, sw);
    const float* src = x + (int64_t)blockIdx.y * Hin * Win;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float* row = src + (int64_t)ty.idx[i] * Win;
        float r = row[tx.idx[0]] * tx.w[0];
        r += row[tx.idx[1]] * tx.w[1];
        r += row[tx.idx[2]] * tx.w[2];
        r += row[tx.idx[3]] * tx.w[3];
        acc = i == 0 ? r * ty.w[0] : acc + r * ty.w[i];
    }
    y[(int64_t)blockIdx.y * Hout * Wout + o] = acc;
}

// gx += adjoint: every output pixel scatters its gradient to its 16 taps (gx zeroed by the caller side of the entry point)
__global__ void __launch_bounds__(256) bicubic2d_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int Hin, int Win,
                                                            int Hout, int Wout, float sh, float sw) {
    pdl_sync();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= Hout * Wout) return;
    const int oy = o / Wout, ox = o - oy * Wout;
    const CubicTaps ty = cubic_taps(oy, Hin, sh), tx = cubic_taps(ox, Win, sw);
    const float g = gy[(int64_t)blockIdx.y * Hout * Wout + o];
    float* dst = gx + (int64_t)blockIdx.y * Hin * Win;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dst + (int64_t)ty.idx[i] * Win + tx.idx[j], g * ty.w[i] * tx.w[j]);
}

static inline float cubic_scale(int in_size, int out_size) { return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f; }

}  // namespace dd

using namespace dd;

extern "C" {

int dd_bicubic2d(const float* x, float* y, int planes, int Hin, int Win, int Hout, int Wout, void* stream) {
    DD_REQUIRE(planes > 0 && planes <= 65535 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "bicubic2d: bad shape (%d planes)", planes);
    dim3 grid((Hout * Wout + 255) / 256, planes);
    launch_pdl(bicubic2d_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, y, Hin, Win, Hout, Wout, cubic_scale(Hin, Hout),
               cubic_scale(Win, Wout));
    return check_launch("bicubic2d");
}

int dd_bicubic2d_bwd(const float* gy, float* gx, int planes, int Hin, int Win, int Hout, int Wout, void* stream) {
    DD_REQUIRE(planes > 0 && planes <= 65535 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "bicubic2d_bwd: bad shape (%d planes)", planes);
    cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)planes * Hin * Win, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("bicubic2d_bwd: memset: %s", cudaGetErrorString(e)); return DD_ERR_CUDA; }
    dim3 grid((Hout * Wout + 255) / 256, planes);
    launch_pdl(bicubic2d_bwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream, gy, gx, Hin, Win, Hout, Wout, cubic_scale(Hin, Hout),
               cubic_scale(Win, Wout));
    return check_launch("bicubic2d_bwd");
}

}  // extern "C"
