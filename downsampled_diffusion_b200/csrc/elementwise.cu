// HBM-bound elementwise / layout kernels of the dDDPM hot path.
// Arithmetic mirrors ATen's op-by-op rounding (explicit __fmul_rn/__fadd_rn, no FMA contraction)
// so the posterior update, q_sample and EMA are bit-identical to the reference's fp32 results.
#include "common.cuh"
#include <stdarg.h>

namespace dd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DD_ERR_CUDA;
    }
    return DD_OK;
}

// ---------------------------------------------------------------------------------------------
// q_sample  (ddpm.py:270-273)
// ---------------------------------------------------------------------------------------------
__global__ void q_sample_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                                const int64_t* __restrict__ t, const float* __restrict__ sa,
                                const float* __restrict__ sb, float4* __restrict__ out, int64_t chw4) {
    pdl_sync();
    const int b = blockIdx.y;
    const float a = sa[t[b]], c = sb[t[b]];
    const int64_t base = (int64_t)b * chw4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chw4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 xv = x[base + i], ev = eps[base + i], o;
        o.x = __fadd_rn(__fmul_rn(a, xv.x), __fmul_rn(c, ev.x));
        o.y = __fadd_rn(__fmul_rn(a, xv.y), __fmul_rn(c, ev.y));
        o.z = __fadd_rn(__fmul_rn(a, xv.z), __fmul_rn(c, ev.z));
        o.w = __fadd_rn(__fmul_rn(a, xv.w), __fmul_rn(c, ev.w));
        out[base + i] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// posterior step (ddpm.py:149-158, 177-185, 217-227): 4 fp32 streams = 16 B / element.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float post1(float xt, float e, float z, float c0, float c1, float c2, float c3,
                                       float sig, int clip) {
    float x0 = __fsub_rn(__fmul_rn(c0, xt), __fmul_rn(c1, e));
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float mean = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, xt));
    return __fadd_rn(mean, __fmul_rn(sig, z));
}

__global__ void posterior_step_kernel(const float4* __restrict__ xt, const float4* __restrict__ eh,
                                      const float4* __restrict__ noise, const float* __restrict__ coef,
                                      const int32_t* __restrict__ t_idx, int t_stride, int64_t noise_step_stride4,
                                      int T, int noise_period, int clip, float4* __restrict__ out, int64_t chw4) {
    pdl_sync();
    const int b = blockIdx.y;
    const int t = t_idx[(int64_t)b * t_stride];
    const float* c = coef + (int64_t)t * 5;
    const float c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    const float sig = (t == 0) ? 0.f : c[4];   // nonzero_mask * exp(0.5*logvar), ddpm.py:224-227
    int step = T - 1 - t;                      // index of this step's pre-drawn noise (ddpm.py:223 draw order)
    if (noise_period > 0) step %= noise_period;
    const float4* nz = noise + (int64_t)step * noise_step_stride4;
    const int64_t base = (int64_t)b * chw4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chw4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = xt[base + i], e = eh[base + i], z = nz[base + i], o;
        o.x = post1(x.x, e.x, z.x, c0, c1, c2, c3, sig, clip);
        o.y = post1(x.y, e.y, z.y, c0, c1, c2, c3, sig, clip);
        o.z = post1(x.z, e.z, z.z, c0, c1, c2, c3, sig, clip);
        o.w = post1(x.w, e.w, z.w, c0, c1, c2, c3, sig, clip);
        out[base + i] = o;
    }
}

__global__ void predict_x0_kernel(const float4* __restrict__ xt, const float4* __restrict__ eps,
                                  const int64_t* __restrict__ t, const float* __restrict__ ra,
                                  const float* __restrict__ rb, int clip, float4* __restrict__ out, int64_t chw4) {
    pdl_sync();
    const int b = blockIdx.y;
    const float a = ra[t[b]], c = rb[t[b]];
    const int64_t base = (int64_t)b * chw4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chw4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = xt[base + i], e = eps[base + i], o;
        o.x = __fsub_rn(__fmul_rn(a, x.x), __fmul_rn(c, e.x));
        o.y = __fsub_rn(__fmul_rn(a, x.y), __fmul_rn(c, e.y));
        o.z = __fsub_rn(__fmul_rn(a, x.z), __fmul_rn(c, e.z));
        o.w = __fsub_rn(__fmul_rn(a, x.w), __fmul_rn(c, e.w));
        if (clip) {
            o.x = fminf(fmaxf(o.x, -1.f), 1.f); o.y = fminf(fmaxf(o.y, -1.f), 1.f);
            o.z = fminf(fmaxf(o.z, -1.f), 1.f); o.w = fminf(fmaxf(o.w, -1.f), 1.f);
        }
        out[base + i] = o;
    }
}

__global__ void tick_kernel(int32_t* t, int n) {
    pdl_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] -= 1;
}

// ---------------------------------------------------------------------------------------------
// per-sample squared error sum (ddpm.py:279-283)
// ---------------------------------------------------------------------------------------------
__global__ void mse_rowsum_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                  float* __restrict__ out, int64_t chw4, float scale) {
    pdl_sync();
    const int row = blockIdx.x;
    const int64_t base = (int64_t)row * chw4;
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < chw4; i += blockDim.x) {
        float4 x = a[base + i], y = b[base + i];
        float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    __shared__ double red[32];
    double v = (double)warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        out[row] = (float)(s * (double)scale);
    }
}

__global__ void mse_rowsum_bwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                      const float* __restrict__ w, float4* __restrict__ g, int64_t chw4, float gscale) {
    pdl_sync();
    const int row = blockIdx.y;
    const float s = 2.f * gscale * (w ? w[row] : 1.f);
    const int64_t base = (int64_t)row * chw4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chw4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = a[base + i], y = b[base + i], o;
        o.x = (y.x - x.x) * s; o.y = (y.y - x.y) * s; o.z = (y.z - x.z) * s; o.w = (y.w - x.w) * s;
        g[base + i] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// multi-tensor EMA (trainers/ema.py:36-44): 12 B / parameter, one launch for every tensor.
// ---------------------------------------------------------------------------------------------
__global__ void ema_update_kernel(const uint64_t* __restrict__ table, const int32_t* __restrict__ chunks,
                                  int chunk_elems, float decay, float omd) {
    pdl_sync();
    const int ti = chunks[2 * blockIdx.x], ci = chunks[2 * blockIdx.x + 1];
    float* __restrict__ s = reinterpret_cast<float*>(table[3 * ti]);
    const float* __restrict__ p = reinterpret_cast<const float*>(table[3 * ti + 1]);
    const int64_t n = (int64_t)table[3 * ti + 2];
    const int64_t lo = (int64_t)ci * chunk_elems;
    const int64_t hi = min(lo + (int64_t)chunk_elems, n);
    const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(p)) & 15) == 0;
    if (vec) {
        const int64_t hi4 = lo + ((hi - lo) & ~(int64_t)3);
        for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi4; i += 4 * (int64_t)blockDim.x) {
            float4 sv = *reinterpret_cast<float4*>(s + i);
            const float4 pv = *reinterpret_cast<const float4*>(p + i);
            sv.x = __fadd_rn(__fmul_rn(sv.x, decay), __fmul_rn(omd, pv.x));
            sv.y = __fadd_rn(__fmul_rn(sv.y, decay), __fmul_rn(omd, pv.y));
            sv.z = __fadd_rn(__fmul_rn(sv.z, decay), __fmul_rn(omd, pv.z));
            sv.w = __fadd_rn(__fmul_rn(sv.w, decay), __fmul_rn(omd, pv.w));
            *reinterpret_cast<float4*>(s + i) = sv;
        }
        for (int64_t i = hi4 + threadIdx.x; i < hi; i += blockDim.x)
            s[i] = __fadd_rn(__fmul_rn(s[i], decay), __fmul_rn(omd, p[i]));
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
            s[i] = __fadd_rn(__fmul_rn(s[i], decay), __fmul_rn(omd, p[i]));
    }
}

// ---------------------------------------------------------------------------------------------
// layout kernels
// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC T.  One thread per (b, pixel); reads are coalesced per channel plane.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int C, int HW) {
    pdl_sync();
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const float* xb = x + (int64_t)b * C * HW + p;
    T* yb = y + ((int64_t)b * HW + p) * C;
    for (int c = 0; c < C; ++c) yb[c] = from_f<T>(xb[(int64_t)c * HW]);
}

// NCHW fp32 (C <= Cp channels) -> NHWC bf16 with the channels zero-padded to Cp (a multiple of 8): the U-Net's input as an
// ordinary 64-channel activation, so that its first convolution is a regular 3x3 tensor-core layer with the fused GroupNorm
// epilogue (one 16-byte store per thread, consecutive threads consecutive channel groups of a pixel).
__global__ void __launch_bounds__(256) nchw_to_nhwc_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int HW, int Cp,
                                                               int64_t total_vec) {
    pdl_sync();
    const int groups = Cp >> 3;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total_vec; v += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(v % groups);
        const int64_t pixg = v / groups;                    // b * HW + pixel
        const int64_t b = pixg / HW;
        const int pix = (int)(pixg - b * HW);
        Vec<__nv_bfloat16> o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            o.v[j] = c < C ? __ldg(x + (b * C + c) * HW + pix) : 0.f;
        }
        o.store(y + v * 8);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int C, int HW) {
    pdl_sync();
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const T* xb = x + ((int64_t)b * HW + p) * C;
    float* yb = y + (int64_t)b * C * HW + p;
    for (int c = 0; c < C; ++c) yb[(int64_t)c * HW] = to_f(xb[c]);
}

// fp32 space-to-depth / depth-to-space, NHWC: full (B, 2h, 2w, C) <-> packed (B, h, w, 4C) with packed channel
// (py*2 + px)*C + c = full[2i + py][2j + px][c].  Lets the stride-2 and the transposed convolutions of the U-Net
// (blocks.py:35,44) run as dense 3x3 stride-1 convolutions on the tensor cores (autograd.py).
__global__ void s2d_f32_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int h, int w, int cv, int to_packed,
                               int64_t total) {
    pdl_sync();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // i indexes the PACKED tensor: (b, i_, j_, plane, c4)
        const int c4 = (int)(i % cv);
        int64_t r = i / cv;
        const int pl = (int)(r & 3); r >>= 2;
        const int j = (int)(r % w); r /= w;
        const int ii = (int)(r % h);
        const int64_t b = r / h;
        const int64_t full = (((b * (2 * h) + 2 * ii + (pl >> 1)) * (2 * w)) + 2 * j + (pl & 1)) * cv + c4;
        if (to_packed) dst[i] = src[full];
        else dst[full] = src[i];
    }
}

// NHWC fp32 -> channel-major fp32 with padding and column shifts:  y[s][b][c][h + hpad][w] = x[b][h][w + s - nshift/2][c]
// (zero outside the map), rows Wp >= W wide, s < nshift (1 or 3).  This is the K-major operand layout of the TF32
// weight-gradient GEMM (pixels = K): a row of 32 pixels is one 128-byte operand row.  TMA needs 16-byte aligned box
// origins, so the +-1 column shift of a 3x3 tap cannot be a coordinate of the innermost (pixel) dimension -- it is baked
// into three copies written from one read; the row shift stays a coordinate (one zero row above and below).
// 32 x 32 (pixel x channel) tiles through shared memory: coalesced 128-byte reads along c, 128-byte writes along w.
constexpr int CHWP_ROWS = 4;          // image rows per CTA
__global__ void __launch_bounds__(256) nhwc_to_chw_pad_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W,
                                                              int Wp, int hpad, int nshift, float* __restrict__ colsum) {
    pdl_sync();
    __shared__ float tile[CHWP_ROWS][34][33];                         // [row][pixel w0 - 1 .. w0 + 32][channel]
    __shared__ float s_red[8][32];
    const int w0 = blockIdx.x * 32, hh0 = blockIdx.y * CHWP_ROWS, Hp = H + 2 * hpad;
    const int cblocks = C >> 5, b = blockIdx.z / cblocks, c0 = (blockIdx.z % cblocks) * 32;
    // ---- reads: 16 bytes (4 channels) per thread, 8 threads per pixel: 34 pixels x CHWP_ROWS rows
    for (int i = threadIdx.x; i < CHWP_ROWS * 34 * 8; i += 256) {
        const int c4 = i & 7, pix = (i >> 3) % 34, r = i / (34 * 8);
        const int h = hh0 + r - hpad, w = w0 - 1 + pix;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h >= 0 && h < H && w >= 0 && w < W) v = *reinterpret_cast<const float4*>(x + (((int64_t)b * H + h) * W + w) * C + c0 + c4 * 4);
        float* t = &tile[r][pix][c4 * 4];
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
    __syncthreads();
    if (colsum) {
        // per-channel sum of this tile's own pixels (the bias gradient when x is an output gradient: ddpm autograd of
        // conv + bias), fused here because this kernel reads every element of dY anyway
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < CHWP_ROWS; ++r) {
            const int h = hh0 + r - hpad;
            if (h < 0 || h >= H) continue;
            for (int pix = 1 + wrp; pix <= 32; pix += 8)
                if (w0 - 1 + pix < W) a += tile[r][pix][lane];
        }
        s_red[wrp][lane] = a;
        __syncthreads();
        if (wrp == 0) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += s_red[k][lane];
            atomicAdd(colsum + c0 + lane, t);
        }
    }
    // ---- writes: 16 bytes (4 pixels of one channel) per thread; the column shift is the pixel offset into the tile
    const int64_t copy = (int64_t)B * C * Hp * Wp;
    const int per_copy = CHWP_ROWS * 32 * 8;
    for (int i = threadIdx.x; i < nshift * per_copy; i += 256) {
        const int w4 = i & 7, c = (i >> 3) & 31, r = (i >> 8) % CHWP_ROWS, s = i / per_copy;
        const int hh = hh0 + r;
        if (hh >= Hp) continue;
        const int p0 = w4 * 4 + 1 + s - (nshift >> 1);                // tile pixel of output column w0 + 4*w4
        const float4 v = make_float4(tile[r][p0][c], tile[r][p0 + 1][c], tile[r][p0 + 2][c], tile[r][p0 + 3][c]);
        *reinterpret_cast<float4*>(y + s * copy + (((int64_t)b * C + c0 + c) * Hp + hh) * Wp + w0 + w4 * 4) = v;
    }
}

// NCHW fp32 -> bf16 im2col rows (B*H*W, kpad), 3x3 pad 1.  One CTA per 32-pixel row segment: the 3 x 34 x C input
// halo is read with coalesced loads into shared memory, the segment's 32 x kpad output block is contiguous in
// memory and written with consecutive 16-byte stores (the first version read 8 strided floats per thread: 23 us
// for 17 MB -- profiles/README.md).
constexpr int IM2COL_TW = 32;
template <int CT>          // CT > 0: compile-time channel count (no divisions in the fill loop); 0: run-time C
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        int C_rt, int H, int W, int kpad) {
    pdl_sync();
    const int C = CT > 0 ? CT : C_rt;
    extern __shared__ float s_halo[];                  // [C][3][IM2COL_TW + 2]
    constexpr int HP = IM2COL_TW + 2;
    const int b = blockIdx.y;
    const int segs = (W + IM2COL_TW - 1) / IM2COL_TW;
    const int h = blockIdx.x / segs, w0 = (blockIdx.x % segs) * IM2COL_TW;
    const int HW = H * W;
    const float* xb = x + (int64_t)b * C * HW;
    for (int i = threadIdx.x; i < C * 3 * HP; i += 256) {
        const int c = i / (3 * HP), rr = (i / HP) % 3, cc = i % HP;
        const int hh = h + rr - 1, ww = w0 + cc - 1;
        s_halo[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xb + (int64_t)c * HW + hh * W + ww) : 0.f;
    }
    __syncthreads();
    const int groups = kpad >> 3;
    const int npix = min(IM2COL_TW, W - w0);
    __nv_bfloat16* yb = y + ((int64_t)b * HW + (int64_t)h * W + w0) * kpad;
    for (int v = threadIdx.x; v < npix * groups; v += 256) {
        const int pix = v / groups, g = v % groups;
        Vec<__nv_bfloat16> o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = g * 8 + j;
            float val = 0.f;
            if (k < 9 * C) {
                const int tap = k / C, c = k - tap * C;
                val = s_halo[(c * 3 + tap / 3) * HP + pix + tap % 3];
            }
            o.v[j] = val;
        }
        o.store(yb + (int64_t)v * 8);
    }
}

// C = 8 (the dDDPM latent): output vector g of a pixel is exactly tap g's eight channels (k = 8 g + c), so a thread gathers its
// 16 bytes straight from the eight NCHW planes (2 MB input: L2-resident, reads coalesced along w) -- no shared memory, no barrier,
// one 16-byte store per thread, consecutive threads consecutive vectors.  Columns 72 .. kpad-1 are written as zeros.
__global__ void __launch_bounds__(256) im2col3x3_c8_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int H, int W, int kpad,
                                                           int64_t total_vec) {
    pdl_sync();
    const int groups = kpad >> 3, HW = H * W;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total_vec; v += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(v % groups);
        const int64_t pixg = v / groups;                    // b * HW + h * W + w
        const int w = (int)(pixg % W), h = (int)((pixg / W) % H);
        const int64_t b = pixg / HW;
        Vec<__nv_bfloat16> o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
        if (g < 9) {
            const int hh = h + g / 3 - 1, ww = w + g % 3 - 1;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                const float* src = x + (b * 8) * HW + hh * W + ww;
#pragma unroll
                for (int c = 0; c < 8; ++c) o.v[c] = __ldg(src + (int64_t)c * HW);
            }
        }
        o.store(y + v * 8);
    }
}

template <typename T>
__global__ void avgpool2_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, int64_t total_vec) {
    pdl_sync();
    constexpr int VN = Vec<T>::N;
    const int Ho = H >> 1, Wo = W >> 1, cv = C / VN;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VN;
        int64_t r = i / cv;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int64_t b = r / Ho;
        const T* p = x + (((b * H + 2 * ho) * W + 2 * wo) * (int64_t)C) + c;
        Vec<T> a, bq, cq, d, o;
        a.load(p); bq.load(p + C); cq.load(p + (int64_t)W * C); d.load(p + (int64_t)W * C + C);
#pragma unroll
        for (int j = 0; j < VN; ++j) o.v[j] = (a.v[j] + bq.v[j] + cq.v[j] + d.v[j]) * 0.25f;
        o.store(y + (((b * Ho + ho) * Wo + wo) * (int64_t)C) + c);
    }
}

template <typename T>
__global__ void upsample2_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, int64_t total_vec) {
    pdl_sync();
    constexpr int VN = Vec<T>::N;
    const int Ho = H * 2, Wo = W * 2, cv = C / VN;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VN;
        int64_t r = i / cv;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int64_t b = r / Ho;
        const uint4 v = *reinterpret_cast<const uint4*>(x + (((b * H + (ho >> 1)) * W + (wo >> 1)) * (int64_t)C) + c);
        *reinterpret_cast<uint4*>(y + (((b * Ho + ho) * Wo + wo) * (int64_t)C) + c) = v;
    }
}

// NHWC bf16 -> (4, B, H/2, W/2, C) parity planes.
__global__ void s2d_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                           int B, int H, int W, int C, int64_t total_vec) {
    pdl_sync();
    const int cv = C >> 3, Ho = H >> 1, Wo = W >> 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * 8;
        int64_t r = i / cv;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int64_t b = r / H;
        const int plane = (h & 1) * 2 + (w & 1);
        const uint4 v = *reinterpret_cast<const uint4*>(x + i * 8);
        *reinterpret_cast<uint4*>(y + ((((int64_t)plane * B + b) * Ho + (h >> 1)) * Wo + (w >> 1)) * (int64_t)C + c) = v;
    }
}

static inline int grid_for(int64_t n, int threads, int cap_mult = 8) {
    int64_t g = (n + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * cap_mult;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace dd

using namespace dd;

extern "C" {

const char* dd_last_error(void) { return dd::g_err; }
int dd_version(void) { return 100; }

int dd_device_ok(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

int dd_q_sample(const float* x, const float* eps, const int64_t* t, const float* sa, const float* sb,
                float* out, int B, int64_t chw, void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0, "q_sample: chw=%lld must be a multiple of 4", (long long)chw);
    dim3 grid(grid_for(chw / 4, 256, 4), B);
    launch_pdl(q_sample_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (const float4*)eps, t, sa, sb,
                                                           (float4*)out, chw / 4);
    return check_launch("q_sample");
}

int dd_posterior_step(const float* x_t, const float* eps_hat, const float* noise, const float* coef,
                      const int32_t* t_idx, int t_stride, int64_t noise_step_stride, int T, int noise_period, int clip,
                      float* x_out, int B, int64_t chw, void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0 && noise_step_stride % 4 == 0, "posterior_step: chw=%lld must be a multiple of 4",
               (long long)chw);
    dim3 grid(grid_for(chw / 4, 256, 4), B);
    launch_pdl(posterior_step_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
        (const float4*)x_t, (const float4*)eps_hat, (const float4*)noise, coef, t_idx, t_stride,
        noise_step_stride / 4, T, noise_period, clip, (float4*)x_out, chw / 4);
    return check_launch("posterior_step");
}

int dd_predict_x0(const float* x_t, const float* eps, const int64_t* t, const float* ra, const float* rb, int clip,
                  float* out, int B, int64_t chw, void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0, "predict_x0: chw=%lld must be a multiple of 4", (long long)chw);
    dim3 grid(grid_for(chw / 4, 256, 4), B);
    launch_pdl(predict_x0_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float4*)x_t, (const float4*)eps, t, ra, rb, clip,
                                                             (float4*)out, chw / 4);
    return check_launch("predict_x0");
}

int dd_tick(int32_t* t_idx, int n, void* stream) {
    launch_pdl(tick_kernel, dim3((n + 127) / 128), dim3(128), 0, (cudaStream_t)stream, t_idx, n);
    return check_launch("tick");
}

int dd_mse_rowsum(const float* a, const float* b, float* out, int B, int64_t chw, float scale, void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0, "mse_rowsum: chw=%lld must be a multiple of 4", (long long)chw);
    launch_pdl(mse_rowsum_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, (const float4*)a, (const float4*)b, out, chw / 4, scale);
    return check_launch("mse_rowsum");
}

int dd_mse_rowsum_bwd(const float* a, const float* b, const float* w, float* grad_b, int B, int64_t chw, float gscale,
                      void* stream) {
    DD_REQUIRE(chw % 4 == 0 && B > 0, "mse_rowsum_bwd: chw=%lld must be a multiple of 4", (long long)chw);
    dim3 grid(grid_for(chw / 4, 256, 4), B);
    launch_pdl(mse_rowsum_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float4*)a, (const float4*)b, w,
                                                                 (float4*)grad_b, chw / 4, gscale);
    return check_launch("mse_rowsum_bwd");
}

int dd_ema_update(const uint64_t* table, const int32_t* chunks, int n_chunks, int chunk_elems, float decay,
                  float one_minus_decay, void* stream) {
    DD_REQUIRE(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0, "ema_update: bad chunking");
    launch_pdl(ema_update_kernel, dim3(n_chunks), dim3(256), 0, (cudaStream_t)stream, table, chunks, chunk_elems, decay, one_minus_decay);
    return check_launch("ema_update");
}

int dd_nchw_to_nhwc(const float* x, void* y, int dtype, int B, int C, int H, int W, void* stream) {
    dim3 grid((H * W + 127) / 128, B);
    DD_DISPATCH_DTYPE(dtype, T, (launch_pdl(nchw_to_nhwc_kernel<T>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, x, (T*)y, C, H * W)));
    return check_launch("nchw_to_nhwc");
}

int dd_nchw_to_nhwc_pad(const float* x, void* y_bf16, int B, int C, int H, int W, int Cp, void* stream) {
    DD_REQUIRE(B > 0 && C > 0 && Cp >= C && Cp % 8 == 0, "nchw_to_nhwc_pad: C=%d, Cp=%d (Cp must be a multiple of 8 >= C)", C, Cp);
    const int64_t total = (int64_t)B * H * W * (Cp >> 3);
    launch_pdl(nchw_to_nhwc_pad_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y_bf16, C, H * W, Cp, total);
    return check_launch("nchw_to_nhwc_pad");
}

int dd_nhwc_to_nchw(const void* x, int dtype, float* y, int B, int C, int H, int W, void* stream) {
    dim3 grid((H * W + 127) / 128, B);
    DD_DISPATCH_DTYPE(dtype, T, (launch_pdl(nhwc_to_nchw_kernel<T>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (const T*)x, y, C, H * W)));
    return check_launch("nhwc_to_nchw");
}

int dd_s2d_f32(const float* src, float* dst, int B, int h, int w, int C, int to_packed, void* stream) {
    DD_REQUIRE(C % 4 == 0 && B > 0 && h > 0 && w > 0, "s2d_f32: C=%d must be a multiple of 4", C);
    const int64_t total = (int64_t)B * h * w * 4 * (C / 4);
    launch_pdl(s2d_f32_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)src, (float4*)dst, h, w, C / 4,
               to_packed, total);
    return check_launch("s2d_f32");
}

int dd_nhwc_to_chw_pad(const float* x, float* y, int B, int C, int H, int W, int Wp, int hpad, int nshift, float* colsum, void* stream) {
    DD_REQUIRE(C % 32 == 0 && Wp % 32 == 0 && Wp >= W && hpad >= 0, "nhwc_to_chw_pad: C=%d must be a multiple of 32, Wp=%d a multiple of 32 >= W", C, Wp);
    DD_REQUIRE(nshift == 1 || nshift == 3, "nhwc_to_chw_pad: nshift must be 1 or 3 (got %d)", nshift);
    DD_REQUIRE((int64_t)B * (C / 32) <= 65535, "nhwc_to_chw_pad: B*C/32 exceeds the grid limit");
    dim3 grid(Wp / 32, (H + 2 * hpad + CHWP_ROWS - 1) / CHWP_ROWS, B * (C / 32));
    launch_pdl(nhwc_to_chw_pad_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, y, B, C, H, W, Wp, hpad, nshift, colsum);
    return check_launch("nhwc_to_chw_pad");
}

int dd_im2col3x3_nchw(const float* x, void* y, int B, int C, int H, int W, int kpad, void* stream) {
    DD_REQUIRE(kpad % 64 == 0 && kpad >= 9 * C, "im2col3x3: kpad=%d must be a multiple of 64 and >= 9*C", kpad);
    DD_REQUIRE(C > 0 && C <= 64, "im2col3x3: C=%d out of range (1..64)", C);
    dim3 grid((unsigned)(H * ((W + IM2COL_TW - 1) / IM2COL_TW)), B);
    const size_t smem = (size_t)C * 3 * (IM2COL_TW + 2) * sizeof(float);
    if (C == 8 && !getenv("DD_IM2COL_STAGED")) {
        const int64_t total = (int64_t)B * H * W * (kpad >> 3);
        launch_pdl(im2col3x3_c8_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, H, W, kpad, total);
    }
    else if (C == 8) launch_pdl(im2col3x3_kernel<8>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, x, (__nv_bfloat16*)y, C, H, W, kpad);
    else launch_pdl(im2col3x3_kernel<0>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, x, (__nv_bfloat16*)y, C, H, W, kpad);
    return check_launch("im2col3x3");
}

int dd_avgpool2(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream) {
    DD_REQUIRE(H % 2 == 0 && W % 2 == 0, "avgpool2: odd spatial size");
    DD_DISPATCH_DTYPE(dtype, T, {
        DD_REQUIRE(C % Vec<T>::N == 0, "avgpool2: C=%d not vectorisable", C);
        int64_t n = (int64_t)B * (H / 2) * (W / 2) * (C / Vec<T>::N);
        launch_pdl(avgpool2_kernel<T>, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (T*)y, H, W, C, n);
    });
    return check_launch("avgpool2");
}

int dd_upsample_nearest2(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream) {
    DD_DISPATCH_DTYPE(dtype, T, {
        DD_REQUIRE(C % Vec<T>::N == 0, "upsample2: C=%d not vectorisable", C);
        int64_t n = (int64_t)B * (H * 2) * (W * 2) * (C / Vec<T>::N);
        launch_pdl(upsample2_kernel<T>, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const T*)x, (T*)y, H, W, C, n);
    });
    return check_launch("upsample_nearest2");
}

int dd_space_to_depth2(const void* x, void* y, int B, int H, int W, int C, void* stream) {
    DD_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "space_to_depth2: bad shape");
    int64_t n = (int64_t)B * H * W * (C / 8);
    launch_pdl(s2d_kernel, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W,
                                                                  C, n);
    return check_launch("space_to_depth2");
}

}  // extern "C"
