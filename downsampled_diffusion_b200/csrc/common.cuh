// Shared device helpers for libddb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/ddb200.h"

namespace dd {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define DD_REQUIRE(cond, ...)                       \
    do {                                            \
        if (!(cond)) {                              \
            dd::set_error(__VA_ARGS__);             \
            return DD_ERR_ARG;                      \
        }                                           \
    } while (0)

// Mish(x) = x * tanh(softplus(x))  (blocks.py:80, convblocks.py:110).
// tanh(log(1+e^x)) = n/(n+2), n = e^x (e^x + 2): one exp, no cancellation for x << 0.
__device__ __forceinline__ float mish_f(float x) {
    if (x > 20.f) return x;
    float e = expf(x);
    float n = e * (e + 2.f);
    return x * (n / (n + 2.f));
}

// fast variant for the bf16 path: two MUFU operations (ex2, rcp) and five FP32 instructions per value.  The flush-to-zero forms
// are spelled out: without -ftz `__expf` / `__fdividef` wrap each MUFU in range-scaling code (2 FSETP + ~6 FMUL/FADD per value,
// measured in SASS), and these kernels are bound by instruction issue.  n / (n + 2) = 1 - 2 / (n + 2); absolute error ~1e-6 |x|.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mish_fast(float x) {
    // no clamp: for large x, e (and n) overflow to +inf, rcp(inf) = +0 and the result is x; for very negative x, e flushes to 0
    const float e = ex2_ftz(x * 1.4426950408889634f);
    const float n = e * (e + 2.f);
    return x * fmaf(-2.f, rcp_ftz(n + 2.f), 1.f);
}
template <typename T> __device__ __forceinline__ float mish_t(float x);
template <> __device__ __forceinline__ float mish_t<float>(float x) { return mish_f(x); }
template <> __device__ __forceinline__ float mish_t<__nv_bfloat16>(float x) { return mish_fast(x); }

// d/dx mish(x) = tanh(sp) + x * sigmoid(x) * (1 - tanh(sp)^2)
__device__ __forceinline__ float mish_grad_f(float x) {
    if (x > 20.f) return 1.f;
    float e = expf(x);
    float n = e * (e + 2.f);
    float th = n / (n + 2.f);
    float sg = e / (1.f + e);
    return th + x * sg * (1.f - th * th);
}

// fast form for fused epilogues (ex2 / rcp approximations, ~1e-6 relative)
__device__ __forceinline__ float mish_grad_fast(float x) {
    const float e = __expf(fminf(x, 20.f));
    const float n = e * (e + 2.f);
    const float th = __fdividef(n, n + 2.f);
    const float sg = __fdividef(e, 1.f + e);
    return th + x * sg * (1.f - th * th);
}

template <typename T> struct Vec;   // 16-byte vector of activations
template <> struct Vec<float> {
    static constexpr int N = 4;
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __device__ __forceinline__ void from_raw(const uint4& t) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
    __device__ __forceinline__ uint4 to_raw() const {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    float v[8];
    __device__ __forceinline__ void load(const __nv_bfloat16* p) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
    __device__ __forceinline__ void from_raw(const uint4& t) {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    __device__ __forceinline__ uint4 to_raw() const {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        return t;
    }
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the library starts with pdl_sync(): wait until the preceding kernel in the stream has
// completed and flushed its writes, then allow the NEXT kernel to be scheduled.  Launched through
// launch_pdl(), a kernel's launch latency and prologue (barrier init, TMEM allocation, descriptor
// prefetch) overlap the tail of its predecessor; inside a captured graph the edges become programmatic
// dependencies.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_sync() {
    // (launch_dependents BEFORE the wait, so that the launch after next may become resident as well, is slower:
    //  0.585 against 0.543 ms per step at 8 samples per GPU -- profiles/README.md, round 2 pass ah)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
    static const bool on = getenv("DD_NO_PDL") == nullptr;
    return on;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace dd

// dispatch on activation dtype
#define DD_DISPATCH_DTYPE(dtype, T, ...)                                   \
    do {                                                                   \
        if ((dtype) == DD_F32) { using T = float; __VA_ARGS__; }           \
        else if ((dtype) == DD_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else { dd::set_error("bad dtype %d", (int)(dtype)); return DD_ERR_ARG; } \
    } while (0)
