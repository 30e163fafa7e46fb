// Micro-benchmark for next round's persistent convolution kernel: steady-state issue rate of tcgen05.mma (kind::f16, bf16
// operands in shared memory, K-major, SWIZZLE_128B, fp32 accumulators in TMEM) for
//   cta_group::1  M = 128, N = 64 / 128 / 256          (one CTA, what conv_tc_kernel / conv_tc_halo_kernel issue today)
//   cta_group::2  M = 256, N = 128 / 256               (a cluster of two CTAs, each holding its 128 A rows and half of B)
// Operands stay resident (no TMA refill), the issuing thread alternates between two accumulators, so the number measures
// the tensor pipe + the shared-memory operand reads only.  Measured before (profiles/README.md): cta_group::1 costs about
// (A bytes + B bytes) / 64 B per clock, i.e. M128 x N128 x K16 = 128 clk = 50 % of the pipe.  Expectation for the pair:
// each CTA reads 4 KB of A + N/2 rows of B per K16 step, M256 x N256 x K16 = 8 KB per CTA = 128 clk = the pipe time itself.
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pair_mma scripts/micro/pair_mma.cu
// Run:    ./pair_mma [iters]        (one cluster / CTA per SM, all SMs busy, per-CTA clock64 around the issue loop)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok = 0;
    for (uint32_t tries = 0; !ok; ++tries) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
        if (tries > (1u << 26)) __trap();       // bounded: a protocol mistake must not hang the GPU box
    }
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (the descriptor conv_tc_common.cuh builds)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int CTAS>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CTAS == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");      // the form conv_tc_common.cuh uses (verified on the GPU)
}
template <int CTAS>
__device__ __forceinline__ void commit(uint32_t bar) {
    if (CTAS == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");   // both CTAs' barriers, as in the library
}

// smem: A tile 128 rows x 64 K (16 KB) + B tile (N / CTAS) rows x 64 K, both zero-filled; TMEM: 512 columns (two accumulators)
template <int CTAS, int N>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (su32(smem) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 16384, bar = sB + (N / CTAS) * 128, tptr = bar + 8;
    volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem + (tptr - su32(smem)));
    for (uint32_t i = threadIdx.x; i < (16384 + (N / CTAS) * 128) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem + (base - su32(smem)))[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    const bool leader = CTAS == 1 || ctarank() == 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if (CTAS == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CTAS == 1) __syncthreads(); else cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tptr_gen;
    long long dt = 0;
    if (warp == 1 && leader) {
        if ((threadIdx.x & 31) == 0) {
            constexpr int M = 128 * CTAS;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
            const uint64_t ad = umma_desc(sA), bd = umma_desc(sB);
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tmem + (uint32_t)((i & 1) * 256);        // two accumulators of <= 256 columns
#pragma unroll
                for (int k = 0; k < 4; ++k) umma<CTAS>(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i > 1 || k) ? 1u : 0u);
            }
            commit<CTAS>(bar);
            mbar_wait(bar, 0);
            dt = clock64() - t0;
            out[blockIdx.x / CTAS] = dt;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CTAS == 1) __syncthreads(); else cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

template <int CTAS, int N>
static void run(int iters, int sms, long long* d_out) {
    const size_t smem = 16384 + (N / CTAS) * 128 + 1024 + 64;
    CK(cudaFuncSetAttribute(mma_rate<CTAS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    const int ctas = CTAS == 1 ? sms : (sms / 2) * 2;
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {          // second run is the warm one
        CK(cudaLaunchKernelEx(&cfg, mma_rate<CTAS, N>, iters, d_out));
        CK(cudaDeviceSynchronize());
    }
    std::vector<long long> h(ctas / CTAS);
    CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (long long v : h) mean += (double)v;
    mean /= (double)h.size();
    const double clk = mean / ((double)iters * 4.0);
    const double flop = 2.0 * (128.0 * CTAS) * N * 16.0;
    // dense bf16 pipe: 8192 FLOP per clock and SM (2.25 PFLOP/s nominal over 148 SMs at ~1.86 GHz)
    printf("cta_group::%d  M=%3d N=%3d K=16: %7.1f clk per MMA, %6.0f FLOP/clk/SM = %5.1f %% of 8192\n", CTAS, 128 * CTAS, N, clk,
           flop / clk / CTAS, 100.0 * flop / clk / CTAS / 8192.0);
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    long long* d_out;
    CK(cudaMalloc(&d_out, sizeof(long long) * sms));
    printf("%s, %d SMs, %d x 4 MMAs per CTA\n", prop.name, sms, iters);
    run<1, 64>(iters, sms, d_out);
    run<1, 128>(iters, sms, d_out);
    run<1, 256>(iters, sms, d_out);
    run<2, 128>(iters, sms, d_out);
    run<2, 256>(iters, sms, d_out);
    return 0;
}
