// Micro-benchmark: how fast can ONE SM pull L2-resident data into shared memory?
//   mode 0: TMA 2D boxes (128 rows x 128 B, SWIZZLE_128B), `stages` in flight, one issuing thread
//   mode 1: 128 threads of LDG.128 -> STS (generic path)
//   mode 2: both at the same time (TMA from buffer A, LDG from buffer B)
// Reports bytes/clk per SM for `ctas` CTAs (1 per SM).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_tx(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
constexpr int BOX = 16384;
__constant__ int c_boxrows = 128;
__global__ void __launch_bounds__(160, 1) ingest(const __grid_constant__ CUtensorMap tm, const uint4* __restrict__ g, int iters, int stages,
                                                int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (su32(smem) + 1023u) & ~1023u;
    const uint32_t bars = base + 8 * BOX;
    uint4* lds = reinterpret_cast<uint4*>(smem + (base - su32(smem)) + 9 * BOX);
    if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) mbar_init(bars + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x < 32) {
        const int nissue = (mode >= 10) ? 2 : 1;
        if (mode != 1 && (int)threadIdx.x < nissue) {
            const int rows_total = 8192;
            const int me = threadIdx.x;
            const int my_stages = stages / nissue;
            int issued = 0, done = 0;
            uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int row0 = (blockIdx.x * 1031 + me * 2048) % 4096;
            const int my_iters = iters / nissue;
            const uint32_t boxbytes = c_boxrows * 128;
            while (done < my_iters) {
                while (issued < my_iters && issued - done < my_stages) {
                    const int s = me * my_stages + issued % my_stages;
                    mbar_tx(bars + 8 * s, boxbytes);
                    tma2d(&tm, bars + 8 * s, base + s * BOX, 0, (row0 + issued * c_boxrows) % (rows_total - 256));
                    ++issued;
                }
                const int s = me * my_stages + done % my_stages;
                mbar_wait(bars + 8 * s, ph[s]); ph[s] ^= 1; ++done;
            }
        }
    } else if (mode != 0) {
        const int t = threadIdx.x - 32;                  // 128 loader threads, 16 KB per iteration = 8 x 16 B each
        const uint4* src = g + (size_t)((blockIdx.x * 977) % 2048) * 1024;
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int it = 0; it < iters; ++it) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + ((size_t)(it & 1023) * 1024 + j * 128 + t));
#pragma unroll
            for (int j = 0; j < 8; ++j) { lds[j * 128 + t] = v[j]; acc.x ^= v[j].x; }
        }
        if (acc.x == 0x12345678u) out[1000 + threadIdx.x] = acc.x;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
int main(int argc, char** argv) {
    const int iters = 512;
    void* ptr; size_t bytes = (size_t)64 << 20;        // 64 MB, L2 resident after the warm-up pass
    CK(cudaMalloc(&ptr, bytes)); CK(cudaMemset(ptr, 1, bytes));
    long long* out; CK(cudaMalloc(&out, 4096 * 8)); long long host[4096];
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    const int smem = 1024 + 9 * BOX + 16384 + 256;
    CK(cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    struct Cfg { int boxrows; int swz; int pitch; const char* name; };
    Cfg cfgs[] = {{128, 1, 128, "box 128x128B sw128 pitch128"}, {64, 1, 128, "box 64x128B sw128 pitch128"}, {128, 0, 128, "box 128x128B noswz"},
                  {128, 1, 512, "box 128x128B sw128 pitch512"}, {32, 1, 128, "box 32x128B sw128"}};
    for (auto& c : cfgs) {
        CUtensorMap tm; cuuint64_t dims[2] = {64, 8192}; cuuint64_t str[1] = {(cuuint64_t)c.pitch}; cuuint32_t box[2] = {64, (cuuint32_t)c.boxrows}; cuuint32_t es[2] = {1, 1};
        CUresult r = ((Enc)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        CK(cudaMemcpyToSymbol(c_boxrows, &c.boxrows, sizeof(int)));
        for (int mode : {0, 10})
            for (int stages : {2, 8}) {
                const int ctas = 32;
                for (int rep = 0; rep < 2; ++rep) { ingest<<<ctas, 160, smem>>>(tm, (const uint4*)ptr, iters, stages, mode, out); CK(cudaDeviceSynchronize()); }
                CK(cudaMemcpy(host, out, ctas * 8, cudaMemcpyDeviceToHost));
                double mx = 0; for (int i = 0; i < ctas; ++i) mx = host[i] > mx ? host[i] : mx;
                printf("%-32s issuers %d stages %d: %.1f B/clk per SM, %.0f clk per box\n", c.name, mode == 10 ? 2 : 1, stages,
                       (double)iters * c.boxrows * 128 / mx, mx / iters);
            }
    }
    return 0;
}
