// tma_bench.cu -- can TMA stage the key patch?  NCHW planes with W=63 have row strides that are not
// multiples of 16 bytes, so the only legal tensor map is 1-D over the whole array (any element offset
// is a legal box start).  One cp.async.bulk.tensor.1d per patch row (32 floats = 128 B).  This measures
// (a) correctness with unaligned starts and (b) sustained small-box throughput per SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const CUtensorMap* map, int x, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(x), "r"(smem_u32(bar))
                 : "memory");
}

// each CTA: `iters` rounds; per round one thread issues `nrows` 128-byte boxes, all threads wait, repeat.
__global__ void __launch_bounds__(256) tma_kernel(const __grid_constant__ CUtensorMap map, float* out, int iters, int nrows,
                                                  int rowStride, long long* cycles, int check, int pitch, int alignX) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ __align__(8) uint64_t bar[2];
    float* buf = sm;  // 2 stages x nrows x 36 floats
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int base = alignX ? blockIdx.x * 7920 : blockIdx.x * 7919 + 3;  // element offset (unaligned unless alignX)
    float acc = 0.f;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], nrows * 128);
        for (int r = 0; r < nrows; ++r) tma_load_1d(buf + r * pitch, &map, base + r * rowStride, &bar[0]);
    }
    for (int it = 0; it < iters; ++it) {
        const int s = it & 1;
        if (threadIdx.x == 0 && it + 1 < iters) {
            mbar_expect_tx(&bar[s ^ 1], nrows * 128);
            for (int r = 0; r < nrows; ++r)
                tma_load_1d(buf + ((s ^ 1) * nrows + r) * pitch, &map, base + (it + 1) * 64 + r * rowStride, &bar[s ^ 1]);
        }
        mbar_wait(&bar[s], (it >> 1) & 1);
        for (int e = threadIdx.x; e < nrows * 32; e += blockDim.x) {
            const int r = e >> 5, x = e & 31;
            const float v = buf[(s * nrows + r) * pitch + x];
            acc += v;
            if (check) {
                const float want = (float)((base + it * 64 + r * rowStride + x) % 9973);
                if (v != want) acc = -1e30f;
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
    const int pitch = argc > 1 ? atoi(argv[1]) : 32;
    const int alignX = argc > 2 ? atoi(argv[2]) : 0;
    const int rowStride = alignX ? 64 : 63;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    const size_t N = (size_t)64 << 20;  // 64 Mi floats = 256 MB
    float* d;
    CK(cudaMalloc(&d, N * sizeof(float)));
    {
        std::vector<float> h(N);
        for (size_t i = 0; i < N; ++i) h[i] = (float)(i % 9973);
        CK(cudaMemcpy(d, h.data(), N * sizeof(float), cudaMemcpyHostToDevice));
    }
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap map;
    cuuint64_t dims[1] = {N};
    cuuint64_t strides[1] = {0};
    cuuint32_t box[1] = {32};
    cuuint32_t estr[1] = {1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled 1-D, box 32 floats: result %d (pitch %d floats, alignX %d)\n", (int)r, pitch, alignX);
    if (r != CUDA_SUCCESS) return 1;
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * nsm * 256));
    CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
    for (int nrows : {31, 62, 124, 248}) {
        const size_t smem = (size_t)2 * nrows * pitch * sizeof(float);
        CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int check = 1; check >= 0; --check) {
            const int iters = 2000;
            tma_kernel<<<nsm, 256, smem>>>(map, out, iters, nrows, rowStride, cyc, check, pitch, alignX);
            CK(cudaDeviceSynchronize());
            std::vector<long long> h(nsm);
            std::vector<float> ho(nsm * 256);
            CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ho.data(), out, sizeof(float) * nsm * 256, cudaMemcpyDeviceToHost));
            double avg = 0;
            for (int i = 0; i < nsm; ++i) avg += h[i];
            avg /= nsm;
            bool ok = true;
            for (float v : ho) ok = ok && v > -1e29f;
            printf("TMA 128-B boxes, %3d per round, check=%d: %.1f cycles per box per SM, %.2f B/cycle/SM, data %s\n", nrows, check,
                   avg / ((double)iters * nrows), 128.0 * iters * nrows / avg, ok ? "OK" : "MISMATCH");
        }
    }
    return 0;
}
