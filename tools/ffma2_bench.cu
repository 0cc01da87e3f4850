#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ float lo(u64 v) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return x + y; }

template <int ILP>
__global__ void __launch_bounds__(256) ffma2_chain(float* out, int iters, float a, float b) {
    u64 acc[ILP];
    const u64 A = pack(a, a), B = pack(b, b);
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = pack(threadIdx.x * 0.001f + i, i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = ffma2(acc[i], A, B);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += lo(acc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// outer product 8 queries x 16 keys as 8 x 8 pairs: acc2[i][j] += (q_i, q_i) * (k_2j, k_2j+1)
template <int NQ, int NK2>
__global__ void __launch_bounds__(256) ffma2_outer(float* out, int iters, float seed) {
    u64 acc[NQ][NK2], qq[NQ], kk[NK2];
    float q[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) q[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int j = 0; j < NK2; ++j) kk[j] = pack(seed * 0.5f + j, seed * 0.25f + j);
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
        for (int j = 0; j < NK2; ++j) acc[i][j] = 0ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) qq[i] = pack(q[i], q[i]);
#pragma unroll
        for (int i = 0; i < NQ; ++i)
#pragma unroll
            for (int j = 0; j < NK2; ++j) acc[i][j] = ffma2(qq[i], kk[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < NQ; ++i) q[i] += 1.0f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
        for (int j = 0; j < NK2; ++j) s += lo(acc[i][j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the same with 8 LDS.128 per 64 FFMA2 (the band kernel's ratio), operands from shared memory
__global__ void __launch_bounds__(256) ffma2_outer_lds(float* out, int iters, float seed) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = seed + i * 0.001f;
    __syncthreads();
    u64 acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0ull;
    const int lane = threadIdx.x & 31;
    const float4* base = reinterpret_cast<const float4*>(sm) + (lane >> 1);
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = base[(u * 37 + it * 8) & 511];
        const float* f = reinterpret_cast<const float*>(v);
        u64 qq[8], kk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) qq[i] = pack(f[i], f[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) kk[j] = pack(f[8 + 2 * j], f[9 + 2 * j]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = ffma2(qq[i], kk[j], acc[i][j]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += lo(acc[i][j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) ffma_outer_lds(float* out, int iters, float seed) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = seed + i * 0.001f;
    __syncthreads();
    float acc[8][16];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    const int lane = threadIdx.x & 31;
    const float4* base = reinterpret_cast<const float4*>(sm) + (lane >> 1);
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = base[(u * 37 + it * 8) & 511];
        const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(f[i], f[8 + j], acc[i][j]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    float* out; CK(cudaMalloc(&out, sizeof(float) * nsm * 8 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int variant = 0; variant < 6; ++variant) {
        const int iters = 20000; float best = 1e30f; double flops = 0;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            if (variant == 0) { ffma2_chain<16><<<nsm * 8, 256>>>(out, iters, 1.0001f, 0.5f); flops = 4.0 * 16 * iters * 256.0 * nsm * 8; }
            if (variant == 1) { ffma2_outer<8, 8><<<nsm, 256>>>(out, iters / 4, 1.5f); flops = 4.0 * 64 * (iters / 4) * 256.0 * nsm; }
            if (variant == 2) { ffma2_outer<8, 8><<<nsm * 2, 256>>>(out, iters / 4, 1.5f); flops = 4.0 * 64 * (iters / 4) * 256.0 * nsm * 2; }
            if (variant == 3) { ffma2_outer_lds<<<nsm, 256>>>(out, iters / 4, 1.5f); flops = 4.0 * 64 * (iters / 4) * 256.0 * nsm; }
            if (variant == 4) { ffma_outer_lds<<<nsm, 256>>>(out, iters / 4, 1.5f); flops = 2.0 * 128 * (iters / 4) * 256.0 * nsm; }
            if (variant == 5) { ffma2_chain<8><<<nsm * 8, 256>>>(out, iters, 1.0001f, 0.5f); flops = 4.0 * 8 * iters * 256.0 * nsm * 8; }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        const char* names[] = {"ffma2 chain ILP16 pairs, 8 CTA/SM", "ffma2 outer 8x(8 pairs), 1 CTA/SM", "ffma2 outer 8x(8 pairs), 2 CTA/SM",
                               "ffma2 outer + 8 LDS.128 per 64 FFMA2, 1 CTA/SM", "ffma  outer + 8 LDS.128 per 128 FFMA, 1 CTA/SM", "ffma2 chain ILP8 pairs, 8 CTA/SM"};
        printf("FP32 %-50s : %.2f TFLOP/s (%.3f ms)\n", names[variant], flops / best * 1e-9, best);
    }
    return 0;
}
