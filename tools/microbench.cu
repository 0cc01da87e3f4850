// microbench.cu -- measured denominators / design inputs for the SIMT correlation kernels.
//   1. FP32 FFMA throughput (the roofline the SIMT band kernel is held against; BASELINE.md section 2
//      lists it as "to be confirmed with a measured FMA micro-benchmark").
//   2. LDS.128 cost per warp instruction for the address patterns the band kernel produces
//      (broadcast across quarter-warps, strided, shared rows).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// ---- 1. FFMA peak -----------------------------------------------------------------------------
template <int ILP>
__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// two distinct multiplicands per FFMA (acc += x*y with x,y in registers): the band kernel's shape
template <int NQ, int NK>
__global__ void __launch_bounds__(256) ffma_outer_kernel(float* out, int iters, float seed) {
    float acc[NQ][NK];
    float q[NQ], k[NK];
#pragma unroll
    for (int i = 0; i < NQ; ++i) q[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int j = 0; j < NK; ++j) k[j] = seed * 0.5f + j;
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
        for (int j = 0; j < NK; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NQ; ++i)
#pragma unroll
            for (int j = 0; j < NK; ++j) acc[i][j] = fmaf(q[i], k[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < NQ; ++i) q[i] += 1.0f;  // keep the loop from being hoisted
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
        for (int j = 0; j < NK; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 2. LDS.128 patterns ------------------------------------------------------------------------
__device__ __forceinline__ int chunk_of(int pattern, int lane) {
    switch (pattern) {
        case 0: return lane;                                   // 32 distinct contiguous chunks (512 B)
        case 1: return 0;                                      // full broadcast
        case 2: return lane & 7;                               // 8 distinct, repeated in every quarter-warp
        case 3: return (lane & 7) * 2;                         // 8 distinct, 32-B stride (2-way bank overlap)
        case 4: return (lane >> 3) * 20 + (lane & 7) * 2;      // 4 rows (pitch 80 floats), 32-B stride: 32 distinct
        case 5: return (lane >> 4) * 20 + (lane & 7) * 2;      // 2 rows shared by quarter-warp pairs: 16 distinct
        case 6: return (lane >> 3) * 21 + (lane & 7) * 2;      // pitch 84 floats
        case 7: return (lane >> 4) * 21 + (lane & 7) * 2;
        case 8: return (lane >> 2);                            // 8 distinct, each read by 4 adjacent lanes
        case 9: return (lane & 3) * 21 + (lane >> 2) * 2;      // bwd pattern: 4 channel groups x 8 column blocks, pitch 84
        case 10: return lane >> 1;                             // 4 distinct per quarter-warp (16 total), contiguous
        case 11: return (lane & 7) >> 1;                       // the same 4 chunks in every quarter-warp
        case 12: return ((lane >> 2) & 1) * 8 + (lane >> 3);   // 2 distinct per quarter, SAME bank group (2-way)
        case 13: return ((lane >> 1) & 3) * 9 + (lane >> 3) * 36;  // 4 distinct per quarter, rows 9 chunks apart
        case 14: return (lane & 7) * 9 + (lane >> 3) * 2;      // 8 distinct per quarter, pitch-36 rows, conflict-free
        case 15: return (lane & 7) * 65 + (lane >> 3) * 4;     // 8 distinct per quarter, pitch-260 blocks
        default: return lane;
    }
}

__global__ void __launch_bounds__(256) lds128_kernel(float* out, int iters, int pattern, long long* cycles) {
    __shared__ __align__(16) float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int base = chunk_of(pattern, lane);
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float4 v = sm[(base + u * 96 + (it & 1)) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clockKHz = 0;
    CK(cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, 0));
    printf("device: %s, %d SMs, max clock %.0f MHz\n", prop.name, prop.multiProcessorCount, clockKHz / 1000.0);
    const int nsm = prop.multiProcessorCount;
    float* out;
    CK(cudaMalloc(&out, sizeof(float) * nsm * 8 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    // FFMA peak: 256 threads x (4 or 8) blocks per SM
    for (int variant = 0; variant < 3; ++variant) {
        const int iters = 20000;
        float best = 1e30f;
        const int blocks = nsm * 8;
        double flops = 0;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            if (variant == 0) { ffma_kernel<16><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f); flops = 2.0 * 16 * iters * 256.0 * blocks; }
            if (variant == 1) { ffma_outer_kernel<8, 8><<<blocks, 256>>>(out, iters / 4, 1.5f); flops = 2.0 * 64 * (iters / 4) * 256.0 * blocks; }
            if (variant == 2) { ffma_outer_kernel<8, 16><<<nsm, 256>>>(out, iters / 4, 1.5f); flops = 2.0 * 128 * (iters / 4) * 256.0 * nsm; }
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const char* names[] = {"ffma chain ILP16, 8 CTA/SM", "ffma outer 8x8, 8 CTA/SM", "ffma outer 8x16, 1 CTA/SM (8 warps)"};
        printf("FP32 %-40s : %.2f TFLOP/s (%.3f ms)\n", names[variant], flops / best * 1e-9, best);
    }

    // LDS.128 patterns: 1 CTA per SM, 8 warps
    long long* cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
    for (int pattern = 0; pattern < 16; ++pattern) {
        const int iters = 4000;
        lds128_kernel<<<nsm, 256>>>(out, iters, pattern, cyc);
        CK(cudaDeviceSynchronize());
        lds128_kernel<<<nsm, 256>>>(out, iters, pattern, cyc);
        CK(cudaDeviceSynchronize());
        long long h[1024];
        CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < nsm; ++i) avg += h[i];
        avg /= nsm;
        const double insts = 8.0 * iters * 16;  // warp-level LDS.128 per SM
        printf("LDS.128 pattern %d: %.2f cycles per warp-instruction per SM (8 warps)\n", pattern, avg / insts);
    }
    return 0;
}
