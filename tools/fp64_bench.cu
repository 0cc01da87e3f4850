// fp64_bench.cu -- how wide is B200's FP64 pipe?  DFMA / DADD / F2F throughput per SM and dependent-issue latency.
// Why: the ROI kernels evaluate the reference's bin-edge expression, which mixes a double literal into float arithmetic
// (roipool_cuda.cu:38-50), i.e. a handful of FP64 instructions per (RoI, bin); and a double-precision
// summed-area table was tried for the batched PSROIPool forward (profiles/r2_psf_experiment.txt).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_bench tools/fp64_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP, int MODE>   // MODE 0: DFMA, 1: DADD, 2: F2F.F64.F32 + DADD, 3: FFMA (reference)
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b, long long* cyc) {
    double acc[ILP];
    float facc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x * 0.001 + i; facc[i] = threadIdx.x * 0.001f + i; }
    const float fa = (float)a, fb = (float)b;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) acc[i] = fma(acc[i], a, b);
            else if (MODE == 1) acc[i] = acc[i] + b;
            else if (MODE == 2) { facc[i] = fmaf(facc[i], fa, fb); acc[i] += (double)facc[i]; }
            else facc[i] = fmaf(facc[i], fa, fb);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i] + facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int MODE>
static void run(const char* name, int nsm, int warpsPerSM, int opsPerIter) {
    double* out; long long* cyc;
    const int blocks = nsm, threads = warpsPerSM * 32;
    CK(cudaMalloc(&out, sizeof(double) * blocks * 256));
    CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
    const int iters = 20000;
    k<ILP, MODE><<<blocks, threads>>>(out, 10, 1.0000001, 1e-9, cyc);
    CK(cudaDeviceSynchronize());
    k<ILP, MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    const double warpInstr = (double)iters * ILP * opsPerIter * warpsPerSM;
    printf("%-34s ILP %2d, %2d warps/SM: %7.2f cycles per warp-instruction per SM  (%6.2f lanes/clk/SM)\n", name, ILP, warpsPerSM,
           avg / warpInstr, 32.0 * warpInstr / avg);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    printf("device: %s, %d SMs\n", prop.name, nsm);
    run<8, 3>("FFMA (reference)", nsm, 8, 1);
    run<8, 0>("DFMA throughput", nsm, 8, 1);
    run<8, 1>("DADD throughput", nsm, 8, 1);
    run<8, 2>("FFMA + F2F.F64.F32 + DADD", nsm, 8, 3);
    run<1, 0>("DFMA dependent chain (latency)", nsm, 1, 1);
    run<1, 1>("DADD dependent chain (latency)", nsm, 1, 1);
    run<1, 3>("FFMA dependent chain (latency)", nsm, 1, 1);
    return 0;
}
