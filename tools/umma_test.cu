// umma_test.cu -- unit test of the tcgen05 building blocks used by the tensor-core correlation kernel:
// tf32 MMA (M=128, N<=256, K=8) from shared memory operands in MN-major, non-swizzled canonical layout
// (core matrix = 8 K-rows x 16 bytes of 4 consecutive MN elements), accumulators in TMEM, read back
// with tcgen05.ld; 3xTF32 split (hi*hi + hi*lo + lo*hi) accuracy against an fp64 reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_test tools/umma_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// shared-memory matrix descriptor, MN-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (sm_100)
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                     // c_format = F32
    d |= 2u << 7;                     // a_format = TF32
    d |= 2u << 10;                    // b_format = TF32
    if (mn_major) d |= (1u << 15) | (1u << 16);  // a_major, b_major = MN
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A: [128][8], B: [N][8] (row-major in global), D: [128][N].  split: 0 = plain tf32, 1 = 3xTF32
__global__ void __launch_bounds__(128) umma_kernel(const float* Ag, const float* Bg, float* Dg, int N, int ksteps, int split, int mn_major) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* Ah = reinterpret_cast<float*>(smem);   // 128*8 floats
    float* Al = Ah + 128 * 8;
    float* Bh = Al + 128 * 8;                     // N*8
    float* Bl = Bh + N * 8;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    uint32_t phase = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
        // stage operands of this k-step in the canonical MN-major layout: (mn, k) -> (mn/4)*32 + k*4 + mn%4 floats
        for (int e = tid; e < 128 * 8; e += 128) {
            const int m = e >> 3, k = e & 7;
            const float v = Ag[(size_t)ks * 128 * 8 + e];
            const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            const int off = mn_major ? (m >> 2) * 32 + k * 4 + (m & 3) : (m >> 3) * 64 + (k >> 2) * 32 + (m & 7) * 4 + (k & 3);
            Ah[off] = split ? hi : v;
            Al[off] = v - hi;
        }
        for (int e = tid; e < N * 8; e += 128) {
            const int n = e >> 3, k = e & 7;
            const float v = Bg[(size_t)ks * N * 8 + e];
            const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            const int off = mn_major ? (n >> 2) * 32 + k * 4 + (n & 3) : (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
            Bh[off] = split ? hi : v;
            Bl[off] = v - hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int n0 = 0; n0 < N; n0 += 256) {
                const int nn = (N - n0) < 256 ? (N - n0) : 256;
                const uint32_t idesc = make_idesc_tf32(128, nn, mn_major);
                const uint32_t sbo = mn_major ? 128 : 256, lbo = 128;
                const uint64_t dAh = make_desc(smem_u32(Ah), sbo, lbo), dAl = make_desc(smem_u32(Al), sbo, lbo);
                const uint64_t dBh = make_desc(smem_u32(Bh + n0 * 8), sbo, lbo), dBl = make_desc(smem_u32(Bl + n0 * 8), sbo, lbo);
                umma_tf32(tmem_base + n0, dAh, dBh, idesc, ks > 0);
                if (split) {
                    umma_tf32(tmem_base + n0, dAh, dBl, idesc, 1);
                    umma_tf32(tmem_base + n0, dAl, dBh, idesc, 1);
                }
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);  // MMAs of this k-step are done reading shared memory
        phase ^= 1;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back: warp w owns TMEM lanes 32w .. 32w+31 (= rows of D)
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
            "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = 32 * warp + lane;
        for (int x = 0; x < 32; ++x)
            if (c0 + x < N) Dg[(size_t)m * N + c0 + x] = __uint_as_float(r[x]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

int main(int argc, char** argv) {
    const int M = 128;
    const int ksteps_arg = argc > 1 ? atoi(argv[1]) : 1;
    const int mn_major = argc > 2 ? atoi(argv[2]) : 1;
    printf("mn_major=%d\n", mn_major);
    for (int N : {256, 96, 384}) {
        for (int split = 0; split <= 1; ++split) {
            const int ksteps = ksteps_arg;
            std::vector<float> A((size_t)ksteps * M * 8), B((size_t)ksteps * N * 8);
            srand(123 + N);
            for (auto& v : A) v = (float)rand() / RAND_MAX - 0.3f;
            for (auto& v : B) v = (float)rand() / RAND_MAX - 0.3f;
            float *dA, *dB, *dD;
            CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
            CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
            const size_t smem = (size_t)(2 * 128 * 8 + 2 * N * 8) * 4;
            CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            umma_kernel<<<1, 128, smem>>>(dA, dB, dD, N, ksteps, split, mn_major);
            CK(cudaDeviceSynchronize());
            std::vector<float> D((size_t)M * N);
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double maxrel = 0, maxabs = 0, scale = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0, mag = 0;
                    for (int ks = 0; ks < ksteps; ++ks)
                        for (int k = 0; k < 8; ++k) {
                            const double a = A[((size_t)ks * M + m) * 8 + k], b = B[((size_t)ks * N + n) * 8 + k];
                            ref += a * b; mag += fabs(a * b);
                        }
                    const double err = fabs(D[(size_t)m * N + n] - ref);
                    if (err > maxabs) maxabs = err;
                    if (err / mag > maxrel) maxrel = err / mag;
                    if (fabs(ref) > scale) scale = fabs(ref);
                }
            {
                int nz = 0; for (float v : D) nz += (v != 0.f);
                double r00 = 0; for (int ks = 0; ks < ksteps; ++ks) for (int k = 0; k < 8; ++k) r00 += (double)A[((size_t)ks * M + 0) * 8 + k] * B[((size_t)ks * N + 0) * 8 + k];
                printf("  nonzero %d / %d; D[0][0..3] = %g %g %g %g (ref00 %g); D[1][0]=%g D[64][5]=%g\n", nz, M * N, D[0], D[1], D[2], D[3], r00, D[N], D[(size_t)64 * N + 5]);
            }
            printf("N=%3d split=%d K=%d: max |err| = %.3e (max|ref| %.2f), max |err|/sum|a*b| = %.3e  %s\n", N, split, ksteps * 8, maxabs,
                   scale, maxrel, maxrel < (split ? 2e-6 : 2e-3) ? "OK" : "FAIL");
            cudaFree(dA); cudaFree(dB); cudaFree(dD);
        }
    }
    return 0;
}
