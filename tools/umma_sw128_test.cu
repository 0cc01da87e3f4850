// umma_sw128_test.cu -- unit test of the operand layout used by the tensor-core correlation BACKWARD kernel:
// tf32 MMA (M=128, N=256, K=8) from shared-memory operands in K-major SWIZZLE_128B layout (one row = 32 consecutive
// K elements = 128 bytes, 8-row atoms of 1024 bytes, 16-byte chunks XOR-swizzled by row & 7), K advanced inside the
// swizzle atom by moving the descriptor start address 32 bytes per k-step; 3xTF32 with a round-to-nearest hi part.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_sw128_test tools/umma_sw128_test.cu
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


__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                    // version = 1 (sm_100)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// byte offset of element (row, k) of a [rows][32] fp32 operand in K-major SWIZZLE_128B
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + (k & 3) * 4);
}
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

// A: [chunks][128][32], B: [chunks][N][32] (row-major in global), D: [128][N]
__global__ void __launch_bounds__(128) umma_kernel(const float* Ag, const float* Bg, float* Dg, int N, int chunks, int split) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* Ah = smem;                 // 128 rows x 128 B
    unsigned char* Al = Ah + 128 * 128;
    unsigned char* Bh = Al + 128 * 128;       // N rows x 128 B
    unsigned char* Bl = Bh + N * 128;
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
    for (int ch = 0; ch < chunks; ++ch) {
        // lane = k: a warp writes one 128-byte row per store (conflict-free)
        for (int row = warp; row < 128; row += 4) {
            const float v = Ag[((size_t)ch * 128 + row) * 32 + lane];
            const float hi = split ? tf32_rn(v) : v;
            *reinterpret_cast<float*>(Ah + sw128_off(row, lane)) = hi;
            *reinterpret_cast<float*>(Al + sw128_off(row, lane)) = v - hi;
        }
        for (int row = warp; row < N; row += 4) {
            const float v = Bg[((size_t)ch * N + row) * 32 + lane];
            const float hi = split ? tf32_rn(v) : v;
            *reinterpret_cast<float*>(Bh + sw128_off(row, lane)) = hi;
            *reinterpret_cast<float*>(Bl + sw128_off(row, lane)) = v - hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = make_idesc_tf32(128, N, 0);
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ko = ks * 32;  // 8 tf32 = 32 bytes inside the 128-byte swizzle row
                const uint64_t dAh = make_desc_sw128(smem_u32(Ah) + ko), dAl = make_desc_sw128(smem_u32(Al) + ko);
                const uint64_t dBh = make_desc_sw128(smem_u32(Bh) + ko), dBl = make_desc_sw128(smem_u32(Bl) + ko);
                umma_tf32(tmem_base, dAh, dBh, idesc, (ch > 0 || ks > 0));
                if (split) {
                    umma_tf32(tmem_base, dAh, dBl, idesc, 1);
                    umma_tf32(tmem_base, dAl, dBh, idesc, 1);
                }
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
    const int chunks = argc > 1 ? atoi(argv[1]) : 3;
    int fails = 0;
    for (int N : {256, 128}) {
        for (int split = 0; split <= 1; ++split) {
            std::vector<float> A((size_t)chunks * M * 32), B((size_t)chunks * N * 32);
            srand(123 + N);
            for (auto& v : A) v = (float)rand() / RAND_MAX - 0.3f;
            for (auto& v : B) v = (float)rand() / RAND_MAX - 0.3f;
            float *dA, *dB, *dD;
            CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
            CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
            const size_t smem = (size_t)(2 * 128 + 2 * N) * 128 + 1024;
            CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            umma_kernel<<<1, 128, smem>>>(dA, dB, dD, N, chunks, split);
            CK(cudaDeviceSynchronize());
            std::vector<float> D((size_t)M * N);
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double maxrel = 0, maxabs = 0, scale = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0, mag = 0;
                    for (int ch = 0; ch < chunks; ++ch)
                        for (int k = 0; k < 32; ++k) {
                            const double a = A[((size_t)ch * M + m) * 32 + k], b = B[((size_t)ch * N + n) * 32 + k];
                            ref += a * b; mag += fabs(a * b);
                        }
                    const double err = fabs(D[(size_t)m * N + n] - ref);
                    if (err > maxabs) maxabs = err;
                    if (err / mag > maxrel) maxrel = err / mag;
                    if (fabs(ref) > scale) scale = fabs(ref);
                }
            const bool ok = maxrel < (split ? 2e-6 : 2e-3);
            fails += !ok;
            printf("sw128 N=%3d split=%d K=%d: max |err| = %.3e (max|ref| %.2f), max |err|/sum|a*b| = %.3e  %s\n", N, split, chunks * 32,
                   maxabs, scale, maxrel, ok ? "OK" : "FAIL");
            cudaFree(dA); cudaFree(dB); cudaFree(dD);
        }
    }
    return fails;
}
