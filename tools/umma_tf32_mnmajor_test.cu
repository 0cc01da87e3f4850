// umma_tf32_mnmajor_test.cu -- tcgen05.mma kind::tf32 with BOTH operands in the MN-major SWIZZLE_128B canonical layout
// (round 2).  cute/arch/mma_sm100_desc.hpp:427 says MN-major is valid for TF32; round 1's tools/umma_test.cu got zeros
// (a descriptor error).  A correlation FORWARD on tensor cores contracts over channels (K) while NCHW memory is contiguous
// over positions (M / N): MN-major operands are the layout the maps already have.
//   operand tile of one MMA (K = 8 tf32): [MN atom (32 positions)][2 K atoms][4 k-rows of 128 bytes], the 32-byte chunks
//   of a row XOR-swizzled with the row index (SWIZZLE_128B_BASE32B, layout type 1); descriptor: LBO = bytes between MN
//   atoms, SBO = bytes between K atoms (cute/atom/mma_traits_sm100.hpp, Layout_MN_SW128_32B_Atom, make_umma_desc<Major::MN>).
//   With the plain SWIZZLE_128B layout type (2) the instruction produces zeros (round 1, and re-checked in round 2).
//   A: [K][M = 128], B: [K][N]; FP32 values split hi = tf32_rn(v), lo = v - hi; products hi*hi (+ hi*lo + lo*hi).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_tf32_mnmajor_test tools/umma_tf32_mnmajor_test.cu
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
// SWIZZLE_128B_BASE32B descriptor (layout type 1: the MN-major layout of 32-bit operands,
// cute/atom/mma_traits_sm100.hpp Layout_MN_SW128_32B_Atom), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_bf16_mn(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                  // c_format = F32
    d |= 2u << 7;                  // a_format = TF32
    d |= 2u << 10;                 // b_format = TF32
    d |= (1u << 15) | (1u << 16);  // a_major = b_major = MN
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// byte offset of element (k, mn) inside an operand tile [K = 8][MN]: atoms of 32 positions x 4 k-rows of 128 bytes whose
// 32-byte chunks are XOR-swizzled with the row index (Swizzle<2,5,2> on byte addresses); the two K atoms of an MN atom are
// adjacent (SBO = 512), MN atoms are 1024 bytes apart (LBO)
__device__ __forceinline__ uint32_t tile_off(int k, int mn, int nAtoms) {
    const int ma = mn >> 5, mr = mn & 31, ka = k >> 2, kr = k & 3;
    const int chunk = (mr >> 3) ^ kr;
    (void)nAtoms;
    return (uint32_t)((ma * 2 + ka) * 512 + kr * 128 + chunk * 32 + (mr & 7) * 4);
}
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

// Ag: [ksteps][8][128], Bg: [ksteps][8][N] (K-rows of contiguous positions, like NCHW), D: [128][N]
__global__ void __launch_bounds__(128) mn_kernel(const float* Ag, const float* Bg, float* Dg, int N, int ksteps, int nprod) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int aBytes = 8 * 128 * 4, bBytes = 8 * N * 4;
    unsigned char* A3 = smem;               // hi, lo
    unsigned char* B3 = smem + 2 * aBytes;  // hi, lo (aBytes = 4 KB keeps 1 KB alignment)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
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
        for (int e = tid; e < 8 * 128; e += 128) {
            const int k = e >> 7, m = e & 127;
            const float x = Ag[(size_t)ks * 8 * 128 + e];
            const float h = tf32_rn(x);
            const uint32_t off = tile_off(k, m, 4);
            *reinterpret_cast<float*>(A3 + off) = h;
            *reinterpret_cast<float*>(A3 + aBytes + off) = x - h;
        }
        for (int e = tid; e < 8 * N; e += 128) {
            const int k = e / N, n = e - k * N;
            const float x = Bg[(size_t)ks * 8 * N + e];
            const float h = tf32_rn(x);
            const uint32_t off = tile_off(k, n, N / 32);
            *reinterpret_cast<float*>(B3 + off) = h;
            *reinterpret_cast<float*>(B3 + bBytes + off) = x - h;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = make_idesc_bf16_mn(128, N);
            // (a piece, b piece) products in decreasing magnitude
            const int pa[3] = {0, 0, 1}, pb[3] = {0, 1, 0};
            for (int t = 0; t < nprod; ++t) {
                const uint64_t da = make_desc(smem_u32(A3 + pa[t] * aBytes), 512, 1024);
                const uint64_t db = make_desc(smem_u32(B3 + pb[t] * bBytes), 512, 1024);
                umma_f16(tmem_base, da, db, idesc, (ks > 0 || t > 0) ? 1u : 0u);
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
}

int main() {
    const int M = 128;
    int fails = 0;
    for (int N : {256, 128, 64}) {
        for (int ksteps : {1, 8}) {
            for (int nprod : {1, 3}) {
                std::vector<float> A((size_t)ksteps * 8 * M), B((size_t)ksteps * 8 * N);
                srand(11 + N + ksteps);
                for (auto& v : A) v = (float)rand() / RAND_MAX - 0.3f;
                for (auto& v : B) v = (float)rand() / RAND_MAX - 0.3f;
                float *dA, *dB, *dD;
                CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
                CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemset(dD, 0, (size_t)M * N * 4));
                const size_t smem = 2 * (size_t)8 * 128 * 4 + 2 * (size_t)8 * N * 4 + 1024;
                CK(cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                mn_kernel<<<1, 128, smem>>>(dA, dB, dD, N, ksteps, nprod);
                CK(cudaDeviceSynchronize());
                std::vector<float> D((size_t)M * N);
                CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
                double maxrel = 0, maxabs = 0;
                int nz = 0;
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < N; ++n) {
                        double ref = 0, mag = 0;
                        for (int ks = 0; ks < ksteps; ++ks)
                            for (int k = 0; k < 8; ++k) {
                                const double a = A[((size_t)ks * 8 + k) * M + m], b = B[((size_t)ks * 8 + k) * N + n];
                                ref += a * b; mag += fabs(a * b);
                            }
                        const double err = fabs(D[(size_t)m * N + n] - ref);
                        nz += D[(size_t)m * N + n] != 0.f;
                        if (err > maxabs) maxabs = err;
                        if (err / mag > maxrel) maxrel = err / mag;
                    }
                const double tol = nprod == 3 ? 2e-6 : 2e-3;
                const bool ok = maxrel < tol;
                fails += !ok;
                printf("MN-major sw128 N=%3d K=%3d products=%d: nonzero %d/%d, max |err| = %.3e, max |err|/sum|a*b| = %.3e  %s\n", N,
                       ksteps * 8, nprod, nz, M * N, maxabs, maxrel, ok ? "OK" : "FAIL");
                cudaFree(dA); cudaFree(dB); cudaFree(dD);
            }
        }
    }
    printf(fails ? "FAILED %d cases\n" : "all cases OK\n", fails);
    return fails ? 1 : 0;
}
