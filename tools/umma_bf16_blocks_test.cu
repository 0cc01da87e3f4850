// umma_bf16_blocks_test.cu -- building block for a tensor-core ROIPool backward:
// tcgen05.mma kind::f16 with BF16 operands in the K-major, NON-swizzled canonical layout, where the two 8-element
// K halves of one MMA (K = 16) are two independent 16-byte-per-row blocks that may lie anywhere in shared memory
// (the descriptor's leading byte offset is the distance between them).  That is what lets one MMA contract two
// arbitrary (RoI, bin row) blocks:  D[pixel][channel] += sum_j A_blk[pixel][j] * B_blk[channel][j].
//   A block: [128 rows][8 bf16] = 2 KB, 0/1 indicators (exact in bf16);  B block: [N rows][8 bf16], an FP32 value split
//   into three BF16 pieces (hi + mid + lo, 24 mantissa bits) => three MMAs per block pair, FP32 accumulation in TMEM.
// Checks: (1) block pairs at irregular distances give the same result as the dense reference, (2) accuracy of the
// 3-piece split against an FP64 reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_bf16_blocks_test tools/umma_bf16_blocks_test.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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
// K-major, SWIZZLE_NONE: rows 16 bytes apart inside an 8-row core matrix, 8-row groups `sbo` bytes apart, the second
// K half `lbo` bytes after the first (cute/arch/mma_sm100_desc.hpp; canonical layout ((8,n),2):((1,SBO),LBO) in uint128)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (sm_100)
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;   // c_format = F32
    d |= 1u << 7;   // a_format = BF16
    d |= 1u << 10;  // b_format = BF16
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Ag: [nb][128][8] floats (0/1), Bg: [nb][N][8] floats, slot[nb]: position of block b in the shared-memory block arrays
// (irregular on purpose).  D: [128][N].
__global__ void __launch_bounds__(128) blocks_kernel(const float* Ag, const float* Bg, const int* slot, float* Dg, int N, int nb,
                                                     int nslots, int nsplit) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int aBytes = 128 * 16, bBytes = N * 16;
    unsigned char* Ablk = smem;                                // [nslots][128][16 B]
    unsigned char* Bblk = smem + (size_t)nslots * aBytes;      // [3][nslots][N][16 B]
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
    // zero everything (unused slots, the zero block), then stage the blocks at their slots
    for (int i = tid; i < (nslots * aBytes + 3 * nslots * bBytes) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int b = 0; b < nb; ++b) {
        const int s = slot[b];
        for (int r = tid; r < 128; r += 128) {
            __nv_bfloat16 v[8];
            for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(Ag[((size_t)b * 128 + r) * 8 + j]);
            *reinterpret_cast<uint4*>(Ablk + (size_t)s * aBytes + r * 16) = *reinterpret_cast<uint4*>(v);
        }
        for (int r = tid; r < N; r += 128) {
            __nv_bfloat16 p0[8], p1[8], p2[8];
            for (int j = 0; j < 8; ++j) {
                const float x = Bg[((size_t)b * N + r) * 8 + j];
                p0[j] = __float2bfloat16(x);
                const float r1 = x - __bfloat162float(p0[j]);
                p1[j] = __float2bfloat16(r1);
                const float r2 = r1 - __bfloat162float(p1[j]);
                p2[j] = __float2bfloat16(r2);
            }
            *reinterpret_cast<uint4*>(Bblk + ((size_t)0 * nslots + s) * bBytes + r * 16) = *reinterpret_cast<uint4*>(p0);
            *reinterpret_cast<uint4*>(Bblk + ((size_t)1 * nslots + s) * bBytes + r * 16) = *reinterpret_cast<uint4*>(p1);
            *reinterpret_cast<uint4*>(Bblk + ((size_t)2 * nslots + s) * bBytes + r * 16) = *reinterpret_cast<uint4*>(p2);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const int zeroSlot = nslots - 1;  // never used by a block: all zeros
        bool first = true;
        for (int b = 0; b < nb; b += 2) {
            // blocks are paired in ascending slot order so that the second K half lies at a higher address
            int s0 = slot[b], s1 = (b + 1 < nb) ? slot[b + 1] : zeroSlot;
            if (s1 < s0) { const int t = s0; s0 = s1; s1 = t; }
            const uint64_t da = make_desc(smem_u32(Ablk + (size_t)s0 * aBytes), 128, (uint32_t)(s1 - s0) * aBytes);
            for (int sp = 0; sp < nsplit; ++sp) {
                const uint64_t db = make_desc(smem_u32(Bblk + ((size_t)sp * nslots + s0) * bBytes), 128, (uint32_t)(s1 - s0) * bBytes);
                umma_f16(tmem_base, da, db, idesc, first ? 0u : 1u);
                first = false;
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
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
    for (int N : {192, 256, 64}) {
        for (int nb : {2, 7, 12}) {
            for (int nsplit : {1, 3}) {
                const int nslots = 16;
                std::vector<float> A((size_t)nb * M * 8), B((size_t)nb * N * 8);
                std::vector<int> slot(nb);
                srand(7 + N + nb);
                for (auto& v : A) v = (rand() % 3 == 0) ? 1.f : 0.f;
                for (auto& v : B) v = ((float)rand() / RAND_MAX - 0.4f) * 3.f;
                for (int b = 0; b < nb; ++b) slot[b] = (b * 4 + 3) % (nslots - 1);  // irregular, distinct (4 is prime to 15), never the zero slot
                float *dA, *dB, *dD;
                int* dS;
                CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
                CK(cudaMalloc(&dS, nb * sizeof(int)));
                CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(dS, slot.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
                const size_t smem = (size_t)nslots * 128 * 16 + (size_t)3 * nslots * N * 16;
                CK(cudaFuncSetAttribute(blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                blocks_kernel<<<1, 128, smem>>>(dA, dB, dS, dD, N, nb, nslots, nsplit);
                CK(cudaDeviceSynchronize());
                std::vector<float> D((size_t)M * N);
                CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
                double maxrel = 0, maxabs = 0;
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < N; ++n) {
                        double ref = 0, mag = 0;
                        for (int b = 0; b < nb; ++b)
                            for (int j = 0; j < 8; ++j) {
                                const double a = A[((size_t)b * M + m) * 8 + j], x = B[((size_t)b * N + n) * 8 + j];
                                ref += a * x; mag += fabs(a * x);
                            }
                        const double err = fabs(D[(size_t)m * N + n] - ref);
                        if (err > maxabs) maxabs = err;
                        if (mag > 0 && err / mag > maxrel) maxrel = err / mag;
                        if (mag == 0 && err > 0) maxrel = 1;
                    }
                const double tol = nsplit == 3 ? 1e-6 : 6e-3;  // 3 pieces: FP32 accumulation rounding; 1 piece: bf16 operand rounding
                const bool ok = maxrel < tol;
                fails += !ok;
                printf("N=%3d blocks=%2d pieces=%d: max |err| = %.3e, max |err|/sum|a*b| = %.3e  %s\n", N, nb, nsplit, maxabs, maxrel, ok ? "OK" : "FAIL");
                cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
            }
        }
    }
    printf(fails ? "FAILED %d cases\n" : "all cases OK\n", fails);
    return fails ? 1 : 0;
}
