// tma_bench2.cu -- can TMA stage the correlation operands straight from the UNPADDED NCHW maps (38 x 63, rows 252 B apart)?
//
// Round 1 (tools/tma_bench.cu) found that a 1-D tensor map is the only "natural" legal map over such planes and that its
// 128-byte boxes are issue-bound.  This bench tries the plane-PAIR view instead: two planes are 2 * 2394 floats = 19152 B
// = 19 lines of 252 floats (1008 B, a multiple of 16), and every image row (63 floats) lies inside one line.  So the
// tensor  {252 (inner), 19 lines (stride 1008 B), B*C/2 plane pairs (stride 19152 B)}  is a legal 3-D map, a patch row of a
// channel is a box row at inner coordinate 63 * (y & 3) (+126 / wrap for the odd plane of a pair) + x0 -- NOT 16-byte
// aligned in general -- image rows y and y + 4 are adjacent lines, and even / odd channels are separate boxes.
//
//   map B  dims {252, pairs, lines}, box {32, 16, 2}: 2 patch rows x 16 same-parity channels, lands as [row][channel][128 B]
//          with SWIZZLE_128B_ATOM_32B = the MN-major tf32 operand layout of csrc/corr_umma_fwd.cu (K atoms of 4 channels).
//   map A  dims {252, lines, pairs}, box {16, 2, 16}: query rows y, y+4 x 16 columns x 16 same-parity channels, lands as
//          [channel][row y | row y+4] = 128-byte operand rows of 32 query positions.
//
// Measures: legality of unaligned inner coordinates, data correctness (incl. the swizzle), and sustained throughput of a
// stage = 8 B boxes (4 KB each) + 8 A boxes (2 KB each) = 48 KB, the staging of one 32-channel k-block of the forward.
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

constexpr int H = 38, W = 63, PLANE = H * W, LINE = 4 * W, LINES = 19;
constexpr int STAGE_BYTES = 48 * 1024, NSTAGE = 3;

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
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// flattened float offset of image row y of plane parity par inside a plane pair -> (line, inner)
__device__ __forceinline__ void row_coord(int y, int par, int& line, int& inner) {
    const int off = par * PLANE + W * y;
    line = off / LINE;
    inner = off - line * LINE;
}

// expected value of pixel (pair, par, y, x): index pattern, or 0 outside what the tensor map can see
__host__ __device__ __forceinline__ float pattern(long long idx) { return (float)(idx % 9973); }

struct Job { int b, i0, j0, g, ch; };

__device__ __forceinline__ Job job_of(int it, int C) {
    // walk tiles / chunks / channel blocks in some deterministic order that exercises every alignment
    Job j;
    const int nch = C / 32;
    j.ch = it % nch;
    int r = (it / nch) + blockIdx.x;
    j.g = r % 3; r /= 3;
    j.j0 = (r % 4) * 16; r /= 4;
    j.i0 = (r % 5) * 8; r /= 5;
    j.b = r % 8;
    return j;
}

__global__ void __launch_bounds__(256) stage_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    int C, int iters, int check, int inflight, int alignOnly, long long* cycles, int* bad) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar[NSTAGE];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t base = smem_u32(smem);
    int nbad = 0;
    float sink = 0.f;

    auto issue = [&](int it) {
        const Job j = job_of(it, C);
        const int s = it % NSTAGE;
        const uint32_t st = base + s * STAGE_BYTES;
        mbar_expect_tx(&bar[s], STAGE_BYTES);
        const int pair0 = (j.b * C + j.ch * 32) / 2;
        // A: [parity][atom g4 (rows i0+g4, i0+g4+4)][16 channels][128 B]  = 2 x 4 x 2 KB = 16 KB
        for (int par = 0; par < 2; ++par)
            for (int g4 = 0; g4 < 4; ++g4) {
                int line, inner;
                row_coord(j.i0 + g4, par, line, inner);
                tma_3d(st + (par * 4 + g4) * 2048, &mapA, alignOnly ? ((inner + j.j0) & ~3) : inner + j.j0, line, pair0, &bar[s]);
            }
        // B: [parity][g4][row sel][16 channels][128 B] = 2 x 4 x 4 KB = 32 KB; patch rows y0 + g4 + 4 * sel
        const int y0 = j.i0 - 8 + 8 * j.g;
        for (int par = 0; par < 2; ++par)
            for (int g4 = 0; g4 < 4; ++g4) {
                int line, inner;
                int y = y0 + g4;
                // rows above the image: clamp to a legal row for the bench (the kernel masks them in the epilogue)
                if (y < 0) y += 8;
                row_coord(y, par, line, inner);
                tma_3d(st + 16384 + (par * 4 + g4) * 4096, &mapB, alignOnly ? ((inner + j.j0 - 8) & ~3) : inner + j.j0 - 8, pair0, line, &bar[s]);
            }
    };

    long long t0 = clock64();
    if (threadIdx.x == 0)
        for (int it = 0; it < inflight && it < iters; ++it) issue(it);
    for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        mbar_wait(&bar[s], (it / NSTAGE) & 1);
        if (check) {
            const Job j = job_of(it, C);
            const unsigned char* st = smem + s * STAGE_BYTES;
            const int y0 = j.i0 - 8 + 8 * j.g;
            // B elements: par, g4, sel, k (16), x (32)
            for (int e = threadIdx.x; e < 2 * 4 * 2 * 16 * 32; e += blockDim.x) {
                const int x = e & 31, k = (e >> 5) & 15, sel = (e >> 9) & 1, g4 = (e >> 10) & 3, par = e >> 12;
                const int off = (par * 4 + g4) * 4096 + sel * 2048 + k * 128 + ((((x >> 3) ^ (k & 3)) & 3) * 32) + (x & 7) * 4;
                const float v = *reinterpret_cast<const float*>(st + 16384 + off);
                int y = y0 + g4;
                if (y < 0) y += 8;
                int line, inner;
                row_coord(y, par, line, inner);
                line += sel;
                const int xi = (alignOnly ? ((inner + j.j0 - 8) & ~3) : inner + j.j0 - 8) + x;
                const long long pair = (j.b * C + j.ch * 32) / 2 + k;
                float want = 0.f;
                if (xi >= 0 && xi < LINE && line < LINES) want = pattern(pair * 2 * PLANE + (long long)line * LINE + xi);
                if (v != want) ++nbad;
            }
            // A elements: par, g4, k (16), sel, x (16)
            for (int e = threadIdx.x; e < 2 * 4 * 16 * 2 * 16; e += blockDim.x) {
                const int x = e & 15, sel = (e >> 4) & 1, k = (e >> 5) & 15, g4 = (e >> 9) & 3, par = e >> 11;
                const int mn = sel * 16 + x;
                const int off = (par * 4 + g4) * 2048 + k * 128 + ((((mn >> 3) ^ (k & 3)) & 3) * 32) + (mn & 7) * 4;
                const float v = *reinterpret_cast<const float*>(st + off);
                int line, inner;
                row_coord(j.i0 + g4, par, line, inner);
                line += sel;
                const int xi = (alignOnly ? ((inner + j.j0) & ~3) : inner + j.j0) + x;
                const long long pair = (j.b * C + j.ch * 32) / 2 + k;
                float want = 0.f;
                if (xi >= 0 && xi < LINE && line < LINES) want = pattern(pair * 2 * PLANE + (long long)line * LINE + xi);
                if (v != want) ++nbad;
            }
        } else {
            // touch the stage so the load cannot be elided
            sink += *reinterpret_cast<const float*>(smem + s * STAGE_BYTES + (threadIdx.x & 255) * 16);
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + inflight < iters) issue(it + inflight);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (nbad) atomicAdd(bad, nbad);
    if (sink == 12345.678f) bad[1] = 1;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    const int B = 8, C = 2048;
    const size_t N = (size_t)B * C * PLANE;
    float* d;
    CK(cudaMalloc(&d, N * sizeof(float)));
    {
        std::vector<float> h(N);
        for (size_t i = 0; i < N; ++i) h[i] = pattern((long long)i);
        CK(cudaMemcpy(d, h.data(), N * sizeof(float), cudaMemcpyHostToDevice));
    }
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const cuuint64_t pairs = (cuuint64_t)B * C / 2;
    CUtensorMap mapA, mapB;
    {
        cuuint64_t dims[3] = {LINE, LINES, pairs};
        cuuint64_t strides[2] = {LINE * 4, 2 * PLANE * 4};
        cuuint32_t box[3] = {16, 2, 16};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("map A {252, 19 lines, pairs} box {16, 2, 16} SWIZZLE_128B_ATOM_32B: result %d\n", (int)r);
        if (r != CUDA_SUCCESS) return 1;
    }
    {
        cuuint64_t dims[3] = {LINE, pairs, LINES};
        cuuint64_t strides[2] = {2 * PLANE * 4, LINE * 4};
        cuuint32_t box[3] = {32, 16, 2};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("map B {252, pairs, 19 lines} box {32, 16, 2} SWIZZLE_128B_ATOM_32B (line stride < pair stride): result %d\n", (int)r);
        if (r != CUDA_SUCCESS) return 1;
    }
    long long* cyc;
    int* bad;
    CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
    CK(cudaMalloc(&bad, sizeof(int) * 2));
    const size_t smem = (size_t)NSTAGE * STAGE_BYTES + 1024;
    CK(cudaFuncSetAttribute(stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int alignOnly = 1; alignOnly >= 0; --alignOnly)
    for (int check = 1; check >= 0; --check)
        for (int inflight = 1; inflight <= 2; ++inflight) {
            const int iters = check ? 512 : 4096;
            CK(cudaMemset(bad, 0, sizeof(int) * 2));
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            stage_kernel<<<nsm, 256, smem>>>(mapA, mapB, C, iters, check, inflight, alignOnly, cyc, bad);
            CK(cudaEventRecord(e1));
            { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s box starts: CUDA error \"%s\" -- TMA rejects a box whose first byte is not 16-byte aligned\n", alignOnly ? "16-byte-ALIGNED" : "natural (UNALIGNED)", cudaGetErrorString(e)); return 0; } }
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<long long> h(nsm);
            int hb[2];
            CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hb, bad, sizeof(int) * 2, cudaMemcpyDeviceToHost));
            double avg = 0;
            for (int i = 0; i < nsm; ++i) avg += h[i];
            avg /= nsm;
            printf("%s box starts: stage = 16 boxes / 48 KB, %d in flight, check=%d: %.0f cycles per stage per SM (%.1f B/cycle/SM), %.2f us per stage, "
                   "%.2f TB/s chip, mismatches %d\n",
                   alignOnly ? "16-byte-ALIGNED" : "natural (UNALIGNED)", inflight, check, avg / iters, STAGE_BYTES * (double)iters / avg, ms * 1e3 / iters,
                   (double)STAGE_BYTES * iters * nsm / (ms * 1e-3) / 1e12, hb[0]);
        }
    return 0;
}
