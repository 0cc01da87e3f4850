// corr_umma.cu -- PointwiseCorrelation forward on the 5th-generation tensor cores (tcgen05 + TMEM), d_max = 8.
//
// Formulation: one tile of 8x16 query positions (M = 128) against its 23x32 key halo patch (N = 736 key
// positions) is a GEMM over channels, D[m][n] = sum_c Q[c][m] * K[c][n]; the correlation map of query m is the
// 16x16 band {n = (qrow+ti)*32 + qcol+tj} of row m.  The dense tile does 2.9x the useful MACs -- the price of
// feeding a GEMM engine -- and FP32 inputs are split 3xTF32 (hi*hi + hi*lo + lo*hi, hi = top 19 bits) so the
// result keeps ~22 bits (measured: |err| <= 4e-6 * sum|a||b|, tools/umma_test.cu), i.e. ~8.6x the useful MACs at
// the TF32 rate, still ~3x faster than the FP32-pipe band kernel (corr_tile.cu).
//
//   operands : K-major canonical UMMA layout without swizzle (core matrix = 8 positions x 4 channels).  Rows of a
//              38x63 NCHW map are not 16-byte aligned, so neither TMA nor wide copies apply (tools/tma_bench.cu) and
//              4-byte cp.async is LSU-bound; operands go global -> registers (coalesced LDG, prefetched two stages
//              ahead) -> hi/lo split in registers -> STS.128 straight into the UMMA layout.
//              MN-major without swizzle silently yields zeros for kind::tf32 (tools/umma_test.cu), so the natural
//              position-contiguous layout is not usable.
//   TMEM     : 512 columns hold at most N = 512, so the patch is processed in two passes over the channel range
//              (key rows 0-11: N = 384, key rows 12-22: N = 352); per 8-channel block 2 x 3 MMAs (N = 256 + rest).
//   pipeline : 3-stage ring of 16-channel stages (64 KB each); all threads stage, one thread issues the MMAs and
//              commits them to the stage's mbarrier, which gates the refill of that stage; one barrier per stage.
//   epilogue : tcgen05.ld (32 lanes x 32 columns = one key row), band extraction into a per-query-row buffer in the
//              final (17x17 per position) layout, coalesced copy-out; dead row/column 16 written as zeros.
//   grid     : the same stream-K plan and partial-slot / finalize machinery as the SIMT kernel.
#include <stdlib.h>

#include "corr_common.cuh"

namespace d2t {

namespace {

constexpr int UD = 8;                 // d_max
constexpr int UKC = 16;               // channels per stage
constexpr int UNS = 3;                // stages
constexpr int UM = 128;               // queries per tile
constexpr int UNMAX = 384;            // key positions per pass (max)
constexpr int UTHREADS = 256;
constexpr int UKB = UKC / 8;          // 8-channel MMA k-blocks per stage
constexpr int UBLK_A = UM * 8;        // floats of one A k-block
constexpr int UBLK_B = UNMAX * 8;     // floats of one B k-block
constexpr int UHI_FLOATS = UKB * (UBLK_A + UBLK_B);  // hi region of a stage (lo region has the same shape)
constexpr int USTAGE_FLOATS = 2 * UHI_FLOATS;
constexpr int URP = 290;              // row-buffer pitch per query (289 used): odd (URP - 1) => conflict-free band stores
constexpr int UROWBUF_FLOATS = 16 * URP;
constexpr int UPASS_ROWS0 = 12;       // key rows 0..11 in pass 0, 12..22 in pass 1

__device__ __forceinline__ uint32_t u_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void u_cp_async4(uint32_t dst, const float* src, uint32_t bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void u_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void u_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void u_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(u_smem(bar)), "r"(count));
}
__device__ __forceinline__ void u_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            u_smem(bar)),
        "r"(parity)
        : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrices 128 B apart along K (LBO), 256 B along M/N (SBO)
__device__ __forceinline__ uint64_t u_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t u_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
}
__device__ __forceinline__ void u_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void u_mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_smem(bar)) : "memory");
}
// element (pos, c) of a k-block, K-major canonical layout, in floats
__device__ __forceinline__ int u_off(int pos, int c) { return (pos >> 3) * 64 + (c >> 2) * 32 + (pos & 7) * 4 + (c & 3); }

__global__ void __launch_bounds__(UTHREADS, 1)
corr_fwd_umma_kernel(const float* __restrict__ fm0, const float* __restrict__ fm1, float* __restrict__ out,
                     float* __restrict__ partial, CorrPlan p) {
    extern __shared__ __align__(128) float smem[];
    float* rowbuf = smem + UNS * USTAGE_FLOATS;
    __shared__ __align__(8) uint64_t bar_stage[UNS];
    __shared__ __align__(8) uint64_t bar_accum;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, C = p.C;
    const size_t plane = (size_t)H * W;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(u_smem(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < UNS; ++s) u_mbar_init(&bar_stage[s], 1);
        u_mbar_init(&bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    uint32_t stage_phase = 0;  // bit s: parity to wait for on bar_stage[s]
    uint32_t accum_phase = 0;
    uint32_t stage_used = 0;   // bit s: stage s has an un-waited commit

    long long it = (long long)blockIdx.x * p.ipc;
    const long long itEnd = min((long long)p.T * p.NI, it + p.ipc);

    while (it < itEnd) {
        const int tile = (int)(it / p.NI);
        const int chunkBeg = (int)(it - (long long)tile * p.NI);
        const int chunkEnd = (int)min((long long)p.NI, chunkBeg + (itEnd - it));
        it += chunkEnd - chunkBeg;
        const int nChunks = chunkEnd - chunkBeg;

        const int b = tile / (p.tilesX * p.tilesY);
        const int trem = tile - b * p.tilesX * p.tilesY;
        const int i0 = (trem / p.tilesX) * 8;
        const int j0 = (trem % p.tilesX) * 16;
        const float* q_img = fm0 + (size_t)b * C * plane;
        const float* k_img = fm1 + (size_t)b * C * plane;
        const bool whole = (chunkBeg == 0 && chunkEnd == p.NI);
        float* slot = partial + (size_t)(blockIdx.x + tile) * (UM * 289);

        for (int pass = 0; pass < 2; ++pass) {
            const int kr0 = pass == 0 ? 0 : UPASS_ROWS0;
            const int nRows = pass == 0 ? UPASS_ROWS0 : 23 - UPASS_ROWS0;
            const int Npass = nRows * 32;  // 384 or 352

            // ---- staging: global -> registers -> (hi, lo) -> shared, no cp.async ------------------------------------------
            // 4-byte cp.async costs ~8 LSU cycles per warp instruction on this part (256 of them per stage would
            // outlast the stage's MMAs); LDG + STS.128 is ~2.7x cheaper and lets the 3xTF32 split happen in registers.
            // A thread owns two positions of the stage (lane = position inside a 32-wide row => coalesced 128-byte
            // loads): row group `warp` (queries for warps 0-3, key rows 4.. for warps 4-7) and row group `warp + 8`
            // (key rows), each for all 16 channels of the stage.
            const float* ptr[2];
            uint32_t dstf[2];  // float offset of (pos, channel 0) inside a k-block
            bool ok[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int rg = warp + 8 * u;  // 0..3 queries, 4..15 key rows (kr0 + rg - 4)
                int yy, xx, pos, opbase;
                const float* img;
                if (rg < 4) {
                    pos = rg * 32 + lane;
                    yy = i0 + (pos >> 4); xx = j0 + (pos & 15);
                    ok[u] = yy < H && xx < W;
                    img = q_img; opbase = 0;
                } else {
                    pos = (rg - 4) * 32 + lane;
                    yy = i0 - UD + kr0 + (rg - 4); xx = j0 - UD + lane;
                    ok[u] = (rg - 4) < nRows && lane < 31 && yy >= 0 && yy < H && xx >= 0 && xx < W;
                    img = k_img; opbase = UBLK_A;
                }
                dstf[u] = (uint32_t)(opbase + (pos >> 3) * 64 + (pos & 7) * 4);
                ptr[u] = img + (size_t)chunkBeg * UKC * plane + (ok[u] ? yy * W + xx : 0);
            }

            auto load_regs = [&](float (&r)[2][UKC], int k) {  // chunk k (relative) -> registers
                if (k < nChunks) {
                    const int nvalid = C - (chunkBeg + k) * UKC;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
#pragma unroll
                        for (int cc = 0; cc < UKC; ++cc)
                            r[u][cc] = (ok[u] && cc < nvalid && !(p.dbg & 4)) ? __ldg(ptr[u] + (size_t)cc * plane) : 0.f;
                        ptr[u] += (size_t)UKC * plane;
                    }
                }
            };
            auto convert_store = [&](const float (&r)[2][UKC], int s) {  // 3xTF32 split in registers, STS.128
                float* hiS = smem + s * USTAGE_FLOATS;
                float* loS = hiS + UHI_FLOATS;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int quad = 0; quad < UKC / 4; ++quad) {
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(r[u][quad * 4 + 0]) & 0xFFFFE000u);
                        h.y = __uint_as_float(__float_as_uint(r[u][quad * 4 + 1]) & 0xFFFFE000u);
                        h.z = __uint_as_float(__float_as_uint(r[u][quad * 4 + 2]) & 0xFFFFE000u);
                        h.w = __uint_as_float(__float_as_uint(r[u][quad * 4 + 3]) & 0xFFFFE000u);
                        l.x = r[u][quad * 4 + 0] - h.x; l.y = r[u][quad * 4 + 1] - h.y;
                        l.z = r[u][quad * 4 + 2] - h.z; l.w = r[u][quad * 4 + 3] - h.w;
                        const int off = (quad >> 1) * (UBLK_A + UBLK_B) + (quad & 1) * 32 + dstf[u];
                        *reinterpret_cast<float4*>(hiS + off) = h;
                        *reinterpret_cast<float4*>(loS + off) = l;
                    }
                }
            };
            const uint32_t smemBase = u_smem(smem);
            auto run_chunk = [&](float (&r)[2][UKC], int k) {
                const int s = k % UNS;
                if ((stage_used >> s) & 1u) {  // the MMAs that last read this stage must have completed
                    u_mbar_wait(&bar_stage[s], (stage_phase >> s) & 1u);
                    stage_phase ^= 1u << s;
                    stage_used &= ~(1u << s);
                }
                if (!(p.dbg & 1)) convert_store(r, s);
                load_regs(r, k + 2);  // prefetch two chunks ahead into the registers just freed
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    if (p.dbg & 2) {
                        u_mma_commit(&bar_stage[s]);
                        if (k == nChunks - 1) u_mma_commit(&bar_accum);
                    } else {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sbase = smemBase + (uint32_t)s * (USTAGE_FLOATS * 4u);
#pragma unroll
                        for (int kb = 0; kb < UKB; ++kb) {
                            const uint32_t aHi = sbase + (uint32_t)(kb * (UBLK_A + UBLK_B)) * 4u;
                            const uint32_t bHi = aHi + UBLK_A * 4u;
                            const uint32_t aLo = aHi + UHI_FLOATS * 4u, bLo = bHi + UHI_FLOATS * 4u;
                            const uint32_t acc = (k > 0 || kb > 0) ? 1u : 0u;
                            for (int n0 = 0; n0 < Npass; n0 += 256) {
                                const int nn = min(256, Npass - n0);
                                const uint32_t idesc = u_idesc(nn);
                                const uint32_t boff = (uint32_t)n0 * 32u;  // n0 positions * 8 channels * 4 B
                                u_mma(tmem_base + n0, u_desc(aHi), u_desc(bHi + boff), idesc, acc);
                                u_mma(tmem_base + n0, u_desc(aHi), u_desc(bLo + boff), idesc, 1u);
                                u_mma(tmem_base + n0, u_desc(aLo), u_desc(bHi + boff), idesc, 1u);
                            }
                        }
                        u_mma_commit(&bar_stage[s]);
                        if (k == nChunks - 1) u_mma_commit(&bar_accum);
                    }
                }
                stage_used |= 1u << s;
            };

            float regA[2][UKC], regB[2][UKC];
            load_regs(regA, 0);
            load_regs(regB, 1);
            for (int k = 0; k < nChunks; k += 2) {
                run_chunk(regA, k);
                if (k + 1 < nChunks) run_chunk(regB, k + 1);
            }

            // ---- epilogue of this pass --------------------------------------------------------------------------
            u_mbar_wait(&bar_accum, accum_phase);
            accum_phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            const int quarter = warp & 3;            // TMEM lanes 32*quarter .. +31
            const int m = quarter * 32 + lane;       // query index in the tile
            const int qrow = m >> 4, qcol = m & 15;
            const int par = warp >> 2;               // this warp takes key rows of parity `par` within the pass
            for (int qr = 0; qr < 8; ++qr) {
                // ti range of query row qr produced by this pass (row 16 of every map is dead: zeros, with pass 1)
                const int tiLo = max(0, kr0 - qr), tiHi = min(15, kr0 + nRows - 1 - qr);
                const int tiEnd = (pass == 1) ? 17 : tiHi + 1;  // exclusive, in 17-float rows
                const bool mine = (qr >> 1) == quarter;         // warps holding this query row
                if (mine) {
                    for (int kr = max(kr0, qr); kr <= min(kr0 + nRows - 1, qr + 15); ++kr) {
                        if (((kr - kr0) & 1) != par) continue;  // warp-uniform: the two warps of a quarter split the rows
                        uint32_t r[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((kr - kr0) * 32);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (qrow == qr) {
                            float* dstq = rowbuf + qcol * URP + (kr - qr) * 17 - qcol;
#pragma unroll
                            for (int x = 0; x < 32; ++x)
                                if (x >= qcol && x < qcol + 16) dstq[x] = __uint_as_float(r[x]);
                            rowbuf[qcol * URP + (kr - qr) * 17 + 16] = 0.f;  // dead column 16
                        }
                    }
                    if (pass == 1 && qrow == qr && par == 0) {
#pragma unroll
                        for (int x = 0; x < 17; ++x) rowbuf[qcol * URP + 16 * 17 + x] = 0.f;  // dead row 16
                    }
                }
                __syncthreads();
                // copy rows [tiLo, tiEnd) of the 16 queries of this row to their destination
                {
                    const int len = (tiEnd - tiLo) * 17;
                    const int gi = i0 + qr;
                    const int ncols = whole ? min(16, W - j0) : 16;
                    if (len > 0 && (!whole || gi < H)) {
                        float* dbase = whole ? out + (((size_t)b * H + gi) * W + j0) * 289 : slot + (size_t)qr * 16 * 289;
                        const int q = tid >> 4;  // 16 threads per query
                        if (q < ncols)
                            for (int o = tid & 15; o < len; o += 16) dbase[q * 289 + tiLo * 17 + o] = rowbuf[q * URP + tiLo * 17 + o];
                    }
                }
                __syncthreads();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();  // TMEM reads done before the next pass overwrites the accumulators
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

}  // namespace

bool corr_umma_supported(int B, int C, int H, int W, int d, int stride) {
    if (stride != 1 || d != 8) return false;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return false;
    if ((long long)B * C * H * W >= (1ll << 31)) return false;
    return true;
}

int corr_umma_fwd_launch(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
    CorrPlan p;
    int rc = make_plan<8>(B, C, H, W, UKC, &p);
    if (rc) return rc;
    if (const char* e = getenv("D2T_UMMA_DBG")) p.dbg = atoi(e);
    const size_t need = (size_t)(p.G + p.T) * FwdCfg<8>::TILE_FLOATS * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
        set_error("corr_fwd(umma): workspace too small (%zu < %zu)", ws_bytes, need);
        return D2T_ERR_WORKSPACE;
    }
    const size_t smem = ((size_t)UNS * USTAGE_FLOATS + UROWBUF_FLOATS) * sizeof(float);
    D2T_CUDA_TRY(cudaFuncSetAttribute(corr_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    corr_fwd_umma_kernel<<<p.G, UTHREADS, smem, st>>>(fm0, fm1, out, static_cast<float*>(ws), p);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    if (p.ipc % p.NI != 0) return corr_fwd_finalize8_launch(static_cast<const float*>(ws), out, p, st);
    return D2T_OK;
}

}  // namespace d2t
