// corr_tile.cu -- tuned float32 PointwiseCorrelation for sm_100a (stride 1, d_max in {4, 8}).
//
// The reference (pointwise_correlation_cuda.cu:63-111) runs one thread per query
// position, accumulating (2d)^2 * C products in GLOBAL memory: 3 thread blocks on a
// 148-SM part for a 38x63 map.  Here the correlation is treated as what it is, a
// banded contraction over channels, and laid out for the FP32 pipe:
//
//   tile    : QROWS x 16 query positions (8x16 for d=8) and their (QROWS+2d-1) x (16+2d-1)
//             key halo patch, staged per channel chunk in shared memory (NCHW is already
//             the right operand layout: for one channel both operands are row vectors
//             over positions, so every FMA is an outer-product term, no transposes).
//   thread  : 8 consecutive queries of one row x all 2d column displacements of ONE row
//             displacement = 8 x 2d accumulators in registers (128 for d=8).  Per channel
//             it needs 8 query values and 8+2d-1 key values (Toeplitz reuse): 8 LDS.128
//             for 128 FFMA.
//   warp    : 4 query rows x 4 row displacements x 2 column halves, so its lanes share
//             4 query rows and 7 key rows (shared-memory broadcast).
//   grid    : stream-K.  The (tile, channel-chunk) iteration space is cut into equal
//             contiguous ranges, one per SM, so B=1 (19-20 tiles) fills 148 SMs and
//             B=8 has no wave-quantisation tail.  A tile whose channels are split over
//             several CTAs goes through fixed-order partial buffers (deterministic, no
//             atomics); tiles owned by one CTA are written straight to `out`.
//   output  : accumulators are transposed through shared memory into the final
//             (B,H,W,2d+1,2d+1) layout, dead entries (row/column 2d, out-of-image
//             displacements: SURVEY.md F4) written as exact zeros, and streamed out as
//             long coalesced runs.  No memset of `out`.
#include <stdlib.h>
#include <string.h>

#include "corr_common.cuh"

namespace d2t {

// shared-memory offset of key-patch row r (rows >= 16 skewed by two 16-byte chunks so that rows r and
// r+16, which one quarter-warp can touch together, fall in different bank groups)
template <int D>
__device__ __forceinline__ int krow_off(int r) {
    return r * FwdCfg<D>::KP + ((r >> 4) << 3);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 4-byte asynchronous global->shared copy, zero-filled when !valid (src must still be a mapped address)
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4s(uint32_t dst, const float* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kCorrBwdUmmaMinC = 128;  // = kCorrTensorMinC of api.cu  // tensor-core backward from this many channels (see use_umma_bwd)
// packed FP32x2 FMA (Blackwell FFMA2): two independent fused multiply-adds per issue slot, same rounding as fmaf.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float2 unpack2(f32x2_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

constexpr int kStages = 3;  // operand ring depth (cp.async groups in flight: kStages - 1)


template <int D, int CK>
__global__ void __launch_bounds__(kCorrThreads, 1)
corr_fwd_tile_kernel(const float* __restrict__ fm0, const float* __restrict__ fm1, float* __restrict__ out,
                     float* __restrict__ partial, CorrPlan p) {
    using Cfg = FwdCfg<D>;
    constexpr int TD = Cfg::TD, K1 = Cfg::K1, KK = Cfg::KK;
    constexpr int STAGE_FLOATS = CK * Cfg::CH_FLOATS;
    extern __shared__ __align__(16) float smem[];  // 2 operand stages, later aliased by the output tile

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // Task = (query row qrow, row displacement ti); its key row inside the staged patch is
    // kr = qrow + ti.  Measured on B200 (profiles/r1_microbench.txt): an LDS.128 costs 4 cycles when
    // the 8 lanes of a quarter-warp read 8 different 16-byte chunks, 2.5 when they read <= 4, and
    // doubles on a bank-group clash.  So a quarter-warp holds 4 query rows x 2 adjacent anti-diagonal
    // classes (qrow + ti == dcls mod 16) of one column half l: its lanes read 4 query chunks and
    // 2-4 key chunks, all in different bank groups.
    const int mlo = tid & 3;
    const int dl = (tid >> 2) & 1;
    const int mhi = (tid >> 3) & 1;
    const int l = (tid >> 4) & 1;
    const int dcls = 2 * warp + dl;  // 0..15
    const int m = mlo + 4 * mhi;     // 0..7
    const int qrow = (Cfg::QROWS == 8) ? m : ((dcls - m) & 15);
    const int ti = (Cfg::QROWS == 8) ? ((dcls - m) & 15) : m;
    const int kr = qrow + ti;  // = dcls or dcls + 16

    const int H = p.H, W = p.W, C = p.C;
    const size_t plane = (size_t)H * W;

    int rnd = 0;
    long long it = (long long)blockIdx.x * p.ipcL;
    const long long itEnd = min((long long)p.left * p.NI, it + p.ipcL);

    while (true) {
        int tile, chunkBeg, chunkEnd, lt = 0;
        if (rnd < p.rounds) {  // whole tiles, in step with the other CTAs
            tile = rnd * p.G + blockIdx.x;
            chunkBeg = 0;
            chunkEnd = p.NI;
            ++rnd;
        } else if (it < itEnd) {  // this CTA's share of the left-over tiles
            lt = (int)(it / p.NI);
            chunkBeg = (int)(it - (long long)lt * p.NI);
            chunkEnd = (int)min((long long)p.NI, chunkBeg + (itEnd - it));
            it += chunkEnd - chunkBeg;
            tile = p.rounds * p.G + lt;
        } else {
            break;
        }

        const int b = tile / (p.tilesX * p.tilesY);
        const int trem = tile - b * p.tilesX * p.tilesY;
        const int i0 = (trem / p.tilesX) * Cfg::QROWS;
        const int j0 = (trem % p.tilesX) * Cfg::QCOLS;

        const float* q_img = fm0 + (size_t)b * C * plane;
        const float* k_img = fm1 + (size_t)b * C * plane;

        // ---- staging map: this thread's key / query elements inside one channel plane -------
        // Loop-invariant per segment: a running global pointer (advanced one plane per channel), the
        // shared-memory byte offset inside a channel block and the copy size (4, or 0 = zero-fill).
        const float* kptr[Cfg::KPASS];
        uint32_t kdst[Cfg::KPASS], ksz[Cfg::KPASS];
#pragma unroll
        for (int ps = 0; ps < Cfg::KPASS; ++ps) {
            const int r = ps * 8 + warp, x = lane;
            const int gi = i0 - D + r, gj = j0 - D + x;
            const bool inPatch = r < Cfg::KROWS && x < Cfg::KP;
            const bool inImg = inPatch && x < Cfg::KCOLS && gi >= 0 && gi < H && gj >= 0 && gj < W;
            kdst[ps] = inPatch ? (uint32_t)(Cfg::QROWS * Cfg::QP + krow_off<D>(r) + x) * 4u : 0xffffffffu;
            ksz[ps] = inImg ? 4u : 0u;
            kptr[ps] = k_img + (size_t)chunkBeg * CK * plane + (inImg ? gi * W + gj : 0);
        }
        const float* qptr[Cfg::QPASS];
        uint32_t qdst[Cfg::QPASS], qsz[Cfg::QPASS];
#pragma unroll
        for (int ps = 0; ps < Cfg::QPASS; ++ps) {
            const int e = ps * kCorrThreads + tid;
            const int r = e / Cfg::QCOLS, x = e % Cfg::QCOLS;
            const bool inTile = e < Cfg::QROWS * Cfg::QCOLS;
            const bool inImg = inTile && i0 + r < H && j0 + x < W;
            qdst[ps] = inTile ? (uint32_t)(r * Cfg::QP + x) * 4u : 0xffffffffu;
            qsz[ps] = inImg ? 4u : 0u;
            qptr[ps] = q_img + (size_t)chunkBeg * CK * plane + (inImg ? (i0 + r) * W + (j0 + x) : 0);
        }

        const bool taskLive = (i0 + qrow < H) && (i0 - D + kr >= 0) && (i0 - D + kr < H);
        const bool warpLive = __any_sync(0xffffffffu, taskLive);

        // Accumulators as FP32 pairs along the key index e = a + t, so that every pair multiplies one (q, q) by one
        // naturally aligned (kv[2m], kv[2m+1]) register pair from the LDS.128: even queries a hold t = (2j, 2j+1),
        // j < TD/2; odd queries hold t = (2j-1, 2j), j <= TD/2, whose first and last halves (t = -1, t = TD) are
        // unused.  68 FFMA2 per channel instead of 128 FFMA (d = 8): half the issue slots for the same FP32-pipe work.
        constexpr int NP = TD / 2;
        f32x2_t acc2[8][NP + 1];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int jj = 0; jj <= NP; ++jj) acc2[a][jj] = 0ull;

        // asynchronous staging of one channel chunk (cp.async: no staging registers, no scoreboard
        // coupling with the LDS of the compute loop).  Chunks are issued strictly in order, so the
        // running pointers are simply advanced.
        const uint32_t smemBase = smem_u32(smem);
        auto issue_chunk = [&](int chunk) {
            if (chunk < chunkEnd) {
                const uint32_t stage = smemBase + (uint32_t)((chunk - chunkBeg) % kStages) * (STAGE_FLOATS * 4u);
                const int nvalid = min(CK, C - chunk * CK);  // channels of this chunk that exist
#pragma unroll
                for (int cc = 0; cc < CK; ++cc) {
                    const bool cv = cc < nvalid;  // uniform; false only in the last chunk of a ragged C
#pragma unroll
                    for (int ps = 0; ps < Cfg::KPASS; ++ps) {
                        if (kdst[ps] != 0xffffffffu)
                            cp_async4s(stage + cc * (Cfg::CH_FLOATS * 4u) + kdst[ps], kptr[ps], cv ? ksz[ps] : 0u);
                        if (cv) kptr[ps] += plane;
                    }
#pragma unroll
                    for (int ps = 0; ps < Cfg::QPASS; ++ps) {
                        if (qdst[ps] != 0xffffffffu)
                            cp_async4s(stage + cc * (Cfg::CH_FLOATS * 4u) + qdst[ps], qptr[ps], cv ? qsz[ps] : 0u);
                        if (cv) qptr[ps] += plane;
                    }
                }
            }
            cp_async_commit();  // always commit: keeps the group count in step with the chunk index
        };

        __syncthreads();  // previous segment's epilogue has finished reading the aliased tile
#pragma unroll
        for (int s0 = 0; s0 < kStages - 1; ++s0) issue_chunk(chunkBeg + s0);

        for (int chunk = chunkBeg; chunk < chunkEnd; ++chunk) {
            cp_async_wait<kStages - 2>();  // this thread's copies of `chunk` have landed
            __syncthreads();               // ... everyone's have; and compute(chunk-1) is finished
            issue_chunk(chunk + kStages - 1);  // refills the stage compute(chunk-1) just released
            const float* stage = smem + ((chunk - chunkBeg) % kStages) * STAGE_FLOATS;
            if (warpLive) {
#pragma unroll
                for (int cc = 0; cc < CK; ++cc) {
                    const float* s = stage + cc * Cfg::CH_FLOATS;
                    const float4* sq = reinterpret_cast<const float4*>(s + qrow * Cfg::QP + 8 * l);
                    const float4* sk = reinterpret_cast<const float4*>(s + Cfg::QROWS * Cfg::QP + krow_off<D>(kr) + 8 * l);
                    float q[8];
                    f32x2_t kk[2 * Cfg::KV];
                    *reinterpret_cast<float4*>(q) = sq[0];
                    *reinterpret_cast<float4*>(q + 4) = sq[1];
#pragma unroll
                    for (int v = 0; v < Cfg::KV; ++v) *reinterpret_cast<ulonglong2*>(kk + 2 * v) = reinterpret_cast<const ulonglong2*>(sk)[v];
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const f32x2_t qq = pack2(q[a], q[a]);
                        const int m0 = a >> 1;               // first key pair used by query a
                        const int np = (a & 1) ? NP + 1 : NP;
#pragma unroll
                        for (int jj = 0; jj < np; ++jj) acc2[a][jj] = ffma2(qq, kk[m0 + jj], acc2[a][jj]);
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // all warps are done with the operand ring before it is reused as the output tile

        // ---- epilogue: registers -> shared tile in final layout -> global ----------------------
        float* tileS = smem;  // [QROWS][QCOLS][K1][K1]
        {
            float* base = tileS + ((qrow * Cfg::QCOLS + 8 * l) * K1 + ti) * K1;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int gj = j0 + 8 * l + a;
#pragma unroll
                for (int t = 0; t < TD; ++t) {
                    const int dj = gj - D + t;
                    const bool live = taskLive && dj >= 0 && dj < W;
                    const int u = (a & 1) ? t + 1 : t;  // position inside the query's pair sequence
                    const float2 pr = unpack2(acc2[a][u >> 1]);
                    base[a * KK + t] = live ? ((u & 1) ? pr.y : pr.x) : 0.f;
                }
            }
            // dead row 2d and dead column 2d of every map in the tile
            for (int e = tid; e < Cfg::QROWS * Cfg::QCOLS * (2 * K1 - 1); e += kCorrThreads) {
                const int pos = e / (2 * K1 - 1), z = e % (2 * K1 - 1);
                const int off = z < K1 ? (K1 - 1) * K1 + z : (z - K1) * K1 + (K1 - 1);
                tileS[pos * KK + off] = 0.f;
            }
        }
        __syncthreads();

        const bool whole = (chunkBeg == 0 && chunkEnd == p.NI);
        const bool dense = (p.os.st == 1 && p.os.sp == KK);   // the reference layout: a tile row is one contiguous run
        if (whole && dense) {
            const int ncols = min(Cfg::QCOLS, W - j0);
            const int run = ncols * KK;
#pragma unroll 1
            for (int r = 0; r < Cfg::QROWS; ++r) {
                if (i0 + r >= H) break;
                float* dst = out + (size_t)b * p.os.sb + ((size_t)(i0 + r) * W + j0) * KK;
                const float* src = tileS + r * Cfg::QCOLS * KK;
                for (int e = tid; e < run; e += kCorrThreads) dst[e] = src[e];
            }
        } else if (whole) {
            // strided output (e.g. channel-major into the tracker's concatenated map): column fastest, so a warp writes
            // two 16-pixel runs and reads tileS at bank (col + t) % 32 (KK = 1 mod 32): conflict-free
            const int ncols = min(Cfg::QCOLS, W - j0), nrows = min(Cfg::QROWS, H - i0);
            float* base = out + (size_t)b * p.os.sb;
            for (int e = tid; e < KK * nrows * Cfg::QCOLS; e += kCorrThreads) {
                const int col = e % Cfg::QCOLS, rt = e / Cfg::QCOLS;
                const int rr = rt % nrows, t = rt / nrows;
                if (col < ncols)
                    base[((long long)(i0 + rr) * W + j0 + col) * p.os.sp + (long long)t * p.os.st] =
                        tileS[(rr * Cfg::QCOLS + col) * KK + t];
            }
        } else {
            // fixed slot per (cta, left-over tile): slot = cta + lt (unique because lt is monotone in cta)
            float4* dst = reinterpret_cast<float4*>(partial + (size_t)(blockIdx.x + lt) * Cfg::TILE_FLOATS);
            const float4* src = reinterpret_cast<const float4*>(tileS);
            for (int e = tid; e < Cfg::TILE_FLOATS / 4; e += kCorrThreads) dst[e] = src[e];
        }
    }
}

// Sum the partial slots of every tile that was split over several CTAs (ascending CTA = ascending
// channel order: deterministic) and write the result in the final layout.
template <int D>
__global__ void __launch_bounds__(256)
corr_fwd_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, CorrPlan p) {
    using Cfg = FwdCfg<D>;
    constexpr int KK = Cfg::KK;
    const int lt = blockIdx.x;  // left-over tile
    const int tile = p.rounds * p.G + lt;
    const long long itBeg = (long long)lt * p.NI, itLast = itBeg + p.NI - 1;
    const int gFirst = (int)(itBeg / p.ipcL), gLast = (int)(itLast / p.ipcL);
    if (gFirst == gLast) return;  // written directly by its only CTA
    const int b = tile / (p.tilesX * p.tilesY);
    const int trem = tile - b * p.tilesX * p.tilesY;
    const int i0 = (trem / p.tilesX) * Cfg::QROWS;
    const int j0 = (trem % p.tilesX) * Cfg::QCOLS;
    const int ncols = min(Cfg::QCOLS, p.W - j0);
    const int run = ncols * KK;
    const bool dense = (p.os.st == 1 && p.os.sp == KK);
    for (int r = blockIdx.y; r < Cfg::QROWS; r += gridDim.y) {
        if (i0 + r >= p.H) break;
        const float* src = partial + (size_t)r * Cfg::QCOLS * KK;
        if (dense) {
            float* dst = out + (size_t)b * p.os.sb + ((size_t)(i0 + r) * p.W + j0) * KK;
            for (int e = threadIdx.x; e < run; e += blockDim.x) {
                float s = 0.f;
                for (int g = gFirst; g <= gLast; ++g) s += src[(size_t)(g + lt) * Cfg::TILE_FLOATS + e];
                dst[e] = s;
            }
        } else {
            float* base = out + (size_t)b * p.os.sb + ((long long)(i0 + r) * p.W + j0) * p.os.sp;
            for (int e = threadIdx.x; e < KK * Cfg::QCOLS; e += blockDim.x) {
                const int col = e % Cfg::QCOLS, t = e / Cfg::QCOLS;
                if (col >= ncols) continue;
                float s = 0.f;
                for (int g = gFirst; g <= gLast; ++g) s += src[(size_t)(g + lt) * Cfg::TILE_FLOATS + col * KK + t];
                base[(long long)col * p.os.sp + (long long)t * p.os.st] = s;
            }
        }
    }
}


// =================================================================================================
// Backward.  Both gradients are one "banded apply":
//     OUT[c, pos] = sum_{si,sj in [0,2d)} G[pos, si, sj] * X[c, pos + (si,sj) - OFF]
//   grad_FM0 (MODE 0): X = FM1, OFF = d,   G[pos,si,sj] = gradOut[pos, si, sj]
//   grad_FM1 (MODE 1): X = FM0, OFF = d-1, G[pos,si,sj] = gradOut[pos + (si,sj) - (d-1), 2d-1-si, 2d-1-sj]
//                      (the queries whose window contains key `pos`; asymmetric because of F4)
// so grad_FM1 is a GATHER too: no atomicAdd (the reference needs C*P of them,
// pointwise_correlation_cuda.cu:169), bitwise reproducible.
//
// Same tile, same key-patch staging and same thread <-> (row, column half, row displacement) map as
// the forward kernel, with the roles turned round: the thread's 8 x 2d block of G is loaded ONCE per
// tile and stays in registers for the whole channel loop; per channel it reads its 8+2d-1 patch
// values, forms 8 partial sums over the column displacements, and the 2d row-displacement partials
// of every output are summed in a fixed order through shared memory.  Every channel chunk produces
// final values, so the (tile, chunk) space is stream-K partitioned with no partial buffers at all.
template <int D, int CK, int MODE>
__global__ void __launch_bounds__(kCorrThreads, 1)
corr_bwd_tile_kernel(const float* __restrict__ go, const float* __restrict__ xsrc, float* __restrict__ gout,
                     CorrPlan p) {
    using Cfg = FwdCfg<D>;
    constexpr int TD = Cfg::TD, K1 = Cfg::K1, KK = Cfg::KK;
    constexpr int OFF = (MODE == 0) ? D : D - 1;
    constexpr int XCH = Cfg::KPATCH;
    constexpr int STAGE_FLOATS = CK * XCH;
    constexpr int RP = CK * 8 + 4;  // reduce-block pitch in floats; RP/4 odd => conflict-free 16-byte rows
    static_assert((RP / 4) % 2 == 1, "reduce pitch");
    extern __shared__ __align__(16) float smem[];
    float* red = smem + kStages * STAGE_FLOATS;  // 2 x [256 threads][RP]: partials of chunk n / n+1
    constexpr int RED_FLOATS = kCorrThreads * RP;
    // the grad_FM1 launch reads nothing this one writes: it may start as soon as SMs are free (programmatic dependent launch)
    if (MODE == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int mlo = tid & 3, dl = (tid >> 2) & 1, mhi = (tid >> 3) & 1, l = (tid >> 4) & 1;
    const int dcls = 2 * warp + dl;
    const int m = mlo + 4 * mhi;
    const int qrow = (Cfg::QROWS == 8) ? m : ((dcls - m) & 15);
    const int ti = (Cfg::QROWS == 8) ? ((dcls - m) & 15) : m;
    const int kr = qrow + ti;

    // reducer role: this thread sums the 2d partials of RV*4 outputs
    constexpr int RV = (Cfg::QROWS == 8) ? 1 : 2;  // float4 per reducer thread
    const int r_ql = (Cfg::QROWS == 8) ? (tid >> 4) : (tid >> 3);
    const int r_cc = (Cfg::QROWS == 8) ? ((tid >> 1) & 7) : (tid & 7);
    const int r_h = (Cfg::QROWS == 8) ? (tid & 1) : 0;
    const int r_qrow = r_ql >> 1, r_l = r_ql & 1;

    const int H = p.H, W = p.W, C = p.C;
    const size_t plane = (size_t)H * W;

    long long it = (long long)blockIdx.x * p.ipc;
    const long long itEnd = min((long long)p.T * p.NI, it + p.ipc);

    // Every chunk of the backward produces final values, so the (tile, chunk) space may be walked in any order.
    // It can be walked channel-group-major (all tiles for channels [g*Gc, (g+1)*Gc) before the next group) so that
    // concurrently running CTAs share channel planes in L2.  Measured on B200 (c5, B=8): Gc = 16/32/64 chunks gives
    // 1975/1873/1787 us against 1725 us for one group -- the kernel is issue-bound, not HBM-bound, and every extra
    // segment pays a G reload and a pipeline refill -- so the default is a single group (D2T_BWD_GC overrides).
    const int Gc = p.dbg > 0 ? p.dbg : p.NI;  // chunks per channel group (host guarantees NI % Gc == 0)
    const long long perGroup = (long long)p.T * Gc;

    while (it < itEnd) {
        const int grp = (int)(it / perGroup);
        const long long rem = it - (long long)grp * perGroup;
        const int tile = (int)(rem / Gc);
        const int cIn = (int)(rem - (long long)tile * Gc);
        const int chunkBeg = grp * Gc + cIn;
        const int chunkEnd = (int)min((long long)(grp + 1) * Gc, chunkBeg + (itEnd - it));
        it += chunkEnd - chunkBeg;

        const int b = tile / (p.tilesX * p.tilesY);
        const int trem = tile - b * p.tilesX * p.tilesY;
        const int i0 = (trem / p.tilesX) * Cfg::QROWS;
        const int j0 = (trem % p.tilesX) * Cfg::QCOLS;
        const float* x_img = xsrc + (size_t)b * C * plane;

        // ---- this thread's block of G, register-resident for the whole channel loop --------------
        float g[8][TD];
        const int pi = i0 + qrow;
        const int xi = pi + ti - OFF;  // patch row in the image
        const bool taskLive = pi < H && xi >= 0 && xi < H;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int pj = j0 + 8 * l + a;
#pragma unroll
            for (int t = 0; t < TD; ++t) {
                const int xj = pj + t - OFF;
                const bool live = taskLive && pj < W && xj >= 0 && xj < W;
                size_t idx;
                if (MODE == 0)
                    idx = (((size_t)b * H + pi) * W + pj) * KK + ti * K1 + t;
                else
                    idx = (((size_t)b * H + xi) * W + xj) * KK + (TD - 1 - ti) * K1 + (TD - 1 - t);
                g[a][t] = live ? __ldg(go + idx) : 0.f;
            }
        }
        const bool warpLive = __any_sync(0xffffffffu, taskLive);

        // ---- staging map for the patch (see the forward kernel) -----------------------------------------
        const float* kptr[Cfg::KPASS];
        uint32_t kdst[Cfg::KPASS], ksz[Cfg::KPASS];
#pragma unroll
        for (int ps = 0; ps < Cfg::KPASS; ++ps) {
            const int r = ps * 8 + warp, x = lane;
            const int gi = i0 - OFF + r, gj = j0 - OFF + x;
            const bool inPatch = r < Cfg::KROWS && x < Cfg::KP;
            const bool inImg = inPatch && x < Cfg::KCOLS && gi >= 0 && gi < H && gj >= 0 && gj < W;
            kdst[ps] = inPatch ? (uint32_t)(krow_off<D>(r) + x) * 4u : 0xffffffffu;
            ksz[ps] = inImg ? 4u : 0u;
            kptr[ps] = x_img + (size_t)chunkBeg * CK * plane + (inImg ? gi * W + gj : 0);
        }
        const uint32_t smemBase = smem_u32(smem);
        auto issue_chunk = [&](int chunk) {
            if (chunk < chunkEnd) {
                const uint32_t stage = smemBase + (uint32_t)((chunk - chunkBeg) % kStages) * (STAGE_FLOATS * 4u);
                const int nvalid = min(CK, C - chunk * CK);
#pragma unroll
                for (int cc = 0; cc < CK; ++cc) {
                    const bool cv = cc < nvalid;
#pragma unroll
                    for (int ps = 0; ps < Cfg::KPASS; ++ps) {
                        if (kdst[ps] != 0xffffffffu)
                            cp_async4s(stage + cc * (XCH * 4u) + kdst[ps], kptr[ps], cv ? ksz[ps] : 0u);
                        if (cv) kptr[ps] += plane;
                    }
                }
            }
            cp_async_commit();
        };

        // reducer: fixed-order sum over the 2d row displacements of chunk `chunk`, then store
        const int o_row = i0 + r_qrow;
        const int o_col = j0 + 8 * r_l + 4 * r_h;
        auto reduce_chunk = [&](int chunk) {
            const float* rbuf = red + ((chunk - chunkBeg) & 1) * RED_FLOATS;
            const int c = chunk * CK + r_cc;
#pragma unroll
            for (int hv = 0; hv < RV; ++hv) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int t = 0; t < TD; ++t) {
                    // thread id holding (r_qrow, r_l, row displacement t)
                    const int dc = (r_qrow + t) & 15;
                    const int mm = (Cfg::QROWS == 8) ? r_qrow : t;
                    const int src = (mm & 3) | ((dc & 1) << 2) | ((mm >> 2) << 3) | (r_l << 4) | ((dc >> 1) << 5);
                    const float4 v = *reinterpret_cast<const float4*>(rbuf + src * RP + r_cc * 8 + 4 * (r_h + hv));
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                if (c < C && o_row < H) {
                    const int col = o_col + 4 * hv;
                    float* dst = gout + ((size_t)b * C + c) * plane + (size_t)o_row * W + col;
                    if (col + 0 < W) dst[0] = s.x;
                    if (col + 1 < W) dst[1] = s.y;
                    if (col + 2 < W) dst[2] = s.z;
                    if (col + 3 < W) dst[3] = s.w;
                }
            }
        };

        __syncthreads();  // the previous segment's last reduce has finished reading `red`
#pragma unroll
        for (int s0 = 0; s0 < kStages - 1; ++s0) issue_chunk(chunkBeg + s0);

        for (int chunk = chunkBeg; chunk < chunkEnd; ++chunk) {
            cp_async_wait<kStages - 2>();
            __syncthreads();  // patch of `chunk` visible; partials of chunk-1 complete; compute(chunk-1) done
            issue_chunk(chunk + kStages - 1);
            if (chunk > chunkBeg) reduce_chunk(chunk - 1);  // overlaps with the FMAs below (no barrier between)
            const float* stage = smem + ((chunk - chunkBeg) % kStages) * STAGE_FLOATS;
            float4* myred = reinterpret_cast<float4*>(red + ((chunk - chunkBeg) & 1) * RED_FLOATS + tid * RP);
#pragma unroll
            for (int cc = 0; cc < CK; ++cc) {
                float part[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) part[a] = 0.f;
                if (warpLive) {
                    const float4* sk = reinterpret_cast<const float4*>(stage + cc * XCH + krow_off<D>(kr) + 8 * l);
                    float kv[4 * Cfg::KV];
#pragma unroll
                    for (int v = 0; v < Cfg::KV; ++v) *reinterpret_cast<float4*>(kv + 4 * v) = sk[v];
#pragma unroll
                    for (int t = 0; t < TD; ++t)
#pragma unroll
                        for (int a = 0; a < 8; ++a) part[a] = fmaf(g[a][t], kv[a + t], part[a]);
                }
                myred[2 * cc] = make_float4(part[0], part[1], part[2], part[3]);
                myred[2 * cc + 1] = make_float4(part[4], part[5], part[6], part[7]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        reduce_chunk(chunkEnd - 1);
    }
}

// ---- host side -------------------------------------------------------------------------------
constexpr int kFwdCK = 8;

int corr_fwd_finalize8_launch(const float* partial, float* out, const CorrPlan& p, cudaStream_t st) {
    if (p.left <= 0) return D2T_OK;
    dim3 grid(p.left, FwdCfg<8>::QROWS);
    corr_fwd_finalize_kernel<8><<<grid, 256, 0, st>>>(partial, out, p);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

bool corr_tile_supported(int B, int C, int H, int W, int d, int stride) {
    if (stride != 1 || (d != 4 && d != 8)) return false;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return false;
    if ((long long)B * C * H * W >= (1ll << 31)) return false;
    return true;
}

template <int D>
static size_t fwd_ws_bytes(int B, int C, int H, int W) {
    CorrPlan p;
    if (make_plan<D>(B, C, H, W, kFwdCK, &p)) return 0;
    DeviceInfo di;
    if (device_info(&di)) return 0;
    return (size_t)(di.sm_count + p.T) * FwdCfg<D>::TILE_FLOATS * sizeof(float);  // enough for any stream-K split
}

size_t corr_tile_fwd_ws_bytes(int B, int C, int H, int W, int d) {
    return d == 8 ? fwd_ws_bytes<8>(B, C, H, W) : fwd_ws_bytes<4>(B, C, H, W);
}
bool corr_tile_bwd_supported(int B, int C, int H, int W, int d, int stride) {
    return corr_tile_supported(B, C, H, W, d, stride);
}

// tensor-core backward (corr_umma_bwd.cu)
bool corr_umma_bwd_supported(int B, int C, int H, int W, int d, int stride);
size_t corr_umma_bwd_ws_bytes(int B, int C, int H, int W);
int corr_umma_bwd_launch(const float*, const float*, const float*, float*, float*, int, int, int, int, void*, size_t,
                         cudaStream_t);

// Which kernel family d2t_corr_bwd_f32 runs.  The rule depends on (C, d_max, stride) ONLY -- never on B, H or W -- so
// that gradients are batch-invariant: the tensor-core kernel (3xTF32) for d_max = 8, stride = 1 and at least
// kCorrBwdUmmaMinC channels, the FP32-pipe kernel otherwise.  d2t_corr_bwd_f32_simt / d2t_corr_bwd_f32_tc select a
// family explicitly.  Measured on B200 (38x63, d=8; profiles/r1_time_ops_v4.txt): B=8 c3/c4/c5 253/386/620 us against
// 437/861/1729 us for the FP32-pipe kernel; B=1: c4 112 vs 142, c5 185 vs 236, c3 112 vs 95.
static bool use_umma_bwd(int B, int C, int H, int W, int d) {
    return corr_umma_bwd_supported(B, C, H, W, d, 1) && C >= kCorrBwdUmmaMinC;
}
size_t corr_tile_bwd_ws_bytes(int B, int C, int H, int W, int d) {
    return use_umma_bwd(B, C, H, W, d) ? corr_umma_bwd_ws_bytes(B, C, H, W) : 0;
}

template <int D>
static int fwd_launch(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, const CorrOutStrides* os,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
    using Cfg = FwdCfg<D>;
    constexpr int CK = kFwdCK;
    CorrPlan p;
    int rc = make_plan<D>(B, C, H, W, CK, &p);
    if (rc) return rc;
    if (os) p.os = *os;
    // Whole-tile rounds (all CTAs in step, halos shared through L2) were measured SLOWER than one contiguous
    // (tile, chunk) range per CTA at B = 8 (c5: 587 vs ~510 us): with every CTA on the same channel planes at the
    // same time the memory system is hit in bursts.  So the schedule is the staggered stream-K walk.
    p.rounds = 0;
    p.left = p.T;
    p.ipcL = p.ipc;
    const size_t need = (size_t)(p.G + p.T) * Cfg::TILE_FLOATS * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
        set_error("corr_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
        return D2T_ERR_WORKSPACE;
    }
    const size_t opBytes = (size_t)kStages * CK * Cfg::CH_FLOATS * sizeof(float);
    const size_t tileBytes = (size_t)Cfg::TILE_FLOATS * sizeof(float);
    const size_t smem = opBytes > tileBytes ? opBytes : tileBytes;
    auto kern = corr_fwd_tile_kernel<D, CK>;
    D2T_SMEM_OPTIN(kern, smem);
    kern<<<p.G, kCorrThreads, smem, st>>>(fm0, fm1, out, static_cast<float*>(ws), p);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    if (p.left > 0 && p.ipcL % p.NI != 0) {  // some left-over tile is split over CTAs
        dim3 grid(p.left, Cfg::QROWS);
        corr_fwd_finalize_kernel<D><<<grid, 256, 0, st>>>(static_cast<const float*>(ws), out, p);
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
    }
    return D2T_OK;
}

int corr_tile_fwd_launch(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, int d,
                         const CorrOutStrides* os, void* ws, size_t ws_bytes, cudaStream_t st) {
    return d == 8 ? fwd_launch<8>(fm0, fm1, out, B, C, H, W, os, ws, ws_bytes, st)
                  : fwd_launch<4>(fm0, fm1, out, B, C, H, W, os, ws, ws_bytes, st);
}

int corr_tile_bwd_simt_launch(const float*, const float*, const float*, float*, float*, int, int, int, int, int,
                              cudaStream_t);

template <int D>
static int bwd_launch(const float* go, const float* fm0, const float* fm1, float* g0, float* g1, int B, int C, int H,
                      int W, cudaStream_t st) {
    using Cfg = FwdCfg<D>;
    constexpr int CK = 8;
    CorrPlan p;
    int rc = make_plan<D>(B, C, H, W, CK, &p);
    if (rc) return rc;
    p.dbg = p.NI;  // one channel group: grouping the walk for L2 reuse was measured slower (segment restarts)
    const size_t smem = ((size_t)kStages * CK * Cfg::KPATCH + (size_t)2 * kCorrThreads * (CK * 8 + 4)) * sizeof(float);
    auto k0 = corr_bwd_tile_kernel<D, CK, 0>;
    auto k1 = corr_bwd_tile_kernel<D, CK, 1>;
    D2T_SMEM_OPTIN(k0, smem);
    D2T_SMEM_OPTIN(k1, smem);
    k0<<<p.G, kCorrThreads, smem, st>>>(go, fm1, g0, p);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(p.G);
        cfg.blockDim = dim3(kCorrThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        D2T_CUDA_TRY(cudaLaunchKernelEx(&cfg, k1, go, fm0, g1, p));
    }
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

int corr_tile_bwd_launch(const float* go, const float* fm0, const float* fm1, float* g0, float* g1, int B, int C,
                         int H, int W, int d, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (use_umma_bwd(B, C, H, W, d)) {
        if (ws == nullptr || ws_bytes < corr_umma_bwd_ws_bytes(B, C, H, W)) {
            set_error("corr_bwd: workspace too small (%zu < %zu)", ws_bytes, corr_umma_bwd_ws_bytes(B, C, H, W));
            return D2T_ERR_WORKSPACE;  // never a silent change of kernel family (and of numerics)
        }
        return corr_umma_bwd_launch(go, fm0, fm1, g0, g1, B, C, H, W, ws, ws_bytes, st);
    }
    return corr_tile_bwd_simt_launch(go, fm0, fm1, g0, g1, B, C, H, W, d, st);
}

int corr_tile_bwd_simt_launch(const float* go, const float* fm0, const float* fm1, float* g0, float* g1, int B, int C,
                              int H, int W, int d, cudaStream_t st) {
    return d == 8 ? bwd_launch<8>(go, fm0, fm1, g0, g1, B, C, H, W, st)
                  : bwd_launch<4>(go, fm0, fm1, g0, g1, B, C, H, W, st);
}

}  // namespace d2t
