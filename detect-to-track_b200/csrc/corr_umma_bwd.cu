// corr_umma_bwd.cu -- PointwiseCorrelation backward on the 5th-generation tensor cores (tcgen05 + TMEM), d_max = 8.
//
// Both gradients are the banded apply of corr_tile.cu,
//     OUT[c, pos] = sum_{si,sj in [0,16)} G[pos, si, sj] * X[c, pos + (si,sj) - OFF]
//   grad_FM0: X = FM1, OFF = 8, G[pos,si,sj] = gradOut[pos, si, sj]                                 (reference
//   grad_FM1: X = FM0, OFF = 7, G[pos,si,sj] = gradOut[pos + (si,sj) - 7, 15 - si, 15 - sj]         pointwise_correlation_cuda.cu:145-172)
// which for one tile of 8x16 positions is a GEMM over the tile's 23x31 halo patch of X:
//     D[m = position (128)][n = channel] = sum_{k = patch element} S[m][k] * X[n][k],
//     S[(qrow,qcol)][(r,x)] = G[(qrow,qcol), r - qrow, x - qcol]  inside the 16x16 band, 0 outside.
// K runs over POSITIONS, which are contiguous in an NCHW plane, so X is K-major as it lies in memory: one patch row
// (31 keys + 1 zero pad = 32 tf32 = 128 bytes) is exactly one row of the K-major SWIZZLE_128B operand layout.  S is
// built on the fly from gradOut (L2-resident).  FP32 inputs are split 3xTF32 (hi*hi + hi*lo + lo*hi, hi rounded to
// nearest) -- measured |err| <= 9e-7 * sum|a||b| (tools/umma_sw128_test.cu); the dense tile does ~2.8x the band's MACs.
// Against the FP32-pipe band kernel (corr_tile.cu: 128 FFMA + 10 shared-memory accesses per thread and channel, bound by
// the FMA and shared-memory pipes together) this moves the contraction to the tensor pipe and leaves the CUDA cores
// only the operand staging.
//
//   work item  (image b, block of 256 channels, tile): M = 128, N = 256 accumulators = 256 TMEM columns; K walks the
//              tile's live patch rows (rows outside the image are skipped, not multiplied).  Items are numbered with
//              the tile fastest, so CTAs running together share an image and a channel block in L2.
//   chunk      one patch row: A = S[128][32], B = X[256][32], each as hi and lo (96 KB per stage, 2 stages);
//              4 k-steps x 3 MMAs (128x256x8, kind::tf32) per chunk.
//   warps 0-7  producers: lane = key column.  Warp w stages channels 32w..32w+31 of B (one coalesced 124-byte LDG and
//              two conflict-free 128-byte-row STS per channel) and query row w of A.  Loads run two chunks ahead in
//              registers, across item boundaries, and never stop for an epilogue.
//   warp 8     issues the MMAs (one lane) once the stage's `full` barrier completes; tcgen05.commit -> `empty` barrier
//              of the stage / `accum_full` of the item.  It must be a warp of its own: tcgen05.mma issue blocks while
//              the tensor core's queue is full, and a producer warp that blocks there serialises staging and MMAs
//              (measured: stage + MMA + handshake times added up exactly).  Registers are allocated per 4 warps, so the
//              third warpgroup (warps 8-11) hands its registers to the producers with setmaxnreg (40 / 200).
//   warps 12-15 epilogue: tcgen05.ld of the finished item's accumulators -> grad[b][n][position].  Two 256-column
//              TMEM buffers alternate between items, so draining item t overlaps the MMAs of item t+1.
//   grad_FM1   needs gradOut transposed (the queries whose window contains a key): corr_bwd_flip_kernel writes
//              GT[b,p,si,sj] = gradOut[b, p + (si,sj) - 7, 15 - si, 15 - sj] (0 outside the image) into the workspace
//              once per call (42 MB of traffic at B = 8), so that both gradients use the same staging code.
#include <stdlib.h>
#include <string.h>

#include "corr_common.cuh"

// Timing ablation only (tools/experiments/staging_ablation.sh; NEVER defined by csrc/Makefile -- results are garbage):
//   1  the producers skip the global loads of the X operand (what the LDGs cost)
//   2  ... and replace its scalar hi/lo stores by what the split warps of a TMA-fed kernel would execute if the raw X tile
//      had landed in shared memory by itself: LDS.128 raw -> lo = v - trunc_tf32(v) -> STS.128 lo  (raw doubles as hi)
//   3  as 2, with the rounded hi written back in place as well (the gemm_tf32x3 recipe)
// The S operand (built from gradOut) is staged as in the product in every mode.
#ifndef D2T_ABLATE_STAGING
#define D2T_ABLATE_STAGING 0
#endif

namespace d2t {

namespace {

constexpr int XD = 8;                     // d_max
constexpr int XTD = 16;                   // live displacements per axis
constexpr int XK1 = 17, XKK = 289;        // gradOut map side / size
constexpr int XM = 128;                   // positions per tile = UMMA M
constexpr int XN = 256;                   // channels per item = UMMA N = TMEM columns
constexpr int XQROWS = 8, XQCOLS = 16;
constexpr int XROWS = XQROWS + XTD - 1;   // 23 patch rows
constexpr int XPROD_WARPS = 8;
constexpr int XEPI_WARP0 = XPROD_WARPS + 4;        // warps 12-15: epilogue (TMEM lane quarter = warp % 4)
constexpr int XTHREADS = (XPROD_WARPS + 8) * 32;  // + MMA issuer warp 8 (9-11 idle) + four epilogue warps
constexpr int XA_BYTES = XM * 128;        // one A operand (hi or lo) of a chunk
constexpr int XB_BYTES = XN * 128;        // one B operand (hi or lo) of a chunk
constexpr int XSTAGE_BYTES = 2 * XA_BYTES + 2 * XB_BYTES;
constexpr int XSTAGES = 2;
constexpr int XNA = XQCOLS, XNB = XN / XPROD_WARPS;  // A / B elements a producer thread stages per chunk
constexpr int XFLIP_THREADS = 512;
constexpr int XFLIP_PITCH = 290;          // shared-memory pitch of one gradOut map in the flip kernel

struct XPlan {
    int B, C, H, W;
    int tilesX, tilesY, nCb, nItems;
};

__device__ __forceinline__ uint32_t x_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void x_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(x_smem(bar)), "r"(count));
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires)
// instead of polling -- every poll is a shared-memory access, and the kernel is bound by the shared-memory data pipe
// (ncu: 7.7 M of 28.9 M LSU shared wavefronts per launch were barrier polls).
__device__ __forceinline__ void x_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            x_smem(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void x_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(x_smem(bar)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): rows of 128 bytes,
// 8-row atoms 1024 bytes apart (SBO); the start address may point 32*ks bytes into the row to select a k-step.
__device__ __forceinline__ uint64_t x_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t kXIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(XN >> 3) << 17) | ((uint32_t)(XM >> 4) << 24);
__device__ __forceinline__ void x_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kXIdesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void x_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(x_smem(bar)) : "memory");
}
__device__ __forceinline__ void x_sts(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float x_ldg_stream(uint64_t addr) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(addr));
    return v;
}
// round-to-nearest split: hi keeps 10 explicit mantissa bits (tf32), lo = v - hi is exact in fp32 and |lo| <= 2^-11 |v|
__device__ __forceinline__ float x_tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

template <int V>
struct XInt { static constexpr int value = V; };

// walks the (item, live patch row) sequence of one CTA
struct XCursor {
    int item, r, rLo, rHi;
    int b, cb, i0, j0;
    __device__ __forceinline__ bool valid(const XPlan& p) const { return item < p.nItems; }
    __device__ __forceinline__ void decode(const XPlan& p, int off) {
        if (item >= p.nItems) return;
        const int tpi = p.tilesX * p.tilesY;
        const int t = item % tpi;
        const int rest = item / tpi;
        cb = rest % p.nCb;
        b = rest / p.nCb;
        i0 = (t / p.tilesX) * XQROWS;
        j0 = (t % p.tilesX) * XQCOLS;
        const int nq = min(XQROWS, p.H - i0);           // live query rows of the tile
        rLo = max(0, off - i0);                           // first patch row inside the image
        rHi = min(min(XROWS, p.H + off - i0), nq + XTD - 1);  // one past the last row that is in the image and in a band
        r = rLo;
    }
    __device__ __forceinline__ void start(const XPlan& p, int off) { item = blockIdx.x; decode(p, off); }
    __device__ __forceinline__ bool last() const { return r == rHi - 1; }
    __device__ __forceinline__ void advance(const XPlan& p, int off) {
        if (++r >= rHi) { item += gridDim.x; decode(p, off); }
    }
};

// MODE 0: grad_FM0 (G = gradOut, 17x17 maps, OFF = 8).  MODE 1: grad_FM1 (G = flipped gradOut, 16x16 maps, OFF = 7).
template <int MODE>
__global__ void __launch_bounds__(XTHREADS, 1)
corr_bwd_umma_kernel(const float* __restrict__ gsrc, const float* __restrict__ xsrc, float* __restrict__ gout, XPlan p) {
    constexpr int OFF = MODE == 0 ? XD : XD - 1;
    constexpr int GMAP = MODE == 0 ? XKK : XTD * XTD;
    constexpr int GROW = MODE == 0 ? XK1 : XTD;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[XSTAGES], bar_empty[XSTAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, C = p.C;
    const size_t plane = (size_t)H * W;
    const uint64_t planeBytes = (uint64_t)plane * sizeof(float);

    // The grad_FM1 kernel does not read anything this kernel writes (its input is the flipped gradOut of the flip kernel, which
    // is complete before this kernel starts): let it be launched as this kernel's CTAs exit, so that the SMs whose CTA had one
    // item less than the others start on grad_FM1 instead of idling through the tail (programmatic dependent launch)
    if (MODE == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(x_smem(&tmem_base_s)), "r"((uint32_t)(2 * XN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < XSTAGES; ++s) {
            x_mbar_init(&bar_full[s], XPROD_WARPS);
            x_mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            x_mbar_init(&bar_acc_full[a], 1);
            x_mbar_init(&bar_acc_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smemBase = x_smem(smem);

    // Row `row` of an operand lives at (row>>3)*1024 + (row&7)*128 and its 16-byte chunks are XOR-swizzled with row&7, so
    // key column `lane` of a row with row&7 == t sits at byte (4*lane) ^ (t << 4).  This warp's rows start at a multiple
    // of 8, so eight per-thread base addresses (stage 0) cover everything else with compile-time offsets.
    uint32_t stsA[8], stsB[8];
#pragma unroll
    for (int t8 = 0; t8 < 8; ++t8) {
        const uint32_t x = (uint32_t)((4 * lane) ^ (t8 << 4)) + t8 * 128;
        stsA[t8] = smemBase + warp * (XNA * 128) + x;
        stsB[t8] = smemBase + 2 * XA_BYTES + warp * (XNB * 128) + x;
    }
    // A element j of this lane is inside the band iff 0 <= lane - j < 16
    uint32_t amask = 0;
#pragma unroll
    for (int j = 0; j < XNA; ++j) amask |= (lane - j >= 0 && lane - j < XTD) ? (1u << j) : 0u;

    if (warp >= XEPI_WARP0) {
        // ================================ epilogue (warps 12-15) ================================================
        // accumulators of a finished item -> grad[b][c][position]; items alternate between the two 256-column TMEM
        // buffers, so the MMAs of item t+1 run while item t is drained and the producers never stop staging
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n" ::);
        const int quarter = warp - XEPI_WARP0;
        const int m = quarter * 32 + lane;
        XCursor c;
        c.start(p, OFF);
        uint32_t t = 0;
        while (c.valid(p)) {
            const uint32_t ab = t & 1u;
            x_mbar_wait(&bar_acc_full[ab], (t >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int gi = c.i0 + (m >> 4), gj = c.j0 + (m & 15);
            const bool pok = gi < H && gj < W;
            const int cbase = c.cb * XN;
            float* dst = gout + ((size_t)c.b * C + cbase) * plane + (size_t)gi * W + gj;
#pragma unroll 1
            for (int q = 0; q < XN / 16; ++q) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * XN + (uint32_t)(q * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int nch = C - (cbase + q * 16);
                if (pok) {
                    if (nch >= 16) {
#pragma unroll
                        for (int x = 0; x < 16; ++x) {
                            *dst = __uint_as_float(r[x]);
                            dst += plane;
                        }
                    } else {
#pragma unroll
                        for (int x = 0; x < 16; ++x)
                            if (x < nch) dst[(size_t)x * plane] = __uint_as_float(r[x]);
                        dst += 16 * plane;
                    }
                } else {
                    dst += 16 * plane;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) x_mbar_arrive(&bar_acc_empty[ab]);
            ++t;
            c.item += gridDim.x;  // next item of this CTA
            c.decode(p, OFF);
        }
    } else if (warp >= XPROD_WARPS) {
        // ================================ MMA issuer (warp 8; warps 9-11 only donate registers) ================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n" ::);
        if (warp == XPROD_WARPS) {
            XCursor c;
            c.start(p, OFF);
            uint32_t k = 0, t = 0;
            while (c.valid(p)) {
                const uint32_t s = k & 1u;
                const uint32_t ab = t & 1u;
                const bool first = c.r == c.rLo, last = c.last();
                if (first) x_mbar_wait(&bar_acc_empty[ab], ((t >> 1) & 1u) ^ 1u);  // this buffer's previous item is drained
                x_mbar_wait(&bar_full[s], (k >> 1) & 1u);                           // all producer warps have staged chunk k
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t aHi = smemBase + s * XSTAGE_BYTES, aLo = aHi + XA_BYTES;
                    const uint32_t bHi = aHi + 2 * XA_BYTES, bLo = bHi + XB_BYTES;
                    const uint32_t dcol = tmem_base + ab * XN;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t ko = ks * 32;  // 8 tf32 = 32 bytes inside the 128-byte row
                        x_mma(dcol, x_desc(aHi + ko), x_desc(bHi + ko), (first && ks == 0) ? 0u : 1u);
                        x_mma(dcol, x_desc(aHi + ko), x_desc(bLo + ko), 1u);
                        x_mma(dcol, x_desc(aLo + ko), x_desc(bHi + ko), 1u);
                    }
                    x_commit(&bar_empty[s]);
                    if (last) x_commit(&bar_acc_full[ab]);
                }
                __syncwarp();
                if (last) ++t;
                ++k;
                c.advance(p, OFF);
            }
        }
    } else {
    // ================================ producers (warps 0-7) ==================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;\n" ::);
    XCursor ld, st;
    ld.start(p, OFF);
    st.start(p, OFF);
    uint32_t k = 0;  // chunks stored

    // chunk at cursor c -> registers: v[0..31] = B (channel 32*warp + j, key column lane), v[32..47] = A (query
    // (warp, j), key column lane)
    auto load = [&](float (&v)[XNB + XNA], const XCursor& c) {
        const int gi = c.i0 - OFF + c.r, gj = c.j0 - OFF + lane;
        const int c0 = c.cb * XN + warp * XNB;
        const bool bok = lane < XQCOLS + XTD - 1 && gj >= 0 && gj < W;
        const int nch = C - c0;
        // byte addresses, advanced one plane per channel (kept as integers so that the compiler does not fall back to
        // element-index arithmetic: 2 instructions per load instead of 4)
        uint64_t ba = (uint64_t)(xsrc + ((size_t)c.b * C + c0) * plane + (size_t)gi * W + gj);
#if D2T_ABLATE_STAGING
#pragma unroll
        for (int j = 0; j < XNB; ++j) v[j] = 0.f;
        (void)ba; (void)bok; (void)nch;
#else
        if (bok && nch >= XNB) {
#pragma unroll
            for (int j = 0; j < XNB; ++j) {
                v[j] = x_ldg_stream(ba);
                ba += planeBytes;
            }
        } else {
#pragma unroll
            for (int j = 0; j < XNB; ++j) {
                v[j] = (bok && j < nch) ? x_ldg_stream(ba) : 0.f;
                ba += planeBytes;
            }
        }
#endif
        const int si = c.r - warp;  // row displacement of query row `warp` for this patch row (warp-uniform)
        uint32_t m = 0;
        if (si >= 0 && si < XTD && c.i0 + warp < H) {
            const int nc = W - c.j0;
            m = nc < XQCOLS ? (amask & ((1u << nc) - 1u)) : amask;
        }
        const float* ap = gsrc + (((size_t)c.b * H + c.i0 + warp) * W + c.j0) * GMAP + si * GROW + lane;
#pragma unroll
        for (int j = 0; j < XNA; ++j) v[XNB + j] = ((m >> j) & 1u) ? __ldg(ap + j * (GMAP - 1)) : 0.f;
    };
    // registers -> (hi, lo) -> stage S
    // aZero bit s: this warp's A rows in stage s are all zero already (its query row had no live displacement for the
    // patch row staged there last) -> a dead chunk need not store them again
    uint32_t aZero = 0;
    auto store = [&](const float (&v)[XNB + XNA], auto S, bool aDead) {
        constexpr uint32_t so = decltype(S)::value * XSTAGE_BYTES;
        constexpr uint32_t sbit = 1u << decltype(S)::value;
        const bool skipA = aDead && (aZero & sbit);
        aZero = aDead ? (aZero | sbit) : (aZero & ~sbit);
        if (!skipA)
#pragma unroll
        for (int j = 0; j < XNA; ++j) {
            const float hi = x_tf32_rn(v[XNB + j]);
            const uint32_t ad = stsA[j & 7] + so + (j >> 3) * 1024;
            x_sts(ad, hi);
            x_sts(ad + XA_BYTES, v[XNB + j] - hi);
        }
#if D2T_ABLATE_STAGING >= 2
        {
            const uint32_t hiBase = smemBase + so + 2 * XA_BYTES, loBase = hiBase + XB_BYTES;
            for (int e = tid * 16; e < XB_BYTES; e += XPROD_WARPS * 32 * 16) {
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(hiBase + e));
                const float4 h = make_float4(x_tf32_rn(q.x), x_tf32_rn(q.y), x_tf32_rn(q.z), x_tf32_rn(q.w));
                if (D2T_ABLATE_STAGING >= 3)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hiBase + e), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(loBase + e), "f"(q.x - h.x), "f"(q.y - h.y), "f"(q.z - h.z),
                             "f"(q.w - h.w)
                             : "memory");
            }
            return;
        }
#endif
#pragma unroll
        for (int j = 0; j < XNB; ++j) {
            const float hi = x_tf32_rn(v[j]);
            const uint32_t ad = stsB[j & 7] + so + (j >> 3) * 1024;
            x_sts(ad, hi);
            x_sts(ad + XB_BYTES, v[j] - hi);
        }
    };
    auto step = [&](float (&v)[XNB + XNA], auto S) {
        constexpr int s = decltype(S)::value;
        x_mbar_wait(&bar_empty[s], ((k >> 1) & 1u) ^ 1u);  // the MMAs that read this stage two chunks ago are done
        {
            const int si = st.r - warp;  // same (warp-uniform) liveness test as in load()
            const bool aDead = !(si >= 0 && si < XTD && st.i0 + warp < H);
            store(v, S, aDead);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) x_mbar_arrive(&bar_full[s]);
        ++k;
        if (ld.valid(p)) {
            load(v, ld);
            ld.advance(p, OFF);
        }
        st.advance(p, OFF);
    };

    // chunk k always uses stage k & 1, so the two register sets are tied to one stage each.  (A third set, loads three
    // chunks ahead, was measured slower: the loads are bound by L1 throughput, not latency.)
    float va[XNB + XNA], vb[XNB + XNA];
    if (ld.valid(p)) { load(va, ld); ld.advance(p, OFF); }
    if (ld.valid(p)) { load(vb, ld); ld.advance(p, OFF); }
    while (st.valid(p)) {
        step(va, XInt<0>{});
        if (!st.valid(p)) break;
        step(vb, XInt<1>{});
    }
    }  // producers

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * XN)));
    }
}

// GT[b, p, si, sj] = gradOut[b, p + (si,sj) - 7, 15 - si, 15 - sj] if that query is inside the image, else 0.
// One CTA per (b, query row xi): the row's W maps are staged in shared memory, then scattered to the <= 16 key rows
// pi = xi + 7 - si they contribute to (64-byte runs).  xi runs over [-7, H + 7] so that out-of-image rows are zeroed.
__global__ void __launch_bounds__(XFLIP_THREADS)
corr_bwd_flip_kernel(const float* __restrict__ go, float* __restrict__ gt, int B, int H, int W) {
    extern __shared__ float fs[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x / (H + 2 * (XD - 1) + 1);
    const int xi = blockIdx.x % (H + 2 * (XD - 1) + 1) - (XD - 1);
    const bool rowIn = xi >= 0 && xi < H;
    if (rowIn) {  // warp per map: 289 contiguous floats -> fs[xj][0..288]
        const float* src = go + ((size_t)b * H + xi) * W * XKK;
        for (int xj = warp; xj < W; xj += XFLIP_THREADS / 32) {
            const float* m = src + (size_t)xj * XKK;
            float* d = fs + xj * XFLIP_PITCH;
#pragma unroll
            for (int u = 0; u < (XKK + 31) / 32; ++u) {
                const int e = lane + 32 * u;
                if (e < XKK) d[e] = __ldg(m + e);
            }
        }
    }
    __syncthreads();
    for (int si = 0; si < XTD; ++si) {
        const int pi = xi + (XD - 1) - si;
        if (pi < 0 || pi >= H) continue;
        float* dst = gt + (((size_t)b * H + pi) * W * XTD + si) * XTD;   // + pj * 256 + sj
        const float* srow = fs + (XTD - 1 - si) * XK1 + (XTD - 1);
        for (int e = threadIdx.x; e < W * XTD; e += XFLIP_THREADS) {
            const int pj = e >> 4, sj = e & 15;
            const int xj = pj + sj - (XD - 1);
            float v = 0.f;
            if (rowIn && xj >= 0 && xj < W) v = srow[xj * XFLIP_PITCH - sj];
            dst[(size_t)pj * (XTD * XTD) + sj] = v;
        }
    }
}

}  // namespace

bool corr_umma_bwd_supported(int B, int C, int H, int W, int d, int stride) {
    if (stride != 1 || d != XD) return false;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return false;
    if ((long long)B * C * H * W >= (1ll << 31)) return false;
    if ((long long)B * H * W * XKK >= (1ll << 31)) return false;
    if ((size_t)W * XFLIP_PITCH * sizeof(float) > 200 * 1024) return false;  // flip kernel stages one gradOut row
    return true;
}

size_t corr_umma_bwd_ws_bytes(int B, int C, int H, int W) {
    (void)C;
    return (size_t)B * H * W * XTD * XTD * sizeof(float);
}

int corr_umma_bwd_launch(const float* go, const float* fm0, const float* fm1, float* g0, float* g1, int B, int C, int H,
                         int W, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = corr_umma_bwd_ws_bytes(B, C, H, W);
    if (ws == nullptr || ws_bytes < need) {
        set_error("corr_bwd(umma): workspace too small (%zu < %zu)", ws_bytes, need);
        return D2T_ERR_WORKSPACE;
    }
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    float* gt = static_cast<float*>(ws);

    XPlan p;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.tilesX = ceil_div(W, XQCOLS);
    p.tilesY = ceil_div(H, XQROWS);
    p.nCb = ceil_div(C, XN);
    p.nItems = B * p.nCb * p.tilesX * p.tilesY;
    const int grid = p.nItems < di.sm_count ? p.nItems : di.sm_count;
    const size_t smem = (size_t)XSTAGES * XSTAGE_BYTES + 1024;
    D2T_SMEM_OPTIN(corr_bwd_umma_kernel<0>, smem);
    D2T_SMEM_OPTIN(corr_bwd_umma_kernel<1>, smem);

    const size_t fsmem = (size_t)W * XFLIP_PITCH * sizeof(float);
    D2T_SMEM_OPTIN(corr_bwd_flip_kernel, fsmem);
    corr_bwd_flip_kernel<<<B * (H + 2 * (XD - 1) + 1), XFLIP_THREADS, fsmem, st>>>(go, gt, B, H, W);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();

    corr_bwd_umma_kernel<0><<<grid, XTHREADS, smem, st>>>(go, fm1, g0, p);   // grad_FM0: G = gradOut, X = FM1
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    {   // grad_FM1: G = flipped gradOut, X = FM0; may start while grad_FM0's last CTAs are still running (see the kernel)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(XTHREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const float* gtc = gt;
        D2T_CUDA_TRY(cudaLaunchKernelEx(&cfg, corr_bwd_umma_kernel<1>, gtc, fm0, g1, p));
    }
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
