// corr_umma.cu -- PointwiseCorrelation forward on the 5th-generation tensor cores (tcgen05 + TMEM), d_max = 8.
//
// Formulation: one tile of 8x16 query positions (M = 128) against 8 rows of its 23x31 key halo patch (N = 8 x 32 = 256
// key positions, 31 used per row) is a GEMM over channels, D[m][n] = sum_c Q[c][m] * K[c][n]; the correlation map of
// query (qrow, qcol) is the band {n = (kr - kr0)*32 + qcol + tj : kr = qrow + ti} of row m.  A tile is three work items
// (key rows 0-7, 8-15, 16-22), each over all channels, writing disjoint displacement rows of the tile's maps.  The dense
// tile does ~2.9x the band's MACs and FP32 inputs are split 3xTF32 (hi*hi + hi*lo + lo*hi, hi rounded to nearest,
// measured |err| <= 9e-7 * sum|a||b|, tools/umma_sw128_test.cu).
//
//   operands : K-major SWIZZLE_128B (row = position, 32 channels = 128 bytes).  The maps are NCHW, so the staging
//              transposes: a producer thread owns ONE position (lane = position inside a 32-wide row group: coalesced
//              LDG per channel), collects 32 channels in registers, splits hi/lo and writes its whole 128-byte operand
//              row with 8 + 8 STS.128 (the XOR swizzle makes the 8 lanes of a quarter-warp hit 8 bank groups).
//   warps    : 0-3 stage the queries (A, 128 rows), 4-11 stage one key row each (B, 256 rows); loads run two chunks
//              ahead in registers.  Warp 12 issues the MMAs (4 k-steps x 3 per 32-channel chunk) and commits to the
//              stage's `empty` / the item's `accum_full` barrier; warps 13-15 only donate their registers (setmaxnreg).
//              Why a separate MMA warp: see corr_umma_bwd.cu.
//   epilogue : tcgen05.ld (32 lanes x 32 columns = two query rows x one key row), band extraction into a per-query-row
//              buffer in the final (17x17 per position) layout, coalesced copy-out; dead row / column 16 written as
//              zeros.  Out-of-image keys are staged as zeros, so dead border entries come out as exact 0.
#include <stdlib.h>
#include <string.h>

#include "corr_common.cuh"

namespace d2t {

namespace {

constexpr int UD = 8;
constexpr int UM = 128;                   // queries per tile = UMMA M
constexpr int UN = 256;                   // key positions per item = UMMA N = TMEM columns
constexpr int UCH = 32;                   // channels per chunk (one 128-byte operand row)
constexpr int UA_WARPS = 4, UB_WARPS = 8;
constexpr int UPROD_WARPS = UA_WARPS + UB_WARPS;
constexpr int UTHREADS = (UPROD_WARPS + 4) * 32;
constexpr int UPROD_THREADS = UPROD_WARPS * 32;
constexpr int UA_BYTES = UM * 128;
constexpr int UB_BYTES = UN * 128;
constexpr int USTAGE_BYTES = 2 * UA_BYTES + 2 * UB_BYTES;
constexpr int USTAGES = 2;
constexpr int URP = 290;                  // row-buffer pitch per query (289 used)
constexpr int UROWBUF_FLOATS = 16 * URP;
constexpr int UGROUPS = 3;                // key-row groups per tile: rows [0,8) [8,16) [16,23)

struct UPlan {
    int B, C, H, W;
    int tilesX, tilesY, nItems, nChunks;
    // Work units of CTA `cta`: the whole items  round * G + cta  (round < rounds), then -- so that the last, partly
    // filled wave does not leave most SMs idle -- one channel range (1 / split of the chunks) of a left-over item:
    // unit  cta < left * split  ->  item rounds * G + cta / split, chunks [q, q + 1) * nChunks / split, q = cta % split,
    // written to partial slot `cta` and summed in fixed order by corr_fwd_umma_finalize_kernel.
    int G, rounds, left, split;
};

__device__ __forceinline__ uint32_t u_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void u_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(u_smem(bar)), "r"(count));
}
__device__ __forceinline__ void u_mbar_wait(uint64_t* bar, uint32_t parity) {  // suspend-time hint: see corr_umma_bwd.cu
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            u_smem(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void u_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(u_smem(bar)) : "memory");
}
__device__ __forceinline__ uint64_t u_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, 8-row atoms 1024 bytes apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
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
__device__ __forceinline__ void u_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_smem(bar)) : "memory");
}
__device__ __forceinline__ float u_ldg_stream(uint64_t addr) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(addr));
    return v;
}
__device__ __forceinline__ void u_sts4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float u_tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void u_prod_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(UPROD_THREADS) : "memory"); }

template <int V>
struct UInt { static constexpr int value = V; };

// walks the (item, channel chunk) sequence of one CTA; item = (image, tile, key-row group)
struct UCursor {
    int rnd, item;          // rnd <= rounds: rnd == rounds is the split unit; item < 0: no more work
    int chunk, chunkBeg, chunkEnd;
    int b, i0, j0, kr0, nRows;
    bool partial;
    __device__ __forceinline__ bool valid(const UPlan&) const { return item >= 0; }
    __device__ __forceinline__ void decode(const UPlan& p) {
        const int cta = blockIdx.x;
        if (rnd < p.rounds) {
            item = rnd * p.G + cta;
            chunkBeg = 0;
            chunkEnd = p.nChunks;
            partial = false;
        } else if (rnd == p.rounds && cta < p.left * p.split) {
            item = p.rounds * p.G + cta / p.split;
            const int q = cta % p.split;
            chunkBeg = (int)((long long)q * p.nChunks / p.split);
            chunkEnd = (int)((long long)(q + 1) * p.nChunks / p.split);
            partial = p.split > 1;
        } else {
            item = -1;
            return;
        }
        const int g = item % UGROUPS;
        const int tile = item / UGROUPS;
        const int tpi = p.tilesX * p.tilesY;
        b = tile / tpi;
        const int t = tile - b * tpi;
        i0 = (t / p.tilesX) * 8;
        j0 = (t % p.tilesX) * 16;
        kr0 = 8 * g;
        nRows = g == UGROUPS - 1 ? 7 : 8;
        chunk = chunkBeg;
    }
    __device__ __forceinline__ void start(const UPlan& p) { rnd = 0; decode(p); }
    __device__ __forceinline__ bool first() const { return chunk == chunkBeg; }
    __device__ __forceinline__ bool last(const UPlan&) const { return chunk == chunkEnd - 1; }
    __device__ __forceinline__ void advance(const UPlan& p) {
        if (++chunk >= chunkEnd) { ++rnd; decode(p); }
    }
};

__global__ void __launch_bounds__(UTHREADS, 1)
corr_fwd_umma_kernel(const float* __restrict__ fm0, const float* __restrict__ fm1, float* __restrict__ out,
                     float* __restrict__ partial, UPlan p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* rowbuf = reinterpret_cast<float*>(smem + USTAGES * USTAGE_BYTES);
    __shared__ __align__(8) uint64_t bar_full[USTAGES], bar_empty[USTAGES], bar_acc_full, bar_acc_empty;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, C = p.C;
    const size_t plane = (size_t)H * W;
    const uint64_t planeBytes = (uint64_t)plane * sizeof(float);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(u_smem(&tmem_base_s)), "r"((uint32_t)UN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < USTAGES; ++s) {
            u_mbar_init(&bar_full[s], UPROD_WARPS);
            u_mbar_init(&bar_empty[s], 1);
        }
        u_mbar_init(&bar_acc_full, 1);
        u_mbar_init(&bar_acc_empty, UPROD_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smemBase = u_smem(smem);

    if (warp >= UPROD_WARPS) {
        // ================================ MMA issuer (warp 12; warps 13-15 only donate registers) ===============
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n" ::);
        if (warp == UPROD_WARPS) {
            UCursor c;
            c.start(p);
            uint32_t k = 0, t = 0;
            while (c.valid(p)) {
                const uint32_t s = k & 1u;
                const bool first = c.first(), last = c.last(p);
                const uint32_t idesc = u_idesc(c.nRows * 32);
                if (first) u_mbar_wait(&bar_acc_empty, (t & 1u) ^ 1u);  // previous item's accumulators are drained
                u_mbar_wait(&bar_full[s], (k >> 1) & 1u);               // all producer warps have staged chunk k
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t aHi = smemBase + s * USTAGE_BYTES, aLo = aHi + UA_BYTES;
                    const uint32_t bHi = aHi + 2 * UA_BYTES, bLo = bHi + UB_BYTES;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t ko = ks * 32;  // 8 tf32 = 32 bytes inside the 128-byte row
                        u_mma(tmem_base, u_desc(aHi + ko), u_desc(bHi + ko), idesc, (first && ks == 0) ? 0u : 1u);
                        u_mma(tmem_base, u_desc(aHi + ko), u_desc(bLo + ko), idesc, 1u);
                        u_mma(tmem_base, u_desc(aLo + ko), u_desc(bHi + ko), idesc, 1u);
                    }
                    u_commit(&bar_empty[s]);
                    if (last) u_commit(&bar_acc_full);
                }
                __syncwarp();
                if (last) ++t;
                ++k;
                c.advance(p);
            }
        }
    } else {
        // ================================ producers / epilogue (warps 0-11) ====================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;\n" ::);
        const bool isA = warp < UA_WARPS;
        const int rowInOp = (isA ? warp : warp - UA_WARPS) * 32 + lane;  // operand row staged by this thread
        // byte address (stage 0, hi part) of this thread's operand row, and its eight swizzled 16-byte chunk offsets
        const uint32_t rowAddr = smemBase + (isA ? 0u : 2u * UA_BYTES) + (uint32_t)((rowInOp >> 3) * 1024 + (rowInOp & 7) * 128);
        const uint32_t loOff = isA ? UA_BYTES : UB_BYTES;
        const uint32_t sw = (uint32_t)(lane & 7) << 4;

        UCursor ld, st;
        ld.start(p);
        st.start(p);
        uint32_t k = 0, t = 0;

        // chunk at cursor c -> registers: 32 channels of this thread's position
        auto load = [&](float (&v)[UCH], const UCursor& c) {
            int gi, gj;
            bool ok;
            const float* img;
            if (isA) {
                const int m = warp * 32 + lane;
                gi = c.i0 + (m >> 4); gj = c.j0 + (m & 15);
                ok = gi < H && gj < W;
                img = fm0;
            } else {
                const int kk = warp - UA_WARPS;
                gi = c.i0 - UD + c.kr0 + kk; gj = c.j0 - UD + lane;
                ok = kk < c.nRows && lane < 31 && gi >= 0 && gi < H && gj >= 0 && gj < W;
                img = fm1;
            }
            const int c0 = c.chunk * UCH;
            const int nch = C - c0;
            uint64_t a = (uint64_t)(img + ((size_t)c.b * C + c0) * plane + (ok ? (size_t)gi * W + gj : 0));
            if (ok && nch >= UCH) {
#pragma unroll
                for (int j = 0; j < UCH; ++j) {
                    v[j] = u_ldg_stream(a);
                    a += planeBytes;
                }
            } else {
#pragma unroll
                for (int j = 0; j < UCH; ++j) {
                    v[j] = (ok && j < nch) ? u_ldg_stream(a) : 0.f;
                    a += planeBytes;
                }
            }
        };
        auto store = [&](const float (&v)[UCH], auto S) {
            constexpr uint32_t so = decltype(S)::value * USTAGE_BYTES;
#pragma unroll
            for (int j = 0; j < UCH / 4; ++j) {
                const float h0 = u_tf32_rn(v[4 * j]), h1 = u_tf32_rn(v[4 * j + 1]);
                const float h2 = u_tf32_rn(v[4 * j + 2]), h3 = u_tf32_rn(v[4 * j + 3]);
                const uint32_t ad = rowAddr + so + (((uint32_t)j << 4) ^ sw);
                u_sts4(ad, h0, h1, h2, h3);
                u_sts4(ad + loOff, v[4 * j] - h0, v[4 * j + 1] - h1, v[4 * j + 2] - h2, v[4 * j + 3] - h3);
            }
        };
        // accumulators of the finished item -> out: band extraction through the row buffer
        auto epilogue = [&](const UCursor& c) {
            u_mbar_wait(&bar_acc_full, t & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int quarter = warp & 3, third = warp >> 2;  // TMEM lanes 32*quarter..+31; key rows kk % 3 == third
            const int m = quarter * 32 + lane;
            const int qrow = m >> 4, qcol = m & 15;
            const bool lastGroup = c.kr0 + c.nRows == 23;
            const int ptid = tid;  // 0 .. UPROD_THREADS-1
            for (int qr = 0; qr < 8; ++qr) {
                // ti range of query row qr produced by this item (row 16 of every map is dead: zeros, with the last group)
                const int tiLo = max(0, c.kr0 - qr), tiHi = min(15, c.kr0 + c.nRows - 1 - qr);
                const int tiEnd = lastGroup ? 17 : tiHi + 1;  // exclusive, in 17-float rows
                if ((qr >> 1) == quarter) {
                    for (int kr = max(c.kr0, qr); kr <= min(c.kr0 + c.nRows - 1, qr + 15); ++kr) {
                        if ((kr - c.kr0) % 3 != third) continue;  // warp-uniform: the three warps of a quarter split the rows
                        uint32_t r[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((kr - c.kr0) * 32);
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
                    if (lastGroup && qrow == qr && third == 0) {
#pragma unroll
                        for (int x = 0; x < 17; ++x) rowbuf[qcol * URP + 16 * 17 + x] = 0.f;  // dead row 16
                    }
                }
                u_prod_barrier();
                {   // copy rows [tiLo, tiEnd) of the 16 queries of this tile row to their maps (24 threads per query)
                    const int len = (tiEnd - tiLo) * 17;
                    const int gi = c.i0 + qr;
                    const int ncols = c.partial ? 16 : min(16, W - c.j0);
                    if (len > 0 && (c.partial || gi < H)) {
                        float* dbase = c.partial ? partial + ((size_t)blockIdx.x * UM + qr * 16) * 289
                                                 : out + (((size_t)c.b * H + gi) * W + c.j0) * 289;
                        const int q = ptid / 24;
                        if (q < ncols)
                            for (int o = ptid - q * 24; o < len; o += 24) dbase[q * 289 + tiLo * 17 + o] = rowbuf[q * URP + tiLo * 17 + o];
                    }
                }
                u_prod_barrier();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) u_mbar_arrive(&bar_acc_empty);
            ++t;
        };
        auto step = [&](float (&v)[UCH], auto S) {
            constexpr int s = decltype(S)::value;
            u_mbar_wait(&bar_empty[s], ((k >> 1) & 1u) ^ 1u);  // the MMAs that read this stage two chunks ago are done
            store(v, S);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) u_mbar_arrive(&bar_full[s]);
            ++k;
            if (ld.valid(p)) {
                load(v, ld);
                ld.advance(p);
            }
            if (st.last(p)) epilogue(st);
            st.advance(p);
        };

        float va[UCH], vb[UCH];
        if (ld.valid(p)) { load(va, ld); ld.advance(p); }
        if (ld.valid(p)) { load(vb, ld); ld.advance(p); }
        while (st.valid(p)) {
            step(va, UInt<0>{});
            if (!st.valid(p)) break;
            step(vb, UInt<1>{});
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)UN));
    }
}

// out <- sum over the `split` channel ranges of every left-over item (ascending range = ascending channel order)
__global__ void __launch_bounds__(256)
corr_fwd_umma_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, UPlan p) {
    const int lt = blockIdx.x, qr = blockIdx.y;  // left-over item, query row of its tile
    const int item = p.rounds * p.G + lt;
    const int g = item % UGROUPS, tile = item / UGROUPS;
    const int tpi = p.tilesX * p.tilesY;
    const int b = tile / tpi, t = tile - b * tpi;
    const int i0 = (t / p.tilesX) * 8, j0 = (t % p.tilesX) * 16;
    const int kr0 = 8 * g, nRows = g == UGROUPS - 1 ? 7 : 8;
    const int gi = i0 + qr;
    if (gi >= p.H) return;
    const int tiLo = max(0, kr0 - qr), tiHi = min(15, kr0 + nRows - 1 - qr);
    const int tiEnd = (kr0 + nRows == 23) ? 17 : tiHi + 1;
    const int len = (tiEnd - tiLo) * 17;
    if (len <= 0) return;
    const int ncols = min(16, p.W - j0);
    float* dbase = out + (((size_t)b * p.H + gi) * p.W + j0) * 289 + tiLo * 17;
    const float* src = partial + ((size_t)(lt * p.split) * UM + qr * 16) * 289 + tiLo * 17;
    for (int e = threadIdx.x; e < ncols * len; e += blockDim.x) {
        const int q = e / len, o = e - q * len;
        float acc = 0.f;
        for (int k = 0; k < p.split; ++k) acc += src[((size_t)k * UM + q) * 289 + o];
        dbase[q * 289 + o] = acc;
    }
}

}  // namespace

bool corr_umma_supported(int B, int C, int H, int W, int d, int stride) {
    if (stride != 1 || d != UD) return false;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return false;
    if ((long long)B * C * H * W >= (1ll << 31)) return false;
    return true;
}

static int umma_fwd_plan(int B, int C, int H, int W, UPlan* p) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    p->B = B; p->C = C; p->H = H; p->W = W;
    p->tilesX = ceil_div(W, 16);
    p->tilesY = ceil_div(H, 8);
    p->nItems = B * p->tilesX * p->tilesY * UGROUPS;
    p->nChunks = ceil_div(C, UCH);
    p->G = p->nItems < di.sm_count ? p->nItems : di.sm_count;
    p->rounds = p->nItems / p->G;
    p->left = p->nItems - p->rounds * p->G;
    p->split = 1;
    if (p->left > 0) {
        p->split = p->G / p->left;
        if (p->split > p->nChunks) p->split = p->nChunks;
        if (p->split > 8) p->split = 8;
        if (p->split < 1) p->split = 1;
    }
    return 0;
}

size_t corr_umma_fwd_ws_bytes(int B, int C, int H, int W) {
    UPlan p;
    if (umma_fwd_plan(B, C, H, W, &p)) return 0;
    return p.split > 1 ? (size_t)p.left * p.split * UM * 289 * sizeof(float) : 0;
}

int corr_umma_fwd_launch(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
    UPlan p;
    int rc = umma_fwd_plan(B, C, H, W, &p);
    if (rc) return rc;
    const size_t need = p.split > 1 ? (size_t)p.left * p.split * UM * 289 * sizeof(float) : 0;
    if (need > 0 && (ws == nullptr || ws_bytes < need)) {
        set_error("corr_fwd(umma): workspace too small (%zu < %zu)", ws_bytes, need);
        return D2T_ERR_WORKSPACE;
    }
    const size_t smem = (size_t)USTAGES * USTAGE_BYTES + UROWBUF_FLOATS * sizeof(float) + 1024;
    D2T_CUDA_TRY(cudaFuncSetAttribute(corr_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    corr_fwd_umma_kernel<<<p.G, UTHREADS, smem, st>>>(fm0, fm1, out, static_cast<float*>(ws), p);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    if (p.split > 1) {
        corr_fwd_umma_finalize_kernel<<<dim3(p.left, 8), 256, 0, st>>>(static_cast<const float*>(ws), out, p);
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
    }
    return D2T_OK;
}

}  // namespace d2t
