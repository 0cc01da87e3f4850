// corr_umma_fwd.cu -- PointwiseCorrelation FORWARD on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
// d_max = 8, stride 1, float32.
//
//   out[b, i, j, ci, cj] = sum_c fm0[b, c, i, j] * fm1[b, c, i - 8 + ci, j - 8 + cj]      (ci, cj < 16)
//
// is, for a tile of 8 x 16 query positions, the product  D[m = query][n = key] = sum_c Q[c][m] * K[c][n]  of the tile's
// queries with its 23 x 31 halo patch of keys, of which the 16 x 16 band of every query is kept.  The contraction runs
// over CHANNELS while NCHW memory is contiguous over POSITIONS, i.e. both operands are "MN-major" exactly as the maps lie
// in memory.  Round 1 believed kind::tf32 could not take MN-major operands (its test produced zeros) and built a forward
// that transposed both operands while staging; it lost to the FP32-pipe kernel and was dropped.  The zeros were a
// descriptor error: 32-bit operands need the SWIZZLE_128B_BASE32B layout (type 1: atoms of 32 positions x 4 k-rows of
// 128 bytes, 32-byte chunks XOR-swizzled with the row; cute Layout_MN_SW128_32B_Atom) -- tools/umma_tf32_mnmajor_test.cu,
// profiles/r2_umma_tf32_mnmajor_test.txt.  With it a patch row of one channel (32 consecutive floats of an NCHW plane) is
// ONE operand row: a warp loads it with one coalesced LDG and stores it with one conflict-free STS per plane.
//
//   work item    (image, tile, chunk g): chunk g = patch rows 8g .. 8g+7 (N = 8 rows x 32 columns = 256 TMEM columns);
//                chunks that lie wholly outside the image are not items (52 instead of 60 per 38 x 63 image), so a
//                batch of 8 images is 416 items = 2.81 per SM.  M = 128 queries.  K = all channels.
//   precision    3xTF32: hi = tf32_rn(v), lo = v - hi; hi*hi + hi*lo + lo*hi per k-step (as the backward kernel).
//   stage        32 channels: A = Q[32][128] and B = K[32][256], hi and lo each: 96 KB; 2 stages.  4 k-steps x 3 MMAs of
//                128 x 256 x 8 per stage.
//   warps        0-7 producers (lane = position inside a 32-position row; warp w stages patch row w of B for the 32
//                channels and 16 channels of one A atom; register double buffering two stages ahead), warp 8 MMA issuer,
//                warps 12-15 epilogue (TMEM -> band -> global) on the other of two 256-column accumulators, so the
//                epilogue of item t overlaps the main loop of item t + 1.  setmaxnreg moves registers to the producers.
//   epilogue     thread m owns query m: for each of the chunk's 8 patch rows it reads 32 accumulator columns and keeps
//                the 16 that start at its query column.  The first valid chunk of a tile also writes the structural zeros
//                (row / column 2d of every map, and the entries of chunks that are not items).  Strided output supported
//                (tracker glue fusion).
//   determinism  fixed channel order, no partial sums, no atomics: bitwise reproducible.
#include "corr_common.cuh"

// Timing ablation only (tools/experiments/staging_ablation.sh; NEVER defined by csrc/Makefile -- results are garbage):
//   1  the producers skip the global loads of both operands (what the LDGs cost)
//   2  ... and replace their scalar hi/lo stores by what the split warps of a TMA-fed kernel would execute if the raw tile
//      had landed in shared memory by itself: LDS.128 raw -> lo = v - trunc_tf32(v) -> STS.128 lo  (raw doubles as hi)
//   3  as 2, with the rounded hi written back in place as well (the gemm_tf32x3 recipe)
#ifndef D2T_ABLATE_STAGING
#define D2T_ABLATE_STAGING 0
#endif

namespace d2t {

namespace {

constexpr int FD = 8, FTD = 16, FK1 = 17, FKK = 289;
constexpr int FM = 128;
constexpr int FQROWS = 8, FQCOLS = 16;
constexpr int FKC = 32;                       // channels per stage
constexpr int FPROD_WARPS = 8;
constexpr int FEPI_WARP0 = FPROD_WARPS + 4;
constexpr int FTHREADS = (FPROD_WARPS + 8) * 32;
constexpr int FA_BYTES = FKC * FM * 4;        // one plane (hi or lo) of A per stage: 16 KB
constexpr int FSTAGES = 2;
constexpr int FLBO = (FKC / 4) * 512;         // bytes between MN atoms (32 positions) of a plane
constexpr int FSBO = 512;                     // bytes between K atoms (4 channels)
constexpr int FNA = FKC / 2;                  // A elements a producer thread stages per stage

// CR = patch rows per chunk: 8 (N = 256) is the efficient shape; 4 (N = 128) doubles the number of items for problems that
// would otherwise leave most SMs idle (one image: 52 items of 8 rows on 148 SMs).  The channel order of every output is
// the same in both, so the choice does not change a single bit of the result.
template <int CR>
struct FCfg {
    static constexpr int N = CR * 32;                       // accumulator columns per item
    static constexpr int NCH = 24 / CR;                     // chunks per tile (patch rows 0 .. 23)
    static constexpr int B_BYTES = FKC * N * 4;             // one plane of B per stage
    static constexpr int STAGE_BYTES = 2 * FA_BYTES + 2 * B_BYTES;
    static constexpr int NB = FKC * CR / 8;                 // B elements a producer thread stages per stage
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                                      ((uint32_t)(FM >> 4) << 24);   // F32 acc, TF32 a/b, MN-major a/b, N, M
};

struct FPlan {
    int B, C, H, W;
    int tilesX, tilesY, itemsPerImage, nItems, nChunks;
    CorrOutStrides os;
};

__device__ __forceinline__ uint32_t f_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem(bar)), "r"(count));
}
__device__ __forceinline__ void f_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            f_smem(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void f_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(f_smem(bar)) : "memory");
}
// MN-major SWIZZLE_128B_BASE32B shared-memory matrix descriptor (layout type 1), version 1
__device__ __forceinline__ uint64_t f_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((FLBO >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((FSBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ void f_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void f_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(f_smem(bar)) : "memory");
}
__device__ __forceinline__ void f_sts(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float f_ldg_stream(uint64_t addr) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(addr));
    return v;
}
__device__ __forceinline__ float f_tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

template <int V>
struct FInt { static constexpr int value = V; };

// chunk g of the tile whose first query row is i0 holds patch rows i0 - 8 + CR*g .. + CR-1: an item iff some row is in the image
template <int CR>
__host__ __device__ __forceinline__ bool f_chunk_valid(int i0, int g, int H) {
    const int r0 = i0 - FD + CR * g;
    return r0 + CR - 1 >= 0 && r0 < H;
}

// walks the (item, channel chunk) sequence of one CTA
template <int CR>
struct FCursor {
    static constexpr int NCH = FCfg<CR>::NCH;
    int item, ch;            // ch: channel chunk inside the item
    int b, i0, j0, g, gFirst;
    unsigned gmask;          // bit g set = chunk g of this tile is an item
    __device__ __forceinline__ bool valid(const FPlan& p) const { return item < p.nItems; }
    __device__ __forceinline__ void decode(const FPlan& p) {
        if (item >= p.nItems) return;
        b = item / p.itemsPerImage;
        int rem = item - b * p.itemsPerImage;
        // items of an image: for ty, for tx, for valid g
        int ty = 0;
        for (; ty < p.tilesY; ++ty) {
            const int ii = ty * FQROWS;
            int ng = 0;
            for (int gg = 0; gg < NCH; ++gg) ng += (int)f_chunk_valid<CR>(ii, gg, p.H);
            if (rem < ng * p.tilesX) {
                const int tx = rem / ng;
                int k = rem - tx * ng;
                i0 = ii;
                j0 = tx * FQCOLS;
                gmask = 0;
                gFirst = -1;
                g = 0;
                for (int gg = 0; gg < NCH; ++gg) {
                    if (!f_chunk_valid<CR>(ii, gg, p.H)) continue;
                    gmask |= 1u << gg;
                    if (gFirst < 0) gFirst = gg;
                    if (k == 0) g = gg;
                    --k;
                }
                break;
            }
            rem -= ng * p.tilesX;
        }
        ch = 0;
    }
    __device__ __forceinline__ void start(const FPlan& p) { item = blockIdx.x; decode(p); }
    __device__ __forceinline__ bool last(const FPlan& p) const { return ch == p.nChunks - 1; }
    __device__ __forceinline__ void advance(const FPlan& p) {
        if (++ch >= p.nChunks) { item += gridDim.x; decode(p); }
    }
};

template <int CR>
__global__ void __launch_bounds__(FTHREADS, 1)
corr_fwd_umma_kernel(const float* __restrict__ fm0, const float* __restrict__ fm1, float* __restrict__ out, FPlan p) {
    using Cfg = FCfg<CR>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[FSTAGES], bar_empty[FSTAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, C = p.C;
    const size_t plane = (size_t)H * W;
    const uint64_t planeBytes = (uint64_t)plane * sizeof(float);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(f_smem(&tmem_base_s)), "r"((uint32_t)(2 * Cfg::N)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < FSTAGES; ++s) {
            f_mbar_init(&bar_full[s], FPROD_WARPS);
            f_mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            f_mbar_init(&bar_acc_full[a], 1);
            f_mbar_init(&bar_acc_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smemBase = f_smem(smem);

    if (warp >= FEPI_WARP0) {
        // ================================ epilogue (warps 12-15) ================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n" ::);
        const int quarter = warp - FEPI_WARP0;
        const int m = quarter * 32 + lane;
        const int qrow = m >> 4, qcol = m & 15;
        FCursor<CR> c;
        c.start(p);
        uint32_t t = 0;
        while (c.valid(p)) {
            const uint32_t ab = t & 1u;
            const int gi = c.i0 + qrow, gj = c.j0 + qcol;
            const bool pok = gi < H && gj < W;
            float* o = out + (long long)c.b * p.os.sb + ((long long)gi * W + gj) * p.os.sp;
            if (c.g == c.gFirst && pok) {
                // structural zeros of this position, written once per tile: row / column 2d of the map and
                // the entries whose key row belongs to a chunk that is not an item (wholly outside the image)
                for (int e = 0; e < FKK; ++e) {
                    const int ci = e / FK1, cj = e - ci * FK1;
                    const bool produced = ci < FTD && cj < FTD && ((c.gmask >> ((qrow + ci) / CR)) & 1u);
                    if (!produced) o[(long long)e * p.os.st] = 0.f;
                }
            }
            f_mbar_wait(&bar_acc_full[ab], (t >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int prl = 0; prl < CR; ++prl) {
                const int ci = CR * c.g + prl - qrow;     // row displacement of this patch row for this query
                // the loads are warp-collective: every lane executes them, the stores are predicated
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * Cfg::N + (uint32_t)(prl * 32 + hf * 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                        "%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (pok && ci >= 0 && ci < FTD) {
                        float* orow = o + (long long)(ci * FK1) * p.os.st;
#pragma unroll
                        for (int x = 0; x < 16; ++x) {
                            const int cj = hf * 16 + x - qcol;   // patch column hf*16 + x is key column j0 - 8 + ..., i.e. cj = column - qcol
                            if (cj >= 0 && cj < FTD) orow[(long long)cj * p.os.st] = __uint_as_float(r[x]);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) f_mbar_arrive(&bar_acc_empty[ab]);
            ++t;
            c.item += gridDim.x;
            c.decode(p);
        }
    } else if (warp >= FPROD_WARPS) {
        // ================================ MMA issuer (warp 8; warps 9-11 only donate registers) ================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n" ::);
        if (warp == FPROD_WARPS) {
            FCursor<CR> c;
            c.start(p);
            uint32_t k = 0, t = 0;
            while (c.valid(p)) {
                const uint32_t s = k & 1u;
                const uint32_t ab = t & 1u;
                const bool first = c.ch == 0, last = c.last(p);
                if (first) f_mbar_wait(&bar_acc_empty[ab], ((t >> 1) & 1u) ^ 1u);   // this buffer's previous item is drained
                f_mbar_wait(&bar_full[s], (k >> 1) & 1u);                            // all producer warps have staged the chunk
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t aHi = smemBase + s * Cfg::STAGE_BYTES, aLo = aHi + FA_BYTES;
                    const uint32_t bHi = aHi + 2 * FA_BYTES, bLo = bHi + Cfg::B_BYTES;
                    const uint32_t dcol = tmem_base + ab * Cfg::N;
#pragma unroll
                    for (int ks = 0; ks < FKC / 8; ++ks) {
                        const uint32_t ko = ks * 2 * FSBO;   // 8 channels = two K atoms
                        f_mma(dcol, f_desc(aHi + ko), f_desc(bHi + ko), Cfg::IDESC, (first && ks == 0) ? 0u : 1u);
                        f_mma(dcol, f_desc(aHi + ko), f_desc(bLo + ko), Cfg::IDESC, 1u);
                        f_mma(dcol, f_desc(aLo + ko), f_desc(bHi + ko), Cfg::IDESC, 1u);
                    }
                    f_commit(&bar_empty[s]);
                    if (last) f_commit(&bar_acc_full[ab]);
                }
                __syncwarp();
                if (last) ++t;
                ++k;
                c.advance(p);
            }
        }
    } else {
        // ================================ producers (warps 0-7) ==================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;\n" ::);
        // element (channel k of the stage, position mn) of a plane lives at
        //   ((mn >> 5) * 8 + (k >> 2)) * 512 + (k & 3) * 128 + ((((mn & 31) >> 3) ^ (k & 3)) * 32) + (mn & 7) * 4
        // B: this warp stages patch row brow (= MN atom brow) for channels bch0 .. bch0 + NB - 1, lane = column;
        // A: atom warp >> 1, channels (warp & 1) * 16 ...
        const int brow = warp % CR, bch0 = (warp / CR) * Cfg::NB;
        uint32_t stsA[4], stsB[4];
#pragma unroll
        for (int kr = 0; kr < 4; ++kr) {
            const uint32_t x = (uint32_t)(kr * 128 + (((lane >> 3) ^ kr) * 32) + (lane & 7) * 4);
            stsB[kr] = smemBase + 2 * FA_BYTES + brow * FLBO + (bch0 / 4) * FSBO + x;
            stsA[kr] = smemBase + (warp >> 1) * FLBO + (warp & 1) * (FNA / 4) * FSBO + x;
        }
        const int aqrow = 2 * (warp >> 1) + (lane >> 4), aqcol = lane & 15;   // query staged by this lane for A

        FCursor<CR> ld, st;
        ld.start(p);
        st.start(p);
        uint32_t k = 0;   // chunks stored

        auto load = [&](float (&v)[Cfg::NB + FNA], const FCursor<CR>& c) {
#if D2T_ABLATE_STAGING
#pragma unroll
            for (int j = 0; j < Cfg::NB + FNA; ++j) v[j] = 0.f;
            return;
#endif
            const int c0 = c.ch * FKC;
            {   // B: patch row (CR*g + brow), column lane (column 31 is padding), channels c0 + bch0 + j
                const int gi = c.i0 - FD + CR * c.g + brow, gj = c.j0 - FD + lane;
                const bool ok = lane < FQCOLS + FTD - 1 && gi >= 0 && gi < H && gj >= 0 && gj < W;
                const int nch = C - (c0 + bch0);
                uint64_t ba = (uint64_t)(fm1 + ((size_t)c.b * C + (nch > 0 ? c0 + bch0 : 0)) * plane + (size_t)(ok ? gi * W + gj : 0));
                if (ok && nch >= Cfg::NB) {
#pragma unroll
                    for (int j = 0; j < Cfg::NB; ++j) {
                        v[j] = f_ldg_stream(ba);
                        ba += planeBytes;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < Cfg::NB; ++j) {
                        v[j] = (ok && j < nch) ? f_ldg_stream(ba) : 0.f;
                        ba += planeBytes;
                    }
                }
            }
            {   // A: query (aqrow, aqcol), channels c0 + (warp & 1) * 16 + j
                const int gi = c.i0 + aqrow, gj = c.j0 + aqcol;
                const bool ok = gi < H && gj < W;
                const int ca = c0 + (warp & 1) * FNA;
                const int nch = C - ca;
                uint64_t ba = (uint64_t)(fm0 + ((size_t)c.b * C + (nch > 0 ? ca : 0)) * plane + (size_t)(ok ? gi * W + gj : 0));
#pragma unroll
                for (int j = 0; j < FNA; ++j) {
                    v[Cfg::NB + j] = (ok && j < nch) ? __ldg(reinterpret_cast<const float*>(ba)) : 0.f;
                    ba += planeBytes;
                }
            }
        };
        auto store = [&](const float (&v)[Cfg::NB + FNA], auto S) {
            constexpr uint32_t so = decltype(S)::value * Cfg::STAGE_BYTES;
#if D2T_ABLATE_STAGING >= 2
            {   // [A hi | A lo | B hi | B lo]: vector split of the A and B tiles by the 256 producer threads
                const uint32_t st0 = smemBase + so;
                auto split = [&](uint32_t hiBase, uint32_t loBase, int bytes) {
                    for (int e = tid * 16; e < bytes; e += FPROD_WARPS * 32 * 16) {
                        float4 q;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(hiBase + e));
                        const float4 h = make_float4(f_tf32_rn(q.x), f_tf32_rn(q.y), f_tf32_rn(q.z), f_tf32_rn(q.w));
                        if (D2T_ABLATE_STAGING >= 3)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hiBase + e), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(loBase + e), "f"(q.x - h.x), "f"(q.y - h.y),
                                     "f"(q.z - h.z), "f"(q.w - h.w)
                                     : "memory");
                    }
                };
                split(st0, st0 + FA_BYTES, FA_BYTES);
                split(st0 + 2 * FA_BYTES, st0 + 2 * FA_BYTES + Cfg::B_BYTES, Cfg::B_BYTES);
                (void)v;
                return;
            }
#endif
#pragma unroll
            for (int j = 0; j < FNA; ++j) {
                const float hi = f_tf32_rn(v[Cfg::NB + j]);
                const uint32_t ad = stsA[j & 3] + so + (j >> 2) * FSBO;
                f_sts(ad, hi);
                f_sts(ad + FA_BYTES, v[Cfg::NB + j] - hi);
            }
#pragma unroll
            for (int j = 0; j < Cfg::NB; ++j) {
                const float hi = f_tf32_rn(v[j]);
                const uint32_t ad = stsB[j & 3] + so + (j >> 2) * FSBO;
                f_sts(ad, hi);
                f_sts(ad + Cfg::B_BYTES, v[j] - hi);
            }
        };
        auto step = [&](float (&v)[Cfg::NB + FNA], auto S) {
            constexpr int s = decltype(S)::value;
            f_mbar_wait(&bar_empty[s], ((k >> 1) & 1u) ^ 1u);   // the MMAs that read this stage two chunks ago are done
            store(v, S);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) f_mbar_arrive(&bar_full[s]);
            ++k;
            if (ld.valid(p)) {
                load(v, ld);
                ld.advance(p);
            }
            st.advance(p);
        };

        float va[Cfg::NB + FNA], vb[Cfg::NB + FNA];
        if (ld.valid(p)) { load(va, ld); ld.advance(p); }
        if (ld.valid(p)) { load(vb, ld); ld.advance(p); }
        while (st.valid(p)) {
            step(va, FInt<0>{});
            if (!st.valid(p)) break;
            step(vb, FInt<1>{});
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * Cfg::N)));
    }
}

}  // namespace

bool corr_umma_fwd_supported(int B, int C, int H, int W, int d, int stride) {
    if (stride != 1 || d != FD) return false;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return false;
    if ((long long)B * C * H * W >= (1ll << 31)) return false;
    if ((long long)B * H * W * FKK >= (1ll << 31)) return false;
    return true;
}

template <int CR>
static int f_count_items(int H, int tilesX, int tilesY) {
    int perImage = 0;
    for (int ty = 0; ty < tilesY; ++ty) {
        int ng = 0;
        for (int g = 0; g < FCfg<CR>::NCH; ++g) ng += f_chunk_valid<CR>(ty * FQROWS, g, H) ? 1 : 0;
        perImage += ng * tilesX;
    }
    return perImage;
}

template <int CR>
static int f_launch(const float* fm0, const float* fm1, float* out, FPlan p, int sms, cudaStream_t st) {
    p.itemsPerImage = f_count_items<CR>(p.H, p.tilesX, p.tilesY);
    p.nItems = p.B * p.itemsPerImage;
    const size_t smem = (size_t)FSTAGES * FCfg<CR>::STAGE_BYTES + 1024;
    D2T_SMEM_OPTIN(corr_fwd_umma_kernel<CR>, smem);
    const int grid = p.nItems < sms ? p.nItems : sms;
    corr_fwd_umma_kernel<CR><<<grid, FTHREADS, smem, st>>>(fm0, fm1, out, p);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

int corr_umma_fwd_launch(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, const CorrOutStrides* os,
                         cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    FPlan p;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.tilesX = ceil_div(W, FQCOLS);
    p.tilesY = ceil_div(H, FQROWS);
    p.nChunks = ceil_div(C, FKC);
    if (os) {
        p.os = *os;
    } else {
        p.os.sb = (long long)H * W * FKK; p.os.sp = FKK; p.os.st = 1;
    }
    // 8-row chunks unless they would leave more than half of the SMs without an item (scheduling only: same bits)
    if (2 * B * f_count_items<8>(H, p.tilesX, p.tilesY) <= di.sm_count) return f_launch<4>(fm0, fm1, out, p, di.sm_count, st);
    return f_launch<8>(fm0, fm1, out, p, di.sm_count, st);
}

}  // namespace d2t
