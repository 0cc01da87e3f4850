// pool_tc.cu -- float32 ROIPool backward on the 5th-generation tensor cores (tcgen05 + TMEM).  sm_100a.  EXPERIMENT
// (opt-in, D2T_ROIPOOL_BWD=tc) -- see DESIGN.md section 2.3 for the measurements.
//
// The backward of average pooling is separable:
//     grad_fm[c, y, x] = sum_{r,i,j} [I0_ri <= y < I1_ri] * [J0_rj <= x < J1_rj] * grad_out[r, c, i, j] / numel_rij
// (reference: roipool_cuda.cu:115-125, one atomicAdd per bin pixel).  For ONE pair of pixel rows the sum over the
// (RoI, bin row) pairs that touch it is a GEMM with a 0/1 operand:
//     D[m = pixel (2 rows x 64 columns)][n = channel] = sum_{k = (r, i, j)} A[m][k] * B[n][k]
//     A[(y, x)][(r, i, j)] = 1 if pixel (y, x) lies in bin (i, j) of RoI r, else 0        (exact in BF16)
//     B[c][(r, i, j)]      = grad_out[r, c, i, j] / numel_rij, split into three BF16 pieces (hi + mid + lo = 24 bits)
// and K only runs over the (r, i) that intersect the row pair (a fraction 3.8 / 38 of all bin rows), so the row
// sparsity of the scatter is kept and the column expansion rides on the tensor core.  FP32 accumulation in TMEM in
// ascending (r, i) order: bitwise reproducible, no atomics.
//
//   block      one (r, i): 8 K entries (7 bin columns + a zero pad) = one 16-byte row per pixel (A) / channel (B) in the
//              K-major non-swizzled canonical layout; an MMA (K = 16) takes two blocks, the second one `LBO` bytes
//              after the first (building block verified in tools/umma_bf16_blocks_test.cu).
//   item       (row pair, tile of 192 channels): M = 128, N = 192 accumulators = 192 TMEM columns.  Items are ordered
//              centre rows first (most RoIs) and dealt to the CTAs boustrophedon, so every CTA gets a heavy and a light one.
//   chunk      8 blocks: A 16 KB + B 3 x 24 KB per stage, 2 stages.
//   warps 0-15 B producers: lane = (channel of a quad, bin column); 4-byte loads of grad_out (the 7 floats of a bin row
//              are contiguous, 4 channels per instruction), scale, split, 2-byte stores into the operand layout.
//   warps 16-19 A producers (thread = pixel): indicator rows from the packed edge table; after the last chunk they are the
//              epilogue: tcgen05.ld -> grad_fm[c][y][x] (lane = x: coalesced rows).
//   warp 20    issues the MMAs, tcgen05.commit -> the stage's `empty` barrier / the item's accumulator barrier.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

namespace {

constexpr int TK = 7, TKK = 49;
constexpr int TN = 192;                         // channels per item = UMMA N = TMEM columns
constexpr int TM = 128;                         // 2 pixel rows x 64 columns = UMMA M
constexpr int TKB = 4;                          // blocks per chunk
constexpr int TSTAGES = 3;
constexpr int TA_HALF = TM * 16;                // bytes of one K half (4 tf32 per row) of an A block
constexpr int TB_HALF = TN * 16 + 64;           // ... of a B block, plus 64 bytes: the lanes of bin columns 4-7 store into the
                                                // second half and must not land on the banks of columns 0-3 (LBO is free)
constexpr int TA_BLK = 2 * TA_HALF;             // 4 KB
constexpr int TB_BLK = 2 * TB_HALF;             // 6 KB per piece
constexpr int TA_BYTES = TKB * TA_BLK;          // 16 KB
constexpr int TB_BYTES = TKB * TB_BLK;          // 24 KB per piece
constexpr int TSTAGE_BYTES = TA_BYTES + 2 * TB_BYTES;  // 64 KB
constexpr int TPROD_WARPS = 16;
constexpr int TA_WARP0 = TPROD_WARPS;           // warps 16-19 (TMEM lane quarter = warp % 4)
constexpr int TMMA_WARP = TA_WARP0 + 4;
constexpr int TTHREADS = (TMMA_WARP + 1) * 32;  // 672
constexpr int TMAXR = 320;
constexpr int THDR = 480;                        // list entries with a precomputed header (offset + 8 reciprocals)
constexpr int TQUADS = TN / 4;                  // 48 channel quads per item
constexpr int TCQ = TQUADS / TPROD_WARPS;       // 3 quads per producer warp and chunk
// kind::tf32: FP32 accumulate, TF32 x TF32, both K-major
constexpr uint32_t kTIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t t_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_smem(bar)), "r"(count));
}
// plain try_wait spin (no suspend-time hint): chunks are short here and the wake-up latency of a suspended wait would be
// exposed once per chunk
__device__ __forceinline__ void t_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            t_smem(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void t_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(t_smem(bar)) : "memory");
}
// K-major SWIZZLE_NONE descriptor: 16-byte rows, 8-row groups 128 bytes apart (SBO), second K half `lbo` bytes on
__device__ __forceinline__ uint64_t t_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void t_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kTIdesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void t_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t_smem(bar)) : "memory");
}
__device__ __forceinline__ void t_sts32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// round-to-nearest TF32 (10 explicit mantissa bits); v - hi is exact in FP32 and |v - hi| <= 2^-11 |v|
__device__ __forceinline__ float t_tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

// I0 | I1<<8 | J0<<16 | J1<<24 of bin index b (row edges from H, column edges from W); reference roipool_cuda.cu:38-50
__device__ __forceinline__ uint32_t t_pack_edges(const float* __restrict__ roi, int b, int H, int W) {
    int i0, i1, j0, j1;
    bin_edge<float, true>(roi[0], roi[2], b, TK, H, i0, i1);
    bin_edge<float, true>(roi[1], roi[3], b, TK, W, j0, j1);
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)j0 << 16) | ((uint32_t)j1 << 24);
}

__global__ void __launch_bounds__(TTHREADS, 1)
roipool_tc_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R,
                      int C, int H, int W, int nTiles, int nPairs, int dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint32_t* edgeS = reinterpret_cast<uint32_t*>(smem + TSTAGES * TSTAGE_BYTES);   // [TMAXR][7]
    uint16_t* listS = reinterpret_cast<uint16_t*>(edgeS + TMAXR * TK);              // [TMAXR * 7]  (r << 3 | i)
    int* hdrSrc = reinterpret_cast<int*>(listS + TMAXR * TK);                       // [THDR] grad_out offset of (c0, i, j=0)
    float* hdrInv = reinterpret_cast<float*>(hdrSrc + THDR);                        // [THDR][8] 1 / numel (0: empty bin / pad)
    __shared__ __align__(8) uint64_t bar_full[TSTAGES], bar_empty[TSTAGES], bar_acc;
    __shared__ uint32_t tmem_base_s;
    __shared__ int kcntS;
    __shared__ int wcntS[TTHREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int HW = H * W;
    const int nItems = nTiles * nPairs;
    const int G = gridDim.x;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t_smem(&tmem_base_s)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) {
            t_mbar_init(&bar_full[s], TPROD_WARPS + 4);
            t_mbar_init(&bar_empty[s], 1);
        }
        t_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int idx = tid; idx < R * TK; idx += TTHREADS) {
        const int rr = idx / TK, b = idx - rr * TK;
        edgeS[idx] = t_pack_edges(rois + (size_t)rr * 4, b, H, W);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t smemBase = t_smem(smem);

    uint32_t kg = 0;       // chunks processed so far by this CTA (all roles count alike): stage = kg & 1
    uint32_t accItems = 0;  // items with at least one chunk so far: phase of bar_acc

    for (int round = 0;; ++round) {
        const int q = round * G + ((round & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x);
        if (round * G >= nItems) break;
        if (q >= nItems) continue;  // (uniform per CTA; the next round starts past nItems and ends the loop)
        const int rank = q / nTiles, tile = q - rank * nTiles;
        // row pairs, centre first (most RoIs): c, c+1, c-1, c+2, ... then walk down to pair 0
        const int cP = nPairs >> 1, U2 = 2 * (nPairs - 1 - cP);
        const int p = rank < U2 ? ((rank & 1) ? cP + 1 + (rank >> 1) : cP - (rank >> 1)) : cP - (U2 >> 1) - (rank - U2);
        const int y0 = 2 * p;
        const int c0 = tile * TN;
        const int cb = min(TN, C - c0);

        // ---- the (RoI, bin row) blocks that touch rows y0, y0+1, ascending ------------------------------------
        // every warp compacts one contiguous segment of the R x 7 candidates: count, prefix over the warps, write
        {
            constexpr int NW = TTHREADS / 32;
            const int total = (dbg & 32) ? 0 : R * TK;
            const int seg = ((total + NW - 1) / NW + 31) & ~31;  // candidates per warp, a multiple of 32
            const int lo = warp * seg, hi = min(total, lo + seg);
            constexpr int MAXR = (TMAXR * TK / NW + 31) / 32 + 1;  // ballot rounds per warp (R <= TMAXR)
            unsigned bits[MAXR];
            int cnt = 0;
#pragma unroll
            for (int nr = 0; nr < MAXR; ++nr) {
                const int e = lo + nr * 32 + lane;
                bool f = false;
                if (e < hi) {
                    const uint32_t ed = edgeS[e];
                    const int i0 = ed & 255, i1 = (ed >> 8) & 255;
                    f = i1 > i0 && i0 < y0 + 2 && i1 > y0;
                }
                bits[nr] = __ballot_sync(0xffffffffu, f);
                cnt += __popc(bits[nr]);
            }
            if (lane == 0) wcntS[warp] = cnt;
            __syncthreads();
            int off = 0, all = 0;
            for (int w = 0; w < NW; ++w) {
                const int cw = wcntS[w];
                off += w < warp ? cw : 0;
                all += cw;
            }
#pragma unroll
            for (int nr = 0; nr < MAXR; ++nr) {
                const int e = lo + nr * 32 + lane;
                const unsigned bal = bits[nr];
                if ((bal >> lane) & 1u) {
                    const int rr = e / TK, i = e - rr * TK;
                    listS[off + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)((rr << 3) | i);
                }
                off += __popc(bal);
            }
            if (tid == 0) kcntS = all;
        }
        __syncthreads();
        const int kp = (dbg & 16) ? 0 : kcntS;
        // per-entry headers, once per item: what every producer lane would otherwise recompute for every chunk
        for (int idx = tid; idx < min(kp, THDR) * 8; idx += TTHREADS) {
            const int e = idx >> 3, jj = idx & 7;
            const int ent = listS[e];
            const int rr = ent >> 3, i = ent & 7;
            const uint32_t ei = edgeS[rr * TK + i];
            const int hI = (int)((ei >> 8) & 255) - (int)(ei & 255);
            float inv = 0.f;
            if (jj < TK) {
                const uint32_t ej = edgeS[rr * TK + jj];
                const int wJ = (int)(ej >> 24) - (int)((ej >> 16) & 255);
                if (hI > 0 && wJ > 0) inv = 1.0f / (float)(hI * wJ);
            }
            hdrInv[idx] = inv;
            if (jj == 0) hdrSrc[e] = (rr * C + c0) * TKK + i * TK;
        }
        __syncthreads();
        const int nch = (kp + TKB - 1) / TKB;

        if (warp < TPROD_WARPS) {
            // ================================ B producers ===========================================================
            // All 16 warps stage every chunk: 12 values per thread, loaded before the stage wait.  What bounds this kernel is
            // the number of instructions a warp executes per chunk (each warp issues once per ~8 cycles with 21 warps on
            // the SM), so the per-block set-up comes from the item's header table instead of being recomputed per lane.
            const int j = lane & 7, chl = lane >> 3;
            // loop-invariant per item: one base pointer per channel quad (channels past the tile's end are clamped to its
            // last channel -- their operand rows feed accumulator columns that are never stored), one operand offset
            const float* base[TCQ];
            uint32_t dstOff[TCQ];
#pragma unroll
            for (int cq = 0; cq < TCQ; ++cq) {
                const int ch = (cq * TPROD_WARPS + warp) * 4 + chl;
                base[cq] = go + (size_t)min(ch, cb - 1) * TKK + min(j, TK - 1);
                dstOff[cq] = (uint32_t)(TA_BYTES + (j >> 2) * TB_HALF + ch * 16 + (j & 3) * 4);
            }
            // operands of chunk c -> registers: per block the header's offset and reciprocal (0 for the pad column, an
            // empty bin or a block past the list's end), then 12 unconditional loads
            float vN[TCQ][TKB], invN[TKB];
            auto fetch = [&](int c) {
                int off[TKB];
#pragma unroll
                for (int b = 0; b < TKB; ++b) {
                    const int e = c * TKB + b;
                    off[b] = 0;
                    invN[b] = 0.f;
                    if (e < kp) {
                        if (e < THDR) {
                            off[b] = hdrSrc[e];
                            invN[b] = hdrInv[e * 8 + j];
                        } else {  // more blocks than header slots (very many RoIs on one row pair)
                            const int ent = listS[e];
                            const int rr = ent >> 3, i = ent & 7;
                            const uint32_t ei = edgeS[rr * TK + i], ej = edgeS[rr * TK + min(j, TK - 1)];
                            const int hI = (int)((ei >> 8) & 255) - (int)(ei & 255);
                            const int wJ = (int)(ej >> 24) - (int)((ej >> 16) & 255);
                            if (j < TK && hI > 0 && wJ > 0) invN[b] = 1.0f / (float)(hI * wJ);
                            off[b] = (rr * C + c0) * TKK + i * TK;
                        }
                    }
                }
                if (!(dbg & 2))
#pragma unroll
                for (int cq = 0; cq < TCQ; ++cq)
#pragma unroll
                    for (int b = 0; b < TKB; ++b) vN[cq][b] = __ldg(base[cq] + off[b]);
            };
#pragma unroll
            for (int cq = 0; cq < TCQ; ++cq)
#pragma unroll
                for (int b = 0; b < TKB; ++b) vN[cq][b] = 0.f;
            if (nch > 0) fetch(0);
            for (int c = 0; c < nch; ++c) {
                const uint32_t k = kg + c, s = k % TSTAGES;
                float v[TCQ][TKB], inv[TKB];
#pragma unroll
                for (int b = 0; b < TKB; ++b) {
                    inv[b] = invN[b];
#pragma unroll
                    for (int cq = 0; cq < TCQ; ++cq) v[cq][b] = vN[cq][b];
                }
                if (c + 1 < nch) fetch(c + 1);  // the next chunk's loads fly while this one is converted
                t_mbar_wait(&bar_empty[s], ((k / TSTAGES) & 1u) ^ 1u);
                const uint32_t sBase = smemBase + s * TSTAGE_BYTES;
                if (!(dbg & 4))
#pragma unroll
                for (int cq = 0; cq < TCQ; ++cq) {
                    const uint32_t dst = sBase + dstOff[cq];
#pragma unroll
                    for (int b = 0; b < TKB; ++b) {
                        // inv = 0 must give 0 whatever was loaded (0 * Inf would poison the whole tile)
                        const float x = inv[b] != 0.f ? v[cq][b] * inv[b] : 0.f;
                        const float hi = t_tf32_rn(x);
                        t_sts32(dst + b * TB_BLK, hi);
                        t_sts32(dst + b * TB_BLK + TB_BYTES, x - hi);
                    }
                }
                if (!(dbg & 64)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) t_mbar_arrive(&bar_full[s]);
            }
        } else if (warp < TMMA_WARP) {
            // ================================ A producers, then epilogue ============================================
            const int m = tid - TA_WARP0 * 32;  // pixel of the tile = TMEM lane
            const int y = y0 + (m >> 6), x = m & 63;
            const bool pixOk = y < H && x < W;
            // staging map: thread = (pixel column ax, half of the chunk's blocks), both pixel rows of the pair -- the column
            // mask of a RoI is computed once for the two rows and reused while consecutive blocks belong to the same RoI
            const int ax = m & 63, ahalf = m >> 6;
            int lastR = -1;
            unsigned lastMask = 0;
            for (int c = 0; c < nch; ++c) {
                const uint32_t k = kg + c, s = k % TSTAGES;
                t_mbar_wait(&bar_empty[s], ((k / TSTAGES) & 1u) ^ 1u);
                const uint32_t aBase = smemBase + s * TSTAGE_BYTES + ax * 16;
                if (!(dbg & 8))
#pragma unroll
                for (int bb = 0; bb < TKB / 2; ++bb) {
                    const int b = ahalf * (TKB / 2) + bb;
                    const int e = c * TKB + b;
                    unsigned mask = 0;  // bin columns of block e that contain pixel column ax
                    bool cov0 = false, cov1 = false;
                    if (e < kp && ax < W) {
                        const int ent = listS[e];
                        const int rr = ent >> 3, i = ent & 7;
                        const uint32_t ei = edgeS[rr * TK + i];
                        const int I0 = ei & 255, I1 = (ei >> 8) & 255;
                        cov0 = y0 >= I0 && y0 < I1;
                        cov1 = y0 + 1 >= I0 && y0 + 1 < I1 && y0 + 1 < H;
                        if (rr != lastR) {
                            lastMask = 0;
#pragma unroll
                            for (int jj = 0; jj < TK; ++jj) {
                                const uint32_t ej = edgeS[rr * TK + jj];
                                lastMask |= (ax >= (int)((ej >> 16) & 255) && ax < (int)(ej >> 24)) ? (1u << jj) : 0u;
                            }
                            lastR = rr;
                        }
                        mask = lastMask;
                    }
                    const uint32_t one = 0x3F800000u;  // 1.0f: exact in TF32
                    const unsigned m0 = cov0 ? mask : 0u, m1 = cov1 ? mask : 0u;
                    const uint32_t a0 = aBase + b * TA_BLK, a1 = a0 + 64 * 16;  // rows y0 and y0 + 1 of the tile
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"((m0 & 1u) ? one : 0u), "r"((m0 & 2u) ? one : 0u),
                                 "r"((m0 & 4u) ? one : 0u), "r"((m0 & 8u) ? one : 0u) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0 + TA_HALF), "r"((m0 & 16u) ? one : 0u),
                                 "r"((m0 & 32u) ? one : 0u), "r"((m0 & 64u) ? one : 0u), "r"(0u) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"((m1 & 1u) ? one : 0u), "r"((m1 & 2u) ? one : 0u),
                                 "r"((m1 & 4u) ? one : 0u), "r"((m1 & 8u) ? one : 0u) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1 + TA_HALF), "r"((m1 & 16u) ? one : 0u),
                                 "r"((m1 & 32u) ? one : 0u), "r"((m1 & 64u) ? one : 0u), "r"(0u) : "memory");
                }
                if (!(dbg & 64)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) t_mbar_arrive(&bar_full[s]);
            }
            // epilogue: accumulators -> grad_fm[c0 + n][y][x]; lane = x, so every store instruction writes one row segment
            if (nch > 0) {
                t_mbar_wait(&bar_acc, accItems & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            float* dst = gin + (size_t)c0 * HW + (size_t)y * W + x;
            const int quarter = warp - TA_WARP0;
#pragma unroll 1
            for (int n0 = 0; n0 < TN; n0 += 32) {
                uint32_t r[32];
                if (nch > 0) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)n0;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int t = 0; t < 32; ++t) r[t] = 0u;  // no RoI touches this row pair
                }
                if (pixOk) {
#pragma unroll
                    for (int t = 0; t < 32; ++t)
                        if (n0 + t < cb) dst[(size_t)(n0 + t) * HW] = __uint_as_float(r[t]);
                }
            }
        } else {
            // ================================ MMA issuer ============================================================
            for (int c = 0; c < nch; ++c) {
                const uint32_t k = kg + c, s = k % TSTAGES;
                t_mbar_wait(&bar_full[s], (k / TSTAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t aS = smemBase + s * TSTAGE_BYTES, bS = aS + TA_BYTES;
                    const int nb = min(TKB, kp - c * TKB);  // live blocks of this chunk
#pragma unroll
                    for (int b = 0; b < TKB; ++b) {
                        if (b < nb && !(dbg & 1)) {  // one MMA (K = 8) per block and piece
                            const uint64_t da = t_desc(aS + b * TA_BLK, TA_HALF);
                            t_mma(tmem_base, da, t_desc(bS + b * TB_BLK, TB_HALF), (c == 0 && b == 0) ? 0u : 1u);
                            t_mma(tmem_base, da, t_desc(bS + TB_BYTES + b * TB_BLK, TB_HALF), 1u);
                        }
                    }
                    if (dbg & 128) {  // experiment: software arrives instead of tcgen05.commit (only meaningful with dbg & 1)
                        t_mbar_arrive(&bar_empty[s]);
                        if (c == nch - 1) t_mbar_arrive(&bar_acc);
                    } else {
                        t_commit(&bar_empty[s]);
                        if (c == nch - 1) t_commit(&bar_acc);
                    }
                }
                __syncwarp();
            }
        }
        kg += (uint32_t)nch;
        if (nch > 0) ++accItems;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // the item's accumulators are drained and its list is no longer read
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
    }
}

size_t tc_smem_bytes() {
    return (size_t)TSTAGES * TSTAGE_BYTES + (size_t)TMAXR * TK * sizeof(uint32_t) + (size_t)TMAXR * TK * sizeof(uint16_t) +
           (size_t)THDR * 9 * sizeof(float) + 128;
}

}  // namespace

bool roipool_tc_bwd_supported(int R, int C, int H, int W, int k) {
    if (k != TK || R <= 0 || R > TMAXR || C <= 0 || H <= 0 || W <= 0 || H > 255 || W > 64) return false;
    if ((long long)R * C * TKK >= (1ll << 31)) return false;
    const char* e = getenv("D2T_ROIPOOL_BWD");  // opt-in
    if (!(e && e[0] == 't')) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    return tc_smem_bytes() <= (size_t)di.max_smem_optin;
}

int roipool_tc_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W, cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int nTiles = ceil_div(C, TN), nPairs = ceil_div(H, 2);
    const int nItems = nTiles * nPairs;
    const int grid = nItems < di.sm_count ? nItems : di.sm_count;
    const size_t smem = tc_smem_bytes();
    D2T_CUDA_TRY(cudaFuncSetAttribute(roipool_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dbg = 0;  // experiment switches (D2T_TC_DBG; results are wrong when set): 1 no MMAs, 2 no loads, 4 no B stores, 8 no A stores
    if (const char* e = getenv("D2T_TC_DBG")) dbg = atoi(e);
    roipool_tc_bwd_kernel<<<grid, TTHREADS, smem, st>>>(go, rois, gin, R, C, H, W, nTiles, nPairs, dbg);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
