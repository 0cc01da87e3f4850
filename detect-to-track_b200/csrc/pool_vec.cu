// pool_vec.cu -- float32 ROIPool forward/backward, third generation: [pixel][16 channel] slabs walked with
// 128-bit shared-memory accesses, row prefix sums (forward) / row difference arrays (backward).  sm_100a.
//
// Why: ROIPool at the D&T track-head size (C=1891, 38x63, R=300, k=7) moves 129 MB per direction, 20 us at the HBM
// roof.  The earlier kernels (pool.cu, pool_fast.cu) were instruction-bound at 150-370 us because every lane carried
// one channel.  Here a lane carries FOUR channels of one bin column, so a bin row costs two LDS.128 per lane and the
// kernels are bound by the shared-memory pipe instead (profiles/).
//
//   slab      a CTA owns <= 16 consecutive channels for all RoIs.  Shared memory holds them as P[y][x][16] with
//             x in [0, W] (one extra column), 64 B per pixel; the four 16-byte channel quads of a pixel are XOR-
//             swizzled with (x >> 1) & 3 so that the transposing loads / stores of the prologue and epilogue
//             (8 consecutive pixels, one quad) hit 8 distinct bank groups.
//   forward   P[y][x] = sum_{x' < x} fm[y][x'] (exclusive row prefix, built once per CTA).  A warp takes one RoI at a
//             time (dynamic queue); lane (j, q) owns bin column j and channel quad q, walks the RoI's rows once and
//             forms  d(y) = P[y][J1_j] - P[y][J0_j]  (2 LDS.128), summing d over the rows of each bin row i.  Adjacent
//             bin rows overlap by at most one pixel row (floor/ceil edges); that row's d is reused, not reloaded.
//             Results are staged per warp and leave as one contiguous run  out[r, c0:c0+cb, :, :].
//             Differs from the reference's left-to-right pixel sum (roipool_cuda.cu:52-61) only by float rounding
//             (tested at rtol 1e-4); d2t_roipool_fwd_f32_exact keeps the bit-identical kernel.
//   backward  the adjoint: D[y][J0_j] += t, D[y][J1_j] -= t with t = sum_{i covers y} grad_out[r,c,i,j] / numel_ij,
//             then grad_fm[y][x] = sum_{x' <= x} D[y][x'] (inclusive row prefix in the epilogue).  Every WARP owns
//             pixel rows (y % nWarps == warp) for all RoIs and walks the RoIs in ascending order, so each row is
//             updated by exactly one warp, in a fixed order: no atomics (reference: atomicAdd per bin pixel,
//             roipool_cuda.cu:119-125), bitwise reproducible.  grad_out blocks of RoI groups are prefetched through
//             registers into a double-buffered shared stage.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

constexpr int kVecFwdWarps = 16;
constexpr int kVecFwdThreads = kVecFwdWarps * 32;
constexpr int kVecSlots = 16;     // channel slots per CTA (4 quads of 4)
constexpr int kVecRChunk = 512;   // RoIs per edge-table chunk
constexpr int kVecBwdMaxWarps = 20;
constexpr int kVecBwdRG = 8;      // RoIs per staged grad_out group

// float offset of channel quad q of pixel column x inside a row
__device__ __forceinline__ int vec_pix_off(int x, int q) { return x * kVecSlots + ((q ^ ((x >> 1) & 3)) << 2); }
__host__ __device__ constexpr int vec_row_pitch(int W) { return (W + 1) * kVecSlots + kVecSlots; }  // +64 B: odd/even rows in opposite bank halves
// stage pitch per channel: >= k*k and == 2 (mod 8) so that the 4 quads of a lane group land in distinct banks
__host__ __device__ constexpr int vec_stage_pitch(int kk) { return ((kk + 5) / 8) * 8 + 2; }

// 1/x to ~1 ulp; 1/0 = +inf
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// packed edges of bin index b (row edges from H, column edges from W): I0 | I1<<8 | J0<<16 | J1<<24
__device__ __forceinline__ uint32_t vec_pack_edges(const float* __restrict__ roi, int b, int k, int H, int W) {
    int i0, i1, j0, j1;
    bin_edge<float, true>(roi[0], roi[2], b, k, H, i0, i1);
    bin_edge<float, true>(roi[1], roi[3], b, k, W, j0, j1);
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)j0 << 16) | ((uint32_t)j1 << 24);
}

// ----------------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kVecFwdThreads, 1)
roipool_vec_fwd_kernel(const float* __restrict__ fm, const float* __restrict__ rois, float* __restrict__ out, int R,
                       int C, int H, int W, int CB) {
    constexpr int KK = K * K;
    constexpr int SP = vec_stage_pitch(KK);
    constexpr int NT = kVecFwdThreads;
    constexpr int NOUT = (kVecSlots * KK + 31) / 32;  // copy-out trips per lane
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int rowPitch = vec_row_pitch(W);
    float* P = reinterpret_cast<float*>(smem_raw);
    float* stageAll = P + (size_t)H * rowPitch;
    uint32_t* edgeS = reinterpret_cast<uint32_t*>(stageAll + kVecFwdWarps * kVecSlots * SP);
    int* counter = reinterpret_cast<int*>(edgeS + kVecRChunk * K);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);
    const int HW = H * W;

    // ---- prologue: slab load (coalesced along x, 4 planes per thread -> one swizzled STS.128) ----------------
    for (int idx = tid; idx < H * kVecSlots; idx += NT) P[(idx >> 4) * rowPitch + (idx & 15)] = 0.f;  // column 0
    {
        const int total = 4 * HW;
#pragma unroll 2
        for (int idx = tid; idx < total; idx += NT) {
            const int q = idx / HW, pix = idx - q * HW;
            const int y = pix / W, x = pix - y * W;
            const float* src = fm + (size_t)(c0 + 4 * q) * HW + pix;
            float4 v;
            v.x = (4 * q + 0 < cb) ? __ldg(src) : 0.f;
            v.y = (4 * q + 1 < cb) ? __ldg(src + HW) : 0.f;
            v.z = (4 * q + 2 < cb) ? __ldg(src + 2 * HW) : 0.f;
            v.w = (4 * q + 3 < cb) ? __ldg(src + 3 * HW) : 0.f;
            st4(P + y * rowPitch + vec_pix_off(x + 1, q), v);
        }
    }
    __syncthreads();
    // in-place row scan: afterwards P[y][x] = sum of the row left of x (P[y][0] = 0)
    for (int t = tid; t < H * 4; t += NT) {
        const int y = t >> 2, q = t & 3;
        float* row = P + y * rowPitch;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int x = 1; x <= W; ++x) {
            float* p = row + vec_pix_off(x, q);
            const float4 v = ld4(p);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
            st4(p, acc);
        }
    }

    // per-lane constants
    const int j = lane >> 2, q = lane & 3;
    const bool jact = j < K;
    const int jc = jact ? j : K - 1;
    float* stage = stageAll + warp * kVecSlots * SP;
    int soff[NOUT];
#pragma unroll
    for (int t = 0; t < NOUT; ++t) {
        const int o = lane + 32 * t;
        const int ch = o / KK;
        soff[t] = ch * SP + (o - ch * KK);
    }
    const int nOut = cb * KK;

    for (int r0 = 0; r0 < R; r0 += kVecRChunk) {
        const int nr = min(kVecRChunk, R - r0);
        __syncthreads();  // previous chunk consumed (and, first time, the scan is complete)
        if (tid == 0) *counter = 0;
        for (int idx = tid; idx < nr * K; idx += NT) {
            const int rr = idx / K, b = idx - rr * K;
            edgeS[idx] = vec_pack_edges(rois + (size_t)(r0 + rr) * 4, b, K, H, W);
        }
        __syncthreads();

        while (true) {
            int rr = 0;
            if (lane == 0) rr = atomicAdd(counter, 1);
            rr = __shfl_sync(0xffffffffu, rr, 0);
            if (rr >= nr) break;
            const uint32_t* ed = edgeS + rr * K;
            const uint32_t ej = ed[jc];
            const int J0 = (ej >> 16) & 255, J1 = ej >> 24;
            const int wj = J1 - J0;
            const float* pA = P + vec_pix_off(J0, q);
            const float* pB = P + vec_pix_off(J1, q);
            float4 lastd = make_float4(0.f, 0.f, 0.f, 0.f);
            int lasty = -1;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const uint32_t ei = ed[i];
                const int I0 = ei & 255, I1 = (ei >> 8) & 255;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int y = I0;
                if (I0 == lasty && I1 > I0) {  // the previous bin row's last pixel row is this one's first
                    acc = lastd;
                    y = I0 + 1;
                }
                const float* pa = pA + y * rowPitch;
                const float* pb = pB + y * rowPitch;
#pragma unroll 1
                for (; y < I1; ++y, pa += rowPitch, pb += rowPitch) {
                    const float4 b = ld4(pb);
                    const float4 a = ld4(pa);
                    lastd.x = b.x - a.x;
                    lastd.y = b.y - a.y;
                    lastd.z = b.z - a.z;
                    lastd.w = b.w - a.w;
                    acc.x += lastd.x;
                    acc.y += lastd.y;
                    acc.z += lastd.z;
                    acc.w += lastd.w;
                }
                lasty = I1 > I0 ? I1 - 1 : -1;
                const float inv = rcp_approx((float)((I1 - I0) * wj));  // 1/0 = inf; 0 * inf = NaN like the reference's 0/0 (F7)
                if (jact) {
                    float* s = stage + (4 * q) * SP + i * K + j;
                    s[0] = acc.x * inv;
                    s[SP] = acc.y * inv;
                    s[2 * SP] = acc.z * inv;
                    s[3 * SP] = acc.w * inv;
                }
            }
            __syncwarp();
            float* dst = out + ((size_t)(r0 + rr) * C + c0) * KK;  // out[r, c0:c0+cb, :, :] is one contiguous run
#pragma unroll
            for (int t = 0; t < NOUT; ++t) {
                const int o = lane + 32 * t;
                if (o < nOut) dst[o] = stage[soff[t]];
            }
            __syncwarp();
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------------------------------
// smem: D[H][rowPitch] | gstage[2][RG][KK][16] | edges[RCH][K] | simple[RCH] | cover[2][H][RG] | off[2][RG][32] | counter[2]
//
// Work unit = (RoI group, pixel row).  Within a group every pixel row is claimed by exactly one warp (dynamic queue),
// which applies the group's RoIs to that row in ascending RoI order; groups are separated by one CTA barrier.  So each
// D row sees its updates in a fixed order whichever warp applies them: deterministic without atomics.
template <int K>
__global__ void __launch_bounds__(kVecBwdMaxWarps * 32, 1)
roipool_vec_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R,
                       int C, int H, int W, int CB) {
    constexpr int KK = K * K;
    constexpr int SP = vec_stage_pitch(KK);
    constexpr int RG = kVecBwdRG;
    static_assert(RG == 8, "cover masks of a row are read as one 64-bit word");
    constexpr int GSZ = RG * kVecSlots * SP;  // floats per stage buffer
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int rowPitch = vec_row_pitch(W);
    float* D = reinterpret_cast<float*>(smem_raw);
    float* gS = D + (size_t)H * rowPitch;
    uint32_t* edgeS = reinterpret_cast<uint32_t*>(gS + 2 * GSZ);                 // [kVecRChunk][K]
    unsigned char* simpleS = reinterpret_cast<unsigned char*>(edgeS + kVecRChunk * K);  // [kVecRChunk]
    unsigned char* coverS = simpleS + kVecRChunk;                                // [2][H][RG], 8-byte aligned rows
    uint32_t* offS = reinterpret_cast<uint32_t*>(coverS + (size_t)2 * ((H * RG + 15) / 16 * 16));  // [2][RG][32]
    int* counter = reinterpret_cast<int*>(offS + 2 * RG * 32);
    constexpr int NITEM = 4;  // (RoI, quad, bin) staging items per thread and group: the host guarantees 4 * threads >= RG * 4 * KK

    const int NT = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);
    const int HW = H * W;
    const int coverBuf = (H * RG + 15) / 16 * 16;

    for (int idx = tid; idx < H * rowPitch / 4; idx += NT) st4(D + idx * 4, make_float4(0.f, 0.f, 0.f, 0.f));

    const int j = lane >> 2, q = lane & 3;
    const bool jact = j < K;

    for (int rc0 = 0; rc0 < R; rc0 += kVecRChunk) {  // RoI chunks: one edge table each (a single chunk for R <= 512)
        const int nrc = min(kVecRChunk, R - rc0);
        __syncthreads();
        for (int idx = tid; idx < nrc * K; idx += NT) {
            const int rr = idx / K, b = idx - rr * K;
            edgeS[idx] = vec_pack_edges(rois + (size_t)(rc0 + rr) * 4, b, K, H, W);
        }
        __syncthreads();
        // "simple" RoI: column edges strictly increasing => the lanes of one update instruction touch distinct pixels
        for (int rr = tid; rr < nrc; rr += NT) {
            const uint32_t* ed = edgeS + rr * K;
            bool simple = true;
#pragma unroll
            for (int b = 1; b < K; ++b) {
                const uint32_t a = ed[b - 1], c = ed[b];
                simple = simple && (((c >> 16) & 255) > ((a >> 16) & 255)) && ((c >> 24) > (a >> 24));
            }
            simpleS[rr] = simple ? 1 : 0;
        }

        // ---- staging of grad_out, pre-scaled and transposed ------------------------------------------------------
        // An item is (RoI rr of the group, channel quad qq, bin): 4 coalesced loads (lanes = consecutive bins of one
        // channel), scaled by 1 / (bin rows x bin columns) (0 for an empty bin), one STS.128 into
        //   gS[buf][rr][bin][slot], slot = qq ^ (j & 3)   (16-byte slots; j = bin column)
        // so that lane (j, q) of the update loop reads its four channels of bin (i, j) with ONE conflict-free LDS.128
        // and needs no edge arithmetic or reciprocal per pixel row.
        int itRR[NITEM], itBin[NITEM], itQ[NITEM];
#pragma unroll
        for (int n = 0; n < NITEM; ++n) {
            const int it = tid + n * NT;
            itRR[n] = it / (4 * KK);
            const int rem = it - itRR[n] * (4 * KK);
            itQ[n] = rem / KK;
            itBin[n] = rem - itQ[n] * KK;
            if (itRR[n] >= RG) itRR[n] = -1;
        }
        float4 pre[NITEM];
        auto prefetch = [&](int grp) {
            const int r0 = rc0 + grp * RG;
#pragma unroll
            for (int n = 0; n < NITEM; ++n) {
                const int rr = itRR[n];
                const bool ok = rr >= 0 && grp * RG + rr < nrc;
                const int ch = 4 * itQ[n];
                const float* src = go + ((size_t)(ok ? r0 + rr : rc0) * C + c0 + ch) * KK + itBin[n];
                pre[n].x = (ok && ch + 0 < cb) ? __ldg(src) : 0.f;
                pre[n].y = (ok && ch + 1 < cb) ? __ldg(src + KK) : 0.f;
                pre[n].z = (ok && ch + 2 < cb) ? __ldg(src + 2 * KK) : 0.f;
                pre[n].w = (ok && ch + 3 < cb) ? __ldg(src + 3 * KK) : 0.f;
            }
        };
        // stage the prefetched grad_out of group `grp`; build its row cover masks
        // (cover[y][rr] bit i set <=> bin row i of RoI rr contains pixel row y) and its per-lane update offsets
        auto commit = [&](int grp, int buf) {
            float* g = gS + buf * GSZ;
            const int nr = min(RG, nrc - grp * RG);
#pragma unroll
            for (int n = 0; n < NITEM; ++n) {
                const int rr = itRR[n];
                if (rr < 0 || rr >= nr) continue;
                const int bi = itBin[n] / K, bj = itBin[n] - bi * K;
                const uint32_t* ed = edgeS + (grp * RG + rr) * K;
                const uint32_t ei = ed[bi], ej = ed[bj];
                const int hI = (int)((ei >> 8) & 255) - (int)(ei & 255);
                const int wJ = (int)(ej >> 24) - (int)((ej >> 16) & 255);
                const float inv = (hI > 0 && wJ > 0) ? rcp_approx((float)(hI * wJ)) : 0.f;
                float4 v = pre[n];
                v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                st4(g + ((rr * KK + itBin[n]) * 4 + (itQ[n] ^ (bj & 3))) * 4, v);
            }
            for (int idx = tid; idx < H * RG; idx += NT) {
                const int y = idx / RG, rr = idx - y * RG;
                unsigned m = 0;
                if (rr < nr) {
                    const uint32_t* ed = edgeS + (grp * RG + rr) * K;
#pragma unroll
                    for (int b = 0; b < K; ++b) {
                        const uint32_t e = ed[b];
                        const int i0 = e & 255, i1 = (e >> 8) & 255;
                        m |= (i0 <= y && y < i1) ? (1u << b) : 0u;
                    }
                }
                coverS[buf * coverBuf + idx] = (unsigned char)m;
            }
            // per (RoI, lane): byte offsets inside a D row of the lane's two updates (+t at J0_j, -t at J1_j), bit 31 = simple
            for (int idx = tid; idx < RG * 32; idx += NT) {
                const int rr = idx >> 5, ln = idx & 31;
                uint32_t w = 0;
                if (rr < nr) {
                    const int jj = min(ln >> 2, K - 1), qq = ln & 3;
                    const uint32_t ej = edgeS[(grp * RG + rr) * K + jj];
                    const int J0 = (ej >> 16) & 255, J1 = ej >> 24;
                    w = (uint32_t)(vec_pix_off(J0, qq) * 4) | ((uint32_t)(vec_pix_off(J1, qq) * 4) << 14) |
                        (simpleS[grp * RG + rr] ? 0x80000000u : 0u);
                }
                offS[buf * RG * 32 + idx] = w;
            }
            if (tid == 0) counter[buf] = 0;
        };

        const int nGroups = (nrc + RG - 1) / RG;
        prefetch(0);
        commit(0, 0);
        const int laneG = jact ? (j * 4 + (q ^ (j & 3))) * 4 : 0;  // float offset of this lane's slot inside a bin row

        for (int grp = 0; grp < nGroups; ++grp) {
            const int buf = grp & 1;
            __syncthreads();  // stage / cover / offsets / counter of `buf` complete; everyone is done with group grp-1
            if (grp + 1 < nGroups) prefetch(grp + 1);
            const float* gB = gS + buf * GSZ + laneG;
            const unsigned char* cov = coverS + buf * coverBuf;
            const uint32_t* offB = offS + buf * RG * 32 + lane;
            while (true) {
                int task = 0;
                if (lane == 0) task = atomicAdd(&counter[buf], 1);
                task = __shfl_sync(0xffffffffu, task, 0);
                if (task >= H) break;
                // centre rows first (they are covered by most RoIs): better tail balance.  c = H/2; tasks alternate
                // c, c+1, c-1, c+2, ... while rows above c last (there are U = H-1-c <= c of them), then walk down to 0.
                const int cRow = H >> 1, U2 = 2 * (H - 1 - cRow);
                const int y = task < U2 ? ((task & 1) ? cRow + 1 + (task >> 1) : cRow - (task >> 1))
                                        : cRow - (U2 >> 1) - (task - U2);
                const uint2 masks2 = *reinterpret_cast<const uint2*>(cov + y * RG);  // one cover byte per RoI of the group
                char* row = reinterpret_cast<char*>(D + y * rowPitch);
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                unsigned masks = half ? masks2.y : masks2.x;
                while (masks != 0u) {  // RoIs of the group that cover this row, in ascending order
                    const int r4 = (__ffs(masks) - 1) >> 3;
                    unsigned cover = (masks >> (8 * r4)) & 0xffu;
                    masks &= ~(0xffu << (8 * r4));
                    const int rr = r4 + 4 * half;
                    const uint32_t w = offB[rr * 32];
                    const float* gR = gB + rr * (KK * 16);
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    while (cover) {  // bin rows containing y (1, or 2 where floor/ceil edges overlap)
                        const int i = __ffs(cover) - 1;
                        cover &= cover - 1;
                        const float4 v = ld4(gR + i * (K * 16));
                        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
                    }
                    float* pA = reinterpret_cast<float*>(row + (w & 0x3fffu));
                    float* pB = reinterpret_cast<float*>(row + ((w >> 14) & 0x3fffu));
                    if (w & 0x80000000u) {
                        if (jact) {
                            float4 a = ld4(pA);
                            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                            st4(pA, a);
                        }
                        __syncwarp();
                        if (jact) {
                            float4 b = ld4(pB);
                            b.x -= t.x; b.y -= t.y; b.z -= t.z; b.w -= t.w;
                            st4(pB, b);
                        }
                        __syncwarp();
                    } else {
                        for (int jj = 0; jj < K; ++jj) {  // degenerate RoI (bins thinner than a pixel / clamped): one bin column at a time
                            if (j == jj) {
                                float4 a = ld4(pA);
                                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                                st4(pA, a);
                            }
                            __syncwarp();
                            if (j == jj) {
                                float4 b = ld4(pB);
                                b.x -= t.x; b.y -= t.y; b.z -= t.z; b.w -= t.w;
                                st4(pB, b);
                            }
                            __syncwarp();
                        }
                    }
                }
                }
            }
            if (grp + 1 < nGroups) commit(grp + 1, buf ^ 1);
        }
    }
    __syncthreads();

    // ---- epilogue: inclusive row scan, then transposed write-out (LDS.128 -> 4 coalesced plane stores) -------
    for (int t = tid; t < H * 4; t += NT) {
        const int y = t >> 2, qq = t & 3;
        float* row = D + y * rowPitch;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int x = 0; x < W; ++x) {
            float* p = row + vec_pix_off(x, qq);
            const float4 v = ld4(p);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
            st4(p, acc);
        }
    }
    __syncthreads();
    {
        const int total = 4 * HW;
        for (int idx = tid; idx < total; idx += NT) {
            const int qq = idx / HW, pix = idx - qq * HW;
            const int y = pix / W, x = pix - y * W;
            const float4 v = ld4(D + y * rowPitch + vec_pix_off(x, qq));
            float* dst = gin + (size_t)(c0 + 4 * qq) * HW + pix;
            if (4 * qq + 0 < cb) dst[0] = v.x;
            if (4 * qq + 1 < cb) dst[HW] = v.y;
            if (4 * qq + 2 < cb) dst[2 * HW] = v.z;
            if (4 * qq + 3 < cb) dst[3 * HW] = v.w;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
static int vec_pick_cb(int C, int sms) {
    int CB = ceil_div(C, sms);
    if (CB > kVecSlots) {
        const int waves = ceil_div(ceil_div(C, kVecSlots), sms);
        CB = ceil_div(C, waves * sms);
        if (CB > kVecSlots) CB = kVecSlots;
    }
    if (CB < 1) CB = 1;
    return CB;
}

static size_t vec_fwd_smem(int H, int W, int k) {
    const int SP = vec_stage_pitch(k * k);
    return (size_t)H * vec_row_pitch(W) * sizeof(float) + (size_t)kVecFwdWarps * kVecSlots * SP * sizeof(float) +
           (size_t)kVecRChunk * k * sizeof(uint32_t) + 16;
}
static size_t vec_bwd_smem(int H, int W, int k) {
    const int SP = vec_stage_pitch(k * k);
    return (size_t)H * vec_row_pitch(W) * sizeof(float) + (size_t)2 * kVecBwdRG * kVecSlots * SP * sizeof(float) +
           (size_t)kVecRChunk * k * sizeof(uint32_t) + kVecRChunk + (size_t)2 * ((H * kVecBwdRG + 15) / 16 * 16) +
           (size_t)2 * kVecBwdRG * 32 * sizeof(uint32_t) + 16;
}
static int vec_bwd_warps(int H) {
    const int rowsPerWarp = ceil_div(H, kVecBwdMaxWarps);
    return ceil_div(H, rowsPerWarp);
}

bool roipool_vec_supported(int R, int C, int H, int W, int k) {
    if (k != 7 || R <= 0 || C <= 0 || H > 255 || W > 255) return false;
    static int off = -1;
    if (off < 0) {
        const char* e = getenv("D2T_ROIPOOL_VEC");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    if (off) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    if (vec_fwd_smem(H, W, k) > (size_t)di.max_smem_optin || vec_bwd_smem(H, W, k) > (size_t)di.max_smem_optin) return false;
    // backward staging map: two elements per thread must cover a RoI's 16*k*k floats
    int nw = vec_bwd_warps(H);
    if (nw < 13) nw = 13;
    return 4 * nw * 32 >= kVecBwdRG * 4 * k * k;  // backward staging: 4 items per thread cover a group
}

int roipool_vec_fwd_launch(const float* fm, const float* rois, float* out, int R, int C, int H, int W, int k,
                           cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int CB = vec_pick_cb(C, di.sm_count);
    const size_t smem = vec_fwd_smem(H, W, k);
    auto kern = roipool_vec_fwd_kernel<7>;
    D2T_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ceil_div(C, CB), kVecFwdThreads, smem, st>>>(fm, rois, out, R, C, H, W, CB);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

int roipool_vec_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W, int k,
                           cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int CB = vec_pick_cb(C, di.sm_count);
    const size_t smem = vec_bwd_smem(H, W, k);
    int nw = vec_bwd_warps(H);
    if (nw < 13) nw = 13;  // staging needs 4 * threads >= 8 * 4 * 49
    auto kern = roipool_vec_bwd_kernel<7>;
    D2T_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ceil_div(C, CB), nw * 32, smem, st>>>(go, rois, gin, R, C, H, W, CB);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
