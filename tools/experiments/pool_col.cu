// pool_col.cu -- float32 ROIPool backward, fourth generation: pixel-COLUMN owners with register accumulators.  sm_100a.
//
// Why: the third-generation kernel (pool_vec.cu) keeps grad_fm of a 16-channel slab in shared memory as row difference
// arrays and pays, per (RoI, pixel row), two dependent LDS.128 -> FADD -> STS.128 round trips plus the bookkeeping that
// keeps them ordered (cover masks, row queues, a CTA barrier per RoI group): 700 k warp instructions and 156 k
// shared-memory wavefronts per SM for 20 us worth of HBM traffic (profiles/r1_ncu_pool_v4_summary.txt).  Here nothing
// of grad_fm lives in shared memory:
//
//   thread    owns one pixel column x of the slab for four channels and ALL rows: acc[y][4] in registers (H <= 38 ->
//             152 accumulators).  A warp is a strip of 8 columns x 4 channel quads.
//   per RoI   (ascending, so every accumulator sees a fixed order: deterministic, no atomics; reference: atomicAdd per
//             bin pixel, roipool_cuda.cu:119-125)  a warp skips the RoI if its columns miss the strip (one packed
//             word); otherwise each lane finds the bin column(s) j that contain its x -- one, or two where the
//             floor/ceil edges of neighbours overlap -- and for each bin row i forms
//                 v = sum_j grad_out[r, c, i, j] / ((I1_i - I0_i) * (J1_j - J0_j))
//             and adds v to acc[I0_i .. I1_i - 1].  Registers cannot be indexed dynamically, so the row range is
//             entered through a jump table into an unrolled ladder (switch with fall-through).
//   grad_out  reaches shared memory untransformed: the slab of one RoI, grad_out[r, c0:c0+cb, :, :], is one contiguous
//             run, copied with 4-byte cp.async into a 3-stage ring of 8-RoI groups (no registers held, no transpose
//             pass); lanes read it with scalar LDS (bank = 4 * quad + bin: conflict-free, same (quad, bin) broadcast).
//   output    every grad_fm element of the slab is written exactly once from registers (no memset, no scan), and the
//             sums are plain ascending-RoI sums: no difference-array cancellation.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

namespace {

constexpr int kColK = 7, kColKK = 49;
constexpr int kColSlots = 16;                      // channel slots per CTA (4 quads of 4)
constexpr int kColThreads = 256;                   // 8 strips of 8 columns x 4 quads
constexpr int kColRG = 8;                          // RoIs per staged group
constexpr int kColStages = 3;
constexpr int kColRChunk = 1024;                   // RoIs per edge-table chunk
constexpr int kColSlab = kColSlots * kColKK;       // floats of one RoI's staged slab
constexpr int kColHMax = 38;

__device__ __forceinline__ void col_cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void col_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void col_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// I0 | I1<<8 | J0<<16 | J1<<24 of bin index b (row edges from H, column edges from W); reference roipool_cuda.cu:38-50
__device__ __forceinline__ uint32_t col_pack_edges(const float* __restrict__ roi, int b, int H, int W) {
    int i0, i1, j0, j1;
    bin_edge<float, true>(roi[0], roi[2], b, kColK, H, i0, i1);
    bin_edge<float, true>(roi[1], roi[3], b, kColK, W, j0, j1);
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)j0 << 16) | ((uint32_t)j1 << 24);
}

#define D2T_COL_ROW(Y)                       \
    case (Y):                                \
        acc[(Y)].x += v.x;                   \
        acc[(Y)].y += v.y;                   \
        acc[(Y)].z += v.z;                   \
        acc[(Y)].w += v.w;                   \
        if (I1 <= (Y) + 1) break;
#define D2T_COL_ROW2(Y) D2T_COL_ROW(Y) D2T_COL_ROW((Y) + 1)
#define D2T_COL_ROW8(Y) D2T_COL_ROW2(Y) D2T_COL_ROW2((Y) + 2) D2T_COL_ROW2((Y) + 4) D2T_COL_ROW2((Y) + 6)

__global__ void __launch_bounds__(kColThreads, 1)
roipool_col_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R,
                       int C, int H, int W, int CB) {
    static_assert(kColHMax == 38, "the ladder below has 38 rungs");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* rawS = reinterpret_cast<float*>(smem_raw);                         // [stages][RG][16*49]
    float* invS = rawS + kColStages * kColRG * kColSlab;                      // [stages][RG][49]
    uint32_t* edgeS = reinterpret_cast<uint32_t*>(invS + kColStages * kColRG * kColKK);  // [RCH][7]
    uint32_t* extS = edgeS + kColRChunk * kColK;                              // [RCH] column extent Jmin | Jmax << 8

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: keeps the ladder's branches uniform
    const int q = lane & 3;
    const int x0 = warp * 8;
    const int x = x0 + (lane >> 2);
    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);
    const int slabN = cb * kColKK;  // floats of one RoI's slab that exist in grad_out

    float4 acc[kColHMax];
#pragma unroll
    for (int y = 0; y < kColHMax; ++y) acc[y] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int rc0 = 0; rc0 < R; rc0 += kColRChunk) {
        const int nrc = min(kColRChunk, R - rc0);
        const int nGroups = (nrc + kColRG - 1) / kColRG;
        __syncthreads();  // the previous chunk is consumed
        for (int idx = tid; idx < nrc * kColK; idx += kColThreads) {
            const int rr = idx / kColK, b = idx - rr * kColK;
            edgeS[idx] = col_pack_edges(rois + (size_t)(rc0 + rr) * 4, b, H, W);
        }
        __syncthreads();
        for (int rr = tid; rr < nrc; rr += kColThreads) {
            int jmin = 255, jmax = 0;
#pragma unroll
            for (int b = 0; b < kColK; ++b) {
                const uint32_t e = edgeS[rr * kColK + b];
                const int j0 = (e >> 16) & 255, j1 = e >> 24;
                if (j1 > j0) {
                    jmin = min(jmin, j0);
                    jmax = max(jmax, j1);
                }
            }
            extS[rr] = (uint32_t)jmin | ((uint32_t)jmax << 8);
        }
        __syncthreads();

        // stage group g: raw grad_out slabs (cp.async) and the reciprocal bin sizes
        auto issue = [&](int g) {
            if (g < nGroups) {
                const int s = g % kColStages;
                const int nr = min(kColRG, nrc - g * kColRG);
                float* dstG = rawS + s * (kColRG * kColSlab);
                const float* srcG = go + ((size_t)(rc0 + g * kColRG) * C + c0) * kColKK;
                for (int rr = 0; rr < nr; ++rr) {
                    const float* src = srcG + (size_t)rr * C * kColKK;
                    float* dst = dstG + rr * kColSlab;
                    for (int e = tid; e < slabN; e += kColThreads) col_cp_async4(dst + e, src + e);
                }
                float* invG = invS + s * (kColRG * kColKK);
                for (int idx = tid; idx < nr * kColKK; idx += kColThreads) {
                    const int rr = idx / kColKK, b = idx - rr * kColKK;
                    const int bi = b / kColK, bj = b - bi * kColK;
                    const uint32_t* ed = edgeS + (g * kColRG + rr) * kColK;
                    const uint32_t ei = ed[bi], ej = ed[bj];
                    const int hI = (int)((ei >> 8) & 255) - (int)(ei & 255);
                    const int wJ = (int)(ej >> 24) - (int)((ej >> 16) & 255);
                    invG[idx] = (hI > 0 && wJ > 0) ? 1.0f / (float)(hI * wJ) : 0.f;
                }
            }
            col_cp_commit();
        };
        issue(0);
        issue(1);

        for (int g = 0; g < nGroups; ++g) {
            col_cp_wait<1>();  // this thread's copies of group g have landed (group g+1 may still be in flight)
            __syncthreads();   // ... and everyone else's; everyone is done with group g-1, whose stage is refilled next
            issue(g + 2);
            const int s = g % kColStages;
            const int nr = min(kColRG, nrc - g * kColRG);
            const float* rawG = rawS + s * (kColRG * kColSlab) + (4 * q) * kColKK;
            const float* invG = invS + s * (kColRG * kColKK);
#pragma unroll 1
            for (int rr = 0; rr < nr; ++rr) {
                const int rIdx = g * kColRG + rr;
                const uint32_t ext = extS[rIdx];
                if (x0 + 8 <= (int)(ext & 255) || x0 >= (int)((ext >> 8) & 255)) continue;  // the RoI misses this strip
                const uint32_t* ed = edgeS + rIdx * kColK;
                unsigned mask = 0;  // bin columns that contain x
#pragma unroll
                for (int j = 0; j < kColK; ++j) {
                    const uint32_t e = ed[j];
                    const int j0 = (e >> 16) & 255, j1 = e >> 24;
                    mask |= (x >= j0 && x < j1) ? (1u << j) : 0u;
                }
                const unsigned m2 = mask & (mask - 1u);
                const unsigned rest = m2 & (m2 - 1u);
                const int jA = mask ? __ffs(mask) - 1 : 0;
                const int jB = m2 ? __ffs(m2) - 1 : 0;
                const bool anyRest = __any_sync(0xffffffffu, rest != 0u);  // > 2 bin columns on one pixel: sub-pixel bins
                const float* rawR = rawG + rr * kColSlab;
                const float* invR = invG + rr * kColKK;
                // Software pipeline over the bin rows: the shared-memory loads of bin row i+1 are issued before the
                // ladder of bin row i and consumed after it, so their latency hides behind the accumulator updates.
                float gA0 = 0.f, gA1 = 0.f, gA2 = 0.f, gA3 = 0.f, iA = 0.f, gB0 = 0.f, gB1 = 0.f, gB2 = 0.f, gB3 = 0.f, iB = 0.f;
                uint32_t eNext = ed[0];
                auto load_row = [&](int i) {
                    if (mask) {
                        const int b = i * kColK + jA;
                        iA = invR[b];
                        gA0 = rawR[b];
                        gA1 = rawR[kColKK + b];
                        gA2 = rawR[2 * kColKK + b];
                        gA3 = rawR[3 * kColKK + b];
                    }
                    if (m2) {
                        const int b = i * kColK + jB;
                        iB = invR[b];
                        gB0 = rawR[b];
                        gB1 = rawR[kColKK + b];
                        gB2 = rawR[2 * kColKK + b];
                        gB3 = rawR[3 * kColKK + b];
                    }
                };
                load_row(0);
#pragma unroll 1
                for (int i = 0; i < kColK; ++i) {
                    const uint32_t ei = eNext;
                    const int I0 = ei & 255, I1 = (ei >> 8) & 255;
                    // lanes outside every bin column keep g = 0, inv = 0: v = 0
                    float4 v = make_float4(gA0 * iA, gA1 * iA, gA2 * iA, gA3 * iA);
                    if (m2) {
                        v.x += gB0 * iB;
                        v.y += gB1 * iB;
                        v.z += gB2 * iB;
                        v.w += gB3 * iB;
                    }
                    if (anyRest) {
                        unsigned m = rest;
                        while (m) {
                            const int b = i * kColK + __ffs(m) - 1;
                            m &= m - 1u;
                            const float inv = invR[b];
                            v.x += rawR[b] * inv;
                            v.y += rawR[kColKK + b] * inv;
                            v.z += rawR[2 * kColKK + b] * inv;
                            v.w += rawR[3 * kColKK + b] * inv;
                        }
                    }
                    if (i + 1 < kColK) {
                        eNext = ed[i + 1];
                        load_row(i + 1);
                    }
                    if (I1 <= I0) continue;
                    // acc[I0 .. I1-1] += v
                    switch (I0) {
                        D2T_COL_ROW8(0)
                        D2T_COL_ROW8(8)
                        D2T_COL_ROW8(16)
                        D2T_COL_ROW8(24)
                        D2T_COL_ROW2(32)
                        D2T_COL_ROW2(34)
                        D2T_COL_ROW2(36)
                        default:
                            break;
                    }
                }
            }
        }
        col_cp_wait<0>();
    }

    // ---- write-out: a warp store covers 4 channel planes x 8 consecutive pixels ---------------------------------
    if (x < W) {
        const size_t HW = (size_t)H * W;
        float* dst = gin + (size_t)(c0 + 4 * q) * HW + x;
        const int nch = cb - 4 * q;  // live channels of this quad
#pragma unroll
        for (int y = 0; y < kColHMax; ++y) {
            if (y < H) {
                float* p = dst + (size_t)y * W;
                if (nch > 0) p[0] = acc[y].x;
                if (nch > 1) p[HW] = acc[y].y;
                if (nch > 2) p[2 * HW] = acc[y].z;
                if (nch > 3) p[3 * HW] = acc[y].w;
            }
        }
    }
}

#undef D2T_COL_ROW8
#undef D2T_COL_ROW2
#undef D2T_COL_ROW

size_t col_smem_bytes() {
    return (size_t)kColStages * kColRG * kColSlab * sizeof(float) + (size_t)kColStages * kColRG * kColKK * sizeof(float) +
           (size_t)kColRChunk * kColK * sizeof(uint32_t) + (size_t)kColRChunk * sizeof(uint32_t);
}

}  // namespace

bool roipool_col_bwd_supported(int R, int C, int H, int W, int k) {
    if (k != kColK || R <= 0 || C <= 0 || H <= 0 || W <= 0 || H > kColHMax || W > 64) return false;
    // opt-in (D2T_ROIPOOL_BWD=col): correct and reproducible but measured 446 us against 168 us for pool_vec.cu at the
    // track-head size -- ~700 instructions per (RoI, strip) visit (edge scan, per-bin-row gathers, ladder dispatch) on 8
    // warps per SM issue at 25 %, and a third of the time is the group barrier (edge strips idle)
    const char* e = getenv("D2T_ROIPOOL_BWD");
    if (!(e && e[0] == 'c')) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    return col_smem_bytes() <= (size_t)di.max_smem_optin;
}

int roipool_col_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W,
                           cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    // channels per CTA: one wave of CTAs, at most 16 channel slots each
    int CB = ceil_div(C, di.sm_count);
    if (CB > kColSlots) {
        const int waves = ceil_div(ceil_div(C, kColSlots), di.sm_count);
        CB = ceil_div(C, waves * di.sm_count);
        if (CB > kColSlots) CB = kColSlots;
    }
    if (CB < 1) CB = 1;
    const size_t smem = col_smem_bytes();
    D2T_CUDA_TRY(cudaFuncSetAttribute(roipool_col_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roipool_col_bwd_kernel<<<ceil_div(C, CB), kColThreads, smem, st>>>(go, rois, gin, R, C, H, W, CB);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
