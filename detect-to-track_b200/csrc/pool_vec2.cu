// pool_vec2.cu -- float32 ROIPool backward, [pixel][16 channel] difference-array slab (the scheme of pool_vec.cu) with
// the per-update instruction count halved.  sm_100a.
//
// ncu on pool_vec.cu's backward at the track-head size (C=1891, R=300, 38x63; profiles/r1_ncu_pool_v4_summary.txt and
// the per-instruction counts behind it) showed an instruction-bound kernel: 700 k warp instructions per SM, of which
//   * 400 k in the update loop = 3841 (RoI, pixel row) updates x 104 instructions: 52 of them only to walk the 64-bit
//     cover mask of the row and unpack offsets, 17 to sum the covering bin rows, 18 for the two read-modify-writes of a
//     "simple" RoI -- and 110 for the 15 % of updates whose RoI has bins thinner than a pixel (7 serialised phases);
//   * 196 k to stage grad_out: register prefetch, scale, transpose to [RoI][bin][16 ch], cover masks, offsets.
// This version keeps the data structure (row difference arrays D[y][x][16 ch] in shared memory, one warp per pixel row
// and RoI group, ascending RoI order => deterministic, no atomics; reference: atomicAdd per bin pixel,
// roipool_cuda.cu:119-125) and changes what surrounds it:
//   staging   grad_out goes to shared memory untransformed with 4-byte cp.async (the slab grad_out[r, c0:c0+cb] of a RoI
//             is one contiguous run; a thread's copy addresses are the same for every group and live in registers), as
//             [RoI][channel][50]: channel pitch 50 puts lane (bin column j, channel quad q) on bank 8q + j + const, so the
//             four scalar loads of a bin row are conflict-free.  The 1 / (bin rows x bin columns) factors go to a
//             49-entry table per RoI and are applied as FFMA while the covering bin rows are summed.
//   row lists per (group, pixel row): the RoIs of the group that cover the row, ascending, as 16-bit (RoI, cover bits)
//             entries compacted with warp ballots -- the update loop reads one entry instead of scanning a mask.
//   classes   per RoI: 0 = column edges strictly increasing (one phase), 1 = strictly increasing within the even and
//             within the odd bin columns (bins at least half a pixel wide: two phases), 2 = anything (seven phases).
//   edges     computed per group, two groups ahead, by 56 threads: no edge table, any R.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

namespace {

constexpr int V2K = 7, V2KK = 49;
constexpr int V2Slots = 16;                 // channel slots per CTA (4 quads of 4)
constexpr int V2RG = 8;                     // RoIs per staged group
constexpr int V2ChPitch = 50;               // staged floats per channel
constexpr int V2Threads = 640;
constexpr int V2Copies = (V2RG * V2Slots * V2KK + V2Threads - 1) / V2Threads;  // cp.async per thread and group
constexpr int V2EdgeSlot = 64;              // words per edge slot (56 used)

__device__ __forceinline__ int v2_pix_off(int x, int q) { return x * V2Slots + ((q ^ ((x >> 1) & 3)) << 2); }
__host__ __device__ constexpr int v2_row_pitch(int W) { return (W + 1) * V2Slots + V2Slots; }
__device__ __forceinline__ float4 v2_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void v2_st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void v2_cp_async4(uint32_t dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void v2_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void v2_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// I0 | I1<<8 | J0<<16 | J1<<24 of bin index b (row edges from H, column edges from W); reference roipool_cuda.cu:38-50
__device__ __forceinline__ uint32_t v2_pack_edges(const float* __restrict__ roi, int b, int H, int W) {
    int i0, i1, j0, j1;
    bin_edge<float, true>(roi[0], roi[2], b, V2K, H, i0, i1);
    bin_edge<float, true>(roi[1], roi[3], b, V2K, W, j0, j1);
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)j0 << 16) | ((uint32_t)j1 << 24);
}

struct V2Smem {
    size_t d, raw, inv, off, edge, list, cnt, counter, total;
};
// A RoI's staged slab holds CB channels (not 16 slots): with CB = 13 at the track-head size that leaves ~20 KB of the SM's
// shared memory free, enough for a PSROIPool CTA of another stream to be co-resident.  Lanes of dead channel slots read
// past their slab (another RoI's values or table bytes, always inside the allocation); nothing they compute is stored.
__host__ __device__ inline V2Smem v2_layout(int H, int W, int CB) {
    V2Smem s;
    size_t o = 0;
    s.d = o;       o += (size_t)H * v2_row_pitch(W) * sizeof(float);
    s.raw = o;     o += (size_t)2 * V2RG * CB * V2ChPitch * sizeof(float);
    s.inv = o;     o += (size_t)2 * V2RG * V2KK * sizeof(float);
    s.off = o;     o += (size_t)2 * V2RG * 32 * sizeof(uint32_t);
    s.edge = o;    o += (size_t)3 * V2EdgeSlot * sizeof(uint32_t);
    s.list = o;    o += (size_t)2 * H * V2RG * sizeof(uint16_t);
    s.cnt = o;     o += ((size_t)2 * H + 15) / 16 * 16;
    s.counter = o; o += 16;
    s.total = o;
    return s;
}

__global__ void __launch_bounds__(V2Threads, 1)
roipool_vec2_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R,
                        int C, int H, int W, int CB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const V2Smem L = v2_layout(H, W, CB);
    const int slab = CB * V2ChPitch;   // staged floats per RoI
    const int stage = V2RG * slab;     // floats per stage
    const int rowPitch = v2_row_pitch(W);
    float* D = reinterpret_cast<float*>(smem_raw + L.d);
    float* rawS = reinterpret_cast<float*>(smem_raw + L.raw);           // [2][RG][CB][50]
    float* invS = reinterpret_cast<float*>(smem_raw + L.inv);           // [2][RG][49]
    uint32_t* offS = reinterpret_cast<uint32_t*>(smem_raw + L.off);     // [2][RG][32]
    uint32_t* edgeG = reinterpret_cast<uint32_t*>(smem_raw + L.edge);   // [3][64]
    uint16_t* listS = reinterpret_cast<uint16_t*>(smem_raw + L.list);   // [2][H][RG]
    unsigned char* cntS = smem_raw + L.cnt;                             // [2][H]
    int* counter = reinterpret_cast<int*>(smem_raw + L.counter);        // [2]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);
    const int HW = H * W;
    const int slabN = cb * V2KK;  // floats of one RoI's slab that exist in grad_out
    const int nG = (R + V2RG - 1) / V2RG;
    const uint32_t rawAddr = (uint32_t)__cvta_generic_to_shared(rawS);

    for (int idx = tid; idx < H * rowPitch / 4; idx += V2Threads) v2_st4(D + idx * 4, make_float4(0.f, 0.f, 0.f, 0.f));

    // this thread's copies of a group: element idx = tid + n * threads of the group's RG x slabN floats
    int cpSrc[V2Copies];       // float offset from the group's first slab; -1 = none
    uint32_t cpDst[V2Copies];  // byte offset inside a stage | RoI slot << 16
#pragma unroll
    for (int n = 0; n < V2Copies; ++n) {
        const int idx = tid + n * V2Threads;
        cpSrc[n] = -1;
        cpDst[n] = 0;
        if (idx < V2RG * slabN) {
            const int rr = idx / slabN, e = idx - rr * slabN;
            cpSrc[n] = rr * C * V2KK + e;
            cpDst[n] = (uint32_t)((rr * slab + e + e / V2KK) * 4) | ((uint32_t)rr << 16);
        }
    }

    // packed edges of group g -> slot g % 3 (RoIs past the end: empty bins)
    auto edges = [&](int g) {
        if (tid < V2RG * V2K) {
            const int rr = tid / V2K, b = tid - rr * V2K;
            const int r = g * V2RG + rr;
            edgeG[(g % 3) * V2EdgeSlot + tid] = r < R ? v2_pack_edges(rois + (size_t)r * 4, b, H, W) : 0u;
        }
    };
    // raw grad_out slabs of group g -> stage buf
    auto issue = [&](int g, int buf) {
        const float* srcG = go + ((size_t)g * V2RG * C + c0) * V2KK;
        const uint32_t dstG = rawAddr + (uint32_t)buf * (uint32_t)(stage * 4);
        const int nr = min(V2RG, R - g * V2RG);
#pragma unroll
        for (int n = 0; n < V2Copies; ++n)
            if (cpSrc[n] >= 0 && (int)(cpDst[n] >> 16) < nr) v2_cp_async4(dstG + (cpDst[n] & 0xffffu), srcG + cpSrc[n]);
        v2_cp_commit();
    };
    // reciprocal bin sizes, per-lane update offsets + class, per-row RoI lists and the row queue of group g -> buf
    auto tables = [&](int g, int buf) {
        const uint32_t* ed = edgeG + (g % 3) * V2EdgeSlot;
        for (int idx = tid; idx < V2RG * V2KK; idx += V2Threads) {
            const int rr = idx / V2KK, b = idx - rr * V2KK;
            const int bi = b / V2K, bj = b - bi * V2K;
            const uint32_t ei = ed[rr * V2K + bi], ej = ed[rr * V2K + bj];
            const int hI = (int)((ei >> 8) & 255) - (int)(ei & 255);
            const int wJ = (int)(ej >> 24) - (int)((ej >> 16) & 255);
            invS[buf * (V2RG * V2KK) + idx] = (hI > 0 && wJ > 0) ? 1.0f / (float)(hI * wJ) : 0.f;
        }
        for (int idx = tid; idx < V2RG * 32; idx += V2Threads) {
            const int rr = idx >> 5, ln = idx & 31;
            const int jj = min(ln >> 2, V2K - 1), qq = ln & 3;
            const uint32_t* e = ed + rr * V2K;
            bool simple = true, semi = true;
#pragma unroll
            for (int b = 1; b < V2K; ++b) {
                const uint32_t a = e[b - 1], c = e[b];
                simple = simple && (((c >> 16) & 255) > ((a >> 16) & 255)) && ((c >> 24) > (a >> 24));
            }
#pragma unroll
            for (int b = 2; b < V2K; ++b) {
                const uint32_t a = e[b - 2], c = e[b];
                semi = semi && (((c >> 16) & 255) > ((a >> 16) & 255)) && ((c >> 24) > (a >> 24));
            }
            const uint32_t ej = e[jj];
            const int J0 = (ej >> 16) & 255, J1 = ej >> 24;
            const uint32_t cls = simple ? 0u : (semi ? 1u : 2u);
            offS[buf * (V2RG * 32) + idx] =
                (uint32_t)(v2_pix_off(J0, qq) * 4) | ((uint32_t)(v2_pix_off(J1, qq) * 4) << 14) | (cls << 28);
        }
        // entry (y, rr): cover bits = bin rows of RoI rr that contain pixel row y; the 8 entries of a row sit in 8
        // consecutive lanes and are compacted in place with a ballot
        for (int base = warp * 32; base < H * V2RG; base += V2Threads) {
            const int idx = base + lane;
            const int y = idx >> 3, rr = idx & 7;
            unsigned m = 0;
            if (idx < H * V2RG) {
#pragma unroll
                for (int b = 0; b < V2K; ++b) {
                    const uint32_t e = ed[rr * V2K + b];
                    const int i0 = e & 255, i1 = (e >> 8) & 255;
                    m |= (i0 <= y && y < i1) ? (1u << b) : 0u;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, m != 0u);
            const unsigned gb = (bal >> (lane & 24)) & 0xffu;
            if (m) listS[(buf * H + y) * V2RG + __popc(gb & ((1u << (lane & 7)) - 1u))] = (uint16_t)((rr << 8) | m);
            if ((lane & 7) == 0 && idx < H * V2RG) cntS[buf * H + y] = (unsigned char)__popc(gb);
        }
        if (tid == 0) counter[buf] = 0;
    };

    edges(0);
    if (nG > 1) edges(1);
    __syncthreads();
    tables(0, 0);
    issue(0, 0);

    const int j = lane >> 2, q = lane & 3;
    const bool jact = j < V2K;
    const int jc = jact ? j : V2K - 1;
    const int laneRaw = (4 * q) * V2ChPitch + jc;  // float offset of (channel 4q, bin column j) inside a RoI slab

    for (int g = 0; g < nG; ++g) {
        const int buf = g & 1;
        v2_cp_wait_all();  // this thread's copies of group g have landed
        __syncthreads();   // ... everyone's, and the tables of `buf`; everyone is done with group g-1
        if (g + 1 < nG) issue(g + 1, buf ^ 1);
        if (g + 2 < nG) edges(g + 2);

        const float* rawB = rawS + buf * stage + laneRaw;
        const float* invB = invS + buf * (V2RG * V2KK) + jc;
        const uint32_t* offB = offS + buf * (V2RG * 32) + lane;
        const uint16_t* listB = listS + buf * H * V2RG;
        const unsigned char* cntB = cntS + buf * H;
        while (true) {
            int task = 0;
            if (lane == 0) task = atomicAdd(&counter[buf], 1);
            task = __shfl_sync(0xffffffffu, task, 0);
            if (task >= H) break;
            // centre rows first (most RoIs cover them): tasks alternate c, c+1, c-1, c+2, ... then walk down to row 0
            const int cRow = H >> 1, U2 = 2 * (H - 1 - cRow);
            const int y = task < U2 ? ((task & 1) ? cRow + 1 + (task >> 1) : cRow - (task >> 1))
                                    : cRow - (U2 >> 1) - (task - U2);
            const int n = cntB[y];
            const uint16_t* lst = listB + y * V2RG;
            char* row = reinterpret_cast<char*>(D + y * rowPitch);
#pragma unroll 1
            for (int e = 0; e < n; ++e) {  // RoIs of the group that cover this row, ascending
                const unsigned ent = lst[e];
                const int rr = ent >> 8;
                unsigned cover = ent & 0xffu;
                const uint32_t w = offB[rr * 32];
                const float* gR = rawB + rr * slab;
                const float* iR = invB + rr * V2KK;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                do {  // bin rows containing y (1, or 2 where floor/ceil edges overlap)
                    const int i7 = (__ffs(cover) - 1) * V2K;
                    cover &= cover - 1u;
                    const float inv = iR[i7];
                    const float* p = gR + i7;
                    t.x = fmaf(p[0], inv, t.x);
                    t.y = fmaf(p[V2ChPitch], inv, t.y);
                    t.z = fmaf(p[2 * V2ChPitch], inv, t.z);
                    t.w = fmaf(p[3 * V2ChPitch], inv, t.w);
                } while (cover);
                float* pA = reinterpret_cast<float*>(row + (w & 0x3fffu));
                float* pB = reinterpret_cast<float*>(row + ((w >> 14) & 0x3fffu));
                const unsigned cls = w >> 28;
                if (cls == 0u) {
                    if (jact) {
                        float4 a = v2_ld4(pA);
                        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                        v2_st4(pA, a);
                    }
                    __syncwarp();
                    if (jact) {
                        float4 b = v2_ld4(pB);
                        b.x -= t.x; b.y -= t.y; b.z -= t.z; b.w -= t.w;
                        v2_st4(pB, b);
                    }
                    __syncwarp();
                } else {
                    // bins thinner than a pixel: lanes of different bin columns may address the same pixel.  Class 1:
                    // even and odd bin columns in turn; class 2: one bin column at a time.
                    const int phases = cls == 1u ? 2 : V2K;
                    for (int ph = 0; ph < phases; ++ph) {
                        const bool sel = jact && (cls == 1u ? ((j & 1) == ph) : (j == ph));
                        if (sel) {
                            float4 a = v2_ld4(pA);
                            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                            v2_st4(pA, a);
                        }
                        __syncwarp();
                        if (sel) {
                            float4 b = v2_ld4(pB);
                            b.x -= t.x; b.y -= t.y; b.z -= t.z; b.w -= t.w;
                            v2_st4(pB, b);
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (g + 1 < nG) tables(g + 1, buf ^ 1);
    }
    __syncthreads();

    // ---- epilogue: inclusive row scan, then transposed write-out (LDS.128 -> 4 coalesced plane stores) -------
    for (int t = tid; t < H * 4; t += V2Threads) {
        const int y = t >> 2, qq = t & 3;
        float* row = D + y * rowPitch;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int x = 0; x < W; ++x) {
            float* p = row + v2_pix_off(x, qq);
            const float4 v = v2_ld4(p);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
            v2_st4(p, acc);
        }
    }
    __syncthreads();
    {
        const int total = 4 * HW;
        for (int idx = tid; idx < total; idx += V2Threads) {
            const int qq = idx / HW, pix = idx - qq * HW;
            const int y = pix / W, x = pix - y * W;
            const float4 v = v2_ld4(D + y * rowPitch + v2_pix_off(x, qq));
            float* dst = gin + (size_t)(c0 + 4 * qq) * HW + pix;
            if (4 * qq + 0 < cb) dst[0] = v.x;
            if (4 * qq + 1 < cb) dst[HW] = v.y;
            if (4 * qq + 2 < cb) dst[2 * HW] = v.z;
            if (4 * qq + 3 < cb) dst[3 * HW] = v.w;
        }
    }
}

}  // namespace

bool roipool_vec2_bwd_supported(int R, int C, int H, int W, int k) {
    if (k != V2K || R <= 0 || C <= 0 || H <= 0 || W <= 0 || H > 255 || W > 254) return false;
    if ((long long)C * V2KK * V2RG >= (1ll << 31)) return false;  // copy offsets are ints
    DeviceInfo di;
    if (device_info(&di)) return false;
    return v2_layout(H, W, V2Slots).total <= (size_t)di.max_smem_optin;
}

int roipool_vec2_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W,
                            cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    // channels per CTA: one wave of CTAs where possible, at most 16 channel slots each
    int CB = ceil_div(C, di.sm_count);
    if (CB > V2Slots) {
        const int waves = ceil_div(ceil_div(C, V2Slots), di.sm_count);
        CB = ceil_div(C, waves * di.sm_count);
        if (CB > V2Slots) CB = V2Slots;
    }
    if (CB < 1) CB = 1;
    const size_t smem = v2_layout(H, W, CB).total;
    D2T_SMEM_OPTIN(roipool_vec2_bwd_kernel, smem);
    roipool_vec2_bwd_kernel<<<ceil_div(C, CB), V2Threads, smem, st>>>(go, rois, gin, R, C, H, W, CB);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
