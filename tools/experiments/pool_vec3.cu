// pool_vec3.cu -- float32 ROIPool backward: the difference-array slab of pool_vec2.cu at TWO CTAs per SM.  sm_100a.
//
// pool_vec2.cu runs one CTA of 20 warps per SM (a 16-channel slab needs 158 KB of shared memory) and sits at 53 % issue
// with five warps per scheduler: each warp is a chain of short shared-memory round trips.  Here a CTA owns an 8-channel
// slab (80 KB), so two CTAs are resident, and a warp works on TWO pixel rows at once: lanes 0-15 walk the RoI list of
// row A, lanes 16-31 that of row B (a half-warp = 7 bin columns x 2 channel quads), in lock step.  The instruction
// stream serves the same number of lanes as before, but the SM holds 32 warps and a warp has one task per RoI group
// instead of two.  Everything else (raw cp.async staging, 1/numel table, ballot-compacted row lists, update classes,
// per-group edges, ascending RoI order per row => deterministic, no atomics on grad_fm) is pool_vec2.cu's.
// MEASURED: 215 us against 156 us for pool_vec2.cu at the track-head size (C=1891, R=300, 38x63) -- opt-in experiment
// (D2T_ROIPOOL_BWD=v3).  With half the channels per CTA the SM executes as many row-pair iterations as pool_vec2.cu
// executes row updates, each with the extra votes / shuffles of the lock step, and builds every per-group table twice.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

namespace {

constexpr int V3K = 7, V3KK = 49;
constexpr int V3Slots = 8;                  // channel slots per CTA (2 quads of 4)
constexpr int V3RG = 8;                     // RoIs per staged group
constexpr int V3ChPitch = 50;               // staged floats per channel
constexpr int V3Threads = 512;
constexpr int V3Copies = (V3RG * V3Slots * V3KK + V3Threads - 1) / V3Threads;  // cp.async per thread and group
constexpr int V3EdgeSlot = 64;              // words per edge slot (56 used)

__device__ __forceinline__ int v3_pix_off(int x, int q) { return x * V3Slots + ((q ^ ((x >> 2) & 1)) << 2); }
__host__ __device__ constexpr int v3_row_pitch(int W) { return (W + 1) * V3Slots + 16; }  // + 64 B: rows at odd distance in opposite bank halves
__device__ __forceinline__ float4 v3_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void v3_st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void v3_cp_async4(uint32_t dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void v3_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void v3_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// I0 | I1<<8 | J0<<16 | J1<<24 of bin index b (row edges from H, column edges from W); reference roipool_cuda.cu:38-50
__device__ __forceinline__ uint32_t v3_pack_edges(const float* __restrict__ roi, int b, int H, int W) {
    int i0, i1, j0, j1;
    bin_edge<float, true>(roi[0], roi[2], b, V3K, H, i0, i1);
    bin_edge<float, true>(roi[1], roi[3], b, V3K, W, j0, j1);
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)j0 << 16) | ((uint32_t)j1 << 24);
}

struct V3Smem {
    size_t d, raw, inv, off, edge, list, cnt, counter, total;
};
// A RoI's staged slab holds CB channels (not 8 slots).  Lanes of dead channel slots read past their slab (another RoI's
// values or table bytes, always inside the allocation); nothing they compute is stored.
__host__ __device__ inline V3Smem v3_layout(int H, int W, int CB) {
    V3Smem s;
    size_t o = 0;
    s.d = o;       o += (size_t)H * v3_row_pitch(W) * sizeof(float);
    s.raw = o;     o += (size_t)2 * V3RG * CB * V3ChPitch * sizeof(float);
    s.inv = o;     o += (size_t)2 * V3RG * V3KK * sizeof(float);
    s.off = o;     o += (size_t)2 * V3RG * 16 * sizeof(uint32_t);
    s.edge = o;    o += (size_t)3 * V3EdgeSlot * sizeof(uint32_t);
    s.list = o;    o += (size_t)2 * H * V3RG * sizeof(uint16_t);
    s.cnt = o;     o += ((size_t)2 * H + 15) / 16 * 16;
    s.counter = o; o += 16;
    s.total = o;
    return s;
}

__global__ void __launch_bounds__(V3Threads, 2)
roipool_vec3_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R,
                        int C, int H, int W, int CB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const V3Smem L = v3_layout(H, W, CB);
    const int slab = CB * V3ChPitch;   // staged floats per RoI
    const int stage = V3RG * slab;     // floats per stage
    const int rowPitch = v3_row_pitch(W);
    float* D = reinterpret_cast<float*>(smem_raw + L.d);
    float* rawS = reinterpret_cast<float*>(smem_raw + L.raw);           // [2][RG][CB][50]
    float* invS = reinterpret_cast<float*>(smem_raw + L.inv);           // [2][RG][49]
    uint32_t* offS = reinterpret_cast<uint32_t*>(smem_raw + L.off);     // [2][RG][16]
    uint32_t* edgeG = reinterpret_cast<uint32_t*>(smem_raw + L.edge);   // [3][64]
    uint16_t* listS = reinterpret_cast<uint16_t*>(smem_raw + L.list);   // [2][H][RG]
    unsigned char* cntS = smem_raw + L.cnt;                             // [2][H]
    int* counter = reinterpret_cast<int*>(smem_raw + L.counter);        // [2]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);
    const int HW = H * W;
    const int slabN = cb * V3KK;  // floats of one RoI's slab that exist in grad_out
    const int nG = (R + V3RG - 1) / V3RG;
    const uint32_t rawAddr = (uint32_t)__cvta_generic_to_shared(rawS);

    for (int idx = tid; idx < H * rowPitch / 4; idx += V3Threads) v3_st4(D + idx * 4, make_float4(0.f, 0.f, 0.f, 0.f));

    // this thread's copies of a group: element idx = tid + n * threads of the group's RG x slabN floats
    int cpSrc[V3Copies];       // float offset from the group's first slab; -1 = none
    uint32_t cpDst[V3Copies];  // byte offset inside a stage | RoI slot << 16
#pragma unroll
    for (int n = 0; n < V3Copies; ++n) {
        const int idx = tid + n * V3Threads;
        cpSrc[n] = -1;
        cpDst[n] = 0;
        if (idx < V3RG * slabN) {
            const int rr = idx / slabN, e = idx - rr * slabN;
            cpSrc[n] = rr * C * V3KK + e;
            cpDst[n] = (uint32_t)((rr * slab + e + e / V3KK) * 4) | ((uint32_t)rr << 16);
        }
    }

    // packed edges of group g -> slot g % 3 (RoIs past the end: empty bins)
    auto edges = [&](int g) {
        if (tid < V3RG * V3K) {
            const int rr = tid / V3K, b = tid - rr * V3K;
            const int r = g * V3RG + rr;
            edgeG[(g % 3) * V3EdgeSlot + tid] = r < R ? v3_pack_edges(rois + (size_t)r * 4, b, H, W) : 0u;
        }
    };
    // raw grad_out slabs of group g -> stage buf
    auto issue = [&](int g, int buf) {
        const float* srcG = go + ((size_t)g * V3RG * C + c0) * V3KK;
        const uint32_t dstG = rawAddr + (uint32_t)buf * (uint32_t)(stage * 4);
        const int nr = min(V3RG, R - g * V3RG);
#pragma unroll
        for (int n = 0; n < V3Copies; ++n)
            if (cpSrc[n] >= 0 && (int)(cpDst[n] >> 16) < nr) v3_cp_async4(dstG + (cpDst[n] & 0xffffu), srcG + cpSrc[n]);
        v3_cp_commit();
    };
    // reciprocal bin sizes, per-lane update offsets + class, per-row RoI lists and the row queue of group g -> buf
    auto tables = [&](int g, int buf) {
        const uint32_t* ed = edgeG + (g % 3) * V3EdgeSlot;
        for (int idx = tid; idx < V3RG * V3KK; idx += V3Threads) {
            const int rr = idx / V3KK, b = idx - rr * V3KK;
            const int bi = b / V3K, bj = b - bi * V3K;
            const uint32_t ei = ed[rr * V3K + bi], ej = ed[rr * V3K + bj];
            const int hI = (int)((ei >> 8) & 255) - (int)(ei & 255);
            const int wJ = (int)(ej >> 24) - (int)((ej >> 16) & 255);
            invS[buf * (V3RG * V3KK) + idx] = (hI > 0 && wJ > 0) ? 1.0f / (float)(hI * wJ) : 0.f;
        }
        for (int idx = tid; idx < V3RG * 16; idx += V3Threads) {
            const int rr = idx >> 4, ln = idx & 15;
            const int jj = min(ln >> 1, V3K - 1), qq = ln & 1;
            const uint32_t* e = ed + rr * V3K;
            bool simple = true, semi = true;
#pragma unroll
            for (int b = 1; b < V3K; ++b) {
                const uint32_t a = e[b - 1], c = e[b];
                simple = simple && (((c >> 16) & 255) > ((a >> 16) & 255)) && ((c >> 24) > (a >> 24));
            }
#pragma unroll
            for (int b = 2; b < V3K; ++b) {
                const uint32_t a = e[b - 2], c = e[b];
                semi = semi && (((c >> 16) & 255) > ((a >> 16) & 255)) && ((c >> 24) > (a >> 24));
            }
            const uint32_t ej = e[jj];
            const int J0 = (ej >> 16) & 255, J1 = ej >> 24;
            const uint32_t cls = simple ? 0u : (semi ? 1u : 2u);
            offS[buf * (V3RG * 16) + idx] =
                (uint32_t)(v3_pix_off(J0, qq) * 4) | ((uint32_t)(v3_pix_off(J1, qq) * 4) << 14) | (cls << 28);
        }
        // entry (y, rr): cover bits = bin rows of RoI rr that contain pixel row y; the 8 entries of a row sit in 8
        // consecutive lanes and are compacted in place with a ballot
        for (int base = warp * 32; base < H * V3RG; base += V3Threads) {
            const int idx = base + lane;
            const int y = idx >> 3, rr = idx & 7;
            unsigned m = 0;
            if (idx < H * V3RG) {
#pragma unroll
                for (int b = 0; b < V3K; ++b) {
                    const uint32_t e = ed[rr * V3K + b];
                    const int i0 = e & 255, i1 = (e >> 8) & 255;
                    m |= (i0 <= y && y < i1) ? (1u << b) : 0u;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, m != 0u);
            const unsigned gb = (bal >> (lane & 24)) & 0xffu;
            if (m) listS[(buf * H + y) * V3RG + __popc(gb & ((1u << (lane & 7)) - 1u))] = (uint16_t)((rr << 8) | m);
            if ((lane & 7) == 0 && idx < H * V3RG) cntS[buf * H + y] = (unsigned char)__popc(gb);
        }
        if (tid == 0) counter[buf] = 0;
    };

    edges(0);
    if (nG > 1) edges(1);
    __syncthreads();
    tables(0, 0);
    issue(0, 0);

    const int half = lane >> 4, l16 = lane & 15;
    const int j = l16 >> 1, q = l16 & 1;
    const bool jact = j < V3K;
    const int jc = jact ? j : V3K - 1;
    const int laneRaw = (4 * q) * V3ChPitch + jc;  // float offset of (channel 4q, bin column j) inside a RoI slab
    const int cRow = H >> 1, U2 = 2 * (H - 1 - cRow);

    for (int g = 0; g < nG; ++g) {
        const int buf = g & 1;
        v3_cp_wait_all();  // this thread's copies of group g have landed
        __syncthreads();   // ... everyone's, and the tables of `buf`; everyone is done with group g-1
        if (g + 1 < nG) issue(g + 1, buf ^ 1);
        if (g + 2 < nG) edges(g + 2);

        const float* rawB = rawS + buf * stage + laneRaw;
        const float* invB = invS + buf * (V3RG * V3KK) + jc;
        const uint32_t* offB = offS + buf * (V3RG * 16) + l16;
        const uint16_t* listB = listS + buf * H * V3RG;
        const unsigned char* cntB = cntS + buf * H;
        while (true) {
            // a warp claims two consecutive tasks = the two rows at the same distance from the centre row (their lists
            // are about equally long); half-warp h works on task + h
            int task = 0;
            if (lane == 0) task = atomicAdd(&counter[buf], 2);
            task = __shfl_sync(0xffffffffu, task, 0);
            if (task >= H) break;
            const int tk = task + half;
            const bool rowOk = tk < H;
            const int tkc = rowOk ? tk : task;
            // centre rows first (most RoIs cover them): tasks alternate c, c+1, c-1, c+2, ... then walk down to row 0
            const int y = tkc < U2 ? ((tkc & 1) ? cRow + 1 + (tkc >> 1) : cRow - (tkc >> 1)) : cRow - (U2 >> 1) - (tkc - U2);
            const int n = rowOk ? (int)cntB[y] : 0;
            const int nMax = max(__shfl_sync(0xffffffffu, n, 0), __shfl_sync(0xffffffffu, n, 16));
            const uint16_t* lst = listB + y * V3RG;
            char* row = reinterpret_cast<char*>(D + y * rowPitch);
#pragma unroll 1
            for (int e = 0; e < nMax; ++e) {  // RoIs of the group that cover the rows, ascending per row
                const bool act = e < n;
                const unsigned ent = act ? (unsigned)lst[e] : 0u;
                const int rr = ent >> 8;
                unsigned cover = ent & 0xffu;
                const uint32_t w = act ? offB[rr * 16] : 0u;
                const float* gR = rawB + rr * slab;
                const float* iR = invB + rr * V3KK;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                while (__any_sync(0xffffffffu, cover != 0u)) {  // bin rows containing the row (1, or 2 at overlaps)
                    if (cover) {
                        const int i7 = (__ffs(cover) - 1) * V3K;
                        cover &= cover - 1u;
                        const float inv = iR[i7];
                        const float* p = gR + i7;
                        t.x = fmaf(p[0], inv, t.x);
                        t.y = fmaf(p[V3ChPitch], inv, t.y);
                        t.z = fmaf(p[2 * V3ChPitch], inv, t.z);
                        t.w = fmaf(p[3 * V3ChPitch], inv, t.w);
                    }
                }
                float* pA = reinterpret_cast<float*>(row + (w & 0x3fffu));
                float* pB = reinterpret_cast<float*>(row + ((w >> 14) & 0x3fffu));
                const unsigned cls = w >> 28;
                // phases: class 0 one, class 1 two (even / odd bin columns), class 2 seven (one bin column at a time);
                // the warp runs as many as its two halves need
                const unsigned clsMax = max(__shfl_sync(0xffffffffu, cls, 0), __shfl_sync(0xffffffffu, cls, 16));
                const int phases = clsMax == 0u ? 1 : (clsMax == 1u ? 2 : V3K);
                for (int ph = 0; ph < phases; ++ph) {
                    const bool sel = act && jact && (cls == 0u ? ph == 0 : (cls == 1u ? (ph < 2 && (j & 1) == ph) : j == ph));
                    if (sel) {
                        float4 a = v3_ld4(pA);
                        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                        v3_st4(pA, a);
                    }
                    __syncwarp();
                    if (sel) {
                        float4 b = v3_ld4(pB);
                        b.x -= t.x; b.y -= t.y; b.z -= t.z; b.w -= t.w;
                        v3_st4(pB, b);
                    }
                    __syncwarp();
                }
            }
        }
        if (g + 1 < nG) tables(g + 1, buf ^ 1);
    }
    __syncthreads();

    // ---- epilogue: inclusive row scan, then transposed write-out (LDS.128 -> 4 coalesced plane stores) -------
    for (int t = tid; t < H * 2; t += V3Threads) {
        const int y = t >> 1, qq = t & 1;
        float* row = D + y * rowPitch;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int x = 0; x < W; ++x) {
            float* p = row + v3_pix_off(x, qq);
            const float4 v = v3_ld4(p);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
            v3_st4(p, acc);
        }
    }
    __syncthreads();
    {
        const int total = 2 * HW;
        for (int idx = tid; idx < total; idx += V3Threads) {
            const int qq = idx / HW, pix = idx - qq * HW;
            const int y = pix / W, x = pix - y * W;
            const float4 v = v3_ld4(D + y * rowPitch + v3_pix_off(x, qq));
            float* dst = gin + (size_t)(c0 + 4 * qq) * HW + pix;
            if (4 * qq + 0 < cb) dst[0] = v.x;
            if (4 * qq + 1 < cb) dst[HW] = v.y;
            if (4 * qq + 2 < cb) dst[2 * HW] = v.z;
            if (4 * qq + 3 < cb) dst[3 * HW] = v.w;
        }
    }
}

}  // namespace

bool roipool_vec3_bwd_supported(int R, int C, int H, int W, int k) {
    if (k != V3K || R <= 0 || C <= 0 || H <= 0 || W <= 0 || H > 255 || W > 254) return false;
    if ((long long)C * V3KK * V3RG >= (1ll << 31)) return false;  // copy offsets are ints
    const char* e = getenv("D2T_ROIPOOL_BWD");  // opt-in: D2T_ROIPOOL_BWD=v3
    if (!(e && e[0] == 'v' && e[1] == '3')) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    return v3_layout(H, W, V3Slots).total <= (size_t)di.max_smem_optin;
}

int roipool_vec3_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W,
                            cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    // channels per CTA: two CTAs per SM, one wave of CTAs where possible, at most 8 channel slots each
    const int seats = 2 * di.sm_count;
    int CB = ceil_div(C, seats);
    if (CB > V3Slots) {
        const int waves = ceil_div(ceil_div(C, V3Slots), seats);
        CB = ceil_div(C, waves * seats);
        if (CB > V3Slots) CB = V3Slots;
    }
    if (CB < 1) CB = 1;
    const size_t smem = v3_layout(H, W, CB).total;
    D2T_CUDA_TRY(cudaFuncSetAttribute(roipool_vec3_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roipool_vec3_bwd_kernel<<<ceil_div(C, CB), V3Threads, smem, st>>>(go, rois, gin, R, C, H, W, CB);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
