// pool_rows.cu -- float32 ROIPool backward, fourth generation: lanes own PIXEL ROWS.  sm_100a, k = 7, H <= 64.
//
// The adjoint of average RoI pooling in the x-difference domain is, per RoI r, channel c and pixel row y of the RoI,
//     D[y][J0_j] += t_j ,  D[y][J1_j] -= t_j ,   t_j = sum_{i : y in rows(i)} grad_out[r,c,i,j] / numel_ij        (j < 7)
// followed by an inclusive scan of D along x (reference: atomicAdd per bin pixel, roipool_cuda.cu:115-125).  The third-
// generation kernel (pool_vec.cu) put the bin columns j on the lanes, so one (RoI, pixel row) pair cost a whole warp
// iteration (~120 instructions, two dependent read-modify-writes fenced by __syncwarp) and a CTA barrier per 8 RoIs:
// 171 us at the D&T track-head size, held by latency (ncu: issue 56 %, 19 warps per SM).  This kernel is an EXPERIMENT
// (D2T_ROIPOOL_ROWS=1; correct, deterministic, but 209 us -- see roipool_rows_bwd_supported).  Here
//
//   slab     a CTA owns 8 consecutive channels for ALL RoIs: D[x][y][8 ch] in shared memory (x in [0, W], 32 bytes per
//            pixel; x-pitch 8H+4 floats so that the transposing write-out is conflict-free), 2 CTAs per SM.
//   apply    warps 0-3 each own a BAND of pixel rows (H/4 each, <= 16).  Lane (y', h) owns pixel row band0+y' and
//            channel quad h, so for one RoI the warp applies all 14 column updates of all its rows at once:
//            two batches (the seven J0, then the seven J1) of 7 independent LDS.128 / 4 FADD / STS.128 chains.  Every D
//            element has exactly one owner lane and the RoIs are walked in ascending order: no atomics, no intra-warp
//            hazards, no __syncwarp, bitwise reproducible.  RoIs whose column edges are not strictly increasing (bins
//            thinner than a pixel) take the same path with the 14 updates in program order.
//   staging  warps 4-7 prefetch grad_out of the next group of 8 RoIs (4 coalesced loads per (RoI, quad, bin) item),
//            scale by 1/numel (0 for empty bins) and store it transposed as gs[RoI][bin][8 ch]; they also build the
//            group's edge words, column byte offsets and per-row cover masks.  One CTA barrier per group.
//   epilogue inclusive scan along x (one thread per (row, quad)), then LDS.128 -> four coalesced plane stores.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

namespace {

constexpr int RK = 7, RKK = 49;
constexpr int RCB = 8;             // channels per CTA
constexpr int RRG = 8;             // RoIs per staged group
constexpr int RAPPLY_WARPS = 4, RSTAGE_WARPS = 4;
constexpr int RTHREADS = (RAPPLY_WARPS + RSTAGE_WARPS) * 32;
constexpr int RSTAGE_THREADS = RSTAGE_WARPS * 32;
constexpr int RITEMS = (RRG * 2 * RKK + RSTAGE_THREADS - 1) / RSTAGE_THREADS;  // (RoI, quad, bin) items per staging thread
constexpr int RGS_FLOATS = RRG * RKK * RCB;   // one grad_out stage buffer
constexpr int RMETA_WORDS = 24;               // per RoI: 7 edge words, flags, 14 column byte offsets (+ pad)

__device__ __forceinline__ uint32_t r_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory accesses of the update chains: volatile asm keeps their order (all loads of a batch, then all stores)
// without a "memory" clobber, so the compiler may still move ordinary loads of the next RoI across them
__device__ __forceinline__ float4 r_lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void r_sts4(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ float r_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t r_pack_edges(const float* __restrict__ roi, int b, int H, int W) {
    int i0, i1, j0, j1;
    bin_edge<float, true>(roi[0], roi[2], b, RK, H, i0, i1);
    bin_edge<float, true>(roi[1], roi[3], b, RK, W, j0, j1);
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)j0 << 16) | ((uint32_t)j1 << 24);
}

__global__ void __launch_bounds__(RTHREADS, 2)
roipool_rows_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R,
                        int C, int H, int W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int xs = H * RCB + 4;                       // floats per pixel column (odd number of 16-byte units)
    float* D = reinterpret_cast<float*>(smem_raw);    // [(W + 1)][xs]
    float* gS = D + (size_t)(W + 1) * xs;             // [2][RRG][49][8]
    uint32_t* metaS = reinterpret_cast<uint32_t*>(gS + 2 * RGS_FLOATS);          // [2][RRG][RMETA_WORDS]
    unsigned char* coverS = reinterpret_cast<unsigned char*>(metaS + 2 * RRG * RMETA_WORDS);  // [2][RRG][64]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * RCB;
    const int cb = min(RCB, C - c0);
    const int HW = H * W;

    for (int idx = tid; idx < (W + 1) * xs / 4; idx += RTHREADS)
        *reinterpret_cast<float4*>(D + idx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);

    const int nGroups = (R + RRG - 1) / RRG;
    const bool stager = warp >= RAPPLY_WARPS;
    const int stid = tid - RAPPLY_WARPS * 32;  // 0 .. RSTAGE_THREADS-1 for staging threads

    // ---- staging side ---------------------------------------------------------------------------------------------
    int itRR[RITEMS], itQ[RITEMS], itBin[RITEMS];
    float4 pre[RITEMS];
    if (stager) {
#pragma unroll
        for (int n = 0; n < RITEMS; ++n) {
            const int it = stid + n * RSTAGE_THREADS;
            itRR[n] = it / (2 * RKK);
            const int rem = it - itRR[n] * (2 * RKK);
            itQ[n] = rem / RKK;
            itBin[n] = rem - itQ[n] * RKK;
            if (itRR[n] >= RRG) itRR[n] = -1;
        }
    }
    auto prefetch = [&](int grp) {
#pragma unroll
        for (int n = 0; n < RITEMS; ++n) {
            const int rr = itRR[n];
            const int r = grp * RRG + rr;
            const bool ok = rr >= 0 && r < R;
            const int ch = 4 * itQ[n];
            const float* src = go + ((size_t)(ok ? r : 0) * C + c0 + ch) * RKK + itBin[n];
            pre[n].x = (ok && ch + 0 < cb) ? __ldg(src) : 0.f;
            pre[n].y = (ok && ch + 1 < cb) ? __ldg(src + RKK) : 0.f;
            pre[n].z = (ok && ch + 2 < cb) ? __ldg(src + 2 * RKK) : 0.f;
            pre[n].w = (ok && ch + 3 < cb) ? __ldg(src + 3 * RKK) : 0.f;
        }
    };
    // meta + cover of group `grp` into buffer `buf` (must precede commit: commit reads the edge words)
    auto build_meta = [&](int grp, int buf) {
        uint32_t* meta = metaS + buf * RRG * RMETA_WORDS;
        for (int idx = stid; idx < RRG * RK; idx += RSTAGE_THREADS) {
            const int rr = idx / RK, b = idx - rr * RK;
            const int r = grp * RRG + rr;
            uint32_t e = 0;
            if (r < R) e = r_pack_edges(rois + (size_t)r * 4, b, H, W);
            meta[rr * RMETA_WORDS + b] = e;
        }
    };
    auto finish_meta = [&](int grp, int buf) {
        uint32_t* meta = metaS + buf * RRG * RMETA_WORDS;
        unsigned char* cov = coverS + buf * RRG * 64;
        for (int idx = stid; idx < RRG * 64; idx += RSTAGE_THREADS) {  // cover[rr][y]: bit i <=> bin row i contains y
            const int rr = idx >> 6, y = idx & 63;
            unsigned m = 0;
            if (grp * RRG + rr < R) {
#pragma unroll
                for (int b = 0; b < RK; ++b) {
                    const uint32_t e = meta[rr * RMETA_WORDS + b];
                    m |= ((int)(e & 255) <= y && y < (int)((e >> 8) & 255)) ? (1u << b) : 0u;
                }
            }
            cov[idx] = (unsigned char)m;
        }
        for (int rr = stid; rr < RRG; rr += RSTAGE_THREADS) {  // flags and column byte offsets
            uint32_t* m = meta + rr * RMETA_WORDS;
            bool simple = true;
            int rowLo = 255, rowHi = 0;
#pragma unroll
            for (int b = 0; b < RK; ++b) {
                const uint32_t e = m[b];
                const int J0 = (e >> 16) & 255, J1 = e >> 24;
                m[8 + b] = (uint32_t)(J0 * xs * 4);
                m[16 + b] = (uint32_t)(J1 * xs * 4);
                if (b > 0) {
                    const uint32_t a = m[b - 1];
                    simple = simple && J0 > (int)((a >> 16) & 255) && J1 > (int)(a >> 24);
                }
                const int I0 = e & 255, I1 = (e >> 8) & 255;
                if (I1 > I0) { rowLo = min(rowLo, I0); rowHi = max(rowHi, I1); }
            }
            const bool live = grp * RRG + rr < R && rowHi > rowLo;
            m[7] = (live ? 1u : 0u) | (simple ? 2u : 0u) | ((uint32_t)(live ? rowLo : 0) << 8) | ((uint32_t)(live ? rowHi : 0) << 16);
        }
    };
    auto commit = [&](int grp, int buf) {
        float* g = gS + buf * RGS_FLOATS;
        const uint32_t* meta = metaS + buf * RRG * RMETA_WORDS;
#pragma unroll
        for (int n = 0; n < RITEMS; ++n) {
            const int rr = itRR[n];
            if (rr < 0 || grp * RRG + rr >= R) continue;
            const int bi = itBin[n] / RK, bj = itBin[n] - bi * RK;
            const uint32_t ei = meta[rr * RMETA_WORDS + bi], ej = meta[rr * RMETA_WORDS + bj];
            const int hI = (int)((ei >> 8) & 255) - (int)(ei & 255);
            const int wJ = (int)(ej >> 24) - (int)((ej >> 16) & 255);
            const float inv = (hI > 0 && wJ > 0) ? r_rcp((float)(hI * wJ)) : 0.f;
            float4 v = pre[n];
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
            *reinterpret_cast<float4*>(g + ((rr * RKK + itBin[n]) * 2 + itQ[n]) * 4) = v;
        }
    };
    auto stage_barrier = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(RSTAGE_THREADS) : "memory"); };

    // ---- apply side -----------------------------------------------------------------------------------------------
    // row bands: boundaries at 0.29 / 0.5 / 0.71 of H (the middle rows are covered by more RoIs)
    const int bnd1 = (29 * H + 50) / 100, bnd2 = H / 2, bnd3 = H - bnd1;
    const int band0 = warp == 0 ? 0 : warp == 1 ? bnd1 : warp == 2 ? bnd2 : bnd3;
    const int band1 = warp == 0 ? bnd1 : warp == 1 ? bnd2 : warp == 2 ? bnd3 : H;
    const int yl = lane >> 1, hq = lane & 1;     // row slot inside the band, channel quad
    const int y = band0 + yl;
    const bool rowMine = !stager && y < band1;
    const uint32_t dLane = r_smem(D) + (uint32_t)((y * RCB + 4 * hq) * 4);

    if (stager) {
        build_meta(0, 0);
        prefetch(0);
        stage_barrier();
        finish_meta(0, 0);
        commit(0, 0);
    }

    for (int grp = 0; grp < nGroups; ++grp) {
        const int buf = grp & 1;
        __syncthreads();  // stage / meta / cover of `buf` complete; the apply warps are done with group grp-1
        if (stager) {
            if (grp + 1 < nGroups) {
                prefetch(grp + 1);
                build_meta(grp + 1, buf ^ 1);
                stage_barrier();
                finish_meta(grp + 1, buf ^ 1);
                commit(grp + 1, buf ^ 1);
            }
        } else {
            const uint32_t* meta = metaS + buf * RRG * RMETA_WORDS;
            const unsigned char* cov = coverS + buf * RRG * 64;
            const float* gB = gS + buf * RGS_FLOATS + 4 * hq;
#pragma unroll 1
            for (int rr = 0; rr < RRG; ++rr) {
                const uint32_t* m = meta + rr * RMETA_WORDS;
                const uint32_t flags = m[7];
                const int rowLo = (flags >> 8) & 255, rowHi = (flags >> 16) & 255;
                if (!(flags & 1u) || rowHi <= band0 || rowLo >= band1) continue;  // warp-uniform: RoI misses this band
                unsigned cover = rowMine ? cov[rr * 64 + y] : 0u;
                const bool rowHit = cover != 0u;  // this lane's row belongs to the RoI
                // t_j for this lane's row: the (one or two) bin rows that contain it
                float4 t[RK];
#pragma unroll
                for (int j = 0; j < RK; ++j) t[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* gR = gB + rr * (RKK * RCB);
                while (cover) {
                    const int i = __ffs(cover) - 1;
                    cover &= cover - 1;
#pragma unroll
                    for (int j = 0; j < RK; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(gR + (i * RK + j) * RCB);
                        t[j].x += v.x; t[j].y += v.y; t[j].z += v.z; t[j].w += v.w;
                    }
                }
                if (!rowHit) continue;
                if (flags & 2u) {
                    float4 d[RK];
#pragma unroll
                    for (int j = 0; j < RK; ++j) d[j] = r_lds4(dLane + m[8 + j]);
#pragma unroll
                    for (int j = 0; j < RK; ++j) {
                        d[j].x += t[j].x; d[j].y += t[j].y; d[j].z += t[j].z; d[j].w += t[j].w;
                    }
#pragma unroll
                    for (int j = 0; j < RK; ++j) r_sts4(dLane + m[8 + j], d[j]);
#pragma unroll
                    for (int j = 0; j < RK; ++j) d[j] = r_lds4(dLane + m[16 + j]);
#pragma unroll
                    for (int j = 0; j < RK; ++j) {
                        d[j].x -= t[j].x; d[j].y -= t[j].y; d[j].z -= t[j].z; d[j].w -= t[j].w;
                    }
#pragma unroll
                    for (int j = 0; j < RK; ++j) r_sts4(dLane + m[16 + j], d[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < RK; ++j) {  // degenerate column edges: strictly in program order
                        float4 a = r_lds4(dLane + m[8 + j]);
                        a.x += t[j].x; a.y += t[j].y; a.z += t[j].z; a.w += t[j].w;
                        r_sts4(dLane + m[8 + j], a);
                        float4 b = r_lds4(dLane + m[16 + j]);
                        b.x -= t[j].x; b.y -= t[j].y; b.z -= t[j].z; b.w -= t[j].w;
                        r_sts4(dLane + m[16 + j], b);
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- epilogue: inclusive scan along x, then transposed write-out -------------------------------------------------
    for (int t = tid; t < H * 2; t += RTHREADS) {
        const int yy = t >> 1, qq = t & 1;
        float* p = D + yy * RCB + 4 * qq;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int x = 0; x < W; ++x) {
            float4 v = *reinterpret_cast<float4*>(p + (size_t)x * xs);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            *reinterpret_cast<float4*>(p + (size_t)x * xs) = acc;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * HW; idx += RTHREADS) {
        const int qq = idx / HW, pix = idx - qq * HW;
        const int yy = pix / W, x = pix - yy * W;
        const float4 v = *reinterpret_cast<const float4*>(D + (size_t)x * xs + yy * RCB + 4 * qq);
        float* dst = gin + (size_t)(c0 + 4 * qq) * HW + pix;
        if (4 * qq + 0 < cb) dst[0] = v.x;
        if (4 * qq + 1 < cb) dst[HW] = v.y;
        if (4 * qq + 2 < cb) dst[2 * HW] = v.z;
        if (4 * qq + 3 < cb) dst[3 * HW] = v.w;
    }
}

size_t rows_smem(int H, int W) {
    return (size_t)(W + 1) * (H * RCB + 4) * sizeof(float) + (size_t)2 * RGS_FLOATS * sizeof(float) +
           (size_t)2 * RRG * RMETA_WORDS * sizeof(uint32_t) + (size_t)2 * RRG * 64 + 16;
}

}  // namespace

bool roipool_rows_bwd_supported(int R, int C, int H, int W, int k) {
    if (k != RK || R <= 0 || C <= 0 || H < 4 || H > 64 || W > 255) return false;
    const int b1 = (29 * H + 50) / 100, b2 = H / 2, b3 = H - b1;
    if (b1 > 16 || b2 - b1 > 16 || b3 - b2 > 16 || b1 < 1 || b2 <= b1) return false;  // a band is at most 16 rows
    // opt-in: measured 209 us against 171 us for the third-generation kernel at the track-head size (two apply warps
    // per scheduler walk 300 RoIs one after the other; each RoI is a chain of ~5 shared-memory round trips)
    const char* e = getenv("D2T_ROIPOOL_ROWS");
    if (!(e && e[0] == '1')) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    return rows_smem(H, W) <= (size_t)di.max_smem_optin;
}

int roipool_rows_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W, cudaStream_t st) {
    const size_t smem = rows_smem(H, W);
    D2T_CUDA_TRY(cudaFuncSetAttribute(roipool_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roipool_rows_bwd_kernel<<<ceil_div(C, RCB), RTHREADS, smem, st>>>(go, rois, gin, R, C, H, W);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
