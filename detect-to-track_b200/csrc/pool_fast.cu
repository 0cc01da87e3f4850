// pool_fast.cu -- float32 ROIPool forward/backward with warp-uniform control flow (sm_100a).
//
// The first-generation slab kernels in pool.cu gave each LANE its own (channel, bin/column) item, so
// every lane carried its own loop bounds and the warp executed the union of all paths: ~200
// instructions per output, issue-bound far from the HBM roof (profiles/r1_ncu_pool_v2_summary.txt).
// Here the lanes of a half-warp are 16 CHANNELS of the same pixel / bin, so all control flow (RoI, bin
// edges, pixel loops) is warp-uniform and is paid once per 16 channels:
//
//   slab     : a CTA owns 16 channels.  The planes live in shared memory transposed to [pixel][16]
//              (64 B per pixel): a half-warp touching one pixel hits 16 consecutive banks, the two
//              half-warps of a warp work on adjacent pixels / bins => conflict-free.
//   backward : every WARP owns pixel rows (row % 32 == warp) for all RoIs, walks the RoIs in order and
//              adds each RoI's contribution to its own rows: exclusive read-modify-write, no atomics,
//              no barrier per RoI, fixed summation order (bitwise reproducible).  grad_out blocks of a
//              group of RoIs are staged [bin][channel] by cp.async, double buffered.
//   forward  : every warp owns whole RoIs; a half-warp sums one bin rows-then-columns (the reference's
//              order => bit-identical results), outputs are staged per warp and written as one
//              contiguous run per RoI.
#include "common.cuh"

namespace d2t {

constexpr int kFastCB = 16;        // channels per CTA
constexpr int kFastThreads = 1024; // 32 warps
constexpr int kFastMaxK = 15;      // bin index must fit 4 bits
constexpr int kGP = 17;            // staged grad_out pitch per bin (floats): 16 channels + 1 => conflict-free transposing store

__device__ __forceinline__ uint32_t pf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pf_cp_async4(uint32_t dst, const float* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void pf_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pf_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct FastLayout {
    int RG;           // RoIs per staged group
    size_t accOff, gOff, invOff, edgeOff, rowbinOff, total;
    int tmpChunk;     // pixels per transposition chunk (aliases the g buffers)
};

__host__ __device__ inline FastLayout fast_layout(int H, int W, int k, int RG) {
    FastLayout L;
    L.RG = RG;
    const int kk = k * k;
    size_t off = 0;
    L.accOff = off;   off += (size_t)H * W * kFastCB * sizeof(float);
    L.gOff = off;     off += (size_t)2 * RG * kk * kGP * sizeof(float);
    L.invOff = off;   off += (size_t)2 * RG * kk * sizeof(float);
    L.edgeOff = off;  off += (size_t)2 * RG * k * 4 * sizeof(short);
    L.rowbinOff = off; off += ((size_t)2 * RG * H + 15) / 16 * 16;
    L.total = off;
    const size_t gBytes = (size_t)2 * RG * kk * kGP * sizeof(float);
    int chunk = (int)(gBytes / sizeof(float) / kFastCB) - 1;
    L.tmpChunk = chunk < 32 ? 0 : (chunk / 32) * 32;
    return L;
}

// ----------------------------------------------------------------------------------------------------
// per-group tables: bin edges, 1/numel, and for every pixel row the range of row bins that cover it
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fast_tables(const float* __restrict__ rois, int r0, int nr, int k, int H, int W,
                                            short* edges, float* inv, unsigned char* rowbin, bool wantInv) {
    const int kk = k * k;
    // edges[(rr*k + b)*4 + {0,1,2,3}] = I0, I1, J0, J1
    for (int idx = threadIdx.x; idx < nr * k; idx += blockDim.x) {
        const int rr = idx / k, b = idx - rr * k;
        const float* roi = rois + (size_t)(r0 + rr) * 4;
        int e0, e1;
        bin_edge<float, true>(roi[0], roi[2], b, k, H, e0, e1);
        edges[idx * 4 + 0] = (short)e0;
        edges[idx * 4 + 1] = (short)e1;
        bin_edge<float, true>(roi[1], roi[3], b, k, W, e0, e1);
        edges[idx * 4 + 2] = (short)e0;
        edges[idx * 4 + 3] = (short)e1;
    }
    if (wantInv) {
        for (int idx = threadIdx.x; idx < nr * kk; idx += blockDim.x) {
            const int rr = idx / kk, b = idx - rr * kk;
            const int i = b / k, j = b - i * k;
            const float* roi = rois + (size_t)(r0 + rr) * 4;
            int i0, i1, j0, j1;
            bin_edge<float, true>(roi[0], roi[2], i, k, H, i0, i1);
            bin_edge<float, true>(roi[1], roi[3], j, k, W, j0, j1);
            inv[idx] = 1.0f / (float)((i1 - i0) * (j1 - j0));
        }
        // rowbin[rr*H + y] = ilo | ihi << 4 (bins covering row y are contiguous), 0xFF if none
        for (int idx = threadIdx.x; idx < nr * H; idx += blockDim.x) {
            const int rr = idx / H, y = idx - rr * H;
            const float* roi = rois + (size_t)(r0 + rr) * 4;
            int lo = 15, hi = -1;
            for (int i = 0; i < k; ++i) {
                int i0, i1;
                bin_edge<float, true>(roi[0], roi[2], i, k, H, i0, i1);
                if (i0 <= y && y < i1) {
                    lo = min(lo, i);
                    hi = i;
                }
            }
            rowbin[idx] = hi < 0 ? 0xFF : (unsigned char)(lo | (hi << 4));
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------------------------------
template <int KT>  // KT = r_hw when specialised (fully unrolled bin loops), 0 = runtime r_hw
__global__ void __launch_bounds__(kFastThreads, 1)
roipool_fast_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R,
                        int C, int H, int W, int k, int RG) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FastLayout L = fast_layout(H, W, k, RG);
    float* acc = reinterpret_cast<float*>(smem_raw + L.accOff);
    float* gS = reinterpret_cast<float*>(smem_raw + L.gOff);
    float* invS = reinterpret_cast<float*>(smem_raw + L.invOff);
    short* edgeS = reinterpret_cast<short*>(smem_raw + L.edgeOff);
    unsigned char* rowbinS = smem_raw + L.rowbinOff;

    const int kk = k * k;
    const int HW = H * W;
    const int c0 = blockIdx.x * kFastCB;
    const int cb = min(kFastCB, C - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, c = lane & 15;
    const int nGroups = (R + RG - 1) / RG;
    const int gStride = RG * kk * kGP;  // floats per g buffer

    for (int idx = tid; idx < HW * kFastCB; idx += kFastThreads) acc[idx] = 0.f;

    // stage grad_out of a RoI group: go[r][c0 + cc][b] -> gS[buf][rr][b][cc]
    auto stage_group = [&](int grp, int buf) {
        const int r0 = grp * RG, nr = min(RG, R - r0);
        const uint32_t base = pf_smem_u32(gS + buf * gStride);
        const int perRoi = kFastCB * kk;
        for (int e = tid; e < nr * perRoi; e += kFastThreads) {
            const int rr = e / perRoi, rem = e - rr * perRoi;
            const int cc = rem / kk, b = rem - cc * kk;
            const bool ok = cc < cb;
            const float* src = go + ((size_t)(r0 + rr) * C + c0 + (ok ? cc : 0)) * kk + b;
            pf_cp_async4(base + (uint32_t)((rr * kk + b) * kGP + cc) * 4u, src, ok ? 4u : 0u);
        }
        pf_cp_async_commit();
    };
    auto tables = [&](int grp, int buf) {
        const int r0 = grp * RG, nr = min(RG, R - r0);
        fast_tables(rois, r0, nr, k, H, W, edgeS + buf * RG * k * 4, invS + buf * RG * kk, rowbinS + buf * RG * H, true);
    };

    stage_group(0, 0);
    tables(0, 0);

    for (int grp = 0; grp < nGroups; ++grp) {
        const int buf = grp & 1;
        pf_cp_async_wait_all();
        __syncthreads();  // group `grp` staged + its tables visible; everyone is done with group grp-1
        if (grp + 1 < nGroups) {
            stage_group(grp + 1, buf ^ 1);
            tables(grp + 1, buf ^ 1);
        }
        const int nr = min(RG, R - grp * RG);
        const float* gB = gS + buf * gStride;
        const float* invB = invS + buf * RG * kk;
        const short* edB = edgeS + buf * RG * k * 4;
        const unsigned char* rbB = rowbinS + buf * RG * H;
        for (int rr = 0; rr < nr; ++rr) {
            const short* ed = edB + rr * k * 4;
            const float* gR = gB + rr * kk * kGP + c;
            const float* invR = invB + rr * kk;
            for (int pi = warp; pi < H; pi += 32) {  // rows owned by this warp
                const unsigned rb = rbB[rr * H + pi];
                if (rb == 0xFF) continue;
                const int ilo = rb & 15, ihi = rb >> 4;
                float* arow = acc + (size_t)pi * W * kFastCB + c;
                if (KT > 0) {
                    float v[KT > 0 ? KT : 1];
#pragma unroll
                    for (int j = 0; j < KT; ++j) v[j] = 0.f;
#pragma unroll 1
                    for (int i = ilo; i <= ihi; ++i) {
#pragma unroll
                        for (int j = 0; j < KT; ++j) v[j] = fmaf(gR[(i * KT + j) * kGP], invR[i * KT + j], v[j]);
                    }
#pragma unroll
                    for (int j = 0; j < KT; ++j) {
                        const int j0 = ed[j * 4 + 2] + half, j1 = ed[j * 4 + 3];
                        float* a = arow + j0 * kFastCB;
#pragma unroll
                        for (int u = 0; u < 4; ++u)  // bins up to 8 pixels wide: straight-line, predicated
                            if (j0 + 2 * u < j1) a[2 * u * kFastCB] += v[j];
#pragma unroll 1
                        for (int pj = j0 + 8; pj < j1; pj += 2) arow[pj * kFastCB] += v[j];
                    }
                } else {
                    for (int j = 0; j < k; ++j) {
                        float v = 0.f;
                        for (int i = ilo; i <= ihi; ++i) v = fmaf(gR[(i * k + j) * kGP], invR[i * k + j], v);
                        const int j0 = ed[j * 4 + 2], j1 = ed[j * 4 + 3];
                        for (int pj = j0 + half; pj < j1; pj += 2) arow[pj * kFastCB] += v;
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- write-out: acc[px][c] -> gin[c0+c][px], transposed through the (now free) g buffers ----------
    float* tmp = gS;  // [16][CH + 1]
    const int CH = L.tmpChunk;
    const int TP = CH + 1;
    for (int p0 = 0; p0 < HW; p0 += CH) {
        const int n = min(CH, HW - p0);
        for (int e = tid; e < n * kFastCB; e += kFastThreads) {
            const int px = e >> 4, cc = e & 15;
            tmp[cc * TP + px] = acc[(size_t)(p0 + px) * kFastCB + cc];
        }
        __syncthreads();
        for (int e = tid; e < cb * n; e += kFastThreads) {
            const int cc = e / n, px = e - cc * n;
            gin[(size_t)(c0 + cc) * HW + p0 + px] = tmp[cc * TP + px];
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------------
// forward (row prefix sums, any r_hw <= 32)
// ----------------------------------------------------------------------------------------------------
struct FastDivP {
    uint32_t m, s;
};
static FastDivP make_fastdiv_p(uint32_t d) {
    FastDivP f;
    uint32_t l = 0;
    while ((1u << l) < d) ++l;
    f.s = l;
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fdiv_p(uint32_t n, const FastDivP& f) { return (__umulhi(f.m, n) + n) >> f.s; }

__global__ void __launch_bounds__(kFastThreads, 1)
roipool_prefix_fwd_kernel(const float* __restrict__ fm, const float* __restrict__ rois, float* __restrict__ out, int R,
                          int C, int H, int W, int k, int CB, int RG, int rowPitch, int planePitch, FastDivP dCBkk,
                          FastDivP dkk, FastDivP dk, FastDivP dW) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* P = reinterpret_cast<float*>(smem_raw);                    // [CB][H][rowPitch], rowPitch odd, >= W + 1
    short* edgeS = reinterpret_cast<short*>(P + (size_t)CB * planePitch);  // [RG][k][4]
    float* invS = reinterpret_cast<float*>(edgeS + RG * k * 4);             // [RG][kk]
    uint32_t* decodeS = reinterpret_cast<uint32_t*>(invS + RG * k * k);      // [CB*kk]
    const int kk = k * k;
    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);
    const int tid = threadIdx.x;

    // slab load (coalesced): element x of a row goes to column x + 1, column 0 stays for the leading zero
    for (int idx = tid; idx < cb * H * W; idx += kFastThreads) {
        const int row = fdiv_p(idx, dW);  // (cc*H + y)
        const int x = idx - row * W;
        const int cc = row / H, y = row - cc * H;
        P[(size_t)cc * planePitch + y * rowPitch + x + 1] = __ldg(fm + (size_t)c0 * H * W + idx);
    }
    __syncthreads();
    // in-place inclusive scan per row: P[y][x] = sum_{x' < x} row[x'] (one thread per row; odd pitch => no conflicts)
    for (int row = tid; row < cb * H; row += kFastThreads) {
        const int cc = row / H, y = row - cc * H;
        float* p = P + (size_t)cc * planePitch + y * rowPitch;
        float acc = 0.f;
        p[0] = 0.f;
        for (int x = 1; x <= W; ++x) {
            acc += p[x];
            p[x] = acc;
        }
    }

    // decode table: item o in [0, CB*kk) -> (cc, i, j) packed, so the hot loop has no divisions
    for (int o = tid; o < CB * kk; o += kFastThreads) {
        const int cc = fdiv_p(o, dkk), b = o - cc * kk;
        const int i = fdiv_p(b, dk), j = b - i * k;
        decodeS[o] = (uint32_t)cc | ((uint32_t)i << 8) | ((uint32_t)j << 16);
    }

    for (int r0 = 0; r0 < R; r0 += RG) {
        const int nr = min(RG, R - r0);
        __syncthreads();
        fast_tables(rois, r0, nr, k, H, W, edgeS, invS, nullptr, false);
        for (int idx = tid; idx < nr * kk; idx += kFastThreads) {  // 1/numel per (roi, bin); inf/NaN for empty bins
            const int rr = idx / kk, b = idx - rr * kk;
            const int i = b / k, j = b - i * k;
            const float* roi = rois + (size_t)(r0 + rr) * 4;
            int i0, i1, j0, j1;
            bin_edge<float, true>(roi[0], roi[2], i, k, H, i0, i1);
            bin_edge<float, true>(roi[1], roi[3], j, k, W, j0, j1);
            invS[idx] = 1.0f / (float)((i1 - i0) * (j1 - j0));
        }
        __syncthreads();
        for (int rr = 0; rr < nr; ++rr) {
            const uint32_t* ed32 = reinterpret_cast<const uint32_t*>(edgeS + rr * k * 4);  // per bin index: {I0|I1<<16, J0|J1<<16}
            const float* inv = invS + rr * kk;
            float* orow = out + ((size_t)(r0 + rr) * C + c0) * kk;
            for (int o = tid; o < cb * kk; o += kFastThreads) {
                const uint32_t d = decodeS[o];
                const int cc = d & 0xff, i = (d >> 8) & 0xff, j = d >> 16;
                const uint32_t ei = ed32[i * 2], ej = ed32[j * 2 + 1];
                const int i0 = (short)(ei & 0xffff), i1 = (short)(ei >> 16);
                const int j0 = (short)(ej & 0xffff), j1 = (short)(ej >> 16);
                const float* p = P + (size_t)cc * planePitch + i0 * rowPitch;
                float acc = 0.f;
#pragma unroll 1
                for (int y = i0; y < i1; ++y, p += rowPitch) acc += p[j1] - p[j0];  // 2-4 trips: keep it a plain loop
                orow[o] = acc * inv[i * k + j];  // 0 * inf = NaN on empty bins, like the reference's 0/0 (F7)
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
static bool pick_rg_bwd(int H, int W, int k, size_t budget, int* RG) {
    for (int rg = 8; rg >= 2; rg /= 2) {
        FastLayout L = fast_layout(H, W, k, rg);
        if (L.total <= budget && L.tmpChunk >= 32) {
            *RG = rg;
            return true;
        }
    }
    return false;
}

bool roipool_fast_supported(int R, int C, int H, int W, int k) {
    if (R <= 0 || C <= 0 || k > kFastMaxK || H >= 32768 || W >= 32768) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    const size_t budget = (size_t)di.max_smem_optin;
    int rg;
    return pick_rg_bwd(H, W, k, budget, &rg);
}

int roipool_fast_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W, int k,
                            cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    int RG = 0;
    if (!pick_rg_bwd(H, W, k, (size_t)di.max_smem_optin, &RG)) {
        set_error("roipool_fast_bwd: plane does not fit shared memory");
        return D2T_ERR_BAD_ARG;
    }
    const FastLayout L = fast_layout(H, W, k, RG);
    auto kern = (k == 7) ? roipool_fast_bwd_kernel<7> : roipool_fast_bwd_kernel<0>;
    D2T_SMEM_OPTIN(kern, L.total);
    kern<<<ceil_div(C, kFastCB), kFastThreads, L.total, st>>>(go, rois, gin, R, C, H, W, k, RG);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

bool roipool_prefix_supported(int R, int C, int H, int W, int k) {
    if (R <= 0 || C <= 0 || k > 32 || H >= 32768 || W >= 32768) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    const int rowPitch = (W + 1) | 1;
    const size_t plane = (size_t)H * rowPitch * sizeof(float);
    return k <= 32 && plane + 64 * k * 4 * sizeof(short) + 64 * k * k * sizeof(float) + 8192 + 1024 <= (size_t)di.max_smem_optin;
}

int roipool_prefix_fwd_launch(const float* fm, const float* rois, float* out, int R, int C, int H, int W, int k,
                              cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int RG = 64;
    const int rowPitch = (W + 1) | 1;
    const int planePitch = H * rowPitch;
    const size_t tabBytes = (size_t)RG * k * 4 * sizeof(short) + (size_t)RG * k * k * sizeof(float);
    const size_t budget = (size_t)di.max_smem_optin - tabBytes - 8192;  // 8 KB reserve for the decode table
    int maxCB = (int)(budget / ((size_t)planePitch * sizeof(float)));
    int CB = ceil_div(C, di.sm_count);
    if (CB > maxCB) {
        const int waves = ceil_div(ceil_div(C, maxCB), di.sm_count);
        CB = ceil_div(C, waves * di.sm_count);
        if (CB > maxCB) CB = maxCB;
    }
    if (CB < 1) CB = 1;
    if ((size_t)CB * k * k * 4 > 8192) CB = (int)(8192 / ((size_t)k * k * 4));
    if (CB < 1) CB = 1;
    const size_t smem = (size_t)CB * planePitch * sizeof(float) + tabBytes + (size_t)CB * k * k * sizeof(uint32_t);
    D2T_SMEM_OPTIN(roipool_prefix_fwd_kernel, smem);
    roipool_prefix_fwd_kernel<<<ceil_div(C, CB), kFastThreads, smem, st>>>(
        fm, rois, out, R, C, H, W, k, CB, RG, rowPitch, planePitch, make_fastdiv_p(CB * k * k), make_fastdiv_p(k * k),
        make_fastdiv_p(k), make_fastdiv_p(W));
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
