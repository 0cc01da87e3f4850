// pool_vec4.cu -- float32 ROIPool backward: a WARP owns a channel plane, lanes are the bins of a RoI, two-dimensional
// difference array in shared memory, double-precision two-dimensional scan.  sm_100a.
// Reference: atomicAdd per bin pixel, roipool_cuda.cu:68-127.
//
// pool_vec2.cu (row difference arrays over [pixel][16 channel] slabs) executes one update per (RoI, pixel row): 16.4 rows per
// RoI x 77 instructions for 16 channels, plus the staging of grad_out through shared memory, row lists and a barrier per RoI
// group: 615 k warp instructions per SM, 155 us at the track-head size.  The adjoint of an average over a rectangle is a FOUR-
// point update of a 2-D difference array, whatever the rectangle's size:
//
//     D[I0][J0] += v,  D[I0][J1] -= v,  D[I1][J0] -= v,  D[I1][J1] += v      v = grad_out[r, c, bin] / numel(bin)
//     grad_fm[c, y, x] = sum_{y' <= y, x' <= x} D[y'][x']
//
//   records  the four offsets of a bin do not depend on the channel: `rp4_records_kernel` (a warp per RoI) writes, per (RoI,
//            bin), {offset of (I0, J0) | dJ << 16 | dI << 24,  phase | phases << 8 | skip << 31} with the reference's edge
//            arithmetic
//   main     a warp owns ONE channel: its (H+1) x (W+1) difference plane lives in shared memory, its lanes are the bins of the
//            current RoI (bins 0..31, then 32..48).  grad_out[r, c, 0..48] is one contiguous 196-byte run, read straight from
//            global memory (no staging, every byte of grad_out crosses HBM once), a ring of eight RoIs in registers (gradients
//            and records are requested seven RoIs before they are used).  Nothing is
//            shared between warps: no barrier, no atomics, and the update order of every plane element is fixed (RoIs
//            ascending, corner type, phase) => bitwise reproducible.
//   phases   two bins of a RoI collide on a corner only if a row edge AND a column edge repeat, i.e. for bins thinner than a
//            pixel.  Edges are monotone, so equal edges are consecutive: with pI / pJ the longest runs of equal row / column
//            edges, bins with the same (i mod pI, j mod pJ) never collide; the record carries that phase and a RoI takes
//            pI * pJ passes (1 for 2 / 3 of the RoIs of the benchmark set, 1.5 on average).
//   scan     per warp: lanes own columns, a row's prefix comes from a shuffle scan, the column sums are carried in registers --
//            all in DOUBLE precision, so the rounding residue of the difference array does not grow with the map (a float32
//            2-D scan would put ~1e-4 relative error on all-positive gradients); rows leave as coalesced stores.
#include "common.cuh"

namespace d2t {

namespace {

constexpr int V4MaxWarps = 16;
constexpr int V4PD = 8;          // RoIs whose gradients are in flight per warp
constexpr int V4MaxKK = 64;      // two lane passes

__device__ __forceinline__ float v4_ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__host__ __device__ inline int v4_pitch(int W) { return (W + 1) | 1; }
__host__ __device__ inline int v4_plane_words(int H, int W) { return ((H + 1) * v4_pitch(W) + 3) & ~3; }

// longest run of equal consecutive values among lanes [base, base + k) of `v` (uniform result)
__device__ __forceinline__ int v4_max_run(int v, int base, int k) {
    int best = 1, run = 1, prev = __shfl_sync(0xffffffffu, v, base);
    for (int i = 1; i < k; ++i) {
        const int cur = __shfl_sync(0xffffffffu, v, base + i);
        run = cur == prev ? run + 1 : 1;
        best = max(best, run);
        prev = cur;
    }
    return best;
}

// one warp per RoI
__global__ void __launch_bounds__(256)
rp4_records_kernel(const float* __restrict__ rois, uint2* __restrict__ rec, int R, int H, int W, int k) {
    const int lane = threadIdx.x & 31;
    const int r = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= R) return;
    const float* roi = rois + (size_t)r * 4;
    int e0 = 0, e1 = 0;
    if (lane < k) bin_edge<float, true>(__ldg(roi), __ldg(roi + 2), lane, k, H, e0, e1);               // roipool_cuda.cu:38-50
    else if (lane < 2 * k) bin_edge<float, true>(__ldg(roi + 1), __ldg(roi + 3), lane - k, k, W, e0, e1);
    const int pI = max(v4_max_run(e0, 0, k), v4_max_run(e1, 0, k));
    const int pJ = max(v4_max_run(e0, k, k), v4_max_run(e1, k, k));
    const int P = v4_pitch(W), kk = k * k;
    for (int b0 = 0; b0 < kk; b0 += 32) {
        const int b = b0 + lane;
        const int i = b < kk ? b / k : 0, j = b < kk ? b - i * k : 0;
        const int i0 = __shfl_sync(0xffffffffu, e0, i), i1 = __shfl_sync(0xffffffffu, e1, i);
        const int j0 = __shfl_sync(0xffffffffu, e0, k + j), j1 = __shfl_sync(0xffffffffu, e1, k + j);
        if (b < kk) {
            const int dI = i1 - i0, dJ = j1 - j0;
            // an empty bin receives nothing (the reference's loops do not run); the range checks hold for every finite RoI
            const bool ok = dI > 0 && dJ > 0 && i0 >= 0 && i1 <= H && j0 >= 0 && j1 <= W;
            uint2 w;
            w.x = ok ? ((uint32_t)(i0 * P + j0) | ((uint32_t)dJ << 16) | ((uint32_t)dI << 24)) : 0u;
            w.y = (ok ? (uint32_t)((i % pI) * pJ + (j % pJ)) : 0x80000000u) | ((uint32_t)(pI * pJ) << 8);   // phase | phases << 8
            rec[(size_t)r * kk + b] = w;
        }
    }
}

template <int PASSES>
__global__ void __launch_bounds__(V4MaxWarps * 32, 1)
rp4_bwd_kernel(const float* __restrict__ go, const uint2* __restrict__ rec, float* __restrict__ gin,
               int R, int C, int H, int W, int kk, int CPB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * CPB + warp;
    if (warp >= CPB || c >= C) return;   // warps are independent: no block-wide barrier anywhere
    const int P = v4_pitch(W);
    const int words = v4_plane_words(H, W);
    float* D = reinterpret_cast<float*>(smem_raw) + (size_t)warp * words;
    for (int idx = lane; idx < words / 4; idx += 32) reinterpret_cast<float4*>(D)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const size_t stride = (size_t)C * kk;
    const float* g = go + (size_t)c * kk + lane;
    const uint2* rp = rec + lane;
    bool has[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) has[p] = lane + 32 * p < kk;

    // ring of V4PD RoIs in registers: slot q holds the gradients and records of RoI r0 + q; it is refilled with those of RoI
    // r0 + q + V4PD right after its values are taken, so every load has V4PD - 1 RoIs of work to land
    float gv[V4PD][PASSES];
    uint2 wv[V4PD][PASSES];
#pragma unroll
    for (int q = 0; q < V4PD; ++q)
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const bool ok = q < R && has[p];
            gv[q][p] = ok ? v4_ldg_stream(g + (size_t)q * stride + 32 * p) : 0.f;
            wv[q][p] = ok ? __ldg(rp + (size_t)q * kk + 32 * p) : make_uint2(0u, 0x80000000u);
        }

    for (int r0 = 0; r0 < R; r0 += V4PD) {
#pragma unroll
        for (int q = 0; q < V4PD; ++q) {
            const int r = r0 + q;
            if (r < R) {   // warp-uniform
                float v[PASSES];
                float* a00[PASSES];
                float* a10[PASSES];
                int dJ[PASSES], ph[PASSES];
#pragma unroll
                for (int p = 0; p < PASSES; ++p) {
                    const uint2 w = wv[q][p];
                    dJ[p] = (int)((w.x >> 16) & 255u);
                    const int dI = (int)(w.x >> 24);
                    v[p] = __fdividef(gv[q][p], (float)(dI * dJ[p]));   // roipool_cuda.cu:123 (go / numel)
                    a00[p] = D + (w.x & 0xffffu);
                    a10[p] = a00[p] + dI * P;
                    ph[p] = (int)(w.y & 0x800000ffu);   // negative: the bin is empty (or the lane has no bin)
                }
                const int np = (int)((__shfl_sync(0xffffffffu, wv[q][0].y, 0) >> 8) & 0xffffu);
                {   // refill the slot
                    const int rr = r + V4PD;
#pragma unroll
                    for (int p = 0; p < PASSES; ++p) {
                        const bool ok = rr < R && has[p];
                        gv[q][p] = ok ? v4_ldg_stream(g + (size_t)rr * stride + 32 * p) : 0.f;
                        wv[q][p] = ok ? __ldg(rp + (size_t)rr * kk + 32 * p) : make_uint2(0u, 0x80000000u);
                    }
                }
                for (int s = 0; s < np; ++s) {
                    // bins of one phase never share a corner, whichever lane pass they belong to: all loads of a step go out
                    // before its stores (the compiler cannot know that the addresses differ and would chain the passes).
                    // The four corner TYPES are separate warp-wide steps: the (I0, J1) corner of a bin is the (I0, J0) corner
                    // of its right neighbour whenever the shared edge falls on a pixel boundary
                    bool on[PASSES];
                    float o[PASSES];
#pragma unroll
                    for (int p = 0; p < PASSES; ++p) on[p] = ph[p] == s;
#define D2T_V4_STEP(PTR, OFF, SIGN)                                              \
    _Pragma("unroll") for (int p = 0; p < PASSES; ++p) o[p] = on[p] ? PTR[p][OFF] : 0.f; \
    _Pragma("unroll") for (int p = 0; p < PASSES; ++p) if (on[p]) PTR[p][OFF] = o[p] SIGN v[p]; \
    __syncwarp();
                    D2T_V4_STEP(a00, 0, +)
                    D2T_V4_STEP(a00, dJ[p], -)
                    D2T_V4_STEP(a10, 0, -)
                    D2T_V4_STEP(a10, dJ[p], +)
#undef D2T_V4_STEP
                }
            }
        }
    }
    __syncwarp();

    // two-dimensional inclusive scan in double precision: lane owns columns lane, lane + 32, ...; a row's prefix by a shuffle
    // scan, the sum over rows in a register per owned column
    float* out = gin + (size_t)c * H * W;
    constexpr int MAXSEG = 8;   // W <= 255
    double col[MAXSEG];
#pragma unroll
    for (int s = 0; s < MAXSEG; ++s) col[s] = 0.0;
    for (int y = 0; y < H; ++y) {
        const float* row = D + y * P;
        double carry = 0.0;
#pragma unroll
        for (int s = 0; s < MAXSEG; ++s) {
            const int x = s * 32 + lane;
            if (s * 32 < W) {   // uniform
                double v = x < W ? (double)row[x] : 0.0;
#pragma unroll
                for (int sh = 1; sh < 32; sh <<= 1) {
                    const double o = __shfl_up_sync(0xffffffffu, v, sh);
                    if (lane >= sh) v += o;
                }
                v += carry;
                carry = __shfl_sync(0xffffffffu, v, 31);
                col[s] += v;
                if (x < W) out[(size_t)y * W + x] = (float)col[s];
            }
        }
    }
}

}  // namespace

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
constexpr size_t kV4SmemCap = (size_t)227 * 1024;   // sm_100a (a constant: workspace queries need no device)

bool roipool_vec4_bwd_supported(int R, int C, int H, int W, int k) {
    if (R <= 0 || C <= 0 || H <= 0 || W <= 0 || k <= 0 || k * k > V4MaxKK || H > 255 || W > 255) return false;
    if ((long long)(H + 1) * v4_pitch(W) > 65535 || (long long)R * k * k > 0x7fffffffLL) return false;
    return (size_t)v4_plane_words(H, W) * 4 <= kV4SmemCap;
}

size_t roipool_vec4_bwd_ws_bytes(int R, int C, int H, int W, int k) {
    (void)C;
    (void)H;
    (void)W;
    return align_up((size_t)R * k * k * sizeof(uint2), 256);
}

int roipool_vec4_bwd_launch(const float* go, const float* rois, float* gin, int R, int C, int H, int W, int k, void* ws,
                            size_t ws_bytes, cudaStream_t st) {
    const size_t need = roipool_vec4_bwd_ws_bytes(R, C, H, W, k);
    if (!ws || ws_bytes < need) {
        set_error("roipool_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return D2T_ERR_WORKSPACE;
    }
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int kk = k * k;
    uint2* rec = static_cast<uint2*>(ws);
    rp4_records_kernel<<<ceil_div(R, 8), 256, 0, st>>>(rois, rec, R, H, W, k);
    D2T_CUDA_TRY(cudaGetLastError());
    // channels (= warps) per CTA: one balanced wave of CTAs where possible
    const size_t planeBytes = (size_t)v4_plane_words(H, W) * 4;
    int cpbMax = (int)(kV4SmemCap / planeBytes);
    if (cpbMax > V4MaxWarps) cpbMax = V4MaxWarps;
    const int waves = ceil_div(ceil_div(C, cpbMax), di.sm_count);
    int CPB = ceil_div(C, waves * di.sm_count);
    if (CPB > cpbMax) CPB = cpbMax;
    if (CPB < 1) CPB = 1;
    const size_t smem = planeBytes * CPB;
    const int grid = ceil_div(C, CPB);
    if (kk <= 32) {
        D2T_SMEM_OPTIN(rp4_bwd_kernel<1>, smem);
        rp4_bwd_kernel<1><<<grid, CPB * 32, smem, st>>>(go, rec, gin, R, C, H, W, kk, CPB);
    } else {
        D2T_SMEM_OPTIN(rp4_bwd_kernel<2>, smem);
        rp4_bwd_kernel<2><<<grid, CPB * 32, smem, st>>>(go, rec, gin, R, C, H, W, kk, CPB);
    }
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch(2);
    return D2T_OK;
}

}  // namespace d2t
