// pool.cu -- ROIPool / PSROIPool forward + backward for sm_100a.
//
// Replaces roipool/roipool_cuda.cu and ps_roipool/ps_roipool_cuda.cu of the
// reference (one thread per output element, global-memory accumulators,
// atomicAdd backward).  Design here:
//
//  ROIPool fwd  : "channel-owner" CTAs.  A CTA loads a slab of CB channel planes
//                 into shared memory ONCE (coalesced), then walks all RoIs and
//                 emits out[r, c0:c0+CB, :, :], so the feature map is read from HBM
//                 once.  1024 threads per CTA; a thread owns one (RoI, channel,
//                 column-bin) and produces its r_hw outputs, summing bin pixels
//                 rows-then-columns -- the reference's order -- so float results
//                 are bit-identical to the reference kernel.
//  ROIPool bwd  : same slab ownership, and inside the CTA every thread OWNS one pixel
//                 column of one channel for ALL RoIs.  It walks the RoIs in order
//                 and adds each RoI's contribution to its own column: no atomics, no
//                 barrier per RoI, a fixed summation order (deterministic), and
//                 grad_fm written exactly once (no zero-fill pass).
//  PSROIPool fwd: one thread per output, target index fastest so the lanes of a
//                 warp share a bin (same trip count).
//  PSROIPool bwd: pixel-owner gather.  A tiny prep pass builds (a) the inverse of the
//                 many-to-one channel map (which (target, bin) pairs read channel ch:
//                 SURVEY.md F6) and (b) per bin row / bin column index, bitmasks over
//                 RoIs ("which RoIs' bin i covers pixel row y").  Each output pixel
//                 ANDs a row mask with a column mask and visits the surviving RoIs in
//                 ascending order: deterministic, no atomics, every pixel written
//                 exactly once; channels nobody reads are just zero-filled.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

// ---- fast division by a runtime constant (n < 2^31) ---------------------------
struct FastDiv {
    uint32_t d, m, s;
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1u << l) < d) ++l;
    f.s = l;
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
    return (__umulhi(f.m, n) + n) >> f.s;
}

constexpr int kPoolThreads = 256;
constexpr int kSlabThreads = 1024;  // channel-owner kernels: one CTA per SM, 32 warps to hide latency
constexpr int kMaxK = 32;  // largest supported r_hw
__host__ __device__ __forceinline__ size_t align_up_dev(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- bin edges for a chunk of RoIs into shared memory -------------------------
template <typename T, bool kClampStart>
__device__ __forceinline__ void edges_to_smem(const T* __restrict__ rois, int r0, int nr, int k, int H, int W,
                                              short* sI0, short* sI1, short* sJ0, short* sJ1) {
    for (int idx = threadIdx.x; idx < nr * k; idx += blockDim.x) {
        const int rr = idx / k, b = idx - rr * k;
        const T* roi = rois + (size_t)(r0 + rr) * 4;
        int e0, e1;
        bin_edge<T, kClampStart>(roi[0], roi[2], b, k, H, e0, e1);
        sI0[idx] = (short)e0;
        sI1[idx] = (short)e1;
        bin_edge<T, kClampStart>(roi[1], roi[3], b, k, W, e0, e1);
        sJ0[idx] = (short)e0;
        sJ1[idx] = (short)e1;
    }
}

// =================================================================================
// ROIPool forward
// =================================================================================
// smem: T plane[CB][H*W] | short edges[4][RCH*k]
template <typename T>
__global__ void __launch_bounds__(kSlabThreads)
roipool_fwd_kernel(const T* __restrict__ fm, const T* __restrict__ rois, T* __restrict__ out, int R, int C, int H,
                   int W, int k, int CB, int RCH, FastDiv dCBk, FastDiv dk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* plane = reinterpret_cast<T*>(smem_raw);
    const int HW = H * W;
    short* sI0 = reinterpret_cast<short*>(smem_raw + align_up_dev((size_t)CB * HW * sizeof(T), 16));
    short* sI1 = sI0 + RCH * k;
    short* sJ0 = sI1 + RCH * k;
    short* sJ1 = sJ0 + RCH * k;

    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);
    const int kk = k * k;

    // slab load: cb*HW contiguous elements
    {
        const T* src = fm + (size_t)c0 * HW;
        const int n = cb * HW;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) plane[idx] = __ldg(src + idx);
    }

    for (int rbase = 0; rbase < R; rbase += RCH) {
        const int nr = min(RCH, R - rbase);
        __syncthreads();
        edges_to_smem<T, true>(rois, rbase, nr, k, H, W, sI0, sI1, sJ0, sJ1);
        __syncthreads();

        // item = (roi, channel, column bin); column bin fastest
        const int total = nr * CB * k;
        for (int item = threadIdx.x; item < total; item += blockDim.x) {
            const int rr = fdiv(item, dCBk);
            const int rem = item - rr * (CB * k);
            const int cc = fdiv(rem, dk);
            if (cc >= cb) continue;
            const int j = rem - cc * k;
            const int j0 = sJ0[rr * k + j], j1 = sJ1[rr * k + j];
            const T* p = plane + (size_t)cc * HW;
            T* o = out + ((size_t)(rbase + rr) * C + c0 + cc) * kk + j;
            for (int i = 0; i < k; ++i) {
                const int i0 = sI0[rr * k + i], i1 = sI1[rr * k + i];
                T acc = 0;
                for (int pi = i0; pi < i1; ++pi) {
                    const T* row = p + pi * W;
                    for (int pj = j0; pj < j1; ++pj) acc += row[pj];
                }
                const int numel = (i1 - i0) * (j1 - j0);
                acc /= numel;  // no empty-bin guard: 0/0 = NaN like roipool_cuda.cu:61
                o[i * k] = acc;
            }
        }
    }
}

// =================================================================================
// ROIPool backward
// =================================================================================
// smem: T acc[CB][H*W] | T inv[RCH][k*k] | short edges[4][RCH*k]
template <typename T>
__global__ void __launch_bounds__(kSlabThreads)
roipool_bwd_kernel(const T* __restrict__ go, const T* __restrict__ rois, T* __restrict__ gin, int R, int C, int H,
                   int W, int k, int CB, int RCH, FastDiv dW, int bandRows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* acc = reinterpret_cast<T*>(smem_raw);
    // blockIdx.y selects a band of `bandRows` pixel rows (one band = the whole plane unless the plane is too large for
    // shared memory): the CTA then owns rows [y0, y0 + HB) of its channels and clips every bin to them
    const int y0 = blockIdx.y * bandRows;
    const int HB = min(bandRows, H - y0);
    const int HW = HB * W;   // pixels of the band
    const int kk = k * k;
    T* sinv = reinterpret_cast<T*>(smem_raw + align_up_dev((size_t)CB * bandRows * W * sizeof(T), 16));
    short* sI0 = reinterpret_cast<short*>(reinterpret_cast<unsigned char*>(sinv) + align_up_dev((size_t)RCH * kk * sizeof(T), 16));
    short* sI1 = sI0 + RCH * k;
    short* sJ0 = sI1 + RCH * k;
    short* sJ1 = sJ0 + RCH * k;

    const int c0 = blockIdx.x * CB;
    const int cb = min(CB, C - c0);

    for (int idx = threadIdx.x; idx < cb * HW; idx += blockDim.x) acc[idx] = 0;

    for (int rbase = 0; rbase < R; rbase += RCH) {
        const int nr = min(RCH, R - rbase);
        __syncthreads();
        edges_to_smem<T, true>(rois, rbase, nr, k, H, W, sI0, sI1, sJ0, sJ1);
        __syncthreads();
        for (int idx = threadIdx.x; idx < nr * kk; idx += blockDim.x) {
            const int rr = idx / kk, b = idx - rr * kk;
            const int i = b / k, j = b - i * k;
            const int numel = (sI1[rr * k + i] - sI0[rr * k + i]) * (sJ1[rr * k + j] - sJ0[rr * k + j]);
            sinv[idx] = static_cast<T>(1) / static_cast<T>(numel);
        }
        __syncthreads();

        // this thread owns pixel column pj of channel cc for every RoI: exclusive read-modify-write
        for (int item = threadIdx.x; item < cb * W; item += blockDim.x) {
            const int cc = fdiv(item, dW);
            const int pj = item - cc * W;
            T* a = acc + (size_t)cc * HW + pj;
            const T* gch = go + ((size_t)rbase * C + c0 + cc) * kk;
            for (int rr = 0; rr < nr; ++rr) {
                const short* J0 = sJ0 + rr * k;
                const short* J1 = sJ1 + rr * k;
                if (pj < J0[0] || pj >= J1[k - 1]) continue;
                // column bins covering pj form a contiguous range [jlo, jhi]
                int jlo = k, jhi = -1;
                for (int j = 0; j < k; ++j) {
                    if (J0[j] <= pj && pj < J1[j]) {
                        jlo = min(jlo, j);
                        jhi = j;
                    }
                }
                if (jhi < 0) continue;
                const T* g = gch + (size_t)rr * C * kk;
                const T* inv = sinv + rr * kk;
                const short* I0 = sI0 + rr * k;
                const short* I1 = sI1 + rr * k;
                for (int i = 0; i < k; ++i) {
                    const int r0 = max((int)I0[i], y0), r1 = min((int)I1[i], y0 + HB);
                    if (r0 >= r1) continue;
                    T u = 0;
                    for (int j = jlo; j <= jhi; ++j) u += __ldg(g + i * k + j) * inv[i * k + j];
                    for (int pi = r0; pi < r1; ++pi) a[(pi - y0) * W] += u;
                }
            }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < cb * HW; idx += blockDim.x) {
        const int cc = idx / HW, rem = idx - cc * HW;
        gin[((size_t)(c0 + cc) * H + y0) * W + rem] = acc[idx];
    }
}

// ROIPool forward straight from global memory, one thread per output element like the reference kernel
// (roipool_cuda.cu:6-63): only for planes too large for the shared-memory slab kernels.
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
roipool_fwd_global_kernel(const T* __restrict__ fm, const T* __restrict__ rois, T* __restrict__ out, long long total, int C,
                          int H, int W, int k) {
    const int kk = k * k;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx % kk);
        const long long rc = idx / kk;
        const int c = (int)(rc % C), r = (int)(rc / C);
        const int i = b / k, j = b - i * k;
        const T* roi = rois + (size_t)r * 4;
        int i0, i1, j0, j1;
        bin_edge<T, true>(roi[0], roi[2], i, k, H, i0, i1);
        bin_edge<T, true>(roi[1], roi[3], j, k, W, j0, j1);
        const T* p = fm + (size_t)c * H * W;
        T acc = 0;
        for (int pi = i0; pi < i1; ++pi)
            for (int pj = j0; pj < j1; ++pj) acc += __ldg(p + (size_t)pi * W + pj);
        const int numel = (i1 - i0) * (j1 - j0);
        acc /= numel;  // 0/0 = NaN like roipool_cuda.cu:61
        out[idx] = acc;
    }
}

// =================================================================================
// PSROIPool forward
// =================================================================================
__device__ __forceinline__ int ps_channel(int t, int i, int j, int k, bool canonical) {
    // reference map (ps_roipool_cuda.cu:58): (t+1)*(i*k+j)   [SURVEY.md F6]
    return canonical ? (t * k * k + i * k + j) : (t + 1) * (i * k + j);
}

template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
psroipool_fwd_kernel(const T* __restrict__ fm, const T* __restrict__ rois, T* __restrict__ out, int R, int nT,
                     int H, int W, int k, bool canonical, FastDiv dnT, FastDiv dk, FastDiv dkk) {
    const int kk = k * k;
    const int total = R * kk * nT;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        // idx = ((r*k + i)*k + j)*nT + t   (t fastest: a warp shares bins)
        const int rb = fdiv(idx, dnT);
        const int t = idx - rb * nT;
        const int r = fdiv(rb, dkk);
        const int b = rb - r * kk;
        const int i = fdiv(b, dk);
        const int j = b - i * k;
        const T* roi = rois + (size_t)r * 4;
        int i0, i1, j0, j1;
        bin_edge<T, false>(roi[0], roi[2], i, k, H, i0, i1);
        bin_edge<T, false>(roi[1], roi[3], j, k, W, j0, j1);
        const T* ch = fm + (size_t)ps_channel(t, i, j, k, canonical) * H * W;
        T acc = 0;
        for (int pi = i0; pi < i1; ++pi)
            for (int pj = j0; pj < j1; ++pj) acc += __ldg(ch + pi * W + pj);
        const int numel = (i1 - i0) * (j1 - j0);
        if (numel > 0) acc /= numel;
        out[((size_t)r * nT + t) * kk + b] = acc;
    }
}

// =================================================================================
// PSROIPool backward: prep (edges + RoI bitmasks) and pixel-owner gather
// =================================================================================
// workspace layout (see psroipool_bwd_ws_layout):
//   short4-like edges: eI0,eI1,eJ0,eJ1 : [R*k] int16 each
//   rowmask : [k][H][NW] uint32   bit r%32 of word r/32 set iff I0[r][i] <= y < I1[r][i]
//   colmask : [k][W][NW] uint32
struct PsBwdWs {
    short4* edges;    // [R][k]       {I0, I1, J0, J1} of row-bin b / column-bin b
    uint2* list;      // [k*k][H][R]  per (bin b, pixel row y): ascending RoIs whose row-bin covers y and whose
                      //              column-bin is non-empty, packed {r, J0 | J1 << 16}
    int* cnt;         // [k*k][H]
    int* entCount;    // [nCh]        how many (target, bin) pairs read channel ch
    uint32_t* ent;    // [nCh][nT]    those pairs, ascending target: (t << 16) | bin ; t = 0xFFFF => "all targets"
    void* gs;         // [R][nT][k*k] grad_out / cell size (the value each covered pixel receives)
    void* gs0;        // [R]          sum over targets of gs[r,t,0]  (reference map: every target's bin 0 reads channel 0)
};

// launch 1 (tiny): bin edges of every RoI, computed once
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
psroipool_bwd_edges_kernel(const T* __restrict__ rois, PsBwdWs ws, int R, int H, int W, int k) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * k; idx += gridDim.x * blockDim.x) {
        const int r = idx / k, b = idx - r * k;
        const T* roi = rois + (size_t)r * 4;
        int i0, i1, j0, j1;
        bin_edge<T, false>(roi[0], roi[2], b, k, H, i0, i1);
        bin_edge<T, false>(roi[1], roi[3], b, k, W, j0, j1);
        ws.edges[idx] = make_short4((short)i0, (short)i1, (short)j0, (short)j1);
    }
}

// launch 2: inverse channel map, pre-scaled gradients, per-(bin, row) RoI lists
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
psroipool_bwd_prep_kernel(const T* __restrict__ go, PsBwdWs ws, int R, int nT, int H, int W, int k, bool canonical) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    const int kk = k * k;
    // (a) inverse channel map: the reference map (t+1)*(i*k+j) is many-to-one
    for (int ch = tid; ch < nT * kk; ch += nth) {
        int n = 0;
        uint32_t* e = ws.ent + (size_t)ch * nT;
        if (canonical) {
            const int t = ch / kk;
            e[n++] = ((uint32_t)t << 16) | (uint32_t)(ch - t * kk);
        } else if (ch == 0) {
            e[n++] = 0xFFFF0000u;  // bin 0 of EVERY target reads channel 0: one merged entry (gs0)
        } else {
            for (int t = 0; t < nT; ++t) {
                if (ch % (t + 1) == 0) {
                    const int sidx = ch / (t + 1);
                    if (sidx < kk) e[n++] = ((uint32_t)t << 16) | (uint32_t)sidx;
                }
            }
        }
        ws.entCount[ch] = n;
    }
    // (b) pre-scaled gradients: gs[r,t,b] = grad_out[r,t,b] / cell size   (ps_roipool_cuda.cu:134-137)
    T* gs = static_cast<T*>(ws.gs);
    for (int idx = tid; idx < R * nT * kk; idx += nth) {
        const int b = idx % kk, r = idx / (nT * kk);
        const int i = b / k, j = b - i * k;
        const short4 ei = ws.edges[r * k + i], ej = ws.edges[r * k + j];
        const int numel = (ei.y - ei.x) * (ej.w - ej.z);
        T v = __ldg(go + idx);
        if (numel > 0) v /= numel;
        gs[idx] = v;
    }
    // (b') merged value for channel 0 of the reference map: one warp per RoI, lanes over targets, fixed-order
    //      shuffle tree (deterministic)
    T* gs0 = static_cast<T*>(ws.gs0);
    {
        const int lane0 = threadIdx.x & 31;
        const int wg = tid >> 5, nw = nth >> 5;
        for (int r = wg; r < R; r += nw) {
            const short4 e = ws.edges[r * k];
            const int numel = (e.y - e.x) * (e.w - e.z);
            T sum = 0;
            for (int t = lane0; t < nT; t += 32) {
                T v = __ldg(go + ((size_t)r * nT + t) * kk);
                if (numel > 0) v /= numel;
                sum += v;
            }
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
            if (lane0 == 0) gs0[r] = sum;
        }
    }
    // (c) per (bin b, pixel row y): ascending list of the RoIs that can contribute, with their column range.
    //     One warp per list; ballot compaction keeps the order deterministic.
    const int lane = threadIdx.x & 31;
    const int warpGlobal = tid >> 5, nWarps = nth >> 5;
    for (int l = warpGlobal; l < kk * H; l += nWarps) {
        const int b = l / H, y = l - b * H;
        const int i = b / k, j = b - i * k;
        uint2* list = ws.list + (size_t)l * R;
        int n = 0;
        for (int r0 = 0; r0 < R; r0 += 32) {
            const int r = r0 + lane;
            bool in = false;
            int j0 = 0, j1 = 0;
            if (r < R) {
                const short4 ei = ws.edges[r * k + i], ej = ws.edges[r * k + j];
                j0 = ej.z; j1 = ej.w;
                in = ei.x <= y && y < ei.y && j1 > j0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in)
                list[n + __popc(bal & ((1u << lane) - 1))] = make_uint2((unsigned)r, (unsigned)j0 | ((unsigned)j1 << 16));
            n += __popc(bal);
        }
        if (lane == 0) ws.cnt[l] = n;
    }
}

// grid: (H * ceil(W/64), nChannels), ONE warp per block: (part of) one pixel row of one channel, a lane owns the
// pixels x and x + 32.  For every (target, bin) pair that reads this channel the warp walks the ascending list of
// RoIs that cover this row in that bin; a lane adds the RoI's pre-scaled gradient if its column lies in the RoI's
// column range.  Control flow is uniform, the order of additions per pixel is fixed => deterministic, no atomics;
// every pixel is written exactly once.
constexpr int kPsRowSpan = 64;
template <typename T>
__global__ void __launch_bounds__(32)
psroipool_bwd_kernel(PsBwdWs ws, T* __restrict__ gin, int R, int nT, int H, int W, int k, int colTiles) {
    const int ch = blockIdx.y;
    const int kk = k * k;
    const int y = blockIdx.x / colTiles;
    const int xa = (blockIdx.x - y * colTiles) * kPsRowSpan + threadIdx.x, xb = xa + 32;
    T* dst = gin + (size_t)ch * H * W + (size_t)y * W;
    const int nEnt = ws.entCount[ch];
    if (nEnt == 0) {
        if (xa < W) dst[xa] = 0;
        if (xb < W) dst[xb] = 0;
        return;
    }
    const uint32_t* ent = ws.ent + (size_t)ch * nT;
    T acca = 0, accb = 0;
    for (int en = 0; en < nEnt; ++en) {
        const uint32_t pk = __ldg(ent + en);
        const int t = pk >> 16, b = pk & 0xffff;
        const int l = b * H + y;
        const int cnt = __ldg(ws.cnt + l);
        const uint2* list = ws.list + (size_t)l * R;
        if (t == 0xFFFF) {
            const T* g = static_cast<const T*>(ws.gs0);
#pragma unroll 4
            for (int n = 0; n < cnt; ++n) {
                const uint2 e = __ldg(list + n);
                const int j0 = e.y & 0xffff, j1 = e.y >> 16;
                const T v = __ldg(g + e.x);
                if (xa >= j0 && xa < j1) acca += v;
                if (xb >= j0 && xb < j1) accb += v;
            }
        } else {
            const T* g = static_cast<const T*>(ws.gs) + (size_t)t * kk + b;
#pragma unroll 4
            for (int n = 0; n < cnt; ++n) {
                const uint2 e = __ldg(list + n);
                const int j0 = e.y & 0xffff, j1 = e.y >> 16;
                const T v = __ldg(g + (size_t)e.x * nT * kk);
                if (xa >= j0 && xa < j1) acca += v;
                if (xb >= j0 && xb < j1) accb += v;
            }
        }
    }
    if (xa < W) dst[xa] = acca;
    if (xb < W) dst[xb] = accb;
}

// =================================================================================
// bin-edge instrumentation
// =================================================================================
template <typename T>
__global__ void pool_bins_kernel(const T* __restrict__ rois, int32_t* __restrict__ edges, int R, int H, int W, int k,
                                 int clampStart) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * k; idx += gridDim.x * blockDim.x) {
        const int r = idx / k, b = idx - r * k;
        const T* roi = rois + (size_t)r * 4;
        int i0, i1, j0, j1;
        if (clampStart) {
            bin_edge<T, true>(roi[0], roi[2], b, k, H, i0, i1);
            bin_edge<T, true>(roi[1], roi[3], b, k, W, j0, j1);
        } else {
            bin_edge<T, false>(roi[0], roi[2], b, k, H, i0, i1);
            bin_edge<T, false>(roi[1], roi[3], b, k, W, j0, j1);
        }
        int32_t* e = edges + (size_t)idx * 4;
        e[0] = i0;
        e[1] = i1;
        e[2] = j0;
        e[3] = j1;
    }
}

// =================================================================================
// host launchers
// =================================================================================
// float32 fast paths for r_hw <= 15 (pool_fast.cu)
bool roipool_fast_supported(int R, int C, int H, int W, int k);
int roipool_fast_bwd_launch(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
bool roipool_prefix_supported(int R, int C, int H, int W, int k);
int roipool_prefix_fwd_launch(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
// float32, r_hw = 7: [pixel][16 channel] slabs -- forward pool_vec.cu, backward pool_vec2.cu
bool roipool_vec_supported(int R, int C, int H, int W, int k);
int roipool_vec_fwd_launch(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
bool roipool_vec2_bwd_supported(int R, int C, int H, int W, int k);
int roipool_vec2_bwd_launch(const float*, const float*, float*, int, int, int, int, cudaStream_t);
template <typename T>
struct FastPath {
    static bool fwd(const T*, const T*, T*, int, int, int, int, int, cudaStream_t, int*) { return false; }
    static bool bwd(const T*, const T*, T*, int, int, int, int, int, cudaStream_t, int*) { return false; }
};
template <>
struct FastPath<float> {
    // forward: row-prefix kernels (rounding-level differences from the reference's summation order);
    // d2t_roipool_fwd_f32_exact bypasses them and keeps the bit-identical slab kernel below
    static bool fwd(const float* fm, const float* rois, float* out, int R, int C, int H, int W, int k, cudaStream_t st,
                    int* rc) {
        if (roipool_vec_supported(R, C, H, W, k)) {
            *rc = roipool_vec_fwd_launch(fm, rois, out, R, C, H, W, k, st);
            return true;
        }
        if (!roipool_prefix_supported(R, C, H, W, k)) return false;
        *rc = roipool_prefix_fwd_launch(fm, rois, out, R, C, H, W, k, st);
        return true;
    }
    static bool bwd(const float* go, const float* rois, float* gin, int R, int C, int H, int W, int k, cudaStream_t st,
                    int* rc) {
        if (roipool_vec2_bwd_supported(R, C, H, W, k)) {
            *rc = roipool_vec2_bwd_launch(go, rois, gin, R, C, H, W, st);
            return true;
        }
        if (!roipool_fast_supported(R, C, H, W, k)) return false;
        *rc = roipool_fast_bwd_launch(go, rois, gin, R, C, H, W, k, st);
        return true;
    }
};
struct SlabPlan {
    int CB;       // channels per CTA
    int RCH;      // RoIs per edge-table chunk
    size_t smem;  // dynamic shared memory bytes
    int grid;
    int bandRows; // pixel rows per CTA (H unless the plane does not fit shared memory) and number of bands
    int bands;
};

// Choose the channel slab so that the grid is one balanced wave when possible.
// shared memory = CB planes + per-RoI tables for a chunk of RCH RoIs (4 int16 edge arrays + `per_roi` bytes).
// D2T_ERR_WORKSPACE (without an error message) = the plane does not fit: the caller falls back (forward: global-memory
// kernel; backward: row bands, allow_bands)
static int plan_slab(int R, int C, int H, int W, int k, size_t elem, size_t per_roi, SlabPlan* plan, bool allow_bands = false) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const size_t budget = (size_t)di.max_smem_optin - 1024;
    int RCH = R < 256 ? (R > 0 ? R : 1) : 256;
    size_t perCB = (size_t)H * W * elem;
    auto tables = [&](int rch) { return align_up((size_t)rch * per_roi, 16) + align_up((size_t)4 * rch * k * sizeof(short), 16); };
    while (RCH > 8 && perCB + tables(RCH) > budget) RCH /= 2;
    plan->bandRows = H;
    plan->bands = 1;
    if (perCB + tables(RCH) > budget) {
        if (!allow_bands) return D2T_ERR_WORKSPACE;
        RCH = R < 64 ? (R > 0 ? R : 1) : 64;
        const size_t rowBytes = (size_t)W * elem;
        if (rowBytes + tables(RCH) > budget) {
            set_error("roipool: one %d-pixel row (%zu B) does not fit the %zu B shared-memory slab", W, rowBytes, budget);
            return D2T_ERR_BAD_ARG;
        }
        plan->bandRows = (int)((budget - tables(RCH)) / rowBytes);
        plan->bands = ceil_div(H, plan->bandRows);
        plan->bandRows = ceil_div(H, plan->bands);   // balanced bands
        perCB = (size_t)plan->bandRows * rowBytes;
    }
    int maxCB = (int)((budget - tables(RCH)) / perCB);
    int CB = ceil_div(C, di.sm_count);  // one wave, one CTA per SM
    if (CB > maxCB) {
        const int waves = ceil_div(ceil_div(C, maxCB), di.sm_count);  // several waves: balance them
        CB = ceil_div(C, waves * di.sm_count);
        if (CB > maxCB) CB = maxCB;
    }
    if (CB < 1) CB = 1;
    plan->CB = CB;
    plan->RCH = RCH;
    plan->smem = align_up((size_t)CB * perCB, 16) + tables(RCH);
    plan->grid = ceil_div(C, CB);
    return 0;
}

template <typename T>
int roipool_fwd_launch(const T* fm, const T* rois, T* out, int R, int C, int H, int W, int k, cudaStream_t st,
                       bool force_exact) {
    D2T_REQUIRE(R >= 0 && C >= 0 && H > 0 && W > 0 && k > 0 && k <= kMaxK, "roipool_fwd: bad shape R=%d C=%d H=%d W=%d r_hw=%d", R,
                C, H, W, k);
    D2T_REQUIRE(H < 32768 && W < 32768, "roipool_fwd: H, W must be < 32768");
    if (R == 0 || C == 0) return D2T_OK;
    int frc = 0;
    if (!force_exact && FastPath<T>::fwd(fm, rois, out, R, C, H, W, k, st, &frc)) return frc;
    SlabPlan p;
    int rc = plan_slab(R, C, H, W, k, sizeof(T), 0, &p);
    if (rc == D2T_ERR_WORKSPACE) {   // plane larger than shared memory: per-output kernel on global memory (same order => same bits)
        const long long total = (long long)R * C * k * k;
        DeviceInfo di;
        if ((rc = device_info(&di))) return rc;
        long long grid = (total + kPoolThreads - 1) / kPoolThreads;
        if (grid > (long long)di.sm_count * 16) grid = (long long)di.sm_count * 16;
        roipool_fwd_global_kernel<T><<<(int)grid, kPoolThreads, 0, st>>>(fm, rois, out, total, C, H, W, k);
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
        return D2T_OK;
    }
    if (rc) return rc;
    D2T_SMEM_OPTIN(roipool_fwd_kernel<T>, p.smem);
    roipool_fwd_kernel<T><<<p.grid, kSlabThreads, p.smem, st>>>(fm, rois, out, R, C, H, W, k, p.CB, p.RCH,
                                                                 make_fastdiv(p.CB * k), make_fastdiv(k));
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

template <typename T>
int roipool_bwd_launch(const T* go, const T* rois, T* gin, int R, int C, int H, int W, int k, cudaStream_t st) {
    D2T_REQUIRE(R >= 0 && C >= 0 && H > 0 && W > 0 && k > 0 && k <= kMaxK, "roipool_bwd: bad shape R=%d C=%d H=%d W=%d r_hw=%d", R,
                C, H, W, k);
    D2T_REQUIRE(H < 32768 && W < 32768, "roipool_bwd: H, W must be < 32768");
    if (C == 0) return D2T_OK;
    if (R == 0) {
        D2T_CUDA_TRY(cudaMemsetAsync(gin, 0, (size_t)C * H * W * sizeof(T), st));
        return D2T_OK;
    }
    int frc = 0;
    if (FastPath<T>::bwd(go, rois, gin, R, C, H, W, k, st, &frc)) return frc;
    SlabPlan p;
    int rc = plan_slab(R, C, H, W, k, sizeof(T), (size_t)k * k * sizeof(T), &p, true);
    if (rc) return rc;
    D2T_SMEM_OPTIN(roipool_bwd_kernel<T>, p.smem);
    roipool_bwd_kernel<T><<<dim3(p.grid, p.bands), kSlabThreads, p.smem, st>>>(go, rois, gin, R, C, H, W, k, p.CB, p.RCH,
                                                                                make_fastdiv(W), p.bandRows);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

template <typename T>
int psroipool_fwd_launch(const T* fm, const T* rois, T* out, int R, int nT, int H, int W, int k, int flags,
                         cudaStream_t st) {
    D2T_REQUIRE(R >= 0 && nT > 0 && H > 0 && W > 0 && k > 0 && k <= kMaxK, "psroipool_fwd: bad shape R=%d nT=%d H=%d W=%d r_hw=%d",
                R, nT, H, W, k);
    if (R == 0) return D2T_OK;
    const long long total = (long long)R * k * k * nT;
    D2T_REQUIRE(total < (1ll << 31), "psroipool_fwd: output too large");
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    int grid = (int)((total + kPoolThreads - 1) / kPoolThreads);
    const int cap = di.sm_count * 8;
    if (grid > cap) grid = cap;
    psroipool_fwd_kernel<T><<<grid, kPoolThreads, 0, st>>>(fm, rois, out, R, nT, H, W, k,
                                                            (flags & D2T_PS_CANONICAL_MAP) != 0, make_fastdiv(nT),
                                                            make_fastdiv(k), make_fastdiv(k * k));
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

static size_t psroipool_bwd_ws_layout(int R, int nT, int H, int W, int k, size_t elem, void* base, PsBwdWs* ws) {
    const size_t Rn = R > 0 ? R : 1;
    const size_t kk = (size_t)k * k;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    size_t oedge = take(Rn * k * sizeof(short4));
    size_t olist = take(kk * H * Rn * sizeof(uint2)), ocnt2 = take(kk * H * sizeof(int));
    size_t ocnt = take((size_t)nT * kk * sizeof(int)), oent = take((size_t)nT * kk * nT * sizeof(uint32_t));
    size_t ogs = take(Rn * nT * kk * elem), ogs0 = take(Rn * elem);
    if (ws) {
        char* b = static_cast<char*>(base);
        ws->edges = (short4*)(b + oedge);
        ws->list = (uint2*)(b + olist);
        ws->cnt = (int*)(b + ocnt2);
        ws->entCount = (int*)(b + ocnt);
        ws->ent = (uint32_t*)(b + oent);
        ws->gs = b + ogs;
        ws->gs0 = b + ogs0;
    }
    return off;
}

size_t psroipool_bwd_ws_bytes(int R, int nT, int H, int W, int k, int elem) {
    return psroipool_bwd_ws_layout(R, nT, H, W, k, (size_t)elem, nullptr, nullptr);
}

template <typename T>
int psroipool_bwd_launch(const T* go, const T* rois, T* gin, int R, int nT, int H, int W, int k, int flags, void* wsp,
                         size_t ws_bytes, cudaStream_t st) {
    D2T_REQUIRE(R >= 0 && nT > 0 && H > 0 && W > 0 && k > 0 && k <= kMaxK, "psroipool_bwd: bad shape R=%d nT=%d H=%d W=%d r_hw=%d",
                R, nT, H, W, k);
    D2T_REQUIRE(H < 32768 && W < 32768, "psroipool_bwd: H, W must be < 32768");
    const int nCh = nT * k * k;
    D2T_REQUIRE(nCh <= 65535, "psroipool_bwd: n_targets*r_hw^2 must be <= 65535");
    D2T_REQUIRE((long long)R * nCh < (1ll << 31), "psroipool_bwd: grad_out too large");
    if (R == 0) {
        D2T_CUDA_TRY(cudaMemsetAsync(gin, 0, (size_t)nCh * H * W * sizeof(T), st));
        return D2T_OK;
    }
    const size_t need = psroipool_bwd_ws_bytes(R, nT, H, W, k, (int)sizeof(T));
    if (wsp == nullptr || ws_bytes < need) {
        set_error("psroipool_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
        return D2T_ERR_WORKSPACE;
    }
    PsBwdWs ws;
    psroipool_bwd_ws_layout(R, nT, H, W, k, sizeof(T), wsp, &ws);
    DeviceInfo di;
    int drc = device_info(&di);
    if (drc) return drc;
    psroipool_bwd_edges_kernel<T><<<ceil_div(R * k, kPoolThreads), kPoolThreads, 0, st>>>(rois, ws, R, H, W, k);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    const int prepItems = R * nCh;
    int prepGrid = ceil_div(prepItems, kPoolThreads);
    if (prepGrid > di.sm_count * 8) prepGrid = di.sm_count * 8;
    psroipool_bwd_prep_kernel<T><<<prepGrid, kPoolThreads, 0, st>>>(go, ws, R, nT, H, W, k,
                                                                     (flags & D2T_PS_CANONICAL_MAP) != 0);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    const int colTiles = ceil_div(W, kPsRowSpan);
    dim3 grid(H * colTiles, nCh);
    psroipool_bwd_kernel<T><<<grid, 32, 0, st>>>(ws, gin, R, nT, H, W, k, colTiles);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

template <typename T>
int pool_bins_launch(const T* rois, int32_t* edges, int R, int H, int W, int k, int clampStart, cudaStream_t st) {
    D2T_REQUIRE(R >= 0 && H > 0 && W > 0 && k > 0, "pool_bins: bad shape");
    if (R == 0) return D2T_OK;
    pool_bins_kernel<T><<<ceil_div(R * k, 256), 256, 0, st>>>(rois, edges, R, H, W, k, clampStart);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

// explicit instantiations used by api.cu
template int roipool_fwd_launch<float>(const float*, const float*, float*, int, int, int, int, int, cudaStream_t, bool);
template int roipool_fwd_launch<double>(const double*, const double*, double*, int, int, int, int, int, cudaStream_t, bool);
template int roipool_bwd_launch<float>(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
template int roipool_bwd_launch<double>(const double*, const double*, double*, int, int, int, int, int, cudaStream_t);
template int psroipool_fwd_launch<float>(const float*, const float*, float*, int, int, int, int, int, int, cudaStream_t);
template int psroipool_fwd_launch<double>(const double*, const double*, double*, int, int, int, int, int, int,
                                          cudaStream_t);
template int psroipool_bwd_launch<float>(const float*, const float*, float*, int, int, int, int, int, int, void*, size_t,
                                         cudaStream_t);
template int psroipool_bwd_launch<double>(const double*, const double*, double*, int, int, int, int, int, int, void*,
                                          size_t, cudaStream_t);
template int pool_bins_launch<float>(const float*, int32_t*, int, int, int, int, int, cudaStream_t);
template int pool_bins_launch<double>(const double*, int32_t*, int, int, int, int, int, cudaStream_t);

}  // namespace d2t
