// pool_ps.cu -- float32 PSROIPool forward/backward over a BATCH of frames (N >= 1), channel-owner CTAs.  sm_100a.
//
// The per-output kernels in pool.cu put the target index on the lanes, so every load instruction touches 32
// different channel planes (32 L1 wavefronts per LDG) and a call is a handful of small latency-bound launches:
// 20 + 90 us per frame for the R-FCN class head (31 targets, 300 RoIs), 16 frames per training step on one GPU.
// Here one launch covers all frames and a CTA owns one (frame, channel) plane:
//
//   forward   the CTA copies its plane into shared memory (coalesced), finds the (target, bin) pairs that read this
//             channel -- the reference map (t+1)*(i*k+j) is many-to-one, SURVEY.md F6 -- and lets a lane take one
//             RoI: bin edges come from a packed table, the bin is summed rows-then-columns out of shared memory in
//             the reference's order (bit-identical results, ps_roipool_cuda.cu:60-69).
//   backward  row-owner difference arrays, no atomics (reference: atomicAdd per bin pixel, ps_roipool_cuda.cu:124-139):
//               1. edges    packed bin edges of every (frame, RoI, bin index), plus a bin-major copy
//               2. scale    Vt[n][b][t][r] = grad_out[n][r][t][b] / cell size      (r fastest: coalesced later)
//               3. rowlists for every (frame, bin row i, pixel row y): the RoIs whose bin row i contains y, ascending
//                           (the bin geometry is shared by all targets and all bin columns: 7 x H lists per frame)
//               4. main     CTA = (frame, channel), a THREAD owns a pixel row of the plane.  For each (target, bin)
//                           pair that reads the channel, in ascending bin order, the thread walks its row list and
//                           applies  D[y][J0] += v, D[y][J1] -= v  (v, J0, J1 of the RoI staged in shared memory);
//                           an inclusive scan along x then turns D into the gradient row, written out coalesced.
//                           One owner per row, fixed order => bitwise reproducible; every pixel written once,
//                           channels nobody reads are zero-filled.
#include <stdlib.h>

#include "common.cuh"

namespace d2t {

constexpr int kPsbThreads = 256;
constexpr int kPsbFwdThreads = 128;

__device__ __forceinline__ uint32_t psb_pack_edges(const float* __restrict__ roi, int b, int k, int H, int W) {
    int i0, i1, j0, j1;
    bin_edge<float, false>(roi[0], roi[2], b, k, H, i0, i1);
    bin_edge<float, false>(roi[1], roi[3], b, k, W, j0, j1);
    return (uint32_t)i0 | ((uint32_t)i1 << 8) | ((uint32_t)j0 << 16) | ((uint32_t)j1 << 24);
}

// edges[(n*R + r)*k + b] = I0 | I1<<8 | J0<<16 | J1<<24 of bin index b (row edges from H, column edges from W);
// edgesT[(n*k + b)*R + r] = the same word, bin-major (optional)
__global__ void __launch_bounds__(256)
psb_edges_kernel(const float* __restrict__ rois, uint32_t* __restrict__ edges, uint32_t* __restrict__ edgesT, int N,
                 int R, int k, int H, int W) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * R * k; idx += gridDim.x * blockDim.x) {
        const int nr = idx / k, b = idx - nr * k;
        const uint32_t e = psb_pack_edges(rois + (size_t)nr * 4, b, k, H, W);
        edges[idx] = e;
        if (edgesT) {
            const int n = nr / R, r = nr - n * R;
            edgesT[((size_t)n * k + b) * R + r] = e;
        }
    }
}

// Users of channel ch, ascending bin index, into shared memory: us[u] = t << 16 | b.  Returns the count (uniform).
// Reference map: ch = (t+1)*b.  ch == 0 <=> b == 0 for EVERY target: reported as one merged user with t = 0xFFFF.
__device__ __forceinline__ int psb_users(int ch, int nT, int kk, bool canonical, uint32_t* us, int* cnt) {
    if (canonical) {
        if (threadIdx.x == 0) {
            const int t = ch / kk;
            us[0] = ((uint32_t)t << 16) | (uint32_t)(ch - t * kk);
        }
        __syncthreads();
        return 1;
    }
    if (ch == 0) {
        if (threadIdx.x == 0) us[0] = 0xFFFF0000u;
        __syncthreads();
        return 1;
    }
    // thread b tests bin b; ordered compaction with warp ballots (kk <= blockDim.x is guaranteed by the host)
    const int b = threadIdx.x;
    bool is = false;
    int t = 0;
    if (b >= 1 && b < kk && ch % b == 0) {
        t = ch / b - 1;
        is = t < nT;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, is);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) cnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
    const int nw = (kk + 31) >> 5;
    for (int w = 0; w < nw; ++w) {
        const int c = cnt[w];
        if (w < warp) base += c;
        total += c;
    }
    if (is) us[base + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)t << 16) | (uint32_t)b;
    __syncthreads();
    return total;
}

// ----------------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------------
// order[n*R + p] = p-th RoI of frame n in ascending (largest bin height, largest bin width) order.  The 32 RoIs a warp of the
// forward kernel walks together then have cells of nearly the same shape, so the trip counts of its two loops agree across
// the lanes (unsorted, a warp runs the tallest cell's rows times the widest cell's columns: 3.5x the mean).  One CTA per
// frame, bitonic sort of (key << 16 | RoI) words in shared memory; R <= kPsbSortMax, else the identity.
constexpr int kPsbSortThreads = 512;
constexpr int kPsbSortMax = 2048;
__global__ void __launch_bounds__(kPsbSortThreads)
psb_order_kernel(const uint32_t* __restrict__ edges, uint16_t* __restrict__ order, int R, int k, int M) {
    __shared__ uint32_t key[kPsbSortMax];
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int r = tid; r < M; r += kPsbSortThreads) {
        uint32_t w = 0xFFFFFFFFu;
        if (r < R) {
            int hb = 0, wb = 0;
            for (int b = 0; b < k; ++b) {
                const uint32_t e = __ldg(edges + ((size_t)n * R + r) * k + b);
                hb = max(hb, (int)((e >> 8) & 255) - (int)(e & 255));
                wb = max(wb, (int)(e >> 24) - (int)((e >> 16) & 255));
            }
            w = ((uint32_t)(min(hb, 15) * 16 + min(wb, 15)) << 16) | (uint32_t)r;   // cells beyond 15 pixels share a class
        }
        key[r] = w;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= M; k2 <<= 1) {
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (int t = tid; t < M / 2; t += kPsbSortThreads) {
                const int lo = ((t & ~(j2 - 1)) << 1) | (t & (j2 - 1)), hi = lo | j2;
                const uint32_t a = key[lo], b = key[hi];
                const bool up = (lo & k2) == 0;
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < R; r += kPsbSortThreads) order[(size_t)n * R + r] = (uint16_t)(key[r] & 0xffffu);
}

__global__ void __launch_bounds__(kPsbFwdThreads)
psb_fwd_kernel(const float* __restrict__ fm, const uint32_t* __restrict__ edges, const uint16_t* __restrict__ order,
               float* __restrict__ out, int R, int nT, int H, int W, int k, int canonical) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* plane = reinterpret_cast<float*>(smem_raw);           // [H*W]
    uint32_t* us = reinterpret_cast<uint32_t*>(plane + H * W);   // [kk]
    int* cnt = reinterpret_cast<int*>(us + k * k);               // [8]
    const int kk = k * k;
    const int nCh = nT * kk;
    const int ch = blockIdx.x, n = blockIdx.y;
    const int HW = H * W;
    const int nU = psb_users(ch, nT, kk, canonical != 0, us, cnt);
    if (nU == 0) return;
    const float* src = fm + ((size_t)n * nCh + ch) * HW;
    for (int idx = threadIdx.x; idx < HW; idx += kPsbFwdThreads) plane[idx] = __ldg(src + idx);
    __syncthreads();
    const uint32_t* ed = edges + (size_t)n * R * k;
    const uint16_t* ord = order ? order + (size_t)n * R : nullptr;
    float* o = out + (size_t)n * R * nCh;
    for (int u = 0; u < nU; ++u) {
        const uint32_t pk = us[u];
        const int t = pk >> 16, b = pk & 0xffff;
        const int i = b / k, j = b - i * k;
        for (int p = threadIdx.x; p < R; p += kPsbFwdThreads) {
            const int r = ord ? (int)__ldg(ord + p) : p;
            const uint32_t ei = __ldg(ed + r * k + i), ej = __ldg(ed + r * k + j);
            const int i0 = ei & 255, i1 = (ei >> 8) & 255, j0 = (ej >> 16) & 255, j1 = ej >> 24;
            float acc = 0.f;
            // rows-then-columns, ONE accumulator: the reference's order (ps_roipool_cuda.cu:60-66).  Cells are a few pixels
            // wide, so the row loop is written out -- blocks of four loads, then the 2 / 1 tail -- instead of leaving the
            // compiler's unroll-by-four with its remainder dispatch (14 instructions per pixel in the ncu source view)
            const int wdt = j1 - j0;
            const float* row = plane + i0 * W + j0;
#pragma unroll 1
            for (int pi = i0; pi < i1; ++pi, row += W) {
                const float* q = row;
                int w4 = wdt;
#pragma unroll 1
                for (; w4 >= 4; w4 -= 4, q += 4) {
                    const float a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3];
                    acc += a0;
                    acc += a1;
                    acc += a2;
                    acc += a3;
                }
                if (w4 & 2) {
                    const float a0 = q[0], a1 = q[1];
                    acc += a0;
                    acc += a1;
                    q += 2;
                }
                if (w4 & 1) acc += q[0];
            }
            const int numel = (i1 - i0) * wdt;
            if (numel > 0) acc /= numel;
            if (t == 0xFFFF) {  // channel 0 of the reference map: bin 0 of every target reads it
                for (int tt = 0; tt < nT; ++tt) o[((size_t)r * nT + tt) * kk] = acc;
            } else {
                o[((size_t)r * nT + t) * kk + b] = acc;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// backward 2: Vt[((n*kk + b)*nT + t)*R + r] = grad_out[((n*R + r)*nT + t)*kk + b] / cell size
// ----------------------------------------------------------------------------------------------------
// grid (ceil(R/32), nT, N), 128 threads: a 32-RoI x kk tile goes through shared memory
constexpr int kPsbScaleThreads = 128;
__global__ void __launch_bounds__(kPsbScaleThreads)
psb_scale_kernel(const float* __restrict__ go, const uint32_t* __restrict__ edges, float* __restrict__ vt, int R, int nT,
                 int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int kk = k * k;
    const int pitch = kk | 1;
    float* tile = reinterpret_cast<float*>(smem_raw);                 // [32][kk | 1]
    short* hS = reinterpret_cast<short*>(tile + 32 * pitch);          // [32][k] bin heights
    short* wS = hS + 32 * k;                                          // [32][k] bin widths
    unsigned char* bi = reinterpret_cast<unsigned char*>(wS + 32 * k);  // [kk] bin row of bin b
    const int r0 = blockIdx.x * 32, t = blockIdx.y, n = blockIdx.z;
    const int nr = min(32, R - r0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NWARP = kPsbScaleThreads / 32;
    for (int e = threadIdx.x; e < nr * k; e += kPsbScaleThreads) {
        const uint32_t w = __ldg(edges + ((size_t)n * R + r0) * k + e);
        hS[e] = (short)((int)((w >> 8) & 255) - (int)(w & 255));
        wS[e] = (short)((int)(w >> 24) - (int)((w >> 16) & 255));
    }
    for (int b = threadIdx.x; b < kk; b += kPsbScaleThreads) bi[b] = (unsigned char)(b / k);
    __syncthreads();
    // a warp takes whole RoIs: its lanes read the RoI's kk gradients as one contiguous run
    for (int rr = warp; rr < nr; rr += NWARP) {
        const float* src = go + (((size_t)n * R + r0 + rr) * nT + t) * kk;
        for (int b = lane; b < kk; b += 32) {
            const int i = bi[b], j = b - i * k;
            const int numel = (int)hS[rr * k + i] * (int)wS[rr * k + j];
            float v = __ldg(src + b);
            if (numel > 0) v /= numel;  // ps_roipool_cuda.cu:134-137
            tile[rr * pitch + b] = v;
        }
    }
    __syncthreads();
    if (lane < nr)
        for (int b = warp; b < kk; b += NWARP)
            vt[(((size_t)n * kk + b) * nT + t) * R + r0 + lane] = tile[lane * pitch + b];
}

// ----------------------------------------------------------------------------------------------------
// backward 3: row lists.  One warp per (frame n, bin row i, pixel row y)
// ----------------------------------------------------------------------------------------------------
// rowlist[((n*k + i)*R + e)*H + y] = e-th RoI (ascending) whose bin row i contains pixel row y   (y fastest: the row
// threads of the main kernel read entry e of their lists with one coalesced load);  rowcnt[(n*k + i)*H + y] = length
__global__ void __launch_bounds__(kPsbThreads)
psb_rowlists_kernel(const uint32_t* __restrict__ edgesT, uint16_t* __restrict__ rowlist, int* __restrict__ rowcnt, int N,
                    int R, int H, int k) {
    const int lane = threadIdx.x & 31;
    const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int l = warpGlobal; l < N * k * H; l += nWarps) {
        const int ni = l / H, y = l - ni * H;  // ni = n*k + i
        const uint32_t* ed = edgesT + (size_t)ni * R;
        uint16_t* list = rowlist + (size_t)ni * R * H + y;
        int cnt = 0;
        for (int r0 = 0; r0 < R; r0 += 32) {
            const int r = r0 + lane;
            bool in = false;
            if (r < R) {
                const uint32_t e = __ldg(ed + r);
                in = (int)(e & 255) <= y && y < (int)((e >> 8) & 255);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) list[(size_t)(cnt + __popc(bal & ((1u << lane) - 1u))) * H] = (uint16_t)r;
            cnt += __popc(bal);
        }
        if (lane == 0) rowcnt[l] = cnt;
    }
}

// ----------------------------------------------------------------------------------------------------
// backward 4: main.  grid (nCh, N): CTA = (channel, frame), 64 threads, a thread owns pixel rows
// ----------------------------------------------------------------------------------------------------
constexpr int kPsbRowThreads = 64;
__global__ void __launch_bounds__(kPsbRowThreads)
psb_bwd_kernel(const float* __restrict__ vt, const uint16_t* __restrict__ rowlist, const int* __restrict__ rowcnt,
               const uint32_t* __restrict__ edgesT, float* __restrict__ gin, int R, int nT, int H, int W, int k,
               int canonical) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int kk = k * k;
    const int pitch = (W + 1) | 1;  // odd: rows land in different banks
    float* D = reinterpret_cast<float*>(smem_raw);                 // [H][pitch]
    uint2* vj = reinterpret_cast<uint2*>(D + ((H * pitch + 1) & ~1));  // [R] {value bits, packed edges of the user's bin column}
    uint32_t* us = reinterpret_cast<uint32_t*>(vj + R);            // [kk]
    int* cnt = reinterpret_cast<int*>(us + kk);                    // [8]
    const int nCh = nT * kk;
    const int ch = blockIdx.x, n = blockIdx.y;
    const int HW = H * W;
    float* dst = gin + ((size_t)n * nCh + ch) * HW;
    const int nU = psb_users(ch, nT, kk, canonical != 0, us, cnt);
    if (nU == 0) {
        for (int px = threadIdx.x; px < HW; px += kPsbRowThreads) dst[px] = 0.f;
        return;
    }
    for (int idx = threadIdx.x; idx < H * pitch; idx += kPsbRowThreads) D[idx] = 0.f;
    // thread -> row map: warp 0 takes even rows, warp 1 odd rows (both warps busy when H < 64)
    const int yFirst = (threadIdx.x & 31) * 2 + (threadIdx.x >> 5);
    for (int u = 0; u < nU; ++u) {
        const uint32_t pk = us[u];
        const int t = pk >> 16, b = pk & 0xffff;
        const int i = b / k, j = b - i * k;
        __syncthreads();  // the previous user's values are no longer read (first time: D is zeroed)
        for (int r = threadIdx.x; r < R; r += kPsbRowThreads) {
            float v;
            if (t == 0xFFFF) {  // merged channel 0: sum over targets, ascending (fixed order)
                v = 0.f;
                for (int tt = 0; tt < nT; ++tt) v += __ldg(vt + (((size_t)n * kk) * nT + tt) * R + r);
            } else {
                v = __ldg(vt + (((size_t)n * kk + b) * nT + t) * R + r);
            }
            vj[r] = make_uint2(__float_as_uint(v), __ldg(edgesT + ((size_t)n * k + j) * R + r));
        }
        __syncthreads();
        const uint16_t* lists = rowlist + ((size_t)n * k + i) * R * H;
        const int* cnts = rowcnt + ((size_t)n * k + i) * H;
        for (int y = yFirst; y < H; y += kPsbRowThreads) {
            float* row = D + y * pitch;
            const int c = __ldg(cnts + y);
            const uint16_t* lp = lists + y;
            for (int e = 0; e < c; e += 4) {  // four entries in flight: list loads, then value loads, then the updates
                int r[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) r[s4] = (e + s4 < c) ? (int)__ldg(lp + (e + s4) * H) : -1;
                uint2 q[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) q[s4] = r[s4] >= 0 ? vj[r[s4]] : make_uint2(0u, 0u);
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int j0 = (q[s4].y >> 16) & 255, j1 = q[s4].y >> 24;
                    if (j1 > j0) {  // an empty bin column receives nothing (and (a + v) - v need not give a back)
                        const float v = __uint_as_float(q[s4].x);
                        row[j0] += v;
                        row[j1] -= v;
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int y = yFirst; y < H; y += kPsbRowThreads) {  // inclusive scan along x
        float* row = D + y * pitch;
        float acc = 0.f;
#pragma unroll 4
        for (int x = 0; x < W; ++x) {
            acc += row[x];
            row[x] = acc;
        }
    }
    __syncthreads();
    for (int x0 = 0; x0 < W; x0 += kPsbRowThreads) {  // thread = column: conflict-free reads, coalesced stores
        const int x = x0 + threadIdx.x;
        if (x < W)
            for (int y = 0; y < H; ++y) dst[y * W + x] = D[y * pitch + x];
    }
}

// ----------------------------------------------------------------------------------------------------
// backward, second generation: ONE launch, no workspace.  grid (nCh, N): CTA = (channel, frame), 128 threads.
// ----------------------------------------------------------------------------------------------------
// Round 1's backward (psb_edges -> psb_scale -> psb_rowlists -> psb_bwd_kernel, kept below as the fallback for RoI counts
// whose bitmasks do not fit shared memory) spent its time in latency: four launches, row lists and scaled gradients
// fetched from global memory in dependent steps, one thread per pixel row doing a 63-step serial scan.  Here a CTA
// derives everything it needs itself:
//   per user (target t, bin b = (i, j)) of its channel
//     A  every thread takes RoIs r = tid, tid + 128, ...: bin-row-i and bin-column-j edges with the reference's
//        expression (ps_roipool_cuda.cu:42-54), v = grad_out[r, t, b] / cell size, stored as {v, j0 | j1 << 16};
//        the RoI sets bit r of rowmask[y] for its rows y in [i0, i1) with atomicOr -- an INTEGER atomic, so the result
//        does not depend on arrival order
//     B  task (y, sign): walks the set bits of rowmask[y] in ascending RoI order and applies  Dp[y][j0] += v  (sign +)
//        or  Dm[y][j1] += v  (sign -): one owner per (row, sign), fixed order => bitwise reproducible, no FP atomics
//        (reference: atomicAdd per bin pixel, ps_roipool_cuda.cu:124-139)
//   then  grad[y][x] = sum_{x' <= x} (Dp[y][x'] - Dm[y][x'])  by a warp-shuffle scan, written coalesced.
// Channels nobody reads (911 of 1519 for the class head) are zero-filled by their CTA.
constexpr int kPsb2Threads = 128;
constexpr int kPsb2UB = 8;   // users whose gradients are prefetched together (one global-memory latency per batch)
__global__ void __launch_bounds__(kPsb2Threads)
psb2_bwd_kernel(const float* __restrict__ go, const float* __restrict__ rois, float* __restrict__ gin, int R, int nT, int H,
                int W, int k, int canonical, int vote) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int kk = k * k;
    const int pitch = W + 1;
    const int MW = (R + 31) >> 5;                                     // mask words per pixel row
    float4* roiS = reinterpret_cast<float4*>(smem_raw);               // [R] the frame's RoIs
    float* Dp = reinterpret_cast<float*>(roiS + R);                   // [H][pitch]
    float* Dm = Dp + H * pitch;                                       // [H][pitch]
    float* gv = Dm + H * pitch;                                       // [kPsb2UB][R] raw gradients of a batch of users
    uint32_t* rowmask = reinterpret_cast<uint32_t*>(gv + kPsb2UB * R);    // [H][MW]
    uint2* vj = reinterpret_cast<uint2*>(rowmask + ((H * MW + 1) & ~1));  // [R] {value bits, j0 | j1 << 16}
    uint32_t* us = reinterpret_cast<uint32_t*>(vj + R);               // [kk]
    int* cnt = reinterpret_cast<int*>(us + kk);                       // [8]
    const int nCh = nT * kk;
    const int ch = blockIdx.x, n = blockIdx.y;
    const int HW = H * W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* dst = gin + ((size_t)n * nCh + ch) * HW;
    const int nU = psb_users(ch, nT, kk, canonical != 0, us, cnt);
    if (nU == 0) {
        for (int px = tid; px < HW; px += kPsb2Threads) dst[px] = 0.f;
        return;
    }
    const float4* roiG = reinterpret_cast<const float4*>(rois) + (size_t)n * R;
    for (int r = tid; r < R; r += kPsb2Threads) roiS[r] = __ldg(roiG + r);
    for (int idx = tid; idx < 2 * H * pitch; idx += kPsb2Threads) Dp[idx] = 0.f;   // Dp and Dm are contiguous
    for (int idx = tid; idx < H * MW; idx += kPsb2Threads) rowmask[idx] = 0u;
    // vote != 0: `go` is the gradient of the VOTE (N, R, nT) -- the mean over the k x k bins (rfcn.py:41) -- so every bin of
    // (r, t) receives go[r, t] / kk; otherwise it is the gradient of the pooled tensor (N, R, nT, k, k)
    const float* goBase = go + (size_t)n * R * (vote ? nT : nCh);
    const int gstride = vote ? 1 : kk;
    const float gscale = vote ? 1.f / (float)kk : 1.f;
    for (int u0 = 0; u0 < nU; u0 += kPsb2UB) {
        const int nb = min(kPsb2UB, nU - u0);
        __syncthreads();   // gv of the previous batch is no longer read
        // all global loads of the batch are issued back to back: one memory latency, not one per user
        for (int e = tid; e < nb * R; e += kPsb2Threads) {
            const int uu = e / R, r = e - uu * R;
            const uint32_t pk = us[u0 + uu];
            const int t = pk >> 16, b = pk & 0xffff;
            float g = 0.f;
            if (t == 0xFFFF) {   // merged channel 0 of the reference map: bin 0 of every target, ascending target
#pragma unroll 4
                for (int tt = 0; tt < nT; ++tt) g += __ldg(goBase + ((size_t)r * nT + tt) * gstride);
            } else {
                g = __ldg(goBase + ((size_t)r * nT + t) * gstride + (vote ? 0 : b));
            }
            gv[uu * R + r] = g * gscale;
        }
        for (int uu = 0; uu < nb; ++uu) {
            const uint32_t pk = us[u0 + uu];
            const int b = pk & 0xffff;
            const int i = b / k, j = b - i * k;
            __syncthreads();   // gv visible; masks cleared; previous user's walk finished
            for (int r = tid; r < R; r += kPsb2Threads) {
                const float4 roi = roiS[r];
                int i0, i1, j0, j1;
                bin_edge<float, false>(roi.x, roi.z, i, k, H, i0, i1);
                bin_edge<float, false>(roi.y, roi.w, j, k, W, j0, j1);
                const int numel = (i1 - i0) * (j1 - j0);
                float v = 0.f;
                if (numel > 0) {
                    v = gv[uu * R + r] / numel;   // ps_roipool_cuda.cu:134-137 (the merged channel 0 divides the target sum)
                    const uint32_t bit = 1u << (r & 31);
                    for (int y = i0; y < i1; ++y) atomicOr(&rowmask[y * MW + (r >> 5)], bit);
                }
                vj[r] = make_uint2(__float_as_uint(v), (uint32_t)j0 | ((uint32_t)j1 << 16));
            }
            __syncthreads();
            // (one task per row doing both updates was measured slower: 99 vs 90 us single-frame class head)
            for (int task = tid; task < 2 * H; task += kPsb2Threads) {
                const int y = task >> 1, minus = task & 1;
                float* row = (minus ? Dm : Dp) + y * pitch;
                const uint32_t* mrow = rowmask + y * MW;
                for (int wd = 0; wd < MW; ++wd) {
                    uint32_t m = mrow[wd];
                    while (m) {
                        const int r = (wd << 5) + __ffs(m) - 1;
                        m &= m - 1;
                        const uint2 q = vj[r];
                        const int x = minus ? (int)(q.y >> 16) : (int)(q.y & 0xffff);
                        row[x] += __uint_as_float(q.x);
                    }
                }
            }
            if (u0 + uu + 1 < nU) {
                __syncthreads();
                for (int idx = tid; idx < H * MW; idx += kPsb2Threads) rowmask[idx] = 0u;
            }
        }
    }
    __syncthreads();
    // inclusive scan along x, a warp per row, 32 columns per step with a carry
    for (int y = warp; y < H; y += kPsb2Threads / 32) {
        const float* rp = Dp + y * pitch;
        const float* rm = Dm + y * pitch;
        float carry = 0.f;
        for (int x0 = 0; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            float v = x < W ? rp[x] - rm[x] : 0.f;
#pragma unroll
            for (int sh = 1; sh < 32; sh <<= 1) {
                const float o = __shfl_up_sync(0xffffffffu, v, sh);
                if (lane >= sh) v += o;
            }
            v += carry;
            if (x < W) dst[y * W + x] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
}

static size_t psb2_smem(int R, int H, int W, int k) {
    const int MW = (R + 31) >> 5;
    return (size_t)R * 16 + (size_t)2 * H * (W + 1) * 4 + (size_t)kPsb2UB * R * 4 + (size_t)((H * MW + 1) & ~1) * 4 +
           (size_t)R * 8 + (size_t)k * k * 4 + 64;
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
struct PsbLayout {
    size_t edgesOff, orderOff, edgesTOff, vtOff, listOff, cntOff, total;
};
static PsbLayout psb_layout(int N, int R, int nT, int H, int W, int k, bool bwd) {
    (void)W;
    PsbLayout L;
    size_t off = 0;
    L.edgesOff = off;
    off += align_up((size_t)N * R * k * sizeof(uint32_t), 256);
    L.orderOff = off;
    if (!bwd) off += align_up((size_t)N * R * sizeof(uint16_t), 256);
    L.edgesTOff = off;
    if (bwd) off += align_up((size_t)N * R * k * sizeof(uint32_t), 256);
    L.vtOff = off;
    if (bwd) off += align_up((size_t)N * k * k * nT * R * sizeof(float), 256);
    L.listOff = off;
    if (bwd) off += align_up((size_t)N * k * R * H * sizeof(uint16_t), 256);
    L.cntOff = off;
    if (bwd) off += align_up((size_t)N * k * H * sizeof(int), 256);
    L.total = off;
    return L;
}

bool psb_supported(int N, int R, int nT, int H, int W, int k) {
    if (N <= 0 || R <= 0 || nT <= 0 || k <= 0 || H <= 0 || W <= 0) return false;
    if (H > 255 || W > 255 || R >= 65534 || k * k > kPsbFwdThreads || nT >= 0xFFFF || N > 65535) return false;
    if ((long long)nT * k * k > 0x7fffffffLL / 4) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    const size_t fwdSmem = (size_t)H * W * 4 + (size_t)k * k * 4 + 64;
    const size_t bwdSmem = (size_t)H * ((W + 1) | 1) * 4 + 8 + (size_t)R * 8 + (size_t)k * k * 4 + 64;
    const size_t cap = (size_t)di.max_smem_optin;
    return fwdSmem <= cap && bwdSmem <= cap;
}

static bool psb2_ok(int R, int H, int W, int k);
// the round-1 row-list BACKWARD finds the users of a channel with one thread per bin and runs kPsbRowThreads threads
// (the forward runs kPsbFwdThreads): r_hw <= 8.  (r_hw = 9 used to reach it and read past its tables -- found by the
// r_hw = 9 cases of test_psroipool_batched_equals_per_frame.)
static bool psb_rowlist_bwd_ok(int N, int R, int nT, int H, int W, int k) {
    return k * k <= kPsbRowThreads && psb_supported(N, R, nT, H, W, k);
}
// third generation (pool_ps3.cu): targets on the lanes, CTA = (frame, pixel row, column block); the default wherever it applies
bool psb3_supported(int N, int R, int nT, int H, int W, int k);
size_t psb3_ws_bytes(int N, int R, int nT, int H, int W, int k);
int psb3_bwd_launch(const float*, const float*, float*, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t);
// Where the third generation runs (measured on B200, 300 RoIs on 38x63; profiles/r2_time_psroipool_v3.txt): a single frame
// always (class head 39 us against 91 us for psb2 and 58 us for the reference's atomic kernel; box head 22.5 / 40 / 26 us);
// a batch of frames when the targets fill at least half a warp (class head, 16 frames: 213 us against 257 us), while for
// few targets its per-CTA fixed work outweighs the row-list kernels (box head, 16 frames: 97 us against 71 us).
static bool psb3_use(int N, int R, int nT, int H, int W, int k) {
    if (!psb3_supported(N, R, nT, H, W, k)) return false;
    return N == 1 || nT > 8 || !psb_rowlist_bwd_ok(N, R, nT, H, W, k);
}
// Which of the OLDER backward kernels runs otherwise: the one-launch kernel for a single frame (latency matters: 87 us against 116 us for the four
// round-1 launches at the class-head size) and for maps the round-1 kernels cannot pack (H or W above 255); a BATCH of
// frames fills the chip and is throughput-bound, where the round-1 row-list kernels execute fewer instructions
// (254 us against 400 us for 16 frames; profiles/r2_ncu_psb2_summary.txt).
static bool psb2_use(int N, int R, int nT, int H, int W, int k) {
    return psb2_ok(R, H, W, k) && (N == 1 || !psb_rowlist_bwd_ok(N, R, nT, H, W, k));
}

bool psb_bwd_supported(int N, int R, int nT, int H, int W, int k) {
    if (N <= 0 || R <= 0 || nT <= 0 || k <= 0 || H <= 0 || W <= 0 || N > 65535 || nT >= 0xFFFF) return false;
    if ((long long)nT * k * k > 0x7fffffffLL / 4 || (long long)R * nT * k * k >= (1ll << 31)) return false;
    return psb3_supported(N, R, nT, H, W, k) || psb2_ok(R, H, W, k) || psb_rowlist_bwd_ok(N, R, nT, H, W, k);
}

size_t psb_ws_bytes(int N, int R, int nT, int H, int W, int k, bool bwd) {
    if (bwd && psb3_use(N, R, nT, H, W, k)) return psb3_ws_bytes(N, R, nT, H, W, k);
    if (bwd && psb2_use(N, R, nT, H, W, k)) return 0;   // the one-launch backward needs no workspace
    return psb_layout(N, R, nT, H, W, k, bwd).total;
}

static int psb_edges_launch(const float* rois, uint32_t* edges, uint32_t* edgesT, int N, int R, int k, int H, int W,
                            cudaStream_t st) {
    const int total = N * R * k;
    psb_edges_kernel<<<ceil_div(total, 256), 256, 0, st>>>(rois, edges, edgesT, N, R, k, H, W);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

int psb_fwd_launch(const float* fm, const float* rois, float* out, int N, int R, int nT, int H, int W, int k, int flags,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
    const PsbLayout L = psb_layout(N, R, nT, H, W, k, false);
    if (!ws || ws_bytes < L.total) {
        set_error("psroipool_fwd (batched): workspace too small (%zu < %zu bytes)", ws_bytes, L.total);
        return D2T_ERR_WORKSPACE;
    }
    uint32_t* edges = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + L.edgesOff);
    int rc = psb_edges_launch(rois, edges, nullptr, N, R, k, H, W, st);
    if (rc) return rc;
    // cell-size order of the RoIs: pays once a frame's live planes outnumber the sort -- a batch of class-head frames
    // (16 frames, 31 targets: 140 -> 124 us with the written-out row loop; 4 targets: 40 -> 42 us, so not there).
    // Always bit-identical: every output is computed by one lane in the reference's order
    uint16_t* order = nullptr;
    if (N > 1 && nT >= 8 && R > 32 && R <= kPsbSortMax) {
        order = reinterpret_cast<uint16_t*>(static_cast<char*>(ws) + L.orderOff);
        int M = 64;
        while (M < R) M <<= 1;
        psb_order_kernel<<<N, kPsbSortThreads, 0, st>>>(edges, order, R, k, M);
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
    }
    const size_t smem = (size_t)H * W * 4 + (size_t)k * k * 4 + 64;
    D2T_SMEM_OPTIN(psb_fwd_kernel, smem);
    psb_fwd_kernel<<<dim3(nT * k * k, N), kPsbFwdThreads, smem, st>>>(fm, edges, order, out, R, nT, H, W, k,
                                                                     (flags & D2T_PS_CANONICAL_MAP) ? 1 : 0);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

// ----------------------------------------------------------------------------------------------------
// PSROIPool + vote (rfcn.py:40-41: pooled.mean(-1).mean(-1)) in one pass: out[n, r, t] = mean over the k x k bins of the
// bin means.  One warp per (frame, RoI, target); a lane takes bins b = lane, lane + 32, ..., sums each bin rows-then-
// columns from its channel plane (L2-resident) and the warp adds the bin means with a fixed-shape shuffle tree.  The
// (R, nT, k, k) tensor and the two reduction kernels of the composition never exist.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
psb_vote_fwd_kernel(const float* __restrict__ fm, const float* __restrict__ rois, float* __restrict__ out, int N, int R, int nT,
                    int H, int W, int k, int canonical) {
    const int kk = k * k, nCh = nT * kk;
    const int lane = threadIdx.x & 31;
    const long long total = (long long)N * R * nT;
    for (long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < total;
         wid += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int t = (int)(wid % nT);
        const long long nr = wid / nT;       // n * R + r
        const int n = (int)(nr / R);
        const float* roi = rois + nr * 4;
        const float r0 = __ldg(roi), r1 = __ldg(roi + 1), r2 = __ldg(roi + 2), r3 = __ldg(roi + 3);
        const float* base = fm + (size_t)n * nCh * H * W;
        float s = 0.f;
        for (int b = lane; b < kk; b += 32) {
            const int i = b / k, j = b - i * k;
            int i0, i1, j0, j1;
            bin_edge<float, false>(r0, r2, i, k, H, i0, i1);
            bin_edge<float, false>(r1, r3, j, k, W, j0, j1);
            const float* ch = base + (size_t)(canonical ? t * kk + b : (t + 1) * b) * H * W;
            float acc = 0.f;
            for (int pi = i0; pi < i1; ++pi)
                for (int pj = j0; pj < j1; ++pj) acc += __ldg(ch + pi * W + pj);
            const int numel = (i1 - i0) * (j1 - j0);
            if (numel > 0) acc /= numel;   // empty cell -> 0 (ps_roipool_cuda.cu:67-69)
            s += acc;
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
        if (lane == 0) out[wid] = s / (float)kk;
    }
}

bool psb_vote_supported(int N, int R, int nT, int H, int W, int k) {
    if (N <= 0 || R <= 0 || nT <= 0 || k <= 0 || H <= 0 || W <= 0 || N > 65535 || nT >= 0xFFFF) return false;
    if ((long long)R * nT * k * k >= (1ll << 31) || (long long)nT * k * k * H * W >= (1ll << 31)) return false;
    return psb2_ok(R, H, W, k);
}

int psb_vote_fwd_launch(const float* fm, const float* rois, float* out, int N, int R, int nT, int H, int W, int k, int flags,
                        cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const long long warps = (long long)N * R * nT;
    long long grid = (warps + 7) / 8;
    if (grid > (long long)di.sm_count * 32) grid = (long long)di.sm_count * 32;
    psb_vote_fwd_kernel<<<(int)grid, 256, 0, st>>>(fm, rois, out, N, R, nT, H, W, k, (flags & D2T_PS_CANONICAL_MAP) ? 1 : 0);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

size_t psb_vote_bwd_ws_bytes(int N, int R, int nT, int H, int W, int k) {
    return psb3_supported(N, R, nT, H, W, k) ? psb3_ws_bytes(N, R, nT, H, W, k) : 0;
}

int psb_vote_bwd_launch(const float* go, const float* rois, float* gin, int N, int R, int nT, int H, int W, int k, int flags,
                        void* ws, size_t ws_bytes, cudaStream_t st) {
    if (psb3_supported(N, R, nT, H, W, k)) return psb3_bwd_launch(go, rois, gin, N, R, nT, H, W, k, flags, 1, ws, ws_bytes, st);
    const size_t smem2 = psb2_smem(R, H, W, k);
    D2T_SMEM_OPTIN(psb2_bwd_kernel, smem2);
    psb2_bwd_kernel<<<dim3(nT * k * k, N), kPsb2Threads, smem2, st>>>(go, rois, gin, R, nT, H, W, k,
                                                                      (flags & D2T_PS_CANONICAL_MAP) ? 1 : 0, 1);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

// the one-launch backward needs its RoI bitmasks and both difference arrays in shared memory, a few CTAs per SM
static bool psb2_ok(int R, int H, int W, int k) {
    return W < 65535 && k * k <= kPsb2Threads && psb2_smem(R, H, W, k) <= 56 * 1024;
}

int psb_bwd_launch(const float* go, const float* rois, float* gin, int N, int R, int nT, int H, int W, int k, int flags,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
    if (psb3_use(N, R, nT, H, W, k)) return psb3_bwd_launch(go, rois, gin, N, R, nT, H, W, k, flags, 0, ws, ws_bytes, st);
    if (psb2_use(N, R, nT, H, W, k)) {
        const size_t smem2 = psb2_smem(R, H, W, k);
        D2T_SMEM_OPTIN(psb2_bwd_kernel, smem2);
        psb2_bwd_kernel<<<dim3(nT * k * k, N), kPsb2Threads, smem2, st>>>(go, rois, gin, R, nT, H, W, k,
                                                                          (flags & D2T_PS_CANONICAL_MAP) ? 1 : 0, 0);
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
        return D2T_OK;
    }
    const PsbLayout L = psb_layout(N, R, nT, H, W, k, true);
    if (!ws || ws_bytes < L.total) {
        set_error("psroipool_bwd (batched): workspace too small (%zu < %zu bytes)", ws_bytes, L.total);
        return D2T_ERR_WORKSPACE;
    }
    char* base = static_cast<char*>(ws);
    uint32_t* edges = reinterpret_cast<uint32_t*>(base + L.edgesOff);
    uint32_t* edgesT = reinterpret_cast<uint32_t*>(base + L.edgesTOff);
    float* vt = reinterpret_cast<float*>(base + L.vtOff);
    uint16_t* rowlist = reinterpret_cast<uint16_t*>(base + L.listOff);
    int* rowcnt = reinterpret_cast<int*>(base + L.cntOff);
    const bool canonical = (flags & D2T_PS_CANONICAL_MAP) != 0;
    const int kk = k * k;
    int rc = psb_edges_launch(rois, edges, edgesT, N, R, k, H, W, st);
    if (rc) return rc;
    psb_scale_kernel<<<dim3(ceil_div(R, 32), nT, N), kPsbScaleThreads, (size_t)32 * (kk | 1) * 4 + (size_t)64 * k * 2 + kk + 16, st>>>(
        go, edges, vt, R, nT, k);
    D2T_CUDA_TRY(cudaGetLastError());
    const int lists = N * k * H;
    psb_rowlists_kernel<<<ceil_div(lists, kPsbThreads / 32), kPsbThreads, 0, st>>>(edgesT, rowlist, rowcnt, N, R, H, k);
    D2T_CUDA_TRY(cudaGetLastError());
    const size_t smem = (size_t)H * ((W + 1) | 1) * 4 + 8 + (size_t)R * 8 + (size_t)kk * 4 + 64;
    D2T_SMEM_OPTIN(psb_bwd_kernel, smem);
    psb_bwd_kernel<<<dim3(nT * kk, N), kPsbRowThreads, smem, st>>>(vt, rowlist, rowcnt, edgesT, gin, R, nT, H, W, k,
                                                                  canonical ? 1 : 0);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch(3);
    return D2T_OK;
}

}  // namespace d2t
