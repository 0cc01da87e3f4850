// pool_ps3.cu -- float32 PSROIPool backward, third generation: TARGETS on the lanes, a CTA per (frame, pixel row, block of
// 16 or 32 pixel columns).  sm_100a.  Reference: atomicAdd per bin pixel, ps_roipool_cuda.cu:76-141.
//
// The bin geometry of a RoI does not depend on the target t, only the gradient value and the destination channel
// (t+1)*(i*k+j) do.  The channel-owner kernels of pool_ps.cu could not use that: the users of ONE channel are
// (target, bin) pairs with different bins, hence unrelated geometry, so every user was walked serially by a few threads
// (87 us for one frame of the class head against 56 us for the reference's atomic kernel, 254 us for 16 frames).  Here the
// contraction is split the other way round:
//
//   prep   (one CTA per (frame, RoI))  bin edges with the reference's expression; the scaled gradients transposed to
//          gT[n][r][b][t] = grad_out[n][r][t][b] / cell size   (t fastest, padded to TP = 2^ceil(log2 nT) lanes): the 31
//          targets of a (RoI, bin) become ONE coalesced 128-byte load; zeros for the channels nobody reads (911 of 1519
//          for the class head), which the main kernel then never visits
//   main   CTA = (frame n, pixel row y, columns x0 .. x0+XB-1), shared memory U[b][x][t]: the gradient row of every USER plane
//            A  the RoIs that touch (y, x block) are listed once, ascending (ordered ballot compaction);  a warp owns bins
//               b = warp, warp + XB, ...: the listed RoIs whose bin row i(b) contains y and whose bin column j(b) meets the
//               block are gathered into the warp's buffer, then lane t adds gT[n][r][b][t] to U[b][x][t] for the cell's
//               columns x -- uniform trip counts (the geometry is the same for all lanes), one owner per U element,
//               ascending RoI order, four gradient loads in flight per warp
//            B  channel ch = sum over its users (t, b) with (t+1)*b == ch, ascending t, read from U through a table of
//               shared-memory offsets that the host computes per call and passes BY VALUE (kernel parameters: no device
//               table to build, nothing to cache between calls); XB lanes write the row segment of one channel
//
// No floating-point atomics, no difference arrays: bitwise reproducible, pixels outside every cell are exact zeros, a
// non-finite gradient stays inside its cell -- the properties of the reference's kernel that the scan-based kernels gave up.
// A batch of frames and a single frame give bit-identical results (same per-element order); nT <= 32 and r_hw <= 8,
// everything else keeps pool_ps.cu / pool.cu.
#include "common.cuh"

namespace d2t {

constexpr int kP3Buf = 64;        // entries a warp gathers before it flushes (+ 4 zero slots behind them)
constexpr int kP3PrepThreads = 256;
constexpr int kP3MaxK = 8;
constexpr int kP3MaxCh = 32 * kP3MaxK * kP3MaxK;   // 2048

static int p3_tp(int nT) {
    int tp = 1;
    while (tp < nT) tp <<= 1;
    return tp;
}

struct P3Layout {
    size_t erowOff, ecolOff, gtOff, total;
};
static P3Layout p3_layout(int N, int R, int nT, int k) {
    P3Layout L;
    size_t off = 0;
    L.erowOff = off;
    off += align_up((size_t)N * k * R * sizeof(uint32_t), 256);
    L.ecolOff = off;
    off += align_up((size_t)N * k * R * sizeof(uint32_t), 256);
    L.gtOff = off;
    off += align_up((size_t)N * R * k * k * p3_tp(nT) * sizeof(float), 256);
    L.total = off;
    return L;
}

// U[kk][XB][TP + 1]: the odd pitch makes both access patterns conflict-free (lanes = targets in A, lanes = columns in B)
static size_t p3_smem(int R, int nT, int k, int XB) {
    const int warps = XB;   // 16 columns: 16 warps, 32 columns: 32 warps
    return align_up((size_t)k * k * XB * (p3_tp(nT) + 1) * 4, 16) + (size_t)warps * (kP3Buf + 4) * 4 + (size_t)warps * 4 +
           (size_t)R * (4 + 2 * k) + 16;
}

// the channel map, per call, by value: live channels (those with users), and per live channel the shared-memory offsets
// b * XB * (TP + 1) + t of its users (t, b), ascending t
struct P3Tab {
    int nLive;
    uint32_t rec[kP3MaxCh];    // live channel l: first user (11 bits) | number of users << 11 (6 bits) | channel << 17
    uint32_t user[kP3MaxCh];   // byte offsets into U, grouped by channel, ascending t inside a channel
};

static void p3_build_tab(P3Tab& tab, int nT, int k, int canonical, int XB) {
    // counting sort of the nT * kk users (t, b) by channel, ascending t inside a channel; values are BYTE offsets into U
    const int kk = k * k, nCh = nT * kk, P = p3_tp(nT) + 1;
    uint16_t cnt[kP3MaxCh + 1] = {0};
    for (int t = 0; t < nT; ++t)
        for (int b = 0; b < kk; ++b) ++cnt[canonical ? t * kk + b : (t + 1) * b];
    uint16_t pos[kP3MaxCh];
    int nU = 0, nLive = 0;
    for (int ch = 0; ch < nCh; ++ch) {
        pos[ch] = (uint16_t)nU;
        if (cnt[ch]) tab.rec[nLive++] = (uint32_t)nU | ((uint32_t)cnt[ch] << 11) | ((uint32_t)ch << 17);
        nU += cnt[ch];
    }
    tab.nLive = nLive;
    for (int t = 0; t < nT; ++t)
        for (int b = 0; b < kk; ++b) tab.user[pos[canonical ? t * kk + b : (t + 1) * b]++] = (uint32_t)((b * XB * P + t) * 4);
}

// which targets read channel ch: bit t of the result (lanes = targets; every lane returns the mask)
__device__ __forceinline__ uint32_t p3_user_mask(int ch, int nT, int kk, int canonical, int lane) {
    bool is;
    if (canonical) is = ch / kk == lane;
    else if (ch == 0) is = true;
    else is = ch % (lane + 1) == 0 && ch / (lane + 1) < kk;
    return __ballot_sync(0xffffffffu, is && lane < nT);
}

// ----------------------------------------------------------------------------------------------------
// prep: edges, scaled + transposed gradients, zero fill of the channels nobody reads.  One CTA per (frame, RoI).
// ----------------------------------------------------------------------------------------------------
// erow[(n*k + i)*R + r] = I0 | I1 << 16 (bin row i), ecol[(n*k + j)*R + r] = J0 | J1 << 16 (bin column j): r fastest, so the
// main kernel's thread-per-RoI prologue loads are coalesced.
__global__ void __launch_bounds__(kP3PrepThreads)
psb3_prep_kernel(const float* __restrict__ go, const float* __restrict__ rois, uint32_t* __restrict__ erow,
                 uint32_t* __restrict__ ecol, float* __restrict__ gT, float* __restrict__ gin, int R, int nT, int TP, int tpShift,
                 int H, int W, int k, int canonical, int vote) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int kk = k * k, nCh = nT * kk;
    const int pitch = kk | 1;
    float* tile = reinterpret_cast<float*>(smem_raw);        // [nT][kk | 1]
    int* hS = reinterpret_cast<int*>(tile + nT * pitch);      // [k] bin heights
    int* wS = hS + k;                                         // [k] bin widths
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t nr = blockIdx.x;                             // n * R + r
    const int n = (int)(nr / R), r = (int)(nr - (size_t)n * R);
    const float* roi = rois + nr * 4;
    if (tid < 2 * k) {
        const bool col = tid >= k;
        const int b = col ? tid - k : tid;
        int e0, e1;
        if (col) bin_edge<float, false>(__ldg(roi + 1), __ldg(roi + 3), b, k, W, e0, e1);
        else bin_edge<float, false>(__ldg(roi), __ldg(roi + 2), b, k, H, e0, e1);
        (col ? wS : hS)[b] = e1 - e0;
        // an inverted (empty) range is stored as the empty range 0..0: the fields are unsigned 16-bit
        const uint32_t w = e1 > e0 ? ((uint32_t)e0 | ((uint32_t)e1 << 16)) : 0u;
        (col ? ecol : erow)[((size_t)n * k + b) * R + r] = w;
    }
    // channels r, r + R, ... of frame n that nobody reads: zeros, a warp per plane (the main kernel visits live channels only)
    {
        const int HW = H * W;
        for (int ch = r + warp * R; ch < nCh; ch += (kP3PrepThreads / 32) * R) {
            if (p3_user_mask(ch, nT, kk, canonical, lane) == 0u) {
                float* z = gin + ((size_t)n * nCh + ch) * HW;
                for (int px = lane; px < HW; px += 32) z[px] = 0.f;
            }
        }
    }
    __syncthreads();
    // vote != 0: `go` is the gradient of the VOTE (N, R, nT) -- the mean over the k x k bins (rfcn.py:41) -- so every bin of
    // (r, t) receives go[r, t] / kk; otherwise it is the gradient of the pooled tensor (N, R, nT, k, k)
    {
        const float gscale = 1.f / (float)kk;
        const float* src = go + nr * (size_t)(vote ? nT : nCh);
        const int b = tid & 63, g = tid >> 6;   // kk <= 64: thread (g, b) takes targets g, g + 4, ...
        if (b < kk) {
            const int i = b / k, j = b - i * k;
            const int hh = hS[i], ww = wS[j];
            const bool live = hh > 0 && ww > 0;
            const float fn = (float)(hh * ww);
            // nT <= 32: eight targets per thread, all loads requested before the first division (one memory latency)
            float gv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int t = g + q * (kP3PrepThreads / 64);
                gv[q] = (live && t < nT) ? (vote ? __ldg(src + t) * gscale : __ldg(src + t * kk + b)) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int t = g + q * (kP3PrepThreads / 64);
                if (t < nT) tile[t * pitch + b] = live ? gv[q] / fn : 0.f;   // ps_roipool_cuda.cu:134-137
            }
        }
    }
    __syncthreads();
    float* dst = gT + nr * (size_t)kk * TP;
    for (int idx = tid; idx < kk * TP; idx += kP3PrepThreads) {
        const int b = idx >> tpShift, t = idx & (TP - 1);
        dst[idx] = t < nT ? tile[t * pitch + b] : 0.f;
    }
}

// ----------------------------------------------------------------------------------------------------
// main
// ----------------------------------------------------------------------------------------------------
// TP lanes along the targets, XB pixel columns and XB warps per CTA
template <int TP, int XB>
__global__ void __launch_bounds__(XB * 32, XB == 16 ? 2 : 1)
psb3_bwd_kernel(const float* __restrict__ gT, const uint32_t* __restrict__ erow, const uint32_t* __restrict__ ecol,
                float* __restrict__ gin, int R, int nT, int H, int W, int k, int NXB, const __grid_constant__ P3Tab tab) {
    constexpr int XL = 32 / TP;       // lanes along x
    constexpr int NW = XB;            // warps
    constexpr int NT = NW * 32;
    constexpr int P = TP + 1;         // pitch of a pixel's targets
    constexpr int KM = kP3MaxK;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int kk = k * k, nCh = nT * kk;
    float* U = reinterpret_cast<float*>(smem_raw);                         // [kk][XB][P]
    const int uWords = (kk * XB * P + 3) & ~3;
    uint32_t* ebuf = reinterpret_cast<uint32_t*>(U + uWords);              // [warps][kP3Buf + 4] entries of the current bin
    int* wcnt = reinterpret_cast<int*>(ebuf + NW * (kP3Buf + 4));          // [warps]
    uint16_t* aR = reinterpret_cast<uint16_t*>(wcnt + NW);                 // [R] listed RoIs, ascending
    uint16_t* aRow = aR + R;                                              // [R] bit i: bin row i contains y
    uint16_t* aCol = aRow + R;                                            // [R][k] xa | xb << 8 (clipped to the block)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int y = blockIdx.x / NXB, x0 = (blockIdx.x - y * NXB) * XB, n = blockIdx.y;

    for (int idx = tid; idx < uWords / 4; idx += NT) reinterpret_cast<float4*>(U)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);

    // ---- the RoIs that touch (y, block), ascending.  All 2k edge words of a RoI are requested before the first is used
    // (one memory latency per round, not 2k)
    const uint32_t* er = erow + (size_t)n * k * R;
    const uint32_t* ec = ecol + (size_t)n * k * R;
    int nA = 0;
    for (int r0 = 0; r0 < R; r0 += NT) {
        const int r = r0 + tid;
        uint32_t rowbits = 0;
        bool any = false;
        uint32_t cw[KM];
        if (r0 + (warp << 5) < R) {   // warp-uniform: warps beyond the RoI count skip the edge arithmetic
            uint32_t rw[KM];
#pragma unroll
            for (int i = 0; i < KM; ++i) {
                const bool ok = r < R && i < k;
                rw[i] = ok ? __ldg(er + (size_t)i * R + r) : 0u;
                cw[i] = ok ? __ldg(ec + (size_t)i * R + r) : 0u;
            }
#pragma unroll
            for (int i = 0; i < KM; ++i) {
                if ((int)(rw[i] & 0xffff) <= y && y < (int)(rw[i] >> 16)) rowbits |= 1u << i;
                const int xa = min(max((int)(cw[i] & 0xffff) - x0, 0), XB), xb = min(max((int)(cw[i] >> 16) - x0, 0), XB);
                any |= xb > xa;
                cw[i] = (uint32_t)(xa | (xb << 8));
            }
        }
        const bool act = rowbits != 0 && any;
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int base = nA, total = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int c = wcnt[w];
            if (w < warp) base += c;
            total += c;
        }
        if (act) {
            const int a = base + __popc(bal & ((1u << lane) - 1u));
            aR[a] = (uint16_t)r;
            aRow[a] = (uint16_t)rowbits;
#pragma unroll
            for (int j = 0; j < KM; ++j)
                if (j < k) aCol[a * k + j] = (uint16_t)cw[j];
        }
        nA += total;
        __syncthreads();
    }

    // ---- A: lane = (column phase dx, target t); a warp owns whole bins.  The entries of a bin are gathered into the warp's
    // buffer first, so that their gradient loads go out four at a time whatever round of the list found them
    {
        const int t = lane & (TP - 1), dx = lane / TP;
        const int gstride = kk * TP;
        const float* gbase = gT + (size_t)n * R * gstride + t;
        uint32_t* buf = ebuf + warp * (kP3Buf + 4);
        int i = warp / k, j = warp - i * k;
        for (int b = warp; b < kk; b += NW) {
            float* Ub = U + (size_t)b * XB * P + dx * P + t;
            const float* gb = gbase + (size_t)b * TP;
            int cnt = 0;
            for (int a0 = 0; a0 < nA; a0 += 32) {
                const int a = a0 + lane;
                uint32_t ent = 0;
                bool in = false;
                if (a < nA) {
                    const uint32_t c = aCol[a * k + j];
                    in = ((aRow[a] >> i) & 1u) && (c >> 8) > (c & 255u);
                    ent = (uint32_t)aR[a] | (c << 16);   // r | xa << 16 | xb << 24
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) buf[cnt + __popc(m & ((1u << lane) - 1u))] = ent;   // cnt <= kP3Buf - 32 here
                cnt += __popc(m);
                // flush when another round might not fit, or after the last round
                if (cnt > kP3Buf - 32 || a0 + 32 >= nA) {
                    if (lane < 4) buf[cnt + lane] = 0u;   // empty entries (xa = xb = 0) pad the last group of four
                    __syncwarp();
                    for (int q0 = 0; q0 < cnt; q0 += 4) {
                        const uint32_t e0 = buf[q0], e1 = buf[q0 + 1], e2 = buf[q0 + 2], e3 = buf[q0 + 3];
                        const float v0 = __ldg(gb + (size_t)((e0 & 0xffffu) * gstride));
                        const float v1 = __ldg(gb + (size_t)((e1 & 0xffffu) * gstride));
                        const float v2 = __ldg(gb + (size_t)((e2 & 0xffffu) * gstride));
                        const float v3 = __ldg(gb + (size_t)((e3 & 0xffffu) * gstride));
                        // lane (dx, t) takes the columns xa + dx, xa + dx + XL, ... of the cell
#define D2T_P3_UPD(E, V)                                                                                  \
    {                                                                                                     \
        float* p = Ub + (int)(((E) >> 16) & 255u) * P;                                                    \
        for (int c = (int)((E) >> 24) - (int)(((E) >> 16) & 255u) - dx; c > 0; c -= XL, p += XL * P) *p += (V); \
    }
                        D2T_P3_UPD(e0, v0)
                        D2T_P3_UPD(e1, v1)
                        D2T_P3_UPD(e2, v2)
                        D2T_P3_UPD(e3, v3)
#undef D2T_P3_UPD
                    }
                    cnt = 0;
                    __syncwarp();
                }
            }
            // bin b + NW
            j += NW;
            while (j >= k) {
                j -= k;
                ++i;
            }
        }
    }
    __syncthreads();

    // ---- B: XB lanes per live channel, lane = column; the users' shared-memory offsets come from the parameter table
    // (warp-uniform reads of the constant bank).  Channels without users were zero-filled by the prep kernel
    {
        constexpr int CPW = 32 / XB;   // channels per step
        const int x = lane & (XB - 1), sub = lane / XB;
        const bool xin = x0 + x < W;
        const uint32_t ux = (uint32_t)__cvta_generic_to_shared(U + x * P);   // shared-window address of this lane's column
        float* dst = gin + ((size_t)n * nCh * H + y) * W + x0 + x;
        const unsigned HW = (unsigned)(H * W);   // nCh * H * W < 2^31 (psb3_supported)
        const int nLive = tab.nLive;
        for (int l0 = warp * CPW; l0 < nLive; l0 += NW * CPW) {
            const int l = l0 + sub;
            const bool ok = l < nLive;
            const uint32_t rec = tab.rec[ok ? l : l0];
            const uint32_t* up = tab.user + (rec & 2047u);
            int nu = ok ? (int)((rec >> 11) & 63u) : 0;
            float acc = 0.f;
#pragma unroll 1
            for (; nu > 0; --nu, ++up) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ux + *up) : "memory");
                acc += v;
            }
            if (ok && xin) dst[(rec >> 17) * HW] = acc;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
constexpr size_t kP3SmemCap = (size_t)227 * 1024;   // sm_100a: dynamic shared memory per CTA (a constant: workspace queries need no device)

bool psb3_supported(int N, int R, int nT, int H, int W, int k) {
    if (N <= 0 || R <= 0 || nT <= 0 || k <= 0 || H <= 0 || W <= 0) return false;
    if (nT > 32 || k > kP3MaxK || R > 65535 || H > 65535 || W > 65535 || N > 65535) return false;
    if ((long long)H * ceil_div(W, 16) > 0x7fffffffLL || (long long)N * R > 0x7fffffffLL) return false;
    if ((long long)R * k * k * p3_tp(nT) > 0x7fffffffLL || (long long)nT * k * k * H * W > 0x7fffffffLL) return false;
    return p3_smem(R, nT, k, 16) <= kP3SmemCap;
}

size_t psb3_ws_bytes(int N, int R, int nT, int H, int W, int k) {
    (void)H;
    (void)W;
    return p3_layout(N, R, nT, k).total;
}

template <int TP, int XB>
static int p3_main_launch(const float* gT, const uint32_t* erow, const uint32_t* ecol, float* gin, int N, int R, int nT, int H,
                          int W, int k, int canonical, cudaStream_t st) {
    const size_t smem = p3_smem(R, nT, k, XB);
    D2T_SMEM_OPTIN((psb3_bwd_kernel<TP, XB>), smem);
    P3Tab tab;
    p3_build_tab(tab, nT, k, canonical, XB);
    const int NXB = ceil_div(W, XB);
    psb3_bwd_kernel<TP, XB><<<dim3(H * NXB, N), XB * 32, smem, st>>>(gT, erow, ecol, gin, R, nT, H, W, k, NXB, tab);
    D2T_CUDA_TRY(cudaGetLastError());
    return D2T_OK;
}

template <int TP>
static int p3_main_dispatch(const float* gT, const uint32_t* erow, const uint32_t* ecol, float* gin, int N, int R, int nT, int H,
                            int W, int k, int canonical, cudaStream_t st) {
    // 32-column blocks (one CTA per SM, half the per-CTA fixed work per pixel) once the grid fills the chip with them
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const bool wide = W > 16 && p3_smem(R, nT, k, 32) <= kP3SmemCap && (long long)N * H * ceil_div(W, 32) >= 2LL * di.sm_count;
    if (wide) return p3_main_launch<TP, 32>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st);
    return p3_main_launch<TP, 16>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st);
}

// vote != 0: grad_out is (N, R, nT), the gradient of the vote (see psb3_prep_kernel)
int psb3_bwd_launch(const float* go, const float* rois, float* gin, int N, int R, int nT, int H, int W, int k, int flags, int vote,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
    const P3Layout L = p3_layout(N, R, nT, k);
    if (!ws || ws_bytes < L.total) {
        set_error("psroipool_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, L.total);
        return D2T_ERR_WORKSPACE;
    }
    char* base = static_cast<char*>(ws);
    uint32_t* erow = reinterpret_cast<uint32_t*>(base + L.erowOff);
    uint32_t* ecol = reinterpret_cast<uint32_t*>(base + L.ecolOff);
    float* gT = reinterpret_cast<float*>(base + L.gtOff);
    const int canonical = (flags & D2T_PS_CANONICAL_MAP) ? 1 : 0;
    const int TP = p3_tp(nT), kk = k * k;
    int tpShift = 0;
    while ((1 << tpShift) < TP) ++tpShift;
    const size_t prepSmem = ((size_t)nT * (kk | 1) + 2 * k) * 4 + 16;
    psb3_prep_kernel<<<(unsigned)((long long)N * R), kP3PrepThreads, prepSmem, st>>>(go, rois, erow, ecol, gT, gin, R, nT, TP, tpShift,
                                                                                    H, W, k, canonical, vote);
    D2T_CUDA_TRY(cudaGetLastError());
    int rc;
    switch (TP) {
        case 1: rc = p3_main_dispatch<1>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st); break;
        case 2: rc = p3_main_dispatch<2>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st); break;
        case 4: rc = p3_main_dispatch<4>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st); break;
        case 8: rc = p3_main_dispatch<8>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st); break;
        case 16: rc = p3_main_dispatch<16>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st); break;
        default: rc = p3_main_dispatch<32>(gT, erow, ecol, gin, N, R, nT, H, W, k, canonical, st); break;
    }
    if (rc) return rc;
    note_launch(2);
    return D2T_OK;
}

}  // namespace d2t
