// pool_vec.cu -- float32 ROIPool forward: [pixel][16 channel] slabs walked with 128-bit shared-memory accesses,
// row prefix sums.  sm_100a.  (The backward built on the same slab layout is pool_vec2.cu.)
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
#include "common.cuh"

namespace d2t {

constexpr int kVecFwdWarps = 16;
constexpr int kVecFwdThreads = kVecFwdWarps * 32;
constexpr int kVecSlots = 16;     // channel slots per CTA (4 quads of 4)
constexpr int kVecRChunk = 512;   // RoIs per edge-table chunk

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
bool roipool_vec_supported(int R, int C, int H, int W, int k) {
    if (k != 7 || R <= 0 || C <= 0 || H > 255 || W > 255) return false;
    DeviceInfo di;
    if (device_info(&di)) return false;
    return vec_fwd_smem(H, W, k) <= (size_t)di.max_smem_optin;
}

int roipool_vec_fwd_launch(const float* fm, const float* rois, float* out, int R, int C, int H, int W, int k,
                           cudaStream_t st) {
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    const int CB = vec_pick_cb(C, di.sm_count);
    const size_t smem = vec_fwd_smem(H, W, k);
    auto kern = roipool_vec_fwd_kernel<7>;
    D2T_SMEM_OPTIN(kern, smem);
    kern<<<ceil_div(C, CB), kVecFwdThreads, smem, st>>>(fm, rois, out, R, C, H, W, CB);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

}  // namespace d2t
