// psf_fwd.cu -- EXPERIMENT (round 2, not part of the library): batched PSROIPool forward on per-plane prefix sums.
//
// Three versions were measured against the bit-identical per-pixel kernel psb_fwd_kernel (pool_ps.cu), class head,
// 16 frames x 300 RoIs (profiles/r2_psf_experiment.txt):
//   psb_fwd_kernel (shipped)                                     140 us   105 M warp instructions, issue 83 %
//   double-precision summed-area table, column scan from global  263 us   latency: 5 serial load batches per plane
//   same, cooperative prefetched plane loads, 256 threads        132 us   two serial scans + 5 barriers per plane
//   float row prefix sums (below)                                128 us   64 M warp instructions, barrier-bound
// The per-OUTPUT work drops 3x (160 vs 460 instructions per 32 outputs), but a plane-owner kernel pays a fixed cost per
// plane -- load, scan (a 63-step serial chain on 38 threads while the other warps wait at the barrier), users -- for only
// ~750 outputs, and 9728 live planes make that cost the kernel.  Box head: 46 us against 40 us.  Not adopted.
// To build it again: paste the kernel and its launch branch back into pool_ps.cu (git show ae8fae4 has the wiring).
// ----------------------------------------------------------------------------------------------------
// forward, second generation: row prefix sums.  grid (ceil(nCh / kPsfPlanes), N): CTA = (frame, 8 consecutive channels)
// ----------------------------------------------------------------------------------------------------
// psb_fwd_kernel above is instruction-bound (ncu: issue 83 %, 70 % of the instructions in the cell loop): a lane sums its
// cell pixel by pixel, and the 32 RoIs of a warp have unrelated cell sizes, so every warp walks the largest cell.  Here the
// CTA turns its plane into exclusive ROW prefix sums  P[y][x] = sum_{x' < x} fm[y][x']  and a cell is
// sum_{y in [i0, i1)} P[y][j1] - P[y][j0]: two loads per cell ROW (cells are ~3 rows x ~4 columns at the R-FCN sizes)
// instead of one per pixel, and the trip count only follows the cell HEIGHT.  Differs from the reference's left-to-right
// pixel sum (ps_roipool_cuda.cu:60-69) by float rounding only -- a row prefix is at most W terms long, so the cancellation
// in P[j1] - P[j0] costs ~W * 2^-24 of the row's magnitude (tested at rtol 1e-4 + atol 1e-5 max|ref|, and at 3e-6 max|ref|
// against a float64 evaluation).  D2T_PS_EXACT_ORDER keeps the bit-identical kernel.
// (A full summed-area table -- four loads per cell, no loop at all -- needs DOUBLE precision to be safe against maps with a
// DC offset; it was measured first: 263 us, 132 us after tuning, against 140 us for the kernel above: two serial scans
// per plane instead of one.  profiles/r2_psf_experiment.txt)
//   planes   the CTA's live channels (host-computed bitmask, SURVEY.md F6: 608 of 1519) are walked in turn; the loads of
//            plane p + 1 are in flight (registers) while plane p is scanned and looked up.
//   scan     thread y runs along row y (W steps of LDS / FADD / STS; the reads are independent of the running sum).
//   outputs  per user (target, bin) of the channel, thread <- RoI: packed bin edges from a shared copy of the frame's
//            table (the edge kernel's output), the row loop, one float division, one store.
// Deviation: a non-finite value reaches every cell to its right in the same pixel row (Inf - Inf), not only the cells that
// contain it; the exact kernel keeps the reference's behaviour.
constexpr int kPsfThreads = 256;
constexpr int kPsfPlanes = 8;
constexpr int kPsfMaxLoads = 12;   // plane elements a thread holds in flight (H * W <= 12 * 256)
constexpr int kPsfMaskWords = 128; // live-channel bitmask passed by value: n_targets * k * k <= 4096
struct PsfLive {
    uint32_t w[kPsfMaskWords];
};
__host__ __device__ constexpr int psf_pitch(int W) { return (W + 1) | 1; }   // floats per prefix row; odd => row-walkers hit distinct banks
__global__ void __launch_bounds__(kPsfThreads)
psf_fwd_kernel(const float* __restrict__ fm, const uint32_t* __restrict__ edges, float* __restrict__ out, int R, int nT, int H,
               int W, int k, int canonical, const __grid_constant__ PsfLive live) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int kk = k * k, nCh = nT * kk, HW = H * W;
    const int pitch = psf_pitch(W);
    float* P = reinterpret_cast<float*>(smem_raw);                             // [H][pitch] exclusive row prefixes
    float* F = P + (size_t)H * pitch;                                          // [H * W] the raw plane
    uint32_t* ed = reinterpret_cast<uint32_t*>(F + HW);                        // [R * k] packed edges of the frame
    uint32_t* us = ed + (size_t)R * k;                                         // [kk]
    int* cnt = reinterpret_cast<int*>(us + kk);                                // [8]
    const int tid = threadIdx.x, n = blockIdx.y;
    const int ch0 = blockIdx.x * kPsfPlanes;
    uint32_t mask = (live.w[ch0 >> 5] >> (ch0 & 31)) & ((1u << kPsfPlanes) - 1u);   // kPsfPlanes divides 32: no straddling
    if (mask == 0) return;
    const float* fmN = fm + (size_t)n * nCh * HW;
    float v[kPsfMaxLoads];
    auto fetch = [&](int ch) {
        const float* src = fmN + (size_t)ch * HW + tid;
#pragma unroll
        for (int q = 0; q < kPsfMaxLoads; ++q) v[q] = (tid + q * kPsfThreads < HW) ? __ldg(src + q * kPsfThreads) : 0.f;
    };
    fetch(ch0 + __ffs(mask) - 1);
    {
        const uint32_t* eg = edges + (size_t)n * R * k;
        for (int idx = tid; idx < R * k; idx += kPsfThreads) ed[idx] = __ldg(eg + idx);
    }
    float* o = out + (size_t)n * R * nCh;
    while (mask) {
        const int ch = ch0 + __ffs(mask) - 1;
        mask &= mask - 1;
        __syncthreads();   // the previous plane's lookups (P, us) are finished; first time: ed is written
#pragma unroll
        for (int q = 0; q < kPsfMaxLoads; ++q)
            if (tid + q * kPsfThreads < HW) F[tid + q * kPsfThreads] = v[q];
        if (mask) fetch(ch0 + __ffs(mask) - 1);   // lands while this plane is scanned and looked up
        const int nU = psb_users(ch, nT, kk, canonical != 0, us, cnt);   // contains the barrier that publishes F
        if (tid < H) {   // exclusive row prefixes (H <= 255 < kPsfThreads)
            const float* src = F + tid * W;
            float* row = P + tid * pitch;
            float acc = 0.f;
#pragma unroll 8
            for (int x = 0; x < W; ++x) {
                row[x] = acc;
                acc += src[x];
            }
            row[W] = acc;
        }
        __syncthreads();
        for (int u = 0; u < nU; ++u) {
            const uint32_t pk = us[u];
            const int t = pk >> 16, b = pk & 0xffff;
            const int i = b / k, j = b - i * k;   // uniform
            float* ou = o + (t == 0xFFFF ? 0 : t * kk + b);
            for (int r = tid; r < R; r += kPsfThreads) {
                const uint32_t ei = ed[r * k + i], ej = ed[r * k + j];
                const int i0 = ei & 255, i1 = (ei >> 8) & 255, j0 = (ej >> 16) & 255, j1 = ej >> 24;
                float acc = 0.f;
                if (j1 > j0) {   // an empty cell sums nothing (the reference's loops do not run)
                    const float* pa = P + i0 * pitch + j0;
                    const int dj = j1 - j0;
                    for (int y = i0; y < i1; ++y, pa += pitch) acc += pa[dj] - pa[0];
                }
                const int numel = (i1 - i0) * (j1 - j0);
                if (numel > 0) acc /= numel;
                float* dst = ou + (size_t)(r * nT) * kk;
                if (t == 0xFFFF) {  // channel 0 of the reference map: bin 0 of every target reads it
                    for (int tt = 0; tt < nT; ++tt) dst[tt * kk] = acc;
                } else {
                    *dst = acc;
                }
            }
        }
    }
}
// which channels does some (target, bin) read?  reference map: channel (t + 1) * b; canonical map: all of them
static void psf_live_mask(PsfLive* m, int nT, int kk, bool canonical) {
    for (int w = 0; w < kPsfMaskWords; ++w) m->w[w] = 0u;
    const int nCh = nT * kk;
    if (canonical) {
        for (int ch = 0; ch < nCh; ++ch) m->w[ch >> 5] |= 1u << (ch & 31);
        return;
    }
    for (int t = 0; t < nT; ++t)
        for (int b = 0; b < kk; ++b) {
            const int ch = (t + 1) * b;
            m->w[ch >> 5] |= 1u << (ch & 31);
        }
}
static size_t psf_smem(int R, int H, int W, int k) {
    return (size_t)H * psf_pitch(W) * 4 + (size_t)H * W * 4 + (size_t)R * k * 4 + (size_t)k * k * 4 + 64;
}

