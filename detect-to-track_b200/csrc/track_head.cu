// track_head.cu -- the D&T track-regression head, ROIPool -> view -> Linear(C*k*k, n_out), as ONE fused operator
// (forward + backward) that never materialises the pooled (R, C, k, k) tensor.  sm_100a.
//
// Reference composition (correlation_tracker.py:82-85):
//     pooled = ROIPool(k)(track_feats, rois)            (R, C, k, k)   111 MB at C = 1891, R = 300, k = 7
//     t_hat  = Linear(C*k*k, n_out)(pooled.view(R, -1))  (R, n_out)
// Both steps are linear in track_feats and the pooling weights do not depend on the channel, so the contraction over
// channels can be done FIRST, on the un-pooled map:
//     Z[p, (o,ij)]   = sum_c X[c, p] * W[o, c*kk + ij]                      GEMM  P x (n_out*kk) x C      (forward 1)
//     t_hat[r, o]    = b[o] + sum_ij (1/n_rij) sum_{p in bin_rij} Z[p,(o,ij)]   position-sensitive pooling of Z (forward 2)
// and the backward is its transpose:
//     gZ[p, (o,ij)]  = sum_r g[r, o] * [p in bin_rij] / n_rij                (a 196-channel, R-RoI scatter, done as a gather)
//     gX[c, p]       = sum_(o,ij) gZ[p,(o,ij)] * W[o, c*kk + ij]             GEMM  P x C x (n_out*kk)
//     gW[o, c*kk+ij] = sum_p gZ[p,(o,ij)] * X[c, p]                          GEMM  C x (n_out*kk) x P
//     gb[o]          = sum_r g[r, o]
// 1.77 GFLOP per GEMM at the BASELINE track-head size instead of 129 MB of HBM traffic per direction plus a
// (300 x 92659) x (92659 x 4) library GEMM.  Bins, clamped RoI start and the empty-bin NaN (0/0) follow ROIPool
// (roipool_cuda.cu:26-61): an empty bin makes every output of that RoI NaN, exactly like pooled -> Linear would.
//
// The GEMMs run on tcgen05 in 3xTF32 (gemm_tf32x3.cu) from K-major FP32 operands written by the layout kernels below
// (TMA needs 16-byte-aligned row pitches, which the reference's 38x63 maps do not have); all reductions have a fixed order (split-K slabs summed ascending, RoI lists ascending): bitwise reproducible.
#include "gemm_tf32x3.cuh"

namespace d2t {

namespace {

constexpr int kThMaxN = 256;   // n_out * r_hw^2 must fit one UMMA N tile

struct ThDims {
    int NB;                 // images per call (the track features of NB frame pairs share one weight)
    int R, C, H, W, k, nO;
    int P, KK, N1;          // pixels per image, bins, n_out * bins
    int ldc, ldp, ldn;      // pitches (floats, multiples of 4) of K = C, K = P and K = N1 operands / N1-wide outputs
    int bn;                 // N tile for the N = N1 GEMMs
    int s1, s3;             // split-K factors of forward GEMM / weight-gradient GEMM
};

static int th_dims(int NB, int R, int C, int H, int W, int k, int nO, ThDims* d) {
    D2T_REQUIRE(NB > 0 && R >= 0 && C > 0 && H > 0 && W > 0 && k > 0 && nO > 0,
                "trackhead: bad shape N=%d R=%d C=%d H=%d W=%d r_hw=%d n_out=%d", NB, R, C, H, W, k, nO);
    D2T_REQUIRE(nO * k * k <= kThMaxN, "trackhead: n_out * r_hw^2 must be <= %d", kThMaxN);
    D2T_REQUIRE(H < 32768 && W < 32768 && (long long)NB * C * H * W < (1ll << 31) && (long long)NB * R < (1ll << 24),
                "trackhead: map too large");
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    d->NB = NB; d->R = R; d->C = C; d->H = H; d->W = W; d->k = k; d->nO = nO;
    d->P = H * W; d->KK = k * k; d->N1 = nO * k * k;
    d->ldc = (int)align_up(C, 4); d->ldp = (int)align_up(d->P, 4); d->ldn = (int)align_up(d->N1, 4);
    d->bn = d->N1 <= 64 ? 64 : d->N1 <= 208 ? 208 : 256;
    auto splits = [&](int tiles, int K) {
        const int kb = ceil_div(K, 32);
        int s = di.sm_count / (tiles > 0 ? tiles : 1);
        if (s > kb) s = kb;
        return s < 1 ? 1 : s;
    };
    d->s1 = splits(ceil_div(NB * d->P, 128), C);
    d->s3 = splits(ceil_div(C, 128), NB * d->ldp);
    return D2T_OK;
}

// workspace carving (all offsets multiples of 256 bytes)
struct ThFwdWs {
    float *xt, *wt, *zpart, *z;
    size_t total;
};
static ThFwdWs th_fwd_ws(const ThDims& d, void* base) {
    ThFwdWs w;
    size_t off = 0;
    auto take = [&](size_t floats) {
        float* p = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
        off = align_up(off + floats * sizeof(float), 256);
        return p;
    };
    w.xt = take((size_t)d.NB * d.P * d.ldc);
    w.wt = take((size_t)d.N1 * d.ldc);
    w.zpart = take((size_t)d.s1 * d.NB * d.P * d.ldn);
    w.z = take((size_t)d.NB * d.P * d.ldn);
    w.total = off;
    return w;
}
struct ThBwdWs {
    float *gz, *gzt, *xc, *wc, *gwpart;
    size_t total;
};
static ThBwdWs th_bwd_ws(const ThDims& d, void* base) {
    ThBwdWs w;
    size_t off = 0;
    auto take = [&](size_t floats) {
        float* p = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
        off = align_up(off + floats * sizeof(float), 256);
        return p;
    };
    w.gz = take((size_t)d.NB * d.P * d.ldn);
    w.gzt = take((size_t)d.N1 * d.NB * d.ldp);   // [n][image b][ldp]: K of the weight-gradient GEMM runs over all images
    w.xc = take((size_t)d.C * d.NB * d.ldp);     // used unless NB == 1 and the map itself is TMA-legal (H*W % 4 == 0, aligned base)
    w.wc = take((size_t)d.C * d.ldn);
    w.gwpart = take((size_t)d.s3 * d.C * d.ldn);
    w.total = off;
    return w;
}

// ---- layout kernels ---------------------------------------------------------------------------------------------
// X (C, P) -> XT (P, ldc): 32 x 32 tiles through shared memory, both sides coalesced
__global__ void __launch_bounds__(256)
th_transpose_kernel(const float* __restrict__ x, float* __restrict__ xt, int C, int P, int ldc) {
    __shared__ float tile[32][33];
    x += (size_t)blockIdx.z * C * P;          // image blockIdx.z: rows b*P .. b*P + P-1 of the tall XT
    xt += (size_t)blockIdx.z * P * ldc;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, p = p0 + tx;
        tile[r][tx] = (c < C && p < P) ? __ldg(x + (size_t)c * P + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, c = c0 + tx;
        if (p < P && c < C) xt[(size_t)p * ldc + c] = tile[tx][r];
    }
}

// X (NB, C, P) -> Xc (C, NB * ldp): channel-major over all images, every image block ldp floats wide (16-byte-aligned pitch),
// pad columns zero (they are inside the K range of the weight-gradient GEMM when NB > 1)
__global__ void __launch_bounds__(256)
th_pad_copy_kernel(const float* __restrict__ x, float* __restrict__ xc, int C, int P, int ldp, int NB) {
    const int c = blockIdx.y, b = blockIdx.z;
    const float* src = x + ((size_t)b * C + c) * P;
    float* dst = xc + (size_t)c * NB * ldp + (size_t)b * ldp;
    if ((P & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0) {
        // even plane size: every source row starts on an 8-byte boundary (the destination pitch is a multiple of 16 bytes)
        const float2* s2 = reinterpret_cast<const float2*>(src);
        float2* d2 = reinterpret_cast<float2*>(dst);
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ldp / 2; q += gridDim.x * blockDim.x)
            d2[q] = 2 * q < P ? __ldg(s2 + q) : make_float2(0.f, 0.f);
        return;
    }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < ldp; p += gridDim.x * blockDim.x) dst[p] = p < P ? __ldg(src + p) : 0.f;
}

// The n_out * kk "position-sensitive channels" are numbered n = ij * nO + o (outputs of one bin adjacent), so that the
// pooling and gZ kernels move the nO values of a bin with one vector access.
// W (nO, C*KK) -> Wt (n, ldc)   [K = c]   (by_channel == 0)
//              -> Wc (c, ldn)   [K = n]   (by_channel == 1)
__global__ void __launch_bounds__(256)
th_weight_prep_kernel(const float* __restrict__ w, float* __restrict__ out, int C, int KK, int nO, int ld, int by_channel) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see launch_overlapped
    const int total = nO * C * KK;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        int o, c, ij;
        size_t dst;
        if (by_channel) {          // e = (c, n): writes contiguous
            c = e / (nO * KK);
            const int n = e - c * (nO * KK);
            ij = n / nO; o = n - ij * nO;
            dst = (size_t)c * ld + n;
        } else {                   // e = (n, c): writes contiguous
            const int n = e / C;
            c = e - n * C;
            ij = n / nO; o = n - ij * nO;
            dst = (size_t)n * ld + c;
        }
        out[dst] = __ldg(w + (size_t)o * C * KK + (size_t)c * KK + ij);
    }
}

// Z[p][n] = sum_s Zpart[s][p][n], ascending s
__global__ void __launch_bounds__(256)
th_reduce_slabs_kernel(const float4* __restrict__ part, float4* __restrict__ z, int n4, int splits) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += gridDim.x * blockDim.x) {
        float4 s = part[e];
        for (int k = 1; k < splits; ++k) {
            const float4 v = part[(size_t)k * n4 + e];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        z[e] = s;
    }
}

// ---- forward 2: position-sensitive pooling of Z over ROIPool bins ----------------------------------------------------
// one CTA per RoI, 4 threads per bin: thread (ij, q) sums the nO adjacent values Z[., ij*nO .. ij*nO+nO-1] (one 16-byte
// load per pixel for nO = 4) over the pixels q, q + 4, ... of its bin (row-major inside the bin), the four partial sums are
// combined in a fixed order and divided by the bin size (0/0 = NaN for an empty bin, like roipool_cuda.cu:61); thread o
// then adds the kk bin means of its output in ascending bin order.  All loads of a thread are independent (L2 hits).
constexpr int kPoolThreads2 = 256;
template <int NO>
__global__ void __launch_bounds__(kPoolThreads2)
th_pool_kernel(const float* __restrict__ z, const float* __restrict__ rois, const float* __restrict__ bias, float* __restrict__ out,
               int R, int H, int W, int k, int ldn) {
    extern __shared__ float th_part[];   // [KK][NO]
    const int KK = k * k;
    const int r = blockIdx.x;
    z += (size_t)blockIdx.y * H * W * ldn;        // image blockIdx.y
    rois += (size_t)blockIdx.y * R * 4;
    out += (size_t)blockIdx.y * R * NO;
    const float* roi = rois + (size_t)r * 4;
    const float r0 = __ldg(roi), r1 = __ldg(roi + 1), r2 = __ldg(roi + 2), r3 = __ldg(roi + 3);
    for (int base = 0; base < KK * 4; base += kPoolThreads2) {   // uniform trip count: every lane reaches the shuffles
        const int item = base + threadIdx.x;
        const bool valid = item < KK * 4;
        const int ij = valid ? item >> 2 : 0, q = item & 3;
        const int i = ij / k, j = ij - i * k;
        int i0, i1, j0, j1;
        bin_edge<float, true>(r0, r2, i, k, H, i0, i1);
        bin_edge<float, true>(r1, r3, j, k, W, j0, j1);
        const int bw = j1 - j0, numel = (i1 - i0) * bw;
        const int lim = (valid && i1 > i0 && bw > 0) ? numel : 0;
        float acc[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = 0.f;
        const float* zb = z + (size_t)ij * NO;
        for (int e = q; e < lim; e += 4) {
            const int pi = i0 + e / bw, pj = j0 + e % bw;
            const float* src = zb + (size_t)(pi * W + pj) * ldn;
            if (NO == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            } else {
#pragma unroll
                for (int o = 0; o < NO; ++o) acc[o] += __ldg(src + o);
            }
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) {   // (q0 + q1) + (q2 + q3): fixed shape
            acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
            acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
        }
        if (valid && q == 0) {
#pragma unroll
            for (int o = 0; o < NO; ++o) th_part[ij * NO + o] = acc[o] / numel;
        }
    }
    __syncthreads();
    if (threadIdx.x < NO) {
        const int o = threadIdx.x;
        float s = bias ? bias[o] : 0.f;
        for (int b = 0; b < KK; ++b) s += th_part[b * NO + o];
        out[(size_t)r * NO + o] = s;
    }
}

// ---- backward 1: gZ (both layouts) -------------------------------------------------------------------------------------
// one CTA per (pixel row y, bin row i).  Phase 1: every RoI's bin-row-i edges; warp 0 compacts, in ascending RoI order,
// the RoIs whose bin row covers y.  Phase 2: column edges and scaled gradients g[r, o] / n_rij of the listed RoIs.
// Phase 3: thread (x, j) adds, in list order, the entries whose column bin j covers x -- a gather: fixed order, no atomics.
constexpr int kGzThreads = 256;
constexpr int kGzMaxO = 8;
__global__ void __launch_bounds__(kGzThreads)
th_gz_kernel(const float* __restrict__ g, const float* __restrict__ rois, float* __restrict__ gz, float* __restrict__ gzt, int R,
             int H, int W, int k, int nO, int ldn, int ldp, int ldpT) {
    extern __shared__ unsigned char th_smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see launch_overlapped
    // list[R] (int) | rowext[R] (int: I1 - I0 or 0) | cj[R*k] (short2 J0, J1) | val[R*k*nO] (float)
    int* list = reinterpret_cast<int*>(th_smem);
    int* rowext = list + R;
    short2* cj = reinterpret_cast<short2*>(rowext + R);
    float* val = reinterpret_cast<float*>(cj + (size_t)R * k);
    __shared__ int nlist;
    const int y = blockIdx.x / k, i = blockIdx.x - y * k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    g += (size_t)blockIdx.y * R * nO;              // image blockIdx.y
    rois += (size_t)blockIdx.y * R * 4;
    gz += (size_t)blockIdx.y * H * W * ldn;
    gzt += (size_t)blockIdx.y * ldp;               // this image's column block of the (N1, ldpT) transposed layout
    if (y == H - 1)                                // pad columns of the block: inside the K range of the weight-gradient GEMM
        for (int e = tid; e < k * nO * (ldp - H * W); e += kGzThreads) {
            const int n = i * k * nO + e / (ldp - H * W), p = H * W + e % (ldp - H * W);
            gzt[(size_t)n * ldpT + p] = 0.f;
        }
    for (int r = tid; r < R; r += kGzThreads) {
        const float* roi = rois + (size_t)r * 4;
        int i0, i1;
        bin_edge<float, true>(roi[0], roi[2], i, k, H, i0, i1);
        rowext[r] = (i0 <= y && y < i1) ? (i1 - i0) : 0;
    }
    __syncthreads();
    if (warp == 0) {
        int cnt = 0;
        for (int r0 = 0; r0 < R; r0 += 32) {
            const int r = r0 + lane;
            const bool in = r < R && rowext[r] > 0;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) list[cnt + __popc(bal & ((1u << lane) - 1))] = r;
            cnt += __popc(bal);
        }
        if (lane == 0) nlist = cnt;
    }
    __syncthreads();
    const int nl = nlist;
    for (int e = tid; e < nl * k; e += kGzThreads) {
        const int l = e / k, j = e - l * k;
        const int r = list[l];
        const float* roi = rois + (size_t)r * 4;
        int j0, j1;
        bin_edge<float, true>(roi[1], roi[3], j, k, W, j0, j1);
        cj[e] = make_short2((short)j0, (short)j1);
        const int numel = rowext[r] * (j1 - j0);
        for (int o = 0; o < nO; ++o) val[(size_t)e * nO + o] = numel > 0 ? __ldg(g + (size_t)r * nO + o) / numel : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < W * k; e += kGzThreads) {
        const int j = e / W, x = e - j * W;   // x fastest: the transposed-layout stores are coalesced
        float acc[kGzMaxO];
#pragma unroll
        for (int o = 0; o < kGzMaxO; ++o) acc[o] = 0.f;
        for (int l = 0; l < nl; ++l) {
            const short2 c = cj[l * k + j];
            if (x >= c.x && x < c.y) {
                const float* v = val + (size_t)(l * k + j) * nO;
#pragma unroll
                for (int o = 0; o < kGzMaxO; ++o)
                    if (o < nO) acc[o] += v[o];
            }
        }
        const int p = y * W + x;
        const int nb = (i * k + j) * nO;
        if (nO == 4)   // ldn and nb are multiples of 4: one 16-byte store
            *reinterpret_cast<float4*>(gz + (size_t)p * ldn + nb) = make_float4(acc[0], acc[1], acc[2], acc[3]);
#pragma unroll
        for (int o = 0; o < kGzMaxO; ++o) {
            if (o < nO) {
                if (nO != 4) gz[(size_t)p * ldn + nb + o] = acc[o];
                gzt[(size_t)(nb + o) * ldpT + p] = acc[o];
            }
        }
    }
}

// ---- backward 3: gW[o][c*KK + ij] = sum_s gWpart[s][c][ij*nO + o];  gb[o] = sum_r g[r][o] -----------------------------
__global__ void __launch_bounds__(256)
th_reduce_w_kernel(const float* __restrict__ part, const float* __restrict__ g, float* __restrict__ gw, float* __restrict__ gb,
                   int R, int C, int KK, int nO, int ldn, int splits) {
    const int N1 = nO * KK;
    const size_t slab = (size_t)C * ldn;
    if (gw) {
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < C * N1; e += gridDim.x * blockDim.x) {
            const int c = e / N1, n = e - c * N1;
            const int ij = n / nO, o = n - ij * nO;
            float s = 0.f;
            for (int k = 0; k < splits; ++k) s += part[(size_t)k * slab + (size_t)c * ldn + n];
            gw[(size_t)o * C * KK + (size_t)c * KK + ij] = s;
        }
    }
    // bias gradient: the last block, one warp per output; lanes take RoIs r = lane, lane + 32, ... (ascending), then a
    // fixed-shape shuffle tree: deterministic
    if (gb && blockIdx.x == gridDim.x - 1) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = warp; o < nO; o += blockDim.x >> 5) {
            float s = 0.f;
            for (int r = lane; r < R; r += 32) s += __ldg(g + (size_t)r * nO + o);
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
            if (lane == 0) gb[o] = s;
        }
    }
}

// launch `kernel` so that it may start before the previous kernel of the stream has finished (programmatic dependent launch):
// only for a kernel that reads nothing the previous one writes, after a kernel that executes griddepcontrol.launch_dependents
template <typename... KArgs, typename... Args>
static cudaError_t launch_overlapped(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int grid_for(size_t n, int block, int cap) {
    size_t g = (n + block - 1) / block;
    if (g > (size_t)cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

}  // namespace

size_t trackhead_fwd_ws_bytes(int NB, int R, int C, int H, int W, int k, int nO) {
    ThDims d;
    if (th_dims(NB, R, C, H, W, k, nO, &d)) return 0;
    return th_fwd_ws(d, nullptr).total;
}
size_t trackhead_bwd_ws_bytes(int NB, int R, int C, int H, int W, int k, int nO) {
    ThDims d;
    if (th_dims(NB, R, C, H, W, k, nO, &d)) return 0;
    return th_bwd_ws(d, nullptr).total;
}

// fm (NB, C, H, W), rois (NB, R, 4), out (NB, R, nO): the NB images share weight and bias; one set of launches, the
// forward GEMM runs over all NB * H * W positions at once
int trackhead_fwd_launch(const float* fm, const float* rois, const float* weight, const float* bias, float* out, int NB, int R, int C,
                         int H, int W, int k, int nO, void* wsp, size_t ws_bytes, cudaStream_t st) {
    ThDims d;
    int rc = th_dims(NB, R, C, H, W, k, nO, &d);
    if (rc) return rc;
    if (R == 0) return D2T_OK;
    D2T_REQUIRE(nO <= 8, "trackhead_fwd: n_out must be <= 8");
    D2T_REQUIRE(fm && rois && weight && out, "trackhead_fwd: null pointer");
    const ThFwdWs w = th_fwd_ws(d, wsp);
    if (wsp == nullptr || ws_bytes < w.total) {
        set_error("trackhead_fwd: workspace too small (%zu < %zu)", ws_bytes, w.total);
        return D2T_ERR_WORKSPACE;
    }
    DeviceInfo di;
    if ((rc = device_info(&di))) return rc;
    const int cap = di.sm_count * 8;
    th_weight_prep_kernel<<<grid_for((size_t)d.N1 * C, 256, cap), 256, 0, st>>>(weight, w.wt, C, d.KK, nO, d.ldc, 0);
    D2T_CUDA_TRY(cudaGetLastError());
    // independent of the weight re-layout before it
    D2T_CUDA_TRY(launch_overlapped(th_transpose_kernel, dim3(ceil_div(d.P, 32), ceil_div(C, 32), NB), dim3(256), 0, st, fm, w.xt, C, d.P,
                                   d.ldc));
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch(2);
    const int PT = NB * d.P;
    GemmOperand A{w.xt, PT, d.ldc}, B{w.wt, d.N1, d.ldc};
    // Wave quantisation: 8 pairs of 38x63 positions are 150 tiles of 128 rows on 148 SMs -- as one launch the last two tiles
    // are a second wave as long as the first (122 us instead of 61).  When a few tiles spill over a whole number of waves,
    // the first waves * SMs tiles run un-split and the spill runs as a short split-K problem whose slabs (in w.z) are summed
    // into the tail rows of the result.
    const int tiles = ceil_div(PT, 128), kb = ceil_div(C, 32);
    const int spill = tiles % di.sm_count;
    if (d.s1 == 1 && tiles > di.sm_count && spill > 0 && spill * 8 <= di.sm_count && kb >= 8) {
        const int M1 = (tiles - spill) * 128, Mt = PT - M1;
        int stail = di.sm_count / (4 * spill);
        if (stail > kb / 2) stail = kb / 2;
        if (stail < 1) stail = 1;
        if ((long long)stail * Mt <= (long long)PT) {
            GemmOperand A1{w.xt, M1, d.ldc}, At{w.xt + (size_t)M1 * d.ldc, Mt, d.ldc};
            if ((rc = gemm_tf32x3(A1, B, w.zpart, M1, d.N1, C, d.ldn, GEMM_EPI_ROW, 1, M1, d.bn, st))) return rc;
            float* tailOut = w.zpart + (size_t)M1 * d.ldn;
            if (stail == 1) {
                if ((rc = gemm_tf32x3(At, B, tailOut, Mt, d.N1, C, d.ldn, GEMM_EPI_ROW, 1, Mt, d.bn, st))) return rc;
            } else {
                if ((rc = gemm_tf32x3(At, B, w.z, Mt, d.N1, C, d.ldn, GEMM_EPI_ROW, stail, Mt, d.bn, st))) return rc;
                const int n4t = Mt * d.ldn / 4;
                th_reduce_slabs_kernel<<<grid_for(n4t, 256, cap), 256, 0, st>>>(reinterpret_cast<const float4*>(w.z),
                                                                                reinterpret_cast<float4*>(tailOut), n4t, stail);
                D2T_CUDA_TRY(cudaGetLastError());
                note_launch();
            }
        } else if ((rc = gemm_tf32x3(A, B, w.zpart, PT, d.N1, C, d.ldn, GEMM_EPI_ROW, d.s1, PT, d.bn, st))) {
            return rc;
        }
    } else if ((rc = gemm_tf32x3(A, B, w.zpart, PT, d.N1, C, d.ldn, GEMM_EPI_ROW, d.s1, PT, d.bn, st))) {
        return rc;
    }
    const float* z = w.zpart;
    if (d.s1 > 1) {
        const int n4 = PT * d.ldn / 4;
        th_reduce_slabs_kernel<<<grid_for(n4, 256, cap), 256, 0, st>>>(reinterpret_cast<const float4*>(w.zpart),
                                                                       reinterpret_cast<float4*>(w.z), n4, d.s1);
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
        z = w.z;
    }
    {
        const size_t psm = (size_t)d.KK * nO * sizeof(float);
        switch (nO) {
            case 1: th_pool_kernel<1><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            case 2: th_pool_kernel<2><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            case 3: th_pool_kernel<3><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            case 4: th_pool_kernel<4><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            case 5: th_pool_kernel<5><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            case 6: th_pool_kernel<6><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            case 7: th_pool_kernel<7><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            case 8: th_pool_kernel<8><<<dim3(R, NB), kPoolThreads2, psm, st>>>(z, rois, bias, out, R, H, W, k, d.ldn); break;
            default: set_error("trackhead_fwd: n_out must be <= 8"); return D2T_ERR_BAD_ARG;
        }
    }
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

// go (NB, R, nO), gfm (NB, C, H, W); gw / gb are the sums over the NB images (one GEMM with K over all images / one pass)
int trackhead_bwd_launch(const float* go, const float* fm, const float* rois, const float* weight, float* gfm, float* gw, float* gb,
                         int NB, int R, int C, int H, int W, int k, int nO, void* wsp, size_t ws_bytes, cudaStream_t st) {
    ThDims d;
    int rc = th_dims(NB, R, C, H, W, k, nO, &d);
    if (rc) return rc;
    D2T_REQUIRE(nO <= kGzMaxO, "trackhead_bwd: n_out must be <= %d", kGzMaxO);
    DeviceInfo di;
    if ((rc = device_info(&di))) return rc;
    const int cap = di.sm_count * 8;
    if (R == 0) {   // no RoIs: every gradient is zero
        if (gfm) D2T_CUDA_TRY(cudaMemsetAsync(gfm, 0, (size_t)NB * C * d.P * sizeof(float), st));
        if (gw) D2T_CUDA_TRY(cudaMemsetAsync(gw, 0, (size_t)nO * C * d.KK * sizeof(float), st));
        if (gb) D2T_CUDA_TRY(cudaMemsetAsync(gb, 0, (size_t)nO * sizeof(float), st));
        return D2T_OK;
    }
    D2T_REQUIRE(go && fm && rois && weight, "trackhead_bwd: null pointer");
    const ThBwdWs w = th_bwd_ws(d, wsp);
    if (wsp == nullptr || ws_bytes < w.total) {
        set_error("trackhead_bwd: workspace too small (%zu < %zu)", ws_bytes, w.total);
        return D2T_ERR_WORKSPACE;
    }
    const size_t gzSmem = (size_t)R * 8 + (size_t)R * k * 4 + (size_t)R * k * nO * 4;
    D2T_REQUIRE(gzSmem <= (size_t)di.max_smem_optin - 1024, "trackhead_bwd: too many RoIs for one call (%d)", R);
    D2T_SMEM_OPTIN(th_gz_kernel, gzSmem);
    const int ldpT = NB * d.ldp;
    th_gz_kernel<<<dim3(H * k, NB), kGzThreads, gzSmem, st>>>(go, rois, w.gz, w.gzt, R, H, W, k, nO, d.ldn, d.ldp, ldpT);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    if (gfm) {
        // independent of the gZ gather before it
        D2T_CUDA_TRY(launch_overlapped(th_weight_prep_kernel, dim3(grid_for((size_t)d.N1 * C, 256, cap)), dim3(256), 0, st, weight, w.wc,
                                       C, d.KK, nO, d.ldn, 1));
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
        GemmOperand A{w.gz, NB * d.P, d.ldn}, B{w.wc, C, d.ldn};
        // N tile 208 (10 x 19 = 190 CTAs, three raw stages) beats 256 (8 x 19 = 152 CTAs: also two waves, two stages)
        if ((rc = gemm_tf32x3(A, B, gfm, NB * d.P, C, d.N1, d.P, GEMM_EPI_COL, 1, 0, 208, st, NB > 1 ? d.P : 0, (long long)C * d.P)))
            return rc;
    }
    if (gw) {
        const float* xc = fm;
        if (NB > 1 || d.P != d.ldp || (reinterpret_cast<uintptr_t>(fm) & 15) != 0) {
            // independent of the grad_X GEMM before it (reads fm, writes w.xc): an HBM-bound copy under a tensor-bound GEMM
            if (gfm) {
                D2T_CUDA_TRY(launch_overlapped(th_pad_copy_kernel, dim3(ceil_div(d.ldp, 2048), C, NB), dim3(256), 0, st, fm, w.xc, C,
                                               d.P, d.ldp, NB));
            } else {
                th_pad_copy_kernel<<<dim3(ceil_div(d.ldp, 2048), C, NB), 256, 0, st>>>(fm, w.xc, C, d.P, d.ldp, NB);
            }
            D2T_CUDA_TRY(cudaGetLastError());
            note_launch();
            xc = w.xc;
        }
        // K runs over the pixels of all images (pad columns are zero in both operands; for one image K = P and TMA zero-fills)
        const int Kw = NB > 1 ? ldpT : d.P;
        GemmOperand A{xc, C, ldpT}, B{w.gzt, d.N1, ldpT};
        if ((rc = gemm_tf32x3(A, B, w.gwpart, C, d.N1, Kw, d.ldn, GEMM_EPI_ROW, d.s3, C, d.bn, st))) return rc;
    }
    if (gw || gb) {
        th_reduce_w_kernel<<<grid_for((size_t)C * d.N1, 256, cap), 256, 0, st>>>(w.gwpart, go, gw, gb, NB * R, C, d.KK, nO, d.ldn, d.s3);
        D2T_CUDA_TRY(cudaGetLastError());
        note_launch();
    }
    return D2T_OK;
}

}  // namespace d2t
