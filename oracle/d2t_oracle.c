/*
 * d2t_oracle.c -- CPU restatement of detect-to-track's three custom ops.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (detect-to-track_b200/) never links, loads or
 * calls it, and has no CPU fallback.
 *
 * Pinning: the reference ships no golden values for these ops except one
 * known-answer test (fully out-of-bounds RoI => zeros,
 * tests/test_ps_roipool.py:33-44) and float64 gradcheck self-consistency
 * (tests/test_pointwise_correlation.py:20, tests/test_roipool.py:25,
 * tests/test_ps_roipool.py:28).  Both are re-run against this file in
 * tests/test_oracle.py.  Forward VALUES are pinned by tests/golden/*.npz,
 * which were produced on a B200 by the reference's own unmodified kernels
 * (oracle/_ref/libd2t_ref_cuda.so, built by oracle/Makefile from
 * /root/reference; generator: tools/make_golden.py).
 *
 * Every function follows the reference line by line; file:line citations are
 * relative to /root/reference/detect_to_track/models/.  Floating-point
 * contraction is OFF for this file (-ffp-contract=off); the places where the
 * reference's nvcc build fuses a multiply-add (checked in its SASS) use an
 * explicit fma()/fmaf(), so the integer bin edges are reproduced bit for bit.
 *
 * Loop ORDER per output element is the reference's (channels ascending for
 * the correlation; rows then columns for the pools); loops over independent
 * outputs are re-nested / OpenMP-parallel, which cannot change any value.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int d2t_oracle_version(void) { return 1; }

int d2t_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void d2t_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* common/cuda_common.cuh:9-13 : clamp to [0,1] = max(0, min(1, x)) */
static inline float clamp_f32(float x) { return fmaxf(0.0f, fminf(1.0f, x)); }
static inline double clamp_f64(double x) { return fmax(0.0, fmin(1.0, x)); }

/* ------------------------------------------------------------------------- */
/* Bin edges                                                                  */
/* ------------------------------------------------------------------------- */

/*
 * roipool_cuda.cu:38-50.  `start` = clamp(rI - rH/2)  (ROIPool: the RoI start
 * is clamped, F7) ; bin centre bI = start + (scalar(i) + 0.5) * bH where the
 * literal 0.5 is a double, so for float the sum is formed in double and
 * narrowed on initialisation (F8).  The nvcc build contracts that
 * multiply-add into one DFMA (seen in the SASS of both instantiations); for
 * float the product is exact in double so the contraction is invisible, for
 * double it is not, hence fma().
 */
static inline void roipool_edge_f32(float rI, float rH, int i, int k, int n, int* e0, int* e1) {
    const float bH = rH / (float)k;
    const float start = clamp_f32(rI - rH / 2);
    const float bI = (float)fma((double)(float)i + 0.5, (double)bH, (double)start);
    *e0 = (int)floorf(clamp_f32(bI - bH / 2) * (float)n);
    *e1 = (int)ceilf(clamp_f32(bI + bH / 2) * (float)n);
}
static inline void roipool_edge_f64(double rI, double rH, int i, int k, int n, int* e0, int* e1) {
    const double bH = rH / (double)k;
    const double start = clamp_f64(rI - rH / 2);
    const double bI = fma((double)i + 0.5, bH, start);
    *e0 = (int)floor(clamp_f64(bI - bH / 2) * (double)n);
    *e1 = (int)ceil(clamp_f64(bI + bH / 2) * (double)n);
}

/*
 * ps_roipool_cuda.cu:42-54.  Same shape but the RoI start is NOT clamped:
 * cI = rI - rH/2 + (scalar(i) + 0.5) * cH, i.e. (rI - rH/2) is formed in
 * scalar_t, then added to the double product.
 */
static inline void psroipool_edge_f32(float rI, float rH, int i, int k, int n, int* e0, int* e1) {
    const float cH = rH / (float)k;
    const float start = rI - rH / 2;
    const float cI = (float)fma((double)(float)i + 0.5, (double)cH, (double)start);
    *e0 = (int)floorf(clamp_f32(cI - cH / 2) * (float)n);
    *e1 = (int)ceilf(clamp_f32(cI + cH / 2) * (float)n);
}
static inline void psroipool_edge_f64(double rI, double rH, int i, int k, int n, int* e0, int* e1) {
    const double cH = rH / (double)k;
    const double start = rI - rH / 2;
    const double cI = fma((double)i + 0.5, cH, start);
    *e0 = (int)floor(clamp_f64(cI - cH / 2) * (double)n);
    *e1 = (int)ceil(clamp_f64(cI + cH / 2) * (double)n);
}

/*
 * Integer bin edges for every (roi, bin index): edges[r][b] = {I0, I1, J0, J1}
 * where (I0,I1) is row-bin b and (J0,J1) is column-bin b.  `clamp_start` != 0
 * selects the ROIPool rule, 0 the PSROIPool rule.  This is the quantity the
 * parity tests require bit-exact.
 */
#define DEFINE_BINS(SUFFIX, T)                                                                       \
    void d2t_oracle_bins_##SUFFIX(const T* rois, int32_t* edges, int R, int H, int W, int k,        \
                                  int clamp_start) {                                                 \
        for (int r = 0; r < R; ++r) {                                                                \
            const T rI = rois[r * 4 + 0], rJ = rois[r * 4 + 1], rH = rois[r * 4 + 2],                \
                    rW = rois[r * 4 + 3];                                                            \
            for (int b = 0; b < k; ++b) {                                                            \
                int i0, i1, j0, j1;                                                                  \
                if (clamp_start) {                                                                   \
                    roipool_edge_##SUFFIX(rI, rH, b, k, H, &i0, &i1);                                \
                    roipool_edge_##SUFFIX(rJ, rW, b, k, W, &j0, &j1);                                \
                } else {                                                                             \
                    psroipool_edge_##SUFFIX(rI, rH, b, k, H, &i0, &i1);                              \
                    psroipool_edge_##SUFFIX(rJ, rW, b, k, W, &j0, &j1);                              \
                }                                                                                    \
                int32_t* e = edges + ((size_t)r * k + b) * 4;                                        \
                e[0] = i0; e[1] = i1; e[2] = j0; e[3] = j1;                                          \
            }                                                                                        \
        }                                                                                            \
    }
DEFINE_BINS(f32, float)
DEFINE_BINS(f64, double)

/* ------------------------------------------------------------------------- */
/* PointwiseCorrelation                                                       */
/* ------------------------------------------------------------------------- */

#define FMA_f32(a, b, c) fmaf((a), (b), (c))
#define FMA_f64(a, b, c) fma((a), (b), (c))

/*
 * pointwise_correlation_cuda.cu:63-111.
 *   out[b,i,j, di-i+d, dj-j+d] += FM0[b,c,i,j] * FM1[b,c,di,dj]
 *   for di in range(max(0,i-d), min(i+d,H), stride)   (exclusive upper bound, F4;
 *   for dj in range(max(0,j-d), min(j+d,W), stride)    phase tied to the clamped start, F5)
 *   c ascending, accumulated with one fused multiply-add per channel (the
 *   reference's `*outDisp += a*b` compiles to FFMA/DFMA).
 * Output is (B,H,W,2d+1,2d+1), fully written (dead entries = 0, :192).
 */
#define DEFINE_CORR_FWD(SUFFIX, T)                                                                   \
    void d2t_oracle_corr_fwd_##SUFFIX(const T* fm0, const T* fm1, T* out, int B, int C, int H,      \
                                      int W, int d, int stride) {                                    \
        const int k = 2 * d + 1;                                                                     \
        const size_t plane = (size_t)H * W;                                                          \
        memset(out, 0, sizeof(T) * (size_t)B * plane * k * k);                                       \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
        for (int b = 0; b < B; ++b) {                                                                \
            for (int i = 0; i < H; ++i) {                                                            \
                T* orow = out + ((size_t)b * plane + (size_t)i * W) * k * k;                         \
                const int di0 = imax(0, i - d), di1 = imin(i + d, H);                                \
                for (int c = 0; c < C; ++c) {                                                        \
                    const T* q = fm0 + ((size_t)b * C + c) * plane + (size_t)i * W;                  \
                    const T* key = fm1 + ((size_t)b * C + c) * plane;                                \
                    for (int j = 0; j < W; ++j) {                                                    \
                        const T qv = q[j];                                                           \
                        const int dj0 = imax(0, j - d), dj1 = imin(j + d, W);                        \
                        for (int di = di0; di < di1; di += stride) {                                 \
                            T* o = orow + ((size_t)j * k + (di - i + d)) * k + (d - j);              \
                            const T* krow = key + (size_t)di * W;                                    \
                            for (int dj = dj0; dj < dj1; dj += stride)                               \
                                o[dj] = FMA_##SUFFIX(qv, krow[dj], o[dj]);                           \
                        }                                                                            \
                    }                                                                                \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
    }
DEFINE_CORR_FWD(f32, float)
DEFINE_CORR_FWD(f64, double)

/*
 * pointwise_correlation_cuda.cu:121-174.
 *   gFM0[b,c,i,j]   += gO[b,i,j,ci,cj] * FM1[b,c,di,dj]   (thread-owned, (di,dj) ascending, :168)
 *   gFM1[b,c,di,dj] += gO[b,i,j,ci,cj] * FM0[b,c,i,j]     (atomicAdd, :169 -- the reference's
 *                      order over (i,j) is non-deterministic; the oracle uses (i,j) ascending)
 */
#define DEFINE_CORR_BWD(SUFFIX, T)                                                                   \
    void d2t_oracle_corr_bwd_##SUFFIX(const T* go, const T* fm0, const T* fm1, T* g0, T* g1,        \
                                      int B, int C, int H, int W, int d, int stride) {               \
        const int k = 2 * d + 1;                                                                     \
        const size_t plane = (size_t)H * W;                                                          \
        memset(g0, 0, sizeof(T) * (size_t)B * C * plane);                                            \
        memset(g1, 0, sizeof(T) * (size_t)B * C * plane);                                            \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
        for (int b = 0; b < B; ++b) {                                                                \
            for (int c = 0; c < C; ++c) {                                                            \
                const size_t off = ((size_t)b * C + c) * plane;                                      \
                const T* q = fm0 + off;                                                              \
                const T* key = fm1 + off;                                                            \
                T* gq = g0 + off;                                                                    \
                T* gk = g1 + off;                                                                    \
                for (int i = 0; i < H; ++i) {                                                        \
                    const int di0 = imax(0, i - d), di1 = imin(i + d, H);                            \
                    for (int j = 0; j < W; ++j) {                                                    \
                        const int dj0 = imax(0, j - d), dj1 = imin(j + d, W);                        \
                        const T qv = q[(size_t)i * W + j];                                           \
                        T acc = gq[(size_t)i * W + j];                                               \
                        const T* gbase = go + (((size_t)b * H + i) * W + j) * k * k;                 \
                        for (int di = di0; di < di1; di += stride) {                                 \
                            const T* grow = gbase + (size_t)(di - i + d) * k + (d - j);              \
                            const T* krow = key + (size_t)di * W;                                    \
                            T* gkrow = gk + (size_t)di * W;                                          \
                            for (int dj = dj0; dj < dj1; dj += stride) {                             \
                                const T g = grow[dj];                                                \
                                acc = FMA_##SUFFIX(g, krow[dj], acc);                                \
                                gkrow[dj] = FMA_##SUFFIX(g, qv, gkrow[dj]);                          \
                            }                                                                        \
                        }                                                                            \
                        gq[(size_t)i * W + j] = acc;                                                 \
                    }                                                                                \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
    }
DEFINE_CORR_BWD(f32, float)
DEFINE_CORR_BWD(f64, double)

/* ------------------------------------------------------------------------- */
/* ROIPool (average pooling, F2)                                              */
/* ------------------------------------------------------------------------- */

/*
 * roipool_cuda.cu:6-63.  out[r,c,i,j] = (sum over rows then columns of the
 * bin) / binNumel with binNumel = (BI1-BI0)*(BJ1-BJ0) converted to scalar_t;
 * no empty-bin guard, so an empty bin yields 0/0 = NaN (F7, :61).
 */
#define DEFINE_ROIPOOL_FWD(SUFFIX, T)                                                                \
    void d2t_oracle_roipool_fwd_##SUFFIX(const T* fm, const T* rois, T* out, int R, int C, int H,   \
                                         int W, int k) {                                             \
        _Pragma("omp parallel for schedule(dynamic, 1)")                                             \
        for (int r = 0; r < R; ++r) {                                                                \
            const T rI = rois[r * 4 + 0], rJ = rois[r * 4 + 1], rH = rois[r * 4 + 2],                \
                    rW = rois[r * 4 + 3];                                                            \
            for (int c = 0; c < C; ++c) {                                                            \
                const T* ch = fm + (size_t)c * H * W;                                                \
                for (int i = 0; i < k; ++i) {                                                        \
                    int i0, i1;                                                                      \
                    roipool_edge_##SUFFIX(rI, rH, i, k, H, &i0, &i1);                                \
                    for (int j = 0; j < k; ++j) {                                                    \
                        int j0, j1;                                                                  \
                        roipool_edge_##SUFFIX(rJ, rW, j, k, W, &j0, &j1);                            \
                        const int numel = (i1 - i0) * (j1 - j0);                                     \
                        T acc = 0;                                                                   \
                        for (int pi = i0; pi < i1; ++pi)                                             \
                            for (int pj = j0; pj < j1; ++pj) acc += ch[(size_t)pi * W + pj];         \
                        out[(((size_t)r * C + c) * k + i) * k + j] = acc / (T)numel;                 \
                    }                                                                                \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
    }
DEFINE_ROIPOOL_FWD(f32, float)
DEFINE_ROIPOOL_FWD(f64, double)

/*
 * roipool_cuda.cu:68-127.  gradIn[c,pI,pJ] += gradOut[r,c,i,j] / binNumel for
 * every pixel of the bin (atomicAdd in the reference; here in (r,i,j)
 * ascending order).  Parallel over channels: each channel plane is private.
 */
#define DEFINE_ROIPOOL_BWD(SUFFIX, T)                                                                \
    void d2t_oracle_roipool_bwd_##SUFFIX(const T* go, const T* rois, T* gin, int R, int C, int H,   \
                                         int W, int k) {                                             \
        memset(gin, 0, sizeof(T) * (size_t)C * H * W);                                               \
        _Pragma("omp parallel for schedule(static)")                                                 \
        for (int c = 0; c < C; ++c) {                                                                \
            T* ch = gin + (size_t)c * H * W;                                                         \
            for (int r = 0; r < R; ++r) {                                                            \
                const T rI = rois[r * 4 + 0], rJ = rois[r * 4 + 1], rH = rois[r * 4 + 2],            \
                        rW = rois[r * 4 + 3];                                                        \
                for (int i = 0; i < k; ++i) {                                                        \
                    int i0, i1;                                                                      \
                    roipool_edge_##SUFFIX(rI, rH, i, k, H, &i0, &i1);                                \
                    for (int j = 0; j < k; ++j) {                                                    \
                        int j0, j1;                                                                  \
                        roipool_edge_##SUFFIX(rJ, rW, j, k, W, &j0, &j1);                            \
                        const int numel = (i1 - i0) * (j1 - j0);                                     \
                        const T add = go[(((size_t)r * C + c) * k + i) * k + j] / (T)numel;          \
                        for (int pi = i0; pi < i1; ++pi)                                             \
                            for (int pj = j0; pj < j1; ++pj) ch[(size_t)pi * W + pj] += add;         \
                    }                                                                                \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
    }
DEFINE_ROIPOOL_BWD(f32, float)
DEFINE_ROIPOOL_BWD(f64, double)

/* ------------------------------------------------------------------------- */
/* PSROIPool                                                                  */
/* ------------------------------------------------------------------------- */

/*
 * ps_roipool_cuda.cu:10-71.  Channel map targetChannel = (t+1)*(i*k+j) (:58,
 * F6 -- NOT t*k*k + i*k + j); divide only when the cell is non-empty (:67-69).
 * `canonical_map` != 0 selects the textbook R-FCN map t*k*k + i*k + j instead
 * (opt-in extension of the product, never the default).
 */
static inline int ps_channel(int t, int i, int j, int k, int canonical_map) {
    return canonical_map ? (t * k * k + i * k + j) : (t + 1) * (i * k + j);
}

#define DEFINE_PSROIPOOL_FWD(SUFFIX, T)                                                              \
    void d2t_oracle_psroipool_fwd_##SUFFIX(const T* fm, const T* rois, T* out, int R, int nT,       \
                                           int H, int W, int k, int canonical_map) {                 \
        _Pragma("omp parallel for schedule(dynamic, 1)")                                             \
        for (int r = 0; r < R; ++r) {                                                                \
            const T rI = rois[r * 4 + 0], rJ = rois[r * 4 + 1], rH = rois[r * 4 + 2],                \
                    rW = rois[r * 4 + 3];                                                            \
            for (int t = 0; t < nT; ++t)                                                             \
                for (int i = 0; i < k; ++i) {                                                        \
                    int i0, i1;                                                                      \
                    psroipool_edge_##SUFFIX(rI, rH, i, k, H, &i0, &i1);                              \
                    for (int j = 0; j < k; ++j) {                                                    \
                        int j0, j1;                                                                  \
                        psroipool_edge_##SUFFIX(rJ, rW, j, k, W, &j0, &j1);                          \
                        const T* ch = fm + (size_t)ps_channel(t, i, j, k, canonical_map) * H * W;    \
                        T acc = 0;                                                                   \
                        for (int pi = i0; pi < i1; ++pi)                                             \
                            for (int pj = j0; pj < j1; ++pj) acc += ch[(size_t)pi * W + pj];         \
                        const int numel = (i1 - i0) * (j1 - j0);                                     \
                        if (numel > 0) acc /= (T)numel;                                              \
                        out[(((size_t)r * nT + t) * k + i) * k + j] = acc;                           \
                    }                                                                                \
                }                                                                                    \
        }                                                                                            \
    }
DEFINE_PSROIPOOL_FWD(f32, float)
DEFINE_PSROIPOOL_FWD(f64, double)

/*
 * ps_roipool_cuda.cu:76-141.  gradIn[ch,pI,pJ] += gradOut[r,t,i,j] / roiNumel
 * (guarded), atomicAdd in the reference; (r,t,i,j) ascending here.  Serial:
 * several (t,i,j) share a channel (F6) and the op is tiny.
 */
#define DEFINE_PSROIPOOL_BWD(SUFFIX, T)                                                              \
    void d2t_oracle_psroipool_bwd_##SUFFIX(const T* go, const T* rois, T* gin, int R, int nT,       \
                                           int H, int W, int k, int canonical_map) {                 \
        memset(gin, 0, sizeof(T) * (size_t)nT * k * k * H * W);                                      \
        for (int r = 0; r < R; ++r) {                                                                \
            const T rI = rois[r * 4 + 0], rJ = rois[r * 4 + 1], rH = rois[r * 4 + 2],                \
                    rW = rois[r * 4 + 3];                                                            \
            for (int t = 0; t < nT; ++t)                                                             \
                for (int i = 0; i < k; ++i) {                                                        \
                    int i0, i1;                                                                      \
                    psroipool_edge_##SUFFIX(rI, rH, i, k, H, &i0, &i1);                              \
                    for (int j = 0; j < k; ++j) {                                                    \
                        int j0, j1;                                                                  \
                        psroipool_edge_##SUFFIX(rJ, rW, j, k, W, &j0, &j1);                          \
                        T* ch = gin + (size_t)ps_channel(t, i, j, k, canonical_map) * H * W;         \
                        const int numel = (i1 - i0) * (j1 - j0);                                     \
                        T add = go[(((size_t)r * nT + t) * k + i) * k + j];                          \
                        if (numel > 0) add /= (T)numel;                                              \
                        for (int pi = i0; pi < i1; ++pi)                                             \
                            for (int pj = j0; pj < j1; ++pj) ch[(size_t)pi * W + pj] += add;         \
                    }                                                                                \
                }                                                                                    \
        }                                                                                            \
    }
DEFINE_PSROIPOOL_BWD(f32, float)
DEFINE_PSROIPOOL_BWD(f64, double)
