// corr_generic.cu -- PointwiseCorrelation for any dtype / d_max / stride.
//
// Correctness path: float64 (gradcheck parity with the reference's own tests)
// and any configuration the tuned float32 kernels in corr_tile.cu do not cover.
// Gather formulation, one thread per OUTPUT element, so every output is written
// exactly once: no zero-fill, no atomics (the reference accumulates in global
// memory and uses atomicAdd for grad_FM1, pointwise_correlation_cuda.cu:106,169).
#include "common.cuh"

namespace d2t {

constexpr int kThreads = 256;

// out[b,i,j,ci,cj] = sum_c fm0[b,c,i,j] * fm1[b,c,i-d+ci,j-d+cj]   (live entries; else 0)
template <typename T>
__global__ void __launch_bounds__(kThreads)
corr_fwd_generic_kernel(const T* __restrict__ fm0, const T* __restrict__ fm1, T* __restrict__ out, int B, int C,
                        int H, int W, int d, int stride, CorrOutStrides os) {
    const int k = 2 * d + 1;
    const long long total = (long long)B * H * W * k * k;
    const size_t plane = (size_t)H * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cj = (int)(idx % k);
        long long r = idx / k;
        const int ci = (int)(r % k);
        r /= k;
        const int j = (int)(r % W);
        r /= W;
        const int i = (int)(r % H);
        const int b = (int)(r / H);
        const int di = i - d + ci, dj = j - d + cj;
        T acc = 0;
        if (corr_live(i, di, d, H, stride) && corr_live(j, dj, d, W, stride)) {
            const T* q = fm0 + (size_t)b * C * plane + (size_t)i * W + j;
            const T* key = fm1 + (size_t)b * C * plane + (size_t)di * W + dj;
            for (int c = 0; c < C; ++c) acc += __ldg(q + c * plane) * __ldg(key + c * plane);
        }
        out[(long long)b * os.sb + ((long long)i * W + j) * os.sp + (long long)(ci * k + cj) * os.st] = acc;
    }
}

// one thread per (b,c,i,j): both gradients by gather
template <typename T>
__global__ void __launch_bounds__(kThreads)
corr_bwd_generic_kernel(const T* __restrict__ go, const T* __restrict__ fm0, const T* __restrict__ fm1,
                        T* __restrict__ g0, T* __restrict__ g1, int B, int C, int H, int W, int d, int stride) {
    const int k = 2 * d + 1;
    const long long total = (long long)B * C * H * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % W);
        long long r = idx / W;
        const int i = (int)(r % H);
        r /= H;
        const int c = (int)(r % C);
        const int b = (int)(r / C);
        const T* q = fm0 + ((size_t)b * C + c) * H * W;
        const T* key = fm1 + ((size_t)b * C + c) * H * W;
        const T* gb = go + (size_t)b * H * W * k * k;

        // grad_fm0[b,c,i,j] = sum over live (di,dj) of gO[b,i,j,ci,cj] * fm1[b,c,di,dj]
        T a0 = 0;
        {
            const int di0 = max(0, i - d), di1 = min(i + d, H);
            const int dj0 = max(0, j - d), dj1 = min(j + d, W);
            const T* g = gb + ((size_t)i * W + j) * k * k;
            for (int di = di0; di < di1; di += stride)
                for (int dj = dj0; dj < dj1; dj += stride)
                    a0 += __ldg(g + (di - i + d) * k + (dj - j + d)) * __ldg(key + di * W + dj);
        }
        g0[idx] = a0;

        // grad_fm1[b,c,p,q] with (p,q) = (i,j): queries (qi,qj) whose live set contains (p,q)
        T a1 = 0;
        {
            const int p = i, qq = j;
            const int qi0 = max(0, p - d + 1), qi1 = min(H - 1, p + d);
            const int qj0 = max(0, qq - d + 1), qj1 = min(W - 1, qq + d);
            for (int qi = qi0; qi <= qi1; ++qi) {
                if (!corr_live(qi, p, d, H, stride)) continue;
                for (int qj = qj0; qj <= qj1; ++qj) {
                    if (!corr_live(qj, qq, d, W, stride)) continue;
                    a1 += __ldg(gb + (((size_t)qi * W + qj) * k + (p - qi + d)) * k + (qq - qj + d)) *
                          __ldg(q + qi * W + qj);
                }
            }
        }
        g1[idx] = a1;
    }
}

template <typename T>
int corr_fwd_generic_launch(const T* fm0, const T* fm1, T* out, int B, int C, int H, int W, int d, int stride,
                            CorrOutStrides os, cudaStream_t st) {
    const long long total = (long long)B * H * W * (2 * d + 1) * (2 * d + 1);
    if (total == 0) return D2T_OK;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    long long blocks = (total + kThreads - 1) / kThreads;
    const long long cap = (long long)di.sm_count * 32;
    if (blocks > cap) blocks = cap;
    corr_fwd_generic_kernel<T><<<(int)blocks, kThreads, 0, st>>>(fm0, fm1, out, B, C, H, W, d, stride, os);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

template <typename T>
int corr_bwd_generic_launch(const T* go, const T* fm0, const T* fm1, T* g0, T* g1, int B, int C, int H, int W, int d,
                            int stride, cudaStream_t st) {
    const long long total = (long long)B * C * H * W;
    if (total == 0) return D2T_OK;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    long long blocks = (total + kThreads - 1) / kThreads;
    const long long cap = (long long)di.sm_count * 32;
    if (blocks > cap) blocks = cap;
    corr_bwd_generic_kernel<T><<<(int)blocks, kThreads, 0, st>>>(go, fm0, fm1, g0, g1, B, C, H, W, d, stride);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

template int corr_fwd_generic_launch<float>(const float*, const float*, float*, int, int, int, int, int, int,
                                            CorrOutStrides, cudaStream_t);
template int corr_fwd_generic_launch<double>(const double*, const double*, double*, int, int, int, int, int, int,
                                             CorrOutStrides, cudaStream_t);
template int corr_bwd_generic_launch<float>(const float*, const float*, const float*, float*, float*, int, int, int, int,
                                            int, int, cudaStream_t);
template int corr_bwd_generic_launch<double>(const double*, const double*, const double*, double*, double*, int, int,
                                             int, int, int, int, cudaStream_t);

}  // namespace d2t
