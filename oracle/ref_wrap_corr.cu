/*
 * TEST INFRASTRUCTURE ONLY.  C-ABI door onto the reference's own, UNMODIFIED
 * correlation kernels (compiled from /root/reference where they lie; see
 * oracle/Makefile).  Used by tests/ and bench.py to run the real reference
 * kernels on the B200 as a parity witness and a "reference kernel on B200"
 * timing row.  Nothing in the product links or loads this.
 *
 * Wraps pointwiseCorrelationCudaForward / ...Backward
 * (pointwise_correlation_cuda.cu:178-249).
 */
#include "pointwise_correlation/pointwise_correlation_cuda.cu"

template <typename T> static at::ScalarType st();
template <> at::ScalarType st<float>() { return at::ScalarType::Float; }
template <> at::ScalarType st<double>() { return at::ScalarType::Double; }

template <typename T>
static int corr_fwd(const T* fm0, const T* fm1, T* out, int B, int C, int H, int W, int d, int stride) {
    auto t0 = at::Tensor::borrow((void*)fm0, {B, C, H, W}, st<T>());
    auto t1 = at::Tensor::borrow((void*)fm1, {B, C, H, W}, st<T>());
    at::Tensor o = pointwiseCorrelationCudaForward(t0, t1, d, stride);
    cudaMemcpyAsync(out, o.raw(), o.nbytes(), cudaMemcpyDeviceToDevice, 0);
    return (int)cudaGetLastError();
}

template <typename T>
static int corr_bwd(const T* go, const T* fm0, const T* fm1, T* g0, T* g1, int B, int C, int H, int W, int d,
                    int stride) {
    const int k = 2 * d + 1;
    auto tg = at::Tensor::borrow((void*)go, {B, H, W, k, k}, st<T>());
    auto t0 = at::Tensor::borrow((void*)fm0, {B, C, H, W}, st<T>());
    auto t1 = at::Tensor::borrow((void*)fm1, {B, C, H, W}, st<T>());
    auto r = pointwiseCorrelationCudaBackward(tg, t0, t1, d, stride);
    cudaMemcpyAsync(g0, std::get<0>(r).raw(), std::get<0>(r).nbytes(), cudaMemcpyDeviceToDevice, 0);
    cudaMemcpyAsync(g1, std::get<1>(r).raw(), std::get<1>(r).nbytes(), cudaMemcpyDeviceToDevice, 0);
    return (int)cudaGetLastError();
}

extern "C" {
int ref_corr_fwd_f32(const float* a, const float* b, float* o, int B, int C, int H, int W, int d, int s) {
    return corr_fwd<float>(a, b, o, B, C, H, W, d, s);
}
int ref_corr_fwd_f64(const double* a, const double* b, double* o, int B, int C, int H, int W, int d, int s) {
    return corr_fwd<double>(a, b, o, B, C, H, W, d, s);
}
int ref_corr_bwd_f32(const float* g, const float* a, const float* b, float* g0, float* g1, int B, int C, int H,
                     int W, int d, int s) {
    return corr_bwd<float>(g, a, b, g0, g1, B, C, H, W, d, s);
}
int ref_corr_bwd_f64(const double* g, const double* a, const double* b, double* g0, double* g1, int B, int C,
                     int H, int W, int d, int s) {
    return corr_bwd<double>(g, a, b, g0, g1, B, C, H, W, d, s);
}
}
