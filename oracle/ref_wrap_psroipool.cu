/*
 * TEST INFRASTRUCTURE ONLY.  C-ABI door onto the reference's own, UNMODIFIED
 * PSROIPool kernels (ps_roipool_cuda.cu:144-204), compiled from /root/reference.
 */
#include "ps_roipool/ps_roipool_cuda.cu"

template <typename T> static at::ScalarType st();
template <> at::ScalarType st<float>() { return at::ScalarType::Float; }
template <> at::ScalarType st<double>() { return at::ScalarType::Double; }

template <typename T>
static int ps_fwd(const T* fm, const T* rois, T* out, int R, int nT, int H, int W, int k) {
    auto tf = at::Tensor::borrow((void*)fm, {nT * k * k, H, W}, st<T>());
    auto tr = at::Tensor::borrow((void*)rois, {R, 4}, st<T>());
    at::Tensor o = psROIPoolCudaForward(tf, tr, nT, k);
    cudaMemcpyAsync(out, o.raw(), o.nbytes(), cudaMemcpyDeviceToDevice, 0);
    return (int)cudaGetLastError();
}
template <typename T>
static int ps_bwd(const T* go, const T* rois, T* gin, int R, int nT, int H, int W, int k) {
    auto tg = at::Tensor::borrow((void*)go, {R, nT, k, k}, st<T>());
    auto tr = at::Tensor::borrow((void*)rois, {R, 4}, st<T>());
    at::Tensor g = psROIPoolCudaBackward(tg, tr, H, W);
    cudaMemcpyAsync(gin, g.raw(), g.nbytes(), cudaMemcpyDeviceToDevice, 0);
    return (int)cudaGetLastError();
}

extern "C" {
int ref_psroipool_fwd_f32(const float* fm, const float* rois, float* out, int R, int nT, int H, int W, int k) {
    return ps_fwd<float>(fm, rois, out, R, nT, H, W, k);
}
int ref_psroipool_fwd_f64(const double* fm, const double* rois, double* out, int R, int nT, int H, int W, int k) {
    return ps_fwd<double>(fm, rois, out, R, nT, H, W, k);
}
int ref_psroipool_bwd_f32(const float* go, const float* rois, float* gin, int R, int nT, int H, int W, int k) {
    return ps_bwd<float>(go, rois, gin, R, nT, H, W, k);
}
int ref_psroipool_bwd_f64(const double* go, const double* rois, double* gin, int R, int nT, int H, int W, int k) {
    return ps_bwd<double>(go, rois, gin, R, nT, H, W, k);
}
}
