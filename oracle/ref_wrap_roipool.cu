/*
 * TEST INFRASTRUCTURE ONLY.  C-ABI door onto the reference's own, UNMODIFIED
 * ROIPool kernels (roipool_cuda.cu:130-190), compiled from /root/reference.
 */
#include "roipool/roipool_cuda.cu"

template <typename T> static at::ScalarType st();
template <> at::ScalarType st<float>() { return at::ScalarType::Float; }
template <> at::ScalarType st<double>() { return at::ScalarType::Double; }

template <typename T>
static int rp_fwd(const T* fm, const T* rois, T* out, int R, int C, int H, int W, int k) {
    auto tf = at::Tensor::borrow((void*)fm, {C, H, W}, st<T>());
    auto tr = at::Tensor::borrow((void*)rois, {R, 4}, st<T>());
    at::Tensor o = ROIPoolCudaForward(tf, tr, k);
    cudaMemcpyAsync(out, o.raw(), o.nbytes(), cudaMemcpyDeviceToDevice, 0);
    return (int)cudaGetLastError();
}
template <typename T>
static int rp_bwd(const T* go, const T* rois, T* gin, int R, int C, int H, int W, int k) {
    auto tg = at::Tensor::borrow((void*)go, {R, C, k, k}, st<T>());
    auto tr = at::Tensor::borrow((void*)rois, {R, 4}, st<T>());
    at::Tensor g = ROIPoolCudaBackward(tg, tr, H, W);
    cudaMemcpyAsync(gin, g.raw(), g.nbytes(), cudaMemcpyDeviceToDevice, 0);
    return (int)cudaGetLastError();
}

extern "C" {
int ref_roipool_fwd_f32(const float* fm, const float* rois, float* out, int R, int C, int H, int W, int k) {
    return rp_fwd<float>(fm, rois, out, R, C, H, W, k);
}
int ref_roipool_fwd_f64(const double* fm, const double* rois, double* out, int R, int C, int H, int W, int k) {
    return rp_fwd<double>(fm, rois, out, R, C, H, W, k);
}
int ref_roipool_bwd_f32(const float* go, const float* rois, float* gin, int R, int C, int H, int W, int k) {
    return rp_bwd<float>(go, rois, gin, R, C, H, W, k);
}
int ref_roipool_bwd_f64(const double* go, const double* rois, double* gin, int R, int C, int H, int W, int k) {
    return rp_bwd<double>(go, rois, gin, R, C, H, W, k);
}
}
