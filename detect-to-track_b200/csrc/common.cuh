// common.cuh -- shared host/device helpers for libd2t_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "d2t_b200.h"

namespace d2t {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void note_launch(int n = 1);  // kernel-launch counter behind d2t_launch_count()

#define D2T_CUDA_TRY(expr)                                         \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::d2t::cuda_fail(_e, #expr); \
    } while (0)

#define D2T_REQUIRE(cond, ...)             \
    do {                                   \
        if (!(cond)) {                     \
            ::d2t::set_error(__VA_ARGS__); \
            return D2T_ERR_BAD_ARG;        \
        }                                  \
    } while (0)

// ---- device attributes (cached per device) -----------------------------------
struct DeviceInfo {
    int sm_count;
    int max_smem_optin;  // bytes of dynamic shared memory a block may opt into
};
int device_info(DeviceInfo* out);

// opt a kernel into `bytes` of dynamic shared memory; the attribute is set once per (device, kernel) and raised only
// when a later launch needs more (cudaFuncSetAttribute is not free: it takes the context lock on every call)
int ensure_dyn_smem(const void* func, size_t bytes);
#define D2T_SMEM_OPTIN(func, bytes)                                                   \
    do {                                                                              \
        int _rc = ::d2t::ensure_dyn_smem(reinterpret_cast<const void*>(func), bytes); \
        if (_rc) return _rc;                                                          \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- [0,1] clamp; reference: common/cuda_common.cuh:9-13 ---------------------
template <typename T>
__device__ __forceinline__ T clamp01(T x) {
    return max(static_cast<T>(0), min(static_cast<T>(1), x));
}

// ---- integer bin edges --------------------------------------------------------
// Written with the reference's exact expression shape (same literal types, same
// operation order) so that nvcc's default contraction yields the same SASS
// arithmetic and the integer edges are bit-identical:
//   ROIPool   roipool_cuda.cu:38-50     start = clamp(r - len/2)            (clamped, F7)
//   PSROIPool ps_roipool_cuda.cu:42-54  start =       r - len/2             (not clamped)
//   centre = start + (scalar(b) + 0.5) * binLen      <- 0.5 is a double literal (F8)
//   e0 = floor(clamp(centre - binLen/2) * n) ; e1 = ceil(clamp(centre + binLen/2) * n)
template <typename T, bool kClampStart>
__device__ __forceinline__ void bin_edge(T r, T len, int b, int k, int n, int& e0, int& e1) {
    const T binLen(len / k);
    T centre;
    if (kClampStart) {
        const T c(clamp01(r - len / 2) + (static_cast<T>(b) + 0.5) * binLen);
        centre = c;
    } else {
        const T c(r - len / 2 + (static_cast<T>(b) + 0.5) * binLen);
        centre = c;
    }
    e0 = static_cast<int>(floor(clamp01(centre - binLen / 2) * n));
    e1 = static_cast<int>(ceil(clamp01(centre + binLen / 2) * n));
}

// ---- correlation output addressing ----------------------------------------------------------------------------
// element (b, pos = i*W + j, t = ci*(2d+1) + cj) of the correlation output lives at out[b*sb + pos*sp + t*st].
// The reference layout (B, H, W, 2d+1, 2d+1) is {H*W*kk, kk, 1}; {*, 1, H*W} is the channel-major ((2d+1)^2, H, W) map
// the tracker concatenates (correlation_tracker.py:64-80), written in place by d2t_corr_fwd_strided_*.
struct CorrOutStrides {
    long long sb, sp, st;
};

// ---- correlation liveness -----------------------------------
// key index p is sampled from query index i iff
//   lo <= p < min(i+d, n)  and  (p - lo) % stride == 0   with lo = max(0, i-d)
__host__ __device__ __forceinline__ bool corr_live(int i, int p, int d, int n, int stride) {
    const int lo = i - d > 0 ? i - d : 0;
    const int hi = i + d < n ? i + d : n;
    return p >= lo && p < hi && ((p - lo) % stride) == 0;
}

}  // namespace d2t
