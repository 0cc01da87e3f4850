/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Minimal stand-in for <ATen/ATen.h>, just large enough that the reference's
 * three *_cuda.cu files compile UNMODIFIED, from where they lie under
 * /root/reference/detect_to_track/models, with plain nvcc and no libtorch.
 * (The real ATen of torch 2.11 rejects the reference's
 * `AT_DISPATCH_FLOATING_TYPES(x.type(), ...)` call sites -- SURVEY.md F9 -- so
 * building against the shim is also the only way to build them untouched.)
 *
 * What the reference uses and therefore what exists here:
 *   at::Tensor           .size(i) .numel() .options() .type() .data<T>()
 *   at::zeros({..}, options), at::zeros_like(t)
 *   AT_DISPATCH_FLOATING_TYPES(type, name, lambda)   (float + double)
 *
 * Tensors are thin handles over device memory.  `at::zeros` takes its block
 * from a tiny size-keyed pool (so repeated calls do not pay cudaMalloc, like
 * torch's caching allocator) and clears it with cudaMemsetAsync on the legacy
 * default stream, which is the stream the reference launches on (F11).
 */
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

namespace at {

enum class ScalarType { Float, Double };

struct TensorOptions {
    ScalarType dtype;
};

/* what `.type()` returns in the reference's dispatch call sites */
struct DeprecatedTypeProperties {
    ScalarType dtype;
    bool is_cuda() const { return true; }
};

namespace shim_detail {
inline size_t itemsize(ScalarType t) { return t == ScalarType::Float ? 4 : 8; }

struct Pool {
    std::multimap<size_t, void*> free_blocks;
    void* take(size_t bytes) {
        auto it = free_blocks.find(bytes);
        if (it != free_blocks.end()) {
            void* p = it->second;
            free_blocks.erase(it);
            return p;
        }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
            std::fprintf(stderr, "ref_shim: cudaMalloc(%zu) failed\n", bytes);
            std::abort();
        }
        return p;
    }
    void give(size_t bytes, void* p) { free_blocks.emplace(bytes, p); }
    static Pool& get() {
        static Pool pool;
        return pool;
    }
};
}  // namespace shim_detail

class Tensor {
  public:
    Tensor() = default;

    /* borrow caller-owned device memory */
    static Tensor borrow(void* ptr, std::vector<int64_t> sizes, ScalarType dtype) {
        Tensor t;
        t.sizes_ = std::move(sizes);
        t.dtype_ = dtype;
        t.ptr_ = std::shared_ptr<void>(ptr, [](void*) {});
        return t;
    }

    /* pool-owned, zero-initialised */
    static Tensor zeros(std::vector<int64_t> sizes, ScalarType dtype) {
        Tensor t;
        t.sizes_ = std::move(sizes);
        t.dtype_ = dtype;
        const size_t bytes = static_cast<size_t>(t.numel()) * shim_detail::itemsize(dtype);
        void* p = shim_detail::Pool::get().take(bytes);
        cudaMemsetAsync(p, 0, bytes, 0);
        t.ptr_ = std::shared_ptr<void>(p, [bytes](void* q) { shim_detail::Pool::get().give(bytes, q); });
        return t;
    }

    int64_t size(int i) const { return sizes_[i]; }
    const std::vector<int64_t>& sizes() const { return sizes_; }
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : sizes_) n *= s;
        return n;
    }
    TensorOptions options() const { return TensorOptions{dtype_}; }
    DeprecatedTypeProperties type() const { return DeprecatedTypeProperties{dtype_}; }
    ScalarType scalar_type() const { return dtype_; }
    size_t nbytes() const { return static_cast<size_t>(numel()) * shim_detail::itemsize(dtype_); }

    template <typename T>
    T* data() const {
        return static_cast<T*>(ptr_.get());
    }
    void* raw() const { return ptr_.get(); }

  private:
    std::vector<int64_t> sizes_;
    ScalarType dtype_ = ScalarType::Float;
    std::shared_ptr<void> ptr_;
};

inline Tensor zeros(std::initializer_list<int64_t> sizes, TensorOptions o) {
    return Tensor::zeros(std::vector<int64_t>(sizes), o.dtype);
}
inline Tensor zeros_like(const Tensor& t) { return Tensor::zeros(t.sizes(), t.scalar_type()); }

}  // namespace at

#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...)          \
    [&] {                                                    \
        const ::at::ScalarType _st = (TYPE).dtype;           \
        if (_st == ::at::ScalarType::Float) {                \
            using scalar_t = float;                          \
            return __VA_ARGS__();                            \
        } else {                                             \
            using scalar_t = double;                         \
            return __VA_ARGS__();                            \
        }                                                    \
    }()
