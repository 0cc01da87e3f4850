// gemm_tf32x3.cuh -- internal interface of the TMA + tcgen05 3xTF32 GEMM (gemm_tf32x3.cu).
#pragma once
#include "common.cuh"

namespace d2t {

enum { GEMM_EPI_ROW = 0, GEMM_EPI_COL = 1 };

// One K-major operand: rows x K FP32, K contiguous, row pitch `ld` floats (multiple of 4; base 16-byte aligned).
struct GemmOperand {
    const float* ptr;
    int rows;
    int ld;
};

struct GemmArgs {
    int M, N, K;
    int kblocks, splits;
    int ldo, epilogue, slab_rows;
    int col_rows;            // COL epilogue over a batch: row m = b * col_rows + p goes to out[b * col_stride + n * ldo + p]
    long long col_stride;    // (col_rows == 0: plain out[n * ldo + m])
};

// out = A (M x K) * B^T (N x K).  splits > 1: split s writes its partial sum to the slab starting at row
// s * slab_rows (ROW) / element s * slab_rows * ldo (COL).  bn: N tile (64, 208 or 256).
int gemm_tf32x3(const GemmOperand& A, const GemmOperand& B, float* out, int M, int N, int K, int ldo, int epilogue, int splits,
                int slab_rows, int bn, cudaStream_t st, int col_rows = 0, long long col_stride = 0);

// round to the nearest tf32 (10 explicit mantissa bits), result kept in an FP32 word
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }

}  // namespace d2t
