// corr_tile.cu -- tuned float32 PointwiseCorrelation kernels (placeholder: not yet enabled).
#include "common.cuh"

namespace d2t {

bool corr_tile_supported(int, int, int, int, int, int) { return false; }
size_t corr_tile_fwd_ws_bytes(int, int, int, int, int) { return 0; }
size_t corr_tile_bwd_ws_bytes(int, int, int, int, int) { return 0; }
int corr_tile_fwd_launch(const float*, const float*, float*, int, int, int, int, int, void*, size_t, cudaStream_t) {
    set_error("corr_tile_fwd: not built");
    return D2T_ERR_BAD_ARG;
}
int corr_tile_bwd_launch(const float*, const float*, const float*, float*, float*, int, int, int, int, int, void*,
                         size_t, cudaStream_t) {
    set_error("corr_tile_bwd: not built");
    return D2T_ERR_BAD_ARG;
}

}  // namespace d2t
