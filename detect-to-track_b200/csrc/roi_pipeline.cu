// roi_pipeline.cu -- device-side region-proposal pipeline: anchor decode + confidence filter + NMS.  sm_100a.
//
// The reference runs this step on the HOST between the RPN and the R-FCN heads (trainer.py:178-190, inference.py:78-84):
// both RPN outputs are copied to numpy, decoded by ml_utils' `frcnn_box_decode`, filtered by its confidence / NMS
// `region_filter`, and copied back.  Those two device<->host round trips are what bounds the training step once the ops
// are fast.  ml_utils' source is not available, so its exact semantics are UNPINNED; the ones
// implemented here are the standard Faster R-CNN ones for fractional (centre_i, centre_j, height, width) boxes:
//     decode   i = a_i + d_i * a_h,  j = a_j + d_j * a_w,  h = a_h * exp(d_h),  w = a_w * exp(d_w)
//     filter   keep score > conf_thresh                                      (cfg/default.yaml:23, 0.3)
//     NMS      candidates in descending score order (ties: lower anchor index first); a candidate is dropped when its IoU
//              with an already kept one exceeds iou_thresh (0.5 / 0.3); at most max_rois survive (3000)
// Everything stays on the device and every step has a fixed order => deterministic:
//     1. rp_decode_kernel      one thread per anchor: box + score (filtered anchors get score -inf)
//     2. (caller)              sort by score, descending, stable -- torch.sort on the device
//     3. rp_iou_mask_kernel    64 x 64 candidate blocks: bit c of mask[r][cb] = IoU(r, 64*cb + c) > thresh, c after r
//     4. rp_nms_scan_kernel    ONE warp walks the candidates in order; lane l holds word l, l + 32, ... of the "removed" bitset
//     5. rp_gather_kernel      kept boxes, in score order, compacted into the (max_rois, 4) output; the count stays on the device
#include "common.cuh"

namespace d2t {

namespace {

__device__ __forceinline__ float4 rp_corners(float4 b) {   // ijhw -> (i0, j0, i1, j1)
    return make_float4(b.x - 0.5f * b.z, b.y - 0.5f * b.w, b.x + 0.5f * b.z, b.y + 0.5f * b.w);
}
__device__ __forceinline__ float rp_iou(float4 a, float4 b) {
    const float4 ca = rp_corners(a), cb = rp_corners(b);
    const float ih = fminf(ca.z, cb.z) - fmaxf(ca.x, cb.x);
    const float iw = fminf(ca.w, cb.w) - fmaxf(ca.y, cb.y);
    const float inter = fmaxf(ih, 0.f) * fmaxf(iw, 0.f);
    const float uni = a.z * a.w + b.z * b.w - inter;
    return uni > 0.f ? inter / uni : 0.f;
}

__global__ void __launch_bounds__(256)
rp_decode_kernel(const float4* __restrict__ anchors, const float4* __restrict__ offsets, const float* __restrict__ conf,
                 float4* __restrict__ boxes, float* __restrict__ scores, int A, float thresh) {
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < A; a += gridDim.x * blockDim.x) {
        const float4 an = __ldg(anchors + a), d = __ldg(offsets + a);
        boxes[a] = make_float4(an.x + d.x * an.z, an.y + d.y * an.w, an.z * expf(d.z), an.w * expf(d.w));
        const float s = __ldg(conf + a);
        scores[a] = s > thresh ? s : -INFINITY;
    }
}

// sorted candidate n = boxes[order[n]]; M candidates, MW = ceil(M / 64) words per row
__global__ void __launch_bounds__(64)
rp_iou_mask_kernel(const float4* __restrict__ boxes, const long long* __restrict__ order, unsigned long long* __restrict__ mask,
                   int M, int MW, float thresh) {
    __shared__ float4 colb[64];
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;   // only candidates after r can be suppressed by r
    const int c = cb * 64 + threadIdx.x;
    if (c < M) colb[threadIdx.x] = __ldg(boxes + order[c]);
    __syncthreads();
    const int r = rb * 64 + threadIdx.x;
    if (r >= M) return;
    const float4 me = __ldg(boxes + order[r]);
    unsigned long long bits = 0ull;
    const int nc = min(64, M - cb * 64);
    for (int x = (cb == rb ? threadIdx.x + 1 : 0); x < nc; ++x)
        if (rp_iou(me, colb[x]) > thresh) bits |= 1ull << x;
    mask[(size_t)r * MW + cb] = bits;
}

// one warp; the "removed" bitset is distributed over the lanes (lane l holds words l, l + 32, ...).  Candidates are
// resolved strictly in order, eight at a time: the eight mask rows are fetched first (independent loads, one memory
// latency per chunk instead of one per candidate), then applied from registers.
// keep[n] = 1 for survivors, count = number kept.
constexpr int kRpWordsPerLane = 8;   // M <= 32 * 8 * 64 = 16384 candidates
constexpr int kRpChunk = 8;
__global__ void __launch_bounds__(32)
rp_nms_scan_kernel(const unsigned long long* __restrict__ mask, const float* __restrict__ sorted_scores, int* __restrict__ keep,
                   int* __restrict__ count, int M, int MW, int max_rois) {
    const int lane = threadIdx.x;
    // filtered anchors (score -inf) sort to the end: the valid candidates are a prefix
    int nValid = M;
    for (int n = lane; n < M; n += 32)
        if (!(__ldg(sorted_scores + n) > -INFINITY)) { nValid = n; break; }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) nValid = min(nValid, __shfl_xor_sync(0xffffffffu, nValid, sh));
    for (int n = nValid + lane; n < M; n += 32) keep[n] = 0;
    unsigned long long removed[kRpWordsPerLane];
#pragma unroll
    for (int w = 0; w < kRpWordsPerLane; ++w) removed[w] = 0ull;
    int kept = 0;
    for (int n0 = 0; n0 < nValid; n0 += kRpChunk) {
        unsigned long long rows[kRpChunk][kRpWordsPerLane];
#pragma unroll
        for (int c = 0; c < kRpChunk; ++c) {
            const int n = n0 + c;
#pragma unroll
            for (int w = 0; w < kRpWordsPerLane; ++w) {
                const int wd = lane + 32 * w;
                // words before the candidate's own are never written by rp_iou_mask_kernel (nothing earlier is suppressed)
                rows[c][w] = (n < nValid && wd < MW && wd >= (n >> 6)) ? __ldg(mask + (size_t)n * MW + wd) : 0ull;
            }
        }
#pragma unroll
        for (int c = 0; c < kRpChunk; ++c) {
            const int n = n0 + c;
            if (n >= nValid) break;
            const int word = n >> 6, owner = word & 31, slot = word >> 5;
            unsigned long long mine = 0ull;
#pragma unroll
            for (int w = 0; w < kRpWordsPerLane; ++w)
                if (w == slot) mine = removed[w];
            const unsigned long long rw = __shfl_sync(0xffffffffu, mine, owner);
            const bool alive = !((rw >> (n & 63)) & 1ull) && kept < max_rois;
            if (lane == 0) keep[n] = alive ? 1 : 0;
            if (alive) {
                ++kept;
#pragma unroll
                for (int w = 0; w < kRpWordsPerLane; ++w) removed[w] |= rows[c][w];
            }
        }
    }
    if (lane == 0) *count = kept;
}

// exclusive prefix of keep (single CTA, M <= 16384) and gather into rois[0 .. count); the tail is zero-filled
__global__ void __launch_bounds__(1024)
rp_gather_kernel(const float4* __restrict__ boxes, const long long* __restrict__ order, const int* __restrict__ keep,
                 float4* __restrict__ rois, int M, int max_rois) {
    __shared__ int warpsum[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base_s = 0;
    for (int r = tid; r < max_rois; r += blockDim.x) rois[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int n0 = 0; n0 < M; n0 += blockDim.x) {
        const int n = n0 + tid;
        const int k = n < M ? keep[n] : 0;
        int incl = k;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, sh);
            if (lane >= sh) incl += o;
        }
        if (lane == 31) warpsum[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warpsum[w];
        const int pos = base_s + woff + incl - k;
        if (k && pos < max_rois) rois[pos] = __ldg(boxes + order[n]);
        __syncthreads();
        if (tid == blockDim.x - 1) base_s = pos + k;
        __syncthreads();
    }
}

}  // namespace

size_t roi_pipeline_ws_bytes(int A, int pre_nms) {
    const int M = pre_nms < A ? pre_nms : A;
    const size_t MW = (size_t)(M + 63) / 64;
    return align_up((size_t)M * MW * 8, 256) + align_up((size_t)M * 4, 256);
}

int roi_decode_launch(const float* anchors, const float* offsets, const float* conf, float* boxes, float* scores, int A,
                      float thresh, cudaStream_t st) {
    D2T_REQUIRE(A >= 0, "roi_decode: bad anchor count");
    if (A == 0) return D2T_OK;
    rp_decode_kernel<<<ceil_div(A, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(anchors), reinterpret_cast<const float4*>(offsets),
                                                        conf, reinterpret_cast<float4*>(boxes), scores, A, thresh);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch();
    return D2T_OK;
}

int roi_nms_launch(const float* boxes, const long long* order, const float* sorted_scores, float* rois, int* count, int A,
                   int pre_nms, int max_rois, float iou_thresh, void* ws, size_t ws_bytes, cudaStream_t st) {
    D2T_REQUIRE(A >= 0 && pre_nms > 0 && max_rois > 0, "roi_nms: bad arguments");
    const int M = pre_nms < A ? pre_nms : A;
    D2T_REQUIRE(M <= 16384, "roi_nms: at most 16384 candidates enter the NMS (got %d)", M);
    if (M == 0) {
        D2T_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), st));
        D2T_CUDA_TRY(cudaMemsetAsync(rois, 0, (size_t)max_rois * 16, st));
        return D2T_OK;
    }
    const int MW = (M + 63) / 64;
    if (ws == nullptr || ws_bytes < roi_pipeline_ws_bytes(A, pre_nms)) {
        set_error("roi_nms: workspace too small (%zu < %zu)", ws_bytes, roi_pipeline_ws_bytes(A, pre_nms));
        return D2T_ERR_WORKSPACE;
    }
    unsigned long long* mask = static_cast<unsigned long long*>(ws);
    int* keep = reinterpret_cast<int*>(static_cast<char*>(ws) + align_up((size_t)M * MW * 8, 256));
    rp_iou_mask_kernel<<<dim3(MW, MW), 64, 0, st>>>(reinterpret_cast<const float4*>(boxes), order, mask, M, MW, iou_thresh);
    D2T_CUDA_TRY(cudaGetLastError());
    rp_nms_scan_kernel<<<1, 32, 0, st>>>(mask, sorted_scores, keep, count, M, MW, max_rois);
    D2T_CUDA_TRY(cudaGetLastError());
    rp_gather_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const float4*>(boxes), order, keep, reinterpret_cast<float4*>(rois), M,
                                         max_rois);
    D2T_CUDA_TRY(cudaGetLastError());
    note_launch(3);
    return D2T_OK;
}

}  // namespace d2t
