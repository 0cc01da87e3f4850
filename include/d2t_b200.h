/*
 * d2t_b200.h -- C ABI of libd2t_b200.so: detect-to-track's custom-op hot path
 * (PointwiseCorrelation, ROIPool, PSROIPool; forward + backward) as
 * hand-written CUDA for sm_100a.
 *
 * This is the drop-in boundary.  Each entry point replaces one function of the
 * reference's pybind surface (paths relative to
 * /root/reference/detect_to_track/models/):
 *
 *   d2t_corr_fwd_*        <- pointwise_correlation_forward   pointwise_correlation/pointwise_correlation.cpp:23-33,52-56
 *                            (launcher pointwise_correlation_cuda.cu:178-210, kernel :63-111)
 *   d2t_corr_bwd_*        <- pointwise_correlation_backward  pointwise_correlation.cpp:36-48,57-61
 *                            (launcher pointwise_correlation_cuda.cu:214-249, kernel :121-174)
 *   d2t_roipool_fwd_*     <- roipool_forward                 roipool/roipool.cpp:22-31,49-53
 *                            (launcher roipool_cuda.cu:130-158, kernel :6-63)
 *   d2t_roipool_bwd_*     <- roipool_backward                roipool/roipool.cpp:34-45,54-58
 *                            (launcher roipool_cuda.cu:161-190, kernel :68-127)
 *   d2t_psroipool_fwd_*   <- ps_roipool_forward              ps_roipool/ps_roipool.cpp:23-33,51-55
 *                            (launcher ps_roipool_cuda.cu:144-174, kernel :10-71)
 *   d2t_psroipool_bwd_*   <- ps_roipool_backward             ps_roipool/ps_roipool.cpp:36-47,56-60
 *                            (launcher ps_roipool_cuda.cu:177-204, kernel :76-141)
 *
 * Conventions
 *   - Plain pointers and ints only; no torch / ATen types.  All data pointers
 *     are DEVICE pointers to dense, C-contiguous arrays of the stated shape.
 *   - `_f32` = float, `_f64` = double (the reference dispatches exactly these
 *     two, AT_DISPATCH_FLOATING_TYPES).
 *   - The CALLER owns every buffer.  Outputs need NOT be zero-initialised: the
 *     kernels define every output element (dead correlation entries, untouched
 *     gradient pixels and empty PSROI cells are written as 0).  The reference
 *     instead allocates zeroed outputs itself (e.g. pointwise_correlation_cuda.cu:192).
 *   - `stream` is a cudaStream_t (NULL = legacy default stream, which is what
 *     the reference always uses).  Calls are asynchronous.
 *   - `ws` / `ws_bytes`: caller-provided scratch of at least
 *     d2t_*_workspace_bytes(...) bytes, 256-byte aligned; may be NULL when that
 *     function returns 0.  Contents are undefined afterwards.
 *   - Return value: 0 on success; non-zero on bad arguments or a CUDA launch
 *     error, with a message retrievable from d2t_last_error() (thread-local).
 *   - Re-entrant; no global mutable state besides per-device caches of device attributes and of the kernels'
 *     shared-memory opt-in, and the launch counter.  No environment variable is read anywhere.
 *   - Results are deterministic: no floating-point atomics anywhere.
 *
 * RoIs are fractional (centre_i, centre_j, height, width) in [0,1] map units,
 * one image per call, exactly like the reference.
 */
#ifndef D2T_B200_H
#define D2T_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D2T_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define D2T_API __attribute__((visibility("default")))
#else
#define D2T_API
#endif

/* status codes */
#define D2T_OK 0
#define D2T_ERR_BAD_ARG 1
#define D2T_ERR_CUDA 2
#define D2T_ERR_WORKSPACE 3

D2T_API int d2t_abi_version(void);
D2T_API const char* d2t_last_error(void);
/* number of kernels this library has launched in this process (monotonic; instrumentation for bench.py) */
D2T_API unsigned long long d2t_launch_count(void);

/* ---- PointwiseCorrelation ------------------------------------------------
 * fm0, fm1 : (B, C, H, W)            feature maps at t and t+tau
 * out      : (B, H, W, 2d+1, 2d+1)   out[b,i,j,ci,cj] = sum_c fm0[b,c,i,j] * fm1[b,c,i-d+ci,j-d+cj]
 *            for the reference's live displacement set (exclusive upper bound,
 *            stride phase tied to the clamped start: SURVEY.md F4/F5); 0 elsewhere.
 */
D2T_API size_t d2t_corr_fwd_workspace_bytes(int B, int C, int H, int W, int d_max, int stride, int elem_size);
D2T_API int d2t_corr_fwd_f32(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, int d_max,
                     int stride, void* ws, size_t ws_bytes, void* stream);
D2T_API int d2t_corr_fwd_f64(const double* fm0, const double* fm1, double* out, int B, int C, int H, int W, int d_max,
                     int stride, void* ws, size_t ws_bytes, void* stream);

/* Kernel family of d2t_corr_fwd_f32 / d2t_corr_fwd_strided_f32 -- like the backward, a function of (C, d_max, stride) only:
 *   d_max == 8, stride == 1, C >= 128 : tensor cores (tcgen05 + TMEM, 3xTF32; csrc/corr_umma_fwd.cu: both operands
 *       MN-major, i.e. NCHW rows as they lie in memory).  |err| <= (2e-6 + 1.5e-8 * C) * sum_c |fm0 * fm1|: the split
 *       itself is good to 7e-7, the C term is the tensor core's FP32 accumulator, which TRUNCATES -- sums of same-signed
 *       terms (post-ReLU maps) come out low by about 0.7e-8 * C relative (-1.4e-5 at C = 2048, measured).  Inside rtol
 *       1e-4, looser than FP32 FMAs (3e-7); d2t_corr_fwd_f32_simt keeps those.  No workspace.
 *       Dead entries are exact zeros.  Non-finite inputs: a NaN/Inf key or query poisons the 128 x 256 accumulator tiles
 *       it is staged into (the positions of its 8 x 16 tile / 8-row patch chunk), not only the windows containing it.
 *   otherwise : FP32 FMAs (SIMT band kernel for d_max in {4, 8} with stride 1, generic gather kernel for the rest).
 * d2t_corr_fwd_f32_simt / d2t_corr_fwd_f32_tc select a family explicitly (_simt takes the workspace reported by
 * d2t_corr_fwd_simt_workspace_bytes; _tc requires d_max == 8, stride == 1). */
D2T_API size_t d2t_corr_fwd_simt_workspace_bytes(int B, int C, int H, int W, int d_max, int stride);
D2T_API int d2t_corr_fwd_f32_simt(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, int d_max,
                          int stride, void* ws, size_t ws_bytes, void* stream);
D2T_API int d2t_corr_fwd_f32_tc(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, int d_max,
                        int stride, void* ws, size_t ws_bytes, void* stream);

/* Strided output (tracker glue fusion, correlation_tracker.py:64-80): element (b, pos = i*W + j, t = ci*(2d+1) + cj)
 * is written to out[b*batch_stride + pos*pos_stride + t*disp_stride] (strides in elements).  {H*W*kk, kk, 1} is the
 * reference layout; {anything, 1, H*W} writes the ((2d+1)^2, H, W) channel-major map the tracker feeds to ROIPool
 * straight into a slice of its concatenated feature buffer -- bit-identical to
 * out.squeeze(0).view(H, W, -1).permute(2, 0, 1) of d2t_corr_fwd_*, without the permute copy and the torch.cat.
 * Same workspace as d2t_corr_fwd_*. */
D2T_API int d2t_corr_fwd_strided_f32(const float* fm0, const float* fm1, float* out, int B, int C, int H, int W, int d_max,
                             int stride, long long batch_stride, long long pos_stride, long long disp_stride, void* ws,
                             size_t ws_bytes, void* stream);
D2T_API int d2t_corr_fwd_strided_f64(const double* fm0, const double* fm1, double* out, int B, int C, int H, int W, int d_max,
                             int stride, long long batch_stride, long long pos_stride, long long disp_stride, void* ws,
                             size_t ws_bytes, void* stream);

/* grad_out : (B, H, W, 2d+1, 2d+1);  grad_fm0, grad_fm1 : (B, C, H, W)
 *
 * Kernel family of d2t_corr_bwd_f32 -- a function of (C, d_max, stride) only, never of B, H or W, so the gradients of
 * an image do not depend on the batch it is in:
 *   d_max == 8, stride == 1, C >= 128 : tensor cores (tcgen05 + TMEM, 3xTF32 split with a round-to-nearest hi part).
 *       Each gradient element agrees with the exact sum to |err| <= 2e-6 * sum |grad_out * fm| over its 256 terms
 *       (measured 9e-7, tools/umma_sw128_test.cu): inside rtol 1e-4 of the result's scale, looser than FP32 FMAs.
 *       Needs the workspace d2t_corr_bwd_workspace_bytes reports (a flipped copy of grad_out); a missing workspace is
 *       an error (D2T_ERR_WORKSPACE), never a silent change of kernel.  Non-finite inputs: the hi/lo split turns
 *       +-Inf into NaN and the dense 128-position tile product spreads a NaN/Inf of a 23x31 halo patch to every
 *       position of that tile, whereas the reference only poisons the windows that contain the value.
 *   otherwise : FP32 FMAs (SIMT band kernel for d_max in {4, 8} with stride 1, generic gather kernel for the rest).
 * d2t_corr_bwd_f32_simt / d2t_corr_bwd_f32_tc select a family explicitly (same contract; _tc requires d_max == 8,
 * stride == 1 and uses d2t_corr_bwd_tc_workspace_bytes).  All families are bitwise reproducible run to run (no atomics).
 */
D2T_API size_t d2t_corr_bwd_workspace_bytes(int B, int C, int H, int W, int d_max, int stride, int elem_size);
D2T_API int d2t_corr_bwd_f32(const float* grad_out, const float* fm0, const float* fm1, float* grad_fm0,
                     float* grad_fm1, int B, int C, int H, int W, int d_max, int stride, void* ws,
                     size_t ws_bytes, void* stream);
D2T_API int d2t_corr_bwd_f64(const double* grad_out, const double* fm0, const double* fm1, double* grad_fm0,
                     double* grad_fm1, int B, int C, int H, int W, int d_max, int stride, void* ws,
                     size_t ws_bytes, void* stream);
D2T_API int d2t_corr_bwd_f32_simt(const float* grad_out, const float* fm0, const float* fm1, float* grad_fm0,
                        float* grad_fm1, int B, int C, int H, int W, int d_max, int stride, void* ws,
                        size_t ws_bytes, void* stream);
D2T_API size_t d2t_corr_bwd_tc_workspace_bytes(int B, int C, int H, int W, int d_max, int stride);
D2T_API int d2t_corr_bwd_f32_tc(const float* grad_out, const float* fm0, const float* fm1, float* grad_fm0,
                        float* grad_fm1, int B, int C, int H, int W, int d_max, int stride, void* ws,
                        size_t ws_bytes, void* stream);

/* ---- ROIPool (average pooling) ------------------------------
 * fm   : (C, H, W);  rois : (R, 4) same dtype;  out : (R, C, r_hw, r_hw)
 * Empty bins give 0/0 = NaN like the reference (F7).
 */
D2T_API size_t d2t_roipool_fwd_workspace_bytes(int R, int C, int H, int W, int r_hw, int elem_size);
D2T_API int d2t_roipool_fwd_f32(const float* fm, const float* rois, float* out, int R, int C, int H, int W, int r_hw,
                        void* ws, size_t ws_bytes, void* stream);
D2T_API int d2t_roipool_fwd_f64(const double* fm, const double* rois, double* out, int R, int C, int H, int W,
                        int r_hw, void* ws, size_t ws_bytes, void* stream);

/* d2t_roipool_fwd_f32 sums each bin through per-row prefix sums (rounding-level differences from the reference's
 * left-to-right order, |err| <~ 1e-6 * row magnitude).  This variant keeps the reference's summation order and is
 * bit-identical to roipool_cuda.cu:56-61 (several times slower). */
D2T_API int d2t_roipool_fwd_f32_exact(const float* fm, const float* rois, float* out, int R, int C, int H, int W, int r_hw,
                        void* ws, size_t ws_bytes, void* stream);

/* grad_out : (R, C, r_hw, r_hw);  grad_fm : (C, H, W) */
D2T_API size_t d2t_roipool_bwd_workspace_bytes(int R, int C, int H, int W, int r_hw, int elem_size);
D2T_API int d2t_roipool_bwd_f32(const float* grad_out, const float* rois, float* grad_fm, int R, int C, int H, int W,
                        int r_hw, void* ws, size_t ws_bytes, void* stream);
D2T_API int d2t_roipool_bwd_f64(const double* grad_out, const double* rois, double* grad_fm, int R, int C, int H,
                        int W, int r_hw, void* ws, size_t ws_bytes, void* stream);

/* ---- PSROIPool -------------------------------------------------------------
 * fm : (n_targets*r_hw^2, H, W);  rois : (R, 4);  out : (R, n_targets, r_hw, r_hw)
 * flags: bit 0 (D2T_PS_CANONICAL_MAP) selects the textbook R-FCN channel map
 *        t*k*k + i*k + j; 0 (default) = the reference's (t+1)*(i*k+j) (F6).
 */
#define D2T_PS_CANONICAL_MAP 1
D2T_API size_t d2t_psroipool_fwd_workspace_bytes(int R, int n_targets, int H, int W, int r_hw, int elem_size);
D2T_API int d2t_psroipool_fwd_f32(const float* fm, const float* rois, float* out, int R, int n_targets, int H, int W,
                          int r_hw, int flags, void* ws, size_t ws_bytes, void* stream);
D2T_API int d2t_psroipool_fwd_f64(const double* fm, const double* rois, double* out, int R, int n_targets, int H,
                          int W, int r_hw, int flags, void* ws, size_t ws_bytes, void* stream);

/* grad_out : (R, n_targets, r_hw, r_hw);  grad_fm : (n_targets*r_hw^2, H, W) */
D2T_API size_t d2t_psroipool_bwd_workspace_bytes(int R, int n_targets, int H, int W, int r_hw, int elem_size);
D2T_API int d2t_psroipool_bwd_f32(const float* grad_out, const float* rois, float* grad_fm, int R, int n_targets,
                          int H, int W, int r_hw, int flags, void* ws, size_t ws_bytes, void* stream);
D2T_API int d2t_psroipool_bwd_f64(const double* grad_out, const double* rois, double* grad_fm, int R, int n_targets,
                          int H, int W, int r_hw, int flags, void* ws, size_t ws_bytes, void* stream);

/* ---- PSROIPool over a batch of frames (float32) ---------------------------
 * Extension: the reference pools one frame per call (rfcn.py:36-41, called per frame from trainer.py:208-209).
 * fm : (N, n_targets*r_hw^2, H, W);  rois : (N, R, 4);  out / grad_out : (N, R, n_targets, r_hw, r_hw);
 * grad_fm : (N, n_targets*r_hw^2, H, W).  Frame n uses rois[n].  One set of launches covers all frames.  The forward is
 * bit-identical to N single-frame calls (and to the reference kernel); the backward is bit-identical to N single-frame
 * calls for more than 8 targets (both run pool_ps3.cu: targets on the lanes, a CTA per (frame, pixel row, column block), no
 * floating-point atomics, no difference arrays) and agrees within FP32 rounding otherwise (row-list kernels).  Always
 * bitwise reproducible run to run.
 */
D2T_API size_t d2t_psroipool_fwd_batched_workspace_bytes(int N, int R, int n_targets, int H, int W, int r_hw, int elem_size);
D2T_API int d2t_psroipool_fwd_batched_f32(const float* fm, const float* rois, float* out, int N, int R, int n_targets,
                          int H, int W, int r_hw, int flags, void* ws, size_t ws_bytes, void* stream);
D2T_API size_t d2t_psroipool_bwd_batched_workspace_bytes(int N, int R, int n_targets, int H, int W, int r_hw, int elem_size);
D2T_API int d2t_psroipool_bwd_batched_f32(const float* grad_out, const float* rois, float* grad_fm, int N, int R,
                          int n_targets, int H, int W, int r_hw, int flags, void* ws, size_t ws_bytes, void* stream);

/* ---- PSROIPool + vote over a batch of frames (float32) ---------------------------
 * Extension: the R-FCN heads follow PSROIPool by the vote `pooled.mean(-1).mean(-1)` (rfcn.py:40-41).  These entry points
 * return the vote directly -- out / grad_out : (N, R, n_targets) -- without forming the (N, R, n_targets, r_hw, r_hw)
 * tensor or launching the two reductions (forward: one warp per (frame, RoI, target); backward: the PSROIPool backward
 * kernels reading grad_out / r_hw^2 per bin).  fm, rois, grad_fm as in the batched PSROIPool.  Values agree with the
 * composition to FP32 rounding; bitwise reproducible.  d2t_psroipool_vote_supported tells whether the shape fits
 * (RoI bitmasks and one plane in shared memory); the forward needs no workspace, the backward
 * d2t_psroipool_vote_bwd_workspace_bytes (ABI version 2). */
D2T_API int d2t_psroipool_vote_supported(int N, int R, int n_targets, int H, int W, int r_hw);
D2T_API int d2t_psroipool_vote_fwd_f32(const float* fm, const float* rois, float* out, int N, int R, int n_targets, int H,
                               int W, int r_hw, int flags, void* stream);
D2T_API size_t d2t_psroipool_vote_bwd_workspace_bytes(int N, int R, int n_targets, int H, int W, int r_hw);
D2T_API int d2t_psroipool_vote_bwd_f32(const float* grad_out, const float* rois, float* grad_fm, int N, int R, int n_targets,
                               int H, int W, int r_hw, int flags, void* ws, size_t ws_bytes, void* stream);

/* ---- Fused track head: ROIPool -> view -> Linear (float32) -----------------------
 * Extension beside the API-parity ops: the reference's track-regression head (correlation_tracker.py:82-85)
 *     t_hat = Linear(C*r_hw^2, n_out)( ROIPool(r_hw)(fm, rois).view(R, -1) )
 * computed without materialising the pooled (R, C, r_hw, r_hw) tensor (111 MB at the D&T size): the channel
 * contraction runs first, on the un-pooled map, as a tcgen05 3xTF32 GEMM fed by TMA, followed by a position-
 * sensitive pooling of its n_out*r_hw^2-channel result; the backward is the transpose (csrc/track_head.cu).
 *   fm : (C, H, W);  rois : (R, 4);  weight : (n_out, C*r_hw^2) row-major as torch.nn.Linear;  bias : (n_out) or NULL
 *   out / grad_out : (R, n_out);  grad_fm : (C, H, W);  grad_weight : like weight;  grad_bias : (n_out)
 * Any of grad_fm / grad_weight / grad_bias may be NULL (not computed).  Requires n_out * r_hw^2 <= 256, n_out <= 8.
 * Same bins, clamped RoI start and empty-bin NaN as d2t_roipool_fwd_f32; values agree with the composition to FP32
 * rounding (3xTF32 with the tensor core's FP32 accumulator: measured |err| <= 2.5e-6 * sum |a||b| per contraction for the
 * longest, 1891-term chains, tools/gemm_sweep.py; inside rtol 1e-4 of the result's scale).  Bitwise reproducible.
 */
D2T_API size_t d2t_trackhead_fwd_workspace_bytes(int R, int C, int H, int W, int r_hw, int n_out);
D2T_API int d2t_trackhead_fwd_f32(const float* fm, const float* rois, const float* weight, const float* bias, float* out,
                          int R, int C, int H, int W, int r_hw, int n_out, void* ws, size_t ws_bytes, void* stream);
D2T_API size_t d2t_trackhead_bwd_workspace_bytes(int R, int C, int H, int W, int r_hw, int n_out);
D2T_API int d2t_trackhead_bwd_f32(const float* grad_out, const float* fm, const float* rois, const float* weight,
                          float* grad_fm, float* grad_weight, float* grad_bias, int R, int C, int H, int W, int r_hw,
                          int n_out, void* ws, size_t ws_bytes, void* stream);

/* The same operator over N images that share weight and bias -- the track features of N frame pairs (extension: the
 * reference runs its tracker one pair per call, correlation_tracker.py:56-87).  fm : (N, C, H, W);  rois : (N, R, 4);
 * out / grad_out : (N, R, n_out);  grad_fm : (N, C, H, W);  grad_weight / grad_bias : sums over the N images.  One set of
 * launches: the forward and grad_fm GEMMs run over all N*H*W positions, the grad_weight GEMM contracts over them. */
D2T_API size_t d2t_trackhead_fwd_batched_workspace_bytes(int N, int R, int C, int H, int W, int r_hw, int n_out);
D2T_API int d2t_trackhead_fwd_batched_f32(const float* fm, const float* rois, const float* weight, const float* bias, float* out,
                          int N, int R, int C, int H, int W, int r_hw, int n_out, void* ws, size_t ws_bytes, void* stream);
D2T_API size_t d2t_trackhead_bwd_batched_workspace_bytes(int N, int R, int C, int H, int W, int r_hw, int n_out);
D2T_API int d2t_trackhead_bwd_batched_f32(const float* grad_out, const float* fm, const float* rois, const float* weight,
                          float* grad_fm, float* grad_weight, float* grad_bias, int N, int R, int C, int H, int W, int r_hw,
                          int n_out, void* ws, size_t ws_bytes, void* stream);

/* ---- 3xTF32 GEMM building block (float32) ------------------------------------------
 * out = A (M x K, row pitch lda) * B^T (N x K, row pitch ldb), both K-major with 16-byte-aligned bases and pitches, on
 * tcgen05 fed by TMA (csrc/gemm_tf32x3.cu; measured |err| <= 2.5e-6 * sum |a||b| at K = 1891, 9e-7 at K = 256).  The contraction kernel of the fused track head,
 * exported for its own parity and timing tests.  splits > 1: split s of the K range writes its partial product to rows
 * [s*M, (s+1)*M) of `out` (row-major, ldo) / to out + s*M*ldo (col_major_out: out[n*ldo + m]); n_tile in {64, 208, 256}. */
D2T_API int d2t_gemm_tf32x3_f32(const float* A, const float* B, float* out, int M, int N, int K, int lda, int ldb, int ldo,
                        int col_major_out, int splits, int n_tile, void* stream);

/* ---- Device-side RoI pipeline (float32) -----------------------------------------
 * Extension replacing the host round trip between the RPN and the R-FCN heads (trainer.py:178-190, inference.py:78-84:
 * RPN outputs -> numpy -> ml_utils frcnn_box_decode + region_filter -> device).  ml_utils' source is unavailable, so the
 * semantics are the standard Faster R-CNN ones (parity UNPINNED, see csrc/roi_pipeline.cu): decode
 *   i = a_i + d_i*a_h, j = a_j + d_j*a_w, h = a_h*exp(d_h), w = a_w*exp(d_w), keep conf > conf_thresh, then greedy NMS
 * over the candidates in descending score order (IoU > iou_thresh suppresses), at most max_rois survivors.
 *   d2t_roi_decode_filter_f32 : anchors, offsets (A, 4) ijhw; conf (A) -> boxes (A, 4), scores (A) (-inf when filtered)
 *   [caller sorts scores descending, stable, on the device -> sorted_scores (A), order (A) int64]
 *   d2t_roi_nms_f32           : the first min(A, pre_nms) candidates (<= 16384) -> rois (max_rois, 4) in score order,
 *                               zero-filled past *count (device int32).  No device->host synchronisation anywhere.
 */
D2T_API size_t d2t_roi_nms_workspace_bytes(int n_anchors, int pre_nms);
D2T_API int d2t_roi_decode_filter_f32(const float* anchors, const float* offsets, const float* conf, float* boxes,
                              float* scores, int n_anchors, float conf_thresh, void* stream);
D2T_API int d2t_roi_nms_f32(const float* boxes, const long long* order, const float* sorted_scores, float* rois, int* count,
                    int n_anchors, int pre_nms, int max_rois, float iou_thresh, void* ws, size_t ws_bytes, void* stream);

/* ---- integer bin edges (parity instrumentation) ----------------------------
 * edges : (R, r_hw, 4) int32 = (I0, I1, J0, J1) of row-bin / column-bin b,
 * computed on the device by the same code the pooling kernels use.
 * clamp_start != 0 -> ROIPool rule (roipool_cuda.cu:38-50), 0 -> PSROIPool
 * rule (ps_roipool_cuda.cu:42-54).  Must be bit-exact against the reference.
 */
D2T_API int d2t_pool_bins_f32(const float* rois, int32_t* edges, int R, int H, int W, int r_hw, int clamp_start,
                      void* stream);
D2T_API int d2t_pool_bins_f64(const double* rois, int32_t* edges, int R, int H, int W, int r_hw, int clamp_start,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D2T_B200_H */
