/*
 * ebos.h -- C-ABI of libebos.so: B200 (sm_100a) kernels for the contrast-maximisation hot path of
 * tub-rip/event_based_bos (warp -> image of warped events -> cost -> analytic backward -> Adam).
 *
 * The reference is pure Python and has no FFI; the interface each entry point replaces is the
 * Python method cited beside it (paths relative to the reference tree).  The ctypes binding a
 * maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions (all entry points):
 *  - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *    the caller owns every buffer, the library allocates nothing persistent;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *    stream-ordered, nothing synchronises unless documented;
 *  - return value: EBOS_OK (0) or a negative error code; ebos_last_error() gives the text
 *    (thread-local);  no exceptions, no aborts;
 *  - `dtype`: EBOS_F32 or EBOS_F64 selects the element type behind `void*` buffers;
 *  - events are [N,4] array-of-structs rows (x = row, y = col, t, p) exactly as the reference
 *    passes them; images are row-major [H,W]; flow is [2,H,W] (channel 0 = row flow).
 *  - arithmetic on coordinates is unfused IEEE (mul then sub, true division) so that pixel
 *    indices and in-bounds masks are bit-identical to the reference's torch ops.
 */
#ifndef EBOS_H_
#define EBOS_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define EBOS_API __attribute__((visibility("default")))
#else
#define EBOS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EBOS_VERSION 200

/* error codes */
#define EBOS_OK 0
#define EBOS_ERR_BAD_ARG (-1)
#define EBOS_ERR_CUDA (-2)
#define EBOS_ERR_WORKSPACE (-3)
#define EBOS_ERR_UNSUPPORTED (-4)

/* dtypes */
#define EBOS_F32 0
#define EBOS_F64 1

/* reference-time selection: Warp.calculate_reftime, src/warp.py:230-262 */
#define EBOS_DIR_FIRST 0  /* min t */
#define EBOS_DIR_LAST 1   /* max t */
#define EBOS_DIR_FRAC 2   /* min t + frac * (max t - min t); 'middle'=0.5, 'before'=-1, 'after'=2 */

/* data objectives (not present upstream) */
/* doubles in the `acc` scratch of the cost / TV / fused entries: [0] sum(IWE) [1] sum(IWE^2), gradient-magnitude
 * sum in [2] + [8..23], TV sum in [3] + [24..39] (per-CTA partial sums are spread over 16 slots: same-address
 * atomics serialise in L2).  Read the totals through ebos_loss_finalize. */
#define EBOS_ACC_DOUBLES 40

#define EBOS_COST_NONE 0
#define EBOS_COST_VARIANCE 1  /* L = -var(IWE), unbiased */
#define EBOS_COST_GRADMAG 2   /* L = -mean((Sobel_x/8)^2 + (Sobel_y/8)^2), replicate border */

/* status word bits written by the kernels into the caller's `status` int32 (device) */
#define EBOS_STATUS_PIXEL_OOB 1 /* an event's integer pixel is outside the flow grid (reference: gather raises) */
#define EBOS_STATUS_PACKED 2    /* ebos_window_prepare stored the window in the packed (row,col,dt) layout */

/* `flags` of the per-iteration window calls */
#define EBOS_WIN_HAS_WEIGHT 1 /* the window was prepared with per-event weights */
#define EBOS_WIN_PACKED 2     /* the window is in the packed layout (status had EBOS_STATUS_PACKED) */

EBOS_API int ebos_version(void);
EBOS_API const char* ebos_last_error(void);
/* 1 when a CUDA device is usable from this process, else 0 (never fails). */
EBOS_API int ebos_device_available(void);

/* ------------------------------------------------------------------------------------------
 * Operator level (drop-in for the individual reference operators)
 * ---------------------------------------------------------------------------------------- */

/* min/max of the timestamp column.  Replaces nt_min/nt_max over events[...,2]
 * (src/warp.py:218,245-252; src/types/__init__.py:20-47).
 * out_min_max: [batch,2] of dtype.  events: [batch, n, 4]. */
EBOS_API int ebos_time_stats(const void* events, int64_t n, int batch, int dtype, void* out_min_max, void* stream);

/* Warp.warp_event(events, flow, "dense-flow", direction) -- src/warp.py:193-228, 264-288, 330-342.
 * tstats: [batch,2] (min,max) from ebos_time_stats.  dt = t - t_ref; if normalize_t:
 * dt /= (max(dt)-min(dt)).  k = trunc(x)*W + trunc(y);  x' = x - dt*flow0[k];  y' = y - dt*flow1[k].
 * warped: [batch,n,4] = (x', y', dt, p).  status: int32[1], OR-ed with EBOS_STATUS_PIXEL_OOB
 * (such events are passed through un-warped; the Python layer raises like torch.gather does).
 * flow_batch_stride: elements between consecutive batch flows (0 = shared flow). */
EBOS_API int ebos_warp_dense_flow(const void* events, int64_t n, int batch, const void* flow, int64_t flow_batch_stride,
                         int H, int W, const void* tstats, int direction, double direction_frac,
                         int normalize_t, int dtype, void* warped, int32_t* status, void* stream);

/* Backward of ebos_warp_dense_flow w.r.t. flow: dflow[c][k_i] += -dt_i * grad_warped[i][c], c=0,1.
 * dflow [batch or 1, 2,H,W] must be zeroed (or hold a running sum) by the caller. */
EBOS_API int ebos_warp_dense_flow_bwd(const void* events, int64_t n, int batch, int H, int W, const void* tstats,
                             int direction, double direction_frac, int normalize_t, int dtype,
                             const void* grad_warped, void* dflow, int64_t dflow_batch_stride, void* stream);

/* Warp.warp_event(events, theta, "2d-translation") -- src/warp.py:344-383:  x' = x + dt*theta0. */
EBOS_API int ebos_warp_2dof(const void* events, int64_t n, int batch, const void* theta, const void* tstats,
                   int direction, double direction_frac, int normalize_t, int dtype, void* warped, void* stream);

/* EventImageConverter.bilinear_vote_tensor -- src/event_image_converter.py:562-620.
 * events: [batch,n,4] (only x,y used).  image: [batch, Hp, Wp] with Hp = H+2*pad_h, Wp = W+2*pad_w
 * (the CALLER passes the padded size, as EventImageConverter.image_size holds it); fully overwritten.
 * weight: NULL (1.0) or [batch,n].
 * floor_bias: the bias added before floor(): 1e-6 for the tensor branch (src/event_image_converter.py:586),
 * 1e-8 for the numpy branch (src/event_image_converter.py:528).
 * mode 0 = atomic (fp add order unspecified);  mode 1 = deterministic: bit-identical to the
 * reference's sequential tap-major scatter_add_ (needs workspace, see ebos_splat_workspace_bytes).
 * dbg_idx (int64 [batch,4n]) / dbg_mask (uint8 [batch,4n]): optional, the reference's `inds`/`inds_mask`
 * (src/event_image_converter.py:592-617) for bit-exact checking; NULL to skip. */
EBOS_API int ebos_iwe_splat(const void* events, int64_t n, int batch, int Hp, int Wp, int pad_h, int pad_w,
                   const void* weight, double floor_bias, int dtype, int mode, void* image, int64_t* dbg_idx, uint8_t* dbg_mask,
                   void* workspace, size_t workspace_bytes, void* stream);
EBOS_API size_t ebos_splat_workspace_bytes(int64_t n, int batch, int Hp, int Wp, int dtype, int mode);

/* Backward of ebos_iwe_splat: grad_events[i] = (dL/dx', dL/dy', 0, 0) and (optional) grad_weight[i],
 * from grad_image [batch,Hp,Wp].  What autograd derives for src/event_image_converter.py:586-619. */
EBOS_API int ebos_iwe_splat_bwd(const void* events, int64_t n, int batch, int Hp, int Wp, int pad_h, int pad_w,
                       const void* weight, double floor_bias, int dtype, const void* grad_image, void* grad_events,
                       void* grad_weight, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused contrast-maximisation path: events of one window are prepared ONCE (sorted by origin
 * pixel, time-normalised), then every solver iteration runs
 *    splat (warp+vote fused, warped events never materialised) -> cost -> backward -> [Adam].
 * dtype EBOS_F32 is the fast path; EBOS_F64 is the dtype the reference's solvers run in
 * (src/solver/patch_eklt_pyramid2.py:253) and is used for solve-level parity.  All planes, the flow,
 * the events and the loss share `dtype`; `acc` is always double[EBOS_ACC_DOUBLES].
 * ---------------------------------------------------------------------------------------- */

/* Bytes of the caller-owned window buffer / of the temporary workspace used by prepare. */
EBOS_API size_t ebos_window_bytes(int64_t n, int H, int W, int dtype);
EBOS_API size_t ebos_window_workspace_bytes(int64_t n, int H, int W);

/* Build a window from raw events [n,4]: computes t_ref/period on the device
 * (src/warp.py:230-288), dt_i, k_i, and stores (x, y, dt[, weight]) sorted by (32x32 tile of the origin
 * pixel, pixel inside the tile) -- stable, so the sensor's time order is kept inside a pixel -- plus the
 * per-tile offsets and work items used by the shared-memory tile kernels.  weight: NULL or [n].  status as above.
 * tminmax: NULL, or [2] (device, dtype) = (min t, max t) to use instead of this call's own
 * reduction -- for a window whose events are sharded over several GPUs (global min/max).
 * allow_packed: when non-zero (fp32 only) and every event is valid with integer coordinates below
 * 65536 -- raw sensor events -- the window is stored as (row<<16|col, dt) = 8 bytes per event instead
 * of 12, and EBOS_STATUS_PACKED is OR-ed into `status`; the caller must then pass EBOS_WIN_PACKED in
 * the `flags` of the per-iteration calls (read `status` once after prepare).
 * window must be 256-byte aligned. */
EBOS_API int ebos_window_prepare(const void* events, int64_t n, int H, int W, int direction, double direction_frac,
                        int normalize_t, const void* weight, const void* tminmax, int allow_packed, int dtype,
                        void* window, void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* Copy the window's event permutation (int32[n]: sorted position -> original event index) and its
 * time statistics (double[4]: t_ref, period, t_min, t_max) out of the opaque buffer (either may be NULL). */
EBOS_API int ebos_window_info(const void* window, int64_t n, int H, int W, int dtype, int32_t* perm_out,
                     double* tinfo_out, void* stream);

/* Fused warp + bilinear vote of a prepared window into iwe [Hp,Wp] (fully overwritten).
 * `n`, `flags` (EBOS_WIN_*) and `dtype` must describe the window as it was prepared. */
EBOS_API int ebos_window_splat(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                      int pad_w, int dtype, void* iwe, void* stream);

/* Data objective on the IWE: value and gradient scaled by `scale`.
 * acc: double[EBOS_ACC_DOUBLES] device scratch (the data-cost entries are zeroed by this call).  grad_iwe: [Hp,Wp], required for
 * GRADMAG; for VARIANCE it may be NULL -- the backward then derives dL/dIWE = c*(IWE-mean) on the
 * fly from `acc` (pass the same `acc` and `iwe` to ebos_window_backward). */
EBOS_API int ebos_iwe_cost(int kind, const void* iwe, int Hp, int Wp, int omit_boundary, double scale, int dtype,
                  double* acc, void* grad_iwe, void* stream);

/* ImageGradient.calculate_torch -- src/costs/image_gradient.py:60-75 (TV-L1 of the flow with
 * torch.gradient semantics) value and gradient:  dflow = tv_scale * dTV/dflow  (OVERWRITES dflow,
 * so it doubles as the zero-fill of the gradient buffer; tv_scale = 0 just zeroes).
 * weights: NULL (1.0) or [H,W].  the TV entries of acc (zeroed by this call) accumulate the un-normalised |.| sum. */
EBOS_API int ebos_flow_tv(const void* flow, const void* weights, int H, int W, double tv_scale, int dtype, double* acc,
                 void* dflow, void* stream);

/* Analytic backward of the fused splat: re-warps every event, gathers dL/dIWE at
 * its four taps and accumulates -dt*dL/dx' into dflow[:, k] (ACCUMULATES; run ebos_flow_tv first).
 * grad_iwe: [Hp,Wp] or NULL with kind == EBOS_COST_VARIANCE (then iwe+acc+scale are used). */
EBOS_API int ebos_window_backward(const void* window, int64_t n, int flags, const void* flow, int H, int W,
                         int pad_h, int pad_w, int dtype, const void* grad_iwe, int kind, const void* iwe,
                         const double* acc, int omit_boundary, double scale, void* dflow, void* stream);

/* loss[0] = data_scale * L_data + tv_scale * TV  from `acc` (device scalar of dtype). */
EBOS_API int ebos_loss_finalize(int kind, const double* acc, int Hp, int Wp, int H, int W, int omit_boundary,
                       double data_scale, double tv_scale, int dtype, void* loss, void* stream);

/* One complete objective evaluation: [TV | splat] -> cost -> [backward | loss], stream-ordered, no host sync
 * (CUDA-graph capturable; the TV kernel and the scalar-loss kernel run on an internal auxiliary stream that is
 * forked from and joined back into `stream`).  iwe [Hp,Wp], grad_iwe [Hp,Wp] (scratch), dflow [2,H,W], loss [1],
 * acc double[EBOS_ACC_DOUBLES].  When `iwe` starts exactly at acc + EBOS_ACC_DOUBLES (one allocation, accumulators
 * first) both are zeroed by a single memset node.
 * clean_workspace != 0: the caller guarantees that acc and iwe are ALL ZERO on entry and receives them all zero again
 * (the IWE of this evaluation is then not available afterwards): the zero-fill for the next evaluation runs
 * concurrently with the backward instead of in front of the splat.  clean_workspace == 0: nothing is assumed, the IWE
 * of this evaluation is left in `iwe`.
 * blur_sigma > 0: the objective is evaluated on gaussian_blur(IWE, kernel_size=3, sigma) -- the `iwe.blur_sigma` of the
 * solver configs, src/event_image_converter.py:399-404 -- and back-propagated through the blur's exact adjoint;
 * blur_plane [Hp,Wp] (scratch, required then) holds the blurred IWE and afterwards dL/dIWE. */
EBOS_API int ebos_cmax_value_and_grad(const void* window, int64_t n, int flags, const void* flow, int H, int W,
                             int pad_h, int pad_w, int kind, int omit_boundary, double data_scale, double tv_scale,
                             const void* tv_weights, int dtype, void* iwe, void* grad_iwe, void* dflow, void* loss,
                             double* acc, int clean_workspace, double blur_sigma, void* blur_plane, void* stream);

/* One complete SOLVER iteration (src/solver/patch_eklt_pyramid2.py:267-285: zero_grad / loss / backward / step):
 * ebos_cmax_value_and_grad followed by the Adam update of `flow`, as six graph nodes
 *   [TV + step counter | splat] [cost] [backward | IWE memset] [Adam + loss + accumulator reset].
 * `acc` (double[EBOS_ACC_DOUBLES]) AND `iwe` must be ZERO on entry and are left zero on exit (the IWE is zero-filled for
 * the next iteration concurrently with the backward); `step_dev` (int32[1]) counts the
 * iterations done (0 before the first call) and is advanced by the call; `loss` receives this iteration's
 * objective value (before the update).  Capture once in a CUDA graph, replay n_iter times. */
EBOS_API int ebos_cmax_adam_iteration(const void* window, int64_t n, int flags, void* flow, int H, int W, int pad_h,
                             int pad_w, int kind, int omit_boundary, double data_scale, double tv_scale,
                             const void* tv_weights, int dtype, void* iwe, void* grad_iwe, void* dflow, void* loss,
                             double* acc, void* exp_avg, void* exp_avg_sq, double lr, double beta1, double beta2,
                             double eps, int32_t* step_dev, double blur_sigma, void* blur_plane, void* stream);

/* The same solver iteration with the TV term folded into the Adam kernel (fp32, unit TV weights): five graph nodes
 *   [splat] [cost] [backward | IWE memset] [Adam + lambda dTV + loss + accumulator reset + step counter].
 * The Adam kernel evaluates the TV stencil (src/costs/image_gradient.py:60-75) on `flow_in` itself, adds it to the data
 * gradient the backward accumulated in `dflow`, and writes the updated flow to `flow_out` (a second [2,H,W] plane, must
 * not alias flow_in: neighbouring pixels still read the old values).  Contract on top of ebos_cmax_adam_iteration's:
 * `dflow` must be ZERO on entry and is left zero on exit.  Callers alternate the two flow planes between iterations.
 * Returns EBOS_ERR_UNSUPPORTED (nothing enqueued) unless dtype is fp32, W % 4 == 0, W >= 12, H >= 5 and the planes are
 * 16-byte aligned -- use ebos_cmax_adam_iteration then. */
EBOS_API int ebos_cmax_adam_iteration_fused_tv(const void* window, int64_t n, int flags, const void* flow_in, void* flow_out,
                             int H, int W, int pad_h, int pad_w, int kind, int omit_boundary, double data_scale,
                             double tv_scale, int dtype, void* iwe, void* grad_iwe, void* dflow, void* loss, double* acc,
                             void* exp_avg, void* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                             int32_t* step_dev, double blur_sigma, void* blur_plane, void* stream);

/* torch.optim.Adam step (src/solver/patch_eklt_pyramid2.py:262-264,284), in place.  `step` is
 * 1-based.  Elementwise over n values. */
EBOS_API int ebos_adam_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, double lr,
                   double beta1, double beta2, double eps, int step, int dtype, void* stream);

/* Same, with the step counter read from device memory (`step_dev` int32[1], incremented by the
 * call) so that a captured CUDA graph can be replayed for every iteration. */
EBOS_API int ebos_adam_step_graph(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, double lr,
                         double beta1, double beta2, double eps, int32_t* step_dev, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Event-sharded windows: one-shot reductions over peer memory.  `peers` is a HOST array of n_peers
 * (<= 8) device pointers, one per rank in rank order, each valid in this process (the own buffer, and the other
 * ranks' buffers mapped over NVLink -- e.g. torch symmetric memory); the caller provides the cross-rank barriers
 * before (all partial planes complete) and after (nobody overwrites a plane a peer still reads).
 *   ebos_iwe_cost_peers  = ebos_iwe_cost(EBOS_COST_GRADMAG) on the SUM of the partial IWE planes, summed in rank
 *                          order while the tiles are loaded: the all-reduce of the partial IWEs fused into the cost
 *                          kernel; bit-identical on every rank.
 *   ebos_sum_peers       : out[i] = sum_r peers[r][i] (rank order) -- the partial flow gradients. */
EBOS_API int ebos_iwe_cost_peers(int kind, const void* const* iwe_peers, int n_peers, int Hp, int Wp, int omit_boundary,
                        double scale, int dtype, double* acc, void* grad_iwe, void* stream);
EBOS_API int ebos_sum_peers(const void* const* peers, int n_peers, int64_t n, int dtype, void* out, void* stream);
/* Two-shot form for larger rank counts (the one-shot pass reads n_peers whole planes per rank):
 *   ebos_reduce_peers_slice   dst[i] = sum_r peers[r][i] for i in [begin, end) (rank order).  dst may be peers[own rank]
 *                             (in place: rank r reduces slice r into its own buffer, which only rank r reads there);
 *   ebos_gather_peers_slices  out[i] = peers[min(i / slice, n_peers - 1)][i]: collect the reduced slices (after a barrier).
 * Every rank ends with bit-identical values (one rank computes each element). */
/* In-switch form (NVLS multicast): multicast_base is the MULTICAST mapping of a symmetric fp32 buffer (one address for the
 * same offset in every rank's copy, e.g. torch symmetric memory's `multicast_ptr`).  For i in [begin, end): the NVSwitch
 * sums the ranks' copies of element i (multimem.ld_reduce) and the result is written into every copy (multimem.st).  Rank r
 * calls it for slice r between the same two barriers as above: afterwards every rank's own buffer holds the reduced plane,
 * bit-identical everywhere.  begin / end multiples of 4 elements, 16-byte aligned base; EBOS_ERR_UNSUPPORTED otherwise. */
EBOS_API int ebos_multimem_allreduce_slice(void* multicast_base, int64_t begin, int64_t end, int dtype, void* stream);
EBOS_API int ebos_reduce_peers_slice(const void* const* peers, int n_peers, int64_t begin, int64_t end, int dtype, void* dst,
                            void* stream);
EBOS_API int ebos_gather_peers_slices(const void* const* peers, int n_peers, int64_t n, int64_t slice, int dtype, void* out,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * Event ingestion: raw sensor arrays (x:int16 sensor column, y:int16 sensor row, t:int32 us,
 * p:bool as uint8; 9 B/event, time ordered -- the `raw_events` datasets of src/data_loader/ccs.py:50-69) ->
 * the reference's event rows [n,4] = (row = y, col = x, t_us / 1e6, p) of CcsDataLoader.load_event
 * (src/data_loader/ccs.py:288-296), optionally through the CROP filter row0 <= row < row1, col0 <= col < col1
 * (order preserved, coordinates not shifted: src/utils/event_utils.py:109-129, src/solver/base.py:123-139).
 * All pointers are device pointers, already offset to the first event of the window; n < 2^31.
 * rebase != 0: t = (t_us - t_origin_us) / 1e6 (integer subtraction first; required for useful fp32 rows);
 * rebase == 0 with EBOS_F64 reproduces the loader's rows bit for bit.
 * events_out: [n,4] of dtype (16-byte aligned; the first *n_kept rows are written), n_kept: device int64.
 * workspace: >= ebos_ingest_workspace_bytes(n) bytes (only needed when crop != 0). */
EBOS_API size_t ebos_ingest_workspace_bytes(int64_t n);
EBOS_API int ebos_ingest_raw(const int16_t* x, const int16_t* y, const int32_t* t_us, const uint8_t* p, int64_t n, int crop,
                    int row0, int row1, int col0, int col1, int64_t t_origin_us, int rebase, int dtype, void* events_out,
                    int64_t* n_kept, void* workspace, size_t workspace_bytes, void* stream);

/* CcsDataLoader.time_to_index (src/data_loader/ccs.py:345-357): *index_out = searchsorted(t_us / 1e6, time) - 1
 * (left side, float64 comparison like the loader's `_time_cache`); index_out: device int64. */
EBOS_API int ebos_time_to_index(const int32_t* t_us, int64_t n, double time, int64_t* index_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * EKLT inner loop of PatchEkltPyramid2 (what configs/hot_plate1.yaml runs: poisson_model +
 * optimize_warp with the costs diff_norm + image_gradient + flow_norm_pxy).  One pyramid level optimises
 * theta [3,ph,pw] = (intensity, p_row, p_col) on the patch grid ph = ceil(H/patch), pw = ceil(W/patch):
 *
 *   pf = Sobel(theta[0]) / 8                       poisson_to_flow, src/solver/patch_eklt_dependent.py:259-281
 *   f = up(pf), t = up(theta[1:3])                 interpolate_dense_flow_from_patch_tensor, src/solver/patch_eklt.py:173-204
 *                                                  (replicate pad 1, bilinear x patch, centre crop)
 *   q = f0 * warp(grad_x, t) + f1 * warp(grad_y, t)   warp_image_forward, src/utils/frame_utils.py:56-89 (grid_sample,
 *                                                  align_corners=True, zeros; base grid built in float32 as upstream)
 *   pred = q / (||q||_F + 1e-4) * M                _make_prediction_torch, src/solver/patch_eklt_pyramid2.py:345-365
 *   L = w_data * max_j sum_i |pred - measured|_ij  DifferenceNorm: torch.linalg.norm(ord=1) of a MATRIX, src/costs/diff_norm.py:52
 *     + w_tv * mean(|d_row(f M)| w_inv + |d_col(f M)| w_inv)        ImageGradient, src/costs/image_gradient.py:60-75
 *     + w_pxy * mean_ij sqrt((t0 M)^2 + (t1 M)^2)                   FlowNormPxy, src/costs/flow_norm.py:52
 *
 * replaces PatchEkltPyramid2._objective_scipy + loss.backward() (src/solver/patch_eklt_pyramid2.py:267-285, 368-392).
 * M is the ROI rectangle rows [roi_x0,roi_x1) x cols [roi_y0,roi_y1) (estimate_mask_dense, :50-51).
 * grad_x / grad_y: [H,W] frame gradients (row / column derivative; cv2.Sobel of _set_frame,
 * src/solver/generative_max_likelihood.py:194-213); measured: [H,W] normalised event histogram ALREADY multiplied by M;
 * weight_inverse: [H,W].  All of `dtype`, device pointers.  loss: [1], grad: like theta (fully overwritten).
 * flags (generative_ml.* of the yaml; hot_plate1 = POISSON | WARP) select the reference's other objectives:
 *   EBOS_EKLT_POISSON      theta[0] is an intensity and the patch flow its Sobel/8; without it theta[0:2] IS the patch
 *                          flow (`_get_patch_flow`, src/solver/patch_eklt_pyramid2.py:290-303)
 *   EBOS_EKLT_WARP         the last two channels of theta translate the frame gradients and carry the pxy term; without
 *                          it the gradients are used as they are and w_pxy is ignored (:354-357)
 *   EBOS_EKLT_NO_POLARITY  q = |q| (:360-361)
 * theta: [(POISSON ? 1 : 2) + (WARP ? 2 : 0), ph, pw].  weights: NULL or [H,W] event-histogram weights multiplied into q
 * before the normalisation (weight_loss_by_event_hist, :363-364).
 * workspace: >= ebos_eklt_workspace_bytes(...) bytes, 256-byte aligned, contents irrelevant on entry.
 * Stream-ordered, CUDA-graph capturable (no host synchronisation, no allocation). */
#define EBOS_EKLT_POISSON 1
#define EBOS_EKLT_WARP 2
#define EBOS_EKLT_NO_POLARITY 4
EBOS_API size_t ebos_eklt_workspace_bytes(int H, int W, int ph, int pw, int patch, int dtype);
EBOS_API int ebos_eklt_value_and_grad(const void* theta, int flags, const void* grad_x, const void* grad_y,
                             const void* measured, const void* weight_inverse, const void* weights, int H, int W, int ph,
                             int pw, int patch, int roi_x0, int roi_x1, int roi_y0, int roi_y1, double w_data, double w_tv,
                             double w_pxy, int dtype, void* workspace, size_t workspace_bytes, void* loss, void* grad,
                             void* stream);

/* One solver iteration of a level (src/solver/patch_eklt_pyramid2.py:265-285: zero_grad / loss / backward / Adam step):
 * ebos_eklt_value_and_grad followed by ebos_adam_step_graph on all of theta.  `step_dev` (int32[1]) counts the
 * iterations done and is advanced by the call; `loss` receives the objective BEFORE the update. */
EBOS_API int ebos_eklt_adam_iteration(void* theta, int flags, const void* grad_x, const void* grad_y, const void* measured,
                             const void* weight_inverse, const void* weights, int H, int W, int ph, int pw, int patch,
                             int roi_x0, int roi_x1, int roi_y0, int roi_y1, double w_data, double w_tv, double w_pxy,
                             int dtype, void* workspace, size_t workspace_bytes, void* loss, void* grad, void* exp_avg,
                             void* exp_avg_sq, double lr, double beta1, double beta2, double eps, int32_t* step_dev,
                             void* stream);

/* interpolate_dense_flow_from_patch_tensor on its own (src/solver/patch_eklt.py:173-204): patch_values [channels,ph,pw]
 * -> dense [channels,H,W]; and poisson_to_flow (src/solver/patch_eklt_dependent.py:259-281): intensity [ph,pw] ->
 * patch_flow [2,ph,pw] = Sobel/8 with replicate padding. */
EBOS_API int ebos_eklt_upsample(const void* patch_values, int channels, int H, int W, int ph, int pw, int patch, int dtype,
                       void* dense, void* stream);
EBOS_API int ebos_eklt_patch_flow(const void* intensity, int ph, int pw, int dtype, void* patch_flow, void* stream);

/* Separable 2-D correlation with mirrored borders -- the per-window preprocessing of the EKLT solvers:
 *   cv2.Sobel(frame, CV_64F, 0|1, 1|0, ksize=3)        _set_frame, src/solver/generative_max_likelihood.py:194-213
 *   cv2.GaussianBlur(hist, None, sigma)                calculate_iwe_cache, src/solver/patch_eklt.py:271-293
 *   scipy.ndimage.gaussian_filter(|hist|, 10)          weight_inverse, src/solver/patch_eklt.py:295-304
 * out[i,j] = sum_u sum_v taps_rows[u] * taps_cols[v] * image[b(i+u-ru), b(j+v-rv)], columns pass first.
 * taps_*: HOST arrays (odd length <= 127); border 0 = reflect-101 (cv2 default), 1 = reflect (scipy 'reflect').
 * image, tmp, out: [H,W] of dtype on the device (tmp is scratch; out may not alias image or tmp). */
EBOS_API int ebos_sepconv2d(const void* image, int H, int W, const double* taps_rows, int n_rows, const double* taps_cols,
                   int n_cols, int border, int dtype, void* tmp, void* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Either side of the path
 * ---------------------------------------------------------------------------------------- */

/* calculate_flow_error_numpy / calculate_flow_error_tensor (src/utils/flow_utils.py:769-821, 705-766; call site
 * SolverBase.calculate_flow_error, src/solver/base.py:289-317): flow_gt, flow_pred [batch,2,H,W] of dtype; event_mask NULL or uint8 (bool)
 * [batch,H,W] (mask_batched != 0) / [H,W] shared by the batch (mask_batched == 0) -- the squeezed [b,1,H,W] mask of
 * EventImageConverter.create_eventmask.  A pixel counts when the ground truth is not +-inf and non-zero in BOTH
 * channels and the mask is set; like upstream the flows are MULTIPLIED by that mask (an infinite value at a
 * masked-out pixel therefore gives NaN metrics, as it does upstream).  time_scale: NULL, or [batch] of dtype
 * multiplied into both masked flows (the tensor variant's optional argument); n_points_fp32 != 0 rounds the
 * denominator n_points + 1e-5 to float32 the way the tensor variant does (an int64 tensor plus a Python float is a
 * float32 tensor in torch, :741).  errors: device double[8] =
 * EPE, 1PE, 2PE, 3PE, 5PE, 10PE, 20PE, AE -- each sum / (n_points + 1e-5) per batch row, then the batch mean.
 * workspace: device double[ebos_flow_error_workspace_doubles(batch)] (zeroed by the call).  Sums in double. */
EBOS_API size_t ebos_flow_error_workspace_doubles(int batch);
EBOS_API int ebos_flow_error(const void* flow_gt, const void* flow_pred, const uint8_t* event_mask, int mask_batched,
                    const void* time_scale, int n_points_fp32, int batch, int H, int W, int dtype, double* workspace,
                    double* errors, void* stream);

/* torchvision gaussian_blur(image[b,1,H,W], kernel_size=3, sigma) as EventImageConverter.create_image_from_events_tensor
 * applies it to the IWE (src/event_image_converter.py:399-404): kernel2d = k k^T with k = exp(-x^2 / (2 sigma^2)) /
 * sum over x = -1, 0, 1 (computed in dtype), reflect padding.  adjoint != 0 applies the TRANSPOSED operator (what
 * autograd's backward of that call computes; the reflect border makes the operator non-symmetric).
 * image, out: [batch,H,W] of dtype, out must not alias image; H, W >= 2. */
EBOS_API int ebos_blur3(const void* image, int batch, int H, int W, double sigma, int adjoint, int dtype, void* out,
               void* stream);

/* Launch accounting (bench.py's `gpu_launches`): capture what the calls made between the two functions enqueue on
 * `stream` (cudaStreamBeginCapture / EndCapture, relaxed mode) and count the graph's nodes -- kernel nodes (this
 * library's kernels) and other nodes (memsets).  Nothing is executed; the graph is destroyed.  host pointers. */
EBOS_API int ebos_capture_begin(void* stream);
EBOS_API int ebos_capture_end_count(void* stream, int32_t* n_kernel_nodes, int32_t* n_other_nodes);

/* Replayable launch sequences for the solver loops (the Adam loop idiom of src/solver/patch_eklt_pyramid2.py:259-288 runs
 * the same launches n_iter times): ebos_capture_begin(stream) ... the calls of `unroll` iterations ... then
 * ebos_capture_end_exec turns what was captured into an executable graph.  *exec_inout is a SLOT: NULL on first use; on
 * later calls the existing executable is updated in place to the new window's pointers and sizes
 * (cudaGraphExecUpdate; *updated_in_place = 1) or, if the launch sequence changed shape, rebuilt (= 0).  No device
 * synchronisation in the update path.  ebos_exec_launch enqueues one replay on `stream`; ebos_exec_destroy frees the slot
 * (may wait for work in flight on the device).  exec_inout / updated_in_place are host pointers. */
EBOS_API int ebos_capture_end_exec(void* stream, void** exec_inout, int32_t* updated_in_place);
EBOS_API int ebos_exec_launch(void* exec, void* stream);
EBOS_API int ebos_exec_destroy(void* exec);

#ifdef __cplusplus
}
#endif
#endif /* EBOS_H_ */
