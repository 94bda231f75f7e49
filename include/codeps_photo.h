/*
 * codeps_photo.h -- C ABI of libcodeps_photo.so: the CoDEPS photometric reprojection loss hot
 * path as hand-written CUDA for sm_100a (B200).
 *
 * The reference (robot-learning-freiburg/CoDEPS) is pure Python/PyTorch and has no FFI; each
 * entry point below replaces a Python class on the hot path and cites it (paths relative to the
 * reference root).  A maintainer binds these with ctypes (INTEGRATION.md shows the stub);
 * codeps_b200/_native.py is exactly that binding.
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch types.  All tensors fp32, NCHW, contiguous, on the
 *     current CUDA device, unless a parameter says HOST.
 *   - The caller owns every buffer, including outputs, scratch and saved-for-backward state,
 *     sized with the *_bytes() helpers.  The library never allocates or frees device memory and
 *     keeps no pointer after a call returns.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation.
 *   - Every function returns CDP_OK (0) or a negative cdp_status; cdp_last_error() returns a
 *     thread-local message for the most recent failure on the calling thread.
 *   - Reductions are fixed-order (no atomics): results are bit-identical run to run.
 *   - There is no CPU implementation behind this ABI.
 */
#ifndef CODEPS_PHOTO_H_
#define CODEPS_PHOTO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDP_ABI_VERSION 6
#define CDP_MAX_LEVELS 6           /* pyramid levels per call (reference uses 5) */
#define CDP_MAX_BATCH_PER_LAUNCH 32 /* intrinsics travel in kernel-parameter (constant) space */

typedef enum cdp_status {
  CDP_OK = 0,
  CDP_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, ...) */
  CDP_ERR_UNSUPPORTED = -2, /* valid request the library does not implement */
  CDP_ERR_CUDA = -3,        /* a CUDA runtime call or launch failed */
  CDP_ERR_WORKSPACE = -4    /* caller-provided buffer too small */
} cdp_status;

typedef void* cdp_stream_t; /* cudaStream_t */

int cdp_version(void);
const char* cdp_last_error(void);
/* CDP_OK when the current device can run the sm_100a kernels. */
int cdp_device_check(void);

/* ---------------------------------------------------------------------------------------------
 * Optional per-kernel timing for benchmarks: while enabled, every kernel of the fused loss is
 * bracketed by CUDA events on its launch stream.  cdp_profile_read synchronises on the recorded
 * events and returns the summed duration and launch count of one kernel since the last
 * cdp_profile_enable().  Do not enable during CUDA-graph capture.
 * ------------------------------------------------------------------------------------------- */
enum cdp_kernel_id {
  CDP_KERNEL_PYRAMID = 0, CDP_KERNEL_PHOTO = 1, CDP_KERNEL_FINALIZE = 2, CDP_KERNEL_DEPTH_GRAD = 3,
  CDP_KERNEL_SMOOTH_SUM = 4, CDP_KERNEL_SMOOTH_MAIN = 5, CDP_KERNEL_SMOOTH_FINALIZE = 6,
  CDP_KERNEL_SMOOTH_BWD = 7, CDP_KERNEL_COUNT = 8
};
int cdp_profile_enable(int32_t enable);
int cdp_profile_read(int32_t kernel_id, double* total_ms, int32_t* launches);

/* ---------------------------------------------------------------------------------------------
 * Bilinear resize tables (Interpolate / F.interpolate(mode="bilinear", align_corners=False),
 * algos/depth.py:158-173, always from full resolution: algos/depth.py:280-281,295,313).
 * Built once per (H, W, num_levels) on the HOST, uploaded by the caller, shared by forward and
 * backward so that both use identical taps and weights.
 * ------------------------------------------------------------------------------------------- */
size_t cdp_resize_tables_bytes(int32_t height, int32_t width, int32_t num_levels);
int cdp_resize_tables_build(int32_t height, int32_t width, int32_t num_levels, void* host_out,
                            size_t host_bytes);

/* ---------------------------------------------------------------------------------------------
 * ReconstructionLoss.__call__ (algos/depth.py:239-326) on top of ImageWarper.forward
 * (misc/image_warper.py:153-184), CameraModel (misc/camera_model.py:36-71), SSIMLoss
 * (algos/depth.py:128-155) and _compute_loss (algos/depth.py:221-237): multi-scale
 * min-reprojection with identity auto-mask.
 * ------------------------------------------------------------------------------------------- */
/* Fused network-head conversions: the op takes what the heads emit --
 * the sigmoid disparity (DepthHead.disp_to_depth, models/depth_head.py:49-54) and per source frame
 * the 6-DoF pose parameters (PoseHead.transformation_from_parameters, models/pose_head.py:56-137,
 * invert = 1 for the t -> t-1 pose) -- and returns gradients with respect to them.  The depth map
 * and the two 4x4 matrices are still produced (DepthAlgo returns them): in this mode
 * cdp_photo_args.depth / pose0 / pose1 are caller-owned OUTPUT buffers the library fills. */
typedef struct cdp_photo_heads {
  const float* disp;           /* [B,1,H,W] in (0,1) */
  float min_depth, max_depth;  /* reference defaults 0.1, 100 */
  const float* axisangle[2];   /* [B,3] per source frame */
  const float* translation[2]; /* [B,3] per source frame */
  int32_t invert[2];
} cdp_photo_heads;

typedef struct cdp_photo_args {
  int32_t batch, height, width; /* full resolution */
  int32_t num_levels;           /* level s has size (height >> s, width >> s) */
  float alpha;                  /* SSIM weight, reference default 0.85 */
  int32_t with_grad;            /* also produce the state cdp_photo_bwd needs */
  /* HOST [num_levels][batch][4] = fx,fy,cx,cy already rescaled to each level
   * (CameraModel.get_scaled_model_image_size, misc/camera_model.py:36-41). */
  const float* intrinsics_host;
  const float* target;  /* [B,3,H,W] frame t   */
  const float* source0; /* [B,3,H,W] frame t-1 */
  const float* source1; /* [B,3,H,W] frame t+1 */
  const float* depth;   /* [B,1,H,W] */
  const float* pose0;   /* [B,4,4] t -> t-1, 16-byte aligned */
  const float* pose1;   /* [B,4,4] t -> t+1, 16-byte aligned */
  /* Tie-break noise (algos/depth.py:316-318).  noise[s] = [B,2,H_s,W_s] standard normal draws
   * (the library multiplies by 1e-5), or all NULL to use the built-in counter-based generator
   * seeded with noise_seed. */
  const float* noise[CDP_MAX_LEVELS];
  uint64_t noise_seed;
  const void* resize_tables; /* device copy of cdp_resize_tables_build output */
  /* outputs */
  float* loss;                      /* [1] */
  uint8_t* argmin[CDP_MAX_LEVELS];  /* [B,H_s,W_s]: 0/1 reprojection from t-1/t+1, 2/3 identity
                                       (pixel auto-masked); may be NULL per level */
  void* scratch; size_t scratch_bytes; /* pyramid levels; dead after the call */
  void* saved;   size_t saved_bytes;   /* with_grad: state for cdp_photo_bwd */
  /* Optional object-motion maps [B,3,H,W] per source frame (object_motion_maps of
   * ReconstructionLoss.__call__, algos/depth.py:296-303; added to the transformed point,
   * misc/image_warper.py:133-134).  Both or neither. */
  const float* motion0;
  const float* motion1;
  /* Alternative to intrinsics_host (exactly one of the two is set): DEVICE [batch][4] = fx,fy,cx,cy
   * at full resolution, e.g. the batch's "camera_model" tensor as the data loader delivers it
   * (codeps/online_adap.py:95-100).  The kernels rescale per level themselves (same arithmetic as
   * misc/camera_model.py:36-41), so the per-sample CameraModel.from_tensor read-back
   * (misc/camera_model.py:27, one host synchronisation per sample) is not needed. */
  const float* intrinsics_dev;
  /* Optional cudaEvent_t (as void*): if set, `stream` waits on it before the first kernel that
   * reads the noise tensors.  Lets the caller produce the noise on another stream while the pyramid
   * kernel runs (both are memory-bound and independent). */
  void* noise_ready;
  /* Optional: fused head conversions (see cdp_photo_heads); depth / pose0 / pose1 are then outputs. */
  const cdp_photo_heads* heads;
  /* Optional DEVICE counter for the built-in generator (noise all NULL): if set, the kernels use
   * *noise_seed_dev instead of noise_seed and add 1 to it once the call's tile kernel has finished,
   * so that a CUDA graph that captured this call draws fresh noise on every replay (a host-side
   * seed would be frozen into the graph). */
  uint64_t* noise_seed_dev;
} cdp_photo_args;

size_t cdp_photo_scratch_bytes(int32_t batch, int32_t height, int32_t width, int32_t num_levels,
                               int32_t with_motion);
size_t cdp_photo_saved_bytes(int32_t batch, int32_t height, int32_t width, int32_t num_levels,
                             int32_t with_motion);
int cdp_photo_fwd(const cdp_photo_args* args, cdp_stream_t stream);
/* Backward of the above: grad_loss is a DEVICE scalar (dL/d loss).  Writes dL/d depth [B,1,H,W]
 * and dL/dT for both poses [B,4,4] each (autograd of torch.bmm, misc/image_warper.py:129); with
 * motion maps also dL/d motion [B,3,H,W] for both (with_motion must match the forward call). */
int cdp_photo_bwd(int32_t batch, int32_t height, int32_t width, int32_t num_levels,
                  const void* saved, size_t saved_bytes, const void* resize_tables,
                  const float* grad_loss, float* grad_depth, float* grad_pose0, float* grad_pose1,
                  int32_t with_motion, float* grad_motion0, float* grad_motion1, cdp_stream_t stream);
/* Backward of a forward call made with cdp_photo_args.heads: `depth` is the map that call wrote.
 * Writes dL/d disp [B,1,H,W] and dL/d axis-angle / dL/d translation [B,3] per source frame (the
 * chain through disp_to_depth and transformation_from_parameters is applied inside the same
 * kernel that assembles dL/d depth and scales dL/dT). */
int cdp_photo_bwd_heads(int32_t batch, int32_t height, int32_t width, int32_t num_levels,
                        const void* saved, size_t saved_bytes, const void* resize_tables,
                        const float* grad_loss, const cdp_photo_heads* heads, const float* depth,
                        float* grad_disp, float* grad_axisangle0, float* grad_translation0,
                        float* grad_axisangle1, float* grad_translation1,
                        int32_t with_motion, float* grad_motion0, float* grad_motion1, cdp_stream_t stream);
/* Number of kernels one cdp_photo_fwd / cdp_photo_bwd call launches (for launch accounting). */
/* The draws of the built-in tie-break generator (cdp_photo_args.noise all NULL) for one level:
 * out = [B,2,H_s,W_s] standard normal values, exactly what cdp_photo_fwd adds (times 1e-5) to the two
 * identity candidates of that level when called with the same noise_seed.  Test / inspection aid:
 * lets a caller reproduce a noise="fused" evaluation with explicit noise tensors. */
int cdp_tiebreak_noise(int32_t batch, int32_t level_height, int32_t level_width, int32_t level,
                       uint64_t noise_seed, float* out, cdp_stream_t stream);
int cdp_photo_fwd_launches(int32_t batch, int32_t num_levels);
int cdp_photo_bwd_launches(int32_t batch, int32_t num_levels, int32_t with_motion);

/* ---------------------------------------------------------------------------------------------
 * EdgeAwareSmoothnessLoss.__call__ (algos/depth.py:58-107).
 * ------------------------------------------------------------------------------------------- */
size_t cdp_smooth_saved_bytes(int32_t batch, int32_t height, int32_t width);
int cdp_smooth_fwd(const float* image /*[B,3,H,W]*/, const float* disp /*[B,1,H,W]*/,
                   int32_t batch, int32_t height, int32_t width, int32_t with_grad,
                   float* loss /*[1]*/, void* saved, size_t saved_bytes, cdp_stream_t stream);
int cdp_smooth_bwd(const void* saved, size_t saved_bytes, const float* grad_loss /*device [1]*/,
                   int32_t batch, int32_t height, int32_t width, float* grad_disp /*[B,1,H,W]*/,
                   cdp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone operators (same classes used outside the fused loss).
 * intrinsics_host: HOST [batch][4] for this image size.
 * ------------------------------------------------------------------------------------------- */
/* CoordinateWarper.forward (misc/image_warper.py:100-144): normalised grid [B,H,W,2]. */
int cdp_warp_grid_fwd(const float* depth, const float* pose, const float* motion /*nullable*/,
                      const float* intrinsics_host, int32_t batch, int32_t height, int32_t width,
                      float* grid, cdp_stream_t stream);
/* ImageWarper.forward (misc/image_warper.py:153-184); mode 0 = bilinear, 1 = nearest. */
int cdp_warp_image_fwd(const float* src /*[B,C,H,W]*/, int32_t channels, const float* depth,
                       const float* pose, const float* motion /*nullable [B,3,H,W]*/,
                       const float* intrinsics_host, int32_t batch, int32_t height, int32_t width,
                       int32_t mode, float* out /*[B,C,H,W]*/, cdp_stream_t stream);
/* Backward of the bilinear warp w.r.t. depth, pose (and motion when given).
 * partials: device scratch of cdp_warp_bwd_scratch_bytes(). */
size_t cdp_warp_bwd_scratch_bytes(int32_t batch, int32_t height, int32_t width);
int cdp_warp_image_bwd(const float* grad_out /*[B,C,H,W]*/, const float* src, int32_t channels,
                       const float* depth, const float* pose, const float* motion,
                       const float* intrinsics_host, int32_t batch, int32_t height, int32_t width,
                       float* grad_depth /*[B,1,H,W]*/, float* grad_pose /*[B,4,4]*/,
                       float* grad_motion /*nullable [B,3,H,W]*/, void* scratch,
                       size_t scratch_bytes, cdp_stream_t stream);
/* SSIMLoss.__call__ (algos/depth.py:128-155) on `planes` = B*C images of H x W. */
int cdp_ssim_fwd(const float* x, const float* y, int32_t planes, int32_t height, int32_t width,
                 float* out, cdp_stream_t stream);
/* scratch: 4 * planes*H*W floats. */
size_t cdp_ssim_bwd_scratch_bytes(int32_t planes, int32_t height, int32_t width);
int cdp_ssim_bwd(const float* grad_out, const float* x, const float* y, int32_t planes,
                 int32_t height, int32_t width, float* grad_x /*nullable*/,
                 float* grad_y /*nullable*/, void* scratch, size_t scratch_bytes,
                 cdp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The two conversions that feed the loss.
 * ------------------------------------------------------------------------------------------- */
/* PoseHead.transformation_from_parameters (models/pose_head.py:56-137): axis-angle [B,3] and
 * translation [B,3] -> T [B,4,4]; invert != 0 gives R^T T(-t) (the t -> t-1 pose). */
int cdp_pose_fwd(const float* axisangle, const float* translation, int32_t batch, int32_t invert,
                 float* T, cdp_stream_t stream);
int cdp_pose_bwd(const float* grad_T /*[B,4,4]*/, const float* axisangle, const float* translation,
                 int32_t batch, int32_t invert, float* grad_axisangle /*[B,3]*/,
                 float* grad_translation /*[B,3]*/, cdp_stream_t stream);
/* DepthHead.disp_to_depth (models/depth_head.py:49-54): depth = 1 / (1/max + (1/min - 1/max) disp). */
int cdp_disp_to_depth_fwd(const float* disp, size_t count, float min_depth, float max_depth,
                          float* depth, cdp_stream_t stream);
int cdp_disp_to_depth_bwd(const float* grad_depth, const float* depth, size_t count, float min_depth,
                          float max_depth, float* grad_disp, cdp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Regularisers of the object-motion maps: FlowSmoothnessLoss
 * (algos/depth.py:15-34) and FlowSparsityLoss (algos/depth.py:37-52), called next to the
 * reconstruction loss at algos/depth.py:483-485.
 *
 * maps: HOST array of n_maps (<= CDP_MAX_FLOW_MAPS) device pointers, each [planes, H, W] with
 * planes = B * C.  loss: device float[1] = (1/n_maps) sum_maps mean(...).  unit_grad: null, or
 * device [n_maps, planes, H, W] receiving d loss / d map (the losses are terminal scalars; backward
 * is cdp_scale_fwd with the upstream scalar).  scratch: cdp_flow_scratch_bytes(...) bytes.
 * ------------------------------------------------------------------------------------------- */
#define CDP_MAX_FLOW_MAPS 4
size_t cdp_flow_scratch_bytes(int32_t n_maps, int32_t planes, int32_t height, int32_t width,
                              int32_t sparsity /* 0: smoothness, 1: sparsity */);
/* mean sqrt((f - roll_x f)^2 + (f - roll_y f)^2 + 1e-7); wrap_around == 0 crops the first row and
 * column first (algos/depth.py:20-27).  2 launches. */
int cdp_flow_smooth_fwd(const float* const* maps, int32_t n_maps, int32_t planes, int32_t height,
                        int32_t width, int32_t wrap_around, float* loss, float* unit_grad,
                        void* scratch, size_t scratch_bytes, cdp_stream_t stream);
/* mean 2 m sqrt(|f| / (m + 1e-7) + 1) with m = mean_{H,W} |f| per plane, detached
 * (algos/depth.py:39-44).  3 launches. */
int cdp_flow_sparsity_fwd(const float* const* maps, int32_t n_maps, int32_t planes, int32_t height,
                          int32_t width, float* loss, float* unit_grad, void* scratch,
                          size_t scratch_bytes, cdp_stream_t stream);
/* out[i] = in[i] * scalar[0] (device scalar): backward of a terminal loss whose unit gradient was
 * produced in the forward pass.  in / out 16-byte aligned. */
int cdp_scale_fwd(const float* in, const float* scalar, size_t count, float* out, cdp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Camera-to-camera warp at constant depth: Mixup.warp_c2c
 * (datasets/mixup.py:211-229) = viewing rays of the TARGET camera on its pixel grid
 * (misc/image_warper.py:54-87, fp32), point at depth_val, projection with the SOURCE camera
 * (datasets/mixup.py:29-66, fp64) and F.grid_sample(align_corners=True) of the source image in
 * fp64.  src: device [B,C,Hs,Ws] float (src_is_f64 == 0) or double; out: device double
 * [B,C,Ht,Wt]; K_src / K_tgt: HOST double [B,4] = fx, fy, cx, cy.  nearest: 0 bilinear, 1 nearest;
 * padding_zeros: 0 "border", 1 "zeros".  No gradient.  1 launch per 32 samples.
 * ------------------------------------------------------------------------------------------- */
int cdp_warp_c2c_fwd(const void* src, int32_t src_is_f64, int32_t batch, int32_t channels,
                     int32_t src_height, int32_t src_width, int32_t out_height, int32_t out_width,
                     const double* K_src, const double* K_tgt, double depth_val, int32_t nearest,
                     int32_t padding_zeros, double* out, cdp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Depth metrics: DepthEvaluator.compute_depth_metrics
 * (eval/depth.py:21-70, statistics :109-133), called inside DepthAlgo.training
 * (algos/depth.py:468-469), and the per-class variant (eval/depth.py:72-106).
 *
 * A unit is one image (units = B, count = H*W) or the whole batch restricted to one class
 * (units = 1, count = B*H*W, labels + class_id).  Per unit over the elements with depth_gt > 0
 * (and inside the Garg crop when garg_crop != 0, which needs height * width == count): optional
 * median scaling pred *= median(gt) / median(pred) (lower median, as torch.median), clamp of both
 * to [min_depth, max_depth], then d_a1, d_a2, d_a3, d_rmse, d_rmse_log, d_abs_rel, d_sq_rel;
 * out[0..6] = mean of the unit values, out[7] = number of units that had ground truth (a unit
 * without any makes the means NaN; the reference raises there).  No host synchronisation.
 * depth_gt / depth_pred: device float [units, count]; labels: device int64 [units, count] or null;
 * out: device float[8]; scratch: cdp_depth_metrics_scratch_bytes(units, count) bytes.
 * 1 memset + 6 launches.
 * ------------------------------------------------------------------------------------------- */
size_t cdp_depth_metrics_scratch_bytes(int32_t units, int32_t count);
int cdp_depth_metrics_fwd(const float* depth_gt, const float* depth_pred, const int64_t* labels,
                          int64_t class_id, int32_t units, int32_t count, int32_t height, int32_t width,
                          int32_t garg_crop, float min_depth, float max_depth, int32_t use_gt_scale,
                          float* out, void* scratch, size_t scratch_bytes, cdp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CODEPS_PHOTO_H_ */
