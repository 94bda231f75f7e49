// cdp_plan.h -- host-side construction of the kernel parameter blocks from the C-ABI arguments.
// Shared by cdp_api.cu (which launches the kernels) and tests/emu (which loops over them on the
// CPU), so that the launch planning itself is covered by the CPU-side tests.
#pragma once

#include "cdp_kernels.h"

static inline int cdp_chunk_size(int B, int b0) {
  return B - b0 < CDP_MAX_BATCH_PER_LAUNCH ? B - b0 : CDP_MAX_BATCH_PER_LAUNCH;
}

static inline void cdp_fill_pyr_params(const CdpPlan& plan, const cdp_photo_args* a, CdpPyrParams* pp) {
  memset(pp, 0, sizeof(*pp));
  float* scratch = static_cast<float*>(a->scratch);
  pp->in[0] = a->target; pp->in[1] = a->source0; pp->in[2] = a->source1; pp->in[3] = a->depth;
  pp->in[4] = a->motion0; pp->in[5] = a->motion1;
  pp->nt = plan.has_motion ? 6 : 4;
  const CdpResizeTap* tab = static_cast<const CdpResizeTap*>(a->resize_tables);
  for (int s = 1; s < plan.L; ++s) {
    pp->out[0][s] = scratch + plan.off_tgt[s]; pp->out[1][s] = scratch + plan.off_src0[s];
    pp->out[2][s] = scratch + plan.off_src1[s]; pp->out[3][s] = scratch + plan.off_depth[s];
    if (plan.has_motion) { pp->out[4][s] = scratch + plan.off_mot[0][s]; pp->out[5][s] = scratch + plan.off_mot[1][s]; }
    pp->tab_x[s] = tab + plan.tab_fwd_x[s]; pp->tab_y[s] = tab + plan.tab_fwd_y[s];
  }
  for (int s = 0; s < plan.L; ++s) { pp->Ws[s] = plan.Ws[s]; pp->Hs[s] = plan.Hs[s]; }
  pp->W = plan.W; pp->H = plan.H; pp->L = plan.L;
  // level 1 with two 16-byte loads per channel and output pair: needs every bound input 16-byte
  // aligned (a contiguous fp32 view with an odd storage offset is not), else the table path runs
  bool aligned = true;
  for (int t = 0; t < pp->nt; ++t) aligned = aligned && (reinterpret_cast<uintptr_t>(pp->in[t]) & 15) == 0;
  if (a->heads && a->heads->disp) aligned = aligned && (reinterpret_cast<uintptr_t>(a->heads->disp) & 15) == 0;
  pp->fast1 = (plan.L > 1 && plan.W % 4 == 0 && plan.H % 2 == 0 && aligned) ? 1 : 0;
  // fused heads: the pyramid launch converts the disparity (needs the level-1 fast path, whose
  // items cover every full-resolution pixel exactly once) and builds the pose matrices
  pp->B = plan.B;
  if (a->heads) {
    const cdp_photo_heads* hd = a->heads;
    if (hd->disp && pp->fast1) {
      pp->in[3] = hd->disp;
      pp->depth_out = const_cast<float*>(a->depth);
      pp->min_disp = 1.0f / hd->max_depth;
      pp->disp_span = 1.0f / hd->min_depth - 1.0f / hd->max_depth;
    }
    if (hd->axisangle[0]) {
      for (int k = 0; k < 2; ++k) {
        pp->axisangle[k] = hd->axisangle[k]; pp->translation[k] = hd->translation[k]; pp->invert[k] = hd->invert[k];
      }
      pp->pose_out[0] = const_cast<float*>(a->pose0); pp->pose_out[1] = const_cast<float*>(a->pose1);
    }
  }
  int off = 0;
  pp->begin[0] = pp->begin[1] = 0;
  for (int s = 1; s < plan.L; ++s) {
    off += (s == 1 && pp->fast1) ? plan.Hs[s] * (plan.Ws[s] / 2) : plan.Hs[s] * plan.Ws[s];
    pp->begin[s + 1] = off;
  }
}

static inline void cdp_fill_photo_params(const CdpPlan& plan, const cdp_photo_args* a, int b0, int nb,
                                         CdpPhotoParams* kp) {
  memset(kp, 0, sizeof(*kp));
  float* scratch = static_cast<float*>(a->scratch);
  float* saved = static_cast<float*>(a->saved);
  const bool G = a->with_grad != 0;
  for (int s = 0; s < plan.L; ++s) {
    CdpLevel& lv = kp->lv[s];
    lv.tgt = s == 0 ? a->target : scratch + plan.off_tgt[s];
    lv.src0 = s == 0 ? a->source0 : scratch + plan.off_src0[s];
    lv.src1 = s == 0 ? a->source1 : scratch + plan.off_src1[s];
    lv.depth = s == 0 ? a->depth : scratch + plan.off_depth[s];
    lv.noise = a->noise[s];
    lv.gdepth = G ? saved + plan.off_gdepth[s] : nullptr;
    lv.argmin = a->argmin[s];
    if (plan.has_motion) {
      lv.mot0 = s == 0 ? a->motion0 : scratch + plan.off_mot[0][s];
      lv.mot1 = s == 0 ? a->motion1 : scratch + plan.off_mot[1][s];
      lv.gmot0 = G ? saved + plan.off_gmot[0][s] : nullptr;
      lv.gmot1 = G ? saved + plan.off_gmot[1][s] : nullptr;
    }
    lv.W = plan.Ws[s]; lv.H = plan.Hs[s];
    lv.tiles_x = plan.tiles_x[s]; lv.tiles_y = plan.tiles_y[s];
    lv.tiles_x_rcp = (uint32_t)((0x100000000ull + (uint64_t)plan.tiles_x[s] - 1) / (uint64_t)plan.tiles_x[s]);
    lv.block_begin = plan.block_begin[s];
    // mean over B*H_s*W_s, / 2^s, / num_levels (algos/depth.py:325-326)
    lv.weight = (float)(1.0 / ((double)plan.B * plan.Hs[s] * plan.Ws[s] * (double)(1 << s) * plan.L));
  }
  kp->K_tab = scratch + plan.off_ktab;
  kp->batch_total = plan.B;
  kp->pose0 = a->pose0; kp->pose1 = a->pose1;
  kp->partials = scratch + plan.off_partials;
  kp->seed = a->noise_seed;
  kp->seed_dev = a->noise_seed_dev;
  kp->num_levels = plan.L; kp->batch_begin = b0; kp->blocks_per_image = plan.blocks_per_image;
  kp->alpha = a->alpha;
}

// parameters of the intrinsics-table kernel for samples [b0, b0 + nb)
static inline void cdp_fill_k_table_params(const CdpPlan& plan, const cdp_photo_args* a, int b0, int nb,
                                           CdpKTableParams* tp) {
  memset(tp, 0, sizeof(*tp));
  tp->K_full = a->intrinsics_host ? nullptr : a->intrinsics_dev;
  tp->K_tab = static_cast<float*>(a->scratch) + plan.off_ktab;
  tp->B = plan.B; tp->L = plan.L; tp->batch_begin = b0; tp->batch_count = nb;
  for (int s = 0; s < plan.L; ++s) {
    // scale of CameraModel.get_scaled_model_image_size: python float ratio, rounded to fp32 on use
    tp->su[s] = (float)((double)plan.Ws[s] / (double)plan.W);
    tp->sv[s] = (float)((double)plan.Hs[s] / (double)plan.H);
    if (a->intrinsics_host)
      for (int i = 0; i < nb; ++i)
        for (int j = 0; j < 4; ++j) tp->K[s][i][j] = a->intrinsics_host[((size_t)s * plan.B + b0 + i) * 4 + j];
  }
}

static inline void cdp_fill_finalize_params(const CdpPlan& plan, const cdp_photo_args* a, CdpFinalizeParams* fp) {
  float* scratch = static_cast<float*>(a->scratch);
  float* saved = static_cast<float*>(a->saved);
  fp->partials = scratch + plan.off_partials;
  fp->loss = a->loss;
  fp->pose_unit = a->with_grad ? saved + plan.off_pose_unit : nullptr;
  bool own_noise = true;
  for (int s = 0; s < plan.L; ++s) own_noise = own_noise && a->noise[s] == nullptr;
  fp->seed_dev = own_noise ? a->noise_seed_dev : nullptr;
  fp->B = plan.B; fp->blocks_per_image = plan.blocks_per_image;
}

static inline void cdp_fill_depth_grad_params(const CdpPlan& plan, const void* saved_, const void* resize_tables,
                                              const float* grad_loss, float* grad_depth, float* grad_pose0,
                                              float* grad_pose1, CdpDepthGradParams* p) {
  memset(p, 0, sizeof(*p));
  const float* saved = static_cast<const float*>(saved_);
  const CdpResizeInv* tab = static_cast<const CdpResizeInv*>(resize_tables);
  for (int s = 0; s < plan.L; ++s) {
    p->gdepth[s] = saved + plan.off_gdepth[s];
    p->Ws[s] = plan.Ws[s]; p->Hs[s] = plan.Hs[s];
    if (s > 0) { p->inv_x[s] = tab + plan.tab_inv_x[s]; p->inv_y[s] = tab + plan.tab_inv_y[s]; }
    p->exact_x[s] = (plan.Ws[s] << s) == plan.W ? 1 : 0;
    p->exact_y[s] = (plan.Hs[s] << s) == plan.H ? 1 : 0;
  }
  p->grad_loss = grad_loss;
  p->pose_unit = saved + plan.off_pose_unit;
  p->grad_depth = grad_depth;
  p->grad_pose[0] = grad_pose0; p->grad_pose[1] = grad_pose1;
  p->B = plan.B; p->H = plan.H; p->W = plan.W; p->L = plan.L;
  p->scale_pose = 1;
}

// fused heads: dL/d disp instead of dL/d depth, dL/d (axis-angle, translation) instead of dL/dT
static inline void cdp_depth_grad_params_heads(const cdp_photo_heads* hd, const float* depth, float* grad_axisangle0,
                                               float* grad_translation0, float* grad_axisangle1,
                                               float* grad_translation1, CdpDepthGradParams* p) {
  p->depth_vals = depth;
  p->disp_span = 1.0f / hd->min_depth - 1.0f / hd->max_depth;
  for (int k = 0; k < 2; ++k) {
    p->axisangle[k] = hd->axisangle[k]; p->translation[k] = hd->translation[k]; p->invert[k] = hd->invert[k];
  }
  p->grad_axisangle[0] = grad_axisangle0; p->grad_translation[0] = grad_translation0;
  p->grad_axisangle[1] = grad_axisangle1; p->grad_translation[1] = grad_translation1;
}

// dL/d motion_k [B,3,H,W]: the same adjoint over 3B planes of the per-level motion gradients
static inline void cdp_fill_motion_grad_params(const CdpPlan& plan, const void* saved_, const void* resize_tables,
                                               const float* grad_loss, int k, float* grad_motion, CdpDepthGradParams* p) {
  cdp_fill_depth_grad_params(plan, saved_, resize_tables, grad_loss, grad_motion, nullptr, nullptr, p);
  const float* saved = static_cast<const float*>(saved_);
  for (int s = 0; s < plan.L; ++s) p->gdepth[s] = saved + plan.off_gmot[k][s];
  p->B = 3 * plan.B;
  p->scale_pose = 0;
}

static inline bool cdp_build_resize_tables(const CdpPlan& plan, void* host_out, int* bad_level) {
  char* base = static_cast<char*>(host_out);
  for (int s = 1; s < plan.L; ++s) {
    bool ok = cdp_resize_axis(plan.W, plan.Ws[s], reinterpret_cast<CdpResizeTap*>(base) + plan.tab_fwd_x[s],
                              reinterpret_cast<CdpResizeInv*>(base) + plan.tab_inv_x[s]);
    ok = cdp_resize_axis(plan.H, plan.Hs[s], reinterpret_cast<CdpResizeTap*>(base) + plan.tab_fwd_y[s],
                         reinterpret_cast<CdpResizeInv*>(base) + plan.tab_inv_y[s]) && ok;
    if (!ok) { *bad_level = s; return false; }
  }
  return true;
}

struct CdpSmoothLayout { size_t g, part, scal, total; int tiles_x, tiles_y; };
static inline CdpSmoothLayout cdp_smooth_layout(int32_t B, int32_t H, int32_t W) {
  CdpSmoothLayout l;
  l.tiles_x = (W + CDP_SMOOTH_TX - 1) / CDP_SMOOTH_TX;
  l.tiles_y = (H + CDP_SMOOTH_TY - 1) / CDP_SMOOTH_TY;
  // the row-walk kernel (W % 4 == 0) has its own block grid; the record area fits either
  const size_t qx = ((size_t)(W + 3) / 4 + CDP_SMOOTH_Q_THREADS - 1) / CDP_SMOOTH_Q_THREADS;
  const size_t qy = ((size_t)H + CDP_SMOOTH_Q_ROWS - 1) / CDP_SMOOTH_Q_ROWS;
  const size_t tiles = (size_t)l.tiles_x * l.tiles_y > qx * qy ? (size_t)l.tiles_x * l.tiles_y : qx * qy;
  size_t o = 0;
  l.g = o; o += cdp_align_floats((size_t)B * H * W);
  l.part = o; o += cdp_align_floats((size_t)B * tiles * 4);
  l.scal = o; o += cdp_align_floats((size_t)B * 2);
  l.total = o;
  return l;
}

static inline void cdp_fill_smooth_params(const float* image, const float* disp, int B, int H, int W, int with_grad,
                                          float* loss, float* saved, CdpSmoothParams* p) {
  const CdpSmoothLayout l = cdp_smooth_layout(B, H, W);
  p->image = image; p->disp = disp; p->g = saved + l.g; p->part = saved + l.part;
  p->scal = saved + l.scal; p->loss = loss;
  p->B = B; p->H = H; p->W = W; p->with_grad = with_grad;
  p->tiles_x = l.tiles_x; p->tiles_y = l.tiles_y;
  if (cdp_smooth_quad_ok(*p)) {  // block grid of the row-walk kernel
    p->tiles_x = (W / 4 + CDP_SMOOTH_Q_THREADS - 1) / CDP_SMOOTH_Q_THREADS;
    p->tiles_y = (H + CDP_SMOOTH_Q_ROWS - 1) / CDP_SMOOTH_Q_ROWS;
  }
}

static inline void cdp_fill_warp_params(CdpWarpParams* p, const float* src, int C, const float* depth,
                                        const float* pose, const float* motion, const float* K, int b0, int nb,
                                        int H, int W) {
  memset(p, 0, sizeof(*p));
  p->src = src; p->depth = depth; p->pose = pose; p->motion = motion;
  p->batch_begin = b0; p->C = C; p->H = H; p->W = W;
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < 4; ++j) p->K[i][j] = K[(size_t)(b0 + i) * 4 + j];
}
