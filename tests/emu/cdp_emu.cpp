// cdp_emu.cpp -- CPU emulator of the CUDA kernels.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the same kernel bodies (codeps_b200/csrc/cdp_kernels.h, cdp_math.h) and the same
// launch planning (cdp_plan.h) with g++ and runs every (block, thread) sequentially, one phase
// at a time where the CUDA kernel has a __syncthreads().  It lets the tile / halo / reflection
// adjoint / pyramid-table logic be checked against the oracle in the build container, which has
// no GPU.  It is built and loaded only by tests/; the codeps_b200 package never touches it, and
// it is never timed.  Entry points mirror include/codeps_photo.h with an emu_ prefix and HOST
// pointers everywhere.
#include <vector>

#include "cdp_plan.h"

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API size_t emu_resize_tables_bytes(int32_t h, int32_t w, int32_t l) {
  CdpPlan plan;
  if (!cdp_make_plan(1, h, w, l, &plan)) return 0;
  return (plan.tab_records > 0 ? plan.tab_records : 1) * sizeof(CdpResizeTap);
}
EMU_API int emu_resize_tables_build(int32_t h, int32_t w, int32_t l, void* out) {
  CdpPlan plan;
  if (!cdp_make_plan(1, h, w, l, &plan)) return CDP_ERR_INVALID;
  int bad = 0;
  return cdp_build_resize_tables(plan, out, &bad) ? CDP_OK : CDP_ERR_UNSUPPORTED;
}
EMU_API size_t emu_photo_scratch_bytes(int32_t b, int32_t h, int32_t w, int32_t l, int32_t m) {
  CdpPlan plan;
  return cdp_make_plan(b, h, w, l, &plan, m != 0) ? plan.scratch_floats * sizeof(float) : 0;
}
EMU_API size_t emu_photo_saved_bytes(int32_t b, int32_t h, int32_t w, int32_t l, int32_t m) {
  CdpPlan plan;
  return cdp_make_plan(b, h, w, l, &plan, m != 0) ? plan.saved_floats * sizeof(float) : 0;
}

// Tile lookup of the photo kernel for every block of one image: out[3 * bx + {0,1,2}] = level, x0, y0
// (the level search and the reciprocal-multiply division of cdp_tile_ctx, checked against plain
// integer arithmetic by tests/test_kernel_logic_cpu.py).  Returns the number of blocks.
EMU_API int emu_tile_table(int32_t h, int32_t w, int32_t l, int32_t* out, int32_t capacity) {
  CdpPlan plan;
  if (!cdp_make_plan(1, h, w, l, &plan, false)) return -1;
  cdp_photo_args a;
  memset(&a, 0, sizeof(a));
  a.batch = 1; a.height = h; a.width = w; a.num_levels = l;
  static float base[1];  // only addresses are formed from the buffers here, nothing is read or written
  a.scratch = base; a.saved = base;
  CdpPhotoParams kp;
  cdp_fill_photo_params(plan, &a, 0, 1, &kp);
  if (plan.blocks_per_image > capacity) return -2;
  for (int bx = 0; bx < plan.blocks_per_image; ++bx) {
    const CdpTileCtx c = cdp_tile_ctx(kp, bx, 0);
    out[3 * bx] = c.lvl; out[3 * bx + 1] = c.x0; out[3 * bx + 2] = c.y0;
  }
  return plan.blocks_per_image;
}

template <bool G, bool M>
static void emu_photo_block(const CdpPhotoParams& kp, int bx, int by) {
  typedef CdpTileGeom<G> Geo;
  const int nt = CDP_PHOTO_THREADS;
  std::vector<float> sm(Geo::SMEM_BYTES / sizeof(float) + 4, 0.f);
  const CdpTileCtx c = cdp_tile_ctx(kp, bx, by);
  std::vector<float> v((size_t)nt * 33, 0.f);
  // phase S: the box fill the GPU does with TMA (zero fill outside the image), here with plain loads
  for (int t = 0; t < nt; ++t) cdp_photo_stage<G>(kp, c, t, nt, sm.data());
  CdpTileConst kc;
  cdp_tile_const(kp, c, kc);
  for (int t = 0; t < nt; ++t) cdp_photo_phase_a<G, M>(kp, c, t, nt, sm.data(), kc);
  for (int t = 0; t < nt; ++t) cdp_photo_phase_b1<G>(kp, c, t, nt, sm.data(), v[(size_t)t * 33]);
  if (G) {
    for (int t = 0; t < nt; ++t) cdp_photo_phase_b2(kp, c, t, nt, sm.data());
    for (int t = 0; t < nt; ++t) cdp_photo_phase_c1(kp, c, t, nt, sm.data());
    for (int t = 0; t < nt; ++t) cdp_photo_restage_sources(kp, c, t, nt, sm.data());
    for (int t = 0; t < nt; ++t) cdp_photo_phase_c2<M>(kp, c, t, nt, sm.data(), &v[(size_t)t * 33 + 1], kc);
  }
  float* rec = kp.partials + ((size_t)c.b * kp.blocks_per_image + bx) * CDP_PARTIAL_STRIDE;
  for (int j = 0; j < 33; ++j) {
    float acc = 0.f;
    for (int t = 0; t < nt; ++t) acc += (j == 0 ? v[(size_t)t * 33] * kp.lv[c.lvl].weight : v[(size_t)t * 33 + j]);
    rec[j] = acc;
  }
}

EMU_API int emu_photo_fwd(const cdp_photo_args* a) {
  CdpPlan plan;
  if (!cdp_make_plan(a->batch, a->height, a->width, a->num_levels, &plan, a->motion0 != nullptr)) return CDP_ERR_INVALID;
  // fused heads: the same fallbacks as cdp_photo_fwd where the pyramid launch cannot do the conversion
  if (a->heads) {
    const cdp_photo_heads* hd = a->heads;
    CdpPyrParams probe;
    if (plan.L > 1) cdp_fill_pyr_params(plan, a, &probe);
    if (hd->disp && (plan.L == 1 || !probe.depth_out)) {
      const float lo = 1.0f / hd->max_depth, span = 1.0f / hd->min_depth - lo;
      float* out = const_cast<float*>(a->depth);
      for (size_t i = 0; i < (size_t)plan.B * plan.H * plan.W; ++i) out[i] = cdp_disp_to_depth(hd->disp[i], lo, span);
    }
    if (hd->axisangle[0] && plan.L == 1)
      for (int k = 0; k < 2; ++k)
        for (int b = 0; b < plan.B; ++b)
          cdp_pose_fwd_sample(hd->axisangle[k] + 3 * b, hd->translation[k] + 3 * b, hd->invert[k],
                              const_cast<float*>(k == 0 ? a->pose0 : a->pose1) + 16 * b);
  }
  if (plan.L > 1) {
    CdpPyrParams pp;
    cdp_fill_pyr_params(plan, a, &pp);
    if (pp.pose_out[0])
      for (int j = 0; j < 2 * pp.B; ++j) cdp_pyr_pose_item(pp, j);
    for (int b = 0; b < plan.B; ++b)
      for (int i = 0; i < pp.begin[plan.L]; ++i) cdp_pyramid_fwd_item(pp, b, i);
  }
  for (int b0 = 0; b0 < plan.B; b0 += CDP_MAX_BATCH_PER_LAUNCH) {  // per-level intrinsics table
    const int nb = cdp_chunk_size(plan.B, b0);
    CdpKTableParams tp;
    cdp_fill_k_table_params(plan, a, b0, nb, &tp);
    for (int i = 0; i < nb * plan.L; ++i) cdp_k_table_entry(tp, i / nb, i % nb);
  }
  {
    CdpPhotoParams kp;
    cdp_fill_photo_params(plan, a, 0, plan.B, &kp);
    for (int by = 0; by < plan.B; ++by)
      for (int bx = 0; bx < plan.blocks_per_image; ++bx) {
        const int which = (a->with_grad ? 2 : 0) + (plan.has_motion ? 1 : 0);
        switch (which) {
          case 0: emu_photo_block<false, false>(kp, bx, by); break;
          case 1: emu_photo_block<false, true>(kp, bx, by); break;
          case 2: emu_photo_block<true, false>(kp, bx, by); break;
          default: emu_photo_block<true, true>(kp, bx, by); break;
        }
      }
  }
  CdpFinalizeParams fp;
  cdp_fill_finalize_params(plan, a, &fp);
  std::vector<double> sm(2048 + 32);
  for (int b = 0; b < fp.B; ++b) {
    for (int t = 0; t < CDP_FINALIZE_THREADS; ++t) cdp_finalize_phase_a(fp, b, t, sm.data());
    for (int t = 0; t < CDP_FINALIZE_THREADS; ++t) cdp_finalize_phase_b(fp, b, t, sm.data());
    for (int t = 0; t < CDP_FINALIZE_THREADS; ++t) cdp_finalize_phase_c(fp, b, t, sm.data());
  }
  return CDP_OK;
}

EMU_API int emu_photo_bwd(int32_t b, int32_t h, int32_t w, int32_t l, const void* saved, const void* tables,
                          const float* grad_loss, float* grad_depth, float* gp0, float* gp1, int32_t with_motion,
                          float* gm0, float* gm1) {
  CdpPlan plan;
  if (!cdp_make_plan(b, h, w, l, &plan, with_motion != 0)) return CDP_ERR_INVALID;
  CdpDepthGradParams p;
  cdp_fill_depth_grad_params(plan, saved, tables, grad_loss, grad_depth, gp0, gp1, &p);
  for (int i = 0; i < plan.B; ++i)
    for (int pix = 0; pix < h * w; ++pix) cdp_depth_grad_pixel(p, i, pix);
  for (int i = 0; i < 2 * plan.B * 16; ++i) cdp_pose_grad_scale(p, i);
  if (with_motion) {
    float* outs[2] = {gm0, gm1};
    for (int k = 0; k < 2; ++k) {
      CdpDepthGradParams pm;
      cdp_fill_motion_grad_params(plan, saved, tables, grad_loss, k, outs[k], &pm);
      for (int i = 0; i < pm.B; ++i)
        for (int pix = 0; pix < h * w; ++pix) cdp_depth_grad_pixel(pm, i, pix);
    }
  }
  return CDP_OK;
}

EMU_API int emu_photo_bwd_heads(int32_t b, int32_t h, int32_t w, int32_t l, const void* saved, const void* tables,
                                const float* grad_loss, const cdp_photo_heads* heads, const float* depth, float* grad_disp,
                                float* ga0, float* gt0, float* ga1, float* gt1) {
  CdpPlan plan;
  if (!cdp_make_plan(b, h, w, l, &plan, false)) return CDP_ERR_INVALID;
  CdpDepthGradParams p;
  cdp_fill_depth_grad_params(plan, saved, tables, grad_loss, grad_disp, nullptr, nullptr, &p);
  cdp_depth_grad_params_heads(heads, depth, ga0, gt0, ga1, gt1, &p);
  for (int i = 0; i < plan.B; ++i)
    for (int pix = 0; pix < h * w; ++pix) cdp_depth_grad_pixel(p, i, pix);
  for (int j = 0; j < 2 * plan.B; ++j) cdp_pose_grad_heads(p, j);
  return CDP_OK;
}

EMU_API size_t emu_smooth_saved_bytes(int32_t b, int32_t h, int32_t w) {
  return cdp_smooth_layout(b, h, w).total * sizeof(float);
}

EMU_API int emu_smooth_fwd(const float* image, const float* disp, int32_t B, int32_t H, int32_t W, int32_t with_grad,
                           float* loss, void* saved) {
  CdpSmoothParams p;
  cdp_fill_smooth_params(image, disp, B, H, W, with_grad, loss, static_cast<float*>(saved), &p);
  const int nt = CDP_SMOOTH_THREADS, nb = p.tiles_x * p.tiles_y;
  if (cdp_smooth_quad_ok(p)) {  // row-walk kernel
    for (int b = 0; b < B; ++b)
      for (int by = 0; by < p.tiles_y; ++by)
        for (int bx = 0; bx < p.tiles_x; ++bx) {
          float tot[4] = {0.f, 0.f, 0.f, 0.f};
          for (int t = 0; t < CDP_SMOOTH_Q_THREADS; ++t) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            cdp_smooth_quad_thread(p, b, bx, by, t, v);
            for (int j = 0; j < 4; ++j) tot[j] += v[j];
          }
          for (int j = 0; j < 4; ++j) p.part[(((size_t)b * p.tiles_y + by) * p.tiles_x + bx) * 4 + j] = tot[j];
        }
    cdp_smooth_finalize(p);
    return CDP_OK;
  }
  std::vector<float> sm(CDP_SMOOTH_SMEM_FLOATS);
  for (int b = 0; b < B; ++b)
    for (int tile = 0; tile < nb; ++tile) {
      std::vector<float> v((size_t)nt * 4, 0.f);
      for (int t = 0; t < nt; ++t) cdp_smooth_phase_load(p, b, tile, t, nt, sm.data());
      for (int t = 0; t < nt; ++t) cdp_smooth_phase_edges(p, tile, t, nt, sm.data(), &v[(size_t)t * 4]);
      for (int t = 0; t < nt; ++t) cdp_smooth_phase_grad(p, b, tile, t, nt, sm.data(), &v[(size_t)t * 4]);
      for (int j = 0; j < 4; ++j) {
        float tot = 0.f;
        for (int t = 0; t < nt; ++t) tot += v[(size_t)t * 4 + j];
        p.part[((size_t)b * nb + tile) * 4 + j] = tot;
      }
    }
  cdp_smooth_finalize(p);
  return CDP_OK;
}

EMU_API int emu_smooth_bwd(const void* saved_, const float* grad_loss, int32_t B, int32_t H, int32_t W,
                           float* grad_disp) {
  const CdpSmoothLayout l = cdp_smooth_layout(B, H, W);
  const float* saved = static_cast<const float*>(saved_);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H * W; i += 4)
      cdp_smooth_bwd_run(saved + l.g, saved + l.scal, grad_loss, b, (size_t)H * W, i, H * W - i < 4 ? H * W - i : 4, grad_disp);
  return CDP_OK;
}

// object-motion regularisers (cdp_flow.h): kind 0 = smoothness, 1 = sparsity
EMU_API int emu_flow_loss(const float* const* maps, int32_t n_maps, int32_t planes, int32_t H, int32_t W,
                          int32_t kind, int32_t wrap, float* loss, float* unit_grad) {
  CdpFlowParams p;
  const bool sparsity = kind != 0;
  std::vector<float> part(cdp_flow_records(n_maps, planes, H, W, sparsity) + 1, 0.f);
  if (!cdp_fill_flow_params(maps, n_maps, planes, H, W, wrap, sparsity, loss, unit_grad, part.data(), &p))
    return CDP_ERR_INVALID;
  const int nt = CDP_FLOW_THREADS, nz = planes * n_maps;
  if (!sparsity) {
    for (int bz = 0; bz < nz; ++bz)
      for (int by = 0; by < p.blocks_y; ++by)
        for (int bx = 0; bx < p.blocks_x; ++bx) {
          float tot = 0.f;
          for (int t = 0; t < nt; ++t) tot += cdp_flow_smooth_thread(p, bx, by, bz, t);
          part[((size_t)bz * p.blocks_y + by) * p.blocks_x + bx] = tot;
        }
    const size_t count = (size_t)nz * p.blocks_y * p.blocks_x;
    loss[0] = (float)(cdp_flow_sum_records_host(part.data(), count) * (double)p.inv_count);
    return CDP_OK;
  }
  const size_t count = (size_t)nz * p.blocks_x;
  for (int bz = 0; bz < nz; ++bz)
    for (int bx = 0; bx < p.blocks_x; ++bx) {
      float tot = 0.f;
      for (int t = 0; t < nt; ++t) tot += cdp_flow_abs_thread(p, bx, bz, t);
      part[(size_t)bz * p.blocks_x + bx] = tot;
    }
  for (int bz = 0; bz < nz; ++bz) {
    double lanes[32];
    for (int l = 0; l < 32; ++l) lanes[l] = cdp_lane_sum(part.data() + (size_t)bz * p.blocks_x, p.blocks_x, 1, l);
    const float mean = cdp_flow_plane_mean(p, bz, lanes);
    for (int bx = 0; bx < p.blocks_x; ++bx) {
      float tot = 0.f;
      for (int t = 0; t < nt; ++t) tot += cdp_flow_sparsity_thread(p, bx, bz, t, mean);
      part[count + (size_t)bz * p.blocks_x + bx] = tot;
    }
  }
  loss[0] = (float)(cdp_flow_sum_records_host(part.data() + count, count) * (double)p.inv_count);
  return CDP_OK;
}

// depth metrics (cdp_metrics.h): sequential form of the four histogram passes, the statistics
// pass and the finalize step
EMU_API int emu_depth_metrics(const float* gt, const float* pred, const int64_t* labels, int64_t class_id, int32_t units,
                              int32_t n, int32_t H, int32_t W, int32_t garg, float lo, float hi, int32_t use_gt_scale,
                              float* out) {
  std::vector<unsigned char> scratch(cdp_metrics_scratch_total(units, n), 0);
  CdpMetricsParams p;
  if (!cdp_fill_metrics_params(&p, gt, pred, labels, class_id, units, n, W, H, garg, lo, hi, use_gt_scale,
                               scratch.data(), out))
    return CDP_ERR_INVALID;
  for (int pass = 0; pass < CDP_METRICS_PASSES; ++pass)
    for (int u = 0; u < units; ++u) {
      uint32_t pre[2] = {0, 0}, rank, count = pass ? cdp_metrics_count(p, u) : 0;
      for (int arr = 0; arr < 2 && pass; ++arr) cdp_metrics_state(p, u, arr, pass, count, pre[arr], rank);
      uint32_t* h = p.hist + (((size_t)pass * units + u) * 2) * 256;
      for (int i = 0; i < n; ++i) {
        float g;
        if (!cdp_metrics_valid(p, u, i, g)) continue;
        const uint32_t kg = cdp_metrics_key(g), kp = cdp_metrics_key(pred[(size_t)u * n + i]);
        if (cdp_metrics_match(kg, pre[0], pass)) ++h[(kg >> (24 - 8 * pass)) & 255u];
        if (cdp_metrics_match(kp, pre[1], pass)) ++h[256 + ((kp >> (24 - 8 * pass)) & 255u)];
      }
    }
  double total[CDP_METRICS_NSTATS] = {0, 0, 0, 0, 0, 0, 0};
  int with_gt = 0;
  for (int u = 0; u < units; ++u) {
    const uint32_t count = cdp_metrics_count(p, u);
    uint32_t kg, kp, rank;
    cdp_metrics_state(p, u, 0, CDP_METRICS_PASSES, count, kg, rank);
    cdp_metrics_state(p, u, 1, CDP_METRICS_PASSES, count, kp, rank);
    const float ratio = use_gt_scale ? cdp_metrics_unkey(kg) / cdp_metrics_unkey(kp) : 1.0f;
    double sums[CDP_METRICS_NSTATS] = {0, 0, 0, 0, 0, 0, 0};
    for (int blk = 0; blk < p.blocks; ++blk) {  // per-block fp32 records, fp64 combine
      float acc[CDP_METRICS_NSTATS] = {0, 0, 0, 0, 0, 0, 0};
      for (int i = blk * CDP_METRICS_CHUNK; i < n && i < (blk + 1) * CDP_METRICS_CHUNK; ++i) {
        float g;
        if (cdp_metrics_valid(p, u, i, g)) cdp_metrics_element(p, g, pred[(size_t)u * n + i], ratio, acc);
      }
      for (int j = 0; j < CDP_METRICS_NSTATS; ++j) sums[j] += acc[j];
    }
    double st[CDP_METRICS_NSTATS];
    const bool ok = cdp_metrics_unit_stats(sums, count, st);
    with_gt += ok;
    for (int j = 0; j < CDP_METRICS_NSTATS; ++j) total[j] += ok ? st[j] : (double)NAN;
  }
  for (int j = 0; j < CDP_METRICS_NSTATS; ++j) out[j] = (float)(total[j] / units);
  out[CDP_METRICS_NSTATS] = (float)with_gt;
  return CDP_OK;
}

// camera-to-camera warp (cdp_c2c.h)
EMU_API int emu_warp_c2c(const void* src, int32_t src_is_f64, int32_t B, int32_t C, int32_t Hs, int32_t Ws, int32_t Ht,
                         int32_t Wt, const double* Ks, const double* Kt, double depth, int32_t nearest, int32_t zeros,
                         double* out) {
  for (int b0 = 0; b0 < B; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(B, b0);
    CdpC2cParams p;
    cdp_fill_c2c_params(&p, src, out, Ks, Kt, b0, nb, C, Hs, Ws, Ht, Wt, depth, nearest, zeros);
    for (int i = 0; i < nb; ++i)
      for (int pix = 0; pix < Ht * Wt; ++pix) {
        if (src_is_f64) cdp_c2c_pixel<double>(p, i, pix);
        else cdp_c2c_pixel<float>(p, i, pix);
      }
  }
  return CDP_OK;
}

EMU_API int emu_warp_grid_fwd(const float* depth, const float* pose, const float* motion, const float* K, int32_t B,
                              int32_t H, int32_t W, float* grid) {
  for (int b0 = 0; b0 < B; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(B, b0);
    CdpWarpParams p;
    cdp_fill_warp_params(&p, nullptr, 0, depth, pose, motion, K, b0, nb, H, W);
    p.out = grid;
    for (int i = 0; i < nb; ++i)
      for (int pix = 0; pix < H * W; ++pix) cdp_warp_grid_pixel(p, i, pix);
  }
  return CDP_OK;
}

EMU_API int emu_warp_image_fwd(const float* src, int32_t C, const float* depth, const float* pose,
                               const float* motion, const float* K, int32_t B, int32_t H, int32_t W, int32_t mode,
                               float* out) {
  for (int b0 = 0; b0 < B; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(B, b0);
    CdpWarpParams p;
    cdp_fill_warp_params(&p, src, C, depth, pose, motion, K, b0, nb, H, W);
    p.out = out; p.mode = mode;
    for (int i = 0; i < nb; ++i)
      for (int pix = 0; pix < H * W; ++pix) cdp_warp_image_pixel(p, i, pix);
  }
  return CDP_OK;
}

EMU_API int emu_warp_image_bwd(const float* grad_out, const float* src, int32_t C, const float* depth,
                               const float* pose, const float* motion, const float* K, int32_t B, int32_t H,
                               int32_t W, float* grad_depth, float* grad_pose, float* grad_motion) {
  for (int b0 = 0; b0 < B; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(B, b0);
    CdpWarpParams p;
    cdp_fill_warp_params(&p, src, C, depth, pose, motion, K, b0, nb, H, W);
    p.grad_out = grad_out; p.grad_depth = grad_depth; p.grad_motion = grad_motion;
    for (int i = 0; i < nb; ++i) {
      double tot[16] = {0};
      for (int pix = 0; pix < H * W; ++pix) {
        float dT[16] = {0};
        cdp_warp_bwd_pixel(p, i, pix, dT);
        for (int j = 0; j < 16; ++j) tot[j] += dT[j];
      }
      for (int j = 0; j < 16; ++j) grad_pose[(size_t)(b0 + i) * 16 + j] = (float)tot[j];
    }
  }
  return CDP_OK;
}

EMU_API int emu_ssim_fwd(const float* x, const float* y, int32_t planes, int32_t H, int32_t W, float* out) {
  for (int pl = 0; pl < planes; ++pl)
    for (int pix = 0; pix < H * W; ++pix) cdp_ssim_fwd_pixel(x, y, W, H, pl, pix, out);
  return CDP_OK;
}

EMU_API int emu_ssim_bwd(const float* go, const float* x, const float* y, int32_t planes, int32_t H, int32_t W,
                         float* gx, float* gy) {
  const size_t total = (size_t)planes * H * W;
  std::vector<float> scratch(4 * total);
  for (int pl = 0; pl < planes; ++pl)
    for (int pix = 0; pix < H * W; ++pix) cdp_ssim_bwd_coef_pixel(go, x, y, W, H, pl, pix, scratch.data(), total);
  for (int pl = 0; pl < planes; ++pl)
    for (int pix = 0; pix < H * W; ++pix) cdp_ssim_bwd_gather_pixel(x, y, W, H, pl, pix, scratch.data(), total, gx, gy);
  return CDP_OK;
}

EMU_API int emu_pose_fwd(const float* aa, const float* tr, int32_t B, int32_t invert, float* M) {
  for (int b = 0; b < B; ++b) cdp_pose_fwd_sample(aa + 3 * b, tr + 3 * b, invert, M + 16 * b);
  return CDP_OK;
}
EMU_API int emu_pose_bwd(const float* gM, const float* aa, const float* tr, int32_t B, int32_t invert, float* gaa, float* gtr) {
  for (int b = 0; b < B; ++b) cdp_pose_bwd_sample(gM + 16 * b, aa + 3 * b, tr + 3 * b, invert, gaa + 3 * b, gtr + 3 * b);
  return CDP_OK;
}
