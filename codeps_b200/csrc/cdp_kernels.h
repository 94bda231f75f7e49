// cdp_kernels.h -- kernel bodies of the photometric-loss hot path, written per (block, thread).
//
// Every body is a host/device function of an explicit block index, thread index and (where
// used) a shared-memory pointer.  Phases that need a block barrier between them are separate
// functions: cdp_api.cu calls them in order with __syncthreads() in between, tests/emu loops
// over the threads of a block phase by phase.
#pragma once

#include "cdp_math.h"

// ==========================================================================================
// 1. Image pyramid: F.interpolate(bilinear, align_corners=False) from full resolution for the
//    target, both sources and depth, all levels >= 1 in one launch (algos/depth.py:280-281,295).
// ==========================================================================================
struct CdpPyrParams {
  const float* in[4];                  // target, source0, source1 [B,3,H,W]; depth [B,1,H,W]
  float* out[4][CDP_MAX_LEVELS];       // per level (index 0 unused)
  const CdpResizeTap* tab_x[CDP_MAX_LEVELS];
  const CdpResizeTap* tab_y[CDP_MAX_LEVELS];
  int32_t Ws[CDP_MAX_LEVELS], Hs[CDP_MAX_LEVELS];
  int32_t begin[CDP_MAX_LEVELS + 1];   // prefix offsets of level outputs within one image
  int32_t W, H, L;
};

CDP_HD void cdp_pyramid_fwd_item(const CdpPyrParams& p, int b, int item) {
  if (item >= p.begin[p.L]) return;
  int s = 1;
  while (s + 1 < p.L && item >= p.begin[s + 1]) ++s;
  const int local = item - p.begin[s];
  const int ws = p.Ws[s], hs = p.Hs[s];
  const int y = local / ws, x = local - y * ws;
  const CdpResizeTap tx = p.tab_x[s][x], ty = p.tab_y[s][y];
  const size_t in_plane = (size_t)p.W * p.H, out_plane = (size_t)ws * hs;
  const int o00 = ty.i0 * p.W + tx.i0, o01 = ty.i0 * p.W + tx.i1;
  const int o10 = ty.i1 * p.W + tx.i0, o11 = ty.i1 * p.W + tx.i1;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int ch = t == 3 ? 1 : 3;
    for (int c = 0; c < ch; ++c) {
      const float* src = p.in[t] + ((size_t)b * ch + c) * in_plane;
      const float top = CDP_LDG(src + o00) * tx.w0 + CDP_LDG(src + o01) * tx.w1;
      const float bot = CDP_LDG(src + o10) * tx.w0 + CDP_LDG(src + o11) * tx.w1;
      p.out[t][s][((size_t)b * ch + c) * out_plane + local] = top * ty.w0 + bot * ty.w1;
    }
  }
}

// ==========================================================================================
// 2. Fused photometric tile kernel.
// ==========================================================================================
template <bool G>
struct CdpTileGeom {
  static constexpr int HALO = G ? 2 : 1;  // warped / source / target values staged around the tile
  static constexpr int RW = CDP_TILE_X + 2 * HALO, RH = CDP_TILE_Y + 2 * HALO, RN = RW * RH;
  static constexpr int HB = HALO - 1;     // ring on which losses / argmin / coefficients are needed
  static constexpr int BW = CDP_TILE_X + 2 * HB, BH = CDP_TILE_Y + 2 * HB, BN = BW * BH;
  // shared-memory planes of RN floats each
  static constexpr int P_WARP = 0;   // 6: warped source k, channel c at k*3+c
  static constexpr int P_TGT = 6;    // 3
  static constexpr int P_SRC = 9;    // 6: un-warped sources (identity candidates)
  static constexpr int P_ID = 15;    // 2: identity losses per source
  static constexpr int P_EXTRA = 17; // 1 (with grad)
  static constexpr int NPLANES = G ? 18 : 17;
  static constexpr size_t SMEM_BYTES = (size_t)NPLANES * RN * sizeof(float) + ((RN + 15) & ~15);
  // After the identity pass the source planes are dead: the 9 coefficient fields of the winning
  // reprojection (A,B,C per channel) reuse P_SRC+0..5, P_ID+0..1 and P_EXTRA.
  static constexpr CDP_HD int coef_plane(int j) { return j < 6 ? P_SRC + j : (j < 8 ? P_ID + (j - 6) : P_EXTRA); }
};

struct CdpTileCtx {
  int lvl, b, b_local, x0, y0;  // level, sample (global / within launch), tile origin
};

CDP_HD CdpTileCtx cdp_tile_ctx(const CdpPhotoParams& p, int bx, int by) {
  CdpTileCtx c;
  int s = p.num_levels - 1;
  while (s > 0 && bx < p.lv[s].block_begin) --s;
  c.lvl = s;
  const int tile = bx - p.lv[s].block_begin;
  const int ty = tile / p.lv[s].tiles_x;
  c.x0 = (tile - ty * p.lv[s].tiles_x) * CDP_TILE_X;
  c.y0 = ty * CDP_TILE_Y;
  c.b_local = by;
  c.b = p.batch_begin + by;
  return c;
}

CDP_HD CdpCam cdp_tile_cam(const CdpPhotoParams& p, const CdpTileCtx& c) {
  CdpCam k;
  k.fx = p.K[c.lvl][c.b_local][0]; k.fy = p.K[c.lvl][c.b_local][1];
  k.cx = p.K[c.lvl][c.b_local][2]; k.cy = p.K[c.lvl][c.b_local][3];
  return k;
}

// Phase A: warp both sources for every staged position (tile + halo, reflected at the image
// border) and stage target / source values.
template <bool G>
CDP_HD void cdp_photo_phase_a(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads,
                              float* sm) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  const size_t plane = (size_t)W * H;
  const CdpCam cam = cdp_tile_cam(p, c);
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float* T[2] = {p.pose0 + (size_t)c.b * 16, p.pose1 + (size_t)c.b * 16};
  const float* src[2] = {lv.src0 + (size_t)c.b * 3 * plane, lv.src1 + (size_t)c.b * 3 * plane};
  const float* tgt = lv.tgt + (size_t)c.b * 3 * plane;
  for (int idx = tid; idx < Geo::RN; idx += nthreads) {
    const int ry = idx / Geo::RW, rx = idx - ry * Geo::RW;
    const int px = c.x0 - Geo::HALO + rx, py = c.y0 - Geo::HALO + ry;
    if (px < -1 || px > W || py < -1 || py > H) continue;  // never read
    const int u = cdp_reflect(px, W), v = cdp_reflect(py, H);
    const int pix = v * W + u;
    CdpPoint pt;
    cdp_backproject((float)u, (float)v, CDP_LDG(lv.depth + (size_t)c.b * plane + pix), cam, pt);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      sm[(Geo::P_TGT + ch) * Geo::RN + idx] = CDP_LDG(tgt + ch * plane + pix);
      sm[(Geo::P_SRC + ch) * Geo::RN + idx] = CDP_LDG(src[0] + ch * plane + pix);
      sm[(Geo::P_SRC + 3 + ch) * Geo::RN + idx] = CDP_LDG(src[1] + ch * plane + pix);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      CdpProj pr;
      cdp_project(pt.P, T[k], nullptr, cam, wm1, hm1, pr);
      CdpTaps t;
      cdp_taps(pr.ix, pr.iy, W, H, t);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        sm[(Geo::P_WARP + k * 3 + ch) * Geo::RN + idx] = cdp_bilinear(src[k] + ch * plane, t);
    }
  }
}

// 3x3 window sums in the reference's order (avg_pool2d: row-major accumulation, then / 9) for
// two candidate images sharing one target.  xs0/xs1/ys point at the window centre.
struct CdpPairStats {
  float mx[2], exx[2], exy[2], my, eyy;
};

CDP_HD void cdp_window_stats(const float* xs0, const float* xs1, const float* ys, int pitch,
                             CdpPairStats& o) {
  float sx0 = 0.f, sx1 = 0.f, sy = 0.f, sxx0 = 0.f, sxx1 = 0.f, syy = 0.f, sxy0 = 0.f, sxy1 = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int o = dy * pitch + dx;
      const float y = ys[o], a = xs0[o], b = xs1[o];
      sy = CDP_ADD(sy, y); syy = CDP_ADD(syy, CDP_MUL(y, y));
      sx0 = CDP_ADD(sx0, a); sxx0 = CDP_ADD(sxx0, CDP_MUL(a, a)); sxy0 = CDP_ADD(sxy0, CDP_MUL(a, y));
      sx1 = CDP_ADD(sx1, b); sxx1 = CDP_ADD(sxx1, CDP_MUL(b, b)); sxy1 = CDP_ADD(sxy1, CDP_MUL(b, y));
    }
  o.my = sy / 9.0f; o.eyy = syy / 9.0f;
  o.mx[0] = sx0 / 9.0f; o.exx[0] = sxx0 / 9.0f; o.exy[0] = sxy0 / 9.0f;
  o.mx[1] = sx1 / 9.0f; o.exx[1] = sxx1 / 9.0f; o.exy[1] = sxy1 / 9.0f;
}

// ReconstructionLoss._compute_loss (algos/depth.py:234-236) for two candidates at one position.
// xbase = first plane of candidate 0 (candidate 1 follows 3 planes later).  coef (optional,
// [2][9]) receives A,B,C per channel for both candidates.
template <bool WANT_COEF, int RN, int RW, int P_TGT>
CDP_HD void cdp_pair_losses(const float* sm, int xbase, int ridx, float alpha, float loss[2],
                            float (*coef)[9]) {
  float ssim_sum[2] = {0.f, 0.f}, l1_sum[2] = {0.f, 0.f};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float* xs0 = sm + (xbase + ch) * RN + ridx;
    const float* xs1 = sm + (xbase + 3 + ch) * RN + ridx;
    const float* ys = sm + (P_TGT + ch) * RN + ridx;
    CdpPairStats st;
    cdp_window_stats(xs0, xs1, ys, RW, st);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      CdpSsimTerms t;
      cdp_ssim_terms(st.mx[k], st.my, st.exx[k], st.eyy, st.exy[k], t);
      ssim_sum[k] = CDP_ADD(ssim_sum[k], t.loss);
      const float x = k == 0 ? xs0[0] : xs1[0];
      l1_sum[k] = CDP_ADD(l1_sum[k], fabsf(CDP_SUB(x, ys[0])));
      if (WANT_COEF) cdp_ssim_coeffs(st.mx[k], st.my, x, ys[0], t, coef[k][ch * 3], coef[k][ch * 3 + 1], coef[k][ch * 3 + 2]);
    }
  }
  const float one_minus_alpha = (float)(1.0 - (double)alpha);
#pragma unroll
  for (int k = 0; k < 2; ++k)
    loss[k] = CDP_ADD(CDP_MUL(alpha, ssim_sum[k] / 3.0f), CDP_MUL(one_minus_alpha, l1_sum[k] / 3.0f));
}

// Phase B1: identity (un-warped) losses on the statistics ring -> P_ID planes.
template <bool G>
CDP_HD void cdp_photo_phase_b1(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads,
                               float* sm) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  for (int idx = tid; idx < Geo::BN; idx += nthreads) {
    const int by = idx / Geo::BW, bx = idx - by * Geo::BW;
    const int qx = c.x0 - Geo::HB + bx, qy = c.y0 - Geo::HB + by;
    if (qx < 0 || qx >= lv.W || qy < 0 || qy >= lv.H) continue;
    const int ridx = (by + 1) * Geo::RW + bx + 1;
    float id[2];
    cdp_pair_losses<false, Geo::RN, Geo::RW, Geo::P_TGT>(sm, Geo::P_SRC, ridx, p.alpha, id, nullptr);
    sm[(Geo::P_ID + 0) * Geo::RN + ridx] = id[0];
    sm[(Geo::P_ID + 1) * Geo::RN + ridx] = id[1];
  }
}

// Phase B2: reprojection losses, tie-break noise, min / argmin (algos/depth.py:316-323),
// per-thread loss sum over the tile proper, argmin map, and (with grad) the SSIM adjoint
// coefficient fields of the winning reprojection.
template <bool G>
CDP_HD void cdp_photo_phase_b2(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads,
                               float* sm, float& loss_acc) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  uint8_t* kplane = reinterpret_cast<uint8_t*>(sm + (size_t)Geo::NPLANES * Geo::RN);
  for (int idx = tid; idx < Geo::BN; idx += nthreads) {
    const int by = idx / Geo::BW, bx = idx - by * Geo::BW;
    const int qx = c.x0 - Geo::HB + bx, qy = c.y0 - Geo::HB + by;
    const int ridx = (by + 1) * Geo::RW + bx + 1;
    if (qx < 0 || qx >= W || qy < 0 || qy >= H) {
      if (G) kplane[ridx] = 255;
      continue;
    }
    float pe[2];
    float coef[2][9];
    cdp_pair_losses<G, Geo::RN, Geo::RW, Geo::P_TGT>(sm, Geo::P_WARP, ridx, p.alpha, pe, coef);
    float n0, n1;
    if (lv.noise) {
      const size_t plane = (size_t)W * H;
      n0 = CDP_LDG(lv.noise + ((size_t)c.b * 2 + 0) * plane + qy * W + qx);
      n1 = CDP_LDG(lv.noise + ((size_t)c.b * 2 + 1) * plane + qy * W + qx);
    } else {
      cdp_noise_pair(p.seed, (uint32_t)(qy * W + qx), (uint32_t)c.lvl, (uint32_t)c.b, n0, n1);
    }
    const float id0 = CDP_ADD(sm[(Geo::P_ID + 0) * Geo::RN + ridx], CDP_MUL(n0, CDP_NOISE_SCALE));
    const float id1 = CDP_ADD(sm[(Geo::P_ID + 1) * Geo::RN + ridx], CDP_MUL(n1, CDP_NOISE_SCALE));
    float best = pe[0];
    int kb = 0;
    if (pe[1] < best) { best = pe[1]; kb = 1; }
    if (id0 < best) { best = id0; kb = 2; }
    if (id1 < best) { best = id1; kb = 3; }
    const bool in_tile = qx >= c.x0 && qx < c.x0 + CDP_TILE_X && qy >= c.y0 && qy < c.y0 + CDP_TILE_Y;
    if (in_tile) {
      loss_acc += best;
      if (lv.argmin) lv.argmin[(size_t)c.b * W * H + qy * W + qx] = (uint8_t)kb;
    }
    if (G) {
      kplane[ridx] = (uint8_t)kb;
      if (kb < 2) {
#pragma unroll
        for (int j = 0; j < 9; ++j) sm[Geo::coef_plane(j) * Geo::RN + ridx] = kb == 0 ? coef[0][j] : coef[1][j];
      }
    }
  }
}

// Phase C (with grad): gather the SSIM adjoint over the reflected 3x3 neighbourhood, add the L1
// term, chain through the bilinear sampler and the projection to depth and pose.
CDP_HD void cdp_photo_phase_c(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads,
                              const float* sm, float* dT /*[32]*/) {
  typedef CdpTileGeom<true> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  const size_t plane = (size_t)W * H;
  const CdpCam cam = cdp_tile_cam(p, c);
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float* T[2] = {p.pose0 + (size_t)c.b * 16, p.pose1 + (size_t)c.b * 16};
  const float* src[2] = {lv.src0 + (size_t)c.b * 3 * plane, lv.src1 + (size_t)c.b * 3 * plane};
  const uint8_t* kplane = reinterpret_cast<const uint8_t*>(sm + (size_t)Geo::NPLANES * Geo::RN);
  const float w_ssim = p.alpha / 27.0f;                       // alpha * (1/3 channels) * (1/9 window)
  const float w_l1 = (float)(1.0 - (double)p.alpha) / 3.0f;
  for (int idx = tid; idx < CDP_TILE_X * CDP_TILE_Y; idx += nthreads) {
    const int ly = idx / CDP_TILE_X, lx = idx - ly * CDP_TILE_X;
    const int px = c.x0 + lx, py = c.y0 + ly;
    if (px >= W || py >= H) continue;
    const int ridx = (ly + Geo::HALO) * Geo::RW + lx + Geo::HALO;
    const int kown = kplane[ridx];
    float gd = 0.f;
    CdpPoint pt;
    bool have_pt = false;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float acc[3] = {0.f, 0.f, 0.f};  // sum over the window of m * (A0 + 2 (x_p - x_q) B + (y_p - y_q) C)
      bool any = kown == k;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const float my = cdp_reflect_mult(py, dy, H);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const float m = my * cdp_reflect_mult(px, dx, W);
          const int n = ridx + dy * Geo::RW + dx;
          if (m == 0.f || kplane[n] != k) continue;
          any = true;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float dxv = sm[(Geo::P_WARP + k * 3 + ch) * Geo::RN + ridx] - sm[(Geo::P_WARP + k * 3 + ch) * Geo::RN + n];
            const float dyv = sm[(Geo::P_TGT + ch) * Geo::RN + ridx] - sm[(Geo::P_TGT + ch) * Geo::RN + n];
            acc[ch] += m * (sm[Geo::coef_plane(ch * 3 + 0) * Geo::RN + n] +
                            2.f * dxv * sm[Geo::coef_plane(ch * 3 + 1) * Geo::RN + n] +
                            dyv * sm[Geo::coef_plane(ch * 3 + 2) * Geo::RN + n]);
          }
        }
      }
      if (!any) continue;
      if (!have_pt) {
        cdp_backproject((float)px, (float)py, CDP_LDG(lv.depth + (size_t)c.b * plane + py * W + px), cam, pt);
        have_pt = true;
      }
      CdpProj pr;
      cdp_project(pt.P, T[k], nullptr, cam, wm1, hm1, pr);
      CdpTaps t;
      cdp_taps(pr.ix, pr.iy, W, H, t);
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float x = sm[(Geo::P_WARP + k * 3 + ch) * Geo::RN + ridx];
        const float y = sm[(Geo::P_TGT + ch) * Geo::RN + ridx];
        float gw = w_ssim * acc[ch];
        if (kown == k) gw += w_l1 * (x > y ? 1.f : (x < y ? -1.f : 0.f));
        gw *= lv.weight;
        float dix, diy;
        cdp_bilinear_grad(src[k] + ch * plane, t, dix, diy);
        gix += gw * dix;
        giy += gw * diy;
      }
      cdp_warp_adjoint(gix * t.mx, giy * t.my, pr, pt, T[k], cam, gd, dT + 16 * k, nullptr);
    }
    lv.gdepth[(size_t)c.b * plane + py * W + px] = gd;
  }
}

// ==========================================================================================
// 3. Fixed-order reduction of the per-CTA partial records (one block).
//    loss[0] = sum of all weighted loss partials; pose_unit[k][b][16] = sum over the CTAs of
//    image b.  Thread (r = tid / 32, j = tid % 32) strides over CTAs; rows are then combined in
//    index order, so the result does not depend on scheduling.
// ==========================================================================================
struct CdpFinalizeParams {
  const float* partials;  // [B][blocks_per_image][CDP_PARTIAL_STRIDE]
  float* loss;            // [1]
  float* pose_unit;       // [2][B][16] or null
  int32_t B, blocks_per_image;
};

#define CDP_FINALIZE_THREADS 1024

// step 1: per-thread strided sums into sm[r][33] (double precision)
CDP_HD void cdp_finalize_phase_a(const CdpFinalizeParams& p, int b, int tid, double* sm) {
  const int r = tid >> 5, j = tid & 31;
  double acc = 0.0, lacc = 0.0;
  for (int blk = r; blk < p.blocks_per_image; blk += 32) {
    const float* rec = p.partials + ((size_t)b * p.blocks_per_image + blk) * CDP_PARTIAL_STRIDE;
    acc += (double)rec[1 + j];
    if (j == 0) lacc += (double)rec[0];
  }
  sm[r * 33 + j] = acc;
  if (j == 0) sm[r * 33 + 32] = lacc;
}
// step 2: combine the 32 rows in order; threads 0..32 each own one output column
CDP_HD void cdp_finalize_phase_b(const CdpFinalizeParams& p, int b, int tid, const double* sm,
                                 double* loss_acc /* thread 32's running loss over images */) {
  if (tid > 32) return;
  double acc = 0.0;
  for (int r = 0; r < 32; ++r) acc += sm[r * 33 + tid];
  if (tid < 32) {
    if (p.pose_unit) p.pose_unit[((size_t)(tid >> 4) * p.B + b) * 16 + (tid & 15)] = (float)acc;
  } else {
    *loss_acc += acc;
  }
}

// ==========================================================================================
// 4. Backward: dL/d depth = grad_loss * (G_0 + sum_s resize_s^T G_s); poses scaled alongside.
// ==========================================================================================
struct CdpDepthGradParams {
  const float* gdepth[CDP_MAX_LEVELS];  // unit gradients per level [B,H_s,W_s]
  const CdpResizeInv* inv_x[CDP_MAX_LEVELS];
  const CdpResizeInv* inv_y[CDP_MAX_LEVELS];
  int32_t Ws[CDP_MAX_LEVELS], Hs[CDP_MAX_LEVELS];
  const float* grad_loss;  // device scalar
  const float* pose_unit;  // [2][B][16]
  float* grad_depth;       // [B,1,H,W]
  float* grad_pose[2];     // [B,16]
  int32_t B, H, W, L;
};

CDP_HD void cdp_depth_grad_pixel(const CdpDepthGradParams& p, int b, int pix) {
  const int y = pix / p.W, x = pix - y * p.W;
  float acc = CDP_LDG(p.gdepth[0] + (size_t)b * p.W * p.H + pix);
  for (int s = 1; s < p.L; ++s) {
    const CdpResizeInv ex = p.inv_x[s][x], ey = p.inv_y[s][y];
    const float* g = p.gdepth[s] + (size_t)b * p.Ws[s] * p.Hs[s];
    float row_a = 0.f, row_b = 0.f;
    if (ey.ja >= 0) {
      if (ex.ja >= 0) row_a += ex.wa * CDP_LDG(g + ey.ja * p.Ws[s] + ex.ja);
      if (ex.jb >= 0) row_a += ex.wb * CDP_LDG(g + ey.ja * p.Ws[s] + ex.jb);
    }
    if (ey.jb >= 0) {
      if (ex.ja >= 0) row_b += ex.wa * CDP_LDG(g + ey.jb * p.Ws[s] + ex.ja);
      if (ex.jb >= 0) row_b += ex.wb * CDP_LDG(g + ey.jb * p.Ws[s] + ex.jb);
    }
    acc += ey.wa * row_a + ey.wb * row_b;
  }
  p.grad_depth[(size_t)b * p.W * p.H + pix] = CDP_LDG(p.grad_loss) * acc;
}

CDP_HD void cdp_pose_grad_scale(const CdpDepthGradParams& p, int i) {  // i in [0, 2*B*16)
  const int k = i / (p.B * 16), r = i - k * p.B * 16;
  p.grad_pose[k][r] = CDP_LDG(p.grad_loss) * p.pose_unit[i];
}

// ==========================================================================================
// 5. Edge-aware smoothness (algos/depth.py:58-107).
// ==========================================================================================
#define CDP_SMOOTH_BLOCKS 64   // blocks per image for the two reduction passes
#define CDP_SMOOTH_THREADS 256

struct CdpSmoothParams {
  const float* image;  // [B,3,H,W]
  const float* disp;   // [B,1,H,W]
  float* g;            // [B,H,W]   d loss / d normalised disparity (unit)
  float* part_sum;     // [B][CDP_SMOOTH_BLOCKS]          partial sums of disp
  float* part_main;    // [B][CDP_SMOOTH_BLOCKS][4]       sum tx, sum ty, sum g*disp
  float* scal;         // [B][2]  a_b = 1/(mean+eps), c_b = sum(g*disp) a_b^2 / (H W)
  float* loss;         // [1]
  int32_t B, H, W, with_grad;
};

// chunk of pixels handled by block `blk` of an image
CDP_HD void cdp_smooth_chunk(int HW, int blk, int& lo, int& hi) {
  const int per = (HW + CDP_SMOOTH_BLOCKS - 1) / CDP_SMOOTH_BLOCKS;
  lo = blk * per;
  hi = lo + per < HW ? lo + per : HW;
  if (lo > HW) lo = HW;
}

CDP_HD float cdp_smooth_sum_thread(const CdpSmoothParams& p, int b, int blk, int tid, int nthreads) {
  int lo, hi;
  cdp_smooth_chunk(p.H * p.W, blk, lo, hi);
  const float* d = p.disp + (size_t)b * p.H * p.W;
  float acc = 0.f;
  for (int i = lo + tid; i < hi; i += nthreads) acc += CDP_LDG(d + i);
  return acc;
}

// mean disparity of image b from the per-block partial sums (fixed order)
CDP_HD float cdp_smooth_mean(const CdpSmoothParams& p, int b) {
  double acc = 0.0;
  for (int i = 0; i < CDP_SMOOTH_BLOCKS; ++i) acc += (double)p.part_sum[b * CDP_SMOOTH_BLOCKS + i];
  return (float)(acc / (double)((size_t)p.H * p.W));
}

CDP_HD float cdp_edge_weight(const float* img, size_t plane, int a, int bidx) {
  const float s = CDP_ADD(CDP_ADD(fabsf(CDP_LDG(img + a) - CDP_LDG(img + bidx)),
                                  fabsf(CDP_LDG(img + plane + a) - CDP_LDG(img + plane + bidx))),
                          fabsf(CDP_LDG(img + 2 * plane + a) - CDP_LDG(img + 2 * plane + bidx)));
  return expf(-(s / 3.0f));
}

CDP_HD float cdp_sign(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// per-thread partial sums over this block's chunk: acc[0] += sum tx, acc[1] += sum ty,
// acc[2] += sum g*disp; writes g.
CDP_HD void cdp_smooth_main_thread(const CdpSmoothParams& p, int b, int blk, int tid, int nthreads,
                                   float mean, float acc[3]) {
  const int W = p.W, H = p.H;
  int lo, hi;
  cdp_smooth_chunk(H * W, blk, lo, hi);
  const size_t plane = (size_t)H * W;
  const float* d = p.disp + (size_t)b * plane;
  const float* img = p.image + (size_t)b * 3 * plane;
  const float den = CDP_ADD(mean, 1e-7f);
  const float cx = 1.0f / ((float)p.B * (float)H * (float)(W - 1));
  const float cy = 1.0f / ((float)p.B * (float)(H - 1) * (float)W);
  for (int i = lo + tid; i < hi; i += nthreads) {
    const int y = i / W, x = i - y * W;
    const float raw = CDP_LDG(d + i);
    const float dc = raw / den;
    float g = 0.f;
    if (x < W - 1) {
      const float diff = CDP_SUB(dc, CDP_LDG(d + i + 1) / den);
      const float e = cdp_edge_weight(img, plane, i, i + 1);
      acc[0] += CDP_MUL(fabsf(diff), e);
      g += cdp_sign(diff) * e * cx;
    }
    if (y < H - 1) {
      const float diff = CDP_SUB(dc, CDP_LDG(d + i + W) / den);
      const float e = cdp_edge_weight(img, plane, i, i + W);
      acc[1] += CDP_MUL(fabsf(diff), e);
      g += cdp_sign(diff) * e * cy;
    }
    if (p.with_grad) {
      if (x > 0) {
        const float diff = CDP_SUB(CDP_LDG(d + i - 1) / den, dc);
        g -= cdp_sign(diff) * cdp_edge_weight(img, plane, i - 1, i) * cx;
      }
      if (y > 0) {
        const float diff = CDP_SUB(CDP_LDG(d + i - W) / den, dc);
        g -= cdp_sign(diff) * cdp_edge_weight(img, plane, i - W, i) * cy;
      }
      p.g[(size_t)b * plane + i] = g;
      acc[2] += g * raw;
    }
  }
}

// one thread: combine the block partials of all images (fixed order)
CDP_HD void cdp_smooth_finalize(const CdpSmoothParams& p) {
  double sx = 0.0, sy = 0.0;
  for (int b = 0; b < p.B; ++b) {
    double gd = 0.0;
    for (int i = 0; i < CDP_SMOOTH_BLOCKS; ++i) {
      const float* rec = p.part_main + ((size_t)b * CDP_SMOOTH_BLOCKS + i) * 4;
      sx += (double)rec[0]; sy += (double)rec[1]; gd += (double)rec[2];
    }
    if (p.with_grad) {
      const float den = CDP_ADD(cdp_smooth_mean(p, b), 1e-7f);
      const double a = 1.0 / (double)den;
      p.scal[b * 2 + 0] = (float)a;
      p.scal[b * 2 + 1] = (float)(gd * a * a / (double)((size_t)p.H * p.W));
    }
  }
  const double nx = (double)p.B * p.H * (p.W - 1), ny = (double)p.B * (p.H - 1) * p.W;
  p.loss[0] = (float)(sx / nx) + (float)(sy / ny);
}

CDP_HD void cdp_smooth_bwd_pixel(const float* g, const float* scal, const float* grad_loss, int b,
                                 size_t plane, int i, float* grad_disp) {
  grad_disp[(size_t)b * plane + i] =
      CDP_LDG(grad_loss) * (CDP_LDG(g + (size_t)b * plane + i) * scal[b * 2] - scal[b * 2 + 1]);
}

// ==========================================================================================
// 6. Stand-alone operators.
// ==========================================================================================
struct CdpWarpParams {
  const float* src;      // [B,C,H,W] (null for grid output)
  const float* depth;    // [B,1,H,W]
  const float* pose;     // [B,16]
  const float* motion;   // [B,3,H,W] or null
  const float* grad_out; // [B,C,H,W] (backward)
  float* out;            // [B,C,H,W] or grid [B,H,W,2]
  float* grad_depth;     // [B,H,W]
  float* grad_motion;    // [B,3,H,W] or null
  float* partials;       // [B][blocks][16]
  float K[CDP_MAX_BATCH_PER_LAUNCH][4];
  int32_t batch_begin, C, H, W, mode;
};

CDP_HD void cdp_warp_setup(const CdpWarpParams& p, int b_local, int pix, CdpPoint& pt, CdpProj& pr,
                           CdpCam& cam) {
  const int b = p.batch_begin + b_local;
  const size_t plane = (size_t)p.H * p.W;
  const int y = pix / p.W, x = pix - y * p.W;
  cam.fx = p.K[b_local][0]; cam.fy = p.K[b_local][1]; cam.cx = p.K[b_local][2]; cam.cy = p.K[b_local][3];
  cdp_backproject((float)x, (float)y, CDP_LDG(p.depth + (size_t)b * plane + pix), cam, pt);
  float mo[3];
  if (p.motion) {
    for (int c = 0; c < 3; ++c) mo[c] = CDP_LDG(p.motion + ((size_t)b * 3 + c) * plane + pix);
  }
  cdp_project(pt.P, p.pose + (size_t)b * 16, p.motion ? mo : nullptr, cam, (float)(p.W - 1),
              (float)(p.H - 1), pr);
}

CDP_HD void cdp_warp_grid_pixel(const CdpWarpParams& p, int b_local, int pix) {
  CdpPoint pt; CdpProj pr; CdpCam cam;
  cdp_warp_setup(p, b_local, pix, pt, pr, cam);
  const size_t o = (((size_t)(p.batch_begin + b_local)) * p.H * p.W + pix) * 2;
  p.out[o] = pr.gx;
  p.out[o + 1] = pr.gy;
}

CDP_HD void cdp_warp_image_pixel(const CdpWarpParams& p, int b_local, int pix) {
  CdpPoint pt; CdpProj pr; CdpCam cam;
  cdp_warp_setup(p, b_local, pix, pt, pr, cam);
  const int b = p.batch_begin + b_local;
  const size_t plane = (size_t)p.H * p.W;
  if (p.mode == 0) {
    CdpTaps t;
    cdp_taps(pr.ix, pr.iy, p.W, p.H, t);
    for (int c = 0; c < p.C; ++c)
      p.out[((size_t)b * p.C + c) * plane + pix] = cdp_bilinear(p.src + ((size_t)b * p.C + c) * plane, t);
  } else {
    const int o = cdp_nearest_index(pr.ix, pr.iy, p.W, p.H);
    for (int c = 0; c < p.C; ++c)
      p.out[((size_t)b * p.C + c) * plane + pix] = CDP_LDG(p.src + ((size_t)b * p.C + c) * plane + o);
  }
}

// backward of the bilinear warp for one pixel; accumulates this thread's dT[16]
CDP_HD void cdp_warp_bwd_pixel(const CdpWarpParams& p, int b_local, int pix, float* dT) {
  CdpPoint pt; CdpProj pr; CdpCam cam;
  cdp_warp_setup(p, b_local, pix, pt, pr, cam);
  const int b = p.batch_begin + b_local;
  const size_t plane = (size_t)p.H * p.W;
  CdpTaps t;
  cdp_taps(pr.ix, pr.iy, p.W, p.H, t);
  float gix = 0.f, giy = 0.f;
  for (int c = 0; c < p.C; ++c) {
    float dix, diy;
    cdp_bilinear_grad(p.src + ((size_t)b * p.C + c) * plane, t, dix, diy);
    const float go = CDP_LDG(p.grad_out + ((size_t)b * p.C + c) * plane + pix);
    gix += go * dix;
    giy += go * diy;
  }
  float gd = 0.f, gm[3];
  cdp_warp_adjoint(gix * t.mx, giy * t.my, pr, pt, p.pose + (size_t)b * 16, cam, gd, dT,
                   p.grad_motion ? gm : nullptr);
  p.grad_depth[(size_t)b * plane + pix] = gd;
  if (p.grad_motion)
    for (int c = 0; c < 3; ++c) p.grad_motion[((size_t)b * 3 + c) * plane + pix] = gm[c];
}

// SSIM map for one pixel of one plane (reflect-padded 3x3 statistics straight from global memory)
CDP_HD void cdp_ssim_stats_global(const float* x, const float* y, int W, int H, int px, int py,
                                  float& mx, float& my, float& exx, float& eyy, float& exy) {
  float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int o = cdp_reflect(py + dy, H) * W + cdp_reflect(px + dx, W);
      const float a = CDP_LDG(x + o), b = CDP_LDG(y + o);
      sx = CDP_ADD(sx, a); sy = CDP_ADD(sy, b);
      sxx = CDP_ADD(sxx, CDP_MUL(a, a)); syy = CDP_ADD(syy, CDP_MUL(b, b)); sxy = CDP_ADD(sxy, CDP_MUL(a, b));
    }
  mx = sx / 9.0f; my = sy / 9.0f; exx = sxx / 9.0f; eyy = syy / 9.0f; exy = sxy / 9.0f;
}

CDP_HD void cdp_ssim_fwd_pixel(const float* x, const float* y, int W, int H, int plane_idx, int pix,
                               float* out) {
  const size_t base = (size_t)plane_idx * W * H;
  float mx, my, exx, eyy, exy;
  cdp_ssim_stats_global(x + base, y + base, W, H, pix % W, pix / W, mx, my, exx, eyy, exy);
  CdpSsimTerms t;
  cdp_ssim_terms(mx, my, exx, eyy, exy, t);
  out[base + pix] = t.loss;
}

// backward pass 1: coefficient fields scaled by the upstream gradient -> scratch[4][planes*H*W]
// (A0_x, A0_y, B, C in the centred form of cdp_ssim_coeffs)
CDP_HD void cdp_ssim_bwd_coef_pixel(const float* grad_out, const float* x, const float* y, int W,
                                    int H, int plane_idx, int pix, float* scratch, size_t total) {
  const size_t base = (size_t)plane_idx * W * H;
  float mx, my, exx, eyy, exy;
  cdp_ssim_stats_global(x + base, y + base, W, H, pix % W, pix / W, mx, my, exx, eyy, exy);
  CdpSsimTerms t;
  cdp_ssim_terms(mx, my, exx, eyy, exy, t);
  const float xq = CDP_LDG(x + base + pix), yq = CDP_LDG(y + base + pix);
  float Ax, Bx, C, Ay, By;
  cdp_ssim_coeffs(mx, my, xq, yq, t, Ax, Bx, C);
  cdp_ssim_coeffs(my, mx, yq, xq, t, Ay, By, C);  // SSIM is symmetric in (x, y)
  const float go = CDP_LDG(grad_out + base + pix);
  scratch[0 * total + base + pix] = go * Ax;
  scratch[1 * total + base + pix] = go * Ay;
  scratch[2 * total + base + pix] = go * Bx;
  scratch[3 * total + base + pix] = go * C;
}

// backward pass 2: gather over the reflected neighbourhood
CDP_HD void cdp_ssim_bwd_gather_pixel(const float* x, const float* y, int W, int H, int plane_idx,
                                      int pix, const float* scratch, size_t total, float* grad_x,
                                      float* grad_y) {
  const size_t base = (size_t)plane_idx * W * H;
  const int py = pix / W, px = pix - py * W;
  const float xv = CDP_LDG(x + base + pix), yv = CDP_LDG(y + base + pix);
  float gx = 0.f, gy = 0.f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const float m = cdp_reflect_mult(py, dy, H) * cdp_reflect_mult(px, dx, W);
      if (m == 0.f) continue;
      const size_t o = base + (size_t)(py + dy) * W + (px + dx);
      const float ddx = xv - CDP_LDG(x + o), ddy = yv - CDP_LDG(y + o);
      const float b = scratch[2 * total + o], c = scratch[3 * total + o];
      gx += m * (scratch[0 * total + o] + 2.f * ddx * b + ddy * c);
      gy += m * (scratch[1 * total + o] + 2.f * ddy * b + ddx * c);
    }
  if (grad_x) grad_x[base + pix] = gx / 9.0f;
  if (grad_y) grad_y[base + pix] = gy / 9.0f;
}
