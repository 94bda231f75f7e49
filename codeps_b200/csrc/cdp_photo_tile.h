// cdp_photo_tile.h -- the fused photometric tile kernel (host/device bodies, phase by phase).
//
// One CTA = one 32x32 tile of one pyramid level of one sample.  Phases (a block barrier between
// consecutive ones):
//   A   warp both source frames for every staged position (tile + halo, reflected at the image
//       border), stage target / source / warped values in shared memory, centred on a tile
//       constant; the two candidates of a pair are interleaved as float2.
//   B1  min-reprojection: each thread owns a vertical strip of 5 pixels in one column and
//       slides a 3-row window down it (separable 3x3 sums), all four candidates at once, the two
//       candidates of a pair in the two lanes of packed fp32 (FFMA2).  -> candidate losses,
//       tie-break noise, min / argmin, loss partial, argmin map, winner plane.
//   B2  (with grad) same walk over the reprojection pair only, SSIM adjoint coefficients of the
//       winner -> shared memory (the source planes are dead by then and are reused).
//   C   (with grad) gather the adjoint over the reflected 3x3 neighbourhood, L1 term, chain
//       through the bilinear sampler and the projection to depth and pose.
// Reference: algos/depth.py:221-237, 272-325 (ReconstructionLoss), 128-155 (SSIMLoss);
// misc/image_warper.py:100-184.
#pragma once

#include "cdp_math.h"

#ifndef CDP_OPT_PACKED_GATHER
#define CDP_OPT_PACKED_GATHER 0  // float2-packed coefficient gather in phase C: measured slower (spills), off
#endif
#ifndef CDP_OPT_DIRECT_DT
#define CDP_OPT_DIRECT_DT 1  // phase C accumulates dL/dT without a per-pixel temporary
#endif
#ifndef CDP_OPT_PREFETCH_DEPTH
#define CDP_OPT_PREFETCH_DEPTH 1  // phase A requests the next region pixel's depth one iteration ahead: -1 %
#endif
#ifndef CDP_STRIP
#define CDP_STRIP 5  // pixels per thread strip in phases B1/B2
#endif

template <bool G>
struct CdpTileGeom {
  static constexpr int HALO = G ? 2 : 1;  // staged ring around the tile
  static constexpr int RW = CDP_TILE_X + 2 * HALO, RH = CDP_TILE_Y + 2 * HALO, RN = RW * RH;
  static constexpr int HB = HALO - 1;     // ring on which losses / argmin / coefficients are needed
  static constexpr int BW = CDP_TILE_X + 2 * HB, BH = CDP_TILE_Y + 2 * HB;
  static constexpr int NSTRIP = (BH + CDP_STRIP - 1) / CDP_STRIP;
  static constexpr int NITEMS = BW * NSTRIP;  // (column, strip) work items of phases B1/B2
  // shared-memory planes of RN floats
  static constexpr int P_TGT = 0;    // 3 planes: target, channel c
  static constexpr int P_WARP = 3;   // 3 float2 planes (6 floats): warped (source 0, source 1), channel c
  static constexpr int P_SRC = 9;    // 3 float2 planes: un-warped (source 0, source 1) = identity candidates
  // with grad, after B1: adjoint coefficients of the winner as 5 float2 planes
  // (A0,A1) (B0,B1) (C0,C1) (A2,B2) (C2,-)  [letter = coefficient, digit = channel]
  static constexpr int P_COEF = 9;
  static constexpr int NPLANES = G ? (CDP_OPT_PACKED_GATHER ? 19 : 18) : 15;
  static constexpr size_t SMEM_BYTES = (size_t)NPLANES * RN * sizeof(float) + ((RN + 15) & ~15);
  // The B1/B2 strip walk always reads CDP_STRIP + 2 region rows, also for the last, partial strip:
  // the rows past RH belong to the following plane (values discarded) and, for the last plane, to
  // the winner-byte tail, which must therefore be large enough for any CDP_TILE_Y / CDP_STRIP.
  static constexpr int OVER_ROWS = NSTRIP * CDP_STRIP + 2 > RH ? NSTRIP * CDP_STRIP + 2 - RH : 0;
  static_assert((size_t)OVER_ROWS * RW * sizeof(float2) <= ((RN + 15) & ~15) + (size_t)(NPLANES - 15) * RN * sizeof(float),
                "strip over-read leaves the shared-memory allocation: pick CDP_TILE_Y / CDP_STRIP so that it fits");
};

struct CdpTileCtx {
  int lvl, b, x0, y0;  // level, sample, tile origin
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
  c.b = p.batch_begin + by;
  return c;
}

CDP_HD CdpCam cdp_tile_cam(const CdpPhotoParams& p, const CdpTileCtx& c) {
  // per-level intrinsics table (rows are 16-byte aligned)
  const float4 k = CDP_LDG(reinterpret_cast<const float4*>(p.K_tab) + (size_t)c.lvl * p.batch_total + c.b);
  return cdp_make_cam(k.x, k.y, k.z, k.w);
}

// Entry (level s, sample b) of the per-level intrinsics table: the scaling of
// CameraModel.get_scaled_model_image_size (misc/camera_model.py:36-41) -- a python float ratio of
// the two image sizes that meets the fp32 calibration value, i.e. rounded to fp32 before the product.
CDP_HD void cdp_k_table_entry(const CdpKTableParams& p, int s, int b_local) {
  const int b = p.batch_begin + b_local;
  float* o = p.K_tab + ((size_t)s * p.B + b) * 4;
  if (p.K_full) {
    const float* k = p.K_full + (size_t)b * 4;
    o[0] = CDP_MUL(CDP_LDG(k + 0), p.su[s]); o[1] = CDP_MUL(CDP_LDG(k + 1), p.sv[s]);
    o[2] = CDP_MUL(CDP_LDG(k + 2), p.su[s]); o[3] = CDP_MUL(CDP_LDG(k + 3), p.sv[s]);
  } else {
    for (int j = 0; j < 4; ++j) o[j] = p.K[s][b_local][j];
  }
}

// Per-tile, per-channel constant subtracted from every staged image value (target value at the
// tile centre).  SSIM variances / covariances, the L1 term and all value differences are
// invariant to it; it only shrinks the magnitudes that get squared.
CDP_HD void cdp_tile_centre(const CdpLevel& lv, const CdpTileCtx& c, float centre[3]) {
  const int cx = c.x0 + CDP_TILE_X / 2 < lv.W ? c.x0 + CDP_TILE_X / 2 : lv.W - 1;
  const int cy = c.y0 + CDP_TILE_Y / 2 < lv.H ? c.y0 + CDP_TILE_Y / 2 : lv.H - 1;
  const size_t plane = (size_t)lv.W * lv.H;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) centre[ch] = CDP_LDG(lv.tgt + ((size_t)c.b * 3 + ch) * plane + cy * lv.W + cx);
}

template <bool G>
CDP_HD float2* cdp_pair_plane(float* sm, int first_plane, int ch) {
  return reinterpret_cast<float2*>(sm + (size_t)(first_plane + 2 * ch) * CdpTileGeom<G>::RN);
}
template <bool G>
CDP_HD const float2* cdp_pair_plane(const float* sm, int first_plane, int ch) {
  return reinterpret_cast<const float2*>(sm + (size_t)(first_plane + 2 * ch) * CdpTileGeom<G>::RN);
}

// ------------------------------------------------------------------------------------------
// Phase A
// ------------------------------------------------------------------------------------------
template <bool G, bool M>
CDP_HD void cdp_photo_phase_a(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  const size_t plane = (size_t)W * H;
  const CdpCam cam = cdp_tile_cam(p, c);
  CdpPose T[2];
  cdp_load_pose_aligned(p.pose0 + (size_t)c.b * 16, T[0]);  // cdp_photo_fwd requires 16-byte aligned poses
  cdp_load_pose_aligned(p.pose1 + (size_t)c.b * 16, T[1]);
  float centre[3];
  cdp_tile_centre(lv, c, centre);
  const float* src0 = lv.src0 + (size_t)c.b * 3 * plane;
  const float* src1 = lv.src1 + (size_t)c.b * 3 * plane;
  const float* tgt = lv.tgt + (size_t)c.b * 3 * plane;
#if CDP_OPT_PREFETCH_DEPTH
  // the depth of the next region pixel is requested one iteration ahead: the loop body has two
  // dependent global-memory round trips (depth -> sample position -> source taps) otherwise
  float depth_next = 0.f;
  {
    const int ry = tid / Geo::RW, rx = tid - ry * Geo::RW;
    const int px = c.x0 - Geo::HALO + rx, py = c.y0 - Geo::HALO + ry;
    if (tid < Geo::RN && !(px < -1 || px > W || py < -1 || py > H))
      depth_next = CDP_LDG(lv.depth + (size_t)c.b * plane + cdp_reflect(py, H) * W + cdp_reflect(px, W));
  }
#endif
  for (int idx = tid; idx < Geo::RN; idx += nthreads) {
    const int ry = idx / Geo::RW, rx = idx - ry * Geo::RW;
    const int px = c.x0 - Geo::HALO + rx, py = c.y0 - Geo::HALO + ry;
#if CDP_OPT_PREFETCH_DEPTH
    const float depth = depth_next;
    {
      const int nidx = idx + nthreads;
      const int nry = nidx / Geo::RW, nrx = nidx - nry * Geo::RW;
      const int npx = c.x0 - Geo::HALO + nrx, npy = c.y0 - Geo::HALO + nry;
      if (nidx < Geo::RN && !(npx < -1 || npx > W || npy < -1 || npy > H))
        depth_next = CDP_LDG(lv.depth + (size_t)c.b * plane + cdp_reflect(npy, H) * W + cdp_reflect(npx, W));
    }
#endif
    if (px < -1 || px > W || py < -1 || py > H) continue;  // never read
    const int u = cdp_reflect(px, W), v = cdp_reflect(py, H);
    const int pix = v * W + u;
#if !CDP_OPT_PREFETCH_DEPTH
    const float depth = CDP_LDG(lv.depth + (size_t)c.b * plane + pix);
#endif
    float m0[3], m1[3];
    if (M) {  // object-motion maps (make_sflow): added to the transformed point
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        m0[ch] = CDP_LDG(lv.mot0 + ((size_t)c.b * 3 + ch) * plane + pix);
        m1[ch] = CDP_LDG(lv.mot1 + ((size_t)c.b * 3 + ch) * plane + pix);
      }
    }
    CdpWarp w0, w1;
    cdp_warp_point((float)u, (float)v, depth, cam, T[0], M ? m0 : nullptr, w0);
    cdp_warp_point((float)u, (float)v, depth, cam, T[1], M ? m1 : nullptr, w1);
    CdpTaps t0, t1;
    cdp_taps(u, v, w0, W, H, t0);
    cdp_taps(u, v, w1, W, H, t1);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float cc = centre[ch];
      sm[(Geo::P_TGT + ch) * Geo::RN + idx] = CDP_LDG(tgt + ch * plane + pix) - cc;
      float2 s, wv;
      s.x = CDP_LDG(src0 + ch * plane + pix) - cc;
      s.y = CDP_LDG(src1 + ch * plane + pix) - cc;
      wv.x = cdp_bilinear(src0 + ch * plane, t0) - cc;
      wv.y = cdp_bilinear(src1 + ch * plane, t1) - cc;
      cdp_pair_plane<G>(sm, Geo::P_SRC, ch)[idx] = s;
      cdp_pair_plane<G>(sm, Geo::P_WARP, ch)[idx] = wv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Strip walk helpers (phases B1 / B2)
// ------------------------------------------------------------------------------------------
// Horizontal 3-tap sums of one window row for a candidate pair sharing the target row sums.
struct CdpRowPair {
  float2 s, ss, sy;  // sum x, sum x^2, sum x*y   (lane = candidate)
};
struct CdpRowTgt {
  float s, ss;  // sum y, sum y^2
};

CDP_HD void cdp_row_tgt(const float y[3], CdpRowTgt& o) {
  o.s = y[0] + y[1] + y[2];
  o.ss = y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
}
CDP_HD void cdp_row_pair(const float2 x[3], const float y[3], CdpRowPair& o) {
  o.s = cdp_add2(cdp_add2(x[0], x[1]), x[2]);
  o.ss = cdp_fma2(x[2], x[2], cdp_fma2(x[1], x[1], cdp_mul2(x[0], x[0])));
  o.sy = cdp_fma2(x[2], cdp_set2(y[2]), cdp_fma2(x[1], cdp_set2(y[1]), cdp_mul2(x[0], cdp_set2(y[0]))));
}

// SSIM loss of a candidate pair from the 3x3 sums of strip-centred values (algos/depth.py:141-153).
// ct = constant that turns strip-centred values back into true image values (means only).
CDP_HD float2 cdp_ssim_pair_loss(float2 sx, float2 sxx, float2 sxy, float sy, float syy, float ct) {
  const float ninth = 1.0f / 9.0f;
  const float2 n9 = cdp_set2(ninth);
  const float2 mxc = cdp_mul2(sx, n9);
  const float myc = sy * ninth;
  const float vy_c2 = (syy * ninth - myc * myc) + CDP_SSIM_C2;
  const float my = myc + ct;
  const float myy_c1 = my * my + CDP_SSIM_C1;
  const float2 mx = cdp_add2(mxc, cdp_set2(ct));
  const float2 vx = cdp_fma2(cdp_mul2(mxc, mxc), cdp_set2(-1.0f), cdp_mul2(sxx, n9));
  const float2 cov = cdp_fma2(mxc, cdp_set2(-myc), cdp_mul2(sxy, n9));
  const float2 n1 = cdp_fma2(mx, cdp_set2(2.0f * my), cdp_set2(CDP_SSIM_C1));
  const float2 n2 = cdp_fma2(cov, cdp_set2(2.0f), cdp_set2(CDP_SSIM_C2));
  const float2 d1 = cdp_fma2(mx, mx, cdp_set2(myy_c1));
  const float2 d2 = cdp_add2(vx, cdp_set2(vy_c2));
  const float2 num = cdp_mul2(n1, n2), den = cdp_mul2(d1, d2);
  float2 S;
  S.x = cdp_fdiv(num.x, den.x);
  S.y = cdp_fdiv(num.y, den.y);
  float2 l = cdp_fma2(S, cdp_set2(-0.5f), cdp_set2(0.5f));
  l.x = fminf(fmaxf(l.x, 0.f), 1.f);
  l.y = fminf(fmaxf(l.y, 0.f), 1.f);
  return l;
}

// ------------------------------------------------------------------------------------------
// Phase B1: candidate losses, noise, min / argmin (algos/depth.py:294-323)
// ------------------------------------------------------------------------------------------
template <bool G>
CDP_HD void cdp_photo_phase_b1(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm,
                               float& loss_acc) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  uint8_t* kplane = reinterpret_cast<uint8_t*>(sm + (size_t)Geo::NPLANES * Geo::RN);
  float centre[3];
  cdp_tile_centre(lv, c, centre);
  // (computed per thread on purpose: reading them from the parameter bank instead measured 1.7 % slower)
  const float a3 = p.alpha * (1.0f / 3.0f), b3 = (float)(1.0 - (double)p.alpha) * (1.0f / 3.0f);
  for (int item = tid; item < Geo::NITEMS; item += nthreads) {
    const int strip = item / Geo::BW, bx = item - strip * Geo::BW;
    const int by0 = strip * CDP_STRIP;
    const int qx = c.x0 - Geo::HB + bx;
    const int qy0 = c.y0 - Geo::HB + by0;
    // region index of the window centre of output row 0 of this strip
    const int r00 = (by0 + 1) * Geo::RW + bx + 1;
    float2 acc_id[CDP_STRIP], acc_pe[CDP_STRIP];
#pragma unroll
    for (int o = 0; o < CDP_STRIP; ++o) { acc_id[o] = cdp_set2(0.f); acc_pe[o] = cdp_set2(0.f); }
    const bool col_ok = qx >= 0 && qx < W;
    // tie-break noise of the strip's pixels, requested early so the loads overlap the strip walk
    float2 nz[CDP_STRIP];
#pragma unroll
    for (int o = 0; o < CDP_STRIP; ++o) {
      const int qy = qy0 + o;
      nz[o] = cdp_set2(0.f);
      if (lv.noise && col_ok && qy >= 0 && qy < H && by0 + o < Geo::BH) {
        const size_t nplane = (size_t)W * H;
        nz[o].x = CDP_LDG(lv.noise + ((size_t)c.b * 2 + 0) * nplane + qy * W + qx);
        nz[o].y = CDP_LDG(lv.noise + ((size_t)c.b * 2 + 1) * nplane + qy * W + qx);
      }
    }
    const int qc = qy0 + CDP_STRIP / 2 < 0 ? 0 : (qy0 + CDP_STRIP / 2 > H - 1 ? H - 1 : qy0 + CDP_STRIP / 2);
    const int cs_idx = (qc - c.y0 + Geo::HALO) * Geo::RW + bx + 1;
    if (col_ok && qy0 < H && qy0 + CDP_STRIP > 0) {
      // the channel loop stays rolled: unrolled, this phase alone is ~60 KB of SASS that every
      // warp streams through once per tile, and instruction fetch becomes the top stall reason
#pragma unroll 1
      for (int ch = 0; ch < 3; ++ch) {
        const float* ty = sm + (size_t)(Geo::P_TGT + ch) * Geo::RN;
        const float2* ts = cdp_pair_plane<G>(sm, Geo::P_SRC, ch);
        const float2* tw = cdp_pair_plane<G>(sm, Geo::P_WARP, ch);
        // strip constant: target value at the strip's middle pixel (already tile-centred),
        // clamped into the image so that it is always a staged value
        const float cs = ty[cs_idx];
        const float ct = cs + centre[ch];
        const float2 cs2 = cdp_set2(-cs);
        CdpRowTgt hy[3];
        CdpRowPair hs[3], hw[3];
        float yc_prev = 0.f;
        float2 sc_prev = cdp_set2(0.f), wc_prev = cdp_set2(0.f);
#pragma unroll
        for (int r = 0; r < CDP_STRIP + 2; ++r) {
          const int row = r00 + (r - 1) * Geo::RW;  // window row r-1 relative to output row 0
          float y[3];
          float2 s[3], w[3];
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            y[t] = ty[row + t - 1] - cs;
            s[t] = cdp_add2(ts[row + t - 1], cs2);
            w[t] = cdp_add2(tw[row + t - 1], cs2);
          }
          cdp_row_tgt(y, hy[r % 3]);
          cdp_row_pair(s, y, hs[r % 3]);
          cdp_row_pair(w, y, hw[r % 3]);
          if (r >= 2) {
            const int o = r - 2;  // output row: window rows r-2, r-1, r; centre row r-1
            const float sy = hy[0].s + hy[1].s + hy[2].s, syy = hy[0].ss + hy[1].ss + hy[2].ss;
            const float2 l_id = cdp_ssim_pair_loss(cdp_add2(cdp_add2(hs[0].s, hs[1].s), hs[2].s),
                                                   cdp_add2(cdp_add2(hs[0].ss, hs[1].ss), hs[2].ss),
                                                   cdp_add2(cdp_add2(hs[0].sy, hs[1].sy), hs[2].sy), sy, syy, ct);
            const float2 l_pe = cdp_ssim_pair_loss(cdp_add2(cdp_add2(hw[0].s, hw[1].s), hw[2].s),
                                                   cdp_add2(cdp_add2(hw[0].ss, hw[1].ss), hw[2].ss),
                                                   cdp_add2(cdp_add2(hw[0].sy, hw[1].sy), hw[2].sy), sy, syy, ct);
            float2 d_id = cdp_add2(sc_prev, cdp_set2(-yc_prev)), d_pe = cdp_add2(wc_prev, cdp_set2(-yc_prev));
            d_id.x = fabsf(d_id.x); d_id.y = fabsf(d_id.y);
            d_pe.x = fabsf(d_pe.x); d_pe.y = fabsf(d_pe.y);
            acc_id[o] = cdp_fma2(l_id, cdp_set2(a3), cdp_fma2(d_id, cdp_set2(b3), acc_id[o]));
            acc_pe[o] = cdp_fma2(l_pe, cdp_set2(a3), cdp_fma2(d_pe, cdp_set2(b3), acc_pe[o]));
          }
          yc_prev = y[1]; sc_prev = s[1]; wc_prev = w[1];
        }
      }
    }
    // min-reprojection with identity auto-mask for the strip's pixels
#pragma unroll
    for (int o = 0; o < CDP_STRIP; ++o) {
      const int by = by0 + o;
      if (by >= Geo::BH) break;
      const int qy = qy0 + o;
      const int ridx = r00 + o * Geo::RW;
      if (!col_ok || qy < 0 || qy >= H) {
        if (G) kplane[ridx] = 255;
        continue;
      }
      float n0 = nz[o].x, n1 = nz[o].y;
      if (!lv.noise) cdp_noise_pair(p.seed, (uint32_t)(qy * W + qx), (uint32_t)c.lvl, (uint32_t)c.b, n0, n1);
      const float id0 = acc_id[o].x + n0 * CDP_NOISE_SCALE, id1 = acc_id[o].y + n1 * CDP_NOISE_SCALE;
      float best = acc_pe[o].x;
      int kb = 0;
      if (acc_pe[o].y < best) { best = acc_pe[o].y; kb = 1; }
      if (id0 < best) { best = id0; kb = 2; }
      if (id1 < best) { best = id1; kb = 3; }
      const bool in_tile = qx >= c.x0 && qx < c.x0 + CDP_TILE_X && qy >= c.y0 && qy < c.y0 + CDP_TILE_Y;
      if (in_tile) {
        loss_acc += best;
        if (lv.argmin) lv.argmin[(size_t)c.b * W * H + qy * W + qx] = (uint8_t)kb;
      }
      if (G) kplane[ridx] = (uint8_t)kb;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Phase B2 (with grad): SSIM adjoint coefficients of the winning reprojection, in the frame of
// the tile-centred values:  d loss(q) / d x(p) = m/9 * (A + 2 x(p) B + y(p) C).
// ------------------------------------------------------------------------------------------
CDP_HD void cdp_photo_phase_b2(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm) {
  typedef CdpTileGeom<true> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const uint8_t* kplane = reinterpret_cast<const uint8_t*>(sm + (size_t)Geo::NPLANES * Geo::RN);
  float centre[3];
  cdp_tile_centre(lv, c, centre);
  for (int item = tid; item < Geo::NITEMS; item += nthreads) {
    const int strip = item / Geo::BW, bx = item - strip * Geo::BW;
    const int by0 = strip * CDP_STRIP;
    const int r00 = (by0 + 1) * Geo::RW + bx + 1;
    int kk[CDP_STRIP];
    bool any = false;
#pragma unroll
    for (int o = 0; o < CDP_STRIP; ++o) {
      kk[o] = (by0 + o < Geo::BH) ? kplane[r00 + o * Geo::RW] : 255;
      any = any || kk[o] < 2;
    }
    if (!any) continue;  // every pixel of the strip is auto-masked or outside the image
    const int qy0 = c.y0 - Geo::HB + by0;
    const int qc = qy0 + CDP_STRIP / 2 < 0 ? 0 : (qy0 + CDP_STRIP / 2 > lv.H - 1 ? lv.H - 1 : qy0 + CDP_STRIP / 2);
    const int cs_idx = (qc - c.y0 + Geo::HALO) * Geo::RW + bx + 1;
#pragma unroll 1
    for (int ch = 0; ch < 3; ++ch) {
      const float* ty = sm + (size_t)(Geo::P_TGT + ch) * Geo::RN;
      const float2* tw = cdp_pair_plane<true>(sm, Geo::P_WARP, ch);
      const float cs = ty[cs_idx];
      const float ct = cs + centre[ch];
      const float2 cs2 = cdp_set2(-cs);
      CdpRowTgt hy[3];
      CdpRowPair hw[3];
#pragma unroll
      for (int r = 0; r < CDP_STRIP + 2; ++r) {
        const int row = r00 + (r - 1) * Geo::RW;
        float y[3];
        float2 w[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          y[t] = ty[row + t - 1] - cs;
          w[t] = cdp_add2(tw[row + t - 1], cs2);
        }
        cdp_row_tgt(y, hy[r % 3]);
        cdp_row_pair(w, y, hw[r % 3]);
        if (r >= 2) {
          const int o = r - 2;
          if (kk[o] < 2) {
            const bool first = kk[o] == 0;
            const float ninth = 1.0f / 9.0f;
            const float2 sx2 = cdp_add2(cdp_add2(hw[0].s, hw[1].s), hw[2].s);
            const float2 sxx2 = cdp_add2(cdp_add2(hw[0].ss, hw[1].ss), hw[2].ss);
            const float2 sxy2 = cdp_add2(cdp_add2(hw[0].sy, hw[1].sy), hw[2].sy);
            const float mxc = (first ? sx2.x : sx2.y) * ninth;
            const float exx = (first ? sxx2.x : sxx2.y) * ninth;
            const float exy = (first ? sxy2.x : sxy2.y) * ninth;
            const float myc = (hy[0].s + hy[1].s + hy[2].s) * ninth;
            const float eyy = (hy[0].ss + hy[1].ss + hy[2].ss) * ninth;
            CdpSsimTerms t;
            cdp_ssim_terms(mxc, myc, exx, eyy, exy, ct, t);
            float A, B, C;
            // means in the tile-centred frame are mxc + cs, myc + cs
            cdp_ssim_coeffs_abc(t, mxc + cs, myc + cs, A, B, C);
            // packed layout: see CdpTileGeom::P_COEF
#if CDP_OPT_PACKED_GATHER
            float* cf = sm + (size_t)Geo::P_COEF * Geo::RN + 2 * (r00 + o * Geo::RW);
            if (ch < 2) {
              cf[0 * 2 * Geo::RN + ch] = A; cf[1 * 2 * Geo::RN + ch] = B; cf[2 * 2 * Geo::RN + ch] = C;
            } else {
              cf[3 * 2 * Geo::RN + 0] = A; cf[3 * 2 * Geo::RN + 1] = B; cf[4 * 2 * Geo::RN + 0] = C;
            }
#else
            const int ridx = r00 + o * Geo::RW;
            sm[(size_t)(Geo::P_COEF + ch * 3 + 0) * Geo::RN + ridx] = A;
            sm[(size_t)(Geo::P_COEF + ch * 3 + 1) * Geo::RN + ridx] = B;
            sm[(size_t)(Geo::P_COEF + ch * 3 + 2) * Geo::RN + ridx] = C;
#endif
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Phase C (with grad)
// ------------------------------------------------------------------------------------------
template <bool M>
CDP_HD void cdp_photo_phase_c(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads,
                              const float* sm, float* dT /*[32]*/) {
  typedef CdpTileGeom<true> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  const size_t plane = (size_t)W * H;
  const CdpCam cam = cdp_tile_cam(p, c);
  const uint8_t* kplane = reinterpret_cast<const uint8_t*>(sm + (size_t)Geo::NPLANES * Geo::RN);
  const float w_ssim = p.alpha / 27.0f;  // alpha * (1/3 channels) * (1/9 window)
  const float w_l1 = (float)(1.0 - (double)p.alpha) * (1.0f / 3.0f);
  for (int idx = tid; idx < CDP_TILE_X * CDP_TILE_Y; idx += nthreads) {
    const int ly = idx / CDP_TILE_X, lx = idx - ly * CDP_TILE_X;
    const int px = c.x0 + lx, py = c.y0 + ly;
    if (px >= W || py >= H) continue;
    const int ridx = (ly + Geo::HALO) * Geo::RW + lx + Geo::HALO;
    const int kown = kplane[ridx];
    // winners and reflection multiplicities of the 3x3 neighbourhood
    int kn[9];
    float mn[9];
    bool any0 = kown == 0, any1 = kown == 1;
    const bool interior = px >= 2 && px <= W - 3 && py >= 2 && py <= H - 3;
    if (interior) {  // no reflection in reach: weights are 1
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        mn[j] = 1.f;
        kn[j] = (int)kplane[ridx + (j / 3 - 1) * Geo::RW + (j % 3 - 1)];
        any0 = any0 || kn[j] == 0;
        any1 = any1 || kn[j] == 1;
      }
    } else {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const float my = cdp_reflect_mult(py, dy, H);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int j = (dy + 1) * 3 + dx + 1;
          mn[j] = my * cdp_reflect_mult(px, dx, W);
          kn[j] = mn[j] != 0.f ? (int)kplane[ridx + dy * Geo::RW + dx] : 255;
          any0 = any0 || kn[j] == 0;
          any1 = any1 || kn[j] == 1;
        }
      }
    }
    float gd = 0.f;
    if (M) {  // dL/d motion: zero unless the source contributes below
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (!any0) lv.gmot0[((size_t)c.b * 3 + ch) * plane + py * W + px] = 0.f;
        if (!any1) lv.gmot1[((size_t)c.b * 3 + ch) * plane + py * W + px] = 0.f;
      }
    }
    if (any0 || any1) {
      const float depth = CDP_LDG(lv.depth + (size_t)c.b * plane + py * W + px);
#pragma unroll 1
      for (int k = 0; k < 2; ++k) {
        if (!(k == 0 ? any0 : any1)) continue;
#if CDP_OPT_PACKED_GATHER
        // masked, reflection-weighted 3x3 sums of the packed coefficient planes
        float2 acc[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) acc[f] = cdp_set2(0.f);
        const float2* cf = reinterpret_cast<const float2*>(sm + (size_t)Geo::P_COEF * Geo::RN);
        if (interior) {
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            if (kn[j] != k) continue;
            const int n = ridx + (j / 3 - 1) * Geo::RW + (j % 3 - 1);
#pragma unroll
            for (int f = 0; f < 5; ++f) acc[f] = cdp_add2(acc[f], cf[(size_t)f * Geo::RN + n]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            if (kn[j] != k) continue;
            const int n = ridx + (j / 3 - 1) * Geo::RW + (j % 3 - 1);
            const float2 m2 = cdp_set2(mn[j]);
#pragma unroll
            for (int f = 0; f < 5; ++f) acc[f] = cdp_fma2(cf[(size_t)f * Geo::RN + n], m2, acc[f]);
          }
        }
        const float sa[3] = {acc[0].x, acc[0].y, acc[3].x};
        const float sb[3] = {acc[1].x, acc[1].y, acc[3].y};
        const float sc[3] = {acc[2].x, acc[2].y, acc[4].x};
#else
        float sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f}, sc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          if (kn[j] != k) continue;
          const int n = ridx + (j / 3 - 1) * Geo::RW + (j % 3 - 1);
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            sa[ch] += mn[j] * sm[(size_t)(Geo::P_COEF + ch * 3 + 0) * Geo::RN + n];
            sb[ch] += mn[j] * sm[(size_t)(Geo::P_COEF + ch * 3 + 1) * Geo::RN + n];
            sc[ch] += mn[j] * sm[(size_t)(Geo::P_COEF + ch * 3 + 2) * Geo::RN + n];
          }
        }
#endif
        // (the source loop is rolled to keep the code small: no register arrays indexed by k)
        CdpPose T;
        cdp_load_pose_aligned((k == 0 ? p.pose0 : p.pose1) + (size_t)c.b * 16, T);
        const float* srck = (k == 0 ? lv.src0 : lv.src1) + (size_t)c.b * 3 * plane;
        float mo[3], gmo[3];
        const float* motk = k == 0 ? lv.mot0 : lv.mot1;
        if (M) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) mo[ch] = CDP_LDG(motk + ((size_t)c.b * 3 + ch) * plane + py * W + px);
        }
        CdpWarp w;
        cdp_warp_point((float)px, (float)py, depth, cam, T, M ? mo : nullptr, w);
        CdpTaps t;
        cdp_taps(px, py, w, W, H, t);
        float gix = 0.f, giy = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float2 xv = cdp_pair_plane<true>(sm, Geo::P_WARP, ch)[ridx];
          const float x = k == 0 ? xv.x : xv.y;
          const float y = sm[(size_t)(Geo::P_TGT + ch) * Geo::RN + ridx];
          float gw = w_ssim * (sa[ch] + 2.f * x * sb[ch] + y * sc[ch]);
          if (kown == k) gw += w_l1 * (x > y ? 1.f : (x < y ? -1.f : 0.f));
          gw *= lv.weight;
          float dix, diy;
          cdp_bilinear_grad(srck + ch * plane, t, dix, diy);
          gix += gw * dix;
          giy += gw * diy;
        }
#if CDP_OPT_DIRECT_DT
        float* gmk = M ? gmo : nullptr;
        if (k == 0) cdp_warp_adjoint(gix * t.mx, giy * t.my, w, cam, T, gd, dT, gmk);
        else cdp_warp_adjoint(gix * t.mx, giy * t.my, w, cam, T, gd, dT + 16, gmk);
        if (M) {
          float* dst = (k == 0 ? lv.gmot0 : lv.gmot1) + (size_t)c.b * 3 * plane + py * W + px;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) dst[ch * plane] = gmo[ch];
        }
#else
        float dTk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) dTk[i] = 0.f;
        float* gmk = M ? gmo : nullptr;
        cdp_warp_adjoint(gix * t.mx, giy * t.my, w, cam, T, gd, dTk, gmk);
        if (M) {
          float* dst = (k == 0 ? lv.gmot0 : lv.gmot1) + (size_t)c.b * 3 * plane + py * W + px;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) dst[ch * plane] = gmo[ch];
        }
        if (k == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dT[i] += dTk[i];
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) dT[16 + i] += dTk[i];
        }
#endif
      }
    }
    lv.gdepth[(size_t)c.b * plane + py * W + px] = gd;
  }
}
