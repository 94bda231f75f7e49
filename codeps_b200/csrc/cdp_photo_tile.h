// cdp_photo_tile.h -- the fused photometric tile kernel (host/device bodies, phase by phase).
//
// One CTA = one 32x32 tile of one pyramid level of one sample.  Phases (a block barrier between
// consecutive ones):
//   S   stage dense boxes in shared memory: target (tile + 2), depth (tile + 2) and both source
//       frames (tile + 2 + a gather margin of 4).  On the GPU these are four TMA box loads
//       (cp.async.bulk.tensor.3d, zero fill outside the image) issued by one thread and awaited on
//       an mbarrier (cdp_api.cu); levels whose row pitch is not a multiple of 16 bytes -- and the
//       CPU emulator -- fill the same boxes with cdp_photo_stage (plain loads).
//   A   reflect the one-pixel ring outside the image inside shared memory (border tiles only);
//       warp both source frames for every staged position with the two sources in the two lanes
//       of packed fp32, bilinear taps served from the staged source boxes (global loads only for
//       taps that leave the margin), warped values -> shared memory as (source 0, source 1) pairs.
//   B1  min-reprojection: each thread owns a vertical strip of 5 pixels in one column and
//       slides a 3-row window down it (separable 3x3 sums), all four candidates at once, the two
//       candidates of a pair in the two lanes of packed fp32 (FFMA2).  -> candidate losses,
//       tie-break noise, min / argmin, loss partial, argmin map, winner plane.
//   B2  (with grad) same walk over the reprojection pair only, SSIM adjoint coefficients of the
//       winner -> shared memory (the source boxes are dead by then and are reused).
//   C1  (with grad) gather the adjoint over the reflected 3x3 neighbourhood, add the L1 term:
//       dL/d warped value per channel and source, stored over the pixel's own warped values.
//   S2  the two source boxes are staged again (TMA) over the dead coefficient planes.
//   C2  (with grad) chain through the bilinear sampler (taps from the boxes) and the projection to
//       depth and pose, both sources in the two lanes of packed fp32.
// Reference: algos/depth.py:221-237, 272-325 (ReconstructionLoss), 128-155 (SSIMLoss);
// misc/image_warper.py:100-184.
#pragma once

#include "cdp_math.h"

#ifndef CDP_OPT_DIRECT_DT
#define CDP_OPT_DIRECT_DT 1  // phase C accumulates dL/dT without a per-pixel temporary
#endif
#ifndef CDP_STRIP
#define CDP_STRIP 5  // pixels per thread strip in phases B1/B2
#endif
#ifndef CDP_OPT_SSIM_RAW
#define CDP_OPT_SSIM_RAW 1  // phase B1: SSIM ratio from the raw window sums (9 instead of 13 packed operations per pair)
#endif
#ifndef CDP_OPT_SSIM_RAW_B2
#define CDP_OPT_SSIM_RAW_B2 1  // phase B2: adjoint coefficients from the raw window sums
#endif
#ifndef CDP_OPT_NOISE_LATE
#define CDP_OPT_NOISE_LATE 1  // phase B1: 1 = tie-break noise loaded after the strip walk instead of before it
#endif
#ifndef CDP_EXP_A_SKIP_TAIL
#define CDP_EXP_A_SKIP_TAIL 0  // timing experiment only (wrong results): phase A drops its ragged last iteration
#endif
#ifndef CDP_EXP_HALO1
#define CDP_EXP_HALO1 0  // timing experiment only (wrong results): the gradient instantiation computes on halo 1 like the
                         // forward-only one -- the upper bound of what sharing halos between the CTAs of a cluster could save
#endif
#ifndef CDP_A_UNROLL
#define CDP_A_UNROLL 1  // unroll factor of the position loop of phase A
#endif
#ifndef CDP_C2_UNROLL
#define CDP_C2_UNROLL 1  // unroll factor of the pixel loop of phase C2
#endif
#ifndef CDP_OPT_PAIR_SHARE
#define CDP_OPT_PAIR_SHARE 1  // phase B1: consecutive outputs of a strip share the sum of their two common window rows
#endif
#ifndef CDP_OPT_INTERIOR
#define CDP_OPT_INTERIOR 1  // tiles whose source boxes lie inside the image skip the reflection / border-clip logic
#endif
#ifndef CDP_SRC_MARGIN
#define CDP_SRC_MARGIN 4  // gather margin of the staged source boxes (pixels beyond the target box)
#endif

// Shared-memory geometry.  All planes except the source boxes are TBW x TBH with origin
// (x0 - TXO, y0 - TYO), whatever the halo the instantiation computes on, so that one set of TMA
// descriptors serves both; "t-index" = ty * TBW + tx.  The source boxes are SBW x SBH with origin
// (x0 - TXO - SBM, y0 - TYO - SBM); "s-index" = (ty + SBM) * SBW + tx + SBM.  A TMA box must start
// at a multiple of 16 bytes in the innermost dimension (measured: an x coordinate that is not a
// multiple of 4 floats faults with "illegal instruction", tools/experiments/tma_probe.cu), hence
// TXO = 4 and SBM = 4 although the window sums only need two columns left of the tile.
template <bool G>
struct CdpTileGeom {
  static constexpr int HALO = (G && !CDP_EXP_HALO1) ? 2 : 1;  // ring around the tile on which warped values are needed
  static constexpr int TXO = 4, TYO = 2;  // box position of the tile's first column / row
  static constexpr int OFFX = TXO - HALO, OFFY = TYO - HALO;  // first box column / row of that region
  static constexpr int RW = CDP_TILE_X + 2 * HALO, RH = CDP_TILE_Y + 2 * HALO, RN = RW * RH;
  static constexpr int HB = HALO - 1;     // ring on which losses / argmin / coefficients are needed
  static constexpr int BW = CDP_TILE_X + 2 * HB, BH = CDP_TILE_Y + 2 * HB;
  static constexpr int NSTRIP = (BH + CDP_STRIP - 1) / CDP_STRIP;
  static constexpr int NITEMS = BW * NSTRIP;  // (column, strip) work items of phases B1/B2
  static constexpr int TBW = CDP_TILE_X + 2 * TXO, TBH = CDP_TILE_Y + 2 * TYO, TBN = TBW * TBH;
  static constexpr int SBM = CDP_SRC_MARGIN;
  static constexpr int SBW = TBW + 2 * SBM, SBH = TBH + 2 * SBM, SBN = SBW * SBH;
  static constexpr int align32(int n) { return (n + 31) & ~31; }  // TMA destinations: 128-byte aligned
  // float offsets
  static constexpr int O_TGT = 0;                                // [3][TBN] target
  static constexpr int O_DEPTH = align32(O_TGT + 3 * TBN);       // [TBN]
  static constexpr int O_WARP = align32(O_DEPTH + TBN);          // 3 float2 planes [TBN]: warped (source 0, source 1)
  static constexpr int O_SRC = align32(O_WARP + 6 * TBN);        // 2 boxes [3][SBN], SRC_STRIDE apart
  static constexpr int SRC_STRIDE = align32(3 * SBN);
  // with grad, after B1: adjoint coefficients of the winner, 9 planes [TBN] (A, B, C per channel)
  static constexpr int O_COEF = O_SRC;
  static constexpr int SRC_FLOATS = (SRC_STRIDE + 3 * SBN > 9 * TBN) ? SRC_STRIDE + 3 * SBN : 9 * TBN;
  static constexpr int O_K = align32(O_SRC + SRC_FLOATS);        // winner bytes [TBN]
  static constexpr int O_MBAR = O_K + align32((TBN + 3) / 4);    // two 8-byte mbarriers
  static constexpr size_t SMEM_BYTES = (size_t)(O_MBAR + 4) * sizeof(float);
  // barrier 0: depth + both source boxes (what phase A reads; re-armed for the sources before C2);
  // barrier 1: target box (first read in phase B1)
  static constexpr unsigned TMA_BYTES_SRC = (unsigned)(2 * 3 * SBN * sizeof(float));  // both source boxes
  static constexpr unsigned TMA_BYTES_A = (unsigned)(TBN * sizeof(float)) + TMA_BYTES_SRC;
  static constexpr unsigned TMA_BYTES_TGT = (unsigned)(3 * TBN * sizeof(float));
  // The B1/B2 strip walk always reads CDP_STRIP + 2 rows, also for the last, partial strip: the
  // rows past the box belong to the following plane (values discarded); in the source boxes they
  // must stay inside the margin rows.
  static constexpr int LAST_ROW = NSTRIP * CDP_STRIP + OFFY + 1;  // last box row a strip walk reads
  static_assert(LAST_ROW + SBM <= SBH - 1, "strip over-read leaves the source box: pick CDP_TILE_Y / CDP_STRIP so that it fits");
  static_assert((LAST_ROW + 1) * TBW <= 2 * TBN, "strip over-read leaves the shared-memory allocation");
  static_assert(SBM % 4 == 0 && TXO % 4 == 0 && TBW % 4 == 0 && SBW % 4 == 0 && CDP_TILE_X % 4 == 0,
                "TMA boxes must start and end at multiples of 16 bytes");
  static_assert(TXO >= 2 && TYO >= 2, "the window sums of the gradient instantiation need two pixels around the tile");
};

struct CdpTileCtx {
  int lvl, b, x0, y0;  // level, sample, tile origin
};

CDP_HD CdpTileCtx cdp_tile_ctx(const CdpPhotoParams& p, int bx, int by) {
  CdpTileCtx c;
  // level = number of levels s >= 1 that start at or before this block (branch-free: the loads of
  // all block_begin values are independent; this runs on the critical path to the tile's TMA loads)
  int s = 0;
#pragma unroll
  for (int l = 1; l < CDP_MAX_LEVELS; ++l) s += (l < p.num_levels && bx >= p.lv[l].block_begin) ? 1 : 0;
  c.lvl = s;
  const int tile = bx - p.lv[s].block_begin;
  // tile / tiles_x by the host's reciprocal (tiles_x == 1: the reciprocal wraps to 0, ty = tile)
  const int tx_n = p.lv[s].tiles_x;
  const int ty = tx_n == 1 ? tile : (int)(((uint64_t)(uint32_t)tile * p.lv[s].tiles_x_rcp) >> 32);
  c.x0 = (tile - ty * tx_n) * CDP_TILE_X;
  c.y0 = ty * CDP_TILE_Y;
  c.b = p.batch_begin + by;
  return c;
}

CDP_HD CdpCam cdp_tile_cam(const CdpPhotoParams& p, const CdpTileCtx& c) {
  // per-level intrinsics table (rows are 16-byte aligned)
  const float4 k = CDP_LDG(reinterpret_cast<const float4*>(p.K_tab) + (size_t)c.lvl * p.batch_total + c.b);
  return cdp_make_cam(k.x, k.y, k.z, k.w);
}

// Per-tile constants of the warp: intrinsics of (level, sample) and both poses, lane-packed.  Loaded
// by the kernel BEFORE it waits for the TMA boxes, so that these global-memory round trips overlap
// the box loads instead of following them.
struct CdpTileConst {
  CdpCam cam;
  CdpPose2 T;
};
CDP_HD void cdp_tile_const(const CdpPhotoParams& p, const CdpTileCtx& c, CdpTileConst& k) {
  k.cam = cdp_tile_cam(p, c);
  CdpPose t0, t1;
  cdp_load_pose_aligned(p.pose0 + (size_t)c.b * 16, t0);  // cdp_photo_fwd requires 16-byte aligned poses
  cdp_load_pose_aligned(p.pose1 + (size_t)c.b * 16, t1);
  cdp_pack_pose(t0, t1, k.T);
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

// Per-tile, per-channel constant (target value at the tile centre, read from the staged box) on
// which the SSIM adjoint coefficients and the values they multiply are centred in phases B2 / C:
// the three terms A + 2 x B + y C cancel almost completely, and the cancellation costs fewer
// digits the smaller |x|, |y| are.
template <bool G>
CDP_HD void cdp_tile_centre(const CdpLevel& lv, const CdpTileCtx& c, const float* sm, float centre[3]) {
  typedef CdpTileGeom<G> Geo;
  const int cx = c.x0 + CDP_TILE_X / 2 < lv.W ? CDP_TILE_X / 2 : lv.W - 1 - c.x0;
  const int cy = c.y0 + CDP_TILE_Y / 2 < lv.H ? CDP_TILE_Y / 2 : lv.H - 1 - c.y0;
  const int ti = (cy + Geo::TYO) * Geo::TBW + cx + Geo::TXO;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) centre[ch] = sm[Geo::O_TGT + ch * Geo::TBN + ti];
}

template <bool G>
CDP_HD float2* cdp_warp_plane(float* sm, int ch) {
  return reinterpret_cast<float2*>(sm + CdpTileGeom<G>::O_WARP + 2 * ch * CdpTileGeom<G>::TBN);
}
template <bool G>
CDP_HD const float2* cdp_warp_plane(const float* sm, int ch) {
  return reinterpret_cast<const float2*>(sm + CdpTileGeom<G>::O_WARP + 2 * ch * CdpTileGeom<G>::TBN);
}

// ------------------------------------------------------------------------------------------
// Phase S without TMA: fill the boxes with plain loads, zeros outside the image (what the TMA
// box loads do in hardware).
// ------------------------------------------------------------------------------------------
CDP_HD void cdp_stage_box(const float* g, size_t plane, int nplanes, float* dst, int bw, int bh, int ox, int oy,
                          int W, int H, int tid, int nthreads) {
  const int bn = bw * bh;
#pragma unroll 1
  for (int e = tid; e < nplanes * bn; e += nthreads) {
    const int pl = e / bn, r = e - pl * bn;
    const int ty = r / bw, tx = r - ty * bw;
    const int x = ox + tx, y = oy + ty;
    dst[e] = (x >= 0 && x < W && y >= 0 && y < H) ? CDP_LDG(g + pl * plane + (size_t)y * W + x) : 0.f;
  }
}

template <bool G>
CDP_HD void cdp_photo_stage(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const size_t plane = (size_t)lv.W * lv.H;
  const int ox = c.x0 - Geo::TXO, oy = c.y0 - Geo::TYO;
  cdp_stage_box(lv.tgt + (size_t)c.b * 3 * plane, plane, 3, sm + Geo::O_TGT, Geo::TBW, Geo::TBH, ox, oy, lv.W, lv.H,
                tid, nthreads);
  cdp_stage_box(lv.depth + (size_t)c.b * plane, plane, 1, sm + Geo::O_DEPTH, Geo::TBW, Geo::TBH, ox, oy, lv.W, lv.H,
                tid, nthreads);
  cdp_stage_box(lv.src0 + (size_t)c.b * 3 * plane, plane, 3, sm + Geo::O_SRC, Geo::SBW, Geo::SBH, ox - Geo::SBM,
                oy - Geo::SBM, lv.W, lv.H, tid, nthreads);
  cdp_stage_box(lv.src1 + (size_t)c.b * 3 * plane, plane, 3, sm + Geo::O_SRC + Geo::SRC_STRIDE, Geo::SBW, Geo::SBH,
                ox - Geo::SBM, oy - Geo::SBM, lv.W, lv.H, tid, nthreads);
}

// ------------------------------------------------------------------------------------------
// Phase A
// ------------------------------------------------------------------------------------------
// nn.ReflectionPad2d(1) of a staged box: the ring one pixel outside the image takes the value of
// its mirror pixel (-1 <- 1, n <- n-2); sources are interior pixels, so no element is both read
// and written.  bn = plane stride of the box stack.
CDP_HD void cdp_reflect_ring(float* box, int nplanes, int bn, int bw, int bh, int ox, int oy, int W, int H, int tid,
                             int nthreads) {
#pragma unroll 1
  for (int e = tid; e < 2 * bw; e += nthreads) {  // rows -1 and H (corners included)
    const int far = e >= bw ? 1 : 0, tx = e - far * bw;
    const int y = far ? H : -1, ty = y - oy, x = ox + tx;
    if (ty < 0 || ty >= bh || x < -1 || x > W) continue;
    const int s = (cdp_reflect(y, H) - oy) * bw + cdp_reflect(x, W) - ox;
#pragma unroll 1
    for (int pl = 0; pl < nplanes; ++pl) box[pl * bn + ty * bw + tx] = box[pl * bn + s];
  }
#pragma unroll 1
  for (int e = tid; e < 2 * bh; e += nthreads) {  // columns -1 and W, image rows only
    const int far = e >= bh ? 1 : 0, ty = e - far * bh;
    const int x = far ? W : -1, tx = x - ox, y = oy + ty;
    if (tx < 0 || tx >= bw || y < 0 || y > H - 1) continue;
    const int s = ty * bw + cdp_reflect(x, W) - ox;
#pragma unroll 1
    for (int pl = 0; pl < nplanes; ++pl) box[pl * bn + ty * bw + tx] = box[pl * bn + s];
  }
}

// One source's literal-formula warp (depth clamp active or Q_w <= 0): (dx, dy, ix, iy) of source k.
// Cold path, out of line: everything travels by value and the pose is re-read from global memory, so
// that no variable of the hot loop has its address taken.
CDP_COLD float4 cdp_warp_lane_literal(int k, float u, float v, float depth, CdpCam cam, const float* pose /*[16]*/,
                                     bool with_motion, float m0, float m1, float m2) {
  CdpPose Tk;
  cdp_load_pose_aligned(pose, Tk);
  float mk[3] = {m0, m1, m2};
  CdpWarp ws;
  cdp_warp_point(u, v, depth, cam, Tk, with_motion ? mk : nullptr, ws);
  float4 r;
  r.x = ws.dx; r.y = ws.dy; r.z = ws.ix; r.w = ws.iy;
  return r;
}

// bilinear blend of (source 0, source 1) tap pairs with per-lane fractions
CDP_HD float2 cdp_lerp2(float2 nw, float2 ne, float2 sw, float2 se, float2 fx, float2 fy) {
  const float2 neg1 = cdp_set2(-1.0f);
  const float2 top = cdp_fma2(fx, cdp_fma2(nw, neg1, ne), nw);
  const float2 bot = cdp_fma2(fx, cdp_fma2(sw, neg1, se), sw);
  return cdp_fma2(fy, cdp_fma2(top, neg1, bot), top);
}

// Tap cell of one axis when the border clip is known to be inactive (see cdp_tile_interior):
// box-relative lower tap index and fraction; the same values cdp_tap_axis_full produces there.
CDP_HD void cdp_tap_axis_interior(int rel, float d, int& b0, float& frac) {
  const float fl = floorf(d);
  b0 = rel + (int)fl;
  frac = d - fl;
}

// A tile is "interior" when its source boxes lie inside [1, n-2] on both axes.  Then every staged
// position is an image pixel (no reflection), and a 2x2 footprint that lies inside the box has its
// lower tap in [1, n-3], i.e. the sample position is strictly inside (0, n-1): the border clip of
// grid_sample and the clamps of cdp_tap_axis_full are no-ops and the gradient mask is 1.  Footprints
// that leave the box (and NaN displacements, whose integer conversion is 0) take the general path.
template <bool G>
CDP_HD bool cdp_tile_interior(const CdpLevel& lv, const CdpTileCtx& c) {
  typedef CdpTileGeom<G> Geo;
  const int box_x = c.x0 - Geo::TXO - Geo::SBM, box_y = c.y0 - Geo::TYO - Geo::SBM;
  return CDP_OPT_INTERIOR && box_x >= 1 && box_y >= 1 && box_x + Geo::SBW <= lv.W - 1 && box_y + Geo::SBH <= lv.H - 1;
}

// Both sources' tap cells relative to the source boxes; returns true when both footprints lie
// inside the boxes (interior tiles only).
template <bool G>
CDP_HD bool cdp_taps_interior(int relx, int rely, const CdpWarp2& w, int& bx0, int& by0, int& bx1, int& by1,
                              float2& fx, float2& fy) {
  typedef CdpTileGeom<G> Geo;
  cdp_tap_axis_interior(relx, w.dx.x, bx0, fx.x);
  cdp_tap_axis_interior(rely, w.dy.x, by0, fy.x);
  cdp_tap_axis_interior(relx, w.dx.y, bx1, fx.y);
  cdp_tap_axis_interior(rely, w.dy.y, by1, fy.y);
  const float2 t = cdp_add2(fx, fy);
  const bool finite = t.x + t.y >= 0.f;  // false for NaN
  return finite && (unsigned)bx0 <= (unsigned)(Geo::SBW - 2) && (unsigned)by0 <= (unsigned)(Geo::SBH - 2) &&
         (unsigned)bx1 <= (unsigned)(Geo::SBW - 2) && (unsigned)by1 <= (unsigned)(Geo::SBH - 2);
}

// Phase A, cold path: at least one of the two 2x2 footprints leaves the staged source box; that
// source's taps come from global memory.
template <bool G>
CDP_COLD void cdp_phase_a_taps_far(const float* sbox0, const float* sbox1, const float* src0, const float* src1,
                                   size_t plane, int W, bool in0, bool in1, int bx0, int by0, int bx1, int by1, int ax0,
                                   int ay0, int ax1, int ay1, float2 fx, float2 fy, float* sm, int ti) {
  typedef CdpTileGeom<G> Geo;
  // (all taps of the three channels requested before the first use: one memory round trip per
  // position instead of one per channel; the stores come last because the compiler keeps loads
  // behind earlier stores it cannot disambiguate)
  float2 nw[3], ne[3], sw[3], se[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (in0) {
      const float* q = sbox0 + ch * Geo::SBN + by0 * Geo::SBW + bx0;
      nw[ch].x = q[0]; ne[ch].x = q[1]; sw[ch].x = q[Geo::SBW]; se[ch].x = q[Geo::SBW + 1];
    } else {
      const float* g = src0 + ch * plane + (size_t)ay0 * W + ax0;
      nw[ch].x = CDP_LDG(g); ne[ch].x = CDP_LDG(g + 1); sw[ch].x = CDP_LDG(g + W); se[ch].x = CDP_LDG(g + W + 1);
    }
    if (in1) {
      const float* q = sbox1 + ch * Geo::SBN + by1 * Geo::SBW + bx1;
      nw[ch].y = q[0]; ne[ch].y = q[1]; sw[ch].y = q[Geo::SBW]; se[ch].y = q[Geo::SBW + 1];
    } else {
      const float* g = src1 + ch * plane + (size_t)ay1 * W + ax1;
      nw[ch].y = CDP_LDG(g); ne[ch].y = CDP_LDG(g + 1); sw[ch].y = CDP_LDG(g + W); se[ch].y = CDP_LDG(g + W + 1);
    }
  }
  float2 wv[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) wv[ch] = cdp_lerp2(nw[ch], ne[ch], sw[ch], se[ch], fx, fy);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) cdp_warp_plane<G>(sm, ch)[ti] = wv[ch];
}

template <bool G, bool M>
CDP_HD void cdp_photo_phase_a(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm,
                              const CdpTileConst& kc) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  const size_t plane = (size_t)W * H;
  // 1. reflected ring of the target and source boxes (read by the window sums of B1 / B2)
  const int ox = c.x0 - Geo::TXO, oy = c.y0 - Geo::TYO;  // image position of box element (0, 0)
  if (c.x0 == 0 || c.y0 == 0 || ox + Geo::TBW > W || oy + Geo::TBH > H) {
    cdp_reflect_ring(sm + Geo::O_TGT, 3, Geo::TBN, Geo::TBW, Geo::TBH, ox, oy, W, H, tid, nthreads);
    cdp_reflect_ring(sm + Geo::O_SRC, 3, Geo::SBN, Geo::SBW, Geo::SBH, ox - Geo::SBM, oy - Geo::SBM, W, H, tid, nthreads);
    cdp_reflect_ring(sm + Geo::O_SRC + Geo::SRC_STRIDE, 3, Geo::SBN, Geo::SBW, Geo::SBH, ox - Geo::SBM, oy - Geo::SBM,
                     W, H, tid, nthreads);
  }
  // 2. warp + gather
  const CdpCam& cam = kc.cam;
  const CdpPose2& T = kc.T;
  const float* src0 = lv.src0 + (size_t)c.b * 3 * plane;
  const float* src1 = lv.src1 + (size_t)c.b * 3 * plane;
  const float* sdepth = sm + Geo::O_DEPTH;
  const float* sbox0 = sm + Geo::O_SRC;
  const float* sbox1 = sm + Geo::O_SRC + Geo::SRC_STRIDE;
  const int box_x = ox - Geo::SBM, box_y = oy - Geo::SBM;  // image position of source box element (0, 0)
  const bool interior = cdp_tile_interior<G>(lv, c);
  constexpr int kUnrollA = CDP_A_UNROLL;
#pragma unroll kUnrollA
  for (int idx = tid; idx < (CDP_EXP_A_SKIP_TAIL ? Geo::RN / 256 * 256 : Geo::RN); idx += nthreads) {
    const int ry = idx / Geo::RW, rx = idx - ry * Geo::RW;
    const int tx = rx + Geo::OFFX, ty = ry + Geo::OFFY;
    const int px = ox + tx, py = oy + ty;
    int u = px, v = py;
    if (!interior) {
      if (px < -1 || px > W || py < -1 || py > H) continue;  // never read
      u = cdp_reflect(px, W); v = cdp_reflect(py, H);
    }
    const float depth = sdepth[(v - oy) * Geo::TBW + u - ox];
    float2 mo[3];
    if (M) {  // object-motion maps (make_sflow): added to the transformed point
      const int pix = v * W + u;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        mo[ch].x = CDP_LDG(lv.mot0 + ((size_t)c.b * 3 + ch) * plane + pix);
        mo[ch].y = CDP_LDG(lv.mot1 + ((size_t)c.b * 3 + ch) * plane + pix);
      }
    }
    CdpWarp2 w;
    cdp_warp_point2((float)u, (float)v, depth, cam, T, M ? mo : nullptr, w);
    if (!(w.regular[0] && w.regular[1])) {  // depth clamp active or Q_w <= 0: literal formula, per source
      if (!w.regular[0]) {
        const float4 r = cdp_warp_lane_literal(0, (float)u, (float)v, depth, cam, p.pose0 + (size_t)c.b * 16, M,
                                               M ? mo[0].x : 0.f, M ? mo[1].x : 0.f, M ? mo[2].x : 0.f);
        w.dx.x = r.x; w.dy.x = r.y; w.ix.x = r.z; w.iy.x = r.w;
      }
      if (!w.regular[1]) {
        const float4 r = cdp_warp_lane_literal(1, (float)u, (float)v, depth, cam, p.pose1 + (size_t)c.b * 16, M,
                                               M ? mo[0].y : 0.f, M ? mo[1].y : 0.f, M ? mo[2].y : 0.f);
        w.dx.y = r.x; w.dy.y = r.y; w.ix.y = r.z; w.iy.y = r.w;
      }
    }
    int ax0 = 0, ay0 = 0, ax1 = 0, ay1 = 0, bx0, by0, bx1, by1;
    float2 fx, fy;
    bool in0, in1;
    if (interior && cdp_taps_interior<G>(tx + Geo::SBM, ty + Geo::SBM, w, bx0, by0, bx1, by1, fx, fy)) {
      in0 = in1 = true;
    } else {
      float unused;
      cdp_tap_axis_full(u, w.dx.x, w.ix.x, W, ax0, fx.x, unused);
      cdp_tap_axis_full(v, w.dy.x, w.iy.x, H, ay0, fy.x, unused);
      cdp_tap_axis_full(u, w.dx.y, w.ix.y, W, ax1, fx.y, unused);
      cdp_tap_axis_full(v, w.dy.y, w.iy.y, H, ay1, fy.y, unused);
      bx0 = ax0 - box_x; by0 = ay0 - box_y; bx1 = ax1 - box_x; by1 = ay1 - box_y;
      in0 = (unsigned)bx0 <= (unsigned)(Geo::SBW - 2) && (unsigned)by0 <= (unsigned)(Geo::SBH - 2);
      in1 = (unsigned)bx1 <= (unsigned)(Geo::SBW - 2) && (unsigned)by1 <= (unsigned)(Geo::SBH - 2);
    }
    const int ti = ty * Geo::TBW + tx;
    if (in0 && in1) {
      // both 2x2 footprints lie inside the staged source boxes: 24 shared-memory loads at
      // immediate offsets from two addresses
      const float* q0 = sbox0 + by0 * Geo::SBW + bx0;
      const float* q1 = sbox1 + by1 * Geo::SBW + bx1;
      // (all 24 loads before the first store: the compiler cannot tell that the warped planes and the
      // source boxes never overlap, so a store between the channels would serialise their loads)
      float2 nw[3], ne[3], sw[3], se[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        nw[ch].x = q0[ch * Geo::SBN]; ne[ch].x = q0[ch * Geo::SBN + 1];
        sw[ch].x = q0[ch * Geo::SBN + Geo::SBW]; se[ch].x = q0[ch * Geo::SBN + Geo::SBW + 1];
        nw[ch].y = q1[ch * Geo::SBN]; ne[ch].y = q1[ch * Geo::SBN + 1];
        sw[ch].y = q1[ch * Geo::SBN + Geo::SBW]; se[ch].y = q1[ch * Geo::SBN + Geo::SBW + 1];
      }
      float2 wv[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) wv[ch] = cdp_lerp2(nw[ch], ne[ch], sw[ch], se[ch], fx, fy);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) cdp_warp_plane<G>(sm, ch)[ti] = wv[ch];
    } else {
      // a footprint leaves the gather margin: global loads for that source (cold path, out of line)
      cdp_phase_a_taps_far<G>(sbox0, sbox1, src0, src1, plane, W, in0, in1, bx0, by0, bx1, by1, ax0, ay0, ax1, ay1, fx, fy,
                              sm, ti);
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

CDP_HD void cdp_rows_add(const CdpRowTgt& a, const CdpRowTgt& b, CdpRowTgt& o) { o.s = a.s + b.s; o.ss = a.ss + b.ss; }
CDP_HD void cdp_rows_add(const CdpRowPair& a, const CdpRowPair& b, CdpRowPair& o) {
  o.s = cdp_add2(a.s, b.s); o.ss = cdp_add2(a.ss, b.ss); o.sy = cdp_add2(a.sy, b.sy);
}

// SSIM loss of a candidate pair from the 3x3 sums of strip-centred values (algos/depth.py:141-153).
// ct = constant that turns strip-centred values back into true image values (means only).
//
// The second factor n2 / d2 = (2 cov + C2) / (var_x + var_y + C2) is a ratio, so it is formed from the
// raw sums scaled by 81 instead of from means and variances:
//     81 (2 cov + C2)          = 18 Sxy - 2 Sx Sy + 81 C2
//    -81 (var_x + var_y + C2)  = Sx^2 - 9 Sxx + (Sy^2 - 9 Syy - 81 C2)
// (two FMAs each, the same cancellation as E[x^2] - mean^2 with one rounding fewer; the minus sign
// of the denominator is undone in the final FMA).  9 packed operations per pair instead of 13.
#if CDP_OPT_SSIM_RAW
struct CdpSsimTgt {  // target-side terms shared by the two candidate pairs of a pixel
  float my, myy_c1, m2sy, nkd;
};
CDP_HD void cdp_ssim_tgt(float sy, float syy, float ct, CdpSsimTgt& o) {
  const float my = cdp_fmaf(sy, 1.0f / 9.0f, ct);
  o.my = my;
  o.myy_c1 = cdp_fmaf(my, my, CDP_SSIM_C1);
  o.m2sy = -2.0f * sy;
  o.nkd = cdp_fmaf(sy, sy, cdp_fmaf(syy, -9.0f, -81.0f * CDP_SSIM_C2));
}
CDP_HD float2 cdp_ssim_pair_loss(float2 sx, float2 sxx, float2 sxy, const CdpSsimTgt& t, float ct) {
  const float2 mx = cdp_fma2(sx, cdp_set2(1.0f / 9.0f), cdp_set2(ct));
  const float2 n1h = cdp_fma2(mx, cdp_set2(t.my), cdp_set2(0.5f * CDP_SSIM_C1));  // n1 / 2 (exactly)
  const float2 d1 = cdp_fma2(mx, mx, cdp_set2(t.myy_c1));
  const float2 n2 = cdp_fma2(sx, cdp_set2(t.m2sy), cdp_fma2(sxy, cdp_set2(18.0f), cdp_set2(81.0f * CDP_SSIM_C2)));
  const float2 nd2 = cdp_fma2(sx, sx, cdp_fma2(sxx, cdp_set2(-9.0f), cdp_set2(t.nkd)));
  const float2 numh = cdp_mul2(n1h, n2), nden = cdp_mul2(d1, nd2);  // numh / nden = -SSIM / 2
  float2 l;  // clamp((1 - S) / 2, 0, 1): reciprocal + one saturating FMA per lane (NaN -> 0 like fmin(fmax(NaN, 0), 1))
  l.x = cdp_half_plus_ratio_sat(numh.x, nden.x);
  l.y = cdp_half_plus_ratio_sat(numh.y, nden.y);
  return l;
}
#else
// (means / variances form: 13 packed operations per pair)
struct CdpSsimTgt {
  float myc, vy_c2, two_my, myy_c1;
};
CDP_HD void cdp_ssim_tgt(float sy, float syy, float ct, CdpSsimTgt& o) {
  const float ninth = 1.0f / 9.0f;
  o.myc = sy * ninth;
  o.vy_c2 = (syy * ninth - o.myc * o.myc) + CDP_SSIM_C2;
  const float my = o.myc + ct;
  o.two_my = 2.0f * my;
  o.myy_c1 = my * my + CDP_SSIM_C1;
}
CDP_HD float2 cdp_ssim_pair_loss(float2 sx, float2 sxx, float2 sxy, const CdpSsimTgt& t, float ct) {
  const float2 n9 = cdp_set2(1.0f / 9.0f);
  const float2 mxc = cdp_mul2(sx, n9);
  const float2 mx = cdp_add2(mxc, cdp_set2(ct));
  const float2 vx = cdp_fma2(cdp_mul2(mxc, mxc), cdp_set2(-1.0f), cdp_mul2(sxx, n9));
  const float2 cov = cdp_fma2(mxc, cdp_set2(-t.myc), cdp_mul2(sxy, n9));
  const float2 n1 = cdp_fma2(mx, cdp_set2(t.two_my), cdp_set2(CDP_SSIM_C1));
  const float2 n2 = cdp_fma2(cov, cdp_set2(2.0f), cdp_set2(CDP_SSIM_C2));
  const float2 d1 = cdp_fma2(mx, mx, cdp_set2(t.myy_c1));
  const float2 d2 = cdp_add2(vx, cdp_set2(t.vy_c2));
  const float2 num = cdp_mul2(n1, n2), den = cdp_mul2(d1, d2);
  float2 S;
  S.x = cdp_fdiv(num.x, den.x);
  S.y = cdp_fdiv(num.y, den.y);
  float2 l;  // clamp((1 - S) / 2, 0, 1): one saturating FMA per lane (NaN -> 0 like fmin(fmax(NaN, 0), 1))
  l.x = cdp_saturate(cdp_fmaf(S.x, -0.5f, 0.5f));
  l.y = cdp_saturate(cdp_fmaf(S.y, -0.5f, 0.5f));
  return l;
}
#endif

// ------------------------------------------------------------------------------------------
// Phase B1: candidate losses, noise, min / argmin (algos/depth.py:294-323)
// ------------------------------------------------------------------------------------------
template <bool G>
CDP_HD void cdp_photo_phase_b1(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm,
                               float& loss_acc) {
  typedef CdpTileGeom<G> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  uint8_t* kplane = reinterpret_cast<uint8_t*>(sm + Geo::O_K);
  // (computed per thread on purpose: reading them from the parameter bank instead measured 1.7 % slower)
  const float a3 = p.alpha * (1.0f / 3.0f), b3 = (float)(1.0 - (double)p.alpha) * (1.0f / 3.0f);
  const uint64_t seed = (!lv.noise && p.seed_dev) ? *p.seed_dev : p.seed;  // built-in generator: device counter or host value
  for (int item = tid; item < Geo::NITEMS; item += nthreads) {
    const int strip = item / Geo::BW, bx = item - strip * Geo::BW;
    const int by0 = strip * CDP_STRIP;
    const int qx = c.x0 - Geo::HB + bx;
    const int qy0 = c.y0 - Geo::HB + by0;
    // t-index / s-index of the window centre of output row 0 of this strip
    const int r00 = (by0 + 1 + Geo::OFFY) * Geo::TBW + bx + 1 + Geo::OFFX;
    const int s00 = (by0 + 1 + Geo::OFFY + Geo::SBM) * Geo::SBW + bx + 1 + Geo::OFFX + Geo::SBM;
    float2 acc_id[CDP_STRIP], acc_pe[CDP_STRIP];
#pragma unroll
    for (int o = 0; o < CDP_STRIP; ++o) { acc_id[o] = cdp_set2(0.f); acc_pe[o] = cdp_set2(0.f); }
    const bool col_ok = qx >= 0 && qx < W;
    // tie-break noise of the strip's pixels, requested early so the loads overlap the strip walk
    float2 nz[CDP_STRIP];
#if !CDP_OPT_NOISE_LATE
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
#endif
    const int qc = qy0 + CDP_STRIP / 2 < 0 ? 0 : (qy0 + CDP_STRIP / 2 > H - 1 ? H - 1 : qy0 + CDP_STRIP / 2);
    const int cs_idx = (qc - c.y0 + Geo::TYO) * Geo::TBW + bx + 1 + Geo::OFFX;
    if (col_ok && qy0 < H && qy0 + CDP_STRIP > 0) {
      // the channel loop stays rolled: unrolled, this phase alone is ~60 KB of SASS that every
      // warp streams through once per tile, and instruction fetch becomes the top stall reason
#pragma unroll 1
      for (int ch = 0; ch < 3; ++ch) {
        const float* ty = sm + Geo::O_TGT + ch * Geo::TBN;
        const float* ts0 = sm + Geo::O_SRC + ch * Geo::SBN;
        const float* ts1 = ts0 + Geo::SRC_STRIDE;
        const float2* tw = cdp_warp_plane<G>(sm, ch);
        // strip constant: target value at the strip's middle pixel, clamped into the image so
        // that it is always a staged value; every window value is centred on it
        const float cs = ty[cs_idx];
        const float2 cs2 = cdp_set2(-cs);
        CdpRowTgt hy[3], py2;
        CdpRowPair hs[3], hw[3], ps2, pw2;
        py2.s = py2.ss = 0.f;
        ps2.s = ps2.ss = ps2.sy = pw2.s = pw2.ss = pw2.sy = cdp_set2(0.f);
        float yc_prev = 0.f;
        float2 sc_prev = cdp_set2(0.f), wc_prev = cdp_set2(0.f);
#pragma unroll
        for (int r = 0; r < CDP_STRIP + 2; ++r) {
          const int row = r00 + (r - 1) * Geo::TBW;  // window row r-1 relative to output row 0
          const int srow = s00 + (r - 1) * Geo::SBW;
          float y[3];
          float2 s[3], w[3];
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            y[t] = ty[row + t - 1] - cs;
            float2 sv;
            sv.x = ts0[srow + t - 1]; sv.y = ts1[srow + t - 1];
            s[t] = cdp_add2(sv, cs2);
            w[t] = cdp_add2(tw[row + t - 1], cs2);
          }
          cdp_row_tgt(y, hy[r % 3]);
          cdp_row_pair(s, y, hs[r % 3]);
          cdp_row_pair(w, y, hw[r % 3]);
          if (r >= 2) {
            const int o = r - 2;  // output row: window rows r-2, r-1, r; centre row r-1
            // 3-row sums; with CDP_OPT_PAIR_SHARE the sum of the two lower rows of an even output is
            // kept and reused by the odd output below it (3 additions per two outputs instead of 4)
            CdpRowTgt y3;
            CdpRowPair s3, w3;
            if (CDP_OPT_PAIR_SHARE && (o & 1) == 0 && o + 1 < CDP_STRIP) {
              cdp_rows_add(hy[(r + 2) % 3], hy[r % 3], py2);
              cdp_rows_add(hs[(r + 2) % 3], hs[r % 3], ps2);
              cdp_rows_add(hw[(r + 2) % 3], hw[r % 3], pw2);
              cdp_rows_add(hy[(r + 1) % 3], py2, y3);
              cdp_rows_add(hs[(r + 1) % 3], ps2, s3);
              cdp_rows_add(hw[(r + 1) % 3], pw2, w3);
            } else if (CDP_OPT_PAIR_SHARE && (o & 1) == 1) {
              cdp_rows_add(py2, hy[r % 3], y3);
              cdp_rows_add(ps2, hs[r % 3], s3);
              cdp_rows_add(pw2, hw[r % 3], w3);
            } else {
              CdpRowTgt ty; CdpRowPair ts, tw2;
              cdp_rows_add(hy[0], hy[1], ty); cdp_rows_add(ty, hy[2], y3);
              cdp_rows_add(hs[0], hs[1], ts); cdp_rows_add(ts, hs[2], s3);
              cdp_rows_add(hw[0], hw[1], tw2); cdp_rows_add(tw2, hw[2], w3);
            }
            CdpSsimTgt st;
            cdp_ssim_tgt(y3.s, y3.ss, cs, st);
            const float2 l_id = cdp_ssim_pair_loss(s3.s, s3.ss, s3.sy, st, cs);
            const float2 l_pe = cdp_ssim_pair_loss(w3.s, w3.ss, w3.sy, st, cs);
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
#if CDP_OPT_NOISE_LATE
    // (all loads of the strip issued together, after the walk: ten registers fewer live through it)
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
#endif
    // min-reprojection with identity auto-mask for the strip's pixels
#pragma unroll
    for (int o = 0; o < CDP_STRIP; ++o) {
      const int by = by0 + o;
      if (by >= Geo::BH) break;
      const int qy = qy0 + o;
      const int ridx = r00 + o * Geo::TBW;
      if (!col_ok || qy < 0 || qy >= H) {
        if (G) kplane[ridx] = 255;
        continue;
      }
      float n0 = nz[o].x, n1 = nz[o].y;
      if (!lv.noise) cdp_noise_pair(seed, (uint32_t)(qy * W + qx), (uint32_t)c.lvl, (uint32_t)c.b, n0, n1);
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

// SSIM adjoint coefficients (cdp_ssim_terms + cdp_ssim_coeffs_abc of cdp_math.h) from the raw 3x3
// sums of values centred on c, for values measured in the frame centred on `centre`:
//     d loss(q) / d x(p) = m/9 * (A + 2 (x(p) - centre) B + (y(p) - centre) C).
// n2 / d2 is formed from the sums scaled by 81 as in cdp_ssim_pair_loss; 1 / d2 = 81 / D2.
CDP_HD void cdp_ssim_abc_raw(float sx, float sxx, float sxy, float sy, float syy, float c, float centre,
                             float& A, float& B, float& C) {
  const float ninth = 1.0f / 9.0f;
  const float mx = cdp_fmaf(sx, ninth, c), my = cdp_fmaf(sy, ninth, c);  // true means
  const float n1 = cdp_fmaf(mx + mx, my, CDP_SSIM_C1);
  const float d1 = cdp_fmaf(mx, mx, cdp_fmaf(my, my, CDP_SSIM_C1));
  const float N2 = cdp_fmaf(sx, -2.0f * sy, cdp_fmaf(sxy, 18.0f, 81.0f * CDP_SSIM_C2));
  const float D2 = cdp_fmaf(-sx, sx, cdp_fmaf(sxx, 9.0f, cdp_fmaf(-sy, sy, cdp_fmaf(syy, 9.0f, 81.0f * CDP_SSIM_C2))));
  const float id1 = cdp_rcp(d1), iD2 = cdp_rcp(D2);
  const float r1 = n1 * id1, r2 = N2 * iD2;  // the two SSIM factors
  const float S = r1 * r2;
  const float raw = cdp_fmaf(S, -0.5f, 0.5f);
  const float g = (raw >= 0.f && raw <= 1.f) ? 40.5f * iD2 : 0.f;  // -dl / d2 (clamp gradient is inclusive), dl = -0.5
  B = g * S;                  // dl * (-S / d2)
  C = -2.0f * (g * r1);       // dl * 2 n1 / (d1 d2)
  const float h = (raw >= 0.f && raw <= 1.f) ? -id1 : 0.f;  // 2 dl / d1
  const float A1 = h * cdp_fmaf(my, r2, -(mx * S));
  A = cdp_fmaf(-2.0f * (mx - centre), B, cdp_fmaf(-(my - centre), C, A1));
}

// ------------------------------------------------------------------------------------------
// Phase B2 (with grad): SSIM adjoint coefficients of the winning reprojection, in the frame of
// the tile-centred values:  d loss(q) / d x(p) = m/9 * (A + 2 x(p) B + y(p) C).
// ------------------------------------------------------------------------------------------
CDP_HD void cdp_photo_phase_b2(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm) {
  typedef CdpTileGeom<true> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const uint8_t* kplane = reinterpret_cast<const uint8_t*>(sm + Geo::O_K);
  float centre[3];
  cdp_tile_centre<true>(lv, c, sm, centre);
  for (int item = tid; item < Geo::NITEMS; item += nthreads) {
    const int strip = item / Geo::BW, bx = item - strip * Geo::BW;
    const int by0 = strip * CDP_STRIP;
    const int r00 = (by0 + 1 + Geo::OFFY) * Geo::TBW + bx + 1 + Geo::OFFX;
    int kk[CDP_STRIP];
    bool any = false;
#pragma unroll
    for (int o = 0; o < CDP_STRIP; ++o) {
      kk[o] = (by0 + o < Geo::BH) ? kplane[r00 + o * Geo::TBW] : 255;
      any = any || kk[o] < 2;
    }
    if (!any) continue;  // every pixel of the strip is auto-masked or outside the image
    const int qy0 = c.y0 - Geo::HB + by0;
    const int qc = qy0 + CDP_STRIP / 2 < 0 ? 0 : (qy0 + CDP_STRIP / 2 > lv.H - 1 ? lv.H - 1 : qy0 + CDP_STRIP / 2);
    const int cs_idx = (qc - c.y0 + Geo::TYO) * Geo::TBW + bx + 1 + Geo::OFFX;
#pragma unroll 1
    for (int ch = 0; ch < 3; ++ch) {
      const float* ty = sm + Geo::O_TGT + ch * Geo::TBN;
      const float2* tw = cdp_warp_plane<true>(sm, ch);
      const float cs = ty[cs_idx];
      const float2 cs2 = cdp_set2(-cs);
      CdpRowTgt hy[3];
      CdpRowPair hw[3];
      // The coefficients of an output are stored one row later, after that row's loads: the compiler
      // cannot tell that the coefficient planes and the planes the walk reads never overlap, so a
      // store right after the (long) coefficient chain would hold back the next row's loads.
      float pA = 0.f, pB = 0.f, pC = 0.f;
      int pidx = -1;
#pragma unroll
      for (int r = 0; r < CDP_STRIP + 2; ++r) {
        // (the last, partial strip would read one row past the planes -- values that no output uses,
        // but the warped plane of channel 2 is followed by the coefficient planes other threads are
        // writing: the row is clamped into the box so that no such read exists)
        int row = r00 + (r - 1) * Geo::TBW;
        if (Geo::NSTRIP * CDP_STRIP - CDP_STRIP + r + Geo::OFFY >= Geo::TBH) {  // (compile time: only the last row(s) of the walk)
          const int over = by0 + Geo::OFFY + r - (Geo::TBH - 1);  // box rows past the last one
          if (over > 0) row -= over * Geo::TBW;
        }
        float y[3];
        float2 w[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          y[t] = ty[row + t - 1] - cs;
          w[t] = cdp_add2(tw[row + t - 1], cs2);
        }
        if (pidx >= 0) {
          sm[Geo::O_COEF + (ch * 3 + 0) * Geo::TBN + pidx] = pA;
          sm[Geo::O_COEF + (ch * 3 + 1) * Geo::TBN + pidx] = pB;
          sm[Geo::O_COEF + (ch * 3 + 2) * Geo::TBN + pidx] = pC;
          pidx = -1;
        }
        cdp_row_tgt(y, hy[r % 3]);
        cdp_row_pair(w, y, hw[r % 3]);
        if (r >= 2) {
          const int o = r - 2;
          if (kk[o] < 2) {
            const bool first = kk[o] == 0;
            const float2 sx2 = cdp_add2(cdp_add2(hw[0].s, hw[1].s), hw[2].s);
            const float2 sxx2 = cdp_add2(cdp_add2(hw[0].ss, hw[1].ss), hw[2].ss);
            const float2 sxy2 = cdp_add2(cdp_add2(hw[0].sy, hw[1].sy), hw[2].sy);
            float A, B, C;
#if CDP_OPT_SSIM_RAW_B2
            cdp_ssim_abc_raw(first ? sx2.x : sx2.y, first ? sxx2.x : sxx2.y, first ? sxy2.x : sxy2.y,
                             hy[0].s + hy[1].s + hy[2].s, hy[0].ss + hy[1].ss + hy[2].ss, cs, centre[ch], A, B, C);
#else
            const float ninth = 1.0f / 9.0f;
            const float mxc = (first ? sx2.x : sx2.y) * ninth;
            const float exx = (first ? sxx2.x : sxx2.y) * ninth;
            const float exy = (first ? sxy2.x : sxy2.y) * ninth;
            const float myc = (hy[0].s + hy[1].s + hy[2].s) * ninth;
            const float eyy = (hy[0].ss + hy[1].ss + hy[2].ss) * ninth;
            CdpSsimTerms t;
            cdp_ssim_terms(mxc, myc, exx, eyy, exy, cs, t);
            // means in the tile-centred frame are mxc + csc, myc + csc
            const float csc = cs - centre[ch];  // strip constant in the tile-centred frame
            cdp_ssim_coeffs_abc(t, mxc + csc, myc + csc, A, B, C);
#endif
            pA = A; pB = B; pC = C;
            pidx = r00 + o * Geo::TBW;
          }
        }
      }
      if (pidx >= 0) {
        sm[Geo::O_COEF + (ch * 3 + 0) * Geo::TBN + pidx] = pA;
        sm[Geo::O_COEF + (ch * 3 + 1) * Geo::TBN + pidx] = pB;
        sm[Geo::O_COEF + (ch * 3 + 2) * Geo::TBN + pidx] = pC;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Phase C (with grad): both sources in the two lanes of packed fp32.
// ------------------------------------------------------------------------------------------
// One source of one pixel through the literal (depth clamp active / Q_w <= 0) formulas: rare, out of
// line, by value.  gw0..2 = dL/d warped value of this source per channel; returns dL/d depth and the
// pixel's contribution to dL/dT of source k.
struct CdpLiteralOut {
  float gd;
  float dT[16];
};
template <bool M>
CDP_COLD CdpLiteralOut cdp_phase_c_lane_literal(int k, const CdpPhotoParams& p, int lvl, int b, int px, int py,
                                               float depth, CdpCam cam, float gw0, float gw1, float gw2) {
  const CdpLevel& lv = p.lv[lvl];
  const int W = lv.W, H = lv.H;
  const size_t plane = (size_t)W * H;
  CdpPose T;
  cdp_load_pose_aligned((k == 0 ? p.pose0 : p.pose1) + (size_t)b * 16, T);
  const float* srck = (k == 0 ? lv.src0 : lv.src1) + (size_t)b * 3 * plane;
  const float gw[3] = {gw0, gw1, gw2};
  float mo[3], gmo[3];
  if (M) {
    const float* motk = k == 0 ? lv.mot0 : lv.mot1;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) mo[ch] = CDP_LDG(motk + ((size_t)b * 3 + ch) * plane + py * W + px);
  }
  CdpWarp w;
  cdp_warp_point((float)px, (float)py, depth, cam, T, M ? mo : nullptr, w);
  CdpTaps t;
  cdp_taps(px, py, w, W, H, t);
  float gix = 0.f, giy = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < 3; ++ch) {
    float dix, diy;
    cdp_bilinear_grad(srck + ch * plane, t, dix, diy);
    gix += gw[ch] * dix;
    giy += gw[ch] * diy;
  }
  CdpLiteralOut o;
  o.gd = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) o.dT[i] = 0.f;
  cdp_warp_adjoint(gix * t.mx, giy * t.my, w, cam, T, o.gd, o.dT, M ? gmo : nullptr);
  if (M) {
    float* dst = (k == 0 ? lv.gmot0 : lv.gmot1) + (size_t)b * 3 * plane + py * W + px;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) dst[ch * plane] = gmo[ch];
  }
  return o;
}

// Phase C1: per tile pixel, the masked, reflection-weighted 3x3 sums of the coefficient planes
// for both sources and the resulting dL/d warped value per channel (level weight included); the
// pair overwrites the pixel's own entry of the warped planes (only this thread reads it again).
// A pixel that no window selecting a reprojection reaches gets exact zeros.
//
// Each thread owns CDP_C1_ROWS vertically adjacent pixels of one column and walks down their
// window rows: per row the horizontal 3-tap sums of all nine planes (neighbour's winner and the
// horizontal reflection multiplicity folded into a per-source weight pair), kept in a ring of
// three rows; an output is the vertically weighted sum of the ring.  ~2.5x fewer shared-memory
// loads and FMAs than nine separate 3x3 gathers per pixel.
#ifndef CDP_C1_ROWS
#define CDP_C1_ROWS 4
#endif
CDP_HD void cdp_photo_phase_c1(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm) {
  typedef CdpTileGeom<true> Geo;
  static_assert(CDP_TILE_Y % CDP_C1_ROWS == 0, "C1 strips must tile the rows");
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  const uint8_t* kplane = reinterpret_cast<const uint8_t*>(sm + Geo::O_K);
  float centre[3];
  cdp_tile_centre<true>(lv, c, sm, centre);
  const float w_ssim = p.alpha / 27.0f * lv.weight;  // alpha * (1/3 channels) * (1/9 window) * level weight
  const float w_l1 = (float)(1.0 - (double)p.alpha) * (1.0f / 3.0f) * lv.weight;
  for (int item = tid; item < CDP_TILE_X * (CDP_TILE_Y / CDP_C1_ROWS); item += nthreads) {
    const int strip = item / CDP_TILE_X, lx = item - strip * CDP_TILE_X;
    const int ly0 = strip * CDP_C1_ROWS;
    const int px = c.x0 + lx;
    if (px >= W || c.y0 + ly0 >= H) continue;
    // horizontal reflection multiplicities of the three window columns (0 outside the image)
    float mxw[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) mxw[d] = cdp_reflect_mult(px, d - 1, W);
    float2 h[3][9];
    float2 pg[3];   // results of the previous output, stored after the next output's loads
    int pidx = -1;
#pragma unroll
    for (int r = 0; r < CDP_C1_ROWS + 2; ++r) {
      const int ty = ly0 - 1 + r + Geo::TYO;  // box row of window row r
      const int n0 = ty * Geo::TBW + lx + Geo::TXO;
      float2 mk[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int kj = (int)kplane[n0 + d - 1];  // (255 outside the image: matches neither source)
        mk[d].x = kj == 0 ? mxw[d] : 0.f;
        mk[d].y = kj == 1 ? mxw[d] : 0.f;
      }
#pragma unroll
      for (int pl = 0; pl < 9; ++pl) {
        const float* cf = sm + Geo::O_COEF + pl * Geo::TBN + n0;
        h[r % 3][pl] = cdp_fma2(mk[2], cdp_set2(cf[1]), cdp_fma2(mk[1], cdp_set2(cf[0]), cdp_mul2(mk[0], cdp_set2(cf[-1]))));
      }
      if (r >= 2) {
        const int ly = ly0 + r - 2, py = c.y0 + ly;
        if (py < H) {
          const int ridx = (ly + Geo::TYO) * Geo::TBW + lx + Geo::TXO;
          const int kown = kplane[ridx];
          // this output's own values, loaded before the previous output's results are stored (the
          // compiler cannot tell that stores and loads never overlap and keeps their order)
          float2 xr[3];
          float yr[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            xr[ch] = cdp_warp_plane<true>(sm, ch)[ridx];
            yr[ch] = sm[Geo::O_TGT + ch * Geo::TBN + ridx];
          }
          if (pidx >= 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) cdp_warp_plane<true>(sm, ch)[pidx] = pg[ch];
          }
          // vertical weights: rows (r-2, r-1, r) of the ring are window rows dy = -1, 0, +1
          const float2 m0 = cdp_set2(cdp_reflect_mult(py, -1, H)), m2 = cdp_set2(cdp_reflect_mult(py, 1, H));
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            float2 sabc[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const int pl = ch * 3 + q;
              sabc[q] = cdp_fma2(m2, h[r % 3][pl], cdp_fma2(m0, h[(r - 2) % 3][pl], h[(r - 1) % 3][pl]));
            }
            const float2 x = cdp_add2(xr[ch], cdp_set2(-centre[ch]));
            const float y = yr[ch] - centre[ch];
            float2 g = cdp_mul2(cdp_set2(w_ssim), cdp_fma2(cdp_mul2(x, cdp_set2(2.f)), sabc[1], cdp_fma2(cdp_set2(y), sabc[2], sabc[0])));
            if (kown == 0) g.x += w_l1 * (x.x > y ? 1.f : (x.x < y ? -1.f : 0.f));
            if (kown == 1) g.y += w_l1 * (x.y > y ? 1.f : (x.y < y ? -1.f : 0.f));
            pg[ch] = g;
          }
          pidx = ridx;
        }
      }
    }
    if (pidx >= 0) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) cdp_warp_plane<true>(sm, ch)[pidx] = pg[ch];
    }
  }
}

// Phase C2, cold path: at least one footprint leaves the staged source box (taps of that source from
// global memory); accumulates dL/d(ix, iy) of both sources.
CDP_COLD float4 cdp_phase_c2_taps_far(const float* sbox0, const float* sbox1, const float* src0, const float* src1,
                                      size_t plane, int W, bool in0, bool in1, int bx0, int by0, int bx1, int by1, int ax0,
                                      int ay0, int ax1, int ay1, float2 fx, float2 fy, float2 gw0, float2 gw1, float2 gw2) {
  typedef CdpTileGeom<true> Geo;
  const float2 neg1 = cdp_set2(-1.0f);
  const float2 wy0 = cdp_fma2(fy, neg1, cdp_set2(1.0f)), wx0 = cdp_fma2(fx, neg1, cdp_set2(1.0f));
  float2 gix = cdp_set2(0.f), giy = cdp_set2(0.f);
  float2 nw[3], ne[3], sw[3], se[3];  // (all taps requested before the first use: one memory round trip per pixel)
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (in0) {
      const float* q = sbox0 + ch * Geo::SBN + by0 * Geo::SBW + bx0;
      nw[ch].x = q[0]; ne[ch].x = q[1]; sw[ch].x = q[Geo::SBW]; se[ch].x = q[Geo::SBW + 1];
    } else {
      const float* g = src0 + ch * plane + (size_t)ay0 * W + ax0;
      nw[ch].x = CDP_LDG(g); ne[ch].x = CDP_LDG(g + 1); sw[ch].x = CDP_LDG(g + W); se[ch].x = CDP_LDG(g + W + 1);
    }
    if (in1) {
      const float* q = sbox1 + ch * Geo::SBN + by1 * Geo::SBW + bx1;
      nw[ch].y = q[0]; ne[ch].y = q[1]; sw[ch].y = q[Geo::SBW]; se[ch].y = q[Geo::SBW + 1];
    } else {
      const float* g = src1 + ch * plane + (size_t)ay1 * W + ax1;
      nw[ch].y = CDP_LDG(g); ne[ch].y = CDP_LDG(g + 1); sw[ch].y = CDP_LDG(g + W); se[ch].y = CDP_LDG(g + W + 1);
    }
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float2 dix = cdp_fma2(cdp_fma2(nw[ch], neg1, ne[ch]), wy0, cdp_mul2(cdp_fma2(sw[ch], neg1, se[ch]), fy));
    const float2 diy = cdp_fma2(cdp_fma2(nw[ch], neg1, sw[ch]), wx0, cdp_mul2(cdp_fma2(ne[ch], neg1, se[ch]), fx));
    const float2 g = ch == 0 ? gw0 : (ch == 1 ? gw1 : gw2);
    gix = cdp_fma2(g, dix, gix);
    giy = cdp_fma2(g, diy, giy);
  }
  float4 r;
  r.x = gix.x; r.y = gix.y; r.z = giy.x; r.w = giy.y;
  return r;
}

// Phase C2 (the source boxes are staged again: the coefficient planes that overlaid them are dead):
// chain dL/d warped through the bilinear sampler's coordinate derivative and the projection to
// dL/d depth_s (written) and dL/dT (accumulated), both sources in the two lanes of packed fp32.
template <bool M>
CDP_HD void cdp_photo_phase_c2(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads,
                               const float* sm, float* dT /*[32]: source-major 4x4 blocks*/, const CdpTileConst& kc) {
  typedef CdpTileGeom<true> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const int W = lv.W, H = lv.H;
  const size_t plane = (size_t)W * H;
  const CdpCam& cam = kc.cam;
  const CdpPose2& T = kc.T;
  const float* src0 = lv.src0 + (size_t)c.b * 3 * plane;
  const float* src1 = lv.src1 + (size_t)c.b * 3 * plane;
  const float* sbox0 = sm + Geo::O_SRC;
  const float* sbox1 = sm + Geo::O_SRC + Geo::SRC_STRIDE;
  const int box_x = c.x0 - Geo::TXO - Geo::SBM, box_y = c.y0 - Geo::TYO - Geo::SBM;
  const bool interior = cdp_tile_interior<true>(lv, c);
  float2 dT2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dT2[i] = cdp_set2(0.f);
  constexpr int kUnrollC2 = CDP_C2_UNROLL;
#pragma unroll kUnrollC2
  for (int idx = tid; idx < CDP_TILE_X * CDP_TILE_Y; idx += nthreads) {
    const int ly = idx / CDP_TILE_X, lx = idx - ly * CDP_TILE_X;
    const int px = c.x0 + lx, py = c.y0 + ly;
    if (px >= W || py >= H) continue;
    const int ridx = (ly + Geo::TYO) * Geo::TBW + lx + Geo::TXO;
    float2 gw[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) gw[ch] = cdp_warp_plane<true>(sm, ch)[ridx];
    const bool any = gw[0].x != 0.f || gw[0].y != 0.f || gw[1].x != 0.f || gw[1].y != 0.f || gw[2].x != 0.f || gw[2].y != 0.f;
    if (!any) {  // no gradient reaches this pixel (auto-masked neighbourhood)
      if (M) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          lv.gmot0[((size_t)c.b * 3 + ch) * plane + py * W + px] = 0.f;
          lv.gmot1[((size_t)c.b * 3 + ch) * plane + py * W + px] = 0.f;
        }
      }
      lv.gdepth[(size_t)c.b * plane + py * W + px] = 0.f;
      continue;
    }
    const float depth = sm[Geo::O_DEPTH + ridx];
    float2 mo[3];
    if (M) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        mo[ch].x = CDP_LDG(lv.mot0 + ((size_t)c.b * 3 + ch) * plane + py * W + px);
        mo[ch].y = CDP_LDG(lv.mot1 + ((size_t)c.b * 3 + ch) * plane + py * W + px);
      }
    }
    CdpWarp2 w;
    cdp_warp_point2((float)px, (float)py, depth, cam, T, M ? mo : nullptr, w);
    int ax0 = 0, ay0 = 0, ax1 = 0, ay1 = 0, bx0, by0, bx1, by1;
    float2 fx, fy, mx, my;
    bool in0, in1;
    if (interior && cdp_taps_interior<true>(lx + Geo::TXO + Geo::SBM, ly + Geo::TYO + Geo::SBM, w, bx0, by0, bx1, by1, fx, fy)) {
      in0 = in1 = true;
      mx = my = cdp_set2(1.0f);
    } else {
      cdp_tap_axis_full(px, w.dx.x, w.ix.x, W, ax0, fx.x, mx.x);
      cdp_tap_axis_full(py, w.dy.x, w.iy.x, H, ay0, fy.x, my.x);
      cdp_tap_axis_full(px, w.dx.y, w.ix.y, W, ax1, fx.y, mx.y);
      cdp_tap_axis_full(py, w.dy.y, w.iy.y, H, ay1, fy.y, my.y);
      bx0 = ax0 - box_x; by0 = ay0 - box_y; bx1 = ax1 - box_x; by1 = ay1 - box_y;
      in0 = (unsigned)bx0 <= (unsigned)(Geo::SBW - 2) && (unsigned)by0 <= (unsigned)(Geo::SBH - 2);
      in1 = (unsigned)bx1 <= (unsigned)(Geo::SBW - 2) && (unsigned)by1 <= (unsigned)(Geo::SBH - 2);
    }
    float2 gix = cdp_set2(0.f), giy = cdp_set2(0.f);
    const float2 neg1 = cdp_set2(-1.0f);
    const float2 wy0 = cdp_fma2(fy, neg1, cdp_set2(1.0f)), wx0 = cdp_fma2(fx, neg1, cdp_set2(1.0f));
    if (in0 && in1) {
      const float* q0 = sbox0 + by0 * Geo::SBW + bx0;
      const float* q1 = sbox1 + by1 * Geo::SBW + bx1;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float2 nw, ne, sw, se;
        nw.x = q0[ch * Geo::SBN]; ne.x = q0[ch * Geo::SBN + 1];
        sw.x = q0[ch * Geo::SBN + Geo::SBW]; se.x = q0[ch * Geo::SBN + Geo::SBW + 1];
        nw.y = q1[ch * Geo::SBN]; ne.y = q1[ch * Geo::SBN + 1];
        sw.y = q1[ch * Geo::SBN + Geo::SBW]; se.y = q1[ch * Geo::SBN + Geo::SBW + 1];
        const float2 dix = cdp_fma2(cdp_fma2(nw, neg1, ne), wy0, cdp_mul2(cdp_fma2(sw, neg1, se), fy));
        const float2 diy = cdp_fma2(cdp_fma2(nw, neg1, sw), wx0, cdp_mul2(cdp_fma2(ne, neg1, se), fx));
        gix = cdp_fma2(gw[ch], dix, gix);
        giy = cdp_fma2(gw[ch], diy, giy);
      }
    } else {
      const float4 r = cdp_phase_c2_taps_far(sbox0, sbox1, src0, src1, plane, W, in0, in1, bx0, by0, bx1, by1, ax0, ay0, ax1,
                                             ay1, fx, fy, gw[0], gw[1], gw[2]);
      gix.x = r.x; gix.y = r.y; giy.x = r.z; giy.y = r.w;
    }
    float2 gQ[3];
    float2 gdep = cdp_warp_adjoint2(cdp_mul2(gix, mx), cdp_mul2(giy, my), w, cam, T, dT2, gQ);
    if (!(w.regular[0] && w.regular[1])) {
      // a lane with the depth clamp active contributed exactly zero above (iz = 0): redo that
      // source with the literal formulas
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (w.regular[k]) continue;
        const CdpLiteralOut lit = cdp_phase_c_lane_literal<M>(k, p, c.lvl, c.b, px, py, depth, cam, k == 0 ? gw[0].x : gw[0].y,
                                                              k == 0 ? gw[1].x : gw[1].y, k == 0 ? gw[2].x : gw[2].y);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (k == 0) dT2[i].x += lit.dT[i]; else dT2[i].y += lit.dT[i];
        }
        if (k == 0) gdep.x = lit.gd; else gdep.y = lit.gd;
      }
    }
    if (M) {
      float* d0 = lv.gmot0 + (size_t)c.b * 3 * plane + py * W + px;
      float* d1 = lv.gmot1 + (size_t)c.b * 3 * plane + py * W + px;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (w.regular[0]) d0[ch * plane] = gQ[ch].x;
        if (w.regular[1]) d1[ch * plane] = gQ[ch].y;
      }
    }
    lv.gdepth[(size_t)c.b * plane + py * W + px] = gdep.x + gdep.y;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) { dT[i] += dT2[i].x; dT[16 + i] += dT2[i].y; }
}

// Re-staging of the two source boxes for phase C2 without TMA (plain loads).
CDP_HD void cdp_photo_restage_sources(const CdpPhotoParams& p, const CdpTileCtx& c, int tid, int nthreads, float* sm) {
  typedef CdpTileGeom<true> Geo;
  const CdpLevel& lv = p.lv[c.lvl];
  const size_t plane = (size_t)lv.W * lv.H;
  const int ox = c.x0 - Geo::TXO - Geo::SBM, oy = c.y0 - Geo::TYO - Geo::SBM;
  cdp_stage_box(lv.src0 + (size_t)c.b * 3 * plane, plane, 3, sm + Geo::O_SRC, Geo::SBW, Geo::SBH, ox, oy, lv.W, lv.H,
                tid, nthreads);
  cdp_stage_box(lv.src1 + (size_t)c.b * 3 * plane, plane, 3, sm + Geo::O_SRC + Geo::SRC_STRIDE, Geo::SBW, Geo::SBH, ox,
                oy, lv.W, lv.H, tid, nthreads);
}
