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
  // target, source0, source1 [B,3,H,W]; depth [B,1,H,W]; optionally motion0, motion1 [B,3,H,W]
  const float* in[6];
  float* out[6][CDP_MAX_LEVELS];       // per level (index 0 unused)
  int32_t nt;                          // 4, or 6 with object-motion maps
  const CdpResizeTap* tab_x[CDP_MAX_LEVELS];
  const CdpResizeTap* tab_y[CDP_MAX_LEVELS];
  int32_t Ws[CDP_MAX_LEVELS], Hs[CDP_MAX_LEVELS];
  int32_t begin[CDP_MAX_LEVELS + 1];   // prefix offsets of level work items within one image
  int32_t W, H, L;
  int32_t fast1;  // level 1 is an exact 2x2 mean (W % 4 == 0, H even): one item = two outputs
  // fused heads (cdp_photo_heads): in[3] is the DISPARITY; depth = 1 / (min_disp + span * disp) is
  // what gets resized, and the level-1 items also write the full-resolution depth map (depth_out)
  float* depth_out;  // null: in[3] is a depth map
  float min_disp, disp_span;
  // poses from the 6-DoF parameters, written by the first block (null: pose matrices are inputs)
  const float* axisangle[2];
  const float* translation[2];
  float* pose_out[2];
  int32_t invert[2];
  int32_t B;
};

CDP_HD float cdp_disp_to_depth(float disp, float min_disp, float span) { return 1.0f / (min_disp + span * disp); }
CDP_HD float cdp_disp_to_depth_grad(float g_depth, float depth, float span) { return -g_depth * span * depth * depth; }


CDP_HD void cdp_pyramid_fwd_item(const CdpPyrParams& p, int b, int item) {
  if (item >= p.begin[p.L]) return;
  int s = 1;
  while (s + 1 < p.L && item >= p.begin[s + 1]) ++s;
  const int local = item - p.begin[s];
  const int ws = p.Ws[s], hs = p.Hs[s];
  if (s == 1 && p.fast1) {
    // exact ratio 2: outputs (x, x+1) of row y are the 2x2 means of input columns 2x..2x+3, rows
    // 2y, 2y+1 -- two 16-byte loads per channel instead of eight 4-byte ones (same arithmetic
    // as the table path: taps weighted 1/2 horizontally, then 1/2 vertically)
    const int half = ws >> 1;
    const int y = local / half, x = (local - y * half) * 2;
    const size_t in_plane = (size_t)p.W * p.H, out_plane = (size_t)ws * hs;
    const int o0 = (2 * y) * p.W + 2 * x, o1 = o0 + p.W;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      if (t >= p.nt) break;
      const int ch = t == 3 ? 1 : 3;
      // all loads of a tensor are issued before the first use (six 16-byte loads in flight per thread)
      float4 a[3], d[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c >= ch) break;
        const float* src = p.in[t] + ((size_t)b * ch + c) * in_plane;
        a[c] = CDP_LDG(reinterpret_cast<const float4*>(src + o0));
        d[c] = CDP_LDG(reinterpret_cast<const float4*>(src + o1));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c >= ch) break;
        if (t == 3 && p.depth_out) {  // disparity in, depth out (full resolution) and resized
          a[c].x = cdp_disp_to_depth(a[c].x, p.min_disp, p.disp_span); a[c].y = cdp_disp_to_depth(a[c].y, p.min_disp, p.disp_span);
          a[c].z = cdp_disp_to_depth(a[c].z, p.min_disp, p.disp_span); a[c].w = cdp_disp_to_depth(a[c].w, p.min_disp, p.disp_span);
          d[c].x = cdp_disp_to_depth(d[c].x, p.min_disp, p.disp_span); d[c].y = cdp_disp_to_depth(d[c].y, p.min_disp, p.disp_span);
          d[c].z = cdp_disp_to_depth(d[c].z, p.min_disp, p.disp_span); d[c].w = cdp_disp_to_depth(d[c].w, p.min_disp, p.disp_span);
          *reinterpret_cast<float4*>(p.depth_out + (size_t)b * in_plane + o0) = a[c];
          *reinterpret_cast<float4*>(p.depth_out + (size_t)b * in_plane + o1) = d[c];
        }
        float2 r;
        r.x = (a[c].x * 0.5f + a[c].y * 0.5f) * 0.5f + (d[c].x * 0.5f + d[c].y * 0.5f) * 0.5f;
        r.y = (a[c].z * 0.5f + a[c].w * 0.5f) * 0.5f + (d[c].z * 0.5f + d[c].w * 0.5f) * 0.5f;
        *reinterpret_cast<float2*>(p.out[t][s] + ((size_t)b * ch + c) * out_plane + y * ws + x) = r;
      }
    }
    return;
  }
  const int y = local / ws, x = local - y * ws;
  const CdpResizeTap tx = p.tab_x[s][x], ty = p.tab_y[s][y];
  const size_t in_plane = (size_t)p.W * p.H, out_plane = (size_t)ws * hs;
  const int o00 = ty.i0 * p.W + tx.i0, o01 = ty.i0 * p.W + tx.i1;
  const int o10 = ty.i1 * p.W + tx.i0, o11 = ty.i1 * p.W + tx.i1;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    if (t >= p.nt) break;
    const int ch = t == 3 ? 1 : 3;
    float v00[3], v01[3], v10[3], v11[3];  // (all taps of a tensor requested before the first use)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c >= ch) break;
      const float* src = p.in[t] + ((size_t)b * ch + c) * in_plane;
      v00[c] = CDP_LDG(src + o00); v01[c] = CDP_LDG(src + o01); v10[c] = CDP_LDG(src + o10); v11[c] = CDP_LDG(src + o11);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c >= ch) break;
      if (t == 3 && p.depth_out) {  // the taps are disparities: resize the depth they stand for
        v00[c] = cdp_disp_to_depth(v00[c], p.min_disp, p.disp_span); v01[c] = cdp_disp_to_depth(v01[c], p.min_disp, p.disp_span);
        v10[c] = cdp_disp_to_depth(v10[c], p.min_disp, p.disp_span); v11[c] = cdp_disp_to_depth(v11[c], p.min_disp, p.disp_span);
      }
      const float top = v00[c] * tx.w0 + v01[c] * tx.w1;
      const float bot = v10[c] * tx.w0 + v11[c] * tx.w1;
      p.out[t][s][((size_t)b * ch + c) * out_plane + local] = top * ty.w0 + bot * ty.w1;
    }
  }
}

// ==========================================================================================
// 2. Fused photometric tile kernel: see cdp_photo_tile.h
// ==========================================================================================
#include "cdp_photo_tile.h"

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
  uint64_t* seed_dev;     // device seed counter of the built-in tie-break generator: bumped once per call (or null)
  int32_t B, blocks_per_image;
};

#define CDP_FINALIZE_THREADS 1024

// Block b, step 1: per-thread strided sums of image b's records into sm[r][32] (double precision).
// Block 0 additionally strides over ALL records' loss entries into sm[1024 + tid].
CDP_HD void cdp_finalize_phase_a(const CdpFinalizeParams& p, int b, int tid, double* sm) {
  const int r = tid >> 5, j = tid & 31;
  // eight independent accumulators (records r, r+32, ..., r+224 of every group of 256): the loads of
  // one group are in flight together instead of one memory round trip per record (the records were
  // written up to a kernel duration ago: most of them come from DRAM, not L2); fixed combination order
  const float* rec = p.partials + (size_t)b * p.blocks_per_image * CDP_PARTIAL_STRIDE + 1 + j;
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.0;
  for (int blk = r; blk < p.blocks_per_image; blk += 256) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = blk + 32 * k < p.blocks_per_image ? rec[(size_t)(blk + 32 * k) * CDP_PARTIAL_STRIDE] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += (double)v[k];
  }
  sm[r * 32 + j] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  if (b == 0) {
    double l[4] = {0.0, 0.0, 0.0, 0.0};
    const int total = p.B * p.blocks_per_image;
    for (int i = tid; i < total; i += 4 * CDP_FINALIZE_THREADS) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = i + k * CDP_FINALIZE_THREADS < total ? p.partials[(size_t)(i + k * CDP_FINALIZE_THREADS) * CDP_PARTIAL_STRIDE] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) l[k] += (double)v[k];
    }
    sm[1024 + tid] = (l[0] + l[1]) + (l[2] + l[3]);
  }
}
// step 2: combine in index order; threads 0..31 own one pose-gradient column, threads 32..63 of
// block 0 each fold 32 consecutive loss sums (sm needs 2048 + 32 doubles)
CDP_HD void cdp_finalize_phase_b(const CdpFinalizeParams& p, int b, int tid, double* sm) {
  if (tid < 32) {
    double acc = 0.0;
    for (int r = 0; r < 32; ++r) acc += sm[r * 32 + tid];
    if (p.pose_unit) p.pose_unit[((size_t)(tid >> 4) * p.B + b) * 16 + (tid & 15)] = (float)acc;
  } else if (tid < 64 && b == 0) {
    double acc = 0.0;
    for (int i = 0; i < 32; ++i) acc += sm[1024 + (tid - 32) * 32 + i];
    sm[2048 + tid - 32] = acc;
  }
}
// step 3: thread 0 of block 0 adds the 32 group sums (a 32 + 32 deep chain of fp64 additions
// instead of 1024: the serial version cost 8 us)
CDP_HD void cdp_finalize_phase_c(const CdpFinalizeParams& p, int b, int tid, const double* sm) {
  if (tid == 0 && b == 0) {
    double acc = 0.0;
    for (int i = 0; i < 32; ++i) acc += sm[2048 + i];
    p.loss[0] = (float)acc;
    if (p.seed_dev) *p.seed_dev += 1;  // next call (or graph replay) of the built-in generator draws fresh noise
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
  // level s halves the axis exactly s times (in = out << s): the resize taps are then the two
  // middle pixels of every 2^s block with weight 1/2, and the adjoint needs no table
  uint8_t exact_x[CDP_MAX_LEVELS], exact_y[CDP_MAX_LEVELS];
  const float* grad_loss;  // device scalar
  const float* pose_unit;  // [2][B][16]
  float* grad_depth;       // [B,1,H,W]
  float* grad_pose[2];     // [B,16]
  int32_t B, H, W, L;
  int32_t scale_pose;      // also write grad_pose = grad_loss * pose_unit (first launch only)
  // fused heads: grad_depth receives dL/d DISPARITY = -span * depth^2 * dL/d depth, and the first
  // block chains the scaled dL/dT through transformation_from_parameters instead of storing it
  const float* depth_vals; // [B,1,H,W] or null
  float disp_span;
  const float* axisangle[2];   // [B,3] or null
  const float* translation[2];
  int32_t invert[2];
  float* grad_axisangle[2];    // [B,3]
  float* grad_translation[2];
};

// references of input index i into level s along one axis: transpose of the bilinear resize taps
CDP_HD void cdp_inv_taps(bool exact, const CdpResizeInv* table, int i, int s, int& ja, int& jb, float& wa, float& wb) {
  if (exact) {
    const int r = 1 << s, m = i & (r - 1);
    ja = (m == (r >> 1) - 1 || m == (r >> 1)) ? (i >> s) : -1;
    wa = 0.5f; jb = -1; wb = 0.f;
  } else {
    const CdpResizeInv e = table[i];
    ja = e.ja; jb = e.jb; wa = e.wa; wb = e.wb;
  }
}

// contribution of level s to full-resolution pixel (x, y): (resize_s^T G_s)(x, y), table or exact taps per axis
CDP_HD float cdp_depth_grad_level(const CdpDepthGradParams& p, int b, int y, int x, int s) {
  int xa, xb, ya, yb;
  float wxa, wxb, wya, wyb;
  cdp_inv_taps(p.exact_y[s] != 0, p.inv_y[s], y, s, ya, yb, wya, wyb);
  if (ya < 0 && yb < 0) return 0.f;
  cdp_inv_taps(p.exact_x[s] != 0, p.inv_x[s], x, s, xa, xb, wxa, wxb);
  if (xa < 0 && xb < 0) return 0.f;
  const int ws = p.Ws[s];
  const float* g = p.gdepth[s] + (size_t)b * ws * p.Hs[s];
  float t = 0.f;
  if (ya >= 0) {
    const float* r = g + ya * ws;
    float u = 0.f;
    if (xa >= 0) u = wxa * CDP_LDG(r + xa);
    if (xb >= 0) u += wxb * CDP_LDG(r + xb);
    t = wya * u;
  }
  if (yb >= 0) {
    const float* r = g + yb * ws;
    float u = 0.f;
    if (xa >= 0) u = wxa * CDP_LDG(r + xa);
    if (xb >= 0) u += wxb * CDP_LDG(r + xb);
    t += wyb * u;
  }
  return t;
}

// dL/d depth at full-resolution pixel (x, y) = grad_loss * (G_0 + sum_s resize_s^T G_s)
CDP_HD void cdp_depth_grad_px(const CdpDepthGradParams& p, int b, int y, int x) {
  const int W = p.W, pix = y * W + x;
  float acc = CDP_LDG(p.gdepth[0] + (size_t)b * W * p.H + pix);
#pragma unroll
  for (int s = 1; s < CDP_MAX_LEVELS; ++s) {
    if (s >= p.L) break;
    acc += cdp_depth_grad_level(p, b, y, x, s);
  }
  float g = CDP_LDG(p.grad_loss) * acc;
  if (p.depth_vals) g = cdp_disp_to_depth_grad(g, CDP_LDG(p.depth_vals + (size_t)b * W * p.H + pix), p.disp_span);
  p.grad_depth[(size_t)b * W * p.H + pix] = g;
}

// Same for pyramids whose every level is an exact power-of-two reduction on both axes (the common
// case): level s contributes 1/4 of G_s(x >> s, y >> s) iff x and y are one of the two middle
// positions of their 2^s block.
CDP_HD void cdp_depth_grad_px_exact(const CdpDepthGradParams& p, int b, int y, int x) {
  const int W = p.W, pix = y * W + x;
  float acc = CDP_LDG(p.gdepth[0] + (size_t)b * W * p.H + pix);
#pragma unroll
  for (int s = 1; s < CDP_MAX_LEVELS; ++s) {
    if (s >= p.L) break;
    const int r = 1 << s, half = r >> 1;
    // m in {half-1, half}  <=>  (unsigned)(m - (half-1)) < 2
    const bool hy = (unsigned)((y & (r - 1)) - (half - 1)) < 2u;
    const bool hx = (unsigned)((x & (r - 1)) - (half - 1)) < 2u;
    if (hy && hx) acc += 0.25f * CDP_LDG(p.gdepth[s] + (size_t)b * p.Ws[s] * p.Hs[s] + (y >> s) * p.Ws[s] + (x >> s));
  }
  float g = CDP_LDG(p.grad_loss) * acc;
  if (p.depth_vals) g = cdp_disp_to_depth_grad(g, CDP_LDG(p.depth_vals + (size_t)b * W * p.H + pix), p.disp_span);
  p.grad_depth[(size_t)b * W * p.H + pix] = g;
}

// Four consecutive pixels x..x+3 (x % 4 == 0) of row y in one thread, for W % 4 == 0 and 16-byte
// aligned rows: one 16-byte load of G_0 and one 16-byte store.  Levels that halve both axes
// exactly s times take the short path -- one 8-byte load of G_1 (every position is a middle
// position of its 2x2 block), at most one scalar load per coarser level (the middle columns of a
// 2^s block are positions 2^(s-1)-1 and 2^(s-1): pixels 1,2 of the quad for s = 2, pixel 3 or
// pixel 0 of one quad each for s >= 3); other levels (e.g. 376 -> 23 rows at level 4 of KITTI-360)
// go through the transposed tap tables per pixel.  Same per-pixel accumulation order and values
// as cdp_depth_grad_px, so results are bit-identical.
// ALL_EXACT: every level is exact on both axes (no table path compiled in)
template <bool ALL_EXACT>
CDP_HD void cdp_depth_grad_quad(const CdpDepthGradParams& p, int b, int y, int x) {
  const int W = p.W;
  const size_t o0 = (size_t)b * W * p.H + (size_t)y * W + x;
  float4 a = CDP_LDG(reinterpret_cast<const float4*>(p.gdepth[0] + o0));
#pragma unroll
  for (int s = 1; s < CDP_MAX_LEVELS; ++s) {
    if (s >= p.L) break;
    if (!ALL_EXACT && !(p.exact_x[s] && p.exact_y[s])) {
      a.x += cdp_depth_grad_level(p, b, y, x, s);
      a.y += cdp_depth_grad_level(p, b, y, x + 1, s);
      a.z += cdp_depth_grad_level(p, b, y, x + 2, s);
      a.w += cdp_depth_grad_level(p, b, y, x + 3, s);
      continue;
    }
    if (s == 1) {
      const float2 g = CDP_LDG(reinterpret_cast<const float2*>(p.gdepth[1] + (size_t)b * p.Ws[1] * p.Hs[1] +
                                                               (size_t)(y >> 1) * p.Ws[1] + (x >> 1)));
      a.x += 0.25f * g.x; a.y += 0.25f * g.x; a.z += 0.25f * g.y; a.w += 0.25f * g.y;
      continue;
    }
    const int r = 1 << s, half = r >> 1;
    if ((unsigned)((y & (r - 1)) - (half - 1)) >= 2u) continue;  // not a middle row (uniform per block)
    const int m = x & (r - 1);  // position of the quad's first pixel inside its block
    if (s > 2 && m != half - 4 && m != half) continue;
    const float g = 0.25f * CDP_LDG(p.gdepth[s] + (size_t)b * p.Ws[s] * p.Hs[s] + (size_t)(y >> s) * p.Ws[s] + (x >> s));
    if (s == 2) { a.y += g; a.z += g; }
    else if (m == half) a.x += g;
    else a.w += g;
  }
  const float go = CDP_LDG(p.grad_loss);
  a.x *= go; a.y *= go; a.z *= go; a.w *= go;
  if (p.depth_vals) {
    const float4 dv = CDP_LDG(reinterpret_cast<const float4*>(p.depth_vals + o0));
    a.x = cdp_disp_to_depth_grad(a.x, dv.x, p.disp_span); a.y = cdp_disp_to_depth_grad(a.y, dv.y, p.disp_span);
    a.z = cdp_disp_to_depth_grad(a.z, dv.z, p.disp_span); a.w = cdp_disp_to_depth_grad(a.w, dv.w, p.disp_span);
  }
  *reinterpret_cast<float4*>(p.grad_depth + o0) = a;
}

// the quad form applies: rows of level 0 (and of an exact level 1) keep the vector accesses aligned
CDP_HD bool cdp_depth_grad_quad_ok(const CdpDepthGradParams& p) {
  if ((p.W & 3) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.gdepth[0]) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.grad_depth) & 15) != 0) return false;
  if (p.depth_vals && (reinterpret_cast<uintptr_t>(p.depth_vals) & 15) != 0) return false;
  if (p.L > 1 && p.exact_x[1] && p.exact_y[1] && (reinterpret_cast<uintptr_t>(p.gdepth[1]) & 7) != 0) return false;
  return true;
}

CDP_HD bool cdp_depth_grad_all_exact(const CdpDepthGradParams& p) {
  for (int s = 1; s < p.L; ++s)
    if (!p.exact_x[s] || !p.exact_y[s]) return false;
  return true;
}

CDP_HD void cdp_depth_grad_pixel(const CdpDepthGradParams& p, int b, int pix) {
  const int y = pix / p.W, x = pix - y * p.W;
  if (cdp_depth_grad_quad_ok(p)) {
    if ((x & 3) != 0) return;
    if (cdp_depth_grad_all_exact(p)) cdp_depth_grad_quad<true>(p, b, y, x);
    else cdp_depth_grad_quad<false>(p, b, y, x);
  } else if (cdp_depth_grad_all_exact(p)) {
    cdp_depth_grad_px_exact(p, b, y, x);
  } else {
    cdp_depth_grad_px(p, b, y, x);
  }
}

CDP_HD void cdp_pose_grad_scale(const CdpDepthGradParams& p, int i) {  // i in [0, 2*B*16)
  const int k = i / (p.B * 16), r = i - k * p.B * 16;
  p.grad_pose[k][r] = CDP_LDG(p.grad_loss) * p.pose_unit[i];
}

// ==========================================================================================
// 5. Edge-aware smoothness (algos/depth.py:58-107).
// ==========================================================================================
// |d^(p) - d^(q)| = |disp(p) - disp(q)| / (mean + eps): the per-image normalisation factors out of
// both sums, so ONE pass over the image produces the un-normalised edge sums, the un-normalised
// gradient field g and the disparity sum; the normalisation is applied in the tiny finalize step
// (loss) and in the backward kernel (gradient).  Each block handles a 62x14 pixel tile staged in
// shared memory with a one-pixel ring (64x16 positions, no index divisions), so every edge weight
// exp(-mean_c |dI|) is evaluated once.
#define CDP_SMOOTH_TX 62  // staged region is 64 x 16: thread t owns column t & 63, rows (t >> 6) + 4 i
#define CDP_SMOOTH_TY 14
#define CDP_SMOOTH_THREADS 256
#define CDP_SMOOTH_RW (CDP_SMOOTH_TX + 2)
#define CDP_SMOOTH_RH (CDP_SMOOTH_TY + 2)
#define CDP_SMOOTH_RN (CDP_SMOOTH_RW * CDP_SMOOTH_RH)
#define CDP_SMOOTH_SMEM_FLOATS (6 * CDP_SMOOTH_RN)  // image x3, disp, hx, hy

struct CdpSmoothParams {
  const float* image;  // [B,3,H,W]
  const float* disp;   // [B,1,H,W]
  float* g;            // [B,H,W]   un-normalised d loss / d normalised disparity
  float* part;         // [B][tiles][4]  sum disp, sum |dx disp| e_x, sum |dy disp| e_y, sum g*disp
  float* scal;         // [B][2]  |a_b|, sign(a_b) * sum(g*disp) a_b^2 / (H W),  a_b = 1/(mean_b + eps)
  float* loss;         // [1]
  int32_t B, H, W, with_grad;
  int32_t tiles_x, tiles_y;
};

CDP_HD float cdp_sign(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// phase 1: stage image and disparity of the tile + ring (zeros outside the image)
CDP_HD void cdp_smooth_phase_load(const CdpSmoothParams& p, int b, int tile, int tid, int nthreads, float* sm) {
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int x0 = tx * CDP_SMOOTH_TX - 1, y0 = ty * CDP_SMOOTH_TY - 1;
  const size_t plane = (size_t)p.H * p.W;
  const float* img = p.image + (size_t)b * 3 * plane;
  const float* d = p.disp + (size_t)b * plane;
  (void)nthreads;  // CDP_SMOOTH_THREADS
  const int rx = tid & 63, x = x0 + rx;
#pragma unroll
  for (int i = 0; i < CDP_SMOOTH_RH / 4; ++i) {
    const int ry = (tid >> 6) + 4 * i, idx = ry * CDP_SMOOTH_RW + rx;
    const int y = y0 + ry;
    const bool ok = x >= 0 && x < p.W && y >= 0 && y < p.H;
    const int o = y * p.W + x;
    sm[0 * CDP_SMOOTH_RN + idx] = ok ? CDP_LDG(img + o) : 0.f;
    sm[1 * CDP_SMOOTH_RN + idx] = ok ? CDP_LDG(img + plane + o) : 0.f;
    sm[2 * CDP_SMOOTH_RN + idx] = ok ? CDP_LDG(img + 2 * plane + o) : 0.f;
    sm[3 * CDP_SMOOTH_RN + idx] = ok ? CDP_LDG(d + o) : 0.f;
  }
}

// phase 2: signed edge weights hx = sign(disp(p) - disp(p+1x)) e_x(p), hy likewise, for every
// staged position that has its right / lower neighbour staged; per-thread loss sums over the
// tile proper: acc[1] += |dx disp| e_x, acc[2] += |dy disp| e_y   (algos/depth.py:79-87)
CDP_HD void cdp_smooth_phase_edges(const CdpSmoothParams& p, int tile, int tid, int nthreads, float* sm, float acc[4]) {
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int x0 = tx * CDP_SMOOTH_TX - 1, y0 = ty * CDP_SMOOTH_TY - 1;
  const float third = 1.0f / 3.0f;
  (void)nthreads;
  const int rx = tid & 63, x = x0 + rx;
#pragma unroll
  for (int i = 0; i < CDP_SMOOTH_RH / 4; ++i) {
    const int ry = (tid >> 6) + 4 * i, idx = ry * CDP_SMOOTH_RW + rx;
    const int y = y0 + ry;
    const bool ok = x >= 0 && x < p.W && y >= 0 && y < p.H;
    const bool own = ok && rx >= 1 && rx <= CDP_SMOOTH_TX && ry >= 1 && ry <= CDP_SMOOTH_TY;
    float hx = 0.f, hy = 0.f;
    const float dc = sm[3 * CDP_SMOOTH_RN + idx];
    if (ok && rx + 1 < CDP_SMOOTH_RW && x + 1 < p.W) {
      const float s = fabsf(sm[idx] - sm[idx + 1]) + fabsf(sm[CDP_SMOOTH_RN + idx] - sm[CDP_SMOOTH_RN + idx + 1]) +
                      fabsf(sm[2 * CDP_SMOOTH_RN + idx] - sm[2 * CDP_SMOOTH_RN + idx + 1]);
      const float e = cdp_exp(-(s * third));
      const float diff = dc - sm[3 * CDP_SMOOTH_RN + idx + 1];
      hx = cdp_sign(diff) * e;
      if (own) acc[1] += fabsf(diff) * e;
    }
    if (ok && ry + 1 < CDP_SMOOTH_RH && y + 1 < p.H) {
      const int dn = idx + CDP_SMOOTH_RW;
      const float s = fabsf(sm[idx] - sm[dn]) + fabsf(sm[CDP_SMOOTH_RN + idx] - sm[CDP_SMOOTH_RN + dn]) +
                      fabsf(sm[2 * CDP_SMOOTH_RN + idx] - sm[2 * CDP_SMOOTH_RN + dn]);
      const float e = cdp_exp(-(s * third));
      const float diff = dc - sm[3 * CDP_SMOOTH_RN + dn];
      hy = cdp_sign(diff) * e;
      if (own) acc[2] += fabsf(diff) * e;
    }
    sm[4 * CDP_SMOOTH_RN + idx] = hx;
    sm[5 * CDP_SMOOTH_RN + idx] = hy;
  }
}

// phase 3: g(p) = cx (hx(p) - hx(p-1x)) + cy (hy(p) - hy(p-1y)); acc[0] += disp, acc[3] += g disp
CDP_HD void cdp_smooth_phase_grad(const CdpSmoothParams& p, int b, int tile, int tid, int nthreads, const float* sm,
                                  float acc[4]) {
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int x0 = tx * CDP_SMOOTH_TX, y0 = ty * CDP_SMOOTH_TY;
  const float cx = 1.0f / ((float)p.B * (float)p.H * (float)(p.W - 1));
  const float cy = 1.0f / ((float)p.B * (float)(p.H - 1) * (float)p.W);
  (void)nthreads;
  const int lx = tid & 63, x = x0 + lx;
#pragma unroll
  for (int i = 0; i < (CDP_SMOOTH_TY + 3) / 4; ++i) {
    const int ly = (tid >> 6) + 4 * i, y = y0 + ly;
    if (lx >= CDP_SMOOTH_TX || ly >= CDP_SMOOTH_TY || x >= p.W || y >= p.H) continue;
    const int idx = (ly + 1) * CDP_SMOOTH_RW + lx + 1;
    const float dv = sm[3 * CDP_SMOOTH_RN + idx];
    acc[0] += dv;
    if (p.with_grad) {
      const float g = cx * (sm[4 * CDP_SMOOTH_RN + idx] - sm[4 * CDP_SMOOTH_RN + idx - 1]) +
                      cy * (sm[5 * CDP_SMOOTH_RN + idx] - sm[5 * CDP_SMOOTH_RN + idx - CDP_SMOOTH_RW]);
      p.g[(size_t)b * p.H * p.W + (size_t)y * p.W + x] = g;
      acc[3] += g * dv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row-walk form of the same pass for W % 4 == 0 and 16-byte aligned planes: a thread owns four
// consecutive pixels and walks CDP_SMOOTH_Q_ROWS rows down, the previous row and the pending
// gradient terms stay in registers, the two neighbours of the quad come from two scalar loads
// (L1 hits: they are the neighbouring threads' quads).  No shared memory, ~5x fewer instructions
// than the staged tile form above, which remains the general path.
// ------------------------------------------------------------------------------------------
#ifndef CDP_SMOOTH_Q_THREADS
#define CDP_SMOOTH_Q_THREADS 128
#endif
#ifndef CDP_SMOOTH_Q_ROWS
#define CDP_SMOOTH_Q_ROWS 16
#endif

struct CdpSmoothRow {  // positions x-1 .. x+4 of one row: image channels and disparity
  float c[4][6];
};

CDP_HD void cdp_smooth_load_row(const CdpSmoothParams& p, int b, int y, int x, CdpSmoothRow& r) {
  const size_t plane = (size_t)p.H * p.W;
  const size_t o = (size_t)y * p.W + x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float* src = k < 3 ? p.image + ((size_t)b * 3 + k) * plane : p.disp + (size_t)b * plane;
    const float4 v = CDP_LDG(reinterpret_cast<const float4*>(src + o));
    r.c[k][1] = v.x; r.c[k][2] = v.y; r.c[k][3] = v.z; r.c[k][4] = v.w;
    r.c[k][0] = x > 0 ? CDP_LDG(src + o - 1) : 0.f;
    r.c[k][5] = x + 4 < p.W ? CDP_LDG(src + o + 4) : 0.f;
  }
}

// signed edge weight sign(d(p) - d(q)) exp(-mean_c |I(p) - I(q)|) and the edge's loss term
CDP_HD float cdp_smooth_edge(float r0, float g0, float b0, float d0, float r1, float g1, float b1, float d1,
                             float& loss_term) {
  const float s = fabsf(r0 - r1) + fabsf(g0 - g1) + fabsf(b0 - b1);
  const float e = cdp_exp(-(s * (1.0f / 3.0f)));
  const float diff = d0 - d1;
  loss_term = fabsf(diff) * e;
  return cdp_sign(diff) * e;
}

// thread `tid` of block (bx, by) of image b; acc = {sum disp, sum x-edge loss, sum y-edge loss, sum g*disp}
CDP_HD void cdp_smooth_quad_thread(const CdpSmoothParams& p, int b, int bx, int by, int tid, float acc[4]) {
  const int x = (bx * CDP_SMOOTH_Q_THREADS + tid) * 4;
  if (x >= p.W) return;
  const int y0 = by * CDP_SMOOTH_Q_ROWS, y1 = y0 + CDP_SMOOTH_Q_ROWS < p.H ? y0 + CDP_SMOOTH_Q_ROWS : p.H;
  const float cx = 1.0f / ((float)p.B * (float)p.H * (float)(p.W - 1));
  const float cy = 1.0f / ((float)p.B * (float)(p.H - 1) * (float)p.W);
  float prev[4][4];       // row t-1, the quad's own pixels
  float xpart[4], hy1[4], hy2[4];  // row t-1: cx (hx(x) - hx(x-1)); hy of rows t-1 and t-2
#pragma unroll
  for (int i = 0; i < 4; ++i) { xpart[i] = 0.f; hy1[i] = 0.f; hy2[i] = 0.f; prev[0][i] = prev[1][i] = prev[2][i] = prev[3][i] = 0.f; }
  const int t0 = y0 > 0 ? y0 - 1 : 0;
  CdpSmoothRow cur, nxt;
  cdp_smooth_load_row(p, b, t0, x, nxt);
  for (int t = t0; t <= y1; ++t) {  // row t is in registers; row t-1 is completed
    cur = nxt;
    if (t + 1 <= y1 && t + 1 < p.H) cdp_smooth_load_row(p, b, t + 1, x, nxt);  // requested one row ahead
    float xp[4] = {0.f, 0.f, 0.f, 0.f};
    const bool have = t < p.H;
    if (have) {
      if (t >= y0 && t < y1) {  // x edges of an owned row: positions (x-1|x) .. (x+3|x+4)
        float hx[5], lt;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const bool edge = x + i - 1 >= 0 && x + i < p.W;
          hx[i] = 0.f;
          if (edge) {
            hx[i] = cdp_smooth_edge(cur.c[0][i], cur.c[1][i], cur.c[2][i], cur.c[3][i], cur.c[0][i + 1], cur.c[1][i + 1],
                                    cur.c[2][i + 1], cur.c[3][i + 1], lt);
            if (i >= 1) acc[1] += lt;  // the edge (x-1|x) is owned by the quad to the left
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) xp[i] = cx * (hx[i + 1] - hx[i]);
      }
    }
    // y edges (t-1 | t)
#pragma unroll
    for (int i = 0; i < 4; ++i) hy2[i] = hy1[i];
    if (have && t > t0) {
      float lt;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hy1[i] = cdp_smooth_edge(prev[0][i], prev[1][i], prev[2][i], prev[3][i], cur.c[0][i + 1], cur.c[1][i + 1],
                                 cur.c[2][i + 1], cur.c[3][i + 1], lt);
        if (t - 1 >= y0) acc[2] += lt;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) hy1[i] = 0.f;  // no row below the last image row / nothing above the first loaded row
    }
    // complete row t-1
    if (t - 1 >= y0 && t - 1 < y1 && t > t0) {
      float4 g;
      float gv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        gv[i] = xpart[i] + cy * (hy1[i] - hy2[i]);
        acc[0] += prev[3][i];
        acc[3] += gv[i] * prev[3][i];
      }
      if (p.with_grad) {
        g.x = gv[0]; g.y = gv[1]; g.z = gv[2]; g.w = gv[3];
        *reinterpret_cast<float4*>(p.g + (size_t)b * p.H * p.W + (size_t)(t - 1) * p.W + x) = g;
      }
    }
    if (have) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xpart[i] = xp[i];
        prev[0][i] = cur.c[0][i + 1]; prev[1][i] = cur.c[1][i + 1]; prev[2][i] = cur.c[2][i + 1]; prev[3][i] = cur.c[3][i + 1];
      }
    }
  }
}

CDP_HD bool cdp_smooth_quad_ok(const CdpSmoothParams& p) {
  return (p.W & 3) == 0 && (reinterpret_cast<uintptr_t>(p.image) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.disp) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.g) & 15) == 0;
}

// Fixed-order combination of per-block records that one warp can evaluate in parallel: lane l sums
// records l, l+32, ...; lanes are combined by a butterfly (cdp_butterfly_host = same order).
CDP_HD double cdp_lane_sum(const float* part, int count, int stride, int lane) {
  double acc = 0.0;
  for (int i = lane; i < count; i += 32) acc += (double)part[(size_t)i * stride];
  return acc;
}
CDP_HD double cdp_butterfly_host(double v[32]) {
  for (int off = 16; off > 0; off >>= 1)
    for (int l = 0; l < off; ++l) v[l] += v[l + off];
  return v[0];
}

// per image: normalisation scalars; returns the image's contribution to the loss
CDP_HD double cdp_smooth_finalize_image(const CdpSmoothParams& p, int b, double sum_disp, double sx, double sy,
                                        double gd) {
  const double hw = (double)((size_t)p.H * p.W);
  const float mean = (float)(sum_disp / hw);
  const double a = 1.0 / (double)(mean + 1e-7f);  // mean_disparity + 1e-7, algos/depth.py:104-105
  const double aa = a < 0 ? -a : a;
  if (p.with_grad) {
    p.scal[b * 2 + 0] = (float)aa;
    p.scal[b * 2 + 1] = (float)((a < 0 ? -1.0 : 1.0) * gd * a * a / hw);
  }
  const double nx = (double)p.B * p.H * (p.W - 1), ny = (double)p.B * (p.H - 1) * p.W;
  return aa * (sx / nx + sy / ny);
}

// host / single-thread form of the whole finalize step
CDP_HD void cdp_smooth_finalize(const CdpSmoothParams& p) {
  const int nb = p.tiles_x * p.tiles_y;
  double loss = 0.0;
  for (int b = 0; b < p.B; ++b) {
    double v[4][32];
    for (int q = 0; q < 4; ++q) {
      for (int l = 0; l < 32; ++l) v[q][l] = cdp_lane_sum(p.part + ((size_t)b * nb) * 4 + q, nb, 4, l);
    }
    const double s0 = cdp_butterfly_host(v[0]), s1 = cdp_butterfly_host(v[1]);
    const double s2 = cdp_butterfly_host(v[2]), s3 = cdp_butterfly_host(v[3]);
    loss += cdp_smooth_finalize_image(p, b, s0, s1, s2, s3);
  }
  p.loss[0] = (float)loss;
}

// grad_disp = grad_loss * (g * |a_b| - c_b) for `n` (<= 4) consecutive pixels starting at i (multiple of 4)
CDP_HD void cdp_smooth_bwd_run(const float* g, const float* scal, const float* grad_loss, int b,
                               size_t plane, int i, int n, float* grad_disp) {
  const float go = CDP_LDG(grad_loss), a = scal[b * 2], c = scal[b * 2 + 1];
  const float* src = g + (size_t)b * plane + i;
  float* dst = grad_disp + (size_t)b * plane + i;
  if (n == 4 && (plane & 3) == 0) {
    const float4 v = CDP_LDG(reinterpret_cast<const float4*>(src));
    float4 o; o.x = go * (v.x * a - c); o.y = go * (v.y * a - c); o.z = go * (v.z * a - c); o.w = go * (v.w * a - c);
    *reinterpret_cast<float4*>(dst) = o;
  } else {
    for (int k = 0; k < n; ++k) dst[k] = go * (CDP_LDG(src + k) * a - c);
  }
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

CDP_HD void cdp_warp_setup(const CdpWarpParams& p, int b_local, int pix, CdpWarp& w, CdpCam& cam, CdpPose& T) {
  const int b = p.batch_begin + b_local;
  const size_t plane = (size_t)p.H * p.W;
  const int y = pix / p.W, x = pix - y * p.W;
  cam = cdp_make_cam(p.K[b_local][0], p.K[b_local][1], p.K[b_local][2], p.K[b_local][3]);
  cdp_load_pose(p.pose + (size_t)b * 16, T);
  float mo[3];
  if (p.motion) {
    for (int c = 0; c < 3; ++c) mo[c] = CDP_LDG(p.motion + ((size_t)b * 3 + c) * plane + pix);
  }
  cdp_warp_point((float)x, (float)y, CDP_LDG(p.depth + (size_t)b * plane + pix), cam, T, p.motion ? mo : nullptr, w);
}

CDP_HD void cdp_warp_grid_pixel(const CdpWarpParams& p, int b_local, int pix) {
  CdpWarp w; CdpCam cam; CdpPose T;
  cdp_warp_setup(p, b_local, pix, w, cam, T);
  const size_t o = (((size_t)(p.batch_begin + b_local)) * p.H * p.W + pix) * 2;
  // normalisation of _PointcloudToImage (misc/image_warper.py:44-45)
  p.out[o] = (w.ix / (float)(p.W - 1) - 0.5f) * 2.0f;
  p.out[o + 1] = (w.iy / (float)(p.H - 1) - 0.5f) * 2.0f;
}

CDP_HD void cdp_warp_image_pixel(const CdpWarpParams& p, int b_local, int pix) {
  CdpWarp w; CdpCam cam; CdpPose T;
  cdp_warp_setup(p, b_local, pix, w, cam, T);
  const int b = p.batch_begin + b_local;
  const size_t plane = (size_t)p.H * p.W;
  if (p.mode == 0) {
    CdpTaps t;
    cdp_taps(pix % p.W, pix / p.W, w, p.W, p.H, t);
    for (int c = 0; c < p.C; ++c)
      p.out[((size_t)b * p.C + c) * plane + pix] = cdp_bilinear(p.src + ((size_t)b * p.C + c) * plane, t);
  } else {
    const int o = cdp_nearest_index(w.ix, w.iy, p.W, p.H);
    for (int c = 0; c < p.C; ++c)
      p.out[((size_t)b * p.C + c) * plane + pix] = CDP_LDG(p.src + ((size_t)b * p.C + c) * plane + o);
  }
}

// backward of the bilinear warp for one pixel; accumulates this thread's dT[16]
CDP_HD void cdp_warp_bwd_pixel(const CdpWarpParams& p, int b_local, int pix, float* dT) {
  CdpWarp w; CdpCam cam; CdpPose T;
  cdp_warp_setup(p, b_local, pix, w, cam, T);
  const int b = p.batch_begin + b_local;
  const size_t plane = (size_t)p.H * p.W;
  CdpTaps t;
  cdp_taps(pix % p.W, pix / p.W, w, p.W, p.H, t);
  float gix = 0.f, giy = 0.f;
  for (int c = 0; c < p.C; ++c) {
    float dix, diy;
    cdp_bilinear_grad(p.src + ((size_t)b * p.C + c) * plane, t, dix, diy);
    const float go = CDP_LDG(p.grad_out + ((size_t)b * p.C + c) * plane + pix);
    gix += go * dix;
    giy += go * diy;
  }
  float gd = 0.f, gm[3];
  cdp_warp_adjoint(gix * t.mx, giy * t.my, w, cam, T, gd, dT, p.grad_motion ? gm : nullptr);
  p.grad_depth[(size_t)b * plane + pix] = gd;
  if (p.grad_motion)
    for (int c = 0; c < 3; ++c) p.grad_motion[((size_t)b * 3 + c) * plane + pix] = gm[c];
}

// SSIM statistics for one pixel of one plane straight from global memory: reflect-padded 3x3
// window, values centred on the target value at the window centre (returned in c).
CDP_HD void cdp_ssim_stats_global(const float* x, const float* y, int W, int H, int px, int py,
                                  float& mxc, float& myc, float& exx, float& eyy, float& exy, float& c) {
  c = CDP_LDG(y + py * W + px);
  float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int o = cdp_reflect(py + dy, H) * W + cdp_reflect(px + dx, W);
      const float a = CDP_LDG(x + o) - c, b = CDP_LDG(y + o) - c;
      sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
    }
  const float ninth = 1.0f / 9.0f;
  mxc = sx * ninth; myc = sy * ninth; exx = sxx * ninth; eyy = syy * ninth; exy = sxy * ninth;
}

CDP_HD void cdp_ssim_fwd_pixel(const float* x, const float* y, int W, int H, int plane_idx, int pix,
                               float* out) {
  const size_t base = (size_t)plane_idx * W * H;
  float mxc, myc, exx, eyy, exy, c;
  cdp_ssim_stats_global(x + base, y + base, W, H, pix % W, pix / W, mxc, myc, exx, eyy, exy, c);
  CdpSsimTerms t;
  cdp_ssim_terms(mxc, myc, exx, eyy, exy, c, t);
  out[base + pix] = t.loss;
}

// backward pass 1: coefficient fields scaled by the upstream gradient -> scratch[4][planes*H*W]
// (A0_x, A0_y, B, C in the centred form of cdp_ssim_coeffs)
CDP_HD void cdp_ssim_bwd_coef_pixel(const float* grad_out, const float* x, const float* y, int W,
                                    int H, int plane_idx, int pix, float* scratch, size_t total) {
  const size_t base = (size_t)plane_idx * W * H;
  float mxc, myc, exx, eyy, exy, c;
  cdp_ssim_stats_global(x + base, y + base, W, H, pix % W, pix / W, mxc, myc, exx, eyy, exy, c);
  CdpSsimTerms t;
  cdp_ssim_terms(mxc, myc, exx, eyy, exy, c, t);
  const float dxq = (CDP_LDG(x + base + pix) - c) - mxc, dyq = (CDP_LDG(y + base + pix) - c) - myc;
  float Ax, Bx, C, Ay, By;
  cdp_ssim_coeffs(t, dxq, dyq, Ax, Bx, C);
  CdpSsimTerms ts = t;  // SSIM is symmetric in (x, y): swap the roles for d/dy
  ts.mx = t.my; ts.my = t.mx;
  cdp_ssim_coeffs(ts, dyq, dxq, Ay, By, C);
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
  if (grad_x) grad_x[base + pix] = gx * (1.0f / 9.0f);
  if (grad_y) grad_y[base + pix] = gy * (1.0f / 9.0f);
}

// ==========================================================================================
// 7. The two conversions that feed the loss: PoseHead.transformation_from_parameters
//    (models/pose_head.py:56-137) and DepthHead.disp_to_depth (models/depth_head.py:49-54).
// ==========================================================================================
struct CdpRodrigues {
  float x, y, z, ca, sa, C, theta, n;
  float R[9];
};

CDP_HD void cdp_rodrigues(const float v[3], CdpRodrigues& o) {
  o.theta = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  o.n = o.theta + 1e-7f;  // axis = axisangle / (angle + 1e-7), pose_head.py:85
  o.x = v[0] / o.n; o.y = v[1] / o.n; o.z = v[2] / o.n;
  o.ca = cosf(o.theta); o.sa = sinf(o.theta); o.C = 1.0f - o.ca;
  const float x = o.x, y = o.y, z = o.z, C = o.C, sa = o.sa, ca = o.ca;
  o.R[0] = x * (x * C) + ca; o.R[1] = x * (y * C) - z * sa; o.R[2] = z * (x * C) + y * sa;
  o.R[3] = x * (y * C) + z * sa; o.R[4] = y * (y * C) + ca; o.R[5] = y * (z * C) - x * sa;
  o.R[6] = z * (x * C) - y * sa; o.R[7] = y * (z * C) + x * sa; o.R[8] = z * (z * C) + ca;
}

// one sample: M = T(t) R, or (invert) M = R^T T(-t)
CDP_HD void cdp_pose_fwd_sample(const float* axisangle, const float* translation, int invert, float* M) {
  const float v[3] = {axisangle[0], axisangle[1], axisangle[2]};
  const float t[3] = {translation[0], translation[1], translation[2]};
  CdpRodrigues q;
  cdp_rodrigues(v, q);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M[4 * r + c] = invert ? q.R[3 * c + r] : q.R[3 * r + c];
    M[4 * r + 3] = invert ? -(q.R[0 + r] * t[0] + q.R[3 + r] * t[1] + q.R[6 + r] * t[2]) : t[r];
  }
  M[12] = 0.f; M[13] = 0.f; M[14] = 0.f; M[15] = 1.f;
}

CDP_HD void cdp_pose_bwd_sample(const float* gM, const float* axisangle, const float* translation, int invert,
                                float* g_axisangle, float* g_translation) {
  const float v[3] = {axisangle[0], axisangle[1], axisangle[2]};
  const float t[3] = {translation[0], translation[1], translation[2]};
  CdpRodrigues q;
  cdp_rodrigues(v, q);
  float gR[9], gt[3];
  if (!invert) {
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) gR[3 * r + c] = gM[4 * r + c];
      gt[r] = gM[4 * r + 3];
    }
  } else {
    // M[:3,:3] = R^T ; M[:3,3] = -R^T t
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) gR[3 * r + c] = gM[4 * c + r] - gM[4 * c + 3] * t[r];
    for (int j = 0; j < 3; ++j) gt[j] = -(q.R[3 * j + 0] * gM[3] + q.R[3 * j + 1] * gM[7] + q.R[3 * j + 2] * gM[11]);
  }
  const float x = q.x, y = q.y, z = q.z, C = q.C, sa = q.sa;
  float gx = 0.f, gy = 0.f, gz = 0.f, gC = 0.f, gsa = 0.f, gca = 0.f;
  // diagonal: a^2 C + ca
  gx += 2.f * x * C * gR[0]; gC += x * x * gR[0]; gca += gR[0];
  gy += 2.f * y * C * gR[4]; gC += y * y * gR[4]; gca += gR[4];
  gz += 2.f * z * C * gR[8]; gC += z * z * gR[8]; gca += gR[8];
  // R01 = xyC - z sa, R10 = xyC + z sa
  { const float s2 = gR[1] + gR[3], d = gR[3] - gR[1];
    gx += y * C * s2; gy += x * C * s2; gC += x * y * s2; gz += sa * d; gsa += z * d; }
  // R02 = zxC + y sa, R20 = zxC - y sa
  { const float s2 = gR[2] + gR[6], d = gR[2] - gR[6];
    gz += x * C * s2; gx += z * C * s2; gC += z * x * s2; gy += sa * d; gsa += y * d; }
  // R12 = yzC - x sa, R21 = yzC + x sa
  { const float s2 = gR[5] + gR[7], d = gR[7] - gR[5];
    gy += z * C * s2; gz += y * C * s2; gC += y * z * s2; gx += sa * d; gsa += x * d; }
  float gtheta = -sa * gca + q.ca * gsa + sa * gC;     // ca = cos, sa = sin, C = 1 - cos
  const float inv_n = 1.0f / q.n;
  float gv[3] = {gx * inv_n, gy * inv_n, gz * inv_n};  // a = v / n
  gtheta += -(gx * v[0] + gy * v[1] + gz * v[2]) * inv_n * inv_n;  // n = theta + eps
  if (q.theta > 0.f) {                                 // torch.norm backward is 0 at the origin
    const float k = gtheta / q.theta;
    gv[0] += k * v[0]; gv[1] += k * v[1]; gv[2] += k * v[2];
  }
  for (int i = 0; i < 3; ++i) { g_axisangle[i] = gv[i]; g_translation[i] = gt[i]; }
}

// ==========================================================================================
// 8. Object-motion regularisers (FlowSmoothnessLoss / FlowSparsityLoss): see cdp_flow.h
// ==========================================================================================
#include "cdp_flow.h"

// ==========================================================================================
// 9. Camera-to-camera warp at constant depth (Mixup.warp_c2c): see cdp_c2c.h
// ==========================================================================================
#include "cdp_c2c.h"

// ==========================================================================================
// 10. Depth metrics (DepthEvaluator.compute_depth_metrics): see cdp_metrics.h
// ==========================================================================================
#include "cdp_metrics.h"

// ==========================================================================================
// 11. Fused heads (cdp_photo_heads): pose matrices in the pyramid launch, the chain back to the
//     6-DoF parameters in the depth-gradient launch.
// ==========================================================================================
// item j in [0, 2*B): source k = j / B, sample b = j % B
CDP_HD void cdp_pyr_pose_item(const CdpPyrParams& p, int j) {
  const int k = j / p.B, b = j - k * p.B;
  cdp_pose_fwd_sample(p.axisangle[k] + 3 * b, p.translation[k] + 3 * b, p.invert[k], p.pose_out[k] + 16 * b);
}

CDP_HD void cdp_pose_grad_heads(const CdpDepthGradParams& p, int j) {
  const int k = j / p.B, b = j - k * p.B;
  float gM[16];
  const float go = CDP_LDG(p.grad_loss);
  for (int i = 0; i < 16; ++i) gM[i] = go * p.pose_unit[((size_t)k * p.B + b) * 16 + i];
  cdp_pose_bwd_sample(gM, p.axisangle[k] + 3 * b, p.translation[k] + 3 * b, p.invert[k], p.grad_axisangle[k] + 3 * b,
                      p.grad_translation[k] + 3 * b);
}

