// cdp_math.h -- per-pixel math of the photometric loss (host/device).
//
// Forward follows the reference's operation order (ray -> normalise -> scale by depth -> T ->
// /w -> clamp -> project -> normalise -> un-normalise -> clip -> bilinear), SURVEY.md section 3.3;
// backward is the closed form of SURVEY.md section 7 / 8a.
#pragma once

#include "cdp_common.h"

// ------------------------------------------------------------------------------------------
// Fast scalar helpers.  The kernels do not try to reproduce the reference's fp32 rounding
// operation by operation (two fp32 evaluations of this loss differ by ~1e-4 of max-abs in the
// gradients anyway, see tests/helpers.py); instead every quantity is computed in the
// best-conditioned form available, so that results sit closer to the fp64 evaluation of the
// reference than the reference's own fp32 run does.
// ------------------------------------------------------------------------------------------
#ifndef CDP_OPT_FAST_RCP
#define CDP_OPT_FAST_RCP 1  // MUFU.RCP (~1 ulp) instead of the correctly rounded reciprocal: -4 % kernel time
#endif
CDP_HD float cdp_rcp(float x) {
#if defined(__CUDA_ARCH__) && CDP_OPT_FAST_RCP
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#elif defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
#ifndef CDP_OPT_FAST_DIV
#define CDP_OPT_FAST_DIV 1  // __fdividef (~2 ulp) in the SSIM ratio; 0 = IEEE division (A/B record: profiles/r02_parity.json)
#endif
#ifndef CDP_OPT_FAST_EXP
#define CDP_OPT_FAST_EXP 1  // __expf in the smoothness edge weights; 0 = expf
#endif
CDP_HD float cdp_fdiv(float a, float b) {  // a / b to ~2 ulp
#if defined(__CUDA_ARCH__) && CDP_OPT_FAST_DIV
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
// clamp(1/2 + a / b, 0, 1) for the SSIM loss, a = half the numerator: MUFU.RCP and ONE saturating FMA.
// (__fdividef expands to a range check, two predicated rescalings, MUFU.RCP and a multiply -- five
// issue slots per division, three of them on the fp32 pipe; the SSIM denominators are >= 7e-6, so the
// rescaling never triggers.)  One rounding fewer than divide-then-FMA.
CDP_HD float cdp_half_plus_ratio_sat(float a, float b) {
#if defined(__CUDA_ARCH__) && CDP_OPT_FAST_DIV
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return __saturatef(__fmaf_rn(a, r, 0.5f));
#elif defined(__CUDA_ARCH__)
  return __saturatef(a / b + 0.5f);
#else
  return fminf(fmaxf(a / b + 0.5f, 0.f), 1.f);
#endif
}
CDP_HD float cdp_exp(float x) {
#if defined(__CUDA_ARCH__) && CDP_OPT_FAST_EXP
  return __expf(x);
#else
  return expf(x);
#endif
}

CDP_HD float cdp_fmaf(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return a * b + c;  // the emulator is built with -ffp-contract=off (as for cdp_fma2)
#endif
}

CDP_HD float cdp_saturate(float x) {  // clamp to [0, 1]
#if defined(__CUDA_ARCH__)
  return __saturatef(x);
#else
  return fminf(fmaxf(x, 0.f), 1.f);
#endif
}

// Packed fp32 pairs: on sm_100a these are single FADD2 / FMUL2 / FFMA2 instructions working on
// an aligned register pair (two pixels' or two candidates' worth of math per issue slot).
CDP_HD float2 cdp_set2(float v) { float2 r; r.x = v; r.y = v; return r; }
CDP_HD float2 cdp_add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fadd2_rn(a, b);
#else
  float2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r;
#endif
}
CDP_HD float2 cdp_mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fmul2_rn(a, b);
#else
  float2 r; r.x = a.x * b.x; r.y = a.y * b.y; return r;
#endif
}
CDP_HD float2 cdp_fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(a, b, c);
#else
  float2 r; r.x = a.x * b.x + c.x; r.y = a.y * b.y + c.y; return r;
#endif
}

CDP_HD CdpCam cdp_make_cam(float fx, float fy, float cx, float cy) {
  CdpCam k;
  k.fx = fx; k.fy = fy; k.cx = cx; k.cy = cy;
  k.ifx = cdp_rcp(fx); k.ify = cdp_rcp(fy);
  return k;
}

// Pose in the form the warp uses: row-major 4x4 with the identity removed from the rotation
// diagonal (D = T - diag(1,1,1,0)), so that small motions keep full relative precision.
struct CdpPose {
  float d[16];
};
CDP_HD void cdp_load_pose(const float* T, CdpPose& o) {
#pragma unroll
  for (int i = 0; i < 16; ++i) o.d[i] = CDP_LDG(T + i);
  o.d[0] -= 1.0f; o.d[5] -= 1.0f; o.d[10] -= 1.0f;
}
// same for a 16-byte aligned matrix: four 16-byte loads instead of sixteen scalar ones
CDP_HD void cdp_load_pose_aligned(const float* T, CdpPose& o) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float4 v = CDP_LDG(reinterpret_cast<const float4*>(T) + r);
    o.d[4 * r + 0] = v.x; o.d[4 * r + 1] = v.y; o.d[4 * r + 2] = v.z; o.d[4 * r + 3] = v.w;
  }
  o.d[0] -= 1.0f; o.d[5] -= 1.0f; o.d[10] -= 1.0f;
}

// ------------------------------------------------------------------------------------------
// Warp of one pixel: CameraModel.get_viewing_ray + _ImageToPointcloud (misc/camera_model.py:52-71,
// misc/image_warper.py:68-87), CoordinateWarper (misc/image_warper.py:100-144),
// _PointcloudToImage / get_image_point (misc/image_warper.py:20-51, misc/camera_model.py:43-50)
// and grid_sample's un-normalisation.
//
// The reference normalises the viewing ray and rescales it by depth/|ray_z|, which is D*(rx,ry,1)
// with rx = (u-cx)/fx; it then forms absolute pixel coordinates u' = fx*E_x/z + cx and pushes them
// through normalise / un-normalise, all in fp32, which leaves ~1e-4 px of rounding at u ~ 1000.
// Here the sample position is computed as a displacement from the pixel itself:
//     Q = P + a,  a = (T - I) P + t (+ motion),     u' - u = fx * (a_x - rx * a_z) / Q_z
// (Q_w cancels whenever the depth clamp max(E_z, 1e-5) is inactive), which is exact to ~1e-7 of
// the displacement.  The clamped / degenerate case falls back to the literal formula.
// ------------------------------------------------------------------------------------------
struct CdpWarp {
  float rx, ry;        // viewing ray (x/z, y/z)
  float P[3];          // back-projected point
  float Q[4];          // transformed homogeneous point
  float iz;            // 1 / Q_z (regular) or 1 / zt (clamped)
  float ix, iy;        // absolute sample position in source pixels (un-clipped)
  float dx, dy;        // ix - u, iy - v at full relative precision (regular case)
  bool regular;        // depth clamp inactive and Q_w != 0
};

CDP_HD void cdp_warp_point(float u, float v, float depth, const CdpCam& k, const CdpPose& T,
                           const float* motion3, CdpWarp& o) {
  o.rx = (u - k.cx) * k.ifx;
  o.ry = (v - k.cy) * k.ify;
  o.P[0] = depth * o.rx; o.P[1] = depth * o.ry; o.P[2] = depth;
  // (explicit FMA order: the packed two-source version below evaluates the same expression tree)
  float ax = cdp_fmaf(T.d[0], o.P[0], cdp_fmaf(T.d[1], o.P[1], cdp_fmaf(T.d[2], o.P[2], T.d[3])));
  float ay = cdp_fmaf(T.d[4], o.P[0], cdp_fmaf(T.d[5], o.P[1], cdp_fmaf(T.d[6], o.P[2], T.d[7])));
  float az = cdp_fmaf(T.d[8], o.P[0], cdp_fmaf(T.d[9], o.P[1], cdp_fmaf(T.d[10], o.P[2], T.d[11])));
  if (motion3) { ax += motion3[0]; ay += motion3[1]; az += motion3[2]; }
  o.Q[3] = cdp_fmaf(T.d[12], o.P[0], cdp_fmaf(T.d[13], o.P[1], cdp_fmaf(T.d[14], o.P[2], T.d[15])));
  o.Q[0] = o.P[0] + ax; o.Q[1] = o.P[1] + ay; o.Q[2] = o.P[2] + az;
  // E_z = Q_z / Q_w >= 1e-5 without dividing
  o.regular = o.Q[3] > 0.f ? (o.Q[2] >= CDP_Z_MIN * o.Q[3]) : (o.Q[3] < 0.f && o.Q[2] <= CDP_Z_MIN * o.Q[3]);
  if (o.regular) {
    o.iz = cdp_rcp(o.Q[2]);
    o.dx = (k.fx * cdp_fmaf(-o.rx, az, ax)) * o.iz;
    o.dy = (k.fy * cdp_fmaf(-o.ry, az, ay)) * o.iz;
    o.ix = u + o.dx;
    o.iy = v + o.dy;
  } else {  // literal reference formula: E = Q_xyz / Q_w, z = max(E_z, 1e-5)
    const float ex = o.Q[0] / o.Q[3], ey = o.Q[1] / o.Q[3], ez = o.Q[2] / o.Q[3];
    const float zt = fmaxf(ez, CDP_Z_MIN);
    o.iz = 1.0f / zt;
    o.ix = ex * o.iz * k.fx + k.cx;
    o.iy = ey * o.iz * k.fy + k.cy;
    o.dx = o.ix - u;
    o.dy = o.iy - v;
  }
}

// Both source frames at once, one per lane of packed fp32 (FFMA2 / FADD2 / FMUL2): sample
// displacement and absolute sample position only (what the forward gather needs).  Lane results
// are bit-identical to cdp_warp_point's regular branch; a lane whose depth clamp is active
// (regular[k] false) must be redone with cdp_warp_point.
struct CdpPose2 {
  float2 d[16];  // lane x = source 0 (t-1), lane y = source 1 (t+1); T - diag(1,1,1,0) as in CdpPose
};
CDP_HD void cdp_pack_pose(const CdpPose& a, const CdpPose& b, CdpPose2& o) {
#pragma unroll
  for (int i = 0; i < 16; ++i) { o.d[i].x = a.d[i]; o.d[i].y = b.d[i]; }
}
CDP_HD void cdp_unpack_pose(const CdpPose2& p, int k, CdpPose& o) {
#pragma unroll
  for (int i = 0; i < 16; ++i) o.d[i] = k == 0 ? p.d[i].x : p.d[i].y;
}
struct CdpWarp2 {
  float2 dx, dy, ix, iy;
  bool regular[2];  // (only ever indexed with constants: stays in predicate registers)
  // for the adjoint (regular lanes): viewing ray, back-projected point, transformed point, 1 / Q_z
  float rx, ry, P[3];
  float2 Qx, Qy, iz;
};
CDP_HD void cdp_warp_point2(float u, float v, float depth, const CdpCam& k, const CdpPose2& T,
                            const float2* motion3, CdpWarp2& o) {
  o.rx = (u - k.cx) * k.ifx; o.ry = (v - k.cy) * k.ify;
  o.P[0] = depth * o.rx; o.P[1] = depth * o.ry; o.P[2] = depth;
  const float2 Px = cdp_set2(o.P[0]), Py = cdp_set2(o.P[1]), Pz = cdp_set2(o.P[2]);
  float2 ax = cdp_fma2(T.d[0], Px, cdp_fma2(T.d[1], Py, cdp_fma2(T.d[2], Pz, T.d[3])));
  float2 ay = cdp_fma2(T.d[4], Px, cdp_fma2(T.d[5], Py, cdp_fma2(T.d[6], Pz, T.d[7])));
  float2 az = cdp_fma2(T.d[8], Px, cdp_fma2(T.d[9], Py, cdp_fma2(T.d[10], Pz, T.d[11])));
  if (motion3) { ax = cdp_add2(ax, motion3[0]); ay = cdp_add2(ay, motion3[1]); az = cdp_add2(az, motion3[2]); }
  const float2 Qw = cdp_fma2(T.d[12], Px, cdp_fma2(T.d[13], Py, cdp_fma2(T.d[14], Pz, T.d[15])));
  const float2 Qz = cdp_add2(Pz, az);
  o.Qx = cdp_add2(Px, ax); o.Qy = cdp_add2(Py, ay);
  o.regular[0] = Qw.x > 0.f ? (Qz.x >= CDP_Z_MIN * Qw.x) : (Qw.x < 0.f && Qz.x <= CDP_Z_MIN * Qw.x);
  o.regular[1] = Qw.y > 0.f ? (Qz.y >= CDP_Z_MIN * Qw.y) : (Qw.y < 0.f && Qz.y <= CDP_Z_MIN * Qw.y);
  // (an irregular lane gets iz = 0: its packed displacement and adjoint are exactly zero instead of
  // inf / NaN, and the caller redoes that source with the literal formulas)
  o.iz.x = o.regular[0] ? cdp_rcp(Qz.x) : 0.f; o.iz.y = o.regular[1] ? cdp_rcp(Qz.y) : 0.f;
  o.dx = cdp_mul2(cdp_mul2(cdp_set2(k.fx), cdp_fma2(cdp_set2(-o.rx), az, ax)), o.iz);
  o.dy = cdp_mul2(cdp_mul2(cdp_set2(k.fy), cdp_fma2(cdp_set2(-o.ry), az, ay)), o.iz);
  o.ix = cdp_add2(cdp_set2(u), o.dx);
  o.iy = cdp_add2(cdp_set2(v), o.dy);
}

// Adjoint of the regular branch for both sources at once (lanes): gu, gv = dL/d(ix, iy) with the
// clip mask applied.  Adds to dT2[4 r + c] (rows 0..2; row 3 receives nothing when the depth
// clamp is inactive, Q_w drops out) and returns dL/d depth summed over the two sources' lanes as
// a pair (caller adds .x + .y); gQ[3] is handed back for dL/d motion.
CDP_HD float2 cdp_warp_adjoint2(float2 gu, float2 gv, const CdpWarp2& w, const CdpCam& k, const CdpPose2& T,
                                float2* dT2, float2* gQ) {
  gQ[0] = cdp_mul2(cdp_mul2(gu, cdp_set2(k.fx)), w.iz);
  gQ[1] = cdp_mul2(cdp_mul2(gv, cdp_set2(k.fy)), w.iz);
  gQ[2] = cdp_mul2(cdp_fma2(gQ[0], w.Qx, cdp_mul2(gQ[1], w.Qy)), cdp_mul2(w.iz, cdp_set2(-1.0f)));
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    dT2[4 * r + 0] = cdp_fma2(gQ[r], cdp_set2(w.P[0]), dT2[4 * r + 0]);
    dT2[4 * r + 1] = cdp_fma2(gQ[r], cdp_set2(w.P[1]), dT2[4 * r + 1]);
    dT2[4 * r + 2] = cdp_fma2(gQ[r], cdp_set2(w.P[2]), dT2[4 * r + 2]);
    dT2[4 * r + 3] = cdp_add2(dT2[4 * r + 3], gQ[r]);
  }
  // gP = T^T gQ with T = D + diag(1,1,1,0);  dP/d depth = (rx, ry, 1)
  const float2 gPx = cdp_fma2(T.d[0], gQ[0], cdp_fma2(T.d[4], gQ[1], cdp_fma2(T.d[8], gQ[2], gQ[0])));
  const float2 gPy = cdp_fma2(T.d[1], gQ[0], cdp_fma2(T.d[5], gQ[1], cdp_fma2(T.d[9], gQ[2], gQ[1])));
  const float2 gPz = cdp_fma2(T.d[2], gQ[0], cdp_fma2(T.d[6], gQ[1], cdp_fma2(T.d[10], gQ[2], gQ[2])));
  return cdp_fma2(gPx, cdp_set2(w.rx), cdp_fma2(gPy, cdp_set2(w.ry), gPz));
}

// ------------------------------------------------------------------------------------------
// F.grid_sample(mode="bilinear", padding_mode="border", align_corners=True) taps
// (misc/image_warper.py:178-182).
// ------------------------------------------------------------------------------------------
struct CdpTaps {
  int o00, o01, o10, o11;      // plane offsets of the nw, ne, sw, se taps
  float wx0, wx1, wy0, wy1;    // 1-D weights (x1-ix, ix-x0, y1-iy, iy-y0)
  float mx, my;                // 1 where the coordinate gradient passes the border clip
};

// one axis: integer pixel index p, displacement d (= i - p), absolute position i, extent n.
// Outside (0, n-1) the coordinate is clipped to the border (weight 1 on the border pixel, gradient
// mask 0: clip_coordinates_set_grad); NaN compares false and lands on border 0.
#ifndef CDP_OPT_BRANCHFREE_TAPS
#define CDP_OPT_BRANCHFREE_TAPS 1
#endif
CDP_HD void cdp_tap_axis(int p, float d, float i, int n, int& i0, int& i1, float& w0, float& w1, float& m) {
  const float nm1 = (float)(n - 1);
#if CDP_OPT_BRANCHFREE_TAPS
  const bool pos = i > 0.f, inside = pos && i < nm1;
  const float fl = floorf(d);
  const float frac = inside ? d - fl : 0.f;
  int a = inside ? p + (int)fl : (pos ? n - 1 : 0);
  a = a < 0 ? 0 : (a > n - 1 ? n - 1 : a);
  i0 = a;
  m = inside ? 1.f : 0.f;
#else
  float frac;
  if (!(i > 0.f)) { i0 = 0; frac = 0.f; m = 0.f; }
  else if (i >= nm1) { i0 = n - 1; frac = 0.f; m = 0.f; }
  else {
    const float fl = floorf(d);
    i0 = p + (int)fl; frac = d - fl; m = 1.f;
  }
  i0 = i0 < 0 ? 0 : (i0 > n - 1 ? n - 1 : i0);
#endif
  i1 = i0 + 1 > n - 1 ? n - 1 : i0 + 1;  // a tap one past the border has weight exactly 0
  w1 = frac; w0 = 1.0f - frac;
}

CDP_HD void cdp_taps(int u, int v, const CdpWarp& w, int W, int H, CdpTaps& t) {
  int x0, x1, y0, y1;
  cdp_tap_axis(u, w.dx, w.ix, W, x0, x1, t.wx0, t.wx1, t.mx);
  cdp_tap_axis(v, w.dy, w.iy, H, y0, y1, t.wy0, t.wy1, t.my);
  t.o00 = y0 * W + x0; t.o01 = y0 * W + x1; t.o10 = y1 * W + x0; t.o11 = y1 * W + x1;
}

// The same sampler with an always-full 2x2 footprint: lower tap index kept in [0, n-2], upper tap =
// lower + 1, so that the four taps sit at base + {0, 1, stride, stride + 1} (one address and three
// immediate offsets).  A coordinate clipped to the far border becomes (n-2, fraction 1), which is
// the same value as (n-1, fraction 0).
CDP_HD void cdp_tap_axis_full(int p, float d, float i, int n, int& i0, float& frac, float& m) {
  const bool pos = i > 0.f, inside = pos && i < (float)(n - 1);  // NaN: both false -> border 0
  const float fl = floorf(d);
  int a = inside ? p + (int)fl : (pos ? n - 2 : 0);
  a = a < 0 ? 0 : (a > n - 2 ? n - 2 : a);
  i0 = a;
  frac = inside ? d - fl : (pos ? 1.f : 0.f);
  m = inside ? 1.f : 0.f;
}

CDP_HD float cdp_bilinear(const float* plane, const CdpTaps& t) {
  const float nw = CDP_LDG(plane + t.o00), ne = CDP_LDG(plane + t.o01);
  const float sw = CDP_LDG(plane + t.o10), se = CDP_LDG(plane + t.o11);
  const float top = nw + t.wx1 * (ne - nw), bot = sw + t.wx1 * (se - sw);
  return top + t.wy1 * (bot - top);
}

// value and d(sample)/d(ix), d(sample)/d(iy) for one channel
CDP_HD float cdp_bilinear_grad(const float* plane, const CdpTaps& t, float& dix, float& diy) {
  const float nw = CDP_LDG(plane + t.o00), ne = CDP_LDG(plane + t.o01);
  const float sw = CDP_LDG(plane + t.o10), se = CDP_LDG(plane + t.o11);
  dix = (ne - nw) * t.wy0 + (se - sw) * t.wy1;
  diy = (sw - nw) * t.wx0 + (se - ne) * t.wx1;
  const float top = nw + t.wx1 * (ne - nw), bot = sw + t.wx1 * (se - sw);
  return top + t.wy1 * (bot - top);
}

CDP_HD int cdp_nearest_index(float ix, float iy, int W, int H) {
  const float cx = fminf(fmaxf(ix, 0.f), (float)(W - 1));
  const float cy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
  int x = (int)nearbyintf(cx), y = (int)nearbyintf(cy);  // ATen: std::nearbyint
  x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
  y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
  return y * W + x;
}

// ------------------------------------------------------------------------------------------
// Chain dL/d(ix,iy) back to depth, pose (and motion).  gu, gv already include the clip mask.
// T here is the TRUE pose matrix minus diag(1,1,1,0) (CdpPose).
// ------------------------------------------------------------------------------------------
CDP_HD void cdp_warp_adjoint(float gu, float gv, const CdpWarp& w, const CdpCam& k, const CdpPose& T,
                             float& g_depth, float* dT, float* g_motion3) {
  float gQ[4];
  const float ax = gu * k.fx, ay = gv * k.fy;
  if (w.regular) {  // u' = fx Q_x / Q_z + cx: Q_w drops out
    gQ[0] = ax * w.iz;
    gQ[1] = ay * w.iz;
    gQ[2] = -(gQ[0] * w.Q[0] + gQ[1] * w.Q[1]) * w.iz;
    gQ[3] = 0.f;
  } else {  // E = Q/Q_w, zt = max(E_z, eps) = eps or E_z
    const float iw = 1.0f / w.Q[3];
    const float ex = w.Q[0] * iw, ey = w.Q[1] * iw, ez = w.Q[2] * iw;
    const float gEx = ax * w.iz, gEy = ay * w.iz;
    const float gzt = -(ax * ex + ay * ey) * w.iz * w.iz;
    const float gEz = (ez >= CDP_Z_MIN) ? gzt : 0.f;  // clamp(min=) passes gradient on >=
    gQ[0] = gEx * iw; gQ[1] = gEy * iw; gQ[2] = gEz * iw;
    gQ[3] = -(gEx * w.Q[0] + gEy * w.Q[1] + gEz * w.Q[2]) * iw * iw;
  }
  if (g_motion3) { g_motion3[0] = gQ[0]; g_motion3[1] = gQ[1]; g_motion3[2] = gQ[2]; }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    dT[4 * r + 0] += gQ[r] * w.P[0];
    dT[4 * r + 1] += gQ[r] * w.P[1];
    dT[4 * r + 2] += gQ[r] * w.P[2];
    dT[4 * r + 3] += gQ[r];
  }
  // gP = T^T gQ with T = D + diag(1,1,1,0);  dP/d depth = (rx, ry, 1)
  const float gPx = gQ[0] + T.d[0] * gQ[0] + T.d[4] * gQ[1] + T.d[8] * gQ[2] + T.d[12] * gQ[3];
  const float gPy = gQ[1] + T.d[1] * gQ[0] + T.d[5] * gQ[1] + T.d[9] * gQ[2] + T.d[13] * gQ[3];
  const float gPz = gQ[2] + T.d[2] * gQ[0] + T.d[6] * gQ[1] + T.d[10] * gQ[2] + T.d[14] * gQ[3];
  g_depth += gPx * w.rx + gPy * w.ry + gPz;
}

// ------------------------------------------------------------------------------------------
// SSIMLoss (algos/depth.py:141-153) from window statistics of values CENTRED on a constant c
// (variances and the covariance are shift invariant; only the means need c added back).
// Centring removes most of the cancellation in E[x^2] - mean^2 that costs the fp32 reference
// about three digits in smooth image regions.
// ------------------------------------------------------------------------------------------
struct CdpSsimTerms {
  float mx, my;  // true means
  float n1, n2, id1, id2, S;
  float loss;  // clamp((1 - S) / 2, 0, 1)
  float dl;    // d loss / d S: -0.5 inside the clamp, else 0
};

// mxc, myc: means of the centred values; exx, eyy, exy: second moments of the centred values
CDP_HD void cdp_ssim_terms(float mxc, float myc, float exx, float eyy, float exy, float c, CdpSsimTerms& o) {
  o.mx = mxc + c; o.my = myc + c;
  const float vx = exx - mxc * mxc, vy = eyy - myc * myc, cov = exy - mxc * myc;
  o.n1 = 2.0f * o.mx * o.my + CDP_SSIM_C1;
  o.n2 = 2.0f * cov + CDP_SSIM_C2;
  const float d1 = o.mx * o.mx + o.my * o.my + CDP_SSIM_C1;
  const float d2 = vx + vy + CDP_SSIM_C2;
  o.id1 = cdp_rcp(d1);
  o.id2 = cdp_rcp(d2);
  o.S = (o.n1 * o.id1) * (o.n2 * o.id2);
  const float raw = (1.0f - o.S) * 0.5f;
  o.loss = fminf(fmaxf(raw, 0.f), 1.f);
  o.dl = (raw >= 0.f && raw <= 1.f) ? -0.5f : 0.f;  // clamp gradient is inclusive
}

// SSIM adjoint coefficients.  With A = dl/dmean_x, B = dl/dE[x^2], C = dl/dE[xy] the gradient of
// the loss at window centre q w.r.t. a window pixel p is
//     d loss(q) / d x(p) = m(p,q)/9 * (A + 2 x(p) B + y(p) C).
// A, 2xB and yC are each O(1/d2) and cancel almost completely.  Substituting
// A = A1 - 2 mean_x B - mean_y C with A1 = dl * (2 mean_y n2 / (d1 d2) - 2 mean_x S / d1) gives
// the centred, well-conditioned form used by the kernels:
//     d loss(q) / d x(p) = m/9 * (A0 + 2 (x(p) - x(q)) B + (y(p) - y(q)) C),
//     A0 = A1 + 2 (x(q) - mean_x) B + (y(q) - mean_y) C            (gradient at the centre itself).
// dxq = x(q) - mean_x, dyq = y(q) - mean_y.
CDP_HD void cdp_ssim_coeffs(const CdpSsimTerms& t, float dxq, float dyq, float& A0, float& B, float& C) {
  B = t.dl * (-t.S * t.id2);
  C = t.dl * (2.f * t.n1 * t.id1 * t.id2);
  const float A1 = t.dl * 2.f * t.id1 * (t.my * t.n2 * t.id2 - t.mx * t.S);
  A0 = A1 + 2.f * dxq * B + dyq * C;
}

// The same coefficients in un-centred form, A + 2 x(p) B + y(p) C, for values measured in a frame
// where the window means are (mx_f, my_f).  Used with tile-centred values (|x| <~ 1), where the
// cancellation between the three terms costs ~1.5 digits instead of ~3.5.
CDP_HD void cdp_ssim_coeffs_abc(const CdpSsimTerms& t, float mx_f, float my_f, float& A, float& B, float& C) {
  B = t.dl * (-t.S * t.id2);
  C = t.dl * (2.f * t.n1 * t.id1 * t.id2);
  const float A1 = t.dl * 2.f * t.id1 * (t.my * t.n2 * t.id2 - t.mx * t.S);
  A = A1 - 2.f * mx_f * B - my_f * C;
}

// Reflection padding of one pixel (nn.ReflectionPad2d(1), algos/depth.py:123): -1 -> 1, n -> n-2.
CDP_HD int cdp_reflect(int i, int n) { return i < 0 ? -i : (i > n - 1 ? 2 * (n - 1) - i : i); }

// How often the 3x3 window of neighbour q = p + d touches pixel p once the reflected border is
// folded back (1-D factor); 0 when q lies outside the image.
CDP_HD float cdp_reflect_mult(int p, int d, int n) {
  const int q = p + d;
  if (q < 0 || q > n - 1) return 0.f;
  if (d == -1 && p == 1) return 2.f;      // q = 0 also reaches p through its mirrored -1
  if (d == 1 && p == n - 2) return 2.f;   // q = n-1 also reaches p through its mirrored n
  return 1.f;
}

// ------------------------------------------------------------------------------------------
// Counter-based tie-break noise (used when the caller passes no noise tensors, i.e.
// ReconstructionLoss(noise="fused")): the identity candidates get 1e-5 * n, n ~ N(0, 1)
// (algos/depth.py:316-318).  Per pixel one 64-bit draw -- two rounds of a 32-bit multiply-xorshift
// mix of (pixel, level, sample, seed) per word -- and ONE Box-Muller transform, whose two outputs
// are exactly the two normals a pixel needs (one per identity candidate): u1 in (0, 1) with 24 bits,
// r = sqrt(-2 ln u1) <= 5.9, angle = 2 pi u2.  On the GPU the transcendental steps are the
// hardware approximations (lg2 / sqrt / sin / cos, ~1e-6 absolute): ~30 instructions per pixel,
// no noise tensors in HBM and no generator launches.  cdp_tiebreak_noise (codeps_photo.h) writes
// the same draws into a tensor so that tests can hand them to the oracle.  Not torch's Philox
// stream: a loss built with noise="torch" draws torch.randn per level like the reference.
// ------------------------------------------------------------------------------------------
CDP_HD uint32_t cdp_mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu;
  h ^= h >> 13; h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
CDP_HD void cdp_box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-08f;  // (k + 1/2) 2^-24 in (0, 1)
  const float turn = (float)(b >> 8) * 5.9604644775390625e-08f;         // [0, 1)
#if defined(__CUDA_ARCH__)
  float r, sn, cs;
  const float x = -1.3862943611198906f * __log2f(u1);  // -2 ln u1 = -2 ln 2 * log2 u1 >= 6e-8
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  const float ang = 6.283185307179586f * turn;
  sn = __sinf(ang); cs = __cosf(ang);
#else
  const float r = sqrtf(-2.0f * logf(u1));
  const float sn = sinf(6.283185307179586f * turn), cs = cosf(6.283185307179586f * turn);
#endif
  n0 = r * cs; n1 = r * sn;
}
CDP_HD void cdp_noise_pair(uint64_t seed, uint32_t pixel, uint32_t level, uint32_t sample,
                           float& n0, float& n1) {
  const uint32_t key = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u) ^ (level * 0x632BE5ABu) ^ (sample * 0x7F4A7C15u);
  const uint32_t a = cdp_mix32(pixel * 0x9E3779B1u + key);
  cdp_box_muller(a, cdp_mix32(a ^ 0x68E31DA4u), n0, n1);
}
