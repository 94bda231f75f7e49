// cdp_math.h -- per-pixel math of the photometric loss (host/device).
//
// Forward follows the reference's operation order (ray -> normalise -> scale by depth -> T ->
// /w -> clamp -> project -> normalise -> un-normalise -> clip -> bilinear), SURVEY.md section 3.3;
// backward is the closed form of SURVEY.md section 7 / 8a.
#pragma once

#include "cdp_common.h"

// ------------------------------------------------------------------------------------------
// Back-projection: CameraModel.get_viewing_ray (misc/camera_model.py:52-71) followed by
// _ImageToPointcloud.forward (misc/image_warper.py:83-85).
// ------------------------------------------------------------------------------------------
struct CdpPoint {
  float P[3];     // 3-D point
  float dPdD[3];  // dP / d depth  (= unit ray / |ray_z|)
};

CDP_HD void cdp_backproject(float u, float v, float depth, const CdpCam& k, CdpPoint& o) {
  const float rx = (u - k.cx) / k.fx;
  const float ry = (v - k.cy) / k.fy;
  const float nrm = sqrtf(CDP_ADD(CDP_ADD(CDP_MUL(rx, rx), CDP_MUL(ry, ry)), 1.0f));
  const float hx = rx / nrm, hy = ry / nrm, hz = 1.0f / nrm;
  const float az = fabsf(hz);
  const float a = depth / az;
  o.P[0] = CDP_MUL(a, hx);
  o.P[1] = CDP_MUL(a, hy);
  o.P[2] = CDP_MUL(a, hz);
  o.dPdD[0] = hx / az;
  o.dPdD[1] = hy / az;
  o.dPdD[2] = hz / az;
}

// ------------------------------------------------------------------------------------------
// CoordinateWarper.forward after the point cloud (misc/image_warper.py:125-144) and
// _PointcloudToImage / CameraModel.get_image_point (misc/image_warper.py:29-45,
// misc/camera_model.py:43-50), then grid_sample's un-normalisation (align_corners=True).
// ------------------------------------------------------------------------------------------
struct CdpProj {
  float Q[4];    // T * [P;1] (+ motion on xyz)
  float E[3];    // Q_xyz / Q_w
  float zt;      // max(E_z, 1e-5)
  float gx, gy;  // normalised grid coordinates (what CoordinateWarper returns)
  float ix, iy;  // un-normalised, un-clipped sample position in source pixels
};

CDP_HD void cdp_project(const float P[3], const float* T, const float* motion3, const CdpCam& k,
                        float wm1, float hm1, CdpProj& o) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
    o.Q[r] = T[4 * r + 0] * P[0] + T[4 * r + 1] * P[1] + T[4 * r + 2] * P[2] + T[4 * r + 3];
  if (motion3) {
    o.Q[0] += motion3[0]; o.Q[1] += motion3[1]; o.Q[2] += motion3[2];
  }
  o.E[0] = o.Q[0] / o.Q[3];
  o.E[1] = o.Q[1] / o.Q[3];
  o.E[2] = o.Q[2] / o.Q[3];
  o.zt = fmaxf(o.E[2], CDP_Z_MIN);
  const float u2 = CDP_ADD(CDP_MUL(o.E[0] / o.zt, k.fx), k.cx);
  const float v2 = CDP_ADD(CDP_MUL(o.E[1] / o.zt, k.fy), k.cy);
  o.gx = CDP_MUL(CDP_SUB(u2 / wm1, 0.5f), 2.0f);
  o.gy = CDP_MUL(CDP_SUB(v2 / hm1, 0.5f), 2.0f);
  o.ix = CDP_MUL(CDP_MUL(CDP_ADD(o.gx, 1.0f), 0.5f), wm1);
  o.iy = CDP_MUL(CDP_MUL(CDP_ADD(o.gy, 1.0f), 0.5f), hm1);
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

CDP_HD void cdp_taps(float ix, float iy, int W, int H, CdpTaps& t) {
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  t.mx = (ix > 0.f && ix < wm1) ? 1.f : 0.f;  // clip_coordinates_set_grad
  t.my = (iy > 0.f && iy < hm1) ? 1.f : 0.f;
  const float cx = fminf(fmaxf(ix, 0.f), wm1);
  const float cy = fminf(fmaxf(iy, 0.f), hm1);
  const float fx0 = floorf(cx), fy0 = floorf(cy);
  t.wx1 = cx - fx0; t.wx0 = (fx0 + 1.f) - cx;
  t.wy1 = cy - fy0; t.wy0 = (fy0 + 1.f) - cy;
  int x0 = (int)fx0, y0 = (int)fy0;
  // NaN / garbage coordinates must never index out of the plane
  x0 = x0 < 0 ? 0 : (x0 > W - 1 ? W - 1 : x0);
  y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
  // a tap one past the border has weight exactly 0: clamp its index instead of branching
  const int x1 = x0 + 1 > W - 1 ? W - 1 : x0 + 1;
  const int y1 = y0 + 1 > H - 1 ? H - 1 : y0 + 1;
  t.o00 = y0 * W + x0; t.o01 = y0 * W + x1; t.o10 = y1 * W + x0; t.o11 = y1 * W + x1;
}

CDP_HD float cdp_bilinear(const float* plane, const CdpTaps& t) {
  const float nw = CDP_LDG(plane + t.o00), ne = CDP_LDG(plane + t.o01);
  const float sw = CDP_LDG(plane + t.o10), se = CDP_LDG(plane + t.o11);
  return CDP_ADD(CDP_ADD(CDP_ADD(CDP_MUL(nw, CDP_MUL(t.wx0, t.wy0)), CDP_MUL(ne, CDP_MUL(t.wx1, t.wy0))),
                         CDP_MUL(sw, CDP_MUL(t.wx0, t.wy1))),
                 CDP_MUL(se, CDP_MUL(t.wx1, t.wy1)));
}

// d(sample)/d(ix), d(sample)/d(iy) for one channel (grid_sampler_2d_backward w.r.t. the grid).
CDP_HD void cdp_bilinear_grad(const float* plane, const CdpTaps& t, float& dix, float& diy) {
  const float nw = CDP_LDG(plane + t.o00), ne = CDP_LDG(plane + t.o01);
  const float sw = CDP_LDG(plane + t.o10), se = CDP_LDG(plane + t.o11);
  dix = (ne - nw) * t.wy0 + (se - sw) * t.wy1;
  diy = (sw - nw) * t.wx0 + (se - ne) * t.wx1;
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
// Chain dL/d(ix,iy) back to depth, pose (and motion): adjoint of cdp_project / cdp_backproject.
// gu, gv already include the border-clip mask.
// ------------------------------------------------------------------------------------------
CDP_HD void cdp_warp_adjoint(float gu, float gv, const CdpProj& pr, const CdpPoint& pt,
                             const float* T, const CdpCam& k, float& g_depth, float* dT,
                             float* g_motion3) {
  const float izt = 1.0f / pr.zt;
  const float ax = gu * k.fx, ay = gv * k.fy;
  const float gEx = ax * izt, gEy = ay * izt;
  const float gzt = -(ax * pr.E[0] + ay * pr.E[1]) * izt * izt;
  const float gEz = (pr.E[2] >= CDP_Z_MIN) ? gzt : 0.f;  // clamp(min=) passes gradient on >=
  const float iw = 1.0f / pr.Q[3];
  float gQ[4];
  gQ[0] = gEx * iw; gQ[1] = gEy * iw; gQ[2] = gEz * iw;
  gQ[3] = -(gEx * pr.Q[0] + gEy * pr.Q[1] + gEz * pr.Q[2]) * iw * iw;
  if (g_motion3) { g_motion3[0] = gQ[0]; g_motion3[1] = gQ[1]; g_motion3[2] = gQ[2]; }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    dT[4 * r + 0] += gQ[r] * pt.P[0];
    dT[4 * r + 1] += gQ[r] * pt.P[1];
    dT[4 * r + 2] += gQ[r] * pt.P[2];
    dT[4 * r + 3] += gQ[r];
  }
  float gd = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float gP = T[c] * gQ[0] + T[4 + c] * gQ[1] + T[8 + c] * gQ[2] + T[12 + c] * gQ[3];
    gd += gP * pt.dPdD[c];
  }
  g_depth += gd;
}

// ------------------------------------------------------------------------------------------
// SSIMLoss (algos/depth.py:141-153) from the five window means.
// ------------------------------------------------------------------------------------------
struct CdpSsimTerms {
  float n1, n2, d1, d2, S;
  float loss;  // clamp((1 - S) / 2, 0, 1)
  float dl;    // d loss / d S: -0.5 inside the clamp, else 0
};

CDP_HD void cdp_ssim_terms(float mx, float my, float exx, float eyy, float exy, CdpSsimTerms& o) {
  const float mxy = CDP_MUL(mx, my), mxx = CDP_MUL(mx, mx), myy = CDP_MUL(my, my);
  const float vx = CDP_SUB(exx, mxx), vy = CDP_SUB(eyy, myy), cov = CDP_SUB(exy, mxy);
  o.n1 = CDP_ADD(CDP_MUL(2.0f, mxy), CDP_SSIM_C1);
  o.n2 = CDP_ADD(CDP_MUL(2.0f, cov), CDP_SSIM_C2);
  o.d1 = CDP_ADD(CDP_ADD(mxx, myy), CDP_SSIM_C1);
  o.d2 = CDP_ADD(CDP_ADD(vx, vy), CDP_SSIM_C2);
  o.S = CDP_MUL(o.n1, o.n2) / CDP_MUL(o.d1, o.d2);
  const float raw = CDP_MUL(CDP_SUB(1.0f, o.S), 0.5f);
  o.loss = fminf(fmaxf(raw, 0.f), 1.f);
  o.dl = (raw >= 0.f && raw <= 1.f) ? -0.5f : 0.f;  // clamp gradient is inclusive
}

// SSIM adjoint coefficients.  With A = dl/dmean_x, B = dl/dE[x^2], C = dl/dE[xy] the gradient of
// the loss at window centre q w.r.t. a window pixel p is
//     d loss(q) / d x(p) = m(p,q)/9 * (A + 2 x(p) B + y(p) C).
// A, 2xB and yC are each O(1/d2) and cancel almost completely, so in fp32 that form loses ~3
// digits.  Substituting A = A1 - 2 mean_x B - mean_y C with
//     A1 = dl * (2 mean_y n2 / (d1 d2) - 2 mean_x S / d1)          (no large terms)
// gives the centred, well-conditioned form used by the kernels:
//     d loss(q) / d x(p) = m/9 * (A0 + 2 (x(p) - x(q)) B + (y(p) - y(q)) C),
//     A0 = A1 + 2 (x(q) - mean_x) B + (y(q) - mean_y) C            (gradient at the centre itself).
CDP_HD void cdp_ssim_coeffs(float mx, float my, float xq, float yq, const CdpSsimTerms& t, float& A0,
                            float& B, float& C) {
  const float inv = 1.0f / (t.d1 * t.d2);
  B = t.dl * (-t.S / t.d2);
  C = t.dl * (2.f * t.n1 * inv);
  const float A1 = t.dl * (2.f * my * t.n2 * inv - 2.f * mx * t.S / t.d1);
  A0 = A1 + 2.f * (xq - mx) * B + (yq - my) * C;
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
// Counter-based tie-break noise (used only when the caller passes no noise tensors):
// Philox-4x32-10 keyed by the seed, counter = (pixel, channel/level, sample), Box-Muller.
// ------------------------------------------------------------------------------------------
CDP_HD void cdp_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                       uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

CDP_HD void cdp_noise_pair(uint64_t seed, uint32_t pixel, uint32_t level, uint32_t sample,
                           float& n0, float& n1) {
  uint32_t r[4];
  cdp_philox(pixel, level, sample, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float u0 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * logf(u0));
  const float ang = 6.28318530717958647692f * u1;
  n0 = rad * cosf(ang);
  n1 = rad * sinf(ang);
}
