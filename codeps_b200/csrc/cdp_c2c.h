// cdp_c2c.h -- camera-to-camera warp at constant depth: Mixup.warp_c2c
// (/root/reference/datasets/mixup.py:211-229) = _ImageToPointcloud on the TARGET camera's pixel
// grid (/root/reference/misc/image_warper.py:54-87, rays in fp32), Mixup._src_pcl_to_tgt on the
// SOURCE camera (mixup.py:29-66, fp64) and F.grid_sample(..., align_corners=True) in fp64 with
// mode bilinear | nearest and padding zeros | border.  The output resolution is the target's,
// the sampled image has its own resolution.  No gradient (data augmentation).
//
// The op order and the fp32 / fp64 split of the reference are kept, so that nearest-mode label
// warps pick the same pixel.
#pragma once
#include "cdp_common.h"

#if defined(__CUDA_ARCH__)
#define CDP_DMUL(a, b) __dmul_rn((a), (b))
#define CDP_DADD(a, b) __dadd_rn((a), (b))
#else
#define CDP_DMUL(a, b) ((a) * (b))
#define CDP_DADD(a, b) ((a) + (b))
#endif

struct CdpC2cParams {
  const void* src;   // [B,C,Hs,Ws] float or double
  double* out;       // [B,C,Ht,Wt]
  double Ks[CDP_MAX_BATCH_PER_LAUNCH][4];  // source camera fx, fy, cx, cy (projection, fp64)
  float Kt[CDP_MAX_BATCH_PER_LAUNCH][4];   // target camera (viewing rays, fp32 like the reference's grid)
  double depth;
  int32_t batch_begin, C, Hs, Ws, Ht, Wt;
  int32_t nearest, zeros;
};

// source-image coordinates (un-normalised, before any clipping) of target pixel (u, v)
CDP_HD void cdp_c2c_coords(const CdpC2cParams& p, int b_local, int u, int v, double& ix, double& iy) {
  const float* kt = p.Kt[b_local];
  // CameraModel.get_viewing_ray on the fp32 pixel grid (misc/camera_model.py:52-71)
  const float rx = CDP_SUB((float)u, kt[2]) / kt[0];
  const float ry = CDP_SUB((float)v, kt[3]) / kt[1];
  const float norm = sqrtf(CDP_ADD(CDP_ADD(CDP_MUL(rx, rx), CDP_MUL(ry, ry)), 1.0f));
  const float ux = rx / norm, uy = ry / norm, uz = 1.0f / norm;
  // _ImageToPointcloud.forward in the dtype of the depth map (fp64): depth / |rz| * r
  const double scale = p.depth / fabs((double)uz);
  const double x3 = CDP_DMUL(scale, (double)ux), y3 = CDP_DMUL(scale, (double)uy);
  double z3 = CDP_DMUL(scale, (double)uz);
  z3 = z3 < 1e-5 ? 1e-5 : z3;  // clamp(min=1e-5), mixup.py:49
  // CameraModel.get_image_point of the source camera (misc/camera_model.py:43-50)
  const double* ks = p.Ks[b_local];
  const double us = CDP_DADD(CDP_DMUL(x3 / z3, ks[0]), ks[2]);
  const double vs = CDP_DADD(CDP_DMUL(y3 / z3, ks[1]), ks[3]);
  // normalise to [-1, 1] (mixup.py:59-60), then grid_sample's un-normalisation (align_corners)
  const double gx = CDP_DMUL(us / (double)(p.Ws - 1) - 0.5, 2.0);
  const double gy = CDP_DMUL(vs / (double)(p.Hs - 1) - 0.5, 2.0);
  ix = CDP_DMUL(CDP_DADD(gx, 1.0) / 2.0, (double)(p.Ws - 1));
  iy = CDP_DMUL(CDP_DADD(gy, 1.0) / 2.0, (double)(p.Hs - 1));
}

template <typename T>
CDP_HD void cdp_c2c_pixel(const CdpC2cParams& p, int b_local, int pix) {
  const int v = pix / p.Wt, u = pix - v * p.Wt;
  const int b = p.batch_begin + b_local;
  double ix, iy;
  cdp_c2c_coords(p, b_local, u, v, ix, iy);
  const int Ws = p.Ws, Hs = p.Hs;
  if (!p.zeros) {  // padding_mode="border": clip the coordinate (NaN -> 0 like ATen's clip_coordinates)
    ix = fmin((double)(Ws - 1), fmax(ix, 0.0));
    iy = fmin((double)(Hs - 1), fmax(iy, 0.0));
  }
  const size_t splane = (size_t)Hs * Ws, tplane = (size_t)p.Ht * p.Wt;
  const T* src = static_cast<const T*>(p.src) + (size_t)b * p.C * splane;
  double* out = p.out + (size_t)b * p.C * tplane + pix;
  if (p.nearest) {
    const double nx = nearbyint(ix), ny = nearbyint(iy);  // round half to even, as ATen
    const bool inside = nx >= 0.0 && nx <= (double)(Ws - 1) && ny >= 0.0 && ny <= (double)(Hs - 1);
    const size_t o = inside ? (size_t)ny * Ws + (size_t)nx : 0;
    for (int c = 0; c < p.C; ++c) out[c * tplane] = inside ? (double)src[c * splane + o] : 0.0;
    return;
  }
  const double fx = floor(ix), fy = floor(iy);
  const double ex = fx + 1.0 - ix, ey = fy + 1.0 - iy, tx = ix - fx, ty = iy - fy;  // ATen: (ix_se - ix) etc.
  const double w_nw = ex * ey, w_ne = tx * ey, w_sw = ex * ty, w_se = tx * ty;
  // taps outside the image are dropped (zeros padding; with border padding only the +1 taps at
  // the far edge can fall outside, with weight 0)
  const bool x0 = fx >= 0.0 && fx <= (double)(Ws - 1), x1 = fx + 1.0 >= 0.0 && fx + 1.0 <= (double)(Ws - 1);
  const bool y0 = fy >= 0.0 && fy <= (double)(Hs - 1), y1 = fy + 1.0 >= 0.0 && fy + 1.0 <= (double)(Hs - 1);
  const long long xi = x0 || x1 ? (long long)fx : 0, yi = y0 || y1 ? (long long)fy : 0;
  for (int c = 0; c < p.C; ++c) {
    const T* s = src + c * splane;
    double acc = 0.0;
    if (y0 && x0) acc += (double)s[yi * Ws + xi] * w_nw;
    if (y0 && x1) acc += (double)s[yi * Ws + xi + 1] * w_ne;
    if (y1 && x0) acc += (double)s[(yi + 1) * Ws + xi] * w_sw;
    if (y1 && x1) acc += (double)s[(yi + 1) * Ws + xi + 1] * w_se;
    out[c * tplane] = acc;
  }
}

static inline bool cdp_fill_c2c_params(CdpC2cParams* p, const void* src, double* out, const double* Ks,
                                       const double* Kt, int b0, int nb, int C, int Hs, int Ws, int Ht, int Wt,
                                       double depth, int nearest, int zeros) {
  memset(p, 0, sizeof(*p));
  p->src = src; p->out = out; p->depth = depth;
  p->batch_begin = b0; p->C = C; p->Hs = Hs; p->Ws = Ws; p->Ht = Ht; p->Wt = Wt;
  p->nearest = nearest; p->zeros = zeros;
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < 4; ++j) {
      p->Ks[i][j] = Ks[(size_t)(b0 + i) * 4 + j];
      p->Kt[i][j] = (float)Kt[(size_t)(b0 + i) * 4 + j];
    }
  return true;
}
