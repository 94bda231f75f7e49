// cdp_flow.h -- the two regularisers of the object-motion (scene-flow) maps that accompany the
// reconstruction loss when make_sflow is on: FlowSmoothnessLoss and FlowSparsityLoss
// (/root/reference/algos/depth.py:15-52; called at algos/depth.py:483-485).
//
// Both losses are terminal scalars, so the forward kernel also writes the *unit* gradient
// d loss / d map; backward is a scaled copy (cdp_scale_fwd).  Reductions use per-block records and
// a fixed-order combine -- no atomics, run-to-run identical.
//
// Kernel bodies are CDP_HD so that the g++ emulator of tests/emu can run them thread by thread.
#pragma once
#include "cdp_common.h"

#define CDP_FLOW_THREADS 256
#define CDP_FLOW_ROWS 4          // smoothness: rows per block (one column per thread)
#define CDP_FLOW_PER_THREAD 16   // sparsity: elements per thread

struct CdpFlowParams {
  const float* map[CDP_MAX_FLOW_MAPS];  // each [planes, H, W]  (planes = B * C)
  float* grad;      // [n_maps, planes, H, W] unit gradients, or null
  float* part;      // per-block records
  float* loss;
  int32_t n_maps, planes, H, W, wrap;
  int32_t blocks_x, blocks_y;  // smoothness: patches per plane; sparsity: chunks per plane (blocks_y = 1)
  float inv_count;             // 1 / (terms per map * n_maps)
};

// --------------------------------------------------------------------------------------------
// FlowSmoothnessLoss: mean sqrt((f - roll_x f)^2 + (f - roll_y f)^2 + 1e-7)   (depth.py:20-27)
// --------------------------------------------------------------------------------------------
CDP_HD float cdp_flow_edge(const float* f, int W, int H, int x, int y, float& gx, float& gy) {
  const int xm = x == 0 ? W - 1 : x - 1, ym = y == 0 ? H - 1 : y - 1;  // torch.roll(shifts=1)
  const float c = CDP_LDG(f + (size_t)y * W + x);
  gx = c - CDP_LDG(f + (size_t)y * W + xm);
  gy = c - CDP_LDG(f + (size_t)ym * W + x);
  return sqrtf(CDP_ADD(CDP_ADD(CDP_MUL(gx, gx), CDP_MUL(gy, gy)), 1e-7f));
}

// one thread: column x of a CDP_FLOW_ROWS-row patch; returns its share of sum(r)
CDP_HD float cdp_flow_smooth_thread(const CdpFlowParams& p, int bx, int by, int bz, int tid) {
  const int m = bz / p.planes, plane = bz - m * p.planes;
  const int W = p.W, H = p.H;
  const int x = bx * CDP_FLOW_THREADS + tid;
  if (x >= W) return 0.f;
  const float* f = p.map[m] + (size_t)plane * W * H;
  float* g = p.grad ? p.grad + ((size_t)m * p.planes + plane) * W * H : nullptr;
  const bool wrap = p.wrap != 0;
  const int xp = x + 1 == W ? 0 : x + 1;
  const bool has_xp = wrap || x + 1 < W;
  float acc = 0.f;
  for (int r = 0; r < CDP_FLOW_ROWS; ++r) {
    const int y = by * CDP_FLOW_ROWS + r;
    if (y >= H) break;
    float gx, gy, grad = 0.f;
    if (wrap || (x >= 1 && y >= 1)) {  // without wrap-around the first row / column is cropped
      const float rr = cdp_flow_edge(f, W, H, x, y, gx, gy);
      acc += rr;
      grad = (gx + gy) / rr;
    }
    if (g) {
      // f(x, y) is the subtrahend of the x-difference at (x+1, y) and of the y-difference at (x, y+1)
      if (has_xp && (wrap || y >= 1)) {
        const float rr = cdp_flow_edge(f, W, H, xp, y, gx, gy);
        grad -= gx / rr;
      }
      const int yp = y + 1 == H ? 0 : y + 1;
      if ((wrap || y + 1 < H) && (wrap || x >= 1)) {
        const float rr = cdp_flow_edge(f, W, H, x, yp, gx, gy);
        grad -= gy / rr;
      }
      g[(size_t)y * W + x] = grad * p.inv_count;
    }
  }
  return acc;
}

// --------------------------------------------------------------------------------------------
// FlowSparsityLoss: mean 2 m sqrt(|f| / (m + 1e-7) + 1),  m = mean_{H,W} |f| per (b, c), detached
// (depth.py:37-43)
// --------------------------------------------------------------------------------------------
// pass 1: sum |f| over this thread's elements of chunk bx of plane (bz)
CDP_HD float cdp_flow_abs_thread(const CdpFlowParams& p, int bx, int bz, int tid) {
  const int m = bz / p.planes, plane = bz - m * p.planes;
  const size_t n = (size_t)p.W * p.H;
  const float* f = p.map[m] + (size_t)plane * n;
  const size_t base = (size_t)bx * CDP_FLOW_THREADS * CDP_FLOW_PER_THREAD;
  float acc = 0.f;
  for (int k = 0; k < CDP_FLOW_PER_THREAD; ++k) {
    const size_t i = base + (size_t)k * CDP_FLOW_THREADS + tid;  // coalesced
    if (i < n) acc += fabsf(CDP_LDG(f + i));
  }
  return acc;
}

// spatial mean of |f| from the pass-1 records of one plane (fixed order; lanes as in cdp_lane_sum)
CDP_HD float cdp_flow_plane_mean(const CdpFlowParams& p, int bz, const double lane_sums[32]) {
  double v[32];
  for (int l = 0; l < 32; ++l) v[l] = lane_sums[l];
  for (int off = 16; off > 0; off >>= 1)
    for (int l = 0; l < off; ++l) v[l] += v[l + off];
  (void)bz;
  return (float)(v[0] / (double)((size_t)p.W * p.H));
}

// pass 2: loss share and unit gradient of this thread's elements
CDP_HD float cdp_flow_sparsity_thread(const CdpFlowParams& p, int bx, int bz, int tid, float mean) {
  const int m = bz / p.planes, plane = bz - m * p.planes;
  const size_t n = (size_t)p.W * p.H;
  const float* f = p.map[m] + (size_t)plane * n;
  float* g = p.grad ? p.grad + ((size_t)m * p.planes + plane) * n : nullptr;
  const size_t base = (size_t)bx * CDP_FLOW_THREADS * CDP_FLOW_PER_THREAD;
  const float den = mean + 1e-7f, two_m = 2.f * mean;
  const float gscale = mean / den * p.inv_count;  // d/df [2 m sqrt(|f|/den + 1)] = m/den * sign(f) / sqrt(.)
  float acc = 0.f;
  for (int k = 0; k < CDP_FLOW_PER_THREAD; ++k) {
    const size_t i = base + (size_t)k * CDP_FLOW_THREADS + tid;
    if (i >= n) break;
    const float v = CDP_LDG(f + i);
    const float s = sqrtf(fabsf(v) / den + 1.f);
    acc += two_m * s;
    if (g) g[i] = (v > 0.f ? gscale : (v < 0.f ? -gscale : 0.f)) / s;
  }
  return acc;
}

// --------------------------------------------------------------------------------------------
// fixed-order sum of `count` records -> loss (host / single-thread form; the device kernel uses
// the same lane-strided order: thread t sums records t, t+1024, ..., butterfly, warps in order)
// --------------------------------------------------------------------------------------------
CDP_HD double cdp_flow_sum_records_host(const float* part, size_t count) {
  double total = 0.0;
  for (int warp = 0; warp < 32; ++warp) {
    double v[32];
    for (int l = 0; l < 32; ++l) {
      double acc = 0.0;
      for (size_t i = (size_t)warp * 32 + l; i < count; i += 1024) acc += (double)part[i];
      v[l] = acc;
    }
    for (int off = 16; off > 0; off >>= 1)
      for (int l = 0; l < off; ++l) v[l] += v[l + off];
    total += v[0];
  }
  return total;
}

static inline bool cdp_fill_flow_params(const float* const* maps, int32_t n_maps, int32_t planes, int32_t H,
                                        int32_t W, int32_t wrap, bool sparsity, float* loss, float* grad,
                                        float* part, CdpFlowParams* p) {
  if (n_maps < 1 || n_maps > CDP_MAX_FLOW_MAPS || planes < 1 || H < 1 || W < 1) return false;
  if (!sparsity && !wrap && (H < 2 || W < 2)) return false;
  for (int i = 0; i < CDP_MAX_FLOW_MAPS; ++i) p->map[i] = i < n_maps ? maps[i] : nullptr;
  p->grad = grad; p->part = part; p->loss = loss;
  p->n_maps = n_maps; p->planes = planes; p->H = H; p->W = W; p->wrap = wrap;
  double terms;
  if (sparsity) {
    const size_t n = (size_t)H * W, per = (size_t)CDP_FLOW_THREADS * CDP_FLOW_PER_THREAD;
    p->blocks_x = (int32_t)((n + per - 1) / per);
    p->blocks_y = 1;
    terms = (double)planes * (double)n;
  } else {
    p->blocks_x = (W + CDP_FLOW_THREADS - 1) / CDP_FLOW_THREADS;
    p->blocks_y = (H + CDP_FLOW_ROWS - 1) / CDP_FLOW_ROWS;
    terms = (double)planes * (wrap ? (double)H * W : (double)(H - 1) * (W - 1));
  }
  p->inv_count = (float)(1.0 / (terms * n_maps));
  return true;
}

// records: smoothness one per block; sparsity two arrays (pass 1, pass 2) of one per block
static inline size_t cdp_flow_records(int32_t n_maps, int32_t planes, int32_t H, int32_t W, bool sparsity) {
  CdpFlowParams p;
  const float* dummy[CDP_MAX_FLOW_MAPS] = {nullptr, nullptr, nullptr, nullptr};
  if (!cdp_fill_flow_params(dummy, n_maps, planes, H, W, 1, sparsity, nullptr, nullptr, nullptr, &p)) return 0;
  const size_t blocks = (size_t)p.blocks_x * p.blocks_y * planes * n_maps;
  return sparsity ? 2 * blocks : blocks;
}
