// cdp_metrics.h -- DepthEvaluator.compute_depth_metrics (/root/reference/eval/depth.py:21-70,
// stats at :109-133): per image, over the pixels with ground truth (> 0, optional Garg crop):
// median scaling of the prediction (gt.median() / pred.median(), lower median as torch.median),
// clamp to the depth range, the three threshold accuracies, RMSE, RMSE(log), abs-rel, sq-rel;
// then the mean over the batch.  The reference loops over the images in python, compacts with a
// boolean mask and sorts twice per image; here the medians come from a 4-pass radix select on the
// order-preserving integer image of the floats (exact, no sort, no compaction) and the statistics
// from one more pass, all images in every launch, no host synchronisation.
//
// A "unit" is one image (compute_depth_metrics) or the whole batch restricted to one semantic
// class (compute_depth_metrics_per_class, eval/depth.py:72-106): units = B, n = H*W, or units = 1,
// n = B*H*W with `labels` / `class_id` set.
//
// Integer histogram counters are accumulated with atomics (exact, order-independent); floating
// point sums go through per-block records and a fixed-order fp64 combine.
#pragma once
#include "cdp_common.h"

#define CDP_METRICS_THREADS 256
#define CDP_METRICS_PER_THREAD 16
#define CDP_METRICS_CHUNK (CDP_METRICS_THREADS * CDP_METRICS_PER_THREAD)
#define CDP_METRICS_PASSES 4       // 8-bit digits, most significant first
#define CDP_METRICS_NSTATS 7       // a1, a2, a3, rmse, rmse_log, abs_rel, sq_rel
#define CDP_METRICS_REC 8          // per-block record: 7 sums + padding

struct CdpMetricsParams {
  const float* gt;        // [units, n]
  const float* pred;      // [units, n]
  const int64_t* labels;  // [units, n] or null
  int64_t class_id;
  uint32_t* hist;         // [CDP_METRICS_PASSES][units][2][256]
  float* part;            // [units][blocks][CDP_METRICS_REC]
  float* out;             // [CDP_METRICS_NSTATS] (+ [7] = number of units with ground truth)
  int32_t units, n, blocks;
  int32_t W, H;                  // image size when the crop box applies (n = H*W), else 0
  int32_t x0, x1, y0, y1;        // crop box [x0,x1) x [y0,y1)
  float lo, hi;                  // depth range
  int32_t use_gt_scale;
};

// order-preserving map float -> uint32 (negative values below positive ones)
CDP_HD uint32_t cdp_metrics_key(float v) {
  uint32_t b;
#if defined(__CUDA_ARCH__)
  b = __float_as_uint(v);
#else
  memcpy(&b, &v, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
CDP_HD float cdp_metrics_unkey(uint32_t k) {
  const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  float v;
#if defined(__CUDA_ARCH__)
  v = __uint_as_float(b);
#else
  memcpy(&v, &b, 4);
#endif
  return v;
}

// mask of eval/depth.py:34-43 (gt > 0, Garg crop) and :84-89 (class)
CDP_HD bool cdp_metrics_valid(const CdpMetricsParams& p, int unit, int i, float& g) {
  const size_t o = (size_t)unit * p.n + i;
  g = CDP_LDG(p.gt + o);
  if (!(g > 0.f)) return false;
  if (p.labels && CDP_LDG(p.labels + o) != p.class_id) return false;
  if (p.W > 0) {
    const int y = i / p.W, x = i - y * p.W;
    if (x < p.x0 || x >= p.x1 || y < p.y0 || y >= p.y1) return false;
  }
  return true;
}

// Selection state after `pass` completed passes: the key prefix of the median and the rank that
// is left inside that prefix.  Recomputed from the stored histograms (sequential form; the device
// kernel evaluates the same thing with one warp).
CDP_HD void cdp_metrics_scan(const uint32_t* hist256, uint32_t rank, uint32_t& bin, uint32_t& rank_out) {
  uint32_t acc = 0;
  for (uint32_t d = 0; d < 256; ++d) {
    const uint32_t c = hist256[d];
    if (rank < acc + c) { bin = d; rank_out = rank - acc; return; }
    acc += c;
  }
  bin = 255; rank_out = 0;  // empty selection
}

CDP_HD const uint32_t* cdp_metrics_hist(const CdpMetricsParams& p, int pass, int unit, int arr) {
  return p.hist + (((size_t)pass * p.units + unit) * 2 + arr) * 256;
}

// number of valid elements of a unit = total of its pass-0 histogram (either array)
CDP_HD uint32_t cdp_metrics_count(const CdpMetricsParams& p, int unit) {
  const uint32_t* h = cdp_metrics_hist(p, 0, unit, 0);
  uint32_t n = 0;
  for (int d = 0; d < 256; ++d) n += h[d];
  return n;
}

// prefix (top 8*passes bits, in place) and remaining rank of array `arr` after `passes` passes
CDP_HD void cdp_metrics_state(const CdpMetricsParams& p, int unit, int arr, int passes, uint32_t count,
                              uint32_t& prefix, uint32_t& rank) {
  prefix = 0;
  rank = count ? (count - 1) / 2 : 0;  // lower median (torch.median)
  for (int q = 0; q < passes; ++q) {
    uint32_t bin, r;
    cdp_metrics_scan(cdp_metrics_hist(p, q, unit, arr), rank, bin, r);
    prefix |= bin << (24 - 8 * q);
    rank = r;
  }
}

// does `key` continue the prefix found by the first `pass` passes?
CDP_HD bool cdp_metrics_match(uint32_t key, uint32_t prefix, int pass) {
  return pass == 0 || (key >> (32 - 8 * pass)) == (prefix >> (32 - 8 * pass));
}

// statistics of one valid element (eval/depth.py:56-57, 109-133); ratio = 1 without gt scaling
CDP_HD void cdp_metrics_element(const CdpMetricsParams& p, float g, float pr, float ratio, float acc[CDP_METRICS_NSTATS]) {
  pr = CDP_MUL(pr, ratio);
  g = fminf(fmaxf(g, p.lo), p.hi);
  pr = fminf(fmaxf(pr, p.lo), p.hi);
  const float t = fmaxf(g / pr, pr / g);
  acc[0] += t < 1.25f ? 1.f : 0.f;
  acc[1] += t < 1.5625f ? 1.f : 0.f;
  acc[2] += t < 1.953125f ? 1.f : 0.f;
  const float d = CDP_SUB(g, pr), d2 = CDP_MUL(d, d);
  acc[3] += d2;
  const float dl = CDP_SUB(logf(g), logf(pr));
  acc[4] += CDP_MUL(dl, dl);
  acc[5] += fabsf(d) / g;
  acc[6] += d2 / g;
}

// per unit: means and roots (fp64), returns false if the unit has no ground truth
CDP_HD bool cdp_metrics_unit_stats(const double sums[CDP_METRICS_NSTATS], uint32_t count, double out[CDP_METRICS_NSTATS]) {
  if (count == 0) return false;
  const double n = (double)count;
  out[0] = sums[0] / n; out[1] = sums[1] / n; out[2] = sums[2] / n;
  out[3] = sqrt(sums[3] / n);
  out[4] = sqrt(sums[4] / n);
  out[5] = sums[5] / n;
  out[6] = sums[6] / n;
  return true;
}

static inline bool cdp_fill_metrics_params(CdpMetricsParams* p, const float* gt, const float* pred,
                                           const int64_t* labels, int64_t class_id, int32_t units, int32_t n,
                                           int32_t W, int32_t H, int32_t garg_crop, float lo, float hi,
                                           int32_t use_gt_scale, void* scratch, float* out) {
  if (units < 1 || n < 1 || !(hi >= lo)) return false;
  if (garg_crop && ((size_t)W * H != (size_t)n || W < 1 || H < 1)) return false;
  memset(p, 0, sizeof(*p));
  p->gt = gt; p->pred = pred; p->labels = labels; p->class_id = class_id;
  p->units = units; p->n = n;
  p->blocks = (n + CDP_METRICS_CHUNK - 1) / CDP_METRICS_CHUNK;
  if (garg_crop) {  // eval/depth.py:37-43: int() truncation of the fractional bounds
    p->W = W; p->H = H;
    p->y0 = (int32_t)(0.4080 * H); p->y1 = (int32_t)(0.9891 * H);
    p->x0 = (int32_t)(0.0354 * W); p->x1 = (int32_t)(0.9638 * W);
  }
  p->lo = lo; p->hi = hi; p->use_gt_scale = use_gt_scale;
  p->hist = static_cast<uint32_t*>(scratch);
  p->part = reinterpret_cast<float*>(p->hist + (size_t)CDP_METRICS_PASSES * units * 2 * 256);
  p->out = out;
  return true;
}

static inline size_t cdp_metrics_hist_bytes(int32_t units) {
  return (size_t)CDP_METRICS_PASSES * units * 2 * 256 * sizeof(uint32_t);
}
static inline size_t cdp_metrics_scratch_total(int32_t units, int32_t n) {
  if (units < 1 || n < 1) return 0;
  const size_t blocks = ((size_t)n + CDP_METRICS_CHUNK - 1) / CDP_METRICS_CHUNK;
  return cdp_metrics_hist_bytes(units) + (size_t)units * blocks * CDP_METRICS_REC * sizeof(float);
}
