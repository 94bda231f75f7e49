// cdp_common.h -- definitions shared by the CUDA kernels, the C ABI and the host-side planner.
//
// The per-pixel math (cdp_math.h) and the kernel bodies (cdp_kernels.h) are written as
// host/device functions taking an explicit (block, thread) context: nvcc compiles them into
// the sm_100a kernels of cdp_api.cu, and tests/emu compiles the very same code with g++ into a
// CPU emulator, so tile / halo / adjoint logic can be checked against the oracle in the build
// container, which has no GPU.  The emulator is test infrastructure; nothing in the product
// path links or loads it.
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "codeps_photo.h"

#if defined(__CUDACC__)
#define CDP_HD __host__ __device__ __forceinline__
#else
#define CDP_HD inline
#endif
// Rarely executed branches of the tile kernel (literal-formula warp of clamped points, bilinear taps
// that leave the staged boxes): kept out of line so that the hot loops stay compact -- the kernel is
// sensitive to its instruction-cache footprint (unrolling phase A by two cost 7 %).
#ifndef CDP_OPT_OUTLINE_COLD
#define CDP_OPT_OUTLINE_COLD 1
#endif
#if defined(__CUDACC__) && CDP_OPT_OUTLINE_COLD
#define CDP_COLD __host__ __device__ __noinline__
#elif defined(__CUDACC__)
#define CDP_COLD __host__ __device__ __forceinline__
#else
#define CDP_COLD inline
#endif

#if !defined(__CUDACC__)
struct alignas(8) float2 {  // CUDA's vector types, for the host build of the kernel bodies
  float x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
#endif

// Separately rounded fp32 operations.  The reference evaluates SSIM and the warp chain as
// individual ATen ops, each rounding to fp32; nvcc would otherwise contract a*b+c into FMA.
// Where matching that rounding matters for the argmin the kernels use
// these instead of plain operators.
#if defined(__CUDA_ARCH__)
#define CDP_MUL(a, b) __fmul_rn((a), (b))
#define CDP_ADD(a, b) __fadd_rn((a), (b))
#define CDP_SUB(a, b) __fsub_rn((a), (b))
#define CDP_LDG(p) __ldg(p)
#else
#define CDP_MUL(a, b) ((a) * (b))  // the emulator is built with -ffp-contract=off
#define CDP_ADD(a, b) ((a) + (b))
#define CDP_SUB(a, b) ((a) - (b))
#define CDP_LDG(p) (*(p))
#endif

// ------------------------------------------------------------------------------------------
// Fused photometric tile kernel geometry.
// ------------------------------------------------------------------------------------------
#define CDP_TILE_X 32
#ifndef CDP_TILE_Y
#define CDP_TILE_Y 32
#endif
#define CDP_PHOTO_THREADS 256
#ifndef CDP_PHOTO_MIN_CTAS
#define CDP_PHOTO_MIN_CTAS 2  // resident CTAs per SM the register allocation targets
#endif
// Per-CTA partial record: [0] loss sum, [1..16] dL/dT0, [17..32] dL/dT1 (row-major 4x4), padded.
#define CDP_PARTIAL_STRIDE 36
#define CDP_NOISE_SCALE 0.00001f  // algos/depth.py:317-318
#define CDP_SSIM_C1 0.0001f       // 0.01**2, algos/depth.py:125
#define CDP_SSIM_C2 0.0009f       // 0.03**2, algos/depth.py:126
#define CDP_Z_MIN 1e-5f           // misc/image_warper.py:32

struct CdpCam {
  float fx, fy, cx, cy;
  float ifx, ify;  // 1/fx, 1/fy
};

// Bilinear-resize taps (forward) and their transpose (adjoint gather), built on the host by
// cdp_resize_tables_build and read by the pyramid kernels.
struct CdpResizeTap {  // output index j reads input i0, i1 with weights w0, w1
  int32_t i0, i1;
  float w0, w1;
};
struct CdpResizeInv {  // input index i receives wa*out[ja] + wb*out[jb]; j < 0 = none
  int32_t ja, jb;
  float wa, wb;
};

struct CdpLevel {
  const float* tgt;    // [B,3,H,W]
  const float* src0;   // [B,3,H,W]
  const float* src1;   // [B,3,H,W]
  const float* depth;  // [B,1,H,W]
  const float* noise;  // [B,2,H,W] or null
  float* gdepth;       // [B,H,W] unit gradient dL/d depth_s (with grad)
  uint8_t* argmin;     // [B,H,W] or null
  const float* mot0;   // [B,3,H,W] object motion added to the transformed point, or null
  const float* mot1;   //           (misc/image_warper.py:133-134)
  float* gmot0;        // [B,3,H,W] unit gradients dL/d motion_s (with grad and motion)
  float* gmot1;
  int32_t W, H;
  int32_t tiles_x, tiles_y;
  uint32_t tiles_x_rcp;  // ceil(2^32 / tiles_x): tile / tiles_x = (tile * tiles_x_rcp) >> 32 for tile * tiles_x < 2^32
  int32_t block_begin;  // first block index (within one image) belonging to this level
  float weight;         // 1 / (B * H * W * 2^s * num_levels)
  int32_t use_tma;      // 1: the boxes of this level are staged by TMA (descriptors in CdpTmaMaps)
};

struct CdpPhotoParams {
  CdpLevel lv[CDP_MAX_LEVELS];
  // Per-level intrinsics [L][B][4] (fx,fy,cx,cy) in scratch memory.  Written before the tile kernel
  // by cdp_k_table_kernel: from host values handed over by value, or from the caller's device
  // tensor [B,4] (no host copy of the calibration needed).  Reading the table measured 2-3 %
  // faster than indexing the same values in kernel-parameter space.
  const float* K_tab;
  const float* pose0;  // [B,16]
  const float* pose1;
  float* partials;  // [B][blocks_per_image][CDP_PARTIAL_STRIDE]
  uint64_t seed;
  const uint64_t* seed_dev;  // if set: the seed is read from here (cdp_photo_args.noise_seed_dev)
  int32_t num_levels;
  int32_t batch_begin;  // first sample handled by this launch
  int32_t batch_total;  // B (row stride of K_tab)
  int32_t blocks_per_image;
  float alpha;
};

// parameters of cdp_k_table_kernel (host values travel by value, <= 32 samples per launch)
struct CdpKTableParams {
  const float* K_full;  // device [B,4] full-resolution intrinsics, or null: use K below
  float* K_tab;         // [L][B][4]
  int32_t B, L, batch_begin, batch_count;
  float su[CDP_MAX_LEVELS], sv[CDP_MAX_LEVELS];            // level scales (device mode)
  float K[CDP_MAX_LEVELS][CDP_MAX_BATCH_PER_LAUNCH][4];    // per level, per sample of this launch (host mode)
};

// ------------------------------------------------------------------------------------------
// Host-side plan: level sizes, buffer carve-up, grid shapes.  Pure arithmetic, no CUDA.
// ------------------------------------------------------------------------------------------
struct CdpPlan {
  int32_t B, H, W, L;
  int32_t Ws[CDP_MAX_LEVELS], Hs[CDP_MAX_LEVELS];
  int32_t tiles_x[CDP_MAX_LEVELS], tiles_y[CDP_MAX_LEVELS], block_begin[CDP_MAX_LEVELS];
  int32_t blocks_per_image;
  // scratch (float offsets): pyramid levels 1..L-1 and per-CTA partials
  size_t off_tgt[CDP_MAX_LEVELS], off_src0[CDP_MAX_LEVELS], off_src1[CDP_MAX_LEVELS],
      off_depth[CDP_MAX_LEVELS];
  size_t off_partials;
  size_t off_ktab;      // [L][B][4] per-level intrinsics (device-intrinsics mode)
  size_t scratch_floats;
  // saved (float offsets): per-level unit depth gradients, unit pose gradients [2][B][16]
  size_t off_gdepth[CDP_MAX_LEVELS];
  size_t off_pose_unit;
  // optional object-motion maps: pyramid levels (scratch) and per-level unit gradients (saved)
  int32_t has_motion;
  size_t off_mot[2][CDP_MAX_LEVELS], off_gmot[2][CDP_MAX_LEVELS];
  size_t saved_floats;
  // pyramid kernel: output pixels of levels >= 1 per image, with prefix offsets
  int32_t pyr_begin[CDP_MAX_LEVELS + 1];
  // resize tables (element offsets into the table buffer, in 16-byte records)
  size_t tab_fwd_x[CDP_MAX_LEVELS], tab_fwd_y[CDP_MAX_LEVELS], tab_inv_x[CDP_MAX_LEVELS],
      tab_inv_y[CDP_MAX_LEVELS];
  size_t tab_records;
};

static inline size_t cdp_align_floats(size_t n) { return (n + 63) & ~size_t(63); }  // 256 B

static inline bool cdp_make_plan(int32_t B, int32_t H, int32_t W, int32_t L, CdpPlan* p, bool motion = false) {
  if (B <= 0 || H <= 0 || W <= 0 || L <= 0 || L > CDP_MAX_LEVELS) return false;
  memset(p, 0, sizeof(*p));
  p->B = B; p->H = H; p->W = W; p->L = L;
  p->has_motion = motion ? 1 : 0;
  size_t so = 0, sv = 0, tab = 0;
  int32_t blk = 0, pyr = 0;
  for (int s = 0; s < L; ++s) {
    p->Ws[s] = W >> s;  // integer halving, algos/depth.py:211-214
    p->Hs[s] = H >> s;
    if (p->Ws[s] < 2 || p->Hs[s] < 2) return false;  // reflection padding needs >= 2 pixels
    p->tiles_x[s] = (p->Ws[s] + CDP_TILE_X - 1) / CDP_TILE_X;
    p->tiles_y[s] = (p->Hs[s] + CDP_TILE_Y - 1) / CDP_TILE_Y;
    p->block_begin[s] = blk;
    blk += p->tiles_x[s] * p->tiles_y[s];
    size_t px = (size_t)B * p->Hs[s] * p->Ws[s];
    if (s > 0) {
      p->off_tgt[s] = so; so += cdp_align_floats(3 * px);
      p->off_src0[s] = so; so += cdp_align_floats(3 * px);
      p->off_src1[s] = so; so += cdp_align_floats(3 * px);
      p->off_depth[s] = so; so += cdp_align_floats(px);
      if (motion) {
        p->off_mot[0][s] = so; so += cdp_align_floats(3 * px);
        p->off_mot[1][s] = so; so += cdp_align_floats(3 * px);
      }
      p->pyr_begin[s] = pyr;
      pyr += p->Hs[s] * p->Ws[s];
      p->tab_fwd_x[s] = tab; tab += p->Ws[s];
      p->tab_fwd_y[s] = tab; tab += p->Hs[s];
      p->tab_inv_x[s] = tab; tab += W;
      p->tab_inv_y[s] = tab; tab += H;
    }
    p->off_gdepth[s] = sv; sv += cdp_align_floats(px);
    if (motion) {
      p->off_gmot[0][s] = sv; sv += cdp_align_floats(3 * px);
      p->off_gmot[1][s] = sv; sv += cdp_align_floats(3 * px);
    }
  }
  p->pyr_begin[L] = pyr;
  p->pyr_begin[0] = 0;
  p->blocks_per_image = blk;
  p->off_partials = so;
  so += cdp_align_floats((size_t)B * blk * CDP_PARTIAL_STRIDE);
  p->off_ktab = so;
  so += cdp_align_floats((size_t)L * B * 4);
  p->scratch_floats = so;
  p->off_pose_unit = sv;
  sv += cdp_align_floats((size_t)2 * B * 16);
  p->saved_floats = sv;
  p->tab_records = tab;
  return true;
}

// One axis of F.interpolate(mode="bilinear", align_corners=False): ATen's
// area_pixel_compute_source_index / compute_source_index_and_lambda in fp32.
static inline bool cdp_resize_axis(int32_t in, int32_t out, CdpResizeTap* fwd, CdpResizeInv* inv) {
  for (int i = 0; i < in; ++i) { inv[i].ja = inv[i].jb = -1; inv[i].wa = inv[i].wb = 0.f; }
  const float scale = (float)in / (float)out;
  bool ok = true;
  for (int j = 0; j < out; ++j) {
    CdpResizeTap t;
    if (in == out) {
      t.i0 = t.i1 = j; t.w0 = 1.f; t.w1 = 0.f;
    } else {
      volatile float a = (float)j + 0.5f;
      volatile float m = scale * a;  // volatile: keep the product and the subtraction unfused
      float src = m - 0.5f;
      if (src < 0.f) src = 0.f;
      int i0 = (int)floorf(src);
      if (i0 > in - 1) i0 = in - 1;
      float w1 = src - (float)i0;
      if (w1 < 0.f) w1 = 0.f;
      if (w1 > 1.f) w1 = 1.f;
      t.i0 = i0; t.i1 = i0 + (i0 < in - 1 ? 1 : 0); t.w1 = w1; t.w0 = 1.f - w1;
    }
    fwd[j] = t;
    // transpose: every input index keeps at most two (j, weight) references; down-sampling
    // (in >= out) guarantees that is enough, otherwise the table is reported unusable.
    const int idx[2] = {t.i0, t.i1};
    const float wt[2] = {t.w0, t.w1};
    for (int r = 0; r < 2; ++r) {
      if (wt[r] == 0.f) continue;
      CdpResizeInv& e = inv[idx[r]];
      if (e.ja < 0) { e.ja = j; e.wa = wt[r]; }
      else if (e.jb < 0) { e.jb = j; e.wb = wt[r]; }
      else ok = false;
    }
  }
  return ok;
}
