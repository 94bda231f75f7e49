// cdp_api.cu -- sm_100a kernels and the C ABI of libcodeps_photo.so (see include/codeps_photo.h).
//
// Kernel bodies live in cdp_kernels.h / cdp_math.h; this file adds the __global__ wrappers
// (shared-memory staging, block barriers, fixed-order block reductions) and the host-side entry
// points (argument checks, launch planning, error reporting).  There is no CPU code path here:
// every entry point enqueues CUDA kernels on the caller's stream or fails.
#include <cuda.h>  // CUtensorMap, cuTensorMapEncodeTiled (types only: the entry point is fetched at run time)
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "cdp_plan.h"

// ------------------------------------------------------------------------------------------
// error reporting
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int cdp_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CDP_REQUIRE(cond, ...) \
  do { if (!(cond)) return cdp_fail(CDP_ERR_INVALID, __VA_ARGS__); } while (0)

#define CDP_CUDA(expr) \
  do { cudaError_t e_ = (expr); \
       if (e_ != cudaSuccess) return cdp_fail(CDP_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } while (0)

#define CDP_LAUNCH_CHECK(name) \
  do { cudaError_t e_ = cudaGetLastError(); \
       if (e_ != cudaSuccess) return cdp_fail(CDP_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); } while (0)

extern "C" int cdp_version(void) { return CDP_ABI_VERSION; }
extern "C" const char* cdp_last_error(void) { return g_last_error.c_str(); }

extern "C" int cdp_device_check(void) {
  int dev = 0, major = 0;
  CDP_CUDA(cudaGetDevice(&dev));
  CDP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return cdp_fail(CDP_ERR_UNSUPPORTED, "device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// optional per-kernel timing (cdp_profile_*): CUDA events recorded on the launch stream around
// each kernel while enabled.  Off by default; must not be enabled during stream capture.
// ------------------------------------------------------------------------------------------
namespace {
struct ProfRecord { int id; cudaEvent_t start, stop; };
std::mutex g_prof_mutex;
std::vector<ProfRecord> g_prof_records;
bool g_prof_enabled = false;

struct ProfScope {
  bool on;
  ProfRecord rec;
  cudaStream_t stream;
  ProfScope(int id, cudaStream_t s) : on(g_prof_enabled), stream(s) {
    if (!on) return;
    rec.id = id;
    if (cudaEventCreate(&rec.start) != cudaSuccess || cudaEventCreate(&rec.stop) != cudaSuccess) { on = false; return; }
    cudaEventRecord(rec.start, stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(rec.stop, stream);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_records.push_back(rec);
  }
};
}  // namespace

extern "C" int cdp_profile_enable(int32_t enable) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  for (auto& r : g_prof_records) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
  g_prof_records.clear();
  g_prof_enabled = enable != 0;
  return CDP_OK;
}

extern "C" int cdp_profile_read(int32_t kernel_id, double* total_ms, int32_t* launches) {
  CDP_REQUIRE(total_ms && launches, "null pointer");
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  double tot = 0.0;
  int n = 0;
  for (auto& r : g_prof_records) {
    if (r.id != kernel_id) continue;
    CDP_CUDA(cudaEventSynchronize(r.stop));
    float ms = 0.f;
    CDP_CUDA(cudaEventElapsedTime(&ms, r.start, r.stop));
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *launches = n;
  return CDP_OK;
}

// cudaFuncSetAttribute once per device (and never during stream capture after the first call)
template <typename K>
static cudaError_t cdp_allow_smem(K kernel, size_t bytes, unsigned long long* done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && (*done_mask >> dev) & 1ull) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && dev < 64) *done_mask |= 1ull << dev;
  return e;
}

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL) of the seven kernels of a training step.  Every one of them
// starts with cdp_pdl_enter(): griddepcontrol.launch_dependents lets the NEXT kernel of the stream
// be scheduled as soon as all CTAs of this one have started, griddepcontrol.wait then blocks until
// the PREVIOUS kernel has completed and its memory is visible -- before this kernel touches any
// memory.  The data dependencies are therefore exactly those of ordinary stream order; what
// overlaps is the launch latency and the block scheduling of the dependent kernel with the tail of
// its predecessor.  A kernel launched without the attribute, or after a non-participating kernel,
// behaves as usual (the wait returns at once, the trigger is implied by completion).
// ------------------------------------------------------------------------------------------
// Measured (B200, graph replay of the step, 1024x512 batch 8): 0.539 ms with PDL vs 0.533 ms without --
// the dependent grids become resident early and wait, which costs more than the ~1 us of launch latency
// per kernel boundary it hides.  Off by default; -DCDP_OPT_PDL=1 builds it.
#ifndef CDP_OPT_PDL
#define CDP_OPT_PDL 0
#endif
__device__ __forceinline__ void cdp_pdl_enter() {
#if CDP_OPT_PDL
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
static cudaError_t cdp_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CDP_OPT_PDL ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
// Sum N per-thread values over the block in a fixed order (shuffle tree inside a warp, then
// warps in index order) and let thread j < N write total j to out[j].  red: >= nwarps*N floats.
template <int N>
__device__ __forceinline__ void cdp_block_reduce_store(float (&v)[N], float* red, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], off);
    if (lane == 0) red[warp * N + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += red[w * N + threadIdx.x];
    out[threadIdx.x] = acc;
  }
}

// Same result type for many values per thread (N ~ 33): every thread parks its values in shared
// memory (value-major, conflict free); then eight lanes share one value: lane j of the group adds
// the entries of threads 4 (j + 8 k) .. + 3, k = 0, 1, ... (16-byte loads, four independent chains),
// and three shuffle steps combine the eight lanes.  All warps work at once (32 values in flight per
// 256 threads) on a short dependency chain -- a warp per value with one entry per lane and step was
// ~5x longer on the critical path of every tile.  red: >= N * blockDim.x floats, 16-byte aligned;
// blockDim.x a multiple of 32.  Fixed order, no atomics.
// SKIP_AT > 0: value i >= SKIP_AT is stored at out[i + SKIP_BY] (the caller left a gap of entries
// it knows to be zero out of the reduction).
template <int N, int SKIP_AT = 0, int SKIP_BY = 0>
__device__ __forceinline__ void cdp_block_reduce_store_wide(const float (&v)[N], float* red, float* out) {
  const int nt = blockDim.x, j = threadIdx.x & 7, group = threadIdx.x >> 3, ngroups = nt >> 3;
#pragma unroll
  for (int i = 0; i < N; ++i) red[i * nt + threadIdx.x] = v[i];
  __syncthreads();
  const unsigned gmask = 0xffu << (threadIdx.x & 24);  // the eight lanes of this group (a second pass is not warp-uniform)
  for (int i = group; i < N; i += ngroups) {
    const float4* row = reinterpret_cast<const float4*>(red + i * nt) + j;
    float4 acc = row[0];
    for (int k = 1; k < nt / 32; ++k) {
      const float4 e = row[8 * k];
      acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
    }
    float a = (acc.x + acc.y) + (acc.z + acc.w);
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) a += __shfl_down_sync(gmask, a, off, 8);
    if (j == 0) out[(SKIP_AT > 0 && i >= SKIP_AT) ? i + SKIP_BY : i] = a;
  }
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// kt.batch_count > 0: the first block also writes the per-level intrinsics table (saves the
// separate cdp_k_table_kernel launch when the whole batch fits one parameter block)
#ifndef CDP_PYR_MIN_BLOCKS
#define CDP_PYR_MIN_BLOCKS 4  // 64 registers: 4 blocks per SM (72 registers without the bound measured 5 % slower)
#endif
__global__ void __launch_bounds__(256, CDP_PYR_MIN_BLOCKS) cdp_pyramid_fwd_kernel(const __grid_constant__ CdpPyrParams p,
                                                              const __grid_constant__ CdpKTableParams kt) {
  cdp_pdl_enter();
  if (blockIdx.x == 0 && blockIdx.y == 0)
    for (int i = threadIdx.x; i < kt.batch_count * kt.L; i += blockDim.x)
      cdp_k_table_entry(kt, i / kt.batch_count, i % kt.batch_count);
  if (p.pose_out[0] && blockIdx.x == 0 && blockIdx.y == 0)  // fused heads: 6-DoF parameters -> 4x4 matrices
    for (int j = threadIdx.x; j < 2 * p.B; j += blockDim.x) cdp_pyr_pose_item(p, j);
  cdp_pyramid_fwd_item(p, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void cdp_k_table_kernel(const __grid_constant__ CdpKTableParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.batch_count * p.L) cdp_k_table_entry(p, i / p.batch_count, i % p.batch_count);
}

// ------------------------------------------------------------------------------------------
// TMA staging of the tile kernel's boxes.  One descriptor per (level, tensor): rank 3
// [planes = B*3 (or B), H_s, W_s], box = [3 (or 1), TBH, TBW] for target / depth and
// [3, SBH, SBW] for the sources; out-of-image elements are zero-filled by the hardware.
// ------------------------------------------------------------------------------------------
struct CdpTmaMaps {
  CUtensorMap m[CDP_MAX_LEVELS][4];  // target, depth, source0, source1
};

__device__ __forceinline__ uint32_t cdp_smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cdp_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cdp_smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async (TMA) proxy
}
__device__ __forceinline__ void cdp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cdp_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cdp_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(cdp_smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cdp_tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(cdp_smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(cdp_smem_addr(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

__device__ __forceinline__ void cdp_tma_prefetch_3d(const CUtensorMap* map, int x, int y, int z) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z) : "memory");
}

#ifndef CDP_OPT_REVERSE_TILES
#define CDP_OPT_REVERSE_TILES 1
#endif
#ifndef CDP_EXP_TMA_TWICE
#define CDP_EXP_TMA_TWICE 0  // timing experiment: both source boxes are loaded twice, at the start and in phase S2 (+101 KB of
                             // TMA traffic per tile on top of 124 KB: +1.9 % kernel time, i.e. the mbarrier waits are latency)
#endif
#ifndef CDP_OPT_L2_PREFETCH
#define CDP_OPT_L2_PREFETCH 296  // CTAs ahead whose boxes are prefetched into L2 (0 = off); one wave of 2 x 148 (measured: -0.4 %)
#endif

template <bool G, bool M>
__global__ void __launch_bounds__(CDP_PHOTO_THREADS, CDP_PHOTO_MIN_CTAS)
cdp_photo_kernel(const __grid_constant__ CdpPhotoParams p, const __grid_constant__ CdpTmaMaps tm) {
  typedef CdpTileGeom<G> Geo;
  extern __shared__ __align__(128) float sm[];
  cdp_pdl_enter();
  // Blocks are dispatched in blockIdx.x order; the tile list is walked backwards (coarse levels first):
  // the tiles of the coarse levels are all border tiles (reflected ring, general tap path -- the
  // slowest ones) and would otherwise be the last blocks of the launch (CDP_OPT_REVERSE_TILES).
  const int bx = CDP_OPT_REVERSE_TILES ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const CdpTileCtx c = cdp_tile_ctx(p, bx, blockIdx.y);
  float v[G ? 33 : 1];
#pragma unroll
  for (int i = 0; i < (G ? 33 : 1); ++i) v[i] = 0.f;
  const bool tma = p.lv[c.lvl].use_tma != 0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + Geo::O_MBAR);
  CdpTileConst kc;
  if (tma) {
    // phase S by TMA: one thread arms the mbarriers with the byte counts and issues the four box
    // loads; every thread waits for the data it needs next (barrier 0: depth + sources for phase A,
    // barrier 1: target for phase B1) -- after fetching its own per-tile constants
    // (the issuing thread initialises the barriers and starts the loads right away; the block barrier
    // that makes the initialised barriers visible to the waiting threads comes after the issue)
    if (threadIdx.x == 0) {
      cdp_mbar_init(bar, 1); cdp_mbar_init(bar + 1, 1);
      const int ox = c.x0 - Geo::TXO, oy = c.y0 - Geo::TYO;  // (multiples of 4 in x: 16-byte aligned box starts)
      cdp_mbar_expect_tx(bar, Geo::TMA_BYTES_A + (CDP_EXP_TMA_TWICE ? Geo::TMA_BYTES_SRC : 0u));
      cdp_tma_load_3d(sm + Geo::O_DEPTH, &tm.m[c.lvl][1], bar, ox, oy, c.b);
      cdp_tma_load_3d(sm + Geo::O_SRC, &tm.m[c.lvl][2], bar, ox - Geo::SBM, oy - Geo::SBM, c.b * 3);
      cdp_tma_load_3d(sm + Geo::O_SRC + Geo::SRC_STRIDE, &tm.m[c.lvl][3], bar, ox - Geo::SBM, oy - Geo::SBM, c.b * 3);
#if CDP_EXP_TMA_TWICE  // timing experiment: the same source boxes once more (sensitivity of the kernel to TMA bytes)
      cdp_tma_load_3d(sm + Geo::O_SRC, &tm.m[c.lvl][2], bar, ox - Geo::SBM, oy - Geo::SBM, c.b * 3);
      cdp_tma_load_3d(sm + Geo::O_SRC + Geo::SRC_STRIDE, &tm.m[c.lvl][3], bar, ox - Geo::SBM, oy - Geo::SBM, c.b * 3);
#endif
      cdp_mbar_expect_tx(bar + 1, Geo::TMA_BYTES_TGT);
      cdp_tma_load_3d(sm + Geo::O_TGT, &tm.m[c.lvl][0], bar + 1, ox, oy, c.b * 3);
    }
    __syncthreads();
#if CDP_OPT_L2_PREFETCH > 0
    if (threadIdx.x == 32) {
      // the boxes of the tile that will start about two waves from now: L2 prefetch, so that its
      // TMA loads find them in L2 instead of paying the DRAM latency inside its mbarrier wait
      const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x + CDP_OPT_L2_PREFETCH;
      const unsigned fy = lin / gridDim.x, fx = lin - fy * gridDim.x;
      if (fy < gridDim.y) {
        const CdpTileCtx f = cdp_tile_ctx(p, CDP_OPT_REVERSE_TILES ? (int)(gridDim.x - 1 - fx) : (int)fx, (int)fy);
        if (p.lv[f.lvl].use_tma) {
          const int ox = f.x0 - Geo::TXO, oy = f.y0 - Geo::TYO;
          cdp_tma_prefetch_3d(&tm.m[f.lvl][1], ox, oy, f.b);
          cdp_tma_prefetch_3d(&tm.m[f.lvl][2], ox - Geo::SBM, oy - Geo::SBM, f.b * 3);
          cdp_tma_prefetch_3d(&tm.m[f.lvl][3], ox - Geo::SBM, oy - Geo::SBM, f.b * 3);
          cdp_tma_prefetch_3d(&tm.m[f.lvl][0], ox, oy, f.b * 3);
        }
      }
    }
#endif
    cdp_tile_const(p, c, kc);
    cdp_mbar_wait(bar, 0);
    // (the reflected ring of the target box is written in phase A: needs the target box too on border tiles)
    if (c.x0 == 0 || c.y0 == 0 || c.x0 - Geo::TXO + Geo::TBW > p.lv[c.lvl].W || c.y0 - Geo::TYO + Geo::TBH > p.lv[c.lvl].H)
      cdp_mbar_wait(bar + 1, 0);
  } else {
    cdp_tile_const(p, c, kc);
    cdp_photo_stage<G>(p, c, threadIdx.x, blockDim.x, sm);
    __syncthreads();
  }
  cdp_photo_phase_a<G, M>(p, c, threadIdx.x, blockDim.x, sm, kc);
  if (tma) cdp_mbar_wait(bar + 1, 0);
  __syncthreads();
  cdp_photo_phase_b1<G>(p, c, threadIdx.x, blockDim.x, sm, v[0]);
  if constexpr (G) {
    __syncthreads();
    cdp_photo_phase_b2(p, c, threadIdx.x, blockDim.x, sm);
    __syncthreads();
    cdp_photo_phase_c1(p, c, threadIdx.x, blockDim.x, sm);
    // phase S2: the coefficient planes are dead, the source boxes come back for the sampler adjoint.
    // Generic-proxy writes (coefficients) are ordered before the async-proxy (TMA) writes to the
    // same bytes by a proxy fence on every thread + the block barrier.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tma) {
      if (threadIdx.x == 0) {
        const int ox = c.x0 - Geo::TXO - Geo::SBM, oy = c.y0 - Geo::TYO - Geo::SBM;
        cdp_mbar_expect_tx(bar, Geo::TMA_BYTES_SRC * (CDP_EXP_TMA_TWICE ? 2u : 1u));
        cdp_tma_load_3d(sm + Geo::O_SRC, &tm.m[c.lvl][2], bar, ox, oy, c.b * 3);
        cdp_tma_load_3d(sm + Geo::O_SRC + Geo::SRC_STRIDE, &tm.m[c.lvl][3], bar, ox, oy, c.b * 3);
#if CDP_EXP_TMA_TWICE
        cdp_tma_load_3d(sm + Geo::O_SRC, &tm.m[c.lvl][2], bar, ox, oy, c.b * 3);
        cdp_tma_load_3d(sm + Geo::O_SRC + Geo::SRC_STRIDE, &tm.m[c.lvl][3], bar, ox, oy, c.b * 3);
#endif
      }
      cdp_tile_const(p, c, kc);  // (reloaded rather than kept in 38 registers through B1 / B2 / C1)
      cdp_mbar_wait(bar, 1);     // second use of barrier 0: phase parity 1
    } else {
      cdp_tile_const(p, c, kc);
      cdp_photo_restage_sources(p, c, threadIdx.x, blockDim.x, sm);
      __syncthreads();
    }
    cdp_photo_phase_c2<M>(p, c, threadIdx.x, blockDim.x, sm, &v[1], kc);
  }
  v[0] *= p.lv[c.lvl].weight;
  float* rec = p.partials + ((size_t)c.b * p.blocks_per_image + bx) * CDP_PARTIAL_STRIDE;  // (record order = tile order)
  if constexpr (G) {
    // The last row of dL/dT only receives something from pixels whose depth clamp is active (Q_w
    // drops out of the projection otherwise): when no thread of the block saw one, the block
    // reduction covers the loss and rows 0..2 of the two matrices only (25 instead of 33 values).
    bool row3 = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) row3 = row3 || v[13 + i] != 0.f || v[29 + i] != 0.f;
    // (the barrier also retires the tile planes: shared memory is reused for the reduction)
    if (__syncthreads_or(row3)) {
      cdp_block_reduce_store_wide(v, sm, rec);
    } else {
      float r[25];
      r[0] = v[0];
#pragma unroll
      for (int i = 0; i < 12; ++i) { r[1 + i] = v[1 + i]; r[13 + i] = v[17 + i]; }
      cdp_block_reduce_store_wide<25, 13, 4>(r, sm, rec);
      if (threadIdx.x < 8) rec[13 + (threadIdx.x & 3) + (threadIdx.x >> 2) * 16] = 0.f;
    }
  } else {
    __syncthreads();  // tile planes are dead: reuse shared memory for the reduction
    cdp_block_reduce_store(v, sm, rec);
    if (threadIdx.x >= 1 && threadIdx.x < 33) rec[threadIdx.x] = 0.f;
  }
}

__global__ void __launch_bounds__(CDP_FINALIZE_THREADS) cdp_finalize_kernel(const CdpFinalizeParams p) {
  __shared__ double sm[2048 + 32];
  cdp_pdl_enter();
  cdp_finalize_phase_a(p, blockIdx.x, threadIdx.x, sm);
  __syncthreads();
  cdp_finalize_phase_b(p, blockIdx.x, threadIdx.x, sm);
  __syncthreads();
  cdp_finalize_phase_c(p, blockIdx.x, threadIdx.x, sm);
}

__device__ __forceinline__ double cdp_warp_butterfly(double v) {  // same order as cdp_butterfly_host
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

// first block of the depth-gradient launch: dL/dT = grad_loss * unit gradient, or (fused heads) its
// chain to the 6-DoF parameters
__device__ __forceinline__ void cdp_pose_grad_block(const CdpDepthGradParams& p) {
  if (p.axisangle[0]) {
    for (int j = threadIdx.x; j < 2 * p.B; j += blockDim.x) cdp_pose_grad_heads(p, j);
  } else {
    for (int i = threadIdx.x; i < 2 * p.B * 16; i += blockDim.x) cdp_pose_grad_scale(p, i);
  }
}

template <bool EXACT>
__global__ void __launch_bounds__(256) cdp_depth_grad_kernel(const __grid_constant__ CdpDepthGradParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;  // one row per blockIdx.y: no index division
  if (x < p.W) {
    if (EXACT) cdp_depth_grad_px_exact(p, blockIdx.z, blockIdx.y, x);
    else cdp_depth_grad_px(p, blockIdx.z, blockIdx.y, x);
  }
  if (p.scale_pose && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    cdp_pose_grad_block(p);
}

template <bool ALL_EXACT>
__global__ void __launch_bounds__(256) cdp_depth_grad_quad_kernel(const __grid_constant__ CdpDepthGradParams p) {
  cdp_pdl_enter();
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x < p.W) cdp_depth_grad_quad<ALL_EXACT>(p, blockIdx.z, blockIdx.y, x);
  if (p.scale_pose && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    cdp_pose_grad_block(p);
}

__global__ void __launch_bounds__(CDP_SMOOTH_THREADS) cdp_smooth_main_kernel(const CdpSmoothParams p) {
  __shared__ float sm[CDP_SMOOTH_SMEM_FLOATS];
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  cdp_smooth_phase_load(p, blockIdx.y, blockIdx.x, threadIdx.x, blockDim.x, sm);
  __syncthreads();
  cdp_smooth_phase_edges(p, blockIdx.x, threadIdx.x, blockDim.x, sm, v);
  __syncthreads();
  cdp_smooth_phase_grad(p, blockIdx.y, blockIdx.x, threadIdx.x, blockDim.x, sm, v);
  __syncthreads();  // staged planes are dead: reuse them for the reduction
  cdp_block_reduce_store(v, sm, p.part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4);
}

#ifndef CDP_SMOOTH_Q_MIN_BLOCKS
#define CDP_SMOOTH_Q_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(CDP_SMOOTH_Q_THREADS, CDP_SMOOTH_Q_MIN_BLOCKS) cdp_smooth_quad_kernel(const CdpSmoothParams p) {
  __shared__ float red[(CDP_SMOOTH_Q_THREADS / 32) * 4];
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  cdp_pdl_enter();
  cdp_smooth_quad_thread(p, blockIdx.z, blockIdx.x, blockIdx.y, threadIdx.x, v);
  cdp_block_reduce_store(v, red, p.part + (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 4);
}

// four warps per image (one per accumulated quantity, fixed order inside each), then one thread
// per image derives its scalars and thread 0 combines the images in index order
__global__ void __launch_bounds__(1024) cdp_smooth_finalize_kernel(const CdpSmoothParams p) {
  __shared__ double sums[8][4];
  __shared__ double contrib[8];
  cdp_pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.tiles_x * p.tiles_y;
  double loss = 0.0;
  for (int b0 = 0; b0 < p.B; b0 += 8) {  // 32 warps = 8 images x 4 quantities per round
    const int b = b0 + (warp >> 2), q = warp & 3;
    if (b < p.B) {
      const double s = cdp_warp_butterfly(cdp_lane_sum(p.part + (size_t)b * nb * 4 + q, nb, 4, lane));
      if (lane == 0) sums[warp >> 2][q] = s;
    }
    __syncthreads();
    if (threadIdx.x < 8 && b0 + threadIdx.x < p.B)
      contrib[threadIdx.x] = cdp_smooth_finalize_image(p, b0 + threadIdx.x, sums[threadIdx.x][0], sums[threadIdx.x][1],
                                                       sums[threadIdx.x][2], sums[threadIdx.x][3]);
    __syncthreads();
    if (threadIdx.x == 0)
      for (int w = 0; w < 8 && b0 + w < p.B; ++w) loss += contrib[w];
    __syncthreads();
  }
  if (threadIdx.x == 0) p.loss[0] = (float)loss;
}

__global__ void __launch_bounds__(256)
cdp_smooth_bwd_kernel(const float* g, const float* scal, const float* grad_loss, int plane, float* grad_disp) {
  cdp_pdl_enter();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i < plane) cdp_smooth_bwd_run(g, scal, grad_loss, blockIdx.y, (size_t)plane, i, plane - i < 4 ? plane - i : 4, grad_disp);
}

__global__ void __launch_bounds__(256) cdp_warp_grid_kernel(const __grid_constant__ CdpWarpParams p) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < p.H * p.W) cdp_warp_grid_pixel(p, blockIdx.y, pix);
}

__global__ void __launch_bounds__(256) cdp_warp_image_kernel(const __grid_constant__ CdpWarpParams p) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < p.H * p.W) cdp_warp_image_pixel(p, blockIdx.y, pix);
}

__global__ void __launch_bounds__(256) cdp_warp_bwd_kernel(const __grid_constant__ CdpWarpParams p) {
  __shared__ float red[16 * 8];
  float dT[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dT[i] = 0.f;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < p.H * p.W) cdp_warp_bwd_pixel(p, blockIdx.y, pix, dT);
  const int b = p.batch_begin + blockIdx.y;
  cdp_block_reduce_store(dT, red, p.partials + ((size_t)b * gridDim.x + blockIdx.x) * 16);
}

// sum [B][blocks][16] partials over blocks: one block per image, 16 columns x 16 strided rows
__global__ void __launch_bounds__(256)
cdp_pose_partials_kernel(const float* partials, int blocks, float* grad_pose) {
  __shared__ double sm[16 * 16];
  const int r = threadIdx.x >> 4, j = threadIdx.x & 15, b = blockIdx.x;
  double acc = 0.0;
  for (int i = r; i < blocks; i += 16) acc += (double)partials[((size_t)b * blocks + i) * 16 + j];
  sm[r * 16 + j] = acc;
  __syncthreads();
  if (threadIdx.x < 16) {
    double tot = 0.0;
    for (int i = 0; i < 16; ++i) tot += sm[i * 16 + threadIdx.x];
    grad_pose[b * 16 + threadIdx.x] = (float)tot;
  }
}

__global__ void __launch_bounds__(256)
cdp_ssim_fwd_kernel(const float* x, const float* y, int W, int H, float* out) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < W * H) cdp_ssim_fwd_pixel(x, y, W, H, blockIdx.y, pix, out);
}

__global__ void __launch_bounds__(256)
cdp_ssim_bwd_coef_kernel(const float* go, const float* x, const float* y, int W, int H, float* scratch, size_t total) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < W * H) cdp_ssim_bwd_coef_pixel(go, x, y, W, H, blockIdx.y, pix, scratch, total);
}

__global__ void __launch_bounds__(256)
cdp_ssim_bwd_gather_kernel(const float* x, const float* y, int W, int H, const float* scratch, size_t total,
                           float* gx, float* gy) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < W * H) cdp_ssim_bwd_gather_pixel(x, y, W, H, blockIdx.y, pix, scratch, total, gx, gy);
}

// ------------------------------------------------------------------------------------------
// resize tables
// ------------------------------------------------------------------------------------------
extern "C" size_t cdp_resize_tables_bytes(int32_t height, int32_t width, int32_t num_levels) {
  CdpPlan plan;
  if (!cdp_make_plan(1, height, width, num_levels, &plan)) return 0;
  return (plan.tab_records > 0 ? plan.tab_records : 1) * sizeof(CdpResizeTap);
}

extern "C" int cdp_resize_tables_build(int32_t height, int32_t width, int32_t num_levels, void* host_out,
                                       size_t host_bytes) {
  CdpPlan plan;
  CDP_REQUIRE(cdp_make_plan(1, height, width, num_levels, &plan), "invalid pyramid %dx%d, %d levels", width, height, num_levels);
  CDP_REQUIRE(host_out != nullptr, "host_out is null");
  if (host_bytes < cdp_resize_tables_bytes(height, width, num_levels))
    return cdp_fail(CDP_ERR_WORKSPACE, "resize table buffer too small");
  static_assert(sizeof(CdpResizeTap) == 16 && sizeof(CdpResizeInv) == 16, "table records are 16 bytes");
  int bad = 0;
  if (!cdp_build_resize_tables(plan, host_out, &bad))
    return cdp_fail(CDP_ERR_UNSUPPORTED, "resize adjoint needs more than two taps per pixel at level %d", bad);
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// photometric loss
// ------------------------------------------------------------------------------------------
extern "C" size_t cdp_photo_scratch_bytes(int32_t batch, int32_t height, int32_t width, int32_t num_levels,
                                          int32_t with_motion) {
  CdpPlan plan;
  if (!cdp_make_plan(batch, height, width, num_levels, &plan, with_motion != 0)) return 0;
  return plan.scratch_floats * sizeof(float);
}

extern "C" size_t cdp_photo_saved_bytes(int32_t batch, int32_t height, int32_t width, int32_t num_levels,
                                        int32_t with_motion) {
  CdpPlan plan;
  if (!cdp_make_plan(batch, height, width, num_levels, &plan, with_motion != 0)) return 0;
  return plan.saved_floats * sizeof(float);
}

// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup (no link against libcuda)
typedef CUresult (*CdpEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CdpEncodeTiled cdp_encode_tiled_fn() {
  static CdpEncodeTiled fn = []() -> CdpEncodeTiled {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<CdpEncodeTiled>(sym);
  }();
  return fn;
}

// fp32 tensor [planes][H][W] (contiguous) -> descriptor with a [box_p][box_h][box_w] box
static bool cdp_encode_box(CUtensorMap* map, const float* base, int planes, int H, int W, int box_p, int box_h, int box_w) {
  CdpEncodeTiled enc = cdp_encode_tiled_fn();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (W % 4) != 0) return false;  // pitch must be a multiple of 16 bytes
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_p};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Descriptors for every level whose four tensors qualify; the others keep use_tma = 0 and are
// staged with plain loads.  CDP_PHOTO_TMA=0 in the environment disables TMA staging (A/B runs).
static void cdp_make_tma_maps(const CdpPlan& plan, CdpPhotoParams* kp, CdpTmaMaps* tm) {
  typedef CdpTileGeom<true> Geo;  // box sizes do not depend on the instantiation
  static const bool enabled = []() { const char* e = getenv("CDP_PHOTO_TMA"); return !(e && e[0] == '0'); }();
  memset(tm, 0, sizeof(*tm));
  for (int s = 0; s < plan.L; ++s) {
    CdpLevel& lv = kp->lv[s];
    lv.use_tma = 0;
    if (!enabled) continue;
    const bool ok = cdp_encode_box(&tm->m[s][0], lv.tgt, plan.B * 3, lv.H, lv.W, 3, Geo::TBH, Geo::TBW) &&
                    cdp_encode_box(&tm->m[s][1], lv.depth, plan.B, lv.H, lv.W, 1, Geo::TBH, Geo::TBW) &&
                    cdp_encode_box(&tm->m[s][2], lv.src0, plan.B * 3, lv.H, lv.W, 3, Geo::SBH, Geo::SBW) &&
                    cdp_encode_box(&tm->m[s][3], lv.src1, plan.B * 3, lv.H, lv.W, 3, Geo::SBH, Geo::SBW);
    lv.use_tma = ok ? 1 : 0;
  }
}

static int cdp_batch_chunks(int32_t batch) { return (batch + CDP_MAX_BATCH_PER_LAUNCH - 1) / CDP_MAX_BATCH_PER_LAUNCH; }

__global__ void __launch_bounds__(256)
cdp_tiebreak_noise_kernel(int B, int H, int W, int level, uint64_t seed, float* out) {
  const size_t plane = (size_t)H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= plane * B) return;
  const int b = (int)(i / plane);
  const uint32_t pix = (uint32_t)(i - (size_t)b * plane);
  float n0, n1;
  cdp_noise_pair(seed, pix, (uint32_t)level, (uint32_t)b, n0, n1);
  out[((size_t)b * 2 + 0) * plane + pix] = n0;
  out[((size_t)b * 2 + 1) * plane + pix] = n1;
}

extern "C" int cdp_tiebreak_noise(int32_t batch, int32_t level_height, int32_t level_width, int32_t level,
                                  uint64_t noise_seed, float* out, cdp_stream_t stream_) {
  CDP_REQUIRE(batch > 0 && level_height > 0 && level_width > 0 && level >= 0 && level < CDP_MAX_LEVELS, "invalid shape");
  CDP_REQUIRE(out != nullptr, "null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t n = (size_t)batch * level_height * level_width;
  CDP_REQUIRE(n < ((size_t)1 << 31) * 256, "too many elements for one launch");
  cdp_tiebreak_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(batch, level_height, level_width, level, noise_seed, out);
  CDP_LAUNCH_CHECK("cdp_tiebreak_noise_kernel");
  return CDP_OK;
}

extern "C" int cdp_photo_fwd_launches(int32_t batch, int32_t num_levels) {
  // pyramid (+ intrinsics table), or intrinsics table kernels per 32 samples; tile kernel; reduction
  const bool table_in_pyramid = num_levels > 1 && batch <= CDP_MAX_BATCH_PER_LAUNCH;
  return (num_levels > 1 ? 1 : 0) + (table_in_pyramid ? 0 : cdp_batch_chunks(batch)) + 1 + 1;
}
extern "C" int cdp_photo_bwd_launches(int32_t, int32_t, int32_t with_motion) { return with_motion ? 3 : 1; }

// (defined with the stand-alone head conversions further down)
__global__ void cdp_pose_fwd_kernel(const float* aa, const float* tr, int B, int invert, float* M);
__global__ void __launch_bounds__(256) cdp_disp_to_depth_fwd_kernel(const float* disp, size_t n, float lo, float span, float* depth);
static int cdp_elementwise_grid(size_t n);

extern "C" int cdp_photo_fwd(const cdp_photo_args* a, cdp_stream_t stream_) {
  CDP_REQUIRE(a != nullptr, "args is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CdpPlan plan;
  CDP_REQUIRE((a->motion0 == nullptr) == (a->motion1 == nullptr), "motion maps must be given for both sources or for none");
  CDP_REQUIRE(cdp_make_plan(a->batch, a->height, a->width, a->num_levels, &plan, a->motion0 != nullptr),
              "invalid shape: batch %d, %dx%d, %d levels (every level needs >= 2x2 pixels, at most %d levels)",
              a->batch, a->width, a->height, a->num_levels, CDP_MAX_LEVELS);
  CDP_REQUIRE((reinterpret_cast<uintptr_t>(a->pose0) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->pose1) & 15) == 0,
              "pose0 / pose1 must be 16-byte aligned");
  CDP_REQUIRE(a->batch <= 65535, "batch %d exceeds the grid limit of one launch", a->batch);
  CDP_REQUIRE((a->intrinsics_host != nullptr) != (a->intrinsics_dev != nullptr),
              "exactly one of intrinsics_host / intrinsics_dev must be set");
  CDP_REQUIRE(a->target && a->source0 && a->source1 && a->depth && a->pose0 && a->pose1 && a->loss,
              "null tensor pointer");
  CDP_REQUIRE(a->scratch != nullptr, "scratch is null");
  if (a->scratch_bytes < plan.scratch_floats * sizeof(float))
    return cdp_fail(CDP_ERR_WORKSPACE, "scratch too small: %zu < %zu", a->scratch_bytes, plan.scratch_floats * sizeof(float));
  const bool G = a->with_grad != 0;
  if (G) {
    CDP_REQUIRE(a->saved != nullptr, "with_grad needs a saved buffer");
    if (a->saved_bytes < plan.saved_floats * sizeof(float))
      return cdp_fail(CDP_ERR_WORKSPACE, "saved too small: %zu < %zu", a->saved_bytes, plan.saved_floats * sizeof(float));
  }
  CDP_REQUIRE(plan.L == 1 || a->resize_tables != nullptr, "resize_tables is null");
  bool any_noise = false, all_noise = true;
  for (int s = 0; s < plan.L; ++s) { any_noise |= a->noise[s] != nullptr; all_noise &= a->noise[s] != nullptr; }
  CDP_REQUIRE(!any_noise || all_noise, "noise must be given for every level or for none");

  // fused heads: disparity -> depth and 6-DoF -> matrices ride along with the pyramid launch; where
  // that launch cannot do it (no pyramid, or no level-1 fast path) they are separate small kernels
  const cdp_photo_heads* hd = a->heads;
  if (hd) {
    CDP_REQUIRE((hd->axisangle[0] != nullptr) == (hd->axisangle[1] != nullptr) &&
                (hd->axisangle[0] != nullptr) == (hd->translation[0] != nullptr) &&
                (hd->axisangle[1] != nullptr) == (hd->translation[1] != nullptr),
                "heads: axis-angle and translation must be given for both sources or for none");
    CDP_REQUIRE(!hd->disp || (hd->min_depth > 0.f && hd->max_depth > hd->min_depth), "heads: need 0 < min_depth < max_depth");
    CdpPyrParams probe;
    if (plan.L > 1) cdp_fill_pyr_params(plan, a, &probe);
    if (hd->disp && (plan.L == 1 || !probe.depth_out)) {
      const size_t count = (size_t)plan.B * plan.H * plan.W;
      const float lo = 1.0f / hd->max_depth, span = 1.0f / hd->min_depth - lo;
      cdp_disp_to_depth_fwd_kernel<<<cdp_elementwise_grid(count), 256, 0, stream>>>(hd->disp, count, lo, span, const_cast<float*>(a->depth));
      CDP_LAUNCH_CHECK("cdp_disp_to_depth_fwd_kernel");
    }
    if (hd->axisangle[0] && plan.L == 1) {
      for (int k = 0; k < 2; ++k)
        cdp_pose_fwd_kernel<<<(plan.B + 127) / 128, 128, 0, stream>>>(hd->axisangle[k], hd->translation[k], plan.B, hd->invert[k],
                                                                  const_cast<float*>(k == 0 ? a->pose0 : a->pose1));
      CDP_LAUNCH_CHECK("cdp_pose_fwd_kernel");
    }
  }

  // the per-level intrinsics table rides along with the pyramid launch when there is one and the
  // whole batch fits its parameter block; otherwise cdp_k_table_kernel writes it (step 2)
  const bool table_in_pyramid = plan.L > 1 && plan.B <= CDP_MAX_BATCH_PER_LAUNCH;
  // 1. pyramid
  if (plan.L > 1) {
    CdpPyrParams pp;
    cdp_fill_pyr_params(plan, a, &pp);
    dim3 grid((pp.begin[plan.L] + 255) / 256, plan.B);
    CdpKTableParams tp;
    cdp_fill_k_table_params(plan, a, 0, table_in_pyramid ? plan.B : 0, &tp);
    { ProfScope prof_(CDP_KERNEL_PYRAMID, stream); CDP_CUDA(cdp_launch_pdl(cdp_pyramid_fwd_kernel, grid, dim3(256), 0, stream, pp, tp)); }
    CDP_LAUNCH_CHECK("cdp_pyramid_fwd_kernel");
  }

  // 2. fused tile kernel, all levels per launch, <= CDP_MAX_BATCH_PER_LAUNCH samples per launch
  const size_t smem = G ? CdpTileGeom<true>::SMEM_BYTES : CdpTileGeom<false>::SMEM_BYTES;
  static unsigned long long smem_done[4] = {0ull, 0ull, 0ull, 0ull};
  const bool M = plan.has_motion != 0;
  typedef void (*PhotoKernel)(const CdpPhotoParams, const CdpTmaMaps);
  static const PhotoKernel kernels[4] = {cdp_photo_kernel<false, false>, cdp_photo_kernel<false, true>,
                                         cdp_photo_kernel<true, false>, cdp_photo_kernel<true, true>};
  const int which = (G ? 2 : 0) + (M ? 1 : 0);
  const PhotoKernel photo_kernel = kernels[which];
  CDP_CUDA(cdp_allow_smem(photo_kernel, smem, &smem_done[which]));
  // per-level intrinsics table: host values travel by value (<= 32 samples per launch), device
  // values are rescaled per level
  for (int b0 = 0; b0 < plan.B && !table_in_pyramid; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(plan.B, b0);
    CdpKTableParams tp;
    cdp_fill_k_table_params(plan, a, b0, nb, &tp);
    cdp_k_table_kernel<<<(nb * plan.L + 127) / 128, 128, 0, stream>>>(tp);
    CDP_LAUNCH_CHECK("cdp_k_table_kernel");
  }
  if (a->noise_ready) CDP_CUDA(cudaStreamWaitEvent(stream, static_cast<cudaEvent_t>(a->noise_ready), 0));
  {
    CdpPhotoParams kp;
    cdp_fill_photo_params(plan, a, 0, plan.B, &kp);
    CdpTmaMaps tm;
    cdp_make_tma_maps(plan, &kp, &tm);
    dim3 grid(plan.blocks_per_image, plan.B);
    {
      ProfScope prof_(CDP_KERNEL_PHOTO, stream);
      CDP_CUDA(cdp_launch_pdl(photo_kernel, grid, dim3(CDP_PHOTO_THREADS), smem, stream, kp, tm));
    }
    CDP_LAUNCH_CHECK("cdp_photo_kernel");
  }

  // 3. fixed-order reduction of the per-CTA records
  CdpFinalizeParams fp;
  cdp_fill_finalize_params(plan, a, &fp);
  { ProfScope prof_(CDP_KERNEL_FINALIZE, stream); CDP_CUDA(cdp_launch_pdl(cdp_finalize_kernel, dim3(plan.B), dim3(CDP_FINALIZE_THREADS), 0, stream, fp)); }
  CDP_LAUNCH_CHECK("cdp_finalize_kernel");
  return CDP_OK;
}

// adjoint of the pyramid: four pixels per thread where the layout allows it, else one
static void cdp_launch_depth_grad(const CdpDepthGradParams& p, cudaStream_t stream) {
  if (cdp_depth_grad_quad_ok(p)) {
    const int threads = p.W / 4 >= 256 ? 256 : (p.W / 4 >= 128 ? 128 : 64);
    dim3 grid((p.W / 4 + threads - 1) / threads, p.H, p.B);
    if (cdp_depth_grad_all_exact(p)) cdp_launch_pdl(cdp_depth_grad_quad_kernel<true>, grid, dim3(threads), 0, stream, p);
    else cdp_launch_pdl(cdp_depth_grad_quad_kernel<false>, grid, dim3(threads), 0, stream, p);
    return;
  }
  dim3 grid((p.W + 255) / 256, p.H, p.B);
  if (cdp_depth_grad_all_exact(p)) cdp_depth_grad_kernel<true><<<grid, 256, 0, stream>>>(p);
  else cdp_depth_grad_kernel<false><<<grid, 256, 0, stream>>>(p);
}

extern "C" int cdp_photo_bwd(int32_t batch, int32_t height, int32_t width, int32_t num_levels, const void* saved_,
                             size_t saved_bytes, const void* resize_tables, const float* grad_loss,
                             float* grad_depth, float* grad_pose0, float* grad_pose1, int32_t with_motion,
                             float* grad_motion0, float* grad_motion1, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CdpPlan plan;
  CDP_REQUIRE(cdp_make_plan(batch, height, width, num_levels, &plan, with_motion != 0), "invalid shape");
  CDP_REQUIRE(!with_motion || (grad_motion0 && grad_motion1), "with_motion needs grad_motion0 and grad_motion1");
  CDP_REQUIRE(saved_ && grad_loss && grad_depth && grad_pose0 && grad_pose1, "null pointer");
  CDP_REQUIRE(plan.L == 1 || resize_tables != nullptr, "resize_tables is null");
  if (saved_bytes < plan.saved_floats * sizeof(float)) return cdp_fail(CDP_ERR_WORKSPACE, "saved too small");
  CdpDepthGradParams p;
  cdp_fill_depth_grad_params(plan, saved_, resize_tables, grad_loss, grad_depth, grad_pose0, grad_pose1, &p);
  {
    ProfScope prof_(CDP_KERNEL_DEPTH_GRAD, stream);
    cdp_launch_depth_grad(p, stream);
  }
  CDP_LAUNCH_CHECK("cdp_depth_grad_kernel");
  if (with_motion) {
    float* outs[2] = {grad_motion0, grad_motion1};
    for (int k = 0; k < 2; ++k) {
      CdpDepthGradParams pm;
      cdp_fill_motion_grad_params(plan, saved_, resize_tables, grad_loss, k, outs[k], &pm);
      ProfScope prof_(CDP_KERNEL_DEPTH_GRAD, stream);
      cdp_launch_depth_grad(pm, stream);
      CDP_LAUNCH_CHECK("cdp_depth_grad_kernel (motion)");
    }
  }
  return CDP_OK;
}

extern "C" int cdp_photo_bwd_heads(int32_t batch, int32_t height, int32_t width, int32_t num_levels, const void* saved_,
                                   size_t saved_bytes, const void* resize_tables, const float* grad_loss,
                                   const cdp_photo_heads* heads, const float* depth, float* grad_disp,
                                   float* grad_axisangle0, float* grad_translation0, float* grad_axisangle1,
                                   float* grad_translation1, int32_t with_motion, float* grad_motion0,
                                   float* grad_motion1, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CdpPlan plan;
  CDP_REQUIRE(cdp_make_plan(batch, height, width, num_levels, &plan, with_motion != 0), "invalid shape");
  CDP_REQUIRE(heads && heads->disp && heads->axisangle[0] && heads->axisangle[1] && heads->translation[0] && heads->translation[1],
              "cdp_photo_bwd_heads needs the disparity and both sources' 6-DoF parameters");
  CDP_REQUIRE(!with_motion || (grad_motion0 && grad_motion1), "with_motion needs grad_motion0 and grad_motion1");
  CDP_REQUIRE(saved_ && grad_loss && depth && grad_disp && grad_axisangle0 && grad_translation0 && grad_axisangle1 &&
              grad_translation1, "null pointer");
  CDP_REQUIRE(plan.L == 1 || resize_tables != nullptr, "resize_tables is null");
  if (saved_bytes < plan.saved_floats * sizeof(float)) return cdp_fail(CDP_ERR_WORKSPACE, "saved too small");
  CdpDepthGradParams p;
  cdp_fill_depth_grad_params(plan, saved_, resize_tables, grad_loss, grad_disp, nullptr, nullptr, &p);
  cdp_depth_grad_params_heads(heads, depth, grad_axisangle0, grad_translation0, grad_axisangle1, grad_translation1, &p);
  {
    ProfScope prof_(CDP_KERNEL_DEPTH_GRAD, stream);
    cdp_launch_depth_grad(p, stream);
  }
  CDP_LAUNCH_CHECK("cdp_depth_grad_kernel (heads)");
  if (with_motion) {
    float* outs[2] = {grad_motion0, grad_motion1};
    for (int k = 0; k < 2; ++k) {
      CdpDepthGradParams pm;
      cdp_fill_motion_grad_params(plan, saved_, resize_tables, grad_loss, k, outs[k], &pm);
      ProfScope prof_(CDP_KERNEL_DEPTH_GRAD, stream);
      cdp_launch_depth_grad(pm, stream);
      CDP_LAUNCH_CHECK("cdp_depth_grad_kernel (motion)");
    }
  }
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// smoothness
// ------------------------------------------------------------------------------------------
extern "C" size_t cdp_smooth_saved_bytes(int32_t batch, int32_t height, int32_t width) {
  if (batch <= 0 || height < 2 || width < 2) return 0;
  return cdp_smooth_layout(batch, height, width).total * sizeof(float);
}

extern "C" int cdp_smooth_fwd(const float* image, const float* disp, int32_t batch, int32_t height, int32_t width,
                              int32_t with_grad, float* loss, void* saved_, size_t saved_bytes, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0 && height >= 2 && width >= 2, "invalid shape: batch %d, %dx%d (needs >= 2x2)", batch, width, height);
  CDP_REQUIRE(image && disp && loss && saved_, "null pointer");
  const CdpSmoothLayout l = cdp_smooth_layout(batch, height, width);
  if (saved_bytes < l.total * sizeof(float)) return cdp_fail(CDP_ERR_WORKSPACE, "saved too small: %zu < %zu", saved_bytes, l.total * sizeof(float));
  float* saved = static_cast<float*>(saved_);
  CdpSmoothParams p;
  cdp_fill_smooth_params(image, disp, batch, height, width, with_grad, loss, saved, &p);
  if (cdp_smooth_quad_ok(p)) {
    dim3 grid(p.tiles_x, p.tiles_y, batch);
    ProfScope prof_(CDP_KERNEL_SMOOTH_MAIN, stream);
    CDP_CUDA(cdp_launch_pdl(cdp_smooth_quad_kernel, grid, dim3(CDP_SMOOTH_Q_THREADS), 0, stream, p));
  } else {
    dim3 grid(p.tiles_x * p.tiles_y, batch);
    ProfScope prof_(CDP_KERNEL_SMOOTH_MAIN, stream);
    cdp_smooth_main_kernel<<<grid, CDP_SMOOTH_THREADS, 0, stream>>>(p);
  }
  CDP_LAUNCH_CHECK("cdp_smooth_main_kernel");
  { ProfScope prof_(CDP_KERNEL_SMOOTH_FINALIZE, stream); CDP_CUDA(cdp_launch_pdl(cdp_smooth_finalize_kernel, dim3(1), dim3(1024), 0, stream, p)); }
  CDP_LAUNCH_CHECK("cdp_smooth_finalize_kernel");
  return CDP_OK;
}

extern "C" int cdp_smooth_bwd(const void* saved_, size_t saved_bytes, const float* grad_loss, int32_t batch,
                              int32_t height, int32_t width, float* grad_disp, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0 && height >= 2 && width >= 2, "invalid shape");
  CDP_REQUIRE(saved_ && grad_loss && grad_disp, "null pointer");
  const CdpSmoothLayout l = cdp_smooth_layout(batch, height, width);
  if (saved_bytes < l.total * sizeof(float)) return cdp_fail(CDP_ERR_WORKSPACE, "saved too small");
  const float* saved = static_cast<const float*>(saved_);
  const int plane = height * width;
  dim3 grid(((plane + 3) / 4 + 255) / 256, batch);
  { ProfScope prof_(CDP_KERNEL_SMOOTH_BWD, stream); CDP_CUDA(cdp_launch_pdl(cdp_smooth_bwd_kernel, grid, dim3(256), 0, stream, (const float*)(saved + l.g), (const float*)(saved + l.scal), grad_loss, plane, grad_disp)); }
  CDP_LAUNCH_CHECK("cdp_smooth_bwd_kernel");
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// stand-alone operators
// ------------------------------------------------------------------------------------------
extern "C" int cdp_warp_grid_fwd(const float* depth, const float* pose, const float* motion, const float* K,
                                 int32_t batch, int32_t height, int32_t width, float* grid_out, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0 && height >= 2 && width >= 2, "invalid shape");
  CDP_REQUIRE(depth && pose && K && grid_out, "null pointer");
  for (int b0 = 0; b0 < batch; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(batch, b0);
    CdpWarpParams p;
    cdp_fill_warp_params(&p, nullptr, 0, depth, pose, motion, K, b0, nb, height, width);
    p.out = grid_out;
    dim3 grid((height * width + 255) / 256, nb);
    cdp_warp_grid_kernel<<<grid, 256, 0, stream>>>(p);
    CDP_LAUNCH_CHECK("cdp_warp_grid_kernel");
  }
  return CDP_OK;
}

extern "C" int cdp_warp_image_fwd(const float* src, int32_t channels, const float* depth, const float* pose,
                                  const float* motion, const float* K, int32_t batch, int32_t height, int32_t width,
                                  int32_t mode, float* out, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0 && height >= 2 && width >= 2 && channels > 0, "invalid shape");
  CDP_REQUIRE(src && depth && pose && K && out, "null pointer");
  CDP_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (bilinear) or 1 (nearest)");
  for (int b0 = 0; b0 < batch; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(batch, b0);
    CdpWarpParams p;
    cdp_fill_warp_params(&p, src, channels, depth, pose, motion, K, b0, nb, height, width);
    p.out = out; p.mode = mode;
    dim3 grid((height * width + 255) / 256, nb);
    cdp_warp_image_kernel<<<grid, 256, 0, stream>>>(p);
    CDP_LAUNCH_CHECK("cdp_warp_image_kernel");
  }
  return CDP_OK;
}

extern "C" size_t cdp_warp_bwd_scratch_bytes(int32_t batch, int32_t height, int32_t width) {
  if (batch <= 0 || height <= 0 || width <= 0) return 0;
  return (size_t)batch * ((height * width + 255) / 256) * 16 * sizeof(float);
}

extern "C" int cdp_warp_image_bwd(const float* grad_out, const float* src, int32_t channels, const float* depth,
                                  const float* pose, const float* motion, const float* K, int32_t batch,
                                  int32_t height, int32_t width, float* grad_depth, float* grad_pose,
                                  float* grad_motion, void* scratch, size_t scratch_bytes, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0 && height >= 2 && width >= 2 && channels > 0, "invalid shape");
  CDP_REQUIRE(grad_out && src && depth && pose && K && grad_depth && grad_pose && scratch, "null pointer");
  CDP_REQUIRE(!(grad_motion != nullptr && motion == nullptr), "grad_motion without motion");
  if (scratch_bytes < cdp_warp_bwd_scratch_bytes(batch, height, width)) return cdp_fail(CDP_ERR_WORKSPACE, "scratch too small");
  const int blocks = (height * width + 255) / 256;
  for (int b0 = 0; b0 < batch; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(batch, b0);
    CdpWarpParams p;
    cdp_fill_warp_params(&p, src, channels, depth, pose, motion, K, b0, nb, height, width);
    p.grad_out = grad_out; p.grad_depth = grad_depth; p.grad_motion = grad_motion;
    p.partials = static_cast<float*>(scratch);
    dim3 grid(blocks, nb);
    cdp_warp_bwd_kernel<<<grid, 256, 0, stream>>>(p);
    CDP_LAUNCH_CHECK("cdp_warp_bwd_kernel");
  }
  cdp_pose_partials_kernel<<<batch, 256, 0, stream>>>(static_cast<const float*>(scratch), blocks, grad_pose);
  CDP_LAUNCH_CHECK("cdp_pose_partials_kernel");
  return CDP_OK;
}

extern "C" int cdp_ssim_fwd(const float* x, const float* y, int32_t planes, int32_t height, int32_t width, float* out,
                            cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(planes > 0 && height >= 2 && width >= 2, "invalid shape (reflection padding needs >= 2x2)");
  CDP_REQUIRE(x && y && out, "null pointer");
  CDP_REQUIRE(planes <= 65535, "too many planes");
  dim3 grid((height * width + 255) / 256, planes);
  cdp_ssim_fwd_kernel<<<grid, 256, 0, stream>>>(x, y, width, height, out);
  CDP_LAUNCH_CHECK("cdp_ssim_fwd_kernel");
  return CDP_OK;
}

extern "C" size_t cdp_ssim_bwd_scratch_bytes(int32_t planes, int32_t height, int32_t width) {
  if (planes <= 0 || height <= 0 || width <= 0) return 0;
  return (size_t)4 * planes * height * width * sizeof(float);
}

extern "C" int cdp_ssim_bwd(const float* grad_out, const float* x, const float* y, int32_t planes, int32_t height,
                            int32_t width, float* grad_x, float* grad_y, void* scratch, size_t scratch_bytes,
                            cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(planes > 0 && height >= 2 && width >= 2, "invalid shape");
  CDP_REQUIRE(grad_out && x && y && scratch, "null pointer");
  CDP_REQUIRE(planes <= 65535, "too many planes");
  if (scratch_bytes < cdp_ssim_bwd_scratch_bytes(planes, height, width)) return cdp_fail(CDP_ERR_WORKSPACE, "scratch too small");
  const size_t total = (size_t)planes * height * width;
  dim3 grid((height * width + 255) / 256, planes);
  cdp_ssim_bwd_coef_kernel<<<grid, 256, 0, stream>>>(grad_out, x, y, width, height, static_cast<float*>(scratch), total);
  CDP_LAUNCH_CHECK("cdp_ssim_bwd_coef_kernel");
  cdp_ssim_bwd_gather_kernel<<<grid, 256, 0, stream>>>(x, y, width, height, static_cast<const float*>(scratch), total, grad_x, grad_y);
  CDP_LAUNCH_CHECK("cdp_ssim_bwd_gather_kernel");
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// pose / depth conversions
// ------------------------------------------------------------------------------------------
__global__ void cdp_pose_fwd_kernel(const float* aa, const float* tr, int B, int invert, float* M) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) cdp_pose_fwd_sample(aa + 3 * b, tr + 3 * b, invert, M + 16 * b);
}
__global__ void cdp_pose_bwd_kernel(const float* gM, const float* aa, const float* tr, int B, int invert, float* gaa,
                                    float* gtr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) cdp_pose_bwd_sample(gM + 16 * b, aa + 3 * b, tr + 3 * b, invert, gaa + 3 * b, gtr + 3 * b);
}
__global__ void __launch_bounds__(256) cdp_disp_to_depth_fwd_kernel(const float* disp, size_t n, float lo, float span, float* depth) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    depth[i] = cdp_disp_to_depth(__ldg(disp + i), lo, span);
}
__global__ void __launch_bounds__(256) cdp_disp_to_depth_bwd_kernel(const float* gdepth, const float* depth, size_t n, float span, float* gdisp) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    gdisp[i] = cdp_disp_to_depth_grad(__ldg(gdepth + i), __ldg(depth + i), span);
}

extern "C" int cdp_pose_fwd(const float* axisangle, const float* translation, int32_t batch, int32_t invert, float* T,
                            cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0, "invalid batch");
  CDP_REQUIRE(axisangle && translation && T, "null pointer");
  cdp_pose_fwd_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(axisangle, translation, batch, invert, T);
  CDP_LAUNCH_CHECK("cdp_pose_fwd_kernel");
  return CDP_OK;
}

extern "C" int cdp_pose_bwd(const float* grad_T, const float* axisangle, const float* translation, int32_t batch,
                            int32_t invert, float* grad_axisangle, float* grad_translation, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0, "invalid batch");
  CDP_REQUIRE(grad_T && axisangle && translation && grad_axisangle && grad_translation, "null pointer");
  cdp_pose_bwd_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(grad_T, axisangle, translation, batch, invert,
                                                              grad_axisangle, grad_translation);
  CDP_LAUNCH_CHECK("cdp_pose_bwd_kernel");
  return CDP_OK;
}

static int cdp_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// grid-stride elementwise launches: a multiple of the SM count
static int cdp_elementwise_grid(size_t n) {
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)cdp_sm_count() * 16;
  return (int)(blocks < cap ? blocks : cap);
}

extern "C" int cdp_disp_to_depth_fwd(const float* disp, size_t count, float min_depth, float max_depth, float* depth,
                                     cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(count > 0 && min_depth > 0.f && max_depth > min_depth, "invalid arguments");
  CDP_REQUIRE(disp && depth, "null pointer");
  const float lo = 1.0f / max_depth, span = 1.0f / min_depth - 1.0f / max_depth;
  cdp_disp_to_depth_fwd_kernel<<<cdp_elementwise_grid(count), 256, 0, stream>>>(disp, count, lo, span, depth);
  CDP_LAUNCH_CHECK("cdp_disp_to_depth_fwd_kernel");
  return CDP_OK;
}

extern "C" int cdp_disp_to_depth_bwd(const float* grad_depth, const float* depth, size_t count, float min_depth,
                                     float max_depth, float* grad_disp, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(count > 0 && min_depth > 0.f && max_depth > min_depth, "invalid arguments");
  CDP_REQUIRE(grad_depth && depth && grad_disp, "null pointer");
  const float span = 1.0f / min_depth - 1.0f / max_depth;
  cdp_disp_to_depth_bwd_kernel<<<cdp_elementwise_grid(count), 256, 0, stream>>>(grad_depth, depth, count, span, grad_disp);
  CDP_LAUNCH_CHECK("cdp_disp_to_depth_bwd_kernel");
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// object-motion regularisers (cdp_flow.h): FlowSmoothnessLoss / FlowSparsityLoss
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float cdp_block_sum(float v, float* red) {  // fixed order; result valid in thread 0
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float acc = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < nwarps; ++w) acc += red[w];
  return acc;
}

__global__ void __launch_bounds__(CDP_FLOW_THREADS) cdp_flow_smooth_kernel(const __grid_constant__ CdpFlowParams p) {
  __shared__ float red[CDP_FLOW_THREADS / 32];
  const float v = cdp_flow_smooth_thread(p, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
  const float s = cdp_block_sum(v, red);
  if (threadIdx.x == 0) p.part[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

__global__ void __launch_bounds__(CDP_FLOW_THREADS) cdp_flow_abs_kernel(const __grid_constant__ CdpFlowParams p) {
  __shared__ float red[CDP_FLOW_THREADS / 32];
  const float v = cdp_flow_abs_thread(p, blockIdx.x, blockIdx.y, threadIdx.x);
  const float s = cdp_block_sum(v, red);
  if (threadIdx.x == 0) p.part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
}

__global__ void __launch_bounds__(CDP_FLOW_THREADS) cdp_flow_sparsity_kernel(const __grid_constant__ CdpFlowParams p) {
  __shared__ float red[CDP_FLOW_THREADS / 32];
  __shared__ float mean_s;
  if (threadIdx.x < 32) {  // every block re-derives its plane's mean from the pass-1 records (<= a few dozen)
    double acc = cdp_lane_sum(p.part + (size_t)blockIdx.y * gridDim.x, gridDim.x, 1, threadIdx.x);
    acc = cdp_warp_butterfly(acc);
    if (threadIdx.x == 0) mean_s = (float)(acc / (double)((size_t)p.W * p.H));
  }
  __syncthreads();
  const float v = cdp_flow_sparsity_thread(p, blockIdx.x, blockIdx.y, threadIdx.x, mean_s);
  const float s = cdp_block_sum(v, red);
  float* part2 = p.part + (size_t)gridDim.x * gridDim.y;
  if (threadIdx.x == 0) part2[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
}

// loss = scale * sum(records), thread t sums records t, t+1024, ... (= cdp_flow_sum_records_host)
__global__ void __launch_bounds__(1024) cdp_sum_records_kernel(const float* part, size_t count, double scale, float* loss) {
  __shared__ double warp_sum[32];
  double acc = 0.0;
  for (size_t i = threadIdx.x; i < count; i += 1024) acc += (double)part[i];
  acc = cdp_warp_butterfly(acc);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int w = 0; w < 32; ++w) total += warp_sum[w];
    loss[0] = (float)(total * scale);
  }
}

__global__ void __launch_bounds__(256) cdp_scale_kernel(const float* in, const float* scalar, size_t n, float* out) {
  const float s = __ldg(scalar);
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(out)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[n4 * 4 + threadIdx.x] = __ldg(in + n4 * 4 + threadIdx.x) * s;
}

extern "C" size_t cdp_flow_scratch_bytes(int32_t n_maps, int32_t planes, int32_t height, int32_t width, int32_t sparsity) {
  return cdp_flow_records(n_maps, planes, height, width, sparsity != 0) * sizeof(float);
}

static int cdp_flow_common(const float* const* maps, int32_t n_maps, int32_t planes, int32_t height, int32_t width,
                           int32_t wrap, bool sparsity, float* loss, float* unit_grad, void* scratch,
                           size_t scratch_bytes, CdpFlowParams* p) {
  CDP_REQUIRE(maps && loss && scratch, "null pointer");
  CDP_REQUIRE(n_maps >= 1 && n_maps <= CDP_MAX_FLOW_MAPS, "n_maps %d outside [1, %d]", n_maps, CDP_MAX_FLOW_MAPS);
  for (int i = 0; i < n_maps; ++i) CDP_REQUIRE(maps[i], "null map pointer");
  CDP_REQUIRE(cdp_fill_flow_params(maps, n_maps, planes, height, width, wrap, sparsity, loss, unit_grad,
                                   static_cast<float*>(scratch), p),
              "invalid shape: %d planes of %dx%d%s", planes, width, height,
              (!sparsity && !wrap) ? " (without wrap-around needs >= 2x2)" : "");
  const size_t need = cdp_flow_records(n_maps, planes, height, width, sparsity) * sizeof(float);
  if (scratch_bytes < need) return cdp_fail(CDP_ERR_WORKSPACE, "scratch too small: %zu < %zu", scratch_bytes, need);
  CDP_REQUIRE((size_t)planes * n_maps <= 65535, "too many planes for one launch");
  return CDP_OK;
}

extern "C" int cdp_flow_smooth_fwd(const float* const* maps, int32_t n_maps, int32_t planes, int32_t height,
                                   int32_t width, int32_t wrap_around, float* loss, float* unit_grad, void* scratch,
                                   size_t scratch_bytes, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CdpFlowParams p;
  const int rc = cdp_flow_common(maps, n_maps, planes, height, width, wrap_around, false, loss, unit_grad, scratch,
                                 scratch_bytes, &p);
  if (rc != CDP_OK) return rc;
  CDP_REQUIRE(p.blocks_y <= 65535, "image too tall for one launch");
  dim3 grid(p.blocks_x, p.blocks_y, planes * n_maps);
  cdp_flow_smooth_kernel<<<grid, CDP_FLOW_THREADS, 0, stream>>>(p);
  CDP_LAUNCH_CHECK("cdp_flow_smooth_kernel");
  const size_t count = (size_t)grid.x * grid.y * grid.z;
  cdp_sum_records_kernel<<<1, 1024, 0, stream>>>(p.part, count, (double)p.inv_count, loss);
  CDP_LAUNCH_CHECK("cdp_sum_records_kernel");
  return CDP_OK;
}

extern "C" int cdp_flow_sparsity_fwd(const float* const* maps, int32_t n_maps, int32_t planes, int32_t height,
                                     int32_t width, float* loss, float* unit_grad, void* scratch,
                                     size_t scratch_bytes, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CdpFlowParams p;
  const int rc = cdp_flow_common(maps, n_maps, planes, height, width, 1, true, loss, unit_grad, scratch,
                                 scratch_bytes, &p);
  if (rc != CDP_OK) return rc;
  dim3 grid(p.blocks_x, planes * n_maps);
  cdp_flow_abs_kernel<<<grid, CDP_FLOW_THREADS, 0, stream>>>(p);
  CDP_LAUNCH_CHECK("cdp_flow_abs_kernel");
  cdp_flow_sparsity_kernel<<<grid, CDP_FLOW_THREADS, 0, stream>>>(p);
  CDP_LAUNCH_CHECK("cdp_flow_sparsity_kernel");
  const size_t count = (size_t)grid.x * grid.y;
  cdp_sum_records_kernel<<<1, 1024, 0, stream>>>(p.part + count, count, (double)p.inv_count, loss);
  CDP_LAUNCH_CHECK("cdp_sum_records_kernel");
  return CDP_OK;
}

extern "C" int cdp_scale_fwd(const float* in, const float* scalar, size_t count, float* out, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(in && scalar && out && count > 0, "null pointer or empty");
  CDP_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "buffers must be 16-byte aligned");
  cdp_scale_kernel<<<cdp_elementwise_grid((count + 3) / 4), 256, 0, stream>>>(in, scalar, count, out);
  CDP_LAUNCH_CHECK("cdp_scale_kernel");
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// camera-to-camera warp (cdp_c2c.h): Mixup.warp_c2c
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cdp_warp_c2c_kernel(const __grid_constant__ CdpC2cParams p) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix < p.Ht * p.Wt) cdp_c2c_pixel<T>(p, blockIdx.y, pix);
}

extern "C" int cdp_warp_c2c_fwd(const void* src, int32_t src_is_f64, int32_t batch, int32_t channels, int32_t src_height,
                                int32_t src_width, int32_t out_height, int32_t out_width, const double* K_src,
                                const double* K_tgt, double depth_val, int32_t nearest, int32_t padding_zeros,
                                double* out, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(batch > 0 && channels > 0 && src_height > 0 && src_width > 0 && out_height > 0 && out_width > 0,
              "invalid shape");
  CDP_REQUIRE((size_t)out_height * out_width < (1ull << 31) && (size_t)src_height * src_width < (1ull << 31),
              "image too large");
  CDP_REQUIRE(src && K_src && K_tgt && out, "null pointer");
  for (int b0 = 0; b0 < batch; b0 += CDP_MAX_BATCH_PER_LAUNCH) {
    const int nb = cdp_chunk_size(batch, b0);
    CdpC2cParams p;
    cdp_fill_c2c_params(&p, src, out, K_src, K_tgt, b0, nb, channels, src_height, src_width, out_height, out_width,
                        depth_val, nearest != 0, padding_zeros != 0);
    dim3 grid((out_height * out_width + 255) / 256, nb);
    if (src_is_f64) cdp_warp_c2c_kernel<double><<<grid, 256, 0, stream>>>(p);
    else cdp_warp_c2c_kernel<float><<<grid, 256, 0, stream>>>(p);
    CDP_LAUNCH_CHECK("cdp_warp_c2c_kernel");
  }
  return CDP_OK;
}

// ------------------------------------------------------------------------------------------
// depth metrics (cdp_metrics.h): DepthEvaluator.compute_depth_metrics
// ------------------------------------------------------------------------------------------
// warp-parallel form of cdp_metrics_scan: lane l owns bins [8l, 8l+8); all lanes return the result
__device__ __forceinline__ void cdp_metrics_scan_warp(const uint32_t* hist256, uint32_t rank, int lane, uint32_t& bin,
                                                      uint32_t& rank_out, uint32_t& total) {
  uint32_t c[8], mine = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = hist256[lane * 8 + j]; mine += c[j]; }
  uint32_t incl = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  total = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t excl = incl - mine;
  const bool here = rank >= excl && rank < incl;
  uint32_t b = 255, r = 0;
  if (here) {
    uint32_t acc = excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (rank >= acc && rank < acc + c[j]) { b = lane * 8 + j; r = rank - acc; }
      acc += c[j];
    }
  }
  const uint32_t who = __ballot_sync(0xffffffffu, here);
  const int src = who ? __ffs(who) - 1 : 0;
  bin = __shfl_sync(0xffffffffu, b, src);
  rank_out = __shfl_sync(0xffffffffu, r, src);
  if (!who) { bin = 255; rank_out = 0; }
}

__device__ __forceinline__ void cdp_metrics_state_warp(const CdpMetricsParams& p, int unit, int arr, int passes, int lane,
                                                       uint32_t& prefix, uint32_t& rank, uint32_t& count) {
  uint32_t bin, r, total;
  prefix = 0; rank = 0; count = 0;
  for (int q = 0; q < passes; ++q) {
    const uint32_t* h = cdp_metrics_hist(p, q, unit, arr);
    if (q == 0) {
      cdp_metrics_scan_warp(h, 0, lane, bin, r, total);  // total = number of valid elements
      count = total;
      rank = count ? (count - 1) / 2 : 0;
    }
    cdp_metrics_scan_warp(h, rank, lane, bin, r, total);
    prefix |= bin << (24 - 8 * q);
    rank = r;
  }
}

template <int PASS>
__global__ void __launch_bounds__(CDP_METRICS_THREADS) cdp_metrics_hist_kernel(const __grid_constant__ CdpMetricsParams p) {
  __shared__ uint32_t h[CDP_METRICS_THREADS / 32][2][256];  // one private histogram pair per warp
  __shared__ uint32_t prefix_s[2];
  const int unit = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (CDP_METRICS_THREADS / 32) * 2 * 256; i += blockDim.x) (&h[0][0][0])[i] = 0;
  if (PASS > 0 && warp < 2) {
    uint32_t prefix, rank, count;
    cdp_metrics_state_warp(p, unit, warp, PASS, lane, prefix, rank, count);
    if (lane == 0) prefix_s[warp] = prefix;
  }
  __syncthreads();
  const uint32_t pre_g = PASS > 0 ? prefix_s[0] : 0, pre_p = PASS > 0 ? prefix_s[1] : 0;
  const int base = blockIdx.x * CDP_METRICS_CHUNK;
#pragma unroll 4
  for (int k = 0; k < CDP_METRICS_PER_THREAD; ++k) {
    const int i = base + k * CDP_METRICS_THREADS + threadIdx.x;
    float g;
    if (i < p.n && cdp_metrics_valid(p, unit, i, g)) {
      const uint32_t kg = cdp_metrics_key(g), kp = cdp_metrics_key(__ldg(p.pred + (size_t)unit * p.n + i));
      if (cdp_metrics_match(kg, pre_g, PASS)) atomicAdd(&h[warp][0][(kg >> (24 - 8 * PASS)) & 255u], 1u);
      if (cdp_metrics_match(kp, pre_p, PASS)) atomicAdd(&h[warp][1][(kp >> (24 - 8 * PASS)) & 255u], 1u);
    }
  }
  __syncthreads();
  uint32_t* out = p.hist + (((size_t)PASS * p.units + unit) * 2) * 256;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < CDP_METRICS_THREADS / 32; ++w) s += (&h[w][0][0])[i];
    if (s) atomicAdd(out + i, s);  // integer counters: exact and order-independent
  }
}

__global__ void __launch_bounds__(CDP_METRICS_THREADS) cdp_metrics_stats_kernel(const __grid_constant__ CdpMetricsParams p) {
  __shared__ float red[(CDP_METRICS_THREADS / 32) * CDP_METRICS_NSTATS];
  __shared__ float med_s[2];
  const int unit = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 2) {
    uint32_t prefix, rank, count;
    cdp_metrics_state_warp(p, unit, warp, CDP_METRICS_PASSES, lane, prefix, rank, count);
    if (lane == 0) med_s[warp] = cdp_metrics_unkey(prefix);
  }
  __syncthreads();
  const float ratio = p.use_gt_scale ? med_s[0] / med_s[1] : 1.0f;  // gt.median() / pred.median()
  float acc[CDP_METRICS_NSTATS];
#pragma unroll
  for (int j = 0; j < CDP_METRICS_NSTATS; ++j) acc[j] = 0.f;
  const int base = blockIdx.x * CDP_METRICS_CHUNK;
#pragma unroll 2
  for (int k = 0; k < CDP_METRICS_PER_THREAD; ++k) {
    const int i = base + k * CDP_METRICS_THREADS + threadIdx.x;
    float g;
    if (i < p.n && cdp_metrics_valid(p, unit, i, g))
      cdp_metrics_element(p, g, __ldg(p.pred + (size_t)unit * p.n + i), ratio, acc);
  }
  float* rec = p.part + ((size_t)unit * p.blocks + blockIdx.x) * CDP_METRICS_REC;
  cdp_block_reduce_store(acc, red, rec);
}

// one warp per unit (fixed order), thread 0 averages the units in index order
__global__ void __launch_bounds__(1024) cdp_metrics_finalize_kernel(const __grid_constant__ CdpMetricsParams p) {
  __shared__ double unit_stats[32][CDP_METRICS_NSTATS];
  __shared__ int unit_ok[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  double total[CDP_METRICS_NSTATS];
  for (int j = 0; j < CDP_METRICS_NSTATS; ++j) total[j] = 0.0;
  int with_gt = 0;
  for (int u0 = 0; u0 < p.units; u0 += nwarps) {
    const int u = u0 + warp;
    if (u < p.units) {
      double sums[CDP_METRICS_NSTATS];
      for (int j = 0; j < CDP_METRICS_NSTATS; ++j)
        sums[j] = cdp_warp_butterfly(cdp_lane_sum(p.part + (size_t)u * p.blocks * CDP_METRICS_REC + j, p.blocks,
                                                  CDP_METRICS_REC, lane));
      uint32_t bin, r, count;
      cdp_metrics_scan_warp(cdp_metrics_hist(p, 0, u, 0), 0, lane, bin, r, count);
      if (lane == 0) {
        double out[CDP_METRICS_NSTATS];
        unit_ok[warp] = cdp_metrics_unit_stats(sums, count, out) ? 1 : 0;
        for (int j = 0; j < CDP_METRICS_NSTATS; ++j) unit_stats[warp][j] = out[j];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int w = 0; w < nwarps && u0 + w < p.units; ++w) {
        with_gt += unit_ok[w];
        for (int j = 0; j < CDP_METRICS_NSTATS; ++j) total[j] += unit_ok[w] ? unit_stats[w][j] : (double)NAN;
      }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int j = 0; j < CDP_METRICS_NSTATS; ++j) p.out[j] = (float)(total[j] / (double)p.units);
    p.out[CDP_METRICS_NSTATS] = (float)with_gt;
  }
}

extern "C" size_t cdp_depth_metrics_scratch_bytes(int32_t units, int32_t count) {
  return cdp_metrics_scratch_total(units, count);
}

extern "C" int cdp_depth_metrics_fwd(const float* depth_gt, const float* depth_pred, const int64_t* labels,
                                     int64_t class_id, int32_t units, int32_t count, int32_t height, int32_t width,
                                     int32_t garg_crop, float min_depth, float max_depth, int32_t use_gt_scale,
                                     float* out, void* scratch, size_t scratch_bytes, cdp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDP_REQUIRE(depth_gt && depth_pred && out && scratch, "null pointer");
  CDP_REQUIRE(units >= 1 && units <= 65535 && count >= 1, "invalid shape: %d units of %d elements", units, count);
  CdpMetricsParams p;
  CDP_REQUIRE(cdp_fill_metrics_params(&p, depth_gt, depth_pred, labels, class_id, units, count, width, height, garg_crop,
                                      min_depth, max_depth, use_gt_scale, scratch, out),
              "invalid arguments (depth range, or crop without height*width == count)");
  const size_t need = cdp_metrics_scratch_total(units, count);
  if (scratch_bytes < need) return cdp_fail(CDP_ERR_WORKSPACE, "scratch too small: %zu < %zu", scratch_bytes, need);
  CDP_CUDA(cudaMemsetAsync(p.hist, 0, cdp_metrics_hist_bytes(units), stream));
  dim3 grid(p.blocks, units);
  cdp_metrics_hist_kernel<0><<<grid, CDP_METRICS_THREADS, 0, stream>>>(p);
  cdp_metrics_hist_kernel<1><<<grid, CDP_METRICS_THREADS, 0, stream>>>(p);
  cdp_metrics_hist_kernel<2><<<grid, CDP_METRICS_THREADS, 0, stream>>>(p);
  cdp_metrics_hist_kernel<3><<<grid, CDP_METRICS_THREADS, 0, stream>>>(p);
  CDP_LAUNCH_CHECK("cdp_metrics_hist_kernel");
  cdp_metrics_stats_kernel<<<grid, CDP_METRICS_THREADS, 0, stream>>>(p);
  CDP_LAUNCH_CHECK("cdp_metrics_stats_kernel");
  cdp_metrics_finalize_kernel<<<1, 1024, 0, stream>>>(p);
  CDP_LAUNCH_CHECK("cdp_metrics_finalize_kernel");
  return CDP_OK;
}
