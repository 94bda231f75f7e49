// Microbenchmark for the next-round idea recorded in DESIGN.md section 6: serve the 2x2 bilinear
// footprints of the warped sources through tex2Dgather (one TLD4 per channel and source) instead
// of four scalar LDG with their own 64-bit addresses.  Stand-alone; not part of the library.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu && ./gather_bench
//
// Workload: Cityscapes level 0, batch 8 (two sources x three channels per pixel), sample positions
// = pixel + a smooth displacement of a few pixels, border clamp as grid_sample(padding_mode=border).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int B = 8, H = 512, W = 1024;

__device__ __forceinline__ void sample_pos(int b, int x, int y, int k, float& ix, float& iy) {
  // smooth displacement field, different per source
  const float fx = 3.1f * __sinf(0.013f * x + 0.7f * k + b) + 1.3f * k;
  const float fy = 2.2f * __cosf(0.017f * y + 0.3f * k) - 0.6f;
  ix = fminf(fmaxf((float)x + fx, 0.f), (float)(W - 1));
  iy = fminf(fmaxf((float)y + fy, 0.f), (float)(H - 1));
}

__global__ void __launch_bounds__(256) gather_ldg(const float* __restrict__ src0, const float* __restrict__ src1,
                                                  float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const size_t plane = (size_t)H * W;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float ix, iy;
    sample_pos(b, x, y, k, ix, iy);
    const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float wx = ix - x0, wy = iy - y0;
    const int o00 = y0 * W + x0, o01 = y0 * W + x1, o10 = y1 * W + x0, o11 = y1 * W + x1;
    const float* s = (k == 0 ? src0 : src1) + (size_t)b * 3 * plane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = s + c * plane;
      const float nw = __ldg(p + o00), ne = __ldg(p + o01), sw = __ldg(p + o10), se = __ldg(p + o11);
      const float top = nw + wx * (ne - nw), bot = sw + wx * (se - sw);
      acc[c] += top + wy * (bot - top);
    }
  }
  out[((size_t)b * H + y) * W + x] = acc[0] + acc[1] + acc[2];
}

__global__ void __launch_bounds__(256) gather_tex(cudaTextureObject_t t0, cudaTextureObject_t t1, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float ix, iy;
    sample_pos(b, x, y, k, ix, iy);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float wx = ix - fx0, wy = iy - fy0;
    const cudaTextureObject_t t = k == 0 ? t0 : t1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // rows of plane (b, c) start at (b*3 + c) * H in the stacked image; the footprint of the four
      // texels around the corner (x0+1, y0+1) is (x0..x0+1, y0..y0+1), clamped at the edges
      const float4 g = tex2Dgather<float4>(t, fx0 + 1.0f, (float)((b * 3 + c) * H) + fy0 + 1.0f, 0);
      // gather order: x = (i0, j1), y = (i1, j1), z = (i1, j0), w = (i0, j0)
      const float nw = g.w, ne = g.z, sw = g.x, se = g.y;
      const float top = nw + wx * (ne - nw), bot = sw + wx * (se - sw);
      acc[c] += top + wy * (bot - top);
    }
  }
  out[((size_t)b * H + y) * W + x] = acc[0] + acc[1] + acc[2];
}

// channel-packed variant: sources stored as [B,H,W,4] (RGB + pad), one 16-byte load per tap
__global__ void __launch_bounds__(256) gather_rgba(const float4* __restrict__ src0, const float4* __restrict__ src1,
                                                   float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const size_t plane = (size_t)H * W;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float ix, iy;
    sample_pos(b, x, y, k, ix, iy);
    const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float wx = ix - x0, wy = iy - y0;
    const float4* s = (k == 0 ? src0 : src1) + (size_t)b * plane;
    const float4 nw = __ldg(s + y0 * W + x0), ne = __ldg(s + y0 * W + x1);
    const float4 sw = __ldg(s + y1 * W + x0), se = __ldg(s + y1 * W + x1);
    float top, bot;
    top = nw.x + wx * (ne.x - nw.x); bot = sw.x + wx * (se.x - sw.x); acc[0] += top + wy * (bot - top);
    top = nw.y + wx * (ne.y - nw.y); bot = sw.y + wx * (se.y - sw.y); acc[1] += top + wy * (bot - top);
    top = nw.z + wx * (ne.z - nw.z); bot = sw.z + wx * (se.z - sw.z); acc[2] += top + wy * (bot - top);
  }
  out[((size_t)b * H + y) * W + x] = acc[0] + acc[1] + acc[2];
}

__global__ void pack_rgba(const float* __restrict__ planar, float4* __restrict__ packed) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, plane = (size_t)H * W;
  if (i >= (size_t)B * plane) return;
  const size_t b = i / plane, p = i - b * plane;
  packed[i] = make_float4(planar[(b * 3 + 0) * plane + p], planar[(b * 3 + 1) * plane + p], planar[(b * 3 + 2) * plane + p], 0.f);
}

static cudaTextureObject_t make_tex(const float* dev) {
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypePitch2D;
  rd.res.pitch2D.devPtr = const_cast<float*>(dev);
  rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  rd.res.pitch2D.width = W;
  rd.res.pitch2D.height = (size_t)B * 3 * H;
  rd.res.pitch2D.pitchInBytes = (size_t)W * sizeof(float);
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t t = 0;
  CHECK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
  return t;
}

int main() {
  const size_t n = (size_t)B * 3 * H * W;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)((i * 2654435761u) % 1000) * 1e-3f;
  float *s0, *s1, *o_ldg, *o_tex;
  CHECK(cudaMalloc(&s0, n * 4)); CHECK(cudaMalloc(&s1, n * 4));
  CHECK(cudaMalloc(&o_ldg, (size_t)B * H * W * 4)); CHECK(cudaMalloc(&o_tex, (size_t)B * H * W * 4));
  CHECK(cudaMemcpy(s0, h.data(), n * 4, cudaMemcpyHostToDevice));
  for (size_t i = 0; i < n; ++i) h[i] = 1.0f - h[i];
  CHECK(cudaMemcpy(s1, h.data(), n * 4, cudaMemcpyHostToDevice));
  const cudaTextureObject_t t0 = make_tex(s0), t1 = make_tex(s1);
  dim3 grid((W + 255) / 256, H, B);
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  float ms_ldg = 0.f, ms_tex = 0.f;
  for (int rep = 0; rep < 3; ++rep) {
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) gather_ldg<<<grid, 256>>>(s0, s1, o_ldg);
    CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms_ldg, e0, e1));
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) gather_tex<<<grid, 256>>>(t0, t1, o_tex);
    CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms_tex, e0, e1));
  }
  float4 *q0, *q1;
  float* o_rgba;
  CHECK(cudaMalloc(&q0, (size_t)B * H * W * 16)); CHECK(cudaMalloc(&q1, (size_t)B * H * W * 16));
  CHECK(cudaMalloc(&o_rgba, (size_t)B * H * W * 4));
  float ms_pack = 0.f, ms_rgba = 0.f;
  const int pack_blocks = (int)(((size_t)B * H * W + 255) / 256);
  for (int rep = 0; rep < 3; ++rep) {
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) { pack_rgba<<<pack_blocks, 256>>>(s0, q0); pack_rgba<<<pack_blocks, 256>>>(s1, q1); }
    CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms_pack, e0, e1));
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) gather_rgba<<<grid, 256>>>(q0, q1, o_rgba);
    CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms_rgba, e0, e1));
  }
  {
    std::vector<float> a2((size_t)B * H * W), c2((size_t)B * H * W);
    CHECK(cudaMemcpy(a2.data(), o_ldg, a2.size() * 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(c2.data(), o_rgba, c2.size() * 4, cudaMemcpyDeviceToHost));
    size_t differ2 = 0;
    for (size_t i = 0; i < a2.size(); ++i) differ2 += a2[i] != c2[i];
    printf("channel-packed [B,H,W,4] sources: gather %.1f us per launch (+ %.1f us to pack both sources once per step); "
           "outputs differ on %zu pixels\n", ms_rgba * 50.f, ms_pack * 50.f, differ2);
  }
  CHECK(cudaGetLastError());
  std::vector<float> a((size_t)B * H * W), b((size_t)B * H * W);
  CHECK(cudaMemcpy(a.data(), o_ldg, a.size() * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(b.data(), o_tex, b.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0.0;
  size_t differ = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    const double d = fabs((double)a[i] - (double)b[i]);
    if (d > worst) worst = d;
    if (a[i] != b[i]) ++differ;
  }
  printf("gather 2 sources x 3 channels, %dx%d batch %d: LDG %.1f us per launch, TLD4 %.1f us per launch; "
         "outputs differ on %zu of %zu pixels, worst |diff| %.3g\n", W, H, B, ms_ldg * 50.f, ms_tex * 50.f, differ, a.size(), worst);
  return 0;
}
