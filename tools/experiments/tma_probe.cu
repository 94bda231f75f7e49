// tma_probe.cu -- which way of handing a 3-D fp32 TMA descriptor to a kernel works on this
// driver / GPU (experiment behind the staging path of cdp_photo_kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Maps { CUtensorMap m[6][4]; };
struct Params { float* out; int lvl; int x, y, z; char pad[700]; };

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT>
__device__ void body(const CUtensorMap* map, float* out, int x, int y, int z) {
  extern __shared__ __align__(128) float sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 44 * 44 * 3 + 32);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (VARIANT & 4) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sa(bar)) : "memory");
    } else {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(bar)), "r"(44 * 44 * z * 0 + (int)(44 * 44 * 4) * (int)gridDim.y) : "memory");
    if (VARIANT & 1)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(sa(sm)), "l"((uint64_t)map), "r"(sa(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(sa(sm)), "l"((uint64_t)map), "r"(sa(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(sa(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 44 * 44 * (int)gridDim.y; i += blockDim.x) out[i] = sm[i];
}

template <int VARIANT>
__global__ void k_direct(const __grid_constant__ CUtensorMap map, float* out, int x, int y, int z) { body<VARIANT>(&map, out, x, y, z); }
template <int VARIANT>
__global__ void k_struct(const __grid_constant__ Params p, const __grid_constant__ Maps tm) { body<VARIANT>(&tm.m[p.lvl][2], p.out, p.x, p.y, p.z); }

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int boxp = argc > 2 ? atoi(argv[2]) : 3;
  const int cx = argc > 3 ? atoi(argv[3]) : -6, cy = argc > 4 ? atoi(argv[4]) : 30, cz = argc > 5 ? atoi(argv[5]) : 3;
  const int W = 128, H = 72, P = 6;
  std::vector<float> h((size_t)W * H * P);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 44 * 44 * 3 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Maps tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[3] = {W, H, P}, str[2] = {W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {44, 44, (cuuint32_t)boxp}, es[3] = {1, 1, 1};
  CUresult r = ((Enc)sym)(&tm.m[1][2], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode=%d sizeof(Params)=%zu sizeof(Maps)=%zu\n", variant, (int)r, sizeof(Params), sizeof(Maps));
  const size_t smem = 44 * 44 * 3 * 4 + 256;
  const int x = cx, y = cy, z = cz;
  if (variant == 4) {
    cudaFuncSetAttribute(k_direct<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_direct<4><<<dim3(1, boxp), 256, smem>>>(tm.m[1][2], out, x, y, z);
  } else if (variant < 2) {
    auto k = variant == 0 ? k_direct<0> : k_direct<1>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<dim3(1, boxp), 256, smem>>>(tm.m[1][2], out, x, y, z);
  } else {
    auto k = variant == 2 ? k_struct<0> : k_struct<1>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    Params p; memset(&p, 0, sizeof(p)); p.out = out; p.lvl = 1; p.x = x; p.y = y; p.z = z;
    k<<<dim3(1, boxp), 256, smem>>>(p, tm);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("  sync: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> o(44 * 44 * 3);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int pl = 0; pl < boxp; ++pl) for (int ty = 0; ty < 44; ++ty) for (int tx = 0; tx < 44; ++tx) {
      const int gx = x + tx, gy = y + ty, gp = z + pl;
      const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H && gp < P) ? h[((size_t)gp * H + gy) * W + gx] : 0.f;
      bad += o[(pl * 44 + ty) * 44 + tx] != want;
    }
    printf("  mismatches: %d\n", bad);
  }
  return 0;
}
