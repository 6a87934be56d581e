// developer probe 2: which flavour of the bulk-copy engine works on this box?  ./tma_probe2 <test>
//   1: cp.async.bulk 1-D global->shared (UBLKCP)        2: libcu++ cuda::memcpy_async + cuda::barrier (compiler-chosen path)
//   3: tensor 2-D map via inline PTX                     4: tensor 3-D via cuda::device::experimental API
//   5: tensor 3-D inline PTX launched with an explicit 1x1x1 cluster
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier_t = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
               ::"r"(smem_u32(bar)), "r"(0u) : "memory");
}

__global__ void k_bulk1d(const uint8_t* src, uint8_t* out, int bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(&bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(src), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
  }
  wait0(&bar);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

__global__ void k_libcu(const uint8_t* src, uint8_t* out, int bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ barrier_t bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier_t::arrival_token tok;
  if (threadIdx.x == 0) {
    cuda::memcpy_async(smem, src, cuda::aligned_size_t<16>(bytes), bar);
    tok = bar.arrive();
  } else tok = bar.arrive();
  bar.wait(std::move(tok));
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

__global__ void k_tensor2d(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(&bar)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem)), "l"((uint64_t)&tmap), "r"(smem_u32(&bar)), "r"(0), "r"(0) : "memory");
  }
  wait0(&bar);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

__global__ void k_tensor3d_api(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ barrier_t bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier_t::arrival_token tok;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_3d_global_to_shared(smem, &tmap, 0, 0, 0, bar);
    tok = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else tok = bar.arrive();
  bar.wait(std::move(tok));
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

__global__ void k_tensor3d_ptx(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(&bar)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem)), "l"((uint64_t)&tmap), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"(0) : "memory");
  }
  wait0(&bar);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 1;
  const int X = 256, Y = 40, Z = 12;
  std::vector<uint8_t> h((size_t)X * Y * Z);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(1 + i % 31);
  uint8_t *d, *dout;
  cudaMalloc(&d, h.size());
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 8192);
  cudaMemset(dout, 0, 8192);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  int bytes = 0;
  std::vector<uint8_t> want;
  if (test == 1 || test == 2) {
    bytes = 96;
    for (int i = 0; i < bytes; i++) want.push_back(h[i]);
    if (test == 1) k_bulk1d<<<1, 128, 256>>>(d, dout, bytes); else k_libcu<<<1, 128, 256>>>(d, dout, bytes);
  } else if (test == 3) {
    const cuuint64_t gdim[2] = {(cuuint64_t)X, (cuuint64_t)Y * Z};
    const cuuint64_t gstr[1] = {(cuuint64_t)X};
    const cuuint32_t bdim[2] = {64, 8}, estr[2] = {1, 1};
    CUresult r = ((EncodeTiled)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode2d %d\n", (int)r);
    bytes = 64 * 8;
    for (int y = 0; y < 8; y++) for (int x = 0; x < 64; x++) want.push_back(h[(size_t)y * X + x]);
    k_tensor2d<<<1, 128, 1024>>>(tmap, dout, bytes);
  } else {
    const cuuint64_t gdim[3] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z};
    const cuuint64_t gstr[2] = {(cuuint64_t)X, (cuuint64_t)X * Y};
    const cuuint32_t bdim[3] = {64, 4, 2}, estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiled)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode3d %d\n", (int)r);
    bytes = 64 * 4 * 2;
    for (int z = 0; z < 2; z++) for (int y = 0; y < 4; y++) for (int x = 0; x < 64; x++) want.push_back(h[((size_t)z * Y + y) * X + x]);
    if (test == 4) k_tensor3d_api<<<1, 128, 1024>>>(tmap, dout, bytes);
    else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_tensor3d_ptx, tmap, dout, bytes);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("test %d: %s\n", test, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint8_t> o(bytes);
  cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < bytes; i++) bad += o[i] != want[i];
  printf("test %d mismatches %d of %d\n", test, bad, bytes);
  return 0;
}
