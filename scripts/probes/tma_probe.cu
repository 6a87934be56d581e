// developer probe: the TMA box load of csrc/segment_kernels.cu in isolation (variants selected by argv[1])
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu   ;   ./tma_probe [variant]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tmap, int bx, int by, int bz, int x0, int y0, int z0, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  const int box = bx * by * bz;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1), "r"(smem_u32(&bar)) : "memory");
    if (VARIANT != 2) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(box), "r"(smem_u32(&bar)) : "memory");
    if (VARIANT == 1)
      asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(smem)), "l"((uint64_t)&tmap), "r"(smem_u32(&bar)), "r"(x0), "r"(y0), "r"(z0) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(smem)), "l"((uint64_t)&tmap), "r"(smem_u32(&bar)), "r"(x0), "r"(y0), "r"(z0) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < box; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int X = argc > 2 ? atoi(argv[2]) : 256, Y = 40, Z = 12;
  std::vector<uint8_t> h((size_t)X * Y * Z);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(1 + i % 31);
  uint8_t *d, *dout;
  cudaMalloc(&d, h.size());
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  const int bx = argc > 3 ? atoi(argv[3]) : 80, by = argc > 4 ? atoi(argv[4]) : 10, bz = argc > 5 ? atoi(argv[5]) : 6;
  cudaMalloc(&dout, bx * by * bz);
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  const cuuint64_t gdim[3] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z};
  const cuuint64_t gstr[2] = {(cuuint64_t)X, (cuuint64_t)X * Y};
  const cuuint32_t bdim[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((EncodeTiled)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int x0 = argc > 6 ? atoi(argv[6]) : -1, y0 = argc > 7 ? atoi(argv[7]) : -1, z0 = argc > 8 ? atoi(argv[8]) : 7;   // default: box sticks out below (x, y) and above (z)
  printf("box %d x %d x %d at (%d, %d, %d)\n", bx, by, bz, x0, y0, z0);
  if (variant == 1) probe_kernel<1><<<1, 128, bx * by * bz + 1024>>>(tmap, bx, by, bz, x0, y0, z0, dout);
  else if (variant == 2) probe_kernel<2><<<1, 128, bx * by * bz + 1024>>>(tmap, bx, by, bz, x0, y0, z0, dout);
  else probe_kernel<0><<<1, 128, bx * by * bz + 1024>>>(tmap, bx, by, bz, x0, y0, z0, dout);
  e = cudaDeviceSynchronize();
  printf("variant %d kernel: %s\n", variant, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint8_t> o(bx * by * bz);
  cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int p = 0; p < bz; p++) for (int rr = 0; rr < by; rr++) for (int c = 0; c < bx; c++) {
    const int x = x0 + c, y = y0 + rr, z = z0 + p;
    const uint8_t want = (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z) ? h[((size_t)z * Y + y) * X + x] : 0;
    if (o[(p * by + rr) * bx + c] != want) bad++;
  }
  printf("variant %d mismatches: %d of %d\n", variant, bad, bx * by * bz);
  return bad != 0;
}
