// Fast fused voxel-based kernels for the headline configuration (kernelRadius 1, 3-D,
// distances [1], 8-bit levels).  One thread per centre voxel, consecutive threads = consecutive
// x so the 24 float64 map stores of a warp are 256-byte coalesced segments.
#include <map>
#include <mutex>

#include "common.cuh"
#include "glcm_fast.cuh"
#include "host_common.hpp"

namespace rb {

constexpr int GF_THREADS = 128;

// Three phases per tile of GF_THREADS consecutive voxels (block-uniform loop):
//   A  every thread: its voxel's window -> equality masks -> all 13 angles, everything except the
//      MCC eigen-solves, which are queued as (owner thread, angle slot) tasks in shared memory;
//   B  the block drains the queue with ALL lanes busy (any thread can rebuild any task from the
//      owner's window + masks in shared memory) -- eigen-solves are needed by only a few % of the
//      (voxel, angle) pairs on noisy data but by most on smooth data, so leaving them inline would
//      idle most lanes of a warp behind one long solve;
//   C  every thread adds its solved tasks (in slot order: deterministic) and stores 24 coalesced
//      float64 map values.
__global__ void __launch_bounds__(GF_THREADS)
glcm_fast_kernel(const uint8_t* __restrict__ lev, const uint8_t* __restrict__ centers,
                 const __grid_constant__ VoxParams P, const GlcmFastTables* __restrict__ Tg,
                 double* __restrict__ out, long long fstride, int z0, int z1, int out_z0) {
  __shared__ GlcmFastTables T;
  __shared__ uint8_t wbuf[27 * GF_THREADS];
  __shared__ uint32_t eqbuf[27 * GF_THREADS];
  __shared__ double solved[GF_NA * GF_THREADS];
  __shared__ uint16_t queue[GF_NA * GF_THREADS];
  __shared__ int qn;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(Tg);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
    for (int i = threadIdx.x; i < (int)(sizeof(GlcmFastTables) / 4); i += GF_THREADS) dst[i] = src[i];
  }
  const int tid = threadIdx.x;
  const long long plane = (long long)P.Y * P.X;
  const long long total = (long long)(z1 - z0) * plane;
  const long long ntiles = (total + GF_THREADS - 1) / GF_THREADS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid == 0) qn = 0;
    __syncthreads();                       // also covers the table copy on the first pass
    const long long t = tile * GF_THREADS + tid;
    const bool live = t < total;
    int z = 0, rem = 0;
    long long oi = 0;
    bool is_center = false;
    double f[GLCM_NF];
    uint32_t tasks = 0;
    int n_ok = 0;
    if (live) {
      z = z0 + (int)(t / plane);
      rem = (int)(t % plane);
      const int y = rem / P.X, x = rem % P.X;
      const long long vi = (long long)z * P.sz + (long long)y * P.sy + x;
      oi = (long long)(z - out_z0) * plane + rem;
      is_center = centers ? centers[(long long)z * plane + rem] != 0 : lev[vi] != 0;
      if (is_center) {
        uint8_t* w = &wbuf[tid];
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
          for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
              const int zz = z + dz, yy = y + dy, xx = x + dx;
              const bool in = zz >= 0 && zz < P.Z && yy >= 0 && yy < P.Y && xx >= 0 && xx < P.X;
              w[((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)) * GF_THREADS] =
                  in ? lev[vi + (long long)dz * P.sz + (long long)dy * P.sy + dx] : (uint8_t)0;
            }
        tasks = glcm_fast_voxel_phaseA(w, GF_THREADS, &eqbuf[tid], GF_THREADS, T, P, f, &n_ok);
        for (uint32_t m = tasks; m; m &= m - 1) {
          const int s = __ffs((int)m) - 1;
          queue[atomicAdd(&qn, 1)] = (uint16_t)(tid | s << 8);
        }
      }
    }
    __syncthreads();
    const int nq = qn;
    for (int k = tid; k < nq; k += GF_THREADS) {
      const int owner = queue[k] & 0xFF, s = queue[k] >> 8;
      solved[s * GF_THREADS + owner] = glcm_fast_solve_task(&wbuf[owner], GF_THREADS, &eqbuf[owner], GF_THREADS, T, s);
    }
    __syncthreads();
    if (live) {
      if (is_center) {
        f[G_MCC] = glcm_fast_finish_mcc(f[G_MCC], n_ok, tasks, &solved[tid], GF_THREADS);
#pragma unroll
        for (int k = 0; k < GLCM_NF; k++) out[k * fstride + oi] = f[k];
      } else {
#pragma unroll
        for (int k = 0; k < GLCM_NF; k++) out[k * fstride + oi] = P.init_value;
      }
    }
  }
}

// device-resident table cache, one per (device, Ng)
static const GlcmFastTables* glcm_fast_tables_dev(int Ng) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, GlcmFastTables*> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find({dev, Ng});
  if (it != cache.end()) return it->second;
  GlcmFastTables* h = new GlcmFastTables;
  memset(h, 0, sizeof *h);
  glcm_fast_build_tables(*h, Ng);
  GlcmFastTables* d = nullptr;
  if (cudaMalloc(&d, sizeof *h) != cudaSuccess || cudaMemcpy(d, h, sizeof *h, cudaMemcpyHostToDevice) != cudaSuccess) {
    delete h;
    return nullptr;
  }
  delete h;
  cache[{dev, Ng}] = d;
  return d;
}

bool glcm_fast_applicable(int cls, int level_bytes, const VoxParams& P) {
  return cls == C_GLCM && level_bytes == 1 && P.rz == 1 && P.ry == 1 && P.rx == 1 && P.na == 13 && P.symmetric &&
         !P.weighted && P.Ng <= 255;
}

int glcm_fast_launch(const void* lev, const uint8_t* centers, const VoxParams& P, double* out, long long fstride,
                     int z0, int z1, int out_z0, cudaStream_t st) {
  const GlcmFastTables* T = glcm_fast_tables_dev(P.Ng);
  if (!T) return fail(RB_ERR_CUDA, "could not build the GLCM table block on the device");
  const long long total = (long long)(z1 - z0) * P.Y * P.X;
  if (total <= 0) return RB_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long need = (total + GF_THREADS - 1) / GF_THREADS, cap = (long long)sms * 16;
  const int grid = (int)(need < cap ? need : cap);
  glcm_fast_kernel<<<grid, GF_THREADS, 0, st>>>((const uint8_t*)lev, centers, P, T, out, fstride, z0, z1, out_z0);
  RB_LAUNCH_CHECK();
  return RB_OK;
}

}  // namespace rb
