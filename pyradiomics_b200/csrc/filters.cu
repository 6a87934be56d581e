// Pre-filter and discretisation kernels feeding the texture path (a14-a16):
//   * gray-level discretisation: ROI min/max reduction + np.digitize-exact binning
//     (reference radiomics/imageoperations.py:67-174)
//   * level-1 stationary wavelet transform, periodic, one axis per pass, low+high band per pass
//     (reference radiomics/imageoperations.py:899-970 -> pywt.swtn(level=1), PyWavelets >= 1.6)
//   * Laplacian of Gaussian by recursive (IIR) Gaussian filtering, one line per thread
//     (reference radiomics/imageoperations.py:756-836 -> ITK LaplacianRecursiveGaussianImageFilter)
// All are HBM-streaming kernels: every thread handles consecutive x so loads/stores coalesce; the
// wavelet pass reads its 6 taps through L1 (neighbouring threads share them).
#include "common.cuh"

namespace rb {

enum DType { DT_I16 = 0, DT_I32 = 1, DT_F32 = 2, DT_F64 = 3, DT_U8 = 4, DT_U16 = 5, DT_I64 = 6 };

template <typename T> __device__ __forceinline__ double as_f64(const void* p, long long i) { return (double)((const T*)p)[i]; }
__device__ __forceinline__ double load_any(const void* p, int dt, long long i) {
  switch (dt) {
    case DT_I16: return as_f64<int16_t>(p, i);
    case DT_I32: return as_f64<int32_t>(p, i);
    case DT_F32: return as_f64<float>(p, i);
    case DT_F64: return as_f64<double>(p, i);
    case DT_U8: return as_f64<uint8_t>(p, i);
    case DT_U16: return as_f64<uint16_t>(p, i);
    default: return as_f64<long long>(p, i);
  }
}

// order-preserving map double <-> signed 64-bit, so atomicMin/atomicMax work on doubles
__device__ __forceinline__ long long f64_key(double v) {
  long long b = __double_as_longlong(v);
  return b >= 0 ? b : b ^ 0x7FFFFFFFFFFFFFFFll;
}

__global__ void __launch_bounds__(256)
minmax_kernel(const void* __restrict__ img, int dt, const uint8_t* __restrict__ mask, long long n,
              long long* __restrict__ keys /* [0]=min key, [1]=max key, [2]=count */) {
  double lo = 1.0 / 0.0, hi = -1.0 / 0.0;
  long long cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;
    const double v = load_any(img, dt, i);
    lo = v < lo ? v : lo; hi = v > hi ? v : hi; cnt++;
  }
  for (int o = 16; o; o >>= 1) {
    const double l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicMin(&keys[0], f64_key(lo));
    atomicMax(&keys[1], f64_key(hi));
    atomicAdd((unsigned long long*)&keys[2], (unsigned long long)cnt);
  }
}

// out[i] = #{k : edges[k] <= x}  (np.digitize with increasing bins, right=False), 0 outside the mask
__global__ void __launch_bounds__(256)
digitize_kernel(const void* __restrict__ img, int dt, const uint8_t* __restrict__ mask, long long n,
                const double* __restrict__ edges, int ne, int32_t* __restrict__ out) {
  extern __shared__ double s_edges[];
  const bool use_s = ne <= 4096;
  if (use_s) { for (int k = threadIdx.x; k < ne; k += blockDim.x) s_edges[k] = edges[k]; __syncthreads(); }
  const double* e = use_s ? s_edges : edges;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int32_t b = 0;
    if (!mask || mask[i]) {
      const double x = load_any(img, dt, i);
      int lo = 0, hi = ne;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (e[mid] <= x) lo = mid + 1; else hi = mid; }
      b = lo;
    }
    out[i] = b;
  }
}

// ---- stationary wavelet, one axis: lo/hi[n] = sum_j f[j] * x[(n + F/2 - j) mod Np], Np = N
// rounded up to even; the pad sample (index N when N is odd) is a copy of x[0] ("wrap" padding,
// imageoperations.py:914-919) and the padded output sample is never stored (cropped, :947-951).
struct SwtFilters { int F; double lo[24], hi[24]; };

__global__ void __launch_bounds__(256)
swt_axis_kernel(const double* __restrict__ in, int Z, int Y, int X, int axis, const __grid_constant__ SwtFilters W,
                double* __restrict__ out_lo, double* __restrict__ out_hi) {
  const long long n = (long long)Z * Y * X, plane = (long long)Y * X;
  const int N = axis == 0 ? Z : axis == 1 ? Y : X;
  const int Np = N + (N & 1);
  const long long stride = axis == 0 ? plane : axis == 1 ? X : 1;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(t / plane), rem = (int)(t % plane), y = rem / X, x = rem % X;
    const int c = axis == 0 ? z : axis == 1 ? y : x;
    const long long base = t - (long long)c * stride;
    double a = 0, d = 0;
    for (int j = 0; j < W.F; j++) {
      int k = (c + W.F / 2 - j) % Np;
      if (k < 0) k += Np;
      if (k >= N) k = 0;                       // the wrap-padding sample
      const double v = in[base + (long long)k * stride];
      a += W.lo[j] * v; d += W.hi[j] * v;
    }
    out_lo[t] = a; out_hi[t] = d;
  }
}

// ---- recursive Gaussian (Deriche 4th order, as used by ITK's RecursiveGaussianImageFilter), one
// line per thread along `axis`; causal + anti-causal passes summed.  Coefficients are computed on
// the host (rg_coefficients).  `scratch` holds the causal pass (same shape as the volume).
struct RGCoef { double N0, N1, N2, N3, D1, D2, D3, D4, M1, M2, M3, M4, BN1, BN2, BN3, BN4, BM1, BM2, BM3, BM4; };

template <typename TIn>
__global__ void __launch_bounds__(128)
recursive_gauss_axis_kernel(const TIn* __restrict__ in, int Z, int Y, int X, int axis, const __grid_constant__ RGCoef C,
                            float* __restrict__ out, double* __restrict__ scratch, double scale, int accumulate) {
  const int N = axis == 0 ? Z : axis == 1 ? Y : X;
  const long long plane = (long long)Y * X;
  const long long stride = axis == 0 ? plane : axis == 1 ? X : 1;
  const long long nlines = (long long)Z * Y * X / N;
  for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (long long)gridDim.x * blockDim.x) {
    long long base;
    if (axis == 2) base = l * X;
    else if (axis == 1) { const long long z = l / X, x = l % X; base = z * plane + x; }
    else base = l;
    // causal pass; boundary: the edge value is assumed to extend to infinity
    const double v0 = (double)in[base];
    double x1 = v0, x2 = v0, x3 = v0;
    const double sN = C.N0 + C.N1 + C.N2 + C.N3, sD = 1.0 + C.D1 + C.D2 + C.D3 + C.D4;
    double y1 = v0 * sN / sD, y2 = y1, y3 = y1, y4 = y1;
    for (int i = 0; i < N; i++) {
      const double xi = (double)in[base + (long long)i * stride];
      const double y = C.N0 * xi + C.N1 * x1 + C.N2 * x2 + C.N3 * x3 - C.D1 * y1 - C.D2 * y2 - C.D3 * y3 - C.D4 * y4;
      scratch[base + (long long)i * stride] = y;
      x3 = x2; x2 = x1; x1 = xi; y4 = y3; y3 = y2; y2 = y1; y1 = y;
    }
    // anti-causal pass
    const double vN = (double)in[base + (long long)(N - 1) * stride];
    double a1 = vN, a2 = vN, a3 = vN, a4 = vN;
    const double sM = C.M1 + C.M2 + C.M3 + C.M4;
    double b1 = vN * sM / sD, b2 = b1, b3 = b1, b4 = b1;
    for (int i = N - 1; i >= 0; i--) {
      const double xi = (double)in[base + (long long)i * stride];
      const double y = C.M1 * a1 + C.M2 * a2 + C.M3 * a3 + C.M4 * a4 - C.D1 * b1 - C.D2 * b2 - C.D3 * b3 - C.D4 * b4;
      const long long o = base + (long long)i * stride;
      const float r = (float)((scratch[o] + y) * scale);
      out[o] = accumulate ? out[o] + r : r;
      a4 = a3; a3 = a2; a2 = a1; a1 = xi; b4 = b3; b3 = b2; b2 = b1; b1 = y;
    }
  }
}

static int grid_n(long long n, int block, int per_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long need = (n + block - 1) / block, cap = (long long)sms * per_sm;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

int minmax_launch(const void* img, int dt, const uint8_t* mask, long long n, long long* keys, cudaStream_t st) {
  minmax_kernel<<<grid_n(n, 256, 8), 256, 0, st>>>(img, dt, mask, n, keys);
  RB_LAUNCH_CHECK();
  return RB_OK;
}
int digitize_launch(const void* img, int dt, const uint8_t* mask, long long n, const double* edges, int ne, int32_t* out,
                    cudaStream_t st) {
  const size_t sh = ne <= 4096 ? sizeof(double) * ne : 0;
  digitize_kernel<<<grid_n(n, 256, 8), 256, sh, st>>>(img, dt, mask, n, edges, ne, out);
  RB_LAUNCH_CHECK();
  return RB_OK;
}
int swt_axis_launch(const double* in, int Z, int Y, int X, int axis, const double* lo, const double* hi, int F,
                    double* out_lo, double* out_hi, cudaStream_t st) {
  if (F < 2 || F > 24) return fail(RB_ERR_UNSUPPORTED, "wavelet filter length %d outside 2..24", F);
  SwtFilters W;
  W.F = F;
  for (int j = 0; j < F; j++) { W.lo[j] = lo[j]; W.hi[j] = hi[j]; }
  swt_axis_kernel<<<grid_n((long long)Z * Y * X, 256, 8), 256, 0, st>>>(in, Z, Y, X, axis, W, out_lo, out_hi);
  RB_LAUNCH_CHECK();
  return RB_OK;
}
int recursive_gauss_launch(const void* in, int in_is_f32, int Z, int Y, int X, int axis, const double* coef20, float* out,
                           double* scratch, double scale, int accumulate, cudaStream_t st) {
  RGCoef C;
  memcpy(&C, coef20, sizeof C);
  const int N = axis == 0 ? Z : axis == 1 ? Y : X;
  const long long nlines = (long long)Z * Y * X / N;
  const int grid = grid_n(nlines, 128, 8);
  if (in_is_f32) recursive_gauss_axis_kernel<float><<<grid, 128, 0, st>>>((const float*)in, Z, Y, X, axis, C, out, scratch, scale, accumulate);
  else recursive_gauss_axis_kernel<double><<<grid, 128, 0, st>>>((const double*)in, Z, Y, X, axis, C, out, scratch, scale, accumulate);
  RB_LAUNCH_CHECK();
  return RB_OK;
}

}  // namespace rb
