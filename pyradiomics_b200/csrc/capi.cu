// extern "C" surface of libb200radiomics.so (declared in include/b200radiomics.h).
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "host_common.hpp"

namespace rb {

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

int voxel_features_generic(int cls, const void* lev, int level_bytes, const uint8_t* centers, const VoxParams& P,
                           double* out, long long fstride, int z0, int z1, int out_z0, int* status, cudaStream_t st);
int glcm_alive_angles(const void* lev, int level_bytes, const uint8_t* centers, const VoxParams& P, uint32_t* alive,
                      cudaStream_t st);
int pack_levels(const int32_t* image, const uint8_t* mask, long long n, int Ng, void* lev, uint32_t* presence,
                int* status, cudaStream_t st);
bool glcm_fast_applicable(int cls, int level_bytes, const VoxParams& P);
int glcm_release_queues();
int glcm_fast_launch(const void* lev, const uint8_t* centers, const VoxParams& P, double* out, long long fstride,
                     int z0, int z1, int out_z0, cudaStream_t st);

bool glrlm_fast_applicable(int cls, int level_bytes, const VoxParams& P);
int glrlm_fast_launch(const void* lev, const uint8_t* centers, const VoxParams& P, double* out, long long fstride,
                      int z0, int z1, int out_z0, cudaStream_t st);

bool small_fast_applicable(int cls, int level_bytes, const VoxParams& P);
int small_fast_launch(int cls, const void* lev, const uint8_t* centers, const VoxParams& P, double* out, long long fstride,
                      int z0, int z1, int out_z0, cudaStream_t st);

// B200_RADIOMICS_FORCE_GENERIC=1 routes everything through the generic kernels (used by the
// tests to cross-check the fast paths on the GPU)
static bool force_generic() {
  const char* e = getenv("B200_RADIOMICS_FORCE_GENERIC");
  return e && e[0] == '1';
}

int calculate_matrix_host(int mode, const int32_t* image, const uint8_t* mask, const int* size, int nd,
                          const int* distances, int ndist, int Ng, int Nr, int alpha, int force2D, int force2Ddimension,
                          int kernelRadius, const int* voxels, int nvox, double* out_host, int* angles_out, int* na_out,
                          const void* levels_dev = nullptr);
int glszm_zones_host(const int32_t* image, const uint8_t* mask, const int* size, int nd, int Ng, int force2D,
                     int force2Ddimension, int kernelRadius, const int* voxels, int nvox, int* max_region_out,
                     void** handle_out, const void* levels_dev = nullptr);
int segment_tile_matrices(const uint8_t* lev, int nd, int Z, int Y, int X, const int* distances, int ndist, int Ng, int alpha,
                          int force2D, int force2Ddimension, double* glcm_host, double* gldm_host, double* ngtdm_host,
                          int* angles_out, int* na_out, cudaStream_t st);
int glszm_fill_host(void* handle, int Ng, int max_region, double* out_host);
void glszm_release(void* handle);

int minmax_launch(const void* img, int dt, const uint8_t* mask, long long n, long long* keys, cudaStream_t st);
int shape_coefficients_dev(const uint8_t* mask_dev, int Z, int Y, int X, long long sz, long long sy, long long sx,
                           const double* spacing, double* out7, cudaStream_t st);
int shape_moments_dev(const uint8_t* mask_dev, int Z, int Y, int X, unsigned long long* out10, cudaStream_t st);
int shape2d_coefficients_dev(const uint8_t* mask_dev, int Y, int X, long long sy, long long sx, const double* spacing, double* out4,
                             cudaStream_t st);
int digitize_launch(const void* img, int dt, const uint8_t* mask, long long n, const double* edges, int ne, int32_t* out,
                    cudaStream_t st);
int swt_axis_launch(const double* in, int Z, int Y, int X, int axis, const double* lo, const double* hi, int F,
                    double* out_lo, double* out_hi, cudaStream_t st);
int recursive_gauss_launch(const void* in, int in_is_f32, int Z, int Y, int X, int axis, const double* coef20, float* out,
                           double* scratch, double scale, int accumulate, cudaStream_t st);
int swt3d_launch(const double* in, int Z, int Y, int X, const double* lo, const double* hi, int F, double* out,
                 long long band_stride, int z_begin, int z_end, cudaStream_t st);
int bspline_prefilter_launch(double* coeffs, int Z, int Y, int X, cudaStream_t st);
int resample_launch(const void* src, int src_dt, const int* in_size, void* dst, int dst_dt, const int* out_size, const double* start,
                    const double* step, int interp, double default_value, cudaStream_t st);

int firstorder_launch(const void* img, int dtype, const uint8_t* mask, const uint8_t* centers, const void* lev,
                      int level_bytes, int Z, int Y, int X, int rz, int ry, int rx, double shift, double voxel_volume,
                      double init_value, double* out, long long fstride, int z0, int z1, int out_z0, cudaStream_t st);
static const char* kFirstOrderNames[] = {"10Percentile", "90Percentile", "Energy", "Entropy", "InterquartileRange", "Kurtosis",
  "Maximum", "MeanAbsoluteDeviation", "Mean", "Median", "Minimum", "Range", "RobustMeanAbsoluteDeviation", "RootMeanSquared",
  "Skewness", "TotalEnergy", "Uniformity", "Variance"};

static const char* kGlcmNames[] = {"Autocorrelation", "ClusterProminence", "ClusterShade", "ClusterTendency", "Contrast",
  "Correlation", "DifferenceAverage", "DifferenceEntropy", "DifferenceVariance", "Id", "Idm", "Idmn", "Idn", "Imc1", "Imc2",
  "InverseVariance", "JointAverage", "JointEnergy", "JointEntropy", "MCC", "MaximumProbability", "SumAverage", "SumEntropy",
  "SumSquares"};
static const char* kGlrlmNames[] = {"GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
  "HighGrayLevelRunEmphasis", "LongRunEmphasis", "LongRunHighGrayLevelEmphasis", "LongRunLowGrayLevelEmphasis",
  "LowGrayLevelRunEmphasis", "RunEntropy", "RunLengthNonUniformity", "RunLengthNonUniformityNormalized", "RunPercentage",
  "RunVariance", "ShortRunEmphasis", "ShortRunHighGrayLevelEmphasis", "ShortRunLowGrayLevelEmphasis"};
static const char* kGlszmNames[] = {"GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
  "HighGrayLevelZoneEmphasis", "LargeAreaEmphasis", "LargeAreaHighGrayLevelEmphasis", "LargeAreaLowGrayLevelEmphasis",
  "LowGrayLevelZoneEmphasis", "SizeZoneNonUniformity", "SizeZoneNonUniformityNormalized", "SmallAreaEmphasis",
  "SmallAreaHighGrayLevelEmphasis", "SmallAreaLowGrayLevelEmphasis", "ZoneEntropy", "ZonePercentage", "ZoneVariance"};
static const char* kGldmNames[] = {"DependenceEntropy", "DependenceNonUniformity", "DependenceNonUniformityNormalized",
  "DependenceVariance", "GrayLevelNonUniformity", "GrayLevelVariance", "HighGrayLevelEmphasis", "LargeDependenceEmphasis",
  "LargeDependenceHighGrayLevelEmphasis", "LargeDependenceLowGrayLevelEmphasis", "LowGrayLevelEmphasis",
  "SmallDependenceEmphasis", "SmallDependenceHighGrayLevelEmphasis", "SmallDependenceLowGrayLevelEmphasis"};
static const char* kNgtdmNames[] = {"Busyness", "Coarseness", "Complexity", "Contrast", "Strength"};
static const char** kNames[5] = {kGlcmNames, kGlrlmNames, kGlszmNames, kGldmNames, kNgtdmNames};

static VoxSettings to_settings(const rb_voxel_settings* s) {
  VoxSettings v;
  static_assert(sizeof(VoxSettings) == sizeof(rb_voxel_settings), "settings mirror out of sync");
  memcpy(&v, s, sizeof v);
  return v;
}

__global__ void maps_to_f32_kernel(const double* __restrict__ src, long long sp, float* __restrict__ dst, long long dp,
                                   long long width, long long height) {
  const long long total = width * height;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / width, c = i - r * width;
    dst[r * dp + c] = (float)src[r * sp + c];
  }
}

}  // namespace rb

using namespace rb;

extern "C" {

const char* rb_last_error(void) { return last_error_ref().c_str(); }
const char* rb_version(void) { return "b200radiomics 0.1 (sm_100a)"; }

int rb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(RB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

int rb_release_device_caches(void) { return glcm_release_queues(); }

int rb_num_features(int cls) { return (cls < 0 || cls > 4) ? RB_ERR_ARG : kNumFeatures[cls]; }

const char* rb_feature_name(int cls, int idx) {
  if (cls < 0 || cls > 4 || idx < 0 || idx >= kNumFeatures[cls]) return NULL;
  return kNames[cls][idx];
}

int rb_generate_angles(const int* size, int nd, const int* distances, int ndist, int bidirectional, int force2D,
                       int force2Ddimension, int* angles, int max_angles) {
  if (!size || !distances || nd < 1 || nd > 3 || ndist < 1) return fail(RB_ERR_ARG, "bad size/distances");
  std::vector<int> out;
  int na = generate_angles(size, nd, distances, ndist, bidirectional != 0, force2D ? force2Ddimension : -1, out);
  if (na <= 0) return fail(RB_ERR_ARG, "Error getting angle count.");
  if (na > max_angles) return -1000 - na;
  memcpy(angles, out.data(), sizeof(int) * (size_t)na * nd);
  return na;
}

int rb_level_bytes(int Ng) { return Ng <= 255 ? 1 : 2; }

int rb_pack_levels_dev(const int32_t* image_dev, const uint8_t* mask_dev, long long nvoxels, int Ng, void* levels_dev,
                       uint32_t* presence_dev, int* status_dev, void* stream) {
  return pack_levels(image_dev, mask_dev, nvoxels, Ng, levels_dev, presence_dev, status_dev, (cudaStream_t)stream);
}

int rb_glcm_alive_angles_dev(const void* levels_dev, int level_bytes, const uint8_t* centers_dev, int Z, int Y, int X,
                             const rb_voxel_settings* settings, uint32_t* alive_dev, void* stream) {
  VoxParams P;
  if (fill_vox_params(C_GLCM, Z, Y, X, to_settings(settings), P)) return fail(RB_ERR_ARG, "bad voxel settings");
  return glcm_alive_angles(levels_dev, level_bytes, centers_dev, P, alive_dev, (cudaStream_t)stream);
}

int rb_voxel_features_dev(int cls, const void* levels_dev, int level_bytes, const uint8_t* centers_dev, int Z, int Y,
                          int X, int z0, int z1, const rb_voxel_settings* settings, const uint32_t* alive_host,
                          void* out_dev, int out_is_f32, long long out_feature_stride, int out_z0, int* status_dev,
                          void* stream) {
  if (cls < 0 || cls > 4) return fail(RB_ERR_ARG, "unknown texture class %d", cls);
  if (out_is_f32) return fail(RB_ERR_UNSUPPORTED, "float32 maps not implemented yet");
  if (z0 < 0 || z1 > Z || z0 > z1) return fail(RB_ERR_ARG, "bad z range");
  VoxParams P;
  if (fill_vox_params(cls, Z, Y, X, to_settings(settings), P)) return fail(RB_ERR_ARG, "bad voxel settings");
  if (alive_host && cls == C_GLCM) memcpy(P.alive, alive_host, sizeof P.alive);
  if (!force_generic() && glcm_fast_applicable(cls, level_bytes, P))
    return glcm_fast_launch(levels_dev, centers_dev, P, (double*)out_dev, out_feature_stride, z0, z1, out_z0,
                            (cudaStream_t)stream);
  if (!force_generic() && small_fast_applicable(cls, level_bytes, P))
    return small_fast_launch(cls, levels_dev, centers_dev, P, (double*)out_dev, out_feature_stride, z0, z1, out_z0,
                             (cudaStream_t)stream);
  if (!force_generic() && glrlm_fast_applicable(cls, level_bytes, P))
    return glrlm_fast_launch(levels_dev, centers_dev, P, (double*)out_dev, out_feature_stride, z0, z1, out_z0,
                             (cudaStream_t)stream);
  return voxel_features_generic(cls, levels_dev, level_bytes, centers_dev, P, (double*)out_dev, out_feature_stride, z0,
                                z1, out_z0, status_dev, (cudaStream_t)stream);
}

int rb_memcpy2d_async(void* dst, unsigned long long dpitch, const void* src, unsigned long long spitch,
                      unsigned long long width, unsigned long long height, int kind, void* stream) {
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (kind < 1 || kind > 3) return fail(RB_ERR_ARG, "rb_memcpy2d_async: kind must be 1, 2 or 3");
  if (!width || !height) return RB_OK;
  RB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, k, (cudaStream_t)stream));
  return RB_OK;
}

int rb_maps_to_f32_dev(const double* src_dev, long long src_pitch, float* dst_dev, long long dst_pitch, long long width,
                       long long height, void* stream) {
  if (width <= 0 || height <= 0) return RB_OK;
  const long long total = width * height;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  maps_to_f32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src_dev, src_pitch, dst_dev, dst_pitch, width, height);
  RB_LAUNCH_CHECK();
  return RB_OK;
}

int rb_voxel_features_host(int cls, const int32_t* image, const uint8_t* mask, int Z, int Y, int X,
                           const rb_voxel_settings* settings, double* maps) {
  if (cls < 0 || cls > 4) return fail(RB_ERR_ARG, "unknown texture class %d", cls);
  const long long n = (long long)Z * Y * X;
  const int nf = kNumFeatures[cls];
  const int lb = rb_level_bytes(settings->Ng);
  int32_t* d_img = NULL; uint8_t* d_msk = NULL; void* d_lev = NULL; double* d_out = NULL; int* d_status = NULL;
  uint32_t* d_alive = NULL;
  int rc = RB_OK;
  auto cleanup = [&]() { cudaFree(d_img); cudaFree(d_msk); cudaFree(d_lev); cudaFree(d_out); cudaFree(d_status); cudaFree(d_alive); };
#define RB_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { cleanup(); return fail(_e == cudaErrorMemoryAllocation ? RB_ERR_NOMEM : RB_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(_e)); } } while (0)
  RB_TRY(cudaMalloc(&d_img, n * 4));
  RB_TRY(cudaMalloc(&d_msk, n));
  RB_TRY(cudaMalloc(&d_lev, n * lb));
  RB_TRY(cudaMalloc(&d_out, sizeof(double) * n * nf));
  RB_TRY(cudaMalloc(&d_status, 2 * sizeof(int)));
  RB_TRY(cudaMalloc(&d_alive, RB_ALIVE_WORDS * 4));
  RB_TRY(cudaMemsetAsync(d_status, 0, 2 * sizeof(int), 0));
  RB_TRY(cudaMemsetAsync(d_alive, 0, RB_ALIVE_WORDS * 4, 0));
  RB_TRY(cudaMemcpyAsync(d_img, image, n * 4, cudaMemcpyHostToDevice, 0));
  RB_TRY(cudaMemcpyAsync(d_msk, mask, n, cudaMemcpyHostToDevice, 0));
  rc = rb_pack_levels_dev(d_img, d_msk, n, settings->Ng, d_lev, NULL, d_status, 0);
  uint32_t alive[RB_ALIVE_WORDS];
  if (rc == RB_OK && cls == RB_GLCM) {
    rc = rb_glcm_alive_angles_dev(d_lev, lb, NULL, Z, Y, X, settings, d_alive, 0);
    if (rc == RB_OK) RB_TRY(cudaMemcpy(alive, d_alive, sizeof alive, cudaMemcpyDeviceToHost));
  }
  if (rc == RB_OK)
    rc = rb_voxel_features_dev(cls, d_lev, lb, NULL, Z, Y, X, 0, Z, settings, cls == RB_GLCM ? alive : NULL, d_out, 0, n, 0,
                               d_status + 1, 0);
  if (rc == RB_OK) {
    int st[2] = {0, 0};
    RB_TRY(cudaMemcpy(st, d_status, sizeof st, cudaMemcpyDeviceToHost));
    if (st[0] & 1) rc = fail(RB_ERR_LEVEL_RANGE, "gray level outside 1..Ng inside the mask");
    else if (st[1] & 2) rc = fail(RB_ERR_UNSUPPORTED, "weighted GLCM entry list overflow");
    else RB_TRY(cudaMemcpy(maps, d_out, sizeof(double) * n * nf, cudaMemcpyDeviceToHost));
  }
#undef RB_TRY
  cleanup();
  return rc;
}

int rb_calculate_glcm(const int32_t* image, const uint8_t* mask, const int* size, int nd, const int* distances, int ndist,
                      int Ng, int force2D, int force2Ddimension, int kernelRadius, const int* voxels, int nvox,
                      double* glcm, int* angles) {
  return calculate_matrix_host(0, image, mask, size, nd, distances, ndist, Ng, 0, 0, force2D, force2Ddimension,
                               kernelRadius, voxels, nvox, glcm, angles, NULL);
}
int rb_calculate_glrlm(const int32_t* image, const uint8_t* mask, const int* size, int nd, int Ng, int Nr, int force2D,
                       int force2Ddimension, int kernelRadius, const int* voxels, int nvox, double* glrlm, int* angles) {
  return calculate_matrix_host(3, image, mask, size, nd, NULL, 0, Ng, Nr, 0, force2D, force2Ddimension, kernelRadius,
                               voxels, nvox, glrlm, angles, NULL);
}
int rb_calculate_gldm(const int32_t* image, const uint8_t* mask, const int* size, int nd, const int* distances, int ndist,
                      int Ng, int alpha, int force2D, int force2Ddimension, int kernelRadius, const int* voxels, int nvox,
                      double* gldm) {
  return calculate_matrix_host(1, image, mask, size, nd, distances, ndist, Ng, 0, alpha, force2D, force2Ddimension,
                               kernelRadius, voxels, nvox, gldm, NULL, NULL);
}
int rb_calculate_ngtdm(const int32_t* image, const uint8_t* mask, const int* size, int nd, const int* distances,
                       int ndist, int Ng, int force2D, int force2Ddimension, int kernelRadius, const int* voxels,
                       int nvox, double* ngtdm) {
  return calculate_matrix_host(2, image, mask, size, nd, distances, ndist, Ng, 0, 0, force2D, force2Ddimension,
                               kernelRadius, voxels, nvox, ngtdm, NULL, NULL);
}
// ---- segment-mode matrices from a device-resident packed level volume (no host round trip of the image)
int rb_segment_texture_dev(const void* levels_dev, int level_bytes, const int* size, int nd, const int* distances, int ndist,
                           int Ng, int alpha, int force2D, int force2Ddimension, double* glcm, double* gldm, double* ngtdm,
                           int* angles) {
  if (!levels_dev || !size || (nd != 2 && nd != 3)) return fail(RB_ERR_ARG, "levels / size");
  if (level_bytes != rb_level_bytes(Ng)) return fail(RB_ERR_ARG, "level_bytes does not match Ng");
  int dmax = 0;
  for (int i = 0; i < ndist; i++) dmax = distances[i] > dmax ? distances[i] : dmax;
  const int Z = nd == 3 ? size[0] : 1, Y = size[nd - 2], X = size[nd - 1];
  if (level_bytes == 1 && dmax <= 3) {
    const int rc = segment_tile_matrices((const uint8_t*)levels_dev, nd, Z, Y, X, distances, ndist, Ng, alpha, force2D,
                                         force2Ddimension, glcm, gldm, ngtdm, angles, NULL, 0);
    if (rc != RB_ERR_UNSUPPORTED) return rc;
  }
  int rc = RB_OK;        // 16-bit levels / long offsets / very many levels: one class at a time through round 1's kernels
  if (glcm) rc = calculate_matrix_host(0, NULL, NULL, size, nd, distances, ndist, Ng, 0, 0, force2D, force2Ddimension, 0, NULL, 1, glcm, angles, NULL, levels_dev);
  if (!rc && gldm) rc = calculate_matrix_host(1, NULL, NULL, size, nd, distances, ndist, Ng, 0, alpha, force2D, force2Ddimension, 0, NULL, 1, gldm, NULL, NULL, levels_dev);
  if (!rc && ngtdm) rc = calculate_matrix_host(2, NULL, NULL, size, nd, distances, ndist, Ng, 0, 0, force2D, force2Ddimension, 0, NULL, 1, ngtdm, NULL, NULL, levels_dev);
  return rc;
}
int rb_segment_glrlm_dev(const void* levels_dev, int level_bytes, const int* size, int nd, int Ng, int Nr, int force2D,
                         int force2Ddimension, double* glrlm, int* angles) {
  if (!levels_dev || !size || (nd != 2 && nd != 3)) return fail(RB_ERR_ARG, "levels / size");
  if (level_bytes != rb_level_bytes(Ng)) return fail(RB_ERR_ARG, "level_bytes does not match Ng");
  return calculate_matrix_host(3, NULL, NULL, size, nd, NULL, 0, Ng, Nr, 0, force2D, force2Ddimension, 0, NULL, 1, glrlm, angles, NULL, levels_dev);
}
int rb_segment_glszm_dev(const void* levels_dev, int level_bytes, const int* size, int nd, int Ng, int force2D,
                         int force2Ddimension, int* max_region, void** handle) {
  if (!levels_dev || !size || (nd != 2 && nd != 3)) return fail(RB_ERR_ARG, "levels / size");
  if (level_bytes != rb_level_bytes(Ng)) return fail(RB_ERR_ARG, "level_bytes does not match Ng");
  return glszm_zones_host(NULL, NULL, size, nd, Ng, force2D, force2Ddimension, 0, NULL, 1, max_region, handle, levels_dev);
}

int rb_calculate_glszm(const int32_t* image, const uint8_t* mask, const int* size, int nd, int Ng, int force2D,
                       int force2Ddimension, int kernelRadius, const int* voxels, int nvox, int* max_region,
                       void** handle) {
  if (!max_region || !handle) return fail(RB_ERR_ARG, "null output pointer");
  return glszm_zones_host(image, mask, size, nd, Ng, force2D, force2Ddimension, kernelRadius, voxels, nvox, max_region,
                          handle);
}
int rb_fill_glszm(void* handle, int Ng, int max_region, double* glszm) { return glszm_fill_host(handle, Ng, max_region, glszm); }
void rb_glszm_release(void* handle) { glszm_release(handle); }

int rb_minmax_dev(const void* image_dev, int dtype, const uint8_t* mask_dev, long long nvoxels, long long* keys_dev,
                  void* stream) {
  if (dtype < 0 || dtype > 6) return fail(RB_ERR_ARG, "unknown dtype code %d", dtype);
  return minmax_launch(image_dev, dtype, mask_dev, nvoxels, keys_dev, (cudaStream_t)stream);
}
int rb_digitize_dev(const void* image_dev, int dtype, const uint8_t* mask_dev, long long nvoxels, const double* edges_dev,
                    int nedges, int32_t* out_dev, void* stream) {
  if (dtype < 0 || dtype > 6) return fail(RB_ERR_ARG, "unknown dtype code %d", dtype);
  if (nedges < 1) return fail(RB_ERR_ARG, "need at least one bin edge");
  return digitize_launch(image_dev, dtype, mask_dev, nvoxels, edges_dev, nedges, out_dev, (cudaStream_t)stream);
}
int rb_swt_axis_dev(const double* in_dev, int Z, int Y, int X, int axis, const double* dec_lo, const double* dec_hi,
                    int flen, double* out_lo_dev, double* out_hi_dev, void* stream) {
  if (axis < 0 || axis > 2) return fail(RB_ERR_ARG, "axis must be 0..2");
  return swt_axis_launch(in_dev, Z, Y, X, axis, dec_lo, dec_hi, flen, out_lo_dev, out_hi_dev, (cudaStream_t)stream);
}
int rb_bspline_prefilter_dev(double* coeffs_dev, int Z, int Y, int X, void* stream) {
  return bspline_prefilter_launch(coeffs_dev, Z, Y, X, (cudaStream_t)stream);
}
int rb_resample_dev(const void* src_dev, int src_dtype, const int* in_size_zyx, void* dst_dev, int dst_dtype, const int* out_size_zyx,
                    const double* start_zyx, const double* step_zyx, int interpolator, double default_value, void* stream) {
  if (!src_dev || !dst_dev || !in_size_zyx || !out_size_zyx || !start_zyx || !step_zyx) return fail(RB_ERR_ARG, "null argument");
  return resample_launch(src_dev, src_dtype, in_size_zyx, dst_dev, dst_dtype, out_size_zyx, start_zyx, step_zyx, interpolator,
                         default_value, (cudaStream_t)stream);
}

int rb_swt3d_dev(const double* in_dev, int Z, int Y, int X, const double* dec_lo, const double* dec_hi, int flen,
                 double* out_dev, long long band_stride, int z_begin, int z_end, void* stream) {
  return swt3d_launch(in_dev, Z, Y, X, dec_lo, dec_hi, flen, out_dev, band_stride, z_begin, z_end, (cudaStream_t)stream);
}

int rb_recursive_gaussian_axis_dev(const void* in_dev, int in_is_f32, int Z, int Y, int X, int axis, const double* coef20,
                                   float* out_dev, double* scratch_dev, double scale, int accumulate, void* stream) {
  if (axis < 0 || axis > 2) return fail(RB_ERR_ARG, "axis must be 0..2");
  return recursive_gauss_launch(in_dev, in_is_f32, Z, Y, X, axis, coef20, out_dev, scratch_dev, scale, accumulate,
                                (cudaStream_t)stream);
}

int rb_shape_coefficients_dev(const uint8_t* mask_dev, int Z, int Y, int X, const double* spacing_zyx, double* out7,
                              void* stream) {
  if (!mask_dev || !spacing_zyx || !out7 || Z < 1 || Y < 1 || X < 1) return fail(RB_ERR_ARG, "shape: bad arguments");
  return shape_coefficients_dev(mask_dev, Z, Y, X, (long long)Y * X, X, 1, spacing_zyx, out7, (cudaStream_t)stream);
}
int rb_shape_moments_dev(const uint8_t* mask_dev, int Z, int Y, int X, unsigned long long* out10, void* stream) {
  if (!mask_dev || !out10 || Z < 1 || Y < 1 || X < 1) return fail(RB_ERR_ARG, "shape: bad arguments");
  return shape_moments_dev(mask_dev, Z, Y, X, out10, (cudaStream_t)stream);
}
int rb_calculate_coefficients(const char* mask, const int* size, const int* strides, const double* spacing,
                              double* surfaceArea, double* volume, double* diameters) {
  if (!mask || !size || !strides || !spacing || !surfaceArea || !volume || !diameters) return fail(RB_ERR_ARG, "shape: null argument");
  const int Z = size[0], Y = size[1], X = size[2];
  if (Z < 1 || Y < 1 || X < 1) return fail(RB_ERR_ARG, "shape: bad size");
  const size_t n = (size_t)Z * Y * X;
  std::vector<uint8_t> packed(n);
  for (int z = 0; z < Z; z++)
    for (int y = 0; y < Y; y++)
      for (int x = 0; x < X; x++)
        packed[((size_t)z * Y + y) * X + x] = mask[(long long)z * strides[0] + (long long)y * strides[1] + (long long)x * strides[2]] != 0;
  uint8_t* d = nullptr;
  RB_CUDA(cudaMalloc(&d, n));
  cudaError_t e = cudaMemcpy(d, packed.data(), n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return fail(RB_ERR_CUDA, "shape: %s", cudaGetErrorString(e)); }
  double out7[7];
  const int rc = shape_coefficients_dev(d, Z, Y, X, (long long)Y * X, X, 1, spacing, out7, 0);
  cudaFree(d);
  if (rc) return rc;
  *surfaceArea = out7[0]; *volume = out7[1];
  for (int q = 0; q < 4; q++) diameters[q] = out7[2 + q];
  return RB_OK;
}

int rb_calculate_coefficients2D(const char* mask, const int* size, const int* strides, const double* spacing, double* perimeter,
                                double* surface, double* diameter) {
  if (!mask || !size || !strides || !spacing || !perimeter || !surface || !diameter) return fail(RB_ERR_ARG, "null argument");
  const int Y = size[0], X = size[1];
  if (Y < 1 || X < 1) return fail(RB_ERR_ARG, "empty mask");
  // gather the (possibly strided) host mask into a contiguous byte image, then one upload
  std::vector<uint8_t> h((size_t)Y * X);
  for (int y = 0; y < Y; y++)
    for (int x = 0; x < X; x++) h[(size_t)y * X + x] = mask[(long long)y * strides[0] + (long long)x * strides[1]] != 0;
  uint8_t* d = NULL;
  RB_CUDA(cudaMalloc(&d, h.size()));
  cudaError_t e = cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return fail(RB_ERR_CUDA, "shape2D upload: %s", cudaGetErrorString(e)); }
  double out4[4];
  const int rc = shape2d_coefficients_dev(d, Y, X, X, 1, spacing, out4, 0);
  cudaFree(d);
  if (rc) return rc;
  *perimeter = out4[0]; *surface = out4[1]; *diameter = out4[2];
  return RB_OK;
}

int rb_firstorder_num_features(void) { return 18; }
const char* rb_firstorder_feature_name(int idx) { return (idx < 0 || idx >= 18) ? NULL : kFirstOrderNames[idx]; }
int rb_firstorder_voxel_dev(const void* image_dev, int dtype, const uint8_t* mask_dev, const uint8_t* centers_dev,
                            const void* levels_dev, int level_bytes, int Z, int Y, int X, int rz, int ry, int rx,
                            double voxelArrayShift, double voxel_volume, double initValue, double* out_dev,
                            long long out_feature_stride, int z0, int z1, int out_z0, void* stream) {
  if (dtype < 0 || dtype > 6) return fail(RB_ERR_ARG, "unknown dtype code %d", dtype);
  if (level_bytes != 1 && level_bytes != 2) return fail(RB_ERR_ARG, "level_bytes must be 1 or 2");
  if (rz < 0 || ry < 0 || rx < 0 || z0 < 0 || z1 > Z || z0 > z1) return fail(RB_ERR_ARG, "bad window / z range");
  return firstorder_launch(image_dev, dtype, mask_dev, centers_dev, levels_dev, level_bytes, Z, Y, X, rz, ry, rx,
                           voxelArrayShift, voxel_volume, initValue, out_dev, out_feature_stride, z0, z1, out_z0,
                           (cudaStream_t)stream);
}

}  // extern "C"
