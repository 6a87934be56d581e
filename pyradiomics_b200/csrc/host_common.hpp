// Host-side helpers shared by the C-ABI (capi.cu) and the test-only host emulation:
// neighbour-offset ("angle") enumeration and the per-launch parameter block.
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "vox_features.cuh"

namespace rb {

// Enumerate neighbour offsets exactly in the reference's order (reference
// radiomics/src/cmatrices.c:756-892, get_angle_count + build_angles): components run from
// +maxdist down to -maxdist, dimension 0 slowest; an offset is kept when its Chebyshev norm is a
// requested distance, |component| < size in every dimension and it does not leave the force2D
// plane.  The list is point-symmetric, so "unidirectional" is its first half.
// Returns the offsets as (dz,dy,dx) rows for nd == 3 and (dy,dx) rows for nd == 2.
inline int generate_angles(const int* size, int nd, const int* distances, int ndist, bool bidirectional,
                           int force2Ddim /* -1 = off */, std::vector<int>& out) {
  out.clear();
  if (nd < 1 || nd > 3) return -1;
  int D = 0;
  for (int i = 0; i < ndist; i++) { if (distances[i] < 1) return 0; if (distances[i] > D) D = distances[i]; }
  std::vector<int> all;
  int off[3] = {0, 0, 0};
  long long total = 1;
  for (int d = 0; d < nd; d++) total *= (2 * D + 1);
  for (long long c = 0; c < total; c++) {
    long long r = c;
    for (int d = nd - 1; d >= 0; d--) { off[d] = D - (int)(r % (2 * D + 1)); r /= (2 * D + 1); }
    int norm = 0; bool ok = true;
    for (int d = 0; d < nd; d++) {
      int a = off[d] < 0 ? -off[d] : off[d];
      if (a >= size[d] || (d == force2Ddim && a != 0)) ok = false;
      if (a > norm) norm = a;
    }
    if (!ok || norm == 0) continue;
    bool wanted = false;
    for (int i = 0; i < ndist; i++) if (distances[i] == norm) wanted = true;
    if (!wanted) continue;
    for (int d = 0; d < nd; d++) all.push_back(off[d]);
  }
  int na = (int)(all.size() / nd);
  if (!bidirectional) na /= 2;
  out.assign(all.begin(), all.begin() + (size_t)na * nd);
  return na;
}

enum Weighting { W_NONE = 0, W_INFINITY = 1, W_EUCLIDEAN = 2, W_MANHATTAN = 3, W_NO_WEIGHTING = 4 };
enum TexClass { C_GLCM = 0, C_GLRLM = 1, C_GLSZM = 2, C_GLDM = 3, C_NGTDM = 4 };
static const int kNumFeatures[5] = {GLCM_NF, GLRLM_NF, GLSZM_NF, GLDM_NF, NGTDM_NF};

// per-angle weight (reference radiomics/glcm.py:160-181 -> exp(-d^2); glrlm.py:130-150 -> d)
inline double angle_weight(const int* a3, const double* spacing_zyx, int weighting, bool glcm) {
  double v[3];
  for (int d = 0; d < 3; d++) v[d] = fabs((double)a3[d]) * spacing_zyx[d];
  double dist;
  switch (weighting) {
    case W_INFINITY: dist = fmax(v[0], fmax(v[1], v[2])); break;
    case W_EUCLIDEAN: dist = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); break;
    case W_MANHATTAN: dist = v[0] + v[1] + v[2]; break;
    default: return 1.0;
  }
  return glcm ? exp(-dist * dist) : dist;
}

struct VoxSettings {      // mirror of rb_voxel_settings in include/b200radiomics.h
  int kernelRadius;
  int force2D, force2Ddimension;
  int ndist; int distances[8];
  int symmetricalGLCM;
  int weighting;
  double spacing_zyx[3];
  int gldm_a;
  double initValue;
  int Ng;               // max level of the ROI
  int n_roi_levels;
};

// Build the launch parameter block of one class for a (Z,Y,X) volume.  Returns 0 or a negative
// error (-3 bad argument / unsupported size).
inline int fill_vox_params(int cls, int Z, int Y, int X, const VoxSettings& s, VoxParams& P) {
  memset(&P, 0, sizeof(P));
  P.Z = Z; P.Y = Y; P.X = X; P.sy = X; P.sz = (long long)X * Y;
  int f2 = s.force2D ? s.force2Ddimension : -1;
  int r = s.kernelRadius;
  if (r < 1) return -3;
  P.rz = f2 == 0 ? 0 : r; P.ry = f2 == 1 ? 0 : r; P.rx = f2 == 2 ? 0 : r;
  int size[3] = {Z, Y, X};
  int one[1] = {1};
  const int* dist = s.distances; int nd = s.ndist; bool bidir = true;
  if (cls == C_GLCM) bidir = false;
  if (cls == C_GLRLM) { bidir = false; dist = one; nd = 1; }
  if (cls == C_GLSZM) { dist = one; nd = 1; }
  std::vector<int> ang;
  int na = generate_angles(size, 3, dist, nd, bidir, f2, ang);
  if (na <= 0 || na > NA_MAX || (!bidir && na > NW_MAX)) return -3;
  P.na = na;
  for (int a = 0; a < na; a++) for (int d = 0; d < 3; d++) P.ang[a][d] = (int8_t)ang[a * 3 + d];
  P.symmetric = s.symmetricalGLCM; P.alpha = s.gldm_a; P.Ng = s.Ng; P.n_roi_levels = s.n_roi_levels;
  P.init_value = s.initValue;
  P.weighted = (s.weighting != W_NONE && (cls == C_GLCM || cls == C_GLRLM)) ? 1 : 0;
  if (P.weighted)
    for (int a = 0; a < na; a++) P.wgt[a] = angle_weight(&ang[a * 3], s.spacing_zyx, s.weighting, cls == C_GLCM);
  for (int a = 0; a < na && a < NW_MAX; a++) P.alive[a >> 5] |= 1u << (a & 31);
  return 0;
}

inline int window_capacity(const VoxParams& P) { return (2 * P.rz + 1) * (2 * P.ry + 1) * (2 * P.rx + 1); }

}  // namespace rb
