// GLRLM fast path: kernelRadius 1, full 3-D (13 distance-1 angles), unweighted, 8-bit levels.
// Same semantics as glrlm_voxel<> (vox_features.cuh), restructured around bitmasks: the 27 window
// levels are compared once into per-position equality masks; for each angle every masked window
// voxel that ENDS a run (its successor along the angle is outside the window / unmasked / another
// level) contributes one run whose length (1..3) comes from two more bit tests.  Sums that are
// linear in the runs are exact integers; sum_i pg(i)^2 and the run entropy come from popcounts of
// (equality mask & run-end masks) and a log2 table -- no merged run list, no transcendental calls.
// __host__ __device__ (tests/host_emul checks it on the CPU; test-only).
#pragma once
#include "glcm_fast.cuh"

namespace rb {

struct GlrlmFastTables {
  uint32_t step[GF_NA][27];   // nxt | prv << 8 | pprv << 16 (window position, 31 = outside the window)
  uint32_t VA[GF_NA], VA2[GF_NA];   // positions whose successor / second successor is inside the window
  uint8_t delta[GF_NA];       // index offset of the angle (dz*9 + dy*3 + dx > 0 for the 13 angles)
  double log2t[32];
  double inv2[256];           // 1 / level^2
};

inline void glrlm_fast_build_tables(GlrlmFastTables& T) {
  int ang[13][3], k = 0;
  for (int z = 1; z >= -1; z--) for (int y = 1; y >= -1; y--) for (int x = 1; x >= -1; x--)
    if (k < 13) { ang[k][0] = z; ang[k][1] = y; ang[k][2] = x; k++; }
  for (int a = 0; a < 13; a++) {
    T.VA[a] = T.VA2[a] = 0;
    T.delta[a] = (uint8_t)(ang[a][0] * 9 + ang[a][1] * 3 + ang[a][2]);
    for (int z = 0; z < 3; z++) for (int y = 0; y < 3; y++) for (int x = 0; x < 3; x++) {
      const int v = z * 9 + y * 3 + x;
      auto pos = [&](int m) {
        const int z2 = z + m * ang[a][0], y2 = y + m * ang[a][1], x2 = x + m * ang[a][2];
        return (z2 < 0 || z2 > 2 || y2 < 0 || y2 > 2 || x2 < 0 || x2 > 2) ? 31 : z2 * 9 + y2 * 3 + x2;
      };
      const int nx = pos(1), nx2 = pos(2), pv = pos(-1), ppv = pos(-2);
      T.step[a][v] = (uint32_t)nx | (uint32_t)pv << 8 | (uint32_t)ppv << 16;
      if (nx != 31) T.VA[a] |= 1u << v;
      if (nx2 != 31) T.VA2[a] |= 1u << v;
    }
  }
  T.log2t[0] = 0;
  for (int c = 1; c < 32; c++) T.log2t[c] = log2((double)c);
  T.inv2[0] = 0;
  for (int g = 1; g < 256; g++) T.inv2[g] = 1.0 / ((double)g * g);
}

// wl: the 27 window levels (registers), out: 16 features in GlrlmF order
RB_HD void glrlm_fast_voxel(const int* wl, const GlrlmFastTables& T, double* out) {
  uint32_t e[27];
  RB_EQMASKS_27(wl, e);
  uint32_t M = 0;
#pragma unroll
  for (int v = 0; v < 27; v++) if (wl[v]) M |= 1u << v;
  const int Np = RB_POPC(M);
  double sum[GLRLM_NF];
#pragma unroll
  for (int k = 0; k < GLRLM_NF; k++) sum[k] = 0;
  int nang = 0;
  for (int a = 0; a < GF_NA; a++) {
    const int d = T.delta[a];
    // cmatrices.c:524-534: an angle none of whose lines holds two masked voxels is dropped
    if (!(((M & T.VA[a]) & (M >> d)) | ((M & T.VA2[a]) & (M >> (2 * d))))) continue;
    uint32_t ENDS = 0, L1 = 0, L2 = 0;
    int n1 = 0, n2 = 0, n3 = 0, B1 = 0, B2 = 0, B3 = 0, C = 0;
    double A1 = 0, A2 = 0, A3 = 0;
#pragma unroll
    for (int v = 0; v < 27; v++) {
      const uint32_t ev = e[v];
      const uint32_t st = T.step[a][v];
      const bool is_end = ev != 0 && !((ev >> (st & 31)) & 1u);
      if (is_end) {
        const uint32_t ps = (ev >> ((st >> 8) & 31)) & 1u, pps = ps & ((ev >> ((st >> 16) & 31)) & 1u);
        const int g = wl[v], g2 = g * g;
        const double ig = T.inv2[g];
        ENDS |= 1u << v;
        C += g;
        if (!ps) { L1 |= 1u << v; n1++; B1 += g2; A1 += ig; }
        else if (!pps) { L2 |= 1u << v; n2++; B2 += g2; A2 += ig; }
        else { n3++; B3 += g2; A3 += ig; }
      }
    }
    const uint32_t L3 = ENDS & ~(L1 | L2);
    int sg = 0;
    double lg = 0;
#pragma unroll
    for (int v = 0; v < 27; v++) {
      if (ENDS >> v & 1u) {
        sg += RB_POPC(e[v] & ENDS);
        const uint32_t Lm = (L1 >> v & 1u) ? L1 : (L2 >> v & 1u) ? L2 : L3;
        lg += T.log2t[RB_POPC(e[v] & Lm)];
      }
    }
    const int Nr = n1 + n2 + n3;
    const double invNr = 1.0 / Nr, invNr2 = invNr * invNr;
    const int lre_n = n1 + 4 * n2 + 9 * n3, B = B1 + B2 + B3;
    sum[R_ShortRunEmphasis] += (n1 + n2 * 0.25 + n3 * (1.0 / 9.0)) * invNr;
    sum[R_LongRunEmphasis] += lre_n * invNr;
    sum[R_GrayLevelNonUniformity] += sg * invNr;
    sum[R_GrayLevelNonUniformityNormalized] += sg * invNr2;
    const int rl = n1 * n1 + n2 * n2 + n3 * n3;
    sum[R_RunLengthNonUniformity] += rl * invNr;
    sum[R_RunLengthNonUniformityNormalized] += rl * invNr2;
    sum[R_RunPercentage] += (double)Nr / Np;
    sum[R_GrayLevelVariance] += (double)(Nr * B - C * C) * invNr2;
    sum[R_RunVariance] += (double)(Nr * lre_n - Np * Np) * invNr2;
    sum[R_RunEntropy] += T.log2t[Nr] - lg * invNr;
    sum[R_LowGrayLevelRunEmphasis] += (A1 + A2 + A3) * invNr;
    sum[R_HighGrayLevelRunEmphasis] += B * invNr;
    sum[R_ShortRunLowGrayLevelEmphasis] += (A1 + A2 * 0.25 + A3 * (1.0 / 9.0)) * invNr;
    sum[R_ShortRunHighGrayLevelEmphasis] += (B1 + B2 * 0.25 + B3 * (1.0 / 9.0)) * invNr;
    sum[R_LongRunLowGrayLevelEmphasis] += (A1 + 4.0 * A2 + 9.0 * A3) * invNr;
    sum[R_LongRunHighGrayLevelEmphasis] += (double)(B1 + 4 * B2 + 9 * B3) * invNr;
    nang++;
  }
  const double inv = nang ? 1.0 / nang : NAN;
#pragma unroll
  for (int k = 0; k < GLRLM_NF; k++) out[k] = nang ? sum[k] * inv : NAN;
}

}  // namespace rb
