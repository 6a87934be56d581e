// Fast paths of GLSZM, GLDM and NGTDM for the headline configuration (kernelRadius 1, full 3-D,
// distance-1 26-neighbourhood, 8-bit levels).  Same semantics as glszm_voxel / gldm_voxel /
// ngtdm_voxel (vox_features.cuh, the generic fallback and cross-check), built on the 27 x 27-bit
// equality masks of the window like the GLCM / GLRLM fast paths:
//   GLDM  dependence of a voxel = popcount(close-level mask & static neighbour mask); merged
//         (level, dependence) counts = popcount(equal-level mask & equal-dependence mask)
//   NGTDM neighbour sums from static neighbour lists (unmasked voxels carry level 0); the pairwise
//         level loop is only needed for Busyness / Complexity, Contrast and Strength collapse to
//         closed forms in integer moments
//   GLSZM zones of one level = flood fill of its equality mask by separable bitmask dilation
// __host__ __device__ (tests/host_emul checks them on the CPU; test-only).
#pragma once
#include "glcm_fast.cuh"

namespace rb {

struct SmallFastTables {
  double log2t[32];     // log2(c), c = 0..31
  double inv2[256];     // 1 / g^2
  double invsq[32];     // 1 / j^2, j = 1..28
  double rcp[64];       // 1 / c
};

inline void small_fast_build_tables(SmallFastTables& T) {
  T.log2t[0] = 0; T.invsq[0] = 0; T.inv2[0] = 0; T.rcp[0] = 0;
  for (int c = 1; c < 32; c++) { T.log2t[c] = log2((double)c); T.invsq[c] = 1.0 / ((double)c * c); }
  for (int g = 1; g < 256; g++) T.inv2[g] = 1.0 / ((double)g * g);
  for (int c = 1; c < 64; c++) T.rcp[c] = 1.0 / (double)c;
}

// 26-neighbourhood of window position v inside the 3x3x3 window (compile-time constant per v)
RB_HD constexpr uint32_t nb26(int v) {
  uint32_t m = 0;
  const int z = v / 9, y = (v / 3) % 3, x = v % 3;
  for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
    if (!dz && !dy && !dx) continue;
    const int z2 = z + dz, y2 = y + dy, x2 = x + dx;
    if (z2 < 0 || z2 > 2 || y2 < 0 || y2 > 2 || x2 < 0 || x2 > 2) continue;
    m |= 1u << (z2 * 9 + y2 * 3 + x2);
  }
  return m;
}
template <int V> struct NB26 { static constexpr uint32_t value = nb26(V); };

// compile-time unrolled loop over the 27 window positions: F::template run<V>(args...)
template <int V, int END> struct ForPos {
  template <typename F> static RB_HD void go(F& f) { f.template at<V>(); ForPos<V + 1, END>::go(f); }
};
template <int END> struct ForPos<END, END> { template <typename F> static RB_HD void go(F&) {} };

// ---------------------------------------------------------------------------------------- GLDM
struct GldmPass1 {
  const int* wl; const uint32_t* cl; int* dep; uint32_t M;
  template <int V> RB_HD void at() { dep[V] = wl[V] ? (int)RB_POPC(cl[V] & NB26<V>::value) : -1 - V; }
};

RB_HD void gldm_fast_voxel(const int* wl, int alpha, const SmallFastTables& T, double* out) {
  uint32_t eq[27];
  RB_EQMASKS_27(wl, eq);
  uint32_t cl[27];
  if (alpha == 0) {
#pragma unroll
    for (int v = 0; v < 27; v++) cl[v] = eq[v];
  } else {
    // "dependent" relation |g_u - g_v| <= alpha between masked voxels
#pragma unroll
    for (int v = 0; v < 27; v++) cl[v] = 0;
#pragma unroll
    for (int p = 0; p < 27; p++)
#pragma unroll
      for (int q = p + 1; q < 27; q++) {
        const int d = wl[p] - wl[q];
        if (wl[p] && wl[q] && d <= alpha && -d <= alpha) { cl[p] |= 1u << q; cl[q] |= 1u << p; }
      }
  }
  int dep[27];
  GldmPass1 p1{wl, cl, dep, 0};
  ForPos<0, 27>::go(p1);
  // equal-dependence masks (unmasked positions carry distinct negative sentinels)
  uint32_t dq[27];
  RB_EQMASKS_27_KEY(dep, dq);
  int Nz = 0, Sj = 0, Sj2 = 0, B = 0, C = 0, X4 = 0, gl = 0, dn = 0;
  double Sinv = 0, A = 0, X1 = 0, X2 = 0, X3 = 0, lg = 0;
#pragma unroll
  for (int v = 0; v < 27; v++) {
    if (wl[v]) {
      const int g = wl[v], g2 = g * g, j = dep[v] + 1, j2 = j * j;
      const double ig = T.inv2[g], ij = T.invsq[j];
      Nz++; Sj += j; Sj2 += j2; B += g2; C += g; X4 += g2 * j2;
      Sinv += ij; A += ig; X1 += ig * ij; X2 += g2 * ij; X3 += j2 * ig;
      gl += RB_POPC(eq[v]); dn += RB_POPC(dq[v]);
      lg += T.log2t[RB_POPC(eq[v] & dq[v])];
    }
  }
  const double inv = 1.0 / Nz, inv2 = inv * inv;
  out[0] = T.log2t[Nz] - lg * inv;                 // DependenceEntropy
  out[1] = dn * inv;                               // DependenceNonUniformity
  out[2] = dn * inv2;                              // DependenceNonUniformityNormalized
  out[3] = (double)(Nz * Sj2 - Sj * Sj) * inv2;    // DependenceVariance
  out[4] = gl * inv;                               // GrayLevelNonUniformity
  out[5] = (double)(Nz * B - C * C) * inv2;        // GrayLevelVariance
  out[6] = B * inv;                                // HighGrayLevelEmphasis
  out[7] = Sj2 * inv;                              // LargeDependenceEmphasis
  out[8] = X4 * inv;                               // LargeDependenceHighGrayLevelEmphasis
  out[9] = X3 * inv;                               // LargeDependenceLowGrayLevelEmphasis
  out[10] = A * inv;                               // LowGrayLevelEmphasis
  out[11] = Sinv * inv;                            // SmallDependenceEmphasis
  out[12] = X2 * inv;                              // SmallDependenceHighGrayLevelEmphasis
  out[13] = X1 * inv;                              // SmallDependenceLowGrayLevelEmphasis
}

// ---------------------------------------------------------------------------------------- NGTDM
struct NgtdmPass1 {
  const int* wl; uint32_t M; const SmallFastTables* T; double* diff;
  template <int V> RB_HD void at() {
    if (!wl[V]) { diff[V] = 0; return; }
    constexpr uint32_t nb = NB26<V>::value;
    int sum = 0;
#pragma unroll
    for (int u = 0; u < 27; u++) if (nb >> u & 1u) sum += wl[u];      // static neighbour list
    const int cnt = RB_POPC(M & nb);
    diff[V] = cnt ? fabs((double)wl[V] - (double)sum / (double)cnt) : 0.0;
  }
};

// scr_pk / scr_cs: per-thread scratch for the compacted level classes, 27 entries each with element stride st (device:
// shared memory laid out [entry][thread]; round 1 kept them in dynamically indexed local arrays -- 482 M local loads per
// 256^3 volume, the kernel was as slow as GLRLM for 5 features).  pk = class size << 8 | level, cs = class sum of diff.
constexpr int NGTDM_SCR_BYTES = 27 * (int)(sizeof(int) + sizeof(double));
RB_HD void ngtdm_fast_voxel(const int* wl, const SmallFastTables& T, double* out, int* scr_pk, double* scr_cs, int st) {
  uint32_t eq[27];
  RB_EQMASKS_27(wl, eq);
  uint32_t M = 0, rep = 0;
#pragma unroll
  for (int v = 0; v < 27; v++) {
    if (wl[v]) M |= 1u << v;
    if (eq[v] && (eq[v] & ((1u << v) - 1)) == 0) rep |= 1u << v;
  }
  double diff[27];
  NgtdmPass1 p1{wl, M, &T, diff};
  ForPos<0, 27>::go(p1);
  const int Nvp = RB_POPC(M), nlev = RB_POPC(rep);
  // per level: n = class size, i = level, s = class sum of diff -- compacted (data-dependent count)
  int B = 0, C = 0, SL = 0, SL2 = 0, nl = 0;
  double ssum = 0, pw = 0;       // sum_i s_i ; sum_i n_i s_i
#pragma unroll
  for (int v = 0; v < 27; v++) {
    if (wl[v]) { B += wl[v] * wl[v]; C += wl[v]; ssum += diff[v]; pw += diff[v] * (double)RB_POPC(eq[v]); }
    if (rep >> v & 1u) {
      double s = 0;
#pragma unroll
      for (int u = v; u < 27; u++) if (eq[v] >> u & 1u) s += diff[u];      // (v is the lowest position of its class)
      scr_pk[nl * st] = RB_POPC(eq[v]) << 8 | wl[v];
      scr_cs[nl * st] = s;
      nl++;
      SL += wl[v]; SL2 += wl[v] * wl[v];
    }
  }
  // pairwise level terms: Busyness denominator sum_ij |i p_i - j p_j|, Complexity numerator
  double busy = 0, cpx = 0;
  for (int a = 0; a + 1 < nl; a++) {
    const int pa = scr_pk[a * st], na = pa >> 8, ia = pa & 255;
    const double sa = na * scr_cs[a * st];
    const int ina = ia * na;
    double cp = 0;
    int bs = 0;
    for (int b = a + 1; b < nl; b++) {
      const int pb = scr_pk[b * st], nb_ = pb >> 8, ib = pb & 255;
      const int x = ina - ib * nb_;
      bs += x < 0 ? -x : x;
      const int d = ia > ib ? ia - ib : ib - ia;
      cp += (double)d * (sa + nb_ * scr_cs[b * st]) * T.rcp[na + nb_];
    }
    busy += (double)bs;
    cpx += cp;
  }
  const double invN = 1.0 / Nvp;
  busy *= 2.0 * invN;                         // both orders, p = n / Nvp
  cpx *= 2.0 * invN;                          // sum over ordered pairs, then / Nvp below... (see Complexity)
  const double ps = pw * invN;                // sum_i p_i s_i
  out[N_Coarseness] = ps != 0 ? 1.0 / ps : 1e6;
  const double div = (double)nlev * (nlev - 1);
  const double con = 2.0 * (double)(Nvp * B - C * C) * invN * invN;     // sum_ij p_i p_j (i-j)^2
  out[N_Contrast] = div != 0 ? con * ssum * invN / div : 0.0;
  out[N_Busyness] = busy != 0 ? ps / busy : 0.0;
  out[N_Complexity] = cpx;                    // = sum_{i != j} |i-j| (p_i s_i + p_j s_j)/(p_i + p_j) / Nvp
  // Strength = sum_ij (p_i + p_j)(i-j)^2 / sum s = (2/Nvp) (nlev*B - 2*C*SL + Nvp*SL2) / sum s
  const double str = 2.0 * invN * (double)(nlev * B - 2 * C * SL + Nvp * SL2);
  out[N_Strength] = ssum != 0 ? str / ssum : 0.0;
}
// convenience: private scratch (host emulation)
RB_HD void ngtdm_fast_voxel(const int* wl, const SmallFastTables& T, double* out) {
  int pk[27];
  double cs[27];
  ngtdm_fast_voxel(wl, T, out, pk, cs, 1);
}

// ---------------------------------------------------------------------------------------- GLSZM
RB_HD uint32_t dilate26(uint32_t m) {
  constexpr uint32_t X0 = 0x1249249u, X2 = 0x4924924u;       // positions with x == 0 / x == 2
  constexpr uint32_t Y0 = 0x01C0E07u, Y2 = 0x70381C0u;       // y == 0 / y == 2
  m |= ((m & ~X2) << 1) | ((m & ~X0) >> 1);
  m |= ((m & ~Y2) << 3) | ((m & ~Y0) >> 3);
  m |= (m << 9) | (m >> 9);
  return m & 0x7FFFFFFu;
}

struct GlszmAcc {
  int Nz, Sg, Sg2, Ss2, X4, gln;
  double A, Sinv, X1, X2, X3, lg;
  unsigned long long h0, h1, h2;      // zone-size histogram, 5-bit fields: sizes 1..12 | 13..24 | 25..27
};

RB_HD void glszm_fast_voxel(const int* wl, const SmallFastTables& T, double* out) {
  // equality masks of the window (as the other fast paths).  A level that occurs ONCE is one zone of size 1: its share
  // of every sum is added in the static loop below (27 i.i.d. levels out of 32: ~12 such levels); only the levels that
  // occur at least twice (<= 13 of them, ~7) go through the flood fill.
  uint32_t eq[27];
  RB_EQMASKS_27(wl, eq);
  uint32_t M = 0;
  uint32_t cls[13];
  int clg[13];
  int nl = 0, S_n = 0, S_g = 0, S_g2 = 0;
  double S_ig = 0;
#pragma unroll
  for (int v = 0; v < 27; v++) {
    if (wl[v]) M |= 1u << v;
    if (eq[v] && (eq[v] & ((1u << v) - 1)) == 0) {
      if (eq[v] == (1u << v)) { S_n++; S_g += wl[v]; S_g2 += wl[v] * wl[v]; S_ig += T.inv2[wl[v]]; }
      else { cls[nl] = eq[v]; clg[nl] = wl[v]; nl++; }
    }
  }
  GlszmAcc a;
  a.Nz = S_n; a.Sg = S_g; a.Sg2 = S_g2; a.Ss2 = S_n; a.X4 = S_g2; a.gln = S_n;
  a.A = S_ig; a.Sinv = (double)S_n; a.X1 = S_ig; a.X2 = (double)S_g2; a.X3 = S_ig; a.lg = 0;
  a.h0 = (unsigned long long)S_n; a.h1 = a.h2 = 0;
  for (int k = 0; k < nl; k++) {
    const int g = clg[k], g2 = g * g;
    uint32_t m = cls[k];
    const double ig = T.inv2[g];
    int zs[8];                                         // <= 8 mutually non-adjacent zones fit a 3x3x3 window
    int zc = 0;
    while (m && zc < 8) {
      uint32_t comp = m & (0u - m);
      for (;;) {
        const uint32_t nx = dilate26(comp) & m;
        if (nx == comp) break;
        comp = nx;
      }
      m &= ~comp;
      const int sz = RB_POPC(comp), s2 = sz * sz;
      zs[zc++] = sz;
      const double is = T.invsq[sz];
      a.Nz++; a.Sg += g; a.Sg2 += g2; a.Ss2 += s2; a.X4 += g2 * s2;
      a.A += ig; a.Sinv += is; a.X1 += ig * is; a.X2 += g2 * is; a.X3 += s2 * ig;
      if (sz <= 12) a.h0 += 1ull << (5 * (sz - 1));
      else if (sz <= 24) a.h1 += 1ull << (5 * (sz - 13));
      else a.h2 += 1ull << (5 * (sz - 25));
    }
    a.gln += zc * zc;
    // merged (level, size) counts inside this level: zones of equal size
    for (int k = 0; k < zc; k++) {
      int c = 0;
      for (int l = 0; l < zc; l++) c += zs[l] == zs[k];
      a.lg += T.log2t[c];
    }
  }
  int szn = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    const int c0 = (int)(a.h0 >> (5 * k)) & 31, c1 = (int)(a.h1 >> (5 * k)) & 31;
    szn += c0 * c0 + c1 * c1;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { const int c2 = (int)(a.h2 >> (5 * k)) & 31; szn += c2 * c2; }
  const int Np = RB_POPC(M), Nz = a.Nz;
  const double inv = 1.0 / Nz, inv2 = inv * inv;
  out[S_GLN] = a.gln * inv; out[S_GLNN] = a.gln * inv2;
  out[S_GLV] = (double)(Nz * a.Sg2 - a.Sg * a.Sg) * inv2;
  out[S_HGLE] = a.Sg2 * inv; out[S_LargeE] = a.Ss2 * inv; out[S_LargeHGLE] = a.X4 * inv; out[S_LargeLGLE] = a.X3 * inv;
  out[S_LGLE] = a.A * inv; out[S_SizeNU] = szn * inv; out[S_SizeNUN] = szn * inv2; out[S_SmallE] = a.Sinv * inv;
  out[S_SmallHGLE] = a.X2 * inv; out[S_SmallLGLE] = a.X1 * inv;
  out[S_Entropy] = T.log2t[Nz] - a.lg * inv;
  out[S_Percentage] = (double)Nz / Np;
  out[S_SizeVar] = (double)(Nz * a.Ss2 - Np * Np) * inv2;
}

}  // namespace rb
