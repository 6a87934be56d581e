/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference texture-matrix
 * algorithms, used as the checker of the CUDA path (tests/, __graft_entry__.smoke(),
 * bench.py cpu_baseline leg).  Nothing in the product path may link or call this file.
 *
 * Restates, for 2-D/3-D C-contiguous (z,y,x) volumes (2-D = nz 1):
 *   calculate_glcm   reference radiomics/src/cmatrices.c:4-92
 *   calculate_glszm  reference radiomics/src/cmatrices.c:94-279  (+ fill_glszm :281-297)
 *   calculate_glrlm  reference radiomics/src/cmatrices.c:299-541
 *   calculate_ngtdm  reference radiomics/src/cmatrices.c:543-658
 *   calculate_gldm   reference radiomics/src/cmatrices.c:660-754
 *   get_angle_count / build_angles  reference radiomics/src/cmatrices.c:756-892
 *   set_bb + per-voxel driver       reference radiomics/src/_cmatrices.c:203-207,1120-1147
 * Parity pinned: tests/test_oracle.py checks this file against the gcc-built reference
 * (oracle/_ref) and against the reference's golden matrices (tests/golden/segment_cases.npz).
 *
 * Written as plain nested loops over an inclusive box [lo,hi]; the reference's flat-index
 * skipping arithmetic is not reproduced, only its results.
 */
#include <stdlib.h>
#include <string.h>

typedef struct { int lo[3], hi[3]; } box_t;

static inline size_t lin(const int size[3], int z, int y, int x) {
  return ((size_t)z * size[1] + y) * size[2] + x;
}
static inline int inbox(const box_t *b, int z, int y, int x) {
  return z >= b->lo[0] && z <= b->hi[0] && y >= b->lo[1] && y <= b->hi[1] && x >= b->lo[2] && x <= b->hi[2];
}

/* -- angles (cmatrices.c:756-892): offsets run +D..-D per dimension, dimension 0 slowest;
 *    an offset vector is kept when its Chebyshev norm is a requested distance, no component
 *    reaches the image size, and it does not move along force2Ddim.  The enumeration is
 *    point-symmetric, so the unidirectional set is its first half. */
int oracle_build_angles(const int size[3], const int *distances, int ndist, int bidirectional,
                        int force2Ddim, int *angles, int max_na) {
  int D = 0, na = 0, total = 0;
  for (int i = 0; i < ndist; i++) { if (distances[i] < 1) return 0; if (distances[i] > D) D = distances[i]; }
  for (int pass = 0; pass < 2; pass++) {
    int cnt = 0;
    for (int oz = D; oz >= -D; oz--) for (int oy = D; oy >= -D; oy--) for (int ox = D; ox >= -D; ox--) {
      int o[3] = {oz, oy, ox}, norm = 0, ok = 1;
      for (int d = 0; d < 3; d++) {
        int a = o[d] < 0 ? -o[d] : o[d];
        if (a >= size[d] || (d == force2Ddim && a != 0)) ok = 0;
        if (a > norm) norm = a;
      }
      if (!ok || norm == 0) continue;
      int wanted = 0;
      for (int i = 0; i < ndist; i++) if (distances[i] == norm) wanted = 1;
      if (!wanted) continue;
      if (pass == 1 && cnt < na) {
        if (cnt >= max_na) return -1;
        angles[cnt * 3] = oz; angles[cnt * 3 + 1] = oy; angles[cnt * 3 + 2] = ox;
      }
      cnt++;
    }
    if (pass == 0) { total = cnt; na = bidirectional ? total : total / 2; }
  }
  return na;
}

/* -- bounding box of one kernel (_cmatrices.c:1120-1147) */
static void kernel_box(const int size[3], const int *voxels, int nvox, int v, int radius,
                       int force2Ddim, box_t *b) {
  for (int d = 0; d < 3; d++) {
    if (!voxels) { b->lo[d] = 0; b->hi[d] = size[d] - 1; continue; }
    int c = voxels[(size_t)d * nvox + v];
    if (d == force2Ddim) { b->lo[d] = b->hi[d] = c; continue; }
    b->lo[d] = c - radius < 0 ? 0 : c - radius;
    b->hi[d] = c + radius >= size[d] ? size[d] - 1 : c + radius;
  }
}

/* ---- GLCM ------------------------------------------------------------------------ */
static int glcm_box(const int *img, const char *msk, const int size[3], const box_t *b,
                    const int *ang, int na, double *out, int ng) {
  for (int z = b->lo[0]; z <= b->hi[0]; z++) for (int y = b->lo[1]; y <= b->hi[1]; y++)
  for (int x = b->lo[2]; x <= b->hi[2]; x++) {
    size_t i = lin(size, z, y, x);
    if (!msk[i]) continue;
    for (int a = 0; a < na; a++) {
      int z2 = z + ang[a * 3], y2 = y + ang[a * 3 + 1], x2 = x + ang[a * 3 + 2];
      if (!inbox(b, z2, y2, x2)) continue;
      size_t j = lin(size, z2, y2, x2);
      if (!msk[j]) continue;
      int gi = img[i], gj = img[j];
      if (gi <= 0 || gj <= 0 || gi > ng || gj > ng) return 0;
      out[((size_t)(gi - 1) * ng + (gj - 1)) * na + a] += 1.0;
    }
  }
  return 1;
}

/* ---- GLDM ------------------------------------------------------------------------ */
static int gldm_box(const int *img, const char *msk, const int size[3], const box_t *b,
                    const int *ang, int na, double *out, int ng, int alpha) {
  int ncol = 2 * na + 1;  /* over-allocated exactly as the reference does (_cmatrices.c:790) */
  for (int z = b->lo[0]; z <= b->hi[0]; z++) for (int y = b->lo[1]; y <= b->hi[1]; y++)
  for (int x = b->lo[2]; x <= b->hi[2]; x++) {
    size_t i = lin(size, z, y, x);
    if (!msk[i]) continue;
    int dep = 0, gi = img[i];
    for (int a = 0; a < na; a++) {
      int z2 = z + ang[a * 3], y2 = y + ang[a * 3 + 1], x2 = x + ang[a * 3 + 2];
      if (!inbox(b, z2, y2, x2)) continue;
      size_t j = lin(size, z2, y2, x2);
      if (!msk[j]) continue;
      int d = gi - img[j];
      if (d < 0) d = -d;
      if (d <= alpha) dep++;
    }
    if (gi <= 0 || gi > ng) return 0;
    out[(size_t)(gi - 1) * ncol + dep] += 1.0;
  }
  return 1;
}

/* ---- NGTDM ----------------------------------------------------------------------- */
static int ngtdm_box(const int *img, const char *msk, const int size[3], const box_t *b,
                     const int *ang, int na, double *out, int ng) {
  for (int g = 0; g < ng; g++) out[g * 3 + 2] = g + 1;
  for (int z = b->lo[0]; z <= b->hi[0]; z++) for (int y = b->lo[1]; y <= b->hi[1]; y++)
  for (int x = b->lo[2]; x <= b->hi[2]; x++) {
    size_t i = lin(size, z, y, x);
    if (!msk[i]) continue;
    double cnt = 0, sum = 0;
    for (int a = 0; a < na; a++) {
      int z2 = z + ang[a * 3], y2 = y + ang[a * 3 + 1], x2 = x + ang[a * 3 + 2];
      if (!inbox(b, z2, y2, x2)) continue;
      size_t j = lin(size, z2, y2, x2);
      if (!msk[j]) continue;
      cnt += 1; sum += img[j];
    }
    double diff = cnt == 0 ? 0.0 : (double)img[i] - sum / cnt;
    if (diff < 0) diff = -diff;
    if (img[i] <= 0 || img[i] > ng) return 0;
    out[(img[i] - 1) * 3] += 1.0;
    out[(img[i] - 1) * 3 + 1] += diff;
  }
  return 1;
}

/* ---- GLRLM ----------------------------------------------------------------------- */
static int glrlm_box(const int *img, const char *msk, const int size[3], const box_t *b,
                     const int *ang, int na, double *out, int ng, int nr) {
  for (int a = 0; a < na; a++) {
    const int *o = ang + a * 3;
    int multi = 0;
    for (int z = b->lo[0]; z <= b->hi[0]; z++) for (int y = b->lo[1]; y <= b->hi[1]; y++)
    for (int x = b->lo[2]; x <= b->hi[2]; x++) {
      /* a line starts where stepping backwards leaves the box */
      if (inbox(b, z - o[0], y - o[1], x - o[2])) continue;
      int cz = z, cy = y, cx = x, gl = -1, rl = 0, elements = 0;
      while (inbox(b, cz, cy, cx)) {
        size_t j = lin(size, cz, cy, cx);
        if (msk[j]) {
          elements++;
          if (gl == -1) { gl = img[j]; rl = 0; }
          else if (img[j] == gl) rl++;
          else {
            if (gl <= 0 || gl > ng || rl >= nr) return 0;
            out[((size_t)(gl - 1) * nr + rl) * na + a] += 1.0;
            gl = img[j]; rl = 0;
          }
        } else if (gl != -1) {
          if (gl <= 0 || gl > ng || rl >= nr) return 0;
          out[((size_t)(gl - 1) * nr + rl) * na + a] += 1.0;
          gl = -1; rl = 0;
        }
        cz += o[0]; cy += o[1]; cx += o[2];
      }
      if (gl != -1) {
        if (gl <= 0 || gl > ng || rl >= nr) return 0;
        out[((size_t)(gl - 1) * nr + rl) * na + a] += 1.0;
      }
      if (elements > 1) multi = 1;
    }
    if (!multi) for (int g = 0; g < ng; g++) out[((size_t)g * nr) * na + a] = 0.0;
  }
  return 1;
}

/* ---- GLSZM: zones appended to zones[2*k]=gray, zones[2*k+1]=size; returns count or -1 --- */
static long glszm_box(const int *img, char *msk, const int size[3], const box_t *b,
                      const int *ang, int na, int *zones, long zcap, size_t *stack, size_t *touched,
                      int restore, int *max_region) {
  long nz = 0; size_t nt = 0;
  for (int z = b->lo[0]; z <= b->hi[0]; z++) for (int y = b->lo[1]; y <= b->hi[1]; y++)
  for (int x = b->lo[2]; x <= b->hi[2]; x++) {
    size_t i = lin(size, z, y, x);
    if (!msk[i]) continue;
    int gl = img[i], region = 0; size_t top = 0;
    stack[top++] = i; msk[i] = 0; touched[nt++] = i;
    while (top) {
      size_t k = stack[--top];
      region++;
      int kz = (int)(k / ((size_t)size[1] * size[2])), ky = (int)((k / size[2]) % size[1]), kx = (int)(k % size[2]);
      for (int a = 0; a < na; a++) {
        int z2 = kz + ang[a * 3], y2 = ky + ang[a * 3 + 1], x2 = kx + ang[a * 3 + 2];
        if (!inbox(b, z2, y2, x2)) continue;
        size_t j = lin(size, z2, y2, x2);
        if (msk[j] && img[j] == gl) { stack[top++] = j; msk[j] = 0; touched[nt++] = j; }
      }
    }
    if (nz >= zcap) return -1;
    zones[2 * nz] = gl; zones[2 * nz + 1] = region; nz++;
    if (region > *max_region) *max_region = region;
  }
  if (restore) while (nt) msk[touched[--nt]] = 1;
  return nz;
}

/* =========================== exported drivers ===================================== */
/* voxels == NULL -> segment mode (nvox must be 1).  Output layouts are the reference's:
 * glcm [nvox][ng][ng][na], glrlm [nvox][ng][nr][na], gldm [nvox][ng][2*na+1],
 * ngtdm [nvox][ng][3] (zero-filled by the caller).  Return 1 ok, 0 = gray level out of range
 * (the reference raises IndexError there). */
int oracle_glcm(const int *img, const char *msk, const int size[3], const int *ang, int na, int ng,
                const int *voxels, int nvox, int radius, int force2Ddim, double *out) {
  box_t b;
  for (int v = 0; v < nvox; v++) {
    kernel_box(size, voxels, nvox, v, radius, force2Ddim, &b);
    if (!glcm_box(img, msk, size, &b, ang, na, out + (size_t)v * ng * ng * na, ng)) return 0;
  }
  return 1;
}
int oracle_gldm(const int *img, const char *msk, const int size[3], const int *ang, int na, int ng, int alpha,
                const int *voxels, int nvox, int radius, int force2Ddim, double *out) {
  box_t b;
  for (int v = 0; v < nvox; v++) {
    kernel_box(size, voxels, nvox, v, radius, force2Ddim, &b);
    if (!gldm_box(img, msk, size, &b, ang, na, out + (size_t)v * ng * (2 * na + 1), ng, alpha)) return 0;
  }
  return 1;
}
int oracle_ngtdm(const int *img, const char *msk, const int size[3], const int *ang, int na, int ng,
                 const int *voxels, int nvox, int radius, int force2Ddim, double *out) {
  box_t b;
  for (int v = 0; v < nvox; v++) {
    kernel_box(size, voxels, nvox, v, radius, force2Ddim, &b);
    if (!ngtdm_box(img, msk, size, &b, ang, na, out + (size_t)v * ng * 3, ng)) return 0;
  }
  return 1;
}
int oracle_glrlm(const int *img, const char *msk, const int size[3], const int *ang, int na, int ng, int nr,
                 const int *voxels, int nvox, int radius, int force2Ddim, double *out) {
  box_t b;
  for (int v = 0; v < nvox; v++) {
    kernel_box(size, voxels, nvox, v, radius, force2Ddim, &b);
    if (!glrlm_box(img, msk, size, &b, ang, na, out + (size_t)v * ng * nr * na, ng, nr)) return 0;
  }
  return 1;
}
/* GLSZM phase 1: zone list per kernel.  zone_offsets[nvox+1] (prefix), zones [2*total].
 * Returns max region (>=0) or -1 on error; *zones_out malloc'ed (free with oracle_free). */
int oracle_glszm_zones(const int *img, const char *msk_in, const int size[3], const int *ang, int na,
                       const int *voxels, int nvox, int radius, int force2Ddim,
                       long *zone_offsets, int **zones_out) {
  size_t n = (size_t)size[0] * size[1] * size[2];
  char *msk = (char *)malloc(n); memcpy(msk, msk_in, n);
  size_t *stack = (size_t *)malloc(sizeof(size_t) * n), *touched = (size_t *)malloc(sizeof(size_t) * n);
  long cap = 1024, total = 0; int *zones = (int *)malloc(sizeof(int) * 2 * cap);
  int max_region = 0; box_t b;
  zone_offsets[0] = 0;
  for (int v = 0; v < nvox; v++) {
    kernel_box(size, voxels, nvox, v, radius, force2Ddim, &b);
    long bvol = (long)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
    if (total + bvol > cap) { while (total + bvol > cap) cap *= 2; zones = (int *)realloc(zones, sizeof(int) * 2 * cap); }
    long nz = glszm_box(img, msk, size, &b, ang, na, zones + 2 * total, cap - total, stack, touched, voxels != NULL, &max_region);
    if (nz < 0) { free(msk); free(stack); free(touched); free(zones); return -1; }
    total += nz; zone_offsets[v + 1] = total;
  }
  free(msk); free(stack); free(touched);
  *zones_out = zones;
  return max_region;
}
/* GLSZM phase 2 (fill_glszm): out [nvox][ng][max_region], zero-filled by the caller. */
int oracle_glszm_fill(const int *zones, const long *zone_offsets, int nvox, int ng, int max_region, double *out) {
  for (int v = 0; v < nvox; v++)
    for (long k = zone_offsets[v]; k < zone_offsets[v + 1]; k++) {
      int g = zones[2 * k], s = zones[2 * k + 1];
      if (g <= 0 || g > ng || s > max_region) return 0;
      out[((size_t)v * ng + (g - 1)) * max_region + (s - 1)] += 1.0;
    }
  return 1;
}
void oracle_free(void *p) { free(p); }
