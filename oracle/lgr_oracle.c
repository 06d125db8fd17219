/* ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the differentiable Gaussian-splatting rasteriser path that LoG calls through
 * `GaussianRasterizer.__call__` (LoG/render/renderer.py:141-153, LoG/model/level_of_gaussian.py:211) and of
 * `compute_radius` (LoG/cuda/compute_radius_kernel.cu:107-156).  Tile-based, forward + analytic backward,
 * OpenMP over Gaussians / tiles.  Compiled twice: -DREAL=double (lgo64_*) and -DREAL=float (lgo32_*).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load this.
 * The product (log_b200/) never links, loads or calls it.
 *
 * PARITY STATUS
 *   pinned   : rotation / cov3D / EWA cov2D / radius  -> LoG/model/geometry.py:4-41,91-151 and
 *              LoG/cuda/compute_radius_kernel.cu:28-156, checked against golden vectors made by running the
 *              reference (tests/golden/make_golden.py); SH basis -> LoG/model/sh_utils.py:31-68, same.
 *   UNPINNED : the blend (tile rule, depth order, alpha rule, early stop) and its backward.  Their source is the
 *              un-vendored diff_gaussian_rasterization[_wodilate] (docs/install.md:37-43, no commit pinned), absent
 *              from /root/reference.  What is written here restates the published 3DGS algorithm and is validated
 *              against the autograd definition in oracle/torch_dense.py (same constants), not against LoG binaries.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL double
#endif
#ifndef PFX
#define PFX lgo64_
#endif
#ifdef REAL_IS_FLOAT
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#else
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(PFX, name)

#define TILE 16
#define NEAR_Z ((REAL)0.2)
#define ALPHA_MAX ((REAL)0.99)
#define ALPHA_MIN ((REAL)(1.0 / 255.0))
#define T_STOP ((REAL)1e-4)
#define FILTER_VAR ((REAL)0.3) /* compute_radius_kernel.cu:61 */
#define CLAMP_FOV ((REAL)1.3)  /* compute_radius_kernel.cu:71-72 */

enum { FILTER_ADD = 0, FILTER_MAX = 1, FILTER_NONE = 2 };

static const REAL SH_C0 = (REAL)0.28209479177387814;
static const REAL SH_C1 = (REAL)0.4886025119029199;
static const REAL SH_C2[5] = {(REAL)1.0925484305920792, (REAL)-1.0925484305920792, (REAL)0.31539156525252005,
                              (REAL)-1.0925484305920792, (REAL)0.5462742152960396};
static const REAL SH_C3[7] = {(REAL)-0.5900435899266435, (REAL)2.890611442640554, (REAL)-0.4570457994644658,
                              (REAL)0.3731763325901154, (REAL)-0.4570457994644658, (REAL)1.445305721320277,
                              (REAL)-0.5900435899266435};

typedef struct {
  int32_t image_height, image_width;
  REAL tanfovx, tanfovy;
  REAL viewmatrix[16]; /* world_view_transform, row-major memory of the TRANSPOSED matrix (dataset/base.py:40-46) */
  REAL projmatrix[16]; /* full_proj_transform, same convention */
  REAL campos[3];
  REAL bg[3];
  REAL scale_modifier;
  int32_t sh_degree; /* active degree 0..3 */
  int32_t sh_K;      /* coefficients per Gaussian in `shs` (>= (deg+1)^2) */
  int32_t filter_mode;
} FN(camera);

typedef struct {
  REAL x, y, depth;
  REAL con_x, con_y, con_z, opacity;
  REAL rgb[3];
  /* kept for the backward pass */
  REAL a_raw, c_raw, a, b, c; /* cov2D before / after the filter */
  REAL t[3];
  int inx, iny;
  int clamped[3];
  int radius;
  int x0, y0, x1, y1;
  int valid;
} splat_t;

static inline REAL rmin(REAL a, REAL b) { return a < b ? a : b; }
static inline REAL rmax(REAL a, REAL b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static void quat_to_R(const REAL* q, REAL R[9]) {
  /* geometry.py:12-24 without the normalisation (compute_radius_kernel.cu:36) */
  REAL r = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - r * z); R[2] = 2 * (x * z + r * y);
  R[3] = 2 * (x * y + r * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - r * x);
  R[6] = 2 * (x * z - r * y); R[7] = 2 * (y * z + r * x); R[8] = 1 - 2 * (x * x + y * y);
}

/* Sigma = (R S)(R S)^T, geometry.py:27-41 */
static void cov3d(const REAL* s, REAL mod, const REAL* q, REAL Sg[9]) {
  REAL R[9], M[9];
  quat_to_R(q, R);
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) M[i * 3 + k] = R[i * 3 + k] * (s[k] * mod);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      REAL v = 0;
      for (int k = 0; k < 3; k++) v += M[i * 3 + k] * M[j * 3 + k];
      Sg[i * 3 + j] = v;
    }
}

/* EWA: geometry.py:91-130 / compute_radius_kernel.cu:63-105.  T (2x3) returned for the backward. */
static void cov2d(const FN(camera) * cam, const REAL* p, const REAL Sg[9], REAL t[3], int* inx, int* iny, REAL Tm[6],
                  REAL* a, REAL* b, REAL* c) {
  const REAL* V = cam->viewmatrix;
  for (int j = 0; j < 3; j++) t[j] = p[0] * V[0 * 4 + j] + p[1] * V[1 * 4 + j] + p[2] * V[2 * 4 + j] + V[3 * 4 + j];
  const REAL fx = cam->image_width / (2 * cam->tanfovx), fy = cam->image_height / (2 * cam->tanfovy);
  const REAL limx = CLAMP_FOV * cam->tanfovx, limy = CLAMP_FOV * cam->tanfovy;
  const REAL txtz = t[0] / t[2], tytz = t[1] / t[2];
  *inx = (txtz >= -limx) && (txtz <= limx);
  *iny = (tytz >= -limy) && (tytz <= limy);
  const REAL txc = rmin(limx, rmax(-limx, txtz)) * t[2];
  const REAL tyc = rmin(limy, rmax(-limy, tytz)) * t[2];
  REAL J[6] = {fx / t[2], 0, -(fx * txc) / (t[2] * t[2]), 0, fy / t[2], -(fy * tyc) / (t[2] * t[2])};
  /* W (math, world->view) = V[:3,:3]^T ; T = J W :  T[r][j] = sum_k J[r][k] * V[j][k] */
  for (int r = 0; r < 2; r++)
    for (int j = 0; j < 3; j++) {
      REAL v = 0;
      for (int k = 0; k < 3; k++) v += J[r * 3 + k] * V[j * 4 + k];
      Tm[r * 3 + j] = v;
    }
  REAL TS[6];
  for (int r = 0; r < 2; r++)
    for (int j = 0; j < 3; j++) {
      REAL v = 0;
      for (int k = 0; k < 3; k++) v += Tm[r * 3 + k] * Sg[k * 3 + j];
      TS[r * 3 + j] = v;
    }
  *a = TS[0] * Tm[0] + TS[1] * Tm[1] + TS[2] * Tm[2];
  *b = TS[0] * Tm[3] + TS[1] * Tm[4] + TS[2] * Tm[5];
  *c = TS[3] * Tm[3] + TS[4] * Tm[4] + TS[5] * Tm[5];
}

static void apply_filter(int mode, REAL* a, REAL* c) {
  if (mode == FILTER_ADD) { *a += FILTER_VAR; *c += FILTER_VAR; }
  else if (mode == FILTER_MAX) { *a = rmax(*a, FILTER_VAR); *c = rmax(*c, FILTER_VAR); }
}

static REAL radius_from_cov(REAL a, REAL b, REAL c, REAL* det_out) {
  /* compute_radius_kernel.cu:139-152 */
  REAL det = a * c - b * b;
  REAL mid = (REAL)0.5 * (a + c);
  REAL root = R_SQRT(rmax((REAL)0.1, mid * mid - det));
  REAL lam = rmax(mid + root, mid - root);
  *det_out = det;
  return 3 * R_SQRT(lam);
}

/* SH basis values for degree<=3 (16 entries); basis[0] multiplies the DC term.  sh_utils.py:31-58 */
static void sh_basis(int deg, const REAL d[3], REAL B[16]) {
  REAL x = d[0], y = d[1], z = d[2];
  B[0] = SH_C0;
  if (deg > 0) {
    B[1] = -SH_C1 * y; B[2] = SH_C1 * z; B[3] = -SH_C1 * x;
    if (deg > 1) {
      REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      B[4] = SH_C2[0] * xy; B[5] = SH_C2[1] * yz; B[6] = SH_C2[2] * (2 * zz - xx - yy);
      B[7] = SH_C2[3] * xz; B[8] = SH_C2[4] * (xx - yy);
      if (deg > 2) {
        B[9] = SH_C3[0] * y * (3 * xx - yy); B[10] = SH_C3[1] * xy * z; B[11] = SH_C3[2] * y * (4 * zz - xx - yy);
        B[12] = SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy); B[13] = SH_C3[4] * x * (4 * zz - xx - yy);
        B[14] = SH_C3[5] * z * (xx - yy); B[15] = SH_C3[6] * x * (xx - 3 * yy);
      }
    }
  }
}

/* d(basis_k)/d(dir) for k<16 */
static void sh_basis_grad(int deg, const REAL d[3], REAL G[16][3]) {
  REAL x = d[0], y = d[1], z = d[2];
  memset(G, 0, sizeof(REAL) * 48);
  if (deg > 0) {
    G[1][1] = -SH_C1; G[2][2] = SH_C1; G[3][0] = -SH_C1;
    if (deg > 1) {
      REAL xx = x * x, yy = y * y, zz = z * z;
      G[4][0] = SH_C2[0] * y; G[4][1] = SH_C2[0] * x;
      G[5][1] = SH_C2[1] * z; G[5][2] = SH_C2[1] * y;
      G[6][0] = SH_C2[2] * -2 * x; G[6][1] = SH_C2[2] * -2 * y; G[6][2] = SH_C2[2] * 4 * z;
      G[7][0] = SH_C2[3] * z; G[7][2] = SH_C2[3] * x;
      G[8][0] = SH_C2[4] * 2 * x; G[8][1] = SH_C2[4] * -2 * y;
      if (deg > 2) {
        G[9][0] = SH_C3[0] * 6 * x * y; G[9][1] = SH_C3[0] * (3 * xx - 3 * yy);
        G[10][0] = SH_C3[1] * y * z; G[10][1] = SH_C3[1] * x * z; G[10][2] = SH_C3[1] * x * y;
        G[11][0] = SH_C3[2] * -2 * x * y; G[11][1] = SH_C3[2] * (4 * zz - xx - 3 * yy); G[11][2] = SH_C3[2] * 8 * y * z;
        G[12][0] = SH_C3[3] * -6 * x * z; G[12][1] = SH_C3[3] * -6 * y * z; G[12][2] = SH_C3[3] * (6 * zz - 3 * xx - 3 * yy);
        G[13][0] = SH_C3[4] * (4 * zz - 3 * xx - yy); G[13][1] = SH_C3[4] * -2 * x * y; G[13][2] = SH_C3[4] * 8 * x * z;
        G[14][0] = SH_C3[5] * 2 * x * z; G[14][1] = SH_C3[5] * -2 * y * z; G[14][2] = SH_C3[5] * (xx - yy);
        G[15][0] = SH_C3[6] * (3 * xx - 3 * yy); G[15][1] = SH_C3[6] * -6 * x * y;
      }
    }
  }
}

/* ---- compute_radius: LoG/cuda/compute_radius_kernel.cu:107-156, exactly (NDC cull, no near cull, max filter) */
void FN(compute_radius)(const FN(camera) * cam, int64_t N, const REAL* means3D, const REAL* scales,
                        const REAL* rotations, REAL* radii) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; i++) {
    radii[i] = 0;
    const REAL* p = means3D + 3 * i;
    const REAL* P = cam->projmatrix;
    REAL hom[4];
    for (int k = 0; k < 4; k++) hom[k] = p[0] * P[0 * 4 + k] + p[1] * P[1 * 4 + k] + p[2] * P[2 * 4 + k] + P[3 * 4 + k];
    REAL pw = 1 / (hom[3] + (REAL)0.0000001);
    REAL px = hom[0] * pw, py = hom[1] * pw;
    if (px < (REAL)-1.3 || px > (REAL)1.3 || py < (REAL)-1.3 || py > (REAL)1.3) continue;
    REAL Sg[9], t[3], Tm[6], a, b, c, det;
    int inx, iny;
    cov3d(scales + 3 * i, 1, rotations + 4 * i, Sg);
    cov2d(cam, p, Sg, t, &inx, &iny, Tm, &a, &b, &c);
    apply_filter(FILTER_MAX, &a, &c);
    REAL rad = radius_from_cov(a, b, c, &det);
    if (det == 0) continue;
    radii[i] = rad;
  }
}

typedef struct { REAL depth; int32_t idx; } key_t;
static int key_cmp(const void* A, const void* B) {
  const key_t* a = (const key_t*)A; const key_t* b = (const key_t*)B;
  if (a->depth < b->depth) return -1;
  if (a->depth > b->depth) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);
}

/* Forward (+ backward when dL_dimage != NULL).  Pointers may be NULL for outputs that are not wanted.
 * Returns D, the number of (Gaussian, tile) instances, or -1 on allocation failure.
 * Inputs : means3D (N,3) opacities (N) scales (N,3) rotations (N,4) colors_precomp (N,3)|NULL shs (N,K,3)|NULL
 * Outputs: image (3,H,W) radii (N) point_id_pixel (H,W) point_weight_pixel (H,W) point_weight (N) final_T (H,W)
 * Grads  : dmeans3D (N,3) dmeans2D (N,3: d/d(ndc x,y), z = 0) dopacities (N) dscales (N,3) drotations (N,4)
 *          dcolors (N,3) dshs (N,K,3) */
int64_t FN(render)(const FN(camera) * cam, int64_t N, const REAL* means3D, const REAL* opacities, const REAL* scales,
                   const REAL* rotations, const REAL* colors_precomp, const REAL* shs, REAL* image, int32_t* radii,
                   int32_t* point_id_pixel, REAL* point_weight_pixel, REAL* point_weight, REAL* final_T,
                   const REAL* dL_dimage, REAL* dmeans3D, REAL* dmeans2D, REAL* dopacities, REAL* dscales,
                   REAL* drotations, REAL* dcolors, REAL* dshs) {
  const int H = cam->image_height, W = cam->image_width;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int64_t ntiles = (int64_t)gx * gy;
  splat_t* sp = (splat_t*)calloc((size_t)(N > 0 ? N : 1), sizeof(splat_t));
  int64_t* tile_start = (int64_t*)calloc((size_t)ntiles + 1, sizeof(int64_t));
  if (!sp || !tile_start) return -1;

  /* ---------------- stage 1: per-Gaussian projection ---------------- */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; i++) {
    splat_t* s = sp + i;
    s->valid = 0; s->radius = 0;
    const REAL* p = means3D + 3 * i;
    REAL Sg[9], Tm[6], a, b, c, det;
    cov3d(scales + 3 * i, cam->scale_modifier, rotations + 4 * i, Sg);
    cov2d(cam, p, Sg, s->t, &s->inx, &s->iny, Tm, &a, &b, &c);
    s->a_raw = a; s->c_raw = c;
    apply_filter(cam->filter_mode, &a, &c);
    s->a = a; s->b = b; s->c = c;
    if (!(s->t[2] > NEAR_Z)) continue; /* [B] near cull */
    REAL radf = radius_from_cov(a, b, c, &det);
    if (!(det > 0)) continue;
    const REAL* P = cam->projmatrix;
    REAL hom[4];
    for (int k = 0; k < 4; k++) hom[k] = p[0] * P[0 * 4 + k] + p[1] * P[1 * 4 + k] + p[2] * P[2 * 4 + k] + P[3 * 4 + k];
    REAL pw = 1 / (hom[3] + (REAL)0.0000001);
    s->x = ((hom[0] * pw + 1) * W - 1) * (REAL)0.5;
    s->y = ((hom[1] * pw + 1) * H - 1) * (REAL)0.5;
    s->depth = s->t[2];
    s->con_x = c / det; s->con_y = -b / det; s->con_z = a / det;
    s->opacity = opacities[i];
    int rad = (int)R_CEIL(radf);
    s->x0 = imin(gx, imax(0, (int)((s->x - rad) / TILE)));
    s->x1 = imin(gx, imax(0, (int)((s->x + rad + TILE - 1) / TILE)));
    s->y0 = imin(gy, imax(0, (int)((s->y - rad) / TILE)));
    s->y1 = imin(gy, imax(0, (int)((s->y + rad + TILE - 1) / TILE)));
    if ((s->x1 - s->x0) * (s->y1 - s->y0) == 0) continue;
    s->radius = rad; s->valid = 1;
    if (colors_precomp) {
      for (int ch = 0; ch < 3; ch++) { s->rgb[ch] = colors_precomp[3 * i + ch]; s->clamped[ch] = 0; }
    } else {
      REAL d[3] = {p[0] - cam->campos[0], p[1] - cam->campos[1], p[2] - cam->campos[2]};
      REAL n = R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      d[0] /= n; d[1] /= n; d[2] /= n;
      REAL B[16];
      sh_basis(cam->sh_degree, d, B);
      int nb = (cam->sh_degree + 1) * (cam->sh_degree + 1);
      for (int ch = 0; ch < 3; ch++) {
        REAL v = 0;
        for (int k = 0; k < nb; k++) v += B[k] * shs[((int64_t)i * cam->sh_K + k) * 3 + ch];
        v += (REAL)0.5;
        s->clamped[ch] = v < 0;     /* [B] stock clamps at 0; LoG's torch path (activation.py:27-34) does not */
        s->rgb[ch] = v < 0 ? 0 : v;
      }
    }
  }
  if (radii) for (int64_t i = 0; i < N; i++) radii[i] = sp[i].radius;

  /* ---------------- stage 2: tile binning in index order, per-tile (depth, index) sort ---------------- */
  for (int64_t i = 0; i < N; i++)
    if (sp[i].valid)
      for (int ty = sp[i].y0; ty < sp[i].y1; ty++)
        for (int tx = sp[i].x0; tx < sp[i].x1; tx++) tile_start[(int64_t)ty * gx + tx + 1]++;
  for (int64_t t = 0; t < ntiles; t++) tile_start[t + 1] += tile_start[t];
  const int64_t D = tile_start[ntiles];
  key_t* keys = (key_t*)malloc(sizeof(key_t) * (size_t)(D > 0 ? D : 1));
  int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)ntiles);
  if (!keys || !cursor) return -1;
  memcpy(cursor, tile_start, sizeof(int64_t) * (size_t)ntiles);
  for (int64_t i = 0; i < N; i++)
    if (sp[i].valid)
      for (int ty = sp[i].y0; ty < sp[i].y1; ty++)
        for (int tx = sp[i].x0; tx < sp[i].x1; tx++) {
          int64_t k = cursor[(int64_t)ty * gx + tx]++;
          keys[k].depth = sp[i].depth; keys[k].idx = (int32_t)i;
        }
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t t = 0; t < ntiles; t++)
    qsort(keys + tile_start[t], (size_t)(tile_start[t + 1] - tile_start[t]), sizeof(key_t), key_cmp);

  /* ---------------- stage 3: per-tile front-to-back blend ---------------- */
  int32_t* n_contrib = (int32_t*)calloc((size_t)H * W, sizeof(int32_t));
  REAL* Tfin = (REAL*)malloc(sizeof(REAL) * (size_t)H * W);
  REAL* iw = point_weight ? (REAL*)calloc((size_t)(D > 0 ? D : 1), sizeof(REAL)) : NULL; /* per-instance max weight */
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t t = 0; t < ntiles; t++) {
    const int tx = (int)(t % gx), ty = (int)(t / gx);
    const int64_t lo = tile_start[t], hi = tile_start[t + 1];
    for (int py = ty * TILE; py < imin(H, (ty + 1) * TILE); py++)
      for (int px = tx * TILE; px < imin(W, (tx + 1) * TILE); px++) {
        REAL T = 1, C[3] = {0, 0, 0}, wmax = 0;
        int32_t wid = -1, last = 0, cnt = 0;
        for (int64_t k = lo; k < hi; k++) {
          cnt++;
          const splat_t* s = sp + keys[k].idx;
          REAL dx = s->x - (REAL)px, dy = s->y - (REAL)py;
          REAL power = (REAL)-0.5 * (s->con_x * dx * dx + s->con_z * dy * dy) - s->con_y * dx * dy;
          if (power > 0) continue;
          REAL alpha = rmin(ALPHA_MAX, s->opacity * R_EXP(power));
          if (alpha < ALPHA_MIN) continue;
          REAL test_T = T * (1 - alpha);
          if (test_T < T_STOP) break;
          REAL w = alpha * T;
          for (int ch = 0; ch < 3; ch++) C[ch] += s->rgb[ch] * w;
          if (w > wmax) { wmax = w; wid = keys[k].idx; }
          if (iw && w > iw[k]) iw[k] = w;
          T = test_T;
          last = cnt;
        }
        const int64_t pix = (int64_t)py * W + px;
        n_contrib[pix] = last; Tfin[pix] = T;
        if (image) for (int ch = 0; ch < 3; ch++) image[(int64_t)ch * H * W + pix] = C[ch] + T * cam->bg[ch];
        if (final_T) final_T[pix] = T;
        if (point_id_pixel) point_id_pixel[pix] = wid;
        if (point_weight_pixel) point_weight_pixel[pix] = wmax;
      }
  }

  if (point_weight) { /* per-Gaussian max over its instances (sequential, deterministic) */
    memset(point_weight, 0, sizeof(REAL) * (size_t)N);
    for (int64_t k = 0; k < D; k++)
      if (iw[k] > point_weight[keys[k].idx]) point_weight[keys[k].idx] = iw[k];
    free(iw);
  }

  /* ---------------- stage 4/5: backward ---------------- */
  if (dL_dimage) {
    /* per-instance accumulators: dxy(2, pixel units) dconic(3, true derivative) dopacity(1) drgb(3) */
    REAL* ig = (REAL*)calloc((size_t)(D > 0 ? D : 1) * 9, sizeof(REAL));
    if (!ig) return -1;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t t = 0; t < ntiles; t++) {
      const int tx = (int)(t % gx), ty = (int)(t / gx);
      const int64_t lo = tile_start[t];
      for (int py = ty * TILE; py < imin(H, (ty + 1) * TILE); py++)
        for (int px = tx * TILE; px < imin(W, (tx + 1) * TILE); px++) {
          const int64_t pix = (int64_t)py * W + px;
          const REAL T_final = Tfin[pix];
          REAL T = T_final, dpix[3], accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0, bgdot = 0;
          for (int ch = 0; ch < 3; ch++) { dpix[ch] = dL_dimage[(int64_t)ch * H * W + pix]; bgdot += cam->bg[ch] * dpix[ch]; }
          for (int64_t k = lo + n_contrib[pix] - 1; k >= lo; k--) {
            const splat_t* s = sp + keys[k].idx;
            REAL dx = s->x - (REAL)px, dy = s->y - (REAL)py;
            REAL power = (REAL)-0.5 * (s->con_x * dx * dx + s->con_z * dy * dy) - s->con_y * dx * dy;
            if (power > 0) continue;
            REAL G = R_EXP(power);
            REAL alpha = rmin(ALPHA_MAX, s->opacity * G);
            if (alpha < ALPHA_MIN) continue;
            T = T / (1 - alpha);
            REAL dL_dalpha = 0;
            REAL* g = ig + k * 9;
            for (int ch = 0; ch < 3; ch++) {
              accum[ch] = last_alpha * last_color[ch] + (1 - last_alpha) * accum[ch];
              last_color[ch] = s->rgb[ch];
              dL_dalpha += (s->rgb[ch] - accum[ch]) * dpix[ch];
              g[6 + ch] += alpha * T * dpix[ch];
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1 - alpha)) * bgdot;
            /* [B] the 0.99 clamp is straight-through in the published backward */
            REAL dL_dG = s->opacity * dL_dalpha;
            REAL gdx = G * dx, gdy = G * dy;
            REAL dG_ddelx = -gdx * s->con_x - gdy * s->con_y;
            REAL dG_ddely = -gdy * s->con_z - gdx * s->con_y;
            g[0] += dL_dG * dG_ddelx;
            g[1] += dL_dG * dG_ddely;
            g[2] += (REAL)-0.5 * gdx * dx * dL_dG;
            g[3] += -gdx * dy * dL_dG;
            g[4] += (REAL)-0.5 * gdy * dy * dL_dG;
            g[5] += G * dL_dalpha;
          }
        }
    }
    /* reduce instances -> Gaussians (sequential, deterministic) */
    REAL* sg = (REAL*)calloc((size_t)(N > 0 ? N : 1) * 9, sizeof(REAL));
    if (!sg) return -1;
    for (int64_t k = 0; k < D; k++) {
      REAL* dst = sg + (int64_t)keys[k].idx * 9;
      for (int j = 0; j < 9; j++) dst[j] += ig[k * 9 + j];
    }
    free(ig);
    const int K = cam->sh_K;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) {
      const splat_t* s = sp + i;
      REAL* dm = dmeans3D + 3 * i;
      dm[0] = dm[1] = dm[2] = 0;
      if (dmeans2D) dmeans2D[3 * i] = dmeans2D[3 * i + 1] = dmeans2D[3 * i + 2] = 0;
      if (dopacities) dopacities[i] = 0;
      for (int j = 0; j < 3; j++) dscales[3 * i + j] = 0;
      for (int j = 0; j < 4; j++) drotations[4 * i + j] = 0;
      if (dcolors) for (int j = 0; j < 3; j++) dcolors[3 * i + j] = 0;
      if (dshs) for (int j = 0; j < K * 3; j++) dshs[(int64_t)i * K * 3 + j] = 0;
      if (!s->valid) continue;
      const REAL* g = sg + i * 9;
      const REAL* p = means3D + 3 * i;
      /* --- colour --- */
      if (dopacities) dopacities[i] = g[5];
      if (colors_precomp) {
        if (dcolors) for (int ch = 0; ch < 3; ch++) dcolors[3 * i + ch] = g[6 + ch];
      } else {
        REAL v[3] = {p[0] - cam->campos[0], p[1] - cam->campos[1], p[2] - cam->campos[2]};
        REAL n = R_SQRT(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        REAL d[3] = {v[0] / n, v[1] / n, v[2] / n};
        REAL B[16], BG[16][3];
        sh_basis(cam->sh_degree, d, B);
        sh_basis_grad(cam->sh_degree, d, BG);
        int nb = (cam->sh_degree + 1) * (cam->sh_degree + 1);
        REAL ddir[3] = {0, 0, 0};
        for (int ch = 0; ch < 3; ch++) {
          REAL dres = s->clamped[ch] ? 0 : g[6 + ch];
          for (int k = 0; k < nb; k++) {
            REAL coef = shs[((int64_t)i * K + k) * 3 + ch];
            if (dshs) dshs[((int64_t)i * K + k) * 3 + ch] = B[k] * dres;
            for (int a = 0; a < 3; a++) ddir[a] += BG[k][a] * coef * dres;
          }
        }
        REAL dot = d[0] * ddir[0] + d[1] * ddir[1] + d[2] * ddir[2];
        for (int a = 0; a < 3; a++) dm[a] += (ddir[a] - d[a] * dot) / n;
      }
      /* --- conic -> cov2D --- */
      const REAL a = s->a, b = s->b, c = s->c;
      const REAL det = a * c - b * b, idet2 = 1 / (det * det);
      REAL da = idet2 * (-c * c * g[2] + b * c * g[3] + (det - a * c) * g[4]);
      REAL dc = idet2 * (-a * a * g[4] + a * b * g[3] + (det - a * c) * g[2]);
      REAL db = idet2 * (2 * b * c * g[2] - (det + 2 * b * b) * g[3] + 2 * a * b * g[4]);
      if (cam->filter_mode == FILTER_MAX) {
        if (!(s->a_raw >= FILTER_VAR)) da = 0;
        if (!(s->c_raw >= FILTER_VAR)) dc = 0;
      }
      /* --- cov2D = T Sigma T^T --- */
      REAL Sg[9], R[9], Tm[6], t[3], aa, bb, cc;
      int inx, iny;
      cov3d(scales + 3 * i, cam->scale_modifier, rotations + 4 * i, Sg);
      cov2d(cam, p, Sg, t, &inx, &iny, Tm, &aa, &bb, &cc);
      quat_to_R(rotations + 4 * i, R);
      REAL Gm[4] = {da, (REAL)0.5 * db, (REAL)0.5 * db, dc};
      /* dSigma = T^T G T (3x3 symmetric) */
      REAL GT[6];
      for (int r = 0; r < 2; r++)
        for (int j = 0; j < 3; j++) GT[r * 3 + j] = Gm[r * 2 + 0] * Tm[0 * 3 + j] + Gm[r * 2 + 1] * Tm[1 * 3 + j];
      REAL dS[9];
      for (int i2 = 0; i2 < 3; i2++)
        for (int j = 0; j < 3; j++) dS[i2 * 3 + j] = Tm[0 * 3 + i2] * GT[0 * 3 + j] + Tm[1 * 3 + i2] * GT[1 * 3 + j];
      /* Sigma = M M^T, M = R diag(s*mod):  dM = 2 dS M */
      const REAL mod = cam->scale_modifier;
      REAL M[9], dM[9];
      for (int i2 = 0; i2 < 3; i2++)
        for (int k = 0; k < 3; k++) M[i2 * 3 + k] = R[i2 * 3 + k] * scales[3 * i + k] * mod;
      for (int i2 = 0; i2 < 3; i2++)
        for (int k = 0; k < 3; k++) {
          REAL v = 0;
          for (int j = 0; j < 3; j++) v += dS[i2 * 3 + j] * M[j * 3 + k];
          dM[i2 * 3 + k] = 2 * v;
        }
      REAL dR[9];
      for (int k = 0; k < 3; k++) {
        REAL v = 0;
        for (int i2 = 0; i2 < 3; i2++) { v += dM[i2 * 3 + k] * R[i2 * 3 + k]; dR[i2 * 3 + k] = dM[i2 * 3 + k] * scales[3 * i + k] * mod; }
        dscales[3 * i + k] = v * mod;
      }
      {
        const REAL* q = rotations + 4 * i;
        REAL r = q[0], x = q[1], y = q[2], z = q[3];
        drotations[4 * i + 0] = 2 * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
        drotations[4 * i + 1] = 2 * (y * dR[1] + z * dR[2] + y * dR[3] - 2 * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2 * x * dR[8]);
        drotations[4 * i + 2] = 2 * (-2 * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2 * y * dR[8]);
        drotations[4 * i + 3] = 2 * (-2 * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2 * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
      }
      /* dT = 2 G T Sigma ; dJ = dT W^T, W = V[:3,:3]^T  ->  dJ[r][k] = sum_j dT[r][j] * V[j][k] */
      const REAL* V = cam->viewmatrix;
      REAL TS[6], dT[6], dJ[6];
      for (int r = 0; r < 2; r++)
        for (int j = 0; j < 3; j++) {
          REAL v = 0;
          for (int k = 0; k < 3; k++) v += Tm[r * 3 + k] * Sg[k * 3 + j];
          TS[r * 3 + j] = v;
        }
      for (int r = 0; r < 2; r++)
        for (int j = 0; j < 3; j++) dT[r * 3 + j] = 2 * (Gm[r * 2 + 0] * TS[0 * 3 + j] + Gm[r * 2 + 1] * TS[1 * 3 + j]);
      for (int r = 0; r < 2; r++)
        for (int k = 0; k < 3; k++) {
          REAL v = 0;
          for (int j = 0; j < 3; j++) v += dT[r * 3 + j] * V[j * 4 + k];
          dJ[r * 3 + k] = v;
        }
      const REAL fx = W / (2 * cam->tanfovx), fy = H / (2 * cam->tanfovy);
      const REAL limx = CLAMP_FOV * cam->tanfovx, limy = CLAMP_FOV * cam->tanfovy;
      const REAL tz = t[2], itz = 1 / tz, itz2 = itz * itz, itz3 = itz2 * itz;
      const REAL txc = rmin(limx, rmax(-limx, t[0] / tz)) * tz, tyc = rmin(limy, rmax(-limy, t[1] / tz)) * tz;
      REAL dt[3];
      dt[0] = inx ? -fx * itz2 * dJ[2] : 0;
      dt[1] = iny ? -fy * itz2 * dJ[5] : 0;
      dt[2] = -fx * itz2 * dJ[0] - fy * itz2 * dJ[4] + 2 * fx * txc * itz3 * dJ[2] + 2 * fy * tyc * itz3 * dJ[5];
      for (int i2 = 0; i2 < 3; i2++) dm[i2] += V[i2 * 4 + 0] * dt[0] + V[i2 * 4 + 1] * dt[1] + V[i2 * 4 + 2] * dt[2];
      /* --- mean2D (ndc) path --- */
      const REAL* P = cam->projmatrix;
      REAL hom[4];
      for (int k = 0; k < 4; k++) hom[k] = p[0] * P[0 * 4 + k] + p[1] * P[1 * 4 + k] + p[2] * P[2 * 4 + k] + P[3 * 4 + k];
      REAL pw = 1 / (hom[3] + (REAL)0.0000001);
      REAL gnx = g[0] * (REAL)0.5 * W, gny = g[1] * (REAL)0.5 * H;
      if (dmeans2D) { dmeans2D[3 * i] = gnx; dmeans2D[3 * i + 1] = gny; }
      REAL dh0 = gnx * pw, dh1 = gny * pw, dh3 = -(gnx * hom[0] + gny * hom[1]) * pw * pw;
      for (int i2 = 0; i2 < 3; i2++) dm[i2] += P[i2 * 4 + 0] * dh0 + P[i2 * 4 + 1] * dh1 + P[i2 * 4 + 3] * dh3;
    }
    free(sg);
  }
  free(n_contrib); free(Tfin); free(keys); free(cursor); free(tile_start); free(sp);
  return D;
}

int FN(num_threads)(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void FN(set_num_threads)(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
