/*
 * moog_oracle.c -- CPU restatement of MOOG's Environment.step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under moog.github.io_b200/ may link, load
 * or call this file; it is the checker the CUDA path is compared against
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).
 *
 * It is a plain scalar C restatement, one env at a time, float64 throughout,
 * of the reference's Python algorithm.  Each function cites the reference
 * file:line it follows (paths relative to the reference repo root).  The
 * third-party pieces the reference reaches into -- matplotlib 3.10 `_path`
 * (absent from the image) and Pillow's Draw.c -- are restated from their
 * published algorithms; see oracle/README.md for how each is pinned.
 *
 * Pinning: tests/test_oracle_vs_golden.py checks this file against golden
 * vectors produced by running the unmodified reference (oracle/gen_golden.py,
 * through oracle/shims) and, for rendering, against the installed Pillow.
 *
 * Build: make -C oracle      (gcc -O2 -ffp-contract=off; no FMA contraction so
 * that intermediate roundings follow the reference's numpy arithmetic).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/moog_b200_program.h"

#define MAXV MOOG_MAX_VERTS
#define EPS_INTERP 1e-8  /* moog/sprite.py:35  _EPSILON_INTERPOLATION */
#define EPS_COLL 1e-2    /* moog/physics/collisions.py:46  _EPSILON    */

/* ------------------------------------------------------------------------ */
/* matplotlib src/_path.h restatement                                        */
/* ------------------------------------------------------------------------ */

static inline int isclose_(double a, double b) {
  return fabs(a - b) <= fmax(1e-10 * fmax(fabs(a), fabs(b)), 1e-13);
}

/* _path.h segments_intersect */
static int segments_intersect(double x1, double y1, double x2, double y2,
                              double x3, double y3, double x4, double y4) {
  double den = ((y4 - y3) * (x2 - x1)) - ((x4 - x3) * (y2 - y1));
  if (isclose_(den, 0.0)) {
    double t_area = (x2 * y3 - x3 * y2) - x1 * (y3 - y2) + y1 * (x3 - x2);
    if (isclose_(t_area, 0.0)) {
      if (x1 == x2 && x2 == x3) {
        return (fmin(y1, y2) <= fmin(y3, y4) && fmin(y3, y4) <= fmax(y1, y2)) ||
               (fmin(y3, y4) <= fmin(y1, y2) && fmin(y1, y2) <= fmax(y3, y4));
      }
      return (fmin(x1, x2) <= fmin(x3, x4) && fmin(x3, x4) <= fmax(x1, x2)) ||
             (fmin(x3, x4) <= fmin(x1, x2) && fmin(x1, x2) <= fmax(x3, x4));
    }
    return 0;
  }
  double n1 = ((x4 - x3) * (y1 - y3)) - ((y4 - y3) * (x1 - x3));
  double n2 = ((x2 - x1) * (y1 - y3)) - ((y2 - y1) * (x1 - x3));
  double u1 = n1 / den;
  double u2 = n2 / den;
  return ((u1 > 0.0 || isclose_(u1, 0.0)) && (u1 < 1.0 || isclose_(u1, 1.0)) &&
          (u2 > 0.0 || isclose_(u2, 0.0)) && (u2 < 1.0 || isclose_(u2, 1.0)));
}

/* _path.h path_intersects_path: open polylines a[na][2], b[nb][2] */
static int path_intersects_path(const double *a, int na, const double *b, int nb) {
  if (na < 2 || nb < 2) return 0;
  double x11 = a[0], y11 = a[1];
  for (int i = 1; i < na; ++i) {
    double x12 = a[2 * i], y12 = a[2 * i + 1];
    if (isclose_((x11 - x12) * (x11 - x12) + (y11 - y12) * (y11 - y12), 0.0)) continue;
    double x21 = b[0], y21 = b[1];
    for (int j = 1; j < nb; ++j) {
      double x22 = b[2 * j], y22 = b[2 * j + 1];
      if (isclose_((x21 - x22) * (x21 - x22) + (y21 - y22) * (y21 - y22), 0.0)) continue;
      if (segments_intersect(x11, y11, x12, y12, x21, y21, x22, y22)) return 1;
      x21 = x22;
      y21 = y22;
    }
    x11 = x12;
    y11 = y12;
  }
  return 0;
}

/* _path.h point_in_path_impl, one closed sub-path of nv vertices */
void orc_points_in_path(const double *pts, int np_, const double *v, int nv, uint8_t *out) {
  for (int p = 0; p < np_; ++p) {
    out[p] = 0;
    if (nv < 3) continue;
    double tx = pts[2 * p], ty = pts[2 * p + 1];
    if (!(isfinite(tx) && isfinite(ty))) continue;
    int inside = 0;
    for (int i = 0; i < nv; ++i) {
      int k = (i + 1 == nv) ? 0 : i + 1;
      double x0 = v[2 * i], y0 = v[2 * i + 1];
      double x1 = v[2 * k], y1 = v[2 * k + 1];
      int f0 = (y0 >= ty), f1 = (y1 >= ty);
      if (f0 != f1) {
        if ((((y1 - ty) * (x0 - x1)) >= ((x1 - tx) * (y0 - y1))) == f1) inside ^= 1;
      }
    }
    out[p] = (uint8_t)inside;
  }
}

static int all_points_in_path(const double *pts, int np_, const double *v, int nv) {
  uint8_t in;
  if (nv < 3) return 0;
  for (int p = 0; p < np_; ++p) {
    orc_points_in_path(pts + 2 * p, 1, v, nv, &in);
    if (!in) return 0;
  }
  return 1;
}

/* Path.intersects_path(a, b, filled=True) as MOOG calls it (sprite.py:482-483) */
int orc_path_intersects_filled(const double *a, int na, const double *b, int nb) {
  if (path_intersects_path(a, na, b, nb)) return 1;
  if (all_points_in_path(b, nb, a, na)) return 1; /* b inside a */
  if (all_points_in_path(a, na, b, nb)) return 1; /* a inside b */
  return 0;
}

/* ------------------------------------------------------------------------ */
/* env view                                                                  */
/* ------------------------------------------------------------------------ */

typedef struct {
  const int32_t *hdr;
  const moog_op *ops;
  const int32_t *ipool;
  const moog_ex *expr;
  int S, L, K;
  double *dyn, *stat;
  int32_t *meta, *cnt, *envi;
  double *envf;
  double *vtx;          /* [VT][2] cached world vertices of this env */
  const int32_t *voff;  /* [S+1] first vertex of each slot          */
  const double *dpool;  /* shape records / sampler parameters       */
  const double *noise; /* [K][noise_dim] uniforms in [0,1) for this step, or NULL */
  const double *rule_noise; /* [rule_noise_dim] uniforms of the rules pass being made, or NULL */
  int env_id;               /* row of the batch: a key of the counter-based draws */
  int substep;
  /* instrumentation for parity tests */
  int64_t n_overlap_calls, n_overlap_true, n_collisions;
  uint64_t overlap_hash; /* order-sensitive hash of the (a, b) of every overlaps() call that returned True */
} env_t;

#define DYN(e, f, s) ((e)->dyn[(f) * (e)->S + (s)])
#define STAT(e, f, s) ((e)->stat[(f) * (e)->S + (s)])
#define META(e, f, s) ((e)->meta[(f) * (e)->S + (s)])
#define LOFF(e, l) ((e)->hdr[MOOG_H_LAYER_OFF + (l)])

typedef struct {
  int n;               /* number of distinct vertices */
  double v[MOOG_MAX_OUTLINE + 1][2]; /* closed: v[n] == v[0]  (sprite.py:394) */
} poly_t;

static void bind_program(env_t *e, const void *blob) {
  const int32_t *hdr = (const int32_t *)blob;
  e->hdr = hdr;
  e->ops = (const moog_op *)(hdr + MOOG_HDR_WORDS);
  e->ipool = (const int32_t *)(e->ops + hdr[MOOG_H_N_OPS]);
  int npool = (hdr[MOOG_H_N_IPOOL] + 1) & ~1;
  e->expr = (const moog_ex *)(e->ipool + npool);
  e->S = hdr[MOOG_H_N_SLOTS];
  e->voff = e->ipool + hdr[MOOG_H_VOFF];
  e->dpool = (const double *)(e->expr + hdr[MOOG_H_N_EXPR]);
  e->L = hdr[MOOG_H_N_LAYERS];
  e->K = hdr[MOOG_H_K];
}

/* The reference never recomputes world vertices from the factors during an
 * episode: Sprite._path is a cache that the position / angle setters transform
 * incrementally (sprite.py:531-540, 616-633).  Its roundings decide exact ties
 * (two identical circles colliding), so the cache is part of the state. */
static void world_path(const env_t *e, int s, poly_t *P) {
  int n = META(e, MOOG_M_NV, s);
  const double *v = e->vtx + 2 * (size_t)e->voff[s];
  P->n = n;
  memcpy(&P->v[0][0], v, sizeof(double) * 2 * n);
  P->v[n][0] = P->v[0][0];
  P->v[n][1] = P->v[0][1];
}

/* sprite.py:616-633 position setter: translate the cached path by the
 * rounded difference; affine_transform evaluates 1*x + 0*y + tx. */
static void set_position(env_t *e, int s, double nx, double ny) {
  double tx = nx - DYN(e, MOOG_D_X, s), ty = ny - DYN(e, MOOG_D_Y, s);
  int n = META(e, MOOG_M_NV, s);
  double *v = e->vtx + 2 * (size_t)e->voff[s];
  for (int i = 0; i < n; ++i) {
    double x = v[2 * i], y = v[2 * i + 1];
    v[2 * i] = 1.0 * x + 0.0 * y + tx;
    v[2 * i + 1] = 0.0 * x + 1.0 * y + ty;
  }
  DYN(e, MOOG_D_X, s) = nx;
  DYN(e, MOOG_D_Y, s) = ny;
}

/* sprite.py:531-540 angle setter: rotate_around(x, y, a - angle) */
static void aff_identity(double m[9]);
static void aff_rotate_around(double m[9], double x, double y, double theta);
static void set_angle(env_t *e, int s, double a, int a_is_f32) {
  double m[9];
  aff_identity(m);
  /* `a - self._angle`: float32 arithmetic when `a` is np.float32 and the old
   * angle is a python float or np.float32 */
  int old_kind = (META(e, MOOG_M_FLAGS, s) >> MOOG_SF_ANG_SHIFT) & 3;
  double dth = (a_is_f32 && old_kind != 2) ? (double)((float)a - (float)DYN(e, MOOG_D_ANG, s))
                                           : a - DYN(e, MOOG_D_ANG, s);
  aff_rotate_around(m, DYN(e, MOOG_D_X, s), DYN(e, MOOG_D_Y, s), dth);
  int n = META(e, MOOG_M_NV, s);
  double *v = e->vtx + 2 * (size_t)e->voff[s];
  for (int i = 0; i < n; ++i) {
    double x = v[2 * i], y = v[2 * i + 1];
    v[2 * i] = m[0] * x + m[1] * y + m[2];
    v[2 * i + 1] = m[3] * x + m[4] * y + m[5];
  }
  DYN(e, MOOG_D_ANG, s) = a;
}

/* numpy arithmetic conventions that decide exact ties (e.g. two identical
 * circles colliding head-on, where |since_0| == |since_1| up to rounding):
 *   - np.linalg.norm(v, axis=1) and elementwise code: separate mul / add;
 *   - np.dot / np.linalg.norm on 1-D 2-vectors go through BLAS ddot, whose
 *     scalar tail accumulates with fused multiply-adds: fma(y,y', x*x');
 *   - np.dot of 3x3 matrices (Affine2D `a + b`): fma(b2,a2, fma(b1,a1, b0*a0)).
 * Measured on numpy 2.3.5 / OpenBLAS 0.3.30 in this image (see oracle/README.md). */
static inline double norm_ax(double x, double y) { return sqrt(x * x + y * y); }
static inline double dot2(double ax, double ay, double bx, double by) { return fma(ay, by, ax * bx); }
static inline double norm1(double x, double y) { return sqrt(dot2(x, y, x, y)); }

/* ---- NumPy dtype emulation ------------------------------------------------
 * distribs.Continuous samples float32 (distributions.py:81,95-99).  A sprite
 * whose x_vel AND y_vel were sampled keeps a float32 velocity array for the
 * whole episode (`velocity += dv` rounds to float32); a sampled angle_vel is a
 * float32 0-d array, and `angle + dt * angle_vel` then becomes float32 as well
 * (NumPy 2 weak-scalar promotion).  Those roundings are ~6e-8 relative but are
 * amplified past the 1e-5 budget by the collision impulse, so the dtype of
 * velocity / angle_vel / angle is tracked per sprite in meta flags. */
enum { KIND_WEAK = 0, KIND_F32 = 1, KIND_F64 = 2 }; /* python float / np.float32 / np.float64 */
static inline double f32r(double x) { return (double)(float)x; }
static inline double f32mul(double a, double b) { return (double)((float)a * (float)b); }
static inline double f32add(double a, double b) { return (double)((float)a + (float)b); }
static inline double f32sub(double a, double b) { return (double)((float)a - (float)b); }
static inline double f32div(double a, double b) { return (double)((float)a / (float)b); }
static inline int vel32(const env_t *e, int s) { return (META(e, MOOG_M_FLAGS, s) & MOOG_SF_VEL32) != 0; }
static inline int angvel_kind(const env_t *e, int s) { return (META(e, MOOG_M_FLAGS, s) >> MOOG_SF_ANGVEL_SHIFT) & 3; }
static inline int ang_kind(const env_t *e, int s) { return (META(e, MOOG_M_FLAGS, s) >> MOOG_SF_ANG_SHIFT) & 3; }
static inline void set_angvel_kind(env_t *e, int s, int k) {
  META(e, MOOG_M_FLAGS, s) = (META(e, MOOG_M_FLAGS, s) & ~(3 << MOOG_SF_ANGVEL_SHIFT)) | (k << MOOG_SF_ANGVEL_SHIFT);
}
static inline void set_ang_kind(env_t *e, int s, int k) {
  META(e, MOOG_M_FLAGS, s) = (META(e, MOOG_M_FLAGS, s) & ~(3 << MOOG_SF_ANG_SHIFT)) | (k << MOOG_SF_ANG_SHIFT);
}
static inline int valias(const env_t *e, int s) {
  return (META(e, MOOG_M_FLAGS, s) >> MOOG_SF_VALIAS_SHIFT) & MOOG_SF_VALIAS_MASK;
}
/* `sprite.velocity += dv` with a float64 dv (ndarray in-place add keeps the dtype).
 * The array may be shared with other sprites (MOOG_SF_VALIAS_SHIFT): the in-place
 * add then shows in all of them (sprite.py:639-643, tether_physics.py:86-91). */
static inline void add_velocity(env_t *e, int s, double dvx, double dvy) {
  double vx = DYN(e, MOOG_D_VX, s) + dvx, vy = DYN(e, MOOG_D_VY, s) + dvy;
  if (vel32(e, s)) {
    vx = f32r(vx);
    vy = f32r(vy);
  }
  int id = valias(e, s);
  if (id) {
    for (int t = 0; t < e->S; ++t)
      if (valias(e, t) == id) {
        DYN(e, MOOG_D_VX, t) = vx;
        DYN(e, MOOG_D_VY, t) = vy;
      }
    return;
  }
  DYN(e, MOOG_D_VX, s) = vx;
  DYN(e, MOOG_D_VY, s) = vy;
}
/* `sprite.velocity = value` replaces the array by a new float64 one */
static inline void assign_velocity(env_t *e, int s, double vx, double vy) {
  DYN(e, MOOG_D_VX, s) = vx;
  DYN(e, MOOG_D_VY, s) = vy;
  META(e, MOOG_M_FLAGS, s) &= ~(MOOG_SF_VEL32 | (MOOG_SF_VALIAS_MASK << MOOG_SF_VALIAS_SHIFT));
}
/* a fresh alias id for an array object that several sprites are about to share */
static inline int new_valias(env_t *e) {
  int id = e->envi[MOOG_EI_VALIAS_NEXT] + 1;
  if (id > MOOG_SF_VALIAS_MASK || id < 1) id = 1;
  e->envi[MOOG_EI_VALIAS_NEXT] = id;
  return id;
}
/* `sprite.angle_vel += dw` with an np.float64 dw */
static inline void add_angvel(env_t *e, int s, double dw) {
  double w = DYN(e, MOOG_D_ANGVEL, s) + dw;
  if (angvel_kind(e, s) == KIND_F32)
    w = f32r(w); /* 0-d float32 array, in-place */
  else
    set_angvel_kind(e, s, KIND_F64);
  DYN(e, MOOG_D_ANGVEL, s) = w;
}
/* value of `c * sprite.velocity * dt` style products used by the trajectory code:
 * float32 chain when the velocity array is float32 */
static inline double scaled_vel(const env_t *e, int s, int field, double c, double dt) {
  double v = DYN(e, field, s);
  return vel32(e, s) ? f32mul(f32mul(c, v), dt) : c * v * dt;
}
static inline double scaled_angvel(const env_t *e, int s, double c, double dt) {
  double w = DYN(e, MOOG_D_ANGVEL, s);
  return angvel_kind(e, s) == KIND_F32 ? f32mul(f32mul(c, w), dt) : c * w * dt;
}

static inline int is_symmetric_circle(const env_t *e, int s) {
  /* sprite.py:500-502 */
  return (META(e, MOOG_M_FLAGS, s) & MOOG_SF_CIRCLE) && STAT(e, MOOG_S_ASPECT, s) == 1.0;
}

/* Optional call log for parity tests: (a, b, result) per overlaps() call. */
static uint8_t *g_log = NULL;
static int64_t g_log_cap = 0, g_log_len = 0;
void orc_set_overlap_log(uint8_t *buf, int64_t cap) {
  g_log = buf;
  g_log_cap = cap;
  g_log_len = 0;
}
int64_t orc_overlap_log_len(void) { return g_log_len; }

/* Row-order experiment (test infrastructure for the next kernel design, see collision_op_rows
 * below): while a row of a Collision entry is processed out of order, the True events of its
 * overlaps() calls are logged per row instead of being folded into the order-sensitive hash. */
static int g_row_mode = 0, g_row_cur = -1;
static int64_t g_row_stats[4]; /* entries handled, rows, waves, sprites that outran the margin */
typedef struct { int n, cap; int *ab; } row_log_t;
static row_log_t *g_row_logs = NULL;
static void row_log_event(int row, int a, int b) {
  row_log_t *L = &g_row_logs[row];
  if (L->n == L->cap) {
    L->cap = L->cap ? 2 * L->cap : 16;
    L->ab = (int *)realloc(L->ab, sizeof(int) * 2 * (size_t)L->cap);
  }
  L->ab[2 * L->n] = a;
  L->ab[2 * L->n + 1] = b;
  L->n++;
}
void orc_set_row_mode(int mode) {
  g_row_mode = mode;
  memset(g_row_stats, 0, sizeof(g_row_stats));
}
void orc_row_stats(int64_t out[4]) { memcpy(out, g_row_stats, sizeof(g_row_stats)); }

/* sprite.py:462-484 Sprite.overlaps_sprite */
static int overlaps(env_t *e, int a, int b) {
  int r = 0;
  double dx = DYN(e, MOOG_D_X, a) - DYN(e, MOOG_D_X, b);
  double dy = DYN(e, MOOG_D_Y, a) - DYN(e, MOOG_D_Y, b);
  double center_dist = norm1(dx, dy);
  if (!(center_dist > STAT(e, MOOG_S_MAXR, a) + STAT(e, MOOG_S_MAXR, b))) {
    poly_t A, B;
    world_path(e, a, &A);
    world_path(e, b, &B);
    r = orc_path_intersects_filled(&A.v[0][0], A.n + 1, &B.v[0][0], B.n + 1);
  }
  e->n_overlap_calls++;
  e->n_overlap_true += r;
  if (g_log && g_log_len + 3 <= g_log_cap) {
    g_log[g_log_len] = (uint8_t)a;
    g_log[g_log_len + 1] = (uint8_t)b;
    g_log[g_log_len + 2] = (uint8_t)r;
  }
  g_log_len += 3;
  if (r) { /* order-sensitive hash of the True events only (cheap enough to keep on the device) */
    if (g_row_cur >= 0)
      row_log_event(g_row_cur, a, b); /* rows run out of order: folded in row order afterwards */
    else
      e->overlap_hash = (e->overlap_hash ^ (uint64_t)((a * 1315423911u) ^ (b * 2654435761u) ^ 1u)) *
                        1099511628211ull;
  }
  return r;
}

/* sprite.py:442-460 Sprite.contains_points (sprite s contains pts?) */
static void contains_points(const env_t *e, int s, const poly_t *P, const double *pts, int np_,
                            uint8_t *out) {
  if (is_symmetric_circle(e, s)) {
    double px = DYN(e, MOOG_D_X, s), py = DYN(e, MOOG_D_Y, s);
    double r = STAT(e, MOOG_S_MAXR, s);
    for (int i = 0; i < np_; ++i)
      out[i] = norm_ax(pts[2 * i] - px, pts[2 * i + 1] - py) <= r;
  } else {
    orc_points_in_path(pts, np_, &P->v[0][0], P->n + 1, out);
  }
}

/* ------------------------------------------------------------------------ */
/* Affine2D helpers (matplotlib/transforms.py), 3x3 row-major                 */
/* ------------------------------------------------------------------------ */

static void aff_identity(double m[9]) {
  memset(m, 0, 9 * sizeof(double));
  m[0] = m[4] = m[8] = 1.0;
}
static void aff_translate(double m[9], double tx, double ty) {
  m[2] += tx;
  m[5] += ty;
}
static void aff_rotate(double m[9], double theta) {
  double a = cos(theta), b = sin(theta);
  double xx = m[0], xy = m[1], x0 = m[2], yx = m[3], yy = m[4], y0 = m[5];
  m[0] = a * xx - b * yx;
  m[1] = a * xy - b * yy;
  m[2] = a * x0 - b * y0;
  m[3] = b * xx + a * yx;
  m[4] = b * xy + a * yy;
  m[5] = b * x0 + a * y0;
}
static void aff_rotate_around(double m[9], double x, double y, double theta) {
  aff_translate(m, -x, -y);
  aff_rotate(m, theta);
  aff_translate(m, x, y);
}
/* out = b . a   ("a, then b") */
static void aff_then(const double a[9], const double b[9], double out[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      out[3 * r + c] = fma(b[3 * r + 2], a[6 + c], fma(b[3 * r + 1], a[3 + c], b[3 * r] * a[c]));
}

/* ------------------------------------------------------------------------ */
/* collisions.py                                                             */
/* ------------------------------------------------------------------------ */

typedef struct {
  int has_point; /* collision_point is not None */
  int future;    /* normal is the scalar NaN marker (collisions.py:214-217) */
  int has_since; /* since_collision is not None */
  double point[2], normal[2], since[2], perp[2];
} cvec_t;

/* collisions.py:62-98 _relative_motion_trajectory (matrix only) */
static void rel_motion_matrix(const env_t *e, int ps, int as, double dt, double out[9]) {
  double m1[9], m2[9], m3[9], m4[9], t12[9], t123[9];
  aff_identity(m1);
  aff_rotate_around(m1, DYN(e, MOOG_D_X, ps), DYN(e, MOOG_D_Y, ps), scaled_angvel(e, ps, -1.0, dt));
  aff_identity(m2);
  aff_translate(m2, scaled_vel(e, ps, MOOG_D_VX, -1.0, dt), scaled_vel(e, ps, MOOG_D_VY, -1.0, dt));
  aff_identity(m3);
  aff_rotate_around(m3, DYN(e, MOOG_D_X, as), DYN(e, MOOG_D_Y, as),
                    angvel_kind(e, as) == KIND_F32 ? f32mul(DYN(e, MOOG_D_ANGVEL, as), dt)
                                                   : DYN(e, MOOG_D_ANGVEL, as) * dt);
  aff_identity(m4);
  aff_translate(m4, vel32(e, as) ? f32mul(DYN(e, MOOG_D_VX, as), dt) : DYN(e, MOOG_D_VX, as) * dt,
                vel32(e, as) ? f32mul(DYN(e, MOOG_D_VY, as), dt) : DYN(e, MOOG_D_VY, as) * dt);
  aff_then(m1, m2, t12);
  aff_then(t12, m3, t123);
  aff_then(t123, m4, out);
}

/* collisions.py:101-232 _directed_collision_vectors(sprite_0=s0, sprite_1=s1) */
static void directed_collision_vectors(env_t *e, int s0, int s1, double dt, cvec_t *o) {
  poly_t P0, P1;
  memset(o, 0, sizeof(*o));
  world_path(e, s0, &P0);
  world_path(e, s1, &P1);
  int n0 = P0.n, n1 = P1.n;
  uint8_t inside[MAXV + 1];
  contains_points(e, s1, &P1, &P0.v[0][0], n0, inside);
  int idx[MAXV], nc = 0;
  for (int i = 0; i < n0; ++i)
    if (inside[i]) idx[nc++] = i;
  if (nc == 0) return; /* (None, None, None, None) */

  double M[9];
  rel_motion_matrix(e, s0, s1, dt, M);

  static const double NEG_INF = -INFINITY;
  double cross_a_sel[MAXV]; /* cross_a[i, inds_crossings[i]] */
  int ind_sel[MAXV];
  double cp[MAXV][2], diff[MAXV][2], dist[MAXV];
  int any_cross = 0;
  for (int c = 0; c < nc; ++c) {
    double ex = P0.v[idx[c]][0], ey = P0.v[idx[c]][1]; /* traj[:,1] */
    double sx = M[0] * ex + M[1] * ey + M[2];          /* traj[:,0] */
    double sy = M[3] * ex + M[4] * ey + M[5];
    double d0x = ex - sx, d0y = ey - sy;
    double best = 0.0, best_a = 0.0;
    int best_j = -1;
    for (int j = 0; j < n1; ++j) {
      /* sprite.py:145-161 segment_crossing_coefficients */
      double s1x = P1.v[j][0], s1y = P1.v[j][1];
      double d1x = P1.v[j + 1][0] - s1x, d1y = P1.v[j + 1][1] - s1y;
      double den = (d0x * d1y - d0y * d1x) + EPS_INTERP;
      double qx = s1x - sx, qy = s1y - sy;
      double A = (qx * d1y - qy * d1x) / den;
      double B = (qx * d0y - qy * d0x) / den;
      int crossing = (B >= 0) && (B <= 1);
      any_cross |= crossing;
      if (!crossing) A = NEG_INF;
      double ab = fabs(1.0 - A);
      /* np.argmin: first NaN wins, else first minimum */
      if (best_j < 0) {
        best = ab; best_a = A; best_j = j;
      } else if (!isnan(best) && (isnan(ab) || ab < best)) {
        best = ab; best_a = A; best_j = j;
      }
    }
    ind_sel[c] = best_j;
    cross_a_sel[c] = best_a;
    cp[c][0] = sx + best_a * (ex - sx);
    cp[c][1] = sy + best_a * (ey - sy);
    diff[c][0] = ex - cp[c][0];
    diff[c][1] = ey - cp[c][1];
    dist[c] = norm_ax(diff[c][0], diff[c][1]);
    if (dist[c] == INFINITY) dist[c] = 0.0;
  }
  if (!any_cross) return; /* collisions.py:177-179 */

  /* np.argmax: first NaN wins, else first maximum */
  int ci = 0;
  for (int c = 1; c < nc; ++c) {
    if (isnan(dist[ci])) break;
    if (isnan(dist[c]) || dist[c] > dist[ci]) ci = c;
  }
  int ei = ind_sel[ci];
  o->has_point = 1;
  o->has_since = 1;
  o->point[0] = cp[ci][0];
  o->point[1] = cp[ci][1];
  o->since[0] = diff[ci][0];
  o->since[1] = diff[ci][1];
  if (cross_a_sel[ci] > 1) { /* collisions.py:214-217 */
    o->future = 1;
    return;
  }
  double dvx = P1.v[ei + 1][0] - P1.v[ei][0];
  double dvy = P1.v[ei + 1][1] - P1.v[ei][1];
  double nx = dvy, ny = -1.0 * dvx;
  double nn = norm1(nx, ny);
  o->normal[0] = nx / nn;
  o->normal[1] = ny / nn;
  double f = dot2(o->since[0], o->since[1], dvx, dvy) / dot2(dvx, dvy, dvx, dvy);
  o->perp[0] = o->since[0] - dvx * f;
  o->perp[1] = o->since[1] - dvy * f;
}

/* collisions.py:235-289 _get_collision_vectors */
static void get_collision_vectors(env_t *e, int s0, int s1, double dt, cvec_t *o) {
  cvec_t c0, c1;
  directed_collision_vectors(e, s1, s0, dt, &c0);
  directed_collision_vectors(e, s0, s1, dt, &c1);
  double s0x = 0, s0y = 0, s1x = 0, s1y = 0;
  if (c0.has_point) {
    if (!c0.future) { /* -1 * nan stays the nan marker */
      c0.normal[0] = -1.0 * c0.normal[0];
      c0.normal[1] = -1.0 * c0.normal[1];
    }
    c0.since[0] = -1.0 * c0.since[0];
    c0.since[1] = -1.0 * c0.since[1];
    s0x = c0.since[0];
    s0y = c0.since[1];
  }
  if (c1.has_since) {
    s1x = c1.since[0];
    s1y = c1.since[1];
  }
#ifdef ORC_DEBUG
  fprintf(stderr, "ORC dir0 has=%d point=(%.17g,%.17g) since=(%.17g,%.17g) n=%.17g | dir1 has=%d point=(%.17g,%.17g) since=(%.17g,%.17g) n=%.17g\n",
          c0.has_point, c0.point[0], c0.point[1], s0x, s0y, norm1(s0x, s0y), c1.has_point, c1.point[0], c1.point[1], s1x, s1y, norm1(s1x, s1y));
#endif
  if (norm1(s0x, s0y) > norm1(s1x, s1y))
    *o = c0;
  else
    *o = c1;
}

static inline double moment_of_inertia(const env_t *e, int s) {
  /* sprite.py:662-664  sum(mass * [ix, iy]) */
  double m = STAT(e, MOOG_S_MASS, s);
  return 0.0 + m * STAT(e, MOOG_S_IX, s) + m * STAT(e, MOOG_S_IY, s);
}

/* collisions.py:292-350 */
static void collide_without_update_angle_vel(env_t *e, int s0, int s1, const cvec_t *cv,
                                             double elasticity, int symmetric) {
  double nx = cv->normal[0], ny = cv->normal[1];
  double nn = norm1(nx, ny);
  /* np.isclose(nn, 1., atol=1e-4): |nn-1| <= 1e-4 + 1e-5*1 */
  if (!(fabs(nn - 1.0) <= 1e-4 + 1e-5 * 1.0)) {
    e->envi[MOOG_EI_ERR] |= MOOG_ERR_NORMAL_NOT_UNIT;
    return;
  }
  double v0x = DYN(e, MOOG_D_VX, s0), v0y = DYN(e, MOOG_D_VY, s0);
  double v1x = DYN(e, MOOG_D_VX, s1), v1y = DYN(e, MOOG_D_VY, s1);
  double m0 = STAT(e, MOOG_S_MASS, s0), m1 = STAT(e, MOOG_S_MASS, s1);
  double d0 = dot2(v0x, v0y, nx, ny), d1 = dot2(v1x, v1y, nx, ny);
  double v0nx = d0 * nx, v0ny = d0 * ny, v1nx = d1 * nx, v1ny = d1 * ny;
  double cmx, cmy;
  if (symmetric) {
    cmx = (v0nx * m0 + v1nx * m1) / (m0 + m1);
    cmy = (v0ny * m0 + v1ny * m1) / (m0 + m1);
  } else {
    cmx = v1nx;
    cmy = v1ny;
  }
  double k = 1 + elasticity;
  add_velocity(e, s0, k * (cmx - v0nx), k * (cmy - v0ny));
  add_velocity(e, s1, k * (cmx - v1nx), k * (cmy - v1ny));
}

/* collisions.py:353-454 */
static void collide_with_update_angle_vel(env_t *e, int s0, int s1, const cvec_t *cv,
                                          double elasticity, int symmetric) {
  double nx = cv->normal[0], ny = cv->normal[1];
  double m0 = STAT(e, MOOG_S_MASS, s0), m1 = STAT(e, MOOG_S_MASS, s1);
  double w0 = DYN(e, MOOG_D_ANGVEL, s0), w1 = DYN(e, MOOG_D_ANGVEL, s1);
  double i0 = moment_of_inertia(e, s0), i1 = moment_of_inertia(e, s1);
  double v0 = dot2(DYN(e, MOOG_D_VX, s0), DYN(e, MOOG_D_VY, s0), nx, ny);
  double v1 = dot2(DYN(e, MOOG_D_VX, s1), DYN(e, MOOG_D_VY, s1), nx, ny);
  double c0x = cv->point[0] - DYN(e, MOOG_D_X, s0), c0y = cv->point[1] - DYN(e, MOOG_D_Y, s0);
  double c1x = cv->point[0] - DYN(e, MOOG_D_X, s1), c1y = cv->point[1] - DYN(e, MOOG_D_Y, s1);
  double r0 = norm1(c0x, c0y), r1 = norm1(c1x, c1y);
  double sin0 = (c0x * ny - c0y * nx) / r0;
  double sin1 = (c1x * ny - c1y * nx) / r1;
  double sa = r0 * sin0, sb = r1 * sin1;
  double a = m0 + m1 + m0 * m1 * ((sa * sa / i0) + (sb * sb / i1));
  double b = (1 + elasticity) * (v0 - v1 + w0 * sa - w1 * sb);
  double dv0, dv1;
  if (symmetric) {
    dv0 = -1 * m1 * b / a;
    dv1 = m0 * b / a;
  } else {
    dv0 = -1 * m1 * b / (a - m0);
    dv1 = 0.0;
  }
  double dw0 = m0 * dv0 * sa / i0;
  double dw1 = m1 * dv1 * sb / i1;
  add_velocity(e, s0, dv0 * nx, dv0 * ny);
  add_velocity(e, s1, dv1 * nx, dv1 * ny);
  add_angvel(e, s0, dw0);
  add_angvel(e, s1, dw1);
}

typedef struct {
  int K;
  double pt[MAXV * MAXV][2];
  int i0[MAXV * MAXV], i1[MAXV * MAXV];
} crossings_t;

/* sprite.py:166-226 segment_crossings / sprite_edge_crossings (row-major argwhere order) */
static void sprite_edge_crossings(const poly_t *P0, const poly_t *P1, crossings_t *c) {
  c->K = 0;
  for (int i = 0; i < P0->n; ++i) {
    double s0x = P0->v[i][0], s0y = P0->v[i][1];
    double d0x = P0->v[i + 1][0] - s0x, d0y = P0->v[i + 1][1] - s0y;
    for (int j = 0; j < P1->n; ++j) {
      double s1x = P1->v[j][0], s1y = P1->v[j][1];
      double d1x = P1->v[j + 1][0] - s1x, d1y = P1->v[j + 1][1] - s1y;
      double den = (d0x * d1y - d0y * d1x) + EPS_INTERP;
      double qx = s1x - s0x, qy = s1y - s0y;
      double A = (qx * d1y - qy * d1x) / den;
      double B = (qx * d0y - qy * d0x) / den;
      if ((A > 0) && (A < 1) && (B > 0) && (B < 1)) {
        int k = c->K++;
        c->pt[k][0] = s0x + A * d0x;
        c->pt[k][1] = s0y + A * d0y;
        c->i0[k] = i;
        c->i1[k] = j;
      }
    }
  }
}

static inline double sign_(double x) { return isnan(x) ? x : (x > 0) - (x < 0); }

/* collisions.py:658-748 _position_correction.  `me`/`other` are the function's
 * sprite_0/sprite_1; inds_me/inds_other its crossing_inds_0/crossing_inds_1. */
static void position_correction(env_t *e, const crossings_t *c, int me, const poly_t *Pme,
                                const int *inds_me, int other, const poly_t *Pother,
                                const int *inds_other, double out[2]) {
  int K = c->K;
  double px = DYN(e, MOOG_D_X, me), py = DYN(e, MOOG_D_Y, me);
  /* stable argsort of distances (numpy uses insertion sort for n <= 16;
   * larger K falls back to a stable order as well) */
  int order[MAXV * MAXV];
  double d[MAXV * MAXV];
  for (int k = 0; k < K; ++k) {
    d[k] = norm_ax(c->pt[k][0] - px, c->pt[k][1] - py);
    order[k] = k;
  }
  for (int k = 1; k < K; ++k) {
    int o = order[k], m = k;
    while (m > 0 && d[order[m - 1]] > d[o]) {
      order[m] = order[m - 1];
      --m;
    }
    order[m] = o;
  }
  int k0 = order[0];
  double pt0x = c->pt[k0][0], pt0y = c->pt[k0][1];
  /* collisions.py:706 compares a value with itself -> always the edge branch */
  int pt0_ind = inds_other[k0];
  int pt1_ind = (pt0_ind - 1 + Pother->n) % Pother->n;
  /* _get_norm_v(other_vertices[pt1_ind], other_vertices[pt0_ind]) */
  double q0x = Pother->v[pt1_ind][0], q0y = Pother->v[pt1_ind][1];
  double bx = Pother->v[pt0_ind][0] - q0x, by = Pother->v[pt0_ind][1] - q0y;
  double bn = norm1(bx, by);
  bx /= bn;
  by /= bn;
  double sg = sign_(dot2(DYN(e, MOOG_D_X, other) - q0x, DYN(e, MOOG_D_Y, other) - q0y, bx, by));
  double nvx = bx * -1 * sg, nvy = by * -1 * sg;

  int n = Pme->n;
  int ind_forward = inds_me[k0];
  int ind_backward = (ind_forward - 1 + n) % n;
  int parity, cur;
  if (dot2(Pme->v[ind_forward][0] - pt0x, Pme->v[ind_forward][1] - pt0y, nvx, nvy) > 0) {
    parity = 1;
    cur = ind_forward;
  } else if (dot2(Pme->v[ind_backward][0] - pt0x, Pme->v[ind_backward][1] - pt0y, nvx, nvy) > 0) {
    parity = -1;
    cur = ind_backward;
  } else {
    out[0] = out[1] = INFINITY;
    return;
  }
  double worst = 0;
  int guard = 0;
  for (;;) {
    double pen = dot2(Pme->v[cur][0] - pt0x, Pme->v[cur][1] - pt0y, nvx, nvy);
    if (!(pen > 0)) break;
    if (pen > worst) worst = pen;
    cur = (cur + parity + n) % n;
    if (++guard > n) { /* the reference would spin forever here */
      e->envi[MOOG_EI_ERR] |= MOOG_ERR_DISJOINT_LOOP;
      break;
    }
  }
  out[0] = worst * nvx;
  out[1] = worst * nvy;
}

/* collisions.py:586-655 Collision._make_disjoint */
static void make_disjoint(env_t *e, int s0, int s1, int symmetric) {
  poly_t P0, P1;
  crossings_t *cc = (crossings_t *)malloc(sizeof(crossings_t)); /* ~24 KB: keep off the stack */
  world_path(e, s0, &P0);
  world_path(e, s1, &P1);
  sprite_edge_crossings(&P0, &P1, cc);
  if (cc->K <= 1) {
    free(cc);
    return;
  }
  double c0[2], c1[2];
  position_correction(e, cc, s0, &P0, cc->i0, s1, &P1, cc->i1, c0);
  position_correction(e, cc, s1, &P1, cc->i1, s0, &P0, cc->i0, c1);
  double corr[2];
  if (norm1(c0[0], c0[1]) > norm1(c1[0], c1[1])) {
    corr[0] = -1 * (1 + EPS_COLL) * c0[0];
    corr[1] = -1 * (1 + EPS_COLL) * c0[1];
  } else {
    corr[0] = (1 + EPS_COLL) * c0[0];
    corr[1] = (1 + EPS_COLL) * c0[1];
  }
  if (!(isfinite(corr[0]) && isfinite(corr[1]))) corr[0] = corr[1] = 0.0;
  if (symmetric) {
    set_position(e, s0, DYN(e, MOOG_D_X, s0) + 0.5 * corr[0], DYN(e, MOOG_D_Y, s0) + 0.5 * corr[1]);
    set_position(e, s1, DYN(e, MOOG_D_X, s1) - 0.5 * corr[0], DYN(e, MOOG_D_Y, s1) - 0.5 * corr[1]);
  } else {
    set_position(e, s0, DYN(e, MOOG_D_X, s0) + corr[0], DYN(e, MOOG_D_Y, s0) + corr[1]);
  }
  free(cc);
}

/* collisions.py:494-584 Collision.step */
static void collision_step(env_t *e, const moog_op *op, int s0, int s1, int depth) {
  int symmetric = (op->flags & MOOG_FL_SYMMETRIC) != 0;
  for (;;) { /* tail recursion of collisions.py:583-584 */
    if (depth > op->i[2]) return;
    if (s0 == s1) return;
    if (!overlaps(e, s0, s1)) return;
    double dt = 1.0 / e->K;
    cvec_t cv;
    get_collision_vectors(e, s0, s1, dt, &cv);
#ifdef ORC_DEBUG
    fprintf(stderr, "ORC cv s0=%d s1=%d has=%d fut=%d point=(%.8g,%.8g) normal=(%.8g,%.8g) since=(%.8g,%.8g) perp=(%.8g,%.8g)\n",
            s0, s1, cv.has_point, cv.future, cv.point[0], cv.point[1], cv.normal[0], cv.normal[1],
            cv.since[0], cv.since[1], cv.perp[0], cv.perp[1]);
#endif
    if (!cv.has_point) {
      make_disjoint(e, s0, s1, symmetric);
    } else {
      if (cv.future) return;
      e->n_collisions++;
      if (symmetric) {
        set_position(e, s0, DYN(e, MOOG_D_X, s0) - (0.5 + EPS_COLL) * cv.perp[0],
                     DYN(e, MOOG_D_Y, s0) - (0.5 + EPS_COLL) * cv.perp[1]);
        set_position(e, s1, DYN(e, MOOG_D_X, s1) + (0.5 + EPS_COLL) * cv.perp[0],
                     DYN(e, MOOG_D_Y, s1) + (0.5 + EPS_COLL) * cv.perp[1]);
      } else {
        set_position(e, s0, DYN(e, MOOG_D_X, s0) - (1. + EPS_COLL) * cv.perp[0],
                     DYN(e, MOOG_D_Y, s0) - (1. + EPS_COLL) * cv.perp[1]);
      }
      if (op->flags & MOOG_FL_UPDATE_ANGLE_VEL)
        collide_with_update_angle_vel(e, s0, s1, &cv, op->p[0], symmetric);
      else
        collide_without_update_angle_vel(e, s0, s1, &cv, op->p[0], symmetric);
    }
    depth += 1;
  }
}

/* ------------------------------------------------------------------------ */
/* forces (abstract_force.py:64-74 + the individual _compute_forces)         */
/* ------------------------------------------------------------------------ */

static inline void newton(env_t *e, int s, double fx, double fy) {
  double m = STAT(e, MOOG_S_MASS, s);
  if (!isfinite(m)) return;
  double den = m * (double)e->K;
  add_velocity(e, s, fx / den, fy / den);
}

static void force_unary(env_t *e, const moog_op *op, int s, int idx_in_layer) {
  double m = STAT(e, MOOG_S_MASS, s);
  double vx = DYN(e, MOOG_D_VX, s), vy = DYN(e, MOOG_D_VY, s);
  double fx = 0, fy = 0;
  switch (op->kind) {
    case MOOG_F_DRAG: /* friction.py:54-56 */
      if (vel32(e, s)) {
        /* python-float scalars are weak: the whole chain runs in float32 */
        if (!isfinite(m)) return;
        double c = -1 * op->p[0], den = m * (double)e->K;
        DYN(e, MOOG_D_VX, s) = f32add(vx, f32div(f32mul(f32mul(c, vx), m), den));
        DYN(e, MOOG_D_VY, s) = f32add(vy, f32div(f32mul(f32mul(c, vy), m), den));
        return;
      }
      fx = -1 * op->p[0] * vx * m;
      fy = -1 * op->p[0] * vy * m;
      break;
    case MOOG_F_KINETIC_FRICTION: { /* friction.py:25-33 */
      double n = norm1(vx, vy);
      double ux = 0, uy = 0;
      if (n != 0) {
        ux = vx / n;
        uy = vy / n;
      }
      fx = -1 * op->p[0] * ux * m;
      fy = -1 * op->p[0] * uy * m;
      break;
    }
    case MOOG_F_DOWN_GRAVITY: /* gravity.py:21-23 */
      fx = op->p[0] * m * 0;
      fy = op->p[0] * m * 1;
      break;
    case MOOG_F_RANDOM: { /* random_force.py:22-26; uniforms come from the noise tensor */
      double u0 = 0, u1 = 0;
      if (e->noise) {
        const double *nz = e->noise + (size_t)e->substep * e->hdr[MOOG_H_NOISE_DIM] + op->i[2] +
                           2 * idx_in_layer;
        u0 = nz[0];
        u1 = nz[1];
      }
      double r = 0.0 + (op->p[0] - 0.0) * u0;
      double th = 0.0 + (2 * M_PI - 0.0) * u1;
      fx = r * cos(th);
      fy = r * sin(th);
      break;
    }
  }
  newton(e, s, fx, fy);
}

static void force_binary(env_t *e, const moog_op *op, int s0, int s1) {
  /* gravity.py:44-60, distance_fn_force.py:30-45 */
  double dx = DYN(e, MOOG_D_X, s1) - DYN(e, MOOG_D_X, s0);
  double dy = DYN(e, MOOG_D_Y, s1) - DYN(e, MOOG_D_Y, s0);
  double dist = norm1(dx, dy);
  double f0x = 0, f0y = 0, f1x = 0, f1y = 0;
  if (dist != 0.) {
    double ux = dx / dist, uy = dy / dist;
    double mag = 0;
    if (op->kind == MOOG_F_GRAVITY) {
      mag = op->p[0] * STAT(e, MOOG_S_MASS, s0) * STAT(e, MOOG_S_MASS, s1) * dist;
    } else if (op->kind == MOOG_F_DIST_LINEAR) { /* distance_fn_force.py:48-74 */
      mag = op->p[0] + op->p[1] * dist;
      if (!(op->flags & MOOG_FL_APPLY_DISTANT) && dist > op->p[2]) mag = 0;
      if (!(op->flags & MOOG_FL_APPLY_NEARBY) && dist < op->p[2]) mag = 0;
    } else if (op->kind == MOOG_F_DIST_SPRING) { /* distance_fn_force.py:77-89 */
      mag = -1. * op->p[0] * (dist - op->p[1]);
    }
    f1x = mag * ux;
    f1y = mag * uy;
    if (op->flags & MOOG_FL_SYMMETRIC) {
      f0x = -1 * f1x;
      f0y = -1 * f1y;
    }
  }
  newton(e, s0, f0x, f0y);
  newton(e, s1, f1x, f1y);
}

/* ------------------------------------------------------------------------ */
/* corrective physics                                                        */
/* ------------------------------------------------------------------------ */

/* tether_physics.py:16-91 */
static void tether_sprites(env_t *e, const int *sp, int n, int update_angle_vel, int has_anchor,
                           double ax, double ay) {
  if (n == 0) return;
  double total_mass = 0;
  for (int i = 0; i < n; ++i) total_mass = total_mass + STAT(e, MOOG_S_MASS, sp[i]);
  if (isinf(total_mass)) return;
  double cx = 0, cy = 0, mx = 0, my = 0;
  for (int i = 0; i < n; ++i) {
    double m = STAT(e, MOOG_S_MASS, sp[i]);
    cx = cx + m * DYN(e, MOOG_D_X, sp[i]);
    cy = cy + m * DYN(e, MOOG_D_Y, sp[i]);
    mx = mx + m * DYN(e, MOOG_D_VX, sp[i]);
    my = my + m * DYN(e, MOOG_D_VY, sp[i]);
  }
  cx /= total_mass;
  cy /= total_mass;
  double tvx = mx / total_mass, tvy = my / total_mass;
  if (has_anchor) {
    cx = ax;
    cy = ay;
    tvx = tvy = 0;
  }
  if (update_angle_vel) {
    double Ltot = 0, Itot = 0;
    double radius[MOOG_MAX_SLOTS], perpx[MOOG_MAX_SLOTS], perpy[MOOG_MAX_SLOTS];
    for (int i = 0; i < n; ++i) {
      int s = sp[i];
      double vx = DYN(e, MOOG_D_VX, s), vy = DYN(e, MOOG_D_VY, s);
      double dpx = vx / e->K, dpy = vy / e->K;
      double px = (DYN(e, MOOG_D_X, s) + 0.5 * dpx) - cx;
      double py = (DYN(e, MOOG_D_Y, s) + 0.5 * dpy) - cy;
      double r = norm1(px, py);
      px /= r;
      py /= r;
      double qx = 0 * px + -1 * py, qy = 1 * px + 0 * py;
      double perp_vel = dot2(vx - tvx, vy - tvy, qx, qy);
      double m = STAT(e, MOOG_S_MASS, s);
      double I = moment_of_inertia(e, s);
      double L = perp_vel * m * r;
      L += DYN(e, MOOG_D_ANGVEL, s) * I;
      Ltot = Ltot + L;
      Itot = Itot + (I + m * r * r);
      radius[i] = r;
      perpx[i] = qx;
      perpy[i] = qy;
    }
    double w = Ltot / Itot;
    for (int i = 0; i < n; ++i) {
      assign_velocity(e, sp[i], tvx + radius[i] * perpx[i] * w, tvy + radius[i] * perpy[i] * w);
      DYN(e, MOOG_D_ANGVEL, sp[i]) = w;
      set_angvel_kind(e, sp[i], KIND_F64);
    }
  } else {
    /* `s.velocity = total_velocity`: the SAME ndarray object for every sprite */
    int id = new_valias(e);
    for (int i = 0; i < n; ++i) {
      assign_velocity(e, sp[i], tvx, tvy);
      META(e, MOOG_M_FLAGS, sp[i]) |= id << MOOG_SF_VALIAS_SHIFT;
      DYN(e, MOOG_D_ANGVEL, sp[i]) = 0.;
      set_angvel_kind(e, sp[i], KIND_WEAK);
    }
  }
}

static int gather_layers(const env_t *e, int start, int n, int *out) {
  int k = 0;
  for (int q = 0; q < n; ++q) {
    int l = e->ipool[start + q];
    for (int i = 0; i < e->cnt[l]; ++i) out[k++] = LOFF(e, l) + i;
  }
  return k;
}


/* ------------------------------------------------------------------------ */
/* maze_lib.Maze, RandomMazeWalk, MazePhysics                                 */
/* ------------------------------------------------------------------------ */
#define MAZE_EPS 1e-5 /* maze_physics.py:15, maze_walk.py:14 */

typedef struct { int n; const double *rows; double grid, half; } maze_t;

static maze_t maze_of(const env_t *e, int off) {
  maze_t m;
  m.n = (int)e->envf[off];
  m.rows = e->envf + off + 1;
  m.grid = 1. / m.n;        /* maze.py:32 */
  m.half = 0.5 * m.grid;    /* maze.py:33 */
  return m;
}
/* maze.py:113-118 */
static int maze_open(const maze_t *m, int i, int j) {
  if (i < 0 || j < 0 || i >= m->n || j >= m->n) return 0;
  return !(((unsigned long long)m->rows[j] >> i) & 1ull);
}
/* maze.py:120-126: v[axis][dir] */
static void maze_valid(const maze_t *m, int i, int j, double v[2][2]) {
  v[0][0] = maze_open(m, i - 1, j); v[0][1] = maze_open(m, i + 1, j);
  v[1][0] = maze_open(m, i, j - 1); v[1][1] = maze_open(m, i, j + 1);
}
static double np_sign(double x) { return isnan(x) ? x : (double)((x > 0) - (x < 0)); }
/* numpy floor_divide for doubles (npy_divmod) */
static double np_floor_divide(double a, double b) {
  if (b == 0) return a / b;
  double mod = fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0) {
    if ((b < 0) != (mod < 0)) div -= 1.0;
  }
  double fl;
  if (div != 0) {
    fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
  } else {
    fl = copysign(0.0, a / b);
  }
  return fl;
}

/* maze_walk.py:159-196 RandomMazeWalk._step_sprite (+ _get_pos_vel :46-79, _update_valid_directions :122-157) */
static void maze_walk_sprite(env_t *e, const moog_op *op, int s, int idx_in_layer) {
  if (isinf(STAT(e, MOOG_S_MASS, s))) return;
  maze_t m = maze_of(e, op->i[3]);
  const double speed = op->p[0];
  double px = DYN(e, MOOG_D_X, s), py = DYN(e, MOOG_D_Y, s);
  double pos[2] = {px, py};
  double vel[2] = {speed * np_sign(DYN(e, MOOG_D_VX, s)), speed * np_sign(DYN(e, MOOG_D_VY, s))};
  double K = (double)e->K;
  double nxt[2] = {pos[0] + vel[0] / K, pos[1] + vel[1] / K};
  int near_[2];
  double inter[2];
  for (int a = 0; a < 2; ++a) {
    near_[a] = (int)rint(pos[a] / m.grid - 0.5);
    inter[a] = m.grid * near_[a] + m.half;
  }
  double d_next_cur = fabs(nxt[0] - pos[0]) + fabs(nxt[1] - pos[1]);
  double d_int_next = fabs(nxt[0] - inter[0]) + fabs(nxt[1] - inter[1]);
  int entering = d_next_cur > d_int_next; /* the reference compares against the same quantity twice */
  double valid[2][2];
  if (entering) {
    maze_valid(&m, near_[0], near_[1], valid);
    if (op->flags & MOOG_FL_PREVENT_BACKTRACKING) {
      int axis = fabs(vel[1]) > fabs(vel[0]) ? 1 : 0; /* np.argmax: first maximum */
      double direction = np_sign(vel[axis]);
      if (direction != 0) {
        int fwd = (int)(0.5 * (1 + direction)), back = (int)(0.5 * (1 - direction));
        int can_continue = valid[axis][fwd] != 0;
        if (!can_continue && (op->flags & MOOG_FL_ALLOW_WALL_BACKTRACKING)) {
        } else if (can_continue && (op->flags & MOOG_FL_ONLY_TURN_AT_WALL)) {
          valid[0][0] = valid[0][1] = valid[1][0] = valid[1][1] = 0;
          valid[axis][fwd] = 1;
        } else {
          valid[axis][back] = 0;
        }
      }
    }
  } else if (vel[0] == 0. && vel[1] == 0.) {
    int on[2];
    for (int a = 0; a < 2; ++a) on[a] = fabs((m.half + near_[a] * m.grid) - pos[a]) < MAZE_EPS;
    if (on[0] && on[1]) {
      maze_valid(&m, near_[0], near_[1], valid);
    } else {
      valid[0][0] = valid[0][1] = valid[1][0] = valid[1][1] = 0;
      int row = 1 - (on[0] ? 0 : (on[1] ? 1 : 0)); /* 1 - np.argmax(on_grid) */
      valid[row][0] = valid[row][1] = 1;
    }
  } else {
    assign_velocity(e, s, vel[0], vel[1]);
    return;
  }
  /* sample = valid_directions * np.random.rand(2, 2); argmax of the ravel (first maximum) */
  double u[4] = {0, 0, 0, 0};
  if (e->noise) {
    const double *nz = e->noise + (size_t)e->substep * e->hdr[MOOG_H_NOISE_DIM] + op->i[2] + 4 * idx_in_layer;
    for (int k = 0; k < 4; ++k) u[k] = nz[k];
  }
  int best = 0;
  double bv = valid[0][0] * u[0];
  for (int k = 1; k < 4; ++k) {
    double v = valid[k >> 1][k & 1] * u[k];
    if (v > bv) { bv = v; best = k; }
  }
  vel[best / 2] = (1 + MAZE_EPS) * speed * (2 * (best % 2) - 1);
  assign_velocity(e, s, vel[0], vel[1]);
}

/* maze_physics.py:51-112 _get_position_affordances; returns 0 when the position is off the grid */
static int maze_affordances(const maze_t *m, const double pos[2], double aff[2][2]) {
  int near_[2], inds[2], on[2];
  for (int a = 0; a < 2; ++a) {
    near_[a] = (int)rint(pos[a] / m->grid - 0.5);
    double rounded = m->half + near_[a] * m->grid;
    on[a] = fabs(rounded - pos[a]) < MAZE_EPS;
    inds[a] = (int)np_floor_divide(pos[a] - m->half, m->grid);
    if (on[a]) inds[a] = near_[a];
  }
  aff[0][0] = aff[0][1] = aff[1][0] = aff[1][1] = 0;
  if (!on[0] && !on[1]) return 0;
  if (on[0] && on[1]) {
    double v[2][2];
    maze_valid(m, inds[0], inds[1], v);
    for (int a = 0; a < 2; ++a) {
      aff[a][0] = v[a][0] * m->grid * -1.;
      aff[a][1] = v[a][1] * m->grid * 1.;
    }
  } else {
    int i = 1 - (on[0] ? 0 : 1);
    aff[i][0] = inds[i] * m->grid + m->half - pos[i];
    aff[i][1] = (inds[i] + 1) * m->grid + m->half - pos[i];
  }
  return 1;
}

/* maze_physics.py:114-167 _get_new_velocity; `pos` and `vel` are mutated like the reference's arrays.
 * Returns 0 when a vertex reached on the way is off the grid. */
static int maze_new_velocity(const maze_t *m, double pos[2], double vel[2], double aff[2][2], int axis,
                             double out[2], int depth) {
  if (depth > 16) return 0;
  if (axis < 0) axis = fabs(vel[1]) > fabs(vel[0]) ? 1 : 0;
  if (aff[axis][0] <= vel[axis] && vel[axis] <= aff[axis][1]) {
    vel[1 - axis] = 0;
    out[0] = vel[0]; out[1] = vel[1];
    return 1;
  }
  int direction = vel[axis] > 0 ? 1 : 0; /* int(0.5 + 0.5 * sign) */
  if (aff[axis][direction] == 0) {
    axis = 1 - axis;
    direction = vel[axis] > 0 ? 1 : 0;
    if (aff[axis][direction] == 0 || vel[axis] == 0) {
      out[0] = out[1] = 0.;
      return 1;
    }
    return maze_new_velocity(m, pos, vel, aff, axis, out, depth + 1);
  }
  pos[axis] += aff[axis][direction];
  double vaff[2][2];
  if (!maze_affordances(m, pos, vaff)) return 0;
  double scaling = aff[axis][direction] / vel[axis];
  double rem[2] = {(1. - scaling) * vel[0], (1. - scaling) * vel[1]};
  double post[2];
  if (!maze_new_velocity(m, pos, rem, vaff, -1, post, depth + 1)) return 0;
  vel[0] *= scaling; vel[1] *= scaling;
  vel[1 - axis] = 0;
  vel[0] += post[0]; vel[1] += post[1];
  out[0] = vel[0]; out[1] = vel[1];
  return 1;
}

/* maze_physics.py:189-203 _update_sprite_in_maze (+ _update_sprite_angle :169-187) */
static void maze_physics_sprite(env_t *e, const moog_op *op, int s) {
  double vel[2] = {DYN(e, MOOG_D_VX, s), DYN(e, MOOG_D_VY, s)};
  if ((vel[0] == 0 && vel[1] == 0) || isnan(vel[0]) || isnan(vel[1])) return;
  maze_t m = maze_of(e, op->i[2]);
  if (!isnan(op->p[1])) { /* np.clip(velocity, -max_speed, max_speed) */
    for (int a = 0; a < 2; ++a) vel[a] = fmin(fmax(vel[a], -op->p[1]), op->p[1]);
  }
  if (!isnan(op->p[0])) {
    for (int a = 0; a < 2; ++a) {
      vel[a] += np_sign(vel[a]);
      vel[a] *= op->p[0];
    }
  }
  double pos[2] = {DYN(e, MOOG_D_X, s), DYN(e, MOOG_D_Y, s)};
  double aff[2][2];
  if (!maze_affordances(&m, pos, aff)) {
    e->envi[MOOG_EI_ERR] |= MOOG_ERR_OFF_MAZE_GRID;
    return;
  }
  /* sprite.position = np.copy(position): a translation by exactly zero */
  set_position(e, s, pos[0], pos[1]);
  double nv[2];
  if (!maze_new_velocity(&m, pos, vel, aff, -1, nv, 0)) {
    e->envi[MOOG_EI_ERR] |= MOOG_ERR_OFF_MAZE_GRID;
    return;
  }
  double new_angle;
  if (nv[0] == 0 && nv[1] == 0) {
    new_angle = NAN;
  } else if (nv[1] == 0) {
    new_angle = -0.5 * np_sign(nv[0]) * M_PI;
  } else if (np_sign(nv[1]) > 0) {
    new_angle = atan(-nv[0] / nv[1]);
  } else {
    new_angle = M_PI + atan(-nv[0] / nv[1]);
  }
  if (!isnan(new_angle) && fabs(new_angle - DYN(e, MOOG_D_ANG, s)) > MAZE_EPS) {
    set_angle(e, s, new_angle, 0);
    set_ang_kind(e, s, KIND_F64);
  }
  assign_velocity(e, s, nv[0], nv[1]);
}

static void corrective(env_t *e, const moog_op *op) {
  int sp[MOOG_MAX_SLOTS];
  switch (op->kind) {
    case MOOG_C_MAZE_PHYSICS: { /* maze_physics.py:205-211 */
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      for (int i = 0; i < n; ++i) maze_physics_sprite(e, op, sp[i]);
      break;
    }
    case MOOG_C_TETHER: { /* tether_physics.py:126-140 */
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      tether_sprites(e, sp, n, (op->flags & MOOG_FL_UPDATE_ANGLE_VEL) != 0,
                     (op->flags & MOOG_FL_HAS_ANCHOR) != 0, op->p[0], op->p[1]);
      break;
    }
    case MOOG_C_TETHER_ZIPPED: { /* tether_physics.py:186-201 */
      int nl = op->i[1];
      int c0 = nl ? e->cnt[e->ipool[op->i[0]]] : 0;
      for (int q = 1; q < nl; ++q)
        if (e->cnt[e->ipool[op->i[0] + q]] != c0) {
          e->envi[MOOG_EI_ERR] |= MOOG_ERR_TETHER_ZIP;
          return;
        }
      for (int i = 0; i < c0; ++i) {
        for (int q = 0; q < nl; ++q) sp[q] = LOFF(e, e->ipool[op->i[0] + q]) + i;
        tether_sprites(e, sp, nl, (op->flags & MOOG_FL_UPDATE_ANGLE_VEL) != 0,
                       (op->flags & MOOG_FL_HAS_ANCHOR) != 0, op->p[0], op->p[1]);
      }
      break;
    }
    case MOOG_C_CONSTANT_SPEED: { /* constant_speed.py:34-46 */
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      for (int i = 0; i < n; ++i) {
        double vx = DYN(e, MOOG_D_VX, sp[i]), vy = DYN(e, MOOG_D_VY, sp[i]);
        if (vel32(e, sp[i])) {
          /* a float32 velocity array: np.linalg.norm is sqrt(x.dot(x)) in float32 (the two
           * products and the sum each rounded to float32, no fma), and
           * `speed * velocity / norm` stays float32 (the python float is a weak scalar) */
          float fx = (float)vx, fy = (float)vy;
          float xx = fx * fx, yy = fy * fy;
          float nf = sqrtf(xx + yy);
          if (nf != 0) {
            float sp32 = (float)op->p[0];
            float ax = sp32 * fx, ay = sp32 * fy;
            assign_velocity(e, sp[i], (double)(ax / nf), (double)(ay / nf));
            META(e, MOOG_M_FLAGS, sp[i]) |= MOOG_SF_VEL32;
          }
          continue;
        }
        double nv = norm1(vx, vy);
        if (nv != 0) { /* NaN is truthy in Python, and NaN != 0 here too */
          assign_velocity(e, sp[i], op->p[0] * vx / nv, op->p[0] * vy / nv);
        }
      }
      break;
    }
  }
}

/* Row-order experiment: do the rows of one Collision entry commute the way the next kernel
 * design assumes ("Next" (1))?  The reference visits the pairs (i, j) of an entry in
 * itertools.product order (physics.py:92-108).  A visit can only act when the circle test of
 * sprite.py:464-466 passes, it reads both sprites and writes sprite_0 (and sprite_1 when the
 * entry is symmetric).  With near(i) = the sprites of layer 1 within reach of row i's sprite plus
 * a margin for what contacts may move them during this entry, row k may run before / next to an
 * earlier row i when
 *     asymmetric entry:  s0(k) not in near(i)  and  s0(i) not in near(k)
 *     symmetric entry:   ({s0(i)} + near(i)) and ({s0(k)} + near(k)) are disjoint.
 * This function executes the entry in waves: a wave holds every remaining row without an earlier
 * remaining row in conflict, and the rows of a wave are processed in REVERSE order.  If the rule
 * holds the state after the entry is bit-identical to the reference order (the golden tests run
 * in this mode too); the order-sensitive hash is folded from the per-row logs in row order.
 * The margin of a sprite is a fixed allowance plus twice the way it travels per substep;
 * g_row_stats[3] counts sprites that moved further than that during one entry (a kernel would
 * have to fall back to the reference order for such an entry). */
#define ORC_ROW_MARGIN 0.02
static void collision_op_rows(env_t *e, const moog_op *op) {
  const int la = op->i[0], lb = op->i[1];
  const int na = e->cnt[la], nb = e->cnt[lb], sa = LOFF(e, la), sb = LOFF(e, lb), S = e->S;
  const int symmetric = (op->flags & MOOG_FL_SYMMETRIC) != 0;
  if (na <= 1) {
    for (int i = 0; i < na; ++i)
      for (int j = 0; j < nb; ++j) collision_step(e, op, sa + i, sb + j, 0);
    return;
  }
  uint8_t *near = (uint8_t *)calloc((size_t)na * (size_t)S, 1); /* near[i][slot] incl. the row's own sprite */
  uint8_t *done = (uint8_t *)calloc((size_t)na, 1);
  int *wave = (int *)malloc(sizeof(int) * (size_t)na);
  double *p0 = (double *)malloc(sizeof(double) * 3 * (size_t)S); /* position and margin at entry start */
  for (int s = 0; s < S; ++s) {
    p0[3 * s] = DYN(e, MOOG_D_X, s);
    p0[3 * s + 1] = DYN(e, MOOG_D_Y, s);
    /* what a contact may move a sprite: a fixed allowance plus twice the way it travels per substep */
    const double v = norm1(DYN(e, MOOG_D_VX, s), DYN(e, MOOG_D_VY, s));
    p0[3 * s + 2] = ORC_ROW_MARGIN + (isfinite(v) ? 2.0 * v / (double)e->K : INFINITY);
  }
  for (int i = 0; i < na; ++i) {
    const int s0 = sa + i;
    near[(size_t)i * S + s0] = 1;
    for (int j = 0; j < nb; ++j) {
      const int s1 = sb + j;
      if (s1 == s0) continue;
      const double d = norm1(DYN(e, MOOG_D_X, s0) - DYN(e, MOOG_D_X, s1), DYN(e, MOOG_D_Y, s0) - DYN(e, MOOG_D_Y, s1));
      /* NaN compares false: a sprite without a position is near everything */
      if (!(d > STAT(e, MOOG_S_MAXR, s0) + STAT(e, MOOG_S_MAXR, s1) + p0[3 * s0 + 2] + p0[3 * s1 + 2]))
        near[(size_t)i * S + s1] = 1;
    }
  }
  g_row_logs = (row_log_t *)calloc((size_t)na, sizeof(row_log_t));
  int remaining = na;
  while (remaining > 0) {
    int nw = 0;
    for (int k = 0; k < na; ++k) {
      if (done[k]) continue;
      int ok = 1;
      for (int i = 0; i < k && ok; ++i) {
        if (done[i]) continue;
        const uint8_t *ni = near + (size_t)i * S, *nk = near + (size_t)k * S;
        if (symmetric) {
          for (int s = 0; s < S && ok; ++s) ok = !(ni[s] && nk[s]);
        } else {
          ok = !(ni[sa + k] || nk[sa + i]);
        }
      }
      if (ok) wave[nw++] = k;
    }
    for (int w = nw - 1; w >= 0; --w) { /* any order inside a wave must do: take the reverse one */
      const int i = wave[w];
      g_row_cur = i;
      for (int j = 0; j < nb; ++j) collision_step(e, op, sa + i, sb + j, 0);
      g_row_cur = -1;
    }
    for (int w = 0; w < nw; ++w) done[wave[w]] = 1;
    remaining -= nw;
    g_row_stats[2]++;
  }
  for (int i = 0; i < na; ++i) { /* the hash as the reference order would have folded it */
    for (int q = 0; q < g_row_logs[i].n; ++q) {
      const int a = g_row_logs[i].ab[2 * q], b = g_row_logs[i].ab[2 * q + 1];
      e->overlap_hash = (e->overlap_hash ^ (uint64_t)((a * 1315423911u) ^ (b * 2654435761u) ^ 1u)) * 1099511628211ull;
    }
    free(g_row_logs[i].ab);
  }
  free(g_row_logs);
  g_row_logs = NULL;
  for (int s = 0; s < S; ++s) {
    const double d = norm1(DYN(e, MOOG_D_X, s) - p0[3 * s], DYN(e, MOOG_D_Y, s) - p0[3 * s + 1]);
    if (d > p0[3 * s + 2]) g_row_stats[3]++;
  }
  g_row_stats[0]++;
  g_row_stats[1] += na;
  free(near); free(done); free(wave); free(p0);
}

/* physics.py:88-117 Physics.apply_physics (one substep) */
static void apply_physics(env_t *e) {
  const int32_t *h = e->hdr;
  for (int f = 0; f < h[MOOG_H_N_FORCES]; ++f) {
    const moog_op *op = e->ops + h[MOOG_H_FORCES] + f;
    int la = op->i[0], lb = op->i[1];
    if (op->kind == MOOG_F_MAZE_WALK) {
      for (int i = 0; i < e->cnt[la]; ++i) maze_walk_sprite(e, op, LOFF(e, la) + i, i);
    } else if (lb < 0) {
      for (int i = 0; i < e->cnt[la]; ++i) force_unary(e, op, LOFF(e, la) + i, i);
    } else if (op->kind == MOOG_F_COLLISION && g_row_mode) {
      collision_op_rows(e, op);
    } else {
      for (int i = 0; i < e->cnt[la]; ++i)
        for (int j = 0; j < e->cnt[lb]; ++j) {
          int s0 = LOFF(e, la) + i, s1 = LOFF(e, lb) + j;
          if (op->kind == MOOG_F_COLLISION)
            collision_step(e, op, s0, s1, 0);
          else
            force_binary(e, op, s0, s1);
        }
    }
  }
  for (int c = 0; c < h[MOOG_H_N_CORR]; ++c) corrective(e, e->ops + h[MOOG_H_CORR] + c);
  /* physics.py:113-117 + sprite.py:426-430 */
  double dt = 1. / e->K;
  for (int l = 0; l < e->L; ++l)
    for (int i = 0; i < e->cnt[l]; ++i) {
      int s = LOFF(e, l) + i;
      /* sprite.py:426-430; `delta_t * velocity` is a float32 product for a
       * float32 velocity array (dt is a weak python float) */
      double dx = vel32(e, s) ? f32mul(dt, DYN(e, MOOG_D_VX, s)) : dt * DYN(e, MOOG_D_VX, s);
      double dy = vel32(e, s) ? f32mul(dt, DYN(e, MOOG_D_VY, s)) : dt * DYN(e, MOOG_D_VY, s);
      set_position(e, s, DYN(e, MOOG_D_X, s) + dx, DYN(e, MOOG_D_Y, s) + dy);
      double w = DYN(e, MOOG_D_ANGVEL, s);
      if (w != 0) {
        int wk = angvel_kind(e, s), ak = ang_kind(e, s);
        double t = (wk == KIND_F32) ? f32mul(dt, w) : dt * w; /* kind(t) == wk */
        double a = DYN(e, MOOG_D_ANG, s), na;
        int nk;
        if (ak == KIND_F64 || wk == KIND_F64) { na = a + t; nk = KIND_F64; }
        else if (ak == KIND_WEAK && wk == KIND_WEAK) { na = a + t; nk = KIND_WEAK; }
        else { na = f32add(a, t); nk = KIND_F32; }
        set_angle(e, s, na, nk == KIND_F32);
        set_ang_kind(e, s, nk);
      }
    }
}

/* ------------------------------------------------------------------------ */
/* expression VM (config lambdas compiled by the host)                       */
/* ------------------------------------------------------------------------ */

static double *attr_ptr(env_t *e, int s, int at) {
  if (at >= MOOG_AT_META0) /* sprite.metadata[key], column at - MOOG_AT_META0 */
    return &e->envf[e->hdr[MOOG_H_META_OFF] + (at - MOOG_AT_META0) * e->S + s];
  switch (at) {
    case MOOG_AT_X: return &DYN(e, MOOG_D_X, s);
    case MOOG_AT_Y: return &DYN(e, MOOG_D_Y, s);
    case MOOG_AT_X_VEL: return &DYN(e, MOOG_D_VX, s);
    case MOOG_AT_Y_VEL: return &DYN(e, MOOG_D_VY, s);
    case MOOG_AT_ANGLE: return &DYN(e, MOOG_D_ANG, s);
    case MOOG_AT_ANGLE_VEL: return &DYN(e, MOOG_D_ANGVEL, s);
    case MOOG_AT_MASS: return &STAT(e, MOOG_S_MASS, s);
    case MOOG_AT_SCALE: return &STAT(e, MOOG_S_SCALE, s);
    case MOOG_AT_ASPECT_RATIO: return &STAT(e, MOOG_S_ASPECT, s);
    case MOOG_AT_C0: return &STAT(e, MOOG_S_C0, s);
    case MOOG_AT_C1: return &STAT(e, MOOG_S_C1, s);
    case MOOG_AT_C2: return &STAT(e, MOOG_S_C2, s);
    case MOOG_AT_OPACITY: return &STAT(e, MOOG_S_OPACITY, s);
  }
  return NULL;
}

static double py_fmod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0) {
    if ((r < 0) != (b < 0)) r += b;
  } else {
    r = copysign(0.0, b); /* CPython float_rem / numpy remainder */
  }
  return r;
}

/* sprite.py:411-424 Sprite._set_path, what the `scale` / `aspect_ratio` setters (:546-558) call: the
 * outline is re-derived from the COM-centred shape (shape record of the blob),
 *   Affine2D().scale(s, s * aspect) + Affine2D().rotate(angle) + Affine2D().translate(*position),
 * the circumscribed radius is re-measured and the rotational inertia is multiplied by the square
 * of the scales -- AGAIN, on every call: the inertia compounds (Appendix A2 of SURVEY.md). */
static void set_path(env_t *e, int s) {
  const int32_t *shape_off = e->ipool + e->hdr[MOOG_H_SHAPE_TAB];
  const double *R = e->dpool + shape_off[META(e, MOOG_M_SHAPE, s)];
  const int nv = (int)R[0];
  const double sx = STAT(e, MOOG_S_SCALE, s), sy = STAT(e, MOOG_S_SCALE, s) * STAT(e, MOOG_S_ASPECT, s);
  const double px = DYN(e, MOOG_D_X, s), py = DYN(e, MOOG_D_Y, s);
  double S[9], Rm[9], T[9], SR[9], M[9];
  aff_identity(S);
  S[0] *= sx; S[1] *= sx; S[2] *= sx;
  S[3] *= sy; S[4] *= sy; S[5] *= sy;
  aff_identity(Rm);
  aff_rotate(Rm, DYN(e, MOOG_D_ANG, s));
  aff_identity(T);
  aff_translate(T, px, py);
  aff_then(S, Rm, SR);
  aff_then(SR, T, M);
  double *v = e->vtx + 2 * (size_t)e->voff[s];
  double r = -INFINITY;
  for (int i = 0; i < nv; ++i) {
    const double bx = R[6 + 2 * i], by = R[7 + 2 * i];
    const double wx = M[0] * bx + M[1] * by + M[2], wy = M[3] * bx + M[4] * by + M[5];
    v[2 * i] = wx;
    v[2 * i + 1] = wy;
    const double d = norm_ax(wx - px, wy - py);
    if (d > r || isnan(d)) r = d; /* np.max propagates NaN */
  }
  META(e, MOOG_M_NV, s) = nv;
  STAT(e, MOOG_S_MAXR, s) = r;
  STAT(e, MOOG_S_IX, s) *= sx * sx;
  STAT(e, MOOG_S_IY, s) *= sy * sy;
}

static double rule_noise_at(env_t *e, int col);
/* Runs the postfix program at `start`; returns top of stack (or 1.0 when start < 0). */
static double eval_expr(env_t *e, int start, int s0, int s1) {
  if (start < 0) return 1.0;
  double st[16];
  int sp = 0;
  for (const moog_ex *x = e->expr + start; x->op != MOOG_X_END; ++x) {
    double a, b;
    switch (x->op) {
      case MOOG_X_CONST: st[sp++] = x->c; break;
      case MOOG_X_ATTR0: st[sp++] = *attr_ptr(e, s0, x->arg); break;
      case MOOG_X_ATTR1: st[sp++] = *attr_ptr(e, s1, x->arg); break;
      case MOOG_X_NOT: st[sp - 1] = !(st[sp - 1] != 0); break;
      case MOOG_X_NEG: st[sp - 1] = -st[sp - 1]; break;
      case MOOG_X_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
      case MOOG_X_STORE: {
        const double v = st[--sp];
        if (x->arg == MOOG_AT_ANGLE) { /* sprite.py:531-540; x->c: kind of the assigned value */
          const int kept = (META(e, MOOG_M_FLAGS, s0) >> MOOG_SF_ANG_SHIFT) & 3;
          set_angle(e, s0, v, 0);
          /* c = 4: computed from the sprite's own angle and python floats only -- the NumPy kind it has stays */
          set_ang_kind(e, s0, x->c == 4.0 ? kept : (int)x->c);
        } else {
          *attr_ptr(e, s0, x->arg) = v;
          if (x->arg == MOOG_AT_SCALE || x->arg == MOOG_AT_ASPECT_RATIO) set_path(e, s0);
          /* sprite.py:639-643: the velocity setter installs a fresh array; x->c = 3: a float64 one */
          if ((x->arg == MOOG_AT_X_VEL || x->arg == MOOG_AT_Y_VEL) && x->c == 3.0)
            META(e, MOOG_M_FLAGS, s0) &= ~(MOOG_SF_VEL32 | (MOOG_SF_VALIAS_MASK << MOOG_SF_VALIAS_SHIFT));
        }
        break;
      }
      case MOOG_X_STORE_POS: {
        double ny = st[--sp], nx = st[--sp];
        set_position(e, s0, nx, ny);
        break;
      }
      case MOOG_X_RULE_NOISE: st[sp++] = rule_noise_at(e, x->arg); break; /* a draw of a traced rule */
      case MOOG_X_NORM2: { /* np.linalg.norm of a 2-vector */
        const double vy = st[--sp], vx = st[--sp];
        st[sp++] = norm1(vx, vy);
        break;
      }
      case MOOG_X_ENVF: st[sp++] = e->envf[x->arg]; break; /* a user-defined rule's own attribute */
      case MOOG_X_STORE_ENVF: e->envf[x->arg] = st[--sp]; break;
      case MOOG_X_SELECT: { /* an `if` / `else` of the config callable on a per-sprite value */
        const double vb = st[--sp], va = st[--sp], vc = st[--sp];
        st[sp++] = vc != 0 ? va : vb;
        break;
      }
      default:
        b = st[--sp];
        a = st[--sp];
        switch (x->op) {
          case MOOG_X_LT: a = a < b; break;
          case MOOG_X_LE: a = a <= b; break;
          case MOOG_X_GT: a = a > b; break;
          case MOOG_X_GE: a = a >= b; break;
          case MOOG_X_EQ: a = a == b; break;
          case MOOG_X_NE: a = a != b; break;
          case MOOG_X_AND: a = (a != 0) && (b != 0); break;
          case MOOG_X_OR: a = (a != 0) || (b != 0); break;
          case MOOG_X_ADD: a = a + b; break;
          case MOOG_X_SUB: a = a - b; break;
          case MOOG_X_MUL: a = a * b; break;
          case MOOG_X_DIV: a = a / b; break;
          case MOOG_X_MOD: a = py_fmod(a, b); break;
        }
        st[sp++] = a;
    }
  }
  return sp ? st[sp - 1] : 1.0;
}

/* state conditions */
/* A decision tree of lambdas.state_tree / lambdas.trace_rule (MOOG_SC_TREE, MOOG_R_TREE), walked lazily from
 * node 0: the tests made -- overlap calls included -- are the ones Python would make, in its order.  Node: kind,
 * expr, layer / index of sprite 0, layer / index of sprite 1, next if true, next if false.  A sprite index
 * beyond its layer's count is the reference's IndexError (MOOG_ERR_BAD_INDEX). */
static double eval_expr(env_t *e, int start, int s0, int s1);
static double walk_tree(env_t *e, const int32_t *nodes, int n) {
  int j = 0;
  for (int guard = 0; guard <= n; ++guard) {
    const int32_t *nd = nodes + 8 * j;
    if (nd[0] == 3) { /* index < len(state[layer]) */
      j = nd[3] < e->cnt[nd[2]] ? nd[6] : nd[7];
      continue;
    }
    int sl[2] = {0, 0};
    for (int q = 0; q < 2; ++q) {
      const int l = nd[2 + 2 * q], k = nd[3 + 2 * q];
      if (l < 0) continue;
      if (k >= e->cnt[l]) {
        e->envi[MOOG_EI_ERR] |= MOOG_ERR_BAD_INDEX;
        return 0;
      }
      sl[q] = LOFF(e, l) + k;
    }
    if (nd[0] == 0) return nd[1] >= 0 ? eval_expr(e, nd[1], sl[0], sl[1]) : 0.0; /* leaf: the value / end of the rule */
    if (nd[0] == 4) { /* the assignments this path made */
      eval_expr(e, nd[1], sl[0], sl[1]);
      j = nd[6];
      continue;
    }
    const int yes = nd[0] == 2 ? overlaps(e, sl[0], sl[1]) : eval_expr(e, nd[1], sl[0], sl[1]) != 0;
    j = yes ? nd[6] : nd[7];
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* counter-based draws (Philox4x32-10)                                        */
/*                                                                            */
/* The reference draws from np.random's global Mersenne Twister, in Python   */
/* call order; a batch of envs stepped concurrently cannot share that stream, */
/* so the CUDA path keys every draw by (seed, env, counters).  This is the    */
/* same generator restated, so that trajectories with CreateSprites / random  */
/* conditions can be compared draw for draw; that the DISTRIBUTIONS are the   */
/* reference's is what tests/test_create_sprites.py checks against it.        */
/* ------------------------------------------------------------------------ */
static void set_path(env_t *e, int s);
static uint64_t g_seed = 0;
void orc_set_seed(uint64_t seed) { g_seed = seed; }

/* Replay hook of the parity tests: rows of MOOG_Z_N_ATTRS factors that the next generate_sprites tries
 * take, in order, INSTEAD of drawing them -- the factor dicts the reference's own factor_dist.sample()
 * returned on a recorded trajectory (oracle/gen_golden_spawn.py).  With it the oracle follows the
 * reference through CreateSprites draw for draw, which pins everything around the draw itself. */
static const double *g_forced = NULL;
static int g_forced_n = 0, g_forced_pos = 0;
void orc_force_factors(const double *rows, int n_rows) {
  g_forced = rows;
  g_forced_n = n_rows;
  g_forced_pos = 0;
}
int orc_forced_left(void) { return g_forced_n - g_forced_pos; }

static double philox_uniform(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    const uint64_t m0 = (uint64_t)0xD2511F53u * c0, m1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(m1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)m1, n2 = (uint32_t)(m0 >> 32) ^ c3 ^ k1,
                   n3 = (uint32_t)m0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint64_t bits = ((uint64_t)c0 << 21) ^ (uint64_t)(c1 >> 11);
  return (double)(bits & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
}

/* the uniform of one rule-noise column for the rules pass being made */
static double rule_noise_at(env_t *e, int col) {
  if (e->rule_noise) return e->rule_noise[col];
  return philox_uniform(g_seed ^ 0x9E3779B97F4A7C15ull, (uint32_t)e->env_id, (uint32_t)e->envi[MOOG_EI_RULE_PASSES],
                        (uint32_t)e->envi[MOOG_EI_EPISODES], (uint32_t)col);
}

/* one factor from a leaf sampler: a constant, np.float32(rng.uniform(lo, hi)) (distributions.py:91-93),
 * or one of n candidates (:127-129) */
static double sample_leaf(const double *dpool, int kind, int idx, int n, double u) {
  if (kind == MOOG_ZK_CONST) return dpool[idx];
  if (kind == MOOG_ZK_UNIFORM32) return (double)(float)(dpool[idx] + (dpool[idx + 1] - dpool[idx]) * u);
  if (kind == MOOG_ZK_DISCRETE_P) {  /* rng.choice(n, p=probs): the first candidate whose cumulative probability exceeds u */
    int k = 0;
    while (k < n - 1 && !(u < dpool[idx + n + k])) ++k;
    return dpool[idx + k];
  }
  const int pick = (int)(u * n);
  return dpool[idx + (pick < n ? pick : n - 1)];
}

/* a DependentDistribution's function (distributions.py:420-470) over the factors drawn so far */
static double eval_factor_expr(const moog_ex *x, const double *v) {
  double st[16];
  int sp = 0;
  for (; x->op != MOOG_X_END && sp < 15; ++x) {
    double a, b;
    switch (x->op) {
      case MOOG_X_CONST: st[sp++] = x->c; break;
      case MOOG_X_ATTR0: st[sp++] = v[x->arg]; break;
      case MOOG_X_NOT: st[sp - 1] = !(st[sp - 1] != 0); break;
      case MOOG_X_NEG: st[sp - 1] = -st[sp - 1]; break;
      case MOOG_X_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
      default:
        if (sp < 2) return NAN;
        b = st[--sp];
        a = st[--sp];
        switch (x->op) {
          case MOOG_X_LT: a = a < b; break;
          case MOOG_X_LE: a = a <= b; break;
          case MOOG_X_GT: a = a > b; break;
          case MOOG_X_GE: a = a >= b; break;
          case MOOG_X_EQ: a = a == b; break;
          case MOOG_X_NE: a = a != b; break;
          case MOOG_X_AND: a = (a != 0) && (b != 0); break;
          case MOOG_X_OR: a = (a != 0) || (b != 0); break;
          case MOOG_X_ADD: a = a + b; break;
          case MOOG_X_SUB: a = a - b; break;
          case MOOG_X_MUL: a = a * b; break;
          case MOOG_X_DIV: a = a / b; break;
          default: a = NAN; break;
        }
        st[sp++] = a;
    }
  }
  return sp ? st[sp - 1] : NAN;
}

/* One call of the closure generate_sprites returns (sprite_generators.py:75-103): `count` sprites
 * drawn from the op's factor table (i4; extension components: Mixture :278-306, SetMinus :325-349,
 * Selection / Intersection :376-404, DependentDistribution :420-470) go to slots first.. of `layer`;
 * each is built like Sprite.__init__ (sprite.py:261-424) and redrawn while it overlaps a sprite it
 * must avoid -- slots avoid[..], or (avoid_layers, create_sprites.py:31-33) the sprites that the
 * layers avoid[..] held when the call began.  Draw k of sprite s, try t: Philox(key; env, c1, s << 20 | t, tag | k). */
static void generate_sprites(env_t *e, const moog_op *op, uint64_t key, uint32_t c1, int layer, int first,
                             int count, const int32_t *avoid, int n_avoid, int avoid_layers) {
  const double *dpool = e->dpool;
  const int32_t *shape_off = e->ipool + e->hdr[MOOG_H_SHAPE_TAB];
  const int32_t *tab = e->ipool + op->i[4];
  const int max_depth = (int)fmin(op->p[0], 1048575.0);
  int placed = 0;
  for (int k = 0; k < count; ++k) {
    const int s = first + k;
    int stop = 0;
    for (int tries = 0;; ++tries) {
      double v[MOOG_Z_N_ATTRS];
      const uint32_t c2 = ((uint32_t)s << 20) | (uint32_t)tries;
      const int forced = g_forced != NULL && g_forced_pos < g_forced_n;
      if (forced) memcpy(v, g_forced + (size_t)MOOG_Z_N_ATTRS * g_forced_pos++, sizeof(v));
      for (int a = 0; a < MOOG_Z_N_ATTRS && !forced; ++a) {
        const int kind = tab[3 * a], idx = tab[3 * a + 1], n = tab[3 * a + 2];
        const double u = kind == MOOG_ZK_CONST ? 0.0 : philox_uniform(key, (uint32_t)e->env_id, c1, c2, (0x5Au << 24) | (uint32_t)a);
        v[a] = sample_leaf(dpool, kind, idx, n, u);
      }
      const int32_t *x = tab + 3 * MOOG_Z_N_ATTRS;
      const int n_ext = forced ? 0 : *x++;
      uint32_t draw = 0;
      for (int c = 0; c < n_ext; ++c) {
        const int kind = *x++;
        if (kind == 3) {
          const int n_dep = *x++;
          for (int q = 0; q < n_dep; ++q, x += 3) {
            const double val = eval_factor_expr(e->expr + x[1], v);
            v[x[0]] = x[2] ? (double)(float)val : val;
          }
        } else if (kind == 1) {
          const int n_alt = *x++;
          const double *cum = dpool + *x++;
          const double u = philox_uniform(key, (uint32_t)e->env_id, c1, c2, (0x5Bu << 24) | (draw++ & 0xffffffu));
          int pick = 0;
          while (pick < n_alt - 1 && !(u < cum[pick])) ++pick; /* rng.choice(n, p=probs) */
          for (int a = 0; a < n_alt; ++a) {
            const int n_leaves = *x++;
            for (int q = 0; q < n_leaves; ++q, x += 4) {
              if (a != pick) continue;
              const double uu = philox_uniform(key, (uint32_t)e->env_id, c1, c2, (0x5Bu << 24) | (draw++ & 0xffffffu));
              v[x[0]] = sample_leaf(dpool, x[1], x[2], x[3], uu);
            }
          }
        } else {
          const int keep_inside = *x++;
          const int n_leaves = *x++;
          const int32_t *leaves = x;
          x += 4 * n_leaves;
          const int n_box = *x++;
          const int32_t *box = x;
          x += 2 * n_box;
          int accepted = 0; /* at most _MAX_TRIES = 1e5 redraws, then the reference raises */
          for (int inner = 0; inner < 100000 && !accepted; ++inner) {
            for (int q = 0; q < n_leaves; ++q) {
              const double uu = philox_uniform(key, (uint32_t)e->env_id, c1, c2, (0x5Bu << 24) | (draw++ & 0xffffffu));
              v[leaves[4 * q]] = sample_leaf(dpool, leaves[4 * q + 1], leaves[4 * q + 2], leaves[4 * q + 3], uu);
            }
            int inside = 1;
            for (int q = 0; q < n_box; ++q) {
              const double val = v[box[2 * q]], lo = dpool[box[2 * q + 1]], hi = dpool[box[2 * q + 1] + 1];
              inside = inside && val >= lo && val < hi; /* Continuous.contains */
            }
            accepted = inside == (keep_inside != 0);
          }
          if (!accepted) e->envi[MOOG_EI_ERR] |= MOOG_ERR_RESET_REJECTED;
        }
      }
      /* Sprite.__init__ (sprite.py:261-327) then the shape setter (:329-409): the outline is laid out at
       * (x, y) by _set_path (:411-424: circumscribed radius, inertia * scale^2), THEN the position setter
       * (:616-633) moves sprite and cached outline by the shape's raw centroid */
      const double *R = dpool + shape_off[(int)v[MOOG_Z_SHAPE_ATTR]];
      DYN(e, MOOG_D_X, s) = v[MOOG_AT_X]; DYN(e, MOOG_D_Y, s) = v[MOOG_AT_Y];
      DYN(e, MOOG_D_VX, s) = v[MOOG_AT_X_VEL]; DYN(e, MOOG_D_VY, s) = v[MOOG_AT_Y_VEL];
      DYN(e, MOOG_D_ANG, s) = v[MOOG_AT_ANGLE]; DYN(e, MOOG_D_ANGVEL, s) = v[MOOG_AT_ANGLE_VEL];
      STAT(e, MOOG_S_MASS, s) = v[MOOG_AT_MASS]; STAT(e, MOOG_S_SCALE, s) = v[MOOG_AT_SCALE];
      STAT(e, MOOG_S_ASPECT, s) = v[MOOG_AT_ASPECT_RATIO];
      STAT(e, MOOG_S_IX, s) = R[2]; STAT(e, MOOG_S_IY, s) = R[3];
      STAT(e, MOOG_S_C0, s) = v[MOOG_AT_C0]; STAT(e, MOOG_S_C1, s) = v[MOOG_AT_C1];
      STAT(e, MOOG_S_C2, s) = v[MOOG_AT_C2]; STAT(e, MOOG_S_OPACITY, s) = v[MOOG_AT_OPACITY];
      META(e, MOOG_M_SHAPE, s) = (int)v[MOOG_Z_SHAPE_ATTR];
      META(e, MOOG_M_FLAGS, s) = op->i[5] | (R[1] != 0.0 ? MOOG_SF_CIRCLE : 0);
      META(e, MOOG_M_NV, s) = (int)R[0];
      e->cnt[layer] = s - LOFF(e, layer) + 1;
      for (int m = 0; m < e->hdr[MOOG_H_N_META]; ++m) /* a new sprite's metadata is {} */
        e->envf[e->hdr[MOOG_H_META_OFF] + m * e->S + s] = NAN;
      set_path(e, s);
      set_position(e, s, v[MOOG_AT_X] + R[4], v[MOOG_AT_Y] + R[5]);
      int hit = 0; /* every pair is evaluated (no short-circuit), sprite_generators.py:69-74 */
      if (avoid_layers) {
        for (int q = 0; q < n_avoid; ++q) {
          const int la = avoid[q], n_la = la == layer ? first - LOFF(e, la) : e->cnt[la];
          for (int j = 0; j < n_la; ++j) hit |= overlaps(e, s, LOFF(e, la) + j);
        }
      } else {
        for (int q = 0; q < n_avoid; ++q) hit |= overlaps(e, s, avoid[q]);
      }
      if (op->flags & MOOG_FL_DISJOINT)
        for (int j = first; j < s; ++j) hit |= overlaps(e, s, j);
      if (!hit) break;
      if (tries > max_depth) { /* sprite_generators.py:92-98 */
        if (op->flags & MOOG_FL_FAIL_GRACEFULLY)
          stop = 1;
        else
          e->envi[MOOG_EI_ERR] |= MOOG_ERR_RESET_REJECTED;
        break;
      }
    }
    if (stop) break;
    placed = k + 1;
  }
  e->cnt[layer] = first - LOFF(e, layer) + placed;
}

static double eval_condition(env_t *e, int op_index) {
  const moog_op *op = e->ops + op_index;
  int sp[MOOG_MAX_SLOTS], sq[MOOG_MAX_SLOTS];
  switch (op->kind) {
    case MOOG_SC_CONST: return op->p[0];
    case MOOG_SC_BERNOULLI: return rule_noise_at(e, op->i[0]) < op->p[0]; /* np.random.binomial(1, p) */
    case MOOG_SC_TREE: return walk_tree(e, e->ipool + op->i[0], op->i[1]);
    case MOOG_SC_ALL:
    case MOOG_SC_ANY:
    case MOOG_SC_COUNT: {
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      int cnt = 0;
      for (int i = 0; i < n; ++i) cnt += eval_expr(e, op->i[2], sp[i], sp[i]) != 0;
      if (op->kind == MOOG_SC_ALL) return cnt == n;
      if (op->kind == MOOG_SC_ANY) return cnt > 0;
      return cnt;
    }
    case MOOG_SC_CONTACT_COUNT: { /* contact_rules.py:15-51 */
      int la = op->i[0], lb = op->i[1], cnt = 0;
      for (int i = 0; i < e->cnt[la]; ++i)
        for (int j = 0; j < e->cnt[lb]; ++j) cnt += overlaps(e, LOFF(e, la) + i, LOFF(e, lb) + j);
      return cnt;
    }
    case MOOG_SC_CONTACT_ANY_COUNT: {
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      int m = gather_layers(e, op->i[2], op->i[3], sq);
      int cnt = 0;
      for (int i = 0; i < n; ++i) {
        if (eval_expr(e, op->i[4], sp[i], sp[i]) == 0) continue;
        int any = 0;
        for (int j = 0; j < m && !any; ++j) any = overlaps(e, sp[i], sq[j]); /* python `or` short-circuits */
        cnt += any;
      }
      return cnt;
    }
    case MOOG_SC_BINARY: {
      double a = eval_condition(e, op->i[0]);
      if (op->i[2] == MOOG_X_AND) return a != 0 ? eval_condition(e, op->i[1]) : a;
      if (op->i[2] == MOOG_X_OR) return a != 0 ? a : eval_condition(e, op->i[1]);
      double b = eval_condition(e, op->i[1]);
      switch (op->i[2]) {
        case MOOG_X_LT: return a < b;
        case MOOG_X_LE: return a <= b;
        case MOOG_X_GT: return a > b;
        case MOOG_X_GE: return a >= b;
        case MOOG_X_EQ: return a == b;
        case MOOG_X_NE: return a != b;
        case MOOG_X_ADD: return a + b;
        case MOOG_X_SUB: return a - b;
        case MOOG_X_MUL: return a * b;
      }
      return 0;
    }
    case MOOG_SC_NOT: return !(eval_condition(e, op->i[0]) != 0);
    case MOOG_SC_FIRST: {
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      return n > 0 ? eval_expr(e, op->i[2], sp[0], sp[0]) : 0.0;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* rules                                                                     */
/* ------------------------------------------------------------------------ */

static void copy_slot(env_t *e, int dst, int src) {
  for (int f = 0; f < MOOG_DYN_FIELDS; ++f) DYN(e, f, dst) = DYN(e, f, src);
  for (int f = 0; f < MOOG_STAT_FIELDS; ++f) STAT(e, f, dst) = STAT(e, f, src);
  for (int f = 0; f < MOOG_META_FIELDS; ++f) META(e, f, dst) = META(e, f, src);
  for (int m = 0; m < e->hdr[MOOG_H_N_META]; ++m) { /* the sprite's metadata columns travel with it */
    double *col = e->envf + e->hdr[MOOG_H_META_OFF] + m * e->S;
    col[dst] = col[src];
  }
  memcpy(e->vtx + 2 * (size_t)e->voff[dst], e->vtx + 2 * (size_t)e->voff[src],
         sizeof(double) * 2 * META(e, MOOG_M_NV, src));
}

/* vanish.py:31-39: pop the flagged indices of layer l, preserving order */
static void vanish(env_t *e, int l, const uint8_t *gone) {
  int base = LOFF(e, l), n = e->cnt[l], w = 0;
  for (int i = 0; i < n; ++i) {
    if (gone[i]) continue;
    if (w != i) copy_slot(e, base + w, base + i);
    ++w;
  }
  e->cnt[l] = w;
}

static int rule_step(env_t *e, int r, const double *rule_noise); /* returns ops consumed */

static int rule_step(env_t *e, int r, const double *rule_noise) {
  const moog_op *op = e->ops + r;
  int sp[MOOG_MAX_SLOTS], sq[MOOG_MAX_SLOTS];
  uint8_t flag[MOOG_MAX_SLOTS];
  switch (op->kind) {
    case MOOG_R_VANISH_ON_CONTACT: { /* vanish.py:66-86, contact_rules.py:28-35 */
      int la = op->i[0], lb = op->i[1];
      for (int i = 0; i < e->cnt[la]; ++i) {
        flag[i] = 0;
        for (int j = 0; j < e->cnt[lb]; ++j) /* every pair is evaluated (no short-circuit) */
          if (overlaps(e, LOFF(e, la) + i, LOFF(e, lb) + j)) flag[i] = 1;
      }
      vanish(e, la, flag);
      return 1;
    }
    case MOOG_R_VANISH_BY_FILTER: { /* vanish.py:42-63 */
      int la = op->i[0];
      for (int i = 0; i < e->cnt[la]; ++i)
        flag[i] = eval_expr(e, op->i[2], LOFF(e, la) + i, LOFF(e, la) + i) != 0;
      vanish(e, la, flag);
      return 1;
    }
    case MOOG_R_MODIFY_ON_CONTACT: { /* contact_rules.py:86-120 */
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      int m = gather_layers(e, op->i[2], op->i[3], sq);
      const int32_t *q = e->ipool + op->i[4]; /* mod0 filt0 mod1 filt1 */
      for (int pass = 0; pass < 2; ++pass) {
        const int *A = pass ? sq : sp, *B = pass ? sp : sq;
        int na = pass ? m : n, nb = pass ? n : m;
        int mod = q[2 * pass], filt = q[2 * pass + 1];
        if (mod < 0) continue;
        for (int i = 0; i < na; ++i) {
          if (eval_expr(e, filt, A[i], A[i]) == 0) continue;
          int any = 0;
          for (int j = 0; j < nb; ++j)
            if (B[j] != A[i] && overlaps(e, A[i], B[j])) any = 1; /* list comprehension: all evaluated */
          if (any) eval_expr(e, mod, A[i], A[i]);
        }
      }
      return 1;
    }
    case MOOG_R_MODIFY_SPRITES: { /* modify_sprites.py:35-52 */
      int n = gather_layers(e, op->i[0], op->i[1], sp);
      int m = 0;
      for (int i = 0; i < n; ++i)
        if (op->i[3] < 0 || eval_expr(e, op->i[3], sp[i], sp[i]) != 0) sq[m++] = sp[i];
      if (m == 0) return 1;
      if (op->flags & MOOG_FL_SAMPLE_ONE) {
        double u = rule_noise ? rule_noise[op->i[4]] : 0.0;
        int k = (int)(u * m);
        if (k >= m) k = m - 1;
        eval_expr(e, op->i[2], sq[k], sq[k]);
      } else {
        for (int i = 0; i < m; ++i) eval_expr(e, op->i[2], sq[i], sq[i]);
      }
      return 1;
    }
    case MOOG_R_TIMED_BEGIN: { /* timing.py:50-56 */
      double *cd = e->envf + op->i[2];
      int nsub = op->i[1];
      if (cd[0] <= 0 && cd[1] > 0) {
        int q = r + 1;
        while (q < r + 1 + nsub) q += rule_step(e, q, rule_noise);
      }
      cd[0] -= 1;
      cd[1] -= 1;
      return 1 + nsub;
    }
    case MOOG_R_KEEP_NEAR_CENTER: { /* re_center.py:60-76 */
      int la = op->i[0];
      if (e->cnt[la] < 1) return 1;
      int a = LOFF(e, la);
      double d[2];
      for (int c = 0; c < 2; ++c) {
        double pos = DYN(e, c ? MOOG_D_Y : MOOG_D_X, a) - 0.5, g = op->p[c];
        d[c] = -1. * g * (double)(pos > g) + g * (double)(pos < -1. * g);
      }
      if (d[0] != 0 || d[1] != 0) {
        int n = gather_layers(e, op->i[1], op->i[2], sp);
        for (int i = 0; i < n; ++i)
          set_position(e, sp[i], DYN(e, MOOG_D_X, sp[i]) + d[0], DYN(e, MOOG_D_Y, sp[i]) + d[1]);
      }
      return 1;
    }
    case MOOG_R_PORTAL: { /* portal.py:41-76 */
      const int la = op->i[0], lp = op->i[1], np_ = e->cnt[lp];
      if (np_ % 2 != 0) {
        e->envi[MOOG_EI_ERR] |= MOOG_ERR_PORTAL_ODD;
        return 1;
      }
      for (int i = 0; i < e->cnt[la]; ++i) {
        const int s = LOFF(e, la) + i;
        const double pt[2] = {DYN(e, MOOG_D_X, s), DYN(e, MOOG_D_Y, s)};
        int first = -1;
        for (int q = 0; q < np_; ++q) { /* in_portals: every portal is asked (sprite.py:432-440) */
          const int p = LOFF(e, lp) + q;
          int in;
          if (is_symmetric_circle(e, p)) {
            in = norm1(pt[0] - DYN(e, MOOG_D_X, p), pt[1] - DYN(e, MOOG_D_Y, p)) < STAT(e, MOOG_S_MAXR, p);
          } else {
            poly_t P;
            uint8_t r;
            world_path(e, p, &P);
            orc_points_in_path(pt, 1, &P.v[0][0], P.n + 1, &r);
            in = r;
          }
          if (in && first < 0) first = q;
        }
        if (first < 0) {
          META(e, MOOG_M_FLAGS, s) &= ~MOOG_SF_TELEPORTING;
          continue;
        }
        if (META(e, MOOG_M_FLAGS, s) & MOOG_SF_TELEPORTING) continue;
        const int ex = LOFF(e, lp) + ((first % 2) ? first - 1 : first + 1);
        set_position(e, s, DYN(e, MOOG_D_X, ex), DYN(e, MOOG_D_Y, ex));
        META(e, MOOG_M_FLAGS, s) |= MOOG_SF_TELEPORTING;
      }
      return 1;
    }
    case MOOG_R_CHANGE_LAYER: { /* change_layer.py:34-45 */
      const int lo = op->i[0], ln = op->i[1], n = e->cnt[lo];
      const int cap = LOFF(e, ln + 1) - LOFF(e, ln);
      uint8_t gone[MOOG_MAX_SLOTS];
      for (int i = 0; i < n; ++i) /* should_change is evaluated for every sprite before anything moves */
        gone[i] = op->i[2] < 0 || eval_expr(e, op->i[2], LOFF(e, lo) + i, LOFF(e, lo) + i) != 0;
      for (int i = 0; i < n; ++i) {
        if (!gone[i]) continue;
        if (e->cnt[ln] >= cap) {
          e->envi[MOOG_EI_ERR] |= MOOG_ERR_LAYER_OVERFLOW;
          gone[i] = 0;
          continue;
        }
        copy_slot(e, LOFF(e, ln) + e->cnt[ln], LOFF(e, lo) + i);
        e->cnt[ln] += 1;
      }
      vanish(e, lo, gone);
      return 1;
    }
    case MOOG_R_FIXATION: { /* fixation.py:45-54 */
      double *count = e->envf + op->i[2];
      if (e->cnt[op->i[0]] < 1 || e->cnt[op->i[1]] < 1) { /* state[layer][0]: IndexError */
        e->envi[MOOG_EI_ERR] |= MOOG_ERR_BAD_INDEX;
        return 1;
      }
      const int a = LOFF(e, op->i[0]), t = LOFF(e, op->i[1]);
      const double dist = norm1(DYN(e, MOOG_D_X, a) - DYN(e, MOOG_D_X, t), DYN(e, MOOG_D_Y, a) - DYN(e, MOOG_D_Y, t));
      *count = dist < op->p[0] ? *count + 1 : 0;
      return 1;
    }
    case MOOG_R_PHASESEQ_BEGIN: /* task_phases.py:126-141: the phase that is current when the pass begins is stepped */
      e->envf[op->i[0] + 1] = e->envf[op->i[0]];
      return 1;
    case MOOG_R_PHASE_BEGIN: { /* task_phases.py:79-95 */
      const int nsub = op->i[1];
      const double *ph = e->envf + op->i[3];
      const int active = ph[0] == 0 && (op->i[0] < 0 || e->envf[op->i[0] + 1] == (double)op->i[2]);
      if (active) {
        int q = r + 1;
        while (q < r + 1 + nsub) q += rule_step(e, q, rule_noise);
      }
      return 1 + nsub;
    }
    case MOOG_R_PHASE_END: { /* task_phases.py:90-95, 133-141 */
      double *ph = e->envf + op->i[1];
      ph[1] += 1;
      if (ph[1] >= ph[2] || (op->i[0] >= 0 && eval_condition(e, op->i[0]) != 0)) {
        ph[0] = 1;
        if (op->i[2] >= 0) {
          double *seq = e->envf + op->i[2];
          seq[0] += 1;
          const int ind = (int)seq[0];
          if (ind >= op->i[5])
            e->envi[MOOG_EI_ERR] |= MOOG_ERR_BAD_INDEX; /* self._phases[ind]: IndexError */
          else if (op->i[3] >= 0)
            e->envf[op->i[3]] = e->dpool[op->i[4] + ind];
        }
      }
      return 1;
    }
    case MOOG_R_TREE: /* a user-defined rule's step(), path by path */
      walk_tree(e, e->ipool + op->i[0], op->i[1]);
      return 1;
    case MOOG_R_CREATE_SPRITES: { /* create_sprites.py:27-34 */
      const int layer = op->i[0], have = e->cnt[layer], cap = LOFF(e, layer + 1) - LOFF(e, layer);
      int count = op->i[1];
      if (op->p[2] > op->p[1]) { /* num_sprites = lambda: np.random.randint(p1, p2): drawn per call */
        const int lo = (int)op->p[1], hi = (int)op->p[2];
        const int c = lo + (int)(rule_noise_at(e, (int)op->p[3]) * (double)(hi - lo));
        count = c >= hi ? hi - 1 : (c < 0 ? 0 : c);
      }
      const int serial = e->envi[MOOG_EI_CREATED]++;
      if (have + count > cap) { /* the reference's lists grow without bound; a layer of the record does not */
        e->envi[MOOG_EI_ERR] |= MOOG_ERR_LAYER_OVERFLOW;
        count = cap - have;
      }
      generate_sprites(e, op, g_seed ^ 0x3C6EF372FE94F82Bull, (uint32_t)serial, layer, LOFF(e, layer) + have, count,
                       e->ipool + op->i[2], op->i[3], 1);
      return 1;
    }
    case MOOG_R_COND_BEGIN: { /* conditional.py:55-58 */
      int times = (int)eval_condition(e, op->i[0]);
      int nsub = op->i[1];
      for (int t = 0; t < times; ++t) {
        int q = r + 1;
        while (q < r + 1 + nsub) q += rule_step(e, q, rule_noise);
      }
      return 1 + nsub;
    }
  }
  return 1;
}

/* environment.py:86: meta_state = meta_state_initializer() -- its entries are variables of the record */
static void meta_reset(env_t *e) {
  const int32_t *h = e->hdr;
  for (int q = 0; q < h[MOOG_H_N_METAVAR]; ++q) e->envf[h[MOOG_H_METAVAR_OFF] + q] = e->dpool[h[MOOG_H_METAVAR_INIT] + q];
}

/* AbstractRule.reset of the rules that keep state: TimedRule re-arms its interval (timing.py:45-48) */
static void rules_reset(env_t *e) {
  const int32_t *h = e->hdr;
  for (int r = h[MOOG_H_RULES]; r < h[MOOG_H_RULES] + h[MOOG_H_N_RULES]; ++r) {
    const moog_op *op = e->ops + r;
    if (op->kind == MOOG_R_TIMED_BEGIN) {
      e->envf[op->i[2]] = op->p[0];
      e->envf[op->i[2] + 1] = op->p[1];
    }
    if (op->kind == MOOG_R_PORTAL) /* portal.py:36-39: _currently_teleporting = set() */
      for (int s2 = 0; s2 < e->S; ++s2) META(e, MOOG_M_FLAGS, s2) &= ~MOOG_SF_TELEPORTING;
    if (op->kind == MOOG_R_FIXATION) e->envf[op->i[2]] = 0; /* fixation.py:41-43 */
    if (op->kind == MOOG_R_PHASESEQ_BEGIN) { /* task_phases.py:118-124 */
      e->envf[op->i[0]] = 0;
      e->envf[op->i[0] + 1] = 0;
      if (op->i[2] >= 0) e->envf[op->i[2]] = e->dpool[op->i[4]];
    }
    if (op->kind == MOOG_R_PHASE_BEGIN) { /* task_phases.py:71-77: the duration is drawn anew */
      double *ph = e->envf + op->i[3];
      ph[0] = 0;
      ph[1] = 0;
      ph[2] = op->p[0];
      if (op->p[2] > op->p[1]) { /* np.random.randint(p1, p2) */
        const int lo = (int)op->p[1], hi = (int)op->p[2];
        int d = lo + (int)(rule_noise_at(e, op->i[4]) * (double)(hi - lo));
        ph[2] = d >= hi ? hi - 1 : d;
      }
    }
    if (op->kind == MOOG_R_TREE) /* the rule's own reset(): its attributes back to their first values */
      for (int q = 0; q < op->i[3]; ++q) e->envf[op->i[2] + q] = e->dpool[op->i[4] + q];
  }
}

static void rules_step(env_t *e, const double *rule_noise) {
  const int32_t *h = e->hdr;
  int r = h[MOOG_H_RULES], end = h[MOOG_H_RULES] + h[MOOG_H_N_RULES];
  e->envi[MOOG_EI_RULE_PASSES] += 1;
  e->rule_noise = rule_noise;
  while (r < end) r += rule_step(e, r, rule_noise);
}

/* ------------------------------------------------------------------------ */
/* action spaces                                                             */
/* ------------------------------------------------------------------------ */

static void actions_step(env_t *e, const double *action) {
  const int32_t *h = e->hdr;
  int sp[MOOG_MAX_SLOTS];
  for (int a = 0; a < h[MOOG_H_N_ACTIONS]; ++a) {
    const moog_op *op = e->ops + h[MOOG_H_ACTIONS] + a;
    const double *act = action + op->i[2];
    int n = gather_layers(e, op->i[0], op->i[1], sp);
    if (op->kind == MOOG_A_SET_POSITION) { /* set_position.py:34-47 */
      for (int i = 0; i < n; ++i)
        set_position(e, sp[i], op->p[0] * DYN(e, MOOG_D_X, sp[i]) + (1 - op->p[0]) * act[0],
                     op->p[0] * DYN(e, MOOG_D_Y, sp[i]) + (1 - op->p[0]) * act[1]);
      continue;
    }
    double *mem = e->envf + op->i[5];
    double ax, ay;
    if (op->kind == MOOG_A_JOYSTICK) { /* joystick.py:45-65 */
      ax = act[0];
      ay = (op->flags & MOOG_FL_CONSTRAINED_LR) ? 0. : act[1];
      mem[0] = mem[0] * op->p[1] + op->p[0] * ax;
      mem[1] = mem[1] * op->p[1] + op->p[0] * ay;
    } else { /* grid.py:52-70: the UNIT action is added, then clipped */
      static const double G[5][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {0, 0}};
      int k = (int)act[0];
      if (k < 0 || k > 4) k = 4;
      mem[0] = mem[0] * op->p[1] + G[k][0];
      mem[1] = mem[1] * op->p[1] + G[k][1];
    }
    mem[0] = fmin(fmax(mem[0], -op->p[0]), op->p[0]);
    mem[1] = fmin(fmax(mem[1], -op->p[0]), op->p[0]);
    for (int i = 0; i < n; ++i) {
      double m = STAT(e, MOOG_S_MASS, sp[i]);
      if (op->flags & MOOG_FL_CONTROL_VELOCITY)
        assign_velocity(e, sp[i], mem[0] / m, mem[1] / m);
      else
        add_velocity(e, sp[i], mem[0] / m, mem[1] / m);
    }
  }
}

/* ------------------------------------------------------------------------ */
/* tasks                                                                     */
/* ------------------------------------------------------------------------ */

static void tasks_reset(env_t *e) {
  const int32_t *h = e->hdr;
  for (int t = 0; t < h[MOOG_H_N_TASKS]; ++t) {
    const moog_op *op = e->ops + h[MOOG_H_TASKS] + t;
    if (op->kind == MOOG_T_CONTACT_REWARD || op->kind == MOOG_T_RESET)
      e->envf[op->i[5]] = INFINITY; /* contact_reward.py:67-68, reset.py:45-46 */
  }
}

static void actions_reset(env_t *e) {
  const int32_t *h = e->hdr;
  for (int a = 0; a < h[MOOG_H_N_ACTIONS]; ++a) {
    const moog_op *op = e->ops + h[MOOG_H_ACTIONS] + a;
    if (op->kind == MOOG_A_JOYSTICK || op->kind == MOOG_A_GRID) {
      e->envf[op->i[5]] = 0; /* joystick.py:67-70, grid.py:72-75 */
      e->envf[op->i[5] + 1] = 0;
    }
  }
}

/* composite_task.py:32-42 flattened over the task tree */
static void tasks_reward(env_t *e, int step_count, double *reward, int *should_reset) {
  const int32_t *h = e->hdr;
  int sp[MOOG_MAX_SLOTS], sq[MOOG_MAX_SLOTS];
  double total = 0;
  int reset = 0;
  for (int t = 0; t < h[MOOG_H_N_TASKS]; ++t) {
    const moog_op *op = e->ops + h[MOOG_H_TASKS] + t;
    switch (op->kind) {
      case MOOG_T_TIMEOUT:
        if (step_count >= op->p[0]) reset = 1;
        break;
      case MOOG_T_STAY_ALIVE: /* stay_alive.py:22-32 */
        if (((step_count + 1) % (int)op->p[0]) == 0) total += op->p[1];
        break;
      case MOOG_T_CONTACT_REWARD: { /* contact_reward.py:70-102 */
        double r = 0;
        double *cd = e->envf + op->i[5];
        int n = gather_layers(e, op->i[0], op->i[1], sp);
        int m = gather_layers(e, op->i[2], op->i[3], sq);
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < m; ++j) {
            if (eval_expr(e, op->i[4], sp[i], sq[j]) == 0) continue;
            if (overlaps(e, sp[i], sq[j])) {
              r = op->p[2] > 0 ? eval_expr(e, (int)op->p[2] - 1, sp[i], sq[j]) : op->p[0];
              if (*cd == INFINITY) *cd = op->p[1];
            }
          }
        *cd -= 1;
        if (*cd < 0) reset = 1;
        total += r;
        break;
      }
      case MOOG_T_RESET: { /* reset.py:48-61 */
        double r = 0.;
        double *cd = e->envf + op->i[5];
        if (*cd == INFINITY && eval_condition(e, op->i[0]) != 0) {
          r = op->i[1] > 0 ? eval_condition(e, op->i[1] - 1) : op->p[1]; /* reward_fn(state), only now */
          *cd = op->p[0];
        }
        *cd -= 1;
        if (*cd < 0) reset = 1;
        total += r;
        break;
      }
    }
  }
  *reward = total;
  *should_reset = reset;
}

/* ------------------------------------------------------------------------ */
/* entry points                                                              */
/* ------------------------------------------------------------------------ */

typedef struct {
  double *dyn, *stat;
  int32_t *meta, *cnt, *envi;
  double *envf;
  double *vtx;
} orc_state;

static void bind_env(env_t *e, const void *blob, const orc_state *st, int n) {
  memset(e, 0, sizeof(*e));
  bind_program(e, blob);
  e->env_id = n;
  int S = e->S;
  e->dyn = st->dyn + (size_t)n * MOOG_DYN_FIELDS * S;
  e->stat = st->stat + (size_t)n * MOOG_STAT_FIELDS * S;
  e->meta = st->meta + (size_t)n * MOOG_META_FIELDS * S;
  e->cnt = st->cnt + (size_t)n * MOOG_MAX_LAYERS;
  e->envi = st->envi + (size_t)n * MOOG_ENVI_WORDS;
  e->envf = st->envf + (size_t)n * e->hdr[MOOG_H_N_ENVF];
  e->vtx = st->vtx + (size_t)n * 2 * e->hdr[MOOG_H_N_VTX];
}

/* environment.py:82-96 the part of reset() after the state initializer ran:
 * task/action reset, then every rule is reset and stepped once. */
void orc_env_post_reset(const void *blob, const orc_state *st, int n_envs, const double *rule_noise) {
  for (int n = 0; n < n_envs; ++n) {
    env_t e;
    bind_env(&e, blob, st, n);
    e.envi[MOOG_EI_STEP_COUNT] = 0;
    e.envi[MOOG_EI_RESET_NEXT] = 0;
    meta_reset(&e);
    tasks_reset(&e);
    actions_reset(&e);
    e.rule_noise = rule_noise;
    rules_reset(&e);
    int nrn = 0; /* rule noise columns */
    (void)nrn;
    rules_step(&e, rule_noise);
  }
}

/* abstract_physics.py:39-42 Physics.step only (K substeps) */
void orc_physics_step(const void *blob, const orc_state *st, int n_envs, const double *noise,
                      int64_t *counters) {
  for (int n = 0; n < n_envs; ++n) {
    env_t e;
    bind_env(&e, blob, st, n);
    int nd = e.hdr[MOOG_H_NOISE_DIM];
    e.noise = noise ? noise + (size_t)n * e.K * nd : NULL;
    for (int k = 0; k < e.K; ++k) {
      e.substep = k;
      apply_physics(&e);
    }
    if (counters) {
      counters[4 * n + 0] = e.n_overlap_calls;
      counters[4 * n + 1] = e.n_overlap_true;
      counters[4 * n + 2] = e.n_collisions;
      counters[4 * n + 3] = (int64_t)e.overlap_hash;
    }
  }
}

/* environment.py:102-126: one transition of an env that is not pending a reset */
static void env_step_one(env_t *e, const double *action, const double *rule_noise, double *reward,
                         int32_t *step_type) {
  rules_step(e, rule_noise);
  actions_step(e, action);
  for (int k = 0; k < e->K; ++k) {
    e->substep = k;
    apply_physics(e);
  }
  e->envi[MOOG_EI_STEP_COUNT] += 1;
  double r;
  int reset;
  tasks_reward(e, e->envi[MOOG_EI_STEP_COUNT], &r, &reset);
  *reward = r;
  *step_type = reset ? MOOG_STEP_LAST : MOOG_STEP_MID;
  e->envi[MOOG_EI_RESET_NEXT] = reset;
}

static void put_counters(const env_t *e, int64_t *counters, int n) {
  if (!counters) return;
  counters[4 * n + 0] = e->n_overlap_calls;
  counters[4 * n + 1] = e->n_overlap_true;
  counters[4 * n + 2] = e->n_collisions;
  counters[4 * n + 3] = (int64_t)e->overlap_hash;
}

/* environment.py:98-126 Environment.step for envs that are not pending a
 * reset (the host handles `_reset_next_step` by re-initialising the state and
 * calling orc_env_post_reset; orc_env_step_auto below does it itself).  action: [N][action_dim];
 * noise: [N][K][noise_dim] uniforms for RandomForce; rule_noise: [N][..] uniforms for sample_one rules. */
void orc_env_step(const void *blob, const orc_state *st, int n_envs, const double *action,
                  const double *noise, const double *rule_noise, int n_rule_noise, double *reward,
                  int32_t *step_type, int64_t *counters) {
  for (int n = 0; n < n_envs; ++n) {
    env_t e;
    bind_env(&e, blob, st, n);
    int nd = e.hdr[MOOG_H_NOISE_DIM];
    int ad = e.hdr[MOOG_H_ACTION_DIM];
    e.noise = noise ? noise + (size_t)n * e.K * nd : NULL;
    env_step_one(&e, action + (size_t)n * ad, rule_noise ? rule_noise + (size_t)n * n_rule_noise : NULL,
                 &reward[n], &step_type[n]);
    put_counters(&e, counters, n);
  }
}

/* The state initializer drawn per env instead of taken from the pool (moog_step_io.sample_resets): every
 * generate_sprites group the initializer was traced into (MOOG_Z_GENERATE) is drawn afresh into the
 * template, on the Philox stream keyed by (seed, env, episode).  sprite_generators.py:75-103; a count given
 * as `lambda: np.random.randint(lo, hi)` is drawn first. */
static int g_sample_resets = 0;
void orc_set_sample_resets(int on) { g_sample_resets = on; }

static void reset_generate(env_t *e, const moog_op *op) {
  const int first = op->i[0];
  int count = op->i[1];
  const uint64_t key = g_seed ^ 0x6A09E667F3BCC908ull;
  if (op->p[2] > op->p[1]) {
    const double u = philox_uniform(key, (uint32_t)e->env_id, (uint32_t)e->envi[MOOG_EI_EPISODES],
                                    ((uint32_t)first << 20) | 0xfffffu, (0x5Cu << 24));
    const int lo = (int)op->p[1], hi = (int)op->p[2];
    int c = lo + (int)(u * (double)(hi - lo));
    if (c >= hi) c = hi - 1;
    if (c < count) count = c < 0 ? 0 : c;
  }
  int layer = 0;
  for (int l = 0; l < e->L; ++l)
    if (first >= LOFF(e, l) && first < LOFF(e, l + 1)) layer = l;
  generate_sprites(e, op, key, (uint32_t)e->envi[MOOG_EI_EPISODES], layer, first, count, e->ipool + op->i[2],
                   op->i[3], 0);
}

/* Environment.step INCLUDING its first two lines (environment.py:100-101): an env whose previous
 * transition was a termination ignores the action and runs reset() instead (environment.py:82-96).
 * The state initializer's result is row reset_index[n] of a pool of initial states (the batched
 * environment keeps the config's own state_initializer() results in such a pool; which row an env
 * receives is the caller's draw): every array of the record but the env's integer words is
 * replaced, the episode counter advances, then task / action space / rules are reset and every
 * rule is stepped once.  Such a step returns dm_env.restart(): step_type FIRST, reward None (NaN
 * here), discount None (NaN); MID gives discount 1, LAST 0 (dm_env.transition / termination). */
void orc_env_step_auto(const void *blob, const orc_state *st, int n_envs, const orc_state *pool, int pool_size,
                       const int32_t *reset_index, const double *action, const double *noise,
                       const double *rule_noise, int n_rule_noise, double *reward, int32_t *step_type,
                       double *discount, int64_t *counters) {
  for (int n = 0; n < n_envs; ++n) {
    env_t e;
    bind_env(&e, blob, st, n);
    int nd = e.hdr[MOOG_H_NOISE_DIM];
    int ad = e.hdr[MOOG_H_ACTION_DIM];
    const double *rn = rule_noise ? rule_noise + (size_t)n * n_rule_noise : NULL;
    e.noise = noise ? noise + (size_t)n * e.K * nd : NULL;
    if (e.envi[MOOG_EI_RESET_NEXT] != 0) {
      int idx = reset_index ? reset_index[n] : 0;
      idx = idx < 0 ? 0 : (idx >= pool_size ? pool_size - 1 : idx);
      const int sample = g_sample_resets && e.hdr[MOOG_H_N_RESET] > 0;
      if (sample) idx = 0; /* pool row 0 is the template the generated sprites are drawn into */
      env_t p;
      bind_env(&p, blob, pool, idx);
      const int S = e.S, NF = e.hdr[MOOG_H_N_ENVF], VT = e.hdr[MOOG_H_N_VTX];
      memcpy(e.dyn, p.dyn, sizeof(double) * MOOG_DYN_FIELDS * S);
      memcpy(e.stat, p.stat, sizeof(double) * MOOG_STAT_FIELDS * S);
      memcpy(e.meta, p.meta, sizeof(int32_t) * MOOG_META_FIELDS * S);
      memcpy(e.cnt, p.cnt, sizeof(int32_t) * MOOG_MAX_LAYERS);
      memcpy(e.envf, p.envf, sizeof(double) * NF);
      memcpy(e.vtx, p.vtx, sizeof(double) * 2 * VT);
      if (sample)
        for (int z = 0; z < e.hdr[MOOG_H_N_RESET]; ++z) reset_generate(&e, e.ops + e.hdr[MOOG_H_RESET] + z);
      e.envi[MOOG_EI_EPISODES] += 1;
      e.envi[MOOG_EI_STEP_COUNT] = 0;
      e.envi[MOOG_EI_RESET_NEXT] = 0;
      meta_reset(&e);
      tasks_reset(&e);
      actions_reset(&e);
      e.rule_noise = rn;
      rules_reset(&e);
      rules_step(&e, rn);
      reward[n] = NAN;
      step_type[n] = MOOG_STEP_FIRST;
    } else {
      env_step_one(&e, action + (size_t)n * ad, rn, &reward[n], &step_type[n]);
    }
    if (discount)
      discount[n] = step_type[n] == MOOG_STEP_FIRST ? NAN : (step_type[n] == MOOG_STEP_LAST ? 0.0 : 1.0);
    put_counters(&e, counters, n);
  }
}

/* overlap matrix between two layers of every env: out[n][cap_a][cap_b] (uint8) */
void orc_overlap_pairs(const void *blob, const orc_state *st, int n_envs, int la, int lb,
                       uint8_t *out) {
  for (int n = 0; n < n_envs; ++n) {
    env_t e;
    bind_env(&e, blob, st, n);
    int ca = LOFF(&e, la + 1) - LOFF(&e, la), cb = LOFF(&e, lb + 1) - LOFF(&e, lb);
    uint8_t *o = out + (size_t)n * ca * cb;
    memset(o, 0, (size_t)ca * cb);
    for (int i = 0; i < e.cnt[la]; ++i)
      for (int j = 0; j < e.cnt[lb]; ++j)
        o[i * cb + j] = (uint8_t)overlaps(&e, LOFF(&e, la) + i, LOFF(&e, lb) + j);
  }
}

/* world vertices of every slot: out[n][S][MOOG_MAX_OUTLINE][2], nv[n][S] */
void orc_world_vertices(const void *blob, const orc_state *st, int n_envs, double *out, int32_t *nv) {
  for (int n = 0; n < n_envs; ++n) {
    env_t e;
    bind_env(&e, blob, st, n);
    for (int l = 0; l < e.L; ++l)
      for (int i = 0; i < e.cnt[l]; ++i) {
        int s = LOFF(&e, l) + i;
        poly_t P;
        world_path(&e, s, &P);
        nv[(size_t)n * e.S + s] = P.n;
        memcpy(out + ((size_t)n * e.S + s) * MOOG_MAX_OUTLINE * 2, &P.v[0][0], sizeof(double) * 2 * P.n);
      }
  }
}
