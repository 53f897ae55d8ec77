// moog_host_geom.cpp -- host-side polygon predicates of libmoog_b200.
//
// Used only at episode-construction time, by the host `moog.sprite.Sprite`
// (rejection sampling of non-overlapping sprites in
// moog/state_initialization/sprite_generators.py:69-105 calls
// Sprite.overlaps_sprite, moog/sprite.py:462-484).  The step itself never
// comes here: it runs in the CUDA kernels of moog_step.cu.
//
// The predicates restate matplotlib's src/_path.h (segments_intersect,
// path_intersects_path, point_in_path_impl) which the reference reaches through
// matplotlib.path.Path; compiled without FMA contraction.
#include <math.h>
#include <stdint.h>

#include "../../include/moog_b200.h"

namespace {

inline bool close_to(double a, double b) {
  double scale = fmax(fabs(a), fabs(b));
  return fabs(a - b) <= fmax(1e-10 * scale, 1e-13);
}

struct Pt { double x, y; };

bool seg_hit(Pt p1, Pt p2, Pt p3, Pt p4) {
  double den = ((p4.y - p3.y) * (p2.x - p1.x)) - ((p4.x - p3.x) * (p2.y - p1.y));
  if (close_to(den, 0.0)) {
    double t_area = (p2.x * p3.y - p3.x * p2.y) - p1.x * (p3.y - p2.y) + p1.y * (p3.x - p2.x);
    if (!close_to(t_area, 0.0)) return false;
    bool vertical = (p1.x == p2.x && p2.x == p3.x);
    double lo12 = vertical ? fmin(p1.y, p2.y) : fmin(p1.x, p2.x);
    double hi12 = vertical ? fmax(p1.y, p2.y) : fmax(p1.x, p2.x);
    double lo34 = vertical ? fmin(p3.y, p4.y) : fmin(p3.x, p4.x);
    double hi34 = vertical ? fmax(p3.y, p4.y) : fmax(p3.x, p4.x);
    return (lo12 <= lo34 && lo34 <= hi12) || (lo34 <= lo12 && lo12 <= hi34);
  }
  double u1 = (((p4.x - p3.x) * (p1.y - p3.y)) - ((p4.y - p3.y) * (p1.x - p3.x))) / den;
  double u2 = (((p2.x - p1.x) * (p1.y - p3.y)) - ((p2.y - p1.y) * (p1.x - p3.x))) / den;
  auto in01 = [](double u) { return (u > 0.0 || close_to(u, 0.0)) && (u < 1.0 || close_to(u, 1.0)); };
  return in01(u1) && in01(u2);
}

bool polylines_cross(const Pt *a, int na, const Pt *b, int nb) {
  if (na < 2 || nb < 2) return false;
  Pt a1 = a[0];
  for (int i = 1; i < na; ++i) {
    Pt a2 = a[i];
    double la = (a1.x - a2.x) * (a1.x - a2.x) + (a1.y - a2.y) * (a1.y - a2.y);
    if (close_to(la, 0.0)) continue;
    Pt b1 = b[0];
    for (int j = 1; j < nb; ++j) {
      Pt b2 = b[j];
      double lb = (b1.x - b2.x) * (b1.x - b2.x) + (b1.y - b2.y) * (b1.y - b2.y);
      if (close_to(lb, 0.0)) continue;
      if (seg_hit(a1, a2, b1, b2)) return true;
      b1 = b2;
    }
    a1 = a2;
  }
  return false;
}

bool inside(Pt t, const Pt *v, int nv) {
  if (nv < 3 || !(isfinite(t.x) && isfinite(t.y))) return false;
  bool in = false;
  for (int i = 0; i < nv; ++i) {
    Pt p = v[i], q = v[(i + 1 == nv) ? 0 : i + 1];
    bool f0 = p.y >= t.y, f1 = q.y >= t.y;
    if (f0 != f1 && ((((q.y - t.y) * (p.x - q.x)) >= ((q.x - t.x) * (p.y - q.y))) == f1)) in = !in;
  }
  return in;
}

bool all_inside(const Pt *pts, int np, const Pt *v, int nv) {
  if (nv < 3) return false;
  for (int i = 0; i < np; ++i)
    if (!inside(pts[i], v, nv)) return false;
  return true;
}

}  // namespace

extern "C" {

int moog_host_paths_overlap(const double *a, int na, const double *b, int nb) {
  const Pt *A = (const Pt *)a, *B = (const Pt *)b;
  return polylines_cross(A, na, B, nb) || all_inside(B, nb, A, na) || all_inside(A, na, B, nb);
}

void moog_host_points_in_path(const double *pts, int np, const double *path, int nv, uint8_t *out) {
  const Pt *P = (const Pt *)pts, *V = (const Pt *)path;
  for (int i = 0; i < np; ++i) out[i] = (uint8_t)inside(P[i], V, nv);
}

}  // extern "C"
