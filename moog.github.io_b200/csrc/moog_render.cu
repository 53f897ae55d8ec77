// moog_render.cu -- PILRenderer.__call__ for N envs (moog/observers/pil_renderer.py:88-120).
//
// The reference pastes a background, then for every layer and sprite (z-order,
// back to front) calls Pillow's ImageDraw.polygon with an RGBA fill on an RGB
// canvas, resizes (identity for anti_aliasing=1, which every shipped config
// uses) and flips the rows.  Pillow's fill (src/libImaging/Draw.c,
// polygon_generic + hline32rgba) is an integer-scanline algorithm: vertices are
// truncated to int, every scanline collects the float32 x of the edges it
// touches (non-fused (y-y0)*dx + x0), sorts them and blends the spans; rows never
// interact.  So: one thread per scanline, one canvas per env in shared memory
// (packed RGBX words, row stride W+1 -> conflict-free), sprites walked in
// z-order by every row thread, and the finished canvas is written to HBM once,
// flipped, as coalesced 16-byte stores.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "moog_common.cuh"

namespace moog {

#define MAXV MOOG_MAX_OUTLINE
#define MAX_XX (2 * MAXV + 8)

// C `(int)double` as the reference's host executes it (x86-64 cvttsd2si): NaN and
// out-of-range values give INT_MIN (CUDA's conversion would saturate / give 0)
__device__ __forceinline__ int c_int_cast(double v) {
  return (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (int)0x80000000;
}

// Resample.c clip8: fixed point (22 fractional bits) -> uint8
__device__ __forceinline__ unsigned clip8(int v) {
  v >>= 22;
  return (unsigned)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ unsigned div255(unsigned a) { return (((a + 128) >> 8) + (a + 128)) >> 8; }

__device__ __forceinline__ unsigned blend_px(unsigned bg, unsigned ink) {
  unsigned a = ink >> 24;
  unsigned r = div255((bg & 255) * (255 - a) + (ink & 255) * a);
  unsigned g = div255(((bg >> 8) & 255) * (255 - a) + ((ink >> 8) & 255) * a);
  unsigned b = div255(((bg >> 16) & 255) * (255 - a) + ((ink >> 16) & 255) * a);
  return r | (g << 8) | (b << 16);
}

// Draw.c hline32rgba for the row this thread owns (y is already known to be in range)
__device__ __forceinline__ void hline(unsigned *row, int W, int x0, int x1, unsigned ink) {
  if (x0 < 0) x0 = 0; else if (x0 >= W) return;
  if (x1 < 0) return; else if (x1 >= W) x1 = W - 1;
  for (int x = x0; x <= x1; ++x) row[x] = blend_px(row[x], ink);
}

__device__ __forceinline__ int round_up_(float f) { return (int)(f >= 0.0f ? floorf(f + 0.5f) : -floorf(fabsf(f) + 0.5f)); }
__device__ __forceinline__ int round_down_(float f) { return (int)(f >= 0.0f ? ceilf(f - 0.5f) : -ceilf(fabsf(f) - 0.5f)); }

// Draw.c: (y - e->y0) * e->dx + e->x0, float32, no fused multiply-add
__device__ __forceinline__ float edge_x(int x0, int y0, float dx, int y) {
  return __fadd_rn(__fmul_rn((float)(y - y0), dx), (float)x0);
}

struct PEdge { int x0, y0, ymin, ymax; float dx; };

__device__ __forceinline__ PEdge make_edge(int2 a, int2 b) {
  PEdge e;
  e.x0 = a.x; e.y0 = a.y;
  e.ymin = min(a.y, b.y); e.ymax = max(a.y, b.y);
  e.dx = (a.y == b.y) ? 0.0f : __fdiv_rn((float)(b.x - a.x), (float)(b.y - a.y));
  return e;
}

// Draw.c draw_horizontal_lines for row y: walks the edge list ImagingDrawPolygon
// would have built (consecutive collinear horizontal edges merged) in order.
__device__ inline void draw_horizontal_lines(const int2 *xy, int count, bool closing, int y, int *x_pos,
                                             unsigned *row, int W, bool row_visible, unsigned ink) {
  bool pend = false;
  int pmin = 0, pmax = 0;
#define MOOG_FLUSH_PENDING()                                        \
  if (pend) {                                                       \
    pend = false;                                                   \
    int xmin_ = pmin, xmax_ = pmax;                                 \
    bool skip_ = (*x_pos != -1 && *x_pos < xmin_);                  \
    if (!skip_ && *x_pos > xmin_) {                                 \
      xmin_ = *x_pos;                                               \
      if (xmax_ < xmin_) skip_ = true;                              \
    }                                                               \
    if (!skip_) {                                                   \
      if (row_visible && xmin_ <= xmax_) hline(row, W, xmin_, xmax_, ink); \
      *x_pos = xmax_ + 1;                                           \
    }                                                               \
  }
  for (int i = 0; i < count - 1; ++i) {
    int2 a = xy[i], b = xy[i + 1];
    if (a.y == b.y && i != 0 && a.y == xy[i - 1].y) {
      int xp = xy[i - 1].x;
      if (b.x > a.x && a.x > xp) {
        if (pend) pmax = b.x;
        continue;
      } else if (b.x < a.x && a.x < xp) {
        if (pend) pmin = b.x;
        continue;
      }
    }
    MOOG_FLUSH_PENDING();
    if (a.y == b.y && a.y == y) {
      pend = true;
      pmin = min(a.x, b.x);
      pmax = max(a.x, b.x);
    }
  }
  if (closing) {
    MOOG_FLUSH_PENDING();
    int2 a = xy[count - 1], b = xy[0];
    if (a.y == b.y && a.y == y) {
      pend = true;
      pmin = min(a.x, b.x);
      pmax = max(a.x, b.x);
    }
  }
  MOOG_FLUSH_PENDING();
#undef MOOG_FLUSH_PENDING
}

// Draw.c polygon_generic, the iteration of its scanline loop for row y.
// ymax_c = min(polygon ymax, H) as in the reference; row == nullptr-safe via row_visible.
__device__ inline void polygon_row(const int2 *xy, int count, int y, int ymax_c, bool has_horizontal,
                                   unsigned *row, int W, bool row_visible, unsigned ink) {
  float xx[MAX_XX];
  int j = 0;
  bool closing = (xy[count - 1].x != xy[0].x) || (xy[count - 1].y != xy[0].y);
  int n_edges = count - 1 + (closing ? 1 : 0);
  for (int i = 0; i < n_edges; ++i) {
    int2 a = xy[i], b = xy[(i + 1 == count) ? 0 : i + 1];
    if (a.y == b.y) continue;  // horizontal edges are deferred when blending
    PEdge cur = make_edge(a, b);
    if (y >= cur.ymin && y <= cur.ymax) {
      xx[j++] = edge_x(cur.x0, cur.y0, cur.dx, y);
      if (y == cur.ymax && y < ymax_c) {
        xx[j] = xx[j - 1];
        j++;
      } else if ((y == cur.ymin || y == cur.ymax) && cur.dx != 0) {
        for (int k = 0; k < i; ++k) {
          int2 c = xy[k], d = xy[(k + 1 == count) ? 0 : k + 1];
          if (c.y == d.y) continue;
          PEdge oth = make_edge(c, d);
          if ((y != oth.ymin && y != oth.ymax) || oth.dx == 0) continue;
          if (roundf(xx[j - 1]) == roundf(edge_x(oth.x0, oth.y0, oth.dx, y))) {
            int off = (y == ymax_c) ? -1 : 1;
            if (y + off >= oth.ymin && y + off <= oth.ymax) {
              float adj = edge_x(cur.x0, cur.y0, cur.dx, y + off);
              float oadj = edge_x(oth.x0, oth.y0, oth.dx, y + off);
              if (xx[j - 1] > adj + 1 && xx[j - 1] > oadj + 1)
                xx[j - 1] = roundf(fmaxf(adj, oadj)) + 1;
              else if (xx[j - 1] < adj - 1 && xx[j - 1] < oadj - 1)
                xx[j - 1] = roundf(fminf(adj, oadj)) - 1;
              break;
            }
          }
        }
      }
    }
  }
  // qsort ascending
  for (int a = 1; a < j; ++a) {
    float v = xx[a];
    int b = a;
    while (b > 0 && xx[b - 1] > v) {
      xx[b] = xx[b - 1];
      --b;
    }
    xx[b] = v;
  }
  int x_pos = (j == 0) ? -1 : 0;
  for (int i = 1; i < j; i += 2) {
    int x_end = round_down_(xx[i]);
    if (x_end < x_pos) continue;
    if (has_horizontal) draw_horizontal_lines(xy, count, closing, y, &x_pos, row, W, row_visible, ink);
    if (x_end < x_pos) continue;
    int x_start = round_up_(xx[i - 1]);
    if (x_pos > x_start) {
      x_start = x_pos;
      if (x_end < x_start) continue;
    }
    if (row_visible && x_start <= x_end) hline(row, W, x_start, x_end, ink);
    x_pos = x_end + 1;
  }
  if (has_horizontal) draw_horizontal_lines(xy, count, closing, y, &x_pos, row, W, row_visible, ink);
}

// ---------------------------------------------------------------------------
// Pillow's edge list, built ONCE per sprite (ImagingDrawPolygon, Draw.c): the
// consecutive collinear horizontal edges are merged exactly as Pillow merges
// them, dx is divided once.  A row then only reads records.
// ---------------------------------------------------------------------------
struct ERec {
  int x0, y0, ymin, ymax;  // horizontal edge: x0 = xmin, ymin == ymax == y
  float dx;
  int xmax_h;              // horizontal edge: xmax
};

__device__ __forceinline__ ERec make_rec(int2 a, int2 b) {
  ERec r;
  r.ymin = min(a.y, b.y);
  r.ymax = max(a.y, b.y);
  if (a.y == b.y) {
    r.x0 = min(a.x, b.x);
    r.y0 = a.y;
    r.dx = 0.0f;
    r.xmax_h = max(a.x, b.x);
  } else {
    r.x0 = a.x;
    r.y0 = a.y;
    r.dx = __fdiv_rn((float)(b.x - a.x), (float)(b.y - a.y));
    r.xmax_h = 0;
  }
  return r;
}

// returns the number of records written to E (<= count)
__device__ inline int build_edge_list(const int2 *xy, int count, ERec *E, int *has_horizontal) {
  int ne = 0, hz = 0;
  for (int i = 0; i < count - 1; ++i) {
    int2 a = xy[i], b = xy[i + 1];
    if (a.y == b.y) {
      hz = 1;
      if (i != 0 && a.y == xy[i - 1].y) {
        int xp = xy[i - 1].x;
        if (b.x > a.x && a.x > xp) {
          E[ne - 1].xmax_h = b.x;
          continue;
        } else if (b.x < a.x && a.x < xp) {
          E[ne - 1].x0 = b.x;
          continue;
        }
      }
    }
    E[ne++] = make_rec(a, b);
  }
  if (count > 0 && (xy[count - 1].x != xy[0].x || xy[count - 1].y != xy[0].y)) {
    if (xy[count - 1].y == xy[0].y) hz = 1;
    E[ne++] = make_rec(xy[count - 1], xy[0]);
  }
  *has_horizontal = hz;
  return ne;
}

// Where polygon_generic's hline calls go.  BlendSink blends straight into the row
// (hline32rgba); SpanSink records the clipped spans of one (sprite, row) item so
// that the rows can be filled later, in z-order, by another thread.
struct BlendSink {
  unsigned *row;
  int W;
  unsigned ink;
  int xlo, xhi;  // the columns this thread owns
  __device__ __forceinline__ void hline(int x0, int x1) {
    if (x0 < 0) x0 = 0; else if (x0 >= W) return;   // hline32rgba's clipping
    if (x1 < 0) return; else if (x1 >= W) x1 = W - 1;
    x0 = max(x0, xlo);
    x1 = min(x1, xhi);
    for (int x = x0; x <= x1; ++x) row[x] = blend_px(row[x], ink);
  }
};

#define ITEM_SPANS 3          /* spans stored per item; more -> the item is redone directly */
#define ITEM_OVERFLOW 0xffu
struct SpanSink {
  unsigned *item;  // [1 + ITEM_SPANS] words: count, then x0 | x1 << 16
  int W;
  int n;
  __device__ __forceinline__ void hline(int x0, int x1) {
    if (x0 < 0) x0 = 0; else if (x0 >= W) return;   // hline32rgba's clipping
    if (x1 < 0) return; else if (x1 >= W) x1 = W - 1;
    if (x0 > x1) return;
    if (n < ITEM_SPANS) item[1 + n] = (unsigned)x0 | ((unsigned)x1 << 16);
    n++;
  }
};

#define MAX_HROW 8 /* horizontal edges of one polygon on one scanline kept in registers */

// Draw.c draw_horizontal_lines for row y over the horizontal records of that row
// (hrow[0..nh): indices into E, in list order; nh < 0: scan the whole list)
template <class Sink>
__device__ inline void draw_horizontal_lines_rec(const ERec *E, int ne, const unsigned char *hrow, int nh, int y,
                                                 int *x_pos, Sink &sink) {
  const int n = nh >= 0 ? nh : ne;
  for (int q = 0; q < n; ++q) {
    const int i = nh >= 0 ? hrow[q] : q;
    const ERec e = E[i];
    if (e.ymin != e.ymax || e.ymin != y) continue;
    int xmin = e.x0;
    if (*x_pos != -1 && *x_pos < xmin) continue;
    int xmax = e.xmax_h;
    if (*x_pos > xmin) {
      xmin = *x_pos;
      if (xmax < xmin) continue;
    }
    if (xmin <= xmax) sink.hline(xmin, xmax);
    *x_pos = xmax + 1;
  }
}

// Draw.c polygon_generic, the iteration of its scanline loop for row y, on the
// prebuilt edge list.
template <class Sink>
__device__ inline void polygon_row_rec(const ERec *E, int ne, int y, int ymax_c, bool has_horizontal, Sink &sink) {
  float xx[MAX_XX];
  unsigned char hrow[MAX_HROW];
  int j = 0, nh = 0;
  for (int i = 0; i < ne; ++i) {
    const ERec cur = E[i];
    if (cur.ymin == cur.ymax) {  // horizontal edges are deferred when blending
      if (cur.ymin == y) {
        if (nh >= 0 && nh < MAX_HROW) hrow[nh++] = (unsigned char)i; else nh = -1;
      }
      continue;
    }
    if (y >= cur.ymin && y <= cur.ymax) {
      xx[j++] = edge_x(cur.x0, cur.y0, cur.dx, y);
      if (y == cur.ymax && y < ymax_c) {
        xx[j] = xx[j - 1];
        j++;
      } else if ((y == cur.ymin || y == cur.ymax) && cur.dx != 0) {
        for (int k = 0; k < i; ++k) {
          const ERec oth = E[k];
          if (oth.ymin == oth.ymax) continue;
          if ((y != oth.ymin && y != oth.ymax) || oth.dx == 0) continue;
          if (roundf(xx[j - 1]) == roundf(edge_x(oth.x0, oth.y0, oth.dx, y))) {
            int off = (y == ymax_c) ? -1 : 1;
            if (y + off >= oth.ymin && y + off <= oth.ymax) {
              float adj = edge_x(cur.x0, cur.y0, cur.dx, y + off);
              float oadj = edge_x(oth.x0, oth.y0, oth.dx, y + off);
              if (xx[j - 1] > adj + 1 && xx[j - 1] > oadj + 1)
                xx[j - 1] = roundf(fmaxf(adj, oadj)) + 1;
              else if (xx[j - 1] < adj - 1 && xx[j - 1] < oadj - 1)
                xx[j - 1] = roundf(fminf(adj, oadj)) - 1;
              break;
            }
          }
        }
      }
    }
  }
  has_horizontal = has_horizontal && nh != 0;
  // qsort ascending
  for (int a = 1; a < j; ++a) {
    float v = xx[a];
    int b = a;
    while (b > 0 && xx[b - 1] > v) {
      xx[b] = xx[b - 1];
      --b;
    }
    xx[b] = v;
  }
  int x_pos = (j == 0) ? -1 : 0;
  for (int i = 1; i < j; i += 2) {
    int x_end = round_down_(xx[i]);
    if (x_end < x_pos) continue;
    if (has_horizontal) draw_horizontal_lines_rec(E, ne, hrow, nh, y, &x_pos, sink);
    if (x_end < x_pos) continue;
    int x_start = round_up_(xx[i - 1]);
    if (x_pos > x_start) {
      x_start = x_pos;
      if (x_end < x_start) continue;
    }
    if (x_start <= x_end) sink.hline(x_start, x_end);
    x_pos = x_end + 1;
  }
  if (has_horizontal) draw_horizontal_lines_rec(E, ne, hrow, nh, y, &x_pos, sink);
}

// color_maps.py:21-23 (CPython colorsys.hsv_to_rgb, x255, astype(uint8))
__device__ __forceinline__ unsigned to_u8(double v) { return (unsigned)(unsigned char)(long long)v; }

__device__ inline unsigned color_to_ink(int cmap, double c0, double c1, double c2, double opacity) {
  unsigned r8, g8, b8;
  if (cmap != MOOG_CMAP_HSV) {
    r8 = to_u8(c0); g8 = to_u8(c1); b8 = to_u8(c2);
  } else {
    double h = c0, s = c1, v = c2, r, g, b;
    if (s == 0.0) {
      r = g = b = v;
    } else {
      int i = (int)(h * 6.0);
      double f = (h * 6.0) - i;
      double p = v * (1.0 - s);
      double q = v * (1.0 - s * f);
      double t = v * (1.0 - s * (1.0 - f));
      i = ((i % 6) + 6) % 6;
      switch (i) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
      }
    }
    r8 = to_u8(255 * r); g8 = to_u8(255 * g); b8 = to_u8(255 * b);
  }
  return r8 | (g8 << 8) | (b8 << 16) | (to_u8(opacity) << 24);
}

/* (sprite, row) items whose spans are precomputed, per env: 512, more for scenes with many
   sprites (a pacman maze has ~180), the rest is scan-converted by the row threads directly */
__host__ __device__ inline int item_cap(int S) { return S * 8 < 512 ? 512 : (S * 8 > 2048 ? 2048 : S * 8); }
struct RenderLayout { int canvas, ivtx, erec, items, ink, ymin, ymax, horiz, nedge, ibase, tmp, cap, total; };

// H, W: canvas size (anti_aliasing x image size); OW: image width
__host__ __device__ inline RenderLayout render_layout(int H, int W, int S, int VT, int OW) {
  RenderLayout L;
  int o = 0;
  L.canvas = o; o += 4 * H * (W + 1);
  o = (o + 7) & ~7;
  L.erec = o;   o += 24 * VT;
  // the int vertices are dead once the edge lists exist: the item spans reuse their space
  L.ivtx = o;
  L.items = o;
  {
    L.cap = item_cap(S);
    int a = 8 * VT, b = 4 * (1 + ITEM_SPANS) * L.cap;
    o += a > b ? a : b;
  }
  L.ink = o;    o += 4 * S;
  L.ymin = o;   o += 4 * S;
  L.ymax = o;   o += 4 * S;
  L.horiz = o;  o += 4 * S;
  L.nedge = o;  o += 4 * S;
  L.ibase = o;  o += 4 * (S + 1);
  // anti_aliasing > 1: the horizontally resampled image [H][OW] reuses everything
  // after the canvas (dead by then); the final image reuses the canvas
  L.tmp = L.erec;
  if (OW != W) {
    int end = L.tmp + 4 * H * OW;
    if (end > o) o = end;
  }
  L.total = (o + 15) & ~15;
  return L;
}

// blockDim.x = envs_per_block * T, T = threads of one env (>= H, multiple of 32)
__global__ void moog_render_kernel(RenderArgs a, int T, int envs_per_block, int P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ProgramView pv = view_of(a.blob);
  const int32_t *hdr = pv.hdr;
  const int S = hdr[MOOG_H_N_SLOTS], L = hdr[MOOG_H_N_LAYERS], VT = hdr[MOOG_H_N_VTX];
  // pil_renderer.py:65-66: the canvas is anti_aliasing x the image size
  const int OH = hdr[MOOG_H_R_HEIGHT], OW = hdr[MOOG_H_R_WIDTH], aa = hdr[MOOG_H_R_AA];
  const int H = aa * OH, W = aa * OW;
  const unsigned bgc = (unsigned)hdr[MOOG_H_R_BG] & 0xffffffu;
  const int cmap = hdr[MOOG_H_R_COLORMAP], pmod = hdr[MOOG_H_R_MODIFIER], pml = hdr[MOOG_H_R_MOD_LAYER];
  const int g = threadIdx.x / T, t = threadIdx.x - g * T;
  const int n = blockIdx.x * envs_per_block + g;
  const bool live = n < a.n_envs;
  // polygon_modifiers.py:88-96 TorusGeometry: every sprite is drawn as 9 copies shifted by
  // (i, j), i outer / j inner over (-1, 0, 1); copy c of slot s is the virtual slot s * C + c
  // and its vertices live at c * VT + voff[s]
  const int C = pmod == MOOG_PMOD_TORUS ? 9 : 1;
  RenderLayout lay = render_layout(H, W, S * C, (VT > 0 ? VT : 1) * C, OW);
  unsigned char *base = smem_raw + (size_t)g * lay.total;
  unsigned *canvas = (unsigned *)(base + lay.canvas);
  int2 *ivtx = (int2 *)(base + lay.ivtx);
  unsigned *ink = (unsigned *)(base + lay.ink);
  int *symin = (int *)(base + lay.ymin), *symax = (int *)(base + lay.ymax), *shoriz = (int *)(base + lay.horiz);
  int *snedge = (int *)(base + lay.nedge), *sibase = (int *)(base + lay.ibase);
  ERec *erec = (ERec *)(base + lay.erec);
  unsigned *items = (unsigned *)(base + lay.items);
  const int stride = W + 1;

  if (live) {
    const double *dyn = a.st.dyn + (size_t)n * MOOG_DYN_FIELDS * S;
    const double *stat = a.st.stat + (size_t)n * MOOG_STAT_FIELDS * S;
    const int32_t *meta = a.st.meta + (size_t)n * MOOG_META_FIELDS * S;
    const double2 *vtx = (const double2 *)(a.st.vtx + (size_t)n * 2 * VT);
    for (int i = t; i < H * stride; i += T) canvas[i] = bgc;
    double ox = 0, oy = 0;
    if (pmod == MOOG_PMOD_FIRST_PERSON) {  // polygon_modifiers.py:54-63
      int s = hdr[MOOG_H_LAYER_OFF + pml];
      ox = 0.5 - dyn[MOOG_D_X * S + s];
      oy = 0.5 - dyn[MOOG_D_Y * S + s];
    }
    // int-truncated canvas vertices (C cast toward zero), per-slot extents and ink
    for (int v = t; v < C * VT; v += T) {
      const int c = v / VT;
      double2 p = vtx[v - c * VT];
      double x = p.x, y = p.y;
      if (pmod == MOOG_PMOD_TORUS) { x = x + (double)(c / 3 - 1); y = y + (double)(c % 3 - 1); }
      else if (pmod != MOOG_PMOD_NONE) { x = x + ox; y = y + oy; }
      ivtx[v] = make_int2(c_int_cast((double)W * x), c_int_cast((double)H * y));
    }
    for (int s = t; s < S; s += T)
      ink[s] = color_to_ink(cmap, stat[MOOG_S_C0 * S + s], stat[MOOG_S_C1 * S + s], stat[MOOG_S_C2 * S + s],
                            stat[MOOG_S_OPACITY * S + s]);
    (void)meta;
  }
  __syncthreads();
  if (live) {
    const int32_t *meta = a.st.meta + (size_t)n * MOOG_META_FIELDS * S;
    for (int vs = t; vs < S * C; vs += T) {
      const int s = vs / C, vo = (vs - s * C) * VT + pv.voff[s];
      int nv = meta[MOOG_M_NV * S + s];
      const int2 *xy = ivtx + vo;
      int lo = 0x7fffffff, hi = -0x7fffffff, hz = 0;
      for (int i = 0; i < nv; ++i) {
        lo = min(lo, xy[i].y);
        hi = max(hi, xy[i].y);
      }
      snedge[vs] = nv > 0 ? build_edge_list(xy, nv, erec + vo, &hz) : 0;
      symin[vs] = lo; symax[vs] = hi; shoriz[vs] = hz;
    }
  }
  // z-order list of the live sprites and the (sprite, row) item ranges
  if (live && t == 0) {
    const int32_t *meta = a.st.meta + (size_t)n * MOOG_META_FIELDS * S;
    const int32_t *cnt = a.st.cnt + (size_t)n * MOOG_MAX_LAYERS;
    int acc = 0;
    for (int s = 0; s < S * C; ++s) sibase[s] = -1;
    for (int l = 0; l < L; ++l) {
      int c = cnt[l];
      for (int k = 0; k < c; ++k) {
        int s0 = hdr[MOOG_H_LAYER_OFF + l] + k;
        if (meta[MOOG_M_NV * S + s0] <= 0) continue;
        for (int s = s0 * C; s < (s0 + 1) * C; ++s) {
          // Draw.c polygon_generic: ymin = max(ymin, 0); ymax = min(ymax, H); rows >= H are clipped by hline
          int lo = max(symin[s], 0), hi = min(symax[s], H - 1);
          int rows = hi >= lo ? hi - lo + 1 : 0;
          if (rows > 0 && acc + rows <= lay.cap) {
            sibase[s] = acc;
            acc += rows;
          }
        }
      }
    }
    sibase[S * C] = acc;
  }
  __syncthreads();
  // phase 1: one (sprite, row) item per thread pass -> clipped spans
  if (live) {
    const int n_items = sibase[S * C];
    int s = 0;  // virtual slot
    for (int it = t; it < n_items; it += T) {
      // items are sprite-major; find the sprite that owns item `it`
      for (;;) {
        int b0 = sibase[s];
        if (b0 >= 0) {
          int lo = max(symin[s], 0), hi = min(symax[s], H - 1);
          if (it < b0 + (hi - lo + 1)) break;
        }
        ++s;
      }
      const int lo = max(symin[s], 0);
      const int y = lo + (it - sibase[s]);
      SpanSink sink;
      sink.item = items + (size_t)it * (1 + ITEM_SPANS);
      sink.W = W;
      sink.n = 0;
      polygon_row_rec(erec + (s % C) * VT + pv.voff[s / C], snedge[s], y, min(symax[s], H), shoriz[s] != 0, sink);
      sink.item[0] = sink.n <= ITEM_SPANS ? (unsigned)sink.n : ITEM_OVERFLOW;
    }
  }
  __syncthreads();
  // phase 2: one row per thread, sprites in z-order
  if (live && t < H * P) {
    const int32_t *meta = a.st.meta + (size_t)n * MOOG_META_FIELDS * S;
    const int32_t *cnt = a.st.cnt + (size_t)n * MOOG_MAX_LAYERS;
    const int part = t / H, y = t - part * H;
    const int xlo = (W * part) / P, xhi = (W * (part + 1)) / P - 1;  // P threads share a row
    unsigned *row = canvas + y * stride;
    for (int l = 0; l < L; ++l) {
      int c = cnt[l];
      for (int k = 0; k < c; ++k) {
        const int s0 = hdr[MOOG_H_LAYER_OFF + l] + k;
        if (meta[MOOG_M_NV * S + s0] <= 0) continue;
        const unsigned color = ink[s0];
        for (int s = s0 * C; s < (s0 + 1) * C; ++s) {
        int ymin_c = max(symin[s], 0), ymax_c = min(symax[s], H);
        if (y < ymin_c || y > ymax_c) continue;
        const int b0 = sibase[s];
        unsigned cntw = ITEM_OVERFLOW;
        const unsigned *item = nullptr;
        if (b0 >= 0) {
          item = items + (size_t)(b0 + (y - ymin_c)) * (1 + ITEM_SPANS);
          cntw = item[0];
        }
        if (cntw != ITEM_OVERFLOW) {
          for (unsigned q = 0; q < cntw; ++q) {
            unsigned sp = item[1 + q];
            int x0 = max((int)(sp & 0xffffu), xlo), x1 = min((int)(sp >> 16), xhi);
            if ((color >> 24) == 255u) {  // DIV255(fg * 255) == fg: opaque ink overwrites
              for (int x = x0; x <= x1; ++x) row[x] = color & 0xffffffu;
            } else {
              for (int x = x0; x <= x1; ++x) row[x] = blend_px(row[x], color);
            }
          }
        } else {
          BlendSink sink;
          sink.row = row;
          sink.W = W;
          sink.ink = color;
          sink.xlo = xlo;
          sink.xhi = xhi;
          polygon_row_rec(erec + (s - s0 * C) * VT + pv.voff[s0], snedge[s], y, ymax_c, shoriz[s] != 0, sink);
        }
        }
      }
    }
  }
  __syncthreads();
  const unsigned *img = canvas;  // the image to write out: [OH][img_stride] RGBX words
  int img_stride = stride;
  if (aa > 1) {
    // Image.resize(LANCZOS) (Resample.c, 8 bpc): horizontal pass into an 8-bit
    // intermediate, then vertical pass; fixed-point coefficients from the host
    const int *bh = a.resample, *kh = bh + 2 * OW;
    const int *bv = kh + OW * a.ksize_h, *kv = bv + 2 * OH;
    unsigned *tmp = (unsigned *)(base + lay.tmp);
    if (live) {
      for (int i = t; i < H * OW; i += T) {
        const int yy = i / OW, xx = i - yy * OW;
        const int xmin = bh[2 * xx], xmax = bh[2 * xx + 1];
        const int *k = kh + xx * a.ksize_h;
        const unsigned *src = canvas + yy * stride + xmin;
        int r = 1 << 21, gch = 1 << 21, b = 1 << 21;
        for (int x = 0; x < xmax; ++x) {
          const unsigned px = src[x];
          const int c = k[x];
          r += (int)(px & 255u) * c;
          gch += (int)((px >> 8) & 255u) * c;
          b += (int)((px >> 16) & 255u) * c;
        }
        tmp[i] = clip8(r) | (clip8(gch) << 8) | (clip8(b) << 16);
      }
    }
    __syncthreads();
    if (live) {
      unsigned *small = canvas;  // the canvas is dead
      for (int i = t; i < OH * OW; i += T) {
        const int yy = i / OW, xx = i - yy * OW;
        const int ymin = bv[2 * yy], ymax = bv[2 * yy + 1];
        const int *k = kv + yy * a.ksize_v;
        int r = 1 << 21, gch = 1 << 21, b = 1 << 21;
        for (int y = 0; y < ymax; ++y) {
          const unsigned px = tmp[(y + ymin) * OW + xx];
          const int c = k[y];
          r += (int)(px & 255u) * c;
          gch += (int)((px >> 8) & 255u) * c;
          b += (int)((px >> 16) & 255u) * c;
        }
        small[i] = clip8(r) | (clip8(gch) << 8) | (clip8(b) << 16);
      }
    }
    __syncthreads();
    img_stride = OW;
  }
  if (live) {
    // pil_renderer.py:118-120: np.flipud -> output row j is image row OH-1-j
    unsigned char *out = a.frames + (size_t)n * OH * OW * 3;
    if ((OW & 15) == 0) {
      // 16 pixels of one row -> 48 bytes = three 16-byte stores
      uint4 *out4 = (uint4 *)out;
      const int groups = OH * (OW >> 4);
      for (int q = t; q < groups; q += T) {
        const int j = q / (OW >> 4), c0 = (q - j * (OW >> 4)) << 4;
        const unsigned *src = img + (OH - 1 - j) * img_stride + c0;
        unsigned w[12];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // 4 pixels (RGBX words) -> 3 packed words
          const unsigned p0 = src[4 * u] & 0xffffffu, p1 = src[4 * u + 1] & 0xffffffu;
          const unsigned p2 = src[4 * u + 2] & 0xffffffu, p3 = src[4 * u + 3] & 0xffffffu;
          w[3 * u] = p0 | (p1 << 24);
          w[3 * u + 1] = (p1 >> 8) | (p2 << 16);
          w[3 * u + 2] = (p2 >> 16) | (p3 << 8);
        }
        out4[3 * q] = make_uint4(w[0], w[1], w[2], w[3]);
        out4[3 * q + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        out4[3 * q + 2] = make_uint4(w[8], w[9], w[10], w[11]);
      }
    } else {
      const int nbytes = OH * OW * 3;
      for (int b = t; b < nbytes; b += T) {
        int p = b / 3, ch = b - 3 * p;
        int j = p / OW, col = p - j * OW;
        out[b] = (unsigned char)((img[(OH - 1 - j) * img_stride + col] >> (8 * ch)) & 255u);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter
// (SURVEY App. B.6), on the host: the coefficients go through libm's sin() like
// Pillow's, then to 22-bit fixed point.  Layout of the returned table:
//   bounds_h[2*OW] coeff_h[OW*ksize_h] bounds_v[2*OH] coeff_v[OH*ksize_v]
// ---------------------------------------------------------------------------
static double lanczos_weight(double x) {
  if (!(-3.0 <= x && x < 3.0)) return 0.0;
  auto sinc = [](double v) {
    if (v == 0.0) return 1.0;
    v = v * M_PI;
    return sin(v) / v;
  };
  return sinc(x) * sinc(x / 3);
}

static int resample_axis(int in_size, int out_size, std::vector<int> &bounds, std::vector<int> &coeffs) {
  double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
  double support = 3.0 * filterscale;
  int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign(2 * (size_t)out_size, 0);
  coeffs.assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (xx + 0.5) * scale, ww = 0.0, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      k[x] = lanczos_weight((x + xmin - center + 0.5) * ss);
      ww += k[x];
    }
    for (int x = 0; x < xmax; ++x) {
      double w = ww != 0.0 ? k[x] / ww : k[x];
      coeffs[(size_t)xx * ksize + x] = w < 0 ? (int)(-0.5 + w * (1 << 22)) : (int)(0.5 + w * (1 << 22));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

int resample_tables(int H, int W, int OH, int OW, std::vector<int> &table, int *ksize_h, int *ksize_v) {
  std::vector<int> bh, kh, bv, kv;
  *ksize_h = resample_axis(W, OW, bh, kh);
  *ksize_v = resample_axis(H, OH, bv, kv);
  table.clear();
  table.insert(table.end(), bh.begin(), bh.end());
  table.insert(table.end(), kh.begin(), kh.end());
  table.insert(table.end(), bv.begin(), bv.end());
  table.insert(table.end(), kv.begin(), kv.end());
  return (int)table.size();
}

cudaError_t launch_render(const RenderArgs &a, const int32_t *hdr, cudaStream_t stream, int *n_launches) {
  if (a.n_envs <= 0) return cudaSuccess;
  const int OH = hdr[MOOG_H_R_HEIGHT], OW = hdr[MOOG_H_R_WIDTH], aa = hdr[MOOG_H_R_AA];
  const int H = aa * OH, W = aa * OW, S = hdr[MOOG_H_N_SLOTS], VT = hdr[MOOG_H_N_VTX];
  const int P = (H <= 128) ? 2 : 1;  // threads per canvas row (they split its columns)
  const int T = (H * P + 31) & ~31;
  const int C = hdr[MOOG_H_R_MODIFIER] == MOOG_PMOD_TORUS ? 9 : 1;  // TorusGeometry: 9 copies per sprite
  RenderLayout lay = render_layout(H, W, S * C, (VT > 0 ? VT : 1) * C, OW);
  // envs per CTA: the split that keeps the most envs resident per SM (228 KB of shared
  // memory, 1 KB of it reserved per CTA); ties go to the larger CTA
  int epb = 1;
  {
    int best = 0;
    for (int c = 1; c * T <= 256 || c == 1; ++c) {
      const size_t per_cta = (size_t)lay.total * c + 1024;
      if (per_cta > 228 * 1024) break;
      const int envs = (int)(233472 / per_cta) * c;
      if (envs >= best) { best = envs; epb = c; }
    }
    static const char *epb_env = getenv("MOOG_RENDER_EPB");
    if (epb_env && atoi(epb_env) > 0) epb = atoi(epb_env);
  }
  size_t smem = (size_t)lay.total * epb;
  if (smem > 224 * 1024 || T > 1024) return cudaErrorInvalidConfiguration;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(moog_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured = smem;
  }
  int blocks = (a.n_envs + epb - 1) / epb;
  moog_render_kernel<<<blocks, epb * T, smem, stream>>>(a, T, epb, P);
  if (n_launches) *n_launches += 1;
  return cudaGetLastError();
}

}  // namespace moog
