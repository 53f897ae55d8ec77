// moog_render_dev.cuh -- device side of PILRenderer.__call__ (moog/observers/pil_renderer.py:88-120),
// shared by the stand-alone render kernel (moog_render.cu) and the step kernel, which draws the
// frame of the state it leaves an env in while the record is still in shared memory (moog_step.cu).
#pragma once
#include <math.h>

#include "moog_common.cuh"

namespace moog {

#define RENDER_MAXV MOOG_MAX_OUTLINE
#define MAX_XX (2 * RENDER_MAXV + 8)

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

__device__ __forceinline__ int round_up_(float f) { return (int)(f >= 0.0f ? floorf(f + 0.5f) : -floorf(fabsf(f) + 0.5f)); }
__device__ __forceinline__ int round_down_(float f) { return (int)(f >= 0.0f ? ceilf(f - 0.5f) : -ceilf(fabsf(f) - 0.5f)); }

// Draw.c: (y - e->y0) * e->dx + e->x0, float32, no fused multiply-add
__device__ __forceinline__ float edge_x(int x0, int y0, float dx, int y) {
  return __fadd_rn(__fmul_rn((float)(y - y0), dx), (float)x0);
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

// H, W: canvas size (anti_aliasing x image size); OW: image width.  ext_items: bytes of a
// buffer outside this layout that may hold the item spans (the step kernel's vertex cache, dead
// once the int vertices exist); when they fit there, L.items = -1
__host__ __device__ inline RenderLayout render_layout(int H, int W, int S, int VT, int OW, int ext_items = 0) {
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
    if (b <= ext_items) {
      L.items = -1;
      b = 0;
    }
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


// what render_env reads of one env: the SoA record (global or shared memory) and the program
struct RenderSrc {
  const double *dyn, *stat;
  const int32_t *meta, *cnt;
  const double2 *vtx;
  const int32_t *hdr, *voff;  // program header, slot -> first cached vertex
};

// PILRenderer.__call__ for ONE env by the T threads t = 0..T-1 of a CTA (every thread of the
// CTA group that shares `base` calls it with the same arguments; `sync` is a barrier over exactly
// those threads).  base: lay.total bytes of shared memory; ext_items: where the item spans go when
// lay.items < 0.  P threads share a canvas row.  out: [OH][OW][3] bytes of this env's frame.
// HB (0 = the whole canvas): rows per band.  A canvas that does not fit one CTA's shared memory
// (the shipped pacman draws 256 x 256, tests/runtime_benchmark.py up to 1024 x 1024) is drawn in
// bands of HB rows: `lay` is then the layout of an HB-row canvas, the edge lists are built once and
// every band scan-converts, fills and writes out its own rows (anti_aliasing 1 only: the Lanczos
// resize needs the rows around a band).
template <class Sync>
__device__ __forceinline__ void render_env(const RenderSrc &src, const RenderLayout &lay, unsigned char *base,
                                           unsigned *ext_items, int t, int T, int P, bool live, unsigned char *out,
                                           const int *resample, int ksize_h, int ksize_v, Sync sync, int HB = 0) {
  const int32_t *hdr = src.hdr;
  const int S = hdr[MOOG_H_N_SLOTS], L = hdr[MOOG_H_N_LAYERS], VT = hdr[MOOG_H_N_VTX];
  // pil_renderer.py:65-66: the canvas is anti_aliasing x the image size
  const int OH = hdr[MOOG_H_R_HEIGHT], OW = hdr[MOOG_H_R_WIDTH], aa = hdr[MOOG_H_R_AA];
  const int H = aa * OH, W = aa * OW;
  const unsigned bgc = (unsigned)hdr[MOOG_H_R_BG] & 0xffffffu;
  const int cmap = hdr[MOOG_H_R_COLORMAP], pmod = hdr[MOOG_H_R_MODIFIER], pml = hdr[MOOG_H_R_MOD_LAYER];
  // polygon_modifiers.py:88-96 TorusGeometry: every sprite is drawn as 9 copies shifted by
  // (i, j), i outer / j inner over (-1, 0, 1); copy c of slot s is the virtual slot s * C + c
  // and its vertices live at c * VT + voff[s]
  const int C = pmod == MOOG_PMOD_TORUS ? 9 : 1;
  unsigned *canvas = (unsigned *)(base + lay.canvas);
  int2 *ivtx = (int2 *)(base + lay.ivtx);
  unsigned *ink = (unsigned *)(base + lay.ink);
  int *symin = (int *)(base + lay.ymin), *symax = (int *)(base + lay.ymax), *shoriz = (int *)(base + lay.horiz);
  int *snedge = (int *)(base + lay.nedge), *sibase = (int *)(base + lay.ibase);
  ERec *erec = (ERec *)(base + lay.erec);
  unsigned *items = lay.items >= 0 ? (unsigned *)(base + lay.items) : ext_items;
  const int stride = W + 1;
  const double *dyn = src.dyn, *stat = src.stat;
  const int32_t *meta = src.meta, *cnt = src.cnt;
  const double2 *vtx = src.vtx;

  if (HB <= 0 || HB > H || aa > 1) HB = H;
  if (live) {
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
  }
  sync();
  // Draw order = layer order, then list order inside a layer (environment.py:20-25), which is
  // slot order over the LIVE slots (layer l owns the slots [LAYER_OFF[l], LAYER_OFF[l] + cnt[l])).
  // A slot that is not live or has no outline gets an empty row range, so that everything below
  // can walk the slots 0 .. S*C-1 without looking at layers again.
  if (live) {
    for (int vs = t; vs < S * C; vs += T) {
      const int s = vs / C, vo = (vs - s * C) * VT + src.voff[s];
      int nv = meta[MOOG_M_NV * S + s];
      bool drawn = false;
      for (int l = 0; l < L; ++l) {
        const int first = hdr[MOOG_H_LAYER_OFF + l];
        drawn = drawn || (s >= first && s < first + cnt[l]);
      }
      if (!drawn) nv = 0;
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
  sync();  // the scan below reads every slot's extents
#pragma unroll 1
  for (int y0 = 0; y0 < H; y0 += HB) {
  const int y1 = min(H, y0 + HB);  // this band: canvas rows [y0, y1)
  if (live)
    for (int i = t; i < (y1 - y0) * stride; i += T) canvas[i] = bgc;
  // (sprite, row) item ranges: exclusive prefix sum of the visible rows per slot by the first
  // warp; when the items do not all fit (lay.cap) thread 0 assigns them greedily instead and
  // the sprites left without a range are scan-converted by the row threads directly
  if (live && t < 32) {
    int acc = 0;
    for (int b = 0; b < S * C; b += 32) {
      const int vs = b + t;
      int rows = 0;
      if (vs < S * C) {
        // Draw.c polygon_generic: ymin = max(ymin, 0); ymax = min(ymax, H); rows >= H are clipped by hline
        const int lo = max(symin[vs], y0), hi = min(symax[vs], y1 - 1);
        rows = hi >= lo ? hi - lo + 1 : 0;
      }
      int x = rows;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (t >= d) x += y;
      }
      if (vs < S * C) sibase[vs] = rows > 0 ? acc + x - rows : -1;
      acc += __shfl_sync(0xffffffffu, x, 31);
    }
    __syncwarp();
    if (t == 0) {
      if (acc > lay.cap) {
        acc = 0;
        for (int s = 0; s < S * C; ++s) {
          const int lo = max(symin[s], y0), hi = min(symax[s], y1 - 1);
          const int rows = hi >= lo ? hi - lo + 1 : 0;
          if (rows > 0 && acc + rows <= lay.cap) {
            sibase[s] = acc;
            acc += rows;
          } else {
            sibase[s] = -1;
          }
        }
      }
      sibase[S * C] = acc;
    }
  }
  sync();
  // phase 1: one (sprite, row) item per thread pass -> clipped spans
  if (live) {
    const int n_items = sibase[S * C];
    int s = 0;  // virtual slot
    for (int it = t; it < n_items; it += T) {
      // items are sprite-major; find the sprite that owns item `it`
      for (;;) {
        int b0 = sibase[s];
        if (b0 >= 0) {
          int lo = max(symin[s], y0), hi = min(symax[s], y1 - 1);
          if (it < b0 + (hi - lo + 1)) break;
        }
        ++s;
      }
      const int lo = max(symin[s], y0);
      const int y = lo + (it - sibase[s]);
      SpanSink sink;
      sink.item = items + (size_t)it * (1 + ITEM_SPANS);
      sink.W = W;
      sink.n = 0;
      polygon_row_rec(erec + (s % C) * VT + src.voff[s / C], snedge[s], y, min(symax[s], H), shoriz[s] != 0, sink);
      sink.item[0] = sink.n <= ITEM_SPANS ? (unsigned)sink.n : ITEM_OVERFLOW;
    }
  }
  sync();
  // phase 2: P threads per row, sprites in z-order
  if (live) {
    const int HBn = y1 - y0;
    for (int w = t; w < HBn * P; w += T) {
      const int part = w / HBn, y = y0 + (w - part * HBn);
      const int xlo = (W * part) / P, xhi = (W * (part + 1)) / P - 1;  // P threads share a row
      unsigned *row = canvas + (y - y0) * stride;
      for (int s = 0; s < S * C; ++s) {
        const int ymin_c = max(symin[s], 0), ymax_c = min(symax[s], H);
        if (y < ymin_c || y > ymax_c) continue;
        const int s0 = C == 1 ? s : s / C;
        const unsigned color = ink[s0];
        const int b0 = sibase[s];
        unsigned cntw = ITEM_OVERFLOW;
        const unsigned *item = nullptr;
        if (b0 >= 0) {
          item = items + (size_t)(b0 + (y - max(ymin_c, y0))) * (1 + ITEM_SPANS);
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
          polygon_row_rec(erec + (s - s0 * C) * VT + src.voff[s0], snedge[s], y, ymax_c, shoriz[s] != 0, sink);
        }
      }
    }
  }
  sync();
  const unsigned *img = canvas;  // the image to write out: [OH][img_stride] RGBX words
  int img_stride = stride;
  if (aa > 1) {
    // Image.resize(LANCZOS) (Resample.c, 8 bpc): horizontal pass into an 8-bit
    // intermediate, then vertical pass; fixed-point coefficients from the host
    const int *bh = resample, *kh = bh + 2 * OW;
    const int *bv = kh + OW * ksize_h, *kv = bv + 2 * OH;
    unsigned *tmp = (unsigned *)(base + lay.tmp);
    if (live) {
      for (int i = t; i < H * OW; i += T) {
        const int yy = i / OW, xx = i - yy * OW;
        const int xmin = bh[2 * xx], xmax = bh[2 * xx + 1];
        const int *k = kh + xx * ksize_h;
        const unsigned *srcp = canvas + yy * stride + xmin;
        int r = 1 << 21, gch = 1 << 21, b = 1 << 21;
        for (int x = 0; x < xmax; ++x) {
          const unsigned px = srcp[x];
          const int c = k[x];
          r += (int)(px & 255u) * c;
          gch += (int)((px >> 8) & 255u) * c;
          b += (int)((px >> 16) & 255u) * c;
        }
        tmp[i] = clip8(r) | (clip8(gch) << 8) | (clip8(b) << 16);
      }
    }
    sync();
    if (live) {
      unsigned *small = canvas;  // the canvas is dead
      for (int i = t; i < OH * OW; i += T) {
        const int yy = i / OW, xx = i - yy * OW;
        const int ymin = bv[2 * yy], ymax = bv[2 * yy + 1];
        const int *k = kv + yy * ksize_v;
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
    sync();
    img_stride = OW;
  }
  if (live) {
    // pil_renderer.py:118-120: np.flipud -> output row j is image row OH-1-j
    // (a band holds the image rows [y0, y1), i.e. the output rows [OH - y1, OH - y0); with
    // anti-aliasing there is one band and the resized image starts at row 0)
    const int r0 = aa > 1 ? 0 : y0, r1 = aa > 1 ? OH : y1;
    if ((OW & 15) == 0) {
      // consecutive threads store consecutive 16-byte chunks of the frame (a row is OW * 3 / 16
      // of them); packed word m = 3u + r of a row holds bytes r.. of pixel 4u + r and the first
      // bytes of the next pixel
      uint4 *out4 = (uint4 *)out;
      const int cpr = (OW * 3) >> 4;
      for (int q0 = t; q0 < (r1 - r0) * cpr; q0 += T) {
        const int q = q0 + (OH - r1) * cpr;
        const int j = q / cpr, c = q - j * cpr;
        const unsigned *srcp = img + (OH - 1 - j - r0) * img_stride;
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = 4 * c + i, u = m / 3, r = m - 3 * u;
          const unsigned lo = srcp[4 * u + r] & 0xffffffu, hi = srcp[4 * u + r + 1] & 0xffffffu;
          w[i] = (lo >> (8 * r)) | (hi << (24 - 8 * r));
        }
        out4[q] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      const int nbytes = (r1 - r0) * OW * 3;
      for (int b0 = t; b0 < nbytes; b0 += T) {
        const int b = b0 + (OH - r1) * OW * 3;
        int p = b / 3, ch = b - 3 * p;
        int j = p / OW, col = p - j * OW;
        out[b] = (unsigned char)((img[(OH - 1 - j - r0) * img_stride + col] >> (8 * ch)) & 255u);
      }
    }
  }
  if (y0 + HB < H) sync();  // the next band reuses the canvas and the item spans
  }
}

}  // namespace moog
