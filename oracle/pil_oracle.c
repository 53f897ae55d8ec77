/*
 * pil_oracle.c -- CPU restatement of MOOG's PILRenderer.__call__.
 *
 * TEST INFRASTRUCTURE ONLY (see moog_oracle.c header).
 *
 * Follows moog/observers/pil_renderer.py:88-120 (background paste, z-ordered
 * ImageDraw.polygon with an RGBA fill on an RGB canvas, LANCZOS resize, flipud),
 * moog/observers/color_maps.py:21-23 (hsv_to_rgb + uint8 truncation) and
 * moog/observers/polygon_modifiers.py:32-98.  The pixel work the reference
 * delegates to Pillow (not vendored in the reference; unpinned in its setup.py;
 * 12.2.0 in this image) is restated from Pillow's published sources
 * src/libImaging/Draw.c (ImagingDrawPolygon / polygon_generic / hline32rgba)
 * and src/libImaging/Resample.c (ImagingResample, 8 bpc path).
 *
 * Pinning: tests/test_oracle_render.py compares this file bit for bit with the
 * installed Pillow on random polygons / scenes and (in the build container)
 * with the reference's own PILRenderer on live states.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/moog_b200_program.h"

#define MAXV MOOG_MAX_OUTLINE

typedef struct {
  int d;
  int x0, y0;
  int xmin, ymin, xmax, ymax;
  float dx;
} edge_t;

/* Draw.c add_edge */
static void add_edge(edge_t *e, int x0, int y0, int x1, int y1) {
  if (x0 <= x1) { e->xmin = x0; e->xmax = x1; } else { e->xmin = x1; e->xmax = x0; }
  if (y0 <= y1) { e->ymin = y0; e->ymax = y1; } else { e->ymin = y1; e->ymax = y0; }
  if (y0 == y1) {
    e->d = 0;
    e->dx = 0.0f;
  } else {
    e->dx = ((float)(x1 - x0)) / (float)(y1 - y0);
    e->d = (y0 == e->ymin) ? 1 : -1;
  }
  e->x0 = x0;
  e->y0 = y0;
}

/* Draw.c: DIV255 / BLEND as used by hline32rgba in Pillow >= 10 */
static inline unsigned div255(unsigned a) { return (((a + 128) >> 8) + (a + 128)) >> 8; }

typedef struct {
  uint8_t *px; /* [H][W][3] */
  int W, H;
} canvas_t;

/* Draw.c hline32rgba (clipped) */
static void hline(canvas_t *c, int x0, int y, int x1, const uint8_t ink[4]) {
  if (y < 0 || y >= c->H) return;
  if (x0 < 0) x0 = 0; else if (x0 >= c->W) return;
  if (x1 < 0) return; else if (x1 >= c->W) x1 = c->W - 1;
  if (x0 > x1) return;
  unsigned a = ink[3];
  uint8_t *p = c->px + ((size_t)y * c->W + x0) * 3;
  for (int x = x0; x <= x1; ++x, p += 3)
    for (int k = 0; k < 3; ++k) p[k] = (uint8_t)div255(p[k] * (255 - a) + ink[k] * a);
}

static inline int round_up_(float f) { return (int)(f >= 0.0f ? floorf(f + 0.5f) : -floorf(fabsf(f) + 0.5f)); }
static inline int round_down_(float f) { return (int)(f >= 0.0f ? ceilf(f - 0.5f) : -ceilf(fabsf(f) - 0.5f)); }

static inline float edge_x(const edge_t *e, int y) {
  /* Draw.c: (y - e->y0) * e->dx + e->x0 in float, no fused multiply-add */
  volatile float t = (float)(y - e->y0) * e->dx;
  return t + (float)e->x0;
}

static int cmp_float(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

/* Draw.c draw_horizontal_lines */
static void draw_horizontal_lines(canvas_t *c, int n, const edge_t *e, const uint8_t ink[4], int *x_pos,
                                  int y) {
  for (int i = 0; i < n; ++i) {
    if (e[i].ymin == y && e[i].ymin == e[i].ymax) {
      int xmin = e[i].xmin;
      if (*x_pos != -1 && *x_pos < xmin) continue;
      int xmax = e[i].xmax;
      if (*x_pos > xmin) {
        xmin = *x_pos;
        if (xmax < xmin) continue;
      }
      hline(c, xmin, y, xmax, ink);
      *x_pos = xmax + 1;
    }
  }
}

/* Draw.c polygon_generic with a blending hline (RGBA draw mode) */
static void polygon_generic(canvas_t *c, int n, edge_t *e, const uint8_t ink[4]) {
  if (n <= 0) return;
  edge_t *table[2 * MAXV + 4];
  float xx[4 * MAXV + 8];
  int edge_count = 0;
  int ymin = c->H - 1, ymax = 0;
  for (int i = 0; i < n; ++i) {
    if (ymin > e[i].ymin) ymin = e[i].ymin;
    if (ymax < e[i].ymax) ymax = e[i].ymax;
    if (e[i].ymin == e[i].ymax) continue; /* horizontal edges are deferred when blending */
    table[edge_count++] = e + i;
  }
  if (ymin < 0) ymin = 0;
  if (ymax > c->H) ymax = c->H;
  for (; ymin <= ymax; ++ymin) {
    int j = 0, x_pos = 0;
    for (int i = 0; i < edge_count; ++i) {
      edge_t *cur = table[i];
      if (ymin >= cur->ymin && ymin <= cur->ymax) {
        xx[j++] = edge_x(cur, ymin);
        if (ymin == cur->ymax && ymin < ymax) {
          xx[j] = xx[j - 1];
          j++;
        } else if ((ymin == cur->ymin || ymin == cur->ymax) && cur->dx != 0) {
          for (int k = 0; k < i; ++k) {
            edge_t *oth = table[k];
            if ((ymin != oth->ymin && ymin != oth->ymax) || oth->dx == 0) continue;
            if (roundf(xx[j - 1]) == roundf(edge_x(oth, ymin))) {
              int off = (ymin == ymax) ? -1 : 1;
              if (ymin + off >= oth->ymin && ymin + off <= oth->ymax) {
                float adj = edge_x(cur, ymin + off);
                float oadj = edge_x(oth, ymin + off);
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
    qsort(xx, j, sizeof(float), cmp_float);
    x_pos = j == 0 ? -1 : 0;
    for (int i = 1; i < j; i += 2) {
      int x_end = round_down_(xx[i]);
      if (x_end < x_pos) continue;
      draw_horizontal_lines(c, n, e, ink, &x_pos, ymin);
      if (x_end < x_pos) continue;
      int x_start = round_up_(xx[i - 1]);
      if (x_pos > x_start) {
        x_start = x_pos;
        if (x_end < x_start) continue;
      }
      hline(c, x_start, ymin, x_end, ink);
      x_pos = x_end + 1;
    }
    draw_horizontal_lines(c, n, e, ink, &x_pos, ymin);
  }
}

/* Draw.c ImagingDrawPolygon(fill=1) on integer points xy[count] */
static void draw_polygon(canvas_t *c, int count, const int *xy, const uint8_t ink[4]) {
  if (count <= 0) return;
  edge_t e[MAXV + 4];
  int n = 0;
  if (count == 1 || (count == 2 && 0)) { /* single point: degenerate, nothing to fill */ }
  for (int i = 0; i < count - 1; ++i) {
    int x_diff = xy[2 * i + 2] - xy[2 * i];
    int y_diff = xy[2 * i + 3] - xy[2 * i + 1];
    (void)x_diff;
    if (y_diff == 0 && i != 0 && xy[2 * i + 1] == xy[2 * i - 1]) {
      /* Pillow: extend the previous horizontal edge instead of adding one */
      int x0 = xy[2 * i], x1 = xy[2 * i + 2], xp = xy[2 * i - 2];
      if (x1 > x0 && x0 > xp) {
        e[n - 1].xmax = x1;
        continue;
      } else if (x1 < x0 && x0 < xp) {
        e[n - 1].xmin = x1;
        continue;
      }
    }
    add_edge(&e[n++], xy[2 * i], xy[2 * i + 1], xy[2 * i + 2], xy[2 * i + 3]);
  }
  if (xy[2 * count - 2] != xy[0] || xy[2 * count - 1] != xy[1])
    add_edge(&e[n++], xy[2 * count - 2], xy[2 * count - 1], xy[0], xy[1]);
  polygon_generic(c, n, e, ink);
}

/* standalone entry for the Pillow differential test: float vertices, C int cast */
void orc_pil_polygon(uint8_t *px, int W, int H, const double *xy, int count, const uint8_t *ink) {
  canvas_t c = {px, W, H};
  int ixy[2 * (MAXV + 4)];
  if (count > MAXV + 2) count = MAXV + 2;
  for (int i = 0; i < 2 * count; ++i) ixy[i] = (int)xy[i];
  draw_polygon(&c, count, ixy, ink);
}

/* ------------------------------------------------------------------------ */
/* Resample.c: ImagingResample(LANCZOS) for 8-bit RGB                        */
/* ------------------------------------------------------------------------ */

static double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
static double lanczos_filter(double x) {
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}

#define PRECISION_BITS (32 - 8 - 2)

static inline uint8_t clip8(int in) {
  int v = in >> PRECISION_BITS;
  if (v < 0) return 0;
  if (v > 255) return 255;
  return (uint8_t)v;
}

/* precompute_coeffs + normalize_coeffs_8bpc */
static int precompute_coeffs(int inSize, int outSize, int **boundsp, int **kkp) {
  double support, scale, filterscale;
  double center, ww, ss;
  int xx, x, ksize, xmin, xmax;
  filterscale = scale = (double)inSize / outSize;
  if (filterscale < 1.0) filterscale = 1.0;
  support = 3.0 * filterscale;
  ksize = (int)ceil(support) * 2 + 1;
  double *kk = (double *)malloc(sizeof(double) * outSize * ksize);
  int *bounds = (int *)malloc(sizeof(int) * outSize * 2);
  for (xx = 0; xx < outSize; xx++) {
    center = (xx + 0.5) * scale;
    ww = 0.0;
    ss = 1.0 / filterscale;
    xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    xmax = (int)(center + support + 0.5);
    if (xmax > inSize) xmax = inSize;
    xmax -= xmin;
    double *k = &kk[xx * ksize];
    for (x = 0; x < xmax; x++) {
      double w = lanczos_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; x++)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; x++) k[x] = 0;
    bounds[xx * 2 + 0] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  int *ikk = (int *)malloc(sizeof(int) * outSize * ksize);
  for (x = 0; x < outSize * ksize; x++) {
    if (kk[x] < 0)
      ikk[x] = (int)(-0.5 + kk[x] * (1 << PRECISION_BITS));
    else
      ikk[x] = (int)(0.5 + kk[x] * (1 << PRECISION_BITS));
  }
  free(kk);
  *boundsp = bounds;
  *kkp = ikk;
  return ksize;
}

/* in: [Hin][Win][3] -> out: [Hout][Wout][3]; horizontal pass then vertical pass */
void orc_lanczos_resize(const uint8_t *in, int Win, int Hin, uint8_t *out, int Wout, int Hout) {
  if (Win == Wout && Hin == Hout) {
    memcpy(out, in, (size_t)Win * Hin * 3);
    return;
  }
  int *bh, *kh, *bv, *kv;
  int ksh = precompute_coeffs(Win, Wout, &bh, &kh);
  int ksv = precompute_coeffs(Hin, Hout, &bv, &kv);
  /* Resample.c only resamples the rows the vertical pass needs, which does not
   * change any value; resample all rows. */
  uint8_t *tmp = (uint8_t *)malloc((size_t)Hin * Wout * 3);
  const uint8_t *src = in;
  if (Win != Wout) {
    for (int yy = 0; yy < Hin; ++yy)
      for (int xx = 0; xx < Wout; ++xx) {
        int xmin = bh[2 * xx], xmax = bh[2 * xx + 1];
        const int *k = &kh[xx * ksh];
        for (int ch = 0; ch < 3; ++ch) {
          int ss = 1 << (PRECISION_BITS - 1);
          for (int x = 0; x < xmax; ++x) ss += in[((size_t)yy * Win + x + xmin) * 3 + ch] * k[x];
          tmp[((size_t)yy * Wout + xx) * 3 + ch] = clip8(ss);
        }
      }
    src = tmp;
  }
  if (Hin != Hout) {
    for (int yy = 0; yy < Hout; ++yy) {
      int ymin = bv[2 * yy], ymax = bv[2 * yy + 1];
      const int *k = &kv[yy * ksv];
      for (int xx = 0; xx < Wout; ++xx)
        for (int ch = 0; ch < 3; ++ch) {
          int ss = 1 << (PRECISION_BITS - 1);
          for (int y = 0; y < ymax; ++y) ss += src[((size_t)(y + ymin) * Wout + xx) * 3 + ch] * k[y];
          out[((size_t)yy * Wout + xx) * 3 + ch] = clip8(ss);
        }
    }
  } else {
    memcpy(out, src, (size_t)Hout * Wout * 3);
  }
  free(tmp);
  free(bh); free(kh); free(bv); free(kv);
}

/* ------------------------------------------------------------------------ */
/* color_maps.py:21-23 (CPython colorsys.hsv_to_rgb, x255, astype(uint8))    */
/* ------------------------------------------------------------------------ */

static inline uint8_t to_u8(double v) { return (uint8_t)(int64_t)v; }

void orc_color_to_rgb(int cmap, double c0, double c1, double c2, uint8_t rgb[3]) {
  if (cmap != MOOG_CMAP_HSV) {
    rgb[0] = to_u8(c0); rgb[1] = to_u8(c1); rgb[2] = to_u8(c2);
    return;
  }
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
  rgb[0] = to_u8(255 * r); rgb[1] = to_u8(255 * g); rgb[2] = to_u8(255 * b);
}

/* ------------------------------------------------------------------------ */
/* pil_renderer.py:88-120                                                    */
/* ------------------------------------------------------------------------ */

typedef struct {
  double *dyn, *stat;
  int32_t *meta, *cnt, *envi;
  double *envf;
  double *vtx;
} orc_state_r;

void orc_render(const void *blob, const orc_state_r *st, int n_envs, uint8_t *out) {
  const int32_t *hdr = (const int32_t *)blob;
  const moog_op *ops = (const moog_op *)(hdr + MOOG_HDR_WORDS);
  const int32_t *ipool = (const int32_t *)(ops + hdr[MOOG_H_N_OPS]);
  const int32_t *voff = ipool + hdr[MOOG_H_VOFF];
  int S = hdr[MOOG_H_N_SLOTS], L = hdr[MOOG_H_N_LAYERS], VT = hdr[MOOG_H_N_VTX];
  int H = hdr[MOOG_H_R_HEIGHT], W = hdr[MOOG_H_R_WIDTH], aa = hdr[MOOG_H_R_AA];
  int Wc = aa * W, Hc = aa * H;
  unsigned bg = (unsigned)hdr[MOOG_H_R_BG];
  int cmap = hdr[MOOG_H_R_COLORMAP], pmod = hdr[MOOG_H_R_MODIFIER], pml = hdr[MOOG_H_R_MOD_LAYER];
  uint8_t *canvas = (uint8_t *)malloc((size_t)Wc * Hc * 3);
  uint8_t *small = (uint8_t *)malloc((size_t)W * H * 3);
  for (int n = 0; n < n_envs; ++n) {
    const double *dyn = st->dyn + (size_t)n * MOOG_DYN_FIELDS * S;
    const double *stat = st->stat + (size_t)n * MOOG_STAT_FIELDS * S;
    const int32_t *meta = st->meta + (size_t)n * MOOG_META_FIELDS * S;
    const int32_t *cnt = st->cnt + (size_t)n * MOOG_MAX_LAYERS;
    const double *vtx = st->vtx + (size_t)n * 2 * VT;
    for (size_t i = 0; i < (size_t)Wc * Hc; ++i) {
      canvas[3 * i] = bg & 255; canvas[3 * i + 1] = (bg >> 8) & 255; canvas[3 * i + 2] = (bg >> 16) & 255;
    }
    canvas_t c = {canvas, Wc, Hc};
    double dx = 0, dy = 0;
    if (pmod == MOOG_PMOD_FIRST_PERSON) { /* polygon_modifiers.py:54-63 */
      int s = hdr[MOOG_H_LAYER_OFF + pml];
      dx = 0.5 - dyn[MOOG_D_X * S + s];
      dy = 0.5 - dyn[MOOG_D_Y * S + s];
    }
    for (int l = 0; l < L; ++l)
      for (int k = 0; k < cnt[l]; ++k) {
        int s = hdr[MOOG_H_LAYER_OFF + l] + k;
        int nv = meta[MOOG_M_NV * S + s];
        const double *v = vtx + 2 * (size_t)voff[s];
        uint8_t ink[4];
        orc_color_to_rgb(cmap, stat[MOOG_S_C0 * S + s], stat[MOOG_S_C1 * S + s], stat[MOOG_S_C2 * S + s], ink);
        ink[3] = to_u8(stat[MOOG_S_OPACITY * S + s]);
        int ncopy = (pmod == MOOG_PMOD_TORUS) ? 9 : 1;
        for (int q = 0; q < ncopy; ++q) {
          double ox = dx, oy = dy;
          if (pmod == MOOG_PMOD_TORUS) { /* polygon_modifiers.py:88-96: i outer, j inner */
            ox = (double)(q / 3 - 1);
            oy = (double)(q % 3 - 1);
          }
          int ixy[2 * MAXV];
          for (int i = 0; i < nv; ++i) {
            double x = v[2 * i], y = v[2 * i + 1];
            if (pmod != MOOG_PMOD_NONE) { x = x + ox; y = y + oy; }
            ixy[2 * i] = (int)((double)Wc * x);
            ixy[2 * i + 1] = (int)((double)Hc * y);
          }
          draw_polygon(&c, nv, ixy, ink);
        }
      }
    orc_lanczos_resize(canvas, Wc, Hc, small, W, H);
    uint8_t *o = out + (size_t)n * H * W * 3;
    for (int r = 0; r < H; ++r) memcpy(o + (size_t)r * W * 3, small + (size_t)(H - 1 - r) * W * 3, (size_t)W * 3);
  }
  free(canvas);
  free(small);
}
