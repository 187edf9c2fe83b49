/* TEST INFRASTRUCTURE - CPU oracle of HM 16.14's deblocking filter for all-intra pictures (first
 * stage of the per-picture tail).  Plain-C restatement of TComLoopFilter::loopFilterPic
 * (/root/reference/source/Lib/TLibCommon/TComLoopFilter.cpp:130-954) on the records the hot path leaves behind
 * (hm_ctu_out: CU depth, transform index, pre-deblocking reconstruction).  The checker of fhc_deblock_pictures; nothing
 * in the product may include or link it.  PINNED against the reference encoder's reconstruction file written with
 * --SAO=0 (tests/test_deblock.py, tests/golden/deblock_*.npz).
 *
 * What the all-intra cfg leaves of the general filter: every TU / CU edge on the 8x8 luma grid inside the picture has
 * boundary strength 2 (xGetBoundaryStrengthSingle :416: intra on either side); QpP = QpQ = slice QP (no delta QP); beta
 * and tc offsets 0; no PCM / lossless; chroma edges on the 16-sample luma grid, chroma QP through the 4:2:0 table with
 * offset 0.  All vertical edges of the picture are filtered first, then the horizontal ones (:130-157). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "hm_intra.h"

static const uint8_t kTc[54] = {                                                   /* sm_tcTable :58 */
  0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };
static const uint8_t kBeta[52] = {                                                 /* sm_betaTable :63 */
  0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };
static const uint8_t kChromaScale420d[58] = {                                      /* g_aucChromaScale[CHROMA_420], TComRom.cpp:532 */
  0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,
  38,39,40,41,42,43,44,45,46,47,48,49,50,51 };

static int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static int iabs(int v) { return v < 0 ? -v : v; }

typedef struct { int W, H, wc, bd; const hm_ctu_out* ctu; int16_t* pl[3]; int qp; } Dbk;

/* z-scan index of the 4x4 partition at raster position (px, py) inside a CTU (g_auiRasterToZscan) */
static int z_of(int px, int py)
{
  int z = 0, b;
  for (b = 0; b < 4; b++) z |= (((px >> b) & 1) << (2 * b)) | (((py >> b) & 1) << (2 * b + 1));
  return z;
}
/* Is luma position (x, y) on the left (dir 0) / top (dir 1) boundary of the transform unit that contains it?
 * xSetEdgefilterTU :274 + xSetEdgefilterPU :299 (CU edges coincide with TU edges; the internal edge of an NxN 8x8 CU
 * lies off the 8x8 grid); picture boundary: bLeftEdge / bTopEdge false (:362-415). */
static int tu_edge(const Dbk* d, int x, int y, int dir)
{
  const hm_ctu_out* c = &d->ctu[(y >> 6) * d->wc + (x >> 6)];
  const int z = z_of((x & 63) >> 2, (y & 63) >> 2);
  const int tu = 64 >> (c->depth[z] + c->trIdx[z]);
  const int v = dir == 0 ? x : y;
  return v > 0 && (v & (tu - 1)) == 0;
}

/* xPelFilterLuma :830 on one line across the edge; p[0] = first sample of the Q side */
static void pel_luma(int16_t* p, int off, int tc, int sw, int thrCut, int filtP, int filtQ, int bd)
{
  const int m4 = p[0], m3 = p[-off], m5 = p[off], m2 = p[-2 * off], m6 = p[2 * off], m1 = p[-3 * off], m7 = p[3 * off], m0 = p[-4 * off];
  const int maxv = (1 << bd) - 1;
  if (sw)
  {
    p[-off]     = (int16_t)clip3i(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
    p[0]        = (int16_t)clip3i(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
    p[-2 * off] = (int16_t)clip3i(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
    p[off]      = (int16_t)clip3i(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
    p[-3 * off] = (int16_t)clip3i(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
    p[2 * off]  = (int16_t)clip3i(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
  }
  else
  {
    int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
    if (iabs(delta) < thrCut)
    {
      const int tc2 = tc >> 1;
      delta = clip3i(-tc, tc, delta);
      p[-off] = (int16_t)clip3i(0, maxv, m3 + delta);
      p[0] = (int16_t)clip3i(0, maxv, m4 - delta);
      if (filtP) p[-2 * off] = (int16_t)clip3i(0, maxv, m2 + clip3i(-tc2, tc2, ((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1)));
      if (filtQ) p[off] = (int16_t)clip3i(0, maxv, m5 + clip3i(-tc2, tc2, ((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1)));
    }
  }
}
static int use_strong(const int16_t* p, int off, int dd, int beta, int tc)        /* xUseStrongFiltering :933 */
{
  const int m4 = p[0], m3 = p[-off], m7 = p[3 * off], m0 = p[-4 * off];
  return (iabs(m0 - m3) + iabs(m7 - m4) < (beta >> 3)) && (dd < (beta >> 2)) && (iabs(m3 - m4) < ((tc * 5 + 1) >> 1));
}
static int calc_dp(const int16_t* p, int off) { return iabs(p[-3 * off] - 2 * p[-2 * off] + p[-off]); }
static int calc_dq(const int16_t* p, int off) { return iabs(p[0] - 2 * p[off] + p[2 * off]); }

/* one 4-line segment of a luma edge whose first Q-side sample is (x, y): xEdgeFilterLuma :557 */
static void luma_segment(Dbk* d, int x, int y, int dir)
{
  const int stride = d->W, off = dir == 0 ? 1 : stride, step = dir == 0 ? stride : 1;
  int16_t* s = d->pl[0] + y * stride + x;
  const int scale = 1 << (d->bd - 8);
  const int tc = kTc[clip3i(0, 53, d->qp + 2)] * scale, beta = kBeta[clip3i(0, 51, d->qp)] * scale;
  const int side = (beta + (beta >> 1)) >> 3, thrCut = tc * 10;
  const int dp0 = calc_dp(s, off), dq0 = calc_dq(s, off), dp3 = calc_dp(s + 3 * step, off), dq3 = calc_dq(s + 3 * step, off);
  const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, dd = d0 + d3;
  int i;
  if (dd < beta)
  {
    const int sw = use_strong(s, off, 2 * d0, beta, tc) && use_strong(s + 3 * step, off, 2 * d3, beta, tc);
    for (i = 0; i < 4; i++) pel_luma(s + i * step, off, tc, sw, thrCut, dp < side, dq < side, d->bd);
  }
}
/* one 2-line chroma segment (the chroma samples of a 4-line luma segment): xEdgeFilterChroma :676, xPelFilterChroma :901 */
static void chroma_segment(Dbk* d, int comp, int cx, int cy, int dir)
{
  const int stride = d->W >> 1, off = dir == 0 ? 1 : stride, step = dir == 0 ? stride : 1;
  int16_t* s = d->pl[comp] + cy * stride + cx;
  const int maxv = (1 << d->bd) - 1;
  int qp = d->qp, i;                                                  /* ((QpP + QpQ + 1) >> 1) + cQpPicOffset (0) */
  if (qp >= 58) qp -= 6; else if (qp >= 0) qp = kChromaScale420d[qp];
  {
    const int tc = kTc[clip3i(0, 53, qp + 2)] * (1 << (d->bd - 8));
    for (i = 0; i < 2; i++)
    {
      int16_t* p = s + i * step;
      const int m4 = p[0], m3 = p[-off], m5 = p[off], m2 = p[-2 * off];
      const int delta = clip3i(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
      p[-off] = (int16_t)clip3i(0, maxv, m3 + delta);
      p[0] = (int16_t)clip3i(0, maxv, m4 - delta);
    }
  }
}

/* Deblocks one picture.  ctus: the records of the picture (raster order); planes: out, int16 planar Y [H][W], Cb, Cr
 * [H/2][W/2] - the picture as TComPicYuv holds it after loopFilterPic. */
int hm_deblock_picture(int width, int height, int bitDepth, int qp, const hm_ctu_out* ctus, int16_t* planeY, int16_t* planeCb, int16_t* planeCr)
{
  Dbk d;
  int c, x, y, dir;
  d.W = width; d.H = height; d.wc = (width + 63) / 64; d.bd = bitDepth; d.ctu = ctus; d.qp = qp;
  d.pl[0] = planeY; d.pl[1] = planeCb; d.pl[2] = planeCr;
  for (c = 0; c < 3; c++)
  {
    const int sh = c ? 1 : 0, n = 64 >> sh, w = width >> sh, h = height >> sh;
    for (y = 0; y < h; y++)
      for (x = 0; x < w; x++)
      {
        const hm_ctu_out* o = &ctus[(y / n) * d.wc + x / n];
        const int16_t* src = c == 0 ? o->recoY : (c == 1 ? o->recoCb : o->recoCr);
        d.pl[c][y * w + x] = src[(y % n) * n + (x % n)];
      }
  }
  for (dir = 0; dir < 2; dir++)
    for (y = 0; y < height; y += (dir == 0 ? 4 : 8))
      for (x = 0; x < width; x += (dir == 0 ? 8 : 4))
      {
        if (!tu_edge(&d, x, y, dir)) continue;
        luma_segment(&d, x, y, dir);
        /* chroma: edges on the 16-sample luma grid only (:701-712) */
        if (((dir == 0 ? x : y) & 15) == 0)
          for (c = 1; c < 3; c++) chroma_segment(&d, c, x >> 1, y >> 1, dir);
      }
  return 0;
}
