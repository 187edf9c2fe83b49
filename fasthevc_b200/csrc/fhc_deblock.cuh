// fhc_deblock.cuh - deblocking of finished pictures on the device (first stage of the per-picture
// tail): TComLoopFilter::loopFilterPic (source/Lib/TLibCommon/TComLoopFilter.cpp:130-954) for the all-intra cfg, from
// the CTU records the mode decision leaves in device memory.  Included by fasthevc_cuda.cu (device build) and by the
// test-only host build of the kernel source (tests/emu_abi.cpp), which runs the same segment functions.
//
// What the all-intra cfg leaves of the general filter: every TU / CU edge on the 8x8 luma grid inside the picture has
// boundary strength 2 (xGetBoundaryStrengthSingle :416: a side is intra); QpP = QpQ = slice QP; beta / tc offsets 0;
// no PCM / lossless; chroma edges on the 16-sample luma grid with the 4:2:0 chroma QP table.  All vertical edges of
// the picture first, then the horizontal ones (:130-157).  An edge segment (4 lines) reads 4 samples and changes at
// most 3 on each side, and edges are 8 samples apart: the segments of one direction touch disjoint samples, so each
// direction is one pass with a thread per segment.
//
// HBM-bound integer work: per picture the passes read the records' reconstruction (2 B / sample), read and write the
// int16 planes around the edges, and write the packed output (1 or 2 B / sample).
#ifndef FHC_DEBLOCK_CUH
#define FHC_DEBLOCK_CUH
#include <stdint.h>
#include "../../include/fasthevc_cuda.h"

#if defined(__CUDACC__) && defined(FHC_DEVICE_BUILD)
#define FHC_DBK_HD __host__ __device__ __forceinline__
#else
#define FHC_DBK_HD static inline
#endif

namespace fhc {

struct DbkPic                     // one picture of a deblocking call
{
  const fhc_ctu_result* ctus;     // the picture's records, raster order
  int16_t* plane[3];              // working planes Y [H][W], Cb, Cr [H/2][W/2]
  void* out;                      // packed output, layout of the input pictures (Y, Cb, Cr; sampleBytes each)
  int32_t qp, pad;
};

struct DbkTables { uint8_t tc[54], beta[52], chromaScale[58]; };
// sm_tcTable (:58), sm_betaTable (:63), g_aucChromaScale[CHROMA_420] (TComRom.cpp:532)
#define FHC_DBK_TABLES { \
  {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24}, \
  {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64}, \
  {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51} }

FHC_DBK_HD int dbk_clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
FHC_DBK_HD int dbk_abs(int v) { return v < 0 ? -v : v; }

// z-scan index of the 4x4 partition at raster position (px, py) of a CTU (g_auiRasterToZscan): bit interleave
FHC_DBK_HD int dbk_z(int px, int py)
{
  int z = 0;
  for (int b = 0; b < 4; b++) z |= (((px >> b) & 1) << (2 * b)) | (((py >> b) & 1) << (2 * b + 1));
  return z;
}

// Is luma position (x, y) on the left (dir 0) / top (dir 1) boundary of the transform unit containing it, inside the
// picture?  xSetEdgefilterTU :274, xSetEdgefilterPU :299 (CU edges are TU edges; the internal edge of an NxN 8x8 CU lies
// off the 8x8 grid), xSetLoopfilterParam :362 (no filtering across the picture boundary).
FHC_DBK_HD int dbk_tu_edge(const fhc_ctu_result* ctus, int wc, int x, int y, int dir)
{
  const fhc_ctu_result* c = &ctus[(y >> 6) * wc + (x >> 6)];
  const int z = dbk_z((x & 63) >> 2, (y & 63) >> 2);
  const int tu = 64 >> (c->depth[z] + c->trIdx[z]);
  const int v = dir == 0 ? x : y;
  return v > 0 && (v & (tu - 1)) == 0;
}

// xPelFilterLuma :830 on one line across the edge; p[0] = first sample of the Q side
FHC_DBK_HD void dbk_pel_luma(int16_t* p, int off, int tc, int sw, int thrCut, int filtP, int filtQ, int maxv)
{
  const int m4 = p[0], m3 = p[-off], m5 = p[off], m2 = p[-2 * off], m6 = p[2 * off], m1 = p[-3 * off], m7 = p[3 * off], m0 = p[-4 * off];
  if (sw)
  {
    p[-off]     = (int16_t)dbk_clip3(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
    p[0]        = (int16_t)dbk_clip3(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
    p[-2 * off] = (int16_t)dbk_clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
    p[off]      = (int16_t)dbk_clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
    p[-3 * off] = (int16_t)dbk_clip3(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
    p[2 * off]  = (int16_t)dbk_clip3(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
  }
  else
  {
    int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
    if (dbk_abs(delta) < thrCut)
    {
      const int tc2 = tc >> 1;
      delta = dbk_clip3(-tc, tc, delta);
      p[-off] = (int16_t)dbk_clip3(0, maxv, m3 + delta);
      p[0] = (int16_t)dbk_clip3(0, maxv, m4 - delta);
      if (filtP) p[-2 * off] = (int16_t)dbk_clip3(0, maxv, m2 + dbk_clip3(-tc2, tc2, ((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1)));
      if (filtQ) p[off] = (int16_t)dbk_clip3(0, maxv, m5 + dbk_clip3(-tc2, tc2, ((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1)));
    }
  }
}
FHC_DBK_HD int dbk_strong(const int16_t* p, int off, int dd, int beta, int tc)           // xUseStrongFiltering :933
{
  const int m4 = p[0], m3 = p[-off], m7 = p[3 * off], m0 = p[-4 * off];
  return (dbk_abs(m0 - m3) + dbk_abs(m7 - m4) < (beta >> 3)) && (dd < (beta >> 2)) && (dbk_abs(m3 - m4) < ((tc * 5 + 1) >> 1));
}
FHC_DBK_HD int dbk_dp(const int16_t* p, int off) { return dbk_abs(p[-3 * off] - 2 * p[-2 * off] + p[-off]); }
FHC_DBK_HD int dbk_dq(const int16_t* p, int off) { return dbk_abs(p[0] - 2 * p[off] + p[2 * off]); }

// One edge segment: the 4 luma lines whose first Q-side sample is (x, y) and, on the 16-sample grid, the 2 chroma lines of
// both chroma planes (xEdgeFilterLuma :557, xEdgeFilterChroma :676).  (x, y) on the 8x8 grid in the edge's direction,
// on the 4-sample grid along it.
FHC_DBK_HD void dbk_segment(const DbkTables& t, const DbkPic& pic, int W, int wc, int bd, int x, int y, int dir)
{
  if (!dbk_tu_edge(pic.ctus, wc, x, y, dir)) return;
  const int scale = 1 << (bd - 8), maxv = (1 << bd) - 1;
  {
    const int stride = W, off = dir == 0 ? 1 : stride, step = dir == 0 ? stride : 1;
    int16_t* s = pic.plane[0] + (size_t)y * stride + x;
    const int tc = t.tc[dbk_clip3(0, 53, pic.qp + 2)] * scale, beta = t.beta[dbk_clip3(0, 51, pic.qp)] * scale;
    const int side = (beta + (beta >> 1)) >> 3, thrCut = tc * 10;
    const int dp0 = dbk_dp(s, off), dq0 = dbk_dq(s, off), dp3 = dbk_dp(s + 3 * step, off), dq3 = dbk_dq(s + 3 * step, off);
    const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
    if (d0 + d3 < beta)
    {
      const int sw = dbk_strong(s, off, 2 * d0, beta, tc) && dbk_strong(s + 3 * step, off, 2 * d3, beta, tc);
      for (int i = 0; i < 4; i++) dbk_pel_luma(s + i * step, off, tc, sw, thrCut, dp < side, dq < side, maxv);
    }
  }
  if (((dir == 0 ? x : y) & 15) == 0)
  {
    int qp = pic.qp;                                                  // ((QpP + QpQ + 1) >> 1) + cQpPicOffset (0)
    if (qp >= 58) qp -= 6; else if (qp >= 0) qp = t.chromaScale[qp];
    const int tc = t.tc[dbk_clip3(0, 53, qp + 2)] * scale;
    const int stride = W >> 1, off = dir == 0 ? 1 : stride, step = dir == 0 ? stride : 1;
    for (int c = 1; c < 3; c++)
    {
      int16_t* s = pic.plane[c] + (size_t)(y >> 1) * stride + (x >> 1);
      for (int i = 0; i < 2; i++)                                     // xPelFilterChroma :901
      {
        int16_t* p = s + i * step;
        const int m4 = p[0], m3 = p[-off], m5 = p[off], m2 = p[-2 * off];
        const int delta = dbk_clip3(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
        p[-off] = (int16_t)dbk_clip3(0, maxv, m3 + delta);
        p[0] = (int16_t)dbk_clip3(0, maxv, m4 - delta);
      }
    }
  }
}

// The copy passes move groups of four neighbouring samples (8 bytes of int16): picture widths are multiples of 8, so a
// group never straddles a CTU, a row or a plane, and every access is 8-byte aligned (records are 8-byte aligned, a
// record row is 128 / 64 bytes).  Group g of the planar 4:2:0 picture (Y, Cb, Cr) -> component and position.
struct DbkGroup { int comp, x, y, w; };
FHC_DBK_HD DbkGroup dbk_group(int W, int H, int g)
{
  const int lumaGroups = W * H / 4, chromaGroups = lumaGroups / 4;
  DbkGroup r;
  r.comp = g < lumaGroups ? 0 : (g < lumaGroups + chromaGroups ? 1 : 2);
  const int k = g - (r.comp == 0 ? 0 : (r.comp == 1 ? lumaGroups : lumaGroups + chromaGroups));
  r.w = r.comp ? W >> 1 : W;
  const int gw = r.w >> 2;
  r.y = k / gw; r.x = (k - r.y * gw) << 2;
  return r;
}
struct DbkQuad { uint32_t lo, hi; };                 // four int16 samples in two words (little endian: sample 0 in the low half of lo)
FHC_DBK_HD DbkQuad dbk_load4(const int16_t* p)
{
  DbkQuad q;
#if defined(__CUDA_ARCH__)
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  q.lo = v.x; q.hi = v.y;
#else
  q.lo = (uint32_t)(uint16_t)p[0] | ((uint32_t)(uint16_t)p[1] << 16);
  q.hi = (uint32_t)(uint16_t)p[2] | ((uint32_t)(uint16_t)p[3] << 16);
#endif
  return q;
}
FHC_DBK_HD void dbk_store4(int16_t* p, const DbkQuad& q)
{
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<uint2*>(p) = make_uint2(q.lo, q.hi);
#else
  p[0] = (int16_t)(q.lo & 0xFFFF); p[1] = (int16_t)(q.lo >> 16); p[2] = (int16_t)(q.hi & 0xFFFF); p[3] = (int16_t)(q.hi >> 16);
#endif
}
// pass 0: the records' reconstruction (64x64 / 32x32 blocks per CTU) -> planes
FHC_DBK_HD void dbk_gather4(const DbkPic& pic, int W, int H, int wc, int g)
{
  const DbkGroup r = dbk_group(W, H, g);
  const int n = r.comp ? 32 : 64, lg = r.comp ? 5 : 6;
  const fhc_ctu_result* o = &pic.ctus[(r.y >> lg) * wc + (r.x >> lg)];
  const int16_t* src = r.comp == 0 ? o->recoY : (r.comp == 1 ? o->recoCb : o->recoCr);
  dbk_store4(pic.plane[r.comp] + (size_t)r.y * r.w + r.x, dbk_load4(src + (r.y & (n - 1)) * n + (r.x & (n - 1))));
}
// pass 3: planes -> packed output in the layout of the input pictures (4 or 8 bytes per group)
FHC_DBK_HD void dbk_pack4(const DbkPic& pic, int W, int H, int sampleBytes, int g)
{
  const DbkGroup r = dbk_group(W, H, g);
  const DbkQuad q = dbk_load4(pic.plane[r.comp] + (size_t)r.y * r.w + r.x);
  if (sampleBytes == 1)
  {
    const uint32_t packed = (q.lo & 255u) | ((q.lo >> 8) & 0xFF00u) | ((q.hi & 255u) << 16) | ((q.hi << 8) & 0xFF000000u);
    static_cast<uint32_t*>(pic.out)[g] = packed;                  // little endian: sample k at byte k
  }
  else dbk_store4(static_cast<int16_t*>(pic.out) + (size_t)g * 4, q);
}

#if defined(__CUDACC__) && defined(FHC_DEVICE_BUILD)
__constant__ DbkTables g_dbkTables = FHC_DBK_TABLES;
// 128 threads x <= 64 registers = 8 K registers per block: fits on an SM next to a block of the persistent mode-decision
// kernel (320 threads x 168 registers = 53.8 K of 64 K), so finished pictures are filtered while later ones are searched.
#define FHC_DBK_THREADS 128

// pass 0 (blockIdx.y = picture): a thread per group of four samples, consecutive threads on consecutive groups
__global__ void __launch_bounds__(FHC_DBK_THREADS) fhc_dbk_gather_kernel(const DbkPic* pics, int W, int H, int wc)
{
  const DbkPic& pic = pics[blockIdx.y];
  const int total = W * H * 3 / 8;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) dbk_gather4(pic, W, H, wc, g);
}
// passes 1 / 2: one thread per candidate edge segment; neighbouring threads take neighbouring segments along the edge
// direction's perpendicular so that a warp's accesses stay within a few rows.
__global__ void __launch_bounds__(FHC_DBK_THREADS) fhc_dbk_edges_kernel(const DbkPic* pics, int W, int H, int wc, int bd, int dir)
{
  const DbkPic& pic = pics[blockIdx.y];
  const int nx = dir == 0 ? W / 8 : W / 4, ny = dir == 0 ? H / 4 : H / 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx * ny; i += gridDim.x * blockDim.x)
  {
    const int sx = i % nx, sy = i / nx;
    dbk_segment(g_dbkTables, pic, W, wc, bd, dir == 0 ? sx * 8 : sx * 4, dir == 0 ? sy * 4 : sy * 8, dir);
  }
}
// pass 3
__global__ void __launch_bounds__(FHC_DBK_THREADS) fhc_dbk_pack_kernel(const DbkPic* pics, int W, int H, int sampleBytes)
{
  const DbkPic& pic = pics[blockIdx.y];
  const int total = W * H * 3 / 8;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) dbk_pack4(pic, W, H, sampleBytes, g);
}
#endif

}  // namespace fhc
#endif
