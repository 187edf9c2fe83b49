// fhc_sao.cuh - SAO statistics of deblocked pictures on the device (second stage of the per-picture
// tail): TEncSampleAdaptiveOffset::getStatistics / getBlkStats (source/Lib/TLibEncoder/TEncSampleAdaptiveOffset.cpp:287-332,
// :911-1290) for the encoder's default set-up (statistics from the deblocked picture; skip lines 5 / 4 luma, 3 / 2
// chroma, :124-131).  More than half of the SAO encoder's time on the host (about 40 of 70 ms per 1080p picture,
// profiles/r02_host_tail_cpu.txt); the RD decision that consumes the numbers (decideBlkParams :782) stays there.  Included by fasthevc_cuda.cu and by the test-only host builds.
//
// HM walks a CTU line by line with sign line buffers; the result is a pure function of the samples, so here every
// sample decides for itself: for each of the five types, is it counted in this CTU (sao_member: the start / end columns
// and rows of getBlkStats, incl. its first-line rules) and which class does it fall in (sao_class: the two neighbours of
// the edge type, or the band).  Neighbours of counted samples always lie inside the picture.
#ifndef FHC_SAO_CUH
#define FHC_SAO_CUH
#include <stdint.h>
#include "../../include/fasthevc_cuda.h"

#if defined(__CUDACC__) && defined(FHC_DEVICE_BUILD)
#define FHC_SAO_HD __host__ __device__ __forceinline__
#else
#define FHC_SAO_HD static inline
#endif

namespace fhc {

FHC_SAO_HD int sao_sgn(int v) { return (v > 0) - (v < 0); }

// Is sample (x, y) of a CTU block of `w` x `h` samples counted for `type` (0..3 = EO 0 / 90 / 135 / 45 degrees, 4 = BO)?
// L / R / A / B: the CTU has a left / right / above / below neighbour inside the picture; chroma: the chroma planes.
FHC_SAO_HD int sao_member(int type, int x, int y, int w, int h, int chroma, int L, int R, int A, int B)
{
  const int skipR = chroma ? 3 : 5, skipB = chroma ? 2 : 4;
  const int x0 = L ? 0 : 1, x1 = R ? w - skipR : w - 1;              // EO 0 / 135 / 45 (:950, :1076, :1160)
  const int xe = R ? w - skipR : w;                                   // EO 90, BO (:1008, :1246)
  switch (type)
  {
  case 0: return x >= x0 && x < x1 && y < (B ? h - skipB : h);
  case 1: return x < xe && y >= (A ? 0 : 1) && y < (B ? h - skipB : h - 1);
  case 2:
  case 3: return x >= x0 && x < x1 && (y == 0 ? A : y < (B ? h - skipB : h - 1));      // first line only below another CTU
  default: return x < xe && y < (B ? h - skipB : h);
  }
}
// class of picture sample (px, py) of a plane with `stride`: EO -> 0..4 (= 2 + sign sum), BO -> band 0..31
FHC_SAO_HD int sao_class(int type, const int16_t* plane, int stride, int px, int py, int bitDepth)
{
  const int16_t* p = plane + (size_t)py * stride + px;
  const int c = p[0];
  switch (type)
  {
  case 0: return 2 + sao_sgn(c - p[-1]) + sao_sgn(c - p[1]);
  case 1: return 2 + sao_sgn(c - p[-stride]) + sao_sgn(c - p[stride]);
  case 2: return 2 + sao_sgn(c - p[-stride - 1]) + sao_sgn(c - p[stride + 1]);
  case 3: return 2 + sao_sgn(c - p[-stride + 1]) + sao_sgn(c - p[stride - 1]);
  default: return c >> (bitDepth - 5);
  }
}
// original sample (px, py) of component plane `org` stored with sampleBytes per sample
FHC_SAO_HD int sao_org(const void* org, int sampleBytes, int stride, int px, int py)
{
  const size_t i = (size_t)py * stride + px;
  return sampleBytes == 1 ? (int)static_cast<const uint8_t*>(org)[i] : (int)static_cast<const uint16_t*>(org)[i];
}

// Geometry of CTU `addr`, component `comp`: block position / size in the component's plane and the availability flags
struct SaoBlk { int x0, y0, w, h, L, R, A, B, stride; };
FHC_SAO_HD SaoBlk sao_block(int W, int H, int addr, int comp)
{
  const int wc = (W + 63) / 64, yPos = (addr / wc) * 64, xPos = (addr % wc) * 64, sh = comp ? 1 : 0;
  SaoBlk b;
  b.x0 = xPos >> sh; b.y0 = yPos >> sh;
  b.w = (xPos + 64 > W ? W - xPos : 64) >> sh; b.h = (yPos + 64 > H ? H - yPos : 64) >> sh;
  b.L = xPos > 0; b.A = yPos > 0; b.R = xPos + 64 < W; b.B = yPos + 64 < H;
  b.stride = W >> sh;
  return b;
}

#if defined(__CUDACC__) && defined(FHC_DEVICE_BUILD)
struct SaoPic { const int16_t* plane[3]; const void* org[3]; fhc_sao_stats* stats; };     // one picture of a call

// One block per (CTU, component) of a picture (blockIdx.z): counters in shared memory, one atomic per counted sample and
// type; 128 threads so that the block fits on an SM beside the persistent mode-decision kernel.
__global__ void __launch_bounds__(128) fhc_sao_stats_kernel(const SaoPic* pics, int W, int H, int bitDepth, int sampleBytes)
{
  __shared__ int sDiff[4 * 5 + 32], sCount[4 * 5 + 32];
  const SaoPic& pic = pics[blockIdx.z];
  const int addr = blockIdx.x, comp = blockIdx.y;
  const SaoBlk b = sao_block(W, H, addr, comp);
  for (int i = threadIdx.x; i < 52; i += blockDim.x) { sDiff[i] = 0; sCount[i] = 0; }
  __syncthreads();
  const int16_t* plane = pic.plane[comp];
  for (int i = threadIdx.x; i < b.w * b.h; i += blockDim.x)
  {
    const int y = i / b.w, x = i - y * b.w, px = b.x0 + x, py = b.y0 + y;
    const int d = sao_org(pic.org[comp], sampleBytes, b.stride, px, py) - plane[(size_t)py * b.stride + px];
#pragma unroll
    for (int type = 0; type < 5; type++)
      if (sao_member(type, x, y, b.w, b.h, comp != 0, b.L, b.R, b.A, b.B))
      {
        const int k = type * 5 + sao_class(type, plane, b.stride, px, py, bitDepth);
        atomicAdd(&sDiff[k], d);
        atomicAdd(&sCount[k], 1);
      }
  }
  __syncthreads();
  fhc_sao_stats* out = &pic.stats[addr * 3 + comp];
  for (int i = threadIdx.x; i < 52; i += blockDim.x)
  {
    if (i < 20) { out->eoDiff[i / 5][i % 5] = sDiff[i]; out->eoCount[i / 5][i % 5] = sCount[i]; }
    else { out->boDiff[i - 20] = sDiff[i]; out->boCount[i - 20] = sCount[i]; }
  }
}
#endif

// host form (the emulated ABI, measurement aids): the same membership / class functions, sample by sample
static inline void sao_stats_host(int W, int H, int bitDepth, const int16_t* const plane[3], const void* const org[3], int sampleBytes, fhc_sao_stats* stats)
{
  const int numCtus = ((W + 63) / 64) * ((H + 63) / 64);
  for (int addr = 0; addr < numCtus; addr++)
    for (int comp = 0; comp < 3; comp++)
    {
      const SaoBlk b = sao_block(W, H, addr, comp);
      fhc_sao_stats* out = &stats[addr * 3 + comp];
      for (int t = 0; t < 4; t++) for (int k = 0; k < 5; k++) { out->eoDiff[t][k] = 0; out->eoCount[t][k] = 0; }
      for (int k = 0; k < 32; k++) { out->boDiff[k] = 0; out->boCount[k] = 0; }
      for (int y = 0; y < b.h; y++)
        for (int x = 0; x < b.w; x++)
        {
          const int px = b.x0 + x, py = b.y0 + y;
          const int d = sao_org(org[comp], sampleBytes, b.stride, px, py) - plane[comp][(size_t)py * b.stride + px];
          for (int type = 0; type < 5; type++)
            if (sao_member(type, x, y, b.w, b.h, comp != 0, b.L, b.R, b.A, b.B))
            {
              const int k = sao_class(type, plane[comp], b.stride, px, py, bitDepth);
              if (type < 4) { out->eoDiff[type][k] += d; out->eoCount[type][k]++; }
              else { out->boDiff[k] += d; out->boCount[k]++; }
            }
        }
    }
}

}  // namespace fhc
#endif
