// TEST INFRASTRUCTURE (oracle side). Not part of the product path.
//
// Dump hook linked into the *instrumented* build of the reference encoder
// (oracle/_ref/TAppEncoderDump, produced by oracle/build_ref.py).  The build
// recipe inserts one call to fhc_dump_ctu() at the end of
// TEncCu::compressCtu (reference: source/Lib/TLibEncoder/TEncCu.cpp:252-288)
// and one call to fhc_dump_slice() before the CTU loop of
// TEncSlice::compressSlice (source/Lib/TLibEncoder/TEncSlice.cpp:792).
// When the environment variable FHC_DUMP names a file, every finished CTU is
// appended to it as one fixed-size binary record (layout: struct below, also
// parsed by tests/fhc_testlib.py).  This is how per-CTU golden vectors for the
// hot path are obtained from the unmodified reference arithmetic.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>
// private members (chroma distortion weights, RDOQ lambdas) are read for the dump only
#include <sstream>
#include <iostream>
#include <fstream>
#include <vector>
#include <list>
#include <map>
#include <string>
#include <algorithm>
#include <deque>
#include <bitset>
#include <iomanip>
#include <functional>
#define private public
#define protected public
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComSlice.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibCommon/TComTrQuant.h"

static FILE* fhc_fp()
{
  static FILE* fp = NULL;
  static int tried = 0;
  if (!tried)
  {
    tried = 1;
    const char* p = getenv("FHC_DUMP");
    if (p && *p) fp = fopen(p, "wb");
  }
  return fp;
}

#pragma pack(push, 1)
struct FhcSliceRec
{
  uint32_t magic;        // 'FHCS'
  int32_t  poc, qp, width, height, bitDepthY, bitDepthC;
  double   lambda, sqrtLambda, distWeightCb, distWeightCr, rdoqLambda[3];
};
struct FhcCtuRec
{
  uint32_t magic;        // 'FHCC'
  int32_t  poc, ctuRsAddr;
  uint8_t  depth[256], partSize[256], predMode[256], lumaDir[256], chromaDir[256], trIdx[256];
  uint8_t  cbf[3][256], tskip[3][256];
  int32_t  coefY[4096], coefCb[1024], coefCr[1024];
  int16_t  recoY[4096], recoCb[1024], recoCr[1024];
  double   totalCost;
  uint32_t totalDist, totalBits, totalBins;
};
#pragma pack(pop)

void fhc_dump_slice(TComPic* pic, TComSlice* slice, TComRdCost* rd, TComTrQuant* tq)
{
  FILE* fp = fhc_fp();
  if (!fp) return;
  FhcSliceRec r;
  memset(&r, 0, sizeof(r));
  r.magic = 0x53434846u;
  r.poc = slice->getPOC();
  r.qp = slice->getSliceQp();
  r.width = pic->getPicYuvOrg()->getWidth(COMPONENT_Y);
  r.height = pic->getPicYuvOrg()->getHeight(COMPONENT_Y);
  r.bitDepthY = slice->getSPS()->getBitDepth(CHANNEL_TYPE_LUMA);
  r.bitDepthC = slice->getSPS()->getBitDepth(CHANNEL_TYPE_CHROMA);
  r.lambda = rd->getLambda();
  r.sqrtLambda = rd->getSqrtLambda();
  r.distWeightCb = rd->m_distortionWeight[COMPONENT_Cb];
  r.distWeightCr = rd->m_distortionWeight[COMPONENT_Cr];
  for (int c = 0; c < 3; c++) r.rdoqLambda[c] = tq->m_lambdas[c];
  fwrite(&r, sizeof(r), 1, fp);
  fflush(fp);
}

void fhc_dump_ctu(TComDataCU* pCtu)
{
  FILE* fp = fhc_fp();
  if (!fp) return;
  static FhcCtuRec r;
  memset(&r, 0, sizeof(r));
  r.magic = 0x43434846u;
  r.poc = pCtu->getSlice()->getPOC();
  r.ctuRsAddr = pCtu->getCtuRsAddr();
  for (int i = 0; i < 256; i++)
  {
    r.depth[i] = pCtu->getDepth(i);
    r.partSize[i] = (uint8_t)pCtu->getPartitionSize(i);
    r.predMode[i] = (uint8_t)pCtu->getPredictionMode(i);
    r.lumaDir[i] = pCtu->getIntraDir(CHANNEL_TYPE_LUMA, i);
    r.chromaDir[i] = pCtu->getIntraDir(CHANNEL_TYPE_CHROMA, i);
    r.trIdx[i] = pCtu->getTransformIdx(i);
    for (int c = 0; c < 3; c++)
    {
      r.cbf[c][i] = pCtu->getCbf(i, ComponentID(c));
      r.tskip[c][i] = pCtu->getTransformSkip(i, ComponentID(c));
    }
  }
  memcpy(r.coefY, pCtu->getCoeff(COMPONENT_Y), sizeof(r.coefY));
  memcpy(r.coefCb, pCtu->getCoeff(COMPONENT_Cb), sizeof(r.coefCb));
  memcpy(r.coefCr, pCtu->getCoeff(COMPONENT_Cr), sizeof(r.coefCr));
  TComPicYuv* rec = pCtu->getPic()->getPicYuvRec();
  const int W = rec->getWidth(COMPONENT_Y), H = rec->getHeight(COMPONENT_Y);
  const int x0 = pCtu->getCUPelX(), y0 = pCtu->getCUPelY();
  for (int c = 0; c < 3; c++)
  {
    const ComponentID id = ComponentID(c);
    const int sh = c ? 1 : 0, n = 64 >> sh;
    const Pel* p = rec->getAddr(id, pCtu->getCtuRsAddr());
    const int stride = rec->getStride(id);
    int16_t* dst = c == 0 ? r.recoY : (c == 1 ? r.recoCb : r.recoCr);
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++)
        dst[y * n + x] = ((x0 >> sh) + x < (W >> sh) && (y0 >> sh) + y < (H >> sh)) ? p[y * stride + x] : 0;
  }
  r.totalCost = pCtu->getTotalCost();
  r.totalDist = pCtu->getTotalDistortion();
  r.totalBits = pCtu->getTotalBits();
  r.totalBins = pCtu->getTotalBins();
  fwrite(&r, sizeof(r), 1, fp);
  fflush(fp);
}

// CU-split hints for the hinted variant of the instrumented encoder (oracle/_ref/TAppEncoderDumpHinted): the file named
// by FHC_HINTS holds one byte per depth-1 CU, [POC][ctuRsAddr][4 in z-order]: 0 = none, 1 = do not split, 2 = split.
// The build recipe patches the two sites of TEncCu::xCompressCU this value steers (TEncCu.cpp:576 and :892); this is how
// the oracle's hinted mode is pinned against the reference's own arithmetic.
int fhc_hint(TComDataCU* cu)
{
  static std::vector<unsigned char> hints;
  static bool loaded = false;
  if (!loaded)
  {
    loaded = true;
    if (const char* p = getenv("FHC_HINTS"))
    {
      if (FILE* f = fopen(p, "rb"))
      {
        int ch;
        while ((ch = fgetc(f)) != EOF) hints.push_back((unsigned char)ch);
        fclose(f);
      }
    }
  }
  const size_t ctus = cu->getPic()->getNumberOfCtusInFrame();
  const size_t i = ((size_t)cu->getSlice()->getPOC() * ctus + cu->getCtuRsAddr()) * 4 + (cu->getZorderIdxInCtu() >> 6);
  return i < hints.size() ? hints[i] : 0;
}

// SAO statistics of a picture (TEncSampleAdaptiveOffset::getStatistics, TEncSampleAdaptiveOffset.cpp:287): the dump variant
// calls this right after getStatistics in SAOProcess (:253).  FHC_SAO_DUMP=<file>: per picture the raw
// [ctu][component][type] {Int64 diff[32]; Int64 count[32]} records, appended.
void fhc_dump_sao_stats(void* statData, int numCtus)
{
  const char* p = getenv("FHC_SAO_DUMP");
  if (!p) return;
  struct Stat { long long diff[32]; long long count[32]; };
  Stat*** s = (Stat***)statData;
  FILE* f = fopen(p, "ab");
  if (!f) return;
  for (int c = 0; c < numCtus; c++)
    for (int comp = 0; comp < 3; comp++) fwrite(s[c][comp], sizeof(Stat), 5, f);
  fclose(f);
}
