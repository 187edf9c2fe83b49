// TEncCuCUDA.cpp - the TLibEncoderCUDA shim: host side of the drop-in boundary, written against the reference's own
// classes (HM 16.14 as shipped in omricarmi/FastHEVC).  It is what INTEGRATION.md asks a maintainer to add to
// source/Lib/TLibEncoderCUDA/: TEncCu::compressCtu (source/Lib/TLibEncoder/TEncCu.cpp:252-288) keeps its call
// surface, the one call to xCompressCU(depth 0) (TEncCu.cpp:260) becomes fhc_shim_compress_ctu(this, pCtu), and
// TEncTop / TEncGOP / TEncSlice, the cfg options and the bitstream writer stay as they are.
//
// What it does, per CTU:
//   * first CTU of a picture: packs the original planes (TComPicYuv, Pel = int16 with margins) into the planar buffer
//     of fhc_submit_frames and passes the doubles the unchanged host already computed (TEncSlice::setUpLambda,
//     TEncSlice.cpp:112-144) bit for bit;
//   * every CTU: fhc_fetch_ctu, then leaves behind exactly what xCompressCU leaves behind - the picture-level
//     TComDataCU as TComDataCU::copyToPic fills it (TComDataCU.cpp:906-979) and the reconstruction in PicYuvRec as
//     TEncCu::xCopyYuv2Pic does (TEncCu.cpp:1729).  The caller re-codes the CTU on RD[0][CI_CURR_BEST] itself
//     (TEncSlice.cpp:886-893).
// With FHC_SHIM_INPUT set it reads ahead (see lookahead()) so that the device encodes many pictures at once.
//
// Error behaviour mirrors HM: print + exit (TEncSbac.cpp:1220-1235).
//
// Built only by oracle/build_ref.py against /root/reference (no reference source enters this repository); the build
// reads two private members of TEncCu, which a maintainer would expose with a friend declaration instead.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>
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
#include "TLibEncoder/TEncCu.h"
#include "TLibCommon/TComLoopFilter.h"
#undef private
#undef protected
#include "fasthevc_cuda.h"

static fhc_ctx* g_ctx = NULL;
static int g_width = 0, g_height = 0, g_bitDepth = 0, g_maxFrames = 0;
// The pictures the device holds: two batches, as the library keeps two submits (fhc_submit_frames while the previous
// submit is still being consumed).  While the host loop works through batch k - fetching its CTUs and running the
// unchanged per-picture tail (deblocking, SAO, final CABAC, MD5: TEncGOP.cpp:1619-1870) - the device already searches
// batch k+1, so the encoder's wall time is bounded by the slower of the two, not by their sum.
struct ShimBatch
{
  int first, n;                                       // POC range
  fhc_slice_params sp;
  std::vector<unsigned char> yuv;                     // the submitted pictures, one after the other
  ShimBatch() : first(-1), n(0) {}
};
static ShimBatch g_batch[2];
static int g_newest = 1;                              // index of the batch submitted last
static std::vector<unsigned char> g_cur;              // the picture the host is working on, packed the same way

// FHC_SHIM_TIMING=1: wall time spent inside the shim's entry points, printed at exit (where the host side of the
// boundary spends its time; the device calls are included in the entry that makes them)
#include <chrono>
struct ShimTimer
{
  double submit, fetchCopy, filter; bool on;
  ShimTimer() : submit(0), fetchCopy(0), filter(0), on(getenv("FHC_SHIM_TIMING") != NULL) {}
  ~ShimTimer() { if (on) fprintf(stderr, "TLibEncoderCUDA timing: submit (pack / compare / read ahead / fhc_submit) %.3f s, compressCtu (fetch + copy into TComDataCU) %.3f s, loop filter + SAO statistics %.3f s\n", submit, fetchCopy, filter); }
};
static ShimTimer g_timer;
struct ShimScope
{
  double* acc; std::chrono::steady_clock::time_point t0;
  explicit ShimScope(double* a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~ShimScope() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

static void fhc_die(const char* what)
{
  fprintf(stderr, "TLibEncoderCUDA: %s failed: %s\n", what, fhc_last_error());
  exit(1);
}

// Read-ahead.  All-intra pictures are independent, so the device can encode pictures the host loop has not reached yet.
// The unchanged host reads the input one frame at a time (TAppEncTop::encode), so the shim reads the following frames
// itself from the same file (FHC_SHIM_INPUT, planar 4:2:0 in the internal bit depth, no frame skip) and submits up to
// FHC_SHIM_LOOKAHEAD pictures (default fhc_pictures_in_flight) with the slice parameters of the first one - in the
// all-intra cfg every picture has the same QP and lambdas.  Nothing is assumed, though: when the host reaches a
// picture, its original planes and its slice parameters are compared with what was submitted, and the picture is
// re-submitted alone if anything differs.  Without FHC_SHIM_INPUT the shim submits one picture at a time.
static int lookahead()
{
  static int n = -1;
  if (n < 0)
  {
    n = 1;
    if (getenv("FHC_SHIM_INPUT"))
    {
      const char* e = getenv("FHC_SHIM_LOOKAHEAD");
      n = e ? atoi(e) : fhc_pictures_in_flight(0);
      if (n < 1) n = 1;
    }
  }
  return n;
}

static void pack_picture(const TComPicYuv* org, int sampleBytes, std::vector<unsigned char>& out)
{
  out.resize((size_t)org->getWidth(COMPONENT_Y) * org->getHeight(COMPONENT_Y) * 3 / 2 * sampleBytes);
  unsigned char* dst = out.data();
  for (int c = 0; c < 3; c++)
  {
    const ComponentID id = ComponentID(c);
    const int cw = org->getWidth(id), ch = org->getHeight(id), stride = org->getStride(id);
    const Pel* src = org->getAddr(id);
    for (int y = 0; y < ch; y++, src += stride)
    {
      if (sampleBytes == 1) { for (int x = 0; x < cw; x++) dst[x] = (unsigned char)src[x]; dst += cw; }      // row loops the compiler vectorises
      else { uint16_t* d16 = reinterpret_cast<uint16_t*>(dst); for (int x = 0; x < cw; x++) d16[x] = (uint16_t)src[x]; dst += 2 * cw; }
    }
  }
}

// Submits pictures [first, first + n): the first one from `head` when given (the picture the host just handed over),
// the others straight from the input file; a short read just shortens the batch.
static void submit_batch(int first, int n, const std::vector<unsigned char>* head, int sampleBytes, size_t frameBytes, const fhc_slice_params& sp, int w, int h)
{
  ShimBatch& bt = g_batch[g_newest ^ 1];
  bt.n = 0;
  bt.yuv.resize(frameBytes * (size_t)n);
  int got = 0;
  if (head) { memcpy(bt.yuv.data(), head->data(), frameBytes); got = 1; }
  if (got < n)
  {
    FILE* fp = fopen(getenv("FHC_SHIM_INPUT"), "rb");
    if (fp && fseek(fp, (long)((size_t)(first + got) * frameBytes), SEEK_SET) == 0)
      while (got < n && fread(bt.yuv.data() + (size_t)got * frameBytes, 1, frameBytes, fp) == frameBytes) got++;
    if (fp) fclose(fp);
  }
  if (got == 0) return;
  n = got;
  std::vector<fhc_slice_params> sps((size_t)n, sp);
  // CU-split hints of the CNN predictor (fhc_submit_frames_hinted): a file of [POC][numCtus][4] bytes named by
  // FHC_HINTS - what fhc_cnn_predict + the scatter of INTEGRATION.md produce for the sequence; pictures beyond the file
  // are searched in full
  static std::vector<uint8_t> allHints;
  static bool hintsLoaded = false;
  if (!hintsLoaded)
  {
    hintsLoaded = true;
    if (const char* hp = getenv("FHC_HINTS"))
      if (FILE* hf = fopen(hp, "rb"))
      {
        int ch;
        while ((ch = fgetc(hf)) != EOF) allHints.push_back((uint8_t)ch);
        fclose(hf);
      }
  }
  const size_t perPic = 4 * (size_t)((w + 63) / 64) * (size_t)((h + 63) / 64);
  std::vector<uint8_t> hints;
  if (!allHints.empty())
  {
    hints.assign(perPic * (size_t)n, FHC_HINT_NONE);
    for (size_t i = 0; i < hints.size(); i++)
      if ((size_t)first * perPic + i < allHints.size()) hints[i] = allHints[(size_t)first * perPic + i];
  }
  if (fhc_submit_frames_hinted(g_ctx, first, n, bt.yuv.data(), sampleBytes, sps.data(), hints.empty() ? NULL : hints.data()) != FHC_OK)
    fhc_die("fhc_submit_frames");
  bt.first = first; bt.n = n; bt.sp = sp;
  g_newest ^= 1;
}

static void submit_pictures(TEncCu* enc, TComDataCU* pCtu)
{
  TComPic* pic = pCtu->getPic();
  TComSlice* slice = pCtu->getSlice();
  const TComPicYuv* org = pic->getPicYuvOrg();
  const int w = org->getWidth(COMPONENT_Y), h = org->getHeight(COMPONENT_Y);
  const int bd = slice->getSPS()->getBitDepth(CHANNEL_TYPE_LUMA);
  const int sampleBytes = bd > 8 ? 2 : 1, poc = slice->getPOC();
  const size_t frameBytes = (size_t)w * h * 3 / 2 * sampleBytes;
  fhc_slice_params sp;
  memset(&sp, 0, sizeof(sp));
  sp.sliceQp = slice->getSliceQp();
  sp.lambda = enc->m_pcRdCost->getLambda();
  sp.sqrtLambda = enc->m_pcRdCost->getSqrtLambda();
  sp.chromaDistWeight[0] = enc->m_pcRdCost->m_distortionWeight[COMPONENT_Cb];
  sp.chromaDistWeight[1] = enc->m_pcRdCost->m_distortionWeight[COMPONENT_Cr];
  for (int c = 0; c < 3; c++) sp.rdoqLambda[c] = enc->m_pcTrQuant->m_lambdas[c];
  pack_picture(org, sampleBytes, g_cur);

  // (FHC_POC_BASE: this instance encodes a frame group that starts in the middle of the sequence, see build_ref.py)
  const int total = enc->m_pcEncCfg->getFramesToBeEncoded() + (getenv("FHC_POC_BASE") ? atoi(getenv("FHC_POC_BASE")) : 0);
  int n = lookahead();
  // already on the device, with these samples and these parameters?
  ShimBatch* have = NULL;
  for (int k = 0; k < 2; k++)
  {
    ShimBatch& bt = g_batch[k];
    if (g_ctx && bt.n > 0 && poc >= bt.first && poc < bt.first + bt.n && w == g_width && h == g_height && bd == g_bitDepth &&
        memcmp(&sp, &bt.sp, sizeof(sp)) == 0 && memcmp(g_cur.data(), bt.yuv.data() + (size_t)(poc - bt.first) * frameBytes, frameBytes) == 0)
      have = &bt;
  }
  if (have)
  {
    // first picture of a batch: queue the batch after it, so that the device works while the host runs this batch's tail
    const int next = have->first + have->n;
    if (n > g_maxFrames) n = g_maxFrames;
    if (n > 1 && poc == have->first && next < total && g_batch[0].first != next && g_batch[1].first != next)
      submit_batch(next, n < total - next ? n : total - next, NULL, sampleBytes, frameBytes, sp, w, h);
    return;
  }

  const int left = total - poc;
  if (n > left) n = left;
  if (n < 1) n = 1;
  if (g_ctx && (w != g_width || h != g_height || bd != g_bitDepth || n > g_maxFrames)) { fhc_destroy(g_ctx); g_ctx = NULL; g_batch[0].n = g_batch[1].n = 0; }
  if (!g_ctx)
  {
    fhc_seq_params p;
    memset(&p, 0, sizeof(p));
    p.width = w; p.height = h; p.bitDepth = bd; p.chromaFormat = 1;
    p.ctuSize = 64; p.maxCuDepth = 4; p.tuLog2Max = 5; p.tuLog2Min = 2; p.tuMaxDepthIntra = 3;
    p.useRdoq = p.useRdoqTs = p.useTransformSkip = p.useTransformSkipFast = 1;
    p.signHideFlag = p.strongIntraSmoothing = p.fastUdiUseMpm = 1;
    p.maxFramesInFlight = n;
    p.waveFrontSynchro = enc->m_pcEncCfg->getEntropyCodingSyncEnabledFlag() ? 1 : 0;
    if (fhc_create(&p, 0, &g_ctx) != FHC_OK) fhc_die("fhc_create");
    if (n > 1 && fhc_configure(g_ctx, FHC_OPT_PIPELINED_SUBMITS, 1) != FHC_OK) fhc_die("fhc_configure");
    g_width = w; g_height = h; g_bitDepth = bd; g_maxFrames = n;
  }
  // a stale batch that claims this POC (wrong samples / parameters) must not be found again
  for (int k = 0; k < 2; k++) if (g_batch[k].n > 0 && poc >= g_batch[k].first && poc < g_batch[k].first + g_batch[k].n) g_batch[k].n = poc - g_batch[k].first;
  submit_batch(poc, n, &g_cur, sampleBytes, frameBytes, sp, w, h);
  // ... and the batch after it right away: two submits in flight from the start
  const ShimBatch& cur = g_batch[g_newest];
  const int next = cur.first + cur.n;
  if (n > 1 && cur.n == n && next < total) submit_batch(next, n < total - next ? n : total - next, NULL, sampleBytes, frameBytes, sp, w, h);
}

void fhc_shim_compress_ctu(TEncCu* enc, TComDataCU* pCtu)
{
  // one device->host copy per picture (fhc_fetch_picture waits for that picture only), served CTU by CTU
  static std::vector<fhc_ctu_result> picture;
  const int addr = (int)pCtu->getCtuRsAddr();
  ShimScope scope(&g_timer.fetchCopy);
  if (addr == 0)
  {
    { ShimScope s2(&g_timer.submit); submit_pictures(enc, pCtu); }
    picture.resize((size_t)pCtu->getPic()->getNumberOfCtusInFrame());
    if (fhc_fetch_picture(g_ctx, pCtu->getSlice()->getPOC(), picture.data()) != FHC_OK) fhc_die("fhc_fetch_picture");
  }
  const fhc_ctu_result& r = picture[(size_t)addr];

  // TComDataCU::copyToPic: the per-partition arrays an intra CTU decision changes (everything else keeps the values
  // TComDataCU::initCtu gave it just before compressCtu, TEncSlice.cpp:797)
  {
    UChar* depth = pCtu->getDepth(); UChar* wd = pCtu->getWidth(); UChar* ht = pCtu->getHeight();
    SChar* part = pCtu->getPartitionSize(); SChar* pred = pCtu->getPredictionMode();
    for (unsigned i = 0; i < 256; i++)
    {
      depth[i] = r.depth[i];
      wd[i] = ht[i] = (UChar)(64 >> r.depth[i]);
      part[i] = (SChar)r.partSize[i];
      pred[i] = (SChar)r.predMode[i];
    }
    memcpy(pCtu->getIntraDir(CHANNEL_TYPE_LUMA), r.lumaDir, 256);
    memcpy(pCtu->getIntraDir(CHANNEL_TYPE_CHROMA), r.chromaDir, 256);
    memcpy(pCtu->getTransformIdx(), r.trIdx, 256);
    for (unsigned c = 0; c < 3; c++)
    {
      memcpy(pCtu->getCbf(ComponentID(c)), r.cbf[c], 256);
      memcpy(pCtu->getTransformSkip(ComponentID(c)), r.transformSkip[c], 256);
    }
    TCoeff* cy = pCtu->getCoeff(COMPONENT_Y); TCoeff* cb = pCtu->getCoeff(COMPONENT_Cb); TCoeff* cr = pCtu->getCoeff(COMPONENT_Cr);
    for (unsigned i = 0; i < 4096; i++) cy[i] = r.coefY[i];
    for (unsigned i = 0; i < 1024; i++) { cb[i] = r.coefCb[i]; cr[i] = r.coefCr[i]; }
  }
  pCtu->getTotalCost() = r.totalCost;
  pCtu->getTotalDistortion() = r.totalDist;
  pCtu->getTotalBits() = r.totalBits;
  pCtu->getTotalBins() = r.totalBins;

  // TEncCu::xCopyYuv2Pic: reconstruction of the CTU into the picture (samples inside the picture only)
  TComPicYuv* rec = pCtu->getPic()->getPicYuvRec();
  const int x0 = (int)pCtu->getCUPelX(), y0 = (int)pCtu->getCUPelY();
  for (int c = 0; c < 3; c++)
  {
    const ComponentID id = ComponentID(c);
    const int sh = c ? 1 : 0, n = 64 >> sh, stride = rec->getStride(id);
    const int cw = rec->getWidth(id), ch = rec->getHeight(id);
    Pel* dst = rec->getAddr(id, addr);
    const int16_t* src = c == 0 ? r.recoY : (c == 1 ? r.recoCb : r.recoCr);
    for (int y = 0; y < n && (y0 >> sh) + y < ch; y++)
      for (int x = 0; x < n && (x0 >> sh) + x < cw; x++) dst[y * stride + x] = src[y * n + x];
  }
}

// The deblocking filter of a finished picture (TEncGOP.cpp:1619 calls this instead of m_pcLoopFilter->loopFilterPic):
// fhc_deblock_pictures filters the picture on the device from the CTU records it already holds and returns the planes,
// which replace PicYuvRec.  FHC_SHIM_DEBLOCK=0 keeps the host filter.
static std::vector<fhc_sao_stats> g_saoStats;         // statistics of picture g_saoPoc, [ctu][component]
static int g_saoPoc = -1;

// TEncSampleAdaptiveOffset::SAOProcess calls this instead of getStatistics (TEncSampleAdaptiveOffset.cpp:253): fills
// m_statData ([ctu][component][type] {Int64 diff[32], count[32]}) from the device's statistics; returns false (the host
// computes them) when the device did not filter this picture.
bool fhc_shim_sao_statistics(void* statData, TComPic* pic)
{
  if (g_saoPoc != pic->getPOC()) return false;
  struct Stat { long long diff[32]; long long count[32]; };
  Stat*** s = (Stat***)statData;
  const int numCtus = (int)pic->getNumberOfCtusInFrame();
  for (int c = 0; c < numCtus; c++)
    for (int comp = 0; comp < 3; comp++)
    {
      const fhc_sao_stats& d = g_saoStats[(size_t)c * 3 + comp];
      for (int t = 0; t < 5; t++) memset(&s[c][comp][t], 0, sizeof(Stat));
      for (int t = 0; t < 4; t++)
        for (int k = 0; k < 5; k++) { s[c][comp][t].diff[k] = d.eoDiff[t][k]; s[c][comp][t].count[k] = d.eoCount[t][k]; }
      for (int k = 0; k < 32; k++) { s[c][comp][4].diff[k] = d.boDiff[k]; s[c][comp][4].count[k] = d.boCount[k]; }
    }
  g_saoPoc = -1;
  return true;
}

void fhc_shim_loop_filter(TComLoopFilter* lf, TComPic* pic)
{
  static const bool onDevice = !(getenv("FHC_SHIM_DEBLOCK") && atoi(getenv("FHC_SHIM_DEBLOCK")) == 0);
  ShimScope scope(&g_timer.filter);
  if (!onDevice || !g_ctx) { lf->loopFilterPic(pic); return; }
  TComPicYuv* rec = pic->getPicYuvRec();
  const int w = rec->getWidth(COMPONENT_Y), h = rec->getHeight(COMPONENT_Y);
  const int sampleBytes = g_bitDepth > 8 ? 2 : 1;
  static std::vector<unsigned char> planes;
  planes.resize((size_t)w * h * 3 / 2 * sampleBytes);
  // the SAO statistics of the deblocked picture come back with it (fhc_shim_sao_statistics hands them to SAOProcess)
  g_saoStats.resize((size_t)pic->getNumberOfCtusInFrame() * 3);
  g_saoPoc = -1;
  if (fhc_deblock_pictures_stats(g_ctx, pic->getPOC(), 1, planes.data(), sampleBytes, g_saoStats.data()) != FHC_OK) fhc_die("fhc_deblock_pictures_stats");
  g_saoPoc = pic->getPOC();
  size_t pos = 0;
  for (int c = 0; c < 3; c++)
  {
    const ComponentID id = ComponentID(c);
    const int cw = rec->getWidth(id), ch = rec->getHeight(id), stride = rec->getStride(id);
    Pel* dst = rec->getAddr(id);
    for (int y = 0; y < ch; y++, dst += stride)
    {
      if (sampleBytes == 1) { const unsigned char* p8 = planes.data() + pos; for (int x = 0; x < cw; x++) dst[x] = (Pel)p8[x]; pos += (size_t)cw; }
      else { const uint16_t* p16 = reinterpret_cast<const uint16_t*>(planes.data() + pos); for (int x = 0; x < cw; x++) dst[x] = (Pel)p16[x]; pos += 2 * (size_t)cw; }
    }
  }
}
