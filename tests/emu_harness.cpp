// TEST INFRASTRUCTURE: host "emulation build" of the kernel source.
// Compiles fasthevc_b200/csrc/fhc_core.cuh for the CPU (FHC_DEVICE_BUILD undefined): the warp-uniform decision code runs
// as an ordinary function, every COOP phase is executed for all 32 virtual lanes in turn and every LANE task runs on
// its own lane engine.  This checks the kernel's logic (striping of the phases, buffer layouts, decision code)
// against the oracle without a GPU.  It is built only by the tests into tests/_build/ and is never part of
// libfasthevc_cuda.so.
#include <stdlib.h>
#include "../fasthevc_b200/csrc/fhc_tables.h"

extern "C" int fhc_emu_compress_picture(int width, int height, int bitDepth, int sampleBytes, const fhc_slice_params* sp,
                                        const void* y, const void* cb, const void* cr, fhc_ctu_result* out)
{
  using namespace fhc;
  FrameGlobal* g = (FrameGlobal*)calloc(1, sizeof(FrameGlobal));
  Shared* s = (Shared*)calloc(1, sizeof(Shared));
  Tables* t = (Tables*)calloc(1, sizeof(Tables));
  build_tables(t, bitDepth);
  g->org[0] = y; g->org[1] = cb; g->org[2] = cr;
  g->width = width; g->height = height; g->bitDepth = bitDepth; g->sampleBytes = sampleBytes;
  g->wc = (width + 63) / 64; g->hc = (height + 63) / 64;
  g->sp = *sp;
  g->out = out;
  for (int c = 0; c < 3; c++) g->rec[c] = (int16_t*)calloc((size_t)(width * height) >> (c ? 2 : 0), sizeof(int16_t));
  init_engines(g, s, t);
  Ctx ctx = { g, s, &s->master, 0, FHC_GROUP, 0xffffffffu };
  PicDesc pd;
  memset(&pd, 0, sizeof(pd));
  for (int c = 0; c < 3; c++) { pd.org[c] = g->org[c]; pd.rec[c] = g->rec[c]; }
  pd.out = out; pd.sp = *sp; pd.width = width; pd.height = height; pd.bitDepth = bitDepth; pd.sampleBytes = sampleBytes;
  pd.wc = g->wc; pd.hc = g->hc;
  load_picture(ctx, &pd);
  compress_picture(ctx);
  for (int c = 0; c < 3; c++) free(g->rec[c]);
  free(g); free(s); free(t);
  return 0;
}
// work counters since the last reset: out[0..7] = TU-pipeline samples [luma,chroma][4x4,8x8,16x16,32x32], out[8..15] = TU codings
// (same order), out[16] = samples of the rough mode decision, out[17] = CABAC bins counted
extern "C" void fhc_emu_counters(unsigned long long* out, int reset)
{
  using namespace fhc;
  for (int c = 0; c < 2; c++)
    for (int k = 0; k < 4; k++) { out[c * 4 + k] = g_emuCounters.tqSamples[c][k]; out[8 + c * 4 + k] = g_emuCounters.tuCalls[c][k]; }
  out[16] = g_emuCounters.rmdSamples; out[17] = g_emuCounters.bins;
  if (reset) memset(&g_emuCounters, 0, sizeof(g_emuCounters));
}
extern "C" int fhc_emu_shared_bytes() { return (int)sizeof(fhc::Shared); }
