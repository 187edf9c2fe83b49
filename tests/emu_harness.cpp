// TEST INFRASTRUCTURE: host "emulation build" of the kernel source.
// Compiles fasthevc_b200/csrc/fhc_engine.cuh for the CPU (FHC_DEVICE_BUILD undefined): one virtual lane per warp, lane
// groups and role warps executed one after the other in dependency order (a job runs when the chain master joins it,
// i.e. after its sub-tree).  This checks the engine's logic (buffer layouts, decision code, syntax walkers, the
// concurrency assumptions between depths) against the oracle without a GPU.  It is built only by the tests into
// tests/_build/ and is never part of libfasthevc_cuda.so.
#include <stdlib.h>
#include <vector>
#include "../fasthevc_b200/csrc/fhc_engine.cuh"

extern "C" int fhc_emu_compress_picture_hinted(int width, int height, int bitDepth, int sampleBytes, const fhc_slice_params* sp,
                                               const void* y, const void* cb, const void* cr, fhc_ctu_result* out, const uint8_t* hints);
extern "C" int fhc_emu_compress_picture_ex(int width, int height, int bitDepth, int sampleBytes, const fhc_slice_params* sp,
                                           const void* y, const void* cb, const void* cr, fhc_ctu_result* out, const uint8_t* hints,
                                           int wpp, int order);
extern "C" int fhc_emu_compress_picture(int width, int height, int bitDepth, int sampleBytes, const fhc_slice_params* sp,
                                        const void* y, const void* cb, const void* cr, fhc_ctu_result* out)
{
  return fhc_emu_compress_picture_ex(width, height, bitDepth, sampleBytes, sp, y, cb, cr, out, nullptr, 0, 0);
}
extern "C" int fhc_emu_compress_picture_hinted(int width, int height, int bitDepth, int sampleBytes, const fhc_slice_params* sp,
                                               const void* y, const void* cb, const void* cr, fhc_ctu_result* out, const uint8_t* hints)
{
  return fhc_emu_compress_picture_ex(width, height, bitDepth, sampleBytes, sp, y, cb, cr, out, hints, 0, 0);
}
// wpp = 1: wavefront picture (PicDesc::rows).  order: 0 = row after row; 1 / 2 = the tightest interleaving the device may
// produce - every row exactly two CTUs behind the row above, one CTU per call, each row on a block of its own - with
// the rows of a time step visited top-down (1) or bottom-up (2).  The engine aborts if a call breaks the wavefront lag.
extern "C" int fhc_emu_compress_picture_ex(int width, int height, int bitDepth, int sampleBytes, const fhc_slice_params* sp,
                                           const void* y, const void* cb, const void* cr, fhc_ctu_result* out, const uint8_t* hints,
                                           int wpp, int order)
{
  using namespace fhc;
  size_t smemBytes, globalBytes;
  block_footprint(&smemBytes, &globalBytes);
  PicDesc pd;
  memset(&pd, 0, sizeof(pd));
  pd.org[0] = y; pd.org[1] = cb; pd.org[2] = cr;
  pd.out = out; pd.sp = *sp; pd.width = width; pd.height = height; pd.bitDepth = bitDepth; pd.sampleBytes = sampleBytes;
  pd.wc = (width + 63) / 64; pd.hc = (height + 63) / 64;
  pd.hints = hints;
  // blocks that take the CTU rows of the picture in turn, as different SMs may on the device: nothing but PicState /
  // RowState, the line buffers and the results may carry over from one row to the next
  const int nblk = (wpp && order) ? pd.hc : 2;
  std::vector<unsigned char*> sm((size_t)nblk), gm((size_t)nblk);
  std::vector<Blk*> blk((size_t)nblk);
  for (int k = 0; k < nblk; k++)
  {
    sm[k] = (unsigned char*)calloc(1, smemBytes);
    gm[k] = (unsigned char*)calloc(1, globalBytes);
    blk[k] = (Blk*)sm[k];
    build_tables(&blk[k]->t, bitDepth);
    init_block(blk[k], sm[k] + ((sizeof(Blk) + 15) & ~(size_t)15), gm[k]);
    blk[k]->inst = 0; blk[k]->twin = 0;
  }
  std::vector<int16_t> lines((size_t)width * 4);
  std::vector<uint8_t> depthRows((size_t)width);
  for (int k = 0; k < 2; k++)
  {
    pd.line[k][0] = lines.data() + (size_t)k * 2 * width;
    pd.line[k][1] = pd.line[k][0] + width;
    pd.line[k][2] = pd.line[k][1] + width / 2;
    pd.depthRow[k] = depthRows.data() + (size_t)k * (width / 2);
  }
  PicState state;
  memset(&state, 0, sizeof(state));
  pd.state = &state;
  std::vector<RowState> rows((size_t)pd.hc);
  memset(rows.data(), 0, sizeof(RowState) * rows.size());
  if (wpp) pd.rows = rows.data();
  if (!(wpp && order))
    for (int row = 0; row < pd.hc; row++) compress_picture(blk[row & 1], &pd, row * pd.wc, pd.wc, 0, 1, wpp);
  else
  {
    const int lag = pd.wc >= 2 ? 2 : 1;
    for (int t = 0; t < pd.wc + lag * (pd.hc - 1); t++)
      for (int k = 0; k < pd.hc; k++)
      {
        const int row = order == 1 ? k : pd.hc - 1 - k, x = t - lag * row;
        if (x >= 0 && x < pd.wc) compress_picture(blk[row], &pd, row * pd.wc + x, 1, 0, 1, wpp);
      }
  }
  for (int k = 0; k < nblk; k++) { free(sm[k]); free(gm[k]); }
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
// twin_align calls per role since the last reset (roles 3A, 3B, 2, 1, 0)
extern "C" void fhc_emu_align_counts(unsigned long long* out5, int reset)
{
  for (int k = 0; k < 5; k++) out5[k] = fhc::g_emuAlignCalls[k];
  if (reset) memset(fhc::g_emuAlignCalls, 0, sizeof(fhc::g_emuAlignCalls));
}
extern "C" int fhc_emu_shared_bytes() { size_t s, g; fhc::block_footprint(&s, &g); return (int)s; }
extern "C" int fhc_emu_global_bytes() { size_t s, g; fhc::block_footprint(&s, &g); return (int)g; }
