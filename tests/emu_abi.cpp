// TEST INFRASTRUCTURE: the subset of include/fasthevc_cuda.h that the TLibEncoderCUDA shim calls, implemented on the
// host emulation build of the kernel source (tests/emu_harness.cpp).  Lets the CPU suite run the reference encoder with
// the shim (oracle/_ref/TAppEncoderFHC) and compare its bitstream with the stock reference without a GPU.  Built into
// tests/_build/libfasthevc_cuda.so and selected with LD_LIBRARY_PATH by the test only; never shipped.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../include/fasthevc_cuda.h"
#include "../fasthevc_b200/csrc/fhc_deblock.cuh"
#include "../fasthevc_b200/csrc/fhc_sao.cuh"

extern "C" int fhc_emu_compress_picture_ex(int width, int height, int bitDepth, int sampleBytes, const fhc_slice_params* sp,
                                           const void* y, const void* cb, const void* cr, fhc_ctu_result* out, const uint8_t* hints,
                                           int wpp, int order);

struct emu_batch { int poc, n; std::vector<fhc_ctu_result> out; std::vector<int> qps; std::vector<unsigned char> yuv; emu_batch() : poc(-1), n(0) {} };
struct fhc_ctx { fhc_seq_params seq; int numCtus, newest, pipelined; emu_batch b[2]; };     // two submit slots, as in the library

// FHC_EMU_LOG=<file>: one line per submit / per picture's first fetch, so that a test can see the order of the shim's calls
static void emu_log(const char* what, int a, int b)
{
  if (const char* p = getenv("FHC_EMU_LOG")) if (FILE* f = fopen(p, "a")) { fprintf(f, "%s %d %d\n", what, a, b); fclose(f); }
}

extern "C" {
const char* fhc_last_error(void) { return "emulated ABI"; }
int fhc_create(const fhc_seq_params* seq, int, fhc_ctx** out)
{
  fhc_ctx* c = new fhc_ctx();
  c->seq = *seq; c->newest = 1; c->pipelined = 0;
  c->numCtus = ((seq->width + 63) / 64) * ((seq->height + 63) / 64);
  *out = c;
  return FHC_OK;
}
int fhc_destroy(fhc_ctx* c) { delete c; return FHC_OK; }
int fhc_pictures_in_flight(int) { return 3; }
int fhc_submit_frames_hinted(fhc_ctx* c, int first_poc, int n, const void* yuv, int sampleBytes, const fhc_slice_params* sp, const uint8_t* hints);
int fhc_submit_frames(fhc_ctx* c, int first_poc, int n, const void* yuv, int sampleBytes, const fhc_slice_params* sp)
{
  return fhc_submit_frames_hinted(c, first_poc, n, yuv, sampleBytes, sp, NULL);
}
int fhc_submit_frames_hinted(fhc_ctx* c, int first_poc, int n, const void* yuv, int sampleBytes, const fhc_slice_params* sp, const uint8_t* hints)
{
  if (n < 1 || n > c->seq.maxFramesInFlight) return FHC_ERR_ARG;
  const size_t luma = (size_t)c->seq.width * c->seq.height * sampleBytes, frame = luma + luma / 2;
  emu_log("submit", first_poc, n);
  emu_batch& bt = c->b[c->newest ^ 1];
  bt.out.assign((size_t)c->numCtus * n, fhc_ctu_result());
  bt.poc = first_poc; bt.n = n;
  bt.yuv.assign((const unsigned char*)yuv, (const unsigned char*)yuv + frame * n);
  bt.qps.resize((size_t)n);
  for (int k = 0; k < n; k++) bt.qps[(size_t)k] = sp[k].sliceQp;
  c->newest ^= 1;
  for (int k = 0; k < n; k++)
  {
    const unsigned char* p = (const unsigned char*)yuv + frame * k;
    int rc = fhc_emu_compress_picture_ex(c->seq.width, c->seq.height, c->seq.bitDepth, sampleBytes, &sp[k], p, p + luma, p + luma + luma / 4,
                                         bt.out.data() + (size_t)c->numCtus * k, hints ? hints + (size_t)4 * c->numCtus * k : NULL,
                                         c->seq.waveFrontSynchro, 2);
    if (rc) return rc;
  }
  return FHC_OK;
}
int fhc_fetch_ctu(fhc_ctx* c, int poc, int addr, fhc_ctu_result* out)
{
  if (addr < 0 || addr >= c->numCtus) return FHC_ERR_ARG;
  for (int k = 0; k < 2; k++)                       // the newer submit first, as the library does
  {
    const emu_batch& bt = c->b[k ? c->newest ^ 1 : c->newest];
    if (poc >= bt.poc && poc < bt.poc + bt.n) { *out = bt.out[(size_t)(poc - bt.poc) * c->numCtus + addr]; return FHC_OK; }
  }
  return FHC_ERR_ARG;
}
// the deblocking passes of fasthevc_b200/csrc/fhc_deblock.cuh and the SAO statistics of fhc_sao.cuh, run sample by sample on the host
int fhc_deblock_pictures_stats(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes, fhc_sao_stats* saoStats);
int fhc_deblock_pictures(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes) { return fhc_deblock_pictures_stats(c, first_poc, n, out, sampleBytes, NULL); }
int fhc_deblock_pictures_stats(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes, fhc_sao_stats* saoStats)
{
  using namespace fhc;
  static const DbkTables tables = FHC_DBK_TABLES;
  const int W = c->seq.width, H = c->seq.height, wc = (W + 63) / 64;
  const long long luma = (long long)W * H, frame = luma * 3 / 2;
  emu_log("deblock", first_poc, n);
  for (int k = 0; k < 2; k++)
  {
    const emu_batch& bt = c->b[k ? c->newest ^ 1 : c->newest];
    if (first_poc < bt.poc || first_poc + n > bt.poc + bt.n) continue;
    std::vector<int16_t> planes((size_t)frame);
    for (int i = 0; i < n; i++)
    {
      DbkPic pic;
      pic.ctus = bt.out.data() + (size_t)(first_poc - bt.poc + i) * c->numCtus;
      pic.plane[0] = planes.data(); pic.plane[1] = planes.data() + luma; pic.plane[2] = planes.data() + luma + luma / 4;
      pic.qp = bt.qps[(size_t)(first_poc - bt.poc + i)]; pic.pad = 0;
      for (int g = 0; g < (int)(frame / 4); g++) dbk_gather4(pic, W, H, wc, g);
      for (int dir = 0; dir < 2; dir++)
        for (int y = 0; y < H; y += dir == 0 ? 4 : 8)
          for (int x = 0; x < W; x += dir == 0 ? 8 : 4) dbk_segment(tables, pic, W, wc, c->seq.bitDepth, x, y, dir);
      pic.out = (unsigned char*)out + (size_t)i * frame * sampleBytes;
      for (int g = 0; g < (int)(frame / 4); g++) dbk_pack4(pic, W, H, sampleBytes, g);
      if (saoStats)
      {
        const int inSb = c->seq.bitDepth > 8 ? 2 : 1;
        const unsigned char* in = bt.yuv.data() + (size_t)(first_poc - bt.poc + i) * frame * inSb;
        const int16_t* pl[3] = { pic.plane[0], pic.plane[1], pic.plane[2] };
        const void* org[3] = { in, in + luma * inSb, in + (luma + luma / 4) * inSb };
        sao_stats_host(W, H, c->seq.bitDepth, pl, org, inSb, saoStats + (size_t)i * c->numCtus * 3);
      }
    }
    return FHC_OK;
  }
  return FHC_ERR_ARG;
}
int fhc_fetch_picture(fhc_ctx* c, int poc, fhc_ctu_result* out)
{
  emu_log("fetch", poc, c->pipelined);
  for (int k = 0; k < 2; k++)
  {
    const emu_batch& bt = c->b[k ? c->newest ^ 1 : c->newest];
    if (poc >= bt.poc && poc < bt.poc + bt.n) { memcpy(out, bt.out.data() + (size_t)(poc - bt.poc) * c->numCtus, sizeof(fhc_ctu_result) * (size_t)c->numCtus); return FHC_OK; }
  }
  return FHC_ERR_ARG;
}
int fhc_configure(fhc_ctx* c, int option, int value) { if (option != FHC_OPT_PIPELINED_SUBMITS) return FHC_ERR_ARG; c->pipelined = value; return FHC_OK; }
}
