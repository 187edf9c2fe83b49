// TEST / MEASUREMENT INFRASTRUCTURE: a libfasthevc_cuda.so that serves precomputed CTU records from a file
// (FHC_REPLAY=<file>: [pictures][numCtus] fhc_ctu_result, picture k for every POC with POC % pictures == k) without any
// computation.  Linked in place of the CUDA library (LD_LIBRARY_PATH) it turns the integrated encoder
// (oracle/_ref/TAppEncoderFHC) into "the unchanged host with an infinitely fast device": its run time is the host's
// per-picture tail alone (tools/host_tail.py).  Never shipped.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../include/fasthevc_cuda.h"
#include "../fasthevc_b200/csrc/fhc_deblock.cuh"
#include "../fasthevc_b200/csrc/fhc_sao.cuh"

#include <map>
struct fhc_ctx { fhc_seq_params seq; int numCtus; std::vector<fhc_ctu_result> recs; int pictures; std::vector<int> qps; int firstPoc; std::map<int, std::vector<unsigned char> > org; };

extern "C" {
const char* fhc_last_error(void) { return "replay ABI"; }
int fhc_create(const fhc_seq_params* seq, int, fhc_ctx** out)
{
  fhc_ctx* c = new fhc_ctx();
  c->seq = *seq;
  c->numCtus = ((seq->width + 63) / 64) * ((seq->height + 63) / 64);
  const char* p = getenv("FHC_REPLAY");
  FILE* f = p ? fopen(p, "rb") : NULL;
  if (!f) { delete c; return FHC_ERR_STATE; }
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  c->pictures = (int)(bytes / (long)(sizeof(fhc_ctu_result) * c->numCtus));
  c->recs.resize((size_t)c->pictures * c->numCtus);
  if (fread(c->recs.data(), sizeof(fhc_ctu_result), c->recs.size(), f) != c->recs.size()) { fclose(f); delete c; return FHC_ERR_STATE; }
  fclose(f);
  c->firstPoc = 0;
  *out = c;
  return FHC_OK;
}
int fhc_destroy(fhc_ctx* c) { delete c; return FHC_OK; }
int fhc_pictures_in_flight(int) { return 296; }
int fhc_configure(fhc_ctx*, int, int) { return FHC_OK; }
int fhc_submit_frames_hinted(fhc_ctx* c, int first_poc, int n, const void* yuv, int sb, const fhc_slice_params* sp, const uint8_t*)
{
  const size_t frame = (size_t)c->seq.width * c->seq.height * 3 / 2 * sb;
  for (int k = 0; k < n; k++) c->org[first_poc + k].assign((const unsigned char*)yuv + frame * k, (const unsigned char*)yuv + frame * (k + 1));
  while (c->org.size() > 64) c->org.erase(c->org.begin());

  if ((int)c->qps.size() < first_poc + n) c->qps.resize((size_t)(first_poc + n), sp[0].sliceQp);
  for (int k = 0; k < n; k++) c->qps[(size_t)(first_poc + k)] = sp[k].sliceQp;
  return FHC_OK;
}
int fhc_submit_frames(fhc_ctx* c, int first_poc, int n, const void* yuv, int sb, const fhc_slice_params* sp) { return fhc_submit_frames_hinted(c, first_poc, n, yuv, sb, sp, NULL); }
int fhc_fetch_picture(fhc_ctx* c, int poc, fhc_ctu_result* out)
{
  memcpy(out, c->recs.data() + (size_t)(poc % c->pictures) * c->numCtus, sizeof(fhc_ctu_result) * (size_t)c->numCtus);
  return FHC_OK;
}
int fhc_fetch_ctu(fhc_ctx* c, int poc, int addr, fhc_ctu_result* out) { *out = c->recs[(size_t)(poc % c->pictures) * c->numCtus + addr]; return FHC_OK; }
// the deblocking passes of the kernel source on the host (its time is printed, so that it can be taken out of the tail)
int fhc_deblock_pictures_stats(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes, fhc_sao_stats* saoStats);
int fhc_deblock_pictures(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes) { return fhc_deblock_pictures_stats(c, first_poc, n, out, sampleBytes, NULL); }
int fhc_deblock_pictures_stats(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes, fhc_sao_stats* saoStats)
{
  using namespace fhc;
  static const DbkTables tables = FHC_DBK_TABLES;
  const int W = c->seq.width, H = c->seq.height, wc = (W + 63) / 64;
  const long long luma = (long long)W * H, frame = luma * 3 / 2;
  std::vector<int16_t> planes((size_t)frame);
  for (int i = 0; i < n; i++)
  {
    DbkPic pic;
    pic.ctus = c->recs.data() + (size_t)((first_poc + i) % c->pictures) * c->numCtus;
    pic.plane[0] = planes.data(); pic.plane[1] = planes.data() + luma; pic.plane[2] = planes.data() + luma + luma / 4;
    pic.qp = c->qps[(size_t)(first_poc + i)]; pic.pad = 0;
    pic.out = (unsigned char*)out + (size_t)i * frame * sampleBytes;
    for (int g = 0; g < (int)(frame / 4); g++) dbk_gather4(pic, W, H, wc, g);
    for (int dir = 0; dir < 2; dir++)
      for (int y = 0; y < H; y += dir == 0 ? 4 : 8)
        for (int x = 0; x < W; x += dir == 0 ? 8 : 4) dbk_segment(tables, pic, W, wc, c->seq.bitDepth, x, y, dir);
    for (int g = 0; g < (int)(frame / 4); g++) dbk_pack4(pic, W, H, sampleBytes, g);
    if (saoStats && getenv("FHC_REPLAY_FAKE_SAO")) memset(saoStats + (size_t)i * c->numCtus * 3, 0, sizeof(fhc_sao_stats) * (size_t)c->numCtus * 3);     // timing only: statistics at no cost
    else if (saoStats)
    {
      const int inSb = c->seq.bitDepth > 8 ? 2 : 1;
      const unsigned char* in = c->org[first_poc + i].data();
      const int16_t* pl[3] = { pic.plane[0], pic.plane[1], pic.plane[2] };
      const void* org[3] = { in, in + luma * inSb, in + (luma + luma / 4) * inSb };
      sao_stats_host(W, H, c->seq.bitDepth, pl, org, inSb, saoStats + (size_t)i * c->numCtus * 3);
    }
  }
  return FHC_OK;
}
}
