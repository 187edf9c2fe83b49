// fasthevc_spool.cpp - libfasthevc_spool.so: the consumer side of the host pipeline.
//
// The unchanged host encoder is single-threaded, and its per-picture tail (RD re-coding of the CTUs, SAO decision, final
// CABAC, MD5) costs about as much per picture as the B200 needs for the whole mode decision.  All-intra pictures are
// independent and - with the cfg's CRA + ReWriteParamSetsFlag - every access unit is self-contained, so the tail is
// spread over host PROCESSES: one device process (fasthevc_b200/pipeline.py) runs the mode decision, the deblocking
// and the SAO statistics of all pictures on the GPU and writes one spool file per picture; N encoder processes
// (oracle/_ref/TAppEncoderFHC, each on a contiguous frame group: --FrameSkip k, FHC_POC_BASE=k) link this library in
// place of libfasthevc_cuda.so, read the device's results from the spool and run the unchanged tail; their bitstreams
// concatenate to the single-process bitstream byte for byte (tests/test_zpipeline.py).
//
// This library computes nothing: every entry point either hands out what the device wrote or fails.  It implements the
// subset of include/fasthevc_cuda.h that the shim (fasthevc_b200/shim/TEncCuCUDA.cpp) calls.
//
// Spool file of picture p: <FHC_SPOOL_DIR>/p<p, 8 digits>.fhc, written under a temporary name and renamed when complete:
//   SpoolHeader | fhc_slice_params | fhc_ctu_result[numCtus] | deblocked picture (planar, sampleBytes) | fhc_sao_stats[numCtus * 3]
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <string>
#include <vector>
#include "../../include/fasthevc_cuda.h"

struct SpoolHeader { uint32_t magic, width, height, bitDepth, numCtus, sampleBytes, poc, waveFrontSynchro; };
static const uint32_t kMagic = 0x4c505346u;     // "FSPL"

struct fhc_ctx
{
  fhc_seq_params seq;
  int numCtus;
  std::string dir;
  double timeoutS;
  // the picture last read
  int poc;
  fhc_slice_params sp;
  std::vector<fhc_ctu_result> recs;
  std::vector<unsigned char> planes;
  std::vector<fhc_sao_stats> stats;
  std::vector<std::pair<int, fhc_slice_params> > submitted;     // what the host says it would have submitted (checked against the file)
};

static thread_local char g_err[512] = "";
static void set_err(const char* fmt, const char* a, int b) { snprintf(g_err, sizeof(g_err), fmt, a, b); }

static int load_picture(fhc_ctx* c, int poc)
{
  if (c->poc == poc) return FHC_OK;
  char name[64];
  snprintf(name, sizeof(name), "/p%08d.fhc", poc);
  const std::string path = c->dir + name;
  // the device process may still be working on this picture: wait for the file (it appears complete: rename)
  const time_t t0 = time(NULL);
  FILE* f = NULL;
  while (!(f = fopen(path.c_str(), "rb")))
  {
    if (difftime(time(NULL), t0) > c->timeoutS) { set_err("spool: %s did not appear (errno %d)", path.c_str(), errno); return FHC_ERR_STATE; }
    struct timespec ts = {0, 2000000};
    nanosleep(&ts, NULL);
  }
  SpoolHeader h;
  const size_t frame = (size_t)c->seq.width * c->seq.height * 3 / 2;
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && h.magic == kMagic && (int)h.width == c->seq.width && (int)h.height == c->seq.height &&
            (int)h.bitDepth == c->seq.bitDepth && (int)h.numCtus == c->numCtus && (int)h.poc == poc && (int)h.waveFrontSynchro == c->seq.waveFrontSynchro &&
            (int)h.sampleBytes == (c->seq.bitDepth > 8 ? 2 : 1);
  if (ok)
  {
    c->recs.resize((size_t)c->numCtus);
    c->planes.resize(frame * h.sampleBytes);
    c->stats.resize((size_t)c->numCtus * 3);
    ok = fread(&c->sp, sizeof(c->sp), 1, f) == 1 && fread(c->recs.data(), sizeof(fhc_ctu_result), c->recs.size(), f) == c->recs.size() &&
         fread(c->planes.data(), 1, c->planes.size(), f) == c->planes.size() && fread(c->stats.data(), sizeof(fhc_sao_stats), c->stats.size(), f) == c->stats.size();
  }
  fclose(f);
  if (!ok) { set_err("spool: %s does not match this sequence (picture %d)", path.c_str(), poc); c->poc = -1; return FHC_ERR_STATE; }
  // the device process derived the picture's lambdas itself: they must be the ones this host computed
  for (size_t i = 0; i < c->submitted.size(); i++)
    if (c->submitted[i].first == poc && memcmp(&c->submitted[i].second, &c->sp, sizeof(fhc_slice_params)) != 0)
    { set_err("spool: %s was searched with other slice parameters than this host computed (picture %d)", path.c_str(), poc); c->poc = -1; return FHC_ERR_STATE; }
  c->poc = poc;
  return FHC_OK;
}

extern "C" {
const char* fhc_last_error(void) { return g_err; }
int fhc_create(const fhc_seq_params* seq, int, fhc_ctx** out)
{
  const char* dir = getenv("FHC_SPOOL_DIR");
  if (!seq || !out || !dir) { set_err("spool: FHC_SPOOL_DIR is not set%s (%d)", "", 0); return FHC_ERR_STATE; }
  fhc_ctx* c = new fhc_ctx();
  c->seq = *seq;
  c->numCtus = ((seq->width + 63) / 64) * ((seq->height + 63) / 64);
  c->dir = dir;
  c->timeoutS = getenv("FHC_SPOOL_TIMEOUT_S") ? atof(getenv("FHC_SPOOL_TIMEOUT_S")) : 600.0;
  c->poc = -1;
  *out = c;
  return FHC_OK;
}
int fhc_destroy(fhc_ctx* c) { delete c; return FHC_OK; }
int fhc_pictures_in_flight(int) { return 1; }
int fhc_configure(fhc_ctx*, int, int) { return FHC_OK; }
// the host's submit: nothing goes to a device from here; the slice parameters are kept for the check in load_picture
int fhc_submit_frames_hinted(fhc_ctx* c, int first_poc, int n, const void*, int, const fhc_slice_params* sp, const uint8_t* hints)
{
  if (!c || !sp || n <= 0) return FHC_ERR_ARG;
  if (hints) { set_err("spool: hinted submits are decided by the device process%s (%d)", "", 0); return FHC_ERR_ARG; }
  if (c->submitted.size() > 64) c->submitted.erase(c->submitted.begin(), c->submitted.begin() + 32);
  for (int k = 0; k < n; k++) c->submitted.push_back(std::make_pair(first_poc + k, sp[k]));
  return FHC_OK;
}
int fhc_submit_frames(fhc_ctx* c, int first_poc, int n, const void* yuv, int sb, const fhc_slice_params* sp) { return fhc_submit_frames_hinted(c, first_poc, n, yuv, sb, sp, NULL); }
int fhc_fetch_picture(fhc_ctx* c, int poc, fhc_ctu_result* out)
{
  if (!c || !out) return FHC_ERR_ARG;
  const int rc = load_picture(c, poc);
  if (rc != FHC_OK) return rc;
  memcpy(out, c->recs.data(), sizeof(fhc_ctu_result) * c->recs.size());
  return FHC_OK;
}
int fhc_fetch_ctu(fhc_ctx* c, int poc, int addr, fhc_ctu_result* out)
{
  if (!c || !out || addr < 0 || addr >= c->numCtus) return FHC_ERR_ARG;
  const int rc = load_picture(c, poc);
  if (rc != FHC_OK) return rc;
  *out = c->recs[(size_t)addr];
  return FHC_OK;
}
int fhc_deblock_pictures_stats(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes, fhc_sao_stats* stats)
{
  if (!c || !out || n != 1 || sampleBytes != (c->seq.bitDepth > 8 ? 2 : 1)) return FHC_ERR_ARG;
  const int rc = load_picture(c, first_poc);
  if (rc != FHC_OK) return rc;
  memcpy(out, c->planes.data(), c->planes.size());
  if (stats) memcpy(stats, c->stats.data(), sizeof(fhc_sao_stats) * c->stats.size());
  return FHC_OK;
}
int fhc_deblock_pictures(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes) { return fhc_deblock_pictures_stats(c, first_poc, n, out, sampleBytes, NULL); }
}
