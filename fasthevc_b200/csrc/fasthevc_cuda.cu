// fasthevc_cuda.cu - kernel entry + C ABI (include/fasthevc_cuda.h) of libfasthevc_cuda.so.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -shared
// (-fmad=false: every RD cost is IEEE double in the reference's operation order).
#define FHC_DEVICE_BUILD 1
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "fhc_tables.h"

using namespace fhc;
#ifndef FHC_MIN_BLOCKS
#define FHC_MIN_BLOCKS 1
#endif
extern "C" int fhc_debug_profile(fhc_ctx* c, int slot, long long* cycles16, long long* counts16);

// ----------------------------------------------------------------------------------------------
// kernel: one block per picture; thread 0 = master, lanes 1..31 of warp 0 retire, warps 1.. = workers
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 + FHC_NWORK, FHC_MIN_BLOCKS) fhc_compress_pictures_kernel(FrameGlobal* frames, const Tables* tables)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared* s = reinterpret_cast<Shared*>(smem_raw);
  FrameGlobal* g = frames + blockIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s->t);
    for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) init_engines(g, s);
  __syncthreads();
  if (threadIdx.x >= 1 && threadIdx.x < 32) return;      // exited threads are not counted by later barriers
  Ctx c;
  c.g = g; c.s = s; c.e = &s->master;
  if (threadIdx.x == 0)
  {
    master_compress_picture(c);
    s->master.cmd.op = OP_EXIT;
    bar_all();
  }
  else
  {
    const int w = threadIdx.x - 32;
    for (;;)
    {
      bar_all();
      const int op = s->master.cmd.op;
      if (op == OP_EXIT) break;
      if (op == OP_TASKS)
      {
        // every worker warp is an engine: lane 0 walks the candidate's decision code, lanes 1..31 run its sample phases
        const int wi = w >> 5, lane = w & 31;
        for (int t = wi; t < s->task.numTasks; t += FHC_NENG)
        {
          Ctx ec;
          ec.g = g; ec.s = s; ec.e = &s->eng[wi];
          if (lane == 0) { run_task(g, s, wi, t); engine_lanes_release(ec); }
          else engine_lanes_loop(ec, lane);
        }
        __syncwarp();
      }
      else
      {
        const int np = (op == OP_RMD) ? s->master.cmd.a[15] : num_phases(op);
        for (int ph = 0; ph < np; ph++)
        {
          run_phase(c, op, ph, w, FHC_NWORK);
          if (ph + 1 < np) bar_workers();
        }
      }
      bar_all();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static uint64_t g_launches = 0;

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      snprintf(g_err, sizeof(g_err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return FHC_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)

struct fhc_ctx
{
  fhc_seq_params seq;
  int device, numCtus, maxFrames;
  size_t lumaSamples, frameSamples;       // per picture
  FrameGlobal* dFrames;                   // [maxFrames]
  Tables* dTables;
  uint8_t* dIn;                           // maxFrames * frameSamples * 2 bytes
  int16_t* dRec;                          // maxFrames * frameSamples
  fhc_ctu_result* dOut;                   // maxFrames * numCtus
  cudaStream_t stream;
  cudaEvent_t evStart, evStop;
  int firstPoc, nSubmitted, pending;
  size_t headerBytes;
  std::vector<unsigned char> hostHeaders;
};

const char* fhc_last_error(void) { return g_err; }
uint64_t fhc_launch_count(void) { return g_launches; }
int fhc_num_ctus(const fhc_ctx* ctx) { return ctx ? ctx->numCtus : FHC_ERR_ARG; }

int fhc_slice_params_from_qp(int qp, fhc_slice_params* out)
{
  if (!out || qp < 0 || qp > 51) return FHC_ERR_ARG;
  slice_params_from_qp(qp, out);
  return FHC_OK;
}

int fhc_create(const fhc_seq_params* seq, int device, fhc_ctx** out)
{
  if (!seq || !out) return FHC_ERR_ARG;
  if (seq->width <= 0 || seq->height <= 0 || (seq->width & 7) || (seq->height & 7) || (seq->bitDepth != 8 && seq->bitDepth != 10) ||
      seq->chromaFormat != 1 || seq->ctuSize != 64 || seq->maxCuDepth != 4 || seq->tuLog2Max != 5 || seq->tuLog2Min != 2 ||
      seq->tuMaxDepthIntra != 3 || !seq->useRdoq || !seq->useRdoqTs || !seq->useTransformSkip || !seq->useTransformSkipFast ||
      !seq->signHideFlag || !seq->strongIntraSmoothing || !seq->fastUdiUseMpm || seq->maxFramesInFlight <= 0)
  {
    snprintf(g_err, sizeof(g_err), "fhc_create: unsupported configuration (only the encoder_intra_main[10].cfg tool set is implemented)");
    return FHC_ERR_ARG;
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { snprintf(g_err, sizeof(g_err), "fhc_create: no CUDA device %d", device); return FHC_ERR_CUDA; }
  CK(cudaSetDevice(device));
  fhc_ctx* c = new fhc_ctx();
  c->seq = *seq; c->device = device;
  c->numCtus = ((seq->width + 63) / 64) * ((seq->height + 63) / 64);
  c->maxFrames = seq->maxFramesInFlight;
  c->lumaSamples = (size_t)seq->width * seq->height;
  c->frameSamples = c->lumaSamples * 3 / 2;
  c->firstPoc = 0; c->nSubmitted = 0; c->pending = 0;
  c->headerBytes = offsetof(FrameGlobal, cuCoef);
  c->hostHeaders.resize(c->headerBytes * (size_t)c->maxFrames);
  CK(cudaMalloc(&c->dFrames, sizeof(FrameGlobal) * (size_t)c->maxFrames));
  CK(cudaMalloc(&c->dTables, sizeof(Tables)));
  CK(cudaMalloc(&c->dIn, c->frameSamples * 2 * (size_t)c->maxFrames));
  CK(cudaMalloc(&c->dRec, c->frameSamples * sizeof(int16_t) * (size_t)c->maxFrames));
  CK(cudaMalloc(&c->dOut, sizeof(fhc_ctu_result) * (size_t)c->numCtus * (size_t)c->maxFrames));
  CK(cudaStreamCreate(&c->stream));
  CK(cudaEventCreate(&c->evStart));
  CK(cudaEventCreate(&c->evStop));
  Tables t;
  build_tables(&t, seq->bitDepth);
  CK(cudaMemcpy(c->dTables, &t, sizeof(t), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(fhc_compress_pictures_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared)));
  CK(cudaDeviceSetLimit(cudaLimitStackSize, 16384));
  *out = c;
  return FHC_OK;
}

int fhc_destroy(fhc_ctx* c)
{
  if (!c) return FHC_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->dFrames); cudaFree(c->dTables); cudaFree(c->dIn); cudaFree(c->dRec); cudaFree(c->dOut);
  cudaEventDestroy(c->evStart); cudaEventDestroy(c->evStop);
  cudaStreamDestroy(c->stream);
  delete c;
  return FHC_OK;
}

// fills the per-picture headers and launches the kernel on c->stream
static int launch(fhc_ctx* c, int n, int sampleBytes, const fhc_slice_params* sp)
{
  for (int i = 0; i < n; i++)
  {
    const uint8_t* in = c->dIn + (size_t)i * c->frameSamples * (size_t)sampleBytes;
    int16_t* rec = c->dRec + (size_t)i * c->frameSamples;
    unsigned char* hb = c->hostHeaders.data() + c->headerBytes * (size_t)i;   // header prefix of FrameGlobal only
    const void* org[3] = { in, in + c->lumaSamples * (size_t)sampleBytes, in + (c->lumaSamples + c->lumaSamples / 4) * (size_t)sampleBytes };
    int16_t* recp[3] = { rec, rec + c->lumaSamples, rec + c->lumaSamples + c->lumaSamples / 4 };
    fhc_ctu_result* outp = c->dOut + (size_t)i * c->numCtus;
    memcpy(hb + offsetof(FrameGlobal, org), org, sizeof(org));
    memcpy(hb + offsetof(FrameGlobal, rec), recp, sizeof(recp));
    memcpy(hb + offsetof(FrameGlobal, out), &outp, sizeof(outp));
    memcpy(hb + offsetof(FrameGlobal, sp), &sp[i], sizeof(fhc_slice_params));
    int32_t v[6] = { c->seq.width, c->seq.height, c->seq.bitDepth, sampleBytes, (c->seq.width + 63) / 64, (c->seq.height + 63) / 64 };
    memcpy(hb + offsetof(FrameGlobal, width), v, sizeof(v));
  }
  CK(cudaMemcpy2DAsync(c->dFrames, sizeof(FrameGlobal), c->hostHeaders.data(), c->headerBytes, c->headerBytes, (size_t)n,
                       cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemset2DAsync(reinterpret_cast<unsigned char*>(c->dFrames) + offsetof(FrameGlobal, prof), sizeof(FrameGlobal), 0, 64 * sizeof(long long), (size_t)n, c->stream));
  CK(cudaEventRecord(c->evStart, c->stream));
  fhc_compress_pictures_kernel<<<n, 32 + FHC_NWORK, sizeof(Shared), c->stream>>>(c->dFrames, c->dTables);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->evStop, c->stream));
  g_launches++;
  return FHC_OK;
}

int fhc_submit_frames(fhc_ctx* c, int first_poc, int n, const void* yuv, int sampleBytes, const fhc_slice_params* per_frame)
{
  if (!c || !yuv || !per_frame || n <= 0 || n > c->maxFrames || (sampleBytes != 1 && sampleBytes != 2)) return FHC_ERR_ARG;
  if (c->seq.bitDepth > 8 && sampleBytes != 2) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  if (c->pending) CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpyAsync(c->dIn, yuv, c->frameSamples * (size_t)sampleBytes * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  int rc = launch(c, n, sampleBytes, per_frame);
  if (rc != FHC_OK) return rc;
  c->firstPoc = first_poc; c->nSubmitted = n; c->pending = 1;
  return FHC_OK;
}

static int wait_done(fhc_ctx* c)
{
  if (c->pending)
  {
    CK(cudaStreamSynchronize(c->stream));
    c->pending = 0;
  }
  return FHC_OK;
}

int fhc_fetch_ctu(fhc_ctx* c, int poc, int ctuRsAddr, fhc_ctu_result* out)
{
  if (!c || !out) return FHC_ERR_ARG;
  if (c->nSubmitted == 0) return FHC_ERR_STATE;
  const int slot = poc - c->firstPoc;
  if (slot < 0 || slot >= c->nSubmitted || ctuRsAddr < 0 || ctuRsAddr >= c->numCtus) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  int rc = wait_done(c);
  if (rc != FHC_OK) return rc;
  CK(cudaMemcpy(out, c->dOut + (size_t)slot * c->numCtus + ctuRsAddr, sizeof(fhc_ctu_result), cudaMemcpyDeviceToHost));
  return FHC_OK;
}

int fhc_fetch_picture(fhc_ctx* c, int poc, fhc_ctu_result* out)
{
  if (!c || !out) return FHC_ERR_ARG;
  if (c->nSubmitted == 0) return FHC_ERR_STATE;
  const int slot = poc - c->firstPoc;
  if (slot < 0 || slot >= c->nSubmitted) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  int rc = wait_done(c);
  if (rc != FHC_OK) return rc;
  CK(cudaMemcpy(out, c->dOut + (size_t)slot * c->numCtus, sizeof(fhc_ctu_result) * (size_t)c->numCtus, cudaMemcpyDeviceToHost));
  return FHC_OK;
}

int fhc_device_input(fhc_ctx* c, void** dptr, uint64_t* bytes)
{
  if (!c || !dptr) return FHC_ERR_ARG;
  *dptr = c->dIn;
  if (bytes) *bytes = (uint64_t)c->frameSamples * 2 * (uint64_t)c->maxFrames;
  return FHC_OK;
}

int fhc_run_resident(fhc_ctx* c, int n, int sampleBytes, const fhc_slice_params* per_frame, float* kernel_ms)
{
  if (!c || !per_frame || n <= 0 || n > c->maxFrames || (sampleBytes != 1 && sampleBytes != 2)) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  if (c->pending) CK(cudaStreamSynchronize(c->stream));
  int rc = launch(c, n, sampleBytes, per_frame);
  if (rc != FHC_OK) return rc;
  CK(cudaStreamSynchronize(c->stream));
  c->firstPoc = 0; c->nSubmitted = n; c->pending = 0;
  if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, c->evStart, c->evStop));
  return FHC_OK;
}

int fhc_host_alloc(void** p, uint64_t bytes)
{
  if (!p) return FHC_ERR_ARG;
  CK(cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault));
  return FHC_OK;
}
int fhc_host_free(void* p)
{
  CK(cudaFreeHost(p));
  return FHC_OK;
}

// debug: cycle counters of picture `slot` (FHC_PROFILE builds; zeros otherwise)
int fhc_debug_profile(fhc_ctx* c, int slot, long long* cycles16, long long* counts16)
{
  if (!c || slot < 0 || slot >= c->maxFrames) return FHC_ERR_ARG;
  CK(cudaMemcpy(cycles16, reinterpret_cast<unsigned char*>(c->dFrames + slot) + offsetof(FrameGlobal, prof), 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(counts16, reinterpret_cast<unsigned char*>(c->dFrames + slot) + offsetof(FrameGlobal, profN), 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return FHC_OK;
}
extern "C" int fhc_debug_profile2(fhc_ctx* c, int slot, long long* cycles16, long long* counts16)
{
  if (!c || slot < 0 || slot >= c->maxFrames) return FHC_ERR_ARG;
  CK(cudaMemcpy(cycles16, reinterpret_cast<unsigned char*>(c->dFrames + slot) + offsetof(FrameGlobal, prof2), 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(counts16, reinterpret_cast<unsigned char*>(c->dFrames + slot) + offsetof(FrameGlobal, prof2N), 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return FHC_OK;
}
