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
#define FHC_BLOCK_THREADS (32 * FHC_WARPS_PER_BLOCK)
extern "C" int fhc_debug_profile(fhc_ctx* c, int slot, long long* cycles16, long long* counts16);

#if FHC_GROUP != 32
#error "the persistent kernel assumes one picture per warp (FHC_GROUP 32): sub-warp groups serialise, see DESIGN.md 4"
#endif
#define FHC_PICS_PER_WARP (32 / FHC_GROUP)
#define FHC_PICS_PER_BLOCK (FHC_PICS_PER_WARP * FHC_WARPS_PER_BLOCK)
static size_t tables_smem_bytes() { return (sizeof(Tables) + 15) & ~(size_t)15; }
static size_t kernel_smem_bytes() { return tables_smem_bytes() + sizeof(Shared) * FHC_PICS_PER_BLOCK + 16 * FHC_WARPS_PER_BLOCK; }

// ----------------------------------------------------------------------------------------------
// kernel: persistent warps; a warp encodes FHC_PICS_PER_WARP pictures side by side (FHC_GROUP lanes each), the
// groups of a warp start their pictures together so that they walk the same code at the same time
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FHC_BLOCK_THREADS, FHC_MIN_BLOCKS)
fhc_compress_pictures_kernel(PicDesc* pics, int numPics, int* queue, FrameGlobal* scratch, const Tables* tables)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tables* t = reinterpret_cast<Tables*>(smem_raw);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
    uint32_t* dst = reinterpret_cast<uint32_t*>(t);
    for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
  }
  int* rvz = reinterpret_cast<int*>(smem_raw + ((sizeof(Tables) + 15) & ~(size_t)15) + sizeof(Shared) * FHC_PICS_PER_BLOCK);
  if (threadIdx.x < 4 * FHC_WARPS_PER_BLOCK) rvz[threadIdx.x] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / FHC_GROUP, slot = warp * FHC_PICS_PER_WARP + grp;
  Shared* s = reinterpret_cast<Shared*>(smem_raw + ((sizeof(Tables) + 15) & ~(size_t)15)) + slot;
  FrameGlobal* g = scratch + (size_t)blockIdx.x * FHC_PICS_PER_BLOCK + slot;
  Ctx c;
  c.g = g; c.s = s; c.e = &s->master; c.lane = lane % FHC_GROUP; c.nl = FHC_GROUP;
  c.mask = (FHC_GROUP == 32) ? FHC_FULL_MASK : (((1u << FHC_GROUP) - 1u) << (grp * FHC_GROUP));
  if (c.lane == 0) init_engines(g, s, t);
  group_sync(c);
  // Rounds: the block pulls FHC_PICS_PER_BLOCK pictures at a time, so that its pictures start together and can be
  // kept aligned (block_align).  __syncthreads is safe here: every lane is back in this loop between rounds.
  for (;;)
  {
    __syncthreads();
    if (threadIdx.x == 0) rvz[0] = atomicAdd(queue, FHC_PICS_PER_BLOCK);
    __syncthreads();
    const int base = rvz[0];
    if (base >= numPics) break;
    const int active = (numPics - base < FHC_PICS_PER_BLOCK) ? numPics - base : FHC_PICS_PER_BLOCK;
    const int p = base + slot;
    if (p < numPics)
    {
      if (c.lane == 0)
      {
        // the group's leader walks the decision code; the other lanes serve its COOP sections and LANE dispatches
        s->alignThreads = (FHC_GROUP == 32) ? 32 * active : 0;
        load_picture(c, &pics[p]);
        compress_picture(c);
        release_followers(c);
        for (int i = 0; i < 16; i++)
        {
          pics[p].prof[i] = g->prof[i]; pics[p].profN[i] = g->profN[i];
          pics[p].prof2[i] = g->prof2[i]; pics[p].prof2N[i] = g->prof2N[i];
        }
      }
      else follower_loop(c);
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
  PicDesc* dPics;                         // [maxFrames] picture queue of the current submit
  FrameGlobal* dScratch;                  // [maxBlocks * FHC_WARPS_PER_BLOCK] per-warp scratch
  int* dQueue;                            // next picture index
  int maxBlocks;                          // resident blocks of the persistent kernel
  Tables* dTables;
  uint8_t* dIn;                           // maxFrames * frameSamples * 2 bytes
  int16_t* dRec;                          // maxFrames * frameSamples
  fhc_ctu_result* dOut;                   // maxFrames * numCtus
  cudaStream_t stream;
  cudaEvent_t evStart, evStop;
  int firstPoc, nSubmitted, pending;
  std::vector<PicDesc> hostPics;
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
  c->hostPics.resize((size_t)c->maxFrames);
  {
    // persistent grid: as many blocks as fit on the device, but never more warps than pictures
    cudaDeviceProp prop;
    int perSm = 0;
    CK(cudaGetDeviceProperties(&prop, device));
    CK(cudaFuncSetAttribute(fhc_compress_pictures_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kernel_smem_bytes()));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fhc_compress_pictures_kernel, FHC_BLOCK_THREADS, kernel_smem_bytes()));
    if (perSm < 1) { snprintf(g_err, sizeof(g_err), "fhc_create: kernel does not fit an SM"); delete c; return FHC_ERR_CUDA; }
    const int need = (c->maxFrames + FHC_PICS_PER_BLOCK - 1) / FHC_PICS_PER_BLOCK;
    c->maxBlocks = prop.multiProcessorCount * perSm;
    if (c->maxBlocks > need) c->maxBlocks = need;
  }
  CK(cudaMalloc(&c->dPics, sizeof(PicDesc) * (size_t)c->maxFrames));
  CK(cudaMalloc(&c->dScratch, sizeof(FrameGlobal) * (size_t)c->maxBlocks * FHC_PICS_PER_BLOCK));
  CK(cudaMalloc(&c->dQueue, sizeof(int)));
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
  CK(cudaDeviceSetLimit(cudaLimitStackSize, 24576));
  *out = c;
  return FHC_OK;
}

int fhc_destroy(fhc_ctx* c)
{
  if (!c) return FHC_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->dPics); cudaFree(c->dScratch); cudaFree(c->dQueue); cudaFree(c->dTables); cudaFree(c->dIn); cudaFree(c->dRec); cudaFree(c->dOut);
  cudaEventDestroy(c->evStart); cudaEventDestroy(c->evStop);
  cudaStreamDestroy(c->stream);
  delete c;
  return FHC_OK;
}

// fills the picture queue and launches the persistent kernel on c->stream
static int launch(fhc_ctx* c, int n, int sampleBytes, const fhc_slice_params* sp)
{
  for (int i = 0; i < n; i++)
  {
    const uint8_t* in = c->dIn + (size_t)i * c->frameSamples * (size_t)sampleBytes;
    int16_t* rec = c->dRec + (size_t)i * c->frameSamples;
    PicDesc& d = c->hostPics[(size_t)i];
    memset(&d, 0, sizeof(d));
    d.org[0] = in; d.org[1] = in + c->lumaSamples * (size_t)sampleBytes; d.org[2] = in + (c->lumaSamples + c->lumaSamples / 4) * (size_t)sampleBytes;
    d.rec[0] = rec; d.rec[1] = rec + c->lumaSamples; d.rec[2] = rec + c->lumaSamples + c->lumaSamples / 4;
    d.out = c->dOut + (size_t)i * c->numCtus;
    d.sp = sp[i];
    d.width = c->seq.width; d.height = c->seq.height; d.bitDepth = c->seq.bitDepth; d.sampleBytes = sampleBytes;
    d.wc = (c->seq.width + 63) / 64; d.hc = (c->seq.height + 63) / 64;
  }
  CK(cudaMemcpyAsync(c->dPics, c->hostPics.data(), sizeof(PicDesc) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(c->dQueue, 0, sizeof(int), c->stream));
  int blocks = (n + FHC_PICS_PER_BLOCK - 1) / FHC_PICS_PER_BLOCK;
  if (blocks > c->maxBlocks) blocks = c->maxBlocks;
  CK(cudaEventRecord(c->evStart, c->stream));
  fhc_compress_pictures_kernel<<<blocks, FHC_BLOCK_THREADS, kernel_smem_bytes(), c->stream>>>(c->dPics, n, c->dQueue, c->dScratch, c->dTables);
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

int fhc_last_kernel_ms(fhc_ctx* c, float* kernel_ms)
{
  if (!c || !kernel_ms) return FHC_ERR_ARG;
  if (c->nSubmitted == 0) return FHC_ERR_STATE;
  CK(cudaSetDevice(c->device));
  int rc = wait_done(c);
  if (rc != FHC_OK) return rc;
  CK(cudaEventElapsedTime(kernel_ms, c->evStart, c->evStop));
  return FHC_OK;
}

int fhc_pictures_in_flight(int device)
{
  int ndev = 0, perSm = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return FHC_ERR_ARG;
  CK(cudaSetDevice(device));
  CK(cudaGetDeviceProperties(&prop, device));
  CK(cudaFuncSetAttribute(fhc_compress_pictures_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kernel_smem_bytes()));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fhc_compress_pictures_kernel, FHC_BLOCK_THREADS, kernel_smem_bytes()));
  return prop.multiProcessorCount * perSm * FHC_PICS_PER_BLOCK;
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
  CK(cudaMemcpy(cycles16, c->dPics[slot].prof, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(counts16, c->dPics[slot].profN, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return FHC_OK;
}
extern "C" int fhc_debug_profile2(fhc_ctx* c, int slot, long long* cycles16, long long* counts16)
{
  if (!c || slot < 0 || slot >= c->maxFrames) return FHC_ERR_ARG;
  CK(cudaMemcpy(cycles16, c->dPics[slot].prof2, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(counts16, c->dPics[slot].prof2N, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return FHC_OK;
}
// debug: launch geometry {blocks, warps per block, dynamic shared memory bytes, per-warp global scratch bytes}
extern "C" int fhc_debug_geometry(fhc_ctx* c, long long* out4)
{
  if (!c || !out4) return FHC_ERR_ARG;
  out4[0] = c->maxBlocks; out4[1] = FHC_WARPS_PER_BLOCK * 1000 + FHC_GROUP; out4[2] = (long long)kernel_smem_bytes(); out4[3] = (long long)sizeof(FrameGlobal);
  return FHC_OK;
}
