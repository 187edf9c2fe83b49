// fasthevc_cuda.cu - kernel entry + C ABI (include/fasthevc_cuda.h) of libfasthevc_cuda.so.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -shared
// (-fmad=false: every RD cost is IEEE double in the reference's operation order).
#define FHC_DEVICE_BUILD 1
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#include <math.h>
#include <vector>
#include "fhc_engine.cuh"
#include "fhc_deblock.cuh"
#include "fhc_sao.cuh"

using namespace fhc;

#ifndef FHC_BLOCKS_PER_SM
#define FHC_BLOCKS_PER_SM (2 / FHC_PICS_PER_BLOCK)      /* two picture chains per SM either way */
#endif
#define FHC_BATCHES 2                     /* submits that may be in flight: one running, one queued behind it */

// ----------------------------------------------------------------------------------------------
// kernel: persistent blocks; the unit of scheduling is one CTU row of one picture
//
// A picture is a sequential chain (CABAC contexts, neighbours), but nothing ties it to one SM: at a CTU row boundary
// the chain consists of the contexts (PicState), the line buffers and the results, all in global memory.  After every
// row a chain slot gives its picture back and takes the ready picture with the most work left (rows left x the
// host's cost weight): the pictures that would finish last run all the time, the cheap ones fill the gaps, and the
// blocks finish together.  The `lone` most expensive pictures are different: two chains of a block each run at
// about two thirds of the speed of a lone one, and the longest pictures set the makespan - they stay
// on the slot that takes them, alone in their block, until they are finished.
// The same holds dynamically: while some block has found nothing to do (sched[SC_IDLE] > 0), no block takes a second
// chain, so that towards the end of a submit (or with fewer pictures than slots) the chains spread out over the blocks.
// sched[SC_FINISHED] = finished pictures, sched[SC_IDLE] = blocks without a chain, sched[SC_WEIGHT + p] = cost weight of
// picture p | SC_LONE_BIT.
// ----------------------------------------------------------------------------------------------
enum { SC_FINISHED = 0, SC_IDLE = 1, SC_ALIVE = 2, SC_WEIGHT = 8, SC_LONE_BIT = 1 << 30 };

// picture index, -1 = nothing ready now, -2 = every picture is finished
__device__ __forceinline__ int sched_take(PicState* states, int* sched, int numPics, int rowsPerPic, int allowLone, int* lone)
{
  for (;;)
  {
    int best = -1, bestLone = 0;
    long long bestPrio = -1;
    for (int p = 0; p < numPics; p++)
    {
      if (*reinterpret_cast<volatile int32_t*>(&states[p].status) != ST_READY) continue;
      const int w = __ldg(sched + SC_WEIGHT + p), isLone = (w & SC_LONE_BIT) != 0;
      if (isLone && !allowLone) continue;
      const long long prio = (long long)(rowsPerPic - *reinterpret_cast<volatile int32_t*>(&states[p].nextRow)) * (w & (SC_LONE_BIT - 1)) + ((long long)isLone << 50);
      if (prio > bestPrio) { bestPrio = prio; best = p; bestLone = isLone; }
    }
    if (best < 0) return *reinterpret_cast<volatile int*>(sched + SC_FINISHED) >= numPics ? -2 : -1;
    if (atomicCAS(&states[best].status, ST_READY, ST_RUNNING) == ST_READY)
    {
      __threadfence();
      *lone = bestLone;
      return best;
    }
  }
}

__global__ void __launch_bounds__(FHC_BLOCK_THREADS, FHC_BLOCKS_PER_SM)
fhc_ctu_kernel(const PicDesc* pics, PicState* states, int numPics, int rowsPerPic, int ctusPerRow, int* sched, unsigned char* workspace, size_t wsPerChain,
               size_t smemPerChain, const Tables* tables, int chainsPerBlock, int wpp, int packTail)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sItem[FHC_PICS_PER_BLOCK], sShape[FHC_PICS_PER_BLOCK], sPinned, sExit, sIdle;
  // chain = one picture at a time on 5 role warps; a block holds FHC_PICS_PER_BLOCK chains that advance in lockstep
#if FHC_PICS_PER_BLOCK == 2 && !defined(FHC_PLAIN_WARP_ORDER)
  // Warp w runs on sub-partition w % 4 of the SM, each with its own scheduler and first-level instruction cache.  The
  // twins of a role share a sub-partition (they run the same routine at the same time, see twin_align), and the two
  // warps on a chain's critical path - the depth-3 pair - have theirs to themselves:
  //   sub-partition 2: 3A of chain 0 / 1 (warps 2, 6)     sub-partition 0: depth 2 of both chains + depth 0 of chain 0 (warps 0, 4, 8)
  //   sub-partition 3: 3B of chain 0 / 1 (warps 3, 7)     sub-partition 1: depth 1 of both chains + depth 0 of chain 1 (warps 1, 5, 9)
  const int warp = threadIdx.x >> 5;
  const int inst = warp < 8 ? (warp >> 2) : warp - 8;
  const int role = warp >= 8 ? ROLE_0 : ((warp & 3) == 0 ? ROLE_2 : ((warp & 3) == 1 ? ROLE_1 : ((warp & 3) == 2 ? ROLE_3A : ROLE_3B)));
  const int tid = (role << 5) | (threadIdx.x & 31), nthr = FHC_INST_THREADS;
#else
  const int inst = threadIdx.x / FHC_INST_THREADS, tid = threadIdx.x - inst * FHC_INST_THREADS, nthr = FHC_INST_THREADS;
#endif
  const bool scheduler = inst == 0 && tid == 0;            // the block's scheduling thread
  unsigned char* mine = smem_raw + (size_t)inst * smemPerChain;
  Blk* b = reinterpret_cast<Blk*>(mine);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&b->t);
    for (int i = tid; i < (int)(sizeof(Tables) / 4); i += nthr) dst[i] = src[i];
  }
  if (tid == 0)
  {
    init_block(b, mine + ((sizeof(Blk) + 15) & ~(size_t)15), workspace + ((size_t)blockIdx.x * FHC_PICS_PER_BLOCK + inst) * wsPerChain);
    b->inst = inst;
  }
  if (scheduler) { sPinned = 0; sExit = 0; sIdle = 0; for (int k = 0; k < FHC_PICS_PER_BLOCK; k++) sItem[k] = -1; }
  __syncthreads();
  for (;;)
  {
    if (scheduler)
    {
      for (;;)
      {
        int any = 0, never = 0;
        for (int k = 0; k < FHC_PICS_PER_BLOCK; k++)
        {
          if (sPinned) { if (k == 0) any = 1; else sItem[k] = -1; continue; }      // slot 0 keeps its lone picture, the others stay empty
          if (k >= chainsPerBlock) { sItem[k] = -1; continue; }                    // experiments: fewer chains per block
          if (k > 0 && !packTail && *reinterpret_cast<volatile int*>(sched + SC_IDLE) > 0) { sItem[k] = -1; continue; }   // a block without any chain goes first
          int lone = 0;
          const int item = sched_take(states, sched, numPics, rowsPerPic, k == 0, &lone);
          sItem[k] = item >= 0 ? item : -1;
          if (item >= 0)
          {
            any = 1;
            if (lone) sPinned = 1;
            if (sIdle) { sIdle = 0; atomicSub(sched + SC_IDLE, 1); }
          }
          else if (item == -2) never = 1;
        }
        if (any) break;
        if (never) { sExit = 1; break; }
        if (packTail)
        {
          // Pipelined submits: the next submit's kernel is queued behind this one and starts on whatever SMs this one
          // frees.  A block that finds nothing to do leaves as soon as the pictures still unfinished fit into the chain
          // slots of the blocks that stay (the invariant alive * chainsPerBlock >= unfinished holds at every exit).
          const int unfinished = numPics - *reinterpret_cast<volatile int*>(sched + SC_FINISHED);
          const int old = atomicSub(sched + SC_ALIVE, 1);
          if ((old - 1) * chainsPerBlock >= unfinished) { sExit = 1; break; }
          atomicAdd(sched + SC_ALIVE, 1);
        }
        if (!sIdle) { sIdle = 1; atomicAdd(sched + SC_IDLE, 1); }
        __nanosleep(5000);
      }
    }
    __syncthreads();
    if (sExit) break;
    const int p = sItem[inst];
    const PicDesc* pd = p >= 0 ? &pics[p] : nullptr;
    const int row = pd ? *reinterpret_cast<volatile int32_t*>(&states[p].nextRow) : 0;
    // the chains of a block align their role warps at fixed points of a CTU's schedule; the schedule depends on the CTU's
    // shape (cut off at the picture's bottom edge or not), so only chains on rows of the same kind are aligned
    // (a picture with CU-split hints has a content-dependent schedule: it gets a shape of its own and is never aligned)
    if (tid == 0) { b->active = pd != nullptr; sShape[inst] = pd ? (pd->hints ? 2 + inst : (row + 1 == rowsPerPic)) : -1; }
    __syncthreads();
    if (tid == 0)
    {
      int n = 0, same = 1;
      for (int k = 0; k < FHC_PICS_PER_BLOCK; k++)
      {
        n += reinterpret_cast<Blk*>(smem_raw + (size_t)k * smemPerChain)->active;
        if (sShape[k] >= 0 && sShape[k] != sShape[inst]) same = 0;
      }
      b->twin = n > 1 && same;
    }
    compress_picture(b, pd, row * ctusPerRow, ctusPerRow, tid, nthr, wpp);
    if (tid == 0 && pd)
    {
      if (row + 1 == rowsPerPic)
      {
        __threadfence_system();
        *reinterpret_cast<volatile int32_t*>(pd->doneFlag) = 1;
        *reinterpret_cast<volatile int32_t*>(&states[p].status) = ST_FINISHED;
        __threadfence();
        atomicAdd(sched + SC_FINISHED, 1);
        if (inst == 0) sPinned = 0;
      }
      else if (!wpp)                       // (wavefront pictures: the row itself released the next one when it was two CTUs ahead)
      {
        *reinterpret_cast<volatile int32_t*>(&states[p].nextRow) = row + 1;
        if (!(inst == 0 && sPinned)) { __threadfence(); *reinterpret_cast<volatile int32_t*>(&states[p].status) = ST_READY; }
      }
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static uint64_t g_launches = 0;

#define CKG(call, onfail)                                                                                \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      snprintf(g_err, sizeof(g_err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      onfail;                                                                                            \
      return FHC_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)
#define CK(call) CKG(call, (void)0)

// one submit's worth of device state
struct Batch
{
  PicDesc* dPics = nullptr;               // [maxFrames]
  uint8_t* dIn = nullptr;                 // maxFrames pictures, inBytesPerFrame each
  size_t inBytesPerFrame = 0;
  fhc_ctu_result* dOut = nullptr;         // [maxFrames * numCtus]
  int16_t* dLines = nullptr;              // [maxFrames][2][2 * width]
  uint8_t* dDepthRows = nullptr;          // [maxFrames][2][width / 4]
  int32_t* hFlags = nullptr;              // [maxFrames] mapped pinned host memory: picture done
  int32_t* dFlags = nullptr;              // device alias of hFlags
  long long* dProf = nullptr;             // [maxFrames][64] cycle counters (FHC_PROFILE builds)
  int32_t* hTrace = nullptr;              // [maxFrames] mapped pinned host memory: progress words (FHC_TRACE=1 in the environment)
  int32_t* dTrace = nullptr;
  uint8_t* dHints = nullptr;              // [maxFrames][numCtus][4] CU-split hints of the last hinted submit (allocated on first use)
  int* dSched = nullptr;                  // scheduler words + ring of ready rows (see the kernel)
  PicState* dState = nullptr;             // [maxFrames] where every picture chain continues
  RowState* dRows = nullptr;              // wavefront sequences: [maxFrames][rows]
  cudaStream_t compute = nullptr;         // the batch's kernel stream: kernels of two submits may overlap (pipelined submits)
  unsigned char* dWorkspace = nullptr;    // per-chain scratch of the batch's kernel
  cudaEvent_t evCopied = nullptr, evStart = nullptr, evStop = nullptr;
  int firstPoc = 0, n = 0, launched = 0;
  std::vector<int> qps;                   // slice QP per picture of the submit (deblocking)
  uint64_t serial = 0;                    // submit counter value of this batch (0 = never used)
};

struct fhc_ctx
{
  fhc_seq_params seq;
  int device = 0, numCtus = 0, maxFrames = 0, maxBlocks = 0;
  size_t lumaSamples = 0, frameSamples = 0, smemBytes = 0, wsPerBlock = 0;   // smemBytes / wsPerBlock: per picture chain
  Tables* dTables = nullptr;
  cudaStream_t copyIn = nullptr, copyOut = nullptr;
  int pipelined = 0;                       // FHC_OPT_PIPELINED_SUBMITS
  int numSms = 0;
  // deblocking (fhc_deblock_pictures): working planes, packed output and picture table, sized for the largest call so far
  int16_t* dDbkPlanes = nullptr; uint8_t* dDbkOut = nullptr; DbkPic* dDbkPics = nullptr; int dbkCap = 0, dbkSampleBytes = 0;
  cudaEvent_t evDbk0 = nullptr, evDbk1 = nullptr; float dbkMs = 0.f;
  fhc_sao_stats* dSaoStats = nullptr; SaoPic* dSaoPics = nullptr; int saoCap = 0;     // SAO statistics of the last call
  cudaEvent_t evSpan = nullptr;            // fhc_span_begin: recorded before the next launch's kernel
  int spanArmed = 0, spanOpen = 0;
  Batch batch[FHC_BATCHES];
  uint64_t submits = 0;
  std::vector<PicDesc> hostPics;
  std::vector<int> hostSched;
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

static int kernel_geometry(int device, size_t* smemBytes, size_t* wsPerBlock, int* maxBlocks)
{
  cudaDeviceProp prop;
  int perSm = 0;
  block_footprint(smemBytes, wsPerBlock);
  CK(cudaGetDeviceProperties(&prop, device));
  const size_t blockSmem = *smemBytes * FHC_PICS_PER_BLOCK;
  CK(cudaFuncSetAttribute(fhc_ctu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)blockSmem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fhc_ctu_kernel, FHC_BLOCK_THREADS, blockSmem));
  if (perSm < 1) { snprintf(g_err, sizeof(g_err), "kernel does not fit an SM (%zu bytes of shared memory)", blockSmem); return FHC_ERR_CUDA; }
  if (const char* e = getenv("FHC_MAX_BLOCKS_PER_SM")) { const int cap = atoi(e); if (cap >= 1 && cap < perSm) perSm = cap; }   // experiments
  *maxBlocks = prop.multiProcessorCount * perSm;
  return FHC_OK;
}

static void free_batch(Batch& bt)
{
  cudaFree(bt.dPics); cudaFree(bt.dIn); cudaFree(bt.dOut); cudaFree(bt.dLines); cudaFree(bt.dDepthRows); cudaFree(bt.dSched); cudaFree(bt.dState); cudaFree(bt.dRows); cudaFree(bt.dProf); cudaFree(bt.dHints); cudaFree(bt.dWorkspace);
  if (bt.compute) cudaStreamDestroy(bt.compute);
  if (bt.hFlags) cudaFreeHost(bt.hFlags);
  if (bt.hTrace) cudaFreeHost(bt.hTrace);
  if (bt.evCopied) cudaEventDestroy(bt.evCopied);
  if (bt.evStart) cudaEventDestroy(bt.evStart);
  if (bt.evStop) cudaEventDestroy(bt.evStop);
  bt = Batch();
}

int fhc_destroy(fhc_ctx* c)
{
  if (!c) return FHC_ERR_ARG;
  cudaSetDevice(c->device);
  for (int i = 0; i < FHC_BATCHES; i++) if (c->batch[i].compute) cudaStreamSynchronize(c->batch[i].compute);
  if (c->copyIn) cudaStreamSynchronize(c->copyIn);
  if (c->copyOut) cudaStreamSynchronize(c->copyOut);
  for (int i = 0; i < FHC_BATCHES; i++) free_batch(c->batch[i]);
  cudaFree(c->dTables); cudaFree(c->dDbkPlanes); cudaFree(c->dDbkOut); cudaFree(c->dDbkPics); cudaFree(c->dSaoStats); cudaFree(c->dSaoPics);
  if (c->evSpan) cudaEventDestroy(c->evSpan);
  if (c->evDbk0) cudaEventDestroy(c->evDbk0);
  if (c->evDbk1) cudaEventDestroy(c->evDbk1);
  if (c->copyIn) cudaStreamDestroy(c->copyIn);
  if (c->copyOut) cudaStreamDestroy(c->copyOut);
  delete c;
  return FHC_OK;
}

int fhc_create(const fhc_seq_params* seq, int device, fhc_ctx** out)
{
  if (!seq || !out) return FHC_ERR_ARG;
  if (seq->width <= 0 || seq->height <= 0 || (seq->width & 7) || (seq->height & 7) || (seq->bitDepth != 8 && seq->bitDepth != 10) ||
      seq->chromaFormat != 1 || seq->ctuSize != 64 || seq->maxCuDepth != 4 || seq->tuLog2Max != 5 || seq->tuLog2Min != 2 ||
      seq->tuMaxDepthIntra != 3 || !seq->useRdoq || !seq->useRdoqTs || !seq->useTransformSkip || !seq->useTransformSkipFast ||
      !seq->signHideFlag || !seq->strongIntraSmoothing || !seq->fastUdiUseMpm || seq->maxFramesInFlight <= 0 ||
      (seq->waveFrontSynchro != 0 && seq->waveFrontSynchro != 1))
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
  c->hostPics.resize((size_t)c->maxFrames);
  // every failure below releases what was allocated so far (fhc_destroy is null-safe)
  if (kernel_geometry(device, &c->smemBytes, &c->wsPerBlock, &c->maxBlocks) != FHC_OK) { fhc_destroy(c); return FHC_ERR_CUDA; }
  {
    const long long chains = (long long)c->maxFrames * (seq->waveFrontSynchro ? (seq->height + 63) / 64 : 1);    // wavefront: every CTU row is a chain
    if (c->maxBlocks > chains) c->maxBlocks = (int)chains;
  }
  CKG(cudaMalloc(&c->dTables, sizeof(Tables)), fhc_destroy(c));
  CKG(cudaEventCreate(&c->evSpan), fhc_destroy(c));
  CKG(cudaStreamCreateWithFlags(&c->copyIn, cudaStreamNonBlocking), fhc_destroy(c));
  CKG(cudaStreamCreateWithFlags(&c->copyOut, cudaStreamNonBlocking), fhc_destroy(c));
  const size_t inBytes = c->frameSamples * (size_t)(seq->bitDepth > 8 ? 2 : 1);
  for (int i = 0; i < FHC_BATCHES; i++)
  {
    Batch& bt = c->batch[i];
    CKG(cudaMalloc(&bt.dPics, sizeof(PicDesc) * (size_t)c->maxFrames), fhc_destroy(c));
    CKG(cudaMalloc(&bt.dWorkspace, c->wsPerBlock * (size_t)c->maxBlocks * FHC_PICS_PER_BLOCK), fhc_destroy(c));
    CKG(cudaStreamCreateWithFlags(&bt.compute, cudaStreamNonBlocking), fhc_destroy(c));
    CKG(cudaMalloc(&bt.dIn, inBytes * (size_t)c->maxFrames), fhc_destroy(c));
    bt.inBytesPerFrame = inBytes;
    CKG(cudaMalloc(&bt.dOut, sizeof(fhc_ctu_result) * (size_t)c->numCtus * (size_t)c->maxFrames), fhc_destroy(c));
    CKG(cudaMalloc(&bt.dLines, sizeof(int16_t) * 4 * (size_t)seq->width * (size_t)c->maxFrames), fhc_destroy(c));
    CKG(cudaMalloc(&bt.dDepthRows, (size_t)seq->width * (size_t)c->maxFrames), fhc_destroy(c));
    CKG(cudaMalloc(&bt.dSched, sizeof(int) * (SC_WEIGHT + (size_t)c->maxFrames)), fhc_destroy(c));
    CKG(cudaMalloc(&bt.dState, sizeof(PicState) * (size_t)c->maxFrames), fhc_destroy(c));
    if (seq->waveFrontSynchro) CKG(cudaMalloc(&bt.dRows, sizeof(RowState) * (size_t)c->maxFrames * (size_t)((seq->height + 63) / 64)), fhc_destroy(c));
    CKG(cudaMalloc(&bt.dProf, sizeof(long long) * 64 * (size_t)c->maxFrames), fhc_destroy(c));
    CKG(cudaMemset(bt.dProf, 0, sizeof(long long) * 64 * (size_t)c->maxFrames), fhc_destroy(c));
    CKG(cudaHostAlloc(&bt.hFlags, sizeof(int32_t) * (size_t)c->maxFrames, cudaHostAllocMapped), fhc_destroy(c));
    CKG(cudaHostGetDevicePointer(&bt.dFlags, bt.hFlags, 0), fhc_destroy(c));
    if (getenv("FHC_TRACE"))
    {
      CKG(cudaHostAlloc(&bt.hTrace, sizeof(int32_t) * 8 * (size_t)c->maxFrames, cudaHostAllocMapped), fhc_destroy(c));
      CKG(cudaHostGetDevicePointer(&bt.dTrace, bt.hTrace, 0), fhc_destroy(c));
    }
    CKG(cudaEventCreate(&bt.evCopied), fhc_destroy(c));
    CKG(cudaEventCreate(&bt.evStart), fhc_destroy(c));
    CKG(cudaEventCreate(&bt.evStop), fhc_destroy(c));
  }
  Tables t;
  build_tables(&t, seq->bitDepth);
  CKG(cudaMemcpy(c->dTables, &t, sizeof(t), cudaMemcpyHostToDevice), fhc_destroy(c));
  *out = c;
  return FHC_OK;
}

// the batch a new submit goes to: the least recently used one, after its kernel has finished
static int take_batch(fhc_ctx* c, Batch** out)
{
  Batch* bt = &c->batch[0];
  for (int i = 1; i < FHC_BATCHES; i++) if (c->batch[i].serial < bt->serial) bt = &c->batch[i];
  if (bt->launched) CK(cudaEventSynchronize(bt->evStop));
  bt->launched = 0; bt->n = 0;
  *out = bt;
  return FHC_OK;
}

// makes sure the batch's input buffer holds n pictures of sampleBytes samples
static int ensure_input(fhc_ctx* c, Batch* bt, int sampleBytes)
{
  const size_t need = c->frameSamples * (size_t)sampleBytes;
  if (need <= bt->inBytesPerFrame) return FHC_OK;
  CK(cudaFree(bt->dIn));
  bt->dIn = nullptr; bt->inBytesPerFrame = 0;
  CK(cudaMalloc(&bt->dIn, need * (size_t)c->maxFrames));
  bt->inBytesPerFrame = need;
  return FHC_OK;
}

// How many of the most expensive pictures run alone in their block (twin builds).  Two chains that share a block each
// advance at about two thirds (`rate`) of a lone chain's speed; the kernel finishes when the slowest picture does.  Cost model:
// relative time of a picture by QP (measured on the bench clip: the time roughly halves every 12 QP steps).  Pick the
// count that minimises max(longest lone picture, longest shared picture / rate, remaining work / remaining capacity).
static double picture_weight(int qp) { return exp2((32.0 - qp) / 12.0); }
static int lone_pictures(const fhc_slice_params* sp, const std::vector<int>& order, int n, int blocks)
{
  if (FHC_PICS_PER_BLOCK == 1) return 0;
  if (const char* e = getenv("FHC_LONE")) { const int v = atoi(e); return v < 0 ? 0 : (v > blocks ? blocks : (v > n ? n : v)); }
  const double rate = 0.66;
  std::vector<double> w((size_t)n);
  double total = 0;
  for (int i = 0; i < n; i++) { w[(size_t)i] = picture_weight(sp[order[(size_t)i]].sliceQp); total += w[(size_t)i]; }
  int best = 0;
  double bestT = 1e300, loneWork = 0;
  for (int k = 0; k <= n && k <= blocks; k++)
  {
    // k lone pictures on k blocks; the others share the remaining blocks, two chains each
    double t = k ? w[0] : 0.0;
    if (k < n)
    {
      if (k == blocks) break;
      t = std::max(t, std::max(w[(size_t)k] / rate, (total - loneWork) / ((blocks - k) * 2 * rate)));
    }
    if (t < bestT - 1e-12) { bestT = t; best = k; }
    if (k < n) loneWork += w[(size_t)k];
  }
  return best;
}

// fills the ring of ready rows and launches the persistent kernel on the compute stream
static int launch(fhc_ctx* c, Batch* bt, int first_poc, int n, int sampleBytes, const fhc_slice_params* sp, bool hinted = false)
{
  const int W = c->seq.width, wpp = c->seq.waveFrontSynchro;
  const size_t frameBytes = c->frameSamples * (size_t)sampleBytes;
  for (int i = 0; i < n; i++)
  {
    const uint8_t* in = bt->dIn + (size_t)i * frameBytes;
    PicDesc& d = c->hostPics[(size_t)i];
    memset(&d, 0, sizeof(d));
    d.org[0] = in; d.org[1] = in + c->lumaSamples * (size_t)sampleBytes; d.org[2] = in + (c->lumaSamples + c->lumaSamples / 4) * (size_t)sampleBytes;
    d.out = bt->dOut + (size_t)i * c->numCtus;
    int16_t* ln = bt->dLines + (size_t)i * 4 * W;
    uint8_t* dr = bt->dDepthRows + (size_t)i * W;
    for (int k = 0; k < 2; k++)
    {
      d.line[k][0] = ln + (size_t)k * 2 * W; d.line[k][1] = d.line[k][0] + W; d.line[k][2] = d.line[k][1] + W / 2;
      d.depthRow[k] = dr + (size_t)k * (W / 2);
    }
    d.sp = sp[i];
    d.width = W; d.height = c->seq.height; d.bitDepth = c->seq.bitDepth; d.sampleBytes = sampleBytes;
    d.wc = (W + 63) / 64; d.hc = (c->seq.height + 63) / 64;
    d.doneFlag = bt->dFlags + i;
    d.state = bt->dState + i;
    d.prof = bt->dProf + 64 * (size_t)i;
    d.hints = hinted ? bt->dHints + (size_t)i * 4 * (size_t)c->numCtus : nullptr;
    d.rows = wpp ? bt->dRows + (size_t)i * (size_t)((c->seq.height + 63) / 64) : nullptr;
    bt->hFlags[i] = 0;
    if (bt->hTrace) { d.trace = bt->dTrace + 8 * i; memset(bt->hTrace + 8 * i, 0, 32); }
  }
  // picture descriptors travel on the copy stream (pageable source: staged before the call returns), so a submit
  // queued behind a running kernel does not wait for it on the host
  CK(cudaMemcpyAsync(bt->dPics, c->hostPics.data(), sizeof(PicDesc) * (size_t)n, cudaMemcpyHostToDevice, c->copyIn));
  CK(cudaEventRecord(bt->evCopied, c->copyIn));
  CK(cudaStreamWaitEvent(bt->compute, bt->evCopied, 0));
  // Work distribution (see the kernel): the ring starts with every picture's first row, most expensive (lowest QP)
  // picture first; the scheduler words and the ring's initial part travel on the copy stream (pageable source: staged
  // before the call returns, and the copy stream is never behind a running kernel).
  const int rows = (c->seq.height + 63) / 64, ctusPerRow = (W + 63) / 64;
  // blocks: one per picture is enough without wavefronts (a picture is one chain); with them every row is a chain
  const long long chainsWanted = wpp ? (long long)n * rows : n;
  const int blocks = chainsWanted < c->maxBlocks ? (int)chainsWanted : c->maxBlocks;
  // Pipelined submits (FHC_OPT_PIPELINED_SUBMITS): the host keeps a second submit queued, so what counts is the throughput
  // across submits, not the makespan of one.  No picture runs alone in its block (pairs of twin-aligned chains get the
  // most out of an SM), and the blocks of a kernel leave as soon as the remaining pictures fit into the
  // others, so that the next submit's kernel - queued on its own stream, with its own scratch - fills the freed SMs.
  // (not for wavefront pictures: a picture then keeps many chain slots busy, and counting pictures would send blocks away too early)
  const int pack = c->pipelined && !hinted && !wpp ? 1 : 0;
  {
    std::vector<int> order((size_t)n);
    for (int i = 0; i < n; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return sp[a].sliceQp < sp[b2].sliceQp; });    // expensive first
    const int lone = (wpp || pack) ? 0 : lone_pictures(sp, order, n, blocks);         // (wavefront: a picture is many chains, none of them pinned; pipelined: see below)
    c->hostSched.assign((size_t)SC_WEIGHT + (size_t)n, 0);
    c->hostSched[SC_ALIVE] = blocks;
    for (int i = 0; i < n; i++)
      c->hostSched[(size_t)SC_WEIGHT + (size_t)order[(size_t)i]] = (int)(1000.0 * picture_weight(sp[order[(size_t)i]].sliceQp)) | (i < lone ? SC_LONE_BIT : 0);
  }
  CK(cudaMemcpyAsync(bt->dSched, c->hostSched.data(), sizeof(int) * c->hostSched.size(), cudaMemcpyHostToDevice, c->copyIn));
  CK(cudaMemsetAsync(bt->dState, 0, sizeof(PicState) * (size_t)n, c->copyIn));
  if (wpp) CK(cudaMemsetAsync(bt->dRows, 0, sizeof(RowState) * (size_t)n * (size_t)rows, c->copyIn));
#ifdef FHC_PROFILE
  CK(cudaMemsetAsync(bt->dProf, 0, sizeof(long long) * 64 * (size_t)n, c->copyIn));
#endif
  CK(cudaEventRecord(bt->evCopied, c->copyIn));
  CK(cudaStreamWaitEvent(bt->compute, bt->evCopied, 0));
  // The schedule of a hinted picture depends on its content, so its role warps cannot be twin-aligned with another chain's.
  // Measured on the bench workload with CNN hints (256 jobs, profiles/r02_c5_chains.txt): pairs + lone pictures 14.7 s, one
  // chain per block 15.5 s, pairs without lone pictures 15.9 s - the pairing stays (FHC_HINTED_CHAINS=1: one chain per block).
  int chains = FHC_PICS_PER_BLOCK;
  if (hinted) { const char* e = getenv("FHC_HINTED_CHAINS"); if (e && atoi(e) == 1) chains = 1; }
  if (c->spanArmed) { CK(cudaEventRecord(c->evSpan, bt->compute)); c->spanArmed = 0; c->spanOpen = 1; }
  CK(cudaEventRecord(bt->evStart, bt->compute));
  fhc_ctu_kernel<<<blocks, FHC_BLOCK_THREADS, c->smemBytes * FHC_PICS_PER_BLOCK, bt->compute>>>(bt->dPics, bt->dState, n, rows, ctusPerRow, bt->dSched, bt->dWorkspace, c->wsPerBlock,
                                                                                            c->smemBytes, c->dTables, chains, wpp, pack);
  CK(cudaGetLastError());
  CK(cudaEventRecord(bt->evStop, bt->compute));
  g_launches++;
  bt->firstPoc = first_poc; bt->n = n; bt->launched = 1; bt->serial = ++c->submits;
  bt->qps.resize((size_t)n);
  for (int i = 0; i < n; i++) bt->qps[(size_t)i] = sp[i].sliceQp;
  return FHC_OK;
}

int fhc_submit_frames(fhc_ctx* c, int first_poc, int n, const void* yuv, int sampleBytes, const fhc_slice_params* per_frame)
{
  if (!c || !yuv || !per_frame || n <= 0 || n > c->maxFrames || (sampleBytes != 1 && sampleBytes != 2)) return FHC_ERR_ARG;
  if (c->seq.bitDepth > 8 && sampleBytes != 2) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  Batch* bt;
  int rc = take_batch(c, &bt);
  if (rc != FHC_OK) return rc;
  rc = ensure_input(c, bt, sampleBytes);
  if (rc != FHC_OK) return rc;
  // the copy runs on its own stream, so it overlaps the kernel of the previous submit
  CK(cudaMemcpyAsync(bt->dIn, yuv, c->frameSamples * (size_t)sampleBytes * (size_t)n, cudaMemcpyHostToDevice, c->copyIn));
  return launch(c, bt, first_poc, n, sampleBytes, per_frame);
}

int fhc_submit_frames_hinted(fhc_ctx* c, int first_poc, int n, const void* yuv, int sampleBytes, const fhc_slice_params* per_frame,
                             const uint8_t* hints)
{
  if (!hints) return fhc_submit_frames(c, first_poc, n, yuv, sampleBytes, per_frame);
  if (!c || !yuv || !per_frame || n <= 0 || n > c->maxFrames || (sampleBytes != 1 && sampleBytes != 2)) return FHC_ERR_ARG;
  if (c->seq.bitDepth > 8 && sampleBytes != 2) return FHC_ERR_ARG;
  const size_t hintBytes = (size_t)n * 4 * (size_t)c->numCtus;
  for (size_t i = 0; i < hintBytes; i++)
    if (hints[i] > FHC_HINT_SPLIT) { snprintf(g_err, sizeof(g_err), "fhc_submit_frames_hinted: hint %zu has the value %d", i, hints[i]); return FHC_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  Batch* bt;
  int rc = take_batch(c, &bt);
  if (rc != FHC_OK) return rc;
  rc = ensure_input(c, bt, sampleBytes);
  if (rc != FHC_OK) return rc;
  if (!bt->dHints) CK(cudaMalloc(&bt->dHints, (size_t)c->maxFrames * 4 * (size_t)c->numCtus));
  CK(cudaMemcpyAsync(bt->dIn, yuv, c->frameSamples * (size_t)sampleBytes * (size_t)n, cudaMemcpyHostToDevice, c->copyIn));
  CK(cudaMemcpyAsync(bt->dHints, hints, hintBytes, cudaMemcpyHostToDevice, c->copyIn));
  return launch(c, bt, first_poc, n, sampleBytes, per_frame, true);
}

// ----------------------------------------------------------------------------------------------
// deblocking of finished pictures (first stage of the per-picture tail)
// ----------------------------------------------------------------------------------------------
static Batch* find_batch(fhc_ctx* c, int poc, int* slot);
static int wait_picture(fhc_ctx* c, Batch* bt, int slot);

int fhc_deblock_pictures(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes) { return fhc_deblock_pictures_stats(c, first_poc, n, out, sampleBytes, nullptr); }

int fhc_deblock_pictures_stats(fhc_ctx* c, int first_poc, int n, void* out, int sampleBytes, fhc_sao_stats* saoStats)
{
  if (!c || !out || n <= 0 || (sampleBytes != 1 && sampleBytes != 2) || (c->seq.bitDepth > 8 && sampleBytes != 2)) return FHC_ERR_ARG;
  if (c->submits == 0) return FHC_ERR_STATE;
  int slot = 0;
  Batch* bt = find_batch(c, first_poc, &slot);
  if (!bt || slot + n > bt->n) { snprintf(g_err, sizeof(g_err), "fhc_deblock_pictures: pictures %d..%d are not in one submit", first_poc, first_poc + n - 1); return FHC_ERR_ARG; }
  CK(cudaSetDevice(c->device));
  for (int i = 0; i < n; i++) { int rc = wait_picture(c, bt, slot + i); if (rc != FHC_OK) return rc; }
  const int W = c->seq.width, H = c->seq.height, wc = (W + 63) / 64;
  const size_t frame = c->frameSamples;
  if (n > c->dbkCap || sampleBytes > c->dbkSampleBytes)
  {
    cudaFree(c->dDbkPlanes); cudaFree(c->dDbkOut); cudaFree(c->dDbkPics);
    c->dDbkPlanes = nullptr; c->dDbkOut = nullptr; c->dDbkPics = nullptr; c->dbkCap = 0;
    CK(cudaMalloc(&c->dDbkPlanes, sizeof(int16_t) * frame * (size_t)n));
    CK(cudaMalloc(&c->dDbkOut, frame * (size_t)n * (size_t)sampleBytes));
    CK(cudaMalloc(&c->dDbkPics, sizeof(DbkPic) * (size_t)n));
    c->dbkCap = n; c->dbkSampleBytes = sampleBytes;
    if (!c->evDbk0) { CK(cudaEventCreate(&c->evDbk0)); CK(cudaEventCreate(&c->evDbk1)); }
  }
  std::vector<DbkPic> hp((size_t)n);
  for (int i = 0; i < n; i++)
  {
    DbkPic& d = hp[(size_t)i];
    d.ctus = bt->dOut + (size_t)(slot + i) * c->numCtus;
    d.plane[0] = c->dDbkPlanes + frame * (size_t)i;
    d.plane[1] = d.plane[0] + c->lumaSamples;
    d.plane[2] = d.plane[1] + c->lumaSamples / 4;
    d.out = c->dDbkOut + frame * (size_t)i * (size_t)sampleBytes;
    d.qp = bt->qps[(size_t)(slot + i)]; d.pad = 0;
  }
  // on the copy-out stream: behind the device->host copies already queued, next to the running kernel
  CK(cudaMemcpyAsync(c->dDbkPics, hp.data(), sizeof(DbkPic) * (size_t)n, cudaMemcpyHostToDevice, c->copyOut));
  if (c->numSms <= 0) CK(cudaDeviceGetAttribute(&c->numSms, cudaDevAttrMultiProcessorCount, c->device));
  const int perPic = std::max(1, (c->numSms * 16 + n - 1) / n);      // grid = a multiple of the SM count over all pictures
  const dim3 grid((unsigned)perPic, (unsigned)n);
  CK(cudaEventRecord(c->evDbk0, c->copyOut));
  fhc_dbk_gather_kernel<<<grid, FHC_DBK_THREADS, 0, c->copyOut>>>(c->dDbkPics, W, H, wc);
  fhc_dbk_edges_kernel<<<grid, FHC_DBK_THREADS, 0, c->copyOut>>>(c->dDbkPics, W, H, wc, c->seq.bitDepth, 0);
  fhc_dbk_edges_kernel<<<grid, FHC_DBK_THREADS, 0, c->copyOut>>>(c->dDbkPics, W, H, wc, c->seq.bitDepth, 1);
  fhc_dbk_pack_kernel<<<grid, FHC_DBK_THREADS, 0, c->copyOut>>>(c->dDbkPics, W, H, sampleBytes);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->evDbk1, c->copyOut));
  g_launches += 4;
  if (saoStats)
  {
    // SAO statistics from the deblocked planes and the submitted originals (still resident: same lifetime as the records)
    const size_t inSb = (size_t)(c->seq.bitDepth > 8 ? 2 : 1), perPicStats = (size_t)c->numCtus * 3;
    if (frame * inSb > bt->inBytesPerFrame) { snprintf(g_err, sizeof(g_err), "fhc_deblock_pictures_stats: the submitted pictures are not resident"); return FHC_ERR_STATE; }
    if (n > c->saoCap)
    {
      cudaFree(c->dSaoStats); cudaFree(c->dSaoPics);
      c->dSaoStats = nullptr; c->dSaoPics = nullptr; c->saoCap = 0;
      CK(cudaMalloc(&c->dSaoStats, sizeof(fhc_sao_stats) * perPicStats * (size_t)n));
      CK(cudaMalloc(&c->dSaoPics, sizeof(SaoPic) * (size_t)n));
      c->saoCap = n;
    }
    std::vector<SaoPic> sp((size_t)n);
    for (int i = 0; i < n; i++)
    {
      SaoPic& d = sp[(size_t)i];
      const uint8_t* in = bt->dIn + (size_t)(slot + i) * frame * inSb;
      for (int k = 0; k < 3; k++) d.plane[k] = hp[(size_t)i].plane[k];
      d.org[0] = in; d.org[1] = in + c->lumaSamples * inSb; d.org[2] = in + (c->lumaSamples + c->lumaSamples / 4) * inSb;
      d.stats = c->dSaoStats + perPicStats * (size_t)i;
    }
    CK(cudaMemcpyAsync(c->dSaoPics, sp.data(), sizeof(SaoPic) * (size_t)n, cudaMemcpyHostToDevice, c->copyOut));
    fhc_sao_stats_kernel<<<dim3((unsigned)c->numCtus, 3, (unsigned)n), 128, 0, c->copyOut>>>(c->dSaoPics, W, H, c->seq.bitDepth, (int)inSb);
    CK(cudaGetLastError());
    g_launches += 1;
    CK(cudaMemcpyAsync(saoStats, c->dSaoStats, sizeof(fhc_sao_stats) * perPicStats * (size_t)n, cudaMemcpyDeviceToHost, c->copyOut));
  }
  CK(cudaMemcpyAsync(out, c->dDbkOut, frame * (size_t)n * (size_t)sampleBytes, cudaMemcpyDeviceToHost, c->copyOut));
  CK(cudaStreamSynchronize(c->copyOut));
  CK(cudaEventElapsedTime(&c->dbkMs, c->evDbk0, c->evDbk1));
  return FHC_OK;
}

int fhc_last_deblock_ms(fhc_ctx* c, float* ms)
{
  if (!c || !ms) return FHC_ERR_ARG;
  *ms = c->dbkMs;
  return FHC_OK;
}

int fhc_configure(fhc_ctx* c, int option, int value)
{
  if (!c) return FHC_ERR_ARG;
  if (option == FHC_OPT_PIPELINED_SUBMITS) { c->pipelined = value ? 1 : 0; return FHC_OK; }
  snprintf(g_err, sizeof(g_err), "fhc_configure: unknown option %d", option);
  return FHC_ERR_ARG;
}

int fhc_span_begin(fhc_ctx* c)
{
  if (!c) return FHC_ERR_ARG;
  c->spanArmed = 1; c->spanOpen = 0;
  return FHC_OK;
}

int fhc_span_ms(fhc_ctx* c, float* ms)
{
  if (!c || !ms) return FHC_ERR_ARG;
  if (!c->spanOpen) { snprintf(g_err, sizeof(g_err), "fhc_span_ms: no launch since fhc_span_begin"); return FHC_ERR_STATE; }
  Batch* last = nullptr;
  for (int i = 0; i < FHC_BATCHES; i++) if (c->batch[i].launched && (!last || c->batch[i].serial > last->serial)) last = &c->batch[i];
  if (!last) return FHC_ERR_STATE;
  CK(cudaSetDevice(c->device));
  for (int i = 0; i < FHC_BATCHES; i++) if (c->batch[i].launched) CK(cudaEventSynchronize(c->batch[i].evStop));
  // kernels of different submits run on different streams: the span ends with whichever finished last
  float best = 0.f;
  for (int i = 0; i < FHC_BATCHES; i++)
    if (c->batch[i].launched) { float t = 0.f; if (cudaEventElapsedTime(&t, c->evSpan, c->batch[i].evStop) == cudaSuccess && t > best) best = t; }
  *ms = best;
  return FHC_OK;
}

int fhc_device_picture(fhc_ctx* c, int poc, void** dptr, uint64_t* bytes)
{
  if (!c || !dptr) return FHC_ERR_ARG;
  Batch* best = nullptr;
  for (int i = 0; i < FHC_BATCHES; i++)
  {
    Batch* bt = &c->batch[i];
    if (bt->n > 0 && poc >= bt->firstPoc && poc < bt->firstPoc + bt->n && (!best || bt->serial > best->serial)) best = bt;
  }
  if (!best) return FHC_ERR_STATE;
  const size_t frameBytes = c->frameSamples * (size_t)(c->seq.bitDepth > 8 ? 2 : 1);
  if (frameBytes > best->inBytesPerFrame) return FHC_ERR_STATE;
  // the copy of the submit that carried the picture has been queued; make sure it has landed before the caller reads
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(best->evCopied));
  *dptr = best->dIn + (size_t)(poc - best->firstPoc) * frameBytes;
  if (bytes) *bytes = (uint64_t)frameBytes * (uint64_t)(best->firstPoc + best->n - poc);
  return FHC_OK;
}

static Batch* find_batch(fhc_ctx* c, int poc, int* slot)
{
  Batch* best = nullptr;
  for (int i = 0; i < FHC_BATCHES; i++)
  {
    Batch* bt = &c->batch[i];
    if (bt->n > 0 && poc >= bt->firstPoc && poc < bt->firstPoc + bt->n && (!best || bt->serial > best->serial)) best = bt;
  }
  if (best) *slot = poc - best->firstPoc;
  return best;
}

// waits until the device has set the picture's done flag (written by the kernel when the picture's last CTU is stored)
static int wait_picture(fhc_ctx* c, Batch* bt, int slot)
{
  volatile int32_t* f = bt->hFlags + slot;
  unsigned spins = 0;
  const char* wd = getenv("FHC_WATCHDOG_S");
  const double limit = wd ? atof(wd) : 0.0;
  const clock_t t0 = clock();
  while (!*f)
  {
    if ((++spins & 1023) == 0)
    {
      if (limit > 0 && (double)(clock() - t0) / CLOCKS_PER_SEC > limit)
      {
        // debugging aid: report where the picture's block is (FHC_TRACE=1) instead of spinning forever
        const int32_t* tw = bt->hTrace ? bt->hTrace + 8 * slot : nullptr;
        const int tr = tw ? tw[0] : -1;
        snprintf(g_err, sizeof(g_err), "watchdog: picture slot %d not finished after %.0f s; trace ctu %d z %d stage %d; roles %d %d %d %d %d", slot, limit,
                 tr >> 12, (tr >> 4) & 255, tr & 15, tw ? tw[1] : -1, tw ? tw[2] : -1, tw ? tw[3] : -1, tw ? tw[4] : -1, tw ? tw[5] : -1);
        return FHC_ERR_CUDA;
      }
      const cudaError_t q = cudaEventQuery(bt->evStop);
      if (q == cudaSuccess) { if (*f) break; snprintf(g_err, sizeof(g_err), "kernel finished without completing picture slot %d", slot); return FHC_ERR_CUDA; }
      if (q != cudaErrorNotReady) { snprintf(g_err, sizeof(g_err), "kernel failed: %s", cudaGetErrorString(q)); return FHC_ERR_CUDA; }
    }
  }
  (void)c;
  return FHC_OK;
}

int fhc_poll_picture(fhc_ctx* c, int poc)
{
  if (!c) return FHC_ERR_ARG;
  int slot = 0;
  Batch* bt = find_batch(c, poc, &slot);
  if (!bt) return FHC_ERR_STATE;
  return *(volatile int32_t*)(bt->hFlags + slot) ? 1 : 0;
}

int fhc_fetch_ctu(fhc_ctx* c, int poc, int ctuRsAddr, fhc_ctu_result* out)
{
  if (!c || !out) return FHC_ERR_ARG;
  if (c->submits == 0) return FHC_ERR_STATE;
  int slot = 0;
  Batch* bt = find_batch(c, poc, &slot);
  if (!bt || ctuRsAddr < 0 || ctuRsAddr >= c->numCtus) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  int rc = wait_picture(c, bt, slot);
  if (rc != FHC_OK) return rc;
  CK(cudaMemcpyAsync(out, bt->dOut + (size_t)slot * c->numCtus + ctuRsAddr, sizeof(fhc_ctu_result), cudaMemcpyDeviceToHost, c->copyOut));
  CK(cudaStreamSynchronize(c->copyOut));
  return FHC_OK;
}

int fhc_fetch_picture(fhc_ctx* c, int poc, fhc_ctu_result* out)
{
  if (!c || !out) return FHC_ERR_ARG;
  if (c->submits == 0) return FHC_ERR_STATE;
  int slot = 0;
  Batch* bt = find_batch(c, poc, &slot);
  if (!bt) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  int rc = wait_picture(c, bt, slot);
  if (rc != FHC_OK) return rc;
  CK(cudaMemcpyAsync(out, bt->dOut + (size_t)slot * c->numCtus, sizeof(fhc_ctu_result) * (size_t)c->numCtus, cudaMemcpyDeviceToHost, c->copyOut));
  CK(cudaStreamSynchronize(c->copyOut));
  return FHC_OK;
}

int fhc_fetch_picture_async(fhc_ctx* c, int poc, fhc_ctu_result* out)
{
  if (!c || !out) return FHC_ERR_ARG;
  if (c->submits == 0) return FHC_ERR_STATE;
  int slot = 0;
  Batch* bt = find_batch(c, poc, &slot);
  if (!bt) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  int rc = wait_picture(c, bt, slot);
  if (rc != FHC_OK) return rc;
  CK(cudaMemcpyAsync(out, bt->dOut + (size_t)slot * c->numCtus, sizeof(fhc_ctu_result) * (size_t)c->numCtus, cudaMemcpyDeviceToHost, c->copyOut));
  return FHC_OK;
}

int fhc_fetch_sync(fhc_ctx* c)
{
  if (!c) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->copyOut));
  return FHC_OK;
}

int fhc_device_input(fhc_ctx* c, void** dptr, uint64_t* bytes)
{
  if (!c || !dptr) return FHC_ERR_ARG;
  Batch* bt = &c->batch[0];
  for (int i = 1; i < FHC_BATCHES; i++) if (c->batch[i].serial < bt->serial) bt = &c->batch[i];
  *dptr = bt->dIn;
  if (bytes) *bytes = (uint64_t)bt->inBytesPerFrame * (uint64_t)c->maxFrames;
  return FHC_OK;
}

int fhc_run_resident(fhc_ctx* c, int n, int sampleBytes, const fhc_slice_params* per_frame, float* kernel_ms)
{
  if (!c || !per_frame || n <= 0 || n > c->maxFrames || (sampleBytes != 1 && sampleBytes != 2)) return FHC_ERR_ARG;
  if (c->seq.bitDepth > 8 && sampleBytes != 2) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  Batch* bt;
  int rc = take_batch(c, &bt);
  if (rc != FHC_OK) return rc;
  if (c->frameSamples * (size_t)sampleBytes > bt->inBytesPerFrame) return FHC_ERR_ARG;      // nothing resident at that sample size
  rc = launch(c, bt, 0, n, sampleBytes, per_frame);
  if (rc != FHC_OK) return rc;
  CK(cudaEventSynchronize(bt->evStop));
  if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, bt->evStart, bt->evStop));
  return FHC_OK;
}

int fhc_last_kernel_ms(fhc_ctx* c, float* kernel_ms)
{
  if (!c || !kernel_ms) return FHC_ERR_ARG;
  if (c->submits == 0) return FHC_ERR_STATE;
  CK(cudaSetDevice(c->device));
  Batch* bt = &c->batch[0];
  for (int i = 1; i < FHC_BATCHES; i++) if (c->batch[i].serial > bt->serial) bt = &c->batch[i];
  CK(cudaEventSynchronize(bt->evStop));
  CK(cudaEventElapsedTime(kernel_ms, bt->evStart, bt->evStop));
  return FHC_OK;
}

int fhc_pictures_in_flight(int device)
{
  int ndev = 0, maxBlocks = 0;
  size_t smemBytes = 0, ws = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return FHC_ERR_ARG;
  CK(cudaSetDevice(device));
  if (kernel_geometry(device, &smemBytes, &ws, &maxBlocks) != FHC_OK) return FHC_ERR_CUDA;
  return maxBlocks * FHC_PICS_PER_BLOCK;
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

// debug: launch geometry {resident blocks, threads per block, dynamic shared memory bytes, global workspace bytes per block}
// diagnostics: how many of n pictures with the given slice QPs the scheduler would run alone in their block on `blocks` blocks
// (host arithmetic only: no device needed)
extern "C" int fhc_debug_lone_pictures(const int* qps, int n, int blocks)
{
  if (!qps || n <= 0 || blocks <= 0) return FHC_ERR_ARG;
  std::vector<fhc_slice_params> sp((size_t)n);
  std::vector<int> order((size_t)n);
  for (int i = 0; i < n; i++) { memset(&sp[(size_t)i], 0, sizeof(fhc_slice_params)); sp[(size_t)i].sliceQp = qps[i]; order[(size_t)i] = i; }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return sp[(size_t)a].sliceQp < sp[(size_t)b2].sliceQp; });
  return lone_pictures(sp.data(), order, n, blocks < n ? blocks : n);
}

extern "C" int fhc_debug_geometry(fhc_ctx* c, long long* out4)
{
  if (!c || !out4) return FHC_ERR_ARG;
  out4[0] = c->maxBlocks; out4[1] = FHC_BLOCK_THREADS; out4[2] = (long long)c->smemBytes * FHC_PICS_PER_BLOCK; out4[3] = (long long)c->wsPerBlock * FHC_PICS_PER_BLOCK;
  return FHC_OK;
}

// ----------------------------------------------------------------------------------------------
// integer-issue roof (the binding roofline of this path is int32 instruction issue, and
// MEASURED_PEAKS.json has no such figure): a microbenchmark of independent IMAD / IADD3 / LOP3 chains, one third
// each, at full occupancy.  Reported in thread-level integer operations per second.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fhc_int_issue_kernel(int* out, int iters, int seed)
{
  int a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 ^ 0x55, a3 = a0 + 7, b0 = a0 * 5, b1 = a1 * 7, b2 = a2 + 11, b3 = a3 ^ 0x33;
  int c0 = a0 - 9, c1 = a1 + 13, c2 = a2 * 3, c3 = a3 - 5;
  const int m = (blockIdx.x | 1) + seed;
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int u = 0; u < 8; u++)
    {
      asm volatile("mad.lo.s32 %0, %0, %4, %1;\n\tmad.lo.s32 %1, %1, %4, %2;\n\tmad.lo.s32 %2, %2, %4, %3;\n\tmad.lo.s32 %3, %3, %4, %0;"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3) : "r"(m));
      asm volatile("{ .reg .s32 t;\n\tadd.s32 t, %0, %1;\n\tadd.s32 %0, t, %4;\n\tadd.s32 t, %1, %2;\n\tadd.s32 %1, t, %4;\n\t"
                   "add.s32 t, %2, %3;\n\tadd.s32 %2, t, %4;\n\tadd.s32 t, %3, %0;\n\tadd.s32 %3, t, %4; }"
                   : "+r"(b0), "+r"(b1), "+r"(b2), "+r"(b3) : "r"(m));
      asm volatile("lop3.b32 %0, %0, %1, %4, 0x96;\n\tlop3.b32 %1, %1, %2, %4, 0xe8;\n\tlop3.b32 %2, %2, %3, %4, 0x96;\n\tlop3.b32 %3, %3, %0, %4, 0xe8;"
                   : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3) : "r"(m));
    }
  }
  const int v = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3 ^ c0 ^ c1 ^ c2 ^ c3;
  if (v == 0x7fffffff) out[0] = v;
}

// *gops = measured integer operations per second / 1e9 (each IMAD, IADD3 and LOP3 counts as one operation per thread)
extern "C" int fhc_measure_int_issue(int device, double* gops, double* ms_out)
{
  if (!gops) return FHC_ERR_ARG;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  int* dOut = nullptr;
  CK(cudaMalloc(&dOut, sizeof(int)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double best = 0, bestMs = 0;
  for (int rep = 0; rep < 5; rep++)
  {
    CK(cudaEventRecord(e0, 0));
    fhc_int_issue_kernel<<<blocks, threads>>>(dOut, iters, rep);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // per thread and iteration: 8 x (4 IMAD + 4 IADD3 + 4 LOP3); a 3-input add.s32 pair fuses into one IADD3
    const double ops = (double)blocks * threads * (double)iters * 8.0 * 12.0;
    const double g = ops / (ms * 1e-3) / 1e9;
    if (rep > 0 && g > best) { best = g; bestMs = ms; }
    g_launches++;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dOut);
  *gops = best;
  if (ms_out) *ms_out = bestMs;
  return FHC_OK;
}

// debug: cycle counters of picture `poc` (FHC_PROFILE builds; zeros otherwise): out64[role * 8 + stage], out64[40] = picture total
extern "C" int fhc_debug_profile(fhc_ctx* c, int poc, long long* out64)
{
  if (!c || !out64) return FHC_ERR_ARG;
  int slot = 0;
  Batch* bt = find_batch(c, poc, &slot);
  if (!bt) return FHC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(bt->evStop));
  CK(cudaMemcpy(out64, bt->dProf + 64 * (size_t)slot, sizeof(long long) * 64, cudaMemcpyDeviceToHost));
  return FHC_OK;
}
