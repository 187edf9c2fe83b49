// fhc_core.cuh - device code of the B200 HEVC intra mode-decision engine.
//
// A LANE GROUP PER PICTURE.  All-intra pictures are independent; inside a picture the
// CABAC context chain of TEncSlice::compressSlice forces raster CTU order
// (SURVEY.md F5), so a picture is one long dependent chain and the GPU is filled
// by running many pictures at once: persistent warps pull pictures from a queue
// and every warp encodes 32 / FHC_GROUP pictures side by side, FHC_GROUP lanes
// each.  The groups of a warp start their pictures together and follow the same
// (almost static) decision schedule, so they share one instruction stream - the
// kernel is instruction-fetch bound, not ALU or HBM bound (profiles/).
//
// Inside a group the LEADER lane walks the decision recursion (quad-tree,
// candidate loops, RQT, RDOQ's sequential pass, CABAC bit counting) on state kept
// in the group's shared-memory slice.  The other lanes serve it:
//   * COOP sections (par / run_phase): sample-parallel stages of one block
//     (reference build, 35-mode prediction + SATD, residual, DCT/DST, RDOQ
//     pre-quantisation, dequant + inverse, reconstruction, SSE, copies) are striped
//     over the group's lanes and fenced with __syncwarp(group mask);
//   * LANE sections (dispatch_tasks): independent candidates (full-RD luma modes
//     of small PUs, chroma modes of 8x8 CUs) are evaluated one per lane on
//     lane-private scratch, then compared in list order.
//
// The file is written against a small portability layer (FHC_DEVICE_BUILD) so
// that tests can compile the very same source for the host and step through it
// without a GPU ("emulation build", tests/ only - the product library never
// contains it and has no CPU path).
//
// Reference citations are to /root/reference/source/Lib (HM 16.14).
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../include/fasthevc_cuda.h"

#ifdef FHC_DEVICE_BUILD
#define FHC_HD __host__ __device__ __forceinline__
#define FHC_DH __device__ __forceinline__
#define FHC_D __device__ __noinline__   /* one copy of every sizeable function: the kernel is instruction-cache bound otherwise */
#define FHC_NOINLINE __device__ __noinline__
#define FHC_NOUNROLL _Pragma("unroll 1")
#else
#define FHC_HD static inline
#define FHC_DH static inline
#define FHC_D static
#define FHC_NOINLINE static
#define FHC_NOUNROLL
#endif

namespace fhc {

// Work counters of the host emulation build (tests/tools only; SURVEY.md 8d asks for the exact sample counts per
// configuration): samples through the TU pipeline by component class and TU size, samples predicted + SATD'd by the
// rough mode decision, and CABAC bins counted.  Compiled out of the device build.
#ifndef FHC_DEVICE_BUILD
struct EmuCounters { unsigned long long tqSamples[2][4], tuCalls[2][4], rmdSamples, bins; };
static EmuCounters g_emuCounters;
#define FHC_COUNT(stmt) do { stmt; } while (0)
#else
#define FHC_COUNT(stmt) do { } while (0)
#endif

// ----------------------------------------------------------------------------------------------
// constants
// ----------------------------------------------------------------------------------------------
enum { MODE_INTRA = 1, MODE_NONE = 2 };
enum { SIZE_2Nx2N = 0, SIZE_NxN = 3, SIZE_NONE = 8 };
enum { PLANAR_IDX = 0, DC_IDX = 1, HOR_IDX = 10, VER_IDX = 26, DM_CHROMA_IDX = 36 };
enum { SCAN_DIAG = 0, SCAN_HOR = 1, SCAN_VER = 2 };
enum { CI_CURR_BEST = 0, CI_NEXT_BEST, CI_TEMP_BEST, CI_CHROMA_INTRA, CI_QT_TRAFO_TEST, CI_QT_TRAFO_ROOT, CI_NUM };
enum {
  CTX_SPLIT = 0, CTX_PART = 3, CTX_IPRED = 4, CTX_CPRED = 5, CTX_CBF = 7, CTX_SUBDIV = 17, CTX_SIGCG = 20, CTX_SIG = 24,
  CTX_LASTX = 68, CTX_LASTY = 98, CTX_ONE = 128, CTX_ABS = 152, CTX_TS = 158, NUM_CTX = 160
};
#define FHC_MAX_COST 1.7e+308
#ifndef FHC_WARPS_PER_BLOCK
#define FHC_WARPS_PER_BLOCK 20   /* warps (= pictures, FHC_GROUP 32) per thread block; the block keeps them aligned (measured: 16 -> 4495, 20 -> 4833, 24 -> 4643 CTU/s) */
#endif
#ifndef FHC_GROUP
#define FHC_GROUP 32             /* lanes per picture: a warp encodes 32 / FHC_GROUP pictures side by side (sub-warp groups serialise) */
#endif
#define FHC_MAX_MODES 10         /* the longest candidate list (8 RMD survivors + 2 MPMs) */
#define FHC_MAX_TASKS 20         /* ... times the transform-skip pair of a 4x4 PU */
#define FHC_NENG (FHC_GROUP < FHC_MAX_TASKS ? FHC_GROUP : FHC_MAX_TASKS)   /* lane engines per picture */
#define FHC_SMALL_COEFS 64       /* lane engines: largest TU (8x8) */
#ifndef FHC_MASTER_SMALL
#define FHC_MASTER_SMALL 16      /* leader: TUs up to this many coefficients use the warp's shared-memory scratch, larger ones the global one */
#endif
#ifndef FHC_SLOT_SMEM_DEPTH
#define FHC_SLOT_SMEM_DEPTH 2     /* context-set slots of depth >= this live in shared memory, shallower ones in global memory */
#endif
#define FHC_FULL_MASK 0xffffffffu

// layout of the 6144-sample "YUV" buffers: Y at 0, Cb at 4096, Cr at 5120
FHC_DH int comp_off(int comp) { return comp == 0 ? 0 : (comp == 1 ? 4096 : 5120); }

// ----------------------------------------------------------------------------------------------
// tables (built on the host by fhc_build_tables, copied to shared memory once per block)
// ----------------------------------------------------------------------------------------------
struct Tables
{
  uint8_t  z2r[256], r2z[256];
  uint16_t scan[1520];            // diag: 4,8,16,32 ; hor: 4,8 ; ver: 4,8
  uint16_t scanOff[3][4];
  uint8_t  scanCG[3][88];         // per type: 1,4,16,64 entries at offsets 0,1,5,21
  int16_t  T32[32 * 32];
  int32_t  entropyBits[128];
  uint8_t  nextState[128][2];
  uint8_t  ctxInit[NUM_CTX];
  int16_t  dst4[16];              // 4x4 DST-VII (TComRom.cpp g_as_DST_MAT_4)
  int16_t  angTable[9], invAngTable[9];
  uint8_t  chromaScale[58];       // g_aucChromaScale[CHROMA_420]
  uint8_t  ctxIndMap4x4[16];
  uint8_t  sigCnt[4][16];         // getSigCtxInc neighbourhood term by (patternSigCtx, position inside the 4x4 group)
  uint8_t  intraFilterThr[5];
  double   errScale[2][4][6];     // [chType][log2-2][qp rem]
};
FHC_HD int cg_off(int l) { return l == 0 ? 0 : (l == 1 ? 1 : (l == 2 ? 5 : 21)); }

struct Sbac { uint8_t c[NUM_CTX]; uint64_t frac; };

struct CUMeta
{
  int zidx, x, y, depth, nparts, width;
  uint32_t dist, bits, bins, pad;
  double cost;
  // per-partition arrays (z-scan inside the CU); storage is sized by depth: 256 >> 2*depth entries each
  uint8_t *dep, *part, *pred, *ldir, *cdir, *tridx;
  uint8_t* cbf[3];
  uint8_t* ts[3];
};
FHC_HD void cu_bind(CUMeta* cu, uint8_t* store, int cap)
{
  cu->dep = store; cu->part = store + cap; cu->pred = store + 2 * cap; cu->ldir = store + 3 * cap; cu->cdir = store + 4 * cap;
  cu->tridx = store + 5 * cap;
  for (int c = 0; c < 3; c++) { cu->cbf[c] = store + (6 + c) * cap; cu->ts[c] = store + (9 + c) * cap; }
}

enum Op {
  OP_EXIT = 0, OP_REFS, OP_RMD, OP_PRED_RESI, OP_FWD_QUANT, OP_LEVELS_OUT, OP_INV_RECON, OP_COPY2D, OP_COPY16, OP_ZERO16,
  OP_CTU_RECO_OUT, OP_CU_TO_OUT, OP_NZMASK, OP_TASKS, OP_BLOCKBAR
};
enum { ENG_MASTER = 0, ENG_LANE = 1 };
enum { TASK_LUMA_PU = 0, TASK_CHROMA_CU = 1, TASK_CHROMA_TU = 2, TASK_CHROMA_BITS = 3 };

struct Cmd
{
  int op;
  int a[16];
  const void* src;
  void* dst;
};

// per-coefficient-group scratch of RDOQ / coefficient coding, sized by the number of 4x4 groups of the TU
template <int N>
struct GroupScratch
{
  double costCGSig[N];
  uint8_t cgFlag[N], cgFlagCode[N];
  uint16_t cgNzMask[N];
  uint8_t cgAbs[N];
};

// per-coefficient scratch of one engine (TU pipeline + RDOQ), sized by the engine's largest TU
template <int N>
struct CoefScratch
{
  int16_t resi[N];
  int32_t tmp[N], tcoef[N], lvlD[N];
  double cost0[N], costCoeff[N], costSig[N];
  int32_t rateIncUp[N], rateIncDown[N], sigRateDelta[N], deltaU[N];
};

// One evaluation engine.  ENG_MASTER: the warp itself (COOP sections use all lanes); ENG_LANE: one lane evaluating one
// independent candidate on private scratch, every stage executed by that lane alone.
struct Engine
{
  int kind;
  uint32_t bins, sse;
  int refDc;
  double rdoqLambda;
  const Tables* t;
  Sbac* go;                       // the coder under work (shared memory: every bin reads and steps a context)
  Sbac* slots;                    // master: [5][CI_NUM]; lane engine: [3] shared by all depths (TU-level slots only)
  Cmd cmd;
  int16_t *refU, *refF;
  uint8_t* flags;
  int16_t* resi;
  int32_t *tmp, *tcoef, *lvlD;
  double *cost0, *costCoeff, *costSig;
  int32_t *rateIncUp, *rateIncDown, *sigRateDelta, *deltaU;
  int16_t tsCoef[3][16], tsRec[3][16], tsPred[3][16];
  void *csSmall, *csLarge, *gsSmall, *gsLarge;        // coefficient scratch by TU size (master: <= 8x8 shared / larger global; lane: 4x4 shared / 8x8 global)
  // engine buffers: prediction/reconstruction of the CU under test and the per-TU-size layers; layerSize samples per
  // layer, component c at coff[c], luma rows qs apart (chroma qs / 2).  Master: 6144 / {0,4096,5120} / 64 (a CTU);
  // lane engine: 96 / {0,64,80} / 8 (an 8x8 CU, dense)
  int16_t* predYuv;
  int16_t *qtRecBase, *qtCoefBase;
  int layerSize, qs, coff[3];
  int writePic;                   // reconstruct into the picture plane too (master only)
  int refsShared;                 // 1: luma TUs use the reference lines the leader built for the rough mode decision; 2: chroma
                                  // TUs of an 8x8 CU use the lines the leader built before the dispatch (Shared::refC)
  int onlyComp;                   // != 0: recur_chroma_qt codes only this component (lane tasks that split Cb / Cr)
  int forceTs;                    // >= 0: evaluate only this transform_skip_flag of a 4x4 luma TU (lane tasks of a 4x4 PU)
  int noAlign;                    // inside a section whose TU count depends on the picture: no block alignment
  uint32_t cgNZ[2];               // per coefficient group (scan order): any pre-quantised level != 0 (set by OP_FWD_QUANT)
  double* costCGSig;              // RDOQ per-group / per-context scratch
  uint8_t *cgFlag, *cgFlagCode;
  double *sigC0, *sigC1;
  int32_t* sigDelta;
  uint16_t* cgNzMask;             // per coefficient group (scan order): non-zero map of the levels
  uint8_t* cgAbs;                 // ... and the parity of their sum
  // engine result: best task of the current dispatch
  double resCost;
  uint32_t resDist;
  int resTask;
};

template <int N>
FHC_DH void bind_group_scratch(Engine* e, GroupScratch<N>* x)
{
  e->costCGSig = x->costCGSig; e->cgFlag = x->cgFlag; e->cgFlagCode = x->cgFlagCode; e->cgNzMask = x->cgNzMask; e->cgAbs = x->cgAbs;
}
FHC_DH int16_t* qt_rec(const Engine* e, int layer) { return e->qtRecBase + layer * e->layerSize; }
FHC_DH int16_t* qt_coef(const Engine* e, int layer) { return e->qtCoefBase + layer * e->layerSize; }

// What one lane engine leaves behind for the leader (global memory) + the scratch that is too large for the lane's
// stack.  An engine runs its tasks in increasing order and keeps the best (strict <).
struct EngineGlobal
{
  double resCost;
  uint32_t resDist;
  int resTask;
  int16_t bestCoef[3][64], bestReco[3][64];
  uint8_t bestTrIdx[4], bestCbf[3][4], bestTs[3][4];
  CoefScratch<FHC_SMALL_COEFS> cs;     // 8x8 TUs
  double sigC0[28], sigC1[28];
  int32_t sigDelta[28];
};

// Working state of one lane engine: a local variable of run_task, i.e. in the lane's LOCAL memory, where the same
// field of neighbouring lanes is adjacent - the lanes of a dispatch walk the same code, so their accesses coalesce
// (private global structs cost one 32-byte sector per lane and access).  Engines only see depth-3 CUs: 96 samples.
struct LaneLocal
{
  Engine eng;
  Sbac go, slots[3];              // slots: CI_QT_TRAFO_ROOT, CI_TEMP_BEST, CI_QT_TRAFO_TEST
  int16_t refU[36], refF[36];
  uint8_t flags[24];
  CoefScratch<16> cs16;           // 4x4 TUs
  double costCGSig[4];
  uint8_t cgFlag[4], cgFlagCode[4];
  uint16_t cgNzMask[4];
  uint8_t cgAbs[4];
  CUMeta cu;
  uint8_t cuStore[12 * 4];
  int16_t predYuv[96];
  int16_t qtRec[4][96];
  int16_t qtCoef[4][96];
};

// picture descriptor (global memory, filled by the host per submitted picture)
struct PicDesc
{
  const void* org[3];             // input planes (sampleBytes each)
  int16_t* rec[3];                // reconstruction planes, stride = plane width
  fhc_ctu_result* out;            // [numCtus]
  fhc_slice_params sp;
  int32_t width, height, bitDepth, sampleBytes, wc, hc;
  long long prof[16], profN[16];  // cycle counters of the picture's warp (FHC_PROFILE builds)
  long long prof2[16], prof2N[16];
};

// per-warp state in global memory: the picture being encoded + scratch that does not fit the warp's shared-memory slice
struct FrameGlobal
{
  const void* org[3];
  int16_t* rec[3];
  fhc_ctu_result* out;
  fhc_slice_params sp;
  int32_t width, height, bitDepth, sampleBytes, wc, hc;
  long long prof[16], profN[16];
  long long prof2[16], prof2N[16];
  int16_t cuCoef[4][2][6144];     // CU coefficient buffers per depth, best/temp
  int16_t cuReco[4][2][6144];     // CU reconstruction buffers
  int16_t predYuv[6144];          // prediction -> reconstruction scratch of the CU under test
  int16_t qtRec[4][6144];         // per-TU-size reconstruction layers (Y stride 64, C stride 32)
  int16_t qtCoef[4][6144];        // per-TU-size coefficient layers
  int16_t rmdPred[4096];          // one group of predicted modes of the rough mode decision
  uint8_t cuStoreBig[2 * 12 * (256 + 64)];   // CU metadata of depth 0 and 1
  uint8_t qtTrIdx[256], qtCbf[3][256], qtTs[3][256];   // best-so-far TU flags of the PU under test
  CoefScratch<1024> big;          // TU pipeline + RDOQ scratch of 16x16 / 32x32 TUs
  GroupScratch<64> bigGs;
  Sbac slotLow[FHC_SLOT_SMEM_DEPTH + 1][CI_NUM];
  EngineGlobal eng[FHC_NENG];
};

struct TaskDesc
{
  int kind, numTasks;
  int depth, bufIdx;              // the CU under test: s->cu[depth][bufIdx]
  int partOffset, log2pu, initTrDepth;
  int resBase;                     // lane engine that receives the result of task 0
  int tsSplit;                     // 4x4 PU: task t = candidate t >> 1 with transform_skip_flag = t & 1
  int mode[FHC_MAX_MODES];
};

struct FrameParams                 // hot read-only copy of the picture header (shared memory)
{
  int32_t width, height, bitDepth, sampleBytes, wc, hc;
  fhc_slice_params sp;
};

// the warp's shared-memory slice
struct Shared
{
  const Tables* t;
  int alignThreads;               // threads of the block that take part in block_align (32 x pictures of this round)
  FrameParams fp;
  Sbac slot[5 - FHC_SLOT_SMEM_DEPTH][CI_NUM];   // depth FHC_SLOT_SMEM_DEPTH.. ; shallower: FrameGlobal::slotLow
  int ctuAddr, ctuX, ctuY, ctuPelX, ctuPelY;
  uint8_t bestIdx[4];
  CUMeta cu[4][2];
  uint8_t cuStore[2 * 12 * (16 + 4)];        // CU metadata of depth 2 and 3 (depth 0 and 1: FrameGlobal::cuStoreBig)
  Engine master;
  TaskDesc task;
  Sbac masterGo;
  // master scratch
  int16_t refU[260], refF[260];
  uint8_t flags[72];
  union                          // a dispatch and a COOP TU of the master never overlap in time
  {
    CoefScratch<FHC_MASTER_SMALL> cs;         // master: small TUs
  } u;
  uint32_t satd[36];
  int16_t refC[2][20];            // unfiltered reference lines of the 4x4 Cb / Cr block of an 8x8 CU
  int dcC[2];
  double sigC0[28], sigC1[28];
  int32_t sigDelta[28];
  GroupScratch<4> gs;             // per-coefficient-group scratch of TUs up to 8x8 (larger: FrameGlobal::bigGs)
};

// lane/nl: this thread's index and the size of the lane group that executes COOP sections together (32 = the warp,
// 1 = a lane engine on its own); mask = the group's __syncwarp mask
struct Ctx { FrameGlobal* g; Shared* s; Engine* e; int lane, nl; unsigned mask; };

#if defined(FHC_DEVICE_BUILD) && defined(FHC_PROFILE)
#define FHC_TIC(k) const long long tic_##k = clock64()
#define FHC_TOC3(c, nm, k) do { if ((c).e->kind == ENG_MASTER && (c).lane == 0) { (c).g->prof2[k] += clock64() - tic_##nm; (c).g->prof2N[k]++; } } while (0)
#define FHC_TOC2(c, nm, k) do { if ((c).e->kind == ENG_MASTER && (c).lane == 0) { (c).g->prof[k] += clock64() - tic_##nm; (c).g->profN[k]++; } } while (0)
#else
#define FHC_TIC(k) do { } while (0)
#define FHC_TOC3(c, nm, k) do { } while (0)
#define FHC_TOC2(c, nm, k) do { } while (0)
#endif
enum { PF_TOTAL = 0, PF_PAR, PF_RDOQ, PF_COEFBITS, PF_REFS_MASTER, PF_RMD_MASTER, PF_TUBLOCK, PF_BITSQT, PF_CUCOPY, PF_RMD_PAR, PF_ENCODE, PF_TASKS_LUMA, PF_TASKS_CHROMA, PF_LUMA, PF_CHROMA, PF_NUM = 16 };
enum { PF2_CHECK_D0 = 0, PF2_CHECK_D1, PF2_CHECK_D2, PF2_CHECK_D3_2N, PF2_CHECK_D3_NXN, PF2_LUMA2ND, PF2_CHROMA_BIG, PF2_LUMA1ST_BIG, PF2_NUM = 16 };


// ----------------------------------------------------------------------------------------------
// portability layer: COOP sections
// ----------------------------------------------------------------------------------------------
FHC_D void run_phase(const Ctx& c, int op, int phase, int w, int nw);
FHC_D void dispatch_tasks(const Ctx& c, int align);
FHC_D void run_tasks(const Ctx& c);
FHC_DH int num_phases(int op)
{
  switch (op)
  {
  case OP_REFS: return 2;
  case OP_RMD: return 2;            // per mode group, see OP_RMD
  case OP_FWD_QUANT: return 3;
  case OP_INV_RECON: return 4;
  default: return 1;
  }
}

#ifdef FHC_DEVICE_BUILD
FHC_DH void group_sync(const Ctx& c) { if (c.nl > 1) __syncwarp(c.mask); }
// A COOP section, called by the group's LEADER (lane 0 of the group) only: the followers sit in follower_loop.  The
// leader publishes the command (its arguments are already in e->cmd), every phase is striped over the group's lanes
// and fenced.  Nothing in the decision code is executed redundantly, so no lane ever depends on running in lockstep
// with another one.  A lane engine (nl == 1) runs the phases inline.
FHC_DH void par(const Ctx& c, int op)
{
  FHC_TIC(par);
  const int np = (op == OP_RMD) ? c.e->cmd.a[15] : num_phases(op);
  if (c.nl > 1)
  {
    c.e->cmd.op = op;
    __syncwarp(c.mask);
  }
  FHC_NOUNROLL for (int ph = 0; ph < np; ph++)
  {
    run_phase(c, op, ph, c.lane, c.nl);
    group_sync(c);
  }
  FHC_TOC2(c, par, PF_PAR);
}
// lanes 1.. of a group: execute the leader's COOP sections and LANE dispatches until it publishes OP_EXIT
FHC_D void follower_loop(const Ctx& c)
{
  for (;;)
  {
    __syncwarp(c.mask);
    const int op = c.e->cmd.op;
    if (op == OP_EXIT) break;
    if (op == OP_TASKS) { run_tasks(c); __syncwarp(c.mask); continue; }
    if (op == OP_BLOCKBAR) { asm volatile("barrier.sync 1, %0;" ::"r"(c.s->alignThreads) : "memory"); continue; }
    const int np = (op == OP_RMD) ? c.e->cmd.a[15] : num_phases(op);
    FHC_NOUNROLL for (int ph = 0; ph < np; ph++)
    {
      run_phase(c, op, ph, c.lane, c.nl);
      __syncwarp(c.mask);
    }
  }
}
// BLOCK ALIGNMENT.  The kernel is bound by instruction delivery (profiles/): warps that run different routines evict
// each other's code from the SM's instruction caches and the fill path from L2 saturates chip-wide.  All pictures of
// one size walk the same static schedule of routines (candidate counts are padded, see ghost_first_pass), so the
// pictures of a thread block are aligned at the entry of every routine of that schedule with a named barrier: the
// block's warps then execute the same code at the same time.  Sections whose TU count depends on the picture
// (chroma trees, the final CTU coding) set noAlign.  Lane engines never align.
FHC_DH void block_align(const Ctx& c)
{
#ifndef FHC_NO_ALIGN
  if (c.e->kind != ENG_MASTER || c.e->noAlign) return;
  const int n = c.s->alignThreads;
  if (n <= 32) return;
  c.e->cmd.op = OP_BLOCKBAR;
  __syncwarp(c.mask);
  asm volatile("barrier.sync 1, %0;" ::"r"(n) : "memory");
#else
  (void)c;
#endif
}
#ifndef FHC_ALIGN_LEVEL
#define FHC_ALIGN_LEVEL 2        /* 1: routine entries only (-11 %); 2: + RDOQ and bit counting inside a TU */
#endif
#if FHC_ALIGN_LEVEL >= 2
FHC_DH void mid_align(const Ctx& c) { block_align(c); }
#else
FHC_DH void mid_align(const Ctx& c) { (void)c; }
#endif
#ifdef FHC_EXTRA_ALIGN
FHC_DH void extra_align(const Ctx& c) { block_align(c); }
#else
FHC_DH void extra_align(const Ctx& c) { (void)c; }
#endif
#ifndef FHC_FINE_ALIGN
#define FHC_FINE_ALIGN 1         /* also align before every stage of a TU (measured at 20 warps per block: +3.6 %; neutral at 8) */
#endif
#if FHC_FINE_ALIGN
FHC_DH void fine_align(const Ctx& c) { block_align(c); }
#else
FHC_DH void fine_align(const Ctx& c) { (void)c; }
#endif
FHC_DH void release_followers(const Ctx& c)
{
  if (c.nl > 1) { c.e->cmd.op = OP_EXIT; __syncwarp(c.mask); }
}
#else
FHC_DH void group_sync(const Ctx& c) { (void)c; }
FHC_DH void block_align(const Ctx& c) { (void)c; }
FHC_DH void mid_align(const Ctx& c) { (void)c; }
FHC_DH void fine_align(const Ctx& c) { (void)c; }
FHC_DH void extra_align(const Ctx& c) { (void)c; }
// emulation: the group's lanes are executed one after the other, phase by phase
FHC_DH void par(const Ctx& c, int op)
{
  const int np = (op == OP_RMD) ? c.e->cmd.a[15] : num_phases(op);
  for (int ph = 0; ph < np; ph++)
    for (int w = 0; w < c.nl; w++) run_phase(c, op, ph, w, c.nl);
}
#endif

FHC_DH int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
FHC_DH int iabs(int v) { return v < 0 ? -v : v; }
FHC_DH int imin(int a, int b) { return a < b ? a : b; }
FHC_DH int imax(int a, int b) { return a > b ? a : b; }
FHC_DH int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }
#ifdef FHC_DEVICE_BUILD
FHC_DH int msb_index(uint32_t v) { return 31 - __clz((int)v); }
FHC_DH int lsb_index(uint32_t v) { return __ffs((int)v) - 1; }
#else
FHC_DH int msb_index(uint32_t v) { return 31 - __builtin_clz(v); }
FHC_DH int lsb_index(uint32_t v) { return __builtin_ctz(v); }
#endif

FHC_DH int org_at(const FrameGlobal* g, int sampleBytes, int comp, int idx)
{
  return sampleBytes == 1 ? (int)((const uint8_t*)g->org[comp])[idx] : (int)((const uint16_t*)g->org[comp])[idx];
}
FHC_DH int plane_stride(const Shared* s, int comp) { return s->fp.width >> (comp ? 1 : 0); }
FHC_DH const uint16_t* scan_tab(const Tables& t, int type, int log2n) { return t.scan + t.scanOff[type][log2n - 2]; }
FHC_DH const uint8_t* scan_cg(const Tables& t, int type, int log2n) { return t.scanCG[type] + cg_off(log2n - 2); }
// the NxN transform matrix as a strided view: M(k, j) = tab[k * rowStride + j] (4x4 DST-VII or rows of the 32-point DCT)
struct TMat { const int16_t* tab; int rowStride; };
FHC_DH TMat tmat(const Tables& t, int n, int useDst)
{
  TMat m;
  if (useDst) { m.tab = t.dst4; m.rowStride = 4; }
  else { m.tab = t.T32; m.rowStride = (n == 4 ? 8 : n == 8 ? 4 : n == 16 ? 2 : 1) * 32; }
  return m;
}
FHC_DH int tcoefm(const Tables& t, int n, int k, int j, int useDst)
{
  // 4x4 DST-VII matrix or rows of the 32-point DCT matrix
  if (useDst) return t.dst4[k * 4 + j];
  return t.T32[(k * (32 / n)) * 32 + j];
}

// ----------------------------------------------------------------------------------------------
// intra prediction of one sample (TComPrediction.cpp:229-473, :731-817), closed forms of HM's loops
// line: 4n+1 reference samples, index 2n = corner, growing index = up the left column then along the top
// ----------------------------------------------------------------------------------------------
FHC_DH int pred_sample(const Tables& t, const int16_t* line, int n, int log2n, int mode, int x, int y, int isLuma, int bd, int dcval)
{
  const int c = 2 * n;
  if (mode == PLANAR_IDX)
  {
    const int left = line[c - 1 - y], top = line[c + 1 + x], topRight = line[c + 1 + n], bottomLeft = line[c - 1 - n];
    return ((n - 1 - x) * left + (x + 1) * topRight + (n - 1 - y) * top + (y + 1) * bottomLeft + n) >> (log2n + 1);
  }
  if (mode == DC_IDX)
  {
    if (isLuma && n <= 16)
    {
      if (x == 0 && y == 0) return (line[c + 1] + line[c - 1] + 2 * dcval + 2) >> 2;
      if (y == 0) return (line[c + 1 + x] + 3 * dcval + 2) >> 2;
      if (x == 0) return (line[c - 1 - y] + 3 * dcval + 2) >> 2;
    }
    return dcval;
  }
  {
    const int isVer = mode >= 18;
    const int angMode = isVer ? mode - VER_IDX : -(mode - HOR_IDX);
    const int absA = iabs(angMode);
    const int angle = angMode < 0 ? -t.angTable[absA] : t.angTable[absA];
    const int invAngle = t.invAngTable[absA];
    const int s = isVer ? 1 : -1;                 // main(k) = line[c + s*k], side(j) = line[c - s*j]
    const int u = isVer ? x : y, v = isVer ? y : x;
    if (angle == 0)
    {
      int val = line[c + s * (u + 1)];
      if (isLuma && n <= 16 && u == 0) val = clip3(0, (1 << bd) - 1, val + ((line[c - s * (v + 1)] - line[c]) >> 1));
      return val;
    }
    const int deltaPos = (v + 1) * angle, deltaInt = deltaPos >> 5, fract = deltaPos & 31;
    const int k0 = u + deltaInt + 1;
    int a = (k0 >= 0) ? line[c + s * k0] : line[c - s * ((128 + (-k0) * invAngle) >> 8)];
    if (!fract) return a;
    const int k1 = k0 + 1;
    int b = (k1 >= 0) ? line[c + s * k1] : line[c - s * ((128 + (-k1) * invAngle) >> 8)];
    return ((32 - fract) * a + fract * b + 16) >> 5;
  }
}

// TComPrediction::filteringIntraReferenceSamples (TComPattern.cpp:541)
FHC_DH int use_filtered_refs(const Tables& t, int comp, int mode, int log2n)
{
  if (comp != 0 || mode == DC_IDX) return 0;
  const int diff = imin(iabs(mode - HOR_IDX), iabs(mode - VER_IDX));
  return diff > t.intraFilterThr[log2n - 2];
}

// 8x8 / 4x4 Hadamard SATD of (org - pred) (TComRdCost.cpp:1549-1750)
FHC_DH uint32_t had8(int* m)
{
  int s = 0;
  for (int i = 0; i < 8; i++)
    for (int k = 1; k < 8; k <<= 1)
      for (int j = 0; j < 8; j++)
        if (!(j & k)) { int a = m[i * 8 + j], b = m[i * 8 + j + k]; m[i * 8 + j] = a + b; m[i * 8 + j + k] = a - b; }
  for (int j = 0; j < 8; j++)
    for (int k = 1; k < 8; k <<= 1)
      for (int i = 0; i < 8; i++)
        if (!(i & k)) { int a = m[i * 8 + j], b = m[(i + k) * 8 + j]; m[i * 8 + j] = a + b; m[(i + k) * 8 + j] = a - b; }
  for (int i = 0; i < 64; i++) s += iabs(m[i]);
  return (uint32_t)((s + 2) >> 2);
}
FHC_DH uint32_t had4(int* d)
{
  int m[16], s = 0;
  for (int i = 0; i < 4; i++)
  {
    int a = d[i * 4] + d[i * 4 + 3], b = d[i * 4 + 1] + d[i * 4 + 2], cc = d[i * 4 + 1] - d[i * 4 + 2], dd = d[i * 4] - d[i * 4 + 3];
    m[i * 4] = a + b; m[i * 4 + 1] = a - b; m[i * 4 + 2] = dd + cc; m[i * 4 + 3] = dd - cc;
  }
  for (int j = 0; j < 4; j++)
  {
    int a = m[j] + m[12 + j], b = m[4 + j] + m[8 + j], cc = m[4 + j] - m[8 + j], dd = m[j] - m[12 + j];
    s += iabs(a + b) + iabs(a - b) + iabs(dd + cc) + iabs(dd - cc);
  }
  return (uint32_t)((s + 1) >> 1);
}

// accumulations from a striped phase: atomic when several lanes of a group contribute, plain for a lane engine
FHC_DH void atomic_or_u32(uint32_t* p, uint32_t v, int nw)
{
#ifdef FHC_DEVICE_BUILD
  if (nw > 1) { atomicOr(p, v); return; }
#endif
  (void)nw;
  *p |= v;
}
FHC_DH void atomic_add_u32(uint32_t* p, uint32_t v, int nw)
{
#ifdef FHC_DEVICE_BUILD
  if (nw > 1) { atomicAdd(p, v); return; }
#endif
  (void)nw;
  *p += v;
}

// One stage of a separable NxN transform executed by ONE lane (lane engines, N <= 8): out[k][j] = (sum_q M(k,q) * in[j][q]
// + add) >> shift with the input row held in registers (N loads per row instead of N per output).  inverse: M(q,k)
// and clipping to 16 bits (TComTrQuant.cpp:860-985).  Same sums in the same order as the striped phases.
template <typename TIn, typename TOut>
FHC_DH void lane_transform_stage(const Tables& t, int n, int useDst, int inverse, const TIn* in, TOut* out, int add, int shift)
{
  const TMat m = tmat(t, n, useDst);
  const int sk = inverse ? 1 : m.rowStride, sq = inverse ? m.rowStride : 1;       // element (k, q): forward M(k,q), inverse M(q,k)
  for (int j = 0; j < n; j++)
  {
    int r[8];
#ifdef FHC_DEVICE_BUILD
#pragma unroll
#endif
    for (int q = 0; q < 8; q++) r[q] = q < n ? (int)(inverse ? in[q * n + j] : in[j * n + q]) : 0;
    for (int k = 0; k < n; k++)
    {
      int acc = 0;
#ifdef FHC_DEVICE_BUILD
#pragma unroll
#endif
      for (int q = 0; q < 8; q++)
        if (q < n) acc += m.tab[k * sk + q * sq] * r[q];
      const int v = (acc + add) >> shift;
      if (inverse) out[j * n + k] = (TOut)clip3(-32768, 32767, v);
      else out[k * n + j] = (TOut)v;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// COOP phases (striped over the lanes of the calling group: w = lane index, nw = group size)
// ----------------------------------------------------------------------------------------------
// cmd.a[] conventions are documented at each master call site ("par_*" helpers below).
// one copy per heavy operation: a phase call saves only the registers that operation needs
FHC_D void phase_refs(const Ctx& c, int phase, int w, int nw)
{
  FrameGlobal* g = c.g; (void)g;
  Shared* s = c.s; (void)s;
  Engine* e = c.e;
  const int* a = e->cmd.a;
  // a0 comp, a1 n, a2 nu, a3 us, a4 plane offset of the block origin, a5 wantFilt, a6 numAvail, a7 total units
  const int comp = a[0], n = a[1], nu = a[2], us = a[3], org0 = a[4], total = a[7];
  const int stride = plane_stride(s, comp);
  const int16_t* origin = g->rec[comp] + org0;
  if (phase == 0)
  {
    const int dc = 1 << (s->fp.bitDepth - 1);
    for (int i = w; i < 4 * n + 1; i += nw)
    {
      // unit of this sample: left part bottom-up, corner, top part
      const int lus = us == 4 ? 2 : 1;             // units are 4 (luma) or 2 (chroma) samples
      int uidx = i < 2 * n ? i >> lus : (i == 2 * n ? 2 * nu : 2 * nu + 1 + ((i - 2 * n - 1) >> lus));
      int v;
      if (a[6] == 0) v = dc;
      else
      {
        int src = i;
        if (!e->flags[uidx])
        {
          int j = uidx - 1;
          while (j >= 0 && !e->flags[j]) j--;
          if (j >= 0) src = j < 2 * nu ? (j * us + us - 1) : (j == 2 * nu ? 2 * n : 2 * n + 1 + (j - 2 * nu - 1) * us + us - 1);
          else
          {
            j = uidx + 1;
            while (j < total && !e->flags[j]) j++;
            src = j < 2 * nu ? j * us : (j == 2 * nu ? 2 * n : 2 * n + 1 + (j - 2 * nu - 1) * us);
          }
        }
        if (src < 2 * n) v = origin[(2 * n - 1 - src) * stride - 1];
        else if (src == 2 * n) v = origin[-stride - 1];
        else v = origin[-stride + (src - 2 * n - 1)];
      }
      e->refU[i] = (int16_t)v;
    }
  }
  else
  {
    // DC value of the unfiltered line + optional smoothing (TComPattern.cpp:191-296)
    const int16_t* line = e->refU;
    if (w == 0)
    {
      int sum = 0;
      for (int i = 0; i < n; i++) sum += line[2 * n + 1 + i] + line[2 * n - 1 - i];
      e->refDc = (sum + n) >> ilog2(n + n);
    }
    if (a[5])
    {
      const int bl = line[0], tl = line[2 * n], tr = line[4 * n];
      int strong = (comp == 0) && n >= 32;
      if (strong)
      {
        const int thr = 1 << (s->fp.bitDepth - 5);
        strong = (iabs((bl + tl) - 2 * line[n]) < thr) && (iabs((tl + tr) - 2 * line[3 * n]) < thr);
      }
      const int shift = ilog2(2 * n);
      for (int i = w; i < 4 * n + 1; i += nw)
      {
        int v;
        if (i == 0 || i == 4 * n) v = line[i];
        else if (strong)
        {
          if (i < 2 * n) v = (((2 * n - i) * bl) + (i * tl) + n) >> shift;
          else if (i == 2 * n) v = line[i];
          else v = (((2 * n - (i - 2 * n)) * tl) + ((i - 2 * n) * tr) + n) >> shift;
        }
        else v = (line[i - 1] + 2 * line[i] + line[i + 1] + 2) >> 2;
        e->refF[i] = (int16_t)v;
      }
    }
  }
}

// one copy per heavy operation: a phase call saves only the registers that operation needs
FHC_D void phase_rmd(const Ctx& c, int phase, int w, int nw)
{
  FrameGlobal* g = c.g; (void)g;
  Shared* s = c.s; (void)s;
  Engine* e = c.e;
  const int* a = e->cmd.a;
  // a0 log2n, a1 org plane offset, a2 modes per group; phases alternate predict(group) / satd(group)
  const int log2n = a[0], n = 1 << log2n, n2 = n * n, mpg = a[2];
  const int grp = phase >> 1, m0 = grp * mpg, m1 = imin(35, m0 + mpg);
  if (!(phase & 1))
  {
    const int items = (m1 - m0) * n2;
    for (int i = w; i < items; i += nw)
    {
      const int mi = i >> (2 * log2n), sidx = i & (n2 - 1), mode = m0 + mi;
      const int x = sidx & (n - 1), y = sidx >> log2n;
      const int16_t* line = use_filtered_refs(*s->t, 0, mode, log2n) ? e->refF : e->refU;
      g->rmdPred[i] = (int16_t)pred_sample(*s->t, line, n, log2n, mode, x, y, 1, s->fp.bitDepth, e->refDc);
    }
  }
  else
  {
    const int stride = plane_stride(s, 0);
    if (n >= 8)
    {
      const int tps = n >> 3, tiles = tps * tps, items = (m1 - m0) * tiles;
      for (int i = w; i < items; i += nw)
      {
        const int ltps = log2n - 3, mi = i >> (2 * ltps), tl = i & (tiles - 1), ty = tl >> ltps, tx = tl & (tps - 1);
        int m[64];
        for (int yy = 0; yy < 8; yy++)
          for (int xx = 0; xx < 8; xx++)
            m[yy * 8 + xx] = org_at(g, s->fp.sampleBytes, 0, a[1] + (ty * 8 + yy) * stride + tx * 8 + xx) - g->rmdPred[mi * n2 + (ty * 8 + yy) * n + tx * 8 + xx];
        atomic_add_u32(&s->satd[m0 + mi], had8(m), nw);
      }
    }
    else
    {
      for (int mi = w; mi < m1 - m0; mi += nw)
      {
        int d[16];
        for (int yy = 0; yy < 4; yy++)
          for (int xx = 0; xx < 4; xx++) d[yy * 4 + xx] = org_at(g, s->fp.sampleBytes, 0, a[1] + yy * stride + xx) - g->rmdPred[mi * 16 + yy * 4 + xx];
        atomic_add_u32(&s->satd[m0 + mi], had4(d), nw);
      }
    }
  }
}

// one copy per heavy operation: a phase call saves only the registers that operation needs
FHC_D void phase_pred_resi(const Ctx& c, int phase, int w, int nw)
{
  FrameGlobal* g = c.g; (void)g;
  Shared* s = c.s; (void)s;
  Engine* e = c.e;
  const int* a = e->cmd.a;
  // a0 comp, a1 log2n, a2 mode, a3 useFilt, a4 offset into predYuv, a5 stride in predYuv, a6 org plane offset,
  // a7 save/load (0 none, 1 save to tsPred, 2 load from tsPred)
  const int comp = a[0], log2n = a[1], n = 1 << log2n, n2 = n * n;
  int16_t* pred = e->predYuv + a[4];
  const int ps = a[5], stride = plane_stride(s, comp);
  const int16_t* line = a[3] ? e->refF : e->refU;
  for (int i = w; i < n2; i += nw)
  {
    const int x = i & (n - 1), y = i >> log2n;
    int v;
    if (a[7] == 2) v = e->tsPred[comp][i];
    else
    {
      v = pred_sample(*s->t, line, n, log2n, a[2], x, y, comp == 0, s->fp.bitDepth, e->refDc);
      if (a[7] == 1) e->tsPred[comp][i] = (int16_t)v;
    }
    pred[y * ps + x] = (int16_t)v;
    e->resi[i] = (int16_t)(org_at(g, s->fp.sampleBytes, comp, a[6] + y * stride + x) - v);
  }
}

// one copy per heavy operation: a phase call saves only the registers that operation needs
FHC_D void phase_fwd_quant(const Ctx& c, int phase, int w, int nw)
{
  FrameGlobal* g = c.g; (void)g;
  Shared* s = c.s; (void)s;
  Engine* e = c.e;
  const int* a = e->cmd.a;
  // a0 log2n, a1 useDst, a2 transformSkip, a3 scanType, a4 quantCoef, a5 qBits, a6 chType(0/1), a7 qp rem
  // out: tcoef[] raster, lvlD[scanPos], rd.cost0[scanPos], tmp[blkPos] = uiMaxAbsLevel
  const int log2n = a[0], n = 1 << log2n, n2 = n * n, bd = s->fp.bitDepth;
  const TMat fm = tmat(*s->t, n, a[1]);
  if (a[2])
  {
    if (phase == 1)
    {
      const int shift = 15 - bd - log2n;
      for (int i = w; i < n2; i += nw) e->tcoef[i] = (int32_t)e->resi[i] << shift;
    }
  }
  else if (phase == 0)
  {
    // xTrMxN first stage (TComTrQuant.cpp:860): rows -> tmp[k][j]
    const int shift1 = log2n + bd + 6 - 15, add1 = shift1 > 0 ? 1 << (shift1 - 1) : 0;
    if (nw == 1 && n <= 8) { lane_transform_stage(*s->t, n, a[1], 0, e->resi, e->tmp, add1, shift1); return; }
    for (int i = w; i < n2; i += nw)
    {
      const int k = i >> log2n, j = i & (n - 1);
      int acc = 0;
      const int16_t* row = fm.tab + k * fm.rowStride;
      for (int q = 0; q < n; q++) acc += row[q] * e->resi[j * n + q];
      e->tmp[i] = (acc + add1) >> shift1;
    }
  }
  else if (phase == 1)
  {
    const int shift2 = log2n + 6, add2 = 1 << (shift2 - 1);
    if (nw == 1 && n <= 8) lane_transform_stage(*s->t, n, a[1], 0, e->tmp, e->tcoef, add2, shift2);
    else
    for (int i = w; i < n2; i += nw)
    {
      const int k = i >> log2n, j = i & (n - 1);
      int acc = 0;
      const int16_t* row = fm.tab + k * fm.rowStride;
      for (int q = 0; q < n; q++) acc += row[q] * e->tmp[j * n + q];
      e->tcoef[i] = (acc + add2) >> shift2;
    }
  }
  if (phase == 2)
  {
    // RDOQ pre-quantisation (TComTrQuant.cpp:2244-2262), independent per coefficient
    const uint16_t* scan = scan_tab(*s->t, a[3], log2n);
    const int qBits = a[5];
    const double errScale = s->t->errScale[a[6]][log2n - 2][a[7]];
    const int64_t cap = (int64_t)2147483647 - ((int64_t)1 << (qBits - 1));
    for (int p = w; p < n2; p += nw)
    {
      const int blk = scan[p];
      const int64_t tl = (int64_t)iabs(e->tcoef[blk]) * a[4];
      const int ld = (int)(tl < cap ? tl : cap);
      const uint32_t m = (uint32_t)((ld + (1 << (qBits - 1))) >> qBits);
      const double err = (double)ld;
      e->lvlD[p] = ld;
      e->cost0[p] = err * err * errScale;
      e->tmp[blk] = (int32_t)(m < 32767u ? m : 32767u);
      if (m) atomic_or_u32(&e->cgNZ[p >> 9], 1u << ((p >> 4) & 31), nw);
    }
  }
}

// one copy per heavy operation: a phase call saves only the registers that operation needs
FHC_D void phase_inv_recon(const Ctx& c, int phase, int w, int nw)
{
  FrameGlobal* g = c.g; (void)g;
  Shared* s = c.s; (void)s;
  Engine* e = c.e;
  const int* a = e->cmd.a;
  // a0 comp, a1 log2n, a2 useDst, a3 transformSkip, a4 absSum>0, a5 qp per, a6 qp rem,
  // a7 predYuv offset, a8 predYuv stride, a9 qt layer, a10 qtRec offset, a11 rec plane offset, a12 org plane offset
  const int comp = a[0], log2n = a[1], n = 1 << log2n, n2 = n * n, bd = s->fp.bitDepth;
  const TMat im = tmat(*s->t, n, a[2]);
  if (a[4])
  {
    if (phase == 0)
    {
      // xDeQuant (TComTrQuant.cpp:1308), flat scaling; levels are in tmp[] -> tcoef[]
      const int transformShift = 15 - bd - log2n, rightShift = 6 - (transformShift + a[5]);
      const int scale = (a[6] == 0 ? 40 : a[6] == 1 ? 45 : a[6] == 2 ? 51 : a[6] == 3 ? 57 : a[6] == 4 ? 64 : 72);
      const int targetBits = imin(16, 32 + rightShift - 7);
      const int inMin = -(1 << (targetBits - 1)), inMax = (1 << (targetBits - 1)) - 1;
      for (int i = w; i < n2; i += nw)
      {
        const int q = clip3(inMin, inMax, e->tmp[i]);
        int v;
        if (rightShift > 0) v = (q * scale + (1 << (rightShift - 1))) >> rightShift;
        else v = (int)((uint32_t)(q * scale) << (-rightShift));
        e->tcoef[i] = clip3(-32768, 32767, v);
      }
    }
    else if (phase == 1)
    {
      if (a[3])
      {
        const int shift = 15 - bd - log2n, off = shift == 0 ? 0 : (1 << (shift - 1));
        for (int i = w; i < n2; i += nw) e->resi[i] = (int16_t)((e->tcoef[i] + off) >> shift);
      }
      else
      {
        // xITrMxN first stage (TComTrQuant.cpp:927): columns -> lvlD[j][i] (reused as scratch)
        if (nw == 1 && n <= 8) lane_transform_stage(*s->t, n, a[2], 1, e->tcoef, e->lvlD, 64, 7);
        else
        for (int i = w; i < n2; i += nw)
        {
          const int j = i >> log2n, q = i & (n - 1);
          int acc = 0;
          for (int k = 0; k < n; k++) acc += im.tab[k * im.rowStride + q] * e->tcoef[k * n + j];
          e->lvlD[i] = clip3(-32768, 32767, (acc + 64) >> 7);
        }
      }
    }
    else if (phase == 2)
    {
      if (!a[3])
      {
        const int shift2 = 12 - (bd - 8), add2 = 1 << (shift2 - 1);
        if (nw == 1 && n <= 8) lane_transform_stage(*s->t, n, a[2], 1, e->lvlD, e->resi, add2, shift2);
        else
        for (int i = w; i < n2; i += nw)
        {
          const int j = i >> log2n, q = i & (n - 1);
          int acc = 0;
          for (int k = 0; k < n; k++) acc += im.tab[k * im.rowStride + q] * e->lvlD[k * n + j];
          e->resi[i] = (int16_t)clip3(-32768, 32767, (acc + add2) >> shift2);
        }
      }
    }
  }
  if (phase == 3)
  {
    // reconstruction into CU buffer, QT layer and picture (TEncSearch.cpp:1380-1393) + SSE (TComRdCost.cpp:1176)
    int16_t* reco = e->predYuv + a[7];
    int16_t* recQt = qt_rec(e, a[9]) + a[10];
    int16_t* recPic = g->rec[comp] + a[11];
    const int ps = a[8], qs = e->qs >> (comp ? 1 : 0), stride = plane_stride(s, comp), shift = (bd - 8) << 1;
    uint32_t acc = 0;
    for (int i = w; i < n2; i += nw)
    {
      const int x = i & (n - 1), y = i >> log2n;
      const int r = a[4] ? e->resi[i] : 0;
      const int v = clip3(0, (1 << bd) - 1, reco[y * ps + x] + r);
      reco[y * ps + x] = (int16_t)v;
      recQt[y * qs + x] = (int16_t)v;
      if (e->writePic) recPic[y * stride + x] = (int16_t)v;
      const int d = v - org_at(g, s->fp.sampleBytes, comp, a[12] + y * stride + x);
      acc += (uint32_t)((d * d) >> shift);
    }
    if (acc) atomic_add_u32(&e->sse, acc, nw);
  }
}

FHC_D void run_phase(const Ctx& c, int op, int phase, int w, int nw)
{
  FrameGlobal* g = c.g; (void)g;
  Shared* s = c.s; (void)s;
  Engine* e = c.e;
  const int* a = e->cmd.a;
  switch (op)
  {
  case OP_REFS: phase_refs(c, phase, w, nw); break;
  case OP_RMD: phase_rmd(c, phase, w, nw); break;
  case OP_PRED_RESI: phase_pred_resi(c, phase, w, nw); break;
  case OP_FWD_QUANT: phase_fwd_quant(c, phase, w, nw); break;
  case OP_INV_RECON: phase_inv_recon(c, phase, w, nw); break;
  case OP_LEVELS_OUT:
  {
    // a0 n2, a1 qtCoef layer, a2 offset: tmp[] (int32 levels) -> qtCoef (int16)
    int16_t* d = qt_coef(e, a[1]) + a[2];
    for (int i = w; i < a[0]; i += nw) d[i] = (int16_t)e->tmp[i];
    break;
  }
  case OP_NZMASK:
  {
    // a0 log2n, a1 scan type, src = int16 levels (raster): non-zero map of every coefficient group, in scan order
    const uint16_t* scan = scan_tab(*s->t, a[1], a[0]);
    const int16_t* coef = (const int16_t*)e->cmd.src;
    const int ncg = 1 << (2 * a[0] - 4);
    for (int cg = w; cg < ncg; cg += nw)
    {
      uint32_t mask = 0;
      for (int i = 0; i < 16; i++) mask |= (uint32_t)(coef[scan[cg * 16 + i]] != 0) << i;
      e->cgNzMask[cg] = (uint16_t)mask;
    }
    break;
  }
  case OP_COPY2D:
  {
    // int16 block copy: a0 w, a1 h, a2 dst stride, a3 src stride
    const int bw = a[0], bh = a[1];
    int16_t* d = (int16_t*)e->cmd.dst;
    const int16_t* sr = (const int16_t*)e->cmd.src;
    if ((bw & (bw - 1)) == 0)
    {
      const int lg = ilog2(bw);                      // block widths are powers of two except at the picture border
      for (int i = w; i < bw * bh; i += nw)
      {
        const int y = i >> lg, x = i & (bw - 1);
        d[y * a[2] + x] = sr[y * a[3] + x];
      }
    }
    else
      for (int i = w; i < bw * bh; i += nw)
      {
        const int y = i / bw, x = i - y * bw;
        d[y * a[2] + x] = sr[y * a[3] + x];
      }
    break;
  }
  case OP_COPY16:
  {
    int16_t* d = (int16_t*)e->cmd.dst;
    const int16_t* sr = (const int16_t*)e->cmd.src;
    for (int i = w; i < a[0]; i += nw) d[i] = sr[i];
    break;
  }
  case OP_ZERO16:
  {
    int16_t* d = (int16_t*)e->cmd.dst;
    for (int i = w; i < a[0]; i += nw) d[i] = 0;
    break;
  }
  case OP_CTU_RECO_OUT:
  {
    // reconstruction of the finished CTU into its result record (0 outside the picture)
    fhc_ctu_result* o = &g->out[s->ctuAddr];
    for (int i = w; i < 6144; i += nw)
    {
      const int comp = i < 4096 ? 0 : (i < 5120 ? 1 : 2), sh = comp ? 1 : 0, nn = 64 >> sh;
      const int k = i - comp_off(comp), xx = k & (nn - 1), yy = k >> (6 - sh);
      const int px = (s->ctuPelX >> sh) + xx, py = (s->ctuPelY >> sh) + yy;
      int16_t* d = comp == 0 ? o->recoY : (comp == 1 ? o->recoCb : o->recoCr);
      d[k] = (px < (s->fp.width >> sh) && py < (s->fp.height >> sh)) ? g->rec[comp][py * plane_stride(s, comp) + px] : (int16_t)0;
    }
    break;
  }
  case OP_CU_TO_OUT:
  {
    // copyToPic (TComDataCU.cpp:906): a0 depth, a1 buffer index, a2 zidx, a3 nparts, a4 width
    const CUMeta* cu = &s->cu[a[0]][a[1]];
    fhc_ctu_result* o = &g->out[s->ctuAddr];
    const int z = a[2], np = a[3], wdt = a[4];
    for (int i = w; i < np; i += nw)
    {
      o->partSize[z + i] = cu->part[i]; o->predMode[z + i] = cu->pred[i]; o->lumaDir[z + i] = cu->ldir[i];
      o->chromaDir[z + i] = cu->cdir[i]; o->trIdx[z + i] = cu->tridx[i]; o->depth[z + i] = cu->dep[i];
      for (int cc = 0; cc < 3; cc++) { o->cbf[cc][z + i] = cu->cbf[cc][i]; o->transformSkip[cc][z + i] = cu->ts[cc][i]; }
    }
    const int16_t* sc = g->cuCoef[a[0]][a[1]];
    const int ny = wdt * wdt, nc = ny >> 2;
    for (int i = w; i < ny; i += nw) o->coefY[z * 16 + i] = sc[i];
    for (int i = w; i < nc; i += nw) { o->coefCb[z * 4 + i] = sc[4096 + i]; o->coefCr[z * 4 + i] = sc[5120 + i]; }
    break;
  }
  default: break;
  }
}

// ----------------------------------------------------------------------------------------------
// master-side helpers that wrap commands
// ----------------------------------------------------------------------------------------------
FHC_DH void par_copy2d(const Ctx& c, int16_t* dst, int ds, const int16_t* src, int ss, int w, int h)
{
  Cmd& m = c.e->cmd;
  if (w * h <= 16) { for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) dst[y * ds + x] = src[y * ss + x]; return; }
  m.a[0] = w; m.a[1] = h; m.a[2] = ds; m.a[3] = ss; m.dst = dst; m.src = src;
  par(c, OP_COPY2D);
}
FHC_DH void par_copy16(const Ctx& c, int16_t* dst, const int16_t* src, int n)
{
  Cmd& m = c.e->cmd;
  if (n <= 16) { for (int i = 0; i < n; i++) dst[i] = src[i]; return; }
  m.a[0] = n; m.dst = dst; m.src = src;
  par(c, OP_COPY16);
}
FHC_DH void par_zero16(const Ctx& c, int16_t* dst, int n)
{
  Cmd& m = c.e->cmd;
  m.a[0] = n; m.dst = dst;
  par(c, OP_ZERO16);
}
FHC_DH void set_range(uint8_t* a, int start, int count, int v) { for (int i = 0; i < count; i++) a[start + i] = (uint8_t)v; }
FHC_DH void copy_bytes(uint8_t* d, const uint8_t* s, int n) { for (int i = 0; i < n; i++) d[i] = s[i]; }

// ----------------------------------------------------------------------------------------------
// CABAC bit counter (TEncBinCoderCABACCounter.cpp:57-140)
// ----------------------------------------------------------------------------------------------
FHC_DH void enc_bin(Engine* e, int ctx, int bin)
{
  const uint8_t st = e->go->c[ctx];
  e->bins++;
  FHC_COUNT(g_emuCounters.bins++);
  e->go->frac += (uint64_t)e->t->entropyBits[st ^ bin];
  e->go->c[ctx] = e->t->nextState[st][bin];
}
FHC_DH void enc_ep(Engine* e, int n) { e->bins += (uint32_t)n; FHC_COUNT(g_emuCounters.bins += (unsigned)n); e->go->frac += (uint64_t)32768 * (uint64_t)n; }
FHC_DH void enc_trm(Engine* e, int bin) { e->bins++; FHC_COUNT(g_emuCounters.bins++); e->go->frac += (uint64_t)e->t->entropyBits[126 ^ bin]; }
FHC_DH void reset_bits(Engine* e) { e->go->frac &= 32767; e->bins = 0; }
FHC_DH uint32_t num_bits(const Engine* e) { return (uint32_t)(e->go->frac >> 15); }
FHC_DH int ctx_bits(const Engine* e, int ctx, int bin) { return e->t->entropyBits[e->go->c[ctx] ^ bin]; }
FHC_DH void sb_copy(const Ctx& c, Sbac* d, const Sbac* sr)
{
  const uint64_t* a = (const uint64_t*)sr;
  uint64_t* b = (uint64_t*)d;
  (void)c;
#ifdef FHC_DEVICE_BUILD
#pragma unroll
#endif
  for (int i = 0; i < (int)(sizeof(Sbac) / 8); i++) b[i] = a[i];
}
FHC_DH void sbac_reset(const Tables& t, Sbac* sb, int qp)                        // ContextModel::init (ContextModel.cpp:56)
{
  qp = clip3(0, 51, qp);
  FHC_NOUNROLL for (int i = 0; i < NUM_CTX; i++)
  {
    const int iv = t.ctxInit[i], slope = (iv >> 4) * 5 - 45, offset = ((iv & 15) << 3) - 16;
    const int st = imin(imax(1, ((slope * qp) >> 4) + offset), 126);
    const int mps = st >= 64;
    sb->c[i] = (uint8_t)(((mps ? (st - 64) : (63 - st)) << 1) + mps);
  }
  sb->frac = 0;
}
#ifdef FHC_DEVICE_BUILD
FHC_DH double rd_cost(const Ctx& c, uint32_t bits, uint32_t dist) { return __dadd_rn((double)dist, __dmul_rn((double)bits, c.s->fp.sp.lambda)); }
FHC_DH double dmul(double a, double b) { return __dmul_rn(a, b); }
FHC_DH double dadd(double a, double b) { return __dadd_rn(a, b); }
FHC_DH double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
FHC_DH double rd_cost(const Ctx& c, uint32_t bits, uint32_t dist) { return (double)dist + ((double)bits * c.s->fp.sp.lambda); }
FHC_DH double dmul(double a, double b) { return a * b; }
FHC_DH double dadd(double a, double b) { return a + b; }
FHC_DH double dsub(double a, double b) { return a - b; }
#endif

// context-set slots: the chain/CU-level slots belong to the leader (chain_slot); a lane engine aliases every depth of the
// TU-level slots (CI_QT_TRAFO_*, CI_TEMP_BEST) onto its private set - a task only ever uses one depth of them.
FHC_DH Sbac* chain_slot(const Ctx& c, int depth, int ci)
{
  return depth >= FHC_SLOT_SMEM_DEPTH ? &c.s->slot[depth - FHC_SLOT_SMEM_DEPTH][ci] : &c.g->slotLow[depth][ci];
}
FHC_DH Sbac* slot_of(const Ctx& c, int depth, int ci)
{
  if (c.e->kind == ENG_MASTER || ci == CI_CURR_BEST) return chain_slot(c, depth, ci);
  return &c.e->slots[ci == CI_QT_TRAFO_ROOT ? 0 : (ci == CI_TEMP_BEST ? 1 : 2)];
}

// winner of a dispatch: lowest cost, ties to the lowest task index (= the sequential strict-< scan of the reference)
FHC_DH int best_engine(const Ctx& c)
{
  const FrameGlobal* g = c.g;
  const int base = c.s->task.resBase, ne = base + imin(FHC_NENG, c.s->task.numTasks);
  int be = base;
  FHC_NOUNROLL for (int i = base + 1; i < ne; i++)
  {
    const EngineGlobal* a = &g->eng[i];
    const EngineGlobal* b = &g->eng[be];
    if (a->resCost < b->resCost || (a->resCost == b->resCost && a->resTask < b->resTask)) be = i;
  }
  return be;
}

// ----------------------------------------------------------------------------------------------
// neighbours (TComDataCU.cpp:985-1200, :1362-1461)
// ----------------------------------------------------------------------------------------------
enum { F_DEPTH, F_LDIR };
FHC_DH int cu_field(const CUMeta* cu, int idx, int f) { return f == F_DEPTH ? cu->dep[idx] : (cu->pred[idx] == MODE_INTRA ? cu->ldir[idx] : DC_IDX); }
FHC_DH int ctu_field(const fhc_ctu_result* o, int idx, int f) { return f == F_DEPTH ? o->depth[idx] : (o->predMode[idx] == MODE_INTRA ? o->lumaDir[idx] : DC_IDX); }

FHC_DH int nb_left(const Ctx& c, const CUMeta* cu, int zabs, int f)
{
  const Shared* s = c.s;
  Engine* e = c.e; (void)e;
  const int r = s->t->z2r[zabs];
  if (r & 15)
  {
    const int nz = s->t->r2z[r - 1];
    if ((r & 15) == (s->t->z2r[cu->zidx] & 15)) return ctu_field(&c.g->out[s->ctuAddr], nz, f);
    return cu_field(cu, nz - cu->zidx, f);
  }
  if (s->ctuX == 0) return -1;
  return ctu_field(&c.g->out[s->ctuAddr - 1], s->t->r2z[r + 15], f);
}
FHC_DH int nb_above(const Ctx& c, const CUMeta* cu, int zabs, int f, int planarAtCtuBoundary)
{
  const Shared* s = c.s;
  Engine* e = c.e; (void)e;
  const int r = s->t->z2r[zabs];
  if (r >> 4)
  {
    const int nz = s->t->r2z[r - 16];
    if ((r >> 4) == (s->t->z2r[cu->zidx] >> 4)) return ctu_field(&c.g->out[s->ctuAddr], nz, f);
    return cu_field(cu, nz - cu->zidx, f);
  }
  if (planarAtCtuBoundary || s->ctuY == 0) return -1;
  return ctu_field(&c.g->out[s->ctuAddr - c.s->fp.wc], s->t->r2z[r + 240], f);
}
FHC_DH int get_mpm(const Ctx& c, const CUMeta* cu, int relIdx, int mpm[3])
{
  const int zabs = cu->zidx + relIdx;
  int l = nb_left(c, cu, zabs, F_LDIR), a = nb_above(c, cu, zabs, F_LDIR, 1);
  if (l < 0) l = DC_IDX;
  if (a < 0) a = DC_IDX;
  if (l == a)
  {
    if (l > 1) { mpm[0] = l; mpm[1] = ((l + 29) % 32) + 2; mpm[2] = ((l - 1) % 32) + 2; }
    else { mpm[0] = PLANAR_IDX; mpm[1] = DC_IDX; mpm[2] = VER_IDX; }
    return 1;
  }
  mpm[0] = l; mpm[1] = a;
  if (l && a) mpm[2] = PLANAR_IDX;
  else mpm[2] = (l + a) < 2 ? VER_IDX : DC_IDX;
  return 2;
}
FHC_DH int ctx_split(const Ctx& c, const CUMeta* cu, int relIdx, int depth)
{
  const int zabs = cu->zidx + relIdx;
  const int l = nb_left(c, cu, zabs, F_DEPTH), a = nb_above(c, cu, zabs, F_DEPTH, 0);
  return (l >= 0 && l > depth) + (a >= 0 && a > depth);
}
FHC_DH int coef_scan_idx(const CUMeta* cu, int relIdx, int n, int comp)       // TComDataCU.cpp:3109
{
  int dir;
  if (n > (comp ? 4 : 8)) return SCAN_DIAG;
  if (comp == 0) dir = cu->ldir[relIdx];
  else
  {
    dir = cu->cdir[relIdx];
    if (dir == DM_CHROMA_IDX) dir = cu->ldir[relIdx & ~3];
  }
  if (iabs(dir - VER_IDX) <= 4) return SCAN_HOR;
  if (iabs(dir - HOR_IDX) <= 4) return SCAN_VER;
  return SCAN_DIAG;
}

// ----------------------------------------------------------------------------------------------
// syntax element bit counting (TEncSbac.cpp)
// ----------------------------------------------------------------------------------------------
FHC_DH void code_split_flag(const Ctx& c, const CUMeta* cu, int relIdx, int depth)
{
  if (depth == 3) return;
  enc_bin(c.e, CTX_SPLIT + ctx_split(c, cu, relIdx, depth), cu->dep[relIdx] > depth ? 1 : 0);
}
FHC_DH void code_part_size(const Ctx& c, const CUMeta* cu, int relIdx, int depth)
{
  if (depth == 3) enc_bin(c.e, CTX_PART, cu->part[relIdx] == SIZE_2Nx2N ? 1 : 0);
}
FHC_D void code_luma_dir(const Ctx& c, const CUMeta* cu, int relIdx, int isMultiple)          // TEncSbac.cpp:643
{
  const int partNum = isMultiple ? (cu->part[relIdx] == SIZE_NxN ? 4 : 1) : 1;
  const int partOffset = (256 >> (cu->dep[relIdx] << 1)) >> 2;
  int predIdx[4];
  FHC_NOUNROLL for (int j = 0; j < partNum; j++)
  {
    int preds[3];
    const int dir = cu->ldir[relIdx + partOffset * j];
    get_mpm(c, cu, relIdx + partOffset * j, preds);
    predIdx[j] = -1;
    FHC_NOUNROLL for (int i = 0; i < 3; i++) if (dir == preds[i]) predIdx[j] = i;
    enc_bin(c.e, CTX_IPRED, predIdx[j] != -1 ? 1 : 0);
  }
  FHC_NOUNROLL for (int j = 0; j < partNum; j++) enc_ep(c.e, predIdx[j] != -1 ? (predIdx[j] ? 2 : 1) : 5);
}
FHC_DH void code_chroma_dir(const Ctx& c, const CUMeta* cu, int relIdx)                        // TEncSbac.cpp:698
{
  if (cu->cdir[relIdx] == DM_CHROMA_IDX) enc_bin(c.e, CTX_CPRED, 0);
  else { enc_bin(c.e, CTX_CPRED, 1); enc_ep(c.e, 2); }
}
FHC_DH int sig_ctx_inc(const Tables& t, int pat, int firstCtx, int rasterPos, int log2n, int isChroma)   // TComTrQuant.cpp:2707
{
  if (rasterPos == 0) return 0;
  if (log2n == 2) return firstCtx + t.ctxIndMap4x4[rasterPos];
  const int posY = rasterPos >> log2n, posX = rasterPos - (posY << log2n);
  const int cnt = t.sigCnt[pat][((posY & 3) << 2) | (posX & 3)];
  return firstCtx + ((((posX | posY) >> 2) != 0 && !isChroma) ? 3 : 0) + cnt;
}
FHC_DH int first_sig_ctx(int log2n, int scanType, int isChroma)
{
  if (log2n == 2) return 0;
  if (log2n == 3) return 9 + ((scanType != SCAN_DIAG && !isChroma) ? 6 : 0);
  return isChroma ? 12 : 21;
}
FHC_DH int pattern_sig_ctx(const uint8_t* cgFlag, int cgX, int cgY, int wg)
{
  int r = 0, l = 0;
  if (wg <= 1) return 0;
  if (cgX < wg - 1) r = cgFlag[cgY * wg + cgX + 1] != 0;
  if (cgY < wg - 1) l = cgFlag[(cgY + 1) * wg + cgX] != 0;
  return r + (l << 1);
}
FHC_DH int sig_cg_ctx(const uint8_t* cgFlag, int cgX, int cgY, int wg)
{
  int r = 0, l = 0;
  if (cgX < wg - 1) r = cgFlag[cgY * wg + cgX + 1] != 0;
  if (cgY < wg - 1) l = cgFlag[(cgY + 1) * wg + cgX] != 0;
  return (r + l) != 0;
}
FHC_DH void last_ctx_params(int isChroma, int log2n, int* off, int* shift)
{
  const int cc = log2n - 2;
  *off = isChroma ? 0 : (cc * 3 + ((cc + 1) >> 2));
  *shift = isChroma ? cc : ((cc + 3) >> 2);
}
FHC_DH int group_idx(int p) { return p < 4 ? p : (p < 6 ? 4 : (p < 8 ? 5 : (p < 12 ? 6 : (p < 16 ? 7 : (p < 24 ? 8 : 9))))); }

FHC_DH void code_remain(Engine* e, uint32_t symbol, uint32_t rParam)                            // TEncSbac.cpp:337
{
  int codeNumber = (int)symbol;
  if (codeNumber < (3 << rParam)) enc_ep(e, (int)(((uint32_t)codeNumber >> rParam) + 1 + rParam));
  else
  {
    uint32_t length = rParam;
    codeNumber -= (3 << rParam);
    FHC_NOUNROLL while (codeNumber >= (1 << length)) codeNumber -= (1 << (length++));
    enc_ep(e, (int)(3 + length + 1 - rParam + length));
  }
}

// TEncSbac::codeCoeffNxN (TEncSbac.cpp:1181-1538); coef: int16 raster levels of this TU
FHC_NOINLINE void code_coeff_nxn_impl(const Ctx& c, const CUMeta* cu, int relIdxComp, const int16_t* coef, int log2n, int comp)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  const int isChroma = comp != 0, n = 1 << log2n;
  const int scanType = coef_scan_idx(cu, relIdxComp, n, comp);
  const uint16_t* scan = scan_tab(*s->t, scanType, log2n);
  const uint8_t* scanCG = scan_cg(*s->t, scanType, log2n);
  const int wg = n >> 2;
  const int firstSig = first_sig_ctx(log2n, scanType, isChroma);
  const int sigBase = CTX_SIG + (isChroma ? 28 : 0);
  if (e->kind == ENG_MASTER)
  {
    // group scratch by TU size (this function is also called outside code_tu_block, for TUs of any size)
    if (log2n <= 3) bind_group_scratch(e, (GroupScratch<4>*)e->gsSmall); else bind_group_scratch(e, (GroupScratch<64>*)e->gsLarge);
  }
  uint8_t* cgFlag = e->cgFlagCode;
  int scanPosLast = -1, posLast = 0;

  if (log2n == 2) enc_bin(e, CTX_TS + isChroma, cu->ts[comp][relIdxComp]);
  // non-zero map per coefficient group (COOP for TUs of 8x8 and up), then last position and group flags from the maps
  const uint16_t* nzMask = e->cgNzMask;
  {
    Cmd& m = e->cmd;
    m.a[0] = log2n; m.a[1] = scanType; m.src = coef;
    if (log2n >= 3) par(c, OP_NZMASK); else run_phase(c, OP_NZMASK, 0, 0, 1);
  }
  FHC_NOUNROLL for (int i = 0; i < wg * wg; i++) cgFlag[i] = 0;
  FHC_NOUNROLL for (int g4 = wg * wg - 1; g4 >= 0; g4--)
  {
    const uint32_t mask = nzMask[g4];
    if (!mask) continue;
    if (scanPosLast < 0) scanPosLast = g4 * 16 + msb_index(mask);
    cgFlag[scanCG[g4]] = 1;
  }
  posLast = scan[scanPosLast];
  {
    int posY = posLast >> log2n, posX = posLast - (posY << log2n), off, shift, cc;
    if (scanType == SCAN_VER) { const int t = posX; posX = posY; posY = t; }
    const int gx = group_idx(posX), gy = group_idx(posY), gmax = group_idx(n - 1);
    last_ctx_params(isChroma, log2n, &off, &shift);
    const int bx = CTX_LASTX + (isChroma ? 15 : 0) + off, by = CTX_LASTY + (isChroma ? 15 : 0) + off;
    FHC_NOUNROLL for (cc = 0; cc < gx; cc++) enc_bin(e, bx + (cc >> shift), 1);
    if (gx < gmax) enc_bin(e, bx + (cc >> shift), 0);
    FHC_NOUNROLL for (cc = 0; cc < gy; cc++) enc_bin(e, by + (cc >> shift), 1);
    if (gy < gmax) enc_bin(e, by + (cc >> shift), 0);
    if (gx > 3) enc_ep(e, (gx - 2) >> 1);
    if (gy > 3) enc_ep(e, (gy - 2) >> 1);
  }
  {
    const int lastScanSet = scanPosLast >> 4;
    uint32_t c1 = 1;
    int scanPosSig = scanPosLast;
    FHC_NOUNROLL for (int subSet = lastScanSet; subSet >= 0; subSet--)
    {
      int numNonZero = 0, lastNZ = -1, firstNZ = 16, escape = 0;
      const int subPos = subSet << 4;
      int absCoeff[16];
      uint32_t goRice = 0;
      const int cgBlk = scanCG[subSet], cgY = cgBlk >> (log2n - 2), cgX = cgBlk & (wg - 1);
      if (scanPosSig == scanPosLast)
      {
        absCoeff[0] = iabs(coef[posLast]);
        numNonZero = 1; lastNZ = scanPosSig; firstNZ = scanPosSig; scanPosSig--;
      }
      if (subSet == lastScanSet || subSet == 0) cgFlag[cgBlk] = 1;
      else enc_bin(e, CTX_SIGCG + (isChroma ? 2 : 0) + sig_cg_ctx(cgFlag, cgX, cgY, wg), cgFlag[cgBlk] != 0);
      if (cgFlag[cgBlk])
      {
        const int pat = pattern_sig_ctx(cgFlag, cgX, cgY, wg);
        const uint32_t cgMask = nzMask[subSet];
        FHC_NOUNROLL for (; scanPosSig >= subPos; scanPosSig--)
        {
          const int blkPos = scan[scanPosSig], sig = (int)((cgMask >> (scanPosSig - subPos)) & 1u);
          if (scanPosSig > subPos || subSet == 0 || numNonZero)
            enc_bin(e, sigBase + sig_ctx_inc(*s->t, pat, firstSig, blkPos, log2n, isChroma), sig);
          if (sig)
          {
            absCoeff[numNonZero++] = iabs(coef[blkPos]);
            if (lastNZ == -1) lastNZ = scanPosSig;
            firstNZ = scanPosSig;
          }
        }
      }
      else scanPosSig = subPos - 1;
      if (numNonZero > 0)
      {
        const int signHidden = (lastNZ - firstNZ) >= 4;
        const int ctxSet = (isChroma ? 4 : 0) + ((!isChroma && subSet > 0) ? 2 : 0) + (c1 == 0 ? 1 : 0);
        const int numC1 = imin(numNonZero, 8);
        int firstC2 = -1;
        c1 = 1;
        FHC_NOUNROLL for (int idx = 0; idx < numC1; idx++)
        {
          const int sym = absCoeff[idx] > 1;
          enc_bin(e, CTX_ONE + 4 * ctxSet + (int)c1, sym);
          if (sym) { c1 = 0; if (firstC2 == -1) firstC2 = idx; else escape = 1; }
          else if (c1 < 3 && c1 > 0) c1++;
        }
        if (c1 == 0 && firstC2 != -1)
        {
          const int sym = absCoeff[firstC2] > 2;
          enc_bin(e, CTX_ABS + ctxSet, sym);
          if (sym) escape = 1;
        }
        escape = escape || (numNonZero > 8);
        enc_ep(e, signHidden ? numNonZero - 1 : numNonZero);
        if (escape)
        {
          int firstCoeff2 = 1;
          FHC_NOUNROLL for (int idx = 0; idx < numNonZero; idx++)
          {
            const int baseLevel = (idx < 8) ? (2 + firstCoeff2) : 1;
            if (absCoeff[idx] >= baseLevel)
            {
              code_remain(e, (uint32_t)(absCoeff[idx] - baseLevel), goRice);
              if (absCoeff[idx] > (3 << goRice)) goRice = goRice + 1 < 4 ? goRice + 1 : 4;
            }
            if (absCoeff[idx] >= 2) firstCoeff2 = 0;
          }
        }
      }
    }
  }
}

FHC_DH void code_coeff_nxn(const Ctx& c, const CUMeta* cu, int relIdxComp, const int16_t* coef, int log2n, int comp)
{
  FHC_TIC(cc);
  code_coeff_nxn_impl(c, cu, relIdxComp, coef, log2n, comp);
  FHC_TOC2(c, cc, PF_COEFBITS);
}

// ----------------------------------------------------------------------------------------------
// RDOQ sequential pass (TComTrQuant.cpp:2119-2661).  Inputs from OP_FWD_QUANT: tcoef[], lvlD[], cost0[],
// tmp[] = uiMaxAbsLevel; output: levels in tmp[] (raster, int32); returns uiAbsSum.
// ----------------------------------------------------------------------------------------------
struct QpParam { int per, rem; };
FHC_DH QpParam qp_param(const Ctx& c, int comp)                                                // TComTrQuant.cpp:71-119
{
  const int bdOff = 6 * (c.s->fp.bitDepth - 8);
  int base;
  if (comp == 0) base = c.s->fp.sp.sliceQp + bdOff;
  else
  {
    base = clip3(-bdOff, 57, c.s->fp.sp.sliceQp);
    base = base < 0 ? base + bdOff : c.s->t->chromaScale[base] + bdOff;
  }
  QpParam q; q.per = base / 6; q.rem = base % 6;
  return q;
}
FHC_DH int quant_scale(int rem) { return rem == 0 ? 26214 : rem == 1 ? 23302 : rem == 2 ? 20560 : rem == 3 ? 18396 : rem == 4 ? 16384 : 14564; }
FHC_DH int inv_quant_scale(int rem) { return rem == 0 ? 40 : rem == 1 ? 45 : rem == 2 ? 51 : rem == 3 ? 57 : rem == 4 ? 64 : 72; }

FHC_DH int ic_rate(const Engine* e, uint32_t absLevel, int ctxOne, int ctxAbs, uint32_t goRice, uint32_t c1Idx, uint32_t c2Idx)
{
  int rate = 32768;
  const uint32_t baseLevel = (c1Idx < 8) ? (2 + (c2Idx < 1)) : 1;
  if (absLevel >= baseLevel)
  {
    uint32_t symbol = absLevel - baseLevel, length;
    if (symbol < (3u << goRice)) { length = symbol >> goRice; rate += (int)((length + 1 + goRice) << 15); }
    else
    {
      length = goRice;
      symbol = symbol - (3u << goRice);
      FHC_NOUNROLL while (symbol >= (1u << length)) symbol -= (1u << (length++));
      rate += (int)((3 + length + 1 - goRice + length) << 15);
    }
    if (c1Idx < 8)
    {
      rate += ctx_bits(e, ctxOne, 1);
      if (c2Idx < 1) rate += ctx_bits(e, ctxAbs, 1);
    }
  }
  else if (absLevel == 1) rate += ctx_bits(e, ctxOne, 0);
  else if (absLevel == 2) { rate += ctx_bits(e, ctxOne, 1); rate += ctx_bits(e, ctxAbs, 0); }
  else rate = 0;
  return rate;
}
FHC_DH uint32_t coded_level(const Engine* e, double lambda, double* codedCost, double* codedCost0, double* codedCostSig, int levelDouble,
                            uint32_t maxAbsLevel, double sigCost0, double sigCost1, int ctxOne, int ctxAbs, uint32_t goRice, uint32_t c1Idx,
                            uint32_t c2Idx, int qBits, double errScale, int isLast)
{
  // sigCost0/1 = lambda * bits(sig_coeff_flag = 0/1) of this position's context (xGetRateSigCoef)
  double currCostSig = 0;
  uint32_t best = 0;
  if (!isLast && maxAbsLevel < 3)
  {
    *codedCostSig = sigCost0;
    *codedCost = dadd(*codedCost0, *codedCostSig);
    if (maxAbsLevel == 0) return 0;
  }
  else *codedCost = FHC_MAX_COST;
  if (!isLast) currCostSig = sigCost1;
  const uint32_t minAbs = maxAbsLevel > 1 ? maxAbsLevel - 1 : 1;
  FHC_NOUNROLL for (int absLevel = (int)maxAbsLevel; absLevel >= (int)minAbs; absLevel--)
  {
    const double err = (double)(levelDouble - (int)((uint32_t)absLevel << qBits));
    double cur = dadd(dmul(dmul(err, err), errScale), dmul(lambda, (double)ic_rate(e, (uint32_t)absLevel, ctxOne, ctxAbs, goRice, c1Idx, c2Idx)));
    cur = dadd(cur, currCostSig);
    if (cur < *codedCost) { best = (uint32_t)absLevel; *codedCost = cur; *codedCostSig = currCostSig; }
  }
  return best;
}

// rateIncUp / rateIncDown of TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:2316-2326) for one coefficient, from the
// state of the level coder (packed) and the level chosen when the coefficient was quantised.  Evaluated only for the
// coefficient groups that sign bit hiding actually changes; contexts do not move while a TU is being quantised.
FHC_DH void sbh_rate_inc(const Engine* e, int packed, int level, int* incUp, int* incDown)
{
  const int ctxOne = packed & 255, ctxAbs = (packed >> 8) & 255;
  const uint32_t goRice = (uint32_t)(packed >> 16) & 7u, c1Idx = ((packed >> 20) & 1) ? 0u : 8u, c2Idx = ((packed >> 21) & 1) ? 0u : 1u;
  if (level > 0)
  {
    const int rateNow = ic_rate(e, (uint32_t)level, ctxOne, ctxAbs, goRice, c1Idx, c2Idx);
    *incUp = ic_rate(e, (uint32_t)level + 1, ctxOne, ctxAbs, goRice, c1Idx, c2Idx) - rateNow;
    *incDown = ic_rate(e, (uint32_t)level - 1, ctxOne, ctxAbs, goRice, c1Idx, c2Idx) - rateNow;
  }
  else { *incUp = ctx_bits(e, ctxOne, 0); *incDown = 0; }
}

FHC_DH uint32_t rdoq_pass_impl(const Ctx& c, const CUMeta* cu, int relIdxComp, int comp, int log2n, int cbfCtx)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  const int isChroma = comp != 0, n = 1 << log2n, numCoef = n * n, bd = c.s->fp.bitDepth;
  const QpParam q = qp_param(c, comp);
  const int qBits = 14 + q.per + (15 - bd - log2n);
  const double errScale = s->t->errScale[isChroma][log2n - 2][q.rem];
  const double lambda = e->rdoqLambda;
  const int scanType = coef_scan_idx(cu, relIdxComp, n, comp);
  const uint16_t* scan = scan_tab(*s->t, scanType, log2n);
  const uint8_t* scanCG = scan_cg(*s->t, scanType, log2n);
  const int wg = n >> 2, cgNum = numCoef >> 4;
  const int firstSig = first_sig_ctx(log2n, scanType, isChroma);
  const int sigBase = CTX_SIG + (isChroma ? 28 : 0);
  const int32_t* src = e->tcoef;
  int32_t* dst = e->tmp;
  double* costCoeff0 = e->cost0; double* costCoeff = e->costCoeff; double* costSig = e->costSig;
  int32_t* rateIncUp = e->rateIncUp; int32_t* rateIncDown = e->rateIncDown;
  int32_t* sigRateDelta = e->sigRateDelta; int32_t* deltaU = e->deltaU;
  double* costCGSig = e->costCGSig;
  uint8_t* cgFlag = e->cgFlag;
  double blockUncoded = 0, baseCost = 0;
  int cgLastScanPos = -1, lastScanPos = -1, c1 = 1, c2 = 0;
  uint32_t ctxSet = 0, c1Idx = 0, c2Idx = 0, goRice = 0, absSum = 0;
  int bestLastIdxP1 = 0;

  // no candidate level anywhere: every coefficient stays 0 and nothing below is read again (TComTrQuant.cpp:2462
  // returns with uiAbsSum = 0 after summing costs nobody uses)
  if (!(e->cgNZ[0] | e->cgNZ[1])) return 0;
  FHC_NOUNROLL for (int i = 0; i < cgNum; i++) { costCGSig[i] = 0; cgFlag[i] = 0; }

  // lambda * bits of sig_coeff_flag = 0 / 1 and the rate difference, per context of this channel (contexts do not
  // change while a TU is being quantised): replaces estBitsSbacStruct::significantBits (TEncSbac.cpp:1772)
  double* sigC0 = e->sigC0; double* sigC1 = e->sigC1;
  int32_t* sigDelta = e->sigDelta;
  const int useTab = numCoef > 16;          // a 4x4 TU looks up at most 16 contexts: cheaper without the table
  if (useTab)
  {
    const int nSig = isChroma ? 16 : 28;
    FHC_NOUNROLL for (int i = 0; i < nSig; i++)
    {
      const int b0 = ctx_bits(e, sigBase + i, 0), b1 = ctx_bits(e, sigBase + i, 1);
      sigC0[i] = dmul(lambda, (double)b0); sigC1[i] = dmul(lambda, (double)b1); sigDelta[i] = b1 - b0;
    }
  }
#define FHC_SIGC0(sc) (useTab ? sigC0[sc] : dmul(lambda, (double)ctx_bits(e, sigBase + (sc), 0)))
#define FHC_SIGC1(sc) (useTab ? sigC1[sc] : dmul(lambda, (double)ctx_bits(e, sigBase + (sc), 1)))
#define FHC_SIGDELTA(sc) (useTab ? sigDelta[sc] : ctx_bits(e, sigBase + (sc), 1) - ctx_bits(e, sigBase + (sc), 0))

  FHC_NOUNROLL for (int cgScanPos = cgNum - 1; cgScanPos >= 0; cgScanPos--)
  {
    const int cgBlk = scanCG[cgScanPos], cgY = cgBlk >> (log2n - 2), cgX = cgBlk & (wg - 1);
    const int cgHasLevel = (e->cgNZ[cgScanPos >> 5] >> (cgScanPos & 31)) & 1;
    if (!cgHasLevel && lastScanPos < 0)
    {
      // nothing coded yet and nothing to code here: only the uncoded distortion accumulates
      FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
      {
        const double c0 = costCoeff0[cgScanPos * 16 + posInCG];
        blockUncoded = dadd(blockUncoded, c0);
        baseCost = dadd(baseCost, c0);
      }
      continue;
    }
    const int pat = pattern_sig_ctx(cgFlag, cgX, cgY, wg);
    if (!cgHasLevel && cgScanPos > 0)
    {
      // group without any candidate level below the last position: every coefficient takes "level 0" and the group
      // ends up with coded_sub_block_flag = 0; none of the per-coefficient arrays of this group is read again
      double sigCost = 0;
      FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
      {
        const int scanPos = cgScanPos * 16 + posInCG;
        const double c0 = costCoeff0[scanPos];
        const double cs = FHC_SIGC0(sig_ctx_inc(*e->t, pat, firstSig, (int)scan[scanPos], log2n, isChroma));
        blockUncoded = dadd(blockUncoded, c0);
        baseCost = dadd(baseCost, dadd(c0, cs));
        sigCost = dadd(sigCost, cs);
      }
      ctxSet = (uint32_t)((isChroma ? 4 : 0) + ((!isChroma && (cgScanPos - 1) > 0) ? 2 : 0) + (c1 == 0 ? 1 : 0));
      c1 = 1; c2 = 0; c1Idx = 0; c2Idx = 0; goRice = 0;
      const int cx = CTX_SIGCG + (isChroma ? 2 : 0) + sig_cg_ctx(cgFlag, cgX, cgY, wg);
      const double cgc = dmul(lambda, (double)ctx_bits(e, cx, 0));
      baseCost = dadd(baseCost, dsub(cgc, sigCost));
      costCGSig[cgScanPos] = cgc;
      continue;
    }
    double sigCost = 0, sigCost0 = 0, codedLevelAndDist = 0, uncodedDist = 0;
    int nnzBeforePos0 = 0;
    FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
    {
      const int scanPos = cgScanPos * 16 + posInCG;
      const uint32_t blkPos = scan[scanPos];
      const int levelDouble = e->lvlD[scanPos];
      const uint32_t maxAbsLevel = (uint32_t)dst[blkPos];
      blockUncoded = dadd(blockUncoded, costCoeff0[scanPos]);
      if (lastScanPos < 0)
      {
        if (maxAbsLevel == 0)
        {
          baseCost = dadd(baseCost, costCoeff0[scanPos]);
          continue;
        }
        lastScanPos = scanPos;
        ctxSet = (uint32_t)((isChroma ? 4 : 0) + ((!isChroma && (scanPos >> 4) > 0) ? 2 : 0));
        cgLastScanPos = cgScanPos;
      }
      const int ctxOne = CTX_ONE + 4 * (int)ctxSet + c1;
      if (maxAbsLevel == 0)
      {
        // "level 0" is the only candidate (xGetCodedLevel returns early); no state of the level coder changes
        const int sc = sig_ctx_inc(*e->t, pat, firstSig, (int)blkPos, log2n, isChroma);
        const double cs = FHC_SIGC0(sc), cc = dadd(costCoeff0[scanPos], cs);
        costSig[scanPos] = cs; costCoeff[scanPos] = cc;
        sigRateDelta[blkPos] = sc;                                     // context only: the delta is evaluated by sign bit hiding
        deltaU[blkPos] = levelDouble >> (qBits - 8);
        rateIncUp[blkPos] = ctxOne; rateIncDown[blkPos] = 0;          // level-coder state and level: see sbh_rate_inc
        baseCost = dadd(baseCost, cc);
        if (posInCG == 0 && scanPos > 0)
        {
          ctxSet = (uint32_t)((isChroma ? 4 : 0) + ((!isChroma && ((scanPos - 1) >> 4) > 0) ? 2 : 0) + (c1 == 0 ? 1 : 0));
          c1 = 1; c2 = 0; c1Idx = 0; c2Idx = 0; goRice = 0;
        }
        sigCost = dadd(sigCost, cs);
        if (posInCG == 0) sigCost0 = cs;
        continue;
      }
      {
        uint32_t level;
        const int ctxAbs = CTX_ABS + (int)ctxSet + c2;
        costCoeff[scanPos] = 0; costSig[scanPos] = 0;
        sigRateDelta[blkPos] = -1;                                     // last position: no sig_coeff_flag
        if (scanPos == lastScanPos)
          level = coded_level(e, lambda, &costCoeff[scanPos], &costCoeff0[scanPos], &costSig[scanPos], levelDouble, maxAbsLevel,
                              0.0, 0.0, ctxOne, ctxAbs, goRice, c1Idx, c2Idx, qBits, errScale, 1);
        else
        {
          const int sc = sig_ctx_inc(*e->t, pat, firstSig, (int)blkPos, log2n, isChroma);
          level = coded_level(e, lambda, &costCoeff[scanPos], &costCoeff0[scanPos], &costSig[scanPos], levelDouble, maxAbsLevel,
                              FHC_SIGC0(sc), FHC_SIGC1(sc), ctxOne, ctxAbs, goRice, c1Idx, c2Idx, qBits, errScale, 0);
          sigRateDelta[blkPos] = sc;
        }
        deltaU[blkPos] = (levelDouble - (int)(level << qBits)) >> (qBits - 8);
        // rate increments of level +-1 are only needed if sign bit hiding touches this group: keep the level-coder state
        // and the level chosen here (TComTrQuant.cpp:2316-2326 computes them eagerly), sbh_rate_inc evaluates them
        rateIncUp[blkPos] = ctxOne | (ctxAbs << 8) | ((int)goRice << 16) | ((c1Idx < 8) << 20) | ((c2Idx < 1) << 21);
        rateIncDown[blkPos] = (int32_t)level;
        dst[blkPos] = (int32_t)level;
        baseCost = dadd(baseCost, costCoeff[scanPos]);
        const int baseLevel = (c1Idx < 8) ? (2 + (c2Idx < 1)) : 1;
        if ((int)level >= baseLevel)
          if (level > 3u * (1u << goRice)) goRice = goRice + 1 < 4 ? goRice + 1 : 4;
        if (level >= 1) c1Idx++;
        if (level > 1) { c1 = 0; c2 += (c2 < 2); c2Idx++; }
        else if (c1 < 3 && c1 > 0 && level) c1++;
        if ((scanPos % 16 == 0) && scanPos > 0)
        {
          ctxSet = (uint32_t)((isChroma ? 4 : 0) + ((!isChroma && ((scanPos - 1) >> 4) > 0) ? 2 : 0) + (c1 == 0 ? 1 : 0));
          c1 = 1; c2 = 0; c1Idx = 0; c2Idx = 0; goRice = 0;
        }
      }
      sigCost = dadd(sigCost, costSig[scanPos]);
      if (posInCG == 0) sigCost0 = costSig[scanPos];
      if (dst[blkPos])
      {
        cgFlag[cgBlk] = 1;
        codedLevelAndDist = dadd(codedLevelAndDist, dsub(costCoeff[scanPos], costSig[scanPos]));
        uncodedDist = dadd(uncodedDist, costCoeff0[scanPos]);
        if (posInCG != 0) nnzBeforePos0++;
      }
    }
    if (cgLastScanPos >= 0)
    {
      if (cgScanPos)
      {
        if (cgFlag[cgBlk] == 0)
        {
          const int cx = CTX_SIGCG + (isChroma ? 2 : 0) + sig_cg_ctx(cgFlag, cgX, cgY, wg);
          baseCost = dadd(baseCost, dsub(dmul(lambda, (double)ctx_bits(e, cx, 0)), sigCost));
          costCGSig[cgScanPos] = dmul(lambda, (double)ctx_bits(e, cx, 0));
        }
        else if (cgScanPos < cgLastScanPos)
        {
          if (nnzBeforePos0 == 0) { baseCost = dsub(baseCost, sigCost0); sigCost = dsub(sigCost, sigCost0); }
          double costZeroCG = baseCost;
          const int cx = CTX_SIGCG + (isChroma ? 2 : 0) + sig_cg_ctx(cgFlag, cgX, cgY, wg);
          baseCost = dadd(baseCost, dmul(lambda, (double)ctx_bits(e, cx, 1)));
          costZeroCG = dadd(costZeroCG, dmul(lambda, (double)ctx_bits(e, cx, 0)));
          costCGSig[cgScanPos] = dmul(lambda, (double)ctx_bits(e, cx, 1));
          costZeroCG = dadd(costZeroCG, uncodedDist);
          costZeroCG = dsub(costZeroCG, codedLevelAndDist);
          costZeroCG = dsub(costZeroCG, sigCost);
          if (costZeroCG < baseCost)
          {
            cgFlag[cgBlk] = 0;
            baseCost = costZeroCG;
            costCGSig[cgScanPos] = dmul(lambda, (double)ctx_bits(e, cx, 0));
            FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
            {
              const int scanPos = cgScanPos * 16 + posInCG;
              const uint32_t blkPos = scan[scanPos];
              if (dst[blkPos]) { dst[blkPos] = 0; costCoeff[scanPos] = costCoeff0[scanPos]; costSig[scanPos] = 0; }
            }
          }
        }
      }
      else cgFlag[cgBlk] = 1;
    }
  }
  if (lastScanPos < 0) return 0;

  double bestCost = dadd(blockUncoded, dmul(lambda, (double)ctx_bits(e, cbfCtx, 0)));
  baseCost = dadd(baseCost, dmul(lambda, (double)ctx_bits(e, cbfCtx, 1)));
  {
    int foundLast = 0, off, shift;
    last_ctx_params(isChroma, log2n, &off, &shift);
    const int bx0 = CTX_LASTX + (isChroma ? 15 : 0) + off, by0 = CTX_LASTY + (isChroma ? 15 : 0) + off, gmax = group_idx(n - 1);
    FHC_NOUNROLL for (int cgScanPos = cgLastScanPos; cgScanPos >= 0; cgScanPos--)
    {
      const int cgBlk = scanCG[cgScanPos];
      baseCost = dsub(baseCost, costCGSig[cgScanPos]);
      if (cgFlag[cgBlk])
      {
        FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
        {
          const int scanPos = cgScanPos * 16 + posInCG;
          if (scanPos > lastScanPos) continue;
          const uint32_t blkPos = scan[scanPos];
          if (dst[blkPos])
          {
            const uint32_t posY = blkPos >> log2n, posX = blkPos - (posY << log2n);
            const uint32_t lx = scanType == SCAN_VER ? posY : posX, ly = scanType == SCAN_VER ? posX : posY;
            const int gx = group_idx((int)lx), gy = group_idx((int)ly);
            int bx = 0, by = 0, cc;
            FHC_NOUNROLL for (cc = 0; cc < gx; cc++) bx += ctx_bits(e, bx0 + (cc >> shift), 1);
            if (gx < gmax) bx += ctx_bits(e, bx0 + (gx >> shift), 0);
            FHC_NOUNROLL for (cc = 0; cc < gy; cc++) by += ctx_bits(e, by0 + (cc >> shift), 1);
            if (gy < gmax) by += ctx_bits(e, by0 + (gy >> shift), 0);
            double cost = (double)(bx + by);
            if (gx > 3) cost = dadd(cost, dmul(32768.0, (double)((gx - 2) >> 1)));
            if (gy > 3) cost = dadd(cost, dmul(32768.0, (double)((gy - 2) >> 1)));
            const double costLast = dmul(lambda, cost);
            const double totalCost = dsub(dadd(baseCost, costLast), costSig[scanPos]);
            if (totalCost < bestCost) { bestLastIdxP1 = scanPos + 1; bestCost = totalCost; }
            if (dst[blkPos] > 1) { foundLast = 1; break; }
            baseCost = dsub(baseCost, costCoeff[scanPos]);
            baseCost = dadd(baseCost, costCoeff0[scanPos]);
          }
          else baseCost = dsub(baseCost, costSig[scanPos]);
        }
        if (foundLast) break;
      }
    }
  }
  // signs, and per coefficient group of the coded range the non-zero map and the level sum (consumed by sign bit hiding)
  uint16_t* nzMask = e->cgNzMask;
  uint8_t* cgAbs = e->cgAbs;
  const int numCodedCG = (bestLastIdxP1 + 15) >> 4;
  FHC_NOUNROLL for (int cgi = 0; cgi < numCodedCG; cgi++)
  {
    const int p0 = cgi << 4, p1 = imin(bestLastIdxP1, p0 + 16);
    uint32_t mask = 0;
    int sum = 0;
    FHC_NOUNROLL for (int scanPos = p0; scanPos < p1; scanPos++)
    {
      const int blkPos = scan[scanPos];
      const int32_t level = dst[blkPos];
      if (level)
      {
        mask |= 1u << (scanPos - p0);
        sum += level;
        if (src[blkPos] < 0) dst[blkPos] = -level;
      }
    }
    absSum += (uint32_t)sum;
    nzMask[cgi] = (uint16_t)mask; cgAbs[cgi] = (uint8_t)(sum & 1);
  }
  FHC_NOUNROLL for (int scanPos = bestLastIdxP1; scanPos <= lastScanPos; scanPos++) dst[scan[scanPos]] = 0;

  if (absSum >= 2)                                                             // sign bit hiding (TComTrQuant.cpp:2530)
  {
    const double invQ = (double)inv_quant_scale(q.rem);
#ifdef FHC_DEVICE_BUILD
    const double f = __dadd_rn(__ddiv_rn(__ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(invQ, invQ), (double)(1 << (2 * q.per))), lambda), 16.0),
                                         (double)(1 << (2 * (bd - 8)))), 0.5);
#else
    const double f = invQ * invQ * (double)(1 << (2 * q.per)) / lambda / 16 / (double)(1 << (2 * (bd - 8))) + 0.5;
#endif
    const int64_t rdFactor = (int64_t)f;
    int lastCG = -1;
    // groups above the coded range hold no level and change nothing (TComTrQuant.cpp:2541-2564 would find lastNZ = -1)
    FHC_NOUNROLL for (int subSet = numCodedCG - 1; subSet >= 0; subSet--)
    {
      const int subPos = subSet << 4;
      const uint32_t mask = nzMask[subSet];
      int nn;
      if (!mask) continue;
      const int lastNZ = msb_index(mask), firstNZ = lsb_index(mask), sum = cgAbs[subSet];     // sum: only its parity is used
      if (lastCG == -1) lastCG = 1;
      if (lastNZ - firstNZ >= 4)
      {
        const uint32_t signbit = dst[scan[subPos + firstNZ]] > 0 ? 0 : 1;
        if (signbit != (uint32_t)(sum & 1))
        {
          int64_t minCostInc = INT64_MAX, curCost = INT64_MAX;
          int minPos = -1, finalChange = 0, curChange = 0;
          FHC_NOUNROLL for (nn = (lastCG == 1 ? lastNZ : 15); nn >= 0; --nn)
          {
            const uint32_t blkPos = scan[nn + subPos];
            if (dst[blkPos] != 0)
            {
              int incUp, incDown;
              sbh_rate_inc(e, rateIncUp[blkPos], rateIncDown[blkPos], &incUp, &incDown);
              const int64_t costUp = rdFactor * (-deltaU[blkPos]) + incUp;
              const int sdc = sigRateDelta[blkPos], sigD = sdc < 0 ? 0 : FHC_SIGDELTA(sdc);
              int64_t costDown = rdFactor * (deltaU[blkPos]) + incDown - ((iabs(dst[blkPos]) == 1) ? sigD : 0);
              if (lastCG == 1 && lastNZ == nn && iabs(dst[blkPos]) == 1) costDown -= (4 << 15);
              if (costUp < costDown) { curCost = costUp; curChange = 1; }
              else
              {
                curChange = -1;
                if (nn == firstNZ && iabs(dst[blkPos]) == 1) curCost = INT64_MAX;
                else curCost = costDown;
              }
            }
            else
            {
              int incUp, incDown;
              sbh_rate_inc(e, rateIncUp[blkPos], rateIncDown[blkPos], &incUp, &incDown);
              const int sdc = sigRateDelta[blkPos], sigD = sdc < 0 ? 0 : FHC_SIGDELTA(sdc);
              curCost = rdFactor * (-(iabs(deltaU[blkPos]))) + (1 << 15) + incUp + sigD;
              curChange = 1;
              if (nn < firstNZ)
              {
                const uint32_t thisSign = src[blkPos] >= 0 ? 0 : 1;
                if (thisSign != signbit) curCost = INT64_MAX;
              }
            }
            if (curCost < minCostInc) { minCostInc = curCost; finalChange = curChange; minPos = (int)blkPos; }
          }
          if (dst[minPos] == 32767 || dst[minPos] == -32768) finalChange = -1;
          if (src[minPos] >= 0) dst[minPos] += finalChange; else dst[minPos] -= finalChange;
        }
      }
      if (lastCG == 1) lastCG = 0;
    }
  }
  return absSum;
}

FHC_DH uint32_t rdoq_pass(const Ctx& c, const CUMeta* cu, int relIdxComp, int comp, int log2n, int cbfCtx)
{
  FHC_TIC(r);
  const uint32_t v = rdoq_pass_impl(c, cu, relIdxComp, comp, log2n, cbfCtx);
  FHC_TOC2(c, r, PF_RDOQ);
  return v;
}

// ----------------------------------------------------------------------------------------------
// geometry helpers
// ----------------------------------------------------------------------------------------------
FHC_DH int rel_x(const Shared* s, int relIdx) { return (s->t->z2r[relIdx] & 15) * 4; }
FHC_DH int rel_y(const Shared* s, int relIdx) { return (s->t->z2r[relIdx] >> 4) * 4; }
FHC_DH CUMeta* best_cu(Shared* s, int d) { return &s->cu[d][s->bestIdx[d]]; }
FHC_DH CUMeta* temp_cu(Shared* s, int d) { return &s->cu[d][1 - s->bestIdx[d]]; }
FHC_DH int cu_buf_index(const Shared* s, const CUMeta* cu) { return (int)(cu - &s->cu[cu->depth][0]); }

// ----------------------------------------------------------------------------------------------
// reference samples: availability flags by the calling engine, gather/substitute/filter as a COOP section
// (TComPattern.cpp:115-745, TComDataCU.cpp:1090-1200)
// ----------------------------------------------------------------------------------------------
FHC_D void build_refs(const Ctx& c, const CUMeta* cu, int relIdx, int comp, int n, int wantFilt)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  const FrameGlobal* g = c.g;
  const int zLT = cu->zidx + relIdx, rLT = s->t->z2r[zLT];
  const int ux = rLT & 15, uy = rLT >> 4;
  const int us = comp ? 2 : 4, nu = n / us, total = 4 * nu + 1;
  uint8_t* flags = e->flags;
  int numAvail = 0;
  FHC_TIC(rf);
  const int leftOk = (ux > 0 || s->ctuX > 0), aboveOk = (uy > 0 || s->ctuY > 0);
  flags[2 * nu] = (uint8_t)(leftOk && aboveOk);
  FHC_NOUNROLL for (int i = 0; i < nu; i++) { flags[2 * nu + 1 + i] = (uint8_t)aboveOk; flags[2 * nu - 1 - i] = (uint8_t)leftOk; }
  {
    const int rRT = rLT + nu - 1, zRT = s->t->r2z[rRT];
    FHC_NOUNROLL for (int i = 1; i <= nu; i++)
    {
      int av;
      if (s->ctuPelX + (rRT & 15) * 4 + 4 * i >= s->fp.width) av = 0;
      else if ((rRT & 15) < 16 - i) av = (uy != 0) ? (zRT > s->t->r2z[rRT - 16 + i]) : (s->ctuY > 0);
      else if (uy != 0) av = 0;
      else av = (s->ctuY > 0) && (s->ctuX + 1 < s->fp.wc);
      flags[3 * nu + i] = (uint8_t)av;
    }
    const int rLB = rLT + (nu - 1) * 16, zLB = s->t->r2z[rLB];
    FHC_NOUNROLL for (int i = 1; i <= nu; i++)
    {
      int av;
      if (s->ctuPelY + (rLB >> 4) * 4 + 4 * i >= s->fp.height) av = 0;
      else if ((rLB >> 4) < 16 - i) av = (ux != 0) ? (zLB > s->t->r2z[rLB + 16 * i - 1]) : (s->ctuX > 0);
      else av = 0;
      flags[nu - i] = (uint8_t)av;
    }
  }
  FHC_NOUNROLL for (int i = 0; i < total; i++) numAvail += flags[i];
  Cmd& m = e->cmd;
  const int px = (s->ctuPelX + ux * 4) >> (comp ? 1 : 0), py = (s->ctuPelY + uy * 4) >> (comp ? 1 : 0);
  m.a[0] = comp; m.a[1] = n; m.a[2] = nu; m.a[3] = us; m.a[4] = py * plane_stride(s, comp) + px; m.a[5] = wantFilt;
  m.a[6] = numAvail; m.a[7] = total;
  FHC_TOC2(c, rf, PF_REFS_MASTER);
  par(c, OP_REFS);
}

template <int N>
FHC_DH void bind_coef_scratch(Engine* e, CoefScratch<N>* x)
{
  e->resi = x->resi; e->tmp = x->tmp; e->tcoef = x->tcoef; e->lvlD = x->lvlD;
  e->cost0 = x->cost0; e->costCoeff = x->costCoeff; e->costSig = x->costSig;
  e->rateIncUp = x->rateIncUp; e->rateIncDown = x->rateIncDown; e->sigRateDelta = x->sigRateDelta; e->deltaU = x->deltaU;
}

// ----------------------------------------------------------------------------------------------
// xIntraCodingTUBlock (TEncSearch.cpp:1104-1399)
// ----------------------------------------------------------------------------------------------
FHC_D uint32_t code_tu_block(const Ctx& c, CUMeta* cu, int relIdx, int comp, int log2n, int lumaLog2, int trDepthBit, int numParts,
                             int predSaveLoad)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  FrameGlobal* g = c.g;
  const int n = 1 << log2n, isChroma = comp != 0;
  const int cuW = cu->width >> isChroma;
  block_align(c);
  FHC_COUNT(g_emuCounters.tqSamples[isChroma][log2n - 2] += (unsigned)(n * n); g_emuCounters.tuCalls[isChroma][log2n - 2]++);
  const int bx = rel_x(s, relIdx) >> isChroma, by = rel_y(s, relIdx) >> isChroma;
  const int qtLayer = 5 - lumaLog2;
  const int picX = (cu->x >> isChroma) + bx, picY = (cu->y >> isChroma) + by;
  const int stride = plane_stride(s, comp);
  const int predOff = e->coff[comp] + by * cuW + bx;
  const int qtOff = e->coff[comp] + by * (e->qs >> isChroma) + bx;
  const int coefOff = e->coff[comp] + (isChroma ? relIdx * 4 : relIdx * 16);
  const int ts = cu->ts[comp][relIdx];
  const QpParam q = qp_param(c, comp);
  int mode;
  Cmd& m = e->cmd;
  FHC_TIC(tu);
  // coefficient scratch by TU size: the small one lives in the warp's shared-memory slice
  if (e->kind == ENG_MASTER)
  {
    if (n * n <= FHC_MASTER_SMALL) bind_coef_scratch(e, (CoefScratch<FHC_MASTER_SMALL>*)e->csSmall);
    else bind_coef_scratch(e, (CoefScratch<1024>*)e->csLarge);
    if (n * n <= 64) bind_group_scratch(e, (GroupScratch<4>*)e->gsSmall); else bind_group_scratch(e, (GroupScratch<64>*)e->gsLarge);
  }
  else
  {
    if (n * n <= 16) bind_coef_scratch(e, (CoefScratch<16>*)e->csSmall);
    else bind_coef_scratch(e, (CoefScratch<FHC_SMALL_COEFS>*)e->csLarge);
  }

  if (comp == 0) mode = cu->ldir[relIdx];
  else
  {
    mode = cu->cdir[relIdx];
    if (mode == DM_CHROMA_IDX) mode = cu->ldir[relIdx & ~3];
  }
  int useFilt = 0;
  if (predSaveLoad != 2)
  {
    useFilt = use_filtered_refs(*s->t, comp, mode, log2n);
    if (e->refsShared == 2 && comp != 0) { e->refU = s->refC[comp - 1]; e->refDc = s->dcC[comp - 1]; }
    else if (!(e->refsShared == 1 && comp == 0)) build_refs(c, cu, relIdx, comp, n, useFilt);
  }
  m.a[0] = comp; m.a[1] = log2n; m.a[2] = mode; m.a[3] = useFilt; m.a[4] = predOff; m.a[5] = cuW;
  m.a[6] = picY * stride + picX; m.a[7] = predSaveLoad;
  fine_align(c);
  par(c, OP_PRED_RESI);

  // transform + RDOQ (TComTrQuant::transformNxN :1450)
  e->rdoqLambda = s->fp.sp.rdoqLambda[comp];
  const int scanType = coef_scan_idx(cu, relIdx, n, comp);
  m.a[0] = log2n; m.a[1] = (comp == 0 && log2n == 2); m.a[2] = ts; m.a[3] = scanType; m.a[4] = quant_scale(q.rem);
  m.a[5] = 14 + q.per + (15 - s->fp.bitDepth - log2n); m.a[6] = isChroma; m.a[7] = q.rem;
  e->cgNZ[0] = 0; e->cgNZ[1] = 0;
  fine_align(c);
  par(c, OP_FWD_QUANT);
  const int cbfCtx = CTX_CBF + (isChroma ? 5 + trDepthBit : (trDepthBit == 0 ? 1 : 0));
  mid_align(c);
  const uint32_t absSum = rdoq_pass(c, cu, relIdx, comp, log2n, cbfCtx);
  set_range(cu->cbf[comp], relIdx, numParts, (absSum > 0 ? 1 : 0) << trDepthBit);

  // levels -> QT coefficient layer
  fine_align(c);
  if (absSum > 0) { m.a[0] = n * n; m.a[1] = qtLayer; m.a[2] = coefOff; par(c, OP_LEVELS_OUT); }
  else par_zero16(c, qt_coef(e, qtLayer) + coefOff, n * n);

  // inverse + reconstruction + SSE
  fine_align(c);
  e->sse = 0;
  m.a[0] = comp; m.a[1] = log2n; m.a[2] = (comp == 0 && log2n == 2); m.a[3] = ts; m.a[4] = absSum > 0; m.a[5] = q.per; m.a[6] = q.rem;
  m.a[7] = predOff; m.a[8] = cuW; m.a[9] = qtLayer; m.a[10] = qtOff; m.a[11] = picY * stride + picX; m.a[12] = picY * stride + picX;
  par(c, OP_INV_RECON);
  uint32_t dist = e->sse;
  if (isChroma) dist = (uint32_t)dmul(s->fp.sp.chromaDistWeight[comp - 1], (double)dist);          // TComRdCost.cpp:352
  FHC_TOC2(c, tu, PF_TUBLOCK);
  return dist;
}

// ----------------------------------------------------------------------------------------------
// search-time bit counting (TEncSearch.cpp:881-1102)
// ----------------------------------------------------------------------------------------------
FHC_DH int min_tu_log2_in_cu(const CUMeta* cu)                                                   // TComDataCU.cpp:1478
{
  const int log2Cb = ilog2(cu->width), split = cu->part[0] == SIZE_NxN;
  if (log2Cb < 2 + 3 - 1 + split) return 2;
  const int r = log2Cb - (3 - 1 + split);
  return r > 5 ? 5 : r;
}
FHC_DH void enc_intra_header(const Ctx& c, const CUMeta* cu, int trDepth, int relIdx, int bLuma, int bChroma)
{
  if (bLuma)
  {
    if (relIdx == 0) code_part_size(c, cu, 0, cu->depth);
    if (cu->part[0] == SIZE_2Nx2N) { if (relIdx == 0) code_luma_dir(c, cu, 0, 0); }
    else if (trDepth > 0 && (relIdx % (cu->nparts >> 2)) == 0) code_luma_dir(c, cu, relIdx, 0);
  }
  if (bChroma && relIdx == 0) code_chroma_dir(c, cu, relIdx);
}
FHC_DH void code_qt_cbf(const Ctx& c, const CUMeta* cu, int relIdxComp, int comp, int trDepth, int lowestLevel, int compLog2n)
{
  const int isChroma = comp != 0;
  const int ctx = CTX_CBF + (isChroma ? 5 + trDepth : (trDepth == 0 ? 1 : 0));
  const int lowest = trDepth + ((!lowestLevel && compLog2n < 3) ? 1 : 0);
  enc_bin(c.e, ctx, (cu->cbf[comp][relIdxComp] >> lowest) & 1);
}
FHC_D void enc_subdiv_cbf(const Ctx& c, const CUMeta* cu, int relIdx, int log2n, int trDepth, int chromaCodeAll, int bLuma, int bChroma)
{
  const int subdiv = cu->tridx[relIdx] > trDepth;
  if (cu->part[0] == SIZE_NxN && trDepth == 0) { }
  else if (log2n > 5) { }
  else if (log2n == 2) { }
  else if (log2n == min_tu_log2_in_cu(cu)) { }
  else if (bLuma) enc_bin(c.e, CTX_SUBDIV + (5 - log2n), subdiv);
  if (bChroma)
    FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
      if (chromaCodeAll && (trDepth == 0 || ((cu->cbf[cc][relIdx] >> (trDepth - 1)) & 1)))
        code_qt_cbf(c, cu, relIdx, cc, trDepth, subdiv == 0, log2n - 1);
  if (subdiv)
  {
    const int q = (cu->nparts >> (2 * trDepth)) >> 2;
    FHC_NOUNROLL for (int i = 0; i < 4; i++) enc_subdiv_cbf(c, cu, relIdx + i * q, log2n - 1, trDepth + 1, (log2n - 1) > 2, bLuma, bChroma);
  }
  else if (bLuma) code_qt_cbf(c, cu, relIdx, 0, trDepth, 1, log2n);
}
FHC_D void enc_coeff_qt(const Ctx& c, const CUMeta* cu, int relIdx, int log2n, int trDepth, int comp, int procChroma)
{
  if (cu->tridx[relIdx] > trDepth)
  {
    const int q = (cu->nparts >> (2 * trDepth)) >> 2;
    FHC_NOUNROLL for (int i = 0; i < 4; i++) enc_coeff_qt(c, cu, relIdx + i * q, log2n - 1, trDepth + 1, comp, (log2n - 1 > 2) ? 1 : (i == 0));
  }
  else if (comp == 0)
  {
    if ((cu->cbf[0][relIdx] >> trDepth) & 1) code_coeff_nxn(c, cu, relIdx, qt_coef(c.e, 5 - log2n) + relIdx * 16, log2n, 0);
  }
  else if (procChroma)
  {
    const int cl = log2n == 2 ? 2 : log2n - 1, ridx = log2n == 2 ? (relIdx & ~3) : relIdx;
    if ((cu->cbf[comp][relIdx] >> trDepth) & 1) code_coeff_nxn(c, cu, ridx, qt_coef(c.e, 5 - log2n) + c.e->coff[comp] + ridx * 4, cl, comp);
  }
}
FHC_D uint32_t intra_bits_qt(const Ctx& c, const CUMeta* cu, int relIdx, int log2n, int trDepth, int chromaCodeAll, int procChroma,
                             int bLuma, int bChroma)
{
  FHC_TIC(bq);
  mid_align(c);
  reset_bits(c.e);
  enc_intra_header(c, cu, trDepth, relIdx, bLuma, bChroma);
  enc_subdiv_cbf(c, cu, relIdx, log2n, trDepth, chromaCodeAll, bLuma, bChroma);
  if (bLuma) enc_coeff_qt(c, cu, relIdx, log2n, trDepth, 0, procChroma);
  if (bChroma) { enc_coeff_qt(c, cu, relIdx, log2n, trDepth, 1, procChroma); enc_coeff_qt(c, cu, relIdx, log2n, trDepth, 2, procChroma); }
  FHC_TOC2(c, bq, PF_BITSQT);
  return num_bits(c.e);
}

// ----------------------------------------------------------------------------------------------
// luma RQT search (TEncSearch.cpp:1404-1713)
// ----------------------------------------------------------------------------------------------
FHC_D void recur_luma_qt(const Ctx& c, CUMeta* cu, int relIdx, int log2n, int trDepth, int checkFirst, uint32_t* distY, double* rdCost)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  FrameGlobal* g = c.g;
  const int fullDepth = cu->depth + trDepth;
  const int numParts = cu->nparts >> (2 * trDepth);
  const int checkFull = log2n <= 5;
  int checkSplit = log2n > min_tu_log2_in_cu(cu);
  double singleCost = FHC_MAX_COST;
  uint32_t singleDist = 0, singleCbf = 0;
  int bestModeId = 0;
  const int checkTS = (log2n == 2) && (cu->part[relIdx] == SIZE_NxN);
  const int layer = 5 - log2n;
  int16_t* qtCoefBlk = qt_coef(e, layer < 0 ? 0 : layer) + relIdx * 16;
  const int qs = e->qs;
  int16_t* qtRecBlk = qt_rec(e, layer < 0 ? 0 : layer) + rel_y(s, relIdx) * qs + rel_x(s, relIdx);
  int16_t* picBlk = g->rec[0] + (cu->y + rel_y(s, relIdx)) * plane_stride(s, 0) + cu->x + rel_x(s, relIdx);
  if (checkFirst && checkFull) checkSplit = 0;

  if (checkFull)
  {
    if (checkTS && e->forceTs >= 0)
    {
      // one half of the transform-skip pair (TEncSearch.cpp:1482-1560): both start from the same contexts and the same
      // prediction, the pair is compared by the caller in the order of the loop below (DCT first, strict <)
      const int modeId = e->forceTs;
      set_range(cu->ts[0], relIdx, numParts, modeId);
      set_range(cu->tridx, relIdx, numParts, trDepth);
      singleDist = code_tu_block(c, cu, relIdx, 0, log2n, log2n, trDepth, numParts, 0);
      singleCbf = (cu->cbf[0][relIdx] >> trDepth) & 1;
      if (modeId == 1 && singleCbf == 0) singleCost = FHC_MAX_COST;
      else singleCost = rd_cost(c, intra_bits_qt(c, cu, relIdx, log2n, trDepth, 1, 0, 1, 0), singleDist);
      bestModeId = modeId;
    }
    else if (checkTS)
    {
      sb_copy(c, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT), e->go);
      FHC_NOUNROLL for (int modeId = 0; modeId < 2; modeId++)
      {
        double costTmp;
        set_range(cu->ts[0], relIdx, numParts, modeId);
        set_range(cu->tridx, relIdx, numParts, trDepth);
        const uint32_t dTmp = code_tu_block(c, cu, relIdx, 0, log2n, log2n, trDepth, numParts, modeId == 0 ? 1 : 2);
        const uint32_t cbfTmp = (cu->cbf[0][relIdx] >> trDepth) & 1;
        if (modeId == 1 && cbfTmp == 0) costTmp = FHC_MAX_COST;
        else costTmp = rd_cost(c, intra_bits_qt(c, cu, relIdx, log2n, trDepth, 1, 0, 1, 0), dTmp);
        if (costTmp < singleCost)
        {
          singleCost = costTmp; singleDist = dTmp; singleCbf = cbfTmp; bestModeId = modeId;
          if (modeId == 0)
          {
            FHC_NOUNROLL for (int i = 0; i < 16; i++) { e->tsCoef[0][i] = qtCoefBlk[i]; e->tsRec[0][i] = qtRecBlk[(i >> 2) * qs + (i & 3)]; }
            sb_copy(c, slot_of(c, fullDepth, CI_TEMP_BEST), e->go);
          }
        }
        if (modeId == 0) sb_copy(c, e->go, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT));
      }
      set_range(cu->ts[0], relIdx, numParts, bestModeId);
      if (bestModeId == 0)
      {
        const int ps = plane_stride(s, 0);
        FHC_NOUNROLL for (int i = 0; i < 16; i++)
        {
          qtCoefBlk[i] = e->tsCoef[0][i];
          qtRecBlk[(i >> 2) * qs + (i & 3)] = e->tsRec[0][i];
          if (e->writePic) picBlk[(i >> 2) * ps + (i & 3)] = e->tsRec[0][i];
        }
        set_range(cu->cbf[0], relIdx, numParts, (int)(singleCbf << trDepth));
        sb_copy(c, e->go, slot_of(c, fullDepth, CI_TEMP_BEST));
      }
    }
    else
    {
      if (checkSplit) sb_copy(c, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT), e->go);
      set_range(cu->ts[0], relIdx, numParts, 0);
      set_range(cu->tridx, relIdx, numParts, trDepth);
      singleDist = code_tu_block(c, cu, relIdx, 0, log2n, log2n, trDepth, numParts, 0);
      if (checkSplit) singleCbf = (cu->cbf[0][relIdx] >> trDepth) & 1;
      singleCost = rd_cost(c, intra_bits_qt(c, cu, relIdx, log2n, trDepth, 1, 0, 1, 0), singleDist);
    }
  }
  if (checkSplit)
  {
    double splitCost = 0.0;
    uint32_t splitDist = 0, splitCbf = 0;
    const int q = numParts >> 2;
    if (checkFull)
    {
      sb_copy(c, slot_of(c, fullDepth, CI_QT_TRAFO_TEST), e->go);
      sb_copy(c, e->go, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT));
    }
    else sb_copy(c, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT), e->go);
    FHC_NOUNROLL for (int i = 0; i < 4; i++)
    {
      recur_luma_qt(c, cu, relIdx + i * q, log2n - 1, trDepth + 1, checkFirst, &splitDist, &splitCost);
      splitCbf |= (cu->cbf[0][relIdx + i * q] >> (trDepth + 1)) & 1;
    }
    if (splitCbf)
      FHC_NOUNROLL for (int i = 0; i < numParts; i++) cu->cbf[0][relIdx + i] |= (uint8_t)(1 << trDepth);
    sb_copy(c, e->go, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT));
    splitCost = rd_cost(c, intra_bits_qt(c, cu, relIdx, log2n, trDepth, 1, 0, 1, 0), splitDist);
    if (splitCost < singleCost)
    {
      *distY += splitDist;
      *rdCost = dadd(*rdCost, splitCost);
      return;
    }
    sb_copy(c, e->go, slot_of(c, fullDepth, CI_QT_TRAFO_TEST));
    set_range(cu->tridx, relIdx, numParts, trDepth);
    set_range(cu->cbf[0], relIdx, numParts, (int)(singleCbf << trDepth));
    set_range(cu->ts[0], relIdx, numParts, bestModeId);
    par_copy2d(c, picBlk, plane_stride(s, 0), qtRecBlk, qs, 1 << log2n, 1 << log2n);
  }
  *distY += singleDist;
  *rdCost = dadd(*rdCost, singleCost);
}

// Second pass of estIntraPredLumaQT (TEncSearch.cpp:2420-2488) for a PU whose size is a legal TU size (8..32): the
// "single TU" branch of xRecurIntraCodingLumaQT(bCheckFirst=false) repeats, instruction for instruction, the first-pass
// evaluation of the same mode (same start contexts CI_CURR_BEST, same neighbours), so its cost is the first-pass best
// cost.  Only the split branch is new work; it wins iff splitCost < that cost.  Returns 1 when the split wins.
FHC_DH int second_pass_split_only(const Ctx& c, CUMeta* cu, int relIdx, int log2n, int trDepth, double singleCost, uint32_t* dist, double* cost)
{
  Engine* e = c.e;
  if (!(log2n > min_tu_log2_in_cu(cu))) return 0;
  const int fullDepth = cu->depth + trDepth;
  const int numParts = cu->nparts >> (2 * trDepth), q = numParts >> 2;
  double splitCost = 0.0;
  uint32_t splitDist = 0, splitCbf = 0;
  sb_copy(c, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT), e->go);
  FHC_NOUNROLL for (int i = 0; i < 4; i++)
  {
    recur_luma_qt(c, cu, relIdx + i * q, log2n - 1, trDepth + 1, 0, &splitDist, &splitCost);
    splitCbf |= (cu->cbf[0][relIdx + i * q] >> (trDepth + 1)) & 1;
  }
  if (splitCbf)
    FHC_NOUNROLL for (int i = 0; i < numParts; i++) cu->cbf[0][relIdx + i] |= (uint8_t)(1 << trDepth);
  sb_copy(c, e->go, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT));
  splitCost = rd_cost(c, intra_bits_qt(c, cu, relIdx, log2n, trDepth, 1, 0, 1, 0), splitDist);
  if (splitCost < singleCost) { *dist = splitDist; *cost = splitCost; return 1; }
  return 0;
}

// xSetIntraResultLumaQT (TEncSearch.cpp:1716): copy winning coefficients / reconstruction of a PU into the CU buffers
FHC_D void set_luma_result(const Ctx& c, CUMeta* cu, int16_t* cuCoef, int16_t* cuReco, int relIdx, int log2n, int trDepth)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  if (cu->tridx[relIdx] == trDepth)
  {
    const int n = 1 << log2n;
    par_copy16(c, cuCoef + relIdx * 16, qt_coef(c.e, 5 - log2n) + relIdx * 16, n * n);
    par_copy2d(c, cuReco + rel_y(s, relIdx) * cu->width + rel_x(s, relIdx), cu->width,
               qt_rec(c.e, 5 - log2n) + rel_y(s, relIdx) * c.e->qs + rel_x(s, relIdx), c.e->qs, n, n);
  }
  else
  {
    const int q = (cu->nparts >> (2 * trDepth)) >> 2;
    FHC_NOUNROLL for (int i = 0; i < 4; i++) set_luma_result(c, cu, cuCoef, cuReco, relIdx + i * q, log2n - 1, trDepth + 1);
  }
}

// The alignment points of one first-pass candidate (recur_luma_qt with bCheckFirst) without its work: pictures whose
// candidate list is shorter than the longest possible one (3 RMD survivors + 2 MPMs for PUs >= 16) pad with these so
// that every picture of a block passes the same number of block_align calls.
FHC_D void ghost_first_pass(const Ctx& c, int log2n)
{
  if (log2n <= 5)
  {
    block_align(c); mid_align(c); mid_align(c);                             // code_tu_block entry, RDOQ, intra_bits_qt
    fine_align(c); fine_align(c); fine_align(c); fine_align(c);             // the stages inside code_tu_block
  }
  else
  {
    FHC_NOUNROLL for (int i = 0; i < 4; i++) ghost_first_pass(c, log2n - 1);
    mid_align(c);                                                            // intra_bits_qt of the split node
  }
}

// estIntraPredLumaQT (TEncSearch.cpp:2177-2557)
FHC_DH void est_intra_luma(const Ctx& c, CUMeta* cu, int16_t* cuCoef, int16_t* cuReco)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  FrameGlobal* g = c.g;
  const int depth = cu->depth;
  const int initTrDepth = cu->part[0] == SIZE_2Nx2N ? 0 : 1;
  const int numPU = 1 << (2 * initTrDepth);
  const int puParts = cu->nparts >> (2 * initTrDepth);
  const int puSize = cu->width >> initTrDepth;
  const int log2pu = ilog2(puSize);
  const int ps = plane_stride(s, 0);
  uint32_t overallDist = 0;

  FHC_NOUNROLL for (int pu = 0; pu < numPU; pu++)
  {
    const int partOffset = pu * puParts;
    int numModesForFullRD = (log2pu <= 3) ? 8 : 3;                    // g_aucIntraModeNumFast_UseMPM (TComRom.cpp:545)
    uint32_t rdModeList[12];
    double candCost[12];
    uint32_t bestPUMode = 0, bestPUDist = 0;
    double bestPUCost = FHC_MAX_COST;
    const int pox = cu->x + rel_x(s, partOffset), poy = cu->y + rel_y(s, partOffset);

    // ---- rough mode decision: 35 predictions + SATD as a COOP section (TEncSearch.cpp:2271-2295)
    block_align(c);
    build_refs(c, cu, partOffset, 0, puSize, 1);
    {
      Cmd& m = e->cmd;
      const int mpg = imin(35, 4096 >> (2 * log2pu));
      FHC_NOUNROLL for (int i = 0; i < 35; i++) s->satd[i] = 0;
      m.a[0] = log2pu; m.a[1] = poy * ps + pox; m.a[2] = mpg; m.a[15] = 2 * ((35 + mpg - 1) / mpg);
      FHC_TIC(rp);
      FHC_COUNT(g_emuCounters.rmdSamples += 35ull << (2 * log2pu));
      par(c, OP_RMD);
      FHC_TOC2(c, rp, PF_RMD_PAR);
    }
    FHC_TIC(rm);
    extra_align(c);
    int preds[3];
    const int numCand = get_mpm(c, cu, partOffset, preds);
    {
      // xModeBitsIntra (TEncSearch.cpp:5353): bin counter state + one context reloaded from CI_CURR_BEST
      const Sbac* srcSb = slot_of(c, depth, CI_CURR_BEST);
      const uint64_t frac0 = srcSb->frac & 32767;
      const uint8_t st = srcSb->c[CTX_IPRED];
      FHC_NOUNROLL for (int i = 0; i < numModesForFullRD; i++) candCost[i] = FHC_MAX_COST;
      FHC_NOUNROLL for (int mode = 0; mode < 35; mode++)
      {
        int idx = -1;
        FHC_NOUNROLL for (int i = 0; i < 3; i++) if (mode == preds[i]) idx = i;
        const uint64_t frac = frac0 + (uint64_t)s->t->entropyBits[st ^ (idx != -1 ? 1 : 0)] + (uint64_t)32768 * (uint64_t)(idx != -1 ? (idx ? 2 : 1) : 5);
        const uint32_t modeBits = (uint32_t)(frac >> 15);
        const uint32_t sad = s->satd[mode] >> (s->fp.bitDepth - 8);
        const double cost = dadd((double)sad, dmul((double)modeBits, s->fp.sp.sqrtLambda));
        int shift = 0;                                                          // xUpdateCandList (TEncSearch.cpp:5385)
        FHC_NOUNROLL while (shift < numModesForFullRD && cost < candCost[numModesForFullRD - 1 - shift]) shift++;
        if (shift != 0)
        {
          FHC_NOUNROLL for (int i = 1; i < shift; i++)
          {
            rdModeList[numModesForFullRD - i] = rdModeList[numModesForFullRD - 1 - i];
            candCost[numModesForFullRD - i] = candCost[numModesForFullRD - 1 - i];
          }
          rdModeList[numModesForFullRD - shift] = (uint32_t)mode;
          candCost[numModesForFullRD - shift] = cost;
        }
      }
    }
    FHC_NOUNROLL for (int j = 0; j < numCand; j++)
    {
      int included = 0;
      FHC_NOUNROLL for (int i = 0; i < numModesForFullRD; i++) included |= (preds[j] == (int)rdModeList[i]);
      if (!included) rdModeList[numModesForFullRD++] = (uint32_t)preds[j];
    }
    FHC_TOC2(c, rm, PF_RMD_MASTER);

    // ---- full RD over the candidates, then the best mode once more with the full RQT
    int firstSerial = 0;
    if (log2pu <= 3 && e->kind == ENG_MASTER)
    {
      // small PUs (4x4 of NxN, 8x8): every candidate is a single TU, independent of the others (same start contexts,
      // neighbours outside the PU) -> one lane engine per candidate, results compared in list order afterwards
      TaskDesc& td = s->task;
      td.resBase = 0;
      td.tsSplit = (log2pu == 2);                     // a 4x4 PU exists only in an NxN CU: every candidate is a DCT/skip pair
      td.kind = TASK_LUMA_PU; td.numTasks = numModesForFullRD << td.tsSplit; td.depth = depth; td.bufIdx = cu_buf_index(s, cu);
      td.partOffset = partOffset; td.log2pu = log2pu; td.initTrDepth = initTrDepth;
      FHC_NOUNROLL for (int i = 0; i < numModesForFullRD; i++) td.mode[i] = (int)rdModeList[i];
      { FHC_TIC(tl); dispatch_tasks(c, 1); FHC_TOC2(c, tl, PF_TASKS_LUMA); }
      {
        // xSetIntraResultLumaQT + the saved trIdx/cbf/tskip arrays of the winner
        const int be = best_engine(c), n = 1 << log2pu;
        const EngineGlobal* wg = &g->eng[be];
        bestPUCost = wg->resCost; bestPUDist = wg->resDist; bestPUMode = rdModeList[wg->resTask >> td.tsSplit];
        par_copy16(c, cuCoef + partOffset * 16, wg->bestCoef[0], n * n);
        par_copy2d(c, cuReco + rel_y(s, partOffset) * cu->width + rel_x(s, partOffset), cu->width, wg->bestReco[0], n, n, n);
        copy_bytes(g->qtTrIdx, wg->bestTrIdx, puParts);
        FHC_NOUNROLL for (int j = 0; j < 3; j++) { copy_bytes(g->qtCbf[j], wg->bestCbf[j], puParts); copy_bytes(g->qtTs[j], wg->bestTs[j], puParts); }
      }
      // second pass: a 4x4 PU cannot split, re-coding the best mode reproduces the same cost (never strictly smaller)
      firstSerial = (log2pu == 2) ? numModesForFullRD + 1 : numModesForFullRD;
    }
    FHC_NOUNROLL for (int i = firstSerial; i < numModesForFullRD + 1; i++)
    {
      const int second = (i == numModesForFullRD);
      const uint32_t orgMode = second ? bestPUMode : rdModeList[i];
      if (second && log2pu >= 4)
        FHC_NOUNROLL for (int k = numModesForFullRD; k < 5; k++) ghost_first_pass(c, log2pu);   // pad the list to 3 + 2 MPMs
      uint32_t puDist = 0;
      double puCost = 0.0;
      set_range(cu->ldir, partOffset, puParts, (int)orgMode);
      sb_copy(c, e->go, slot_of(c, depth, CI_CURR_BEST));
      if (second && log2pu >= 3 && log2pu <= 5)
      {
        FHC_TIC(sp);
        const int won = second_pass_split_only(c, cu, partOffset, log2pu, initTrDepth, bestPUCost, &puDist, &puCost);
        FHC_TOC3(c, sp, PF2_LUMA2ND);
        if (!won) continue;
      }
      else
      {
        FHC_TIC(fp);
        recur_luma_qt(c, cu, partOffset, log2pu, initTrDepth, second ? 0 : 1, &puDist, &puCost);
        FHC_TOC3(c, fp, (second ? PF2_LUMA2ND : PF2_LUMA1ST_BIG));
      }
      if (puCost < bestPUCost)
      {
        bestPUMode = orgMode; bestPUDist = puDist; bestPUCost = puCost;
        set_luma_result(c, cu, cuCoef, cuReco, partOffset, log2pu, initTrDepth);
        copy_bytes(g->qtTrIdx, cu->tridx + partOffset, puParts);
        FHC_NOUNROLL for (int j = 0; j < 3; j++) { copy_bytes(g->qtCbf[j], cu->cbf[j] + partOffset, puParts); copy_bytes(g->qtTs[j], cu->ts[j] + partOffset, puParts); }
      }
    }
    overallDist += bestPUDist;
    copy_bytes(cu->tridx + partOffset, g->qtTrIdx, puParts);
    FHC_NOUNROLL for (int j = 0; j < 3; j++) { copy_bytes(cu->cbf[j] + partOffset, g->qtCbf[j], puParts); copy_bytes(cu->ts[j] + partOffset, g->qtTs[j], puParts); }
    if (pu != numPU - 1)
      par_copy2d(c, g->rec[0] + poy * ps + pox, ps, cuReco + rel_y(s, partOffset) * cu->width + rel_x(s, partOffset), cu->width, puSize, puSize);
    set_range(cu->ldir, partOffset, puParts, (int)bestPUMode);
  }
  if (numPU > 1)
  {
    uint32_t cY = 0, cU = 0, cV = 0;
    FHC_NOUNROLL for (int i = 0; i < 4; i++)
    {
      cY |= (cu->cbf[0][i * puParts] >> 1) & 1; cU |= (cu->cbf[1][i * puParts] >> 1) & 1; cV |= (cu->cbf[2][i * puParts] >> 1) & 1;
    }
    FHC_NOUNROLL for (int i = 0; i < 4 * puParts; i++) { cu->cbf[0][i] |= (uint8_t)cY; cu->cbf[1][i] |= (uint8_t)cU; cu->cbf[2][i] |= (uint8_t)cV; }
  }
  sb_copy(c, e->go, slot_of(c, depth, CI_CURR_BEST));
  cu->dist = overallDist;
}

// ----------------------------------------------------------------------------------------------
// chroma search (TEncSearch.cpp:1915-2120, :2562-2712)
// ----------------------------------------------------------------------------------------------
FHC_D void recur_chroma_qt(const Ctx& c, CUMeta* cu, int relIdx, int log2n, int trDepth, int procChroma, uint32_t* dist)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  FrameGlobal* g = c.g;
  const int numPartsLuma = cu->nparts >> (2 * trDepth);
  if (cu->tridx[relIdx] == trDepth)
  {
    if (!procChroma) return;
    const int fullDepth = cu->depth + trDepth;
    const int stepped = (log2n == 2);
    const int cl = stepped ? 2 : log2n - 1;
    const int ridx = stepped ? (relIdx & ~3) : relIdx;
    const int nparts = stepped ? numPartsLuma * 4 : numPartsLuma;
    int checkTS = stepped;
    if (checkTS)
    {
      int nb = 0;
      FHC_NOUNROLL for (int k = relIdx; k < relIdx + 4; k++) nb += cu->ts[0][k];
      checkTS = nb > 0;
    }
    FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
    {
      if (e->onlyComp && cc != e->onlyComp) continue;
      double singleCost = FHC_MAX_COST, costTmp = 0;
      uint32_t singleDist = 0, singleCbf = 0;
      int bestMode = 0, bestTS = 0, currMode = 0;
      const int total = checkTS ? 2 : 1;
      const int qsc = e->qs >> 1;
      int16_t* qtCoefBlk = qt_coef(e, 5 - log2n) + e->coff[cc] + ridx * 4;
      int16_t* qtRecBlk = qt_rec(e, 5 - log2n) + e->coff[cc] + (rel_y(s, ridx) >> 1) * qsc + (rel_x(s, ridx) >> 1);
      sb_copy(c, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT), e->go);
      FHC_NOUNROLL for (int tsId = 0; tsId < total; tsId++)
      {
        const int isOne = total == 1;
        set_range(cu->ts[cc], ridx, nparts, tsId);
        currMode++;
        const int isLast = currMode == total;
        const uint32_t dTmp = code_tu_block(c, cu, ridx, cc, cl, log2n, trDepth, nparts, isOne ? 0 : (tsId == 0 ? 1 : 2));
        const uint32_t cbfTmp = (cu->cbf[cc][ridx] >> trDepth) & 1;
        if (tsId == 1 && cbfTmp == 0) costTmp = FHC_MAX_COST;
        else if (!isOne)
        {
          reset_bits(e);                                                         // xGetIntraBitsQTChroma (TEncSearch.cpp:1094)
          if (cbfTmp) code_coeff_nxn(c, cu, ridx, qtCoefBlk, cl, cc);
          costTmp = rd_cost(c, num_bits(e), dTmp);
        }
        if (costTmp < singleCost)
        {
          singleCost = costTmp; singleDist = dTmp; bestTS = tsId; bestMode = currMode; singleCbf = cbfTmp;
          if (!isOne && !isLast)
          {
            FHC_NOUNROLL for (int i = 0; i < 16; i++) { e->tsCoef[cc][i] = qtCoefBlk[i]; e->tsRec[cc][i] = qtRecBlk[(i >> 2) * qsc + (i & 3)]; }
            sb_copy(c, slot_of(c, fullDepth, CI_TEMP_BEST), e->go);
          }
        }
        if (!isOne && !isLast) sb_copy(c, e->go, slot_of(c, fullDepth, CI_QT_TRAFO_ROOT));
      }
      if (bestMode < total)
      {
        const int ps = plane_stride(s, cc);
        int16_t* picBlk = g->rec[cc] + ((cu->y + rel_y(s, ridx)) >> 1) * ps + ((cu->x + rel_x(s, ridx)) >> 1);
        FHC_NOUNROLL for (int i = 0; i < 16; i++)
        {
          qtCoefBlk[i] = e->tsCoef[cc][i];
          qtRecBlk[(i >> 2) * qsc + (i & 3)] = e->tsRec[cc][i];
          if (e->writePic) picBlk[(i >> 2) * ps + (i & 3)] = e->tsRec[cc][i];
        }
        set_range(cu->cbf[cc], ridx, nparts, (int)(singleCbf << trDepth));
        sb_copy(c, e->go, slot_of(c, fullDepth, CI_TEMP_BEST));
      }
      set_range(cu->ts[cc], ridx, nparts, bestTS);
      *dist += singleDist;
    }
  }
  else
  {
    uint32_t splitCbf[3] = { 0, 0, 0 };
    const int q = numPartsLuma >> 2;
    FHC_NOUNROLL for (int i = 0; i < 4; i++)
    {
      const int childProc = (log2n - 1 > 2) ? 1 : (i == 0);
      recur_chroma_qt(c, cu, relIdx + i * q, log2n - 1, trDepth + 1, procChroma && childProc, dist);
      FHC_NOUNROLL for (int cc = 1; cc < 3; cc++) splitCbf[cc] |= (cu->cbf[cc][relIdx + i * q] >> (trDepth + 1)) & 1;
    }
    FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
      if (splitCbf[cc])
        FHC_NOUNROLL for (int i = 0; i < numPartsLuma; i++) cu->cbf[cc][relIdx + i] |= (uint8_t)(1 << trDepth);
  }
}
FHC_D void set_chroma_result(const Ctx& c, CUMeta* cu, int16_t* cuCoef, int16_t* cuReco, int relIdx, int log2n, int trDepth, int procChroma)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  if (!procChroma) return;
  if (cu->tridx[relIdx] == trDepth)
  {
    const int stepped = (log2n == 2), cl = stepped ? 2 : log2n - 1, n = 1 << cl, ridx = stepped ? (relIdx & ~3) : relIdx;
    const int cw = cu->width >> 1;
    FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
    {
      par_copy16(c, cuCoef + comp_off(cc) + ridx * 4, qt_coef(c.e, 5 - log2n) + c.e->coff[cc] + ridx * 4, n * n);
      par_copy2d(c, cuReco + comp_off(cc) + (rel_y(s, ridx) >> 1) * cw + (rel_x(s, ridx) >> 1), cw,
                 qt_rec(c.e, 5 - log2n) + c.e->coff[cc] + (rel_y(s, ridx) >> 1) * (c.e->qs >> 1) + (rel_x(s, ridx) >> 1), c.e->qs >> 1, n, n);
    }
  }
  else
  {
    const int q = (cu->nparts >> (2 * trDepth)) >> 2;
    FHC_NOUNROLL for (int i = 0; i < 4; i++) set_chroma_result(c, cu, cuCoef, cuReco, relIdx + i * q, log2n - 1, trDepth + 1, (log2n - 1 > 2) ? 1 : (i == 0));
  }
}
FHC_DH void est_intra_chroma(const Ctx& c, CUMeta* cu, int16_t* cuCoef, int16_t* cuReco)
{
  Shared* s = c.s;
  FrameGlobal* g = c.g;
  Engine* e = c.e; (void)e;
  const int depth = cu->depth, nparts = cu->nparts, log2cu = ilog2(cu->width);
  int modeList[5] = { PLANAR_IDX, VER_IDX, HOR_IDX, DC_IDX, DM_CHROMA_IDX };
  uint32_t bestMode = 0, bestDist = 0;
  double bestCost = FHC_MAX_COST;
  FHC_NOUNROLL for (int i = 0; i < 4; i++) if (modeList[i] == cu->ldir[0]) { modeList[i] = 34; break; }      // getAllowedChromaDir
  if (log2cu == 3 && e->kind == ENG_MASTER)
  {
    // 8x8 CU: one 4x4 block per chroma component and mode, the five modes are independent -> lane engines
    TaskDesc& td = s->task;
    // the 4x4 Cb / Cr blocks sit at the CU origin whatever the luma TU tree: their reference lines are built once
    FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
    {
      build_refs(c, cu, 0, cc, 4, 0);
      FHC_NOUNROLL for (int i = 0; i < 17; i++) s->refC[cc - 1][i] = s->refU[i];
      s->dcC[cc - 1] = e->refDc;
    }
    td.tsSplit = 0; td.resBase = 0; td.depth = depth; td.bufIdx = cu_buf_index(s, cu);
    td.partOffset = 0; td.log2pu = log2cu; td.initTrDepth = 0;
    FHC_NOUNROLL for (int mi = 0; mi < 5; mi++) td.mode[mi] = modeList[mi];
    // transform skip is tried for chroma only below a 4x4 luma TU that chose it (TEncSearch.cpp:1943-1956); without
    // it Cb and Cr of a mode are independent until their bits are counted: 10 TU tasks, then 5 bit-counting tasks
#ifdef FHC_CHROMA_SPLIT
    const int chromaTS = (cu->tridx[0] == 1) && (cu->ts[0][0] + cu->ts[0][1] + cu->ts[0][2] + cu->ts[0][3] > 0);
#else
    const int chromaTS = 1;       // measured: the two-phase dispatch costs as much as it saves; one task per mode
#endif
    FHC_TIC(tc);
    if (chromaTS) { td.kind = TASK_CHROMA_CU; td.numTasks = 5; dispatch_tasks(c, 1); }
    else
    {
      td.kind = TASK_CHROMA_TU; td.numTasks = 10; dispatch_tasks(c, 1);
      td.kind = TASK_CHROMA_BITS; td.numTasks = 5; td.resBase = 10; dispatch_tasks(c, 0);
    }
    FHC_TOC2(c, tc, PF_TASKS_CHROMA);
    const int be = best_engine(c);
    const EngineGlobal* wg = &g->eng[be];
    bestCost = wg->resCost; bestDist = wg->resDist; bestMode = (uint32_t)modeList[wg->resTask];
    FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
    {
      par_copy16(c, cuCoef + comp_off(cc), wg->bestCoef[cc], 16);
      par_copy2d(c, cuReco + comp_off(cc), 4, wg->bestReco[cc], 4, 4, 4);
      copy_bytes(g->qtCbf[cc], wg->bestCbf[cc], nparts); copy_bytes(g->qtTs[cc], wg->bestTs[cc], nparts);
    }
  }
  else
  FHC_NOUNROLL for (int mi = 0; mi < 5; mi++)
  {
    uint32_t dist = 0;
    FHC_TIC(cb);
    block_align(c);
    e->noAlign = 1;                                  // the chroma TU tree follows the luma tree of this picture
    sb_copy(c, e->go, slot_of(c, depth, CI_CURR_BEST));
    set_range(cu->cdir, 0, nparts, modeList[mi]);
    recur_chroma_qt(c, cu, 0, log2cu, 0, 1, &dist);
    sb_copy(c, e->go, slot_of(c, depth, CI_CURR_BEST));
    const double cost = rd_cost(c, intra_bits_qt(c, cu, 0, log2cu, 0, 1, 1, 0, 1), dist);
    e->noAlign = 0;
    if (cost < bestCost)
    {
      bestCost = cost; bestDist = dist; bestMode = (uint32_t)modeList[mi];
      set_chroma_result(c, cu, cuCoef, cuReco, 0, log2cu, 0, 1);
      FHC_NOUNROLL for (int cc = 1; cc < 3; cc++) { copy_bytes(g->qtCbf[cc], cu->cbf[cc], nparts); copy_bytes(g->qtTs[cc], cu->ts[cc], nparts); }
    }
    FHC_TOC3(c, cb, PF2_CHROMA_BIG);
  }
  FHC_NOUNROLL for (int cc = 1; cc < 3; cc++) { copy_bytes(cu->cbf[cc], g->qtCbf[cc], nparts); copy_bytes(cu->ts[cc], g->qtTs[cc], nparts); }
  set_range(cu->cdir, 0, nparts, (int)bestMode);
  cu->dist += bestDist;
  sb_copy(c, e->go, slot_of(c, depth, CI_CURR_BEST));
}

// ----------------------------------------------------------------------------------------------
// final-syntax coding (TEncEntropy.cpp:200-448)
// ----------------------------------------------------------------------------------------------
FHC_D void encode_transform(const Ctx& c, const CUMeta* cu, const int16_t* coef, int relIdx, int log2n, int trDepth, int chromaCodeAll, int procChroma)
{
  const int subdiv = cu->tridx[relIdx] > trDepth;
  int cbf[3];
  FHC_NOUNROLL for (int cc = 0; cc < 3; cc++) cbf[cc] = (cu->cbf[cc][relIdx] >> trDepth) & 1;
  if (cu->part[relIdx] == SIZE_NxN && trDepth == 0) { }
  else if (log2n > 5) { }
  else if (log2n == 2) { }
  else if (log2n == min_tu_log2_in_cu(cu)) { }
  else enc_bin(c.e, CTX_SUBDIV + (5 - log2n), subdiv);
  FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
    if (trDepth == 0 || chromaCodeAll)
      if (trDepth == 0 || ((cu->cbf[cc][relIdx] >> (trDepth - 1)) & 1))
        code_qt_cbf(c, cu, relIdx, cc, trDepth, subdiv == 0, log2n - 1);
  if (subdiv)
  {
    const int q = (cu->nparts >> (2 * trDepth)) >> 2;
    FHC_NOUNROLL for (int i = 0; i < 4; i++)
      encode_transform(c, cu, coef, relIdx + i * q, log2n - 1, trDepth + 1, (log2n - 1) > 2, (log2n - 1 > 2) ? 1 : (i == 3));
  }
  else
  {
    code_qt_cbf(c, cu, relIdx, 0, trDepth, 1, log2n);
    if (cbf[0]) code_coeff_nxn(c, cu, relIdx, coef + relIdx * 16, log2n, 0);
    if (procChroma)
    {
      const int cl = log2n == 2 ? 2 : log2n - 1, ridx = log2n == 2 ? (relIdx & ~3) : relIdx;
      FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
        if (cbf[cc]) code_coeff_nxn(c, cu, ridx, coef + comp_off(cc) + ridx * 4, cl, cc);
    }
  }
}
FHC_D void encode_cu_syntax(const Ctx& c, const CUMeta* cu, const int16_t* coef)
{
  code_part_size(c, cu, 0, cu->depth);
  code_luma_dir(c, cu, 0, 1);
  code_chroma_dir(c, cu, 0);
  encode_transform(c, cu, coef, 0, ilog2(cu->width), 0, 1, 1);
}

// ----------------------------------------------------------------------------------------------
// CU level (TEncCu.cpp)
// ----------------------------------------------------------------------------------------------
FHC_D void cu_init_est(const Ctx& c, CUMeta* cu)                                                 // initEstData (TComDataCU.cpp:525)
{
  const int n = cu->nparts;
  cu->cost = FHC_MAX_COST; cu->dist = 0; cu->bits = 0; cu->bins = 0;
  set_range(cu->dep, 0, n, cu->depth); set_range(cu->tridx, 0, n, 0); set_range(cu->part, 0, n, SIZE_NONE);
  set_range(cu->pred, 0, n, MODE_NONE); set_range(cu->ldir, 0, n, DC_IDX); set_range(cu->cdir, 0, n, 0);
  FHC_NOUNROLL for (int cc = 0; cc < 3; cc++) { set_range(cu->cbf[cc], 0, n, 0); set_range(cu->ts[cc], 0, n, 0); }
  // coefficient buffers need no clearing here: every coded TU (re)writes its block, and blocks of a CU
  // hypothesis are only read through cbf != 0 (coding) or copied as a whole after being written.
  (void)c;
}
FHC_D void cu_position(const Ctx& c, CUMeta* cu, const CUMeta* parent, int partIdx, int depth)    // initSubCU (TComDataCU.cpp:575)
{
  cu->depth = depth; cu->nparts = parent->nparts >> 2; cu->width = 64 >> depth;
  cu->zidx = parent->zidx + cu->nparts * partIdx;
  cu->x = parent->x + cu->width * (partIdx & 1);
  cu->y = parent->y + cu->width * (partIdx >> 1);
  cu_init_est(c, cu);
}
FHC_D void cu_copy_part_from(const Ctx& c, CUMeta* dst, const CUMeta* src, int partIdx, int withCoef)   // copyPartFrom (TComDataCU.cpp:800)
{
  const int n = src->nparts, off = n * partIdx;
  dst->cost = dadd(dst->cost, src->cost); dst->dist += src->dist; dst->bits += src->bits; dst->bins += src->bins;
  copy_bytes(dst->part + off, src->part, n); copy_bytes(dst->pred + off, src->pred, n); copy_bytes(dst->ldir + off, src->ldir, n);
  copy_bytes(dst->cdir + off, src->cdir, n); copy_bytes(dst->tridx + off, src->tridx, n); copy_bytes(dst->dep + off, src->dep, n);
  FHC_NOUNROLL for (int cc = 0; cc < 3; cc++) { copy_bytes(dst->cbf[cc] + off, src->cbf[cc], n); copy_bytes(dst->ts[cc] + off, src->ts[cc], n); }
  if (withCoef)
  {
    int16_t* dc = c.g->cuCoef[dst->depth][cu_buf_index(c.s, dst)];
    const int16_t* sc = c.g->cuCoef[src->depth][cu_buf_index(c.s, src)];
    const int ny = src->width * src->width, nc = ny >> 2;
    par_copy16(c, dc + partIdx * ny, sc, ny);
    par_copy16(c, dc + 4096 + partIdx * nc, sc + 4096, nc);
    par_copy16(c, dc + 5120 + partIdx * nc, sc + 5120, nc);
  }
}
FHC_D void cu_copy_to_pic(const Ctx& c, const CUMeta* cu, int withCoef)                          // copyToPic (TComDataCU.cpp:906)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  fhc_ctu_result* o = &c.g->out[s->ctuAddr];
  o->totalCost = cu->cost; o->totalDist = cu->dist; o->totalBits = cu->bits; o->totalBins = cu->bins;
  Cmd& m = e->cmd;
  m.a[0] = cu->depth; m.a[1] = cu_buf_index(s, cu); m.a[2] = cu->zidx; m.a[3] = cu->nparts; m.a[4] = withCoef ? cu->width : 0;
  par(c, OP_CU_TO_OUT);
}
FHC_D void copy_reco_to_pic(const Ctx& c, const CUMeta* cu)                                      // xCopyYuv2Pic (TEncCu.cpp:1729)
{
  const FrameGlobal* g = c.g;
  const Shared* s = c.s;
  const int16_t* y = g->cuReco[cu->depth][cu_buf_index(c.s, cu)];
  FHC_NOUNROLL for (int cc = 0; cc < 3; cc++)
  {
    const int sh = cc ? 1 : 0, w = cu->width >> sh, px = cu->x >> sh, py = cu->y >> sh;
    const int cw = imin(w, (s->fp.width >> sh) - px), ch = imin(w, (s->fp.height >> sh) - py);
    par_copy2d(c, g->rec[cc] + py * plane_stride(s, cc) + px, plane_stride(s, cc), y + comp_off(cc), w, cw, ch);
  }
}
FHC_D void copy_reco_to_parent(const Ctx& c, const CUMeta* child, const CUMeta* parent, int partIdx)   // xCopyYuv2Tmp (TEncCu.cpp:1742)
{
  const int16_t* sy = c.g->cuReco[child->depth][cu_buf_index(c.s, child)];
  int16_t* dy = c.g->cuReco[parent->depth][cu_buf_index(c.s, parent)];
  FHC_NOUNROLL for (int cc = 0; cc < 3; cc++)
  {
    const int w = child->width >> (cc ? 1 : 0), pw = parent->width >> (cc ? 1 : 0);
    par_copy2d(c, dy + comp_off(cc) + (partIdx >> 1) * w * pw + (partIdx & 1) * w, pw, sy + comp_off(cc), w, w, w);
  }
}
FHC_DH void check_best_mode(const Ctx& c, int depth)                                             // xCheckBestMode (TEncCu.cpp:1659)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  if (temp_cu(s, depth)->cost < best_cu(s, depth)->cost)
  {
    s->bestIdx[depth] = (uint8_t)(1 - s->bestIdx[depth]);
    sb_copy(c, slot_of(c, depth, CI_NEXT_BEST), slot_of(c, depth, CI_TEMP_BEST));
  }
}
FHC_D void check_rd_cost_intra(const Ctx& c, int depth, int partSize)                            // xCheckRDCostIntra (TEncCu.cpp:1528)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  FrameGlobal* g = c.g;
  CUMeta* cu = temp_cu(s, depth);
  const int bi = cu_buf_index(s, cu);
  int16_t* cuCoef = g->cuCoef[depth][bi];
  int16_t* cuReco = g->cuReco[depth][bi];
  FHC_TIC(chk);
  set_range(cu->part, 0, cu->nparts, partSize);
  set_range(cu->pred, 0, cu->nparts, MODE_INTRA);
  { FHC_TIC(el); est_intra_luma(c, cu, cuCoef, cuReco); FHC_TOC2(c, el, PF_LUMA); }
  par_copy2d(c, g->rec[0] + cu->y * plane_stride(s, 0) + cu->x, plane_stride(s, 0), cuReco, cu->width, cu->width, cu->width);
  { FHC_TIC(ec); est_intra_chroma(c, cu, cuCoef, cuReco); FHC_TOC2(c, ec, PF_CHROMA); }
  block_align(c);
  reset_bits(e);
  encode_cu_syntax(c, cu, cuCoef);
  sb_copy(c, slot_of(c, depth, CI_TEMP_BEST), e->go);
  cu->bits = num_bits(e);
  cu->bins = e->bins;
  cu->cost = rd_cost(c, cu->bits, cu->dist);
  check_best_mode(c, depth);
  FHC_TOC3(c, chk, (depth < 3 ? depth : (partSize == SIZE_NxN ? PF2_CHECK_D3_NXN : PF2_CHECK_D3_2N)));
}

FHC_D void compress_cu(const Ctx& c, int depth)                                                  // xCompressCU (TEncCu.cpp:496-1058)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  const FrameGlobal* g = c.g;
  const int x = best_cu(s, depth)->x, y = best_cu(s, depth)->y, w = best_cu(s, depth)->width;
  const int boundary = !(x + w - 1 < s->fp.width && y + w - 1 < s->fp.height);
  extra_align(c);
  if (!boundary)
  {
    cu_init_est(c, temp_cu(s, depth));
    check_rd_cost_intra(c, depth, SIZE_2Nx2N);
    cu_init_est(c, temp_cu(s, depth));
    if (depth == 3)
    {
      check_rd_cost_intra(c, depth, SIZE_NxN);
      cu_init_est(c, temp_cu(s, depth));
    }
    CUMeta* b = best_cu(s, depth);
    sb_copy(c, e->go, slot_of(c, depth, CI_NEXT_BEST));
    reset_bits(e);
    code_split_flag(c, b, 0, depth);
    b->bits += num_bits(e);
    b->bins += e->bins;
    b->cost = rd_cost(c, b->bits, b->dist);
    sb_copy(c, slot_of(c, depth, CI_NEXT_BEST), e->go);
  }
  if (depth < 3)
  {
    const int nd = depth + 1;
    CUMeta* t = temp_cu(s, depth);
    cu_init_est(c, t);
    FHC_NOUNROLL for (int i = 0; i < 4; i++)
    {
      cu_position(c, &s->cu[nd][0], t, i, nd);
      cu_position(c, &s->cu[nd][1], t, i, nd);
      if (s->cu[nd][0].x < s->fp.width && s->cu[nd][0].y < s->fp.height)
      {
        if (i == 0) sb_copy(c, slot_of(c, nd, CI_CURR_BEST), slot_of(c, depth, CI_CURR_BEST));
        else sb_copy(c, slot_of(c, nd, CI_CURR_BEST), slot_of(c, nd, CI_NEXT_BEST));
        compress_cu(c, nd);
        extra_align(c);
        cu_copy_part_from(c, t, best_cu(s, nd), i, 1);
        copy_reco_to_parent(c, best_cu(s, nd), t, i);
      }
      else
      {
        cu_copy_to_pic(c, best_cu(s, nd), 0);
        cu_copy_part_from(c, t, best_cu(s, nd), i, 0);
        {
          // initSubCU leaves zero coefficients for a sub-CU outside the picture (TComDataCU.cpp:651-660)
          int16_t* dc = c.g->cuCoef[depth][cu_buf_index(s, t)];
          const int ny = (w >> 1) * (w >> 1), nc = ny >> 2;
          par_zero16(c, dc + i * ny, ny); par_zero16(c, dc + 4096 + i * nc, nc); par_zero16(c, dc + 5120 + i * nc, nc);
        }
      }
    }
    sb_copy(c, e->go, slot_of(c, nd, CI_NEXT_BEST));
    if (!boundary)
    {
      reset_bits(e);
      code_split_flag(c, t, 0, depth);
      t->bits += num_bits(e);
      t->bins += e->bins;
    }
    t->cost = rd_cost(c, t->bits, t->dist);
    sb_copy(c, slot_of(c, depth, CI_TEMP_BEST), e->go);
    check_best_mode(c, depth);
  }
  extra_align(c);
  cu_copy_to_pic(c, best_cu(s, depth), 1);
  copy_reco_to_pic(c, best_cu(s, depth));
}

// xEncodeCU on the finished CTU (TEncCu.cpp:1124): advances the chain coder RD[0][CI_CURR_BEST]
FHC_D void encode_ctu_rec(const Ctx& c, CUMeta* view, int zidx, int depth)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  const FrameGlobal* g = c.g;
  const fhc_ctu_result* o = &g->out[s->ctuAddr];
  const int r = s->t->z2r[zidx];
  const int x = s->ctuPelX + (r & 15) * 4, y = s->ctuPelY + (r >> 4) * 4, w = 64 >> depth;
  const int nparts = 256 >> (2 * depth);
  int boundary = 0;
  view->zidx = zidx; view->x = x; view->y = y; view->depth = depth; view->nparts = nparts; view->width = w;
  copy_bytes(view->dep, o->depth + zidx, nparts); copy_bytes(view->part, o->partSize + zidx, nparts);
  copy_bytes(view->pred, o->predMode + zidx, nparts); copy_bytes(view->ldir, o->lumaDir + zidx, nparts);
  copy_bytes(view->cdir, o->chromaDir + zidx, nparts); copy_bytes(view->tridx, o->trIdx + zidx, nparts);
  FHC_NOUNROLL for (int cc = 0; cc < 3; cc++) { copy_bytes(view->cbf[cc], o->cbf[cc] + zidx, nparts); copy_bytes(view->ts[cc], o->transformSkip[cc] + zidx, nparts); }
  if (x + w - 1 < s->fp.width && y + w - 1 < s->fp.height) code_split_flag(c, view, 0, depth);
  else boundary = 1;
  if ((depth < o->depth[zidx] && depth < 3) || boundary)
  {
    const int q = nparts >> 2;
    FHC_NOUNROLL for (int i = 0; i < 4; i++)
    {
      const int cz = zidx + i * q, cr = s->t->z2r[cz];
      if (s->ctuPelX + (cr & 15) * 4 < s->fp.width && s->ctuPelY + (cr >> 4) * 4 < s->fp.height) encode_ctu_rec(c, view, cz, depth + 1);
    }
    return;
  }
  // coefficients of this CU: gather the three planes of the record into a CU-relative view (cuCoef[0][temp] is free here)
  int16_t* cv = c.g->cuCoef[0][1 - s->bestIdx[0]];
  par_copy16(c, cv, o->coefY + zidx * 16, w * w);
  par_copy16(c, cv + 4096, o->coefCb + zidx * 4, (w * w) >> 2);
  par_copy16(c, cv + 5120, o->coefCr + zidx * 4, (w * w) >> 2);
  encode_cu_syntax(c, view, cv);
  const int lastInCtu = ((x + w >= s->fp.width) || (((r & 15) * 4 + w) == 64)) && ((y + w >= s->fp.height) || (((r >> 4) * 4 + w) == 64));
  if (lastInCtu && s->ctuAddr != s->fp.wc * s->fp.hc - 1) enc_trm(e, 0);                               // finishCU (TEncCu.cpp:1066)
}

// ----------------------------------------------------------------------------------------------
// engines and candidate tasks
// ----------------------------------------------------------------------------------------------
// executed by the group's leader: binds the picture's engines to their scratch
FHC_D void init_engines(FrameGlobal* g, Shared* s, const Tables* tables)
{
  Engine* m = &s->master;
  s->t = tables;
  m->kind = ENG_MASTER; m->t = tables; m->slots = 0; m->noAlign = 0; m->forceTs = -1; m->refsShared = 0; m->onlyComp = 0;
  m->refU = s->refU; m->refF = s->refF; m->flags = s->flags; m->go = &s->masterGo;
  m->csSmall = &s->u.cs; m->csLarge = &g->big;
  bind_coef_scratch(m, &s->u.cs);
  m->predYuv = g->predYuv; m->qtRecBase = &g->qtRec[0][0]; m->qtCoefBase = &g->qtCoef[0][0]; m->writePic = 1;
  m->layerSize = 6144; m->qs = 64; m->coff[0] = 0; m->coff[1] = 4096; m->coff[2] = 5120;
  m->sigC0 = s->sigC0; m->sigC1 = s->sigC1; m->sigDelta = s->sigDelta;
  m->gsSmall = &s->gs; m->gsLarge = &g->bigGs;
  bind_group_scratch(m, &s->gs);
  {
    uint8_t* st = g->cuStoreBig;
    FHC_NOUNROLL for (int d = 0; d < 4; d++)
    {
      if (d == 2) st = s->cuStore;
      FHC_NOUNROLL for (int b = 0; b < 2; b++) { cu_bind(&s->cu[d][b], st, 256 >> (2 * d)); st += 12 * (256 >> (2 * d)); }
    }
  }
}

// Candidate t of s->task evaluated by ONE lane on lane engine engIdx (an engine runs its tasks in increasing order and
// keeps the best: strict <).  Everything it writes is private (g->eng[engIdx]); it reads the picture planes, the
// picture-level CTU records, the CU under test and the CU's CI_CURR_BEST contexts, none of which change while a
// dispatch runs.
FHC_D void run_task(const Ctx& mc, int engIdx, int t)
{
  FrameGlobal* g = mc.g;
  Shared* s = mc.s;
  EngineGlobal* eg = &g->eng[s->task.resBase + engIdx];
  const int firstOfDispatch = t < FHC_NENG;
  LaneLocal L;
  Engine* e = &L.eng;
  Ctx c;
  c.g = g; c.s = s; c.e = e; c.lane = 0; c.nl = 1; c.mask = 0;
  e->kind = ENG_LANE; e->t = s->t; e->slots = L.slots; e->go = &L.go; e->noAlign = 1; e->writePic = 0;
  e->flags = L.flags; e->csSmall = &L.cs16; e->csLarge = &eg->cs;
  bind_coef_scratch(e, &L.cs16);
  e->predYuv = L.predYuv; e->qtRecBase = &L.qtRec[0][0]; e->qtCoefBase = &L.qtCoef[0][0];
  e->layerSize = 96; e->qs = 8; e->coff[0] = 0; e->coff[1] = 64; e->coff[2] = 80;
  e->costCGSig = L.costCGSig; e->cgFlag = L.cgFlag; e->cgFlagCode = L.cgFlagCode; e->cgNzMask = L.cgNzMask; e->cgAbs = L.cgAbs;
  e->sigC0 = eg->sigC0; e->sigC1 = eg->sigC1; e->sigDelta = eg->sigDelta;
  cu_bind(&L.cu, L.cuStore, 4);
  const TaskDesc& td = s->task;
  const CUMeta* src = &s->cu[td.depth][td.bufIdx];
  CUMeta* cu = &L.cu;
  const int np = src->nparts;
  cu->zidx = src->zidx; cu->x = src->x; cu->y = src->y; cu->depth = src->depth; cu->nparts = np; cu->width = src->width;
  FHC_NOUNROLL for (int i = 0; i < np; i++)
  {
    cu->dep[i] = src->dep[i]; cu->part[i] = src->part[i]; cu->pred[i] = src->pred[i]; cu->ldir[i] = src->ldir[i];
    cu->cdir[i] = src->cdir[i]; cu->tridx[i] = src->tridx[i];
    FHC_NOUNROLL for (int cc = 0; cc < 3; cc++) { cu->cbf[cc][i] = src->cbf[cc][i]; cu->ts[cc][i] = src->ts[cc][i]; }
  }
  const int mode = td.mode[td.kind == TASK_CHROMA_TU ? (t >> 1) : (t >> td.tsSplit)];
  e->forceTs = td.tsSplit ? (t & 1) : -1;
  e->onlyComp = 0;
  if (td.kind == TASK_LUMA_PU)
  {
    // the PU is one TU at the position and size of the leader's rough mode decision: same reference lines
    e->refsShared = 1; e->refU = s->refU; e->refF = s->refF; e->refDc = s->master.refDc;
  }
  else { e->refsShared = 2; e->refU = L.refU; e->refF = L.refF; }
  if (td.kind == TASK_LUMA_PU)
  {
    // one iteration of the candidate loop of estIntraPredLumaQT (TEncSearch.cpp:2349-2418), bCheckFirst = true
    uint32_t dist = 0;
    double cost = 0.0;
    const int puParts = np >> (2 * td.initTrDepth);
    const int layer = 5 - td.log2pu, n = 1 << td.log2pu;
    set_range(cu->ldir, td.partOffset, puParts, mode);
    sb_copy(c, e->go, chain_slot(c, td.depth, CI_CURR_BEST));
    recur_luma_qt(c, cu, td.partOffset, td.log2pu, td.initTrDepth, 1, &dist, &cost);
    if (firstOfDispatch || cost < eg->resCost)
    {
      eg->resCost = cost; eg->resDist = dist; eg->resTask = t;
      par_copy16(c, eg->bestCoef[0], qt_coef(e, layer) + td.partOffset * 16, n * n);
      par_copy2d(c, eg->bestReco[0], n, qt_rec(e, layer) + rel_y(s, td.partOffset) * e->qs + rel_x(s, td.partOffset), e->qs, n, n);
      FHC_NOUNROLL for (int i = 0; i < puParts; i++)
      {
        eg->bestTrIdx[i] = cu->tridx[td.partOffset + i];
        FHC_NOUNROLL for (int cc = 0; cc < 3; cc++) { eg->bestCbf[cc][i] = cu->cbf[cc][td.partOffset + i]; eg->bestTs[cc][i] = cu->ts[cc][td.partOffset + i]; }
      }
    }
  }
  else if (td.kind == TASK_CHROMA_TU)
  {
    // one component of one iteration of the chroma mode loop: the TU coding only (no context is touched when no
    // transform skip is tried, TEncSearch.cpp:2022-2047), left in this task's engine for the bit-counting task
    uint32_t dist = 0;
    const int cc = 1 + (t & 1);
    sb_copy(c, e->go, chain_slot(c, td.depth, CI_CURR_BEST));
    set_range(cu->cdir, 0, np, mode);
    e->onlyComp = cc;
    recur_chroma_qt(c, cu, 0, ilog2(cu->width), 0, 1, &dist);
    const int layer = 5 - (3 - (int)cu->tridx[0]);
    eg->resDist = dist;
    par_copy16(c, eg->bestCoef[cc], qt_coef(e, layer) + e->coff[cc], 16);
    par_copy2d(c, eg->bestReco[cc], 4, qt_rec(e, layer) + e->coff[cc], e->qs >> 1, 4, 4);
    FHC_NOUNROLL for (int i = 0; i < np; i++) { eg->bestCbf[cc][i] = cu->cbf[cc][i]; eg->bestTs[cc][i] = cu->ts[cc][i]; }
  }
  else
  {
    // chroma mode loop of estIntraPredChromaQT (TEncSearch.cpp:2612-2648): TASK_CHROMA_CU = one whole iteration,
    // TASK_CHROMA_BITS = its bit counting and cost from the two TU tasks of the mode (engines 2t, 2t+1)
    uint32_t dist = 0;
    const int log2cu = ilog2(cu->width);
    const int layer = 5 - (3 - (int)cu->tridx[0]);             // luma TU 8x8 (layer 2) or 4x4 (layer 3); chroma is 4x4 at index 0
    sb_copy(c, e->go, chain_slot(c, td.depth, CI_CURR_BEST));
    set_range(cu->cdir, 0, np, mode);
    if (td.kind == TASK_CHROMA_CU) recur_chroma_qt(c, cu, 0, log2cu, 0, 1, &dist);
    else
      FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
      {
        const EngineGlobal* a = &g->eng[2 * t + cc - 1];
        dist += a->resDist;
        par_copy16(c, qt_coef(e, layer) + e->coff[cc], a->bestCoef[cc], 16);
        par_copy2d(c, qt_rec(e, layer) + e->coff[cc], e->qs >> 1, a->bestReco[cc], 4, 4, 4);
        FHC_NOUNROLL for (int i = 0; i < np; i++) { cu->cbf[cc][i] = a->bestCbf[cc][i]; cu->ts[cc][i] = a->bestTs[cc][i]; }
      }
    sb_copy(c, e->go, chain_slot(c, td.depth, CI_CURR_BEST));
    const double cost = rd_cost(c, intra_bits_qt(c, cu, 0, log2cu, 0, 1, 1, 0, 1), dist);
    if (firstOfDispatch || cost < eg->resCost)
    {
      eg->resCost = cost; eg->resDist = dist; eg->resTask = t;
      FHC_NOUNROLL for (int cc = 1; cc < 3; cc++)
      {
        par_copy16(c, eg->bestCoef[cc], qt_coef(e, layer) + e->coff[cc], 16);
        par_copy2d(c, eg->bestReco[cc], 4, qt_rec(e, layer) + e->coff[cc], e->qs >> 1, 4, 4);
        FHC_NOUNROLL for (int i = 0; i < np; i++) { eg->bestCbf[cc][i] = cu->cbf[cc][i]; eg->bestTs[cc][i] = cu->ts[cc][i]; }
      }
    }
  }
}

// LANE section: lane l of the group runs tasks l, l + FHC_NENG, ... of s->task on lane engine l
FHC_D void run_tasks(const Ctx& c)
{
  const int n = c.s->task.numTasks;
  if (c.lane < FHC_NENG)
    FHC_NOUNROLL for (int t = c.lane; t < n; t += FHC_NENG) run_task(c, c.lane, t);
}
// called by the leader: publishes the dispatch, takes part in it, and waits for the group before comparing results
FHC_D void dispatch_tasks(const Ctx& c, int align)
{
  if (align) block_align(c);       // every picture must pass the same number of alignment points (see block_align)
#ifdef FHC_DEVICE_BUILD
  if (c.nl > 1) { c.e->cmd.op = OP_TASKS; __syncwarp(c.mask); }
  run_tasks(c);
  group_sync(c);
#else
  FHC_NOUNROLL for (int t = 0; t < c.s->task.numTasks; t++) run_task(c, t % FHC_NENG, t);
#endif
}

// ----------------------------------------------------------------------------------------------
// picture loop (TEncSlice::compressSlice, TEncSlice.cpp:698-983), executed by the group leader
// ----------------------------------------------------------------------------------------------
FHC_D void load_picture(const Ctx& c, const PicDesc* p)
{
  FrameGlobal* g = c.g;
  Shared* s = c.s;
  FHC_NOUNROLL for (int i = 0; i < 3; i++) { g->org[i] = p->org[i]; g->rec[i] = p->rec[i]; }
  g->out = p->out; g->sp = p->sp;
  g->width = p->width; g->height = p->height; g->bitDepth = p->bitDepth; g->sampleBytes = p->sampleBytes; g->wc = p->wc; g->hc = p->hc;
  FHC_NOUNROLL for (int i = 0; i < 16; i++) { g->prof[i] = 0; g->profN[i] = 0; g->prof2[i] = 0; g->prof2N[i] = 0; }
  s->fp.width = p->width; s->fp.height = p->height; s->fp.bitDepth = p->bitDepth; s->fp.sampleBytes = p->sampleBytes;
  s->fp.wc = p->wc; s->fp.hc = p->hc; s->fp.sp = p->sp;
}

FHC_D void compress_picture(const Ctx& c)
{
  Shared* s = c.s;
  Engine* e = c.e; (void)e;
  FrameGlobal* g = c.g;
  FHC_TIC(tot);
  FHC_NOUNROLL for (int d = 0; d < 4; d++) s->bestIdx[d] = 0;
  sbac_reset(*s->t, chain_slot(c, 0, CI_CURR_BEST), s->fp.sp.sliceQp);
  const int numCtus = s->fp.wc * s->fp.hc;
  FHC_NOUNROLL for (int addr = 0; addr < numCtus; addr++)
  {
    fhc_ctu_result* o = &g->out[addr];
    s->ctuAddr = addr; s->ctuX = addr % s->fp.wc; s->ctuY = addr / s->fp.wc;
    s->ctuPelX = s->ctuX * 64; s->ctuPelY = s->ctuY * 64;
    // initCtu (TComDataCU.cpp:422): the record doubles as the picture-level CTU
    FHC_NOUNROLL for (int d = 0; d < 2; d++)
    {
      CUMeta* cu = &s->cu[0][d];
      cu->zidx = 0; cu->x = s->ctuPelX; cu->y = s->ctuPelY; cu->depth = 0; cu->nparts = 256; cu->width = 64;
      cu_init_est(c, cu);
    }
    s->bestIdx[0] = 0;
    {
      // preset record: depth 0, partSize/predMode "none", lumaDir DC, everything else 0 (coefficients: zero)
      CUMeta* cu = &s->cu[0][0];
      Cmd& m = e->cmd;
      par_zero16(c, o->coefY, 4096 + 1024 + 1024);
      m.a[0] = 0; m.a[1] = 0; m.a[2] = 0; m.a[3] = 256; m.a[4] = 0;
      set_range(cu->dep, 0, 256, 0);
      par(c, OP_CU_TO_OUT);
      cu_init_est(c, cu);
    }
    sb_copy(c, e->go, slot_of(c, 0, CI_CURR_BEST));
    compress_cu(c, 0);
    sb_copy(c, e->go, slot_of(c, 0, CI_CURR_BEST));
    reset_bits(e);
    {
      FHC_TIC(en);
      extra_align(c);
      encode_ctu_rec(c, &s->cu[0][1 - s->bestIdx[0]], 0, 0);
      FHC_TOC2(c, en, PF_ENCODE);
    }
    sb_copy(c, slot_of(c, 0, CI_CURR_BEST), e->go);
    par(c, OP_CTU_RECO_OUT);
  }
  FHC_TOC2(c, tot, PF_TOTAL);
}

}  // namespace fhc
