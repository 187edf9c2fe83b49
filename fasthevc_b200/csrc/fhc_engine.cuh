// fhc_engine.cuh - device code of the B200 HEVC intra mode-decision engine (round 2 design).
//
// ONE THREAD BLOCK PER PICTURE CHAIN.  Inside a picture the CABAC context chain of TEncSlice::compressSlice forces
// raster CTU order; a block walks the CTUs of its picture in that order and spends its warps on the
// parallelism INSIDE a CTU that HM's recursion hides:
//
//   * depth levels run concurrently.  The 2Nx2N check of a CU at depth d and its whole sub-tree at depth d+1 start
//     from the same contexts ([d][CI_CURR_BEST], TEncCu.cpp:806 vs :917-939) and read only neighbours outside the CU,
//     so one ROLE WARP per depth (0, 1, 2, 3-2Nx2N, 3-NxN) evaluates "its" CU while the chain master (the 3-2Nx2N
//     warp) walks the z-order, starts jobs when a region opens and joins them when it closes (named barriers);
//   * inside a role, the full-RD candidates of a PU (TEncSearch.cpp:2349-2418: all start from CI_CURR_BEST) and the
//     five chroma modes are evaluated side by side by LANE GROUPS (G = 32 / candidates lanes each): sample stages
//     (prediction, residual, DCT/DST, pre-quantisation, dequant, inverse, reconstruction, SSE) are striped over the
//     group's lanes, the sequential stages (RDOQ pass, CABAC bit counting) run on the group's first lane - up to 20
//     of them at once per warp;
//   * the rough mode decision predicts one 8x8 (4x4) tile row per lane and finishes the Hadamard with warp shuffles.
//
// State lives in typed shared memory (source CTU, the CTU's best-so-far reconstruction with its neighbour lines,
// context sets, candidate scratch of depth 2/3); only the large scratch of depth 0/1 candidates is in a per-block
// global workspace (L2 resident).  Nothing is staged through global memory per sample.
//
// The same source compiles for the host (FHC_DEVICE_BUILD undefined; tests only): one virtual lane, groups and roles
// executed one after the other in dependency order - a job is run when it is JOINED, i.e. after its sub-tree, which
// also proves that it does not depend on anything the sub-tree wrote.
//
// Reference citations are to /root/reference/source/Lib (HM 16.14).
#pragma once
#include "fhc_rom.h"

#ifdef FHC_DEVICE_BUILD
#define FHC_DH __device__ __forceinline__
#define FHC_D __device__ __noinline__
#define FHC_SHARED(p) __builtin_assume(__isShared(p))      // address-space hint: the pointer is in shared memory (LDS / STS instead of generic accesses)
#define FHC_LANE ((int)(threadIdx.x & 31))
#define FHC_NL 32
#ifndef FHC_NOUNROLL
#define FHC_NOUNROLL _Pragma("unroll 1")
#endif
#else
#define FHC_DH static inline
#define FHC_D static
#define FHC_SHARED(p) ((void)0)
#define FHC_LANE 0
#define FHC_NL 1
#define FHC_NOUNROLL
#endif

namespace fhc {

#define FHC_ROLES 5              /* warps per block: 0 = chain master + depth-3 2Nx2N, 1 = depth-3 NxN, 2 = depth 2, 3 = depth 1, 4 = depth 0 */
#ifndef FHC_PICS_PER_BLOCK
#define FHC_PICS_PER_BLOCK 2     /* picture chains per thread block; 2 = "twins": the same-role warps of the two pictures are kept on the same routine */
#endif
#define FHC_INST_THREADS (32 * FHC_ROLES)
#define FHC_BLOCK_THREADS (FHC_INST_THREADS * FHC_PICS_PER_BLOCK)
#define FHC_MAX_CAND 10          /* 8 RMD survivors + 2 MPMs */
enum { ROLE_3A = 0, ROLE_3B = 1, ROLE_2 = 2, ROLE_1 = 3, ROLE_0 = 4 };

// work counters of the host emulation build (tools/work_count.py); compiled out of the device build
#ifndef FHC_DEVICE_BUILD
struct EmuCounters { unsigned long long tqSamples[2][4], tuCalls[2][4], rmdSamples, bins; };
static EmuCounters g_emuCounters;
#define FHC_COUNT(stmt) do { stmt; } while (0)
#else
#define FHC_COUNT(stmt) do { } while (0)
#endif

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
FHC_DH int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
FHC_DH int iabs(int v) { return v < 0 ? -v : v; }
FHC_DH int imin(int a, int b) { return a < b ? a : b; }
FHC_DH int imax(int a, int b) { return a > b ? a : b; }
#ifdef FHC_DEVICE_BUILD
FHC_DH int msb_index(uint32_t v) { return 31 - __clz((int)v); }
FHC_DH int lsb_index(uint32_t v) { return __ffs((int)v) - 1; }
FHC_DH int popc(uint32_t v) { return __popc(v); }
FHC_DH double dmul(double a, double b) { return __dmul_rn(a, b); }
FHC_DH double dadd(double a, double b) { return __dadd_rn(a, b); }
FHC_DH double dsub(double a, double b) { return __dsub_rn(a, b); }
FHC_DH void warp_sync() { __syncwarp(); }
// data another block (another SM) may have written during this launch: read from L2, never from this SM's L1
FHC_DH int ld_l2(const int16_t* p) { return __ldcg(p); }
FHC_DH int ld_l2(const uint8_t* p) { return __ldcg(p); }
FHC_DH int ld_l2(const int32_t* p) { return __ldcg(p); }
#else
FHC_DH int msb_index(uint32_t v) { return 31 - __builtin_clz(v); }
FHC_DH int lsb_index(uint32_t v) { return __builtin_ctz(v); }
FHC_DH int popc(uint32_t v) { return __builtin_popcount(v); }
FHC_DH double dmul(double a, double b) { return a * b; }
FHC_DH double dadd(double a, double b) { return a + b; }
FHC_DH double dsub(double a, double b) { return a - b; }
FHC_DH void warp_sync() { }
FHC_DH int ld_l2(const int16_t* p) { return *p; }
FHC_DH int ld_l2(const uint8_t* p) { return *p; }
FHC_DH int ld_l2(const int32_t* p) { return *p; }
#endif
#define FHC_LANES(i, n) for (int i = FHC_LANE; i < (n); i += FHC_NL)
// hides a compile-time constant from the optimiser: keeps ONE copy of the large shared functions instead of a clone per
// constant argument pattern (the kernel is bound by instruction delivery, profiles/)
FHC_DH int opq(int v)
{
#ifdef FHC_DEVICE_BUILD
  asm volatile("" : "+r"(v));
#endif
  return v;
}

// A lane group: G consecutive lanes of a warp that evaluate one candidate together (host: one virtual lane).
struct Grp { int gl, G, idx; unsigned mask; };
FHC_DH void gsync(const Grp& g)
{
#ifdef FHC_DEVICE_BUILD
  if (g.G > 1) __syncwarp(g.mask);
#else
  (void)g;
#endif
}
#define FHC_GLANES(g, i, n) for (int i = (g).gl; i < (n); i += (g).G)
FHC_DH Grp whole_warp() { Grp g; g.gl = FHC_LANE; g.G = FHC_NL; g.idx = 0; g.mask = 0xffffffffu; return g; }
FHC_DH void gadd_u32(const Grp& g, uint32_t* p, uint32_t v)
{
#ifdef FHC_DEVICE_BUILD
  if (g.G > 1) { atomicAdd(p, v); return; }
#endif
  (void)g; *p += v;
}
FHC_DH void gor_u32(const Grp& g, uint32_t* p, uint32_t v)
{
#ifdef FHC_DEVICE_BUILD
  if (g.G > 1) { atomicOr(p, v); return; }
#endif
  (void)g; *p |= v;
}

// ----------------------------------------------------------------------------------------------
// data model
// ----------------------------------------------------------------------------------------------
struct Sbac { uint8_t c[NUM_CTX]; uint64_t frac; };       // TEncSbac contexts + TEncBinCABACCounter::m_fracBits (carry survives resetBits)

// Where a picture chain continues: pictures are scheduled CTU row by CTU row (a row may run on any block), so the one
// thing that crosses a row boundary besides the line buffers - the CABAC contexts - is parked here in between.
struct PicState { int32_t nextRow, status, pad[2]; Sbac ctx; };   // status: see the kernel's scheduler
enum { ST_READY = 0, ST_RUNNING = 1, ST_FINISHED = 2 };
// Wavefront pictures (entropy_coding_sync, HM --WaveFrontSynchro=1): the CTU rows of a picture are chains of their own.
// TEncSlice::compressSlice resets the RD coder at the start of every CTU row - contexts AND the fractional-bit carry
// (TEncSbac::resetEntropy -> TEncBinCABAC::start, TEncBinCoderCABAC.cpp:69-79 with FAST_BIT_EST) - and, when the
// picture is at least two CTUs wide, loads the contexts saved after the second CTU of the row above
// (TEncSlice.cpp:808-823, :925-928).  A row therefore depends on the row above only through those contexts and through
// the neighbouring samples / CU depths of the CTUs up to above-right: CTU (r, x) may start when row r-1 has finished
// min(x + 2, wc) CTUs.  progress = CTUs of the row finished (published after the results are visible).
struct RowState { int32_t progress, pad[3]; Sbac sync; Sbac ctx; };   // ctx: where a row continues when a range ends inside it (emulation only)

struct PicDesc                    // global memory, one per submitted picture (filled by the host)
{
  const void* org[3];             // input planes, sampleBytes each
  fhc_ctu_result* out;            // [numCtus]
  int16_t* line[2][3];            // reconstruction of the CTU row above (ping-pong by CTU row), plane width each
  uint8_t* depthRow[2];           // CU depth of the bottom partition row of the CTU row above (ping-pong), width / 4
  fhc_slice_params sp;
  int32_t width, height, bitDepth, sampleBytes, wc, hc;
  int32_t* doneFlag;              // set to 1 (system scope) when the picture is finished
  PicState* state;                // row-wise scheduling state (device builds; the emulation may pass its own)
  int32_t* trace;                 // optional progress word (host-mapped): ctu << 12 | region z << 4 | stage; 0 = off
  long long* prof;                // [64] clock64 counters of the picture's block (FHC_PROFILE builds): [role][stage]
  const uint8_t* hints;           // CU-split hints of the CNN predictor [numCtus][4] (fasthevc_cuda.h: FHC_HINT_*) or null
  RowState* rows;                 // wavefront pictures: [hc] per-row state (zeroed by the host); null = one chain per picture
};

struct PicParams                  // hot copy of the picture header (shared memory)
{
  const void* org[3];
  fhc_ctu_result* out;
  int16_t* line[2][3];
  uint8_t* depthRow[2];
  fhc_slice_params sp;
  int32_t width, height, bitDepth, sampleBytes, wc, hc;
  int32_t* trace;
  long long* prof;
  const uint8_t* hints;
};
// cycle counters per role (FHC_PROFILE builds; lane 0 of the role's warp): stage indices
enum { PF_RMD = 0, PF_LUMA1, PF_LUMA2, PF_CHROMA, PF_SYNTAX, PF_JOB, PF_WAIT, PF_JOBS, PF_NUM = 8 };
#if defined(FHC_DEVICE_BUILD) && defined(FHC_PROFILE)
#define FHC_TIC(v) const long long v = clock64()
#define FHC_TOC(r, stage, v) do { if (FHC_LANE == 0) (r)->prof[stage] += clock64() - (v); } while (0)
#define FHC_CNT(r, stage) do { if (FHC_LANE == 0) (r)->prof[stage]++; } while (0)
#else
#define FHC_TIC(v) do { } while (0)
#define FHC_TOC(r, stage, v) do { } while (0)
#define FHC_CNT(r, stage) do { } while (0)
#endif
FHC_DH void trace_at(const PicParams& pp, int ctu, int z, int stage)
{
  if (pp.trace && FHC_LANE == 0) *(volatile int32_t*)pp.trace = (ctu << 12) | (z << 4) | stage;
}
FHC_DH void trace_role(const PicParams& pp, int ri, int stage)          // words 1..7 of the picture's trace record: one per role
{
  if (pp.trace && FHC_LANE == 0) *(volatile int32_t*)(pp.trace + 1 + ri) = stage;
}

// One role = one warp = one CU hypothesis under evaluation ("job").  Buffers are carved from the shared-memory arena
// (depth 2, 3; the small hot state of depth 0, 1) or from the block's global workspace (sample / coefficient scratch
// of depth 0, 1) by init_block.  Sample planes of a CU: Y (cuw x cuw) at 0, Cb at cuw*cuw, Cr at cuw*cuw*5/4 ("yuv
// layout"); coefficient buffers use HM's per-TU raster blocks at 16 * relIdx (luma) / 4 * relIdx (chroma) in the
// same three-plane layout.
struct Role
{
  // configuration
  int depth, cuw, log2cu, np, yuvSize, maxGroups, tuCap, refCap;   // np = 4x4 partitions of the CU, tuCap = largest TU in coefficients
  // the job (written by the chain master before the start signal)
  int zidx, x, y, partSize;       // z index of the CU's first partition in the CTU, CTU-relative luma position
  // per-group arenas
  Sbac* go;                       // [maxGroups] coder under work
  int32_t *tuA, *tuB;             // [maxGroups][tuCap]
  double* cgCost;                 // [maxGroups][tuCap / 16] costCGSig of RDOQ
  uint16_t* nzMask;               // [maxGroups][yuvSize / 16] non-zero map per coefficient group (coefficient coding)
  uint32_t* gAcc;                 // [maxGroups][4]: sse accumulator, cgNZ[2], broadcast word
  int16_t *gRec, *gCoef;          // [maxGroups][yuvSize] candidate reconstruction / levels
  uint8_t* gMeta;                 // [maxGroups][4][np]: chroma cbf[2], ts[2] of the candidate chroma mode
  double* gCost; uint32_t* gDist; // [maxGroups] results of a dispatch
  uint32_t* gCbf;                 // [maxGroups] luma cbf bits of the candidate's TUs (bit k = TU k) | ts << 8
  int16_t *gRefU, *gRefF;         // [maxGroups][refCap] reference lines built by a group (TUs inside a 64x64 PU, chroma)
  // hypothesis under construction ("cur": second luma pass) and best-so-far of the job
  int16_t *curRec, *curCoef;      // [yuvSize] second-pass working planes (NxN: the CU's own reconstruction that later PUs predict from)
  int16_t *bestRec, *bestCoef;    // [yuvSize]
  int16_t *altRec, *altCoef;      // [(cuw/2)^2] saved "single" result while the split of an RQT node is evaluated
  uint8_t *tridx, *cbf[3], *ts[3];        // [np] best
  uint8_t *cTridx, *cCbf, *cTs;            // [np] second-pass luma working arrays
  Sbac* slots;                    // [2 (+ 2 per chroma group for the NxN role)]: ROOT/TEST of the inner RQT level; transform-skip root/best
  int16_t *refU, *refF;           // [4*cuw+1] reference lines of the PU (shared by its candidates)
  // job results
  uint8_t ldir[4]; uint8_t cdir; uint8_t pad0[3];
  uint32_t dist, bits, bins, pad1;
  double cost;
  Sbac endCtx;                    // CI_TEMP_BEST of the hypothesis
  // RMD scratch
  uint32_t satd[36];
  double rmdCost[36];
  int cand[FHC_MAX_CAND + 2]; int numCand;
  int mpm[3]; int dcVal;
  long long prof[PF_NUM];
  // scratch that holds nothing once a job has produced its result (everything but best* / nzMask / gMeta): whole 128-byte lines
  unsigned char* dead[10];
  uint32_t deadLines[10];
  int numDead;
};

// chain totals of a decided region
struct Totals { uint32_t dist, bits, bins, pad; double cost; };

struct Blk                        // shared memory of a block
{
  Tables t;
  PicParams pp;
  int ctuAddr, ctuX, ctuY, ctuPelX, ctuPelY;
  int16_t src[6144];              // source CTU, yuv layout of a 64x64 CU
  int16_t pic[6144];              // best-so-far reconstruction of the CTU ("picture" as the search sees it)
  int16_t topY[1 + 128], topC[2][1 + 64];     // row above the CTU, x = -1 .. 127 (luma)
  int16_t leftY[64], leftC[2][32];             // column left of the CTU
  // best-so-far CU data per partition (z-scan) of this CTU = the picture-level TComDataCU (TComDataCU.cpp:906-979)
  uint8_t ctuDepth[256], ctuLdir[256], ctuPart[256], ctuPred[256], ctuCdir[256], ctuTridx[256], ctuCbf[3][256], ctuTs[3][256];
  uint16_t ctuNz[384];            // non-zero maps for the final coding of the CTU
  uint8_t leftDepth[16], leftLdir[16], aboveDepth[16];
  Sbac curr[4], next[4], tmp[4];  // [depth] CI_CURR_BEST / CI_NEXT_BEST of the chain, scratch of the split hypothesis
  Totals tot[4];                  // totals of the last decided region per depth
  uint32_t scal[8];               // scalar broadcast words of the chain master
  Role role[FHC_ROLES];
  int mailbox[FHC_ROLES];         // job command per role: 1 = run, 0 = end of CTU
  int jobSeq[FHC_ROLES], doneSeq[FHC_ROLES];   // jobs started by the master / finished by the role (join = the two are equal)
  int inst, twin, active;         // index of this picture chain inside the block; active: it has a picture; twin: so has the other chain
};

FHC_DH int comp_off64(int comp) { return comp == 0 ? 0 : (comp == 1 ? 4096 : 5120); }
FHC_DH int yuv_off(int cuw, int comp) { return comp == 0 ? 0 : (comp == 1 ? cuw * cuw : (cuw * cuw * 5) >> 2); }

// ----------------------------------------------------------------------------------------------
// named-barrier hand-shake between the chain master and a role warp (device); the host build runs the job at the join
// ----------------------------------------------------------------------------------------------
#ifdef FHC_DEVICE_BUILD
// aligned forms: the whole (converged) warp executes the instruction once and counts as 32 threads
FHC_DH void bar_sync(int id) { __syncwarp(); asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
FHC_DH void bar_arrive(int id) { __syncwarp(); asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
#endif
FHC_DH int bar_start(const Blk* b, int role) { return b->inst * (FHC_ROLES - 1) + role; }      // ids 1 .. 8 (roles 1..4 of chain 0 / 1)
FHC_DH int bar_twin(int role) { return 2 * (FHC_ROLES - 1) + 1 + role; }                       // ids 9 .. 13

// TWIN ALIGNMENT.  The kernel is bound by instruction delivery: an SM hands only ~0.4 instructions per cycle in total to
// warps that run different code, but warps that run the same code share their fetches (profiles/r02_icache_probe.md).
// The two picture chains of a block have the same geometry, are started together and advance CTU by CTU in lockstep,
// so their same-role warps ("twins") walk the same static schedule of routines; a named barrier at the entry of each
// routine of that schedule keeps them on the same code at the same time.  Only call sites whose count per CTU does not
// depend on the picture content may align (jobs, PUs, the four second-pass nodes, chroma and syntax entry).
#ifndef FHC_DEVICE_BUILD
static unsigned long long g_emuAlignCalls[8];       // per role: the emulation counts what the device would wait on (tests: the count of a
#endif                                               // picture depends on its geometry only, or twins would deadlock)
FHC_DH void twin_align(const Blk* b, int role)
{
#if defined(FHC_DEVICE_BUILD) && FHC_PICS_PER_BLOCK > 1
  if (b->twin) bar_sync(bar_twin(role));
#elif !defined(FHC_DEVICE_BUILD)
  (void)b; g_emuAlignCalls[role]++;
#else
  (void)b; (void)role;
#endif
}

// ----------------------------------------------------------------------------------------------
// CABAC bit counter (TEncBinCoderCABACCounter.cpp:57-140)
// ----------------------------------------------------------------------------------------------
// The bit counter of a serial coding section: the fractional-bit accumulator lives in a register while the section
// runs (coder_open .. coder_close); context states are stepped in place.  Functions that are real calls take the
// caller's Coder by reference and work on a local copy, so that the accumulator stays in registers inside them too.
struct Coder { Sbac* sb; const Tables* t; uint32_t bins; uint64_t frac; };
FHC_DH Coder coder_open(Sbac* s, const Tables* tab) { Coder e; e.sb = s; e.t = tab; e.bins = 0; e.frac = s->frac; return e; }
FHC_DH void coder_close(const Coder& e) { e.sb->frac = e.frac; }
FHC_DH void enc_bin(Coder& e, int ctx, int bin)
{
  const uint32_t tr = e.t->cabac[e.sb->c[ctx] * 2 + bin];            // fractional bits | next state << 24
  e.bins++;
  FHC_COUNT(g_emuCounters.bins++);
  e.frac += (uint64_t)(tr & 0xffffffu);
  e.sb->c[ctx] = (uint8_t)(tr >> 24);
}
FHC_DH void enc_ep(Coder& e, int n) { e.bins += (uint32_t)n; FHC_COUNT(g_emuCounters.bins += (unsigned)n); e.frac += (uint64_t)32768 * (uint64_t)n; }
FHC_DH void enc_trm(Coder& e, int bin) { e.bins++; FHC_COUNT(g_emuCounters.bins++); e.frac += (uint64_t)e.t->entropyBits[126 ^ bin]; }
FHC_DH void reset_bits(Coder& e) { e.frac &= 32767; e.bins = 0; }                    // TEncBinCoderCABAC.cpp:161-174
FHC_DH uint32_t num_bits(const Coder& e) { return (uint32_t)(e.frac >> 15); }
FHC_DH int ctx_bits(const Tables& t, const Sbac* sb, int ctx, int bin) { return t.entropyBits[sb->c[ctx] ^ bin]; }
// context-set copies: by one lane, or striped over a group
FHC_DH void sb_copy1(Sbac* d, const Sbac* s)
{
  const uint64_t* a = (const uint64_t*)s; uint64_t* b = (uint64_t*)d;
#ifdef FHC_DEVICE_BUILD
#pragma unroll
#endif
  for (int i = 0; i < (int)(sizeof(Sbac) / 8); i++) b[i] = a[i];
}
FHC_DH void sb_copy_g(const Grp& g, Sbac* d, const Sbac* s)
{
  const uint64_t* a = (const uint64_t*)s; uint64_t* b = (uint64_t*)d;
  FHC_GLANES(g, i, (int)(sizeof(Sbac) / 8)) b[i] = a[i];
}
FHC_DH void sbac_reset(const Tables& t, Sbac* sb, int qp)                             // ContextModel::init (ContextModel.cpp:56)
{
  qp = clip3(0, 51, qp);
  FHC_LANES(i, NUM_CTX)
  {
    const int iv = t.ctxInit[i], slope = (iv >> 4) * 5 - 45, offset = ((iv & 15) << 3) - 16;
    const int st = imin(imax(1, ((slope * qp) >> 4) + offset), 126);
    const int mps = st >= 64;
    sb->c[i] = (uint8_t)(((mps ? (st - 64) : (63 - st)) << 1) + mps);
  }
  if (FHC_LANE == 0) sb->frac = 0;
}
FHC_DH double rd_cost(const Blk* b, uint32_t bits, uint32_t dist)                    // TComRdCost::calcRdCost (TComRdCost.cpp:62)
{
  return dadd((double)dist, dmul((double)bits, b->pp.sp.lambda));
}

// ----------------------------------------------------------------------------------------------
// geometry
// ----------------------------------------------------------------------------------------------
FHC_DH int zx(const Tables& t, int z) { return (t.z2r[z] & 15) * 4; }                // luma x / y of a partition inside the CTU
FHC_DH int zy(const Tables& t, int z) { return (t.z2r[z] >> 4) * 4; }
FHC_DH const uint16_t* scan_tab(const Tables& t, int type, int log2n) { return t.scan + t.scanOff[type][log2n - 2]; }
FHC_DH const uint8_t* scan_cg(const Tables& t, int type, int log2n) { return t.scanCG[type] + cg_off(log2n - 2); }
FHC_DH int scan_idx_for(int dir, int log2n, int isChroma)                            // TComDataCU::getCoefScanIdx (TComDataCU.cpp:3109)
{
  if (log2n > (isChroma ? 2 : 3)) return SCAN_DIAG;
  if (iabs(dir - VER_IDX) <= 4) return SCAN_HOR;
  if (iabs(dir - HOR_IDX) <= 4) return SCAN_VER;
  return SCAN_DIAG;
}

// best-so-far reconstruction of the picture as seen from inside the CTU: component sample at CTU-relative (x, y),
// x >= -1, y >= -1 (row above / column left come from the neighbour lines)
FHC_DH int pic_at(const Blk* b, int comp, int x, int y)
{
  const int sh = comp ? 1 : 0, w = 64 >> sh;
  if (y < 0) return comp == 0 ? b->topY[1 + x] : b->topC[comp - 1][1 + x];
  if (x < 0) return comp == 0 ? b->leftY[y] : b->leftC[comp - 1][y];
  return b->pic[comp_off64(comp) + y * w + x];
}
// a CU's private plane for samples inside the CU, the picture view outside
struct RecView { const int16_t* inner; int cx, cy, cw; };      // CU rectangle in component samples, CTU-relative; inner plane stride = cw
FHC_DH int rec_at(const Blk* b, const RecView& v, int comp, int x, int y)
{
  if (v.inner && x >= v.cx && y >= v.cy && x < v.cx + v.cw && y < v.cy + v.cw) return v.inner[(y - v.cy) * v.cw + (x - v.cx)];
  return pic_at(b, comp, x, y);
}

// ----------------------------------------------------------------------------------------------
// intra reference samples (TComPattern.cpp:115-566).  line: 4n+1 samples, index 0 = bottom-most below-left,
// 2n = corner, 4n = right-most above-right.  (x0, y0): block origin in component samples, CTU-relative.
// Availability follows HM's z-order rules (getPUAboveRight / getPUBelowLeft, TComDataCU.cpp:1090-1200); per side it
// is a prefix of units, so it is five numbers instead of a flag per unit.
// ----------------------------------------------------------------------------------------------
struct RefAvail { int left, corner, above, nBL, nAR; };          // nBL / nAR: available below-left / above-right SAMPLES
FHC_DH RefAvail ref_avail(const Blk* b, int comp, int n, int x0, int y0)
{
  const Tables& t = b->t;
  const int sh = comp ? 1 : 0, us = comp ? 2 : 4, nu = n / us;
  const int ux = (x0 << sh) >> 2, uy = (y0 << sh) >> 2;            // position in 4x4 luma units inside the CTU
  const int rLT = uy * 16 + ux;
  RefAvail a;
  a.left = (ux > 0 || b->ctuX > 0);
  a.above = (uy > 0 || b->ctuY > 0);
  a.corner = a.left && a.above;
  {
    const int rRT = rLT + nu - 1, zRT = t.r2z[rRT];
    int k = 0;
    FHC_NOUNROLL for (int i = 1; i <= nu; i++)
    {
      int av;
      if (b->ctuPelX + (rRT & 15) * 4 + 4 * i >= b->pp.width) av = 0;
      else if ((rRT & 15) < 16 - i) av = (uy != 0) ? (zRT > t.r2z[rRT - 16 + i]) : (b->ctuY > 0);
      else if (uy != 0) av = 0;
      else av = (b->ctuY > 0) && (b->ctuX + 1 < b->pp.wc);
      if (!av) break;
      k = i;
    }
    a.nAR = k * us;
  }
  {
    const int rLB = rLT + (nu - 1) * 16, zLB = t.r2z[rLB];
    int k = 0;
    FHC_NOUNROLL for (int i = 1; i <= nu; i++)
    {
      int av;
      if (b->ctuPelY + (rLB >> 4) * 4 + 4 * i >= b->pp.height) av = 0;
      else if ((rLB >> 4) < 16 - i) av = (ux != 0) ? (zLB > t.r2z[rLB + 16 * i - 1]) : (b->ctuX > 0);
      else av = 0;
      if (!av) break;
      k = i;
    }
    a.nBL = k * us;
  }
  return a;
}

// gathers the 4n+1 samples with HM's substitution process (fillReferenceSamples, TComPattern.cpp:322-539), striped over
// the group's lanes; optionally the smoothed copy (TComPattern.cpp:191-296).  Returns the DC value of the unfiltered line.
FHC_D int build_refs(const Blk* b, const Grp& g, const RecView& v, int comp, int n, int x0, int y0, int wantFilt, int16_t* refU, int16_t* refF)
{
  FHC_SHARED(b);
  const RefAvail a = ref_avail(b, comp, n, x0, y0);
  const int total = 4 * n + 1;
  const int dc = 1 << (b->pp.bitDepth - 1);
  // available index ranges: [n - nBL, n) below-left, [n, 2n) left, {2n} corner, [2n+1, 3n+1) above, [3n+1, 3n+1+nAR) above-right
  const int any = a.left | a.above | a.corner | (a.nBL > 0) | (a.nAR > 0);
  FHC_GLANES(g, k, total)
  {
    int val;
    if (!any) val = dc;
    else
    {
      int src = k;
      // is k available?  otherwise the nearest available index below it, else the first available index above it
      int avail;
      if (k < n) avail = k >= n - a.nBL;
      else if (k < 2 * n) avail = a.left;
      else if (k == 2 * n) avail = a.corner;
      else if (k < 3 * n + 1) avail = a.above;
      else avail = k < 3 * n + 1 + a.nAR;
      if (!avail)
      {
        int lo = -1;
        // highest available index < k
        if (k > 3 * n + 1 && a.nAR > 0) lo = 3 * n + a.nAR;
        else if (k > 2 * n + 1 && a.above) lo = imin(k - 1, 3 * n);
        else if (k > 2 * n && a.corner) lo = 2 * n;
        else if (k > n && a.left) lo = imin(k - 1, 2 * n - 1);
        else if (k > n - a.nBL && a.nBL > 0) lo = imin(k - 1, n - 1);
        if (lo >= 0) src = lo;
        else
        {
          // lowest available index > k
          if (a.nBL > 0 && k < n - a.nBL) src = n - a.nBL;
          else if (a.left && k < n) src = n;
          else if (a.corner && k < 2 * n) src = 2 * n;
          else if (a.above && k < 2 * n + 1) src = 2 * n + 1;
          else src = 3 * n + 1;                      // above-right (nAR > 0 since something is available)
        }
      }
      if (src < 2 * n) val = rec_at(b, v, comp, x0 - 1, y0 + (2 * n - 1 - src));
      else if (src == 2 * n) val = rec_at(b, v, comp, x0 - 1, y0 - 1);
      else val = rec_at(b, v, comp, x0 + (src - 2 * n - 1), y0 - 1);
    }
    refU[k] = (int16_t)val;
  }
  gsync(g);
  int sum = 0;
  FHC_NOUNROLL for (int i = 0; i < n; i++) sum += refU[2 * n + 1 + i] + refU[2 * n - 1 - i];
  int lg = 0;
  while ((1 << lg) < 2 * n) lg++;
  const int dcVal = (sum + n) >> lg;
  if (wantFilt)
  {
    const int bl = refU[0], tl = refU[2 * n], tr = refU[4 * n];
    int strong = (comp == 0) && n >= 32;
    if (strong)
    {
      const int thr = 1 << (b->pp.bitDepth - 5);
      strong = (iabs((bl + tl) - 2 * refU[n]) < thr) && (iabs((tl + tr) - 2 * refU[3 * n]) < thr);
    }
    FHC_GLANES(g, i, total)
    {
      int val;
      if (i == 0 || i == 4 * n) val = refU[i];
      else if (strong)
      {
        if (i < 2 * n) val = (((2 * n - i) * bl) + (i * tl) + n) >> lg;
        else if (i == 2 * n) val = refU[i];
        else val = (((2 * n - (i - 2 * n)) * tl) + ((i - 2 * n) * tr) + n) >> lg;
      }
      else val = (refU[i - 1] + 2 * refU[i] + refU[i + 1] + 2) >> 2;
      refF[i] = (int16_t)val;
    }
    gsync(g);
  }
  return dcVal;
}

// TComPrediction::filteringIntraReferenceSamples (TComPattern.cpp:541)
FHC_DH int use_filtered_refs(const Tables& t, int comp, int mode, int log2n)
{
  if (comp != 0 || mode == DC_IDX) return 0;
  const int diff = imin(iabs(mode - HOR_IDX), iabs(mode - VER_IDX));
  return diff > t.intraFilterThr[log2n - 2];
}

// intra prediction of one sample (TComPrediction.cpp:229-473, :731-817), closed forms of HM's loops
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
    const int a = (k0 >= 0) ? line[c + s * k0] : line[c - s * ((128 + (-k0) * invAngle) >> 8)];
    if (!fract) return a;
    const int k1 = k0 + 1;
    const int bb = (k1 >= 0) ? line[c + s * k1] : line[c - s * ((128 + (-k1) * invAngle) >> 8)];
    return ((32 - fract) * a + fract * bb + 16) >> 5;
  }
}

// The same predictor with everything that depends only on (mode, block size) hoisted out of the sample loop.
struct PredMode { int kind, s, angle, invAngle, edge, shift; };    // kind: 0 planar, 1 DC, 2 pure horizontal / vertical, 3 angular
FHC_DH PredMode pred_mode(const Tables& t, int mode, int n, int log2n, int isLuma)
{
  PredMode p;
  p.edge = isLuma && n <= 16; p.shift = log2n + 1; p.s = 1; p.angle = 0; p.invAngle = 0;
  if (mode == PLANAR_IDX) { p.kind = 0; return p; }
  if (mode == DC_IDX) { p.kind = 1; return p; }
  const int isVer = mode >= 18;
  const int angMode = isVer ? mode - VER_IDX : -(mode - HOR_IDX);
  const int absA = iabs(angMode);
  p.angle = angMode < 0 ? -t.angTable[absA] : t.angTable[absA];
  p.invAngle = t.invAngTable[absA];
  p.s = isVer ? 1 : -1;
  p.kind = p.angle == 0 ? 2 : 3;
  return p;
}
FHC_DH int pred_at(const PredMode& p, const int16_t* line, int n, int x, int y, int bd, int dcval)
{
  const int c = 2 * n;
  if (p.kind == 0)
  {
    const int left = line[c - 1 - y], top = line[c + 1 + x], topRight = line[c + 1 + n], bottomLeft = line[c - 1 - n];
    return ((n - 1 - x) * left + (x + 1) * topRight + (n - 1 - y) * top + (y + 1) * bottomLeft + n) >> p.shift;
  }
  if (p.kind == 1)
  {
    if (p.edge)
    {
      if (x == 0 && y == 0) return (line[c + 1] + line[c - 1] + 2 * dcval + 2) >> 2;
      if (y == 0) return (line[c + 1 + x] + 3 * dcval + 2) >> 2;
      if (x == 0) return (line[c - 1 - y] + 3 * dcval + 2) >> 2;
    }
    return dcval;
  }
  const int s = p.s;                                  // main(k) = line[c + s*k], side(j) = line[c - s*j]
  const int u = s > 0 ? x : y, v = s > 0 ? y : x;
  if (p.kind == 2)
  {
    int val = line[c + s * (u + 1)];
    if (p.edge && u == 0) val = clip3(0, (1 << bd) - 1, val + ((line[c - s * (v + 1)] - line[c]) >> 1));
    return val;
  }
  const int deltaPos = (v + 1) * p.angle, deltaInt = deltaPos >> 5, fract = deltaPos & 31;
  const int k0 = u + deltaInt + 1;
  const int a = (k0 >= 0) ? line[c + s * k0] : line[c - s * ((128 + (-k0) * p.invAngle) >> 8)];
  if (!fract) return a;
  const int k1 = k0 + 1;
  const int bb = (k1 >= 0) ? line[c + s * k1] : line[c - s * ((128 + (-k1) * p.invAngle) >> 8)];
  return ((32 - fract) * a + fract * bb + 16) >> 5;
}

// ----------------------------------------------------------------------------------------------
// quantisation parameters (TComTrQuant.cpp:71-119)
// ----------------------------------------------------------------------------------------------
struct QpParam { int per, rem; };
FHC_DH QpParam qp_param(const Blk* b, int comp)
{
  const int bdOff = 6 * (b->pp.bitDepth - 8);
  int base;
  if (comp == 0) base = b->pp.sp.sliceQp + bdOff;
  else
  {
    base = clip3(-bdOff, 57, b->pp.sp.sliceQp);
    base = base < 0 ? base + bdOff : b->t.chromaScale[base] + bdOff;
  }
  QpParam q; q.per = base / 6; q.rem = base % 6;
  return q;
}
FHC_DH int quant_scale(int rem) { return rem == 0 ? 26214 : rem == 1 ? 23302 : rem == 2 ? 20560 : rem == 3 ? 18396 : rem == 4 ? 16384 : 14564; }
FHC_DH int inv_quant_scale(int rem) { return rem == 0 ? 40 : rem == 1 ? 45 : rem == 2 ? 51 : rem == 3 ? 57 : rem == 4 ? 64 : 72; }

// ----------------------------------------------------------------------------------------------
// coefficient-level context derivation (TComTrQuant.cpp:2672-2760, TComChromaFormat.cpp:100-160)
// ----------------------------------------------------------------------------------------------
FHC_DH int first_sig_ctx(int log2n, int scanType, int isChroma)
{
  if (log2n == 2) return 0;
  if (log2n == 3) return 9 + ((scanType != SCAN_DIAG && !isChroma) ? 6 : 0);
  return isChroma ? 12 : 21;
}
// coefficient-group flags as a 64-bit map indexed by the group's raster position
FHC_DH int cg_right_below(uint64_t cgFlags, int cgX, int cgY, int log2wg, int* right, int* below)
{
  const int wg = 1 << log2wg;
  *right = (cgX < wg - 1) ? (int)((cgFlags >> ((cgY << log2wg) + cgX + 1)) & 1) : 0;
  *below = (cgY < wg - 1) ? (int)((cgFlags >> (((cgY + 1) << log2wg) + cgX)) & 1) : 0;
  return 0;
}
FHC_DH int pattern_sig_ctx(uint64_t cgFlags, int cgX, int cgY, int log2wg)
{
  int r, l;
  if (log2wg == 0) return 0;
  cg_right_below(cgFlags, cgX, cgY, log2wg, &r, &l);
  return r + (l << 1);
}
FHC_DH int sig_cg_ctx(uint64_t cgFlags, int cgX, int cgY, int log2wg)
{
  int r, l;
  cg_right_below(cgFlags, cgX, cgY, log2wg, &r, &l);
  return (r + l) != 0;
}
FHC_DH void last_ctx_params(int isChroma, int log2n, int* off, int* shift)
{
  const int cc = log2n - 2;
  *off = isChroma ? 0 : (cc * 3 + ((cc + 1) >> 2));
  *shift = isChroma ? cc : ((cc + 3) >> 2);
}
FHC_DH int group_idx(int p) { return p < 4 ? p : (p < 6 ? 4 : (p < 8 ? 5 : (p < 12 ? 6 : (p < 16 ? 7 : (p < 24 ? 8 : 9))))); }

// ----------------------------------------------------------------------------------------------
// TEncSbac::codeCoeffNxN (TEncSbac.cpp:1181-1538) incl. transform_skip_flag and last position; coef: int16 raster
// levels of the TU.  Group-level: the non-zero map of every coefficient group is built by all lanes, the bins are
// counted by the group's first lane.
// ----------------------------------------------------------------------------------------------
// nzMask: one 16-bit non-zero map per coefficient group, indexed by the group's RASTER position inside the TU, bit i =
// i-th coefficient of the group in scan order
FHC_DH void nz_masks_tu(const Grp& g, const Tables& t, const int16_t* coef, int log2n, int scanType, uint16_t* nzMask)
{
  const uint16_t* s4 = scan_tab(t, scanType, 2);
  const int n = 1 << log2n, log2wg = log2n - 2, ncg = 1 << (2 * log2wg);
  FHC_GLANES(g, cg, ncg)
  {
    const int16_t* c = coef + ((cg >> log2wg) * 4) * n + (cg & ((1 << log2wg) - 1)) * 4;
    uint32_t mask = 0;
    for (int i = 0; i < 16; i++) { const int v = s4[i]; mask |= (uint32_t)(c[(v >> 2) * n + (v & 3)] != 0) << i; }
    nzMask[cg] = (uint16_t)mask;
  }
}
FHC_DH void code_remain(Coder& e, uint32_t symbol, uint32_t rParam)                            // TEncSbac.cpp:337
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
// serial part; nzMask[] must describe coef (nz_masks_tu / nz_masks_range + group sync before the call)
FHC_D void code_coeff_serial(Coder& eio, const int16_t* coef, int log2n, int isChroma, int scanType, int tsFlag, const uint16_t* nzMask)
{
  Coder e = eio;                                                      // registers inside this call
  FHC_SHARED(e.sb); FHC_SHARED(e.t);
  const Tables& t = *e.t;
  const int n = 1 << log2n, log2wg = log2n - 2, wg = n >> 2;
  const uint16_t* scan = scan_tab(t, scanType, log2n);
  const uint8_t* scanCG = scan_cg(t, scanType, log2n);
  const int firstSig = first_sig_ctx(log2n, scanType, isChroma);
  const int sigBase = CTX_SIG + (isChroma ? 28 : 0);
  uint64_t cgFlags = 0;
  int scanPosLast = -1;
  if (log2n == 2) enc_bin(e, CTX_TS + isChroma, tsFlag);
  FHC_NOUNROLL for (int g4 = wg * wg - 1; g4 >= 0; g4--)
  {
    const uint32_t mask = nzMask[scanCG[g4]];
    if (!mask) continue;
    if (scanPosLast < 0) scanPosLast = g4 * 16 + msb_index(mask);
    cgFlags |= (uint64_t)1 << scanCG[g4];
  }
  const int posLast = scan[scanPosLast];
  {
    int posY = posLast >> log2n, posX = posLast - (posY << log2n), off, shift, cc;
    if (scanType == SCAN_VER) { const int tt = posX; posX = posY; posY = tt; }
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
  const int lastScanSet = scanPosLast >> 4;
  uint32_t c1 = 1;
  FHC_NOUNROLL for (int subSet = lastScanSet; subSet >= 0; subSet--)
  {
    const int subPos = subSet << 4;
    const int cgBlk = scanCG[subSet], cgY = cgBlk >> log2wg, cgX = cgBlk & (wg - 1);
    const uint32_t cgMask = nzMask[cgBlk];
    // coded_sub_block_flag: inferred for the group of the last position and the DC group
    if (subSet == lastScanSet || subSet == 0) cgFlags |= (uint64_t)1 << cgBlk;
    else enc_bin(e, CTX_SIGCG + (isChroma ? 2 : 0) + sig_cg_ctx(cgFlags, cgX, cgY, log2wg), cgMask != 0);
    if (!((cgFlags >> cgBlk) & 1)) continue;
    {
      // sig_coeff_flag of the positions below the last one; the flag at position 0 of a group is inferred when it is
      // the group's only non-zero candidate (no non-zero above it) - except in the DC group
      const int pat = pattern_sig_ctx(cgFlags, cgX, cgY, log2wg);
      const uint8_t* sigRow = t.sigScan[scanType][log2n == 2 ? 4 : pat];
      const int sigAdd = sigBase + firstSig + ((log2n != 2 && cgBlk != 0 && !isChroma) ? 3 : 0);
      int p = (subSet == lastScanSet) ? (scanPosLast - subPos - 1) : 15;
      uint32_t above = (subSet == lastScanSet) ? 1u : 0u;       // a non-zero coefficient was seen in this group
      FHC_NOUNROLL for (; p >= 0; p--)
      {
        const int sig = (int)((cgMask >> p) & 1u);
        if (p > 0 || subSet == 0 || above)
          enc_bin(e, (subPos + p) ? sigAdd + sigRow[p] : sigBase, sig);
        above |= (uint32_t)sig;
      }
    }
    if (cgMask)
    {
      const int lastNZ = msb_index(cgMask), firstNZ = lsb_index(cgMask);
      const int numNonZero = popc(cgMask);
      const int signHidden = (lastNZ - firstNZ) >= 4;
      const int ctxSet = (isChroma ? 4 : 0) + ((!isChroma && subSet > 0) ? 2 : 0) + (c1 == 0 ? 1 : 0);
      int firstC2 = -1, escape = 0, idx = 0;
      uint32_t m = cgMask;
      c1 = 1;
      // coeff_abs_level_greater1_flag of the first 8 non-zero coefficients in reverse scan order
      FHC_NOUNROLL for (; m && idx < 8; idx++)
      {
        const int p = msb_index(m);
        m &= ~(1u << p);
        const int sym = iabs(coef[scan[subPos + p]]) > 1;
        enc_bin(e, CTX_ONE + 4 * ctxSet + (int)c1, sym);
        if (sym) { c1 = 0; if (firstC2 == -1) firstC2 = p; else escape = 1; }
        else if (c1 < 3 && c1 > 0) c1++;
      }
      if (c1 == 0 && firstC2 != -1)
      {
        const int sym = iabs(coef[scan[subPos + firstC2]]) > 2;
        enc_bin(e, CTX_ABS + ctxSet, sym);
        if (sym) escape = 1;
      }
      escape = escape || (numNonZero > 8);
      enc_ep(e, signHidden ? numNonZero - 1 : numNonZero);
      if (escape)
      {
        int firstCoeff2 = 1;
        uint32_t goRice = 0;
        m = cgMask; idx = 0;
        FHC_NOUNROLL for (; m; idx++)
        {
          const int p = msb_index(m);
          m &= ~(1u << p);
          const int a = iabs(coef[scan[subPos + p]]);
          const int baseLevel = (idx < 8) ? (2 + firstCoeff2) : 1;
          if (a >= baseLevel)
          {
            code_remain(e, (uint32_t)(a - baseLevel), goRice);
            if (a > (3 << goRice)) goRice = goRice + 1 < 4 ? goRice + 1 : 4;
          }
          if (a >= 2) firstCoeff2 = 0;
        }
      }
    }
  }
  eio = e;
}

// ----------------------------------------------------------------------------------------------
// RDOQ (TComTrQuant::xRateDistOptQuant, TComTrQuant.cpp:2119-2661), the sequential pass of one TU on one lane.
//
// In:  A[p]   signed pre-quantised magnitude lLevelDouble of scan position p (sign of the coefficient), from the
//             pre-quantisation stage; cgNZ = per coefficient group (scan order): some candidate level is non-zero.
// Out: coef[] int16 levels (raster), B[p] = state of the level coder when position p was quantised (packed),
//      returns uiAbsSum.
// HM keeps seven per-coefficient arrays (costs, rate increments, deltaU, sigRateDelta); here only the packed coder
// state is kept: the costs the last-position search needs and the rate increments sign bit hiding needs are
// re-evaluated from it with the same double operations in the same order (contexts do not move while a TU is being
// quantised), so the results are bit-identical.
// ----------------------------------------------------------------------------------------------
struct RdoqIn { int log2n, isChroma, scanType, cbfCtx, qBits, per, rem, bd; double lambda, errScale; };

FHC_DH int ic_rate(const Tables& t, const Sbac* go, uint32_t absLevel, int ctxOne, int ctxAbs, uint32_t goRice, int c1Lt8, int c2Lt1)   // xGetICRate :2881
{
  int rate = 32768;
  const uint32_t baseLevel = c1Lt8 ? (2u + (uint32_t)c2Lt1) : 1u;
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
    if (c1Lt8)
    {
      rate += ctx_bits(t, go, ctxOne, 1);
      if (c2Lt1) rate += ctx_bits(t, go, ctxAbs, 1);
    }
  }
  else if (absLevel == 1) rate += ctx_bits(t, go, ctxOne, 0);
  else if (absLevel == 2) { rate += ctx_bits(t, go, ctxOne, 1); rate += ctx_bits(t, go, ctxAbs, 0); }
  else rate = 0;
  return rate;
}
// packed coder state: ctxOne - CTX_ONE (5 bits) | ctxAbs - CTX_ABS (3) << 5 | goRice (3) << 8 | c1Idx < 8 << 11 | c2Idx < 1 << 12 |
// sig ctx (6) << 13 | "no sig_coeff_flag" (last position) << 19
FHC_DH int pk_one(int pk) { return CTX_ONE + (pk & 31); }
FHC_DH int pk_abs(int pk) { return CTX_ABS + ((pk >> 5) & 7); }
FHC_DH uint32_t pk_rice(int pk) { return (uint32_t)(pk >> 8) & 7u; }
FHC_DH int pk_c1lt8(int pk) { return (pk >> 11) & 1; }
FHC_DH int pk_c2lt1(int pk) { return (pk >> 12) & 1; }
FHC_DH int pk_sig(int pk) { return (pk >> 13) & 63; }
FHC_DH int pk_nosig(int pk) { return (pk >> 19) & 1; }

struct RdoqState { double blockUncoded, baseCost; int cgLastScanPos, lastScanPos; uint64_t cgFlags; };
FHC_D int ic_rate_call(const Tables& t, const Sbac* go, uint32_t absLevel, int ctxOne, int ctxAbs, uint32_t goRice, int c1Lt8, int c2Lt1)
{
  FHC_SHARED(go); FHC_SHARED(&t);
  return ic_rate(t, go, absLevel, ctxOne, ctxAbs, goRice, c1Lt8, c2Lt1);
}

// best last position (TComTrQuant.cpp:2465-2524): walks down from the last candidate until a level above 1 is met.  The
// per-coefficient costs of the main pass are re-evaluated from the packed coder state (same operations, same order).
FHC_D int rdoq_last_pos(const Blk* b, const Sbac* go, const RdoqIn& q, const int32_t* A, const int32_t* B, const int16_t* coef,
                        const double* costCGSig, const RdoqState& st)
{
  FHC_SHARED(b); FHC_SHARED(go);
  const Tables& t = b->t;
  const int log2n = q.log2n, isChroma = q.isChroma, n = 1 << log2n, qBits = q.qBits;
  const double errScale = q.errScale, lambda = q.lambda;
  const uint16_t* scan = scan_tab(t, q.scanType, log2n);
  const uint8_t* scanCG = scan_cg(t, q.scanType, log2n);
  const int sigBase = CTX_SIG + (isChroma ? 28 : 0);
  const int cgLastScanPos = st.cgLastScanPos, lastScanPos = st.lastScanPos;
  const uint64_t cgFlags = st.cgFlags;
  const double blockUncoded = st.blockUncoded;
  double baseCost = st.baseCost;
  int bestLastIdxP1 = 0;
    double bestCost = dadd(blockUncoded, dmul(lambda, (double)ctx_bits(t, go, q.cbfCtx, 0)));
    baseCost = dadd(baseCost, dmul(lambda, (double)ctx_bits(t, go, q.cbfCtx, 1)));
    int foundLast = 0, off, shift;
    last_ctx_params(isChroma, log2n, &off, &shift);
    const int bx0 = CTX_LASTX + (isChroma ? 15 : 0) + off, by0 = CTX_LASTY + (isChroma ? 15 : 0) + off, gmax = group_idx(n - 1);
    // bits of the last-position prefixes per group index (estLastSignificantPositionBit, TEncSbac.cpp:1862)
    int tabX[10], tabY[10];
    {
      int ax = 0, ay = 0;
      FHC_NOUNROLL for (int gi = 0; gi <= gmax; gi++)
      {
        tabX[gi] = ax + (gi < gmax ? ctx_bits(t, go, bx0 + (gi >> shift), 0) : 0);
        tabY[gi] = ay + (gi < gmax ? ctx_bits(t, go, by0 + (gi >> shift), 0) : 0);
        if (gi < gmax) { ax += ctx_bits(t, go, bx0 + (gi >> shift), 1); ay += ctx_bits(t, go, by0 + (gi >> shift), 1); }
      }
    }
    FHC_NOUNROLL for (int cgScanPos = cgLastScanPos; cgScanPos >= 0 && !foundLast; cgScanPos--)
    {
      const int cgBlk = scanCG[cgScanPos];
      baseCost = dsub(baseCost, costCGSig[cgScanPos]);
      if (!((cgFlags >> cgBlk) & 1)) continue;
      FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
      {
        const int scanPos = cgScanPos * 16 + posInCG;
        if (scanPos > lastScanPos) continue;
        const int blkPos = scan[scanPos];
        const int level = coef[blkPos], pk = B[scanPos];
        if (level)
        {
          const double costSigP = pk_nosig(pk) ? 0.0 : dmul(lambda, (double)ctx_bits(t, go, sigBase + pk_sig(pk), 1));
          const int posY = blkPos >> log2n, posX = blkPos - (posY << log2n);
          const int lx = q.scanType == SCAN_VER ? posY : posX, ly = q.scanType == SCAN_VER ? posX : posY;
          const int gx = group_idx(lx), gy = group_idx(ly);
          double cost = (double)(tabX[gx] + tabY[gy]);
          if (gx > 3) cost = dadd(cost, dmul(32768.0, (double)((gx - 2) >> 1)));
          if (gy > 3) cost = dadd(cost, dmul(32768.0, (double)((gy - 2) >> 1)));
          const double costLast = dmul(lambda, cost);
          const double totalCost = dsub(dadd(baseCost, costLast), costSigP);
          if (totalCost < bestCost) { bestLastIdxP1 = scanPos + 1; bestCost = totalCost; }
          if (level > 1) { foundLast = 1; break; }
          // level == 1: costCoeff of this position as the main pass computed it, and its uncoded cost
          const int ld = iabs(A[scanPos]);
          const double err = (double)(ld - (int)(1u << qBits)), err0 = (double)ld;
          double cur = dadd(dmul(dmul(err, err), errScale),
                            dmul(lambda, (double)ic_rate_call(t, go, 1u, pk_one(pk), pk_abs(pk), pk_rice(pk), pk_c1lt8(pk), pk_c2lt1(pk))));
          cur = dadd(cur, costSigP);
          baseCost = dsub(baseCost, cur);
          baseCost = dadd(baseCost, dmul(dmul(err0, err0), errScale));
        }
        else baseCost = dsub(baseCost, dmul(lambda, (double)ctx_bits(t, go, sigBase + pk_sig(pk), 0)));
      }
    }
  return bestLastIdxP1;
}

// sign bit hiding (TComTrQuant.cpp:2530-2660); rate increments of level +-1 are re-evaluated from the packed coder state
FHC_D void rdoq_sbh(const Blk* b, const Sbac* go, const RdoqIn& q, const int32_t* A, const int32_t* B, int16_t* coef, uint32_t* cgInfo, int numCodedCG)
{
  FHC_SHARED(b); FHC_SHARED(go);
  const Tables& t = b->t;
  const int log2n = q.log2n, isChroma = q.isChroma, qBits = q.qBits;
  const double lambda = q.lambda;
  const uint16_t* scan = scan_tab(t, q.scanType, log2n);
  const int sigBase = CTX_SIG + (isChroma ? 28 : 0);
    const double invQ = (double)inv_quant_scale(q.rem);
#ifdef FHC_DEVICE_BUILD
    const double f = __dadd_rn(__ddiv_rn(__ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(invQ, invQ), (double)(1 << (2 * q.per))), lambda), 16.0),
                                         (double)(1 << (2 * (q.bd - 8)))), 0.5);
#else
    const double f = invQ * invQ * (double)(1 << (2 * q.per)) / lambda / 16 / (double)(1 << (2 * (q.bd - 8))) + 0.5;
#endif
    const int64_t rdFactor = (int64_t)f;
    int lastCG = -1;
    FHC_NOUNROLL for (int subSet = numCodedCG - 1; subSet >= 0; subSet--)
    {
      const int subPos = subSet << 4;
      const uint32_t mask = cgInfo[subSet] & 0xffffu;
      const int sum = (int)(cgInfo[subSet] >> 16);                      // only its parity is used
      if (!mask) continue;
      const int lastNZ = msb_index(mask), firstNZ = lsb_index(mask);
      if (lastCG == -1) lastCG = 1;
      if (lastNZ - firstNZ >= 4)
      {
        const uint32_t signbit = A[subPos + firstNZ] >= 0 ? 0u : 1u;             // sign of the first non-zero level = sign of its coefficient
        if (signbit != (uint32_t)(sum & 1))
        {
          int64_t minCostInc = INT64_MAX, curCost = INT64_MAX;
          int minPos = -1, finalChange = 0, curChange = 0;
          FHC_NOUNROLL for (int nn = (lastCG == 1 ? lastNZ : 15); nn >= 0; --nn)
          {
            const int scanPos = nn + subPos, blkPos = scan[scanPos];
            const int lv = coef[blkPos], pk = B[scanPos], a = A[scanPos], ld = iabs(a);
            const int deltaU = (ld - (int)((uint32_t)lv << qBits)) >> (qBits - 8);
            const int sigD = pk_nosig(pk) ? 0 : ctx_bits(t, go, sigBase + pk_sig(pk), 1) - ctx_bits(t, go, sigBase + pk_sig(pk), 0);
            if (lv != 0)
            {
              const int rateNow = ic_rate_call(t, go, (uint32_t)lv, pk_one(pk), pk_abs(pk), pk_rice(pk), pk_c1lt8(pk), pk_c2lt1(pk));
              const int incUp = ic_rate_call(t, go, (uint32_t)lv + 1, pk_one(pk), pk_abs(pk), pk_rice(pk), pk_c1lt8(pk), pk_c2lt1(pk)) - rateNow;
              const int incDown = ic_rate_call(t, go, (uint32_t)lv - 1, pk_one(pk), pk_abs(pk), pk_rice(pk), pk_c1lt8(pk), pk_c2lt1(pk)) - rateNow;
              const int64_t costUp = rdFactor * (-deltaU) + incUp;
              int64_t costDown = rdFactor * (deltaU) + incDown - ((lv == 1) ? sigD : 0);
              if (lastCG == 1 && lastNZ == nn && lv == 1) costDown -= (4 << 15);
              if (costUp < costDown) { curCost = costUp; curChange = 1; }
              else
              {
                curChange = -1;
                if (nn == firstNZ && lv == 1) curCost = INT64_MAX;
                else curCost = costDown;
              }
            }
            else
            {
              curCost = rdFactor * (-(iabs(deltaU))) + (1 << 15) + ctx_bits(t, go, pk_one(pk), 0) + sigD;
              curChange = 1;
              if (nn < firstNZ)
              {
                const uint32_t thisSign = a >= 0 ? 0u : 1u;
                if (thisSign != signbit) curCost = INT64_MAX;
              }
            }
            if (curCost < minCostInc) { minCostInc = curCost; finalChange = curChange; minPos = scanPos; }
          }
          {
            const int blkPos = scan[minPos];
            int lv = coef[blkPos];
            if (lv == 32767 && A[minPos] >= 0) finalChange = -1;       // HM tests the signed level against 32767 / -32768
            lv += finalChange;
            coef[blkPos] = (int16_t)lv;
            if (lv) cgInfo[subSet] |= 1u << (minPos - subPos); else cgInfo[subSet] &= ~(1u << (minPos - subPos));
          }
        }
      }
      if (lastCG == 1) lastCG = 0;
    }
}

FHC_D uint32_t rdoq_serial(const Blk* b, const Sbac* go, const RdoqIn& q, const int32_t* A, int32_t* B, int16_t* coef, double* costCGSig,
                           uint32_t cgNZ0, uint32_t cgNZ1)
{
  FHC_SHARED(b); FHC_SHARED(go);
  const Tables& t = b->t;
  const int log2n = q.log2n, isChroma = q.isChroma, n = 1 << log2n, numCoef = n * n, log2wg = log2n - 2;
  const int qBits = q.qBits;
  const double errScale = q.errScale, lambda = q.lambda;
  const uint16_t* scan = scan_tab(t, q.scanType, log2n);
  const uint8_t* scanCG = scan_cg(t, q.scanType, log2n);
  const int wg = n >> 2, cgNum = numCoef >> 4;
  const int firstSig = first_sig_ctx(log2n, q.scanType, isChroma);
  const int sigBase = CTX_SIG + (isChroma ? 28 : 0);
  const int cgBase = CTX_SIGCG + (isChroma ? 2 : 0);
  double blockUncoded = 0, baseCost = 0;
  int cgLastScanPos = -1, lastScanPos = -1, c1 = 1, c2 = 0, stPack = 0;   // stPack: the coder state as it is stored in B[] (bits 0..12)
  uint32_t ctxSet = 0, c1Idx = 0, c2Idx = 0, goRice = 0;
  uint64_t cgFlags = 0;

  // no candidate level anywhere: every coefficient stays 0 (TComTrQuant.cpp:2462 returns uiAbsSum = 0 after summing costs nobody uses)
  if (!(cgNZ0 | cgNZ1)) return 0;

  FHC_NOUNROLL for (int cgScanPos = cgNum - 1; cgScanPos >= 0; cgScanPos--)
  {
    const int cgBlk = scanCG[cgScanPos], cgY = cgBlk >> log2wg, cgX = cgBlk & (wg - 1);
    const int cgHasCand = (int)(((cgScanPos < 32 ? cgNZ0 : cgNZ1) >> (cgScanPos & 31)) & 1u);
    costCGSig[cgScanPos] = 0;
    if (!cgHasCand && lastScanPos < 0)
    {
      // nothing coded yet and nothing to code here: only the uncoded distortion accumulates
      FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
      {
        const double err = (double)iabs(A[cgScanPos * 16 + posInCG]);
        const double c0 = dmul(dmul(err, err), errScale);
        blockUncoded = dadd(blockUncoded, c0);
        baseCost = dadd(baseCost, c0);
        coef[scan[cgScanPos * 16 + posInCG]] = 0;
      }
      continue;
    }
    const int pat = pattern_sig_ctx(cgFlags, cgX, cgY, log2wg);
    // sig_coeff_flag context increment of position p of this group = sigAdd + sigRow[p] (0 for the DC coefficient)
    const uint8_t* sigRow = t.sigScan[q.scanType][log2n == 2 ? 4 : pat];
    const int sigAdd = firstSig + ((log2n != 2 && cgBlk != 0 && !isChroma) ? 3 : 0);
    double sigCost = 0, sigCost0 = 0, codedLevelAndDist = 0, uncodedDist = 0;
    int nnzBeforePos0 = 0, cgHasLevel = 0;
    FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--)
    {
      const int scanPos = cgScanPos * 16 + posInCG;
      const int blkPos = scan[scanPos];
      const int levelDouble = iabs(A[scanPos]);
      const uint32_t m0 = (uint32_t)((levelDouble + (1 << (qBits - 1))) >> qBits);
      const uint32_t maxAbsLevel = m0 < 32767u ? m0 : 32767u;
      const double err0 = (double)levelDouble;
      const double c0 = dmul(dmul(err0, err0), errScale);
      blockUncoded = dadd(blockUncoded, c0);
      if (lastScanPos < 0)
      {
        if (maxAbsLevel == 0) { baseCost = dadd(baseCost, c0); coef[blkPos] = 0; continue; }
        lastScanPos = scanPos;
        ctxSet = (uint32_t)((isChroma ? 4 : 0) + ((!isChroma && (scanPos >> 4) > 0) ? 2 : 0));
        cgLastScanPos = cgScanPos;
        stPack = (int)(4 * ctxSet + 1) | ((int)ctxSet << 5) | (3 << 11);   // c1 = 1, c2 = 0, goRice = 0, c1Idx < 8, c2Idx < 1
      }
      const int isLast = scanPos == lastScanPos;
      const int sc = (isLast || scanPos == 0) ? 0 : sigAdd + sigRow[posInCG];
      B[scanPos] = stPack | (sc << 13) | (isLast << 19);
      if (maxAbsLevel == 0)
      {
        // level 0 (never the last position): only the significance flag is paid for; the coder state does not move
        const double s0 = dmul(lambda, (double)ctx_bits(t, go, sigBase + sc, 0));
        coef[blkPos] = 0;
        baseCost = dadd(baseCost, dadd(c0, s0));
        sigCost = dadd(sigCost, s0);
        if (posInCG == 0) sigCost0 = s0;
      }
      else
      {
        const int ctxOne = CTX_ONE + 4 * (int)ctxSet + c1, ctxAbs = CTX_ABS + (int)ctxSet + c2;
        // xGetCodedLevel (TComTrQuant.cpp:2812)
        double codedCost, codedCostSig = 0;
        uint32_t level = 0;
        if (!isLast && maxAbsLevel < 3)
        {
          codedCostSig = dmul(lambda, (double)ctx_bits(t, go, sigBase + sc, 0));
          codedCost = dadd(c0, codedCostSig);
        }
        else codedCost = FHC_MAX_COST;
        {
          const double currCostSig = isLast ? 0.0 : dmul(lambda, (double)ctx_bits(t, go, sigBase + sc, 1));
          const uint32_t minAbs = maxAbsLevel > 1 ? maxAbsLevel - 1 : 1;
          FHC_NOUNROLL for (int absLevel = (int)maxAbsLevel; absLevel >= (int)minAbs; absLevel--)
          {
            const double err = (double)(levelDouble - (int)((uint32_t)absLevel << qBits));
            double cur = dadd(dmul(dmul(err, err), errScale),
                              dmul(lambda, (double)ic_rate(t, go, (uint32_t)absLevel, ctxOne, ctxAbs, goRice, c1Idx < 8, c2Idx < 1)));
            cur = dadd(cur, currCostSig);
            if (cur < codedCost) { level = (uint32_t)absLevel; codedCost = cur; codedCostSig = currCostSig; }
          }
        }
        coef[blkPos] = (int16_t)level;
        baseCost = dadd(baseCost, codedCost);
        sigCost = dadd(sigCost, codedCostSig);
        if (posInCG == 0) sigCost0 = codedCostSig;
        if (level)
        {
          const int baseLevel = (c1Idx < 8) ? (2 + (c2Idx < 1)) : 1;
          if ((int)level >= baseLevel)
            if (level > 3u * (1u << goRice)) goRice = goRice + 1 < 4 ? goRice + 1 : 4;
          c1Idx++;
          if (level > 1) { c1 = 0; c2 += (c2 < 2); c2Idx++; }
          else if (c1 < 3 && c1 > 0) c1++;
          stPack = (4 * (int)ctxSet + c1) | (((int)ctxSet + c2) << 5) | ((int)goRice << 8) | ((c1Idx < 8) << 11) | ((c2Idx < 1) << 12);
          cgHasLevel = 1;
          codedLevelAndDist = dadd(codedLevelAndDist, dsub(codedCost, codedCostSig));
          uncodedDist = dadd(uncodedDist, c0);
          if (posInCG != 0) nnzBeforePos0++;
        }
      }
      if (posInCG == 0 && scanPos > 0)
      {
        ctxSet = (uint32_t)((isChroma ? 4 : 0) + ((!isChroma && ((scanPos - 1) >> 4) > 0) ? 2 : 0) + (c1 == 0 ? 1 : 0));
        c1 = 1; c2 = 0; c1Idx = 0; c2Idx = 0; goRice = 0;
        stPack = (int)(4 * ctxSet + 1) | ((int)ctxSet << 5) | (3 << 11);
      }
    }
    if (cgLastScanPos >= 0)
    {
      if (cgScanPos)
      {
        if (!cgHasLevel)
        {
          const int cx = cgBase + sig_cg_ctx(cgFlags, cgX, cgY, log2wg);
          const double c = dmul(lambda, (double)ctx_bits(t, go, cx, 0));
          baseCost = dadd(baseCost, dsub(c, sigCost));
          costCGSig[cgScanPos] = c;
        }
        else if (cgScanPos < cgLastScanPos)
        {
          if (nnzBeforePos0 == 0) { baseCost = dsub(baseCost, sigCost0); sigCost = dsub(sigCost, sigCost0); }
          double costZeroCG = baseCost;
          const int cx = cgBase + sig_cg_ctx(cgFlags, cgX, cgY, log2wg);
          const double cg1 = dmul(lambda, (double)ctx_bits(t, go, cx, 1)), cg0 = dmul(lambda, (double)ctx_bits(t, go, cx, 0));
          baseCost = dadd(baseCost, cg1);
          costZeroCG = dadd(costZeroCG, cg0);
          costCGSig[cgScanPos] = cg1;
          costZeroCG = dadd(costZeroCG, uncodedDist);
          costZeroCG = dsub(costZeroCG, codedLevelAndDist);
          costZeroCG = dsub(costZeroCG, sigCost);
          if (costZeroCG < baseCost)
          {
            cgHasLevel = 0;
            baseCost = costZeroCG;
            costCGSig[cgScanPos] = cg0;
            FHC_NOUNROLL for (int posInCG = 15; posInCG >= 0; posInCG--) coef[scan[cgScanPos * 16 + posInCG]] = 0;
          }
        }
      }
      else cgHasLevel = 1;
    }
    if (cgHasLevel) cgFlags |= (uint64_t)1 << cgBlk;
  }
  if (lastScanPos < 0) return 0;

  // best last position (TComTrQuant.cpp:2465-2524)
  RdoqState st;
  st.blockUncoded = blockUncoded; st.baseCost = baseCost; st.cgLastScanPos = cgLastScanPos; st.lastScanPos = lastScanPos; st.cgFlags = cgFlags;
  const int bestLastIdxP1 = rdoq_last_pos(b, go, q, A, B, coef, costCGSig, st);
  // per coefficient group of the coded range: non-zero map (bits 0..15) and parity of the level sum (bit 16), kept in
  // the group-cost scratch (no longer needed); consumed by sign bit hiding and by the sign pass
  uint32_t absSum = 0;
  const int numCodedCG = (bestLastIdxP1 + 15) >> 4;
  uint32_t* cgInfo = (uint32_t*)costCGSig;
  FHC_NOUNROLL for (int scanPos = bestLastIdxP1; scanPos <= lastScanPos; scanPos++) coef[scan[scanPos]] = 0;
  FHC_NOUNROLL for (int cg = 0; cg < numCodedCG; cg++)
  {
    const int p0 = cg << 4, cnt = imin(16, bestLastIdxP1 - p0);
    uint32_t mask = 0, sum = 0;
    FHC_NOUNROLL for (int i = 0; i < cnt; i++)
    {
      const uint32_t lv = (uint32_t)coef[scan[p0 + i]];
      if (lv) { mask |= 1u << i; sum += lv; }
    }
    absSum += sum;
    cgInfo[cg] = mask | ((sum & 1u) << 16);
  }
  if (absSum >= 2) rdoq_sbh(b, go, q, A, B, coef, cgInfo, numCodedCG);  // sign bit hiding (TComTrQuant.cpp:2530-2660)
  // apply the signs of the coefficients to the non-zero levels
  FHC_NOUNROLL for (int cg = 0; cg < numCodedCG; cg++)
  {
    uint32_t m = cgInfo[cg] & 0xffffu;
    FHC_NOUNROLL while (m)
    {
      const int i = lsb_index(m);
      m &= m - 1;
      if (A[(cg << 4) + i] < 0) { const int blkPos = scan[(cg << 4) + i]; coef[blkPos] = (int16_t)(-coef[blkPos]); }
    }
  }
  return absSum;
}

// ----------------------------------------------------------------------------------------------
// TU evaluation = xIntraCodingTUBlock (TEncSearch.cpp:1104-1399) for one block of one component, by a lane group:
// prediction, residual, DCT/DST or transform skip, RDOQ, dequantisation, inverse transform, reconstruction, SSE.
// ----------------------------------------------------------------------------------------------
struct GScr                       // a group's scratch (see Role)
{
  Sbac* go;
  int32_t *A, *B;
  double* cgCost; uint16_t* nz; uint32_t* acc;     // acc[0] sse, acc[1..2] cgNZ, acc[3] broadcast word
  int16_t *rec, *coef;            // candidate planes (yuv layout of the role's CU)
  uint8_t* meta;                  // chroma cbf[2][np], ts[2][np]
  int16_t *refU, *refF;
};
// planeStride: elements between the candidate planes of consecutive groups (the role's yuvSize, or 16 for the 4x4
// luma candidates of the NxN role, which only ever hold one 4x4 block)
FHC_DH GScr group_scratch(const Role* r, int idx, int planeStride)
{
  GScr s;
  s.go = r->go + idx;
  s.A = r->tuA + idx * r->tuCap; s.B = r->tuB + idx * r->tuCap;
  s.cgCost = r->cgCost + idx * (r->tuCap >> 4); s.nz = r->nzMask + idx * (r->yuvSize >> 4); s.acc = r->gAcc + idx * 4;
  s.rec = r->gRec + idx * planeStride; s.coef = r->gCoef + idx * planeStride;
  s.meta = r->gMeta + (idx < 5 ? idx : 0) * 4 * r->np;
  if (idx >= 5) idx = 0;                           // only the (at most five) chroma / 64x64 groups build reference lines of their own
  s.refU = r->gRefU + idx * r->refCap; s.refF = r->gRefF + idx * r->refCap;
  FHC_SHARED(s.go); FHC_SHARED(s.acc);
  return s;
}

struct TuArgs
{
  int comp, log2n, mode, scanType, ts, cbfCtx;
  int srcOff, srcStride;          // source block inside Blk::src
  int16_t* rec; int recStride;    // prediction -> reconstruction (block origin)
  int16_t* coef;                  // level block of this TU (n*n, raster)
  const int16_t *refU, *refF; int dcVal;
  int phases;                     // TU_FRONT | TU_RDOQ | TU_BACK: which parts of the evaluation this call runs (state stays in the group scratch)
  int alignRole;                  // >= 0: whole-warp call of that role, twin-aligned before the RDOQ pass and before the reconstruction
};
enum { TU_FRONT = 1, TU_RDOQ = 2, TU_BACK = 4, TU_ALL = 7 };

FHC_DH int tmat_at(const Tables& t, int n, int useDst, int k, int q)
{
  return useDst ? t.dst4[k * 4 + q] : t.T32[(k * (32 / n)) * 32 + q];
}

// One stage of a separable NxN transform, striped over the group's lanes:
//   out[i] = post((sum_q m[q * ms] * in[q * is] + add) >> shift),  i = hi * N + lo,
//   m = mbase + hi * mHi + lo * mLo, in = ibase + hi * iHi + lo * iLo; post = clipping to 16 bits for the inverse stages
// (xTrMxN / xITrMxN, TComTrQuant.cpp:860-985).  One copy serves all four stages.
FHC_D void mat_stage(const Grp& g, int log2n, const int16_t* mbase, int mHi, int mLo, int ms, const int32_t* ibase, int iHi, int iLo, int is,
                     int32_t* out, int add, int shift, int clip)
{
  FHC_SHARED(mbase);
  const int n = 1 << log2n, n2 = n * n;
  FHC_GLANES(g, i, n2)
  {
    const int hi = i >> log2n, lo = i & (n - 1);
    const int16_t* m = mbase + hi * mHi + lo * mLo;
    const int32_t* in = ibase + hi * iHi + lo * iLo;
    int acc = 0;
#ifdef FHC_DEVICE_BUILD
#pragma unroll 4
#endif
    for (int q = 0; q < n; q++) acc += m[q * ms] * in[q * is];
    acc = (acc + add) >> shift;
    out[i] = clip ? clip3(-32768, 32767, acc) : acc;
  }
}

// The DCT matrices are even / odd symmetric, M(k, N-1-q) = (-1)^k M(k, q), which halves the multiplications exactly as
// HM's partial butterflies do (TComTrQuant.cpp:507-840; integer sums, so the result is identical):
// fold_stage replaces (x[q], x[N-1-q]) by (x[q] + x[N-1-q], x[q] - x[N-1-q]) along one axis of an N x N int32 block;
FHC_D void fold_stage(const Grp& g, int log2n, int32_t* buf, int sa, int sq)
{
  const int n = 1 << log2n, half = n >> 1;
  FHC_GLANES(g, i, n * half)
  {
    const int a = i >> (log2n - 1), q = i & (half - 1);
    int32_t* x = buf + a * sa + q * sq;
    int32_t* y = buf + a * sa + (n - 1 - q) * sq;
    const int u = *x, v = *y;
    *x = u + v; *y = u - v;
  }
}
// fwd_stage: out[i] = (sum_{q < N/2} M(k, q) * f[a][k even ? q : N-1-q] + add) >> shift on a folded input f, where
// (k, a) = (lo, hi) or (hi, lo) of i = hi * N + lo; M(k, q) = mbase[k * mk + q * mq], f[a][q'] = in[a * ia + q' * iq];
FHC_D void fwd_stage(const Grp& g, int log2n, int kIsLo, const int16_t* mbase, int mk, int mq, const int32_t* in, int ia, int iq,
                     int32_t* out, int add, int shift)
{
  FHC_SHARED(mbase);
  const int n = 1 << log2n, n2 = n * n, half = n >> 1;
  FHC_GLANES(g, i, n2)
  {
    const int hi = i >> log2n, lo = i & (n - 1), k = kIsLo ? lo : hi, a = kIsLo ? hi : lo;
    const int16_t* m = mbase + k * mk;
    const int32_t* f = in + a * ia + ((k & 1) ? (n - 1) * iq : 0);
    const int fs = (k & 1) ? -iq : iq;
    int acc = 0;
#ifdef FHC_DEVICE_BUILD
#pragma unroll 4
#endif
    for (int q = 0; q < half; q++) acc += m[q * mq] * f[q * fs];
    out[i] = (acc + add) >> shift;
  }
}
// inv_stage: one lane computes out[j][i] and out[j][N-1-i] = clip((E +- O + add) >> shift) with E / O the sums over
// the even / odd rows k of M(k, i) * in[k][j]  (xITrMxN, TComTrQuant.cpp:927; partialButterflyInverse*, :476-840).
FHC_D void inv_stage(const Grp& g, int log2n, const int16_t* mat, int rs, const int32_t* in, int32_t* out, int add, int shift)
{
  FHC_SHARED(mat);
  const int n = 1 << log2n, half = n >> 1;
  FHC_GLANES(g, idx, n * half)
  {
    const int j = idx >> (log2n - 1), i = idx & (half - 1);
    const int16_t* m = mat + i;
    const int32_t* c = in + j;
    int e = 0, o = 0;
#ifdef FHC_DEVICE_BUILD
#pragma unroll 2
#endif
    for (int k = 0; k < n; k += 2) { e += m[k * rs] * c[k * n]; o += m[(k + 1) * rs] * c[(k + 1) * n]; }
    out[j * n + i] = clip3(-32768, 32767, (e + o + add) >> shift);
    out[j * n + n - 1 - i] = clip3(-32768, 32767, (e - o + add) >> shift);
  }
}

// returns the distortion (chroma: weighted, TComRdCost.cpp:352); *cbfOut = uiAbsSum > 0.  All lanes of the group
// return the same values.
FHC_D uint32_t tu_eval(Blk* b, const Grp& g, const GScr& s, const TuArgs& a, int* cbfOut)
{
  FHC_SHARED(b);
  const Tables& t = b->t;
  Sbac* const sgo = s.go;
  uint32_t* const sacc = s.acc;
  FHC_SHARED(sgo); FHC_SHARED(sacc);
  const int comp = a.comp, log2n = a.log2n, n = 1 << log2n, n2 = n * n, bd = b->pp.bitDepth, isChroma = comp != 0;
  const int useDst = (comp == 0 && log2n == 2);
  const int16_t* line = use_filtered_refs(t, comp, a.mode, log2n) ? a.refF : a.refU;
  const int16_t* src = b->src + a.srcOff;
  // M(k, q) = mat[k * rs + q]; its transpose Mt(q, k) = matT[q * rsT + k * ksT] (lanes that differ in k read neighbouring words)
  const int16_t* mat = useDst ? t.dst4 : t.T32;
  const int16_t* matT = useDst ? t.dst4t : t.T32t;
  const int rs = useDst ? 4 : (1 << (10 - log2n)), rsT = useDst ? 4 : 32, ksT = useDst ? 1 : (1 << (5 - log2n));
  const QpParam qp = qp_param(b, comp);
  const int qBits = 14 + qp.per + (15 - bd - log2n);
  if (a.phases & TU_FRONT)
  {
  FHC_COUNT(if (g.gl == 0) { g_emuCounters.tqSamples[isChroma][log2n - 2] += (unsigned)n2; g_emuCounters.tuCalls[isChroma][log2n - 2]++; });
  gsync(g);                                                            // every lane has consumed the accumulators of the previous TU
  if (g.gl == 0) { sacc[0] = 0; sacc[1] = 0; sacc[2] = 0; }
  // prediction + residual -> B
  const PredMode pm = pred_mode(t, a.mode, n, log2n, comp == 0);
  FHC_GLANES(g, i, n2)
  {
    const int x = i & (n - 1), y = i >> log2n;
    const int v = pred_at(pm, line, n, x, y, bd, a.dcVal);
    a.rec[y * a.recStride + x] = (int16_t)v;
    s.B[i] = src[y * a.srcStride + x] - v;
  }
  gsync(g);
  // forward transform (xTrMxN, TComTrQuant.cpp:860) or transform skip (:2011) -> B (raster)
  if (a.ts)
  {
    const int shift = 15 - bd - log2n;
    FHC_GLANES(g, i, n2) s.B[i] = (int32_t)((uint32_t)s.B[i] << shift);
  }
  else
  {
    const int shift1 = log2n + bd + 6 - 15, add1 = shift1 > 0 ? 1 << (shift1 - 1) : 0;
    const int shift2 = log2n + 6, add2 = 1 << (shift2 - 1);
    // rows of the residual: A[j][k] = sum_q M(k, q) * resi[j][q] (stored transposed w.r.t. HM's tmp[k][j]: coalesced both ways),
    // then columns: B[k2][j2] = sum_q M(k2, q) * A[q][j2]
    if (log2n == 2)
    {
      mat_stage(g, log2n, matT, 0, ksT, rsT, s.B, n, 0, 1, s.A, add1, shift1, 0);
      gsync(g);
      mat_stage(g, log2n, mat, rs, 0, 1, s.A, 0, 1, n, s.B, add2, shift2, 0);
    }
    else
    {
      fold_stage(g, log2n, s.B, n, 1);                                  // along the rows of the residual
      gsync(g);
      fwd_stage(g, log2n, 1, matT, ksT, rsT, s.B, n, 1, s.A, add1, shift1);
      gsync(g);
      fold_stage(g, log2n, s.A, 1, n);                                  // along the columns of the intermediate block
      gsync(g);
      fwd_stage(g, log2n, 0, mat, rs, 1, s.A, 1, n, s.B, add2, shift2);
    }
  }
  gsync(g);
  // RDOQ pre-quantisation (TComTrQuant.cpp:2244-2262), independent per coefficient: A[p] = signed lLevelDouble
  {
    const uint16_t* scan = scan_tab(t, a.scanType, log2n);
    const int quantCoef = quant_scale(qp.rem);
    const int64_t cap = (int64_t)2147483647 - ((int64_t)1 << (qBits - 1));
    FHC_GLANES(g, p, n2)
    {
      const int c = s.B[scan[p]];
      const int64_t tl = (int64_t)iabs(c) * quantCoef;
      const int ld = (int)(tl < cap ? tl : cap);
      s.A[p] = c < 0 ? -ld : ld;
      if ((ld + (1 << (qBits - 1))) >> qBits) gor_u32(g, &sacc[1 + (p >> 9)], 1u << ((p >> 4) & 31));
    }
  }
  gsync(g);
  }
  if (a.alignRole >= 0) twin_align(b, a.alignRole);
  if ((a.phases & TU_RDOQ) && g.gl == 0)
  {
    RdoqIn q;
    q.log2n = log2n; q.isChroma = isChroma; q.scanType = a.scanType; q.cbfCtx = a.cbfCtx; q.qBits = qBits; q.per = qp.per; q.rem = qp.rem; q.bd = bd;
    q.lambda = b->pp.sp.rdoqLambda[comp]; q.errScale = t.errScale[isChroma][log2n - 2][qp.rem];
    sacc[3] = rdoq_serial(b, sgo, q, s.A, s.B, a.coef, s.cgCost, sacc[1], sacc[2]);
  }
  if (a.alignRole >= 0) twin_align(b, a.alignRole);
  if (!(a.phases & TU_BACK)) return 0;
  gsync(g);
  const int coded = sacc[3] > 0;
  if (coded)
  {
    // xDeQuant (TComTrQuant.cpp:1308, flat scaling) -> A (raster)
    const int transformShift = 15 - bd - log2n, rightShift = 6 - (transformShift + qp.per);
    const int scale = inv_quant_scale(qp.rem);
    const int targetBits = imin(16, 32 + rightShift - 7);
    const int inMin = -(1 << (targetBits - 1)), inMax = (1 << (targetBits - 1)) - 1;
    FHC_GLANES(g, i, n2)
    {
      const int qv = clip3(inMin, inMax, (int)a.coef[i]);
      int v;
      if (rightShift > 0) v = (qv * scale + (1 << (rightShift - 1))) >> rightShift;
      else v = (int)((uint32_t)(qv * scale) << (-rightShift));
      s.A[i] = clip3(-32768, 32767, v);
    }
    gsync(g);
    // inverse transform skip (:2060) or xITrMxN (:927, columns first, clipping after each stage) -> residual in A
    if (a.ts)
    {
      const int shift = transformShift, off = shift == 0 ? 0 : (1 << (shift - 1));
      FHC_GLANES(g, i, n2) s.A[i] = (int16_t)((s.A[i] + off) >> shift);
    }
    else
    {
      const int shift2 = 12 - (bd - 8);
      if (log2n == 2)
      {
        mat_stage(g, log2n, mat, 0, 1, rs, s.A, 1, 0, n, s.B, 64, 7, 1);
        gsync(g);
        mat_stage(g, log2n, mat, 0, 1, rs, s.B, 1, 0, n, s.A, 1 << (shift2 - 1), shift2, 1);
      }
      else
      {
        inv_stage(g, log2n, mat, rs, s.A, s.B, 64, 7);
        gsync(g);
        inv_stage(g, log2n, mat, rs, s.B, s.A, 1 << (shift2 - 1), shift2);
      }
    }
  }
  else
  {
    FHC_GLANES(g, i, n2) a.coef[i] = 0;
  }
  gsync(g);
  // reconstruction (TEncSearch.cpp:1380-1393) + SSE (TComRdCost.cpp:1176)
  {
    const int shift = (bd - 8) << 1, maxv = (1 << bd) - 1;
    uint32_t acc = 0;
    FHC_GLANES(g, i, n2)
    {
      const int x = i & (n - 1), y = i >> log2n;
      const int v = clip3(0, maxv, a.rec[y * a.recStride + x] + (coded ? s.A[i] : 0));
      a.rec[y * a.recStride + x] = (int16_t)v;
      const int d = v - src[y * a.srcStride + x];
      acc += (uint32_t)((d * d) >> shift);
    }
    if (acc) gadd_u32(g, &sacc[0], acc);
  }
  gsync(g);
  uint32_t dist = sacc[0];
  if (isChroma) dist = (uint32_t)dmul(b->pp.sp.chromaDistWeight[comp - 1], (double)dist);
  *cbfOut = coded;
  return dist;
}

// ----------------------------------------------------------------------------------------------
// hypothesis view used by the syntax walkers: TU tree + flags + levels of (a part of) one CU
// ----------------------------------------------------------------------------------------------
struct Hyp
{
  int depth, np, log2cu, cuw, partSize;
  const uint8_t* tridx;           // [np]
  const uint8_t* cbf[3];          // [np] each, bit t = cbf at transform depth t
  const uint8_t* ts[3];
  const int16_t* coefc[3];        // per component: levels of the CU, per-TU raster blocks at 16 * p (luma) / 4 * p (chroma), p = first partition of the TU
  uint16_t* nzc[3];               // per component: non-zero maps per coefficient group, index = (block offset >> 4) + raster position of the group in its TU
  uint8_t ldir[4]; uint8_t cdir;
};
FHC_DH int min_tu_log2(int log2cu, int isNxN)                                                  // TComDataCU.cpp:1478
{
  if (log2cu < 2 + 3 - 1 + isNxN) return 2;
  const int rr = log2cu - (3 - 1 + isNxN);
  return rr > 5 ? 5 : rr;
}
FHC_DH int chroma_pred_mode(const Hyp& h) { return h.cdir == DM_CHROMA_IDX ? h.ldir[0] : h.cdir; }

// Non-zero maps of every coefficient group of the TUs in partitions [p0, p0 + cnt) of a hypothesis (lanes over 4x4
// luma partitions = luma coefficient groups; every 4th one also does the co-located chroma groups).
FHC_D void nz_masks_range(const Tables& t, const Grp& g, const Hyp& h, int p0, int cnt, int doLuma, int doChroma)
{
  FHC_SHARED(&t);
  FHC_GLANES(g, k, cnt)
  {
    const int p = p0 + k;
    const int T = h.tridx[p], leafParts = h.np >> (2 * T), leaf = p & ~(leafParts - 1), log2n = h.log2cu - T;
    if (doLuma)
    {
      // group position inside the TU
      const int dx = (zx(t, p) - zx(t, leaf)) >> 2, dy = (zy(t, p) - zy(t, leaf)) >> 2, n = 1 << log2n;
      const int pu = h.partSize == SIZE_NxN ? (p / (h.np >> 2)) : 0;
      const uint16_t* s4 = scan_tab(t, scan_idx_for(h.ldir[pu], log2n, 0), 2);
      const int16_t* c = h.coefc[0] + 16 * leaf + (dy * 4) * n + dx * 4;
      uint32_t m = 0;
      for (int i = 0; i < 16; i++) { const int v = s4[i]; m |= (uint32_t)(c[(v >> 2) * n + (v & 3)] != 0) << i; }
      h.nzc[0][leaf + (dy << (log2n - 2)) + dx] = (uint16_t)m;
    }
    if (doChroma && (p & 3) == 0)
    {
      // chroma group co-located with the 8x8 luma area of partitions p .. p+3
      const int cl = log2n == 2 ? 2 : log2n - 1, cn = 1 << cl;
      const int cleaf = log2n == 2 ? p : leaf;
      const int dx = (zx(t, p) - zx(t, cleaf)) >> 3, dy = (zy(t, p) - zy(t, cleaf)) >> 3;
      const uint16_t* s4 = scan_tab(t, scan_idx_for(chroma_pred_mode(h), cl, 1), 2);
      for (int cc = 1; cc < 3; cc++)
      {
        const int16_t* c = h.coefc[cc] + 4 * cleaf + (dy * 4) * cn + dx * 4;
        uint32_t m = 0;
        for (int i = 0; i < 16; i++) { const int v = s4[i]; m |= (uint32_t)(c[(v >> 2) * cn + (v & 3)] != 0) << i; }
        h.nzc[cc][(cleaf >> 2) + (dy << (cl - 2)) + dx] = (uint16_t)m;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// neighbours (TComDataCU.cpp:985-1200, :1362-1461)
// ----------------------------------------------------------------------------------------------
// TComDataCU::getIntraDirPredictor for the PU whose first partition is zabs (CTU z index).  The CU under test spans
// partitions [cuZ, cuZ + cuNp); its own luma modes are cuLdir[pu].  Returns *piMode (1 or 2: the number of MPMs HM appends).
FHC_D int get_mpm(const Blk* b, int cuZ, int cuNp, int isNxN, const uint8_t* cuLdir, int zabs, int mpm[3])
{
  FHC_SHARED(b);
  const Tables& t = b->t;
  const int r = t.z2r[zabs];
  int l, a;
  if (r & 15)
  {
    const int nz = t.r2z[r - 1];
    if (nz >= cuZ && nz < cuZ + cuNp) l = cuLdir[isNxN ? (nz - cuZ) / (cuNp >> 2) : 0];
    else l = b->ctuLdir[nz];
  }
  else l = b->ctuX > 0 ? b->leftLdir[r >> 4] : DC_IDX;
  if (r >> 4)
  {
    const int nz = t.r2z[r - 16];
    if (nz >= cuZ && nz < cuZ + cuNp) a = cuLdir[isNxN ? (nz - cuZ) / (cuNp >> 2) : 0];
    else a = b->ctuLdir[nz];
  }
  else a = DC_IDX;                                     // above the CTU: treated as unavailable (TComDataCU.cpp:1384)
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
FHC_DH int ctx_split(const Blk* b, int zabs, int depth)                                          // getCtxSplitFlag (TComDataCU.cpp:1447)
{
  const Tables& t = b->t;
  const int r = t.z2r[zabs];
  const int l = (r & 15) ? b->ctuDepth[t.r2z[r - 1]] : (b->ctuX > 0 ? b->leftDepth[r >> 4] : -1);
  const int a = (r >> 4) ? b->ctuDepth[t.r2z[r - 16]] : (b->ctuY > 0 ? b->aboveDepth[r & 15] : -1);
  return (l > depth) + (a > depth);
}
FHC_DH void allowed_chroma(int lumaMode, int list[5])                                            // TComDataCU.cpp:1334
{
  list[0] = PLANAR_IDX; list[1] = VER_IDX; list[2] = HOR_IDX; list[3] = DC_IDX; list[4] = DM_CHROMA_IDX;
  for (int i = 0; i < 4; i++) if (list[i] == lumaMode) { list[i] = 34; break; }
}

// ----------------------------------------------------------------------------------------------
// syntax element bit counting (TEncSbac.cpp), serial: called by one lane
// ----------------------------------------------------------------------------------------------
FHC_DH int mpm_index(int dir, const int mpm[3])
{
  int predIdx = -1;
  for (int i = 0; i < 3; i++) if (dir == mpm[i]) predIdx = i;
  return predIdx;
}
FHC_DH void code_luma_dir1(Coder& e, int dir, const int mpm[3])                                   // codeIntraDirLumaAng, one PU (TEncSbac.cpp:643)
{
  const int predIdx = mpm_index(dir, mpm);
  enc_bin(e, CTX_IPRED, predIdx != -1 ? 1 : 0);
  enc_ep(e, predIdx != -1 ? (predIdx ? 2 : 1) : 5);
}
FHC_DH void code_chroma_dir(Coder& e, int cdir)                                                   // TEncSbac.cpp:698
{
  if (cdir == DM_CHROMA_IDX) enc_bin(e, CTX_CPRED, 0);
  else { enc_bin(e, CTX_CPRED, 1); enc_ep(e, 2); }
}
FHC_DH void code_qt_cbf(Coder& e, const Hyp& h, int p, int comp, int tr, int lowestLevel, int compLog2n)   // TEncSbac.cpp:920 (4:2:0)
{
  const int ctx = CTX_CBF + (comp ? 5 + tr : (tr == 0 ? 1 : 0));
  const int lowest = tr + ((!lowestLevel && compLog2n < 3) ? 1 : 0);
  enc_bin(e, ctx, (h.cbf[comp][p] >> lowest) & 1);
}
FHC_DH int subdiv_flag_coded(const Hyp& h, int log2n, int tr)
{
  if (h.partSize == SIZE_NxN && tr == 0) return 0;
  if (log2n > 5 || log2n == 2) return 0;
  return log2n != min_tu_log2(h.log2cu, h.partSize == SIZE_NxN);
}
// first transform depth at which a node starts at partition p (p inside a sub-tree rooted at p0 / depth t0)
FHC_DH int first_node_depth(const Hyp& h, int p, int p0, int t0)
{
  if (p == p0) return t0;
  int tr = t0 + 1;
  while (p & ((h.np >> (2 * tr)) - 1)) tr++;
  return tr;
}
FHC_DH void code_coeff_of(Coder& e, const Hyp& h, int comp, int p, int log2n)                     // the TU of component comp whose block starts at partition p
{
  const int off = comp ? 4 * p : 16 * p;
  const int pu = h.partSize == SIZE_NxN ? (p / (h.np >> 2)) : 0;
  const int dir = comp ? chroma_pred_mode(h) : h.ldir[pu];
  code_coeff_serial(e, h.coefc[comp] + off, log2n, opq(comp != 0), scan_idx_for(dir, log2n, comp != 0), h.ts[comp][p], h.nzc[comp] + (off >> 4));
}

// xEncSubdivCbfQT (TEncSearch.cpp:881) over the sub-tree rooted at partition p0, transform depth t0
FHC_D void enc_subdiv_cbf(Coder& eio, const Hyp& h, int p0, int t0, int bLuma, int bChroma)
{
  Coder e = eio;
  FHC_SHARED(e.sb); FHC_SHARED(e.t);
  const int cnt = h.np >> (2 * t0);
  FHC_NOUNROLL for (int p = p0; p < p0 + cnt;)
  {
    const int T = h.tridx[p];
    FHC_NOUNROLL for (int tr = first_node_depth(h, p, p0, t0); tr <= T; tr++)
    {
      const int log2n = h.log2cu - tr, subdiv = T > tr;
      if (bLuma && subdiv_flag_coded(h, log2n, tr)) enc_bin(e, CTX_SUBDIV + (5 - log2n), subdiv);
      if (bChroma && (tr == t0 || log2n > 2))
        for (int c = 1; c < 3; c++)
          if (tr == 0 || ((h.cbf[c][p] >> (tr - 1)) & 1)) code_qt_cbf(e, h, p, c, tr, subdiv == 0, log2n - 1);
      if (!subdiv && bLuma) code_qt_cbf(e, h, p, 0, tr, 1, log2n);
    }
    p += h.np >> (2 * T);
  }
  eio = e;
}
// xEncCoeffQT (TEncSearch.cpp:958) for one component over a sub-tree; search-time order: the 4x4 chroma block of an
// 8x8 luma area split into four 4x4 TUs goes with the FIRST quadrant
FHC_D void enc_coeff_qt(Coder& eio, const Hyp& h, int p0, int t0, int comp)
{
  Coder e = eio;
  FHC_SHARED(e.sb); FHC_SHARED(e.t);
  const int cnt = h.np >> (2 * t0);
  FHC_NOUNROLL for (int p = p0; p < p0 + cnt;)
  {
    const int T = h.tridx[p], log2n = h.log2cu - T;
    if (comp == 0)
    {
      if ((h.cbf[0][p] >> T) & 1) code_coeff_of(e, h, 0, p, log2n);
    }
    else if (log2n > 2 || (p & 3) == 0)
    {
      if ((h.cbf[comp][p] >> T) & 1) code_coeff_of(e, h, comp, p, log2n == 2 ? 2 : log2n - 1);
    }
    p += h.np >> (2 * T);
  }
  eio = e;
}
// luma-dir part of xEncIntraHeader for the node at partition p0 / depth t0 (TEncSearch.cpp:1002)
FHC_DH void enc_intra_header_luma(Coder& e, const Blk* b, const Hyp& h, int cuZ, int p0, int t0)
{
  if (p0 == 0 && h.depth == 3) enc_bin(e, CTX_PART, h.partSize == SIZE_2Nx2N ? 1 : 0);
  int mpm[3];
  if (h.partSize == SIZE_2Nx2N)
  {
    if (p0 == 0) { get_mpm(b, cuZ, h.np, 0, h.ldir, cuZ, mpm); code_luma_dir1(e, h.ldir[0], mpm); }
  }
  else if (t0 > 0 && (p0 % (h.np >> 2)) == 0)
  {
    const int pu = p0 / (h.np >> 2);
    get_mpm(b, cuZ, h.np, 1, h.ldir, cuZ + p0, mpm);
    code_luma_dir1(e, h.ldir[pu], mpm);
  }
}
// xGetIntraBitsQT, luma only (TEncSearch.cpp:1067)
FHC_D uint32_t intra_bits_luma(Coder& eio, const Blk* b, const Hyp& h, int cuZ, int p0, int t0)
{
  FHC_SHARED(b);
  Coder e = eio;
  FHC_SHARED(e.sb); FHC_SHARED(e.t);
  reset_bits(e);
  enc_intra_header_luma(e, b, h, cuZ, p0, t0);
  enc_subdiv_cbf(e, h, p0, t0, 1, 0);
  enc_coeff_qt(e, h, p0, t0, 0);
  eio = e;
  return num_bits(e);
}
// xGetIntraBitsQT, chroma only, whole CU
FHC_D uint32_t intra_bits_chroma(Coder& eio, const Hyp& h)
{
  Coder e = eio;
  FHC_SHARED(e.sb); FHC_SHARED(e.t);
  reset_bits(e);
  code_chroma_dir(e, h.cdir);
  enc_subdiv_cbf(e, h, 0, 0, 0, 1);
  enc_coeff_qt(e, h, 0, 0, 1);
  enc_coeff_qt(e, h, 0, 0, 2);
  eio = e;
  return num_bits(e);
}
// encodePredInfo + encodeCoeff of a CU in real syntax order (TEncEntropy.cpp:200-448): the CU-level bit count of
// xCheckRDCostIntra (TEncCu.cpp:1564-1582) and the chain advance of xEncodeCU.  The 4x4 chroma block of a split 8x8
// luma area goes with the LAST quadrant here.
FHC_D void encode_cu_syntax(Coder& eio, const Blk* b, const Hyp& h, int cuZ)
{
  FHC_SHARED(b);
  Coder e = eio;
  FHC_SHARED(e.sb); FHC_SHARED(e.t);
  if (h.depth == 3) enc_bin(e, CTX_PART, h.partSize == SIZE_2Nx2N ? 1 : 0);
  {
    const int numPU = h.partSize == SIZE_NxN ? 4 : 1;
    int predIdx[4];
    for (int j = 0; j < numPU; j++)
    {
      int mpm[3];
      get_mpm(b, cuZ, h.np, numPU == 4, h.ldir, cuZ + j * (h.np >> 2), mpm);
      predIdx[j] = mpm_index(h.ldir[j], mpm);
      enc_bin(e, CTX_IPRED, predIdx[j] != -1 ? 1 : 0);
    }
    for (int j = 0; j < numPU; j++) enc_ep(e, predIdx[j] != -1 ? (predIdx[j] ? 2 : 1) : 5);
  }
  code_chroma_dir(e, h.cdir);
  FHC_NOUNROLL for (int p = 0; p < h.np;)
  {
    const int T = h.tridx[p];
    FHC_NOUNROLL for (int tr = first_node_depth(h, p, 0, 0); tr <= T; tr++)
    {
      const int log2n = h.log2cu - tr, subdiv = T > tr;
      if (subdiv_flag_coded(h, log2n, tr)) enc_bin(e, CTX_SUBDIV + (5 - log2n), subdiv);
      if (tr == 0 || log2n > 2)
        for (int c = 1; c < 3; c++)
          if (tr == 0 || ((h.cbf[c][p] >> (tr - 1)) & 1)) code_qt_cbf(e, h, p, c, tr, subdiv == 0, log2n - 1);
      if (!subdiv)
      {
        code_qt_cbf(e, h, p, 0, tr, 1, log2n);
        if ((h.cbf[0][p] >> tr) & 1) code_coeff_of(e, h, 0, p, log2n);
        if (log2n > 2 || (p & 3) == 3)
        {
          const int ridx = log2n == 2 ? (p & ~3) : p;
          for (int c = 1; c < 3; c++)
            if ((h.cbf[c][p] >> tr) & 1) code_coeff_of(e, h, c, ridx, log2n == 2 ? 2 : log2n - 1);
        }
      }
    }
    p += h.np >> (2 * T);
  }
  eio = e;
}

// one shared copy of the luma predictor for the rough mode decision (keeps the instruction footprint small)

// ----------------------------------------------------------------------------------------------
// rough mode decision of one luma PU (TEncSearch.cpp:2271-2295): 35 predictions + Hadamard SATD (xGetHADs,
// TComRdCost.cpp:1549-1824).  Device: a lane predicts one row of an 8x8 (4x4) tile and transforms it in registers,
// the column transform and the sum run over warp shuffles; 4 (8) tiles per pass.  r->refU / refF / dcVal hold the
// PU's reference lines.  Result: r->satd[mode], already shifted by bitDepth - 8.
// ----------------------------------------------------------------------------------------------
#ifndef FHC_DEVICE_BUILD
static uint32_t had8_ref(const int* d)
{
  int m[64], s = 0;
  for (int i = 0; i < 64; i++) m[i] = d[i];
  for (int i = 0; i < 8; i++)
    for (int k = 1; k < 8; k <<= 1)
      for (int j = 0; j < 8; j++)
        if (!(j & k)) { int a = m[i * 8 + j], bb = m[i * 8 + j + k]; m[i * 8 + j] = a + bb; m[i * 8 + j + k] = a - bb; }
  for (int j = 0; j < 8; j++)
    for (int k = 1; k < 8; k <<= 1)
      for (int i = 0; i < 8; i++)
        if (!(i & k)) { int a = m[i * 8 + j], bb = m[(i + k) * 8 + j]; m[i * 8 + j] = a + bb; m[(i + k) * 8 + j] = a - bb; }
  for (int i = 0; i < 64; i++) s += iabs(m[i]);
  return (uint32_t)((s + 2) >> 2);
}
static uint32_t had4_ref(const int* d)
{
  int m[16], s = 0;
  for (int i = 0; i < 4; i++)
  {
    int a = d[i * 4] + d[i * 4 + 3], bb = d[i * 4 + 1] + d[i * 4 + 2], cc = d[i * 4 + 1] - d[i * 4 + 2], dd = d[i * 4] - d[i * 4 + 3];
    m[i * 4] = a + bb; m[i * 4 + 1] = a - bb; m[i * 4 + 2] = dd + cc; m[i * 4 + 3] = dd - cc;
  }
  for (int j = 0; j < 4; j++)
  {
    int a = m[j] + m[12 + j], bb = m[4 + j] + m[8 + j], cc = m[4 + j] - m[8 + j], dd = m[j] - m[12 + j];
    s += iabs(a + bb) + iabs(a - bb) + iabs(dd + cc) + iabs(dd - cc);
  }
  return (uint32_t)((s + 1) >> 1);
}
#endif

FHC_D void rmd_satd(Blk* b, Role* r, int log2n, int x0, int y0)
{
  FHC_SHARED(b); FHC_SHARED(r);
  const Tables& t = b->t;
  const int n = 1 << log2n, bd = b->pp.bitDepth;
  const int16_t* src = b->src + y0 * 64 + x0;
  FHC_LANES(i, 36) r->satd[i] = 0;
  warp_sync();
  FHC_COUNT(g_emuCounters.rmdSamples += 35ull * (unsigned)(n * n));
#ifdef FHC_DEVICE_BUILD
  const int lane = FHC_LANE;
  if (n >= 8)
  {
    // the whole warp works on one 8x8 tile of one mode: lane (rr, c) holds the samples (x = c, y = rr) and (c, rr + 4)
    const int ltpr = log2n - 3, tiles = 1 << (2 * ltpr);
    const int c = lane & 7, rr = lane >> 3;
    FHC_NOUNROLL for (int mode = 0; mode < 35; mode++)
    {
      const int16_t* line = use_filtered_refs(t, 0, mode, log2n) ? r->refF : r->refU;
      const PredMode pm = pred_mode(t, mode, n, log2n, 1);
      const int dcv = r->dcVal;
      uint32_t total = 0;
      FHC_NOUNROLL for (int tile = 0; tile < tiles; tile++)
      {
        const int x = ((tile & ((1 << ltpr) - 1)) << 3) + c, y = ((tile >> ltpr) << 3) + rr;
        int a = src[y * 64 + x] - pred_at(pm, line, n, x, y, bd, dcv);
        int bb = src[(y + 4) * 64 + x] - pred_at(pm, line, n, x, y + 4, bd, dcv);
        // rows: across the 8 lanes of a row
#pragma unroll
        for (int k = 1; k < 8; k <<= 1)
        {
          const int oa = __shfl_xor_sync(0xffffffffu, a, k), ob = __shfl_xor_sync(0xffffffffu, bb, k);
          a = (c & k) ? oa - a : a + oa;
          bb = (c & k) ? ob - bb : bb + ob;
        }
        // columns: rows y and y + 4 are in this lane, the other two stages run across the lanes
        { const int t0 = a + bb, t1 = a - bb; a = t0; bb = t1; }
#pragma unroll
        for (int k = 8; k < 32; k <<= 1)
        {
          const int oa = __shfl_xor_sync(0xffffffffu, a, k), ob = __shfl_xor_sync(0xffffffffu, bb, k);
          a = (lane & k) ? oa - a : a + oa;
          bb = (lane & k) ? ob - bb : bb + ob;
        }
        int sm = iabs(a) + iabs(bb);
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, k);
        total += (uint32_t)((sm + 2) >> 2);
      }
      if (lane == 0) r->satd[mode] = total;
    }
  }
  else
  {
    // two modes per pass, one sample per lane
    const int half = lane >> 4, x = lane & 3, y = (lane >> 2) & 3;
    FHC_NOUNROLL for (int m0 = 0; m0 < 35; m0 += 2)
    {
      const int mode = imin(m0 + half, 34);
      const PredMode pm = pred_mode(t, mode, 4, 2, 1);
      int a = src[y * 64 + x] - pred_at(pm, r->refU, 4, x, y, bd, r->dcVal);
#pragma unroll
      for (int k = 1; k < 16; k <<= 1)
      {
        const int o = __shfl_xor_sync(0xffffffffu, a, k);
        a = (lane & k) ? o - a : a + o;
      }
      int sm = iabs(a);
#pragma unroll
      for (int k = 1; k < 16; k <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, k);
      if ((lane & 15) == 0 && m0 + half < 35) r->satd[mode] = (uint32_t)((sm + 1) >> 1);
    }
  }
#else
  for (int mode = 0; mode < 35; mode++)
  {
    const int16_t* line = use_filtered_refs(t, 0, mode, log2n) ? r->refF : r->refU;
    uint32_t sum = 0;
    if (n >= 8)
    {
      for (int ty = 0; ty < n; ty += 8)
        for (int tx = 0; tx < n; tx += 8)
        {
          int d[64];
          for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) d[y * 8 + x] = src[(ty + y) * 64 + tx + x] - pred_sample(t, line, n, log2n, mode, tx + x, ty + y, 1, bd, r->dcVal);
          sum += had8_ref(d);
        }
    }
    else
    {
      int d[16];
      for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) d[y * 4 + x] = src[y * 64 + x] - pred_sample(t, line, n, log2n, mode, x, y, 1, bd, r->dcVal);
      sum = had4_ref(d);
    }
    r->satd[mode] = sum;
  }
#endif
  warp_sync();
  FHC_LANES(i, 35) r->satd[i] >>= (bd - 8);
  warp_sync();
}

// candidate list of a PU: the N best RMD modes (xUpdateCandList, TEncSearch.cpp:5385: insertion with strict <, i.e. a
// stable sort by cost) plus the first 1 or 2 MPMs (TEncSearch.cpp:2297-2320).  r->cand[], r->numCand.
FHC_D void rmd_candidates(Blk* b, Role* r, int pu, int log2pu)
{
  FHC_SHARED(b); FHC_SHARED(r);
  const int isNxN = r->partSize == SIZE_NxN;
  const int zabs = r->zidx + pu * (r->np >> 2) * (isNxN ? 1 : 0);
  int mpm[3];
  const int numMpm = get_mpm(b, r->zidx, r->np, isNxN, r->ldir, zabs, mpm);
  if (FHC_LANE == 0) { r->mpm[0] = mpm[0]; r->mpm[1] = mpm[1]; r->mpm[2] = mpm[2]; }
  // xModeBitsIntra (TEncSearch.cpp:5353): luma-dir bits on CI_CURR_BEST's prev_intra_luma_pred context and bit counter
  const Sbac* cur = &b->curr[r->depth];
  const uint32_t f0 = (uint32_t)(cur->frac & 32767);
  const int st = cur->c[CTX_IPRED];
  FHC_LANES(m, 35)
  {
    const int pi = mpm_index(m, mpm);
    const uint32_t f = f0 + (uint32_t)b->t.entropyBits[st ^ (pi != -1 ? 1 : 0)] + 32768u * (uint32_t)(pi != -1 ? (pi ? 2 : 1) : 5);
    r->rmdCost[m] = dadd((double)r->satd[m], dmul((double)(f >> 15), b->pp.sp.sqrtLambda));
  }
  warp_sync();
  const int K = (log2pu <= 3) ? 8 : 3;                                 // g_aucIntraModeNumFast_UseMPM (TComRom.cpp:545)
  FHC_LANES(m, 35)
  {
    const double c = r->rmdCost[m];
    int rank = 0;
    for (int o = 0; o < 35; o++) { const double co = r->rmdCost[o]; rank += (co < c) || (co == c && o < m); }
    if (rank < K) r->cand[rank] = m;
  }
  warp_sync();
  if (FHC_LANE == 0)
  {
    int nc = K;
    for (int j = 0; j < numMpm; j++)
    {
      int included = 0;
      for (int i = 0; i < nc; i++) included |= (mpm[j] == r->cand[i]);
      if (!included) r->cand[nc++] = mpm[j];
    }
    r->numCand = nc;
  }
  warp_sync();
}

// ----------------------------------------------------------------------------------------------
// lane-group dispatch: task idx < tasks runs on lanes [idx * G, (idx + 1) * G), G = 32 / tasks; the warp rejoins at the end
// ----------------------------------------------------------------------------------------------
template <class F>
FHC_DH void for_groups(int tasks, F f)
{
#ifdef FHC_DEVICE_BUILD
  const int G = 32 / tasks, lane = FHC_LANE, idx = lane / G;
  if (idx < tasks)
  {
    Grp g;
    g.gl = lane - idx * G; g.G = G; g.idx = idx;
    g.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (idx * G));
    f(g);
  }
  __syncwarp();
#else
  for (int idx = 0; idx < tasks; idx++)
  {
    Grp g;
    g.gl = 0; g.G = 1; g.idx = idx; g.mask = 1u;
    f(g);
  }
#endif
}

FHC_DH Hyp hyp_of(const Role* r, const uint8_t* tridx, const uint8_t* cbf0, const uint8_t* cbf1, const uint8_t* cbf2, const uint8_t* ts0,
                  const uint8_t* ts1, const uint8_t* ts2, const int16_t* coef, uint16_t* nz)
{
  Hyp h;
  h.depth = r->depth; h.np = r->np; h.log2cu = r->log2cu; h.cuw = r->cuw; h.partSize = r->partSize;
  h.tridx = tridx; h.cbf[0] = cbf0; h.cbf[1] = cbf1; h.cbf[2] = cbf2; h.ts[0] = ts0; h.ts[1] = ts1; h.ts[2] = ts2;
  for (int c = 0; c < 3; c++) { h.coefc[c] = coef + yuv_off(r->cuw, c); h.nzc[c] = nz + (yuv_off(r->cuw, c) >> 4); }
  for (int i = 0; i < 4; i++) h.ldir[i] = r->ldir[i];
  h.cdir = r->cdir;
  return h;
}

// ----------------------------------------------------------------------------------------------
// second luma pass of a 2Nx2N PU (TEncSearch.cpp:2420-2488 -> xRecurIntraCodingLumaQT with bCheckFirst = false,
// :1404-1713): the best mode again, now with the residual quad-tree.  Root node = the PU; its four B nodes (half
// size) are each coded as one TU and, where the RQT depth allows, as four C leaves; the cheaper wins (strict <).
// The root's own single-TU branch repeats the first-pass evaluation of the same mode (same cost), so only the split
// branch is evaluated for PUs up to 32x32; a 64x64 PU has no single-TU branch.  Works on r->curRec / curCoef /
// cTridx / cCbf / cTs with the whole warp as one group.  Returns 1 when the split tree beats *bestCost.
// ----------------------------------------------------------------------------------------------
FHC_D int second_pass(Blk* b, Role* r, int mode, double* bestCost, uint32_t* bestDist)
{
  FHC_SHARED(b); FHC_SHARED(r);
  const Tables& t = b->t;
  const Grp w = whole_warp();
  GScr s = group_scratch(r, 0, r->yuvSize);
  const int L = r->log2cu, minTu = min_tu_log2(L, 0), LB = L - 1, nB = 1 << LB, LC = LB - 1, nC = 1 << LC;
  const int splitB = LB > minTu, partsB = r->np >> 2, partsC = partsB >> 2, cuw = r->cuw;
  Hyp h = hyp_of(r, r->cTridx, r->cCbf, r->cCbf, r->cCbf, r->cTs, r->cTs, r->cTs, r->curCoef, s.nz);
  h.ldir[0] = (uint8_t)mode;
  Sbac* rootB = r->slots + 0;
  Sbac* testB = r->slots + 1;
  const RecView view = { r->curRec, r->x, r->y, cuw };
  sb_copy_g(w, s.go, &b->curr[r->depth]);
  warp_sync();
  uint32_t splitDist = 0;
  int rootCbf = 0;
  FHC_NOUNROLL for (int i = 0; i < 4; i++)
  {
    const int pB = i * partsB, bx = (i & 1) * nB, by = (i >> 1) * nB;
    twin_align(b, (int)(r - b->role));
    if (splitB) sb_copy_g(w, rootB, s.go);
    FHC_LANES(k, partsB) { r->cTridx[pB + k] = 1; r->cTs[pB + k] = 0; }
    TuArgs a;
    a.comp = 0; a.log2n = LB; a.mode = mode; a.scanType = scan_idx_for(mode, LB, 0); a.ts = 0; a.cbfCtx = CTX_CBF + 0;
    a.srcOff = (r->y + by) * 64 + r->x + bx; a.srcStride = 64;
    a.rec = r->curRec + by * cuw + bx; a.recStride = cuw; a.coef = r->curCoef + 16 * pB;
    a.refU = r->refU; a.refF = r->refF; a.phases = TU_ALL; a.alignRole = (int)(r - b->role);
    a.dcVal = build_refs(b, w, view, opq(0), nB, r->x + bx, r->y + by, LB >= 3, r->refU, r->refF);
    int cbfS;
    const uint32_t dS = tu_eval(b, w, s, a, &cbfS);
    FHC_LANES(k, partsB) r->cCbf[pB + k] = (uint8_t)(cbfS << 1);
    nz_masks_tu(w, t, a.coef, LB, a.scanType, h.nzc[0] + pB);
    warp_sync();
    if (w.gl == 0) { Coder e = coder_open(s.go, &t); s.acc[3] = intra_bits_luma(e, b, h, r->zidx, pB, 1); coder_close(e); }
    warp_sync();
    const double costS = rd_cost(b, s.acc[3], dS);
    uint32_t distB = dS;
    if (splitB)
    {
      sb_copy_g(w, testB, s.go);
      warp_sync();
      sb_copy_g(w, s.go, rootB);
      // keep the single-TU result while the four leaves overwrite its place
      FHC_LANES(k, nB * nB) { r->altRec[k] = a.rec[(k >> LB) * cuw + (k & (nB - 1))]; r->altCoef[k] = a.coef[k]; }
      warp_sync();
      uint32_t dSp = 0;
      int cbfAny = 0;
      FHC_NOUNROLL for (int j = 0; j < 4; j++)
      {
        const int pC = pB + j * partsC, cx = bx + (j & 1) * nC, cy = by + (j >> 1) * nC;
        FHC_LANES(k, partsC) { r->cTridx[pC + k] = 2; r->cTs[pC + k] = 0; }
        TuArgs c;
        c.comp = 0; c.log2n = LC; c.mode = mode; c.scanType = scan_idx_for(mode, LC, 0); c.ts = 0; c.cbfCtx = CTX_CBF + 0;
        c.srcOff = (r->y + cy) * 64 + r->x + cx; c.srcStride = 64;
        c.rec = r->curRec + cy * cuw + cx; c.recStride = cuw; c.coef = r->curCoef + 16 * pC;
        c.refU = r->refU; c.refF = r->refF; c.phases = TU_ALL; c.alignRole = (int)(r - b->role);
        c.dcVal = build_refs(b, w, view, opq(0), nC, r->x + cx, r->y + cy, LC >= 3, r->refU, r->refF);
        int cbfC;
        dSp += tu_eval(b, w, s, c, &cbfC);
        cbfAny |= cbfC;
        FHC_LANES(k, partsC) r->cCbf[pC + k] = (uint8_t)(cbfC << 2);
        nz_masks_tu(w, t, c.coef, LC, c.scanType, h.nzc[0] + pC);
        warp_sync();
        if (w.gl == 0) { Coder e = coder_open(s.go, &t); intra_bits_luma(e, b, h, r->zidx, pC, 2); coder_close(e); }
        warp_sync();
      }
      if (cbfAny) FHC_LANES(k, partsB) r->cCbf[pB + k] |= 2;
      warp_sync();
      sb_copy_g(w, s.go, rootB);
      warp_sync();
      if (w.gl == 0) { Coder e = coder_open(s.go, &t); s.acc[3] = intra_bits_luma(e, b, h, r->zidx, pB, 1); coder_close(e); }
      warp_sync();
      const double costSp = rd_cost(b, s.acc[3], dSp);
      if (costSp < costS) distB = dSp;
      else
      {
        sb_copy_g(w, s.go, testB);
        FHC_LANES(k, partsB) { r->cTridx[pB + k] = 1; r->cCbf[pB + k] = (uint8_t)(cbfS << 1); r->cTs[pB + k] = 0; }
        FHC_LANES(k, nB * nB) { a.rec[(k >> LB) * cuw + (k & (nB - 1))] = r->altRec[k]; a.coef[k] = r->altCoef[k]; }
      }
      warp_sync();
    }
    splitDist += distB;
    rootCbf |= (r->cCbf[pB] >> 1) & 1;
  }
  if (rootCbf) FHC_LANES(k, r->np) r->cCbf[k] |= 1;
  warp_sync();
  nz_masks_range(t, w, h, 0, r->np, 1, 0);
  sb_copy_g(w, s.go, &b->curr[r->depth]);
  warp_sync();
  if (w.gl == 0) { Coder e = coder_open(s.go, &t); s.acc[3] = intra_bits_luma(e, b, h, r->zidx, 0, 0); coder_close(e); }
  warp_sync();
  const double cost = rd_cost(b, s.acc[3], splitDist);
  if (cost < *bestCost) { *bestCost = cost; *bestDist = splitDist; return 1; }
  return 0;
}

// ----------------------------------------------------------------------------------------------
// luma search of one PU (estIntraPredLumaQT, TEncSearch.cpp:2177-2557): RMD, candidate list, full RD of every
// candidate on its own lane group (4x4 PUs: every (candidate, transform_skip_flag) pair), second pass.
// ----------------------------------------------------------------------------------------------
FHC_D void search_luma_pu(Blk* b, Role* r, int pu)
{
  FHC_SHARED(b); FHC_SHARED(r);
  const Tables& t = b->t;
  const Grp w = whole_warp();
  const int isNxN = r->partSize == SIZE_NxN, depth = r->depth, cuw = r->cuw;
  const int log2pu = r->log2cu - isNxN, n = 1 << log2pu;
  const int puParts = r->np >> (2 * isNxN), p0 = pu * puParts;
  const int inX = isNxN ? (pu & 1) * n : 0, inY = isNxN ? (pu >> 1) * n : 0;         // inside the CU
  const int px = r->x + inX, py = r->y + inY;                                         // CTU-relative
  {
    const RecView view = { r->curRec, r->x, r->y, cuw };
    const int dc = build_refs(b, w, view, opq(0), n, px, py, log2pu >= 3, r->refU, r->refF);
    if (w.gl == 0) r->dcVal = dc;
    warp_sync();
  }
  const int ri = (int)(r - b->role);
  trace_role(b->pp, ri, 16 + pu * 8 + 1);
  twin_align(b, ri);
  FHC_TIC(tRmd);
  rmd_satd(b, r, log2pu, px, py);
  trace_role(b->pp, ri, 16 + pu * 8 + 2);
  rmd_candidates(b, r, pu, log2pu);
  trace_role(b->pp, ri, 16 + pu * 8 + 3);
  FHC_TOC(r, PF_RMD, tRmd);
  twin_align(b, ri);
  FHC_TIC(tL1);
  const int numCand = r->numCand, tasks = numCand * (log2pu == 2 ? 2 : 1);
  // The dispatch runs in three parts - sample stages up to the pre-quantisation, the sequential RDOQ passes, then
  // reconstruction + distortion + bit counting - with a twin alignment between them: the RDOQ passes (most of the
  // instructions of a dispatch) of the two pictures of the block then run at the same time.  The state of a candidate
  // stays in its group's scratch between the parts.
  if (log2pu <= 5)
  {
    auto args = [&](const Grp& g, const GScr& s, int phases)
    {
      const int ci = (log2pu == 2) ? (g.idx >> 1) : g.idx;
      const int mode = r->cand[ci];
      TuArgs a;
      a.comp = 0; a.log2n = log2pu; a.mode = mode; a.scanType = scan_idx_for(mode, log2pu, 0); a.ts = (log2pu == 2) ? (g.idx & 1) : 0;
      a.cbfCtx = CTX_CBF + (isNxN ? 0 : 1);
      a.srcOff = py * 64 + px; a.srcStride = 64;
      if (isNxN) { a.rec = s.rec; a.recStride = 4; a.coef = s.coef; }                  // compact 4x4 candidate block
      else { a.rec = s.rec + inY * cuw + inX; a.recStride = cuw; a.coef = s.coef + 16 * p0; }
      a.refU = r->refU; a.refF = r->refF; a.dcVal = r->dcVal;
      a.phases = phases; a.alignRole = -1;
      return a;
    };
    for_groups(tasks, [&](const Grp& g)
    {
      const GScr s = group_scratch(r, g.idx, isNxN ? 16 : r->yuvSize);
      sb_copy_g(g, s.go, &b->curr[depth]);
      gsync(g);
      int cbf;
      tu_eval(b, g, s, args(g, s, TU_FRONT), &cbf);
    });
    twin_align(b, ri);
    for_groups(tasks, [&](const Grp& g)
    {
      const GScr s = group_scratch(r, g.idx, isNxN ? 16 : r->yuvSize);
      int cbf;
      tu_eval(b, g, s, args(g, s, TU_RDOQ), &cbf);
    });
    twin_align(b, ri);
    for_groups(tasks, [&](const Grp& g)
    {
      const GScr s = group_scratch(r, g.idx, isNxN ? 16 : r->yuvSize);
      const TuArgs a = args(g, s, TU_BACK);
      uint16_t* nzTu = isNxN ? s.nz : s.nz + p0;
      int cbf;
      const uint32_t dist = tu_eval(b, g, s, a, &cbf);
      nz_masks_tu(g, t, a.coef, log2pu, a.scanType, nzTu);
      gsync(g);
      if (g.gl == 0)
      {
        double cost = FHC_MAX_COST;
        if (!(a.ts && !cbf))                                          // transform skip with cbf = 0 is not allowed (TEncSearch.cpp:1515)
        {
          Coder e = coder_open(s.go, &t);
          reset_bits(e);
          if (p0 == 0 && depth == 3) enc_bin(e, CTX_PART, isNxN ? 0 : 1);
          code_luma_dir1(e, a.mode, r->mpm);
          if (!isNxN) enc_bin(e, CTX_SUBDIV + (5 - log2pu), 0);        // split_transform_flag of the PU-sized TU (always coded for 8..32)
          enc_bin(e, a.cbfCtx, cbf);
          if (cbf) code_coeff_serial(e, a.coef, log2pu, opq(0), a.scanType, a.ts, nzTu);
          cost = rd_cost(b, num_bits(e), dist);
          coder_close(e);
        }
        r->gCost[g.idx] = cost; r->gDist[g.idx] = dist; r->gCbf[g.idx] = (uint32_t)cbf | ((uint32_t)a.ts << 8);
      }
    });
  }
  else
  {
    // 64x64 PU: four 32x32 TUs in z-order, contexts chained from TU to TU (xRecurIntraCodingLumaQT at 64: split only)
    auto args = [&](const Grp& g, const GScr& s, int k, int phases)
    {
      const int tx = (k & 1) * 32, ty = (k >> 1) * 32;
      TuArgs a;
      a.comp = 0; a.log2n = 5; a.mode = r->cand[g.idx]; a.scanType = SCAN_DIAG; a.ts = 0; a.cbfCtx = CTX_CBF + 0;
      a.srcOff = (r->y + ty) * 64 + r->x + tx; a.srcStride = 64;
      a.rec = s.rec + ty * 64 + tx; a.recStride = 64; a.coef = s.coef + 16 * 64 * k;
      a.refU = s.refU; a.refF = s.refF; a.dcVal = 0;
      a.phases = phases; a.alignRole = -1;
      return a;
    };
    FHC_LANES(k, tasks) { r->gDist[k] = 0; r->gCbf[k] = 0; }
    for_groups(tasks, [&](const Grp& g)
    {
      const GScr s = group_scratch(r, g.idx, r->yuvSize);
      sb_copy_g(g, s.go, &b->curr[depth]);
    });
    FHC_NOUNROLL for (int k = 0; k < 4; k++)
    {
      for_groups(tasks, [&](const Grp& g)
      {
        const GScr s = group_scratch(r, g.idx, r->yuvSize);
        const RecView gv = { s.rec, r->x, r->y, 64 };
        TuArgs a = args(g, s, k, TU_FRONT);
        a.dcVal = build_refs(b, g, gv, opq(0), opq(32), r->x + (k & 1) * 32, r->y + (k >> 1) * 32, opq(1), s.refU, s.refF);
        int cbf;
        tu_eval(b, g, s, a, &cbf);
      });
      twin_align(b, ri);
      for_groups(tasks, [&](const Grp& g)
      {
        const GScr s = group_scratch(r, g.idx, r->yuvSize);
        int cbf;
        tu_eval(b, g, s, args(g, s, k, TU_RDOQ), &cbf);
      });
      twin_align(b, ri);
      for_groups(tasks, [&](const Grp& g)
      {
        const GScr s = group_scratch(r, g.idx, r->yuvSize);
        const TuArgs a = args(g, s, k, TU_BACK);
        int cbf;
        const uint32_t dist = tu_eval(b, g, s, a, &cbf);
        nz_masks_tu(g, t, a.coef, 5, SCAN_DIAG, s.nz + 64 * k);
        gsync(g);
        if (g.gl == 0)
        {
          r->gDist[g.idx] += dist; r->gCbf[g.idx] |= (uint32_t)cbf << k;
          Coder e = coder_open(s.go, &t);
          reset_bits(e);
          if (k == 0) code_luma_dir1(e, a.mode, r->mpm);
          enc_bin(e, CTX_SUBDIV + 0, 0);
          enc_bin(e, CTX_CBF + 0, cbf);
          if (cbf) code_coeff_serial(e, a.coef, opq(5), opq(0), opq(SCAN_DIAG), opq(0), s.nz + 64 * k);
          coder_close(e);
        }
      });
    }
    for_groups(tasks, [&](const Grp& g)
    {
      const GScr s = group_scratch(r, g.idx, r->yuvSize);
      sb_copy_g(g, s.go, &b->curr[depth]);
      gsync(g);
      if (g.gl == 0)
      {
        const uint32_t cbfBits = r->gCbf[g.idx];
        Coder e = coder_open(s.go, &t);
        reset_bits(e);
        code_luma_dir1(e, r->cand[g.idx], r->mpm);
        for (int k = 0; k < 4; k++) { enc_bin(e, CTX_SUBDIV + 0, 0); enc_bin(e, CTX_CBF + 0, (cbfBits >> k) & 1); }
        for (int k = 0; k < 4; k++)
          if ((cbfBits >> k) & 1) code_coeff_serial(e, s.coef + 16 * 64 * k, opq(5), opq(0), opq(SCAN_DIAG), opq(0), s.nz + 64 * k);
        r->gCost[g.idx] = rd_cost(b, num_bits(e), r->gDist[g.idx]);
        coder_close(e);
      }
    });
  }
  trace_role(b->pp, ri, 16 + pu * 8 + 4);
  FHC_TOC(r, PF_LUMA1, tL1);
  // winner: lowest cost, ties to the lowest task index = HM's sequential scan with strict <
  int best = 0;
  FHC_NOUNROLL for (int i = 1; i < tasks; i++) if (r->gCost[i] < r->gCost[best]) best = i;
  double bestCost = r->gCost[best];
  uint32_t bestDist = r->gDist[best];
  const uint32_t bestCbf = r->gCbf[best];
  const int bestMode = r->cand[(log2pu == 2) ? (best >> 1) : best];
  {
    // xSetIntraResultLumaQT (TEncSearch.cpp:1716) + the TU flags of the PU
    const GScr s = group_scratch(r, best, isNxN ? 16 : r->yuvSize);
    FHC_LANES(k, n * n)
    {
      const int o = (inY + (k >> log2pu)) * cuw + inX + (k & (n - 1));
      r->bestRec[o] = isNxN ? s.rec[k] : s.rec[o];
      r->bestCoef[16 * p0 + k] = isNxN ? s.coef[k] : s.coef[16 * p0 + k];
    }
    FHC_LANES(k, puParts)
    {
      int cb;
      if (log2pu == 6) cb = (int)(((bestCbf >> (k >> 6)) & 1) << 1) | ((bestCbf & 15) ? 1 : 0);
      else cb = (int)(bestCbf & 1) << (isNxN ? 1 : 0);
      r->tridx[p0 + k] = (uint8_t)((isNxN || log2pu == 6) ? 1 : 0);
      r->cbf[0][p0 + k] = (uint8_t)cb;
      r->ts[0][p0 + k] = (uint8_t)((bestCbf >> 8) & 1);
    }
    warp_sync();
  }
  FHC_TIC(tL2);
  twin_align(b, ri);
  if (log2pu >= 3)
  {
    if (second_pass(b, r, bestMode, &bestCost, &bestDist))
    {
      FHC_LANES(k, n * n) { r->bestRec[k] = r->curRec[k]; r->bestCoef[k] = r->curCoef[k]; }
      FHC_LANES(k, r->np) { r->tridx[k] = r->cTridx[k]; r->cbf[0][k] = r->cCbf[k]; r->ts[0][k] = r->cTs[k]; }
      warp_sync();
    }
  }
  else
  {
    // the PU's reconstruction is what the following PUs of the CU predict from (TEncSearch.cpp:2506-2525)
    FHC_LANES(k, n * n)
    {
      const int o = (inY + (k >> log2pu)) * cuw + inX + (k & (n - 1));
      r->curRec[o] = r->bestRec[o];
    }
  }
  if (w.gl == 0) { r->ldir[pu] = (uint8_t)bestMode; r->dist += bestDist; }
  warp_sync();
  FHC_TOC(r, PF_LUMA2, tL2);
}

// ----------------------------------------------------------------------------------------------
// chroma search of a CU (estIntraPredChromaQT, TEncSearch.cpp:2562-2712; xRecurIntraChromaCodingQT :1915-2120):
// the five candidate modes on five lane groups, each walking the luma TU tree of the CU.
// ----------------------------------------------------------------------------------------------
FHC_D void search_chroma(Blk* b, Role* r)
{
  FHC_SHARED(b); FHC_SHARED(r);
  const Tables& t = b->t;
  const int np = r->np, cuw = r->cuw, cw = cuw >> 1, depth = r->depth;
  int modeList[5];
  allowed_chroma(r->ldir[0], modeList);
  for_groups(5, [&](const Grp& g)
  {
    const GScr s = group_scratch(r, g.idx, r->yuvSize);
    const int cdir = modeList[g.idx], predMode = cdir == DM_CHROMA_IDX ? r->ldir[0] : cdir;
    uint8_t* cbfC[2] = { s.meta, s.meta + np };
    uint8_t* tsC[2] = { s.meta + 2 * np, s.meta + 3 * np };
    FHC_GLANES(g, k, 4 * np) s.meta[k] = 0;
    sb_copy_g(g, s.go, &b->curr[depth]);
    gsync(g);
    uint32_t dist = 0;
    int maxT = 0;
    FHC_NOUNROLL for (int p = 0; p < np;)
    {
      const int T = r->tridx[p], leafParts = np >> (2 * T), log2n = r->log2cu - T;
      maxT = imax(maxT, T);
      if (log2n > 2 || (p & 3) == 0)
      {
        // chroma block of this luma leaf; below 8x8 luma the 4x4 chroma block of the parent goes with the first quadrant
        const int cl = log2n == 2 ? 2 : log2n - 1, cn = 1 << cl, nparts = log2n == 2 ? 4 : leafParts;
        const int cxin = zx(t, p) >> 1, cyin = zy(t, p) >> 1;
        int checkTS = 0;
        if (log2n == 2) checkTS = r->ts[0][p] | r->ts[0][p + 1] | r->ts[0][p + 2] | r->ts[0][p + 3];      // TransformSkipFast (TEncSearch.cpp:1943-1956)
        FHC_NOUNROLL for (int c = 1; c < 3; c++)
        {
          int16_t* plane = s.rec + yuv_off(cuw, c);
          const RecView cv = { plane, r->x >> 1, r->y >> 1, cw };
          TuArgs a;
          a.comp = c; a.log2n = cl; a.mode = predMode; a.scanType = scan_idx_for(predMode, cl, 1); a.ts = 0; a.cbfCtx = CTX_CBF + 5 + T;
          a.srcOff = comp_off64(c) + ((r->y >> 1) + cyin) * 32 + (r->x >> 1) + cxin; a.srcStride = 32;
          a.rec = plane + cyin * cw + cxin; a.recStride = cw; a.coef = s.coef + yuv_off(cuw, c) + 4 * p;
          a.refU = s.refU; a.refF = s.refU; a.phases = TU_ALL; a.alignRole = -1;
          a.dcVal = build_refs(b, g, cv, c, cn, (r->x >> 1) + cxin, (r->y >> 1) + cyin, opq(0), s.refU, s.refF);
          uint16_t* nzTu = s.nz + ((yuv_off(cuw, c) + 4 * p) >> 4);
          int cbf, tsBest = 0;
          if (!checkTS) dist += tu_eval(b, g, s, a, &cbf);
          else
          {
            // transform_skip_flag 0 / 1 of a 4x4 chroma block, compared by RD cost of the coefficient bits alone
            Sbac* tsRoot = r->slots + 2 + 2 * g.idx;
            Sbac* tsKeep = tsRoot + 1;
            int16_t* save = s.refF;                                   // chroma never uses the filtered line: 32 samples of save area
            sb_copy_g(g, tsRoot, s.go);
            int cbf0, cbf1;
            const uint32_t d0 = tu_eval(b, g, s, a, &cbf0);
            nz_masks_tu(g, t, a.coef, 2, a.scanType, nzTu);
            gsync(g);
            if (g.gl == 0)
            {
              Coder e = coder_open(s.go, &t);
              reset_bits(e);
              if (cbf0) code_coeff_serial(e, a.coef, opq(2), opq(1), a.scanType, opq(0), nzTu);
              s.acc[3] = num_bits(e);
              coder_close(e);
            }
            gsync(g);
            const double cost0 = rd_cost(b, s.acc[3], d0);
            FHC_GLANES(g, k, 16) { save[k] = a.coef[k]; save[16 + k] = a.rec[(k >> 2) * cw + (k & 3)]; }
            sb_copy_g(g, tsKeep, s.go);
            gsync(g);
            sb_copy_g(g, s.go, tsRoot);
            gsync(g);
            a.ts = 1;
            const uint32_t d1 = tu_eval(b, g, s, a, &cbf1);
            double cost1 = FHC_MAX_COST;
            if (cbf1)
            {
              nz_masks_tu(g, t, a.coef, 2, a.scanType, nzTu);
              gsync(g);
              if (g.gl == 0)
              {
                Coder e = coder_open(s.go, &t);
                reset_bits(e);
                code_coeff_serial(e, a.coef, opq(2), opq(1), a.scanType, opq(1), nzTu);
                s.acc[3] = num_bits(e);
                coder_close(e);
              }
              gsync(g);
              cost1 = rd_cost(b, s.acc[3], d1);
            }
            if (cost1 < cost0) { cbf = cbf1; tsBest = 1; dist += d1; }
            else
            {
              cbf = cbf0; dist += d0;
              FHC_GLANES(g, k, 16) { a.coef[k] = save[k]; a.rec[(k >> 2) * cw + (k & 3)] = save[16 + k]; }
              sb_copy_g(g, s.go, tsKeep);
            }
            gsync(g);
          }
          FHC_GLANES(g, k, nparts) { cbfC[c - 1][p + k] = (uint8_t)(cbf << T); tsC[c - 1][p + k] = (uint8_t)tsBest; }
          gsync(g);
        }
      }
      p += leafParts;
    }
    // cbf of the inner nodes: set when any child has it (TEncSearch.cpp:2096-2115)
    FHC_NOUNROLL for (int tr = maxT - 1; tr >= 0; tr--)
    {
      const int parts = np >> (2 * tr), q = parts >> 2;
      FHC_GLANES(g, node, 1 << (2 * tr))
      {
        const int p = node * parts;
        if (r->tridx[p] > tr)
          for (int c = 0; c < 2; c++)
          {
            const int any = ((cbfC[c][p] | cbfC[c][p + q] | cbfC[c][p + 2 * q] | cbfC[c][p + 3 * q]) >> (tr + 1)) & 1;
            if (any) for (int k = 0; k < parts; k++) cbfC[c][p + k] |= (uint8_t)(1 << tr);
          }
      }
      gsync(g);
    }
    Hyp h = hyp_of(r, r->tridx, r->cbf[0], cbfC[0], cbfC[1], r->ts[0], tsC[0], tsC[1], s.coef, s.nz);
    h.cdir = (uint8_t)cdir;
    nz_masks_range(t, g, h, 0, np, 0, 1);
    sb_copy_g(g, s.go, &b->curr[depth]);
    gsync(g);
    if (g.gl == 0)
    {
      Coder e = coder_open(s.go, &t);
      const uint32_t bits = intra_bits_chroma(e, h);
      coder_close(e);
      r->gCost[g.idx] = rd_cost(b, bits, dist); r->gDist[g.idx] = dist;
    }
  });
  int best = 0;
  FHC_NOUNROLL for (int i = 1; i < 5; i++) if (r->gCost[i] < r->gCost[best]) best = i;
  {
    const GScr s = group_scratch(r, best, r->yuvSize);
    const int c0 = yuv_off(cuw, 1), cnt = (cuw * cuw) >> 1;
    FHC_LANES(k, cnt) { r->bestRec[c0 + k] = s.rec[c0 + k]; r->bestCoef[c0 + k] = s.coef[c0 + k]; }
    FHC_LANES(k, np)
    {
      r->cbf[1][k] = s.meta[k]; r->cbf[2][k] = s.meta[np + k];
      r->ts[1][k] = s.meta[2 * np + k]; r->ts[2][k] = s.meta[3 * np + k];
    }
    if (FHC_LANE == 0) { r->cdir = (uint8_t)modeList[best]; r->dist += r->gDist[best]; }
    warp_sync();
  }
}

// The job's scratch is dead now: tell the L2 not to write those lines back (the chains' scratch is rewritten all the time;
// without this its dirty lines leave the L2 for DRAM whenever the picture / result stream pushes them out).  The host
// emulation fills the ranges with garbage instead, which proves that no job reads what an earlier job left there.
FHC_DH void role_discard(const Role* r)
{
#ifdef FHC_DEVICE_BUILD
  for (int k = 0; k < r->numDead; k++)
    for (uint32_t i = FHC_LANE; i < r->deadLines[k]; i += 32)
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(r->dead[k] + ((size_t)i << 7)) : "memory");
  __syncwarp();
#else
  for (int k = 0; k < r->numDead; k++) memset(r->dead[k], 0xA5, (size_t)r->deadLines[k] << 7);
#endif
}

// ----------------------------------------------------------------------------------------------
// one CU hypothesis = xCheckRDCostIntra (TEncCu.cpp:1528-1593): luma search, chroma search, CU syntax bits, J
// ----------------------------------------------------------------------------------------------
FHC_D void run_job(Blk* b, Role* r)
{
  FHC_SHARED(b); FHC_SHARED(r);
  const Tables& t = b->t;
  const Grp w = whole_warp();
  const int isNxN = r->partSize == SIZE_NxN, np = r->np;
  if (w.gl == 0) { r->dist = 0; r->ldir[0] = r->ldir[1] = r->ldir[2] = r->ldir[3] = DC_IDX; r->cdir = 0; }
  warp_sync();
  const int ri = (int)(r - b->role);
  trace_role(b->pp, ri, 1);
  twin_align(b, ri);
  FHC_TIC(tJob);
  FHC_CNT(r, PF_JOBS);
  for (int pu = 0; pu < (isNxN ? 4 : 1); pu++) { search_luma_pu(b, r, pu); trace_role(b->pp, ri, 2 + pu); }
  if (isNxN)
  {
    // cbf of the CU level: set when any PU has coefficients (TEncSearch.cpp:2531-2548)
    const int q = np >> 2;
    const int any = ((r->cbf[0][0] | r->cbf[0][q] | r->cbf[0][2 * q] | r->cbf[0][3 * q]) >> 1) & 1;
    warp_sync();
    if (any) FHC_LANES(k, np) r->cbf[0][k] |= 1;
    warp_sync();
  }
  twin_align(b, ri);
  FHC_TIC(tCh);
  search_chroma(b, r);
  FHC_TOC(r, PF_CHROMA, tCh);
  twin_align(b, ri);
  FHC_TIC(tSy);
  trace_role(b->pp, ri, 6);
  // bits of the whole CU in real syntax order on the CI_CURR_BEST contexts (TEncCu.cpp:1564-1588)
  GScr s = group_scratch(r, 0, r->yuvSize);
  Hyp h = hyp_of(r, r->tridx, r->cbf[0], r->cbf[1], r->cbf[2], r->ts[0], r->ts[1], r->ts[2], r->bestCoef, s.nz);
  nz_masks_range(t, w, h, 0, np, 1, 1);
  sb_copy_g(w, &r->endCtx, &b->curr[r->depth]);
  warp_sync();
  if (w.gl == 0)
  {
    Coder e = coder_open(&r->endCtx, &t);
    reset_bits(e);
    encode_cu_syntax(e, b, h, r->zidx);
    coder_close(e);
    r->bits = num_bits(e); r->bins = e.bins;
    r->cost = rd_cost(b, r->bits, r->dist);
  }
  warp_sync();
  role_discard(r);
  FHC_TOC(r, PF_SYNTAX, tSy);
  FHC_TOC(r, PF_JOB, tJob);
}

// ----------------------------------------------------------------------------------------------
// chain master (role 0): walks the CTU's quad-tree in z-order (xCompressCU, TEncCu.cpp:496-1058), starts the job of a
// depth when its region opens, joins it when the region closes, decides, and keeps the best-so-far CTU state.
// ----------------------------------------------------------------------------------------------
FHC_DH int role_of_depth(int d) { return d == 0 ? ROLE_0 : (d == 1 ? ROLE_1 : ROLE_2); }

FHC_DH void set_job(Blk* b, int ri, int z, int partSize)
{
  Role* r = &b->role[ri];
  if (FHC_LANE == 0) { r->zidx = z; r->x = zx(b->t, z); r->y = zy(b->t, z); r->partSize = partSize; b->mailbox[ri] = 1; b->jobSeq[ri]++; }
  warp_sync();
}
FHC_DH void start_job(Blk* b, int ri, int z, int partSize)
{
  set_job(b, ri, z, partSize);
#ifdef FHC_DEVICE_BUILD
  bar_arrive(bar_start(b, ri));
#endif
}
FHC_DH void join_job(Blk* b, int ri)
{
#ifdef FHC_DEVICE_BUILD
  FHC_TIC(tW);
  // the role publishes the sequence number of the job it finished; the master rarely gets here first
  const int want = b->jobSeq[ri];
  while (((volatile int*)b->doneSeq)[ri] != want) __nanosleep(40);
  __threadfence_block();
  __syncwarp();
  FHC_TOC(&b->role[ri], PF_WAIT, tW);
#else
  run_job(b, &b->role[ri]);       // host build: the job runs at the join, i.e. after its sub-tree
#endif
}

// the job's hypothesis becomes the best-so-far state of its region: copyToPic + xCopyYuv2Pic (TEncCu.cpp:1043-1045)
FHC_D void commit_job(Blk* b, const Role* r)
{
  FHC_SHARED(b); FHC_SHARED(r);
  const Tables& t = b->t;
  const int cuw = r->cuw, np = r->np, z0 = r->zidx, isNxN = r->partSize == SIZE_NxN, lg = r->log2cu;
  fhc_ctu_result* o = &b->pp.out[b->ctuAddr];
  FHC_LANES(k, np)
  {
    const int z = z0 + k;
    b->ctuDepth[z] = (uint8_t)r->depth; b->ctuPart[z] = (uint8_t)r->partSize; b->ctuPred[z] = MODE_INTRA;
    b->ctuLdir[z] = r->ldir[isNxN ? k / (np >> 2) : 0]; b->ctuCdir[z] = r->cdir; b->ctuTridx[z] = r->tridx[k];
    for (int c = 0; c < 3; c++) { b->ctuCbf[c][z] = r->cbf[c][k]; b->ctuTs[c][z] = r->ts[c][k]; }
  }
  FHC_LANES(k, cuw * cuw)
  {
    b->pic[(r->y + (k >> lg)) * 64 + r->x + (k & (cuw - 1))] = r->bestRec[k];
    o->coefY[16 * z0 + k] = r->bestCoef[k];
  }
  const int cw = cuw >> 1, c1 = cuw * cuw, c2 = (cuw * cuw * 5) >> 2;
  FHC_LANES(k, cw * cw)
  {
    const int o2 = ((r->y >> 1) + (k >> (lg - 1))) * 32 + (r->x >> 1) + (k & (cw - 1));
    b->pic[4096 + o2] = r->bestRec[c1 + k];
    b->pic[5120 + o2] = r->bestRec[c2 + k];
    o->coefCb[4 * z0 + k] = r->bestCoef[c1 + k];
    o->coefCr[4 * z0 + k] = r->bestCoef[c2 + k];
  }
  warp_sync();
  (void)t;
}

FHC_DH void set_totals(Blk* b, int d, uint32_t dist, uint32_t bits, uint32_t bins, double cost)
{
  if (FHC_LANE == 0) { b->tot[d].dist = dist; b->tot[d].bits = bits; b->tot[d].bins = bins; b->tot[d].cost = cost; }
}

template <int D>
FHC_D void region(Blk* b, int z)
{
  FHC_SHARED(b);
  const Tables& t = b->t;
  const Grp w = whole_warp();
  const int W = b->pp.width, H = b->pp.height;
  if (D == 3)
  {
    // 8x8 CU: 2Nx2N here, NxN on its own warp (both start from [3][CI_CURR_BEST], TEncCu.cpp:806-815)
    trace_at(b->pp, b->ctuAddr, z, 1);
    start_job(b, ROLE_3B, z, SIZE_NxN);
    set_job(b, ROLE_3A, z, SIZE_2Nx2N);
    run_job(b, &b->role[ROLE_3A]);
    trace_at(b->pp, b->ctuAddr, z, 2);
    join_job(b, ROLE_3B);
    trace_at(b->pp, b->ctuAddr, z, 3);
    const Role* a = &b->role[ROLE_3A];
    const Role* n = &b->role[ROLE_3B];
    const Role* win = (n->cost < a->cost) ? n : a;                    // xCheckBestMode: strict <, 2Nx2N was first
    commit_job(b, win);
    sb_copy_g(w, &b->next[3], &win->endCtx);
    set_totals(b, 3, win->dist, win->bits, win->bins, win->cost);     // no split flag at depth 3
    warp_sync();
    return;
  }
  const int wd = 64 >> D, np = 256 >> (2 * D);
  const int px = b->ctuPelX + zx(t, z), py = b->ctuPelY + zy(t, z);
  const int inside = (px + wd <= W) && (py + wd <= H);
  const int ri = role_of_depth(D);
  // CU-split hint of the CNN predictor (this repository's hook): depth-1 CUs inside the picture only.
  // FHC_HINT_SPLIT: the CU's own modes are not evaluated (TEncCu.cpp:806, :839-848); FHC_HINT_NOSPLIT: its sub-branch is
  // not searched (TEncCu.cpp:892-894).
  const int hint = (D == 1 && inside && b->pp.hints) ? b->pp.hints[b->ctuAddr * 4 + (z >> 6)] : FHC_HINT_NONE;
  const int own = inside && hint != FHC_HINT_SPLIT;
  if (own) start_job(b, ri, z, SIZE_2Nx2N);
  uint32_t aDist = 0, aBits = 0, aBins = 0;
  const int search = hint != FHC_HINT_NOSPLIT;
  FHC_NOUNROLL for (int i = 0; i < (search ? 4 : 0); i++)
  {
    const int cz = z + i * (np >> 2);
    if (b->ctuPelX + zx(t, cz) < W && b->ctuPelY + zy(t, cz) < H)
    {
      // context hand-off (TEncCu.cpp:917-924)
      sb_copy_g(w, &b->curr[D + 1], i == 0 ? &b->curr[D] : &b->next[D + 1]);
      warp_sync();
      region<(D < 3 ? D + 1 : 3)>(b, cz);
      aDist += b->tot[D + 1].dist; aBits += b->tot[D + 1].bits; aBins += b->tot[D + 1].bins;
    }
    else
    {
      // sub-CU outside the picture: its partitions keep the initialised values at this depth (TEncCu.cpp:949-953)
      FHC_LANES(k, np >> 2) b->ctuDepth[cz + k] = (uint8_t)(D + 1);
    }
  }
  warp_sync();
  // split hypothesis: contexts after the last sub-CU, plus split_cu_flag = 1 (TEncCu.cpp:956-979)
  sb_copy_g(w, &b->tmp[D], &b->next[D + 1]);
  warp_sync();
  if (inside && search)
  {
    if (w.gl == 0)
    {
      Coder e = coder_open(&b->tmp[D], &t);
      reset_bits(e);
      enc_bin(e, CTX_SPLIT + ctx_split(b, z, D), 1);
      coder_close(e);
      b->scal[0] = num_bits(e); b->scal[1] = e.bins;
    }
    warp_sync();
    aBits += b->scal[0]; aBins += b->scal[1];
  }
  const double aCost = rd_cost(b, aBits, aDist);
  int splitWins = search;
  Role* r = &b->role[ri];
  if (own)
  {
    trace_at(b->pp, b->ctuAddr, z, 4 + D);
    join_job(b, ri);
    trace_at(b->pp, b->ctuAddr, z, 8 + D);
    // the 2Nx2N hypothesis + split_cu_flag = 0 on its own contexts (TEncCu.cpp:839-848)
    if (w.gl == 0)
    {
      Coder e = coder_open(&r->endCtx, &t);
      reset_bits(e);
      enc_bin(e, CTX_SPLIT + ctx_split(b, z, D), 0);
      coder_close(e);
      r->bits += num_bits(e); r->bins += e.bins;
      r->cost = rd_cost(b, r->bits, r->dist);
    }
    warp_sync();
    splitWins = search && aCost < r->cost;                            // xCheckBestMode (TEncCu.cpp:1036, :1661)
  }
  if (splitWins)
  {
    sb_copy_g(w, &b->next[D], &b->tmp[D]);
    set_totals(b, D, aDist, aBits, aBins, aCost);
  }
  else
  {
    commit_job(b, r);
    sb_copy_g(w, &b->next[D], &r->endCtx);
    set_totals(b, D, r->dist, r->bits, r->bins, r->cost);
  }
  warp_sync();
}

// xEncodeCU on the finished CTU (TEncCu.cpp:1124-1228, called from TEncSlice.cpp:886-893): advances the chain coder
FHC_D void encode_ctu(Blk* b)
{
  FHC_SHARED(b);
  const Tables& t = b->t;
  const Grp w = whole_warp();
  const int W = b->pp.width, H = b->pp.height;
  fhc_ctu_result* o = &b->pp.out[b->ctuAddr];
  Sbac* go = &b->curr[0];
  if (w.gl == 0) go->frac &= 32767;
  warp_sync();
  FHC_NOUNROLL for (int z = 0; z < 256;)
  {
    const int d = b->ctuDepth[z], np = 256 >> (2 * d), wd = 64 >> d;
    const int x = zx(t, z), y = zy(t, z), px = b->ctuPelX + x, py = b->ctuPelY + y;
    if (px >= W || py >= H) { z += np; continue; }
    Hyp h;
    h.depth = d; h.np = np; h.log2cu = 6 - d; h.cuw = wd; h.partSize = b->ctuPart[z];
    h.tridx = b->ctuTridx + z;
    for (int c = 0; c < 3; c++) { h.cbf[c] = b->ctuCbf[c] + z; h.ts[c] = b->ctuTs[c] + z; }
    h.coefc[0] = o->coefY + 16 * z; h.coefc[1] = o->coefCb + 4 * z; h.coefc[2] = o->coefCr + 4 * z;
    h.nzc[0] = b->ctuNz + z; h.nzc[1] = b->ctuNz + 256 + (z >> 2); h.nzc[2] = b->ctuNz + 320 + (z >> 2);
    for (int i = 0; i < 4; i++) h.ldir[i] = b->ctuLdir[z + (h.partSize == SIZE_NxN ? i * (np >> 2) : 0)];
    h.cdir = b->ctuCdir[z];
    nz_masks_range(t, w, h, 0, np, 1, 1);
    warp_sync();
    if (w.gl == 0)
    {
      Coder e = coder_open(go, &t);
      // split_cu_flag of every quad-tree node that starts here (not coded for nodes crossing the picture edge)
      int dd = 0;
      if (z) { dd = 1; while (z & ((256 >> (2 * dd)) - 1)) dd++; }
      for (; dd <= d && dd < 3; dd++)
      {
        const int wn = 64 >> dd;
        if (px + wn <= W && py + wn <= H) enc_bin(e, CTX_SPLIT + ctx_split(b, z, dd), dd < d ? 1 : 0);
      }
      encode_cu_syntax(e, b, h, z);
      // finishCU (TEncCu.cpp:1066): terminating bin after the last CU of the CTU unless it is the slice's last CTU
      const int lastInCtu = ((px + wd >= W) || (x + wd == 64)) && ((py + wd >= H) || (y + wd == 64));
      if (lastInCtu && b->ctuAddr != b->pp.wc * b->pp.hc - 1) enc_trm(e, 0);
      coder_close(e);
    }
    warp_sync();
    z += np;
  }
}

FHC_D void master_ctu(Blk* b)
{
  FHC_SHARED(b);
  region<0>(b, 0);
  trace_at(b->pp, b->ctuAddr, 0, 12);
#ifdef FHC_DEVICE_BUILD
  for (int ri = 1; ri < FHC_ROLES; ri++)
  {
    if (FHC_LANE == 0) b->mailbox[ri] = 0;
    __syncwarp();
    bar_arrive(bar_start(b, ri));
  }
#endif
  FHC_TIC(tEnc);
  encode_ctu(b);
  FHC_TOC(&b->role[ROLE_3A], PF_WAIT, tEnc);            // role 0 never waits for itself: its slot holds the final CTU coding
}

#ifdef FHC_DEVICE_BUILD
FHC_D void role_loop(Blk* b, int ri)
{
  FHC_SHARED(b);
  for (;;)
  {
    trace_role(b->pp, ri, 100);
    bar_sync(bar_start(b, ri));
    trace_role(b->pp, ri, 101);
    if (!b->mailbox[ri]) break;
    run_job(b, &b->role[ri]);
    if (FHC_LANE == 0) { __threadfence_block(); ((volatile int*)b->doneSeq)[ri] = b->jobSeq[ri]; }
    __syncwarp();
  }
}
#endif

// ----------------------------------------------------------------------------------------------
// block set-up: role configuration and buffer carving.  Contexts, flags, reference lines and accumulators of every role
// live in the shared-memory arena behind Blk; the sample / coefficient scratch is in the chain's global workspace.
// ----------------------------------------------------------------------------------------------
struct Arena { unsigned char* p; size_t used; };
FHC_HD void* arena_take(Arena* a, size_t bytes, size_t align = 16)
{
  const size_t off = (a->used + align - 1) & ~(align - 1);
  a->used = off + bytes;
  return a->p ? (void*)(a->p + off) : (void*)0;
}
FHC_HD void role_config(Role* r, int ri)
{
  const int depth = ri == ROLE_0 ? 0 : (ri == ROLE_1 ? 1 : (ri == ROLE_2 ? 2 : 3));
  r->depth = depth; r->cuw = 64 >> depth; r->log2cu = 6 - depth; r->np = 256 >> (2 * depth); r->yuvSize = (r->cuw * r->cuw * 3) >> 1;
  r->maxGroups = ri == ROLE_3B ? 20 : (ri == ROLE_3A ? 10 : 5);
  r->tuCap = ri == ROLE_3B ? 16 : (r->cuw > 32 ? 1024 : r->cuw * r->cuw);
  r->refCap = depth == 0 ? 132 : (depth == 3 ? 36 : 2 * r->cuw + 4);       // depth 3: 32 samples of transform-skip save area (search_chroma)
}
// carves the buffers of role ri; sm / gm may have p == 0 to only measure sizes
FHC_HD void role_carve(Role* r, int ri, Arena* sm, Arena* gm)
{
  role_config(r, ri);
  // sample / level scratch of every role lives in the chain's global workspace (L1 / L2 resident): measured 2 % faster
  // than keeping the depth 2 / 3 roles' scratch in shared memory, and the shared memory saved becomes L1
  Arena* big = gm;
  const int G = r->maxGroups, np = r->np, yuv = r->yuvSize, cap = r->tuCap;
  r->go = (Sbac*)arena_take(sm, sizeof(Sbac) * G);
  r->gAcc = (uint32_t*)arena_take(sm, 16 * G);
  r->gCost = (double*)arena_take(sm, 8 * G);
  r->gDist = (uint32_t*)arena_take(sm, 4 * G);
  r->gCbf = (uint32_t*)arena_take(sm, 4 * G);
  r->slots = (Sbac*)arena_take(sm, sizeof(Sbac) * (ri == ROLE_3B ? 12 : 2));
  r->refU = (int16_t*)arena_take(sm, 2 * (4 * r->cuw + 4));
  r->refF = (int16_t*)arena_take(sm, 2 * (4 * r->cuw + 4));
  {
    uint8_t* m = (uint8_t*)arena_take(sm, 10 * np);
    r->tridx = m;
    for (int c = 0; c < 3; c++) { r->cbf[c] = m + (1 + c) * np; r->ts[c] = m + (4 + c) * np; }
    r->cTridx = m + 7 * np; r->cCbf = m + 8 * np; r->cTs = m + 9 * np;
  }
  // (line-aligned: dead ranges are dropped from the L2 line by line, see role_discard)
  r->numDead = 0;
  const size_t tuBytes = (size_t)4 * cap * G;
  r->tuA = (int32_t*)arena_take(big, tuBytes, 128);
  r->tuB = (int32_t*)arena_take(big, tuBytes, 128);
  const size_t cgBytes = (size_t)8 * (cap >> 4) * G;
  r->cgCost = (double*)arena_take(big, cgBytes, 128);
  r->nzMask = (uint16_t*)arena_take(big, 2 * (yuv >> 4) * G, 128);
  // candidate planes: full CU planes per group, except the NxN role's twenty 4x4 luma candidates (one 4x4 block each);
  // its five chroma candidates use the same memory with full planes
  const int planes = ri == ROLE_3B ? (5 * yuv > 16 * G ? 5 * yuv : 16 * G) : yuv * G;
  r->gRec = (int16_t*)arena_take(big, 2 * planes, 128);
  r->gCoef = (int16_t*)arena_take(big, 2 * planes, 128);
  r->gMeta = (uint8_t*)arena_take(big, 4 * np * 5, 128);
  r->gRefU = (int16_t*)arena_take(big, 2 * r->refCap * 5, 128);
  r->gRefF = (int16_t*)arena_take(big, 2 * r->refCap * 5, 128);
  r->curRec = (int16_t*)arena_take(big, 2 * yuv, 128);
  r->curCoef = (int16_t*)arena_take(big, 2 * yuv, 128);
  r->bestRec = (int16_t*)arena_take(big, 2 * yuv, 128);
  r->bestCoef = (int16_t*)arena_take(big, 2 * yuv, 128);
  const size_t altBytes = (size_t)2 * (r->cuw >> 1) * (r->cuw >> 1);
  r->altRec = (int16_t*)arena_take(big, altBytes, 128);
  r->altCoef = (int16_t*)arena_take(big, altBytes, 128);
  arena_take(big, 0, 128);                                                 // the next role starts on a line of its own
  {
    struct { void* p; size_t bytes; } d[] = { { r->tuA, tuBytes }, { r->tuB, tuBytes }, { r->cgCost, cgBytes }, { r->gRec, (size_t)2 * planes },
                                              { r->gCoef, (size_t)2 * planes }, { r->curRec, (size_t)2 * yuv }, { r->curCoef, (size_t)2 * yuv },
                                              { r->altRec, altBytes }, { r->altCoef, altBytes } };
    for (int k = 0; k < 9; k++)
      if (d[k].bytes >= 128) { r->dead[r->numDead] = (unsigned char*)d[k].p; r->deadLines[r->numDead] = (uint32_t)(d[k].bytes >> 7); r->numDead++; }
  }
}
// bytes of shared memory (Blk + arena) and of global workspace one block needs
FHC_HD void block_footprint(size_t* smemBytes, size_t* globalBytes)
{
  Arena sm = { 0, (sizeof(Blk) + 15) & ~(size_t)15 }, gm = { 0, 0 };
  Role tmp;
  for (int ri = 0; ri < FHC_ROLES; ri++) role_carve(&tmp, ri, &sm, &gm);
  *smemBytes = (sm.used + 15) & ~(size_t)15;
  *globalBytes = (gm.used + 255) & ~(size_t)255;
}

// ----------------------------------------------------------------------------------------------
// CTU load / store, executed by all threads of the block (tid / nthr: thread index / count)
// ----------------------------------------------------------------------------------------------
FHC_D void load_ctu(Blk* b, int addr, int tid, int nthr)
{
  FHC_SHARED(b);
  const PicParams& pp = b->pp;
  const int W = pp.width, H = pp.height;
  const int cx = addr % pp.wc, cy = addr / pp.wc, x0 = cx * 64, y0 = cy * 64;
  if (tid == 0) { b->ctuAddr = addr; b->ctuX = cx; b->ctuY = cy; b->ctuPelX = x0; b->ctuPelY = y0; }
  // source samples (0 outside the picture: never read by the search)
  for (int i = tid; i < 6144; i += nthr)
  {
    const int comp = i < 4096 ? 0 : (i < 5120 ? 1 : 2), sh = comp ? 1 : 0, nn = 64 >> sh;
    const int k = i - comp_off64(comp), xx = k & (nn - 1), yy = k >> (6 - sh);
    const int px = (x0 >> sh) + xx, py = (y0 >> sh) + yy, pw = W >> sh, ph = H >> sh;
    int v = 0;
    if (px < pw && py < ph)
      v = pp.sampleBytes == 1 ? (int)((const uint8_t*)pp.org[comp])[(size_t)py * pw + px] : (int)((const uint16_t*)pp.org[comp])[(size_t)py * pw + px];
    b->src[i] = (int16_t)v;
    b->pic[i] = 0;
  }
  // row above the CTU from the line buffers of the previous CTU row (x = -1 .. 127 luma, -1 .. 63 chroma)
  const int rb = (cy + 1) & 1;                     // buffer written by CTU row cy - 1
  for (int i = tid; i < 129 + 2 * 65; i += nthr)
  {
    const int comp = i < 129 ? 0 : (i < 194 ? 1 : 2), k = i - (comp == 0 ? 0 : (comp == 1 ? 129 : 194));
    const int sh = comp ? 1 : 0, px = (x0 >> sh) - 1 + k, pw = W >> sh;
    int v = 0;
    if (cy > 0 && px >= 0 && px < pw) v = ld_l2(&pp.line[rb][comp][px]);
    if (comp == 0) b->topY[k] = (int16_t)v; else b->topC[comp - 1][k] = (int16_t)v;
  }
  for (int i = tid; i < 16; i += nthr)
  {
    const int u = (x0 >> 2) + i;
    b->aboveDepth[i] = (cy > 0 && u < (W >> 2)) ? ld_l2(&pp.depthRow[rb][u]) : 0;
  }
  // picture-level CU data of this CTU: initCtu (TComDataCU.cpp:422-468)
  for (int i = tid; i < 256; i += nthr)
  {
    b->ctuDepth[i] = 0; b->ctuLdir[i] = DC_IDX; b->ctuPart[i] = SIZE_NONE; b->ctuPred[i] = MODE_NONE; b->ctuCdir[i] = 0; b->ctuTridx[i] = 0;
    for (int c = 0; c < 3; c++) { b->ctuCbf[c][i] = 0; b->ctuTs[c][i] = 0; }
  }
  fhc_ctu_result* o = &pp.out[addr];
  for (int i = tid; i < 4096; i += nthr) o->coefY[i] = 0;
  for (int i = tid; i < 1024; i += nthr) { o->coefCb[i] = 0; o->coefCr[i] = 0; }
}

FHC_D void store_ctu(Blk* b, int tid, int nthr)
{
  FHC_SHARED(b);
  const PicParams& pp = b->pp;
  const int W = pp.width, H = pp.height, x0 = b->ctuPelX, y0 = b->ctuPelY, cy = b->ctuY;
  fhc_ctu_result* o = &pp.out[b->ctuAddr];
  for (int i = tid; i < 256; i += nthr)
  {
    o->depth[i] = b->ctuDepth[i]; o->partSize[i] = b->ctuPart[i]; o->predMode[i] = b->ctuPred[i]; o->lumaDir[i] = b->ctuLdir[i];
    o->chromaDir[i] = b->ctuCdir[i]; o->trIdx[i] = b->ctuTridx[i];
    for (int c = 0; c < 3; c++) { o->cbf[c][i] = b->ctuCbf[c][i]; o->transformSkip[c][i] = b->ctuTs[c][i]; }
  }
  for (int i = tid; i < 6144; i += nthr)
  {
    const int comp = i < 4096 ? 0 : (i < 5120 ? 1 : 2), sh = comp ? 1 : 0, nn = 64 >> sh;
    const int k = i - comp_off64(comp), xx = k & (nn - 1), yy = k >> (6 - sh);
    const int px = (x0 >> sh) + xx, py = (y0 >> sh) + yy, pw = W >> sh, ph = H >> sh;
    const int in = px < pw && py < ph;
    const int16_t v = in ? b->pic[i] : (int16_t)0;
    (comp == 0 ? o->recoY : (comp == 1 ? o->recoCb : o->recoCr))[k] = v;
    // bottom row -> line buffer of this CTU row; right column -> left neighbour of the next CTU
    if (in && yy == nn - 1) pp.line[cy & 1][comp][px] = v;
  }
  if (tid == 0) { o->totalCost = b->tot[0].cost; o->totalDist = b->tot[0].dist; o->totalBits = b->tot[0].bits; o->totalBins = b->tot[0].bins; }
  for (int i = tid; i < 16; i += nthr)
  {
    const int u = (x0 >> 2) + i;
    if (u < (W >> 2)) pp.depthRow[cy & 1][u] = b->ctuDepth[b->t.r2z[15 * 16 + i]];
  }
}
// after store_ctu and a block barrier: the CTU's right column becomes the left neighbour of the next CTU
FHC_D void shift_left_neighbour(Blk* b, int tid, int nthr)
{
  FHC_SHARED(b);
  for (int i = tid; i < 128; i += nthr)
  {
    if (i < 64) b->leftY[i] = b->pic[i * 64 + 63];
    else if (i < 96) b->leftC[0][i - 64] = b->pic[4096 + (i - 64) * 32 + 31];
    else b->leftC[1][i - 96] = b->pic[5120 + (i - 96) * 32 + 31];
  }
  for (int i = tid; i < 16; i += nthr)
  {
    const int z = b->t.r2z[i * 16 + 15];
    b->leftDepth[i] = b->ctuDepth[z]; b->leftLdir[i] = b->ctuLdir[z];
  }
}

// ----------------------------------------------------------------------------------------------
// one picture on one block
// ----------------------------------------------------------------------------------------------
FHC_D void init_block(Blk* b, unsigned char* smemArena, unsigned char* globalArena)
{
  FHC_SHARED(b);
  Arena sm = { smemArena, 0 }, gm = { globalArena, 0 };
  for (int ri = 0; ri < FHC_ROLES; ri++) role_carve(&b->role[ri], ri, &sm, &gm);
}

// CTUs [firstCtu, firstCtu + count) of one picture on one chain of a block (the kernel passes one CTU row).  Device:
// called by all threads of the block at the same time (tid / nthr are the thread index / count inside the chain); the
// chains of a block advance CTU by CTU in lockstep (__syncthreads), a chain without work (pd == 0) only takes part in
// the barriers.  The contexts come from the slice start (firstCtu == 0) or from pd->state, and go back there unless
// the range ends the picture.
FHC_D void compress_picture(Blk* b, const PicDesc* pd, int firstCtu, int count, int tid, int nthr, int wpp = 0)   // wpp: the same for every chain of the launch
{
  FHC_SHARED(b);
  if (tid == 0 && pd)
  {
    PicParams& pp = b->pp;
    for (int c = 0; c < 3; c++) { pp.org[c] = pd->org[c]; pp.line[0][c] = pd->line[0][c]; pp.line[1][c] = pd->line[1][c]; }
    pp.depthRow[0] = pd->depthRow[0]; pp.depthRow[1] = pd->depthRow[1];
    pp.out = pd->out; pp.sp = pd->sp;
    pp.width = pd->width; pp.height = pd->height; pp.bitDepth = pd->bitDepth; pp.sampleBytes = pd->sampleBytes; pp.wc = pd->wc; pp.hc = pd->hc;
    pp.trace = pd->trace; pp.prof = pd->prof; pp.hints = pd->hints;
    for (int ri = 0; ri < FHC_ROLES; ri++) { for (int k = 0; k < PF_NUM; k++) b->role[ri].prof[k] = 0; b->jobSeq[ri] = 0; b->doneSeq[ri] = 0; }
  }
  const int lastOfPicture = pd && firstCtu + count == pd->wc * pd->hc;
#ifdef FHC_DEVICE_BUILD
  __syncthreads();
#ifdef FHC_PROFILE
  const long long tPic = clock64();
#endif
  RowState* const rows = pd ? pd->rows : nullptr;                              // wavefront picture: this range lies inside one CTU row
  const int wc = pd ? pd->wc : 1, row = firstCtu / wc;
  if (pd && tid < 32)
  {
    if (rows ? firstCtu == row * wc : firstCtu == 0)
    {
      sbac_reset(b->t, &b->curr[0], b->pp.sp.sliceQp);                         // slice start / wavefront row start: resetEntropy (TEncSlice.cpp:719-720, :815)
      if (rows && row > 0 && wc >= 2)                                          // contexts after the second CTU of the row above (TEncSlice.cpp:817-827); the bit carry stays 0
      {
        __syncwarp();
        while (*reinterpret_cast<volatile int32_t*>(&rows[row - 1].progress) < 2) __nanosleep(200);
        __threadfence();
        const int32_t* src = reinterpret_cast<const int32_t*>(rows[row - 1].sync.c);
        int32_t* dst = reinterpret_cast<int32_t*>(b->curr[0].c);
        for (int i = tid; i < NUM_CTX / 4; i += 32) dst[i] = ld_l2(src + i);
        for (int i = (NUM_CTX & ~3) + tid; i < NUM_CTX; i += 32) b->curr[0].c[i] = (uint8_t)ld_l2(&rows[row - 1].sync.c[i]);
      }
    }
    else
    {
      const int32_t* src = reinterpret_cast<const int32_t*>(rows ? &rows[row].ctx : &pd->state->ctx);
      int32_t* dst = reinterpret_cast<int32_t*>(&b->curr[0]);
      for (int i = tid; i < (int)(sizeof(Sbac) / 4); i += 32) dst[i] = ld_l2(src + i);
    }
  }
  for (int addr = firstCtu; addr < firstCtu + count; addr++)
  {
    if (wpp)
    {
      // wavefront lag: the row above must have finished the CTU above-right of this one (the barrier is taken by every
      // chain of the block, with or without a picture: the chains advance in lockstep)
      if (rows && row > 0 && tid == 0)
      {
        const int need = imin(addr - row * wc + 2, wc);
        while (*reinterpret_cast<volatile int32_t*>(&rows[row - 1].progress) < need) __nanosleep(200);
        __threadfence();
      }
      __syncthreads();
    }
    if (pd) load_ctu(b, addr, tid, nthr);
    __syncthreads();
    if (pd) { if (tid < 32) master_ctu(b); else role_loop(b, tid >> 5); }
    __syncthreads();
    if (pd) { store_ctu(b, tid, nthr); shift_left_neighbour(b, tid, nthr); }
    if (wpp)
    {
      const int x = addr - row * wc;
      if (rows && x == 1 && tid < 32)                                          // TEncSlice.cpp:925-928
      {
        const int32_t* src = reinterpret_cast<const int32_t*>(&b->curr[0]);
        int32_t* dst = reinterpret_cast<int32_t*>(&rows[row].sync);
        for (int i = tid; i < (int)(sizeof(Sbac) / 4); i += 32) dst[i] = src[i];
      }
      __threadfence();                   // this CTU's results, line-buffer segment and sync contexts are visible before its progress is
      __syncthreads();
      if (rows && tid == 0)
      {
        *reinterpret_cast<volatile int32_t*>(&rows[row].progress) = x + 1;
        // the picture's next row may start on any chain slot once this one is two CTUs ahead
        if (x + 1 == imin(2, wc) && row + 1 < pd->hc)
        {
          *reinterpret_cast<volatile int32_t*>(&pd->state->nextRow) = row + 1;
          __threadfence();
          *reinterpret_cast<volatile int32_t*>(&pd->state->status) = ST_READY;
        }
      }
    }
    __syncthreads();
  }
  if (pd)
  {
    if (!lastOfPicture && !rows && tid < 32)
    {
      const int32_t* src = reinterpret_cast<const int32_t*>(&b->curr[0]);
      int32_t* dst = reinterpret_cast<int32_t*>(&pd->state->ctx);
      for (int i = tid; i < (int)(sizeof(Sbac) / 4); i += 32) dst[i] = src[i];
    }
    __threadfence();                     // results, line buffers and contexts of this range are visible before the chain is handed on
  }
  __syncthreads();
#ifdef FHC_PROFILE
  if (tid == 0 && pd && b->pp.prof)
  {
    for (int ri = 0; ri < FHC_ROLES; ri++) for (int k = 0; k < PF_NUM; k++) b->pp.prof[ri * PF_NUM + k] += b->role[ri].prof[k];
    b->pp.prof[5 * PF_NUM] += clock64() - tPic;
  }
#endif
#else
  RowState* const rows = pd->rows;
  const int wc = pd->wc, row = firstCtu / wc;
  if (rows ? firstCtu == row * wc : firstCtu == 0)
  {
    sbac_reset(b->t, &b->curr[0], b->pp.sp.sliceQp);
    if (rows && row > 0 && wc >= 2)
    {
      if (rows[row - 1].progress < 2) abort();                                 // the caller broke the wavefront order
      for (int i = 0; i < NUM_CTX; i++) b->curr[0].c[i] = rows[row - 1].sync.c[i];
    }
  }
  else b->curr[0] = rows ? rows[row].ctx : pd->state->ctx;
  for (int addr = firstCtu; addr < firstCtu + count; addr++)
  {
    const int x = addr - row * wc;
    if (rows && row > 0 && rows[row - 1].progress < imin(x + 2, wc)) abort();
    load_ctu(b, addr, 0, 1);
    master_ctu(b);
    store_ctu(b, 0, 1);
    shift_left_neighbour(b, 0, 1);
    if (rows)
    {
      if (x == 1) rows[row].sync = b->curr[0];
      rows[row].progress = x + 1;
    }
  }
  if (rows) { if (firstCtu + count < (row + 1) * wc) rows[row].ctx = b->curr[0]; }
  else if (!lastOfPicture) pd->state->ctx = b->curr[0];
#endif
}

}  // namespace fhc
