/* fasthevc_cuda.h - C ABI of the B200 intra mode-decision engine (libfasthevc_cuda.so).
 *
 * This is the `source/Lib/TLibEncoderCUDA` boundary of BASELINE.json's north
 * star: the unchanged HM/FastHEVC host (TEncTop/TEncGOP/TEncSlice, cfg parsing,
 * bitstream writer) reaches the GPU only through these entry points.  The
 * reference has no plugin/FFI interface for this path (everything is direct
 * C++ member calls wired in TEncCu::init, source/Lib/TLibEncoder/TEncCu.cpp:227-243),
 * so each entry point names the reference code it stands in for.
 *
 * Plain C, POD only, caller-owned host buffers, 0 = success, negative = error.
 * One host thread per context.  See INTEGRATION.md for the TEncCu-side stub.
 */
#ifndef FASTHEVC_CUDA_H
#define FASTHEVC_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHC_OK               0
#define FHC_ERR_ARG         -1   /* bad argument / unsupported configuration */
#define FHC_ERR_CUDA        -2   /* CUDA runtime error (no device, launch failure, ...) */
#define FHC_ERR_STATE       -3   /* call sequence error (e.g. fetch before submit) */

typedef struct fhc_ctx fhc_ctx;

/* Sequence-level configuration: the subset of TEncCfg/TComSPS/TComPPS the hot
 * path reads (reference: TEncTop::xInitSPS/xInitPPS, cfg/encoder_intra_main.cfg).
 * Only the tool set of encoder_intra_main[10].cfg is implemented; anything
 * else is rejected by fhc_create with FHC_ERR_ARG. */
typedef struct fhc_seq_params
{
  int32_t width, height;            /* luma samples; multiples of 8 (MinCU) */
  int32_t bitDepth;                 /* InternalBitDepth: 8 (Main) or 10 (Main10) */
  int32_t chromaFormat;             /* must be 1 (4:2:0) */
  int32_t ctuSize, maxCuDepth;      /* must be 64 and 4 (MaxCUWidth, MaxPartitionDepth) */
  int32_t tuLog2Max, tuLog2Min;     /* must be 5 and 2 */
  int32_t tuMaxDepthIntra;          /* must be 3 */
  int32_t useRdoq, useRdoqTs;       /* must be 1, 1 */
  int32_t useTransformSkip, useTransformSkipFast;   /* must be 1, 1 */
  int32_t signHideFlag;             /* must be 1 */
  int32_t strongIntraSmoothing;     /* must be 1 */
  int32_t fastUdiUseMpm;            /* must be 1 */
  int32_t maxFramesInFlight;        /* device buffers are sized for this many pictures per submit */
  int32_t waveFrontSynchro;         /* 0, or 1 = --WaveFrontSynchro=1 (entropy_coding_sync_enabled_flag): the RD coder is
                                     * reset at every CTU row start and synchronised with the row above
                                     * (TEncSlice.cpp:808-823, :925-928), so the CTU rows of a picture are searched
                                     * concurrently with a two-CTU lag.  Results equal the reference run with the same flag. */
} fhc_seq_params;

/* Per-picture parameters computed by the UNCHANGED host code
 * (TEncSlice::initEncSlice / setUpLambda, TEncSlice.cpp:112-144, :372):
 * the doubles are passed through bit-for-bit, the device never recomputes them. */
typedef struct fhc_slice_params
{
  int32_t sliceQp;
  int32_t reserved;
  double  lambda;                   /* TComRdCost::m_dLambda */
  double  sqrtLambda;               /* TComRdCost::m_sqrtLambda */
  double  chromaDistWeight[2];      /* TComRdCost::m_distortionWeight[Cb,Cr] */
  double  rdoqLambda[3];            /* TComTrQuant::m_lambdas[Y,Cb,Cr] */
} fhc_slice_params;

/* One finished CTU = what TEncCu::compressCtu leaves behind
 * (TComDataCU::copyToPic, TComDataCU.cpp:906-979; TEncCu::xCopyYuv2Pic, TEncCu.cpp:1729).
 * Arrays are indexed by z-scan 4x4 partition; coefficients use HM's per-TU
 * raster blocks at offset 16*absPartIdx (luma) / 4*absPartIdx (chroma). */
typedef struct fhc_ctu_result
{
  uint8_t  depth[256], partSize[256], predMode[256], lumaDir[256], chromaDir[256], trIdx[256];
  uint8_t  cbf[3][256], transformSkip[3][256];
  int16_t  coefY[4096], coefCb[1024], coefCr[1024];
  int16_t  recoY[4096], recoCb[1024], recoCr[1024];   /* pre-deblocking reconstruction; 0 outside the picture */
  double   totalCost;
  uint32_t totalDist, totalBits, totalBins;
} fhc_ctu_result;

/* TEncCu::create + TEncSearch::init equivalent: allocates device state. */
int fhc_create(const fhc_seq_params* seq, int device, fhc_ctx** out);
int fhc_destroy(fhc_ctx* ctx);

/* Hands n consecutive pictures (POC first_poc .. first_poc+n-1) to the device and starts
 * their mode decision.  yuv420_planar: n pictures, each Y then Cb then Cr, sampleBytes 1 or 2
 * (host memory).  Replaces the per-CTU calls of TEncSlice::compressSlice -> compressCtu
 * (TEncSlice.cpp:879) for those pictures.  Asynchronous; two submits may be outstanding (the second
 * one's host->device copy overlaps the first one's kernel), a third waits for the oldest to finish,
 * whose results are then no longer fetchable.  A page-locked yuv buffer must stay untouched until a
 * fetch of one of its pictures has returned (pageable memory is staged before the call returns).
 * The pictures of a submit finish in no particular order: the device works on the pictures with the most work left
 * first (per_frame[i].sliceQp is its cost estimate), so that a submit with more pictures than fhc_pictures_in_flight
 * finishes as a whole as early as possible - its pictures then complete close together, towards the end. */
int fhc_submit_frames(fhc_ctx* ctx, int first_poc, int n, const void* yuv420_planar, int sampleBytes,
                      const fhc_slice_params* per_frame);

/* fhc_submit_frames with CU-split hints from the CNN predictor (include/fasthevc_cnn.h; BASELINE configs[4], SURVEY 8
 * a20 / 8c).  The reference has no such hook; this is the one this repository defines, a two-condition patch
 * of TEncCu::xCompressCU shown in INTEGRATION.md: for a depth-1 CU (32x32) lying completely inside the picture,
 *   FHC_HINT_NOSPLIT  the sub-branch is not searched (TEncCu.cpp:892-894 bSubBranch = false),
 *   FHC_HINT_SPLIT    the CU's own modes are not evaluated (TEncCu.cpp:806 xCheckRDCostIntra and :839-848),
 *   FHC_HINT_NONE     HM's full search.
 * hints: host memory, [n][numCtus][4] bytes, the four depth-1 CUs of a CTU in z-order; entries of CUs cut by the
 * picture edge are ignored.  hints == NULL is fhc_submit_frames.  Results are identical to the reference encoder
 * carrying that patch (oracle/_ref/TAppEncoderDumpHinted), not to the unpatched one. */
#define FHC_HINT_NONE    0
#define FHC_HINT_NOSPLIT 1
#define FHC_HINT_SPLIT   2
int fhc_submit_frames_hinted(fhc_ctx* ctx, int first_poc, int n, const void* yuv420_planar, int sampleBytes,
                             const fhc_slice_params* per_frame, const uint8_t* hints);

/* First stage of the per-picture tail on the device: TComLoopFilter::loopFilterPic
 * (TComLoopFilter.cpp:130, called from TEncGOP.cpp:1619) for pictures [first_poc, first_poc + n) of one submit, from the CTU
 * records in device memory.  Waits for those pictures, filters them (all vertical edges, then all horizontal ones) and
 * writes the deblocked pictures to `out` (host memory) in the layout of fhc_submit_frames (Y, Cb, Cr planes per picture,
 * sampleBytes each).  Equals the reference's reconstruction with --SAO=0 sample for sample. */
int fhc_deblock_pictures(fhc_ctx* ctx, int first_poc, int n, void* yuv420_planar_out, int sampleBytes);
/* Second stage: the SAO statistics of the deblocked pictures (TEncSampleAdaptiveOffset::getStatistics,
 * TEncSampleAdaptiveOffset.cpp:287, the first step of SAOProcess :253) computed on the device in the same call, from the
 * deblocked planes and the submitted originals.  sao_stats: host memory, [n][numCtus][3 components] records (or NULL:
 * fhc_deblock_pictures).  Per record the five types of HM's SAOStatData in compact form: edge offset 0 / 90 / 135 / 45
 * degrees with classes 0..4 (HM's diff[2 + edgeType]), band offset with bands 0..31; diff = sum of (original -
 * deblocked), count = samples.  The RD decision that consumes them (decideBlkParams :782) stays on the host. */
typedef struct fhc_sao_stats
{
  int32_t eoDiff[4][5], eoCount[4][5];
  int32_t boDiff[32], boCount[32];
} fhc_sao_stats;
int fhc_deblock_pictures_stats(fhc_ctx* ctx, int first_poc, int n, void* yuv420_planar_out, int sampleBytes, fhc_sao_stats* sao_stats);
/* CUDA-event time of the four deblocking kernels of the last call. */
int fhc_last_deblock_ms(fhc_ctx* ctx, float* ms);

/* Options.  FHC_OPT_PIPELINED_SUBMITS = 1: the host promises to keep a second submit queued behind the running one
 * (submit k+1 before fetching the results of submit k).  The scheduler then optimises the throughput across submits
 * instead of the completion time of one: no picture runs alone in its block, the blocks of a kernel leave as soon as the
 * pictures still unfinished fit into the others, and the next submit's kernel (own stream, own scratch) starts on the
 * freed SMs.  Results are unaffected. */
#define FHC_OPT_PIPELINED_SUBMITS 1
int fhc_configure(fhc_ctx* ctx, int option, int value);
/* Device time across several submits (their kernels may overlap): fhc_span_begin arms a CUDA event that the next submit
 * records in front of its kernel; fhc_span_ms waits for every kernel launched so far and returns the time from that
 * event to the end of the last one. */
int fhc_span_begin(fhc_ctx* ctx);
int fhc_span_ms(fhc_ctx* ctx, float* ms);

/* Device address of the input planes of submitted picture `poc` (layout as submitted; the pictures of one submit follow
 * each other, *bytes = bytes from there to the end of that submit).  Lets the CNN predictor run on the resident pictures
 * (fhc_cnn_predict_device) without a second host->device copy.  Valid until the submit slot is reused. */
int fhc_device_picture(fhc_ctx* ctx, int poc, void** dptr, uint64_t* bytes);

/* Blocks until picture `poc` is finished (only that picture: the device sets a per-picture flag; other pictures of
 * the submit may still be running), then copies CTU `ctuRsAddr` into *out.
 * This is the body of the replaced TEncCu::xCompressCU(depth 0) (TEncCu.cpp:496). */
int fhc_fetch_ctu(fhc_ctx* ctx, int poc, int ctuRsAddr, fhc_ctu_result* out);

/* Bulk variant: all CTUs of a picture in raster order (out has numCtus entries). */
int fhc_fetch_picture(fhc_ctx* ctx, int poc, fhc_ctu_result* out);
/* 1 when picture `poc` is finished on the device, 0 when not yet, negative = error.  Never blocks. */
int fhc_poll_picture(fhc_ctx* ctx, int poc);
/* Like fhc_fetch_picture, but only queues the device->host copy (out should be page-locked) on the
 * library's copy stream, where it overlaps the kernel still working on other pictures;
 * fhc_fetch_sync waits for all queued copies. */
int fhc_fetch_picture_async(fhc_ctx* ctx, int poc, fhc_ctu_result* out);
int fhc_fetch_sync(fhc_ctx* ctx);

/* Number of CTUs per picture (TComPicSym::getNumberOfCtusInFrame). */
int fhc_num_ctus(const fhc_ctx* ctx);

/* Lambda derivation for callers without an HM host (tests, bench): identical arithmetic to
 * TEncSlice::calculateLambda/setUpLambda for the all-intra cfg, evaluated on the host. */
int fhc_slice_params_from_qp(int qp, fhc_slice_params* out);

/* --- device-resident interface (bench `value`: inputs already in HBM) ------------------ */
/* Device pointer to the input planes of submit slot i (layout as fhc_submit_frames). */
int fhc_device_input(fhc_ctx* ctx, void** dptr, uint64_t* bytes);
/* Runs the decision kernel on n pictures whose samples are already in the device input
 * buffer; returns after the kernel finished; *kernel_ms = CUDA-event time of the launch. */
int fhc_run_resident(fhc_ctx* ctx, int n, int sampleBytes, const fhc_slice_params* per_frame, float* kernel_ms);
/* CUDA-event time of the kernel of the last submit (waits for it to finish). */
int fhc_last_kernel_ms(fhc_ctx* ctx, float* kernel_ms);
/* Number of pictures the persistent kernel encodes concurrently on `device` (SMs x chain slots per SM; a picture is a
 * sequential chain whose CTU rows are scheduled over the slots): submits of at least this many pictures keep the whole
 * device busy.  Negative = error. */
int fhc_pictures_in_flight(int device);
/* Page-locked host buffers for submit/fetch (optional; any host memory works). */
int fhc_host_alloc(void** p, uint64_t bytes);
int fhc_host_free(void* p);
/* Last CUDA error string for diagnostics. */
const char* fhc_last_error(void);
/* Number of kernels launched by this library so far (bench `gpu_launches`). */
uint64_t fhc_launch_count(void);
/* Measured int32 issue roof of `device`: independent IMAD / IADD3 / LOP3 chains, one third each, at full
 * occupancy; *gops = thread-level integer operations per second / 1e9, *ms = duration of the timed launch (may be NULL). */
int fhc_measure_int_issue(int device, double* gops, double* ms);

#ifdef __cplusplus
}
#endif
#endif
