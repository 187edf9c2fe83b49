/* TEST INFRASTRUCTURE - CPU oracle for the HM 16.14 all-intra CTU mode decision.
 *
 * Plain-C restatement of the reference hot path (TEncCu::xCompressCU and
 * everything below it).  It is the checker for the
 * CUDA path; nothing in the product (fasthevc_b200/) may include, link or call
 * it.  Parity of this restatement is PINNED against the real reference built
 * into oracle/_ref/ (per-CTU dumps from TAppEncoderDump, see
 * tests/test_oracle_vs_reference.py and tests/golden/).
 */
#ifndef HM_INTRA_ORACLE_H
#define HM_INTRA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One finished CTU, exactly what TComDataCU::copyToPic leaves in the picture
 * CTU (reference: source/Lib/TLibCommon/TComDataCU.cpp:906-979) plus the
 * reconstruction of the CTU area (TEncCu::xCopyYuv2Pic, TEncCu.cpp:1729). */
typedef struct hm_ctu_out
{
  uint8_t  depth[256], partSize[256], predMode[256], lumaDir[256], chromaDir[256], trIdx[256];
  uint8_t  cbf[3][256], tskip[3][256];
  int16_t  coefY[4096], coefCb[1024], coefCr[1024];
  int16_t  recoY[4096], recoCb[1024], recoCr[1024];   /* samples outside the picture are 0 */
  double   totalCost;
  uint32_t totalDist, totalBits, totalBins;
} hm_ctu_out;

typedef struct hm_params
{
  int32_t width, height;          /* luma samples, multiples of 8 */
  int32_t bitDepth;               /* 8 or 10 (luma == chroma) */
  int32_t qp;                     /* slice QP */
  double  lambda;                 /* TComRdCost::m_dLambda */
  double  sqrtLambda;             /* TComRdCost::m_sqrtLambda */
  double  distWeight[3];          /* [0] unused (1.0), Cb, Cr: TComRdCost::m_distortionWeight */
  double  rdoqLambda[3];          /* TComTrQuant::m_lambdas */
} hm_params;

/* lambda set-up exactly as TEncSlice::calculateLambda / setUpLambda do for the
 * all-intra cfg (reference: TEncSlice.cpp:112-144, :433-527). */
void hm_params_from_qp(hm_params* p, int width, int height, int bitDepth, int qp);

/* Compress one all-intra picture: every CTU in raster order with the CABAC
 * chain of TEncSlice::compressSlice (TEncSlice.cpp:792-961).
 * org*: planar 4:2:0 samples as 16-bit; out: widthInCtu*heightInCtu records.
 * Returns 0. */
int hm_intra_compress_picture(const hm_params* p, const uint16_t* orgY, const uint16_t* orgCb,
                              const uint16_t* orgCr, hm_ctu_out* out);

/* The same with CU-split hints of the CNN predictor (the hook is THIS repository's definition, the reference
 * has none - F1): hints[ctuRsAddr * 4 + k] for the k-th (z-order) depth-1 CU of a CTU, consulted only for CUs that lie
 * completely inside the picture.  HM_HINT_NOSPLIT: the sub-branch of the CU is not searched (TEncCu.cpp:892-894);
 * HM_HINT_SPLIT: its 2Nx2N check is skipped (TEncCu.cpp:806).  NULL or all HM_HINT_NONE = hm_intra_compress_picture.
 * Pinned against the real reference carrying exactly this patch (oracle/_ref/TAppEncoderDumpHinted, build_ref.py). */
#define HM_HINT_NONE    0
#define HM_HINT_NOSPLIT 1
#define HM_HINT_SPLIT   2
int hm_intra_compress_picture_hinted(const hm_params* p, const uint16_t* orgY, const uint16_t* orgCb,
                                     const uint16_t* orgCr, hm_ctu_out* out, const uint8_t* hints);

/* General form: hints as above (or NULL); waveFrontSynchro = 1 restates --WaveFrontSynchro=1
 * (TEncSlice.cpp:808-827, :925-928: the RD coder is reset at every CTU row start and takes the contexts saved after
 * the second CTU of the row above).  Pinned against oracle/_ref/TAppEncoderDump run with that flag. */
int hm_intra_compress_picture_ex(const hm_params* p, const uint16_t* orgY, const uint16_t* orgCb,
                                 const uint16_t* orgCr, hm_ctu_out* out, const uint8_t* hints, int waveFrontSynchro);

/* Stage trace for tests (see hm_intra.c): records go to `path` until it is called again with NULL. */
void hm_intra_set_trace(const char* path);

/* --- leaf kernels exposed for unit tests --------------------------------- */
/* 35-mode prediction of an NxN block from a reference line of 4N+1 samples
 * (index 0 = bottom-most below-left, 2N = corner, 4N = right-most above-right). */
void hm_predict(int mode, int n, int isLuma, int bitDepth, const int16_t* refUnfilt, const int16_t* refFilt,
                int useFilt, int16_t* dst, int dstStride);
uint32_t hm_satd(const int16_t* org, int orgStride, const int16_t* cur, int curStride, int n, int bitDepth);
void hm_fwd_transform(const int16_t* resi, int stride, int32_t* coef, int n, int useDst, int bitDepth);
void hm_inv_transform(const int32_t* coef, int16_t* resi, int stride, int n, int useDst, int bitDepth);

#ifdef __cplusplus
}
#endif
#endif
