/* fasthevc_cnn.h - C ABI of the CNN CU-split predictor (libfasthevc_cnn.so), SURVEY.md 8 row a20.
 *
 * BASELINE.json's north star names "the HARP/OpenCV CNN CU-split predictor" as a changed subsystem.  The reference holds
 * NO inference code, weights or encoder hook for it: only the training-set dumper
 * (Src_HARP/CShow_PredResiReco.h:127-255: 32x32 luma blocks of depth-1 CUs, replicated to three channels, labelled
 * "split further or not", matlab/.../detectAndClassify32Cu.m:57-62) and the MATLAB layer list
 * (matlab/train/TrainABasicConvolutionalNeuralNetworkForClassificationExample.m:70-91).  This library is the inference
 * side of that specification: one batched pass over every depth-1 CU of a batch of pictures, convolutions on the
 * tcgen05 tensor cores (fp16 operands, fp32 accumulators in tensor memory).  Where a maintainer would call it:
 * TEncGOP::compressGOP before TEncSlice::compressSlice (TEncGOP.cpp:1523), with the result consumed at depth 1 of
 * TEncCu::xCompressCU (TEncCu.cpp:806 / :892-894) - see INTEGRATION.md and fhc_submit_frames_hinted.
 *
 * Plain C, POD only, caller-owned host buffers, 0 = success, negative = error (codes of fasthevc_cuda.h).
 */
#ifndef FASTHEVC_CNN_H
#define FASTHEVC_CNN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef FHC_OK
#define FHC_OK               0
#define FHC_ERR_ARG         -1
#define FHC_ERR_CUDA        -2
#define FHC_ERR_STATE       -3
#endif

typedef struct fhc_cnn fhc_cnn;

/* fp32 weights on the host, in the layouts MATLAB / PyTorch export (...Example.m:70-91):
 * layer i = 0,1,2: convolution [Cout][Cin][3][3] with (Cin, Cout) = (3,16), (16,32), (32,64), bias [Cout], batch
 * normalisation gamma / beta / running mean / running variance [Cout]; FC [2][64*8*8] over features flattened
 * channel-major (c, y, x), bias [2].  mean = the zero-centre image of imageInputLayer, [32][32] (the three input
 * channels are replicas of the luma block, so their mean images are identical). */
typedef struct fhc_cnn_weights
{
  const float* mean;
  const float* conv_w[3];
  const float* conv_b[3];
  const float* bn_gamma[3];
  const float* bn_beta[3];
  const float* bn_mean[3];
  const float* bn_var[3];
  float        bn_eps;
  float        reserved;
  const float* fc_w;
  const float* fc_b;
} fhc_cnn_weights;

/* Folds the batch normalisation into the convolutions, rounds the operands to fp16, lays them out for the tensor cores
 * and uploads them.  Fails with FHC_ERR_CUDA when `device` is not an sm_100 GPU: there is no other path. */
int fhc_cnn_create(const fhc_cnn_weights* w, int device, fhc_cnn** out);
int fhc_cnn_destroy(fhc_cnn* ctx);

/* Number of depth-1 CUs (32x32) lying completely inside a width x height picture.  Their order everywhere in this
 * interface: CTUs in raster order, the (up to four) CUs of a CTU in z-order.  CUs cut by the picture edge are split by
 * HM without a decision (TEncCu.cpp:519-521) and are not predicted. */
int fhc_cnn_num_cus(int width, int height);
/* xy[2*i], xy[2*i+1] = luma position of CU i. */
int fhc_cnn_cu_positions(int width, int height, int32_t* xy);

/* One batched inference over every depth-1 CU of n pictures.  yuv420_planar: host memory, layout of
 * fhc_submit_frames (per picture Y, Cb, Cr; sampleBytes 1 or 2); only the luma planes are copied to the device.
 * Samples above 8 bits are reduced by >> (bitDepth - 8), as the reference's 8-bit training-set dump does.
 * logits: [n][numCus][2] fp32 (before the softmax); split: [n][numCus], 1 = "split further" (logit 1 > logit 0).
 * Either output may be NULL. */
int fhc_cnn_predict(fhc_cnn* ctx, const void* yuv420_planar, int sampleBytes, int bitDepth, int width, int height, int n,
                    float* logits, uint8_t* split);
/* Same on pictures already resident in device memory (e.g. the buffer fhc_device_input returns): no host->device
 * copy.  Results as above (device->host copy of the logits / decisions included). */
int fhc_cnn_predict_device(fhc_cnn* ctx, const void* d_yuv420_planar, int sampleBytes, int bitDepth, int width, int height,
                           int n, float* logits, uint8_t* split);
/* CUDA-event time of the inference kernel of the last predict call, and the number of CUs it covered. */
int fhc_cnn_last_kernel_ms(fhc_cnn* ctx, float* kernel_ms, int64_t* cus);
/* Re-runs the kernel of the last predict call `reps` times on the resident pictures (bench `value`): average
 * CUDA-event time per launch. */
int fhc_cnn_rerun_resident(fhc_cnn* ctx, int reps, float* avg_kernel_ms);
const char* fhc_cnn_last_error(void);
uint64_t fhc_cnn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
