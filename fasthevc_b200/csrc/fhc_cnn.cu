// fhc_cnn.cu - CNN CU-split predictor: batched inference over every depth-1 CU of a picture batch,
// convolutions on the sm_100a tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in tensor memory).
//
// Network (the reference only lists it: matlab/train/TrainABasicConvolutionalNeuralNetworkForClassificationExample.m:70-91):
//   32x32 luma block (replicated x3 by the dumper, Src_HARP/CShow_PredResiReco.h:150) - zero-centre
//   conv3x3x16 - BN - ReLU - maxpool2 | conv3x3x32 - BN - ReLU - maxpool2 | conv3x3x64 - BN - ReLU | FC 2
//
// One thread block walks CUs; the whole network of a CU stays in shared / tensor memory:
//   conv1  explicit im2col in shared memory (K = 9 taps -> 16), 8 MMAs of 128x16x16, rows ordered so that the four
//          pixels of a pooling window are four neighbouring lanes of one tensor-memory quarter (pooling = 2 shuffles)
//   conv2  implicit GEMM without any copy: the pooled activations lie as [8-channel chunk][padded pixel] x 16 bytes, i.e.
//          every 8 consecutive padded pixels already form a K-major core matrix (8 rows x 16 bytes), so the A operand
//          of tap (dy,dx) is the same buffer with its start address advanced by (dy*18+dx)*16 bytes; output rows run
//          over the padded width (2 of 18 columns are thrown away).  3 row tiles x 9 taps of 128x32x16.
//   conv3  the same trick on the 10-wide padded 8x8 map: 9 taps x 2 K-steps of 128x64x16.
//   FC     fp32 on the CUDA cores from the fp32 accumulators (8 k MAC).
// BN is folded into the fp16 weights (scale) and an fp32 per-channel offset added in the epilogue.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include <cmath>
#include <cstdlib>

#include "../../include/fasthevc_cnn.h"

namespace {

constexpr int kThreads = 128;
// shared memory map (bytes)
constexpr int A1_OFF = 0;                       // conv1 im2col: [2 chunks][1024 rows] x 16 B; later the conv2 staging
constexpr int A1_CHUNK = 1024 * 16;
constexpr int ACT1_PIX = 432;                   // 18 x 18 padded map + slack read by the junk rows of the last tile
constexpr int ACT1_OFF = A1_OFF + 2 * A1_CHUNK; // [2 chunks][432] x 16 B
constexpr int ACT2_PIX = 160;                   // 10 x 10 padded map + slack
constexpr int ACT2_OFF = ACT1_OFF + 2 * ACT1_PIX * 16;
constexpr int W1_OFF = ACT2_OFF + 4 * ACT2_PIX * 16;   // [2][16] x 16 B
constexpr int W2_OFF = W1_OFF + 2 * 16 * 16;           // [9 taps][2][32] x 16 B
constexpr int W3_OFF = W2_OFF + 9 * 2 * 32 * 16;       // [9 taps][4][64] x 16 B
constexpr int T_OFF = W3_OFF + 9 * 4 * 64 * 16;        // t1[16] t2[32] t3[64] fp32
constexpr int MEAN_OFF = T_OFF + 112 * 4;              // [32][32] fp32
constexpr int INPAD_OFF = MEAN_OFF + 1024 * 4;         // [34][34] fp16, zero border
constexpr int MISC_OFF = INPAD_OFF + 2432;             // mbarrier, tensor-memory address, logits
constexpr int SMEM_BYTES = MISC_OFF + 64;
static_assert(SMEM_BYTES <= 113 * 1024, "two blocks per SM");
constexpr int STAGE2_OFF = A1_OFF;                     // conv2 epilogue staging [384 rows][32 ch] fp16 (aliases A1)
constexpr int kDebugBytes = 2 * ACT1_PIX * 16 + 4 * ACT2_PIX * 16 + 4096 * 4;
constexpr uint32_t TMEM_COLS = 256;                    // conv1: 0..127, conv2: 128..223, conv3: 0..63

struct DevParams
{
  uint4 w1[2 * 16];
  uint4 w2[9 * 2 * 32];
  uint4 w3[9 * 4 * 64];
  float t[112];
  float mean[1024];
  float fcw[2 * 4096];      // [j][pixel y*8+x][channel]
  float fcb[2];
  float pad[2];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous; SBO = distance of 8-row groups, LBO = distance of
// the two 16-byte K chunks of one MMA (cute/atom/mma_traits_sm100.hpp: "((8,n),2):((1,SBO),LBO)" in 16-byte units).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes)
{
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lboBytes >> 4) << 16) | ((uint64_t)(sboBytes >> 4) << 32) |
         (1ull << 46);
}
// kind::f16, A/B = fp16 K-major, D = fp32, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void umma_f16(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dTmem),
      "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void async_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// waits for the phase with the given parity; gives up after 2 s (error flag) instead of hanging the device
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity, int* err)
{
  unsigned long long t0 = 0;
  for (uint32_t spins = 0;; spins++)
  {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((spins & 255) == 255)
    {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (!t0) t0 = t;
      else if (t - t0 > 2000000000ull) break;
    }
  }
  atomicExch(err, 1);
  return false;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r)
{
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t relu_pack(uint32_t a, uint32_t b, float ta, float tb)
{
  __half2 h = __floats2half2_rn(fmaxf(__uint_as_float(a) + ta, 0.f), fmaxf(__uint_as_float(b) + tb, 0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b)
{
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 hmax2_u4(uint4 a, uint4 b)
{
  return make_uint4(hmax2_u32(a.x, b.x), hmax2_u32(a.y, b.y), hmax2_u32(a.z, b.z), hmax2_u32(a.w, b.w));
}

__global__ void __launch_bounds__(kThreads, 2)
fhc_cnn_split_kernel(const DevParams* __restrict__ prm, const uint8_t* __restrict__ pictures, long long pictureStrideBytes,
                     int width, int sampleBytes, int shift, const int2* __restrict__ cuXY, int cusPerPicture, int numCus,
                     float* __restrict__ logits, uint8_t* __restrict__ split, int* __restrict__ err, uint8_t* __restrict__ dbg)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  uint64_t* mbarPtr = reinterpret_cast<uint64_t*>(smem + MISC_OFF);
  uint32_t* tmemPtr = reinterpret_cast<uint32_t*>(smem + MISC_OFF + 8);
  float* logitAcc = reinterpret_cast<float*>(smem + MISC_OFF + 16);
  const uint32_t mbar = smem_u32(mbarPtr);
  const float* tS = reinterpret_cast<const float*>(smem + T_OFF);
  const float* meanS = reinterpret_cast<const float*>(smem + MEAN_OFF);
  __half* inpad = reinterpret_cast<__half*>(smem + INPAD_OFF);

  // ---- once per block: zero the padded maps, stage weights, barrier, tensor memory ----
  for (int i = tid; i < (W1_OFF - A1_OFF) / 16; i += kThreads) reinterpret_cast<uint4*>(smem + A1_OFF)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2432 / 16; i += kThreads) reinterpret_cast<uint4*>(smem + INPAD_OFF)[i] = make_uint4(0, 0, 0, 0);
  {
    const uint4* src = reinterpret_cast<const uint4*>(prm);      // w1, w2, w3, t, mean are contiguous in DevParams and in smem
    uint4* dst = reinterpret_cast<uint4*>(smem + W1_OFF);
    for (int i = tid; i < (INPAD_OFF - W1_OFF) / 16; i += kThreads) dst[i] = src[i];
  }
  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int k = 0; k < 6; k++) logitAcc[k] = 0.f;
  }
  __syncwarp();
  if (warp == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemPtr)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmemPtr;
  const uint32_t tmemLane = tmem + ((uint32_t)(32 * warp) << 16);   // this warp's quarter of the 128 lanes
  uint32_t parity = 0;
  bool alive = true;

  for (int cu = blockIdx.x; cu < numCus && alive; cu += gridDim.x)
  {
    const int pic = cu / cusPerPicture;
    const int2 xy = cuXY[cu - pic * cusPerPicture];
    // ---- 1. the 32x32 luma block, zero-centred, fp16, into the padded input map ----
    {
      const int row = tid >> 2, c8 = (tid & 3) * 8;
      const uint8_t* src = pictures + (long long)pic * pictureStrideBytes +
                           ((long long)(xy.y + row) * width + xy.x + c8) * sampleBytes;
      int px[8];
      if (sampleBytes == 1)
      {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
#pragma unroll
        for (int k = 0; k < 4; k++) { px[k] = (v.x >> (8 * k)) & 255; px[4 + k] = (v.y >> (8 * k)) & 255; }
      }
      else
      {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { px[2 * k] = min((int)((w4[k] & 0xFFFF) >> shift), 255); px[2 * k + 1] = min((int)((w4[k] >> 16) >> shift), 255); }
      }
      __half* d = inpad + (row + 1) * 34 + c8 + 1;
      const float* m = meanS + row * 32 + c8;
#pragma unroll
      for (int k = 0; k < 8; k++) d[k] = __float2half_rn((float)px[k] - m[k]);
    }
    __syncthreads();
    // ---- 2. conv1 im2col: row r = 128*tile + lane l; l = 32*quarter + 4*group + sub; pooled pixel P = 32*tile +
    //         8*quarter + group; source pixel = (2*Py + sub/2, 2*Px + sub%2) ----
#pragma unroll 2
    for (int tile = 0; tile < 8; tile++)
    {
      const int P = tile * 32 + (tid >> 2), sub = tid & 3;
      const int y = 2 * (P >> 4) + (sub >> 1), x = 2 * (P & 15) + (sub & 1);
      const __half* s = inpad + y * 34 + x;            // top-left tap in padded coordinates
      __half2 k01 = __halves2half2(s[0], s[1]), k23 = __halves2half2(s[2], s[34]), k45 = __halves2half2(s[35], s[36]),
              k67 = __halves2half2(s[68], s[69]), k89 = __halves2half2(s[70], __ushort_as_half(0));
      uint4 c0 = make_uint4(*reinterpret_cast<uint32_t*>(&k01), *reinterpret_cast<uint32_t*>(&k23),
                            *reinterpret_cast<uint32_t*>(&k45), *reinterpret_cast<uint32_t*>(&k67));
      uint4 c1 = make_uint4(*reinterpret_cast<uint32_t*>(&k89), 0, 0, 0);
      const int r = tile * 128 + tid;
      *reinterpret_cast<uint4*>(smem + A1_OFF + r * 16) = c0;
      *reinterpret_cast<uint4*>(smem + A1_OFF + A1_CHUNK + r * 16) = c1;
    }
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    if (tid == 0)
    {
      tc_fence_after();
      const uint64_t b = umma_desc(sbase + W1_OFF, 16 * 16, 128);
#pragma unroll
      for (int tile = 0; tile < 8; tile++)
        umma_f16(tmem + tile * 16, umma_desc(sbase + A1_OFF + tile * 2048, A1_CHUNK, 128), b, umma_idesc(16), 0);
      umma_commit(mbar);
    }
    alive = mbar_wait(mbar, parity, err);
    parity ^= 1;
    tc_fence_after();
    // ---- 3. conv1 epilogue: + offset, ReLU, fp16, 2x2 max over the four lanes of a group -> padded 16x16x16 map ----
    if (alive)
    {
#pragma unroll 1
      for (int tile = 0; tile < 8; tile++)
      {
        uint32_t r[16], h[8];
        tmem_ld16(tmemLane + tile * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
          h[k] = relu_pack(r[2 * k], r[2 * k + 1], tS[2 * k], tS[2 * k + 1]);
          h[k] = hmax2_u32(h[k], __shfl_xor_sync(0xffffffffu, h[k], 1));
          h[k] = hmax2_u32(h[k], __shfl_xor_sync(0xffffffffu, h[k], 2));
        }
        if ((lane & 3) == 0)
        {
          const int P = tile * 32 + warp * 8 + (lane >> 2);
          const int idx = ((P >> 4) + 1) * 18 + (P & 15) + 1;
          *reinterpret_cast<uint4*>(smem + ACT1_OFF + idx * 16) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(smem + ACT1_OFF + ACT1_PIX * 16 + idx * 16) = make_uint4(h[4], h[5], h[6], h[7]);
        }
      }
    }
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    if (dbg && cu == 0)      // diagnostics: the pooled conv1 map as the tensor cores will read it
      for (int i = tid; i < 2 * ACT1_PIX; i += kThreads) reinterpret_cast<uint4*>(dbg)[i] = reinterpret_cast<const uint4*>(smem + ACT1_OFF)[i];
    // ---- 4. conv2: implicit GEMM over the padded map, 3 row tiles x 9 taps ----
    if (tid == 0)
    {
      tc_fence_after();
#pragma unroll 1
      for (int tile = 0; tile < 3; tile++)
#pragma unroll
        for (int tap = 0; tap < 9; tap++)
        {
          const uint32_t a = sbase + ACT1_OFF + (tile * 128 + (tap / 3) * 18 + (tap % 3)) * 16;
          umma_f16(tmem + 128 + tile * 32, umma_desc(a, ACT1_PIX * 16, 128),
                   umma_desc(sbase + W2_OFF + tap * (2 * 32 * 16), 32 * 16, 128), umma_idesc(32), tap > 0);
        }
      umma_commit(mbar);
    }
    alive = alive && mbar_wait(mbar, parity, err);
    parity ^= 1;
    tc_fence_after();
    // ---- 5. conv2 epilogue: + offset, ReLU, fp16 -> staging [row][32 ch]; then 2x2 max -> padded 8x8x32 map ----
    if (alive)
    {
#pragma unroll 1
      for (int tile = 0; tile < 3; tile++)
      {
        const int p = tile * 128 + tid;
#pragma unroll
        for (int half = 0; half < 2; half++)
        {
          uint32_t r[16];
          tmem_ld16(tmemLane + 128 + tile * 32 + half * 16, r);
          tmem_ld_wait();
          uint32_t h[8];
#pragma unroll
          for (int k = 0; k < 8; k++) h[k] = relu_pack(r[2 * k], r[2 * k + 1], tS[16 + half * 16 + 2 * k], tS[16 + half * 16 + 2 * k + 1]);
          uint4* d = reinterpret_cast<uint4*>(smem + STAGE2_OFF + p * 64 + half * 32);
          d[0] = make_uint4(h[0], h[1], h[2], h[3]);
          d[1] = make_uint4(h[4], h[5], h[6], h[7]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
#pragma unroll
    for (int it = tid; it < 256; it += kThreads)
    {
      const int P = it >> 2, ch = it & 3;
      const int p00 = (2 * (P >> 3)) * 18 + 2 * (P & 7);
      const uint8_t* s = smem + STAGE2_OFF + ch * 16;
      uint4 v = hmax2_u4(hmax2_u4(*reinterpret_cast<const uint4*>(s + p00 * 64), *reinterpret_cast<const uint4*>(s + (p00 + 1) * 64)),
                         hmax2_u4(*reinterpret_cast<const uint4*>(s + (p00 + 18) * 64), *reinterpret_cast<const uint4*>(s + (p00 + 19) * 64)));
      *reinterpret_cast<uint4*>(smem + ACT2_OFF + ch * (ACT2_PIX * 16) + (((P >> 3) + 1) * 10 + (P & 7) + 1) * 16) = v;
    }
    async_proxy_fence();
    __syncthreads();
    if (dbg && cu == 0)
      for (int i = tid; i < 4 * ACT2_PIX; i += kThreads)
        reinterpret_cast<uint4*>(dbg + 2 * ACT1_PIX * 16)[i] = reinterpret_cast<const uint4*>(smem + ACT2_OFF)[i];
    // ---- 6. conv3: implicit GEMM over the 10-wide padded map, 9 taps x 2 K-steps ----
    if (tid == 0)
    {
      tc_fence_after();
#pragma unroll
      for (int tap = 0; tap < 9; tap++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++)
        {
          const uint32_t a = sbase + ACT2_OFF + ks * 2 * (ACT2_PIX * 16) + ((tap / 3) * 10 + (tap % 3)) * 16;
          umma_f16(tmem, umma_desc(a, ACT2_PIX * 16, 128),
                   umma_desc(sbase + W3_OFF + tap * (4 * 64 * 16) + ks * (2 * 64 * 16), 64 * 16, 128), umma_idesc(64),
                   (tap | ks) > 0);
        }
      umma_commit(mbar);
    }
    alive = alive && mbar_wait(mbar, parity, err);
    parity ^= 1;
    tc_fence_after();
    // ---- 7. conv3 epilogue + FC (fp32): row p = y*10 + x of the padded-width output ----
    if (alive && warp < 3)
    {
      const int p = tid, y = p / 10, x = p - 10 * y;
      const bool valid = p < 80 && x < 8;
      const float* fw = prm->fcw + (valid ? (y * 8 + x) * 64 : 0);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
      for (int q = 0; q < 4; q++)
      {
        uint32_t r[16];
        tmem_ld16(tmemLane + q * 16, r);
        tmem_ld_wait();
        if (valid)
        {
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++)
          {
            const float4 a = __ldg(reinterpret_cast<const float4*>(fw + q * 16 + k4 * 4));
            const float4 b = __ldg(reinterpret_cast<const float4*>(fw + 4096 + q * 16 + k4 * 4));
            const float* t3 = tS + 48 + q * 16 + k4 * 4;
            const float f0 = fmaxf(__uint_as_float(r[4 * k4]) + t3[0], 0.f), f1 = fmaxf(__uint_as_float(r[4 * k4 + 1]) + t3[1], 0.f),
                        f2 = fmaxf(__uint_as_float(r[4 * k4 + 2]) + t3[2], 0.f), f3 = fmaxf(__uint_as_float(r[4 * k4 + 3]) + t3[3], 0.f);
            if (dbg && cu == 0)
              *reinterpret_cast<float4*>(dbg + 2 * ACT1_PIX * 16 + 4 * ACT2_PIX * 16 + ((y * 8 + x) * 64 + q * 16 + k4 * 4) * 4) = make_float4(f0, f1, f2, f3);
            s0 += f0 * a.x + f1 * a.y + f2 * a.z + f3 * a.w;
            s1 += f0 * b.x + f1 * b.y + f2 * b.z + f3 * b.w;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (lane == 0)      // one partial sum per warp, added in a fixed order below: the logits are reproducible bit for bit
      {
        logitAcc[2 * warp] = s0;
        logitAcc[2 * warp + 1] = s1;
      }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0)
    {
      const float l0 = ((logitAcc[0] + logitAcc[2]) + logitAcc[4]) + prm->fcb[0], l1 = ((logitAcc[1] + logitAcc[3]) + logitAcc[5]) + prm->fcb[1];
      logits[2 * (long long)cu] = l0;
      logits[2 * (long long)cu + 1] = l1;
      split[cu] = l1 > l0 ? 1 : 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------ host side ------------------------------------------------
thread_local std::string g_err;
uint64_t g_launches = 0;

#define CK(call)                                                                                                   \
  do                                                                                                               \
  {                                                                                                                \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess)                                                                                         \
    {                                                                                                              \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                                                  \
      return FHC_ERR_CUDA;                                                                                         \
    }                                                                                                              \
  } while (0)

void cu_positions(int width, int height, std::vector<int32_t>& xy)
{
  xy.clear();
  for (int cy = 0; cy < height; cy += 64)
    for (int cx = 0; cx < width; cx += 64)
      for (int k = 0; k < 4; k++)
      {
        const int x = cx + 32 * (k & 1), y = cy + 32 * (k >> 1);
        if (x + 32 <= width && y + 32 <= height) { xy.push_back(x); xy.push_back(y); }
      }
}

// [tap][chunk][n][8] fp16: the K-major no-swizzle operand of one 3x3 convolution with cin input channels (cin = 1: the
// nine taps are the K dimension of a single "tap").
void pack_conv(const std::vector<float>& w /* [cout][cin][9], scale folded */, int cout, int cin, __half* dst)
{
  if (cin == 1)
  {
    for (int c = 0; c < 2; c++)
      for (int n = 0; n < cout; n++)
        for (int j = 0; j < 8; j++)
        {
          const int k = c * 8 + j;
          dst[(c * cout + n) * 8 + j] = __float2half_rn(k < 9 ? w[n * 9 + k] : 0.f);
        }
    return;
  }
  const int chunks = cin / 8;
  for (int tap = 0; tap < 9; tap++)
    for (int c = 0; c < chunks; c++)
      for (int n = 0; n < cout; n++)
        for (int j = 0; j < 8; j++)
          dst[((tap * chunks + c) * cout + n) * 8 + j] = __float2half_rn(w[(n * cin + c * 8 + j) * 9 + tap]);
}

}  // namespace

struct fhc_cnn
{
  int device = 0;
  int numSms = 0;
  DevParams* dPrm = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint8_t* dPics = nullptr;
  size_t picsBytes = 0;
  int2* dXY = nullptr;
  int xyW = 0, xyH = 0, cusPerPic = 0;
  float* dLogits = nullptr;
  uint8_t* dSplit = nullptr;
  int* dErr = nullptr;
  uint8_t* dDbg = nullptr;     // FHC_CNN_DEBUG=1: intermediate maps of CU 0
  size_t outCus = 0;
  // last launch (for rerun)
  const uint8_t* lastPics = nullptr;
  long long lastStride = 0;
  int lastW = 0, lastSb = 0, lastShift = 0, lastCus = 0;
  float lastMs = 0.f;
};

extern "C" {

const char* fhc_cnn_last_error(void) { return g_err.c_str(); }

/* Diagnostics (FHC_CNN_DEBUG=1 at create time): the intermediate maps of CU 0 of the last launch - the pooled conv1 map
 * [2][432][8] fp16, the pooled conv2 map [4][160][8] fp16, the conv3 features [64 pixels][64] fp32. */
int fhc_cnn_debug_dump(fhc_cnn* c, void* out, uint64_t bytes)
{
  if (!c || !c->dDbg || !out || bytes < (uint64_t)kDebugBytes) { g_err = "debug buffer not enabled (FHC_CNN_DEBUG) or too small"; return FHC_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpy(out, c->dDbg, kDebugBytes, cudaMemcpyDeviceToHost));
  return kDebugBytes;
}
uint64_t fhc_cnn_launch_count(void) { return g_launches; }

int fhc_cnn_num_cus(int width, int height)
{
  if (width <= 0 || height <= 0) return FHC_ERR_ARG;
  return ((width / 32) * (height / 32));
}

int fhc_cnn_cu_positions(int width, int height, int32_t* xy)
{
  if (width <= 0 || height <= 0 || !xy) return FHC_ERR_ARG;
  std::vector<int32_t> v;
  cu_positions(width, height, v);
  memcpy(xy, v.data(), v.size() * sizeof(int32_t));
  return (int)(v.size() / 2);
}

int fhc_cnn_destroy(fhc_cnn* c)
{
  if (!c) return FHC_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->dPrm);
  cudaFree(c->dPics);
  cudaFree(c->dXY);
  cudaFree(c->dLogits);
  cudaFree(c->dSplit);
  cudaFree(c->dErr);
  cudaFree(c->dDbg);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return FHC_OK;
}

int fhc_cnn_create(const fhc_cnn_weights* w, int device, fhc_cnn** out)
{
  if (!w || !out || !w->mean || !w->fc_w || !w->fc_b) { g_err = "null argument"; return FHC_ERR_ARG; }
  for (int i = 0; i < 3; i++)
    if (!w->conv_w[i] || !w->conv_b[i] || !w->bn_gamma[i] || !w->bn_beta[i] || !w->bn_mean[i] || !w->bn_var[i]) { g_err = "null weight pointer"; return FHC_ERR_ARG; }
  *out = nullptr;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { g_err = "the split predictor needs an sm_100 device (tcgen05); there is no other path"; return FHC_ERR_CUDA; }

  // fold BN: y = conv(x, w*s) + t, s = gamma / sqrt(var + eps), t = beta + s*(b - mean)
  static const int CIN[3] = {3, 16, 32}, COUT[3] = {16, 32, 64};
  std::vector<DevParams> hp(1);
  DevParams& P = hp[0];
  memset(&P, 0, sizeof(P));
  int tOff = 0;
  for (int i = 0; i < 3; i++)
  {
    const int cin = CIN[i], cout = COUT[i], kin = (i == 0) ? 1 : cin;
    std::vector<float> f((size_t)cout * kin * 9);
    for (int n = 0; n < cout; n++)
    {
      const float s = w->bn_gamma[i][n] / sqrtf(w->bn_var[i][n] + w->bn_eps);
      P.t[tOff + n] = w->bn_beta[i][n] + s * (w->conv_b[i][n] - w->bn_mean[i][n]);
      for (int k = 0; k < 9; k++)
        if (i == 0)
        {
          float a = 0.f;   // the three input channels are replicas of the luma block
          for (int c = 0; c < 3; c++) a += w->conv_w[0][(n * 3 + c) * 9 + k] * s;
          f[n * 9 + k] = a;
        }
        else
          for (int c = 0; c < cin; c++) f[(n * cin + c) * 9 + k] = w->conv_w[i][(n * cin + c) * 9 + k] * s;
    }
    __half* dst = reinterpret_cast<__half*>(i == 0 ? P.w1 : i == 1 ? P.w2 : P.w3);
    pack_conv(f, cout, kin, dst);
    tOff += cout;
  }
  memcpy(P.mean, w->mean, 1024 * sizeof(float));
  for (int j = 0; j < 2; j++)
    for (int c = 0; c < 64; c++)
      for (int p = 0; p < 64; p++) P.fcw[j * 4096 + p * 64 + c] = w->fc_w[j * 4096 + c * 64 + p];
  P.fcb[0] = w->fc_b[0];
  P.fcb[1] = w->fc_b[1];

  fhc_cnn* c = new fhc_cnn();
  c->device = device;
  c->numSms = prop.multiProcessorCount;
  *out = c;   // on failure below the caller destroys what exists
  CK(cudaMalloc(&c->dPrm, sizeof(DevParams)));
  CK(cudaMemcpy(c->dPrm, &P, sizeof(DevParams), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->dErr, sizeof(int)));
  CK(cudaMemset(c->dErr, 0, sizeof(int)));
  if (getenv("FHC_CNN_DEBUG"))
  {
    CK(cudaMalloc(&c->dDbg, kDebugBytes));
    CK(cudaMemset(c->dDbg, 0, kDebugBytes));
  }
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  CK(cudaFuncSetAttribute(fhc_cnn_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return FHC_OK;
}

static int launch(fhc_cnn* c, const uint8_t* dPics, long long stride, int width, int sampleBytes, int shift, int numCus)
{
  const int grid = numCus < 2 * c->numSms ? numCus : 2 * c->numSms;
  CK(cudaEventRecord(c->ev0, c->stream));
  fhc_cnn_split_kernel<<<grid, kThreads, SMEM_BYTES, c->stream>>>(c->dPrm, dPics, stride, width, sampleBytes, shift, c->dXY,
                                                                   c->cusPerPic, numCus, c->dLogits, c->dSplit, c->dErr, c->dDbg);
  g_launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev1, c->stream));
  return FHC_OK;
}

static int predict_common(fhc_cnn* c, const void* yuv, bool onDevice, int sampleBytes, int bitDepth, int width, int height, int n,
                          float* logits, uint8_t* split)
{
  if (!c || !yuv || n <= 0 || width < 32 || height < 32 || (width & 7) || (height & 7) || (sampleBytes != 1 && sampleBytes != 2) ||
      bitDepth < 8 || bitDepth > 16 || (bitDepth > 8 && sampleBytes == 1))
  {
    g_err = "bad argument";
    return FHC_ERR_ARG;
  }
  CK(cudaSetDevice(c->device));
  if (c->xyW != width || c->xyH != height)
  {
    std::vector<int32_t> xy;
    cu_positions(width, height, xy);
    cudaFree(c->dXY);
    c->dXY = nullptr;
    CK(cudaMalloc(&c->dXY, xy.size() * sizeof(int32_t) + 16));
    CK(cudaMemcpy(c->dXY, xy.data(), xy.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    c->xyW = width;
    c->xyH = height;
    c->cusPerPic = (int)(xy.size() / 2);
  }
  const long long numCus = (long long)c->cusPerPic * n;
  if (numCus <= 0 || numCus > 0x7fffffffLL) { g_err = "no depth-1 CU inside the picture"; return FHC_ERR_ARG; }
  if ((size_t)numCus > c->outCus)
  {
    cudaFree(c->dLogits);
    cudaFree(c->dSplit);
    c->dLogits = nullptr;
    c->dSplit = nullptr;
    c->outCus = 0;
    CK(cudaMalloc(&c->dLogits, numCus * 2 * sizeof(float)));
    CK(cudaMalloc(&c->dSplit, numCus));
    c->outCus = (size_t)numCus;
  }
  const long long stride = (long long)width * height * 3 / 2 * sampleBytes;
  const uint8_t* dPics = static_cast<const uint8_t*>(yuv);
  if (!onDevice)
  {
    const size_t need = (size_t)stride * n;
    if (need > c->picsBytes)
    {
      cudaFree(c->dPics);
      c->dPics = nullptr;
      c->picsBytes = 0;
      CK(cudaMalloc(&c->dPics, need));
      c->picsBytes = need;
    }
    // luma planes only: n rows of width*height*sampleBytes bytes, pitch = one whole picture
    CK(cudaMemcpy2DAsync(c->dPics, (size_t)stride, yuv, (size_t)stride, (size_t)width * height * sampleBytes, (size_t)n,
                         cudaMemcpyHostToDevice, c->stream));
    dPics = c->dPics;
  }
  int rc = launch(c, dPics, stride, width, sampleBytes, bitDepth - 8, (int)numCus);
  if (rc != FHC_OK) return rc;
  int herr = 0;
  if (logits) CK(cudaMemcpyAsync(logits, c->dLogits, numCus * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (split) CK(cudaMemcpyAsync(split, c->dSplit, numCus, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&herr, c->dErr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaEventElapsedTime(&c->lastMs, c->ev0, c->ev1));
  c->lastPics = dPics;
  c->lastStride = stride;
  c->lastW = width;
  c->lastSb = sampleBytes;
  c->lastShift = bitDepth - 8;
  c->lastCus = (int)numCus;
  if (herr) { g_err = "tensor-core pipeline timed out (mbarrier never completed)"; return FHC_ERR_CUDA; }
  return FHC_OK;
}

int fhc_cnn_predict(fhc_cnn* c, const void* yuv, int sampleBytes, int bitDepth, int width, int height, int n, float* logits,
                    uint8_t* split)
{
  return predict_common(c, yuv, false, sampleBytes, bitDepth, width, height, n, logits, split);
}

int fhc_cnn_predict_device(fhc_cnn* c, const void* dYuv, int sampleBytes, int bitDepth, int width, int height, int n,
                           float* logits, uint8_t* split)
{
  return predict_common(c, dYuv, true, sampleBytes, bitDepth, width, height, n, logits, split);
}

int fhc_cnn_last_kernel_ms(fhc_cnn* c, float* ms, int64_t* cus)
{
  if (!c || !c->lastCus) { g_err = "no predict call yet"; return FHC_ERR_STATE; }
  if (ms) *ms = c->lastMs;
  if (cus) *cus = c->lastCus;
  return FHC_OK;
}

int fhc_cnn_rerun_resident(fhc_cnn* c, int reps, float* avgMs)
{
  if (!c || !c->lastCus || reps <= 0) { g_err = "no predict call yet"; return FHC_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  double total = 0;
  for (int i = 0; i < reps; i++)
  {
    int rc = launch(c, c->lastPics, c->lastStride, c->lastW, c->lastSb, c->lastShift, c->lastCus);
    if (rc != FHC_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    total += ms;
  }
  if (avgMs) *avgMs = (float)(total / reps);
  return FHC_OK;
}

}  // extern "C"
