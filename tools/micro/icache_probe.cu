// Instruction-cache probe (sm_100a): W warps per SM each loop over their own straight-line body of S KB (K distinct
// bodies, warp w runs body w % K).  Prints cycles per instruction versus the total code footprint K x S.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe icache_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY, int ID>
__device__ __noinline__ int body(int a, int iters)
{
  int b = a + ID, c = a ^ ID, d = a * 3 + ID;
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int u = 0; u < BODY; u++)
    {
      asm volatile("mad.lo.s32 %0, %0, %3, %1;\n\tmad.lo.s32 %1, %1, %3, %2;\n\tmad.lo.s32 %2, %2, %3, %0;\n\tadd.s32 %0, %0, %1;"
                   : "+r"(b), "+r"(c), "+r"(d) : "r"(a));
    }
  }
  return b ^ c ^ d;
}

template <int BODY>
__global__ void probe(int* out, long long* cyc, int iters, int K)
{
  const int w = threadIdx.x >> 5;
  int r = 0;
  const long long t0 = clock64();
  switch (w % K)
  {
  case 0: r = body<BODY, 0>(threadIdx.x, iters); break;
  case 1: r = body<BODY, 1>(threadIdx.x, iters); break;
  case 2: r = body<BODY, 2>(threadIdx.x, iters); break;
  case 3: r = body<BODY, 3>(threadIdx.x, iters); break;
  case 4: r = body<BODY, 4>(threadIdx.x, iters); break;
  case 5: r = body<BODY, 5>(threadIdx.x, iters); break;
  case 6: r = body<BODY, 6>(threadIdx.x, iters); break;
  case 7: r = body<BODY, 7>(threadIdx.x, iters); break;
  case 8: r = body<BODY, 8>(threadIdx.x, iters); break;
  default: r = body<BODY, 9>(threadIdx.x, iters); break;
  }
  const long long t1 = clock64();
  if (r == 0x7fffffff) out[0] = r;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + w] = t1 - t0;
}

template <int BODY>
void run(int warps, int K, int blocks)
{
  int* dOut; long long* dCyc;
  cudaMalloc(&dOut, 4); cudaMalloc(&dCyc, sizeof(long long) * blocks * warps);
  const int instrPerIter = BODY * 4;
  const int iters = 40000000 / (instrPerIter * 50) + 50;
  for (int rep = 0; rep < 2; rep++) { probe<BODY><<<blocks, warps * 32>>>(dOut, dCyc, iters, K); cudaDeviceSynchronize(); }
  long long h[4096];
  cudaMemcpy(h, dCyc, sizeof(long long) * blocks * warps, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks * warps; i++) avg += (double)h[i];
  avg /= blocks * warps;
  printf("body %5.1f KB x %2d distinct = %6.1f KB, %2d warps/SM: %.2f cycles per instruction per warp\n", instrPerIter * 16 / 1024.0, K,
         K * instrPerIter * 16 / 1024.0, warps, avg / ((double)iters * instrPerIter));
  cudaFree(dOut); cudaFree(dCyc);
}

#include <cstdlib>
int main(int argc, char** argv)
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount;
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  if (mode == 0)
  {
    for (int warps : {1, 5, 10})
      for (int K : {1, 5, 10})
      {
        if (K > warps) continue;
        run<32>(warps, K, blocks);     //  2 KB
        run<96>(warps, K, blocks);     //  6 KB
        run<192>(warps, K, blocks);    // 12 KB
        run<384>(warps, K, blocks);    // 24 KB
        run<768>(warps, K, blocks);    // 48 KB
      }
  }
  else
  {
    // small bodies, few streams: is it capacity or the number of distinct streams?
    for (int warps : {2, 4, 8, 16})
      for (int K : {1, 2, 4, 8})
      {
        if (K > warps) continue;
        run<4>(warps, K, blocks);      // 0.25 KB
        run<16>(warps, K, blocks);     // 1 KB
        run<96>(warps, K, blocks);     // 6 KB
      }
  }
  return 0;
}
