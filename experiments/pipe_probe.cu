// pipe_probe.cu -- small probe kernels to be launched BETWEEN the stages of the running model
// (experiments/pipe_probe.py): throughput of DFMA / DADD / DMUL / FFMA / integer chains on random
// data while the GPU is in its sustained (power-managed) state vs after an idle gap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o libpipe_probe.so pipe_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

__device__ unsigned long long g_clk[2];

template <int OP>
__global__ void probe(double *io, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long c0 = clock64(), t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  double x[8], b = io[t], c = io[t + 1];
#pragma unroll
  for (int n = 0; n < 8; ++n) x[n] = io[t + 2 + n];
  float fx[8];
  unsigned ix[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) { fx[n] = (float)x[n]; ix[n] = (unsigned)__double2loint(x[n]); }
  const float fb = (float)b, fc = (float)c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (OP == 0) x[n] = fma(x[n], b, c);
        if (OP == 1) x[n] = x[n] + c;
        if (OP == 2) x[n] = x[n] * b;
        if (OP == 3) fx[n] = fmaf(fx[n], fb, fc);
        if (OP == 4) ix[n] = ix[n] * 1664525u + 1013904223u;
      }
  }
  double s = 0;
#pragma unroll
  for (int n = 0; n < 8; ++n) s += x[n] + fx[n] + ix[n];
  io[t] = s * 1e-300 + b;  // keeps the inputs essentially unchanged
  if (t == 0) {
    unsigned long long c1 = clock64(), t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    g_clk[0] = c1 - c0; g_clk[1] = t1 - t0;
  }
}

extern "C" int probe_launch(int op, double *io, int blocks, int threads, int iters, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (op) {
    case 0: probe<0><<<blocks, threads, 0, st>>>(io, iters); break;
    case 1: probe<1><<<blocks, threads, 0, st>>>(io, iters); break;
    case 2: probe<2><<<blocks, threads, 0, st>>>(io, iters); break;
    case 3: probe<3><<<blocks, threads, 0, st>>>(io, iters); break;
    default: probe<4><<<blocks, threads, 0, st>>>(io, iters); break;
  }
  return (int)cudaGetLastError();
}
extern "C" int probe_clock(unsigned long long out[2]) {
  return (int)cudaMemcpyFromSymbol(out, g_clk, 16);
}
