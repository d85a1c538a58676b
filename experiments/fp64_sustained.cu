// fp64_sustained.cu -- DFMA throughput of the whole GPU over seconds of continuous load (power
// capping shows up after ~100 ms; a burst measurement does not see it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_sustained fp64_sustained.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double *out, double a, double b, int iters, int duty) {
  double x[ILP];
#pragma unroll
  for (int n = 0; n < ILP; ++n) x[n] = a + n + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int n = 0; n < ILP; ++n) x[n] = fma(x[n], b, a);
    if (duty > 0) __nanosleep(duty);
  }
  double s = 0;
#pragma unroll
  for (int n = 0; n < ILP; ++n) s += x[n];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char **argv) {
  const int duty = argc > 1 ? atoi(argv[1]) : 0;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out; cudaMalloc(&out, sizeof(double) * 1024 * p.multiProcessorCount);
  const int iters = 4000, warps = 16;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 40; ++rep) {   // each rep ~ 50-100 ms
    cudaEventRecord(e0);
    for (int l = 0; l < 20; ++l) chain<4><<<p.multiProcessorCount, 32 * warps>>>(out, 1.0, 0.999999, iters, duty);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = 20.0 * iters * 16 * 4 * 32 * warps * p.multiProcessorCount;
    printf("rep %2d: %.1f ms, %.2f T DFMA/s\n", rep, ms, n / (ms * 1e-3) / 1e12);
  }
  return 0;
}
