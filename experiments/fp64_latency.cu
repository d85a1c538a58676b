// fp64_latency.cu -- dependent-issue latency and per-SM throughput of DFMA / DADD / DMUL on the
// device at hand (sizing of the instruction-level parallelism the fp64 kernels need).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double *out, long long *cyc, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int n = 0; n < ILP; ++n) x[n] = a + n + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int n = 0; n < ILP; ++n) x[n] = fma(x[n], b, a);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int n = 0; n < ILP; ++n) s += x[n];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, int sms, double *out, long long *cyc) {
  const int iters = 2000;
  // one block per SM with `warps_per_sm` warps
  chain<ILP><<<sms, 32 * warps_per_sm>>>(out, cyc, 1.0, 0.999999, iters);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chain<ILP><<<sms, 32 * warps_per_sm>>>(out, cyc, 1.0, 0.999999, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double n_inst = (double)iters * 16 * ILP;  // per thread
  printf("ILP %d warps/SM %2d: %.2f cycles per dependent DFMA step (all chains), %.1f thread-DFMA/clk/SM, %.2f T DFMA/s\n",
         ILP, warps_per_sm, (double)c / (iters * 16), n_inst * 32 * warps_per_sm / (double)c,
         n_inst * 32.0 * warps_per_sm * sms / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out; long long *cyc;
  cudaMalloc(&out, sizeof(double) * 1024 * p.multiProcessorCount);
  cudaMalloc(&cyc, 8);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  for (int w : {1, 4, 8, 16, 32}) {
    run<1>(w, p.multiProcessorCount, out, cyc);
    run<2>(w, p.multiProcessorCount, out, cyc);
    run<4>(w, p.multiProcessorCount, out, cyc);
    run<8>(w, p.multiProcessorCount, out, cyc);
  }
  return 0;
}
