// slab_l2.cu -- feasibility probe for L2-resident slab pipelining of the fused RK stage.
// Three memory-traffic-equivalent kernels (A: s-step, B: column scan, C: momentum step) are
// run slab by slab (rows [ja, jb) of every level) so that the hand-off arrays (s_pre, mtg_new)
// live in slab-sized ring buffers and the re-read inputs (s_now, u, v) may still be in L2.
// Prints the time of one "stage" per (slab rows, streams) against the monolithic order.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct G { int nx, ny, nz; long long s1, s2; };

// A: out_s = f(s_now, s_int(5 rows), u, v); ring_s likewise.  thread per (i, j, k)
__global__ void __launch_bounds__(256) kA(G g, int ja, int jb, int ring_rows, const double* __restrict__ s_now,
   const double* __restrict__ s_int, const double* __restrict__ u, const double* __restrict__ v,
   double* __restrict__ s_out, double* __restrict__ ring_s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = ja + blockIdx.y * blockDim.y + threadIdx.y;
  int k = blockIdx.z;
  if (i >= g.nx || j >= jb) return;
  long long o = i + j * g.s1 + k * g.s2;
  int jm = max(j - 3, 0), jp = min(j + 3, g.ny - 1);
  double a = s_int[o] + s_int[i + jm * g.s1 + k * g.s2] + s_int[i + jp * g.s1 + k * g.s2];
  double r = s_now[o] - 0.1 * (a * u[o] + v[o]);
  s_out[o] = r;
  ring_s[i + (long long)(j % ring_rows) * g.s1 + (long long)k * g.s1 * ring_rows] = r;
}
// B: column scan over the ring: reads ring_s (all k), writes ring_m
__global__ void __launch_bounds__(128) kB(G g, int ja, int jb, int ring_rows, const double* __restrict__ ring_s,
   double* __restrict__ ring_m) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = ja + blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= jb) return;
  long long base = i + (long long)(j % ring_rows) * g.s1, ps = (long long)g.s1 * ring_rows;
  double p = 100.0;
  double e[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) e[k] = ring_s[base + k * ps];
#pragma unroll
  for (int k = 0; k < 64; ++k) { p = p + 0.3 * e[k]; e[k] = exp2(0.28 * log2(p)); }
  double m = 0.0;
#pragma unroll
  for (int k = 63; k >= 0; --k) { m = m + 2.0 * e[k]; ring_m[base + k * ps] = m; }
}
// C: momentum-like: 10 reads (5 possibly in L2), 4 writes
__global__ void __launch_bounds__(256) kC(G g, int ja, int jb, int ring_rows, const double* __restrict__ s_now,
   const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ ring_s,
   const double* __restrict__ ring_m, const double* __restrict__ su_now, const double* __restrict__ su_int,
   const double* __restrict__ sv_now, const double* __restrict__ sv_int, const double* __restrict__ mtg_now,
   double* __restrict__ su, double* __restrict__ sv, double* __restrict__ un, double* __restrict__ vn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = ja + blockIdx.y * blockDim.y + threadIdx.y;
  int k = blockIdx.z;
  if (i >= g.nx || j >= jb) return;
  long long o = i + j * g.s1 + k * g.s2;
  long long r = i + (long long)(j % ring_rows) * g.s1 + (long long)k * g.s1 * ring_rows;
  double sp = ring_s[r], mw = ring_m[r];
  double a = su_now[o] - 0.1 * (su_int[o] * u[o] + s_now[o] * mtg_now[o] + sp * mw);
  double b = sv_now[o] - 0.1 * (sv_int[o] * v[o] + s_now[o] * mtg_now[o] + sp * mw);
  su[o] = a; sv[o] = b; un[o] = a / sp; vn[o] = b / sp;
}

int main(int argc, char** argv) {
  G g{1024, 1024, 64, 0, 0};
  g.s1 = 1040; g.s2 = g.s1 * 1025;
  size_t n = (size_t)g.s2 * 65;
  const int NF = 14;
  std::vector<double*> f(NF);
  for (auto& p : f) { CK(cudaMalloc(&p, n * 8)); CK(cudaMemset(p, 0, n * 8)); }
  // s_now 0, s_int 1, u 2, v 3, s_out 4, su_now 5, su_int 6, sv_now 7, sv_int 8, mtg_now 9, su 10, sv 11, un 12, vn 13
  size_t ring_max = (size_t)g.s1 * 1025 * 65;
  double *ring_s, *ring_m;
  CK(cudaMalloc(&ring_s, ring_max * 8)); CK(cudaMalloc(&ring_m, ring_max * 8));
  CK(cudaMemset(ring_s, 0, ring_max * 8)); CK(cudaMemset(ring_m, 0, ring_max * 8));
  // fill s fields with 1.0 to avoid NaNs
  std::vector<double> h(n, 1.0);
  for (int x : {0, 1, 2, 3, 5, 6, 7, 8, 9}) CK(cudaMemcpy(f[x], h.data(), n * 8, cudaMemcpyHostToDevice));
  const int NS = 4;
  cudaStream_t st[NS];
  for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  cudaEvent_t fork, join[NS];
  CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  for (auto& j : join) CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));

  auto run_stage = [&](int rows, int nstreams) {
    // ring holds rows*nstreams*... rows: each concurrent slab needs its own ring rows -> ring_rows = rows * nstreams
    int ring_rows = rows >= g.ny ? g.ny : rows * nstreams;
    int slab = 0;
    CK(cudaEventRecord(fork, st[0]));
    for (int s = 1; s < nstreams; ++s) CK(cudaStreamWaitEvent(st[s], fork, 0));
    for (int ja = 0; ja < g.ny; ja += rows, ++slab) {
      int jb = min(ja + rows, g.ny);
      cudaStream_t s = st[slab % nstreams];
      dim3 b(64, 4), ga((g.nx + 63) / 64, (jb - ja + 3) / 4, g.nz);
      kA<<<ga, b, 0, s>>>(g, ja, jb, ring_rows, f[0], f[1], f[2], f[3], f[4], ring_s);
      dim3 bb(32, 4), gb((g.nx + 31) / 32, (jb - ja + 3) / 4, 1);
      kB<<<gb, bb, 0, s>>>(g, ja, jb, ring_rows, ring_s, ring_m);
      kC<<<ga, b, 0, s>>>(g, ja, jb, ring_rows, f[0], f[2], f[3], ring_s, ring_m, f[5], f[6], f[7], f[8], f[9],
                          f[10], f[11], f[12], f[13]);
    }
    for (int s = 1; s < nstreams; ++s) { CK(cudaEventRecord(join[s], st[s])); CK(cudaStreamWaitEvent(st[0], join[s], 0)); }
  };

  printf("rows streams ms_per_stage  (ideal 14 words = %.3f ms at 6.56 TB/s; 20 words = %.3f ms)\n",
         14.0 * g.nx * g.ny * g.nz * 8 / 6.56e12 * 1e3, 20.0 * g.nx * g.ny * g.nz * 8 / 6.56e12 * 1e3);
  int rows_list[] = {1024, 128, 64, 32, 16, 8};
  for (int rows : rows_list) for (int ns : {1, 2, 4}) {
    if (rows == 1024 && ns > 1) continue;
    for (int w = 0; w < 2; ++w) run_stage(rows, ns);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, st[0]));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) run_stage(rows, ns);
    CK(cudaEventRecord(e1, st[0]));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%5d %3d %9.3f\n", rows, ns, ms / reps);
    fflush(stdout);
  }
  // graph variant for the best-looking candidates
  for (int rows : {32, 16}) for (int ns : {2, 4}) {
    cudaGraph_t graph; cudaGraphExec_t exec;
    CK(cudaStreamBeginCapture(st[0], cudaStreamCaptureModeGlobal));
    run_stage(rows, ns);
    CK(cudaStreamEndCapture(st[0], &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    for (int w = 0; w < 2; ++w) CK(cudaGraphLaunch(exec, st[0]));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, st[0]));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(exec, st[0]));
    CK(cudaEventRecord(e1, st[0]));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("graph %5d %3d %9.3f\n", rows, ns, ms / reps);
  }
  return 0;
}
