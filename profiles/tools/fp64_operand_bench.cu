// Microbenchmark: where should the 20x20 P-matrix live for the protein kernels?
// Each thread keeps SPT child vectors (20 doubles) in registers and evaluates
// y = M x in the reference's 4-lane FMA order; M comes from (a) shared memory
// (broadcast LDS.128) or (b) constant memory (LDCU -> uniform register operand
// of DFMA).  Prints achieved FP64 instruction rate as a fraction of
// 64 lanes/clk/SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double CM[4 * 2 * 400];
template <int SPT> __device__ __forceinline__ void mv(const double* __restrict__ m, double (&c)[SPT][20], double (&y)[SPT][20]) {
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    double a[SPT][4];
#pragma unroll
    for (int s = 0; s < SPT; ++s) a[s][0] = a[s][1] = a[s][2] = a[s][3] = 0;
#pragma unroll
    for (int j = 0; j < 20; j += 4)
#pragma unroll
      for (int s = 0; s < SPT; ++s) {
        a[s][0] = fma(m[i * 20 + j], c[s][j], a[s][0]);
        a[s][1] = fma(m[i * 20 + j + 1], c[s][j + 1], a[s][1]);
        a[s][2] = fma(m[i * 20 + j + 2], c[s][j + 2], a[s][2]);
        a[s][3] = fma(m[i * 20 + j + 3], c[s][j + 3], a[s][3]);
      }
#pragma unroll
    for (int s = 0; s < SPT; ++s) y[s][i] = (a[s][0] + a[s][1]) + (a[s][2] + a[s][3]);
  }
}
template <int SPT, int CONST> __global__ void __launch_bounds__(128) bench(const double* __restrict__ gm, double* out, int iters) {
  __shared__ __align__(16) double sm[4 * 2 * 400];
  for (int e = threadIdx.x; e < 3200; e += blockDim.x) sm[e] = gm[e];
  __syncthreads();
  double c[SPT][20], y[SPT][20];
  for (int s = 0; s < SPT; ++s) for (int j = 0; j < 20; ++j) c[s][j] = 1.0 + 1e-3 * (threadIdx.x + j + s);
  for (int it = 0; it < iters; ++it)
    for (int r = 0; r < 8; ++r) {
      const double* m = CONST ? (CM + r * 400) : (sm + r * 400);
      mv<SPT>(m, c, y);
#pragma unroll
      for (int s = 0; s < SPT; ++s)
#pragma unroll
        for (int j = 0; j < 20; ++j) c[s][j] = y[s][j] * 0.05;
    }
  double t = 0;
  for (int s = 0; s < SPT; ++s) for (int j = 0; j < 20; ++j) t += c[s][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int SPT, int CONST> void run(const char* name, const double* gm, double* out, int ctas_per_sm) {
  int iters = 200; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = 148 * ctas_per_sm;
  bench<SPT, CONST><<<grid, 128>>>(gm, out, 2);
  cudaEventRecord(a);
  bench<SPT, CONST><<<grid, 128>>>(gm, out, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double instr = (double)grid * 128 * SPT * iters * 8 * (20 * 23 + 20);
  int occ; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench<SPT, CONST>, 128, 0);
  printf("%-28s ctas/sm %d (max resident %d): %.3f ms, %.2f T fp64-instr/s, %.1f%% of 64 lanes/clk/SM @1.965GHz\n", name, ctas_per_sm, occ, ms,
         instr / ms / 1e9, 100.0 * instr / (ms * 1e-3) / (148.0 * 64 * 1.965e9));
}
int main() {
  double h[3200]; for (int i = 0; i < 3200; ++i) h[i] = 0.01 + 1e-4 * (i % 97);
  double *gm, *out; cudaMalloc(&gm, sizeof(h)); cudaMalloc(&out, 148 * 8 * 128 * 8);
  cudaMemcpy(gm, h, sizeof(h), cudaMemcpyHostToDevice); cudaMemcpyToSymbol(CM, h, sizeof(h));
  for (int c : {1, 2, 3, 4}) { run<1, 0>("smem  spt1", gm, out, c); run<1, 1>("const spt1", gm, out, c); }
  for (int c : {1, 2, 3}) { run<2, 0>("smem  spt2", gm, out, c); run<2, 1>("const spt2", gm, out, c); }
  for (int c : {1, 2}) { run<3, 0>("smem  spt3", gm, out, c); run<3, 1>("const spt3", gm, out, c); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
