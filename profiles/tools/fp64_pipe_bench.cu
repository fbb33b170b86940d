// Microbenchmark: what is the FP64 ceiling of one B200 SM, and does the FP64
// tensor path (DMMA, mma.sync.m8n8k4.f64) raise it?  Evidence for the
// "tensor cores only if ncu shows compute-bound" decision of the 20-state path.
//   (a) DFMA with register operands, ILP independent chains per thread
//   (b) DMMA m8n8k4 with ILP independent accumulators per warp
// Prints FMA lanes/clk/SM (1 DFMA = 1 lane-FMA, 1 DMMA = 256 FMA per warp)
// assuming the clock stays at 1.965 GHz, for several warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe_bench fp64_pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP> __global__ void k_dfma(double* out, int iters, double x, double y)
{
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], x, y);
  double t = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) t += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int ILP> __global__ void k_dmma(double* out, int iters, double x, double y)
{
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  double a = x + threadIdx.x * 1e-9, b = y;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  double t = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) t += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <typename F> static void run(const char* name, F kernel, int threads, int ctas_per_sm, double fma_per_thread_iter,
                                      double* out)
{
  const int iters = 20000, grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kernel<<<grid, threads>>>(out, 100, 0.999, 1e-3);
  cudaEventRecord(e0);
  kernel<<<grid, threads>>>(out, iters, 0.999, 1e-3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)grid * threads * iters * fma_per_thread_iter;
  printf("%-14s %4d thr x %d cta/SM: %8.3f ms  %6.2f TFLOP/s  %6.1f FMA lanes/clk/SM @1.965GHz\n", name, threads,
         ctas_per_sm, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / (148.0 * 1.965e9));
}

int main()
{
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  for (int w : {128, 256, 512, 1024})
  {
    run("dfma ilp4", k_dfma<4>, w, 1, 4, out);
    run("dfma ilp8", k_dfma<8>, w, 1, 8, out);
  }
  for (int w : {128, 256, 512, 1024})
  {
    run("dmma ilp2", k_dmma<2>, w, 1, 2 * 256 / 32.0, out);
    run("dmma ilp4", k_dmma<4>, w, 1, 4 * 256 / 32.0, out);
    run("dmma ilp8", k_dmma<8>, w, 1, 8 * 256 / 32.0, out);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
