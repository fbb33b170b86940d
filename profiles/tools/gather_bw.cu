// Gather bandwidth microbenchmark: what HBM delivers when 128-byte blocks (one DNA CLV entry at 4 rates) are
// fetched through an index array instead of streamed - the access pattern of the site-repeat CLV kernels
// (class -> site -> child class -> CLV block).  One thread per (block, 32-byte quarter), U blocks in flight per
// thread, persistent grid; every launch reads `n` blocks of a `pool`-block buffer and writes `n` blocks
// sequentially.  Patterns:
//   seq       idx[i] = i                      (streaming reference point)
//   random    idx[i] = uniform random
//   repeats   idx[i] = next unseen block with probability `fresh`, else a uniformly random earlier one
//             (class ids in order of first appearance with back-references: what repeat identifiers look like)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int U, bool INDIRECT2>
__global__ void __launch_bounds__(128, 3)
k_gather(const double4 * __restrict__ src, const unsigned int * __restrict__ idx, const unsigned int * __restrict__ idx2,
         double4 * __restrict__ dst, unsigned int n)
{
  const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int q = tid & 3;
  const unsigned int b0 = tid >> 2;
  const unsigned int pass = (gridDim.x * blockDim.x) >> 2;
  for (unsigned int base = 0; base < n; base += pass * U)
  {
    double4 v[U];
    unsigned int b[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      b[u] = base + u * pass + b0;
      if (b[u] < n)
      {
        unsigned int i = idx[b[u]];
        if (INDIRECT2) i = idx2[i];
        v[u] = src[(size_t)i * 4 + q];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (b[u] < n)
      {
        v[u].x = v[u].x * 1.5 + v[u].y;
        dst[(size_t)b[u] * 4 + q] = v[u];
      }
  }
}

static double run(int mode, unsigned int n, const double4 * src, const unsigned int * idx, const unsigned int * idx2,
                  double4 * dst, int sms)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 3;
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep)
  {
    cudaEventRecord(e0);
    if (mode == 0) k_gather<4, false><<<grid, 128>>>(src, idx, idx2, dst, n);
    else if (mode == 1) k_gather<4, true><<<grid, 128>>>(src, idx, idx2, dst, n);
    else k_gather<8, true><<<grid, 128>>>(src, idx, idx2, dst, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char ** argv)
{
  const unsigned int pool = argc > 1 ? atoi(argv[1]) : (16u << 20); // blocks of 128 B: 2 GB
  const unsigned int n = argc > 2 ? atoi(argv[2]) : (8u << 20);     // blocks read per launch: 1 GB
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double4 * src, * dst;
  unsigned int * idx, * idx2;
  cudaMalloc(&src, (size_t)pool * 128);
  cudaMalloc(&dst, (size_t)n * 128);
  cudaMalloc(&idx, (size_t)n * 4);
  cudaMalloc(&idx2, (size_t)pool * 4);
  cudaMemset(src, 0, (size_t)pool * 128);
  std::vector<unsigned int> h(n), id(pool);
  for (unsigned int i = 0; i < pool; ++i) id[i] = i;
  cudaMemcpy(idx2, id.data(), (size_t)pool * 4, cudaMemcpyHostToDevice);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  printf("pool %.2f GB, %u blocks of 128 B per launch (read %.2f GB + write %.2f GB), %d SMs\n", pool * 128.0 / 1e9, n,
         n * 128.0 / 1e9, n * 128.0 / 1e9, sms);
  const char * names[] = {"seq", "random", "repeats fresh=0.9", "repeats fresh=0.5", "repeats fresh=0.17"};
  const double fresh[] = {0, 0, 0.9, 0.5, 0.17};
  for (int p = 0; p < 5; ++p)
  {
    unsigned int next = 0;
    for (unsigned int i = 0; i < n; ++i)
    {
      if (p == 0) h[i] = i % pool;
      else if (p == 1) h[i] = (unsigned int)(rnd() % pool);
      else
      {
        const bool f = next == 0 || (rnd() % 1000000) < fresh[p] * 1e6;
        h[i] = f ? next++ % pool : (unsigned int)(rnd() % next);
      }
    }
    cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 3; ++mode)
    {
      const double ms = run(mode, n, src, idx, idx2, dst, sms);
      printf("%-20s %-28s %8.3f ms  %7.0f GB/s (data read+written)\n", names[p],
             mode == 0 ? "1 index level, 4 in flight" : mode == 1 ? "2 index levels, 4 in flight" : "2 index levels, 8 in flight",
             ms, 2.0 * n * 128.0 / ms / 1e6);
    }
  }
  return 0;
}
