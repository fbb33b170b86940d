/*
 * plf_peer.cu -- the one exchange of a site-sharded evaluation, as a kernel over peer memory.
 *
 * Each rank (one process per GPU of one node) owns a contiguous site slice; an evaluation ends with the sum over
 * the ranks of at most four doubles: {logL} or {logL, d_f, dd_f} (the reference's clients
 * do this with MPI_Allreduce between their site-split threads/ranks).  Twenty-four bytes do not need a
 * collective library: every rank exposes one small buffer through CUDA IPC, and ONE single-block kernel on the
 * partition's stream
 *     1. stores its values and a sequence number into its slot of EVERY peer's buffer (NVLink / NVSwitch stores),
 *     2. waits until all slots of its own buffer carry this sequence number,
 *     3. adds them in rank order (every rank forms the same sum, bit for bit) and overwrites its values.
 * Slots are double-buffered by the parity of the sequence number: a rank can only be two exchanges ahead of the
 * slowest one after that rank has read the older values (it takes part in the exchange in between).
 * A rank that never shows up does not hang the device: the wait gives up after PEER_TIMEOUT_NS and the call
 * reports failure at the next pll_cuda_peer_group_check().
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "plf_internal.h"

#define PEER_MAX_WORLD 64
#define PEER_MAX_VALUES 4
#define PEER_TIMEOUT_NS 20000000000ull /* 20 s */

struct peer_slot
{
  double v[PEER_MAX_VALUES];
  unsigned long long seq;
  unsigned long long pad[3];
};

struct pll_cuda_peer_group
{
  int device;
  unsigned int rank, world;
  peer_slot * mine;               /* [2][world], written by the peers */
  peer_slot * peers[PEER_MAX_WORLD]; /* every rank's buffer as mapped here (own entry = mine) */
  peer_slot ** d_peers;
  int * d_error;                  /* set by a wait that timed out */
  unsigned long long seq;
  int connected;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long * p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long * p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(PEER_MAX_WORLD)
k_peer_allreduce(peer_slot * const * __restrict__ peers, peer_slot * mine, unsigned int rank, unsigned int world,
                 unsigned long long seq, double * values, unsigned int count, int * error)
{
  __shared__ int failed;
  const unsigned int t = threadIdx.x;
  const unsigned int set = (unsigned int)(seq & 1ull);
  if (t == 0) failed = 0;
  __syncthreads();
  if (t < world)
  {
    /* my values into slot `rank` of peer t's buffer */
    peer_slot * dst = peers[t] + (size_t)set * world + rank;
    for (unsigned int i = 0; i < count; ++i) reinterpret_cast<volatile double *>(dst->v)[i] = values[i];
    st_release_sys(&dst->seq, seq);
    /* and wait for peer t's values in my own buffer */
    const peer_slot * src = mine + (size_t)set * world + t;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(&src->seq) != seq)
      if (global_ns() - t0 > PEER_TIMEOUT_NS)
      {
        failed = 1;
        break;
      }
  }
  __syncthreads();
  if (t == 0)
  {
    if (failed)
    {
      *error = 1;
      for (unsigned int i = 0; i < count; ++i) values[i] = __longlong_as_double(0x7ff8000000000000ll);
    }
    else
      for (unsigned int i = 0; i < count; ++i)
      {
        double acc = 0;
        for (unsigned int r = 0; r < world; ++r)
          acc += reinterpret_cast<const volatile double *>(mine[(size_t)set * world + r].v)[i];
        values[i] = acc;
      }
  }
}

#define PEER_EXPORT extern "C" __attribute__((visibility("default")))

/* rank `rank` of `world` on `device`: allocates the exchange buffer; handle_out receives the 64-byte IPC handle
 * the other ranks need (exchange them by any means, e.g. an all-gather at start-up) */
PEER_EXPORT pll_cuda_peer_group * pll_cuda_peer_group_create(int device, unsigned int rank, unsigned int world,
                                                             void * handle_out)
{
  if (!world || world > PEER_MAX_WORLD || rank >= world || !handle_out) return NULL;
  if (cudaSetDevice(device) != cudaSuccess) return NULL;
  pll_cuda_peer_group * g = (pll_cuda_peer_group *)calloc(1, sizeof(*g));
  if (!g) return NULL;
  g->device = device;
  g->rank = rank;
  g->world = world;
  const size_t bytes = (size_t)2 * world * sizeof(peer_slot);
  cudaIpcMemHandle_t h;
  if (cudaMalloc(&g->mine, bytes) != cudaSuccess || cudaMemset(g->mine, 0, bytes) != cudaSuccess ||
      cudaMalloc(&g->d_peers, PEER_MAX_WORLD * sizeof(peer_slot *)) != cudaSuccess ||
      cudaMalloc(&g->d_error, sizeof(int)) != cudaSuccess || cudaMemset(g->d_error, 0, sizeof(int)) != cudaSuccess ||
      cudaIpcGetMemHandle(&h, g->mine) != cudaSuccess)
  {
    cudaGetLastError();
    cudaFree(g->mine);
    cudaFree(g->d_peers);
    cudaFree(g->d_error);
    free(g);
    return NULL;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, 64);
  cudaDeviceSynchronize();
  return g;
}

/* handles: world x 64 bytes in rank order (this rank's own entry is ignored) */
PEER_EXPORT int pll_cuda_peer_group_connect(pll_cuda_peer_group * g, const void * handles)
{
  if (!g || !handles || cudaSetDevice(g->device) != cudaSuccess) return 0;
  for (unsigned int r = 0; r < g->world; ++r)
  {
    if (r == g->rank)
    {
      g->peers[r] = g->mine;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * 64, 64);
    void * p = NULL;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
    g->peers[r] = (peer_slot *)p;
  }
  if (cudaMemcpy(g->d_peers, g->peers, g->world * sizeof(peer_slot *), cudaMemcpyHostToDevice) != cudaSuccess) return 0;
  g->connected = 1;
  return 1;
}

/* dev_values[0 .. count) <- sum over the ranks, on `stream` (the partition's stream: the values were left there
 * by pll_cuda_edge_loglikelihood_async / pll_cuda_likelihood_derivatives_async).  Every rank must make the same
 * sequence of calls.  Nothing is waited for on the host. */
PEER_EXPORT int pll_cuda_peer_allreduce(pll_cuda_peer_group * g, void * stream, double * dev_values, unsigned int count)
{
  if (!g || !g->connected || !dev_values || !count || count > PEER_MAX_VALUES) return 0;
  if (cudaSetDevice(g->device) != cudaSuccess) return 0;
  ++g->seq;
  k_peer_allreduce<<<1, PEER_MAX_WORLD, 0, (cudaStream_t)stream>>>(g->d_peers, g->mine, g->rank, g->world, g->seq, dev_values,
                                                                    count, g->d_error);
  plf_count_launch();
  return cudaGetLastError() == cudaSuccess;
}

/* 1 when no exchange has timed out so far (synchronises the device) */
PEER_EXPORT int pll_cuda_peer_group_check(pll_cuda_peer_group * g)
{
  int err = 1;
  if (!g || cudaSetDevice(g->device) != cudaSuccess) return 0;
  if (cudaMemcpy(&err, g->d_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return err == 0;
}

PEER_EXPORT void pll_cuda_peer_group_destroy(pll_cuda_peer_group * g)
{
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (unsigned int r = 0; r < g->world; ++r)
    if (g->connected && r != g->rank && g->peers[r]) cudaIpcCloseMemHandle(g->peers[r]);
  cudaFree(g->mine);
  cudaFree(g->d_peers);
  cudaFree(g->d_error);
  free(g);
}
