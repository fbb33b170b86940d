/*
 * plf_device.cuh -- device-side arithmetic shared by the sm_100a kernels.
 *
 * The whole translation unit is compiled with -fmad=false: a*b+c is two
 * roundings unless written as fma().  That is what makes CLVs and integer
 * scalers reproducible bit-for-bit against the reference's AVX (4 states, no
 * FMA) and AVX2 (other state counts, explicit FMA lanes) kernels; the
 * evaluation orders below are those pinned by oracle/plf_oracle.c against the
 * reference (tests/test_oracle_vs_reference.py).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PLF_SCALE_FACTOR 0x1p+256
#define PLF_SCALE_THRESHOLD 0x1p-256
#define PLF_LOG_SCALE_THRESHOLD (-177.445678223345993274) /* log(2^-256) */
#define PLF_MAXDIFF 4

typedef unsigned long long plf_state_t;

/* ---- 256-bit global accesses (Blackwell LDG.E.256 / STG.E.256) ---------- */
struct __align__(32) dbl4 { double x, y, z, w; };

__device__ __forceinline__ dbl4 ld256(const double * p)
{
  dbl4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
/* read-only, streaming: do not keep in L1 (each CLV byte is used once) */
__device__ __forceinline__ dbl4 ld256_stream(const double * p)
{
  dbl4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
/* coherent at L2: for values another CTA of the SAME launch wrote (after an acquire of its flag) */
__device__ __forceinline__ dbl4 ld256_cg(const double * p)
{
  dbl4 v;
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st256(double * p, const dbl4 & v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

/* ---- 4 states: multiplies + pairwise tree, no fma ----------------------- *
 * (core_partials_avx.c:456-524) */
__device__ __forceinline__ double dot4_pairwise(const double * m, const dbl4 & c)
{
  double p0 = m[0] * c.x, p1 = m[1] * c.y, p2 = m[2] * c.z, p3 = m[3] * c.w;
  return (p0 + p1) + (p2 + p3);
}
/* masked pairwise sum of a 4-entry matrix row (core_partials_avx.c:1355-1395) */
__device__ __forceinline__ double masked_sum4(const double * m, unsigned int mask)
{
  double p0 = (mask & 1u) ? m[0] : 0.0, p1 = (mask & 2u) ? m[1] : 0.0;
  double p2 = (mask & 4u) ? m[2] : 0.0, p3 = (mask & 8u) ? m[3] : 0.0;
  return (p0 + p1) + (p2 + p3);
}

/* ---- other state counts: 4 fma lanes over column quads, then pairwise --- *
 * (core_partials_avx2.c:695-771, :1114-1227) */
template <typename MP, typename CP>
__device__ __forceinline__ double dot_lanes_fma(MP m, CP c, int n)
{
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int j = 0; j < n; j += 4)
  {
    a0 = fma(m[j + 0], c[j + 0], a0);
    a1 = fma(m[j + 1], c[j + 1], a1);
    a2 = fma(m[j + 2], c[j + 2], a2);
    a3 = fma(m[j + 3], c[j + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}
/* scalar, increasing column order (core_partials_avx2.c:387-456) */
__device__ __forceinline__ double masked_sum_seq(const double * m, plf_state_t mask, int states)
{
  double t = 0;
  for (int k = 0; k < states; ++k)
    if ((mask >> k) & 1ull) t += m[k];
  return t;
}
/* lane adds over column quads then pairwise (core_partials_avx2.c:159-185) */
__device__ __forceinline__ double masked_sum_lanes(const double * m, plf_state_t mask, int n)
{
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int j = 0; j < n; j += 4)
  {
    if ((mask >> j) & 0xFull)
    {
      a0 = a0 + (((mask >> (j + 0)) & 1ull) ? m[j + 0] : 0.0);
      a1 = a1 + (((mask >> (j + 1)) & 1ull) ? m[j + 1] : 0.0);
      a2 = a2 + (((mask >> (j + 2)) & 1ull) ? m[j + 2] : 0.0);
      a3 = a3 + (((mask >> (j + 3)) & 1ull) ? m[j + 3] : 0.0);
    }
  }
  return (a0 + a1) + (a2 + a3);
}

/* ---- warp helpers -------------------------------------------------------- */
/* AND of `v` over an aligned group of `L` lanes (L power of two <= 32) */
__device__ __forceinline__ int group_and(int v, int L)
{
  for (int o = L >> 1; o > 0; o >>= 1) v &= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned int group_min_u(unsigned int v, int L)
{
  for (int o = L >> 1; o > 0; o >>= 1)
  {
    unsigned int w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
/* fixed-shape sum over an aligned group of L lanes (same result in every lane) */
__device__ __forceinline__ double group_sum(double v, int L)
{
  for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
/* deterministic block sum; result valid in thread 0.  `red` >= 32 doubles. */
__device__ __forceinline__ double block_sum(double v, double * red)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0;
  if (w == 0)
  {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

/* ---- fused grid reduction ------------------------------------------------- *
 * Every block leaves its NV partial sums in partial[v * gridDim.x + block];
 * the block that arrives last (ticket counter) adds them in a fixed order and
 * writes out[v] (and hout[v], pinned host memory, when given): one launch, the
 * same bits run to run for a given grid size.  `red` >= 32 doubles.          */
template <int NV>
__device__ __forceinline__ void grid_reduce_finish(const double (&v)[NV], double * __restrict__ partial,
                                                   unsigned int * ticket, double * out, double * hout, double * red)
{
  __shared__ int s_last;
#pragma unroll
  for (int i = 0; i < NV; ++i)
  {
    const double r = block_sum(v[i], red);
    if (threadIdx.x == 0) partial[(size_t)i * gridDim.x + blockIdx.x] = r;
  }
  if (threadIdx.x == 0)
  {
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NV; ++i)
  {
    double acc = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) acc += __ldcg(partial + (size_t)i * gridDim.x + b);
    const double r = block_sum(acc, red);
    if (threadIdx.x == 0)
    {
      out[i] = r;
      if (hout) hout[i] = r;
    }
  }
  if (threadIdx.x == 0) *ticket = 0;
}
