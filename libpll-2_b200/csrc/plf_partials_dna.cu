/*
 * plf_partials_dna.cu -- the DNA (4-state) CLV update kernels, one per op kind.
 *
 * These are the dominant kernels of a traversal: 396 / 265 / 134 algorithmic
 * bytes per site for inner-inner / tip-inner / tip-tip at 4 rate categories
 * against 240 non-fused FP64 operations, i.e. HBM-bound by a wide margin.  The
 * design is therefore about bytes in flight, not arithmetic:
 *
 *   - one thread per (site, rate): its 32-byte CLV block is ONE 256-bit load
 *     or store, a warp touches 1 KB contiguous per access;
 *   - every thread keeps U independent sites in flight (all loads of an
 *     iteration are issued before the first use), so a resident SM holds
 *     threads x U x 64 B of outstanding reads without needing more warps;
 *   - the two 4x4 P-matrices of the thread's rate stay in registers for the
 *     whole kernel (persistent grid: one wave of CTAs, each striding over the
 *     sites), tip lookup tables live in shared memory;
 *   - the parent scaler (left + right [+1]) is produced by the same pass
 *     (reference: separate pll_fill_parent_scaler, src/pll.c:1202).
 *
 * Arithmetic order is the reference's AVX 4x4 kernels (multiplies, pairwise
 * adds, no FMA; src/core_partials_avx.c:402-563,992-1030,1310-1480): results are
 * bit-identical.  The repeats variants of the reference
 * (src/core_partials_avx.c:567,761) are the same kernels with the three
 * identifier arrays non-NULL.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"
#include "plf_stream.cuh"

#include <stdlib.h>

#define DNA_THREADS 128

struct SiteRef
{
  unsigned int n, lid, rid;
  bool active;
};

__device__ __forceinline__ SiteRef resolve_site(const plf_op_t & op, unsigned int n)
{
  SiteRef s;
  s.n = n;
  s.active = n < op.nsites;
  s.lid = s.rid = n;
  if (s.active && (op.parent_id_site || op.left_site_id || op.right_site_id))
  {
    const unsigned int site = op.parent_id_site ? op.parent_id_site[n] : n;
    s.lid = op.left_site_id ? op.left_site_id[site] : site;
    s.rid = op.right_site_id ? op.right_site_id[site] : site;
  }
  return s;
}

/* scaling test + scaler store for one (site, rate) block `v`; LOG2R = log2(rate_cats) */
template <int LOG2R>
__device__ __forceinline__ void scale_and_store(const plf_op_t & op, const SiteRef & s, int rate, int per_rate,
                                                unsigned int sc_in, dbl4 & v)
{
  constexpr int R = 1 << LOG2R;
  if (op.parent_scaler)
  {
    const int below = (v.x < PLF_SCALE_THRESHOLD) && (v.y < PLF_SCALE_THRESHOLD) && (v.z < PLF_SCALE_THRESHOLD) &&
                      (v.w < PLF_SCALE_THRESHOLD);
    const int fire = per_rate ? below : group_and(below, R); /* all lanes of the warp take part */
    if (fire)
    {
      v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR;
      v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
    }
    if (s.active)
    {
      if (per_rate)
        op.parent_scaler[(size_t)s.n * R + rate] = sc_in + (fire ? 1u : 0u);
      else if (rate == 0)
        op.parent_scaler[s.n] = sc_in + (fire ? 1u : 0u);
    }
  }
  if (s.active) st256(op.parent_clv + ((size_t)s.n * R + rate) * 4, v);
}

/* ---- inner-inner ------------------------------------------------------------ */
template <int LOG2R, int U>
__global__ void __launch_bounds__(DNA_THREADS, 3)
k_clv_dna_ii(const plf_op_t * __restrict__ ops, int per_rate)
{
  constexpr int R = 1 << LOG2R;
  const plf_op_t op = ops[blockIdx.y];
  const unsigned int tid = blockIdx.x * DNA_THREADS + threadIdx.x;
  const int rate = tid & (R - 1);
  const unsigned int site0 = tid >> LOG2R;
  const unsigned int pass = (gridDim.x * DNA_THREADS) >> LOG2R; /* sites one sweep of the grid covers */

  double Lm[16], Rm[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
  {
    Lm[i] = op.left_matrix[rate * 16 + i];
    Rm[i] = op.right_matrix[rate * 16 + i];
  }

  for (unsigned int base = 0; base < op.nsites; base += pass * U)
  {
    SiteRef s[U];
    dbl4 l[U], r[U];
    unsigned int sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      s[u] = resolve_site(op, base + u * pass + site0);
      l[u] = r[u] = dbl4{0, 0, 0, 0};
      sc[u] = 0;
      if (s[u].active)
      {
        l[u] = ld256_stream(op.left_clv + ((size_t)s[u].lid * R + rate) * 4);
        r[u] = ld256_stream(op.right_clv + ((size_t)s[u].rid * R + rate) * 4);
        if (op.parent_scaler)
        {
          if (per_rate)
            sc[u] = (op.left_scaler ? op.left_scaler[(size_t)s[u].lid * R + rate] : 0u) +
                    (op.right_scaler ? op.right_scaler[(size_t)s[u].rid * R + rate] : 0u);
          else if (rate == 0)
            sc[u] = (op.left_scaler ? op.left_scaler[s[u].lid] : 0u) + (op.right_scaler ? op.right_scaler[s[u].rid] : 0u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      dbl4 v;
      v.x = dot4_pairwise(Lm + 0, l[u]) * dot4_pairwise(Rm + 0, r[u]);
      v.y = dot4_pairwise(Lm + 4, l[u]) * dot4_pairwise(Rm + 4, r[u]);
      v.z = dot4_pairwise(Lm + 8, l[u]) * dot4_pairwise(Rm + 8, r[u]);
      v.w = dot4_pairwise(Lm + 12, l[u]) * dot4_pairwise(Rm + 12, r[u]);
      scale_and_store<LOG2R>(op, s[u], rate, per_rate, sc[u], v);
    }
  }
}

/* ---- inner-inner under site repeats: ops of very different sizes in one launch ---- *
 * The ops of a level are compressed to their own class counts (a few dozen to all the   *
 * sites), so giving every op the same share of the grid (gridDim.y = ops) leaves most   *
 * of the chip waiting for the largest one (ncu: a 91-op level took 689 us for 71 MB).    *
 * Here the level is ONE list of tiles (DNA_THREADS x U (site, rate) items each,          *
 * tile_prefix[op] = first tile of the op, built by the host from the class counts) and   *
 * every CTA takes a contiguous share of it, reloading the op descriptor and its two      *
 * P-matrices only when it crosses into the next op.                                      */
template <int LOG2R, int U>
__global__ void __launch_bounds__(DNA_THREADS, 3)
k_clv_dna_ii_balanced(const plf_op_t * __restrict__ ops, int per_rate_and_nops,
                      const unsigned int * __restrict__ tile_prefix)
{
  constexpr int R = 1 << LOG2R;
  constexpr unsigned int TILE_SITES = (DNA_THREADS * U) >> LOG2R;
  const int per_rate = per_rate_and_nops & 1;
  const unsigned int nops = (unsigned int)per_rate_and_nops >> 1;
  const unsigned int total = tile_prefix[nops];
  const unsigned int share = (total + gridDim.x - 1) / gridDim.x;
  const unsigned int lo = blockIdx.x * share;
  const unsigned int hi = min(lo + share, total);
  if (lo >= hi) return;
  const int rate = threadIdx.x & (R - 1);
  const unsigned int site_in_tile = threadIdx.x >> LOG2R;
  constexpr unsigned int PASS = DNA_THREADS >> LOG2R; /* sites one sweep of the CTA covers */

  /* op of the first tile: binary search in the prefix array */
  unsigned int cur = 0;
  {
    unsigned int a = 0, b = nops;
    while (b - a > 1)
    {
      const unsigned int m = (a + b) >> 1;
      if (tile_prefix[m] <= lo)
        a = m;
      else
        b = m;
    }
    cur = a;
  }
  plf_op_t op = ops[cur];
  double Lm[16], Rm[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
  {
    Lm[i] = op.left_matrix[rate * 16 + i];
    Rm[i] = op.right_matrix[rate * 16 + i];
  }
  for (unsigned int t = lo; t < hi; ++t)
  {
    if (t >= tile_prefix[cur + 1])
    {
      do
        ++cur;
      while (t >= tile_prefix[cur + 1]);
      op = ops[cur];
#pragma unroll
      for (int i = 0; i < 16; ++i)
      {
        Lm[i] = op.left_matrix[rate * 16 + i];
        Rm[i] = op.right_matrix[rate * 16 + i];
      }
    }
    const unsigned int base = (t - tile_prefix[cur]) * TILE_SITES;
    SiteRef s[U];
    dbl4 l[U], r[U];
    unsigned int sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      s[u] = resolve_site(op, base + u * PASS + site_in_tile);
      l[u] = r[u] = dbl4{0, 0, 0, 0};
      sc[u] = 0;
      if (s[u].active)
      {
        l[u] = ld256_stream(op.left_clv + ((size_t)s[u].lid * R + rate) * 4);
        r[u] = ld256_stream(op.right_clv + ((size_t)s[u].rid * R + rate) * 4);
        if (op.parent_scaler)
        {
          if (per_rate)
            sc[u] = (op.left_scaler ? op.left_scaler[(size_t)s[u].lid * R + rate] : 0u) +
                    (op.right_scaler ? op.right_scaler[(size_t)s[u].rid * R + rate] : 0u);
          else if (rate == 0)
            sc[u] = (op.left_scaler ? op.left_scaler[s[u].lid] : 0u) + (op.right_scaler ? op.right_scaler[s[u].rid] : 0u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      dbl4 v;
      v.x = dot4_pairwise(Lm + 0, l[u]) * dot4_pairwise(Rm + 0, r[u]);
      v.y = dot4_pairwise(Lm + 4, l[u]) * dot4_pairwise(Rm + 4, r[u]);
      v.z = dot4_pairwise(Lm + 8, l[u]) * dot4_pairwise(Rm + 8, r[u]);
      v.w = dot4_pairwise(Lm + 12, l[u]) * dot4_pairwise(Rm + 12, r[u]);
      scale_and_store<LOG2R>(op, s[u], rate, per_rate, sc[u], v);
    }
  }
}

/* The same tile walk with the two P-matrices (and the op descriptor) staged in shared memory instead of held
 * in 64 + ~28 registers per thread, and U = 2 instead of 4 items in flight: 64 registers, 8 instead of 3 resident
 * CTAs per SM.  The gathers are latency-bound (class -> site -> child class -> CLV block), so resident warps are
 * what buys bandwidth: config 4 traversal 2.23 -> 1.39 ms (3.1 -> 5.0 TB/s algorithmic); sweep over U x CTAs in
 * profiles/r1_notes.md.  Rows are read back with 128-bit shared loads, 8 rows per (site, rate) item; the rows of
 * the rates are 18 doubles apart so that lanes of different rates hit different banks (16 apart: 4-way conflicts,
 * 3.04 ms). */
template <int LOG2R, int U, int CTAS>
__global__ void __launch_bounds__(DNA_THREADS, CTAS)
k_clv_dna_ii_balanced_sm(const plf_op_t * __restrict__ ops, int per_rate_and_nops,
                         const unsigned int * __restrict__ tile_prefix)
{
  constexpr int R = 1 << LOG2R;
  constexpr unsigned int TILE_SITES = (DNA_THREADS * 4) >> LOG2R;
  constexpr unsigned int PASS = DNA_THREADS >> LOG2R;
  constexpr int MS = 18; /* doubles per rate: 16 + 2 of padding, so that the rates' rows fall into different banks */
  __shared__ __align__(16) double Ls[R * MS];
  __shared__ __align__(16) double Rs[R * MS];
  __shared__ plf_op_t s_op;
  const int per_rate = per_rate_and_nops & 1;
  const unsigned int nops = (unsigned int)per_rate_and_nops >> 1;
  const unsigned int total = tile_prefix[nops];
  const unsigned int share = (total + gridDim.x - 1) / gridDim.x;
  const unsigned int lo = blockIdx.x * share;
  const unsigned int hi = min(lo + share, total);
  if (lo >= hi) return;
  const int rate = threadIdx.x & (R - 1);
  const unsigned int site_in_tile = threadIdx.x >> LOG2R;

  unsigned int cur = 0;
  {
    unsigned int a = 0, b = nops;
    while (b - a > 1)
    {
      const unsigned int m = (a + b) >> 1;
      if (tile_prefix[m] <= lo)
        a = m;
      else
        b = m;
    }
    cur = a;
  }
  unsigned int next = tile_prefix[cur + 1], first = tile_prefix[cur];
  bool reload = true;
  for (unsigned int t = lo; t < hi; ++t)
  {
    if (t >= next)
    {
      do
        ++cur;
      while (t >= tile_prefix[cur + 1]);
      next = tile_prefix[cur + 1];
      first = tile_prefix[cur];
      reload = true;
    }
    if (reload) /* the same for every thread of the CTA */
    {
      __syncthreads(); /* nobody still reads the previous op's matrices */
      if (threadIdx.x == 0) s_op = ops[cur];
      {
        const plf_op_t & o = ops[cur];
        for (int i = threadIdx.x; i < R * 16; i += DNA_THREADS)
        {
          Ls[(i >> 4) * MS + (i & 15)] = o.left_matrix[i];
          Rs[(i >> 4) * MS + (i & 15)] = o.right_matrix[i];
        }
      }
      __syncthreads();
      reload = false;
    }
    const plf_op_t & op = s_op;
    const double * Lm = Ls + rate * MS;
    const double * Rm = Rs + rate * MS;
    for (unsigned int part = 0; part < 4 / U; ++part) /* a tile is 4 items per thread, U of them in flight */
    {
    const unsigned int base = (t - first) * TILE_SITES + part * U * PASS;
    SiteRef s[U];
    dbl4 l[U], r[U];
    unsigned int sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      s[u] = resolve_site(op, base + u * PASS + site_in_tile);
      l[u] = r[u] = dbl4{0, 0, 0, 0};
      sc[u] = 0;
      if (s[u].active)
      {
        l[u] = ld256_stream(op.left_clv + ((size_t)s[u].lid * R + rate) * 4);
        r[u] = ld256_stream(op.right_clv + ((size_t)s[u].rid * R + rate) * 4);
        if (op.parent_scaler)
        {
          if (per_rate)
            sc[u] = (op.left_scaler ? op.left_scaler[(size_t)s[u].lid * R + rate] : 0u) +
                    (op.right_scaler ? op.right_scaler[(size_t)s[u].rid * R + rate] : 0u);
          else if (rate == 0)
            sc[u] = (op.left_scaler ? op.left_scaler[s[u].lid] : 0u) + (op.right_scaler ? op.right_scaler[s[u].rid] : 0u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      dbl4 v;
      v.x = dot4_pairwise(Lm + 0, l[u]) * dot4_pairwise(Rm + 0, r[u]);
      v.y = dot4_pairwise(Lm + 4, l[u]) * dot4_pairwise(Rm + 4, r[u]);
      v.z = dot4_pairwise(Lm + 8, l[u]) * dot4_pairwise(Rm + 8, r[u]);
      v.w = dot4_pairwise(Lm + 12, l[u]) * dot4_pairwise(Rm + 12, r[u]);
      scale_and_store<LOG2R>(op, s[u], rate, per_rate, sc[u], v);
    }
    }
  }
}

/* The tile walk once more, for ops that carry a PAIR LIST (plf_op_t::pair_list, written once per identifier
 * update by k_rep_pairs): entry n of the parent reads entry pair[n].x of the left and pair[n].y of the right
 * child.  The chain class -> site -> child class -> CLV block of resolve_site() (three dependent loads, the
 * first two of them 4-byte gathers) becomes ONE coalesced 8-byte load, and that load is issued one step ahead:
 * while the CLV blocks of the current items are in flight the pairs of the next items arrive, so every step
 * waits for a single level of gathers. */
template <int LOG2R, int U, int CTAS>
__global__ void __launch_bounds__(DNA_THREADS, CTAS)
k_clv_dna_ii_pairs(const plf_op_t * __restrict__ ops, int per_rate_and_nops,
                   const unsigned int * __restrict__ tile_prefix)
{
  constexpr int R = 1 << LOG2R;
  constexpr unsigned int TILE_SITES = (DNA_THREADS * 4) >> LOG2R;
  constexpr unsigned int PASS = DNA_THREADS >> LOG2R;
  constexpr unsigned int PARTS = 4 / U;
  constexpr int MS = 18;
  __shared__ __align__(16) double Ls[R * MS];
  __shared__ __align__(16) double Rs[R * MS];
  __shared__ plf_op_t s_op;
  const int per_rate = per_rate_and_nops & 1;
  const unsigned int nops = (unsigned int)per_rate_and_nops >> 1;
  const unsigned int total = tile_prefix[nops];
  const unsigned int share = (total + gridDim.x - 1) / gridDim.x;
  const unsigned int lo = blockIdx.x * share;
  const unsigned int hi = min(lo + share, total);
  if (lo >= hi) return;
  const int rate = threadIdx.x & (R - 1);
  const unsigned int site_in_tile = threadIdx.x >> LOG2R;

  unsigned int cur = 0;
  {
    unsigned int a = 0, b = nops;
    while (b - a > 1)
    {
      const unsigned int m = (a + b) >> 1;
      if (tile_prefix[m] <= lo)
        a = m;
      else
        b = m;
    }
    cur = a;
  }
  unsigned int next = tile_prefix[cur + 1], first = tile_prefix[cur];
  bool reload = true;
  uint2 pr[U];
  bool have = false; /* pr[] already holds the pairs of the step about to run */
  for (unsigned int t = lo; t < hi; ++t)
  {
    if (t >= next)
    {
      do
        ++cur;
      while (t >= tile_prefix[cur + 1]);
      next = tile_prefix[cur + 1];
      first = tile_prefix[cur];
      reload = true;
    }
    if (reload)
    {
      __syncthreads();
      if (threadIdx.x == 0) s_op = ops[cur];
      {
        const plf_op_t & o = ops[cur];
        for (int i = threadIdx.x; i < R * 16; i += DNA_THREADS)
        {
          Ls[(i >> 4) * MS + (i & 15)] = o.left_matrix[i];
          Rs[(i >> 4) * MS + (i & 15)] = o.right_matrix[i];
        }
      }
      __syncthreads();
      reload = false;
      have = false;
    }
    const plf_op_t & op = s_op;
    const uint2 * __restrict__ pairs = reinterpret_cast<const uint2 *>(op.pair_list);
    const double * Lm = Ls + rate * MS;
    const double * Rm = Rs + rate * MS;
    for (unsigned int part = 0; part < PARTS; ++part)
    {
      const unsigned int base = (t - first) * TILE_SITES + part * U * PASS;
      SiteRef s[U];
      dbl4 l[U], r[U];
      unsigned int sc[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        s[u].n = base + u * PASS + site_in_tile;
        s[u].active = s[u].n < op.nsites;
        if (!have) pr[u] = s[u].active ? __ldg(pairs + s[u].n) : make_uint2(0, 0);
        s[u].lid = pr[u].x;
        s[u].rid = pr[u].y;
        l[u] = r[u] = dbl4{0, 0, 0, 0};
        sc[u] = 0;
        if (s[u].active)
        {
          l[u] = ld256_stream(op.left_clv + ((size_t)s[u].lid * R + rate) * 4);
          r[u] = ld256_stream(op.right_clv + ((size_t)s[u].rid * R + rate) * 4);
          if (op.parent_scaler)
          {
            if (per_rate)
              sc[u] = (op.left_scaler ? op.left_scaler[(size_t)s[u].lid * R + rate] : 0u) +
                      (op.right_scaler ? op.right_scaler[(size_t)s[u].rid * R + rate] : 0u);
            else if (rate == 0)
              sc[u] = (op.left_scaler ? op.left_scaler[s[u].lid] : 0u) + (op.right_scaler ? op.right_scaler[s[u].rid] : 0u);
          }
        }
      }
      /* the pairs of the next step of the SAME op, while the CLV blocks above are on their way */
      {
        const bool same_tile = part + 1 < PARTS;
        const unsigned int nbase = same_tile ? base + U * PASS : (t + 1 - first) * TILE_SITES;
        have = same_tile || (t + 1 < hi && t + 1 < next);
        if (have)
        {
#pragma unroll
          for (int u = 0; u < U; ++u)
          {
            const unsigned int nn = nbase + u * PASS + site_in_tile;
            pr[u] = nn < op.nsites ? __ldg(pairs + nn) : make_uint2(0, 0);
          }
        }
      }
      /* ncu (profiles/r2_full_pairs_summary.txt): the L1 data pipe, not DRAM, is what this kernel fills (78 % of
       * its peak: 16 shared-memory row reads per item next to the gathers).  Each matrix row is therefore read
       * ONCE per step and applied to all U items before anything is stored (the stores are asm volatile with a
       * memory clobber: between them the compiler must reload shared memory). */
      dbl4 v[U];
      {
        double ra[U], rb[U];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
          const dbl4 lrow = lds_dbl4(Lm + 4 * i), rrow = lds_dbl4(Rm + 4 * i);
#pragma unroll
          for (int u = 0; u < U; ++u)
          {
            ra[u] = ((lrow.x * l[u].x) + (lrow.y * l[u].y)) + ((lrow.z * l[u].z) + (lrow.w * l[u].w));
            rb[u] = ((rrow.x * r[u].x) + (rrow.y * r[u].y)) + ((rrow.z * r[u].z) + (rrow.w * r[u].w));
            const double e = ra[u] * rb[u];
            if (i == 0) v[u].x = e;
            if (i == 1) v[u].y = e;
            if (i == 2) v[u].z = e;
            if (i == 3) v[u].w = e;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) scale_and_store<LOG2R>(op, s[u], rate, per_rate, sc[u], v[u]);
    }
  }
}

/* ---- tip-inner (the tip is "left") --------------------------------------------- */
template <int LOG2R, int U>
__global__ void __launch_bounds__(DNA_THREADS, 4)
k_clv_dna_ti(const plf_op_t * __restrict__ ops, int per_rate)
{
  constexpr int R = 1 << LOG2R;
  __shared__ __align__(16) double tl[64 * R];
  const plf_op_t op = ops[blockIdx.y];
  build_tip_table(tl, op.left_matrix, R);
  const unsigned int tid = blockIdx.x * DNA_THREADS + threadIdx.x;
  const int rate = tid & (R - 1);
  const unsigned int site0 = tid >> LOG2R;
  const unsigned int pass = (gridDim.x * DNA_THREADS) >> LOG2R;
  double Rm[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) Rm[i] = op.right_matrix[rate * 16 + i];
  __syncthreads();

  for (unsigned int base = 0; base < op.nsites; base += pass * U)
  {
    SiteRef s[U];
    dbl4 r[U];
    unsigned int sc[U], code[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      s[u] = resolve_site(op, base + u * pass + site0);
      r[u] = dbl4{0, 0, 0, 0};
      sc[u] = 0;
      code[u] = 0;
      if (s[u].active)
      {
        r[u] = ld256_stream(op.right_clv + ((size_t)s[u].rid * R + rate) * 4);
        code[u] = op.left_tip[s[u].lid];
        if (op.parent_scaler && op.right_scaler)
        {
          if (per_rate)
            sc[u] = op.right_scaler[(size_t)s[u].rid * R + rate];
          else if (rate == 0)
            sc[u] = op.right_scaler[s[u].rid];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const dbl4 a = lds_dbl4(tl + (code[u] * R + rate) * 4);
      dbl4 v;
      v.x = a.x * dot4_pairwise(Rm + 0, r[u]);
      v.y = a.y * dot4_pairwise(Rm + 4, r[u]);
      v.z = a.z * dot4_pairwise(Rm + 8, r[u]);
      v.w = a.w * dot4_pairwise(Rm + 12, r[u]);
      scale_and_store<LOG2R>(op, s[u], rate, per_rate, sc[u], v);
    }
  }
}

/* ---- tip-tip: table product, never scales, scaler zeroed ------------------------ *
 * (src/core_partials_avx.c:255-400,992-1030: the reference builds the 16x16      *
 * product table once per op; multiplying the two half-tables per site gives the   *
 * same bits)                                                                       */
#define TT_GROUP 8 /* consecutive sites per thread: their 8 tip codes are one 64-bit load per tip */

template <int LOG2R, int U>
__global__ void __launch_bounds__(DNA_THREADS, 6)
k_clv_dna_tt(const plf_op_t * __restrict__ ops, int per_rate_and_nops)
{
  /* A write-only kernel (134 B/site out, 2 B/site in).  One thread owns the
   * (rate) block of TT_GROUP consecutive sites: a single round of load latency
   * (two 64-bit code loads) is followed by 8 table products and 8 256-bit
   * stores.  The two half-tables are read with the 16-byte halves swapped
   * between the even and odd site group of a quarter-warp, which makes the
   * LDS.128 accesses conflict-free. */
  constexpr int R = 1 << LOG2R;
  __shared__ __align__(128) double tl[64 * R];
  __shared__ __align__(128) double tr[64 * R];
  const int per_rate = per_rate_and_nops & 1;
  const unsigned int nops = (unsigned int)per_rate_and_nops >> 1;
  const unsigned int tid = blockIdx.x * DNA_THREADS + threadIdx.x;
  const int rate = tid & (R - 1);
  const unsigned int grp0 = tid >> LOG2R;
  const unsigned int pass = (gridDim.x * DNA_THREADS) >> LOG2R;
  /* gridDim.y == 1: the whole grid sweeps one op after the other, so that the
   * chip writes ONE parent CLV at a time (few open DRAM pages) */
  for (unsigned int o = blockIdx.y; o < nops; o += gridDim.y)
  {
  const plf_op_t op = ops[o];
  __syncthreads();
  build_tip_table(tl, op.left_matrix, R);
  build_tip_table(tr, op.right_matrix, R);
  __syncthreads();
  const unsigned int ngroups = (op.nsites + TT_GROUP - 1) / TT_GROUP;
  const bool direct = !(op.parent_id_site || op.left_site_id || op.right_site_id);
  const int swap = (R <= 4) ? ((threadIdx.x >> LOG2R) & 1) : 0; /* odd group of the quarter-warp */

  /* the codes of the NEXT group are fetched before the current one is
   * processed: the only load latency on this path is hidden behind 8 stores */
  unsigned long long lc8_next = 0, rc8_next = 0;
  if (direct && grp0 < ngroups)
  {
    /* the host layer pads tip buffers by 16 bytes: the last group may read past nsites */
    lc8_next = *reinterpret_cast<const unsigned long long *>(op.left_tip + (size_t)grp0 * TT_GROUP);
    rc8_next = *reinterpret_cast<const unsigned long long *>(op.right_tip + (size_t)grp0 * TT_GROUP);
  }
  for (unsigned int g = grp0; g < ngroups; g += pass)
  {
    const unsigned int first = g * TT_GROUP;
    unsigned long long lc8 = lc8_next, rc8 = rc8_next;
    if (direct)
    {
      if (g + pass < ngroups)
      {
        lc8_next = *reinterpret_cast<const unsigned long long *>(op.left_tip + (size_t)(g + pass) * TT_GROUP);
        rc8_next = *reinterpret_cast<const unsigned long long *>(op.right_tip + (size_t)(g + pass) * TT_GROUP);
      }
    }
    else
    {
#pragma unroll
      for (int j = 0; j < TT_GROUP; ++j)
      {
        const unsigned int n = first + j;
        if (n < op.nsites)
        {
          const unsigned int site = op.parent_id_site ? op.parent_id_site[n] : n;
          lc8 |= (unsigned long long)op.left_tip[op.left_site_id ? op.left_site_id[site] : site] << (8 * j);
          rc8 |= (unsigned long long)op.right_tip[op.right_site_id ? op.right_site_id[site] : site] << (8 * j);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < TT_GROUP; ++j)
    {
      const unsigned int n = first + j;
      if (n >= op.nsites) break;
      const unsigned int lc = (unsigned int)(lc8 >> (8 * j)) & 0xFFu, rc = (unsigned int)(rc8 >> (8 * j)) & 0xFFu;
      const double * pl = tl + (lc * R + rate) * 4;
      const double * pr = tr + (rc * R + rate) * 4;
      const double2 a0 = *reinterpret_cast<const double2 *>(pl + 2 * swap);
      const double2 b0 = *reinterpret_cast<const double2 *>(pr + 2 * swap);
      const double2 a1 = *reinterpret_cast<const double2 *>(pl + 2 * (swap ^ 1));
      const double2 b1 = *reinterpret_cast<const double2 *>(pr + 2 * (swap ^ 1));
      const double p0 = a0.x * b0.x, p1 = a0.y * b0.y, q0 = a1.x * b1.x, q1 = a1.y * b1.y;
      const dbl4 v = swap ? dbl4{q0, q1, p0, p1} : dbl4{p0, p1, q0, q1};
      st256(op.parent_clv + ((size_t)n * R + rate) * 4, v);
      if (op.parent_scaler && per_rate) op.parent_scaler[(size_t)n * R + rate] = 0;
    }
    if (op.parent_scaler && !per_rate && rate == 0)
    {
      if (first + TT_GROUP <= op.nsites)
      {
        uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(op.parent_scaler + first) = z;
        *reinterpret_cast<uint4 *>(op.parent_scaler + first + 4) = z;
      }
      else
        for (unsigned int n = first; n < op.nsites; ++n) op.parent_scaler[n] = 0;
    }
  }
  }
}

/* ---- tip-tip through shared memory and bulk stores ------------------------------ *
 * Same arithmetic, but the parent tile is assembled in shared memory and leaves the  *
 * SM as ONE bulk async store (cp.async.bulk.global.shared::cta, the copy engine       *
 * writes whole lines while the warps build the next tile); NOUT tiles in flight.     */
__device__ __forceinline__ void bulk_s2g(void * gdst, const void * ssrc, unsigned int bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}

template <int LOG2R, int ITEMS, int NOUT>
__global__ void __launch_bounds__(DNA_THREADS)
k_clv_dna_tt_bulk(const plf_op_t * __restrict__ ops, int per_rate_and_nops)
{
  constexpr int R = 1 << LOG2R;
  constexpr int TILE = (DNA_THREADS * ITEMS) >> LOG2R;
  constexpr int OUT_BYTES = TILE * R * 32;
  extern __shared__ __align__(128) unsigned char obuf[]; /* [NOUT][OUT_BYTES] */
  __shared__ __align__(128) double tl[64 * R];
  __shared__ __align__(128) double tr[64 * R];
  const int per_rate = per_rate_and_nops & 1;
  const unsigned int nops = (unsigned int)per_rate_and_nops >> 1;
  const int rate = threadIdx.x & (R - 1);
  const int swap = (R <= 4) ? ((threadIdx.x >> LOG2R) & 1) : 0; /* odd site of the quarter-warp: halves swapped */
  unsigned int ob = 0;
  for (unsigned int o = blockIdx.y; o < nops; o += gridDim.y)
  {
    const plf_op_t op = ops[o];
    __syncthreads();
    build_tip_table(tl, op.left_matrix, R);
    build_tip_table(tr, op.right_matrix, R);
    __syncthreads();
    const unsigned int ntiles = (op.nsites + TILE - 1) / TILE;
    unsigned int lc_next[ITEMS], rc_next[ITEMS];
    auto fetch_codes = [&](unsigned int t) {
#pragma unroll
      for (int u = 0; u < ITEMS; ++u)
      {
        const unsigned int n = t * TILE + ((threadIdx.x + u * DNA_THREADS) >> LOG2R);
        const unsigned int nn = n < op.nsites ? n : op.nsites - 1;
        lc_next[u] = op.left_tip[nn];
        rc_next[u] = op.right_tip[nn];
      }
    };
    if (blockIdx.x < ntiles) fetch_codes(blockIdx.x);
    for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ob)
    {
      unsigned int lc[ITEMS], rc[ITEMS];
#pragma unroll
      for (int u = 0; u < ITEMS; ++u)
      {
        lc[u] = lc_next[u];
        rc[u] = rc_next[u];
      }
      if (t + gridDim.x < ntiles) fetch_codes(t + gridDim.x);
      /* the buffer about to be overwritten must have been read by its bulk store */
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NOUT - 1) : "memory");
      __syncthreads();
      unsigned char * buf = obuf + (size_t)(ob % NOUT) * OUT_BYTES;
#pragma unroll
      for (int u = 0; u < ITEMS; ++u)
      {
        const unsigned int item = threadIdx.x + u * DNA_THREADS;
        const double * pl = tl + (lc[u] * R + rate) * 4;
        const double * pr = tr + (rc[u] * R + rate) * 4;
        const double2 a0 = *reinterpret_cast<const double2 *>(pl + 2 * swap);
        const double2 b0 = *reinterpret_cast<const double2 *>(pr + 2 * swap);
        const double2 a1 = *reinterpret_cast<const double2 *>(pl + 2 * (swap ^ 1));
        const double2 b1 = *reinterpret_cast<const double2 *>(pr + 2 * (swap ^ 1));
        *reinterpret_cast<double2 *>(buf + (size_t)item * 32 + 16 * swap) = make_double2(a0.x * b0.x, a0.y * b0.y);
        *reinterpret_cast<double2 *>(buf + (size_t)item * 32 + 16 * (swap ^ 1)) = make_double2(a1.x * b1.x, a1.y * b1.y);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      const unsigned int first = t * TILE;
      const unsigned int n = min((unsigned int)TILE, op.nsites - first);
      if (threadIdx.x == 0)
      {
        bulk_s2g(op.parent_clv + (size_t)first * R * 4, buf, n * R * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (op.parent_scaler)
      {
        const unsigned int entries = per_rate ? n * R : n;
        unsigned int * sc = op.parent_scaler + (per_rate ? (size_t)first * R : first);
        for (unsigned int e = threadIdx.x; e < entries; e += DNA_THREADS) sc[e] = 0;
      }
    }
  }
  /* shared memory must stay valid until the last stores have read it */
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

/* ------------------------------------------------------------------------ *
 *  Streaming variants for contiguous (non-repeats) CLVs: the child tiles,    *
 *  their scalers and the tip codes are brought into a shared-memory ring by  *
 *  1-D bulk async copies (cp.async.bulk, the non-tensor TMA path; UBLKCP in  *
 *  SASS) that complete on an mbarrier.  Bytes in flight are then set by the  *
 *  ring depth (NSTAGE x ~36 KB per CTA), not by registers or warp count, and *
 *  the copy engine keeps issuing while every warp is busy with arithmetic.   *
 * ------------------------------------------------------------------------ */

/* what a child of a streamed op is: an inner CLV (tile copied into the ring, 4x4 mat-vec with the matrix
 * held in registers), a pattern tip (code copied into the ring, 16-entry lookup table in shared memory) or a
 * virtual cherry (the codes of its two tips copied into the ring, 256-entry lookup table in shared memory) */
enum { CK_I = 0, CK_T = 1, CK_C = 2 };

/* ITEMS = (site, rate) blocks per thread and tile */
template <int LOG2R, int LK, int RK, int ITEMS, int THREADS = DNA_THREADS>
struct StreamLayout
{
  static constexpr int R = 1 << LOG2R;
  static constexpr int TILE = (THREADS * ITEMS) >> LOG2R; /* sites per tile */
  static constexpr int CLV_BYTES = TILE * R * 32;
  static constexpr int SC_BYTES = TILE * R * 4; /* per-rate worst case */
  static constexpr int CODE_BYTES = (TILE + 127) & ~127; /* whole 128-byte lines: the CLV tile behind a code array stays line-aligned */
  static constexpr int side_bytes(int k) { return k == CK_I ? CLV_BYTES + SC_BYTES : k == CK_T ? CODE_BYTES : 2 * CODE_BYTES; }
  /* left side: CLV tile then scalers, or one / two code arrays; the right side follows */
  static constexpr int OFF_L = 0;
  static constexpr int OFF_LSC = CLV_BYTES;
  static constexpr int OFF_LCODE = 0;
  static constexpr int OFF_LCODE2 = CODE_BYTES;
  static constexpr int OFF_R = side_bytes(LK);
  static constexpr int OFF_RSC = OFF_R + CLV_BYTES;
  static constexpr int OFF_RCODE = OFF_R;
  static constexpr int OFF_RCODE2 = OFF_R + CODE_BYTES;
  static constexpr int STAGE_BYTES = (side_bytes(LK) + side_bytes(RK) + 127) & ~127;
  /* lookup tables behind the ring: doubles */
  static constexpr int table_doubles(int k) { return k == CK_T ? 64 * R : k == CK_C ? 1024 * R : 0; }
  static constexpr int TAB_L = 0;
  static constexpr int TAB_R = table_doubles(LK);
  static constexpr int TAB_SCRATCH = TAB_R + table_doubles(RK); /* the two half tables a cherry table is built from */
  static constexpr int TAB_DOUBLES = TAB_SCRATCH + ((LK == CK_C || RK == CK_C) ? 128 * R : 0);
  static constexpr size_t smem_bytes(int nstage) { return (size_t)nstage * STAGE_BYTES + (size_t)TAB_DOUBLES * 8; }
};

/* thread 0: queue the copies of tile `t` of `op` into ring slot `slot` */
template <int LOG2R, int LK, int RK, int ITEMS, int THREADS>
__device__ __forceinline__ void stream_issue(const plf_op_t & op, unsigned int t, unsigned char * slot,
                                             unsigned long long * bar, int per_rate)
{
  typedef StreamLayout<LOG2R, LK, RK, ITEMS, THREADS> Ly;
  const unsigned int first = t * Ly::TILE;
  const unsigned int n = min((unsigned int)Ly::TILE, op.nsites - first);
  const unsigned int clv_bytes = n * Ly::R * 32;
  /* 4-byte scalers / 1-byte codes: sizes rounded up to the 16-byte copy
   * granule; the host layer pads those allocations by 16 bytes */
  const unsigned int sc_bytes = ((per_rate ? n * Ly::R : n) * 4 + 15) & ~15u;
  const unsigned int code_bytes = (n + 15) & ~15u;
  const size_t sc_first = per_rate ? (size_t)first * Ly::R : first;
  const bool lsc = LK == CK_I && op.left_scaler && op.parent_scaler;
  const bool rsc = RK == CK_I && op.right_scaler && op.parent_scaler;
  unsigned int total = 0;
  total += LK == CK_I ? clv_bytes + (lsc ? sc_bytes : 0) : LK == CK_T ? code_bytes : 2 * code_bytes;
  total += RK == CK_I ? clv_bytes + (rsc ? sc_bytes : 0) : RK == CK_T ? code_bytes : 2 * code_bytes;
  mbar_expect_tx(bar, total);
  if (LK == CK_I)
  {
    bulk_g2s(slot + Ly::OFF_L, op.left_clv + (size_t)first * Ly::R * 4, clv_bytes, bar);
    if (lsc) bulk_g2s(slot + Ly::OFF_LSC, op.left_scaler + sc_first, sc_bytes, bar);
  }
  else
  {
    bulk_g2s(slot + Ly::OFF_LCODE, op.left_tip + first, code_bytes, bar);
    if (LK == CK_C) bulk_g2s(slot + Ly::OFF_LCODE2, op.left_tip2 + first, code_bytes, bar);
  }
  if (RK == CK_I)
  {
    bulk_g2s(slot + Ly::OFF_R, op.right_clv + (size_t)first * Ly::R * 4, clv_bytes, bar);
    if (rsc) bulk_g2s(slot + Ly::OFF_RSC, op.right_scaler + sc_first, sc_bytes, bar);
  }
  else
  {
    bulk_g2s(slot + Ly::OFF_RCODE, op.right_tip + first, code_bytes, bar);
    if (RK == CK_C) bulk_g2s(slot + Ly::OFF_RCODE2, op.right_tip2 + first, code_bytes, bar);
  }
}

/* lookup table of a virtual cherry seen through the branch above it:
 *   tab[(codeA << 4 | codeB)][rate][i] = row i of `outer` (pairwise dot, as an inner child is treated,
 *   src/core_partials_avx.c:456-524) times the cherry's CLV entry, which is termA[k] * termB[k] with the
 *   masked pairwise sums of the two tip matrices (the reference's tip-tip table, :295-396,1012-1028).
 * The same operations in the same order as writing the cherry and reading it back: the same bits. */
template <int LOG2R>
__device__ __forceinline__ void build_cherry_table(double * tab, double * scratch, const double * __restrict__ outer,
                                                   const double * __restrict__ cm1, const double * __restrict__ cm2)
{
  constexpr int R = 1 << LOG2R;
  __syncthreads(); /* the scratch may still be read by the previous table's build */
  build_tip_table(scratch, cm1, R);
  build_tip_table(scratch + 64 * R, cm2, R);
  /* entry e = ((codeA * 16 + codeB) * R + rate) * 4 + i: a thread's (rate, i) = e mod 4R never changes
   * (the block size is a multiple of 4R), so its row of `outer` is read from global memory once */
  /* (block sizes are multiples of 4R = 4 .. 16) */
  const int i = threadIdx.x & 3, rate = (threadIdx.x >> 2) & (R - 1);
  const double o0 = outer[rate * 16 + i * 4 + 0], o1 = outer[rate * 16 + i * 4 + 1];
  const double o2 = outer[rate * 16 + i * 4 + 2], o3 = outer[rate * 16 + i * 4 + 3];
  __syncthreads();
#pragma unroll 4
  for (int e = threadIdx.x; e < 1024 * R; e += blockDim.x)
  {
    const int cb = (e >> (2 + LOG2R)) & 15, ca = e >> (6 + LOG2R);
    const double2 a0 = *reinterpret_cast<const double2 *>(scratch + (ca * R + rate) * 4);
    const double2 a1 = *reinterpret_cast<const double2 *>(scratch + (ca * R + rate) * 4 + 2);
    const double2 b0 = *reinterpret_cast<const double2 *>(scratch + 64 * R + (cb * R + rate) * 4);
    const double2 b1 = *reinterpret_cast<const double2 *>(scratch + 64 * R + (cb * R + rate) * 4 + 2);
    /* dot4_pairwise(outer row, cherry entry): multiplies, then the pairwise tree */
    const double p0 = o0 * (a0.x * b0.x), p1 = o1 * (a0.y * b0.y), p2 = o2 * (a1.x * b1.x), p3 = o3 * (a1.y * b1.y);
    tab[e] = (p0 + p1) + (p2 + p3);
  }
}

/* 32-byte table row read as two 16-byte halves; odd site groups take the halves in the other order, so that
 * two neighbouring sites of a warp cover all 32 banks whatever rows they read (a row is 128 bytes apart from
 * the next: without this every site of the warp would hit the same 16 banks) */
__device__ __forceinline__ dbl4 lds_dbl4_swz(const double * p, int swap)
{
  const double2 a = *reinterpret_cast<const double2 *>(p + 2 * swap);
  const double2 b = *reinterpret_cast<const double2 *>(p + 2 * (swap ^ 1));
  return swap ? dbl4{b.x, b.y, a.x, a.y} : dbl4{a.x, a.y, b.x, b.y};
}

template <int LOG2R, int LK, int RK, int NSTAGE, int ITEMS, int THREADS = DNA_THREADS>
__global__ void __launch_bounds__(THREADS)
k_clv_dna_stream(const plf_op_t * __restrict__ ops, int per_rate)
{
  typedef StreamLayout<LOG2R, LK, RK, ITEMS, THREADS> Ly;
  constexpr int R = Ly::R;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) unsigned long long full[NSTAGE];
  double * tables = reinterpret_cast<double *>(ring + (size_t)NSTAGE * Ly::STAGE_BYTES);
  double * tabL = tables + Ly::TAB_L;
  double * tabR = tables + Ly::TAB_R;

  const plf_op_t op = ops[blockIdx.y];
  const unsigned int ntiles = (op.nsites + Ly::TILE - 1) / Ly::TILE;
  const int rate = threadIdx.x & (R - 1);
  const int swz = (R <= 4) ? ((threadIdx.x >> LOG2R) & 1) : 0; /* odd site of a pair: table halves swapped */

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned int t = blockIdx.x;
    for (int s = 0; s < NSTAGE && t < ntiles; ++s, t += gridDim.x)
      stream_issue<LOG2R, LK, RK, ITEMS, THREADS>(op, t, ring + (size_t)s * Ly::STAGE_BYTES, &full[s], per_rate);
  }

  double Lm[LK == CK_I ? 16 : 1], Rm[RK == CK_I ? 16 : 1];
#pragma unroll
  for (int i = 0; i < 16; ++i)
  {
    if (LK == CK_I) Lm[i] = op.left_matrix[rate * 16 + i];
    if (RK == CK_I) Rm[i] = op.right_matrix[rate * 16 + i];
  }
  if (LK == CK_T) build_tip_table(tabL, op.left_matrix, R);
  if (LK == CK_C) build_cherry_table<LOG2R>(tabL, tables + Ly::TAB_SCRATCH, op.left_matrix, op.left_cm1, op.left_cm2);
  if (RK == CK_C) build_cherry_table<LOG2R>(tabR, tables + Ly::TAB_SCRATCH, op.right_matrix, op.right_cm1, op.right_cm2);
  if (LK != CK_I || RK != CK_I) __syncthreads();

  unsigned int it = 0;
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it)
  {
    const int s = it % NSTAGE;
    const unsigned int parity = (it / NSTAGE) & 1u;
    unsigned char * slot = ring + (size_t)s * Ly::STAGE_BYTES;
    while (!mbar_try_wait(&full[s], parity)) {}

    const unsigned int first = t * Ly::TILE;
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
    {
      const unsigned int item = threadIdx.x + u * THREADS; /* (site in tile, rate) */
      const unsigned int ls = item >> LOG2R;
      SiteRef sr;
      sr.n = first + ls;
      sr.lid = sr.rid = sr.n;
      sr.active = sr.n < op.nsites;
      dbl4 a;
      unsigned int sc = 0;
      if (LK == CK_I)
      {
        const dbl4 l = lds_dbl4(reinterpret_cast<const double *>(slot + Ly::OFF_L) + (size_t)item * 4);
        a.x = dot4_pairwise(Lm + 0, l);
        a.y = dot4_pairwise(Lm + 4, l);
        a.z = dot4_pairwise(Lm + 8, l);
        a.w = dot4_pairwise(Lm + 12, l);
      }
      else if (LK == CK_T)
      {
        /* 16 rows, a handful of them hot: neighbouring sites mostly read the same row (broadcast) */
        const unsigned int code = sr.active ? slot[Ly::OFF_LCODE + ls] : 0u;
        a = lds_dbl4(tabL + (code * R + rate) * 4);
      }
      else
      {
        const unsigned int code = sr.active ? (((slot[Ly::OFF_LCODE + ls] & 15u) << 4) | (slot[Ly::OFF_LCODE2 + ls] & 15u)) : 0u;
        a = lds_dbl4_swz(tabL + (code * R + rate) * 4, swz);
      }
      if (op.parent_scaler && sr.active && (per_rate || rate == 0))
      {
        const unsigned int k = per_rate ? item : ls;
        if (LK == CK_I && op.left_scaler) sc += reinterpret_cast<const unsigned int *>(slot + Ly::OFF_LSC)[k];
        if (RK == CK_I && op.right_scaler) sc += reinterpret_cast<const unsigned int *>(slot + Ly::OFF_RSC)[k];
      }
      dbl4 v;
      if (RK == CK_I)
      {
        const dbl4 r = lds_dbl4(reinterpret_cast<const double *>(slot + Ly::OFF_R) + (size_t)item * 4);
        v.x = a.x * dot4_pairwise(Rm + 0, r);
        v.y = a.y * dot4_pairwise(Rm + 4, r);
        v.z = a.z * dot4_pairwise(Rm + 8, r);
        v.w = a.w * dot4_pairwise(Rm + 12, r);
      }
      else
      {
        const unsigned int code = sr.active ? (((slot[Ly::OFF_RCODE + ls] & 15u) << 4) | (slot[Ly::OFF_RCODE2 + ls] & 15u)) : 0u;
        const dbl4 b = lds_dbl4_swz(tabR + (code * R + rate) * 4, swz);
        v.x = a.x * b.x;
        v.y = a.y * b.y;
        v.z = a.z * b.z;
        v.w = a.w * b.w;
      }
      if (!sr.active) v = dbl4{1.0, 1.0, 1.0, 1.0}; /* stale ring bytes must not reach the scaling vote */
      scale_and_store<LOG2R>(op, sr, rate, per_rate, sc, v);
    }
    __syncthreads(); /* every warp is done with this slot */
    const unsigned int tn = t + (unsigned int)NSTAGE * gridDim.x;
    if (threadIdx.x == 0 && tn < ntiles) stream_issue<LOG2R, LK, RK, ITEMS, THREADS>(op, tn, slot, &full[s], per_rate);
  }
}

/* ---- write-only consumers of virtual cherries: tip + cherry, cherry + cherry (PLF_CHERRY_BULK=1) ----- *
 * NOT the default: at 3 CTAs per SM (32 KB of output tiles + a 32 KB table) and two block barriers per tile   *
 * it runs the 21 such ops of a config-2 traversal 0.53 ms slower than the ring kernel (profiles/r2_notes.md). *
 * Neither child is a CLV: the parent is a product of two table rows, 3 or 4 bytes of tip codes in, 132 bytes   *
 * out per site.  Same shape as the tip-tip kernel above (tiles assembled in shared memory, one bulk async      *
 * store each, NOUT in flight, codes of the next tile fetched while the current one is built), plus what a      *
 * tip-inner / inner-inner operation of the reference does and a tip-tip one does not: the scaling test         *
 * (src/core_partials_avx.c:526-563).                                                                            */
template <int LOG2R, int ITEMS, int NOUT, int LC>
__global__ void __launch_bounds__(DNA_THREADS)
k_clv_dna_lookup_bulk(const plf_op_t * __restrict__ ops, int per_rate)
{
  constexpr int R = 1 << LOG2R;
  constexpr int TILE = (DNA_THREADS * ITEMS) >> LOG2R;
  constexpr int OUT_BYTES = TILE * R * 32;
  constexpr int TABL = LC ? 1024 * R : 64 * R; /* left: cherry or tip; right: always a cherry */
  extern __shared__ __align__(128) unsigned char dynbuf[]; /* [NOUT][OUT_BYTES] | tabL | tabR | scratch */
  unsigned char * obuf = dynbuf;
  double * tabL = reinterpret_cast<double *>(dynbuf + (size_t)NOUT * OUT_BYTES);
  double * tabR = tabL + TABL;
  double * scratch = tabR + 1024 * R;
  const plf_op_t op = ops[blockIdx.y];
  const int rate = threadIdx.x & (R - 1);
  const int swz = (R <= 4) ? ((threadIdx.x >> LOG2R) & 1) : 0;
  if (LC)
    build_cherry_table<LOG2R>(tabL, scratch, op.left_matrix, op.left_cm1, op.left_cm2);
  else
    build_tip_table(tabL, op.left_matrix, R);
  build_cherry_table<LOG2R>(tabR, scratch, op.right_matrix, op.right_cm1, op.right_cm2);
  __syncthreads();
  const unsigned int ntiles = (op.nsites + TILE - 1) / TILE;
  unsigned int lc_next[ITEMS], rc_next[ITEMS];
  auto fetch_codes = [&](unsigned int t) {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
    {
      const unsigned int n = t * TILE + ((threadIdx.x + u * DNA_THREADS) >> LOG2R);
      const unsigned int nn = n < op.nsites ? n : op.nsites - 1;
      lc_next[u] = LC ? (((op.left_tip[nn] & 15u) << 4) | (op.left_tip2[nn] & 15u)) : op.left_tip[nn];
      rc_next[u] = ((op.right_tip[nn] & 15u) << 4) | (op.right_tip2[nn] & 15u);
    }
  };
  if (blockIdx.x < ntiles) fetch_codes(blockIdx.x);
  unsigned int ob = 0;
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ob)
  {
    unsigned int lc[ITEMS], rc[ITEMS];
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
    {
      lc[u] = lc_next[u];
      rc[u] = rc_next[u];
    }
    if (t + gridDim.x < ntiles) fetch_codes(t + gridDim.x);
    /* the buffer about to be overwritten must have been read by its bulk store */
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NOUT - 1) : "memory");
    __syncthreads();
    unsigned char * buf = obuf + (size_t)(ob % NOUT) * OUT_BYTES;
    const unsigned int first = t * TILE;
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
    {
      const unsigned int item = threadIdx.x + u * DNA_THREADS;
      const unsigned int n = first + (item >> LOG2R);
      const dbl4 a = LC ? lds_dbl4_swz(tabL + (lc[u] * R + rate) * 4, swz) : lds_dbl4(tabL + (lc[u] * R + rate) * 4);
      const dbl4 b = lds_dbl4_swz(tabR + (rc[u] * R + rate) * 4, swz);
      dbl4 v = dbl4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
      if (op.parent_scaler)
      {
        const int below = (v.x < PLF_SCALE_THRESHOLD) && (v.y < PLF_SCALE_THRESHOLD) && (v.z < PLF_SCALE_THRESHOLD) &&
                          (v.w < PLF_SCALE_THRESHOLD);
        const int fire = per_rate ? below : group_and(below, R);
        if (fire)
        {
          v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR;
          v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
        }
        if (n < op.nsites)
        {
          if (per_rate)
            op.parent_scaler[(size_t)n * R + rate] = fire ? 1u : 0u;
          else if (rate == 0)
            op.parent_scaler[n] = fire ? 1u : 0u;
        }
      }
      /* odd sites store their halves in the other order: each 16-byte store instruction covers all banks */
      const double2 h0 = make_double2(v.x, v.y), h1 = make_double2(v.z, v.w);
      *reinterpret_cast<double2 *>(buf + (size_t)item * 32 + 16 * swz) = swz ? h1 : h0;
      *reinterpret_cast<double2 *>(buf + (size_t)item * 32 + 16 * (swz ^ 1)) = swz ? h0 : h1;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0)
    {
      const unsigned int n = min((unsigned int)TILE, op.nsites - first);
      bulk_s2g(op.parent_clv + (size_t)first * R * 4, buf, n * R * 32);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

/* A virtual cherry's own "operation": its CLV is never written.  The two P-matrices it was asked to be
 * computed with are copied into the node's side buffer (consumers and a later materialisation read them
 * there, so a P-matrix update in between changes nothing), and its scaler is zeroed as the reference's
 * tip-tip kernel does (src/core_partials_avx.c:1005-1006). */
__global__ void __launch_bounds__(256)
k_cherry_prepare(const plf_op_t * __restrict__ ops, int per_rate, int R)
{
  const plf_op_t op = ops[blockIdx.y];
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < R * 16; i += blockDim.x)
    {
      op.parent_clv[i] = op.left_matrix[i];
      op.parent_clv[R * 16 + i] = op.right_matrix[i];
    }
  if (!op.parent_scaler) return;
  const size_t n = per_rate ? (size_t)op.nsites * R : op.nsites;
  const size_t quads = (n + 3) >> 2; /* the allocation is padded by 16 bytes */
  uint4 * dst = reinterpret_cast<uint4 *>(op.parent_scaler);
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x)
    dst[q] = make_uint4(0, 0, 0, 0);
}

/* ------------------------------------------------------------------------ *
 *  Narrow alignments: ONE launch per traversal level.  A 100-taxon traversal  *
 *  is 12 levels but 22-28 launches of the per-kind kernels above, and below   *
 *  ~10^4 sites a launch costs what a CTA's life costs (barrier set-up, table   *
 *  build, one tile), not what its bytes cost.  This kernel takes every op of   *
 *  a level whatever its kind (blockIdx.y = op; the kind is uniform per CTA, so *
 *  the branches do not diverge): children are read straight from global        *
 *  memory (inner), looked up in a 16-row table (tip) or formed from two such   *
 *  tables (virtual cherry: entry k = tA[codeA][k] * tB[codeB][k], then the     *
 *  4x4 product with the matrix of the branch above it); a virtual cherry's own *
 *  "operation" (matrix snapshot, zeroed scaler) rides along.  Same arithmetic, *
 *  same order as the per-kind kernels: the same bits.                          *
 * ------------------------------------------------------------------------ */
template <int LOG2R>
__global__ void __launch_bounds__(DNA_THREADS)
k_clv_dna_level(const plf_op_t * __restrict__ ops, int per_rate)
{
  constexpr int R = 1 << LOG2R;
  __shared__ __align__(16) double tabs[4][64 * R]; /* left A, left B, right A, right B */
  const plf_op_t op = ops[blockIdx.y];
  const unsigned int kind = op.kind;
  if (kind == PLF_OP_TT_VIRTUAL)
  {
    if (blockIdx.x == 0)
      for (int i = threadIdx.x; i < R * 16; i += DNA_THREADS)
      {
        op.parent_clv[i] = op.left_matrix[i];
        op.parent_clv[R * 16 + i] = op.right_matrix[i];
      }
    if (op.parent_scaler)
    {
      const size_t n = per_rate ? (size_t)op.nsites * R : op.nsites;
      for (size_t e = (size_t)blockIdx.x * DNA_THREADS + threadIdx.x; e < n; e += (size_t)gridDim.x * DNA_THREADS)
        op.parent_scaler[e] = 0;
    }
    return;
  }
  /* what each child is: 0 inner CLV, 1 pattern tip, 2 virtual cherry */
  const int lk = (kind == PLF_OP_II) ? CK_I : (kind == PLF_OP_CI || kind == PLF_OP_CC) ? CK_C : CK_T;
  const int rk = (kind == PLF_OP_TT) ? CK_T : (kind == PLF_OP_TC || kind == PLF_OP_CC) ? CK_C : CK_I;
  if (lk == CK_T) build_tip_table(tabs[0], op.left_matrix, R);
  if (lk == CK_C)
  {
    build_tip_table(tabs[0], op.left_cm1, R);
    build_tip_table(tabs[1], op.left_cm2, R);
  }
  if (rk == CK_T) build_tip_table(tabs[2], op.right_matrix, R);
  if (rk == CK_C)
  {
    build_tip_table(tabs[2], op.right_cm1, R);
    build_tip_table(tabs[3], op.right_cm2, R);
  }
  const unsigned int tid = blockIdx.x * DNA_THREADS + threadIdx.x;
  const int rate = tid & (R - 1);
  const unsigned int site0 = tid >> LOG2R;
  const unsigned int pass = (gridDim.x * DNA_THREADS) >> LOG2R;
  double Lm[16], Rm[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
  {
    Lm[i] = (lk != CK_T) ? op.left_matrix[rate * 16 + i] : 0.0;
    Rm[i] = (rk != CK_T) ? op.right_matrix[rate * 16 + i] : 0.0;
  }
  __syncthreads();
  const bool scales = kind != PLF_OP_TT; /* tip-tip never scales and zeroes its scaler */
  for (unsigned int base = 0; base < op.nsites; base += pass)
  {
    SiteRef s;
    s.n = base + site0;
    s.lid = s.rid = s.n;
    s.active = s.n < op.nsites;
    const unsigned int n = s.active ? s.n : op.nsites - 1; /* inactive lanes recompute the last site, store nothing */
    dbl4 a, b;
    unsigned int sc = 0;
    if (lk == CK_T)
      a = lds_dbl4(tabs[0] + (op.left_tip[n] * R + rate) * 4);
    else
    {
      dbl4 l;
      if (lk == CK_I)
        l = ld256_stream(op.left_clv + ((size_t)n * R + rate) * 4);
      else
      {
        const dbl4 x = lds_dbl4(tabs[0] + ((op.left_tip[n] & 15u) * R + rate) * 4);
        const dbl4 y = lds_dbl4(tabs[1] + ((op.left_tip2[n] & 15u) * R + rate) * 4);
        l = dbl4{x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w};
      }
      a.x = dot4_pairwise(Lm + 0, l);
      a.y = dot4_pairwise(Lm + 4, l);
      a.z = dot4_pairwise(Lm + 8, l);
      a.w = dot4_pairwise(Lm + 12, l);
    }
    if (rk == CK_T)
      b = lds_dbl4(tabs[2] + (op.right_tip[n] * R + rate) * 4);
    else
    {
      dbl4 r;
      if (rk == CK_I)
        r = ld256_stream(op.right_clv + ((size_t)n * R + rate) * 4);
      else
      {
        const dbl4 x = lds_dbl4(tabs[2] + ((op.right_tip[n] & 15u) * R + rate) * 4);
        const dbl4 y = lds_dbl4(tabs[3] + ((op.right_tip2[n] & 15u) * R + rate) * 4);
        r = dbl4{x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w};
      }
      b.x = dot4_pairwise(Rm + 0, r);
      b.y = dot4_pairwise(Rm + 4, r);
      b.z = dot4_pairwise(Rm + 8, r);
      b.w = dot4_pairwise(Rm + 12, r);
    }
    if (op.parent_scaler && s.active && (per_rate || rate == 0))
    {
      if (lk == CK_I && op.left_scaler) sc += op.left_scaler[per_rate ? (size_t)n * R + rate : n];
      if (rk == CK_I && op.right_scaler) sc += op.right_scaler[per_rate ? (size_t)n * R + rate : n];
    }
    dbl4 v = dbl4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
    if (scales)
      scale_and_store<LOG2R>(op, s, rate, per_rate, sc, v);
    else
    {
      if (s.active)
      {
        st256(op.parent_clv + ((size_t)s.n * R + rate) * 4, v);
        if (op.parent_scaler)
        {
          if (per_rate)
            op.parent_scaler[(size_t)s.n * R + rate] = 0;
          else if (rate == 0)
            op.parent_scaler[s.n] = 0;
        }
      }
    }
  }
}

/* one launch for every op of a level (any kinds, contiguous CLVs, rate_cats 1, 2 or 4) */
int plf_launch_dna_level(plf_ctx * ctx, const plf_op_t * d_ops, unsigned int nops, unsigned int rate_cats, int per_rate,
                         unsigned int max_sites)
{
  int log2r = 0;
  while ((1u << log2r) < rate_cats) ++log2r;
  const unsigned long long lanes = (unsigned long long)max_sites << log2r;
  unsigned long long bx = (lanes + DNA_THREADS - 1) / DNA_THREADS;
  const unsigned long long cap = ((unsigned long long)ctx->sm_count * 12 + nops - 1) / nops;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const dim3 grid((unsigned int)bx, nops);
  switch (log2r)
  {
    case 0: k_clv_dna_level<0><<<grid, DNA_THREADS, 0, ctx->stream>>>(d_ops, per_rate); break;
    case 1: k_clv_dna_level<1><<<grid, DNA_THREADS, 0, ctx->stream>>>(d_ops, per_rate); break;
    default: k_clv_dna_level<2><<<grid, DNA_THREADS, 0, ctx->stream>>>(d_ops, per_rate); break;
  }
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* ------------------------------------------------------------------------ *
 *  Narrow alignments, one step further: the WHOLE traversal in one launch.    *
 *  Twelve launches of a 100-taxon traversal still cost ~5 us each although a   *
 *  level's work is a fraction of that.  Here the op list becomes a queue of     *
 *  work items (path, chunk of sites).  A PATH is a chain of ops in which each   *
 *  op's parent is the next op's child and nobody else's: the host cuts the      *
 *  tree into such chains along the heavier child (plf_dna_flow_plan), so any    *
 *  root-to-tip walk crosses at most log2(n) of them.  Inside a path the parent  *
 *  stays in the thread's registers (a thread owns one (site, rate) block from   *
 *  the first op of the path to the last; sites and rates are independent), its  *
 *  scaler count with it: no barrier, no trip through L2.  Between paths,        *
 *  persistent CTAs claim items in queue order from an atomic counter, and an    *
 *  op whose other child is written by an earlier item of the same launch waits  *
 *  for exactly that item's flag (same chunk) instead of for a launch boundary.  *
 *  Descriptors, matrices and tip codes of the whole path are staged in shared    *
 *  memory before the first wait.                                                 *
 *  No deadlock: an item's producers have smaller queue positions (paths are      *
 *  queued in the order of their last ops), so they were claimed earlier by CTAs  *
 *  that are running; the unfinished item with the smallest position never waits. *
 *  Flags carry the launch's epoch, the last CTA to leave rewinds the queue and   *
 *  advances the epoch: nothing to reset between launches, so the kernel replays  *
 *  inside a CUDA graph as it is.                                                 *
 *  Arithmetic and order as the per-kind kernels (a tip term is the masked        *
 *  pairwise sum the tip tables hold): the same bits.                             *
 * ------------------------------------------------------------------------ */
struct plf_flow_ctrl
{
  unsigned long long epoch;
  unsigned int next_item;
  unsigned int exited;
};

__device__ __forceinline__ unsigned long long flow_ld_acquire(const unsigned long long * p)
{
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void flow_st_release(unsigned long long * p, unsigned long long v)
{
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* waits for an item's flag.  By construction the wait ends (see above); should it not - a corrupted workspace -
 * the launch fails loudly after FLOW_TIMEOUT_NS instead of hanging the device */
#define FLOW_TIMEOUT_NS 10000000000ull
__device__ __forceinline__ void flow_wait(const unsigned long long * f, unsigned long long epoch1)
{
  unsigned int polls = 0;
  unsigned long long t0 = 0;
  while (flow_ld_acquire(f) != epoch1)
    if ((++polls & 0x3FFu) == 0)
    {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (!t0) t0 = now;
      else if (now - t0 > FLOW_TIMEOUT_NS) __trap();
    }
}

/* 64 registers: in the throughput regime of the kernel resident CTAs count (two blocks per thread at 78 registers
 * and 6 CTAs per SM: 4 - 5 % slower from 10k sites on; profiles/r2_notes.md) */
#define FLOW_MIN_CTAS 8
#define FLOW_CARRY_MASK 3u   /* 1: the left child is the previous op's parent (in registers), 2: the right one */
#define FLOW_CARRY_SCALER 4u /* ... and its scaler counts with it */
#define FLOW_TIP_TIP 8u      /* never scales, zeroes its scaler (src/core_partials_avx.c:1005-1006) */

/* one row of a 4x4 matrix applied to U child blocks: the row is read from shared memory once */
template <int U>
__device__ __forceinline__ void flow_rows_inner(const double * M, const dbl4 (&c)[U], dbl4 (&t)[U])
{
#pragma unroll
  for (int row = 0; row < 4; ++row)
  {
    const double2 a = *reinterpret_cast<const double2 *>(M + row * 4), b = *reinterpret_cast<const double2 *>(M + row * 4 + 2);
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double p0 = a.x * c[u].x, p1 = a.y * c[u].y, p2 = b.x * c[u].z, p3 = b.y * c[u].w;
      const double v = (p0 + p1) + (p2 + p3);
      if (row == 0) t[u].x = v; else if (row == 1) t[u].y = v; else if (row == 2) t[u].z = v; else t[u].w = v;
    }
  }
}
/* ... and the masked pairwise row sums a tip table holds (plf_stream.cuh: build_tip_table) for U codes */
template <int U>
__device__ __forceinline__ void flow_rows_tip(const double * M, const unsigned int (&code)[U], dbl4 (&t)[U])
{
#pragma unroll
  for (int row = 0; row < 4; ++row)
  {
    const double2 a = *reinterpret_cast<const double2 *>(M + row * 4), b = *reinterpret_cast<const double2 *>(M + row * 4 + 2);
    const double m[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double v = masked_sum4(m, code[u]);
      if (row == 0) t[u].x = v; else if (row == 1) t[u].y = v; else if (row == 2) t[u].z = v; else t[u].w = v;
    }
  }
}

template <int LOG2R, int U>
__global__ void __launch_bounds__(DNA_THREADS, FLOW_MIN_CTAS)
k_clv_dna_flow(const plf_flow_op * __restrict__ fops, const unsigned int * __restrict__ path_start, unsigned int npaths,
               unsigned int nchunks, int per_rate, plf_flow_ctrl * ctrl, unsigned long long * ready)
{
  constexpr int R = 1 << LOG2R;
  constexpr unsigned int PASS = DNA_THREADS >> LOG2R; /* sites one sweep of the CTA covers; a work item is U sweeps */
  constexpr int MSTRIDE = 18;                         /* doubles per rate of a staged matrix: rates on different banks */
  constexpr int DWORDS = sizeof(plf_flow_op) / 8;
  __shared__ __align__(16) unsigned long long sdesc_raw[PLF_FLOW_PATH_MAX * DWORDS];
  __shared__ __align__(16) double smat[PLF_FLOW_PATH_MAX][2][R * MSTRIDE];
  __shared__ unsigned char scode[PLF_FLOW_PATH_MAX][2][U * PASS];
  __shared__ unsigned int s_item;
  __shared__ unsigned long long s_epoch;
  const plf_flow_op * sdesc = reinterpret_cast<const plf_flow_op *>(sdesc_raw);
  const unsigned int total = npaths * nchunks;
  if (threadIdx.x == 0)
  {
    s_epoch = *reinterpret_cast<volatile unsigned long long *>(&ctrl->epoch) + 1ull;
    s_item = atomicAdd(&ctrl->next_item, 1u);
  }
  __syncthreads();
  const unsigned long long epoch1 = s_epoch;
  const int rate = threadIdx.x & (R - 1);
  const unsigned int lane_site = threadIdx.x >> LOG2R;
  for (;;)
  {
    const unsigned int item = s_item;
    if (item >= total) break;
    const unsigned int path = item / nchunks, chunk = item - path * nchunks;
    const unsigned int first_op = path_start[path], n = path_start[path + 1] - first_op;
    unsigned int next = 0;
    if (threadIdx.x == 0) next = atomicAdd(&ctrl->next_item, 1u); /* in flight while this item is worked on */
    {
      const unsigned long long * src = reinterpret_cast<const unsigned long long *>(fops + first_op);
      for (unsigned int i = threadIdx.x; i < n * DWORDS; i += DNA_THREADS) sdesc_raw[i] = src[i];
    }
    __syncthreads();
    const unsigned int nsites = sdesc[0].nsites;
    const unsigned int first = chunk * (U * PASS);
    for (unsigned int i = threadIdx.x; i < n * 2 * R * 16; i += DNA_THREADS)
    {
      const unsigned int k = i / (2 * R * 16), rem = i - k * (2 * R * 16), side = rem / (R * 16), e = rem - side * (R * 16);
      smat[k][side][(e >> 4) * MSTRIDE + (e & 15)] = sdesc[k].matrix[side][e];
    }
    for (unsigned int i = threadIdx.x; i < n * 2 * U * PASS; i += DNA_THREADS)
    {
      const unsigned int k = i / (2 * U * PASS), rem = i - k * (2 * U * PASS), side = rem / (U * PASS), ls = rem - side * (U * PASS);
      const unsigned char * tp = sdesc[k].tip[side];
      if (tp) scode[k][side][ls] = tp[min(first + ls, nsites - 1)];
    }
    __syncthreads();
    if (first < nsites)
    {
      unsigned int site[U], nn[U];
      bool active[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        site[u] = first + u * PASS + lane_site;
        active[u] = site[u] < nsites;
        nn[u] = active[u] ? site[u] : nsites - 1; /* inactive lanes recompute the last site, store nothing */
      }
      const bool keeps_count = per_rate || rate == 0;
      dbl4 carry_v[U];
      unsigned int carry_sc[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        carry_v[u] = dbl4{0, 0, 0, 0};
        carry_sc[u] = 0;
      }
#pragma unroll 1
      for (unsigned int k = 0; k < n; ++k)
      {
        const plf_flow_op & d = sdesc[k];
        const unsigned int flags = d.flags;
        dbl4 term[2][U];
        unsigned int sc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) sc[u] = 0;
#pragma unroll
        for (int side = 0; side < 2; ++side)
        {
          const double * M = smat[k][side] + rate * MSTRIDE;
          if (d.tip[side])
          {
            unsigned int code[U];
#pragma unroll
            for (int u = 0; u < U; ++u) code[u] = scode[k][side][u * PASS + lane_site];
            flow_rows_tip<U>(M, code, term[side]);
          }
          else
          {
            dbl4 c[U];
            if ((flags & FLOW_CARRY_MASK) == (unsigned int)(side + 1))
            {
#pragma unroll
              for (int u = 0; u < U; ++u)
              {
                c[u] = carry_v[u];
                if (flags & FLOW_CARRY_SCALER) sc[u] += carry_sc[u];
              }
            }
            else
            {
              if (d.dep[side] >= 0)
              {
                const unsigned long long * f = ready + (size_t)d.dep[side] * nchunks + chunk;
                flow_wait(f, epoch1);
              }
#pragma unroll
              for (int u = 0; u < U; ++u) c[u] = ld256_cg(d.clv[side] + ((size_t)nn[u] * R + rate) * 4);
            }
            if (d.scaler[side] && d.parent_scaler && keeps_count)
            {
#pragma unroll
              for (int u = 0; u < U; ++u) sc[u] += __ldcg(d.scaler[side] + (per_rate ? (size_t)nn[u] * R + rate : nn[u]));
            }
            flow_rows_inner<U>(M, c, term[side]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
          dbl4 v = dbl4{term[0][u].x * term[1][u].x, term[0][u].y * term[1][u].y, term[0][u].z * term[1][u].z,
                        term[0][u].w * term[1][u].w};
          if (d.parent_scaler)
          {
            if (flags & FLOW_TIP_TIP)
              sc[u] = 0;
            else
            {
              const int below = (v.x < PLF_SCALE_THRESHOLD) && (v.y < PLF_SCALE_THRESHOLD) && (v.z < PLF_SCALE_THRESHOLD) &&
                                (v.w < PLF_SCALE_THRESHOLD);
              const int fire = per_rate ? below : group_and(below, R); /* all lanes of the warp take part */
              if (fire)
              {
                v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR;
                v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
                sc[u] += 1u;
              }
            }
            if (active[u] && keeps_count) d.parent_scaler[per_rate ? (size_t)site[u] * R + rate : site[u]] = sc[u];
          }
          if (active[u]) st256(d.parent_clv + ((size_t)site[u] * R + rate) * 4, v);
          carry_v[u] = v;
          carry_sc[u] = sc[u];
        }
      }
    }
    __syncthreads(); /* every store of the item has been issued */
    if (threadIdx.x == 0)
    {
      flow_st_release(ready + item, epoch1);
      s_item = next;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    __threadfence();
    if (atomicAdd(&ctrl->exited, 1u) == gridDim.x - 1)
    {
      /* every CTA has made its last claim: rewind the queue for the next launch */
      ctrl->next_item = 0;
      ctrl->exited = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned long long *>(&ctrl->epoch) = epoch1;
    }
  }
}

/* blocks of (site, rate) a thread carries through a path: 1 while a traversal is pure latency, 2 when the matrix
 * reads from shared memory start to count (each row is read once for both) */
unsigned int plf_dna_flow_unroll(unsigned int max_sites)
{
  const char * v = getenv("PLF_FLOW_UNROLL");
  if (v && (v[0] == '1' || v[0] == '2')) return (unsigned int)(v[0] - '0');
  v = getenv("PLF_FLOW_UNROLL_SITES");
  return max_sites >= ((v && v[0]) ? strtoul(v, nullptr, 10) : 2048ul) ? 2u : 1u;
}

unsigned int plf_dna_flow_chunks(unsigned int rate_cats, unsigned int max_sites)
{
  int log2r = 0;
  while ((1u << log2r) < rate_cats) ++log2r;
  const unsigned int pass = (DNA_THREADS >> log2r) * plf_dna_flow_unroll(max_sites);
  return (max_sites + pass - 1) / pass;
}

/* Cuts a level-sorted, plain op list (kinds II / TI / TT, contiguous CLVs, dep[] filled by the host layer) into
 * paths.  out_ops (nops entries) receives the ops path by path, bottom to top; out_start (nops + 1 entries) the
 * first op of each path.  Returns the number of paths, 0 when the list cannot run as one launch. */
extern "C" unsigned int plf_dna_flow_plan(const plf_op_t * h_ops, unsigned int nops, unsigned int path_max,
                                          plf_flow_op * out_ops, unsigned int * out_start)
{
  if (path_max < 1) path_max = 1;
  if (path_max > PLF_FLOW_PATH_MAX) path_max = PLF_FLOW_PATH_MAX;
  unsigned int * buf = (unsigned int *)malloc((size_t)nops * 6 * sizeof(unsigned int));
  if (!buf) return 0;
  unsigned int * readers = buf, * weight = buf + nops, * path_of = buf + 2 * (size_t)nops, * chain = buf + 3 * (size_t)nops;
  int * down = (int *)(buf + 4 * (size_t)nops); /* the child op whose parent this op keeps in registers, or -1 */
  unsigned char * taken = (unsigned char *)(buf + 5 * (size_t)nops); /* is some op's `down` */
  unsigned int npaths = 0, emitted = 0;
  int ok = 1;
  memset(readers, 0, (size_t)nops * sizeof(unsigned int));
  memset(taken, 0, nops);
  for (unsigned int i = 0; i < nops && ok; ++i)
  {
    const plf_op_t & o = h_ops[i];
    if (o.dep[2] != PLF_DEP_NONE || o.dep[3] != PLF_DEP_NONE) ok = 0; /* a scaler written by another op than the CLV's */
    for (int s = 0; s < 2 && ok; ++s)
      if (o.dep[s] != PLF_DEP_NONE)
      {
        if (o.dep[s] < 0 || (unsigned int)o.dep[s] >= i) ok = 0;
        else ++readers[o.dep[s]];
      }
  }
  for (unsigned int i = 0; i < nops && ok; ++i)
  {
    const plf_op_t & o = h_ops[i];
    weight[i] = 1;
    down[i] = -1;
    for (int s = 0; s < 2; ++s)
      if (o.dep[s] >= 0 && !(s == 1 && o.dep[1] == o.dep[0])) weight[i] += weight[o.dep[s]];
    if (path_max == 1) continue;
    for (int s = 0; s < 2; ++s)
    {
      const int d = o.dep[s];
      if (d < 0 || readers[d] != 1) continue;
      if (down[i] < 0 || weight[d] > weight[down[i]]) down[i] = d;
    }
    if (down[i] >= 0) taken[down[i]] = 1;
  }
  /* chains in the order of their last ops (an op nobody keeps in registers ends a chain), each cut bottom to top
   * into paths of <= path_max ops.  A lower path of a chain is read by the next path of the same chain only, which
   * follows it directly; everything else a chain reads ends before the chain's last op and was queued earlier. */
  for (unsigned int i = 0; i < nops && ok; ++i)
  {
    if (taken[i]) continue;
    unsigned int len = 0;
    for (int k = (int)i; k >= 0; k = down[k]) chain[len++] = (unsigned int)k; /* top to bottom */
    for (unsigned int b = 0; b < len; b += path_max)
    {
      const unsigned int e = (b + path_max < len) ? b + path_max : len;
      out_start[npaths] = emitted;
      for (unsigned int k = b; k < e; ++k, ++emitted) path_of[chain[len - 1 - k]] = npaths;
      ++npaths;
    }
  }
  if (ok && emitted != nops) ok = 0;
  if (ok)
  {
    unsigned int pos = 0;
    out_start[npaths] = nops;
    /* the same walk again, now that every op knows its path: the descriptors */
    unsigned int path = 0;
    for (unsigned int i = 0; i < nops; ++i)
    {
      if (taken[i]) continue;
      unsigned int len = 0;
      for (int k = (int)i; k >= 0; k = down[k]) chain[len++] = (unsigned int)k;
      for (unsigned int b = 0; b < len; b += path_max, ++path)
      {
        const unsigned int e = (b + path_max < len) ? b + path_max : len;
        for (unsigned int k = b; k < e; ++k, ++pos)
        {
          const unsigned int oi = chain[len - 1 - k];
          const plf_op_t & o = h_ops[oi];
          plf_flow_op & f = out_ops[pos];
          memset(&f, 0, sizeof(f));
          f.parent_clv = o.parent_clv;
          f.parent_scaler = o.parent_scaler;
          f.nsites = o.nsites;
          f.matrix[0] = o.left_matrix;
          f.matrix[1] = o.right_matrix;
          f.dep[0] = f.dep[1] = PLF_DEP_NONE;
          if (o.kind == PLF_OP_TT) f.flags |= FLOW_TIP_TIP;
          if (o.kind != PLF_OP_II) f.tip[0] = o.left_tip; else f.clv[0] = o.left_clv;
          if (o.kind == PLF_OP_TT) f.tip[1] = o.right_tip; else f.clv[1] = o.right_clv;
          if (o.kind == PLF_OP_II) f.scaler[0] = o.left_scaler;
          if (o.kind != PLF_OP_TT) f.scaler[1] = o.right_scaler;
          for (int s = 0; s < 2; ++s)
          {
            const int d = o.dep[s];
            if (d < 0) continue;
            /* which side reads that op's parent: the host layer may have swapped the children (tip first) */
            const double * produced = h_ops[d].parent_clv;
            int hit = 0;
            for (int side = 0; side < 2; ++side)
            {
              if (f.clv[side] != produced) continue;
              hit = 1;
              if (k > b && down[oi] == d && (f.flags & FLOW_CARRY_MASK) == 0)
              {
                f.flags |= (unsigned int)(side + 1);
                if (f.scaler[side] && f.scaler[side] == h_ops[d].parent_scaler)
                {
                  f.flags |= FLOW_CARRY_SCALER;
                  f.scaler[side] = nullptr;
                }
                break; /* the other side, should it read the same CLV, could not: readers == 1 */
              }
              f.dep[side] = (int)path_of[d];
            }
            if (!hit) ok = 0;
          }
        }
      }
    }
  }
  free(buf);
  return ok ? npaths : 0;
}

int plf_launch_dna_flow(plf_ctx * ctx, const plf_flow_op * d_fops, const unsigned int * d_path_start, unsigned int npaths,
                        unsigned int rate_cats, int per_rate, unsigned int max_sites, void * flow)
{
  int log2r = 0;
  while ((1u << log2r) < rate_cats) ++log2r;
  const unsigned int nchunks = plf_dna_flow_chunks(rate_cats, max_sites);
  typedef void (*flow_kernel_t)(const plf_flow_op *, const unsigned int *, unsigned int, unsigned int, int, plf_flow_ctrl *,
                                unsigned long long *);
  static const flow_kernel_t kernels[2][3] = {{k_clv_dna_flow<0, 1>, k_clv_dna_flow<1, 1>, k_clv_dna_flow<2, 1>},
                                              {k_clv_dna_flow<0, 2>, k_clv_dna_flow<1, 2>, k_clv_dna_flow<2, 2>}};
  const int two = plf_dna_flow_unroll(max_sites) == 2;
  const flow_kernel_t fn = kernels[two][log2r];
  if (!ctx->dna_flow_occupancy[two][log2r])
  {
    int per_sm = 0;
    PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, DNA_THREADS, 0));
    ctx->dna_flow_occupancy[two][log2r] = per_sm > 0 ? per_sm : 1;
  }
  const unsigned long long items = (unsigned long long)npaths * nchunks;
  unsigned long long grid = (unsigned long long)ctx->sm_count * ctx->dna_flow_occupancy[two][log2r];
  if (grid > items) grid = items;
  plf_flow_ctrl * ctrl = reinterpret_cast<plf_flow_ctrl *>(flow);
  unsigned long long * ready = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(flow) + 64);
  fn<<<(unsigned int)grid, DNA_THREADS, 0, ctx->stream>>>(d_fops, d_path_start, npaths, nchunks, per_rate, ctrl, ready);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* ------------------------------------------------------------------------ */

typedef void (*dna_kernel_t)(const plf_op_t *, int);

template <int LOG2R>
static dna_kernel_t pick_kernel(unsigned int kind)
{
  if (kind == PLF_OP_II) return k_clv_dna_ii<LOG2R, 4>;
  if (kind == PLF_OP_TI) return k_clv_dna_ti<LOG2R, 4>;
  return k_clv_dna_tt<LOG2R, 4>;
}

static const int DNA_UNROLL = 4;

template <int LOG2R, int NSTAGE, int ITEMS>
static dna_kernel_t pick_stream_kernel(unsigned int kind)
{
  if (kind == PLF_OP_II) return k_clv_dna_stream<LOG2R, CK_I, CK_I, NSTAGE, ITEMS>;
  return k_clv_dna_stream<LOG2R, CK_T, CK_I, NSTAGE, ITEMS>;
}

template <int LOG2R, int ITEMS>
static dna_kernel_t pick_stream_kernel_stages(unsigned int kind, int nstage, size_t * smem, unsigned int * tile)
{
  *smem = kind == PLF_OP_II ? StreamLayout<LOG2R, CK_I, CK_I, ITEMS>::smem_bytes(nstage)
                            : StreamLayout<LOG2R, CK_T, CK_I, ITEMS>::smem_bytes(nstage);
  *tile = StreamLayout<LOG2R, CK_I, CK_I, ITEMS>::TILE;
  switch (nstage)
  {
    case 2: return pick_stream_kernel<LOG2R, 2, ITEMS>(kind);
    case 3: return pick_stream_kernel<LOG2R, 3, ITEMS>(kind);
    case 4: return pick_stream_kernel<LOG2R, 4, ITEMS>(kind);
    default: return pick_stream_kernel<LOG2R, 6, ITEMS>(kind);
  }
}

template <int LOG2R>
static dna_kernel_t pick_stream_kernel_items(unsigned int kind, int nstage, int items, size_t * smem, unsigned int * tile)
{
  if (items == 2) return pick_stream_kernel_stages<LOG2R, 2>(kind, nstage, smem, tile);
  if (items == 1) return pick_stream_kernel_stages<LOG2R, 1>(kind, nstage, smem, tile);
  return pick_stream_kernel_stages<LOG2R, 4>(kind, nstage, smem, tile);
}

/* consumers of virtual cherries (rate_cats <= 4): ring of 6 (default) or 4 stages; `items` 2 or 4 blocks per
 * thread at 128 threads, or (items == 1) one block per thread at 256 threads: the same 64-site tiles and the same
 * shared memory per CTA with twice the warps */
template <int LOG2R, int LK, int RK, int NSTAGE>
static dna_kernel_t pick_cherry_items(int items, size_t * smem, unsigned int * tile, unsigned int * threads)
{
  *threads = DNA_THREADS;
  if (items == 4)
  {
    *smem = StreamLayout<LOG2R, LK, RK, 4>::smem_bytes(NSTAGE);
    *tile = StreamLayout<LOG2R, LK, RK, 4>::TILE;
    return k_clv_dna_stream<LOG2R, LK, RK, NSTAGE, 4>;
  }
  if (items == 1)
  {
    *threads = 256;
    *smem = StreamLayout<LOG2R, LK, RK, 1, 256>::smem_bytes(NSTAGE);
    *tile = StreamLayout<LOG2R, LK, RK, 1, 256>::TILE;
    return k_clv_dna_stream<LOG2R, LK, RK, NSTAGE, 1, 256>;
  }
  *smem = StreamLayout<LOG2R, LK, RK, 2>::smem_bytes(NSTAGE);
  *tile = StreamLayout<LOG2R, LK, RK, 2>::TILE;
  return k_clv_dna_stream<LOG2R, LK, RK, NSTAGE, 2>;
}

template <int LOG2R, int NSTAGE>
static dna_kernel_t pick_cherry_kind(unsigned int kind, int items, size_t * smem, unsigned int * tile, unsigned int * threads)
{
  if (kind == PLF_OP_CI) return pick_cherry_items<LOG2R, CK_C, CK_I, NSTAGE>(items, smem, tile, threads);
  if (kind == PLF_OP_TC) return pick_cherry_items<LOG2R, CK_T, CK_C, NSTAGE>(items, smem, tile, threads);
  return pick_cherry_items<LOG2R, CK_C, CK_C, NSTAGE>(items, smem, tile, threads);
}

template <int LOG2R>
static dna_kernel_t pick_cherry_kernel(unsigned int kind, int items, int stages, size_t * smem, unsigned int * tile,
                                       unsigned int * threads)
{
  if (stages == 4) return pick_cherry_kind<LOG2R, 4>(kind, items, smem, tile, threads);
  return pick_cherry_kind<LOG2R, 6>(kind, items, smem, tile, threads);
}

static int env_int(const char * name, int dflt)
{
  const char * v = getenv(name);
  return (v && v[0]) ? atoi(v) : dflt;
}

/* the A/B switches of the 4-state kernels, read from the environment on first use */
static void dna_read_switches(plf_ctx * ctx)
{
  if (ctx->dna_stream >= 0) return;
  ctx->dna_stream = env_int("PLF_DNA_STREAM", 1);
  ctx->dna_stages = env_int("PLF_DNA_STAGES", 6);
  ctx->dna_items = env_int("PLF_DNA_ITEMS", 2);
  ctx->dna_tt_bulk = env_int("PLF_TT_BULK", 1);
  ctx->dna_tt_items = env_int("PLF_TT_ITEMS", 2) == 4 ? 4 : 2;
  ctx->dna_tt_seq = env_int("PLF_TT_SEQ", 1);
  ctx->dna_balanced = env_int("PLF_DNA_BALANCED", 1);
  /* default 1: one block per thread at 256 threads per CTA (2.83 vs 2.89 ms per config-2 traversal) */
  ctx->dna_cherry_items = env_int("PLF_CHERRY_ITEMS", 1);
  if (ctx->dna_cherry_items != 2 && ctx->dna_cherry_items != 4) ctx->dna_cherry_items = 1;
  ctx->dna_cherry_stages = env_int("PLF_CHERRY_STAGES", 6) == 4 ? 4 : 6;
  ctx->dna_cherry_bulk = env_int("PLF_CHERRY_BULK", 0); /* measured slower than the ring kernel: profiles/r2_notes.md */
  if (ctx->dna_stages != 2 && ctx->dna_stages != 3 && ctx->dna_stages != 4) ctx->dna_stages = 6;
}

/* launch one group of same-kind DNA ops (rate_cats a power of two <= 32) as a
 * single persistent wave: gridDim.y = ops, gridDim.x = CTAs striding over the
 * sites of each op.  `contiguous`: no op of the group gathers through repeat
 * identifiers, so the bulk-copy streaming kernels apply. */
typedef void (*dna_balanced_kernel_t)(const plf_op_t *, int, const unsigned int *);

/* tiles of one op for the balanced inner-inner kernel (DNA_THREADS x 4 items per tile) */
unsigned int plf_dna_balanced_tiles(unsigned int nsites, unsigned int rate_cats)
{
  const unsigned int tile_sites = (DNA_THREADS * 4u) / rate_cats;
  return (nsites + tile_sites - 1) / tile_sites;
}

int plf_launch_dna_group(plf_ctx * ctx, const plf_op_t * d_ops, unsigned int nops, unsigned int kind,
                         unsigned int rate_cats, int per_rate, unsigned int max_sites, int contiguous,
                         const unsigned int * d_tile_prefix, unsigned int total_tiles, int pair_lists)
{
  int log2r = 0;
  while ((1u << log2r) < rate_cats) ++log2r;
  dna_read_switches(ctx);

  if (kind == PLF_OP_TT_VIRTUAL)
  {
    /* snapshot of the cherries' P-matrices + zeroed scalers: no CLV is written */
    const unsigned long long entries = (unsigned long long)max_sites * (per_rate ? rate_cats : 1u);
    unsigned long long bx = (entries / 4 + 255) / 256;
    const unsigned long long cap = ((unsigned long long)ctx->sm_count * 8 + nops - 1) / nops;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    k_cherry_prepare<<<dim3((unsigned int)bx, nops), 256, 0, ctx->stream>>>(d_ops, per_rate, (int)rate_cats);
    plf_count_launch();
    PLF_CHECK(ctx, cudaGetLastError());
    return 1;
  }
  if (kind == PLF_OP_CI || kind == PLF_OP_TC || kind == PLF_OP_CC)
  {
    if (!contiguous || !ctx->dna_stream || log2r > 2)
    {
      plf_set_error(ctx, "virtual cherry consumers need the contiguous 4-state streaming path (rate_cats <= 4)");
      return 0;
    }
    dna_kernel_t k = nullptr;
    size_t smem = 0;
    unsigned int tile = 0;
    if (kind != PLF_OP_CI && ctx->dna_cherry_bulk)
    {
      /* no CLV comes in: tiles built in shared memory and sent off by bulk stores */
      const int lc = (kind == PLF_OP_CC);
      switch (log2r)
      {
        case 0: k = lc ? k_clv_dna_lookup_bulk<0, 2, 4, 1> : k_clv_dna_lookup_bulk<0, 2, 4, 0>; break;
        case 1: k = lc ? k_clv_dna_lookup_bulk<1, 2, 4, 1> : k_clv_dna_lookup_bulk<1, 2, 4, 0>; break;
        default: k = lc ? k_clv_dna_lookup_bulk<2, 2, 4, 1> : k_clv_dna_lookup_bulk<2, 2, 4, 0>; break;
      }
      smem = (size_t)4 * 8192 + ((size_t)(lc ? 1024 : 64) + 1024 + 128) * rate_cats * sizeof(double);
      tile = (DNA_THREADS * 2u) >> log2r;
      int & occ = ctx->dna_cherry_occupancy[kind - PLF_OP_CI][log2r];
      if (!occ)
      {
        PLF_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, DNA_THREADS, smem));
        if (occ < 1) occ = 1;
      }
      const unsigned long long ntiles = ((unsigned long long)max_sites + tile - 1) / tile;
      unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
      if (bx < 1) bx = 1;
      if (bx > ntiles) bx = ntiles;
      k<<<dim3((unsigned int)bx, nops), DNA_THREADS, smem, ctx->stream>>>(d_ops, per_rate);
      plf_count_launch();
      PLF_CHECK(ctx, cudaGetLastError());
      return 1;
    }
    unsigned int threads = DNA_THREADS;
    switch (log2r)
    {
      case 0: k = pick_cherry_kernel<0>(kind, ctx->dna_cherry_items, ctx->dna_cherry_stages, &smem, &tile, &threads); break;
      case 1: k = pick_cherry_kernel<1>(kind, ctx->dna_cherry_items, ctx->dna_cherry_stages, &smem, &tile, &threads); break;
      default: k = pick_cherry_kernel<2>(kind, ctx->dna_cherry_items, ctx->dna_cherry_stages, &smem, &tile, &threads); break;
    }
    int & occ = ctx->dna_cherry_occupancy[kind - PLF_OP_CI][log2r];
    if (!occ)
    {
      PLF_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, (int)threads, smem));
      if (occ < 1) occ = 1;
    }
    const unsigned long long ntiles = ((unsigned long long)max_sites + tile - 1) / tile;
    unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
    if (bx < 1) bx = 1;
    if (bx > ntiles) bx = ntiles;
    k<<<dim3((unsigned int)bx, nops), threads, smem, ctx->stream>>>(d_ops, per_rate);
    plf_count_launch();
    PLF_CHECK(ctx, cudaGetLastError());
    return 1;
  }

  /* the ring copies tip codes and scalers in 16-byte granules from tile-aligned offsets: a tile must
   * hold at least 16 sites (32 rate categories with 2 items per thread would give 8) */
  if (contiguous && ctx->dna_stream && kind != PLF_OP_TT && ((DNA_THREADS * (unsigned int)ctx->dna_items) >> log2r) >= 16)
  {
    dna_kernel_t k = nullptr;
    size_t smem = 0;
    unsigned int tile = 0;
    switch (log2r)
    {
      case 0: k = pick_stream_kernel_items<0>(kind, ctx->dna_stages, ctx->dna_items, &smem, &tile); break;
      case 1: k = pick_stream_kernel_items<1>(kind, ctx->dna_stages, ctx->dna_items, &smem, &tile); break;
      case 2: k = pick_stream_kernel_items<2>(kind, ctx->dna_stages, ctx->dna_items, &smem, &tile); break;
      case 3: k = pick_stream_kernel_items<3>(kind, ctx->dna_stages, ctx->dna_items, &smem, &tile); break;
      case 4: k = pick_stream_kernel_items<4>(kind, ctx->dna_stages, ctx->dna_items, &smem, &tile); break;
      default: k = pick_stream_kernel_items<5>(kind, ctx->dna_stages, ctx->dna_items, &smem, &tile); break;
    }
    int & occ = ctx->dna_stream_occupancy[kind == PLF_OP_II ? 0 : 1][log2r];
    if (!occ)
    {
      PLF_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, DNA_THREADS, smem));
      if (occ < 1) occ = 1;
    }
    const unsigned long long ntiles = ((unsigned long long)max_sites + tile - 1) / tile;
    unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
    if (bx < 1) bx = 1;
    if (bx > ntiles) bx = ntiles;
    dim3 grid((unsigned int)bx, nops);
    k<<<grid, DNA_THREADS, smem, ctx->stream>>>(d_ops, per_rate);
    plf_count_launch();
    PLF_CHECK(ctx, cudaGetLastError());
    return 1;
  }

  if (kind == PLF_OP_TT && contiguous && ctx->dna_stream && log2r <= 3 && ctx->dna_tt_bulk)
  {
    const int items = ctx->dna_tt_items;
    dna_kernel_t kb = nullptr;
    switch (log2r)
    {
      case 0: kb = items == 4 ? k_clv_dna_tt_bulk<0, 4, 3> : k_clv_dna_tt_bulk<0, 2, 4>; break;
      case 1: kb = items == 4 ? k_clv_dna_tt_bulk<1, 4, 3> : k_clv_dna_tt_bulk<1, 2, 4>; break;
      case 2: kb = items == 4 ? k_clv_dna_tt_bulk<2, 4, 3> : k_clv_dna_tt_bulk<2, 2, 4>; break;
      default: kb = items == 4 ? k_clv_dna_tt_bulk<3, 4, 3> : k_clv_dna_tt_bulk<3, 2, 4>; break;
    }
    const size_t smem = items == 4 ? (size_t)3 * 16384 : (size_t)4 * 8192;
    const unsigned int tile = (DNA_THREADS * items) >> log2r;
    int & occ = ctx->dna_tt_bulk_occupancy[log2r];
    if (!occ)
    {
      PLF_CHECK(ctx, cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb, DNA_THREADS, smem));
      if (occ < 1) occ = 1;
    }
    const unsigned long long ntiles = ((unsigned long long)max_sites + tile - 1) / tile;
    unsigned long long all = (unsigned long long)ctx->sm_count * occ;
    if (all > ntiles) all = ntiles;
    /* the whole grid sweeps one op after the other (few open DRAM pages) */
    kb<<<dim3((unsigned int)(all < 1 ? 1 : all), 1), DNA_THREADS, smem, ctx->stream>>>(
        d_ops, (per_rate & 1) | (int)(nops << 1));
    plf_count_launch();
    PLF_CHECK(ctx, cudaGetLastError());
    return 1;
  }

  if (kind == PLF_OP_II && !contiguous && d_tile_prefix && total_tiles && ctx->dna_balanced)
  {
    dna_balanced_kernel_t kb = nullptr;
    switch (log2r)
    {
      case 0: kb = k_clv_dna_ii_balanced<0, 4>; break;
      case 1: kb = k_clv_dna_ii_balanced<1, 4>; break;
      case 2: kb = k_clv_dna_ii_balanced<2, 4>; break;
      case 3: kb = k_clv_dna_ii_balanced<3, 4>; break;
      case 4: kb = k_clv_dna_ii_balanced<4, 4>; break;
      default: kb = k_clv_dna_ii_balanced<5, 4>; break;
    }
    if (ctx->dna_balanced != 9) /* PLF_DNA_BALANCED=9 keeps the P-matrices in registers (3 CTAs per SM) for A/B runs */
      switch (log2r)
      {
        case 0: kb = k_clv_dna_ii_balanced_sm<0, 2, 8>; break;
        case 1: kb = k_clv_dna_ii_balanced_sm<1, 2, 8>; break;
        case 2: kb = k_clv_dna_ii_balanced_sm<2, 2, 8>; break;
        case 3: kb = k_clv_dna_ii_balanced_sm<3, 2, 8>; break;
        case 4: kb = k_clv_dna_ii_balanced_sm<4, 2, 8>; break;
        default: kb = k_clv_dna_ii_balanced_sm<5, 2, 8>; break;
      }
    if (pair_lists && ctx->dna_balanced != 9 && ctx->dna_balanced != 8) /* 8: the three-array gather for A/B runs */
    {
      /* items in flight x resident CTAs: 26 = 2 x 6 (80 registers; default: config 4 traversal 1.19 ms), 28 = 2 x 8
       * (64 registers, spills 72 bytes: 1.31 ms), 45 = 4 x 5 (96 registers), 44 = 4 x 4 (128 registers) */
      const int var = env_int("PLF_PAIRS_SHAPE", 26);
#define PAIRS_PICK(L) (var == 28 ? k_clv_dna_ii_pairs<L, 2, 8> : var == 45 ? k_clv_dna_ii_pairs<L, 4, 5> : \
                       var == 44 ? k_clv_dna_ii_pairs<L, 4, 4> : k_clv_dna_ii_pairs<L, 2, 6>)
      switch (log2r)
      {
        case 0: kb = PAIRS_PICK(0); break;
        case 1: kb = PAIRS_PICK(1); break;
        case 2: kb = PAIRS_PICK(2); break;
        case 3: kb = PAIRS_PICK(3); break;
        case 4: kb = PAIRS_PICK(4); break;
        default: kb = PAIRS_PICK(5); break;
      }
#undef PAIRS_PICK
    }
    int & occ = ctx->dna_balanced_occupancy[pair_lists ? 1 : 0][log2r];
    if (!occ)
    {
      PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb, DNA_THREADS, 0));
      if (occ < 1) occ = 1;
    }
    unsigned int grid = (unsigned int)ctx->sm_count * occ;
    if (grid > total_tiles) grid = total_tiles;
    kb<<<grid, DNA_THREADS, 0, ctx->stream>>>(d_ops, (per_rate & 1) | (int)(nops << 1), d_tile_prefix);
    plf_count_launch();
    PLF_CHECK(ctx, cudaGetLastError());
    return 1;
  }

  dna_kernel_t k = nullptr;
  switch (log2r)
  {
    case 0: k = pick_kernel<0>(kind); break;
    case 1: k = pick_kernel<1>(kind); break;
    case 2: k = pick_kernel<2>(kind); break;
    case 3: k = pick_kernel<3>(kind); break;
    case 4: k = pick_kernel<4>(kind); break;
    default: k = pick_kernel<5>(kind); break;
  }
  int & occ = ctx->dna_occupancy[kind < 3 ? kind : 0][log2r];
  if (!occ)
  {
    PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, DNA_THREADS, 0));
    if (occ < 1) occ = 1;
  }
  const unsigned long long lanes = (unsigned long long)max_sites << log2r;
  const unsigned int per_thread = (kind == PLF_OP_TT) ? TT_GROUP : DNA_UNROLL;
  unsigned long long need = (lanes + (unsigned long long)DNA_THREADS * per_thread - 1) / ((unsigned long long)DNA_THREADS * per_thread);
  unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
  if (bx < 1) bx = 1;
  if (bx > need) bx = need;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned int)bx, nops);
  int arg = per_rate;
  if (kind == PLF_OP_TT)
  {
    arg = (per_rate & 1) | (int)(nops << 1);
    if (ctx->dna_tt_seq)
    {
      unsigned long long all = (unsigned long long)ctx->sm_count * occ;
      if (all > need) all = need;
      grid = dim3((unsigned int)(all < 1 ? 1 : all), 1);
    }
  }
  k<<<grid, DNA_THREADS, 0, ctx->stream>>>(d_ops, arg);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* do the kernels that consume virtual cherries serve this shape? (the host layer asks before it leaves a
 * tip-tip parent unwritten) */
int plf_dna_virtual_cherries_supported(plf_ctx * ctx, const plf_shape_t * sh)
{
  dna_read_switches(ctx);
  if (sh->states != 4 || !ctx->dna_stream) return 0;
  if (sh->rate_cats != 1 && sh->rate_cats != 2 && sh->rate_cats != 4) return 0;
  /* consumers that are not cherry-fed go through the ring kernels too: their tiles need >= 16 sites */
  return ((DNA_THREADS * (unsigned int)ctx->dna_items) / sh->rate_cats) >= 16;
}
