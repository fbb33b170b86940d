/*
 * plf_parsimony.cu -- Fitch parsimony on packed bit vectors, sm_100a.
 *
 * Reference: src/fast_parsimony.c (informative sites :128-199,381-413; bit packing :201-379; vector update
 * :461-530 / generic :572-625; edge score :415-459 / generic :627-662) and its SSE/AVX/AVX2 variants, which
 * compute the same integers 4 or 8 words at a time.
 *
 * A node's vector is `states` rows of `words` 32-bit words: bit b of row k is set when state k is possible
 * at (weighted) site b.  All work is integer and HBM/L2-bound; nothing here is GEMM-shaped.
 *
 *   k_pars_informative   one thread per site: 256 two-bit saturating counters in shared memory count how many
 *                        tips carry each tip code (0, 1, "2 or more"), which is all the reference's 256-entry
 *                        histogram is used for
 *   k_pars_informative_wide   > 8 states without pattern tips (codes up to 2^20): O(tips^2) compares per site
 *   k_pars_pack          one thread per (tip, output word): binary search of the word's first site in the
 *                        prefix sum of site widths, then at most 32 sites are OR-ed in
 *   k_pars_update        one thread per word column runs the WHOLE operation list: columns are independent and a
 *                        thread reads back only what it wrote itself, so a traversal is one launch with no
 *                        grid-wide dependency; popcounts are warp-reduced and added with integer atomics
 *                        (order-independent, hence bit-reproducible)
 *   k_pars_edge_scores   a batch of edge scores in one launch (gridDim.y = edges)
 *   k_pars_insert_scan   stepwise addition: for every candidate edge (a,b) the Fitch parent of a and b is formed
 *                        in registers and scored against the subtree to insert; src/stepwise.c:436-530 does this
 *                        with one vector update + one edge score per edge
 */
#include <cub/cub.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "plf_internal.h"

struct plf_pars
{
  plf_ctx * ctx;
  unsigned int * d_small; /* op / pair lists followed by the score cells */
  size_t small_cap;       /* in uints */
  unsigned int * h_pin;   /* pinned staging, same capacity */
};

extern "C" int plf_pars_create(plf_ctx_t * ctx, plf_pars_t ** out)
{
  plf_pars * ps = (plf_pars *)calloc(1, sizeof(plf_pars));
  if (!ps) return 0;
  ps->ctx = ctx;
  *out = ps;
  return 1;
}

extern "C" void plf_pars_destroy(plf_pars_t * ps)
{
  if (!ps) return;
  cudaSetDevice(ps->ctx->device);
  cudaStreamSynchronize(ps->ctx->stream);
  cudaFree(ps->d_small);
  cudaFreeHost(ps->h_pin);
  free(ps);
}

static int pars_reserve(plf_pars * ps, size_t uints)
{
  if (ps->small_cap >= uints) return 1;
  plf_ctx * ctx = ps->ctx;
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ps->d_small);
  cudaFreeHost(ps->h_pin);
  ps->d_small = NULL;
  ps->h_pin = NULL;
  ps->small_cap = 0;
  size_t want = uints < 4096 ? 4096 : uints + uints / 2;
  PLF_CHECK(ctx, cudaMalloc((void **)&ps->d_small, want * sizeof(unsigned int)));
  PLF_CHECK(ctx, cudaMallocHost((void **)&ps->h_pin, want * sizeof(unsigned int)));
  ps->small_cap = want;
  return 1;
}

/* ---- tip codes ------------------------------------------------------------------------------------------ */

struct pars_tips_dev
{
  unsigned int tips, sites, states, span; /* span = states_padded * rate_cats doubles per CLV entry */
  const unsigned char * const * tipchars;
  const double * const * tipclv;
  const unsigned int * const * site_id;
  const unsigned long long * tipmap;
};

/* set of states of tip `t` at `site` as a bit mask (bit k = state k), fast_parsimony.c:282-300 */
__device__ __forceinline__ unsigned long long pars_tip_mask(const pars_tips_dev & a, unsigned int t, unsigned int site)
{
  if (a.tipchars)
  {
    unsigned long long c = a.tipchars[t][site];
    if (a.states != 4) c = a.tipmap[c];
    return c;
  }
  unsigned int entry = site;
  if (a.site_id)
  {
    const unsigned int * ids = a.site_id[t];
    if (ids) entry = ids[site];
  }
  const double * clv = a.tipclv[t] + (size_t)entry * a.span;
  unsigned long long c = 0;
  for (unsigned int k = 0; k < a.states; ++k)
    if ((int)clv[k]) c |= 1ull << k;
  return c;
}

/* ---- informative sites ---------------------------------------------------------------------------------- */

#define PARS_INF_THREADS 128

/* codes < 256: pattern-tip codes, or state masks of <= 8 states.  out: informative flag, the site's width in
 * bits (weight when informative, else 0) and its constant cost (singletons x weight when not informative). */
__global__ void __launch_bounds__(PARS_INF_THREADS)
k_pars_informative(pars_tips_dev a, const unsigned int * __restrict__ weights, int * __restrict__ informative,
                   unsigned int * __restrict__ width, unsigned int * __restrict__ ccost)
{
  __shared__ unsigned int cnt[16][PARS_INF_THREADS]; /* 2-bit saturating counters, 16 per word */
  const unsigned int site = blockIdx.x * PARS_INF_THREADS + threadIdx.x;
  if (site >= a.sites) return;
#pragma unroll
  for (int w = 0; w < 16; ++w) cnt[w][threadIdx.x] = 0;
  for (unsigned int t = 0; t < a.tips; ++t)
  {
    unsigned int c;
    if (a.tipchars)
      c = a.tipchars[t][site];
    else
      c = (unsigned int)pars_tip_mask(a, t, site) & 255u; /* distinct masks <-> distinct reference codes */
    const unsigned int w = c >> 4, sh = (c & 15u) * 2u;
    const unsigned int v = cnt[w][threadIdx.x];
    if (((v >> sh) & 3u) < 2u) cnt[w][threadIdx.x] = v + (1u << sh);
  }
  unsigned int multi = 0, single = 0;
#pragma unroll
  for (int w = 0; w < 16; ++w)
  {
    const unsigned int v = cnt[w][threadIdx.x];
    const unsigned int lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
    multi += __popc(hi);
    single += __popc(lo & ~hi);
  }
  const int inf = multi > 1;
  const unsigned int wgt = weights[site];
  informative[site] = inf;
  width[site] = inf ? wgt : 0u;
  ccost[site] = inf ? 0u : single * wgt;
}

__global__ void k_pars_codes(pars_tips_dev a, unsigned int * __restrict__ codes)
{
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int t = blockIdx.y;
  if (site < a.sites) codes[(size_t)t * a.sites + site] = (unsigned int)pars_tip_mask(a, t, site);
}

__global__ void k_pars_informative_wide(const unsigned int * __restrict__ codes, unsigned int tips, unsigned int sites,
                                        const unsigned int * __restrict__ weights, int * __restrict__ informative,
                                        unsigned int * __restrict__ width, unsigned int * __restrict__ ccost)
{
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= sites) return;
  unsigned int multi = 0, single = 0;
  for (unsigned int i = 0; i < tips; ++i)
  {
    const unsigned int c = codes[(size_t)i * sites + site];
    unsigned int n = 0;
    bool first = true;
    for (unsigned int j = 0; j < tips; ++j)
      if (codes[(size_t)j * sites + site] == c)
      {
        ++n;
        if (j < i) first = false;
      }
    if (first)
    {
      if (n > 1) ++multi;
      else ++single;
    }
  }
  const int inf = multi > 1;
  const unsigned int wgt = weights[site];
  informative[site] = inf;
  width[site] = inf ? wgt : 0u;
  ccost[site] = inf ? 0u : single * wgt;
}

extern "C" int plf_pars_informative(plf_pars_t * ps, const plf_pars_tips_t * tp, int * d_informative,
                                    unsigned int * d_bitpos, unsigned int * h_bitcount, unsigned int * h_const_cost,
                                    unsigned int * h_informative_count)
{
  plf_ctx * ctx = ps->ctx;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const unsigned int S = tp->sites;
  pars_tips_dev a;
  a.tips = tp->tips;
  a.sites = S;
  a.states = tp->states;
  a.span = tp->states_padded * tp->rate_cats;
  a.tipchars = tp->d_tipchars;
  a.tipclv = tp->d_tipclv;
  a.site_id = tp->d_tip_site_id;
  a.tipmap = tp->d_tipmap;

  unsigned int * d_width = NULL, * d_cc = NULL, * d_codes = NULL, * d_sums = NULL;
  void * d_tmp = NULL;
  size_t tmp_bytes = 0, tmp2 = 0;
  int ok = 0;
  const bool wide = !tp->d_tipchars && tp->states > 8;
  cudaError_t e = cudaMalloc((void **)&d_width, ((size_t)S + 1) * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_cc, ((size_t)S + 1) * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_sums, 4 * sizeof(unsigned int));
  if (e == cudaSuccess && wide) e = cudaMalloc((void **)&d_codes, (size_t)S * tp->tips * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_width + S, 0, sizeof(unsigned int), ctx->stream);
  if (e != cudaSuccess) goto fail;
  if (S)
  {
    if (wide)
    {
      k_pars_codes<<<dim3((S + 255) / 256, tp->tips), 256, 0, ctx->stream>>>(a, d_codes);
      k_pars_informative_wide<<<(S + 127) / 128, 128, 0, ctx->stream>>>(d_codes, tp->tips, S, tp->d_weights,
                                                                        d_informative, d_width, d_cc);
      plf_count_launches(2);
    }
    else
    {
      k_pars_informative<<<(S + PARS_INF_THREADS - 1) / PARS_INF_THREADS, PARS_INF_THREADS, 0, ctx->stream>>>(
          a, tp->d_weights, d_informative, d_width, d_cc);
      plf_count_launch();
    }
  }
  /* bit position of every site = exclusive prefix sum of the widths (S+1 entries: the last one is the total) */
  cub::DeviceScan::ExclusiveSum(NULL, tmp_bytes, d_width, d_bitpos, (int)(S + 1), ctx->stream);
  cub::DeviceReduce::Sum(NULL, tmp2, d_cc, d_sums, (int)S, ctx->stream);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  cub::DeviceReduce::Sum(NULL, tmp2, d_informative, (int *)d_sums + 1, (int)S, ctx->stream);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  e = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
  if (e != cudaSuccess) goto fail;
  e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_width, d_bitpos, (int)(S + 1), ctx->stream);
  if (e == cudaSuccess) e = cub::DeviceReduce::Sum(d_tmp, tmp_bytes, d_cc, d_sums, (int)S, ctx->stream);
  if (e == cudaSuccess) e = cub::DeviceReduce::Sum(d_tmp, tmp_bytes, d_informative, (int *)d_sums + 1, (int)S, ctx->stream);
  plf_count_launches(3);
  {
    unsigned int h[2] = {0, 0}, total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_sums, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_bitpos + S, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) goto fail;
    *h_const_cost = h[0];
    *h_informative_count = h[1];
    *h_bitcount = total;
  }
  ok = 1;
fail:
  if (!ok) plf_set_error(ctx, "parsimony informative-site pass failed: %s", cudaGetErrorString(e));
  cudaFree(d_width);
  cudaFree(d_cc);
  cudaFree(d_codes);
  cudaFree(d_sums);
  cudaFree(d_tmp);
  return ok;
}

/* ---- bit packing ---------------------------------------------------------------------------------------- */

/* bits [s, e) of the output word that site j covers, 0 when it covers none; stop = the walk is past the word */
__device__ __forceinline__ unsigned int pars_site_mask(const unsigned int * __restrict__ bitpos, unsigned int j,
                                                       unsigned long long lo, bool & stop)
{
  const unsigned long long b0 = bitpos[j], b1 = bitpos[j + 1];
  stop = b0 >= lo + 32u;
  if (stop || b1 == b0) return 0u;
  const unsigned int s = (unsigned int)((b0 > lo ? b0 : lo) - lo);
  const unsigned int e = (unsigned int)((b1 < lo + 32u ? b1 : lo + 32u) - lo); /* 1..32 */
  const unsigned int upto = e == 32u ? ~0u : ((1u << e) - 1u);
  return upto & ~((1u << s) - 1u);
}

template <int ST>
__global__ void __launch_bounds__(128)
k_pars_pack(pars_tips_dev a, const unsigned int * __restrict__ bitpos, unsigned int bitcount, unsigned int words,
            unsigned int * __restrict__ vec)
{
  const unsigned int word = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int t = blockIdx.y;
  if (word >= words) return;
  unsigned int * out = vec + (size_t)t * a.states * words + word;
  const unsigned long long lo = (unsigned long long)word * 32u;
  if (lo >= bitcount)
  {
    /* padding words are all ones (fast_parsimony.c:336-339) */
    for (unsigned int k = 0; k < a.states; ++k) out[(size_t)k * words] = ~0u;
    return;
  }
  /* first site whose bit range ends after `lo` */
  unsigned int l = 0, r = a.sites;
  while (l < r)
  {
    const unsigned int m = l + ((r - l) >> 1);
    if (bitpos[m + 1] > lo) r = m;
    else l = m + 1;
  }
  /* the unused tail of the last word is filled with ones (fast_parsimony.c:323-334) */
  const unsigned int pad = (bitcount - lo < 32u) ? ~((1u << (unsigned int)(bitcount - lo)) - 1u) : 0u;
  if constexpr (ST > 0)
  {
    constexpr int N = ST > 0 ? ST : 1;
    unsigned int v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = pad;
    for (unsigned int j = l; j < a.sites; ++j)
    {
      bool stop;
      const unsigned int m = pars_site_mask(bitpos, j, lo, stop);
      if (stop) break;
      if (!m) continue;
      const unsigned long long c = pars_tip_mask(a, t, j);
#pragma unroll
      for (int k = 0; k < N; ++k)
        if ((c >> k) & 1ull) v[k] |= m;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) out[(size_t)k * words] = v[k];
  }
  else /* any state count: one walk over the (at most 32) sites of the word per state row */
  for (unsigned int k = 0; k < a.states; ++k)
  {
    unsigned int v = pad;
    for (unsigned int j = l; j < a.sites; ++j)
    {
      bool stop;
      const unsigned int m = pars_site_mask(bitpos, j, lo, stop);
      if (stop) break;
      if (m && ((pars_tip_mask(a, t, j) >> k) & 1ull)) v |= m;
    }
    out[(size_t)k * words] = v;
  }
}

extern "C" int plf_pars_pack(plf_pars_t * ps, const plf_pars_tips_t * tp, const unsigned int * d_bitpos,
                             unsigned int bitcount, unsigned int words, unsigned int * d_vec)
{
  plf_ctx * ctx = ps->ctx;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!words || !tp->tips) return 1;
  pars_tips_dev a;
  a.tips = tp->tips;
  a.sites = tp->sites;
  a.states = tp->states;
  a.span = tp->states_padded * tp->rate_cats;
  a.tipchars = tp->d_tipchars;
  a.tipclv = tp->d_tipclv;
  a.site_id = tp->d_tip_site_id;
  a.tipmap = tp->d_tipmap;
  const dim3 grid((words + 127) / 128, tp->tips);
  if (tp->states == 4)
    k_pars_pack<4><<<grid, 128, 0, ctx->stream>>>(a, d_bitpos, bitcount, words, d_vec);
  else if (tp->states == 20)
    k_pars_pack<20><<<grid, 128, 0, ctx->stream>>>(a, d_bitpos, bitcount, words, d_vec);
  else
    k_pars_pack<0><<<grid, 128, 0, ctx->stream>>>(a, d_bitpos, bitcount, words, d_vec);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* ---- vector updates ------------------------------------------------------------------------------------- */

#define PARS_THREADS 128

/* ST > 0: the state count is known at compile time and both children stay in registers.  The list is walked
 * with one operation of look-ahead: the children of the next operation are fetched before the current one is
 * finished, except a child that IS the current parent (the usual case in a post-order list: a subtree root is
 * consumed right after it was made), which is handed over in registers.  The dependent store -> load round trip
 * through L2 disappears from the chain and the remaining loads overlap the current operation. */
template <int ST>
__global__ void __launch_bounds__(PARS_THREADS)
k_pars_update(unsigned int * vec, size_t node_stride, unsigned int states, unsigned int words,
              const unsigned int * __restrict__ ops, unsigned int count, unsigned int * scores)
{
  const unsigned int word = blockIdx.x * PARS_THREADS + threadIdx.x;
  const bool active = word < words;
  const unsigned int w = active ? word : 0;
  if constexpr (ST > 0)
  {
    constexpr int N = ST > 0 ? ST : 1;
    unsigned int x[N], y[N], nx[N], ny[N];
    {
      const unsigned int * c1 = vec + ops[1] * node_stride + w;
      const unsigned int * c2 = vec + ops[2] * node_stride + w;
#pragma unroll
      for (int j = 0; j < N; ++j)
      {
        x[j] = c1[(size_t)j * words];
        y[j] = c2[(size_t)j * words];
      }
    }
    for (unsigned int o = 0; o < count; ++o)
    {
      const unsigned int p = ops[3 * o];
      unsigned int n1 = p, n2 = p;
      if (o + 1 < count)
      {
        /* plain (coherent) loads: these vectors may have been written by this thread earlier in the list */
        n1 = ops[3 * o + 4];
        n2 = ops[3 * o + 5];
        if (n1 != p)
        {
          const unsigned int * c = vec + n1 * node_stride + w;
#pragma unroll
          for (int j = 0; j < N; ++j) nx[j] = c[(size_t)j * words];
        }
        if (n2 != p)
        {
          const unsigned int * c = vec + n2 * node_stride + w;
#pragma unroll
          for (int j = 0; j < N; ++j) ny[j] = c[(size_t)j * words];
        }
      }
      unsigned int orvand = 0;
#pragma unroll
      for (int j = 0; j < N; ++j) orvand |= x[j] & y[j];
      unsigned int * parent = vec + p * node_stride + w;
#pragma unroll
      for (int j = 0; j < N; ++j)
      {
        const unsigned int v = (x[j] & y[j]) | (~orvand & (x[j] | y[j]));
        if (active) parent[(size_t)j * words] = v;
        x[j] = (n1 == p) ? v : nx[j];
        y[j] = (n2 == p) ? v : ny[j];
      }
      const unsigned int pc = __reduce_add_sync(0xffffffffu, active ? (unsigned int)__popc(~orvand) : 0u);
      if ((threadIdx.x & 31) == 0 && pc) atomicAdd(scores + o, pc);
    }
  }
  else
  for (unsigned int o = 0; o < count; ++o)
  {
    unsigned int * parent = vec + ops[3 * o] * node_stride + w;
    const unsigned int * c1 = vec + ops[3 * o + 1] * node_stride + w;
    const unsigned int * c2 = vec + ops[3 * o + 2] * node_stride + w;
    unsigned int orvand = 0;
    for (unsigned int j = 0; j < states; ++j) orvand |= c1[(size_t)j * words] & c2[(size_t)j * words];
    if (active)
      for (unsigned int j = 0; j < states; ++j)
      {
        const unsigned int a = c1[(size_t)j * words], b = c2[(size_t)j * words];
        parent[(size_t)j * words] = (a & b) | (~orvand & (a | b));
      }
    const unsigned int pc = __reduce_add_sync(0xffffffffu, active ? (unsigned int)__popc(~orvand) : 0u);
    if ((threadIdx.x & 31) == 0 && pc) atomicAdd(scores + o, pc);
  }
}

/* h_ops: count x {parent, child1, child2} vector indices; h_scores[count]: mutations added by every op
 * (the caller chains node costs: fast_parsimony.c:527-529) */
extern "C" int plf_pars_update(plf_pars_t * ps, unsigned int * d_vec, unsigned int states, unsigned int words,
                               const unsigned int * h_ops, unsigned int count, unsigned int * h_scores)
{
  plf_ctx * ctx = ps->ctx;
  if (!count) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!pars_reserve(ps, (size_t)4 * count)) return 0;
  if (!words)
  {
    memset(h_scores, 0, count * sizeof(unsigned int));
    return 1;
  }
  unsigned int * d_ops = ps->d_small, * d_scores = ps->d_small + (size_t)3 * count;
  memcpy(ps->h_pin, h_ops, (size_t)3 * count * sizeof(unsigned int));
  PLF_CHECK(ctx, cudaMemcpyAsync(d_ops, ps->h_pin, (size_t)3 * count * sizeof(unsigned int), cudaMemcpyHostToDevice,
                                 ctx->stream));
  PLF_CHECK(ctx, cudaMemsetAsync(d_scores, 0, count * sizeof(unsigned int), ctx->stream));
  const unsigned int grid = (words + PARS_THREADS - 1) / PARS_THREADS;
  const size_t stride = (size_t)states * words;
  if (states == 4)
    k_pars_update<4><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, d_ops, count, d_scores);
  else if (states == 20)
    k_pars_update<20><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, d_ops, count, d_scores);
  else
    k_pars_update<0><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, d_ops, count, d_scores);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaMemcpyAsync(ps->h_pin + (size_t)3 * count, d_scores, count * sizeof(unsigned int),
                                 cudaMemcpyDeviceToHost, ctx->stream));
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(h_scores, ps->h_pin + (size_t)3 * count, count * sizeof(unsigned int));
  return 1;
}

/* One level of a level-scheduled list: blockIdx.y = operation, independent of all others of the launch.  Wide
 * trees keep the chip busy this way (words x operations of a level threads instead of words threads running a
 * chain as deep as the list); the host picks between the two forms (plf_pars_update_levels). */
template <int ST>
__global__ void __launch_bounds__(PARS_THREADS)
k_pars_level(unsigned int * vec, size_t node_stride, unsigned int states, unsigned int words,
             const unsigned int * __restrict__ ops, unsigned int * scores)
{
  const unsigned int o = blockIdx.y;
  unsigned int * parent = vec + ops[3 * o] * node_stride;
  const unsigned int * c1 = vec + ops[3 * o + 1] * node_stride;
  const unsigned int * c2 = vec + ops[3 * o + 2] * node_stride;
  unsigned int pc = 0;
  for (unsigned int w = blockIdx.x * PARS_THREADS + threadIdx.x; w < words; w += gridDim.x * PARS_THREADS)
  {
    unsigned int orvand = 0;
    if constexpr (ST > 0)
    {
      unsigned int x[ST > 0 ? ST : 1], y[ST > 0 ? ST : 1];
#pragma unroll
      for (int j = 0; j < ST; ++j)
      {
        x[j] = c1[(size_t)j * words + w];
        y[j] = c2[(size_t)j * words + w];
      }
#pragma unroll
      for (int j = 0; j < ST; ++j) orvand |= x[j] & y[j];
#pragma unroll
      for (int j = 0; j < ST; ++j) parent[(size_t)j * words + w] = (x[j] & y[j]) | (~orvand & (x[j] | y[j]));
    }
    else
    {
      for (unsigned int j = 0; j < states; ++j) orvand |= c1[(size_t)j * words + w] & c2[(size_t)j * words + w];
      for (unsigned int j = 0; j < states; ++j)
      {
        const unsigned int a = c1[(size_t)j * words + w], b = c2[(size_t)j * words + w];
        parent[(size_t)j * words + w] = (a & b) | (~orvand & (a | b));
      }
    }
    pc += __popc(~orvand);
  }
  pc = __reduce_add_sync(0xffffffffu, pc);
  if ((threadIdx.x & 31) == 0 && pc) atomicAdd(scores + o, pc);
}

/* h_ops: the list sorted by level (level l = entries [h_level_start[l], h_level_start[l+1])), no entry of a
 * level reads or writes what another entry of the same level writes; h_scores in the same order.  One launch
 * per level, one synchronisation at the end. */
extern "C" int plf_pars_update_levels(plf_pars_t * ps, unsigned int * d_vec, unsigned int states, unsigned int words,
                                      const unsigned int * h_ops, unsigned int count,
                                      const unsigned int * h_level_start, unsigned int nlevels,
                                      unsigned int * h_scores)
{
  plf_ctx * ctx = ps->ctx;
  if (!count) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!pars_reserve(ps, (size_t)4 * count)) return 0;
  if (!words)
  {
    memset(h_scores, 0, count * sizeof(unsigned int));
    return 1;
  }
  unsigned int * d_ops = ps->d_small, * d_scores = ps->d_small + (size_t)3 * count;
  memcpy(ps->h_pin, h_ops, (size_t)3 * count * sizeof(unsigned int));
  PLF_CHECK(ctx, cudaMemcpyAsync(d_ops, ps->h_pin, (size_t)3 * count * sizeof(unsigned int), cudaMemcpyHostToDevice,
                                 ctx->stream));
  PLF_CHECK(ctx, cudaMemsetAsync(d_scores, 0, count * sizeof(unsigned int), ctx->stream));
  const size_t stride = (size_t)states * words;
  const unsigned int full = (words + PARS_THREADS - 1) / PARS_THREADS;
  /* Runs of narrow levels (one or two operations: the top of a tree) are not worth a launch each: a run goes
   * to the chain kernel as one launch, in level order, which is a valid sequential order of its operations. */
  unsigned int seg_start = 0, seg_len = 0;
  auto flush_chain = [&]() {
    if (!seg_len) return;
    const unsigned int * o = d_ops + (size_t)3 * seg_start;
    unsigned int * sc = d_scores + seg_start;
    if (states == 4)
      k_pars_update<4><<<full, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, o, seg_len, sc);
    else if (states == 20)
      k_pars_update<20><<<full, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, o, seg_len, sc);
    else
      k_pars_update<0><<<full, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, o, seg_len, sc);
    plf_count_launch();
    seg_len = 0;
  };
  for (unsigned int l = 0; l < nlevels; ++l)
  {
    const unsigned int a = h_level_start[l], n = h_level_start[l + 1] - a;
    if (!n) continue;
    if (n <= 2)
    {
      if (!seg_len) seg_start = a;
      seg_len += n;
      continue;
    }
    flush_chain();
    /* enough CTAs along x to fill the chip when the level is narrow, one sweep of the vector otherwise */
    unsigned int gx = full;
    const unsigned int want = (unsigned int)(ctx->sm_count * 8 + n - 1) / n;
    if (gx > want) gx = want ? want : 1;
    for (unsigned int done = 0; done < n; done += 65535u) /* gridDim.y limit */
    {
      const unsigned int m = n - done < 65535u ? n - done : 65535u;
      const dim3 grid(gx, m);
      const unsigned int * o = d_ops + (size_t)3 * (a + done);
      unsigned int * sc = d_scores + a + done;
      if (states == 4)
        k_pars_level<4><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, o, sc);
      else if (states == 20)
        k_pars_level<20><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, o, sc);
      else
        k_pars_level<0><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, o, sc);
      plf_count_launch();
    }
  }
  flush_chain();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaMemcpyAsync(ps->h_pin + (size_t)3 * count, d_scores, count * sizeof(unsigned int),
                                 cudaMemcpyDeviceToHost, ctx->stream));
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(h_scores, ps->h_pin + (size_t)3 * count, count * sizeof(unsigned int));
  return 1;
}

/* ---- edge scores and the insertion scan ----------------------------------------------------------------- */

/* MODE 0: score[e] = popcount of the sites where the vectors of pair e share no state.
 * MODE 1: the pair is first merged (Fitch parent, in registers) and the merge is scored against `third`:
 *         score[e] = mutations of the merge + mutations against the third vector. */
template <int ST, int MODE>
__global__ void __launch_bounds__(PARS_THREADS)
k_pars_pairs(const unsigned int * __restrict__ vec, size_t node_stride, unsigned int states, unsigned int words,
             const unsigned int * __restrict__ pairs, unsigned int third, unsigned int * scores)
{
  const unsigned int e = blockIdx.y;
  const unsigned int * n1 = vec + pairs[2 * e] * node_stride;
  const unsigned int * n2 = vec + pairs[2 * e + 1] * node_stride;
  const unsigned int * n3 = vec + third * node_stride;
  unsigned int pc = 0;
  for (unsigned int w = blockIdx.x * PARS_THREADS + threadIdx.x; w < words; w += gridDim.x * PARS_THREADS)
  {
    unsigned int orvand = 0;
    if (ST > 0)
    {
      unsigned int x[ST > 0 ? ST : 1], y[ST > 0 ? ST : 1];
#pragma unroll
      for (int j = 0; j < ST; ++j)
      {
        x[j] = n1[(size_t)j * words + w];
        y[j] = n2[(size_t)j * words + w];
      }
#pragma unroll
      for (int j = 0; j < ST; ++j) orvand |= x[j] & y[j];
      pc += __popc(~orvand);
      if (MODE == 1)
      {
        unsigned int or2 = 0;
#pragma unroll
        for (int j = 0; j < ST; ++j) or2 |= ((x[j] & y[j]) | (~orvand & (x[j] | y[j]))) & n3[(size_t)j * words + w];
        pc += __popc(~or2);
      }
    }
    else
    {
      for (unsigned int j = 0; j < states; ++j) orvand |= n1[(size_t)j * words + w] & n2[(size_t)j * words + w];
      pc += __popc(~orvand);
      if (MODE == 1)
      {
        unsigned int or2 = 0;
        for (unsigned int j = 0; j < states; ++j)
        {
          const unsigned int x = n1[(size_t)j * words + w], y = n2[(size_t)j * words + w];
          or2 |= ((x & y) | (~orvand & (x | y))) & n3[(size_t)j * words + w];
        }
        pc += __popc(~or2);
      }
    }
  }
  pc = __reduce_add_sync(0xffffffffu, pc);
  if ((threadIdx.x & 31) == 0 && pc) atomicAdd(scores + e, pc);
}

template <int MODE>
static int pars_pairs(plf_pars * ps, const unsigned int * d_vec, unsigned int states, unsigned int words,
                      const unsigned int * h_pairs, unsigned int n, unsigned int third, unsigned int * h_scores)
{
  plf_ctx * ctx = ps->ctx;
  if (!n) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!pars_reserve(ps, (size_t)3 * n)) return 0;
  if (!words)
  {
    memset(h_scores, 0, n * sizeof(unsigned int));
    return 1;
  }
  unsigned int * d_pairs = ps->d_small, * d_scores = ps->d_small + (size_t)2 * n;
  memcpy(ps->h_pin, h_pairs, (size_t)2 * n * sizeof(unsigned int));
  PLF_CHECK(ctx, cudaMemcpyAsync(d_pairs, ps->h_pin, (size_t)2 * n * sizeof(unsigned int), cudaMemcpyHostToDevice,
                                 ctx->stream));
  PLF_CHECK(ctx, cudaMemsetAsync(d_scores, 0, n * sizeof(unsigned int), ctx->stream));
  /* enough CTAs along x to fill the chip when the batch is small, one sweep of the vector otherwise */
  unsigned int gx = (words + PARS_THREADS - 1) / PARS_THREADS;
  const unsigned int want = (unsigned int)(ctx->sm_count * 8 + n - 1) / n;
  if (gx > want) gx = want ? want : 1;
  const dim3 grid(gx, n);
  const size_t stride = (size_t)states * words;
  if (states == 4)
    k_pars_pairs<4, MODE><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, d_pairs, third, d_scores);
  else if (states == 20)
    k_pars_pairs<20, MODE><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, d_pairs, third, d_scores);
  else
    k_pars_pairs<0, MODE><<<grid, PARS_THREADS, 0, ctx->stream>>>(d_vec, stride, states, words, d_pairs, third, d_scores);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaMemcpyAsync(ps->h_pin + (size_t)2 * n, d_scores, n * sizeof(unsigned int), cudaMemcpyDeviceToHost,
                                 ctx->stream));
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(h_scores, ps->h_pin + (size_t)2 * n, n * sizeof(unsigned int));
  return 1;
}

extern "C" int plf_pars_edge_scores(plf_pars_t * ps, const unsigned int * d_vec, unsigned int states, unsigned int words,
                                    const unsigned int * h_pairs, unsigned int n, unsigned int * h_scores)
{
  return pars_pairs<0>(ps, d_vec, states, words, h_pairs, n, 0, h_scores);
}

extern "C" int plf_pars_insert_scan(plf_pars_t * ps, const unsigned int * d_vec, unsigned int states, unsigned int words,
                                    const unsigned int * h_pairs, unsigned int n, unsigned int third,
                                    unsigned int * h_scores)
{
  return pars_pairs<1>(ps, d_vec, states, words, h_pairs, n, third, h_scores);
}

/* ---- weighted (Sankoff) parsimony, src/parsimony.c ----------------------------------------------------------
 * Score buffers are [site][state] doubles.  Sites are independent, so as with the bit vectors one thread per
 * site runs the WHOLE operation list: a traversal (and a reconstruction pass) is one launch.  The arithmetic is
 * the reference's, operation for operation (sequential fmin over the child states, child 1 then child 2), so
 * buffers are bit-identical. */

template <int ST>
__global__ void __launch_bounds__(128)
k_wpars_build(double * const * __restrict__ sbuf, const unsigned int * __restrict__ ops, unsigned int count,
              unsigned int states_rt, unsigned int sites, const double * __restrict__ matrix)
{
  extern __shared__ double M[]; /* [states][states] */
  const unsigned int states = ST ? ST : states_rt;
  for (unsigned int i = threadIdx.x; i < states * states; i += blockDim.x) M[i] = matrix[i];
  __syncthreads();
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= sites) return;
  for (unsigned int o = 0; o < count; ++o)
  {
    double * parent = sbuf[ops[3 * o]] + (size_t)site * states;
    const double * c1 = sbuf[ops[3 * o + 1]] + (size_t)site * states;
    const double * c2 = sbuf[ops[3 * o + 2]] + (size_t)site * states;
    if constexpr (ST > 0)
    {
      double a[ST > 0 ? ST : 1], b[ST > 0 ? ST : 1];
#pragma unroll
      for (int k = 0; k < ST; ++k)
      {
        a[k] = c1[k];
        b[k] = c2[k];
      }
#pragma unroll
      for (int n = 0; n < ST; ++n)
      {
        double m1 = a[0] + M[n], m2 = b[0] + M[n];
#pragma unroll
        for (int k = 1; k < ST; ++k)
        {
          m1 = fmin(a[k] + M[k * ST + n], m1);
          m2 = fmin(b[k] + M[k * ST + n], m2);
        }
        parent[n] = m1 + m2;
      }
    }
    else
    {
      /* the parent may alias a child only in lists the reference would also get wrong: no staging */
      for (unsigned int n = 0; n < states; ++n)
      {
        double m1 = c1[0] + M[n], m2 = c2[0] + M[n];
        for (unsigned int k = 1; k < states; ++k)
        {
          m1 = fmin(c1[k] + M[k * states + n], m1);
          m2 = fmin(c2[k] + M[k * states + n], m2);
        }
        parent[n] = m1 + m2;
      }
    }
  }
}

__global__ void k_wpars_tip(double * __restrict__ out, const unsigned char * __restrict__ seq,
                            const unsigned long long * __restrict__ map, unsigned int sites, unsigned int states,
                            double inf)
{
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= sites) return;
  const unsigned long long c = map[seq[site]];
  for (unsigned int j = 0; j < states; ++j) out[(size_t)site * states + j] = ((c >> j) & 1ull) ? 0.0 : inf;
}

__global__ void k_wpars_site_min(const double * __restrict__ buf, unsigned int states, unsigned int sites,
                                 double * __restrict__ out)
{
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= sites) return;
  const double * s = buf + (size_t)site * states;
  double m = s[0];
  for (unsigned int j = 1; j < states; ++j) m = fmin(s[j], m);
  out[site] = m;
}

/* recops: count x {node score, node ancestral, parent score, parent ancestral}; op 0 is the subtree root
 * (src/parsimony.c:301-382).  revmap[state] = character of the one-state code, ctz[ch] = its state. */
__global__ void __launch_bounds__(128)
k_wpars_reconstruct(double * const * __restrict__ sbuf, unsigned int * const * __restrict__ anc,
                    const unsigned int * __restrict__ ops, unsigned int count, unsigned int states, unsigned int sites,
                    const unsigned int * __restrict__ revmap_g, const unsigned long long * __restrict__ map_g)
{
  __shared__ unsigned int revmap[256];
  __shared__ unsigned char ctz[256];
  for (unsigned int i = threadIdx.x; i < 256; i += blockDim.x)
  {
    revmap[i] = revmap_g[i];
    ctz[i] = map_g[i] ? (unsigned char)(__ffsll((long long)map_g[i]) - 1) : 64;
  }
  __syncthreads();
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= sites) return;
  for (unsigned int o = 0; o < count; ++o)
  {
    const double * score = sbuf[ops[4 * o]] + (size_t)site * states;
    unsigned int minindex = 0;
    for (unsigned int j = 1; j < states; ++j)
      if (score[j] < score[minindex]) minindex = j;
    unsigned int * mine = anc[ops[4 * o + 1]] + site;
    if (o == 0)
    {
      *mine = revmap[minindex];
      continue;
    }
    const unsigned int pch = anc[ops[4 * o + 3]][site];
    const double parent_val = (sbuf[ops[4 * o + 2]] + (size_t)site * states)[ctz[pch & 255u] & 63];
    *mine = (score[minindex] + 1 > parent_val) ? pch : revmap[minindex];
  }
}

extern "C" int plf_wpars_tip(plf_ctx_t * ctx, double * d_out, const char * h_seq, const unsigned long long * h_map,
                             unsigned int sites, unsigned int states, double inf)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!sites) return 1;
  unsigned char * d = (unsigned char *)plf_ws_reserve(ctx, &ctx->ws_small, (size_t)sites + 256 * sizeof(unsigned long long) + 64);
  if (!d) return 0;
  unsigned long long * d_map = (unsigned long long *)d;
  unsigned char * d_seq = d + 256 * sizeof(unsigned long long);
  PLF_CHECK(ctx, cudaMemcpyAsync(d_map, h_map, 256 * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  PLF_CHECK(ctx, cudaMemcpyAsync(d_seq, h_seq, sites, cudaMemcpyHostToDevice, ctx->stream));
  k_wpars_tip<<<(sites + 255) / 256, 256, 0, ctx->stream>>>(d_out, d_seq, d_map, sites, states, inf);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 1;
}

extern "C" int plf_wpars_build(plf_ctx_t * ctx, double * const * d_sbuf_table, const unsigned int * h_ops,
                               unsigned int count, unsigned int states, unsigned int sites, const double * d_matrix)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!count || !sites) return 1;
  unsigned int * d_ops = (unsigned int *)plf_ws_reserve(ctx, &ctx->ws_ops, (size_t)3 * count * sizeof(unsigned int));
  if (!d_ops) return 0;
  PLF_CHECK(ctx, cudaMemcpyAsync(d_ops, h_ops, (size_t)3 * count * sizeof(unsigned int), cudaMemcpyHostToDevice,
                                 ctx->stream));
  const size_t smem = (size_t)states * states * sizeof(double);
  if (smem > 48 * 1024)
  {
    plf_set_error(ctx, "weighted parsimony: %u states need a %zu-byte score matrix in shared memory", states, smem);
    return 0;
  }
  const unsigned int grid = (sites + 127) / 128;
  if (states == 4)
    k_wpars_build<4><<<grid, 128, smem, ctx->stream>>>(d_sbuf_table, d_ops, count, states, sites, d_matrix);
  else if (states == 20)
    k_wpars_build<20><<<grid, 128, smem, ctx->stream>>>(d_sbuf_table, d_ops, count, states, sites, d_matrix);
  else
    k_wpars_build<0><<<grid, 128, smem, ctx->stream>>>(d_sbuf_table, d_ops, count, states, sites, d_matrix);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); /* the op list came from pageable memory; results are host-visible */
  return 1;
}

extern "C" int plf_wpars_site_min(plf_ctx_t * ctx, const double * d_buf, unsigned int states, unsigned int sites,
                                  double * h_out)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!sites) return 1;
  double * d_min = (double *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)sites * sizeof(double));
  if (!d_min) return 0;
  k_wpars_site_min<<<(sites + 255) / 256, 256, 0, ctx->stream>>>(d_buf, states, sites, d_min);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaMemcpyAsync(h_out, d_min, (size_t)sites * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 1;
}

extern "C" int plf_wpars_reconstruct(plf_ctx_t * ctx, double * const * d_sbuf_table, unsigned int * const * d_anc_table,
                                     const unsigned int * h_recops, unsigned int count, unsigned int states,
                                     unsigned int sites, const unsigned int * h_revmap,
                                     const unsigned long long * h_map)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (!count || !sites) return 1;
  unsigned int * d_ops = (unsigned int *)plf_ws_reserve(ctx, &ctx->ws_ops, (size_t)4 * count * sizeof(unsigned int));
  unsigned char * d = (unsigned char *)plf_ws_reserve(ctx, &ctx->ws_small, 256 * (sizeof(unsigned long long) + sizeof(unsigned int)));
  if (!d_ops || !d) return 0;
  unsigned long long * d_map = (unsigned long long *)d;
  unsigned int * d_rev = (unsigned int *)(d + 256 * sizeof(unsigned long long));
  PLF_CHECK(ctx, cudaMemcpyAsync(d_ops, h_recops, (size_t)4 * count * sizeof(unsigned int), cudaMemcpyHostToDevice,
                                 ctx->stream));
  PLF_CHECK(ctx, cudaMemcpyAsync(d_map, h_map, 256 * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  PLF_CHECK(ctx, cudaMemcpyAsync(d_rev, h_revmap, 256 * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
  k_wpars_reconstruct<<<(sites + 127) / 128, 128, 0, ctx->stream>>>(d_sbuf_table, d_anc_table, d_ops, count, states,
                                                                     sites, d_rev, d_map);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 1;
}
