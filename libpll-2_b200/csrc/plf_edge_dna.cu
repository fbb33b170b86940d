/*
 * plf_edge_dna.cu -- 4-state (DNA) edge / root log-likelihood, sumtable and
 * derivative kernels for contiguous (non-repeats) CLVs and per-site scalers.
 *
 * Replaces, for 4 states, pll_core_edge_loglikelihood_ii / _ti_4x4 and
 * pll_core_root_loglikelihood (reference src/core_likelihood.c:1192,352,25;
 * AVX: src/core_likelihood_avx.c:1513,286), pll_core_update_sumtable_ii / _ti
 * (src/core_derivatives.c:321,473; AVX: src/core_derivatives_avx.c:25,212) and
 * pll_core_likelihood_derivatives (src/core_derivatives.c:696).  Everything
 * else (site-repeat gathers, per-rate scalers, rate counts that are not a
 * power of two <= 8) stays on the generic kernels of plf_likelihood.cu.
 *
 * All three are streaming reductions bounded by HBM (268 / 137 / 132 B per
 * site against ~150 FP64 operations), so they reuse the design of the
 * dominant CLV kernel (plf_partials_dna.cu): CLV / sumtable tiles, scalers,
 * tip codes, pattern weights and invariant flags are brought into a
 * shared-memory ring by 1-D bulk async copies completing on an mbarrier; one
 * thread per (site, rate) reads its 32-byte block from shared memory, lanes
 * of a site combine with shuffles, and the block that finishes last adds the
 * per-block partial sums in a fixed order (one launch, reproducible bits).
 *
 * The sumtable is not a kernel of its own: sum[j] = (sum_k clvp_k pi_k
 * Vinv_kj) (sum_k V_jk clvc_k) is a CLV update with "P-matrices"
 * (pi Vinv)^T and V and no scaling, so it runs on the streaming CLV kernels
 * with a synthetic operation.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"
#include "plf_stream.cuh"

#include <stdlib.h>

#define EDGE_THREADS 128
enum { EDGE_II = 0, EDGE_TI = 1, EDGE_ROOT = 2, EDGE_DERIV = 3 };

/* one ring stage, only the arrays MODE needs */
template <int LOG2R, int ITEMS, int MODE>
struct EdgeLayout
{
  static constexpr int R = 1 << LOG2R;
  static constexpr int TILE = (EDGE_THREADS * ITEMS) >> LOG2R; /* sites per tile */
  static constexpr int CLV_BYTES = TILE * R * 32;
  static constexpr int OFF_P = 0;
  static constexpr int OFF_C = CLV_BYTES;
  static constexpr int OFF_PSC = (MODE == EDGE_II ? 2 : 1) * CLV_BYTES;
  static constexpr int OFF_CSC = OFF_PSC + (MODE == EDGE_DERIV ? 0 : TILE * 4);
  static constexpr int OFF_W = OFF_CSC + (MODE == EDGE_II ? TILE * 4 : 0);
  static constexpr int OFF_INV = OFF_W + TILE * 4;
  static constexpr int OFF_CODE = OFF_INV + TILE * 4;
  static constexpr int STAGE_BYTES = (OFF_CODE + (MODE == EDGE_TI ? TILE : 0) + 127) & ~127;
};

/* what one tile needs, as the device sees it */
struct edge_src_t
{
  const double * p;            /* parent CLV or sumtable */
  const double * c;            /* child CLV (EDGE_II) */
  const unsigned int * psc;
  const unsigned int * csc;
  const unsigned int * w;
  const int * inv;
  const unsigned char * code;  /* EDGE_TI */
  unsigned int sites;
};

template <int LOG2R, int MODE, int ITEMS>
__device__ __forceinline__ void edge_issue(const edge_src_t & a, unsigned int t, unsigned char * slot,
                                           unsigned long long * bar)
{
  typedef EdgeLayout<LOG2R, ITEMS, MODE> Ly;
  const unsigned int first = t * Ly::TILE;
  const unsigned int n = min((unsigned int)Ly::TILE, a.sites - first);
  const unsigned int clv_bytes = n * Ly::R * 32;
  const unsigned int u32_bytes = (n * 4 + 15) & ~15u; /* allocations carry 16 bytes of slack (plf_alloc) */
  const unsigned int code_bytes = (n + 15) & ~15u;
  unsigned int total = clv_bytes + u32_bytes;
  if (MODE == EDGE_II) total += clv_bytes;
  if (MODE == EDGE_TI) total += code_bytes;
  if (a.psc) total += u32_bytes;
  if (MODE == EDGE_II && a.csc) total += u32_bytes;
  if (a.inv) total += u32_bytes;
  mbar_expect_tx(bar, total);
  bulk_g2s(slot + Ly::OFF_P, a.p + (size_t)first * Ly::R * 4, clv_bytes, bar);
  if (MODE == EDGE_II) bulk_g2s(slot + Ly::OFF_C, a.c + (size_t)first * Ly::R * 4, clv_bytes, bar);
  if (MODE == EDGE_TI) bulk_g2s(slot + Ly::OFF_CODE, a.code + first, code_bytes, bar);
  if (a.psc) bulk_g2s(slot + Ly::OFF_PSC, a.psc + first, u32_bytes, bar);
  if (MODE == EDGE_II && a.csc) bulk_g2s(slot + Ly::OFF_CSC, a.csc + first, u32_bytes, bar);
  bulk_g2s(slot + Ly::OFF_W, a.w + first, u32_bytes, bar);
  if (a.inv) bulk_g2s(slot + Ly::OFF_INV, a.inv + first, u32_bytes, bar);
}

/* ------------------------------------------------------------------------ *
 *  log-likelihood: MODE = EDGE_II (two CLVs), EDGE_TI (child is a pattern    *
 *  tip) or EDGE_ROOT (no P-matrix).  Model block layout: plf_backend.h.      *
 * ------------------------------------------------------------------------ */
template <int LOG2R, int MODE, int NSTAGE, int ITEMS>
__global__ void __launch_bounds__(EDGE_THREADS)
k_lk_dna(plf_lk_t a, double * __restrict__ partial, unsigned int * ticket, double * out, double * hout)
{
  typedef EdgeLayout<LOG2R, ITEMS, MODE> Ly;
  constexpr int R = Ly::R;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) unsigned long long full[NSTAGE];
  __shared__ __align__(16) double tl[MODE == EDGE_TI ? 64 * R : 2];
  __shared__ double red[32];

  edge_src_t src;
  src.p = a.clvp;
  src.c = a.clvc;
  src.psc = a.pscaler;
  src.csc = a.cscaler;
  src.w = a.pattern_weights;
  src.inv = a.invariant;
  src.code = a.tipchars;
  src.sites = a.sites;
  const unsigned int ntiles = (a.sites + Ly::TILE - 1) / Ly::TILE;
  const int rate = threadIdx.x & (R - 1);

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
      edge_issue<LOG2R, MODE, ITEMS>(src, t, ring + (size_t)s * Ly::STAGE_BYTES, &full[s]);
  }

  /* per-rate constants of this thread */
  const double * freqs = a.model + 3 * R + rate * 4;
  const double f0 = freqs[0], f1 = freqs[1], f2 = freqs[2], f3 = freqs[3];
  const double wr = a.model[R + rate], pinv = a.model[2 * R + rate];
  double Pm[MODE == EDGE_II ? 16 : 1];
  if (MODE == EDGE_II)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i) Pm[i] = a.pmatrix[rate * 16 + i];
  }
  if (MODE == EDGE_TI)
  {
    build_tip_table(tl, a.pmatrix, R);
    __syncthreads();
  }
  const bool has_inv = (a.invariant != nullptr);

  double acc = 0;
  unsigned int it = 0;
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it)
  {
    const int s = it % NSTAGE;
    const unsigned int parity = (it / NSTAGE) & 1u;
    unsigned char * slot = ring + (size_t)s * Ly::STAGE_BYTES;
    while (!mbar_try_wait(&full[s], parity)) {}
    const unsigned int first = t * Ly::TILE;
    double ta[ITEMS], ti[ITEMS];
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
    {
      const unsigned int item = threadIdx.x + u * EDGE_THREADS; /* (site in tile, rate) */
      const unsigned int ls = item >> LOG2R;
      const unsigned int n = first + ls;
      const bool active = n < a.sites;
      const dbl4 p = lds_dbl4(reinterpret_cast<const double *>(slot + Ly::OFF_P) + (size_t)item * 4);
      const int inv = (has_inv && active) ? reinterpret_cast<const int *>(slot + Ly::OFF_INV)[ls] : -1;
      double term;
      if (MODE == EDGE_ROOT)
      {
        term = fma(p.w, f3, fma(p.z, f2, fma(p.y, f1, p.x * f0)));
      }
      else
      {
        dbl4 tb;
        if (MODE == EDGE_II)
        {
          const dbl4 c = lds_dbl4(reinterpret_cast<const double *>(slot + Ly::OFF_C) + (size_t)item * 4);
          tb.x = fma(Pm[3], c.w, fma(Pm[2], c.z, fma(Pm[1], c.y, Pm[0] * c.x)));
          tb.y = fma(Pm[7], c.w, fma(Pm[6], c.z, fma(Pm[5], c.y, Pm[4] * c.x)));
          tb.z = fma(Pm[11], c.w, fma(Pm[10], c.z, fma(Pm[9], c.y, Pm[8] * c.x)));
          tb.w = fma(Pm[15], c.w, fma(Pm[14], c.z, fma(Pm[13], c.y, Pm[12] * c.x)));
        }
        else
        {
          const unsigned int code = active ? slot[Ly::OFF_CODE + ls] : 0u;
          tb = lds_dbl4(tl + (code * R + rate) * 4);
        }
        term = fma(p.w * f3, tb.w, fma(p.z * f2, tb.z, fma(p.y * f1, tb.y, (p.x * f0) * tb.x)));
      }
      double terma, terminv = 0;
      if (pinv > 0)
      {
        const double inv_lk = (inv == -1) ? 0.0 : freqs[inv];
        if (MODE == EDGE_ROOT)
          terma = wr * (term * (1.0 - pinv) + inv_lk * pinv); /* core_likelihood.c:179-180 */
        else
        {
          terma = wr * term * (1.0 - pinv);                   /* core_likelihood.c:1445-1452 */
          if (inv != -1) terminv = wr * inv_lk * pinv;
        }
      }
      else
        terma = term * wr;
      ta[u] = group_sum(terma, R);
      ti[u] = (has_inv && MODE != EDGE_ROOT) ? group_sum(terminv, R) : 0.0;
    }
    /* finish the sites (log, scalers, weight): lane l of a site's lane group
     * takes item q*R + l, so the expensive log runs on compacted lanes */
#pragma unroll
    for (int q = 0; q < (ITEMS + R - 1) / R; ++q)
    {
      double terma = 0, terminv = 0;
      int u_mine = -1;
#pragma unroll
      for (int l = 0; l < R; ++l)
      {
        const int u = q * R + l;
        if (u < ITEMS && rate == l)
        {
          terma = ta[u];
          terminv = ti[u];
          u_mine = u;
        }
      }
      const unsigned int ls = (threadIdx.x + (unsigned int)(u_mine < 0 ? 0 : u_mine) * EDGE_THREADS) >> LOG2R;
      const unsigned int n = first + ls;
      if (u_mine >= 0 && n < a.sites)
      {
        unsigned int sc = 0;
        if (a.pscaler) sc += reinterpret_cast<const unsigned int *>(slot + Ly::OFF_PSC)[ls];
        if (MODE == EDGE_II && a.cscaler) sc += reinterpret_cast<const unsigned int *>(slot + Ly::OFF_CSC)[ls];
        double site_lk;
        if (MODE == EDGE_ROOT)
        {
          site_lk = log(terma);
          if (sc) site_lk += sc * PLF_LOG_SCALE_THRESHOLD;
        }
        else if (sc)
        {
          if (terminv > 0.0)
          {
            const unsigned int capped = sc < PLF_MAXDIFF ? sc : PLF_MAXDIFF;
            site_lk = log(ldexp(terma, -256 * (int)capped) + terminv); /* core_likelihood.c:1366-1376 */
          }
          else
            site_lk = log(terma) + sc * PLF_LOG_SCALE_THRESHOLD;
        }
        else
          site_lk = log(terma + terminv);
        site_lk *= (double)reinterpret_cast<const unsigned int *>(slot + Ly::OFF_W)[ls];
        if (a.persite) a.persite[n] = site_lk;
        acc += site_lk;
      }
    }
    __syncthreads(); /* every warp is done with this slot */
    const unsigned int tn = t + (unsigned int)NSTAGE * gridDim.x;
    if (threadIdx.x == 0 && tn < ntiles) edge_issue<LOG2R, MODE, ITEMS>(src, tn, slot, &full[s]);
  }
  const double v[1] = {acc};
  grid_reduce_finish<1>(v, partial, ticket, out, hout, red);
}

/* ------------------------------------------------------------------------ *
 *  derivatives: per site L, L', L'' from the sumtable and                    *
 *  diag[r][j] = {e, lk e, (lk)^2 e}, e = exp(lambda_j k_r t),                 *
 *  k_r = rate_r / (1 - pinv_r)        (core_derivatives.c:757-772,825-848)   *
 * ------------------------------------------------------------------------ */
template <int LOG2R, int NSTAGE, int ITEMS>
__global__ void __launch_bounds__(EDGE_THREADS)
k_deriv_dna(plf_deriv_t a, double * __restrict__ partial, unsigned int * ticket, double * out, double * hout)
{
  typedef EdgeLayout<LOG2R, ITEMS, EDGE_DERIV> Ly;
  constexpr int R = Ly::R;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) unsigned long long full[NSTAGE];
  __shared__ double red[32];

  edge_src_t src;
  src.p = a.sumtable;
  src.c = nullptr;
  src.psc = src.csc = nullptr;
  src.w = a.pattern_weights;
  src.inv = a.invariant;
  src.code = nullptr;
  src.sites = a.sites;
  const unsigned int ntiles = (a.sites + Ly::TILE - 1) / Ly::TILE;
  const int rate = threadIdx.x & (R - 1);

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
      edge_issue<LOG2R, EDGE_DERIV, ITEMS>(src, t, ring + (size_t)s * Ly::STAGE_BYTES, &full[s]);
  }

  const double * freqs = a.model + 3 * R + rate * 4;
  const double * evals = a.model + 3 * R + R * 4 + rate * 4;
  const double wr = a.model[R + rate], pinv = a.model[2 * R + rate];
  const double ki = a.model[rate] / (1.0 - pinv);
  double d0[4], d1[4], d2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
  {
    const double lam = evals[j];
    const double e = exp(lam * ki * a.branch_length);
    d0[j] = e;
    d1[j] = lam * ki * e;
    d2[j] = lam * ki * lam * ki * e;
  }
  const bool has_inv = (a.invariant != nullptr);

  double acc1 = 0, acc2 = 0;
  unsigned int it = 0;
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it)
  {
    const int s = it % NSTAGE;
    const unsigned int parity = (it / NSTAGE) & 1u;
    unsigned char * slot = ring + (size_t)s * Ly::STAGE_BYTES;
    while (!mbar_try_wait(&full[s], parity)) {}
    const unsigned int first = t * Ly::TILE;
    double l0[ITEMS], l1[ITEMS], l2[ITEMS];
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
    {
      const unsigned int item = threadIdx.x + u * EDGE_THREADS;
      const unsigned int ls = item >> LOG2R;
      const unsigned int n = first + ls;
      const bool active = n < a.sites;
      const dbl4 sv = lds_dbl4(reinterpret_cast<const double *>(slot + Ly::OFF_P) + (size_t)item * 4);
      double c0 = fma(sv.w, d0[3], fma(sv.z, d0[2], fma(sv.y, d0[1], fma(sv.x, d0[0], 0.0))));
      double c1 = fma(sv.w, d1[3], fma(sv.z, d1[2], fma(sv.y, d1[1], fma(sv.x, d1[0], 0.0))));
      double c2 = fma(sv.w, d2[3], fma(sv.z, d2[2], fma(sv.y, d2[1], fma(sv.x, d2[0], 0.0))));
      if (pinv > 0)
      {
        const int inv = (has_inv && active) ? reinterpret_cast<const int *>(slot + Ly::OFF_INV)[ls] : -1;
        const double inv_lk = (inv == -1) ? 0.0 : freqs[inv] * pinv;
        c0 = c0 * (1.0 - pinv) + inv_lk;
        c1 = c1 * (1.0 - pinv);
        c2 = c2 * (1.0 - pinv);
      }
      l0[u] = group_sum(c0 * wr, R);
      l1[u] = group_sum(c1 * wr, R);
      l2[u] = group_sum(c2 * wr, R);
    }
    /* lane l of a site's lane group finishes item q*R + l: divisions on compacted lanes */
#pragma unroll
    for (int q = 0; q < (ITEMS + R - 1) / R; ++q)
    {
      double lk0 = 1, lk1 = 0, lk2 = 0;
      int u_mine = -1;
#pragma unroll
      for (int l = 0; l < R; ++l)
      {
        const int u = q * R + l;
        if (u < ITEMS && rate == l)
        {
          lk0 = l0[u];
          lk1 = l1[u];
          lk2 = l2[u];
          u_mine = u;
        }
      }
      const unsigned int ls = (threadIdx.x + (unsigned int)(u_mine < 0 ? 0 : u_mine) * EDGE_THREADS) >> LOG2R;
      if (u_mine >= 0 && first + ls < a.sites)
      {
        const double w = (double)reinterpret_cast<const unsigned int *>(slot + Ly::OFF_W)[ls];
        const double rinv = 1.0 / lk0;
        const double r1 = -lk1 * rinv;
        const double r2 = r1 * r1 - lk2 * rinv;
        acc1 = fma(w, r1, acc1);
        acc2 = fma(w, r2, acc2);
      }
    }
    __syncthreads();
    const unsigned int tn = t + (unsigned int)NSTAGE * gridDim.x;
    if (threadIdx.x == 0 && tn < ntiles) edge_issue<LOG2R, EDGE_DERIV, ITEMS>(src, tn, slot, &full[s]);
  }
  const double v[2] = {acc1, acc2};
  grid_reduce_finish<2>(v, partial, ticket, out, hout, red);
}

/* ------------------------------------------------------------------------ *
 *  sumtable as a CLV update: build the synthetic operation on the device      *
 *  ws = [plf_op_t (256 B)] [left R x 16] [right R x 16]                       *
 *  left[r][j][k] = pi_k Vinv[k][j] ("parent" side), right[r][j][k] = V[j][k]   *
 * ------------------------------------------------------------------------ */
__global__ void k_sumtable_op(plf_sumtable_t a, int R, int st, int sp, unsigned char * ws, unsigned int ntiles)
{
  plf_op_t * op = reinterpret_cast<plf_op_t *>(ws);
  double * lm = reinterpret_cast<double *>(ws + 256);
  double * rm = lm + R * st * sp;
  const double * freqs = a.model + 3 * R;
  const double * evecs = freqs + 2 * R * sp;
  const double * ievecs = evecs + R * st * sp;
  for (int e = threadIdx.x; e < R * st * sp; e += blockDim.x)
  {
    const int r = e / (st * sp), j = (e / sp) % st, k = e % sp;
    lm[e] = (k < st) ? freqs[r * sp + k] * ievecs[r * st * sp + k * sp + j] : 0.0;
    rm[e] = (k < st) ? evecs[r * st * sp + j * sp + k] : 0.0;
  }
  if (threadIdx.x == 0)
  {
    plf_op_t o;
    memset(&o, 0, sizeof(o));
    o.parent_clv = a.sumtable;
    o.left_clv = a.clvp;
    o.right_clv = a.clvc;
    o.left_tip = a.tipchars;
    o.right_tip = nullptr;
    o.left_matrix = lm;
    o.right_matrix = rm;
    o.parent_scaler = nullptr; /* no scaling test, no scaler traffic */
    o.left_scaler = o.right_scaler = nullptr;
    o.parent_id_site = nullptr;
    o.left_site_id = a.p_site_id; /* repeats: full-length output, gathered inputs (core_derivatives.c:25) */
    o.right_site_id = a.c_site_id;
    o.nsites = a.sites;
    o.kind = a.tipchars ? PLF_OP_TI : PLF_OP_II;
    *op = o;
    /* tile list of this one op for the tile-walk gather kernel (last 8 bytes of the descriptor's slot) */
    unsigned int * prefix = reinterpret_cast<unsigned int *>(ws + 248);
    prefix[0] = 0;
    prefix[1] = ntiles;
  }
}

/* 4 states: the streaming / gathering DNA CLV kernels; 20 states: the DMMA kernels */
int plf_sumtable_as_clv(plf_ctx * ctx, const plf_shape_t * sh, const plf_sumtable_t * a,
                        const unsigned long long * d_tipmap, unsigned int maxstates)
{
  const int R = (int)sh->rate_cats, st = (int)sh->states, sp = (int)sh->states_padded;
  unsigned char * ws =
      (unsigned char *)plf_ws_reserve(ctx, &ctx->ws_edge, 256 + (size_t)2 * R * st * sp * sizeof(double));
  if (!ws) return 0;
  static_assert(sizeof(plf_op_t) <= 248, "op descriptor and its two-entry tile list must fit the 256-byte slot");
  const int contiguous = !(a->p_site_id || a->c_site_id);
  const unsigned int kind = a->tipchars ? PLF_OP_TI : PLF_OP_II;
  const int pow2 = sh->rate_cats && !(sh->rate_cats & (sh->rate_cats - 1)) && sh->rate_cats <= 32;
  const unsigned int ntiles = (st == 4 && pow2 && !contiguous && kind == PLF_OP_II)
                                  ? plf_dna_balanced_tiles(a->sites, sh->rate_cats) : 0;
  k_sumtable_op<<<1, 256, 0, ctx->stream>>>(*a, R, st, sp, ws, ntiles);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  if (st == 4)
    return plf_launch_dna_group(ctx, reinterpret_cast<const plf_op_t *>(ws), 1, kind, sh->rate_cats, 0, a->sites,
                                contiguous, ntiles ? reinterpret_cast<const unsigned int *>(ws + 248) : nullptr, ntiles);
  return plf_launch_aa_mma_group(ctx, reinterpret_cast<const plf_op_t *>(ws), 1, kind, sh->rate_cats, 0, a->sites,
                                 d_tipmap, maxstates, contiguous);
}

/* ------------------------------------------------------------------------ */

/* ring shape: (ITEMS, NSTAGE) = (4, 3) finishes one site per lane at 4 rates; PLF_EDGE_ITEMS=2 selects (2, 6) */
static int edge_items(plf_ctx * ctx)
{
  if (!ctx->edge_items)
  {
    const char * v = getenv("PLF_EDGE_ITEMS");
    ctx->edge_items = (v && v[0] == '2') ? 2 : 4;
  }
  return ctx->edge_items;
}

template <typename K>
static int edge_grid(plf_ctx * ctx, K kernel, size_t smem, int * occ_cache, unsigned int sites, unsigned int tile,
                     unsigned int * blocks)
{
  if (!*occ_cache)
  {
    PLF_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_cache, kernel, EDGE_THREADS, smem));
    if (*occ_cache < 1) *occ_cache = 1;
  }
  const unsigned long long ntiles = ((unsigned long long)sites + tile - 1) / tile;
  unsigned long long b = (unsigned long long)ctx->sm_count * *occ_cache;
  if (b > ntiles) b = ntiles;
  *blocks = (unsigned int)(b < 1 ? 1 : b);
  return 1;
}

template <int LOG2R, int MODE, int EDGE_ITEMS, int EDGE_NSTAGE>
static int launch_lk_shape(plf_ctx * ctx, const plf_lk_t * a, double * dst, double * hdst)
{
  typedef EdgeLayout<LOG2R, EDGE_ITEMS, MODE> Ly;
  auto k = k_lk_dna<LOG2R, MODE, EDGE_NSTAGE, EDGE_ITEMS>;
  const size_t smem = (size_t)EDGE_NSTAGE * Ly::STAGE_BYTES;
  unsigned int blocks;
  if (!edge_grid(ctx, k, smem, &ctx->edge_occupancy[MODE][LOG2R], a->sites, Ly::TILE, &blocks)) return 0;
  double * partial = (double *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)blocks * 2 * sizeof(double));
  if (!partial) return 0;
  k<<<blocks, EDGE_THREADS, smem, ctx->stream>>>(*a, partial, ctx->d_ticket, dst, hdst);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

template <int LOG2R, int MODE>
static int launch_lk(plf_ctx * ctx, const plf_lk_t * a, double * dst, double * hdst)
{
  if (edge_items(ctx) == 2) return launch_lk_shape<LOG2R, MODE, 2, 6>(ctx, a, dst, hdst);
  return launch_lk_shape<LOG2R, MODE, 4, 3>(ctx, a, dst, hdst);
}

template <int LOG2R>
static int launch_lk_mode(plf_ctx * ctx, const plf_lk_t * a, double * dst, double * hdst)
{
  if (!a->pmatrix) return launch_lk<LOG2R, EDGE_ROOT>(ctx, a, dst, hdst);
  if (a->tipchars) return launch_lk<LOG2R, EDGE_TI>(ctx, a, dst, hdst);
  return launch_lk<LOG2R, EDGE_II>(ctx, a, dst, hdst);
}

/* returns -1 when the call is not eligible (the generic kernel takes it) */
int plf_loglikelihood_dna(plf_ctx * ctx, const plf_shape_t * sh, const plf_lk_t * a, double * dst, double * hdst)
{
  const unsigned int R = sh->rate_cats;
  if (sh->states != 4 || sh->per_rate_scalers || a->p_site_id || a->c_site_id || !R || (R & (R - 1)) || R > 8 ||
      !a->sites)
    return -1;
  switch (R)
  {
    case 1: return launch_lk_mode<0>(ctx, a, dst, hdst);
    case 2: return launch_lk_mode<1>(ctx, a, dst, hdst);
    case 4: return launch_lk_mode<2>(ctx, a, dst, hdst);
    default: return launch_lk_mode<3>(ctx, a, dst, hdst);
  }
}

template <int LOG2R, int EDGE_ITEMS, int EDGE_NSTAGE>
static int launch_deriv_shape(plf_ctx * ctx, const plf_deriv_t * a, double * dst, double * hdst)
{
  typedef EdgeLayout<LOG2R, EDGE_ITEMS, EDGE_DERIV> Ly;
  auto k = k_deriv_dna<LOG2R, EDGE_NSTAGE, EDGE_ITEMS>;
  const size_t smem = (size_t)EDGE_NSTAGE * Ly::STAGE_BYTES;
  unsigned int blocks;
  if (!edge_grid(ctx, k, smem, &ctx->edge_occupancy[EDGE_DERIV][LOG2R], a->sites, Ly::TILE, &blocks)) return 0;
  double * partial = (double *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)blocks * 2 * sizeof(double));
  if (!partial) return 0;
  k<<<blocks, EDGE_THREADS, smem, ctx->stream>>>(*a, partial, ctx->d_ticket, dst, hdst);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

template <int LOG2R>
static int launch_deriv(plf_ctx * ctx, const plf_deriv_t * a, double * dst, double * hdst)
{
  if (edge_items(ctx) == 2) return launch_deriv_shape<LOG2R, 2, 6>(ctx, a, dst, hdst);
  return launch_deriv_shape<LOG2R, 4, 6>(ctx, a, dst, hdst);
}

int plf_derivatives_dna(plf_ctx * ctx, const plf_shape_t * sh, const plf_deriv_t * a, double * dst, double * hdst)
{
  const unsigned int R = sh->rate_cats;
  if (sh->states != 4 || !R || (R & (R - 1)) || R > 8 || !a->sites) return -1;
  switch (R)
  {
    case 1: return launch_deriv<0>(ctx, a, dst, hdst);
    case 2: return launch_deriv<1>(ctx, a, dst, hdst);
    case 4: return launch_deriv<2>(ctx, a, dst, hdst);
    default: return launch_deriv<3>(ctx, a, dst, hdst);
  }
}
