/*
 * plf_likelihood.cu -- log-likelihood reductions, sumtable and derivative
 * kernels (sm_100a).
 *
 * Replaces pll_core_root_loglikelihood / pll_core_edge_loglikelihood_{ii,ti,
 * repeats} (reference src/core_likelihood.c:25,1192,581,924),
 * pll_core_update_sumtable_{ii,ti,repeats} (src/core_derivatives.c:321,473,25)
 * and pll_core_likelihood_derivatives (src/core_derivatives.c:696).
 *
 * Same thread mapping as the CLV kernels: one thread per (site, rate) when
 * rate_cats is a power of two <= 32, lanes of a site combine with shuffles;
 * per-site values are reduced with a fixed-shape tree (warp shuffles, then
 * shared memory, then a one-block pass over the per-block partials), so the
 * result is reproducible run to run.  The reference sums sites sequentially;
 * the difference is covered by the 1e-10 (logL) / 1e-9 (derivatives)
 * tolerances of the parity contract.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

/* model block accessors (layout in plf_backend.h) */
struct Model
{
  const double * rates, * weights, * pinv, * freqs, * evals, * evecs, * ievecs;
  __device__ Model(const double * m, int R, int st, int sp)
  {
    rates = m;
    weights = m + R;
    pinv = m + 2 * R;
    freqs = m + 3 * R;
    evals = freqs + (size_t)R * sp;
    evecs = evals + (size_t)R * sp;
    ievecs = evecs + (size_t)R * st * sp;
  }
};

/* ------------------------------------------------------------------------ *
 *  log-likelihood (root when a.pmatrix == NULL)                              *
 * ------------------------------------------------------------------------ */
template <int ST>
__global__ void __launch_bounds__(256)
k_loglik(plf_lk_t a, int R, int st_rt, int sp_rt, int per_rate, int L, double * __restrict__ partial,
         unsigned int * ticket, double * out, double * hout)
{
  __shared__ double red[32];
  const int st = ST ? ST : st_rt;
  const int sp = ST ? ((ST + 3) & ~3) : sp_rt;
  const Model M(a.model, R, st, sp);
  const int RT = R / L;
  const bool root = (a.pmatrix == nullptr);
  const bool tip = (a.tipchars != nullptr);
  const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int nthreads = gridDim.x * blockDim.x;
  const unsigned int lane_in_site = tid & (L - 1);
  const unsigned int sites_per_iter = nthreads / L;
  const unsigned int my_site0 = tid / L;
  const unsigned int warp_site0 = (tid & ~31u) / L;
  const size_t span = (size_t)sp * R;
  /* 2^-256k, k = 1..4 (core_likelihood.c:1366-1376) */
  const double minlh[4] = {0x1p-256, 0x1p-512, 0x1p-768, 0x1p-1024};

  double acc = 0;
  for (unsigned int s0 = warp_site0; s0 < a.sites; s0 += sites_per_iter)
  {
    const unsigned int n = s0 + (my_site0 - warp_site0);
    const bool active = n < a.sites;
    unsigned int pid = n, cid = n;
    plf_state_t mask = 0;
    int inv = -1;
    if (active)
    {
      if (a.p_site_id) pid = a.p_site_id[n];
      if (a.c_site_id) cid = a.c_site_id[n];
      if (tip)
      {
        const unsigned int code = a.tipchars[n];
        mask = (st == 4) ? (plf_state_t)code : a.tipmap[code];
      }
      if (a.invariant) inv = a.invariant[n];
    }
    /* scalers */
    unsigned int site_scalings = 0, my_min = 0xFFFFFFFFu;
    const bool use_rate_scalers = per_rate && !root;
    if (active)
    {
      if (use_rate_scalers)
      {
        for (int rr = 0; rr < RT; ++rr)
        {
          const int rate = lane_in_site * RT + rr;
          unsigned int s = (a.pscaler ? a.pscaler[(size_t)pid * R + rate] : 0u) +
                           (a.cscaler ? a.cscaler[(size_t)cid * R + rate] : 0u);
          my_min = s < my_min ? s : my_min;
        }
      }
      else
        site_scalings = (a.pscaler ? a.pscaler[pid] : 0u) + ((!root && a.cscaler) ? a.cscaler[cid] : 0u);
    }
    if (use_rate_scalers) site_scalings = group_min_u(my_min, L);

    double terma = 0, terminv = 0;
    if (active)
    {
      for (int rr = 0; rr < RT; ++rr)
      {
        const int rate = lane_in_site * RT + rr;
        const double * cp = a.clvp + (size_t)pid * span + (size_t)rate * sp;
        const double * freqs = M.freqs + (size_t)rate * sp;
        double term_r = 0;
        if (root)
        {
          for (int k = 0; k < st; ++k) term_r = fma(cp[k], freqs[k], term_r);
        }
        else
        {
          const double * pm = a.pmatrix + (size_t)rate * st * sp;
          const double * cc = tip ? nullptr : a.clvc + (size_t)cid * span + (size_t)rate * sp;
          for (int j = 0; j < st; ++j)
          {
            double termb = 0;
            if (tip)
            {
              for (int k = 0; k < st; ++k)
                if ((mask >> k) & 1ull) termb += pm[j * sp + k];
            }
            else
              for (int k = 0; k < st; ++k) termb = fma(pm[j * sp + k], cc[k], termb);
            term_r = fma(cp[j] * freqs[j], termb, term_r);
          }
          if (use_rate_scalers)
          {
            unsigned int s = (a.pscaler ? a.pscaler[(size_t)pid * R + rate] : 0u) +
                             (a.cscaler ? a.cscaler[(size_t)cid * R + rate] : 0u);
            unsigned int d = s - site_scalings;
            if (d > PLF_MAXDIFF) d = PLF_MAXDIFF;
            if (d > 0) term_r *= minlh[d - 1];
          }
        }
        const double pinv = M.pinv[rate], w = M.weights[rate];
        if (pinv > 0)
        {
          const double inv_lk = (inv == -1) ? 0.0 : freqs[inv];
          if (root)
            terma += w * (term_r * (1.0 - pinv) + inv_lk * pinv); /* core_likelihood.c:179-180 */
          else
          {
            terma += w * term_r * (1.0 - pinv);                   /* core_likelihood.c:1445-1452 */
            if (inv != -1) terminv += w * inv_lk * pinv;
          }
        }
        else
          terma += term_r * w;
      }
    }
    terma = group_sum(terma, L);
    terminv = group_sum(terminv, L);
    if (active && lane_in_site == 0)
    {
      double site_lk;
      if (root)
      {
        site_lk = log(terma);
        if (site_scalings) site_lk += site_scalings * PLF_LOG_SCALE_THRESHOLD;
      }
      else if (site_scalings)
      {
        if (terminv > 0.0)
        {
          const unsigned int capped = site_scalings < PLF_MAXDIFF ? site_scalings : PLF_MAXDIFF;
          site_lk = log(terma * minlh[capped - 1] + terminv);
        }
        else
          site_lk = log(terma) + site_scalings * PLF_LOG_SCALE_THRESHOLD;
      }
      else
        site_lk = log(terma + terminv);
      site_lk *= (double)a.pattern_weights[n];
      if (a.persite) a.persite[n] = site_lk;
      acc += site_lk;
    }
  }
  const double v[1] = {acc};
  grid_reduce_finish<1>(v, partial, ticket, out, hout, red);
}

static unsigned int pick_L(unsigned int R) { return (R && !(R & (R - 1)) && R <= 32) ? R : 1; }

static unsigned int pick_blocks(plf_ctx * ctx, unsigned long long work, int threads, int waves)
{
  unsigned long long b = (work + threads - 1) / threads;
  unsigned long long cap = (unsigned long long)ctx->sm_count * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned int)b;
}

/* results land in `dst` (device) and, when the caller wants them on the host,
 * in the pinned ctx->h_result written by the reducing block itself: the only
 * host-side cost is one stream synchronisation */
int plf_finish_reduction(plf_ctx * ctx, int nvals, double * h_out)
{
  if (h_out)
  {
    PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < nvals; ++i) h_out[i] = ctx->h_result[i];
  }
  return 1;
}

extern "C" int plf_loglikelihood(plf_ctx_t * ctx, const plf_shape_t * sh, const plf_lk_t * a, double * d_out,
                                 double * h_out)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  double * dst = d_out ? d_out : ctx->d_result;
  double * hdst = h_out ? ctx->h_result : nullptr;
  if (ctx->edge_fast)
  {
    int rc = plf_loglikelihood_dna(ctx, sh, a, dst, hdst);
    if (rc < 0 && ctx->aa_fast && ctx->aa_mma) rc = plf_loglikelihood_aa(ctx, sh, a, a->maxstates, dst, hdst);
    if (rc >= 0) return rc ? plf_finish_reduction(ctx, 1, h_out) : 0;
  }
  const int R = (int)sh->rate_cats;
  const int L = (int)pick_L(sh->rate_cats);
  const int threads = 256;
  const unsigned int blocks = pick_blocks(ctx, (unsigned long long)a->sites * L, threads, 8);
  double * partial = (double *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)blocks * 2 * sizeof(double));
  if (!partial) return 0;
  if (sh->states == 4)
    k_loglik<4><<<blocks, threads, 0, ctx->stream>>>(*a, R, 4, 4, sh->per_rate_scalers, L, partial, ctx->d_ticket,
                                                    dst, hdst);
  else if (sh->states == 20)
    k_loglik<20><<<blocks, threads, 0, ctx->stream>>>(*a, R, 20, 20, sh->per_rate_scalers, L, partial,
                                                     ctx->d_ticket, dst, hdst);
  else
    k_loglik<0><<<blocks, threads, 0, ctx->stream>>>(*a, R, (int)sh->states, (int)sh->states_padded,
                                                    sh->per_rate_scalers, L, partial, ctx->d_ticket, dst, hdst);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return plf_finish_reduction(ctx, 1, h_out);
}

/* ------------------------------------------------------------------------ *
 *  sumtable: sum[n][r][j] = (sum_k clvp_k pi_k Vinv_kj) (sum_k V_jk clvc_k)  *
 *            * 2^(-256 min(dscaler_r, 4))   (core_derivatives.c:418-465)     *
 * ------------------------------------------------------------------------ */
template <int ST>
__global__ void __launch_bounds__(256)
k_sumtable(plf_sumtable_t a, int R, int st_rt, int sp_rt, int per_rate, int L)
{
  const int st = ST ? ST : st_rt;
  const int sp = ST ? ((ST + 3) & ~3) : sp_rt;
  const Model M(a.model, R, st, sp);
  const int RT = R / L;
  const bool tip = (a.tipchars != nullptr);
  const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int nthreads = gridDim.x * blockDim.x;
  const unsigned int lane_in_site = tid & (L - 1);
  const unsigned int sites_per_iter = nthreads / L;
  const unsigned int my_site0 = tid / L;
  const unsigned int warp_site0 = (tid & ~31u) / L;
  const size_t span = (size_t)sp * R;
  const double minlh[4] = {0x1p-256, 0x1p-512, 0x1p-768, 0x1p-1024};

  for (unsigned int s0 = warp_site0; s0 < a.sites; s0 += sites_per_iter)
  {
    const unsigned int n = s0 + (my_site0 - warp_site0);
    const bool active = n < a.sites;
    unsigned int pid = n, cid = n;
    plf_state_t mask = 0;
    if (active)
    {
      if (a.p_site_id) pid = a.p_site_id[n];
      if (a.c_site_id) cid = a.c_site_id[n];
      if (tip)
      {
        const unsigned int code = a.tipchars[n];
        mask = (st == 4) ? (plf_state_t)code : a.tipmap[code];
      }
    }
    unsigned int my_min = 0xFFFFFFFFu, site_min = 0;
    if (per_rate)
    {
      if (active)
        for (int rr = 0; rr < RT; ++rr)
        {
          const int rate = lane_in_site * RT + rr;
          unsigned int s = ((a.pscaler && !tip) ? a.pscaler[(size_t)pid * R + rate] : 0u) +
                           (a.cscaler ? a.cscaler[(size_t)cid * R + rate] : 0u);
          my_min = s < my_min ? s : my_min;
        }
      site_min = group_min_u(my_min, L);
    }
    if (!active) continue;
    for (int rr = 0; rr < RT; ++rr)
    {
      const int rate = lane_in_site * RT + rr;
      const double * cp = tip ? nullptr : a.clvp + (size_t)pid * span + (size_t)rate * sp;
      const double * cc = a.clvc + (size_t)cid * span + (size_t)rate * sp;
      const double * f = M.freqs + (size_t)rate * sp;
      const double * ev = M.evecs + (size_t)rate * st * sp;
      const double * iev = M.ievecs + (size_t)rate * st * sp;
      double * out = a.sumtable + (size_t)n * span + (size_t)rate * sp;
      double scale = 1.0;
      if (per_rate)
      {
        unsigned int s = ((a.pscaler && !tip) ? a.pscaler[(size_t)pid * R + rate] : 0u) +
                         (a.cscaler ? a.cscaler[(size_t)cid * R + rate] : 0u);
        unsigned int d = s - site_min;
        if (d > PLF_MAXDIFF) d = PLF_MAXDIFF;
        if (d > 0) scale = minlh[d - 1];
      }
      for (int j = 0; j < st; ++j)
      {
        double l = 0, r = 0;
        for (int k = 0; k < st; ++k)
        {
          const double lk = tip ? (((mask >> k) & 1ull) ? f[k] : 0.0) : cp[k] * f[k];
          l = fma(lk, iev[k * sp + j], l);
          r = fma(ev[j * sp + k], cc[k], r);
        }
        double v = l * r;
        if (scale != 1.0) v *= scale;
        out[j] = v;
      }
      for (int j = st; j < sp; ++j) out[j] = 0.0;
    }
  }
}

extern "C" int plf_update_sumtable(plf_ctx_t * ctx, const plf_shape_t * sh, const plf_sumtable_t * a)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const int R = (int)sh->rate_cats;
  /* sum[j] = (sum_k clvp_k pi_k Vinv_kj)(sum_k V_jk clvc_k) is a CLV update without scaling:
   * 4 states run it on the DNA CLV kernels, 20 states on the DMMA kernels (plf_edge_dna.cu) */
  if (ctx->edge_fast && !sh->per_rate_scalers && a->sites &&
      ((sh->states == 4 && R > 0 && !(R & (R - 1)) && R <= 32) || (sh->states == 20 && ctx->aa_fast && ctx->aa_mma)))
  {
    const int rc = plf_sumtable_as_clv(ctx, sh, a, a->tipmap, a->maxstates);
    if (rc >= 0) return rc;
  }
  const int L = (int)pick_L(sh->rate_cats);
  const int threads = 256;
  const unsigned int blocks = pick_blocks(ctx, (unsigned long long)a->sites * L, threads, 16);
  if (sh->states == 4)
    k_sumtable<4><<<blocks, threads, 0, ctx->stream>>>(*a, R, 4, 4, sh->per_rate_scalers, L);
  else if (sh->states == 20)
    k_sumtable<20><<<blocks, threads, 0, ctx->stream>>>(*a, R, 20, 20, sh->per_rate_scalers, L);
  else
    k_sumtable<0><<<blocks, threads, 0, ctx->stream>>>(*a, R, (int)sh->states, (int)sh->states_padded,
                                                      sh->per_rate_scalers, L);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* per-thread share of sum_sites w (-L'/L) and sum_sites w ((L'/L)^2 - L''/L) for the diag table in shared memory */
template <int ST>
__device__ __forceinline__ void deriv_accumulate(const plf_deriv_t & a, const Model & M, const double * diag, int R, int st,
                                                 int sp, int L, double & acc1, double & acc2)
{
  const int RT = R / L;
  const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int nthreads = gridDim.x * blockDim.x;
  const unsigned int lane_in_site = tid & (L - 1);
  const unsigned int sites_per_iter = nthreads / L;
  const unsigned int my_site0 = tid / L;
  const unsigned int warp_site0 = (tid & ~31u) / L;
  const size_t span = (size_t)sp * R;

  acc1 = 0;
  acc2 = 0;
  if (ST == 4 && RT == 1)
  {
    /* 4 states, one rate per lane: the loads of four sweeps of the grid are issued before the first use (the
     * loads are volatile asm and would otherwise go out one at a time); sums are added in the same order as
     * by the plain loop below, so the result has the same bits */
    const int rate = lane_in_site;
    const double * d = diag + (size_t)rate * 4 * 3;
    const double pinv = M.pinv[rate], wr = M.weights[rate];
    for (unsigned int s0 = warp_site0; s0 < a.sites; s0 += 4 * sites_per_iter)
    {
      dbl4 sv[4];
      double pw[4];
      int inv[4];
      bool active[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const unsigned int n = s0 + u * sites_per_iter + (my_site0 - warp_site0);
        active[u] = n < a.sites;
        sv[u] = dbl4{0, 0, 0, 0};
        pw[u] = 0;
        inv[u] = -1;
        if (active[u])
        {
          sv[u] = ld256_stream(a.sumtable + (size_t)n * span + (size_t)rate * 4);
          if (lane_in_site == 0) pw[u] = (double)a.pattern_weights[n];
          if (a.invariant) inv[u] = a.invariant[n];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        /* warps past the end of the alignment leave together: the shuffles below stay converged */
        if (s0 + u * sites_per_iter >= a.sites) break;
        const double x[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
        double c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          c0 = fma(x[j], d[j * 3 + 0], c0);
          c1 = fma(x[j], d[j * 3 + 1], c1);
          c2 = fma(x[j], d[j * 3 + 2], c2);
        }
        if (pinv > 0)
        {
          const double inv_lk = (inv[u] == -1) ? 0.0 : M.freqs[(size_t)rate * sp + inv[u]] * pinv;
          c0 = c0 * (1.0 - pinv) + inv_lk;
          c1 = c1 * (1.0 - pinv);
          c2 = c2 * (1.0 - pinv);
        }
        double lk0 = active[u] ? fma(c0, wr, 0.0) : 0.0;
        double lk1 = active[u] ? fma(c1, wr, 0.0) : 0.0;
        double lk2 = active[u] ? fma(c2, wr, 0.0) : 0.0;
        lk0 = group_sum(lk0, L);
        lk1 = group_sum(lk1, L);
        lk2 = group_sum(lk2, L);
        if (active[u] && lane_in_site == 0)
        {
          const double d1 = -lk1 / lk0;
          const double d2 = d1 * d1 - lk2 / lk0;
          acc1 = fma(pw[u], d1, acc1);
          acc2 = fma(pw[u], d2, acc2);
        }
      }
    }
    return;
  }
  for (unsigned int s0 = warp_site0; s0 < a.sites; s0 += sites_per_iter)
  {
    const unsigned int n = s0 + (my_site0 - warp_site0);
    const bool active = n < a.sites;
    double lk0 = 0, lk1 = 0, lk2 = 0;
    if (active)
    {
      const int inv = a.invariant ? a.invariant[n] : -1;
      for (int rr = 0; rr < RT; ++rr)
      {
        const int rate = lane_in_site * RT + rr;
        const double * sum = a.sumtable + (size_t)n * span + (size_t)rate * sp;
        const double * d = diag + (size_t)rate * st * 3;
        double c0 = 0, c1 = 0, c2 = 0;
        if (ST == 4)
        {
          const dbl4 s = ld256_stream(sum);
          const double sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
          {
            c0 = fma(sv[j], d[j * 3 + 0], c0);
            c1 = fma(sv[j], d[j * 3 + 1], c1);
            c2 = fma(sv[j], d[j * 3 + 2], c2);
          }
        }
        else
          for (int j = 0; j < st; ++j)
          {
            const double s = sum[j];
            c0 = fma(s, d[j * 3 + 0], c0);
            c1 = fma(s, d[j * 3 + 1], c1);
            c2 = fma(s, d[j * 3 + 2], c2);
          }
        const double pinv = M.pinv[rate], w = M.weights[rate];
        if (pinv > 0)
        {
          const double inv_lk = (inv == -1) ? 0.0 : M.freqs[(size_t)rate * sp + inv] * pinv;
          c0 = c0 * (1.0 - pinv) + inv_lk;
          c1 = c1 * (1.0 - pinv);
          c2 = c2 * (1.0 - pinv);
        }
        lk0 = fma(c0, w, lk0);
        lk1 = fma(c1, w, lk1);
        lk2 = fma(c2, w, lk2);
      }
    }
    lk0 = group_sum(lk0, L);
    lk1 = group_sum(lk1, L);
    lk2 = group_sum(lk2, L);
    if (active && lane_in_site == 0)
    {
      const double w = (double)a.pattern_weights[n];
      const double d1 = -lk1 / lk0;
      const double d2 = d1 * d1 - lk2 / lk0;
      acc1 = fma(w, d1, acc1);
      acc2 = fma(w, d2, acc2);
    }
  }
}

/* diag[r][j] = {e, lk e, (lk)^2 e} for branch length t (every thread of the block takes part) */
__device__ __forceinline__ void deriv_diag_table(const Model & M, double * diag, int R, int st, int sp, double t)
{
  for (int x = threadIdx.x; x < R * st; x += blockDim.x)
  {
    const int r = x / st, j = x % st;
    const double lam = M.evals[(size_t)r * sp + j];
    const double ki = M.rates[r] / (1.0 - M.pinv[r]);
    const double e = exp(lam * ki * t);
    diag[x * 3 + 0] = e;
    diag[x * 3 + 1] = lam * ki * e;
    diag[x * 3 + 2] = lam * ki * lam * ki * e;
  }
}

/* ------------------------------------------------------------------------ *
 *  derivatives: per site L, L', L'' from the sumtable and                    *
 *  diag[r][j] = {e, lk e, (lk)^2 e}, e = exp(lambda_j k_r t),                 *
 *  k_r = rate_r / (1 - pinv_r)        (core_derivatives.c:757-772,825-848)   *
 * ------------------------------------------------------------------------ */
template <int ST>
__global__ void __launch_bounds__(256)
k_derivatives(plf_deriv_t a, int R, int st_rt, int sp_rt, int L, double * __restrict__ partial,
              unsigned int * ticket, double * out, double * hout)
{
  extern __shared__ double diag[]; /* [R][st][3] */
  __shared__ double red[32];
  const int st = ST ? ST : st_rt;
  const int sp = ST ? ((ST + 3) & ~3) : sp_rt;
  const Model M(a.model, R, st, sp);
  deriv_diag_table(M, diag, R, st, sp, a.branch_length);
  __syncthreads();
  double acc1, acc2;
  deriv_accumulate<ST>(a, M, diag, R, st, sp, L, acc1, acc2);
  const double v[2] = {acc1, acc2};
  grid_reduce_finish<2>(v, partial, ticket, out, hout, red);
}

extern "C" int plf_derivatives(plf_ctx_t * ctx, const plf_shape_t * sh, const plf_deriv_t * a, double * d_out2,
                               double * h_out2)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  double * dst = d_out2 ? d_out2 : ctx->d_result;
  double * hdst = h_out2 ? ctx->h_result : nullptr;
  if (ctx->edge_fast)
  {
    const int rc = plf_derivatives_dna(ctx, sh, a, dst, hdst);
    if (rc >= 0) return rc ? plf_finish_reduction(ctx, 2, h_out2) : 0;
  }
  const int R = (int)sh->rate_cats;
  const int L = (int)pick_L(sh->rate_cats);
  const int threads = 256;
  const unsigned int blocks = pick_blocks(ctx, (unsigned long long)a->sites * L, threads, 8);
  double * partial = (double *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)blocks * 2 * sizeof(double));
  if (!partial) return 0;
  const size_t smem = (size_t)R * sh->states * 3 * sizeof(double);
  if (smem > 48 * 1024)
  {
    plf_set_error(ctx, "derivatives: rate_cats*states too large for the diag table (%zu B)", smem);
    return 0;
  }
  if (sh->states == 4)
    k_derivatives<4><<<blocks, threads, smem, ctx->stream>>>(*a, R, 4, 4, L, partial, ctx->d_ticket, dst, hdst);
  else if (sh->states == 20)
    k_derivatives<20><<<blocks, threads, smem, ctx->stream>>>(*a, R, 20, 20, L, partial, ctx->d_ticket, dst, hdst);
  else
    k_derivatives<0><<<blocks, threads, smem, ctx->stream>>>(*a, R, (int)sh->states, (int)sh->states_padded, L,
                                                            partial, ctx->d_ticket, dst, hdst);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return plf_finish_reduction(ctx, 2, h_out2);
}

/* ------------------------------------------------------------------------ *
 *  Newton-Raphson on one branch, entirely on the device (the loop of          *
 *  examples/newton/newton.c:67-96, which calls                                *
 *  pll_compute_likelihood_derivatives up to 32 times per branch and pays a    *
 *  host round trip each time).  One cooperative launch: every iteration        *
 *  re-reads the sumtable (L2-resident up to ~1M DNA sites), leaves per-block   *
 *  partial sums in a ping-pong buffer, crosses ONE grid barrier, and every     *
 *  block then adds all partials in the same fixed order, so all blocks take    *
 *  the same step and agree on convergence without another exchange.            *
 *  out[4] = {length, d_f, dd_f, iterations}.                                   *
 * ------------------------------------------------------------------------ */
struct plf_newton_args
{
  double t0, tmin, tmax, tolerance;
  unsigned int max_iters;
};

/* two deterministic block sums for the price of one (results valid in thread 0); red >= 64 doubles */
__device__ __forceinline__ void block_sum2(double & a, double & b, double * red)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0)
  {
    red[w] = a;
    red[32 + w] = b;
  }
  __syncthreads();
  if (w == 0)
  {
    const int nw = (blockDim.x + 31) >> 5;
    a = warp_sum(lane < nw ? red[lane] : 0.0);
    b = warp_sum(lane < nw ? red[32 + lane] : 0.0);
  }
}

template <int ST>
__global__ void __launch_bounds__(512)
k_newton(plf_deriv_t a, int R, int st_rt, int sp_rt, int L, double * __restrict__ partial, plf_newton_args nw,
         double * out, double * hout)
{
  extern __shared__ double diag2[]; /* two tables [R][st][3], used alternately */
  __shared__ double red[64];
  __shared__ double s_sum[2];
  cg::grid_group grid = cg::this_grid();
  const int st = ST ? ST : st_rt;
  const int sp = ST ? ((ST + 3) & ~3) : sp_rt;
  const Model M(a.model, R, st, sp);
  double t = nw.t0, d1 = 0, d2 = 0;
  unsigned int iters = 0;
  for (unsigned int it = 0; it < nw.max_iters; ++it)
  {
    double * buf = partial + (size_t)(it & 1u) * 2 * gridDim.x;
    double * diag = diag2 + (size_t)(it & 1u) * R * st * 3;
    deriv_diag_table(M, diag, R, st, sp, t);
    __syncthreads();
    double acc1, acc2;
    deriv_accumulate<ST>(a, M, diag, R, st, sp, L, acc1, acc2);
    block_sum2(acc1, acc2, red);
    if (threadIdx.x == 0)
    {
      buf[blockIdx.x] = acc1;
      buf[gridDim.x + blockIdx.x] = acc2;
    }
    grid.sync();
    /* the same additions in the same order in every block */
    double p1 = 0, p2 = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
    {
      p1 += __ldcg(buf + b);
      p2 += __ldcg(buf + gridDim.x + b);
    }
    block_sum2(p1, p2, red);
    if (threadIdx.x == 0)
    {
      s_sum[0] = p1;
      s_sum[1] = p2;
    }
    __syncthreads();
    d1 = s_sum[0];
    d2 = s_sum[1];
    iters = it + 1;
    if (fabs(d1) < nw.tolerance) break;
    double tn = t - d1 / d2;
    if (tn < nw.tmin) tn = nw.tmin;
    if (tn > nw.tmax) tn = nw.tmax;
    if (!(tn == tn) || tn == t) break; /* NaN step, or pinned at a bound */
    t = tn;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    const double r[4] = {t, d1, d2, (double)iters};
    for (int i = 0; i < 4; ++i)
    {
      out[i] = r[i];
      if (hout) hout[i] = r[i];
    }
  }
}

template <int ST>
static int newton_launch(plf_ctx * ctx, const plf_deriv_t * a, int R, int st, int sp, int L, const plf_newton_args & nw,
                         size_t smem, double * dst, double * hdst)
{
  /* few, large blocks: the grid barrier and the re-reduction of the partials grow with the block count
   * (PLF_NEWTON_THREADS / PLF_NEWTON_BPS override for experiments) */
  static int threads = 0, bps = 0;
  if (!threads)
  {
    const char * v = getenv("PLF_NEWTON_THREADS");
    const int t = v ? atoi(v) : 0;
    threads = (t == 256 || t == 512) ? t : 512;
    v = getenv("PLF_NEWTON_BPS");
    bps = (v && atoi(v) > 0) ? atoi(v) : 2;
  }
  int per_sm = 0;
  PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_newton<ST>, threads, smem));
  if (per_sm < 1)
  {
    plf_set_error(ctx, "newton: kernel does not fit an SM");
    return 0;
  }
  if (per_sm > bps) per_sm = bps;
  unsigned int blocks = pick_blocks(ctx, (unsigned long long)a->sites * L, threads, per_sm);
  double * partial = (double *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)blocks * 4 * sizeof(double));
  if (!partial) return 0;
  plf_deriv_t args = *a;
  plf_newton_args nwa = nw;
  void * params[] = {&args, &R, &st, &sp, &L, &partial, &nwa, &dst, &hdst};
  PLF_CHECK(ctx, cudaLaunchCooperativeKernel((const void *)k_newton<ST>, dim3(blocks), dim3(threads), params, smem,
                                             ctx->stream));
  plf_count_launch();
  return 1;
}

/* h_out4 = {length, d_f, dd_f, iterations}; synchronises the stream */
extern "C" int plf_newton_branch(plf_ctx_t * ctx, const plf_shape_t * sh, const plf_deriv_t * a, double t0, double tmin,
                                 double tmax, double tolerance, unsigned int max_iters, double * h_out4)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const int R = (int)sh->rate_cats;
  const int L = (int)pick_L(sh->rate_cats);
  const size_t smem = (size_t)2 * R * sh->states * 3 * sizeof(double);
  if (smem > 48 * 1024)
  {
    plf_set_error(ctx, "newton: rate_cats*states too large for the diag table (%zu B)", smem);
    return 0;
  }
  plf_newton_args nw = {t0, tmin, tmax, tolerance, max_iters};
  int ok;
  if (sh->states == 4)
    ok = newton_launch<4>(ctx, a, R, 4, 4, L, nw, smem, ctx->d_result, ctx->h_result);
  else if (sh->states == 20)
    ok = newton_launch<20>(ctx, a, R, 20, 20, L, nw, smem, ctx->d_result, ctx->h_result);
  else
    ok = newton_launch<0>(ctx, a, R, (int)sh->states, (int)sh->states_padded, L, nw, smem, ctx->d_result,
                          ctx->h_result);
  if (!ok) return 0;
  return plf_finish_reduction(ctx, 4, h_out4);
}
