/*
 * plf_oracle.c -- TEST INFRASTRUCTURE ONLY (see plf_oracle.h).
 *
 * Scalar restatement of the arithmetic the reference performs under
 * PLL_ATTRIB_ARCH_AVX2.  Compile with -ffp-contract=off: every fused
 * multiply-add below is an explicit fma(), everything else rounds once per
 * operation, which is what the reference's intrinsics do.
 *
 * Three arithmetic "profiles", chosen by the state count exactly as the
 * reference's dispatchers do (core_partials_avx2.c:74-107,1031-1076):
 *   4 states : plain multiplies + pairwise tree (p0+p1)+(p2+p3), NO fma
 *              (core_partials_avx.c:456-524; files built with -mavx only)
 *   otherwise: four lane accumulators, lane l chains fma over columns
 *              l, l+4, l+8, ... starting from 0, then (a0+a1)+(a2+a3)
 *              (core_partials_avx2.c:695-771 for 20, :1114-1227 generic)
 * Tip-side sums differ again (masked pairwise / scalar / lane adds), see the
 * comments at each function.
 */
#include "plf_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SCALE_FACTOR 0x1p+256
#define SCALE_THRESHOLD 0x1p-256
#define MAXDIFF 4
#define EMPTY 0xFFFFFFFFu

/* ---- dot products ------------------------------------------------------ */

/* 4-state row times vector: (m0*c0 + m1*c1) + (m2*c2 + m3*c3) */
static double dot4_pairwise(const double * m, const double * c)
{
  double p0 = m[0] * c[0], p1 = m[1] * c[1];
  double p2 = m[2] * c[2], p3 = m[3] * c[3];
  return (p0 + p1) + (p2 + p3);
}

/* lane-chained fma dot product over `n` (multiple of 4) columns */
static double dot_lanes_fma(const double * m, const double * c, unsigned int n)
{
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  unsigned int j;
  for (j = 0; j < n; j += 4)
  {
    a0 = fma(m[j + 0], c[j + 0], a0);
    a1 = fma(m[j + 1], c[j + 1], a1);
    a2 = fma(m[j + 2], c[j + 2], a2);
    a3 = fma(m[j + 3], c[j + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}

/* sum of the matrix-row entries selected by `mask`:
 * 4 states: masked pairwise (core_partials_avx.c:1355-1395) */
static double masked_sum4(const double * m, unsigned int mask)
{
  double p0 = (mask & 1) ? m[0] : 0.0, p1 = (mask & 2) ? m[1] : 0.0;
  double p2 = (mask & 4) ? m[2] : 0.0, p3 = (mask & 8) ? m[3] : 0.0;
  return (p0 + p1) + (p2 + p3);
}

/* scalar, increasing column order (core_partials_avx2.c:387-456,
 * core_partials_avx.c:63-92) */
static double masked_sum_seq(const double * m, orc_state_t mask,
                             unsigned int states)
{
  double t = 0;
  unsigned int k;
  for (k = 0; k < states; ++k)
    if ((mask >> k) & 1)
      t += m[k];
  return t;
}

/* lane adds over column quads, then pairwise (core_partials_avx2.c:159-185) */
static double masked_sum_lanes(const double * m, orc_state_t mask,
                               unsigned int n)
{
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  unsigned int j;
  for (j = 0; j < n; j += 4)
  {
    if ((mask >> j) & 0xF)
    {
      a0 = a0 + (((mask >> (j + 0)) & 1) ? m[j + 0] : 0.0);
      a1 = a1 + (((mask >> (j + 1)) & 1) ? m[j + 1] : 0.0);
      a2 = a2 + (((mask >> (j + 2)) & 1) ? m[j + 2] : 0.0);
      a3 = a3 + (((mask >> (j + 3)) & 1) ? m[j + 3] : 0.0);
    }
  }
  return (a0 + a1) + (a2 + a3);
}

static double row_dot(unsigned int states, unsigned int sp, const double * m,
                      const double * c)
{
  return states == 4 ? dot4_pairwise(m, c) : dot_lanes_fma(m, c, sp);
}

/* ---- P-matrices -------------------------------------------------------- */

void orc_update_pmatrix(double ** pmatrix, unsigned int states,
                        unsigned int sp, unsigned int rate_cats,
                        const double * rates, const double * branch_lengths,
                        const unsigned int * matrix_indices,
                        const unsigned int * params_indices,
                        const double * prop_invar, double * const * eigenvals,
                        double * const * eigenvecs,
                        double * const * inv_eigenvecs, unsigned int count)
{
  unsigned int i, n, j, k, m;
  double * expd = (double *)malloc(sp * sizeof(double));
  double * temp = (double *)malloc((size_t)states * sp * sizeof(double));

  for (i = 0; i < count; ++i)
  {
    double t = branch_lengths[i];
    for (n = 0; n < rate_cats; ++n)
    {
      double * pmat = pmatrix[matrix_indices[i]] + (size_t)n * states * sp;
      double pinv = prop_invar[params_indices[n]];
      const double * evecs = eigenvecs[params_indices[n]];
      const double * ievecs = inv_eigenvecs[params_indices[n]];
      const double * evals = eigenvals[params_indices[n]];

      if (!(t > 0.))
      {
        /* identity; padded columns zero for 4/20, untouched otherwise
         * (core_pmatrix.c:243-248 writes only k < states) */
        for (j = 0; j < states; ++j)
          for (k = 0; k < ((states == 4 || states == 20) ? sp : states); ++k)
            pmat[j * sp + k] = (j == k) ? 1. : 0.;
        continue;
      }

      /* x = (lambda*rate)*t [ /(1-pinv) ]; e = expm1(x) */
      for (j = 0; j < states; ++j)
      {
        double x;
        if (states == 4 || states == 20)
        {
          x = (evals[j] * rates[n]) * t;          /* core_pmatrix_avx.c:99-113 */
          if (pinv > 1e-8) x = x / (1.0 - pinv);
        }
        else
        {
          x = evals[j] * rates[n] * t;            /* core_pmatrix.c:206-216 */
          if (pinv > 1e-8) x = x / (1.0 - pinv);
        }
        expd[j] = expm1(x);
      }

      if (states == 4)
      {
        /* tmp_jm = ievecs_jm * e_m ; P_jk = pairwise_m(tmp_jm*evecs_mk) + I */
        for (j = 0; j < 4; ++j)
        {
          double tm[4];
          for (m = 0; m < 4; ++m) tm[m] = ievecs[j * 4 + m] * expd[m];
          for (k = 0; k < 4; ++k)
          {
            double p0 = tm[0] * evecs[0 * 4 + k], p1 = tm[1] * evecs[1 * 4 + k];
            double p2 = tm[2] * evecs[2 * 4 + k], p3 = tm[3] * evecs[3 * 4 + k];
            pmat[j * 4 + k] = ((p0 + p1) + (p2 + p3)) + ((j == k) ? 1.0 : 0.0);
          }
        }
      }
      else if (states == 20)
      {
        /* core_pmatrix_avx2.c:212-282 */
        for (j = 0; j < 20; ++j)
          for (m = 0; m < 20; ++m)
            temp[j * 20 + m] = expd[m] * ievecs[j * 20 + m];
        for (j = 0; j < 20; ++j)
          for (k = 0; k < 20; ++k)
          {
            double a[4];
            unsigned int l, q;
            for (l = 0; l < 4; ++l)
            {
              a[l] = temp[j * 20 + l] * evecs[l * 20 + k];
              for (q = 1; q < 5; ++q)
                a[l] = fma(temp[j * 20 + l + 4 * q], evecs[(l + 4 * q) * 20 + k], a[l]);
            }
            pmat[j * 20 + k] = (a[0] + a[1]) + (a[2] + a[3]);
          }
        for (j = 0; j < 20; ++j) pmat[j * 20 + j] += 1.0;
      }
      else
      {
        /* core_pmatrix.c:218-240: start from identity, add sequentially */
        for (j = 0; j < states; ++j)
          for (k = 0; k < states; ++k)
            temp[j * states + k] = ievecs[j * sp + k] * expd[k];
        for (j = 0; j < states; ++j)
          for (k = 0; k < states; ++k)
          {
            double acc = (j == k) ? 1.0 : 0;
            for (m = 0; m < states; ++m)
              acc += temp[j * states + m] * evecs[m * sp + k];
            pmat[j * sp + k] = acc;
          }
      }
    }
  }
  free(expd);
  free(temp);
}

/* ---- scaling helpers (Appendix A.3 of SURVEY; core_partials_avx.c:526) -- */

static unsigned int scaler_sum(const unsigned int * a, const unsigned int * b,
                               size_t ia, size_t ib)
{
  return (a ? a[ia] : 0) + (b ? b[ib] : 0);
}

/* After the span of one site has been written unscaled: apply the scaling
 * rule.  below[k] says whether every entry (padded lanes included) of rate k
 * was < 2^-256. */
static void apply_scaling(double * site_clv, unsigned int sp,
                          unsigned int rate_cats, const int * below,
                          unsigned int * parent_scaler, size_t n,
                          int per_rate)
{
  unsigned int k, i;
  if (!parent_scaler) return;
  if (per_rate)
  {
    for (k = 0; k < rate_cats; ++k)
      if (below[k])
      {
        for (i = 0; i < sp; ++i) site_clv[k * sp + i] *= SCALE_FACTOR;
        parent_scaler[n * rate_cats + k] += 1;
      }
  }
  else
  {
    int all = 1;
    for (k = 0; k < rate_cats; ++k) all = all && below[k];
    if (all)
    {
      for (i = 0; i < sp * rate_cats; ++i) site_clv[i] *= SCALE_FACTOR;
      parent_scaler[n] += 1;
    }
  }
}

/* number of parent rows the reference computes: the AVX2 generic kernels
 * treat the matrix as states_padded x states_padded and let the padded rows
 * run into the following matrix ("displacement", core_partials_avx2.c:1101);
 * their products land in the padded CLV lanes and take part in the scaling
 * test.  The 4- and 20-state kernels have no padding. */
static unsigned int rows_computed(unsigned int states, unsigned int sp)
{
  (void)states;
  return sp;
}

/* ---- CLV updates ------------------------------------------------------- */

static void partial_core(unsigned int states, unsigned int sp,
                         unsigned int rate_cats, double * pclv,
                         const double * lclv, const double * rclv,
                         const double * lmat, const double * rmat,
                         int * below)
{
  unsigned int k, i, rows = rows_computed(states, sp);
  for (k = 0; k < rate_cats; ++k)
  {
    const double * lm = lmat + (size_t)k * states * sp;
    const double * rm = rmat + (size_t)k * states * sp;
    int b = 1;
    for (i = 0; i < rows; ++i)
    {
      double a = row_dot(states, sp, lm + (size_t)i * sp, lclv + k * sp);
      double c = row_dot(states, sp, rm + (size_t)i * sp, rclv + k * sp);
      double v = a * c;
      pclv[k * sp + i] = v;
      b = b && (v < SCALE_THRESHOLD);
    }
    below[k] = b;
  }
}

void orc_update_partial_ii(unsigned int states, unsigned int sp,
                           unsigned int sites, unsigned int rate_cats,
                           double * parent_clv, unsigned int * parent_scaler,
                           const double * left_clv, const double * right_clv,
                           const double * left_matrix,
                           const double * right_matrix,
                           const unsigned int * left_scaler,
                           const unsigned int * right_scaler, int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int k;
  int * below = (int *)malloc(rate_cats * sizeof(int));
  for (n = 0; n < sites; ++n)
  {
    if (parent_scaler)
    {
      if (per_rate)
        for (k = 0; k < rate_cats; ++k)
          parent_scaler[n * rate_cats + k] =
              scaler_sum(left_scaler, right_scaler, n * rate_cats + k,
                         n * rate_cats + k);
      else
        parent_scaler[n] = scaler_sum(left_scaler, right_scaler, n, n);
    }
    partial_core(states, sp, rate_cats, parent_clv + n * span,
                 left_clv + n * span, right_clv + n * span, left_matrix,
                 right_matrix, below);
    apply_scaling(parent_clv + n * span, sp, rate_cats, below, parent_scaler,
                  n, per_rate);
  }
  free(below);
}

void orc_update_partial_repeats(unsigned int states, unsigned int sp,
                                unsigned int parent_sites,
                                unsigned int rate_cats, double * parent_clv,
                                unsigned int * parent_scaler,
                                const double * left_clv,
                                const double * right_clv,
                                const double * left_matrix,
                                const double * right_matrix,
                                const unsigned int * left_scaler,
                                const unsigned int * right_scaler,
                                const unsigned int * parent_id_site,
                                const unsigned int * left_site_id,
                                const unsigned int * right_site_id,
                                int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int k;
  int * below = (int *)malloc(rate_cats * sizeof(int));
  for (n = 0; n < parent_sites; ++n)
  {
    size_t site = parent_id_site ? parent_id_site[n] : n;
    size_t lid = left_site_id ? left_site_id[site] : site;
    size_t rid = right_site_id ? right_site_id[site] : site;
    if (parent_scaler)
    {
      /* repeats.c:392-540 */
      if (per_rate)
        for (k = 0; k < rate_cats; ++k)
          parent_scaler[n * rate_cats + k] =
              scaler_sum(left_scaler, right_scaler, lid * rate_cats + k,
                         rid * rate_cats + k);
      else
        parent_scaler[n] = scaler_sum(left_scaler, right_scaler, lid, rid);
    }
    partial_core(states, sp, rate_cats, parent_clv + n * span,
                 left_clv + lid * span, right_clv + rid * span, left_matrix,
                 right_matrix, below);
    apply_scaling(parent_clv + n * span, sp, rate_cats, below, parent_scaler,
                  n, per_rate);
  }
  free(below);
}

/* tip-side term for parent row i of rate k */
static double tip_term(unsigned int states, unsigned int sp,
                       const double * row, orc_state_t mask)
{
  if (states == 4) return masked_sum4(row, (unsigned int)mask);
  if (states == 20) return masked_sum_seq(row, mask, states);
  return masked_sum_lanes(row, mask, sp);
}

void orc_update_partial_ti(unsigned int states, unsigned int sp,
                           unsigned int sites, unsigned int rate_cats,
                           double * parent_clv, unsigned int * parent_scaler,
                           const unsigned char * left_tipchars,
                           const double * right_clv,
                           const double * left_matrix,
                           const double * right_matrix,
                           const unsigned int * right_scaler,
                           const orc_state_t * tipmap, unsigned int maxstates,
                           int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int k, i, rows = rows_computed(states, sp);
  int * below = (int *)malloc(rate_cats * sizeof(int));
  (void)maxstates;
  for (n = 0; n < sites; ++n)
  {
    /* 4 states: tipchars hold the raw mask (pll.c:875-895); otherwise an
     * index into tipmap (pll.c:912-933) */
    orc_state_t mask = (states == 4) ? left_tipchars[n] : tipmap[left_tipchars[n]];
    double * pclv = parent_clv + n * span;
    if (parent_scaler)
    {
      if (per_rate)
        for (k = 0; k < rate_cats; ++k)
          parent_scaler[n * rate_cats + k] =
              right_scaler ? right_scaler[n * rate_cats + k] : 0;
      else
        parent_scaler[n] = right_scaler ? right_scaler[n] : 0;
    }
    for (k = 0; k < rate_cats; ++k)
    {
      const double * lm = left_matrix + (size_t)k * states * sp;
      const double * rm = right_matrix + (size_t)k * states * sp;
      int b = 1;
      for (i = 0; i < rows; ++i)
      {
        double a = tip_term(states, sp, lm + (size_t)i * sp, mask);
        double c = row_dot(states, sp, rm + (size_t)i * sp,
                           right_clv + n * span + k * sp);
        double v = a * c;
        pclv[k * sp + i] = v;
        b = b && (v < SCALE_THRESHOLD);
      }
      below[k] = b;
    }
    apply_scaling(pclv, sp, rate_cats, below, parent_scaler, n, per_rate);
  }
  free(below);
}

void orc_update_partial_tt(unsigned int states, unsigned int sp,
                           unsigned int sites, unsigned int rate_cats,
                           double * parent_clv, unsigned int * parent_scaler,
                           const unsigned char * left_tipchars,
                           const unsigned char * right_tipchars,
                           const double * left_matrix,
                           const double * right_matrix,
                           const orc_state_t * tipmap, unsigned int maxstates,
                           int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int k, i;
  (void)maxstates;
  /* never scales; the parent scaler is zeroed (core_partials_avx.c:1007-1010) */
  if (parent_scaler)
    memset(parent_scaler, 0,
           sizeof(unsigned int) * (per_rate ? (size_t)sites * rate_cats : sites));
  for (n = 0; n < sites; ++n)
  {
    orc_state_t lmask = (states == 4) ? left_tipchars[n] : tipmap[left_tipchars[n]];
    orc_state_t rmask = (states == 4) ? right_tipchars[n] : tipmap[right_tipchars[n]];
    double * pclv = parent_clv + n * span;
    for (k = 0; k < rate_cats; ++k)
    {
      const double * lm = left_matrix + (size_t)k * states * sp;
      const double * rm = right_matrix + (size_t)k * states * sp;
      for (i = 0; i < states; ++i)
      {
        double a, c;
        if (states == 4)
        {
          a = masked_sum4(lm + i * 4, (unsigned int)lmask);
          c = masked_sum4(rm + i * 4, (unsigned int)rmask);
        }
        else
        {
          /* 20 states and generic: scalar sums in increasing column order
           * (core_partials_avx.c:63-92,160-185) */
          a = masked_sum_seq(lm + (size_t)i * sp, lmask, states);
          c = masked_sum_seq(rm + (size_t)i * sp, rmask, states);
        }
        pclv[k * sp + i] = a * c;
      }
      for (i = states; i < sp; ++i) pclv[k * sp + i] = 0.0;
    }
  }
}

/* ---- log-likelihood ---------------------------------------------------- */

static void minlh_table(double * t)
{
  double f = 1.0;
  int i;
  for (i = 0; i < MAXDIFF; ++i)
  {
    f *= SCALE_THRESHOLD;
    t[i] = f;
  }
}

double orc_root_loglikelihood(unsigned int states, unsigned int sp,
                              unsigned int sites, unsigned int rate_cats,
                              const double * clv, const unsigned int * site_id,
                              const unsigned int * scaler,
                              double * const * frequencies,
                              const double * rate_weights,
                              const unsigned int * pattern_weights,
                              const double * invar_proportion,
                              const int * invar_indices,
                              const unsigned int * freqs_indices,
                              double * persite_lnl)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int j, k;
  double logl = 0;
  for (n = 0; n < sites; ++n)
  {
    size_t id = site_id ? site_id[n] : n;
    const double * c = clv + id * span;
    double term = 0;
    for (j = 0; j < rate_cats; ++j)
    {
      const double * freqs = frequencies[freqs_indices[j]];
      double term_r = 0, pinv;
      for (k = 0; k < states; ++k) term_r += c[j * sp + k] * freqs[k];
      pinv = invar_proportion ? invar_proportion[freqs_indices[j]] : 0;
      if (pinv > 0)
      {
        double inv_lk = (invar_indices[n] == -1) ? 0 : freqs[invar_indices[n]];
        term += rate_weights[j] * (term_r * (1 - pinv) + inv_lk * pinv);
      }
      else
        term += term_r * rate_weights[j];
    }
    term = log(term);
    /* per-site scalers only (core_likelihood.c:197-198) */
    if (scaler && scaler[id]) term += scaler[id] * log(SCALE_THRESHOLD);
    term *= pattern_weights[n];
    if (persite_lnl) persite_lnl[n] = term;
    logl += term;
  }
  return logl;
}

/* shared tail of the edge functions (core_likelihood.c:1388-1490) */
static double edge_site(unsigned int states, unsigned int sp,
                        unsigned int rate_cats, const double * cp,
                        const double * cc, orc_state_t tipmask, int is_tip,
                        const double * pmatrix, double * const * frequencies,
                        const double * rate_weights,
                        const double * invar_proportion, int invar_index,
                        const unsigned int * freqs_indices,
                        const unsigned int * pscal, const unsigned int * cscal,
                        size_t pid, size_t cid, int per_rate,
                        const double * minlh, unsigned int * rate_scalings)
{
  unsigned int i, j, k, site_scalings;
  double terma = 0, terminv = 0, site_lk;
  if (per_rate)
  {
    site_scalings = UINT_MAX;
    for (i = 0; i < rate_cats; ++i)
    {
      rate_scalings[i] = (pscal ? pscal[pid * rate_cats + i] : 0) +
                         (cscal ? cscal[cid * rate_cats + i] : 0);
      if (rate_scalings[i] < site_scalings) site_scalings = rate_scalings[i];
    }
    for (i = 0; i < rate_cats; ++i)
    {
      unsigned int d = rate_scalings[i] - site_scalings;
      rate_scalings[i] = d < MAXDIFF ? d : MAXDIFF;
    }
  }
  else
    site_scalings = (pscal ? pscal[pid] : 0) + (cscal ? cscal[cid] : 0);

  for (i = 0; i < rate_cats; ++i)
  {
    const double * freqs = frequencies[freqs_indices[i]];
    const double * pm = pmatrix + (size_t)i * states * sp;
    double terma_r = 0, pinv;
    for (j = 0; j < states; ++j)
    {
      double termb = 0;
      if (is_tip)
      {
        for (k = 0; k < states; ++k)
          if ((tipmask >> k) & 1) termb += pm[j * sp + k];
      }
      else
        for (k = 0; k < states; ++k) termb += pm[j * sp + k] * cc[i * sp + k];
      terma_r += cp[i * sp + j] * freqs[j] * termb;
    }
    if (per_rate && rate_scalings[i] > 0) terma_r *= minlh[rate_scalings[i] - 1];
    pinv = invar_proportion ? invar_proportion[freqs_indices[i]] : 0;
    if (pinv > 0)
    {
      terma += rate_weights[i] * terma_r * (1. - pinv);
      if (invar_index != -1)
        terminv += rate_weights[i] * freqs[invar_index] * pinv;
    }
    else
      terma += terma_r * rate_weights[i];
  }
  if (site_scalings)
  {
    if (terminv > 0.)
    {
      unsigned int capped = site_scalings < MAXDIFF ? site_scalings : MAXDIFF;
      site_lk = log(terma * minlh[capped - 1] + terminv);
    }
    else
      site_lk = log(terma) + site_scalings * log(SCALE_THRESHOLD);
  }
  else
    site_lk = log(terma + terminv);
  return site_lk;
}

double orc_edge_loglikelihood_ii(unsigned int states, unsigned int sp,
                                 unsigned int sites, unsigned int rate_cats,
                                 const double * clvp,
                                 const unsigned int * parent_scaler,
                                 const unsigned int * parent_site_id,
                                 const double * clvc,
                                 const unsigned int * child_scaler,
                                 const unsigned int * child_site_id,
                                 const double * pmatrix,
                                 double * const * frequencies,
                                 const double * rate_weights,
                                 const unsigned int * pattern_weights,
                                 const double * invar_proportion,
                                 const int * invar_indices,
                                 const unsigned int * freqs_indices,
                                 double * persite_lnl, int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  double logl = 0, minlh[MAXDIFF];
  unsigned int * rs = (unsigned int *)calloc(rate_cats, sizeof(unsigned int));
  minlh_table(minlh);
  for (n = 0; n < sites; ++n)
  {
    size_t pid = parent_site_id ? parent_site_id[n] : n;
    size_t cid = child_site_id ? child_site_id[n] : n;
    double lk = edge_site(states, sp, rate_cats, clvp + pid * span,
                          clvc + cid * span, 0, 0, pmatrix, frequencies,
                          rate_weights, invar_proportion,
                          invar_indices ? invar_indices[n] : -1, freqs_indices,
                          parent_scaler, child_scaler, pid, cid, per_rate,
                          minlh, rs);
    lk *= pattern_weights[n];
    if (persite_lnl) persite_lnl[n] = lk;
    logl += lk;
  }
  free(rs);
  return logl;
}

double orc_edge_loglikelihood_ti(unsigned int states, unsigned int sp,
                                 unsigned int sites, unsigned int rate_cats,
                                 const double * clvp,
                                 const unsigned int * parent_scaler,
                                 const unsigned char * tipchars,
                                 const orc_state_t * tipmap,
                                 const double * pmatrix,
                                 double * const * frequencies,
                                 const double * rate_weights,
                                 const unsigned int * pattern_weights,
                                 const double * invar_proportion,
                                 const int * invar_indices,
                                 const unsigned int * freqs_indices,
                                 double * persite_lnl, int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  double logl = 0, minlh[MAXDIFF];
  unsigned int * rs = (unsigned int *)calloc(rate_cats, sizeof(unsigned int));
  minlh_table(minlh);
  for (n = 0; n < sites; ++n)
  {
    orc_state_t mask = (states == 4) ? tipchars[n] : tipmap[tipchars[n]];
    double lk = edge_site(states, sp, rate_cats, clvp + n * span, NULL, mask, 1,
                          pmatrix, frequencies, rate_weights, invar_proportion,
                          invar_indices ? invar_indices[n] : -1, freqs_indices,
                          parent_scaler, NULL, n, n, per_rate, minlh, rs);
    lk *= pattern_weights[n];
    if (persite_lnl) persite_lnl[n] = lk;
    logl += lk;
  }
  free(rs);
  return logl;
}

/* ---- sumtable and derivatives ------------------------------------------ */

static void rate_scalings_for(unsigned int rate_cats, const unsigned int * ps,
                              const unsigned int * cs, size_t pid, size_t cid,
                              unsigned int * rs)
{
  unsigned int i, mn = UINT_MAX;
  for (i = 0; i < rate_cats; ++i)
  {
    rs[i] = (ps ? ps[pid * rate_cats + i] : 0) + (cs ? cs[cid * rate_cats + i] : 0);
    if (rs[i] < mn) mn = rs[i];
  }
  for (i = 0; i < rate_cats; ++i)
  {
    unsigned int d = rs[i] - mn;
    rs[i] = d < MAXDIFF ? d : MAXDIFF;
  }
}

void orc_update_sumtable_ii(unsigned int states, unsigned int sp,
                            unsigned int sites, unsigned int rate_cats,
                            const double * clvp, const unsigned int * parent_site_id,
                            const double * clvc, const unsigned int * child_site_id,
                            const unsigned int * parent_scaler,
                            const unsigned int * child_scaler,
                            double * const * eigenvecs,
                            double * const * inv_eigenvecs,
                            double * const * freqs, double * sumtable,
                            int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int i, j, k;
  double minlh[MAXDIFF];
  unsigned int * rs = (unsigned int *)calloc(rate_cats, sizeof(unsigned int));
  minlh_table(minlh);
  for (n = 0; n < sites; ++n)
  {
    size_t pid = parent_site_id ? parent_site_id[n] : n;
    size_t cid = child_site_id ? child_site_id[n] : n;
    if (per_rate)
      rate_scalings_for(rate_cats, parent_scaler, child_scaler, pid, cid, rs);
    for (i = 0; i < rate_cats; ++i)
    {
      const double * ev = eigenvecs[i], * iev = inv_eigenvecs[i], * f = freqs[i];
      const double * cp = clvp + pid * span + i * sp;
      const double * cc = clvc + cid * span + i * sp;
      double * sum = sumtable + n * span + i * sp;
      for (j = 0; j < states; ++j)
      {
        double l = 0, r = 0;
        for (k = 0; k < states; ++k)
        {
          l += cp[k] * f[k] * iev[k * sp + j];
          r += ev[j * sp + k] * cc[k];
        }
        sum[j] = l * r;
        if (per_rate && rs[i] > 0) sum[j] *= minlh[rs[i] - 1];
      }
      for (j = states; j < sp; ++j) sum[j] = 0;
    }
  }
  free(rs);
}

void orc_update_sumtable_ti(unsigned int states, unsigned int sp,
                            unsigned int sites, unsigned int rate_cats,
                            const double * clv_inner,
                            const unsigned char * tipchars,
                            const orc_state_t * tipmap,
                            const unsigned int * inner_scaler,
                            double * const * eigenvecs,
                            double * const * inv_eigenvecs,
                            double * const * freqs, double * sumtable,
                            int per_rate)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int i, j, k;
  double minlh[MAXDIFF];
  unsigned int * rs = (unsigned int *)calloc(rate_cats, sizeof(unsigned int));
  minlh_table(minlh);
  for (n = 0; n < sites; ++n)
  {
    orc_state_t mask = (states == 4) ? tipchars[n] : tipmap[tipchars[n]];
    if (per_rate) rate_scalings_for(rate_cats, inner_scaler, NULL, n, n, rs);
    for (i = 0; i < rate_cats; ++i)
    {
      const double * ev = eigenvecs[i], * iev = inv_eigenvecs[i], * f = freqs[i];
      const double * cc = clv_inner + n * span + i * sp;
      double * sum = sumtable + n * span + i * sp;
      for (j = 0; j < states; ++j)
      {
        /* tip on the "left": sum_k [bit k] pi_k Vinv_kj (core_derivatives.c:667-676) */
        double l = 0, r = 0;
        for (k = 0; k < states; ++k)
        {
          l += (double)((mask >> k) & 1) * f[k] * iev[k * sp + j];
          r += ev[j * sp + k] * cc[k];
        }
        sum[j] = l * r;
        if (per_rate && rs[i] > 0) sum[j] *= minlh[rs[i] - 1];
      }
      for (j = states; j < sp; ++j) sum[j] = 0;
    }
  }
  free(rs);
}

void orc_likelihood_derivatives(unsigned int states, unsigned int sp,
                                unsigned int sites, unsigned int rate_cats,
                                const double * rate_weights,
                                const int * invariant,
                                const unsigned int * pattern_weights,
                                double branch_length,
                                const double * prop_invar,
                                double * const * freqs, const double * rates,
                                double * const * eigenvals,
                                const double * sumtable, double * d_f,
                                double * dd_f)
{
  size_t n, span = (size_t)sp * rate_cats;
  unsigned int i, j;
  double * diag = (double *)malloc((size_t)rate_cats * states * 3 * sizeof(double));
  double df = 0, ddf = 0;
  for (i = 0; i < rate_cats; ++i)
  {
    double ki = rates[i] / (1.0 - prop_invar[i]);
    for (j = 0; j < states; ++j)
    {
      double lam = eigenvals[i][j];
      double e = exp(lam * ki * branch_length);
      double * d = diag + ((size_t)i * states + j) * 3;
      d[0] = e;
      d[1] = lam * ki * e;
      d[2] = lam * ki * lam * ki * e;
    }
  }
  for (n = 0; n < sites; ++n)
  {
    double lk[3] = {0, 0, 0};
    for (i = 0; i < rate_cats; ++i)
    {
      const double * sum = sumtable + n * span + i * sp;
      double c[3] = {0, 0, 0}, pinv = prop_invar[i];
      for (j = 0; j < states; ++j)
      {
        const double * d = diag + ((size_t)i * states + j) * 3;
        c[0] += sum[j] * d[0];
        c[1] += sum[j] * d[1];
        c[2] += sum[j] * d[2];
      }
      if (pinv > 0)
      {
        double inv_lk = (invariant[n] == -1) ? 0 : freqs[i][invariant[n]] * pinv;
        c[0] = c[0] * (1. - pinv) + inv_lk;
        c[1] = c[1] * (1. - pinv);
        c[2] = c[2] * (1. - pinv);
      }
      lk[0] += c[0] * rate_weights[i];
      lk[1] += c[1] * rate_weights[i];
      lk[2] += c[2] * rate_weights[i];
    }
    {
      double d1 = -lk[1] / lk[0];
      double d2 = d1 * d1 - lk[2] / lk[0];
      df += pattern_weights[n] * d1;
      ddf += pattern_weights[n] * d2;
    }
  }
  *d_f = df;
  *dd_f = ddf;
  free(diag);
}

/* ---- repeat identifiers ------------------------------------------------ */

unsigned int orc_update_repeats(unsigned int sites,
                                const unsigned int * site_id_left,
                                unsigned int ids_left,
                                const unsigned int * site_id_right,
                                unsigned int ids_right,
                                unsigned int * site_id_parent,
                                unsigned int * id_site_parent,
                                unsigned int * lookup,
                                unsigned int lookup_size)
{
  unsigned long long min_size = (unsigned long long)ids_left * ids_right;
  unsigned int s, cur = 0;
  unsigned int * toclean;
  /* enable rule, repeats.c:100-110 */
  if (!min_size || (unsigned long long)lookup_size <= min_size ||
      ids_left > sites / 2 || ids_right > sites / 2)
    return 0;
  toclean = (unsigned int *)malloc(sites * sizeof(unsigned int));
  for (s = 0; s < sites; ++s)
  {
    unsigned int key = site_id_left[s] + site_id_right[s] * ids_left;
    unsigned int id = lookup[key];
    if (id == EMPTY)
    {
      toclean[cur] = key;
      id_site_parent[cur] = s;
      id = cur;
      lookup[key] = cur++;
    }
    site_id_parent[s] = id;
  }
  for (s = 0; s < cur; ++s) lookup[toclean[s]] = EMPTY;
  free(toclean);
  /* "no compression" rule, repeats.c:366-370 */
  if (cur >= sites) return 0;
  return cur;
}

unsigned int orc_update_repeats_tip(unsigned int sites, const orc_state_t * map,
                                    const char * sequence,
                                    unsigned int * site_id,
                                    unsigned int * id_site)
{
  /* repeats.c:28-45: dense class of a character = first character (in ASCII
   * order) that shares its state mask */
  unsigned int cls[256], seen[257];
  unsigned int i, j, maxc = 0, s, cur = 0;
  memset(cls, 0, sizeof(cls));
  for (i = 0; i < 256; ++i)
  {
    for (j = 0; j < i; ++j)
      if (map[i] == map[j])
      {
        cls[i] = cls[j];
        break;
      }
    if (!cls[i]) cls[i] = ++maxc;
  }
  for (i = 0; i < 257; ++i) seen[i] = EMPTY;
  for (s = 0; s < sites; ++s)
  {
    unsigned int c = cls[(unsigned char)sequence[s]];
    if (seen[c] == EMPTY)
    {
      id_site[cur] = s;
      seen[c] = cur++;
    }
    site_id[s] = seen[c];
  }
  return cur;
}
