/*
 * pll_eigen.c -- host-side eigendecomposition of a reversible rate matrix.
 *
 * Stays on the host as in the reference (pll_update_eigen, src/models.c:293-410):
 * it runs once per model change on a states x states matrix.  The method is the
 * classic one the reference uses -- symmetrise Q with sqrt(pi), Householder
 * reduction to tridiagonal form, implicit-shift QL iteration -- written here
 * 0-based from the textbook algorithm.  The order of floating-point operations
 * is kept the same as the reference's routines (models.c:24-178) on purpose:
 * identical eigenvectors mean identical P-matrices, which is what makes CLVs and
 * scalers comparable bit for bit.  Build without FMA contraction
 * (-ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pll_b200.h"

/* Householder tridiagonalisation of the symmetric n x n matrix a (row pointers).
 * On exit d = diagonal, e = sub-diagonal (e[0] = 0), a = accumulated transform. */
static void householder_tridiag(double ** a, int n, double * d, double * e)
{
  int i, j, k, l;
  for (i = n - 1; i >= 1; --i)
  {
    double h = 0.0, scale = 0.0;
    l = i - 1;
    if (l > 0)
    {
      for (k = 0; k <= l; ++k) scale += fabs(a[k][i]);
      if (scale == 0.0)
        e[i] = a[l][i];
      else
      {
        double f, g, hh;
        for (k = 0; k <= l; ++k)
        {
          a[k][i] /= scale;
          h += a[k][i] * a[k][i];
        }
        f = a[l][i];
        g = (f > 0) ? -sqrt(h) : sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        a[l][i] = f - g;
        f = 0.0;
        for (j = 0; j <= l; ++j)
        {
          a[i][j] = a[j][i] / h;
          g = 0.0;
          for (k = 0; k <= j; ++k) g += a[k][j] * a[k][i];
          for (k = j + 1; k <= l; ++k) g += a[j][k] * a[k][i];
          e[j] = g / h;
          f += e[j] * a[j][i];
        }
        hh = f / (h + h);
        for (j = 0; j <= l; ++j)
        {
          f = a[j][i];
          g = e[j] - hh * f;
          e[j] = g;
          for (k = 0; k <= j; ++k) a[k][j] -= (f * e[k] + g * a[k][i]);
        }
      }
    }
    else
      e[i] = a[l][i];
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (i = 0; i < n; ++i)
  {
    l = i - 1;
    if (d[i] != 0.0)
    {
      for (j = 0; j <= l; ++j)
      {
        double g = 0.0;
        for (k = 0; k <= l; ++k) g += a[k][i] * a[j][k];
        for (k = 0; k <= l; ++k) a[j][k] -= g * a[i][k];
      }
    }
    d[i] = a[i][i];
    a[i][i] = 1.0;
    for (j = 0; j <= l; ++j) a[i][j] = a[j][i] = 0.0;
  }
}

/* Implicit-shift QL on the tridiagonal (d, e); rows of z are rotated along. */
static int ql_implicit(double * d, double * e, int n, double ** z)
{
  int m, l, iter, i, k;
  for (i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (l = 0; l < n; ++l)
  {
    iter = 0;
    do
    {
      for (m = l; m < n - 1; ++m)
      {
        double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) + dd == dd) break;
      }
      if (m != l)
      {
        double g, r, s, c, p, f, b;
        if (++iter > 60) return 0;
        g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        r = sqrt((g * g) + 1.0);
        g = d[m] - d[l] + e[l] / (g + ((g < 0) ? -fabs(r) : fabs(r)));
        s = c = 1.0;
        p = 0.0;
        for (i = m - 1; i >= l; --i)
        {
          f = s * e[i];
          b = c * e[i];
          if (fabs(f) >= fabs(g))
          {
            c = g / f;
            r = sqrt((c * c) + 1.0);
            e[i + 1] = f * r;
            c *= (s = 1.0 / r);
          }
          else
          {
            s = f / g;
            r = sqrt((s * s) + 1.0);
            e[i + 1] = g * r;
            s *= (c = 1.0 / r);
          }
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (k = 0; k < n; ++k)
          {
            f = z[i + 1][k];
            z[i + 1][k] = s * z[i][k] + c * f;
            z[i][k] = c * z[i][k] - s * f;
          }
        }
        d[l] = d[l] - p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return 1;
}

/* Symmetrised, mean-rate-1 rate matrix sqrt(pi) Q sqrt(pi)^-1 (models.c:182-256) */
static double ** symmetric_ratematrix(const double * params, const double * freqs, unsigned int states)
{
  unsigned int i, j, k = 0;
  unsigned int np = states * (states - 1) / 2;
  double * pn = (double *)malloc(np * sizeof(double));
  double ** q = (double **)malloc(states * sizeof(double *));
  double mean = 0;
  if (!pn || !q)
  {
    free(pn);
    free(q);
    return NULL;
  }
  for (i = 0; i < states; ++i) q[i] = (double *)calloc(states, sizeof(double));
  memcpy(pn, params, np * sizeof(double));
  if (pn[np - 1] > 0.0)
    for (i = 0; i < np; ++i) pn[i] /= pn[np - 1];
  for (i = 0; i < states; ++i)
    for (j = i + 1; j < states; ++j)
    {
      double factor = (freqs[i] <= PLL_EIGEN_MINFREQ || freqs[j] <= PLL_EIGEN_MINFREQ) ? 0 : pn[k];
      k++;
      q[i][j] = q[j][i] = factor * sqrt(freqs[i] * freqs[j]);
      q[i][i] -= factor * freqs[j];
      q[j][j] -= factor * freqs[i];
    }
  for (i = 0; i < states; ++i) mean += freqs[i] * (-q[i][i]);
  for (i = 0; i < states; ++i)
    for (j = 0; j < states; ++j) q[i][j] /= mean;
  free(pn);
  return q;
}

int pll_cuda_host_eigen(unsigned int states, unsigned int sp, const double * subst_params, const double * freqs,
                        double * eigenvecs, double * inv_eigenvecs, double * eigenvals)
{
  unsigned int i, j, inew, jnew, ns = 0;
  int ok;
  double ** a = symmetric_ratematrix(subst_params, freqs, states);
  double * d = (double *)malloc(states * sizeof(double));
  double * e = (double *)malloc(states * sizeof(double));
  double * sf = (double *)malloc(states * sizeof(double));
  if (!a || !d || !e || !sf)
  {
    free(d);
    free(e);
    free(sf);
    return PLL_FAILURE;
  }
  /* drop states of (near) zero frequency (models.c:258-291) */
  for (i = 0; i < states; ++i)
    if (freqs[i] > PLL_EIGEN_MINFREQ) sf[ns++] = freqs[i];
  if (ns < states)
    for (i = 0, inew = 0; i < states; ++i)
      if (freqs[i] > PLL_EIGEN_MINFREQ)
      {
        for (j = 0, jnew = 0; j < states; ++j)
          if (freqs[j] > PLL_EIGEN_MINFREQ) a[inew][jnew++] = a[i][j];
        inew++;
      }
  householder_tridiag(a, (int)ns, d, e);
  ok = ql_implicit(d, e, (int)ns, a);
  for (i = 0, inew = 0; i < states; ++i) eigenvals[i] = (freqs[i] > PLL_EIGEN_MINFREQ) ? d[inew++] : 0;
  for (i = 0; i < ns; ++i) sf[i] = sqrt(sf[i]);
  if (ns < states)
  {
    memset(eigenvecs, 0, (size_t)sp * states * sizeof(double));
    memset(inv_eigenvecs, 0, (size_t)sp * states * sizeof(double));
    for (i = 0; i < states; ++i) eigenvecs[i * sp + i] = inv_eigenvecs[i * sp + i] = 1.;
    for (i = 0, inew = 0; i < states; ++i)
      if (freqs[i] > PLL_EIGEN_MINFREQ)
      {
        for (j = 0, jnew = 0; j < states; ++j)
          if (freqs[j] > PLL_EIGEN_MINFREQ)
          {
            eigenvecs[i * sp + j] = a[inew][jnew] * sf[jnew];
            inv_eigenvecs[i * sp + j] = a[jnew][inew] / sf[inew];
            jnew++;
          }
        inew++;
      }
  }
  else
    for (i = 0; i < states; ++i)
      for (j = 0; j < states; ++j)
      {
        eigenvecs[i * sp + j] = a[i][j] * sf[j];       /* V  = U sqrt(pi)      */
        inv_eigenvecs[i * sp + j] = a[j][i] / sf[i];   /* V^-1 = sqrt(pi)^-1 U^T */
      }
  free(d);
  free(e);
  free(sf);
  for (i = 0; i < states; ++i) free(a[i]);
  free(a);
  return ok ? PLL_SUCCESS : PLL_FAILURE;
}
