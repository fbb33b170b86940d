/*
 * plf_edge_aa.cu -- 20-state (protein) edge / root log-likelihood for
 * contiguous (non-repeats) CLVs and per-site scalers.
 *
 * Replaces, for 20 states, pll_core_edge_loglikelihood_ii / _ti and
 * pll_core_root_loglikelihood (reference src/core_likelihood.c:1192,581,25;
 * AVX2: src/core_likelihood_avx2.c).  Site-repeat gathers, per-rate scalers
 * and rate counts other than 1, 2, 4, 8 stay on the generic kernel of
 * plf_likelihood.cu.
 *
 * Same machinery as the streaming 20-state CLV kernel
 * (plf_partials_aa_mma.cu): tiles of 32/R sites (all rates, contiguous) arrive
 * in a shared-memory ring through bulk async copies; every warp owns one rate
 * category and one 8-site block; the child side sum_j P_ij c_j runs on the
 * FP64 tensor path (15 DMMA per 8 sites and rate, P^T fragments in registers)
 * and lands in the D-fragment layout (lane = site, two consecutive states),
 * where it meets pi_i p_i read from the parent tile.  A lane-group shuffle
 * gives the per-(site, rate) term; the terms of a batch of 32 sites are parked
 * in shared memory and finished (rate mixture, +I, scalers, log, pattern
 * weight) by one warp with every lane busy.  Per-block partial sums are
 * combined by the last block in a fixed order.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"
#include "plf_mma.cuh"
#include "plf_stream.cuh"

enum { LKA_II = 0, LKA_TI = 1, LKA_ROOT = 2 };
#define LKA_NSTAGE 4

template <int MODE, int LOG2R, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 512 / (NWARPS * 32))
k_lk_aa_mma(plf_lk_t a, int maxstates, double * __restrict__ partial, unsigned int * ticket, double * out,
            double * hout)
{
  constexpr int R = 1 << LOG2R;
  constexpr int SB = NWARPS / R;            /* 8-site blocks per tile */
  constexpr int TILE = 8 * SB;              /* sites per tile */
  constexpr int CH_BYTES = TILE * R * 160;  /* one CLV tile */
  constexpr int NCH = (MODE == LKA_II) ? 2 : 1;
  constexpr int STAGE = NCH * CH_BYTES;
  constexpr int BATCH = 32 / TILE;          /* tiles per finishing pass (32 sites) */
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) unsigned long long full[LKA_NSTAGE];
  __shared__ double terms[2][32][R];        /* [batch parity][site of the batch][rate] */
  __shared__ double red[32];
  unsigned char * ring = dyn;
  double * tl = reinterpret_cast<double *>(dyn + (size_t)LKA_NSTAGE * STAGE); /* TI: [maxstates][R][AAM_TAB_STRIDE] */

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 3, gs = lane >> 2;
  const int rate = warp & (R - 1), sb = warp >> LOG2R;
  const unsigned int ntiles = (a.sites + TILE - 1) / TILE;
  const size_t span = (size_t)R * 20;
  const double * m_weights = a.model + R, * m_pinv = a.model + 2 * R, * m_freqs = a.model + 3 * R;

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < LKA_NSTAGE; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == LKA_TI)
  {
    /* termb[code][rate][i] = sum of the columns of P row i selected by the state mask,
     * increasing column order (core_likelihood.c:581-700) */
    for (int e = threadIdx.x; e < maxstates * R * AAM_TAB_STRIDE; e += blockDim.x)
    {
      const int c = e / (R * AAM_TAB_STRIDE), r = (e / AAM_TAB_STRIDE) % R, i = e % AAM_TAB_STRIDE;
      tl[e] = (i < 20) ? masked_sum_seq(a.pmatrix + r * 400 + i * 20, a.tipmap[c], 20) : 0.0;
    }
  }
  __syncthreads();

  auto issue = [&](unsigned int t, int s) {
    const unsigned int first = t * TILE;
    const unsigned int n = min((unsigned int)TILE, a.sites - first);
    const unsigned int bytes = n * R * 160;
    unsigned char * slot = ring + (size_t)s * STAGE;
    mbar_expect_tx(&full[s], NCH * bytes);
    bulk_g2s(slot, a.clvp + (size_t)first * span, bytes, &full[s]);
    if (MODE == LKA_II) bulk_g2s(slot + CH_BYTES, a.clvc + (size_t)first * span, bytes, &full[s]);
  };
  if (threadIdx.x == 0)
  {
    unsigned int t = blockIdx.x;
    for (int s = 0; s < LKA_NSTAGE && t < ntiles; ++s, t += gridDim.x) issue(t, s);
  }

  /* per-warp constants: P^T fragments and pi of this rate in D-fragment layout */
  double br[MODE == LKA_II ? AAM_FRAGS : 1];
  if (MODE == LKA_II)
  {
#pragma unroll
    for (int f = 0; f < AAM_FRAGS; ++f)
    {
      const int nt = f / 5, kt = f % 5;
      const int i = 8 * nt + gs, j = aam_state(kt, q);
      br[f] = (i < 20) ? a.pmatrix[rate * 400 + i * 20 + j] : 0.0;
    }
  }
  double fq[3][2];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
  {
    const int i = 8 * nt + 2 * q;
    fq[nt][0] = (i < 20) ? m_freqs[rate * 20 + i] : 0.0;
    fq[nt][1] = (i + 1 < 20) ? m_freqs[rate * 20 + i + 1] : 0.0;
  }

  const unsigned int my = sb * 8 + gs; /* site within the tile */
  unsigned int code_next = 0;
  if (MODE == LKA_TI && blockIdx.x < ntiles)
  {
    const unsigned int n0 = blockIdx.x * TILE + my;
    code_next = a.tipchars[n0 < a.sites ? n0 : a.sites - 1];
  }

  /* finishing pass over the 32 sites of batch `b` (tiles b*BATCH .. of this CTA): one lane per site */
  double acc = 0;
  auto finish = [&](unsigned int b, unsigned int tiles_in_batch) {
    const unsigned int k = lane / TILE, i = lane % TILE; /* tile of the batch, site of the tile */
    if (k >= tiles_in_batch) return;
    const unsigned int t = blockIdx.x + (b * BATCH + k) * gridDim.x;
    const unsigned int n = t * TILE + i;
    if (n >= a.sites) return;
    const int inv = a.invariant ? a.invariant[n] : -1;
    double terma = 0, terminv = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
      const double term_r = terms[b & 1][lane][r];
      const double pinv = m_pinv[r], w = m_weights[r];
      if (pinv > 0)
      {
        const double inv_lk = (inv == -1) ? 0.0 : m_freqs[r * 20 + inv];
        if (MODE == LKA_ROOT)
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
    unsigned int sc = a.pscaler ? a.pscaler[n] : 0u;
    if (MODE == LKA_II && a.cscaler) sc += a.cscaler[n];
    double site_lk;
    if (MODE == LKA_ROOT)
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
    site_lk *= (double)a.pattern_weights[n];
    if (a.persite) a.persite[n] = site_lk;
    acc += site_lk;
  };

  unsigned int it = 0;
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it)
  {
    const int s = it % LKA_NSTAGE;
    const unsigned int parity = (it / LKA_NSTAGE) & 1u;
    const unsigned char * slot = ring + (size_t)s * STAGE;
    const unsigned int code = code_next;
    if (MODE == LKA_TI)
    {
      const unsigned int nn = (t + gridDim.x) * TILE + my;
      if (t + gridDim.x < ntiles) code_next = a.tipchars[nn < a.sites ? nn : a.sites - 1];
    }
    while (!mbar_try_wait(&full[s], parity)) {}

    double tb[3][2]; /* child side in D-fragment layout: states 8 nt + 2 q, + 1 of site `my` */
    if (MODE == LKA_II)
    {
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) tb[nt][0] = tb[nt][1] = 0.0;
      const double * pc = reinterpret_cast<const double *>(slot + CH_BYTES) + ((size_t)my * R + rate) * 20;
      const double2 v0 = *reinterpret_cast<const double2 *>(pc + 2 * q);
      const double2 v1 = *reinterpret_cast<const double2 *>(pc + 8 + 2 * q);
      const double af[5] = {v0.x, v0.y, v1.x, v1.y, pc[16 + q]};
#pragma unroll
      for (int kt = 0; kt < 5; ++kt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) dmma(tb[nt], af[kt], br[nt * 5 + kt]);
    }
    else if (MODE == LKA_TI)
    {
      const double * row = tl + ((size_t)code * R + rate) * AAM_TAB_STRIDE + 2 * q;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        tb[nt][0] = tb[nt][1] = 0.0;
        if (nt < 2 || q < 2)
        {
          const double2 tv = *reinterpret_cast<const double2 *>(row + 8 * nt);
          tb[nt][0] = tv.x;
          tb[nt][1] = tv.y;
        }
      }
    }
    else
    {
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) tb[nt][0] = tb[nt][1] = 1.0;
    }
    const double * pp = reinterpret_cast<const double *>(slot) + ((size_t)my * R + rate) * 20 + 2 * q;
    double term = 0;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
      if (nt < 2 || q < 2)
      {
        const double2 pv = *reinterpret_cast<const double2 *>(pp + 8 * nt);
        term = fma(pv.x * fq[nt][0], tb[nt][0], term);
        term = fma(pv.y * fq[nt][1], tb[nt][1], term);
      }
    term += __shfl_xor_sync(0xffffffffu, term, 1);
    term += __shfl_xor_sync(0xffffffffu, term, 2);
    const unsigned int b = it / BATCH, kb = it % BATCH;
    if (q == 0) terms[b & 1][kb * TILE + my][rate] = term;
    __syncthreads(); /* ring slot free, this tile's terms visible */
    {
      const unsigned int tn = t + (unsigned int)LKA_NSTAGE * gridDim.x;
      if (threadIdx.x == 0 && tn < ntiles) issue(tn, s);
    }
    if (kb == BATCH - 1 && warp == (int)(b % NWARPS)) finish(b, BATCH);
  }
  /* the last, incomplete batch of this CTA */
  if (it % BATCH)
  {
    const unsigned int b = it / BATCH;
    if (warp == (int)(b % NWARPS)) finish(b, it % BATCH);
  }
  const double v[1] = {acc};
  grid_reduce_finish<1>(v, partial, ticket, out, hout, red);
}

/* ------------------------------------------------------------------------ */

typedef void (*lka_kernel_t)(plf_lk_t, int, double *, unsigned int *, double *, double *);

template <int LOG2R, int NWARPS>
static lka_kernel_t lka_pick(int mode)
{
  return mode == LKA_II   ? k_lk_aa_mma<LKA_II, LOG2R, NWARPS>
         : mode == LKA_TI ? k_lk_aa_mma<LKA_TI, LOG2R, NWARPS>
                          : k_lk_aa_mma<LKA_ROOT, LOG2R, NWARPS>;
}

/* returns -1 when the call is not eligible (the generic kernel takes it) */
int plf_loglikelihood_aa(plf_ctx * ctx, const plf_shape_t * sh, const plf_lk_t * a, unsigned int maxstates,
                         double * dst, double * hdst)
{
  const unsigned int R = sh->rate_cats;
  if (sh->states != 20 || sh->per_rate_scalers || a->p_site_id || a->c_site_id || !a->sites ||
      !(R == 1 || R == 2 || R == 4 || R == 8))
    return -1;
  const int mode = !a->pmatrix ? LKA_ROOT : a->tipchars ? LKA_TI : LKA_II;
  const int log2r = R == 1 ? 0 : R == 2 ? 1 : R == 4 ? 2 : 3;
  const int nwarps = (log2r == 3) ? 8 : 4;
  lka_kernel_t k = log2r == 0 ? lka_pick<0, 4>(mode) : log2r == 1 ? lka_pick<1, 4>(mode)
                   : log2r == 2 ? lka_pick<2, 4>(mode) : lka_pick<3, 8>(mode);
  size_t smem = (size_t)LKA_NSTAGE * (mode == LKA_II ? 2 : 1) * 1280 * nwarps;
  if (mode == LKA_TI) smem += (size_t)maxstates * R * AAM_TAB_STRIDE * sizeof(double);
  if (smem > ctx->smem_optin) return -1;
  int & occ = ctx->lka_occupancy[mode][log2r];
  if (!occ || smem > ctx->lka_smem_set[mode][log2r])
  {
    PLF_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->lka_smem_set[mode][log2r] = smem;
    PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, nwarps * 32, smem));
    if (occ < 1) occ = 1;
  }
  const unsigned int tile = 8 * nwarps / R;
  const unsigned long long ntiles = ((unsigned long long)a->sites + tile - 1) / tile;
  unsigned long long blocks = (unsigned long long)ctx->sm_count * occ;
  if (blocks > ntiles) blocks = ntiles;
  double * partial = (double *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)blocks * 2 * sizeof(double));
  if (!partial) return 0;
  k<<<(unsigned int)blocks, nwarps * 32, smem, ctx->stream>>>(*a, (int)maxstates, partial, ctx->d_ticket, dst, hdst);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}
