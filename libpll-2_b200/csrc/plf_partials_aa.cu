/*
 * plf_partials_aa.cu -- 20-state (protein) CLV updates, inner-inner and tip-inner,
 * on the FP64 vector pipe (DFMA) in the reference's summation order: CLVs are
 * bit-identical to the AVX2 reference.  Selected with PLF_AA_MMA=0; the default
 * path is the tensor-core one (plf_partials_aa_mma.cu), which ncu showed to be
 * 1.7x faster at the cost of last-bit differences in the CLVs.
 *
 * Replaces pll_core_update_partial_ii_20x20_avx2 / _ti_20x20_avx2 (reference
 * src/core_partials_avx2.c:630,343) and the scaler pass src/pll.c:1202.
 *
 * Unlike DNA this path sits at the FP64 ridge: per (site, rate) block 2 x 20 rows
 * of (20 FMA + 3 ADD) + 20 MUL = 940 FP64 instructions against 480 bytes, i.e.
 * ~63 % of the FP64 pipe is needed to keep HBM busy.  The kernel is therefore
 * built around FP64 issue:
 *   - lanes of a warp are SITES and the whole warp works on ONE rate category at
 *     a time, so every P-matrix element is a warp-uniform shared-memory address:
 *     one broadcast LDS.128 feeds two FMAs of every lane, no bank conflicts;
 *   - each thread owns SITES_PER_THREAD sites, so a matrix element fetched once is
 *     used for 2 x SITES_PER_THREAD FMAs and 4 x SITES_PER_THREAD independent FMA
 *     chains hide the FP64 latency;
 *   - the left products A[0..19] stay in registers while the right child streams in;
 *   - per-site scaling is decided after all rates (the values were stored
 *     unscaled; the rare site that scales is rescaled in place by its own thread).
 *
 * Summation order is the reference's: four lane accumulators over the column
 * quads, fused multiply-adds, then (a0+a1)+(a2+a3):
 * CLVs and scalers are bit-identical.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"

#define AA_THREADS 128
#define AA_SPT 2          /* sites per thread */
#define AA_TAB_STRIDE 22  /* doubles per tip-table row: 16-byte aligned, spreads codes over banks */

/* row . vec for AA_SPT sites at once; m = 20 matrix entries in shared memory */
__device__ __forceinline__ void rows4_fma(const double * __restrict__ m, const double (&c)[AA_SPT][20],
                                          double (&out)[AA_SPT][4])
{
#pragma unroll
  for (int q = 0; q < 4; ++q)
  {
    const double * row = m + q * 20;
    double a[AA_SPT][4];
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s) a[s][0] = a[s][1] = a[s][2] = a[s][3] = 0.0;
#pragma unroll
    for (int j = 0; j < 20; j += 4)
    {
      const double2 m01 = *reinterpret_cast<const double2 *>(row + j);
      const double2 m23 = *reinterpret_cast<const double2 *>(row + j + 2);
#pragma unroll
      for (int s = 0; s < AA_SPT; ++s)
      {
        a[s][0] = fma(m01.x, c[s][j + 0], a[s][0]);
        a[s][1] = fma(m01.y, c[s][j + 1], a[s][1]);
        a[s][2] = fma(m23.x, c[s][j + 2], a[s][2]);
        a[s][3] = fma(m23.y, c[s][j + 3], a[s][3]);
      }
    }
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s) out[s][q] = (a[s][0] + a[s][1]) + (a[s][2] + a[s][3]);
  }
}

__device__ __forceinline__ void load20(double (&v)[20], const double * __restrict__ p, bool streaming)
{
#pragma unroll
  for (int j = 0; j < 20; j += 4)
  {
    const dbl4 t = streaming ? ld256_stream(p + j) : ld256(p + j);
    v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
  }
}

template <int KIND>
__global__ void __launch_bounds__(AA_THREADS)
k_clv_aa(const plf_op_t * __restrict__ ops, int R, int per_rate, const plf_state_t * __restrict__ tipmap,
         int maxstates)
{
  extern __shared__ __align__(16) double smem[];
  const plf_op_t op = ops[blockIdx.y];
  double * lmat = smem;                                   /* [R][400], II only */
  double * rmat = smem + (KIND == PLF_OP_II ? R * 400 : 0); /* [R][400] */
  double * tl = rmat + R * 400;                           /* TI: [maxstates][R][AA_TAB_STRIDE] */
  for (int e = threadIdx.x; e < R * 400; e += blockDim.x)
  {
    if (KIND == PLF_OP_II) lmat[e] = op.left_matrix[e];
    rmat[e] = op.right_matrix[e];
  }
  if (KIND == PLF_OP_TI)
  {
    /* scalar sums in increasing column order (src/core_partials_avx2.c:387-456) */
    for (int e = threadIdx.x; e < maxstates * R * 20; e += blockDim.x)
    {
      const int c = e / (R * 20), r = (e / 20) % R, i = e % 20;
      tl[(c * R + r) * AA_TAB_STRIDE + i] = masked_sum_seq(op.left_matrix + r * 400 + i * 20, tipmap[c], 20);
    }
  }
  __syncthreads();

  const unsigned int nsites = op.nsites;
  const size_t span = (size_t)R * 20;
  const unsigned int chunk = gridDim.x * AA_THREADS; /* sites one sweep covers per slot */
  const unsigned int t0 = blockIdx.x * AA_THREADS + threadIdx.x;
  const bool gather = op.parent_id_site || op.left_site_id || op.right_site_id;

  for (unsigned int base = 0; base < nsites; base += chunk * AA_SPT)
  {
    unsigned int n[AA_SPT], lid[AA_SPT], rid[AA_SPT], code[AA_SPT];
    bool act[AA_SPT];
    int below_all[AA_SPT];
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      n[s] = base + s * chunk + t0;
      act[s] = n[s] < nsites;
      lid[s] = rid[s] = act[s] ? n[s] : 0;
      code[s] = 0;
      below_all[s] = 1;
      if (act[s] && gather)
      {
        const unsigned int site = op.parent_id_site ? op.parent_id_site[n[s]] : n[s];
        lid[s] = op.left_site_id ? op.left_site_id[site] : site;
        rid[s] = op.right_site_id ? op.right_site_id[site] : site;
      }
      if (KIND == PLF_OP_TI && act[s]) code[s] = op.left_tip[lid[s]];
    }

    for (int rate = 0; rate < R; ++rate)
    {
      double c[AA_SPT][20];
      double A[AA_SPT][20];
      if (KIND == PLF_OP_II)
      {
#pragma unroll
        for (int s = 0; s < AA_SPT; ++s) load20(c[s], op.left_clv + (size_t)lid[s] * span + rate * 20, true);
#pragma unroll
        for (int i = 0; i < 20; i += 4)
        {
          double o[AA_SPT][4];
          rows4_fma(lmat + rate * 400 + i * 20, c, o);
#pragma unroll
          for (int s = 0; s < AA_SPT; ++s)
          {
            A[s][i] = o[s][0]; A[s][i + 1] = o[s][1]; A[s][i + 2] = o[s][2]; A[s][i + 3] = o[s][3];
          }
        }
      }
      else
      {
#pragma unroll
        for (int s = 0; s < AA_SPT; ++s)
        {
          const double * row = tl + ((size_t)code[s] * R + rate) * AA_TAB_STRIDE;
#pragma unroll
          for (int i = 0; i < 20; i += 2)
          {
            const double2 t = *reinterpret_cast<const double2 *>(row + i);
            A[s][i] = t.x;
            A[s][i + 1] = t.y;
          }
        }
      }
#pragma unroll
      for (int s = 0; s < AA_SPT; ++s) load20(c[s], op.right_clv + (size_t)rid[s] * span + rate * 20, true);
      int below[AA_SPT];
#pragma unroll
      for (int s = 0; s < AA_SPT; ++s) below[s] = 1;
#pragma unroll
      for (int i = 0; i < 20; i += 4)
      {
        double o[AA_SPT][4];
        rows4_fma(rmat + rate * 400 + i * 20, c, o);
#pragma unroll
        for (int s = 0; s < AA_SPT; ++s)
        {
          dbl4 v;
          v.x = A[s][i] * o[s][0];
          v.y = A[s][i + 1] * o[s][1];
          v.z = A[s][i + 2] * o[s][2];
          v.w = A[s][i + 3] * o[s][3];
          below[s] &= (v.x < PLF_SCALE_THRESHOLD) && (v.y < PLF_SCALE_THRESHOLD) && (v.z < PLF_SCALE_THRESHOLD) &&
                      (v.w < PLF_SCALE_THRESHOLD);
          if (act[s]) st256(op.parent_clv + (size_t)n[s] * span + rate * 20 + i, v);
        }
      }
#pragma unroll
      for (int s = 0; s < AA_SPT; ++s)
      {
        below_all[s] &= below[s];
        if (op.parent_scaler && per_rate && act[s])
        {
          unsigned int sc = 0;
          if (KIND == PLF_OP_II && op.left_scaler) sc += op.left_scaler[(size_t)lid[s] * R + rate];
          if (op.right_scaler) sc += op.right_scaler[(size_t)rid[s] * R + rate];
          if (below[s])
          {
            double * p = op.parent_clv + (size_t)n[s] * span + rate * 20;
#pragma unroll
            for (int i = 0; i < 20; i += 4)
            {
              dbl4 v = ld256(p + i);
              v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR; v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
              st256(p + i, v);
            }
            sc += 1;
          }
          op.parent_scaler[(size_t)n[s] * R + rate] = sc;
        }
      }
    }

    if (op.parent_scaler && !per_rate)
    {
#pragma unroll
      for (int s = 0; s < AA_SPT; ++s)
      {
        if (!act[s]) continue;
        unsigned int sc = 0;
        if (KIND == PLF_OP_II && op.left_scaler) sc += op.left_scaler[lid[s]];
        if (op.right_scaler) sc += op.right_scaler[rid[s]];
        if (below_all[s])
        {
          double * p = op.parent_clv + (size_t)n[s] * span;
          for (int i = 0; i < R * 20; i += 4)
          {
            dbl4 v = ld256(p + i);
            v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR; v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
            st256(p + i, v);
          }
          sc += 1;
        }
        op.parent_scaler[n[s]] = sc;
      }
    }
  }
}

/* one run of same-kind protein ops (ii or ti) as a single persistent wave */
int plf_launch_aa_group(plf_ctx * ctx, const plf_op_t * d_ops, unsigned int nops, unsigned int kind,
                        unsigned int rate_cats, int per_rate, unsigned int max_sites,
                        const unsigned long long * d_tipmap, unsigned int maxstates)
{
  const int ii = (kind == PLF_OP_II);
  size_t smem = (size_t)(ii ? 2 : 1) * rate_cats * 400 * sizeof(double);
  if (!ii) smem += (size_t)maxstates * rate_cats * AA_TAB_STRIDE * sizeof(double);
  if (smem > ctx->smem_optin) return -1; /* caller falls back to the generic kernel */
  void (*k)(const plf_op_t *, int, int, const plf_state_t *, int) = ii ? k_clv_aa<PLF_OP_II> : k_clv_aa<PLF_OP_TI>;
  size_t & set = ctx->aa_smem_set[ii ? 0 : 1];
  if (smem > set)
  {
    PLF_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = smem;
    ctx->aa_occupancy[ii ? 0 : 1] = 0;
  }
  int & occ = ctx->aa_occupancy[ii ? 0 : 1];
  if (!occ)
  {
    PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, AA_THREADS, smem));
    if (occ < 1) occ = 1;
  }
  unsigned long long need = ((unsigned long long)max_sites + AA_THREADS * AA_SPT - 1) / (AA_THREADS * AA_SPT);
  unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
  if (bx < 1) bx = 1;
  if (bx > need) bx = need;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned int)bx, nops);
  k<<<grid, AA_THREADS, smem, ctx->stream>>>(d_ops, (int)rate_cats, per_rate, d_tipmap, (int)maxstates);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}
