/*
 * plf_partials_aa.cu -- 20-state (protein) CLV updates, inner-inner and tip-inner.
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

/* ---- two sites per thread, operands staged by cp.async one phase ahead ---------- *
 * Same arithmetic as k_clv_aa, but the 20-double operand vectors travel           *
 * global -> shared memory with 16-byte cp.async (LDGSTS: no registers, no warp     *
 * stall) into a lane-contiguous layout [chunk][thread], one "phase" (= one child   *
 * of one rate category for the thread's two sites) ahead of the arithmetic.  A     *
 * thread only reads back what it copied itself, so cp.async.wait_group is the only *
 * synchronisation.  The microbenchmark profiles/tools/fp64_operand_bench.cu shows  *
 * this register tiling sustains ~78 % of the FP64 pipe once loads are hidden.      */
#define AA2_STAGES 2
#define AA2_CHUNKS (AA_SPT * 10) /* 16-byte chunks per thread and phase */
#define AA2_STAGE_BYTES (AA2_CHUNKS * AA_THREADS * 16)

__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gmem_src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(AA_THREADS)
k_clv_aa2(const plf_op_t * __restrict__ ops, int R, int per_rate, const plf_state_t * __restrict__ tipmap,
          int maxstates)
{
  extern __shared__ __align__(16) double smem[];
  const plf_op_t op = ops[blockIdx.y];
  unsigned char * stage0 = reinterpret_cast<unsigned char *>(smem); /* AA2_STAGES operand stages first (16-byte aligned) */
  double * lmat = smem + AA2_STAGES * AA2_STAGE_BYTES / 8;
  double * rmat = lmat + (KIND == PLF_OP_II ? R * 400 : 0);
  double * tl = rmat + R * 400;
  for (int e = threadIdx.x; e < R * 400; e += blockDim.x)
  {
    if (KIND == PLF_OP_II) lmat[e] = op.left_matrix[e];
    rmat[e] = op.right_matrix[e];
  }
  if (KIND == PLF_OP_TI)
    for (int e = threadIdx.x; e < maxstates * R * 20; e += blockDim.x)
    {
      const int c = e / (R * 20), r = (e / 20) % R, i = e % 20;
      tl[(c * R + r) * AA_TAB_STRIDE + i] = masked_sum_seq(op.left_matrix + r * 400 + i * 20, tipmap[c], 20);
    }
  __syncthreads();

  const unsigned int nsites = op.nsites;
  const size_t span = (size_t)R * 20;
  const unsigned int chunk = gridDim.x * AA_THREADS;
  const unsigned int t0 = blockIdx.x * AA_THREADS + threadIdx.x;
  const bool gather = op.parent_id_site || op.left_site_id || op.right_site_id;
  const unsigned int iters = (nsites + chunk * AA_SPT - 1) / (chunk * AA_SPT);
  const int phases = (KIND == PLF_OP_II) ? 2 * R : R; /* per iteration */

  unsigned int lid[AA_SPT], rid[AA_SPT], nlid[AA_SPT], nrid[AA_SPT];
  auto resolve = [&](unsigned int it, unsigned int (&l)[AA_SPT], unsigned int (&r)[AA_SPT]) {
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      const unsigned int nn = it * chunk * AA_SPT + s * chunk + t0;
      l[s] = r[s] = nn < nsites ? nn : 0;
      if (nn < nsites && gather)
      {
        const unsigned int site = op.parent_id_site ? op.parent_id_site[nn] : nn;
        l[s] = op.left_site_id ? op.left_site_id[site] : site;
        r[s] = op.right_site_id ? op.right_site_id[site] : site;
      }
    }
  };
  /* queue the copies of phase `ph` (of the iteration whose ids are l/r) into stage `st` */
  auto issue = [&](int ph, const unsigned int (&l)[AA_SPT], const unsigned int (&r)[AA_SPT], int st) {
    const int rate = (KIND == PLF_OP_II) ? (ph >> 1) : ph;
    const bool left = (KIND == PLF_OP_II) && !(ph & 1);
    unsigned char * dst = stage0 + (size_t)st * AA2_STAGE_BYTES + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      const double * src = (left ? op.left_clv + (size_t)l[s] * span : op.right_clv + (size_t)r[s] * span) + rate * 20;
#pragma unroll
      for (int k = 0; k < 10; ++k) cp_async16(dst + (size_t)(s * 10 + k) * AA_THREADS * 16, src + 2 * k);
    }
    cp_async_commit();
  };
  auto fetch = [&](double (&c)[AA_SPT][20], int st) {
    const unsigned char * src = stage0 + (size_t)st * AA2_STAGE_BYTES + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
#pragma unroll
      for (int k = 0; k < 10; ++k)
      {
        const double2 t = *reinterpret_cast<const double2 *>(src + (size_t)(s * 10 + k) * AA_THREADS * 16);
        c[s][2 * k] = t.x;
        c[s][2 * k + 1] = t.y;
      }
  };

  if (!iters) return;
  resolve(0, lid, rid);
  issue(0, lid, rid, 0);
  unsigned int g = 0; /* global phase counter: stage = g % AA2_STAGES */

  for (unsigned int it = 0; it < iters; ++it)
  {
    unsigned int n[AA_SPT], code[AA_SPT];
    bool act[AA_SPT];
    int below_all[AA_SPT];
    resolve(it + 1, nlid, nrid);
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      n[s] = it * chunk * AA_SPT + s * chunk + t0;
      act[s] = n[s] < nsites;
      below_all[s] = 1;
      code[s] = (KIND == PLF_OP_TI && act[s]) ? op.left_tip[lid[s]] : 0u;
    }
    for (int rate = 0; rate < R; ++rate)
    {
      double c[AA_SPT][20];
      double A[AA_SPT][20];
      if (KIND == PLF_OP_II)
      {
        /* left phase: next = right(rate) */
        issue(2 * rate + 1, lid, rid, (g + 1) % AA2_STAGES);
        cp_async_wait<1>();
        fetch(c, g % AA2_STAGES);
        ++g;
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
      /* right phase: next = first phase of the next rate, or of the next iteration */
      {
        const int ph = (KIND == PLF_OP_II) ? 2 * rate + 1 : rate;
        if (ph + 1 < phases)
        {
          issue(ph + 1, lid, rid, (g + 1) % AA2_STAGES);
          cp_async_wait<1>();
        }
        else if (it + 1 < iters)
        {
          issue(0, nlid, nrid, (g + 1) % AA2_STAGES);
          cp_async_wait<1>();
        }
        else
          cp_async_wait<0>();
        fetch(c, g % AA2_STAGES);
        ++g;
      }
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
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      lid[s] = nlid[s];
      rid[s] = nrid[s];
    }
  }
}

/* ---- two sites per thread, register prefetch, left products parked in smem ------- *
 * k_clv_aa keeps A[2][20] in registers and therefore has no room to request the      *
 * next operand early.  Here A goes to a private, lane-contiguous slice of shared     *
 * memory (20 STS.128 + 20 LDS.128 per phase pair) and the freed registers hold the   *
 * operand of the NEXT phase, requested before the arithmetic of the current one.     */
template <int KIND>
__global__ void __launch_bounds__(AA_THREADS)
k_clv_aa3(const plf_op_t * __restrict__ ops, int R, int per_rate, const plf_state_t * __restrict__ tipmap,
          int maxstates)
{
  extern __shared__ __align__(16) double smem[];
  const plf_op_t op = ops[blockIdx.y];
  double2 * sA = reinterpret_cast<double2 *>(smem); /* [AA_SPT*10][AA_THREADS] */
  double * lmat = smem + AA_SPT * 20 * AA_THREADS;
  double * rmat = lmat + (KIND == PLF_OP_II ? R * 400 : 0);
  double * tl = rmat + R * 400;
  for (int e = threadIdx.x; e < R * 400; e += blockDim.x)
  {
    if (KIND == PLF_OP_II) lmat[e] = op.left_matrix[e];
    rmat[e] = op.right_matrix[e];
  }
  if (KIND == PLF_OP_TI)
    for (int e = threadIdx.x; e < maxstates * R * 20; e += blockDim.x)
    {
      const int c = e / (R * 20), r = (e / 20) % R, i = e % 20;
      tl[(c * R + r) * AA_TAB_STRIDE + i] = masked_sum_seq(op.left_matrix + r * 400 + i * 20, tipmap[c], 20);
    }
  __syncthreads();

  const unsigned int nsites = op.nsites;
  const size_t span = (size_t)R * 20;
  const unsigned int chunk = gridDim.x * AA_THREADS;
  const unsigned int t0 = blockIdx.x * AA_THREADS + threadIdx.x;
  const bool gather = op.parent_id_site || op.left_site_id || op.right_site_id;
  const unsigned int iters = (nsites + chunk * AA_SPT - 1) / (chunk * AA_SPT);

  unsigned int lid[AA_SPT], rid[AA_SPT], nlid[AA_SPT], nrid[AA_SPT];
  auto resolve = [&](unsigned int it, unsigned int (&l)[AA_SPT], unsigned int (&r)[AA_SPT]) {
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      const unsigned int nn = it * chunk * AA_SPT + s * chunk + t0;
      l[s] = r[s] = nn < nsites ? nn : 0;
      if (nn < nsites && gather)
      {
        const unsigned int site = op.parent_id_site ? op.parent_id_site[nn] : nn;
        l[s] = op.left_site_id ? op.left_site_id[site] : site;
        r[s] = op.right_site_id ? op.right_site_id[site] : site;
      }
    }
  };
  if (!iters) return;
  resolve(0, lid, rid);
  double cur[AA_SPT][20], nxt[AA_SPT][20];
#pragma unroll
  for (int s = 0; s < AA_SPT; ++s)
    load20(nxt[s], (KIND == PLF_OP_II ? op.left_clv + (size_t)lid[s] * span : op.right_clv + (size_t)rid[s] * span), true);

  for (unsigned int it = 0; it < iters; ++it)
  {
    unsigned int n[AA_SPT], code[AA_SPT];
    bool act[AA_SPT];
    int below_all[AA_SPT];
    resolve(it + 1, nlid, nrid);
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      n[s] = it * chunk * AA_SPT + s * chunk + t0;
      act[s] = n[s] < nsites;
      below_all[s] = 1;
      code[s] = (KIND == PLF_OP_TI && act[s]) ? op.left_tip[lid[s]] : 0u;
    }
    for (int rate = 0; rate < R; ++rate)
    {
      if (KIND == PLF_OP_II)
      {
#pragma unroll
        for (int s = 0; s < AA_SPT; ++s)
        {
#pragma unroll
          for (int j = 0; j < 20; ++j) cur[s][j] = nxt[s][j];
          load20(nxt[s], op.right_clv + (size_t)rid[s] * span + rate * 20, true);
        }
#pragma unroll
        for (int i = 0; i < 20; i += 4)
        {
          double o[AA_SPT][4];
          rows4_fma(lmat + rate * 400 + i * 20, cur, o);
#pragma unroll
          for (int s = 0; s < AA_SPT; ++s)
          {
            sA[(s * 10 + i / 2) * AA_THREADS + threadIdx.x] = make_double2(o[s][0], o[s][1]);
            sA[(s * 10 + i / 2 + 1) * AA_THREADS + threadIdx.x] = make_double2(o[s][2], o[s][3]);
          }
        }
      }
      /* right phase: request the operand after this one */
#pragma unroll
      for (int s = 0; s < AA_SPT; ++s)
      {
#pragma unroll
        for (int j = 0; j < 20; ++j) cur[s][j] = nxt[s][j];
        const double * nsrc = nullptr;
        if (rate + 1 < R)
          nsrc = (KIND == PLF_OP_II ? op.left_clv + (size_t)lid[s] * span : op.right_clv + (size_t)rid[s] * span) + (rate + 1) * 20;
        else if (it + 1 < iters)
          nsrc = (KIND == PLF_OP_II ? op.left_clv + (size_t)nlid[s] * span : op.right_clv + (size_t)nrid[s] * span);
        if (nsrc) load20(nxt[s], nsrc, true);
      }
      int below[AA_SPT];
#pragma unroll
      for (int s = 0; s < AA_SPT; ++s) below[s] = 1;
#pragma unroll
      for (int i = 0; i < 20; i += 4)
      {
        double o[AA_SPT][4];
        rows4_fma(rmat + rate * 400 + i * 20, cur, o);
#pragma unroll
        for (int s = 0; s < AA_SPT; ++s)
        {
          double2 a01, a23;
          if (KIND == PLF_OP_II)
          {
            a01 = sA[(s * 10 + i / 2) * AA_THREADS + threadIdx.x];
            a23 = sA[(s * 10 + i / 2 + 1) * AA_THREADS + threadIdx.x];
          }
          else
          {
            const double * row = tl + ((size_t)code[s] * R + rate) * AA_TAB_STRIDE + i;
            a01 = *reinterpret_cast<const double2 *>(row);
            a23 = *reinterpret_cast<const double2 *>(row + 2);
          }
          dbl4 v;
          v.x = a01.x * o[s][0];
          v.y = a01.y * o[s][1];
          v.z = a23.x * o[s][2];
          v.w = a23.y * o[s][3];
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
#pragma unroll
    for (int s = 0; s < AA_SPT; ++s)
    {
      lid[s] = nlid[s];
      rid[s] = nrid[s];
    }
  }
}

/* ---- one site per thread, operands prefetched one phase ahead ------------------ *
 * The sequence of 20-double operand vectors a thread consumes is                   *
 *   left(rate 0), right(rate 0), left(rate 1), ... (ii)   or  right(rate 0..R-1) (ti) *
 * and the vector of phase p+1 is requested before the arithmetic of phase p         *
 * starts, so global-load latency overlaps ~460 FP64 instructions.  Fewer registers  *
 * than the two-site variant (3 CTAs/SM), at twice the shared-memory reads per FMA.  */
__device__ __forceinline__ void rows_all_fma1(const double * __restrict__ m, const double (&c)[20], double (&out)[20])
{
#pragma unroll
  for (int i = 0; i < 20; ++i)
  {
    const double * row = m + i * 20;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int j = 0; j < 20; j += 4)
    {
      const double2 m01 = *reinterpret_cast<const double2 *>(row + j);
      const double2 m23 = *reinterpret_cast<const double2 *>(row + j + 2);
      a0 = fma(m01.x, c[j + 0], a0);
      a1 = fma(m01.y, c[j + 1], a1);
      a2 = fma(m23.x, c[j + 2], a2);
      a3 = fma(m23.y, c[j + 3], a3);
    }
    out[i] = (a0 + a1) + (a2 + a3);
  }
}

__device__ __forceinline__ void rows4_fma1(const double * __restrict__ m, const double (&c)[20], double (&out)[4])
{
#pragma unroll
  for (int q = 0; q < 4; ++q)
  {
    const double * row = m + q * 20;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int j = 0; j < 20; j += 4)
    {
      const double2 m01 = *reinterpret_cast<const double2 *>(row + j);
      const double2 m23 = *reinterpret_cast<const double2 *>(row + j + 2);
      a0 = fma(m01.x, c[j + 0], a0);
      a1 = fma(m01.y, c[j + 1], a1);
      a2 = fma(m23.x, c[j + 2], a2);
      a3 = fma(m23.y, c[j + 3], a3);
    }
    out[q] = (a0 + a1) + (a2 + a3);
  }
}

template <int KIND>
__global__ void __launch_bounds__(AA_THREADS, 3)
k_clv_aa1(const plf_op_t * __restrict__ ops, int R, int per_rate, const plf_state_t * __restrict__ tipmap,
          int maxstates)
{
  extern __shared__ __align__(16) double smem[];
  const plf_op_t op = ops[blockIdx.y];
  double * lmat = smem;
  double * rmat = smem + (KIND == PLF_OP_II ? R * 400 : 0);
  double * tl = rmat + R * 400;
  for (int e = threadIdx.x; e < R * 400; e += blockDim.x)
  {
    if (KIND == PLF_OP_II) lmat[e] = op.left_matrix[e];
    rmat[e] = op.right_matrix[e];
  }
  if (KIND == PLF_OP_TI)
    for (int e = threadIdx.x; e < maxstates * R * 20; e += blockDim.x)
    {
      const int c = e / (R * 20), r = (e / 20) % R, i = e % 20;
      tl[(c * R + r) * AA_TAB_STRIDE + i] = masked_sum_seq(op.left_matrix + r * 400 + i * 20, tipmap[c], 20);
    }
  __syncthreads();

  const unsigned int nsites = op.nsites;
  const size_t span = (size_t)R * 20;
  const unsigned int chunk = gridDim.x * AA_THREADS;
  const bool gather = op.parent_id_site || op.left_site_id || op.right_site_id;

  unsigned int n = blockIdx.x * AA_THREADS + threadIdx.x;
  unsigned int lid = 0, rid = 0;
  auto resolve = [&](unsigned int nn, unsigned int & l, unsigned int & r) {
    l = r = nn < nsites ? nn : 0;
    if (nn < nsites && gather)
    {
      const unsigned int site = op.parent_id_site ? op.parent_id_site[nn] : nn;
      l = op.left_site_id ? op.left_site_id[site] : site;
      r = op.right_site_id ? op.right_site_id[site] : site;
    }
  };
  resolve(n, lid, rid);
  double ping[20], pong[20];
  /* prologue: first operand of the first site */
  if (KIND == PLF_OP_II)
    load20(ping, op.left_clv + (size_t)lid * span, true);
  else
    load20(pong, op.right_clv + (size_t)rid * span, true);

  const unsigned int iters = (nsites + chunk - 1) / chunk; /* same trip count for every thread */
  for (unsigned int it = 0; it < iters; ++it, n += chunk)
  {
    const bool act = n < nsites;
    unsigned int nlid, nrid;
    resolve(n + chunk, nlid, nrid);
    const bool has_next = it + 1 < iters;
    const unsigned int code = (KIND == PLF_OP_TI && act) ? op.left_tip[lid] : 0u;
    int below_all = 1;
    for (int rate = 0; rate < R; ++rate)
    {
      double A[20];
      if (KIND == PLF_OP_II)
      {
        /* phase "left": request right(rate), compute A from ping */
        load20(pong, op.right_clv + (size_t)rid * span + rate * 20, true);
        rows_all_fma1(lmat + rate * 400, ping, A);
      }
      else
      {
        const double * row = tl + ((size_t)code * R + rate) * AA_TAB_STRIDE;
#pragma unroll
        for (int i = 0; i < 20; i += 2)
        {
          const double2 t = *reinterpret_cast<const double2 *>(row + i);
          A[i] = t.x;
          A[i + 1] = t.y;
        }
      }
      /* phase "right": request the next operand, compute B from pong */
      double cur[20];
      if (KIND == PLF_OP_II)
      {
        if (rate + 1 < R)
          load20(ping, op.left_clv + (size_t)lid * span + (rate + 1) * 20, true);
        else if (has_next)
          load20(ping, op.left_clv + (size_t)nlid * span, true);
#pragma unroll
        for (int j = 0; j < 20; ++j) cur[j] = pong[j];
      }
      else
      {
#pragma unroll
        for (int j = 0; j < 20; ++j) cur[j] = pong[j];
        if (rate + 1 < R)
          load20(pong, op.right_clv + (size_t)rid * span + (rate + 1) * 20, true);
        else if (has_next)
          load20(pong, op.right_clv + (size_t)nrid * span, true);
      }
      int below = 1;
      double * pout = op.parent_clv + (size_t)n * span + rate * 20;
#pragma unroll
      for (int i = 0; i < 20; i += 4)
      {
        double B[4];
        rows4_fma1(rmat + rate * 400 + i * 20, cur, B);
        dbl4 v;
        v.x = A[i] * B[0];
        v.y = A[i + 1] * B[1];
        v.z = A[i + 2] * B[2];
        v.w = A[i + 3] * B[3];
        below &= (v.x < PLF_SCALE_THRESHOLD) && (v.y < PLF_SCALE_THRESHOLD) && (v.z < PLF_SCALE_THRESHOLD) &&
                 (v.w < PLF_SCALE_THRESHOLD);
        if (act) st256(pout + i, v);
      }
      below_all &= below;
      if (op.parent_scaler && per_rate && act)
      {
        unsigned int sc = 0;
        if (KIND == PLF_OP_II && op.left_scaler) sc += op.left_scaler[(size_t)lid * R + rate];
        if (op.right_scaler) sc += op.right_scaler[(size_t)rid * R + rate];
        if (below)
        {
#pragma unroll
          for (int i = 0; i < 20; i += 4)
          {
            dbl4 v = ld256(pout + i);
            v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR; v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
            st256(pout + i, v);
          }
          sc += 1;
        }
        op.parent_scaler[(size_t)n * R + rate] = sc;
      }
    }
    if (op.parent_scaler && !per_rate && act)
    {
      unsigned int sc = 0;
      if (KIND == PLF_OP_II && op.left_scaler) sc += op.left_scaler[lid];
      if (op.right_scaler) sc += op.right_scaler[rid];
      if (below_all)
      {
        double * p = op.parent_clv + (size_t)n * span;
        for (int i = 0; i < R * 20; i += 4)
        {
          dbl4 v = ld256(p + i);
          v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR; v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
          st256(p + i, v);
        }
        sc += 1;
      }
      op.parent_scaler[n] = sc;
    }
    lid = nlid;
    rid = nrid;
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
  const int variant = ctx->aa_spt; /* 1: one site + register prefetch, 2: two sites, 3: two sites + cp.async staging */
  const int spt = variant == 1 ? 1 : 2;
  if (variant == 3) smem += (size_t)AA2_STAGES * AA2_STAGE_BYTES;
  if (variant == 4) smem += (size_t)AA_SPT * 20 * AA_THREADS * sizeof(double);
  if (smem > ctx->smem_optin) return -1; /* caller falls back to the generic kernel */
  void (*k)(const plf_op_t *, int, int, const plf_state_t *, int) =
      variant == 1   ? (ii ? k_clv_aa1<PLF_OP_II> : k_clv_aa1<PLF_OP_TI>)
      : variant == 2 ? (ii ? k_clv_aa<PLF_OP_II> : k_clv_aa<PLF_OP_TI>)
      : variant == 4 ? (ii ? k_clv_aa3<PLF_OP_II> : k_clv_aa3<PLF_OP_TI>)
                     : (ii ? k_clv_aa2<PLF_OP_II> : k_clv_aa2<PLF_OP_TI>);
  size_t & set = ctx->aa_smem_set[ii ? 0 : 1];
  (void)spt;
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
  unsigned long long need = ((unsigned long long)max_sites + AA_THREADS * spt - 1) / (AA_THREADS * spt);
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
