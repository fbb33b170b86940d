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

/* tip lookup: tab[code][rate][i] = sum of the columns of row i selected by the
 * 4-bit state mask, pairwise order (src/core_partials_avx.c:1336-1395) */
__device__ __forceinline__ void build_tip_table(double * tab, const double * __restrict__ matrix, int R)
{
  for (int e = threadIdx.x; e < 64 * R; e += blockDim.x)
  {
    const int code = e / (4 * R), r = (e >> 2) % R, i = e & 3;
    tab[e] = masked_sum4(matrix + r * 16 + i * 4, code);
  }
}

__device__ __forceinline__ dbl4 lds_dbl4(const double * p)
{
  const double2 a = *reinterpret_cast<const double2 *>(p);
  const double2 b = *reinterpret_cast<const double2 *>(p + 2);
  return dbl4{a.x, a.y, b.x, b.y};
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
template <int LOG2R, int U>
__global__ void __launch_bounds__(DNA_THREADS, 4)
k_clv_dna_tt(const plf_op_t * __restrict__ ops, int per_rate)
{
  constexpr int R = 1 << LOG2R;
  __shared__ __align__(16) double tl[64 * R];
  __shared__ __align__(16) double tr[64 * R];
  const plf_op_t op = ops[blockIdx.y];
  build_tip_table(tl, op.left_matrix, R);
  build_tip_table(tr, op.right_matrix, R);
  __syncthreads();
  const unsigned int tid = blockIdx.x * DNA_THREADS + threadIdx.x;
  const int rate = tid & (R - 1);
  const unsigned int site0 = tid >> LOG2R;
  const unsigned int pass = (gridDim.x * DNA_THREADS) >> LOG2R;

  for (unsigned int base = 0; base < op.nsites; base += pass * U)
  {
    unsigned int n[U], lc[U], rc[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      n[u] = base + u * pass + site0;
      lc[u] = rc[u] = 0;
      if (n[u] < op.nsites)
      {
        const unsigned int site = op.parent_id_site ? op.parent_id_site[n[u]] : n[u];
        lc[u] = op.left_tip[op.left_site_id ? op.left_site_id[site] : site];
        rc[u] = op.right_tip[op.right_site_id ? op.right_site_id[site] : site];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      if (n[u] >= op.nsites) continue;
      const dbl4 a = lds_dbl4(tl + (lc[u] * R + rate) * 4);
      const dbl4 b = lds_dbl4(tr + (rc[u] * R + rate) * 4);
      const dbl4 v = {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
      st256(op.parent_clv + ((size_t)n[u] * R + rate) * 4, v);
      if (op.parent_scaler)
      {
        if (per_rate)
          op.parent_scaler[(size_t)n[u] * R + rate] = 0;
        else if (rate == 0)
          op.parent_scaler[n[u]] = 0;
      }
    }
  }
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

/* launch one group of same-kind DNA ops (rate_cats a power of two <= 32) as a
 * single persistent wave: gridDim.y = ops, gridDim.x = CTAs striding over the
 * sites of each op */
int plf_launch_dna_group(plf_ctx * ctx, const plf_op_t * d_ops, unsigned int nops, unsigned int kind,
                         unsigned int rate_cats, int per_rate, unsigned int max_sites)
{
  dna_kernel_t k = nullptr;
  int log2r = 0;
  while ((1u << log2r) < rate_cats) ++log2r;
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
  unsigned long long need = (lanes + (unsigned long long)DNA_THREADS * DNA_UNROLL - 1) / ((unsigned long long)DNA_THREADS * DNA_UNROLL);
  unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
  if (bx < 1) bx = 1;
  if (bx > need) bx = need;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned int)bx, nops);
  k<<<grid, DNA_THREADS, 0, ctx->stream>>>(d_ops, per_rate);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}
