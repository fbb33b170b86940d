/*
 * plf_partials_aa_mma.cu -- 20-state (protein) CLV updates on the FP64 tensor
 * path (DMMA, mma.sync.m8n8k4.f64), and the 20-state tip-tip kernel.
 *
 * Replaces pll_core_update_partial_ii_20x20_avx2 / _ti_20x20_avx2 / _tt
 * (reference src/core_partials_avx2.c:630,343, src/core_partials.c:82) and the
 * scaler pass src/pll.c:1202.
 *
 * Why tensor cores here and nowhere else: one inner-inner site update is
 * 2 x 4 x 400 FMAs against 1932 bytes, 3.3 flop/B.  Measured on this B200
 * (profiles/r1_fp64_pipe_bench.txt): DFMA peaks at 58 FMA lanes/clk/SM and the
 * register-tiled DFMA kernel (plf_partials_aa.cu) reached 37 % of the FP64
 * pipe with 12 % occupancy (255 registers, matrix operands through LDS) --
 * issue/latency-bound at 2.0 TB/s.  DMMA sustains 63.8 lanes/clk/SM from 4
 * warps per SM, needs one instruction per 256 FMAs and takes the P-matrix
 * operand as a 1-register fragment, so the same arithmetic costs ~20 us per
 * 100k-site op and the kernel becomes HBM-bound like the DNA ones.
 *
 * Mapping (m8n8k4, A row-major 8x4, B 4x8, C/D 8x8):
 *   A = child CLV block of 8 sites x 4 states, B = P^T (4 states x 8 parent
 *   states), D[site][i] accumulates over 5 k-tiles; 3 n-tiles cover the 20
 *   parent states (rows 20..23 of the last tile are zero padding).
 *   The k order is permuted so that a lane's A elements of two k-tiles are
 *   adjacent in memory (16-byte loads): k-tile (2p, 2p+1) slot q <-> states
 *   8p + 2q, 8p + 2q + 1; k-tile 4 slot q <-> state 16 + q.
 *   D fragment: lane holds 2 consecutive parent states of one site, stored
 *   with one 16-byte store; the two children's D fragments line up, so the
 *   product and the scaling test need no data movement.
 *
 * Numerics: DMMA accumulates the 20 products of a row in tensor-core order,
 * not in the reference's 4-lane-FMA-then-pairwise order, so CLV entries can
 * differ from the AVX2 reference in the last bits (observed <= 4 ulp).  Integer
 * scalers, which the contract requires to be exact, would only differ if an
 * entry landed within those ulps of 2^-256 (never observed; the bit-exact DFMA
 * kernels stay selectable with PLF_AA_MMA=0 and are what the parity tests pin
 * bit-for-bit).  Log-likelihoods agree to <1e-13 relative.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"

#include <stdlib.h>

#include "plf_mma.cuh"
#include "plf_stream.cuh"

/* L2::256B: the 160-byte block of one (site, rate) straddles 128-byte lines that the
 * neighbouring rate steps of the same site need a moment later; asking L2 to bring the
 * whole 256-byte neighbourhood keeps DRAM reads at the algorithmic byte count */
template <int PF>
__device__ __forceinline__ double2 ldg_stream_v2(const double * p)
{
  double2 v;
  if (PF)
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
template <int PF>
__device__ __forceinline__ double ldg_stream(const double * p)
{
  double v;
  if (PF)
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
struct AamSite
{
  unsigned int n, lid, rid, code;
  bool act;
};

__device__ __forceinline__ AamSite aam_resolve(const plf_op_t & op, unsigned int n, bool gather, bool tip)
{
  AamSite s;
  s.act = n < op.nsites;
  s.n = s.act ? n : op.nsites - 1; /* inactive rows compute a copy of the last site and store nothing */
  s.lid = s.rid = s.n;
  if (gather)
  {
    const unsigned int site = op.parent_id_site ? op.parent_id_site[s.n] : s.n;
    s.lid = op.left_site_id ? op.left_site_id[site] : site;
    s.rid = op.right_site_id ? op.right_site_id[site] : site;
  }
  s.code = tip ? op.left_tip[s.lid] : 0u;
  return s;
}

/* A fragments of one child block (20 doubles at `p`): 2 x 16-byte + 1 x 8-byte load */
template <int PF>
__device__ __forceinline__ void aam_load_t(double (&a)[5], const double * __restrict__ p, int q)
{
  const double2 v0 = ldg_stream_v2<PF>(p + 2 * q);
  const double2 v1 = ldg_stream_v2<PF>(p + 8 + 2 * q);
  a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
  a[4] = ldg_stream<PF>(p + 16 + q);
}
#define aam_load aam_load_t<L2PF>

template <int KIND, int L2PF>
__global__ void __launch_bounds__(AAM_THREADS, 2)
k_clv_aa_mma(const plf_op_t * __restrict__ ops, int R, int per_rate, const plf_state_t * __restrict__ tipmap,
             int maxstates)
{
  extern __shared__ __align__(16) double smem[];
  const plf_op_t op = ops[blockIdx.y];
  constexpr int NMAT = (KIND == PLF_OP_II) ? 2 : 1;
  double * bfr = smem;                                  /* [NMAT][R][AAM_FRAGS][32] B fragments of P^T */
  double * tl = smem + (size_t)NMAT * R * AAM_FRAGS * 32; /* TI: [maxstates][R][AAM_TAB_STRIDE] */
  for (int e = threadIdx.x; e < NMAT * R * AAM_FRAGS * 32; e += blockDim.x)
  {
    const int ln = e & 31, f = (e >> 5) % AAM_FRAGS, rate = (e / (AAM_FRAGS * 32)) % R, mat = e / (AAM_FRAGS * 32 * R);
    const int nt = f / 5, kt = f % 5;
    const int i = 8 * nt + (ln >> 2), j = aam_state(kt, ln & 3);
    const double * M = (KIND == PLF_OP_II && mat == 0) ? op.left_matrix : op.right_matrix;
    bfr[e] = (i < 20) ? M[rate * 400 + i * 20 + j] : 0.0;
  }
  if (KIND == PLF_OP_TI)
  {
    /* scalar sums in increasing column order (src/core_partials_avx2.c:387-456): same table bits as the reference */
    for (int e = threadIdx.x; e < maxstates * R * AAM_TAB_STRIDE; e += blockDim.x)
    {
      const int c = e / (R * AAM_TAB_STRIDE), r = (e / AAM_TAB_STRIDE) % R, i = e % AAM_TAB_STRIDE;
      tl[e] = (i < 20) ? masked_sum_seq(op.left_matrix + r * 400 + i * 20, tipmap[c], 20) : 0.0;
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, q = lane & 3, gs = lane >> 2;
  const unsigned int warps_total = gridDim.x * (AAM_THREADS / 32);
  const unsigned int ngroups = (op.nsites + 7) >> 3; /* 8 sites per warp step */
  const size_t span = (size_t)R * 20;
  const bool gather = op.parent_id_site || op.left_site_id || op.right_site_id;
  const double * bL = bfr + lane;
  const double * bR = bfr + (size_t)(NMAT - 1) * R * AAM_FRAGS * 32 + lane;

  unsigned int g = blockIdx.x * (AAM_THREADS / 32) + (threadIdx.x >> 5);
  if (g >= ngroups) return;
  AamSite s = aam_resolve(op, g * 8 + gs, gather, KIND == PLF_OP_TI);
  double cl[5], cr[5], nl[5], nr[5];
  if (KIND == PLF_OP_II) aam_load(cl, op.left_clv + (size_t)s.lid * span, q);
  aam_load(cr, op.right_clv + (size_t)s.rid * span, q);

  while (true)
  {
    int below_all = 1;
    AamSite s2 = s;
    unsigned int g2 = g;
    for (int rate = 0; rate < R; ++rate)
    {
      /* ---- prefetch the next step's operands ---- */
      if (rate + 1 < R)
      {
        if (KIND == PLF_OP_II) aam_load(nl, op.left_clv + (size_t)s.lid * span + (rate + 1) * 20, q);
        aam_load(nr, op.right_clv + (size_t)s.rid * span + (rate + 1) * 20, q);
      }
      else
      {
        g2 = g + warps_total;
        if (g2 < ngroups)
        {
          s2 = aam_resolve(op, g2 * 8 + gs, gather, KIND == PLF_OP_TI);
          if (KIND == PLF_OP_II) aam_load(nl, op.left_clv + (size_t)s2.lid * span, q);
          aam_load(nr, op.right_clv + (size_t)s2.rid * span, q);
        }
      }
      /* ---- 8 sites x 20 parent states of this rate ---- */
      double accL[3][2], accR[3][2];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) accL[nt][0] = accL[nt][1] = accR[nt][0] = accR[nt][1] = 0.0;
      const double * fl = bL + (size_t)rate * AAM_FRAGS * 32;
      const double * fr = bR + (size_t)rate * AAM_FRAGS * 32;
#pragma unroll
      for (int kt = 0; kt < 5; ++kt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          if (KIND == PLF_OP_II) dmma(accL[nt], cl[kt], fl[(nt * 5 + kt) * 32]);
          dmma(accR[nt], cr[kt], fr[(nt * 5 + kt) * 32]);
        }
      if (KIND == PLF_OP_TI)
      {
        const double * row = tl + ((size_t)s.code * R + rate) * AAM_TAB_STRIDE + 2 * q;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          if (nt < 2 || q < 2)
          {
            const double2 t = *reinterpret_cast<const double2 *>(row + 8 * nt);
            accL[nt][0] = t.x;
            accL[nt][1] = t.y;
          }
        }
      }
      double v[3][2];
      int below = 1;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        v[nt][0] = accL[nt][0] * accR[nt][0];
        v[nt][1] = accL[nt][1] * accR[nt][1];
        if (nt < 2 || q < 2) below &= (v[nt][0] < PLF_SCALE_THRESHOLD) && (v[nt][1] < PLF_SCALE_THRESHOLD);
      }
      if (op.parent_scaler && per_rate)
      {
        below &= __shfl_xor_sync(0xffffffffu, below, 1);
        below &= __shfl_xor_sync(0xffffffffu, below, 2);
        if (below)
        {
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
          {
            v[nt][0] *= PLF_SCALE_FACTOR;
            v[nt][1] *= PLF_SCALE_FACTOR;
          }
        }
        if (s.act && q == 0)
        {
          unsigned int sc = below ? 1u : 0u;
          if (KIND == PLF_OP_II && op.left_scaler) sc += op.left_scaler[(size_t)s.lid * R + rate];
          if (op.right_scaler) sc += op.right_scaler[(size_t)s.rid * R + rate];
          op.parent_scaler[(size_t)s.n * R + rate] = sc;
        }
      }
      below_all &= below;
      if (s.act)
      {
        double * out = op.parent_clv + (size_t)s.n * span + rate * 20 + 2 * q;
        stg_v2(out, v[0][0], v[0][1]);
        stg_v2(out + 8, v[1][0], v[1][1]);
        if (q < 2) stg_v2(out + 16, v[2][0], v[2][1]);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k)
      {
        if (KIND == PLF_OP_II) cl[k] = nl[k];
        cr[k] = nr[k];
      }
    }
    /* ---- per-site scaling: decided after all rates; the rare site that
     * scales is rescaled in place, every lane revisiting what it stored ---- */
    if (op.parent_scaler && !per_rate)
    {
      below_all &= __shfl_xor_sync(0xffffffffu, below_all, 1);
      below_all &= __shfl_xor_sync(0xffffffffu, below_all, 2);
      if (s.act)
      {
        if (below_all)
        {
          for (int rate = 0; rate < R; ++rate)
          {
            double * out = op.parent_clv + (size_t)s.n * span + rate * 20 + 2 * q;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
              if (nt < 2 || q < 2)
              {
                double2 t = *reinterpret_cast<double2 *>(out + 8 * nt);
                stg_v2(out + 8 * nt, t.x * PLF_SCALE_FACTOR, t.y * PLF_SCALE_FACTOR);
              }
          }
        }
        if (q == 0)
        {
          unsigned int sc = below_all ? 1u : 0u;
          if (KIND == PLF_OP_II && op.left_scaler) sc += op.left_scaler[s.lid];
          if (op.right_scaler) sc += op.right_scaler[s.rid];
          op.parent_scaler[s.n] = sc;
        }
      }
    }
    if (g2 >= ngroups || g2 == g) break;
    g = g2;
    s = s2;
  }
}

/* ------------------------------------------------------------------------ *
 *  Streaming variant for contiguous (non-repeats) CLVs and 1, 2, 4 or 8 rate  *
 *  categories.  The direct-load kernel above touches DRAM in 160-byte blocks  *
 *  (one rate of one site), which straddle 128-byte lines: ncu showed 1.6x the *
 *  algorithmic read bytes and long-scoreboard stalls at 16 warps/SM.  Here a  *
 *  tile of 64/R sites (all rates, 10 KB per child, contiguous in memory)      *
 *  lands in a shared-memory ring through ONE bulk async copy per child        *
 *  (cp.async.bulk + mbarrier), 4 tiles in flight per CTA, so DRAM sees whole  *
 *  lines once and bytes in flight do not depend on registers.  Each warp owns *
 *  one rate category and one 8-site block of the tile for the whole kernel:   *
 *  its 2 x 15 B fragments (P^T) stay in registers, A fragments come from the  *
 *  ring with 16-byte LDS.  CTAs are 4 warps (1, 2, 4 rates; the barrier per   *
 *  tile then only joins the warps that must exchange scaling flags and 4      *
 *  independent CTAs per SM keep the DMMA pipe fed) or 8 warps (8 rates).      *
 * ------------------------------------------------------------------------ */

__device__ __forceinline__ void mbar_arrive(unsigned long long * bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

/* No CTA-wide barrier in the tile loop.  Ring slots are handed back through `empty` mbarriers (every
 * warp arrives as soon as its A fragments are in registers, i.e. before its DMMAs), so the copy engine
 * runs up to AAS_NSTAGE tiles ahead of the slowest warp.  Per-site scaling needs the verdict of all
 * rate warps of a site: each warp stores its block unscaled, publishes its flag (`flagbar` mbarrier)
 * and settles the tile one iteration later, when every flag has long arrived: the rare site that
 * scales is rescaled in place by the lanes that stored it. */
/* What a child of a streamed op is: an inner CLV (its tile comes through the ring), a pattern tip (one code
 * per site, 24-entry table of masked row sums in shared memory: the value of the whole term) or a VIRTUAL
 * CHERRY (the codes of its two tips; its CLV entry k is hA[codeA][k] * hB[codeB][k] with
 * the two half tables of the reference's tip-tip kernel, formed in registers as the A fragment of the DMMA
 * with the matrix of the branch above it -- the same operands the kernel would have read back from HBM). */
enum { AK_I = 0, AK_T = 1, AK_C = 2 };

template <int LK, int RK, int LOG2R, int NWARPS, int NST = 4>
__global__ void __launch_bounds__(NWARPS * 32, NST == 4 ? 512 / (NWARPS * 32) : 1)
k_clv_aa_mma_stream(const plf_op_t * __restrict__ ops, int per_rate, const plf_state_t * __restrict__ tipmap,
                    int maxstates)
{
  constexpr int R = 1 << LOG2R;
  constexpr int SB = NWARPS / R;          /* 8-site blocks per tile */
  constexpr int TILE = 8 * SB;            /* sites per tile */
  constexpr int CH_BYTES = TILE * R * 160; /* one child's tile: 5 KB (4 warps) or 10 KB (8 warps) */
  constexpr int NCH = (LK == AK_I ? 1 : 0) + (RK == AK_I ? 1 : 0);
  constexpr int STAGE = NCH * CH_BYTES;
  constexpr int OFF_LCH = (RK == AK_I) ? CH_BYTES : 0; /* the right child's tile comes first */
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) unsigned long long full[NST];
  __shared__ __align__(8) unsigned long long empty[NST];
  __shared__ __align__(8) unsigned long long flagbar[2];
  __shared__ int flags[2][TILE][R];
  unsigned char * ring = dyn;
  /* tables behind the ring: tip table of a left tip [maxstates][R][AAM_TAB_STRIDE], then the half tables of
   * the cherries [maxstates][R][AAM_TAB_STRIDE] each: left tip A, left tip B, right tip A, right tip B */
  double * tl = reinterpret_cast<double *>(dyn + (size_t)NST * STAGE);
  const int half = maxstates * R * AAM_TAB_STRIDE; /* rows 22 doubles apart: codes spread over the banks */
  double * hl1 = tl + (LK == AK_T ? maxstates * R * AAM_TAB_STRIDE : 0);
  double * hl2 = hl1 + (LK == AK_C ? half : 0);
  double * hr1 = hl2 + (LK == AK_C ? half : 0);
  double * hr2 = hr1 + (RK == AK_C ? half : 0);

  const plf_op_t op = ops[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 3, gs = lane >> 2;
  const int rate = warp & (R - 1), sb = warp >> LOG2R;
  const unsigned int ntiles = (op.nsites + TILE - 1) / TILE;
  const size_t span = (size_t)R * 20;
  const bool site_scaling = op.parent_scaler && !per_rate;

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < NST; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NWARPS);
    }
    mbar_init(&flagbar[0], NWARPS);
    mbar_init(&flagbar[1], NWARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (LK == AK_T)
  {
    for (int e = threadIdx.x; e < maxstates * R * AAM_TAB_STRIDE; e += blockDim.x)
    {
      const int c = e / (R * AAM_TAB_STRIDE), r = (e / AAM_TAB_STRIDE) % R, i = e % AAM_TAB_STRIDE;
      tl[e] = (i < 20) ? masked_sum_seq(op.left_matrix + r * 400 + i * 20, tipmap[c], 20) : 0.0;
    }
  }
  if (LK == AK_C || RK == AK_C)
  {
    /* scalar sums in increasing column order, as the reference's tip-tip table (src/core_partials_avx.c:124-253) */
    for (int e = threadIdx.x; e < half; e += blockDim.x)
    {
      const int c = e / (R * AAM_TAB_STRIDE), r = (e / AAM_TAB_STRIDE) % R, i = e % AAM_TAB_STRIDE;
      const plf_state_t mask = tipmap[c];
      if (LK == AK_C)
      {
        hl1[e] = (i < 20) ? masked_sum_seq(op.left_cm1 + r * 400 + i * 20, mask, 20) : 0.0;
        hl2[e] = (i < 20) ? masked_sum_seq(op.left_cm2 + r * 400 + i * 20, mask, 20) : 0.0;
      }
      if (RK == AK_C)
      {
        hr1[e] = (i < 20) ? masked_sum_seq(op.right_cm1 + r * 400 + i * 20, mask, 20) : 0.0;
        hr2[e] = (i < 20) ? masked_sum_seq(op.right_cm2 + r * 400 + i * 20, mask, 20) : 0.0;
      }
    }
  }
  __syncthreads();

  auto issue = [&](unsigned int t, int s) {
    const unsigned int first = t * TILE;
    const unsigned int n = min((unsigned int)TILE, op.nsites - first);
    const unsigned int bytes = n * R * 160;
    unsigned char * slot = ring + (size_t)s * STAGE;
    mbar_expect_tx(&full[s], NCH * bytes);
    if (LK == AK_I) bulk_g2s(slot + OFF_LCH, op.left_clv + (size_t)first * span, bytes, &full[s]);
    if (RK == AK_I) bulk_g2s(slot, op.right_clv + (size_t)first * span, bytes, &full[s]);
  };
  if (NCH > 0 && threadIdx.x == 0)
  {
    unsigned int t = blockIdx.x;
    for (int s = 0; s < NST && t < ntiles; ++s, t += gridDim.x) issue(t, s);
  }

  /* B fragments of this warp's rate: lane holds P[8 nt + gs][state(kt, q)] */
  double bl[LK != AK_T ? AAM_FRAGS : 1], br[AAM_FRAGS];
#pragma unroll
  for (int f = 0; f < AAM_FRAGS; ++f)
  {
    const int nt = f / 5, kt = f % 5;
    const int i = 8 * nt + gs, j = aam_state(kt, q);
    br[f] = (i < 20) ? op.right_matrix[rate * 400 + i * 20 + j] : 0.0;
    if (LK != AK_T) bl[f] = (i < 20) ? op.left_matrix[rate * 400 + i * 20 + j] : 0.0;
  }

  const unsigned int my = sb * 8 + gs; /* site within the tile */
  /* tip codes of the site this lane works on in the NEXT iteration: left tip / left cherry (2), right cherry (2) */
  unsigned int cn_l1 = 0, cn_l2 = 0, cn_r1 = 0, cn_r2 = 0;
  auto fetch_codes = [&](unsigned int t) {
    const unsigned int n0 = t * TILE + my;
    const unsigned int nn = n0 < op.nsites ? n0 : op.nsites - 1;
    if (LK != AK_I) cn_l1 = op.left_tip[nn];
    if (LK == AK_C) cn_l2 = op.left_tip2[nn];
    if (RK == AK_C)
    {
      cn_r1 = op.right_tip[nn];
      cn_r2 = op.right_tip2[nn];
    }
  };
  if ((LK != AK_I || RK != AK_I) && blockIdx.x < ntiles) fetch_codes(blockIdx.x);

  /* A fragment of a virtual cherry: entry k = hA[codeA][rate][k] * hB[codeB][rate][k] for the lane's 5 states */
  auto cherry_frag = [&](double (&a)[5], const double * h1, const double * h2, unsigned int c1, unsigned int c2) {
    const double * p1 = h1 + ((size_t)c1 * R + rate) * AAM_TAB_STRIDE;
    const double * p2 = h2 + ((size_t)c2 * R + rate) * AAM_TAB_STRIDE;
    const double2 u0 = *reinterpret_cast<const double2 *>(p1 + 2 * q), w0 = *reinterpret_cast<const double2 *>(p2 + 2 * q);
    const double2 u1 = *reinterpret_cast<const double2 *>(p1 + 8 + 2 * q), w1 = *reinterpret_cast<const double2 *>(p2 + 8 + 2 * q);
    a[0] = u0.x * w0.x;
    a[1] = u0.y * w0.y;
    a[2] = u1.x * w1.x;
    a[3] = u1.y * w1.y;
    a[4] = p1[16 + q] * p2[16 + q];
  };

  /* settle per-site scaling of the tile processed in iteration `itp` (site n_prev, child scalers sc_prev) */
  unsigned int n_prev = 0, sc_prev = 0;
  bool act_prev = false;
  auto settle = [&](unsigned int itp) {
    while (!mbar_try_wait(&flagbar[itp & 1], (itp >> 1) & 1u)) {}
    int fire = 1;
#pragma unroll
    for (int r = 0; r < R; ++r) fire &= flags[itp & 1][my][r];
    if (!act_prev) return;
    if (fire)
    {
      double * out = op.parent_clv + (size_t)n_prev * span + rate * 20 + 2 * q;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
        if (nt < 2 || q < 2)
        {
          const double2 tv = *reinterpret_cast<double2 *>(out + 8 * nt);
          stg_v2(out + 8 * nt, tv.x * PLF_SCALE_FACTOR, tv.y * PLF_SCALE_FACTOR);
        }
    }
    if (q == 0 && rate == 0) op.parent_scaler[n_prev] = sc_prev + (fire ? 1u : 0u);
  };

  unsigned int it = 0;
  for (unsigned int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it)
  {
    const int s = it % NST;
    const unsigned int parity = (it / NST) & 1u;
    const unsigned char * slot = ring + (size_t)s * STAGE;
    /* producer duty: refill the slot of the previous iteration once every warp has handed it back */
    if (NCH > 0 && threadIdx.x == 0 && it > 0)
    {
      const unsigned int tn = t + (unsigned int)(NST - 1) * gridDim.x;
      if (tn < ntiles)
      {
        const int sp = (it - 1) % NST;
        while (!mbar_try_wait(&empty[sp], ((it - 1) / NST) & 1u)) {}
        issue(tn, sp);
      }
    }
    const unsigned int n = t * TILE + my;
    const bool act = n < op.nsites;
    const unsigned int c_l1 = cn_l1, c_l2 = cn_l2, c_r1 = cn_r1, c_r2 = cn_r2;
    if ((LK != AK_I || RK != AK_I) && t + gridDim.x < ntiles) fetch_codes(t + gridDim.x);
    /* child scalers of this site: needed only after the arithmetic (tips and cherries have none) */
    unsigned int sc = 0;
    if (op.parent_scaler && act && q == 0 && (per_rate || rate == 0))
    {
      const size_t k = per_rate ? (size_t)n * R + rate : n;
      if (LK == AK_I && op.left_scaler) sc += op.left_scaler[k];
      if (RK == AK_I && op.right_scaler) sc += op.right_scaler[k];
    }
    if (NCH > 0)
      while (!mbar_try_wait(&full[s], parity)) {}

    double accL[3][2], accR[3][2];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) accL[nt][0] = accL[nt][1] = accR[nt][0] = accR[nt][1] = 0.0;
    /* A fragments of both children first (the slot is released before any arithmetic) */
    double ar[5], al[LK != AK_T ? 5 : 1];
    if (RK == AK_I)
    {
      const double * pr = reinterpret_cast<const double *>(slot) + ((size_t)my * R + rate) * 20;
      const double2 r0 = *reinterpret_cast<const double2 *>(pr + 2 * q);
      const double2 r1 = *reinterpret_cast<const double2 *>(pr + 8 + 2 * q);
      ar[0] = r0.x; ar[1] = r0.y; ar[2] = r1.x; ar[3] = r1.y;
      ar[4] = pr[16 + q];
    }
    else
      cherry_frag(ar, hr1, hr2, c_r1, c_r2);
    if (LK == AK_I)
    {
      const double * pl = reinterpret_cast<const double *>(slot + OFF_LCH) + ((size_t)my * R + rate) * 20;
      const double2 l0 = *reinterpret_cast<const double2 *>(pl + 2 * q);
      const double2 l1 = *reinterpret_cast<const double2 *>(pl + 8 + 2 * q);
      al[0] = l0.x; al[1] = l0.y; al[2] = l1.x; al[3] = l1.y;
      al[4] = pl[16 + q];
    }
    else if (LK == AK_C)
    {
      double tmp[5];
      cherry_frag(tmp, hl1, hl2, c_l1, c_l2);
#pragma unroll
      for (int k = 0; k < 5; ++k) al[k] = tmp[k];
    }
    if (NCH > 0)
    {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]); /* this warp is done with the slot */
    }
    if (LK != AK_T)
    {
      /* the two DMMA chains interleaved: 6 independent accumulators in flight */
#pragma unroll
      for (int kt = 0; kt < 5; ++kt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          dmma(accR[nt], ar[kt], br[nt * 5 + kt]);
          dmma(accL[nt], al[kt], bl[nt * 5 + kt]);
        }
    }
    else
    {
#pragma unroll
      for (int kt = 0; kt < 5; ++kt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) dmma(accR[nt], ar[kt], br[nt * 5 + kt]);
      const double * row = tl + ((size_t)c_l1 * R + rate) * AAM_TAB_STRIDE + 2 * q;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
        if (nt < 2 || q < 2)
        {
          const double2 tv = *reinterpret_cast<const double2 *>(row + 8 * nt);
          accL[nt][0] = tv.x;
          accL[nt][1] = tv.y;
        }
    }
    /* the previous tile's flags have all arrived by now */
    if (site_scaling && it > 0) settle(it - 1);

    double v[3][2];
    int below = 1;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
    {
      v[nt][0] = accL[nt][0] * accR[nt][0];
      v[nt][1] = accL[nt][1] * accR[nt][1];
      if (nt < 2 || q < 2) below &= (v[nt][0] < PLF_SCALE_THRESHOLD) && (v[nt][1] < PLF_SCALE_THRESHOLD);
    }
    if (op.parent_scaler)
    {
      below &= __shfl_xor_sync(0xffffffffu, below, 1);
      below &= __shfl_xor_sync(0xffffffffu, below, 2);
      if (per_rate)
      {
        if (below)
        {
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
          {
            v[nt][0] *= PLF_SCALE_FACTOR;
            v[nt][1] *= PLF_SCALE_FACTOR;
          }
        }
        if (act && q == 0) op.parent_scaler[(size_t)n * R + rate] = sc + (below ? 1u : 0u);
      }
      else
      {
        if (q == 0) flags[it & 1][my][rate] = below;
        __syncwarp();
        if (lane == 0) mbar_arrive(&flagbar[it & 1]);
      }
    }
    if (act)
    {
      double * out = op.parent_clv + (size_t)n * span + rate * 20 + 2 * q;
      stg_v2(out, v[0][0], v[0][1]);
      stg_v2(out + 8, v[1][0], v[1][1]);
      if (q < 2) stg_v2(out + 16, v[2][0], v[2][1]);
    }
    n_prev = n;
    sc_prev = sc;
    act_prev = act;
  }
  if (site_scaling && it > 0) settle(it - 1);
}

/* a virtual cherry's own "operation" (20 states): snapshot of its two P-matrices into the node's side buffer
 * (parent_clv), scaler zeroed (src/core_partials.c:82-200 zeroes it too) */
__global__ void __launch_bounds__(256)
k_cherry_prepare_aa(const plf_op_t * __restrict__ ops, int per_rate, int R, int msz)
{
  const plf_op_t op = ops[blockIdx.y];
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < msz; i += blockDim.x)
    {
      op.parent_clv[i] = op.left_matrix[i];
      op.parent_clv[msz + i] = op.right_matrix[i];
    }
  if (!op.parent_scaler) return;
  const size_t n = per_rate ? (size_t)op.nsites * R : op.nsites;
  const size_t quads = (n + 3) >> 2; /* the allocation is padded by 16 bytes */
  uint4 * dst = reinterpret_cast<uint4 *>(op.parent_scaler);
  for (size_t qd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; qd < quads; qd += (size_t)gridDim.x * blockDim.x)
    dst[qd] = make_uint4(0, 0, 0, 0);
}

/* ------------------------------------------------------------------------ *
 *  tip-tip, 20 states: parent = tl[code_l] * tr[code_r], never scales, the   *
 *  scaler is zeroed (src/core_partials.c:82-200).  Write-only, 646 B/site:   *
 *  one thread per 32-byte chunk (4 states of one rate), 256-bit stores, a     *
 *  warp writes 1 KB contiguous.                                               *
 * ------------------------------------------------------------------------ */
#define AATT_THREADS 256
#define AATT_U 4

__global__ void __launch_bounds__(AATT_THREADS)
k_clv_aa_tt(const plf_op_t * __restrict__ ops, int R, int per_rate, const plf_state_t * __restrict__ tipmap,
            int maxstates)
{
  extern __shared__ __align__(16) double smem[];
  const plf_op_t op = ops[blockIdx.y];
  double * tl = smem;                                  /* [maxstates][R][20] */
  double * tr = smem + (size_t)maxstates * R * 20;
  for (int e = threadIdx.x; e < maxstates * R * 20; e += blockDim.x)
  {
    const int c = e / (R * 20), r = (e / 20) % R, i = e % 20;
    const plf_state_t mask = tipmap[c];
    tl[e] = masked_sum_seq(op.left_matrix + r * 400 + i * 20, mask, 20);
    tr[e] = masked_sum_seq(op.right_matrix + r * 400 + i * 20, mask, 20);
  }
  __syncthreads();
  const bool gather = op.parent_id_site || op.left_site_id || op.right_site_id;
  const unsigned int cps = (unsigned int)R * 5; /* chunks per site */
  const unsigned long long nchunks = (unsigned long long)op.nsites * cps;
  const unsigned long long stride = (unsigned long long)gridDim.x * AATT_THREADS;
  for (unsigned long long c0 = (unsigned long long)blockIdx.x * AATT_THREADS + threadIdx.x; c0 < nchunks;
       c0 += stride * AATT_U)
  {
    unsigned int lc[AATT_U], rc[AATT_U];
#pragma unroll
    for (int u = 0; u < AATT_U; ++u)
    {
      const unsigned long long c = c0 + u * stride;
      lc[u] = rc[u] = 0;
      if (c < nchunks)
      {
        const unsigned int n = (unsigned int)(c / cps);
        unsigned int lid = n, rid = n;
        if (gather)
        {
          const unsigned int site = op.parent_id_site ? op.parent_id_site[n] : n;
          lid = op.left_site_id ? op.left_site_id[site] : site;
          rid = op.right_site_id ? op.right_site_id[site] : site;
        }
        lc[u] = op.left_tip[lid];
        rc[u] = op.right_tip[rid];
      }
    }
#pragma unroll
    for (int u = 0; u < AATT_U; ++u)
    {
      const unsigned long long c = c0 + u * stride;
      if (c >= nchunks) break;
      const unsigned int n = (unsigned int)(c / cps), k = (unsigned int)(c % cps); /* k = rate * 5 + quad */
      const double * a = tl + (size_t)lc[u] * R * 20 + k * 4;
      const double * b = tr + (size_t)rc[u] * R * 20 + k * 4;
      const double2 a0 = *reinterpret_cast<const double2 *>(a), a1 = *reinterpret_cast<const double2 *>(a + 2);
      const double2 b0 = *reinterpret_cast<const double2 *>(b), b1 = *reinterpret_cast<const double2 *>(b + 2);
      const dbl4 v = {a0.x * b0.x, a0.y * b0.y, a1.x * b1.x, a1.y * b1.y};
      st256(op.parent_clv + c * 4, v);
      if (op.parent_scaler)
      {
        if (per_rate)
        {
          if ((k % 5) == 0) op.parent_scaler[(size_t)n * R + k / 5] = 0;
        }
        else if (k == 0)
          op.parent_scaler[n] = 0;
      }
    }
  }
}

/* ------------------------------------------------------------------------ */

/* returns 1 launched, 0 error, -1 not applicable (caller falls back) */
typedef void (*aas_kernel_t)(const plf_op_t *, int, const plf_state_t *, int);
template <int LOG2R, int NWARPS, int NST>
static aas_kernel_t aas_pick(unsigned int kind)
{
  switch (kind)
  {
    case PLF_OP_II: return k_clv_aa_mma_stream<AK_I, AK_I, LOG2R, NWARPS, NST>;
    case PLF_OP_TI: return k_clv_aa_mma_stream<AK_T, AK_I, LOG2R, NWARPS, NST>;
    case PLF_OP_CI: return k_clv_aa_mma_stream<AK_C, AK_I, LOG2R, NWARPS, NST>;
    case PLF_OP_TC: return k_clv_aa_mma_stream<AK_T, AK_C, LOG2R, NWARPS, NST>;
    default: return k_clv_aa_mma_stream<AK_C, AK_C, LOG2R, NWARPS, NST>;
  }
}

template <int NST>
static aas_kernel_t aas_pick_shape(unsigned int kind, int log2r, int nwarps)
{
  return log2r == 3   ? aas_pick<3, 8, NST>(kind)
         : nwarps == 8 ? (log2r == 0 ? aas_pick<0, 8, NST>(kind) : log2r == 1 ? aas_pick<1, 8, NST>(kind) : aas_pick<2, 8, NST>(kind))
                       : (log2r == 0 ? aas_pick<0, 4, NST>(kind) : log2r == 1 ? aas_pick<1, 4, NST>(kind) : aas_pick<2, 4, NST>(kind));
}

/* dynamic shared memory of the streaming kernel for one op kind: ring + tip table + cherry half tables */
static size_t aas_smem_bytes(unsigned int kind, unsigned int rate_cats, int nwarps, unsigned int maxstates, int nstage = 4)
{
  const int left_inner = (kind == PLF_OP_II), right_inner = (kind == PLF_OP_II || kind == PLF_OP_TI || kind == PLF_OP_CI);
  const int left_tip = (kind == PLF_OP_TI || kind == PLF_OP_TC);
  const int cherries = (kind == PLF_OP_CI || kind == PLF_OP_TC) ? 1 : kind == PLF_OP_CC ? 2 : 0;
  size_t smem = (size_t)nstage * (left_inner + right_inner) * 1280 * nwarps;
  if (left_tip) smem += (size_t)maxstates * rate_cats * AAM_TAB_STRIDE * sizeof(double);
  smem += (size_t)cherries * 2 * maxstates * rate_cats * AAM_TAB_STRIDE * sizeof(double);
  return smem;
}

static int aas_kind_slot(unsigned int kind)
{
  switch (kind)
  {
    case PLF_OP_II: return 0;
    case PLF_OP_TI: return 1;
    case PLF_OP_CI: return 2;
    case PLF_OP_TC: return 3;
    default: return 4;
  }
}

static int launch_aa_stream(plf_ctx * ctx, const plf_op_t * d_ops, unsigned int nops, unsigned int kind,
                            unsigned int rate_cats, int per_rate, unsigned int max_sites,
                            const unsigned long long * d_tipmap, unsigned int maxstates)
{
  int log2r = 0;
  while ((1u << log2r) < rate_cats) ++log2r;
  const int nwarps = (log2r == 3 || ctx->aa_warps8) ? 8 : 4; /* PLF_AA_WARPS=8: 8-warp CTAs for every rate count */
  /* PLF_AA_STAGES=6: a deeper ring for the ops that read CLVs (A/B; the default 4 keeps 4 CTAs per SM) */
  const int nst = (ctx->aa_stages == 6 && (kind == PLF_OP_II || kind == PLF_OP_TI || kind == PLF_OP_CI)) ? 6 : 4;
  aas_kernel_t k = nst == 6 ? aas_pick_shape<6>(kind, log2r, nwarps) : aas_pick_shape<4>(kind, log2r, nwarps);
  const size_t smem = aas_smem_bytes(kind, rate_cats, nwarps, maxstates, nst);
  if (smem > ctx->smem_optin) return -1;
  const int slot = aas_kind_slot(kind);
  if (smem > ctx->aas_smem_set[slot] || ctx->aas_log2r[slot] != log2r)
  {
    PLF_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->aas_smem_set[slot] = smem;
    ctx->aas_log2r[slot] = log2r;
    ctx->aas_occupancy[slot] = 0;
  }
  int & occ = ctx->aas_occupancy[slot];
  if (!occ)
  {
    PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, nwarps * 32, smem));
    if (occ < 1) occ = 1;
  }
  const unsigned int tile = 8 * nwarps / rate_cats;
  const unsigned long long need = ((unsigned long long)max_sites + tile - 1) / tile;
  unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
  if (bx < 1) bx = 1;
  if (bx > need) bx = need;
  dim3 grid((unsigned int)bx, nops);
  k<<<grid, nwarps * 32, smem, ctx->stream>>>(d_ops, per_rate, d_tipmap, (int)maxstates);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* `contiguous`: no op of the group gathers through repeat identifiers */
int plf_launch_aa_mma_group(plf_ctx * ctx, const plf_op_t * d_ops, unsigned int nops, unsigned int kind,
                            unsigned int rate_cats, int per_rate, unsigned int max_sites,
                            const unsigned long long * d_tipmap, unsigned int maxstates, int contiguous)
{
  if (ctx->aa_stream < 0)
  {
    const char * v = getenv("PLF_AA_STREAM");
    ctx->aa_stream = !(v && v[0] == '0');
    v = getenv("PLF_AA_WARPS");
    ctx->aa_warps8 = (v && v[0] == '8');
    v = getenv("PLF_AAM_L2PF");
    ctx->aa_l2pf = !(v && v[0] == '0');
    v = getenv("PLF_AA_STAGES");
    ctx->aa_stages = (v && v[0] == '6') ? 6 : 4;
  }
  if (kind == PLF_OP_TT_VIRTUAL)
  {
    const unsigned long long entries = (unsigned long long)max_sites * (per_rate ? rate_cats : 1u);
    unsigned long long bx = (entries / 4 + 255) / 256;
    const unsigned long long cap = ((unsigned long long)ctx->sm_count * 8 + nops - 1) / nops;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    k_cherry_prepare_aa<<<dim3((unsigned int)bx, nops), 256, 0, ctx->stream>>>(d_ops, per_rate, (int)rate_cats,
                                                                               (int)rate_cats * 400);
    plf_count_launch();
    PLF_CHECK(ctx, cudaGetLastError());
    return 1;
  }
  if (kind != PLF_OP_TT && contiguous && ctx->aa_stream &&
      (rate_cats == 1 || rate_cats == 2 || rate_cats == 4 || rate_cats == 8))
  {
    const int rc = launch_aa_stream(ctx, d_ops, nops, kind, rate_cats, per_rate, max_sites, d_tipmap, maxstates);
    if (rc >= 0) return rc;
  }
  if (kind == PLF_OP_CI || kind == PLF_OP_TC || kind == PLF_OP_CC)
  {
    plf_set_error(ctx, "virtual cherry consumers need the contiguous 20-state streaming path");
    return 0;
  }
  void (*k)(const plf_op_t *, int, int, const plf_state_t *, int);
  size_t smem;
  int threads, slot;
  unsigned long long need;
  if (kind == PLF_OP_TT)
  {
    k = k_clv_aa_tt;
    smem = (size_t)2 * maxstates * rate_cats * 20 * sizeof(double);
    threads = AATT_THREADS;
    slot = 2;
    need = ((unsigned long long)max_sites * rate_cats * 5 + (unsigned long long)AATT_THREADS * AATT_U - 1) /
           ((unsigned long long)AATT_THREADS * AATT_U);
  }
  else
  {
    const int ii = (kind == PLF_OP_II);
    k = ctx->aa_l2pf ? (ii ? k_clv_aa_mma<PLF_OP_II, 1> : k_clv_aa_mma<PLF_OP_TI, 1>)
             : (ii ? k_clv_aa_mma<PLF_OP_II, 0> : k_clv_aa_mma<PLF_OP_TI, 0>);
    smem = (size_t)(ii ? 2 : 1) * rate_cats * AAM_FRAGS * 32 * sizeof(double);
    if (!ii) smem += (size_t)maxstates * rate_cats * AAM_TAB_STRIDE * sizeof(double);
    threads = AAM_THREADS;
    slot = ii ? 0 : 1;
    need = ((unsigned long long)max_sites + 8 * (AAM_THREADS / 32) - 1) / (8 * (AAM_THREADS / 32));
  }
  if (smem > ctx->smem_optin) return -1;
  if (smem > ctx->aam_smem_set[slot])
  {
    PLF_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->aam_smem_set[slot] = smem;
    ctx->aam_occupancy[slot] = 0;
  }
  int & occ = ctx->aam_occupancy[slot];
  if (!occ)
  {
    PLF_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem));
    if (occ < 1) occ = 1;
  }
  unsigned long long bx = ((unsigned long long)ctx->sm_count * occ) / nops;
  if (bx < 1) bx = 1;
  if (bx > need) bx = need;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned int)bx, nops);
  k<<<grid, threads, smem, ctx->stream>>>(d_ops, (int)rate_cats, per_rate, d_tipmap, (int)maxstates);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* do the 20-state streaming kernels serve consumers of virtual cherries for this shape and tip alphabet? */
int plf_aa_virtual_cherries_supported(plf_ctx * ctx, const plf_shape_t * sh, unsigned int maxstates)
{
  if (ctx->aa_stream < 0)
  {
    const char * v = getenv("PLF_AA_STREAM");
    ctx->aa_stream = !(v && v[0] == '0');
    v = getenv("PLF_AA_WARPS");
    ctx->aa_warps8 = (v && v[0] == '8');
    v = getenv("PLF_AAM_L2PF");
    ctx->aa_l2pf = !(v && v[0] == '0');
    v = getenv("PLF_AA_STAGES");
    ctx->aa_stages = (v && v[0] == '6') ? 6 : 4;
  }
  if (sh->states != 20 || !ctx->aa_fast || !ctx->aa_mma || !ctx->aa_stream) return 0;
  if (sh->rate_cats != 1 && sh->rate_cats != 2 && sh->rate_cats != 4 && sh->rate_cats != 8) return 0;
  const int nwarps = (sh->rate_cats == 8 || ctx->aa_warps8) ? 8 : 4;
  /* the widest table set (two cherries) must leave room for at least two CTAs per SM */
  return 2 * (aas_smem_bytes(PLF_OP_CC, sh->rate_cats, nwarps, maxstates) + 2048) <= ctx->smem_optin &&
         2 * (aas_smem_bytes(PLF_OP_CI, sh->rate_cats, nwarps, maxstates) + 2048) <= ctx->smem_optin;
}
