/*
 * plf_partials.cu -- CLV update kernels (sm_100a), one launch per LEVEL of the
 * operation list: blockIdx.y selects the op, blockIdx.x strides over its sites.
 *
 * Replaces pll_core_update_partial_{ii,ti,tt} and the repeats variants
 * (reference src/core_partials.c:82,354,510,612 and their AVX/AVX2 versions)
 * plus the separate scaler pass pll_fill_parent_scaler (src/pll.c:1202), which
 * is fused here.
 *
 * Thread mapping: one thread per (site, rate) block of states_padded doubles
 * when rate_cats is a power of two <= 32 (the R lanes of a site are adjacent
 * lanes of one warp and agree on per-site scaling with shuffles); otherwise one
 * thread per site looping over the rates.  For DNA a (site, rate) block is 32
 * bytes = one 256-bit load/store, a warp touches 1 KB contiguous per access,
 * and the two 4x4 P-matrices of the thread's rate live in registers for the
 * whole kernel.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ *
 *  DNA (4 states)                                                           *
 * ------------------------------------------------------------------------ */

struct DnaItem
{
  dbl4 l, r;
  unsigned int n, lid, rid, lcode, rcode;
  bool active;
};

template <int ONE_RATE>
__global__ void __launch_bounds__(256, 2)
k_partials_dna(const plf_op_t * __restrict__ ops, int R, int per_rate)
{
  extern __shared__ double smem[];
  const plf_op_t op = ops[blockIdx.y];
  const int kind = (int)op.kind;
  const unsigned int nsites = op.nsites;

  /* per-call tip lookup: tab[code][rate][i] = masked pairwise sum of row i
   * (core_partials_avx.c:1336-1395 for ti, :295-351 for the two tt halves) */
  double * tl = smem;
  double * tr = smem + 64 * R;
  if (kind != PLF_OP_II)
  {
    for (int e = threadIdx.x; e < 64 * R; e += blockDim.x)
    {
      const int code = e / (4 * R), r = (e >> 2) % R, i = e & 3;
      tl[e] = masked_sum4(op.left_matrix + r * 16 + i * 4, code);
      if (kind == PLF_OP_TT)
        tr[e] = masked_sum4(op.right_matrix + r * 16 + i * 4, code);
    }
    __syncthreads();
  }

  const int L = ONE_RATE ? R : 1;
  const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int nthreads = gridDim.x * blockDim.x;
  const unsigned int lane_in_site = tid & (L - 1);
  const unsigned int sites_per_iter = nthreads / L;
  const unsigned int my_site0 = tid / L;
  const unsigned int warp_site0 = (tid & ~31u) / L;

  double Lm[16], Rm[16];
  if (ONE_RATE)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
      Lm[i] = (kind == PLF_OP_II) ? op.left_matrix[lane_in_site * 16 + i] : 0.0;
      Rm[i] = (kind != PLF_OP_TT) ? op.right_matrix[lane_in_site * 16 + i] : 0.0;
    }
  }

  for (unsigned int s0 = warp_site0; s0 < nsites; s0 += sites_per_iter)
  {
    const unsigned int n = s0 + (my_site0 - warp_site0);
    const bool active = n < nsites;
    unsigned int site = n, lid = n, rid = n, lcode = 0, rcode = 0;
    if (active)
    {
      if (op.parent_id_site) site = op.parent_id_site[n];
      lid = op.left_site_id ? op.left_site_id[site] : site;
      rid = op.right_site_id ? op.right_site_id[site] : site;
      if (kind != PLF_OP_II) lcode = op.left_tip[site];
      if (kind == PLF_OP_TT) rcode = op.right_tip[site];
    }

    if (ONE_RATE)
    {
      const unsigned int rate = lane_in_site;
      dbl4 l = {0, 0, 0, 0}, r = {0, 0, 0, 0};
      if (active)
      {
        if (kind == PLF_OP_II) l = ld256_stream(op.left_clv + ((size_t)lid * R + rate) * 4);
        if (kind != PLF_OP_TT) r = ld256_stream(op.right_clv + ((size_t)rid * R + rate) * 4);
      }
      double a[4], b[4];
      if (kind == PLF_OP_II)
      {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = dot4_pairwise(Lm + 4 * i, l);
      }
      else
      {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = tl[(lcode * R + rate) * 4 + i];
      }
      if (kind != PLF_OP_TT)
      {
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = dot4_pairwise(Rm + 4 * i, r);
      }
      else
      {
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = tr[(rcode * R + rate) * 4 + i];
      }
      dbl4 v = {a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3]};

      if (op.parent_scaler)
      {
        /* tip-tip never scales and zeroes the scaler (core_partials_avx.c:1007) */
        int below = (kind != PLF_OP_TT) && (v.x < PLF_SCALE_THRESHOLD) && (v.y < PLF_SCALE_THRESHOLD) &&
                    (v.z < PLF_SCALE_THRESHOLD) && (v.w < PLF_SCALE_THRESHOLD);
        if (per_rate)
        {
          if (active)
          {
            unsigned int sc = 0;
            if (kind == PLF_OP_II && op.left_scaler) sc += op.left_scaler[(size_t)lid * R + rate];
            if (kind != PLF_OP_TT && op.right_scaler) sc += op.right_scaler[(size_t)rid * R + rate];
            if (below)
            {
              v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR;
              v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
              sc += 1;
            }
            op.parent_scaler[(size_t)n * R + rate] = sc;
          }
        }
        else
        {
          const int all = group_and(below, L); /* every lane of the warp takes part */
          if (all)
          {
            v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR;
            v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
          }
          if (active && lane_in_site == 0)
          {
            unsigned int sc = all ? 1u : 0u;
            if (kind == PLF_OP_II && op.left_scaler) sc += op.left_scaler[lid];
            if (kind != PLF_OP_TT && op.right_scaler) sc += op.right_scaler[rid];
            op.parent_scaler[n] = sc;
          }
        }
      }
      if (active) st256(op.parent_clv + ((size_t)n * R + rate) * 4, v);
    }
    else
    {
      /* rate_cats not a power of two: one thread per site, rates in a loop,
       * matrices through L1; outputs are stored unscaled and rescaled in the
       * (rare) case the whole site falls below the threshold */
      if (!active) continue;
      int site_below = 1;
      for (int rate = 0; rate < R; ++rate)
      {
        dbl4 l = {0, 0, 0, 0}, r = {0, 0, 0, 0};
        if (kind == PLF_OP_II) l = ld256(op.left_clv + ((size_t)lid * R + rate) * 4);
        if (kind != PLF_OP_TT) r = ld256(op.right_clv + ((size_t)rid * R + rate) * 4);
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
          a[i] = (kind == PLF_OP_II) ? dot4_pairwise(op.left_matrix + rate * 16 + 4 * i, l)
                                     : tl[(lcode * R + rate) * 4 + i];
          b[i] = (kind != PLF_OP_TT) ? dot4_pairwise(op.right_matrix + rate * 16 + 4 * i, r)
                                     : tr[(rcode * R + rate) * 4 + i];
        }
        dbl4 v = {a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3]};
        const int below = (kind != PLF_OP_TT) && (v.x < PLF_SCALE_THRESHOLD) && (v.y < PLF_SCALE_THRESHOLD) &&
                          (v.z < PLF_SCALE_THRESHOLD) && (v.w < PLF_SCALE_THRESHOLD);
        if (op.parent_scaler && per_rate)
        {
          unsigned int sc = 0;
          if (kind == PLF_OP_II && op.left_scaler) sc += op.left_scaler[(size_t)lid * R + rate];
          if (kind != PLF_OP_TT && op.right_scaler) sc += op.right_scaler[(size_t)rid * R + rate];
          if (below)
          {
            v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR;
            v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
            sc += 1;
          }
          op.parent_scaler[(size_t)n * R + rate] = sc;
        }
        site_below &= below;
        st256(op.parent_clv + ((size_t)n * R + rate) * 4, v);
      }
      if (op.parent_scaler && !per_rate)
      {
        unsigned int sc = 0;
        if (kind == PLF_OP_II && op.left_scaler) sc += op.left_scaler[lid];
        if (kind != PLF_OP_TT && op.right_scaler) sc += op.right_scaler[rid];
        if (site_below)
        {
          for (int rate = 0; rate < R; ++rate)
          {
            double * p = op.parent_clv + ((size_t)n * R + rate) * 4;
            dbl4 v = ld256(p);
            v.x *= PLF_SCALE_FACTOR; v.y *= PLF_SCALE_FACTOR;
            v.z *= PLF_SCALE_FACTOR; v.w *= PLF_SCALE_FACTOR;
            st256(p, v);
          }
          sc += 1;
        }
        op.parent_scaler[n] = sc;
      }
    }
  }
}

/* ------------------------------------------------------------------------ *
 *  Other state counts.  ST = 20: matrices and tip tables staged in shared    *
 *  memory, child vectors in registers.  ST = 0: run-time state count,        *
 *  matrices through L1, and -- as the reference's generic AVX2 kernels do    *
 *  (core_partials_avx2.c:1101-1227) -- all states_padded rows are computed,  *
 *  the padded ones running into the following matrix; they land in the       *
 *  padded CLV lanes and take part in the scaling test.                       *
 * ------------------------------------------------------------------------ */

template <int ST>
__global__ void __launch_bounds__(128)
k_partials_gen(const plf_op_t * __restrict__ ops, int R, int per_rate, int st_rt, int sp_rt,
               const plf_state_t * __restrict__ tipmap, int maxstates, int L)
{
  extern __shared__ double smem[];
  const plf_op_t op = ops[blockIdx.y];
  const int kind = (int)op.kind;
  const unsigned int nsites = op.nsites;
  const int st = ST ? ST : st_rt;
  const int sp = ST ? ((ST + 3) & ~3) : sp_rt;
  const int msz = st * sp; /* one rate's matrix */

  /* ST=20 staging: [Lmat R*400][Rmat R*400][tl maxstates*R*20][tr ...] */
  const double * lmat = op.left_matrix;
  const double * rmat = op.right_matrix;
  double * tl = nullptr;
  double * tr = nullptr;
  if (ST == 20)
  {
    double * sl = smem;
    double * sr = smem + R * msz;
    for (int e = threadIdx.x; e < R * msz; e += blockDim.x)
    {
      sl[e] = op.left_matrix[e];
      sr[e] = op.right_matrix[e];
    }
    lmat = sl;
    rmat = sr;
    if (kind != PLF_OP_II)
    {
      /* scalar sums in increasing column order (core_partials_avx2.c:387-456,
       * core_partials_avx.c:160-185) */
      tl = smem + 2 * R * msz;
      tr = tl + maxstates * R * sp;
      for (int e = threadIdx.x; e < maxstates * R * sp; e += blockDim.x)
      {
        const int c = e / (R * sp), r = (e / sp) % R, i = e % sp;
        const plf_state_t mask = tipmap[c];
        tl[e] = masked_sum_seq(op.left_matrix + r * msz + i * sp, mask, st);
        if (kind == PLF_OP_TT) tr[e] = masked_sum_seq(op.right_matrix + r * msz + i * sp, mask, st);
      }
    }
    __syncthreads();
  }

  const int RT = R / L;
  const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int nthreads = gridDim.x * blockDim.x;
  const unsigned int lane_in_site = tid & (L - 1);
  const unsigned int sites_per_iter = nthreads / L;
  const unsigned int my_site0 = tid / L;
  const unsigned int warp_site0 = (tid & ~31u) / L;
  const size_t span = (size_t)sp * R;

  for (unsigned int s0 = warp_site0; s0 < nsites; s0 += sites_per_iter)
  {
    const unsigned int n = s0 + (my_site0 - warp_site0);
    const bool active = n < nsites;
    unsigned int site = n, lid = n, rid = n, lcode = 0, rcode = 0;
    plf_state_t lmask = 0, rmask = 0;
    if (active)
    {
      if (op.parent_id_site) site = op.parent_id_site[n];
      lid = op.left_site_id ? op.left_site_id[site] : site;
      rid = op.right_site_id ? op.right_site_id[site] : site;
      if (kind != PLF_OP_II) { lcode = op.left_tip[site]; lmask = tipmap[lcode]; }
      if (kind == PLF_OP_TT) { rcode = op.right_tip[site]; rmask = tipmap[rcode]; }
    }
    int my_below = 1;
    if (active)
    {
      for (int rr = 0; rr < RT; ++rr)
      {
        const int rate = lane_in_site * RT + rr;
        double * pout = op.parent_clv + (size_t)n * span + (size_t)rate * sp;
        const double * lc = (kind == PLF_OP_II) ? op.left_clv + (size_t)lid * span + (size_t)rate * sp : nullptr;
        const double * rc = (kind != PLF_OP_TT) ? op.right_clv + (size_t)rid * span + (size_t)rate * sp : nullptr;
        int below = 1;
        if (ST == 20)
        {
          double l[20], r[20];
          if (lc)
          {
#pragma unroll
            for (int j = 0; j < 20; j += 4)
            {
              dbl4 t = ld256_stream(lc + j);
              l[j] = t.x; l[j + 1] = t.y; l[j + 2] = t.z; l[j + 3] = t.w;
            }
          }
          if (rc)
          {
#pragma unroll
            for (int j = 0; j < 20; j += 4)
            {
              dbl4 t = ld256_stream(rc + j);
              r[j] = t.x; r[j + 1] = t.y; r[j + 2] = t.z; r[j + 3] = t.w;
            }
          }
#pragma unroll 1
          for (int i = 0; i < 20; i += 4)
          {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
              const double * lrow = lmat + rate * 400 + (i + q) * 20;
              const double * rrow = rmat + rate * 400 + (i + q) * 20;
              const double a = (kind == PLF_OP_II) ? dot_lanes_fma(lrow, l, 20)
                                                   : tl[((size_t)lcode * R + rate) * 20 + i + q];
              const double b = (kind != PLF_OP_TT) ? dot_lanes_fma(rrow, r, 20)
                                                   : tr[((size_t)rcode * R + rate) * 20 + i + q];
              v[q] = a * b;
              below &= (v[q] < PLF_SCALE_THRESHOLD);
            }
            dbl4 o = {v[0], v[1], v[2], v[3]};
            st256(pout + i, o);
          }
        }
        else
        {
          for (int i = 0; i < sp; ++i)
          {
            const double * lrow = lmat + (size_t)rate * msz + (size_t)i * sp;
            const double * rrow = rmat + (size_t)rate * msz + (size_t)i * sp;
            double v;
            if (kind == PLF_OP_TT)
            {
              /* generic tip-tip table: real rows only, padded lanes zero
               * (core_partials_avx.c:63-101) */
              v = (i < st) ? masked_sum_seq(lrow, lmask, st) * masked_sum_seq(rrow, rmask, st) : 0.0;
            }
            else
            {
              const double a = (kind == PLF_OP_II) ? dot_lanes_fma(lrow, lc, sp) : masked_sum_lanes(lrow, lmask, sp);
              const double b = dot_lanes_fma(rrow, rc, sp);
              v = a * b;
            }
            below &= (v < PLF_SCALE_THRESHOLD);
            pout[i] = v;
          }
        }
        if (kind == PLF_OP_TT) below = 0;
        if (op.parent_scaler && per_rate)
        {
          unsigned int sc = 0;
          if (kind == PLF_OP_II && op.left_scaler) sc += op.left_scaler[(size_t)lid * R + rate];
          if (kind != PLF_OP_TT && op.right_scaler) sc += op.right_scaler[(size_t)rid * R + rate];
          if (below)
          {
            for (int i = 0; i < sp; ++i) pout[i] *= PLF_SCALE_FACTOR;
            sc += 1;
          }
          op.parent_scaler[(size_t)n * R + rate] = sc;
        }
        my_below &= below;
      }
    }
    if (op.parent_scaler && !per_rate)
    {
      const int all = group_and(my_below, L);
      if (active)
      {
        if (all)
          for (int rr = 0; rr < RT; ++rr)
          {
            double * pout = op.parent_clv + (size_t)n * span + (size_t)(lane_in_site * RT + rr) * sp;
            for (int i = 0; i < sp; ++i) pout[i] *= PLF_SCALE_FACTOR;
          }
        if (lane_in_site == 0)
        {
          unsigned int sc = all ? 1u : 0u;
          if (kind == PLF_OP_II && op.left_scaler) sc += op.left_scaler[lid];
          if (kind != PLF_OP_TT && op.right_scaler) sc += op.right_scaler[rid];
          op.parent_scaler[n] = sc;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------ */

/* one kernel launch serves a run of same-kind ops of a level, the op selected by blockIdx.y: a run ends at
 * the first op of another kind (or, under site repeats, at the first op that gathers / does not gather
 * through identifiers like the first one), at the end of the level `b`, or after PLF_MAX_RUN_OPS ops
 * (gridDim.y limit) */
extern "C" unsigned int plf_run_end(const plf_op_t * h_ops, unsigned int i, unsigned int b,
                                    unsigned int * max_sites, int * contiguous)
{
  unsigned int j = i;
  const bool gathers = h_ops[i].parent_id_site || h_ops[i].left_site_id || h_ops[i].right_site_id;
  while (j < b && h_ops[j].kind == h_ops[i].kind && j - i < PLF_MAX_RUN_OPS &&
         (bool)(h_ops[j].parent_id_site || h_ops[j].left_site_id || h_ops[j].right_site_id) == gathers)
  {
    if (max_sites && h_ops[j].nsites > *max_sites) *max_sites = h_ops[j].nsites;
    if (contiguous && (h_ops[j].parent_id_site || h_ops[j].left_site_id || h_ops[j].right_site_id)) *contiguous = 0;
    ++j;
  }
  return j;
}

/* do the kernels that consume virtual cherries serve this shape (and, for 20 states, this many tip codes)?
 * The host layer asks before it leaves a tip-tip parent unwritten. */
extern "C" int plf_virtual_cherries_supported(plf_ctx_t * ctx, const plf_shape_t * sh, unsigned int maxstates)
{
  const char * v = getenv("PLF_VIRTUAL_CHERRIES");
  if (v && v[0] == '0') return 0;
  if (sh->states == 4) return plf_dna_virtual_cherries_supported(ctx, sh);
  if (sh->states == 20) return plf_aa_virtual_cherries_supported(ctx, sh, maxstates);
  return 0;
}

static int is_pow2(unsigned int x) { return x && !(x & (x - 1)); }

/* queues the launches of a level-sorted op list.  `upload` = 0: the op descriptors and tile-prefix arrays
 * are already in the workspace (identical call being captured into a CUDA graph): no host-to-device copy */
static int enqueue_levels(plf_ctx_t * ctx, const plf_shape_t * sh, const plf_op_t * h_ops, unsigned int nops,
                          const unsigned int * h_level_start, unsigned int nlevels,
                          const unsigned long long * d_tipmap, unsigned int maxstates, int upload,
                          plf_ws * ws = nullptr)
{
  if (!ws) ws = &ctx->ws_ops;
  /* workspace: the op descriptors, then one tile-prefix array per (level, kind) run of gathering
   * 4-state inner-inner ops (site repeats; see k_clv_dna_ii_balanced) */
  const size_t prefix_entries = 2 * (size_t)nops + 1; /* one array of (ops + 1) entries per run; runs <= ops */
  const size_t ops_bytes = (size_t)nops * sizeof(plf_op_t);
  plf_op_t * d_ops = (plf_op_t *)plf_ws_reserve(ctx, ws, ops_bytes + prefix_entries * sizeof(unsigned int));
  if (!d_ops) return 0;
  unsigned int * d_prefix = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned char *>(d_ops) + ops_bytes);
  if (upload) PLF_CHECK(ctx, cudaMemcpyAsync(d_ops, h_ops, ops_bytes, cudaMemcpyHostToDevice, ctx->stream));
  unsigned int * h_prefix = nullptr;
  size_t prefix_used = 0;
  const int R = (int)sh->rate_cats;
  const int one_rate = is_pow2(sh->rate_cats) && sh->rate_cats <= 32;
  if (ctx->dna_level_max_sites < 0)
  {
    const char * v = getenv("PLF_LEVEL_MAX_SITES");
    ctx->dna_level_max_sites = (v && v[0]) ? atoi(v) : 2048; /* measured crossover ~2500 sites, profiles/r2_notes.md */
  }
  const unsigned int level_max_sites = (unsigned int)ctx->dna_level_max_sites;
  const int L = one_rate ? R : 1;
  if (ctx->dna_flow < 0)
  {
    const char * v = getenv("PLF_FLOW");
    ctx->dna_flow = !(v && v[0] == '0');
    /* 100 taxa: equal to the ring kernels near 80k sites = 7.8M site-updates (profiles/r2_notes.md) */
    v = getenv("PLF_FLOW_MAX_SITES");
    ctx->dna_flow_max_sites = (v && v[0]) ? atoi(v) : 65536;
    v = getenv("PLF_FLOW_MAX_UPDATES");
    ctx->dna_flow_max_updates = (v && v[0]) ? strtoull(v, nullptr, 10) : 6500000ull;
    v = getenv("PLF_FLOW_PATH_MAX");
    ctx->dna_flow_path_max = (v && atoi(v) >= 1 && atoi(v) <= PLF_FLOW_PATH_MAX) ? atoi(v) : PLF_FLOW_PATH_MAX;
  }
  if (ctx->dna_flow && ws == &ctx->ws_ops && nlevels > 1 && nops > 1 && sh->states == 4 && one_rate &&
      sh->rate_cats <= 4 && h_ops[0].dep[0] != PLF_DEP_ORDERED && h_ops[0].nsites <= (unsigned int)ctx->dna_flow_max_sites &&
      (unsigned long long)nops * h_ops[0].nsites <= ctx->dna_flow_max_updates)
  {
    /* narrow alignment, plain list (no buffer recycled, no virtual cherries, no repeat identifiers): the whole
     * traversal as one launch of k_clv_dna_flow */
    const unsigned int sites = h_ops[0].nsites;
    int plain = sites > 0;
    for (unsigned int i = 0; i < nops && plain; ++i)
    {
      const plf_op_t & o = h_ops[i];
      plain = o.nsites == sites && (o.kind == PLF_OP_II || o.kind == PLF_OP_TI || o.kind == PLF_OP_TT) &&
              o.dep[0] != PLF_DEP_ORDERED && !(o.parent_id_site || o.left_site_id || o.right_site_id);
    }
    const unsigned long long flags = (unsigned long long)nops * plf_dna_flow_chunks(sh->rate_cats, sites);
    if (plain && flags < (1ull << 30))
    {
      const size_t plan_at = 64 + (size_t)flags * sizeof(unsigned long long);
      const size_t start_at = plan_at + (size_t)nops * sizeof(plf_flow_op);
      const size_t bytes = start_at + ((size_t)nops + 1) * sizeof(unsigned int);
      unsigned char * h_plan = (unsigned char *)malloc(bytes - plan_at);
      const unsigned int npaths =
          h_plan ? plf_dna_flow_plan(h_ops, nops, (unsigned int)ctx->dna_flow_path_max, (plf_flow_op *)h_plan,
                                     (unsigned int *)(h_plan + (start_at - plan_at)))
                 : 0;
      if (npaths)
      {
        unsigned char * flow = (unsigned char *)plf_ws_reserve(ctx, &ctx->ws_flow, bytes);
        cudaError_t e = flow ? cudaSuccess : cudaErrorMemoryAllocation;
        if (flow && (ctx->ws_flow_zeroed != flow || ctx->ws_flow_zeroed_bytes != ctx->ws_flow.bytes))
        {
          /* a new allocation (it may sit at the old address): epoch 0, no flag set.  Never inside a capture: a
           * captured list ran once before with the same sizes */
          e = cudaMemsetAsync(flow, 0, ctx->ws_flow.bytes, ctx->stream);
          ctx->ws_flow_zeroed = flow;
          ctx->ws_flow_zeroed_bytes = ctx->ws_flow.bytes;
        }
        if (e == cudaSuccess && upload)
          e = cudaMemcpyAsync(flow + plan_at, h_plan, bytes - plan_at, cudaMemcpyHostToDevice, ctx->stream);
        /* (pageable source: staged by the driver before the call returns, like the descriptors above) */
        free(h_plan);
        if (!flow) return 0;
        PLF_CHECK(ctx, e);
        return plf_launch_dna_flow(ctx, (const plf_flow_op *)(flow + plan_at), (const unsigned int *)(flow + start_at),
                                   npaths, sh->rate_cats, sh->per_rate_scalers, sites, flow);
      }
      free(h_plan);
    }
  }

  for (unsigned int lv = 0; lv < nlevels; ++lv)
  {
    const unsigned int a = h_level_start[lv], b = h_level_start[lv + 1];
    if (b <= a) continue;
    unsigned int max_sites = 0;
    int any_tip = 0;
    for (unsigned int i = a; i < b; ++i)
    {
      if (h_ops[i].nsites > max_sites) max_sites = h_ops[i].nsites;
      any_tip |= (h_ops[i].kind != PLF_OP_II);
    }
    if (!max_sites) continue;
    if (sh->states == 4 && one_rate && sh->rate_cats <= 4 && max_sites <= level_max_sites)
    {
      /* narrow alignment: the whole level in one launch (k_clv_dna_level), unless an op gathers through
       * repeat identifiers */
      int contiguous = 1;
      for (unsigned int i = a; i < b && contiguous; ++i)
        contiguous = !(h_ops[i].parent_id_site || h_ops[i].left_site_id || h_ops[i].right_site_id);
      if (contiguous)
      {
        for (unsigned int c0 = a; c0 < b; c0 += PLF_MAX_RUN_OPS)
          if (!plf_launch_dna_level(ctx, d_ops + c0, (b - c0 < PLF_MAX_RUN_OPS) ? b - c0 : PLF_MAX_RUN_OPS, sh->rate_cats,
                                    sh->per_rate_scalers, max_sites))
          {
            free(h_prefix);
            return 0;
          }
        continue;
      }
    }
    if (sh->states == 4 && one_rate)
    {
      /* specialised kernels per op kind (plf_partials_dna.cu): one launch per
       * run of same-kind ops; the host layer sorts a level by kind */
      for (unsigned int i = a; i < b;)
      {
        unsigned int run_sites = 0;
        int contiguous = 1;
        const unsigned int j = plf_run_end(h_ops, i, b, &run_sites, &contiguous);
        const unsigned int * d_run_prefix = nullptr;
        unsigned int total_tiles = 0;
        if (!contiguous && h_ops[i].kind == PLF_OP_II) /* single ops too: the tile-walk kernel is the faster gather */
        {
          if (!h_prefix) h_prefix = (unsigned int *)malloc(prefix_entries * sizeof(unsigned int));
          if (h_prefix && prefix_used + (j - i) + 1 <= prefix_entries)
          {
            unsigned int * pre = h_prefix + prefix_used;
            for (unsigned int k = i; k < j; ++k)
            {
              pre[k - i] = total_tiles;
              total_tiles += plf_dna_balanced_tiles(h_ops[k].nsites, sh->rate_cats);
            }
            pre[j - i] = total_tiles;
            if (!upload || cudaMemcpyAsync(d_prefix + prefix_used, pre, (size_t)(j - i + 1) * sizeof(unsigned int),
                                           cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess)
              d_run_prefix = d_prefix + prefix_used;
            prefix_used += (j - i) + 1;
          }
        }
        int pair_lists = !contiguous;
        for (unsigned int k = i; k < j && pair_lists; ++k) pair_lists = h_ops[k].pair_list != nullptr;
        if (run_sites && !plf_launch_dna_group(ctx, d_ops + i, j - i, h_ops[i].kind, sh->rate_cats,
                                               sh->per_rate_scalers, run_sites, contiguous, d_run_prefix, total_tiles,
                                               pair_lists))
        {
          free(h_prefix);
          return 0;
        }
        i = j;
      }
      continue;
    }
    if (sh->states == 4)
    {
      const int threads = 256;
      for (unsigned int c0 = a; c0 < b; c0 += PLF_MAX_RUN_OPS)
      {
        const unsigned int nrun = (b - c0 < PLF_MAX_RUN_OPS) ? b - c0 : PLF_MAX_RUN_OPS;
        unsigned int bx = (unsigned int)(((unsigned long long)max_sites + threads - 1) / threads);
        unsigned int cap = (unsigned int)(ctx->sm_count * 2 * 4) / nrun;
        if (cap < 1) cap = 1;
        if (bx > cap) bx = cap;
        dim3 grid(bx, nrun);
        const size_t smem = any_tip ? (size_t)2 * 64 * R * sizeof(double) : 0;
        k_partials_dna<0><<<grid, threads, smem, ctx->stream>>>(d_ops + c0, R, sh->per_rate_scalers);
        if (c0 + PLF_MAX_RUN_OPS < b) /* the last chunk is counted and checked below */
        {
          plf_count_launch();
          PLF_CHECK(ctx, cudaGetLastError());
        }
      }
    }
    else
    {
      /* runs of same-kind ops: 20-state ii / ti runs go to the register-tiled
       * kernels of plf_partials_aa.cu, everything else to the generic kernel */
      for (unsigned int i = a; i < b;)
      {
        unsigned int run_sites = 0;
        int run_tip = (h_ops[i].kind != PLF_OP_II);
        int contiguous = 1;
        const unsigned int j = plf_run_end(h_ops, i, b, &run_sites, &contiguous);
        int done = 0;
        if (run_sites && sh->states == 20 && ctx->aa_fast)
        {
          /* tensor-core (DMMA) kernels by default, tip-tip always on its own write-only kernel;
           * PLF_AA_MMA=0 keeps inner-inner / tip-inner on the bit-exact DFMA kernels */
          int rc = -1;
          if (ctx->aa_mma || h_ops[i].kind == PLF_OP_TT)
            rc = plf_launch_aa_mma_group(ctx, d_ops + i, j - i, h_ops[i].kind, sh->rate_cats, sh->per_rate_scalers,
                                         run_sites, d_tipmap, maxstates, contiguous);
          if (rc == -1 && h_ops[i].kind != PLF_OP_TT)
            rc = plf_launch_aa_group(ctx, d_ops + i, j - i, h_ops[i].kind, sh->rate_cats, sh->per_rate_scalers,
                                     run_sites, d_tipmap, maxstates);
          if (rc == 0) return 0;
          done = (rc == 1);
        }
        if (run_sites && !done)
        {
          const int threads = 128;
          const unsigned long long work = (unsigned long long)run_sites * L;
          unsigned int bx = (unsigned int)((work + threads - 1) / threads);
          unsigned int cap = (unsigned int)(ctx->sm_count * 4 * 2) / (j - i);
          if (cap < 1) cap = 1;
          if (bx > cap) bx = cap;
          dim3 grid(bx, j - i);
          if (sh->states == 20)
          {
            size_t smem = (size_t)2 * R * 400 * sizeof(double);
            if (run_tip) smem += (size_t)2 * maxstates * R * 20 * sizeof(double);
            if (smem > ctx->smem_optin)
            {
              plf_set_error(ctx, "protein tip tables need %zu B of shared memory (> %zu)", smem, ctx->smem_optin);
              return 0;
            }
            if (smem > ctx->gen20_smem_set)
            {
              PLF_CHECK(ctx, cudaFuncSetAttribute(k_partials_gen<20>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)smem));
              ctx->gen20_smem_set = smem;
            }
            k_partials_gen<20><<<grid, threads, smem, ctx->stream>>>(d_ops + i, R, sh->per_rate_scalers, 20, 20,
                                                                      d_tipmap, (int)maxstates, L);
          }
          else
            k_partials_gen<0><<<grid, threads, 0, ctx->stream>>>(d_ops + i, R, sh->per_rate_scalers,
                                                                 (int)sh->states, (int)sh->states_padded, d_tipmap,
                                                                 (int)maxstates, L);
          plf_count_launch();
          PLF_CHECK(ctx, cudaGetLastError());
        }
        i = j;
      }
      continue;
    }
    plf_count_launch();
    PLF_CHECK(ctx, cudaGetLastError());
  }
  free(h_prefix);
  return 1;
}

/* ------------------------------------------------------------------------ *
 *  One CUDA graph per traversal.  Clients evaluate the same operation list   *
 *  over and over (branch-length and model optimisation change P-matrices,    *
 *  not the list), and on narrow alignments a traversal is bound by the       *
 *  launch path: 12 launches of ~5 us kernels.  The resolved descriptors      *
 *  (device pointers included) of the last call are kept; the second          *
 *  identical call is captured, later ones replay the graph.                  *
 * ------------------------------------------------------------------------ */
static void graph_cache_drop(plf_ctx * ctx)
{
  if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
  ctx->graph_exec = nullptr;
  ctx->graph_valid = 0;
}

void plf_graph_cache_destroy(plf_ctx * ctx)
{
  graph_cache_drop(ctx);
  free(ctx->graph_ops);
  free(ctx->graph_levels);
  ctx->graph_ops = nullptr;
  ctx->graph_levels = nullptr;
}

extern "C" int plf_update_partials_once(plf_ctx_t * ctx, const plf_shape_t * sh, const plf_op_t * h_ops,
                                        unsigned int nops, const unsigned long long * d_tipmap,
                                        unsigned int maxstates)
{
  if (!nops) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  /* its own descriptor workspace: the cached graph of the last traversal keeps reading ws_ops */
  const unsigned int level_start[2] = {0, nops};
  return enqueue_levels(ctx, sh, h_ops, nops, level_start, 1, d_tipmap, maxstates, 1, &ctx->ws_once);
}

extern "C" int plf_update_partials(plf_ctx_t * ctx, const plf_shape_t * sh, const plf_op_t * h_ops,
                                   unsigned int nops, const unsigned int * h_level_start,
                                   unsigned int nlevels, const unsigned long long * d_tipmap,
                                   unsigned int maxstates)
{
  if (!nops) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (ctx->graph_mode < 0)
  {
    const char * v = getenv("PLF_GRAPH");
    ctx->graph_mode = !(v && v[0] == '0');
  }
  if (!ctx->graph_mode) return enqueue_levels(ctx, sh, h_ops, nops, h_level_start, nlevels, d_tipmap, maxstates, 1);

  const int same = ctx->graph_valid && ctx->graph_nops == nops && ctx->graph_nlevels == nlevels &&
                   ctx->graph_tipmap == d_tipmap && ctx->graph_maxstates == maxstates &&
                   !memcmp(&ctx->graph_shape, sh, sizeof(*sh)) &&
                   !memcmp(ctx->graph_ops, h_ops, (size_t)nops * sizeof(plf_op_t)) &&
                   !memcmp(ctx->graph_levels, h_level_start, ((size_t)nlevels + 1) * sizeof(unsigned int));
  if (same && ctx->graph_exec)
  {
    PLF_CHECK(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
    plf_count_launches(ctx->graph_launches);
    return 1;
  }
  if (same)
  {
    /* second identical call: capture it (descriptors are resident, nothing is uploaded) */
    cudaGraph_t graph = nullptr;
    const unsigned long long before = plf_kernel_launches();
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
    {
      const int ok = enqueue_levels(ctx, sh, h_ops, nops, h_level_start, nlevels, d_tipmap, maxstates, 0);
      const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
      const unsigned long long captured = plf_kernel_launches() - before;
      if (ok && e == cudaSuccess && graph &&
          cudaGraphInstantiate(&ctx->graph_exec, graph, 0) == cudaSuccess)
      {
        cudaGraphDestroy(graph);
        ctx->graph_launches = captured;
        PLF_CHECK(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
        return 1; /* the launches were counted while they were captured */
      }
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      ctx->graph_exec = nullptr;
      plf_count_launches(0ull - captured); /* nothing ran */
    }
    ctx->graph_mode = 0; /* capture is not available here: plain launches from now on */
    return enqueue_levels(ctx, sh, h_ops, nops, h_level_start, nlevels, d_tipmap, maxstates, 1);
  }
  /* a new list: run it and remember it */
  graph_cache_drop(ctx);
  if (ctx->graph_cap_ops < nops)
  {
    free(ctx->graph_ops);
    ctx->graph_ops = (plf_op_t *)malloc((size_t)nops * sizeof(plf_op_t));
    ctx->graph_cap_ops = ctx->graph_ops ? nops : 0;
  }
  if (ctx->graph_cap_levels < nlevels + 1)
  {
    free(ctx->graph_levels);
    ctx->graph_levels = (unsigned int *)malloc(((size_t)nlevels + 1) * sizeof(unsigned int));
    ctx->graph_cap_levels = ctx->graph_levels ? nlevels + 1 : 0;
  }
  if (!enqueue_levels(ctx, sh, h_ops, nops, h_level_start, nlevels, d_tipmap, maxstates, 1)) return 0;
  if (ctx->graph_ops && ctx->graph_levels)
  {
    memcpy(ctx->graph_ops, h_ops, (size_t)nops * sizeof(plf_op_t));
    memcpy(ctx->graph_levels, h_level_start, ((size_t)nlevels + 1) * sizeof(unsigned int));
    ctx->graph_nops = nops;
    ctx->graph_nlevels = nlevels;
    ctx->graph_shape = *sh;
    ctx->graph_tipmap = d_tipmap;
    ctx->graph_maxstates = maxstates;
    ctx->graph_valid = 1;
  }
  return 1;
}
