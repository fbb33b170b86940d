/*
 * pll_host.c -- the C host layer of the B200-native likelihood engine.
 *
 * Implements the partition API of libpll-2 (reference src/pll.h:636-849,
 * front ends in src/pll.c, models.c, partials.c, likelihood.c, derivatives.c,
 * repeats.c) for partitions created with PLL_ATTRIB_ARCH_CUDA.  This file owns
 * the host-side state (the public pll_partition_t plus a private tail), decides
 * WHAT to run -- which kernel variant, on which buffers, in which launch level
 * -- and calls the sm_100a kernels through the thin C-ABI of plf_backend.h.
 * There is no arithmetic on CLVs here and no CPU fallback: without a CUDA
 * device pll_partition_create fails.
 *
 * Memory model
 *   device-canonical : clv[], scale_buffer[], pmatrix[] (the pointer fields
 *                      hold HBM addresses), tipchars mirrors, repeat-id arrays,
 *                      sumtables (keyed by the caller's host pointer)
 *   host-canonical   : rates, weights, frequencies, subst_params, eigen*,
 *                      prop_invar, pattern_weights, tipchars[], charmap, tipmap
 *                      and every field of pll_repeats_t.  Small model arrays are
 *                      packed per call and re-uploaded only when their bytes
 *                      changed, so direct writes to the struct fields are seen.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll_b200.h"
#include "plf_backend.h"
#include "pll_host_internal.h"

__thread int pll_errno;
__thread char pll_errmsg[200] = {0};
__thread pll_hardware_t pll_hardware = {0};

static __thread int tls_device = -1;

#define PLL_CUDA_MAGIC 0xB200C0DEu
#define EMPTY_ELEMENT 0xFFFFFFFFu
/* sumtables live in HBM, keyed by the caller's host pointer: this many before the least recently used one
 * is dropped ($PLL_CUDA_MAX_SUMTABLES); the slot array grows on demand */
#define DEFAULT_MAX_SUMTABLES 64
/* pll_cuda_newton_branch: tables up to this size use the one-launch loop (they stay in the 126 MB L2) */
#define NEWTON_FUSED_MAX_BYTES ((size_t)96 << 20)
/* the streaming kernels fetch scalers and tip codes with 16-byte bulk copies:
 * the last copy of a buffer may read up to 15 bytes past its logical end */
#define BULK_PAD 16
/* site-repeat identifiers, batched per traversal level: lookup pool = this many
 * times lookup_buffer_size entries; workspace budget for the rank arrays */
#define REPEATS_POOL_FACTOR 16
#define REPEATS_BATCH_WS_BYTES ((size_t)1 << 30)

typedef struct sumtable_slot
{
  const double * key; /* caller's host pointer */
  double * dev;
  size_t doubles;
  unsigned long long stamp;
  double * asc_host; /* ascertainment bias: host copy of the `states` pseudo-site blocks */
} sumtable_slot_t;

typedef struct rid_memo
{
  unsigned int c1, c2, lookup_size;
  unsigned long long v1, v2;
  void * enable;
  int valid;
} rid_memo_t;

typedef struct pair_state
{
  unsigned int * dev;
  unsigned int cap, entries;
  unsigned int c1, c2;
  unsigned long long vp, v1, v2; /* identifier versions of parent and children the list was built from */
  int valid;
} pair_state_t;

typedef struct cherry_state
{
  unsigned int tip1, tip2;
  int scaler_index;
  int is_virtual; /* the node's CLV buffer does not hold its value: (tip1, tip2, snapshot matrices) do */
} cherry_state_t;

typedef struct cuda_partition
{
  pll_partition_t pub; /* MUST be first: callers hold &pub */
  unsigned long long partials_generation; /* bumped by every pll_update_partials */
  unsigned int * asc_sc;                  /* cached pseudo-site scalers [parent states][child states] */
  int asc_sc_valid, asc_sc_parent, asc_sc_child;
  unsigned long long asc_sc_generation;
  unsigned int magic;
  plf_ctx_t * ctx;
  plf_shape_t shape;

  double * d_pmatrix_block;
  size_t pmatrix_doubles;

  double * h_model;      /* packed per-rate-category model block */
  double * h_model_sent; /* bytes last uploaded                  */
  double * d_model;
  size_t model_doubles;
  int model_sent_valid;

  unsigned int * d_pattern_weights;
  int weights_dirty;
  int * d_invariant;

  unsigned char ** d_tipchars; /* [tips] device mirrors of tipchars[] */
  unsigned long long * d_tipmap;
  int tipmap_dirty;

  unsigned int * clv_entries;    /* site entries allocated per node          */
  unsigned int * scaler_entries; /* uints allocated per scale buffer          */

  sumtable_slot_t * sumtabs;
  unsigned int n_sumtabs, max_sumtabs;
  const double ** evicted_keys; /* tables this library computed and then dropped: their host bytes were never written */
  unsigned int n_evicted, cap_evicted;
  unsigned long long stamp;
  int sumtable_mirror;
  double * d_persite;
  size_t persite_cap;

  /* scratch for pll_set_tip_states */
  unsigned char * tip_scratch; /* page-locked host buffer: codes of the tip being set */
  int tip_scratch_busy;        /* its last upload may still be in flight */
  unsigned char * d_seq;
  unsigned long long * d_map;
  /* pattern-tip codes are formed on the device from the raw characters; the host copies tipchars[] are
   * brought up to date on request (pll_cuda_host_tipchars) or at once with $PLL_CUDA_TIPCHARS_MIRROR=1 */
  unsigned char * d_codes;       /* [sites + pseudo-sites] codes of the tip being set */
  unsigned short * d_tip_lut;    /* 256 entries + the first-illegal-site word behind them */
  unsigned char * tipchars_stale; /* [tips] */
  int tipchars_mirror, tip_host_map; /* $PLL_CUDA_TIP_HOST_MAP=1: the round-1 host mapping loop */

  /* site repeats: device-canonical identifier arrays */
  unsigned int ** d_site_id;    /* [nodes] -> device [sites]               */
  unsigned int ** d_id_site;    /* [nodes] -> device [id_site_count]       */
  unsigned int * id_site_count; /* classes found at the last id computation */
  unsigned char * ids_stale;    /* host mirrors older than the device copy */
  unsigned int * d_lookup;
  unsigned int * d_lookup_pool;           /* slices for the nodes of one traversal level */
  unsigned long long lookup_pool_entries;
  unsigned int * d_keys;
  unsigned char * d_rep_charmap;
  int repeats_mirror;
  /* identifiers of a whole operation list without a host round trip per level (default enable rule) */
  unsigned int * d_node_ids;              /* [nodes] device copy of pernode_ids */
  unsigned int * d_raw_ids;               /* [rid_cap] classes found per job */
  plf_rid_job_t * d_rid_jobs;             /* [rid_cap] */
  unsigned int rid_cap;
  unsigned long long * d_lookup64;        /* tagged lookup pool */
  unsigned int * d_rank_pool;             /* class number per lookup key, same index space */
  unsigned long long lookup64_entries;
  unsigned int rid_tag;                   /* host view of *d_rid_tag: tags below it are unused */
  unsigned int * d_rid_tag;               /* tag base in device memory (read by k_rid_min, lowered once per update) */
  /* one CUDA graph per repeated identifier update (same jobs, same pool): the ~5 launches per level are replayed */
  plf_rid_job_t * rid_graph_jobs;         /* host copy of the jobs of the last update */
  unsigned int * rid_graph_parts;         /* its part boundaries */
  unsigned int rid_graph_count, rid_graph_nparts, rid_graph_cap;
  unsigned long long * rid_graph_pool;
  void * rid_graph_scratch;
  void * rid_graph_exec;
  unsigned long long rid_graph_launches;
  int rid_graph_mode;                     /* $PLF_GRAPH=0: plain launches */
  void * d_rid_scratch;
  size_t rid_scratch_bytes;
  int rid_fast;                           /* $PLL_CUDA_REPEATS_LEVEL_SYNC=1 keeps one host synchronisation per level */
  /* identifiers are a pure function of the children's identifiers: an operation whose parent was last numbered
   * from the same children at the same identifier versions keeps what it has (rid_memo, by parent node) */
  struct rid_memo * rid_memo;
  int rid_memo_on;                        /* $PLL_CUDA_REPEATS_MEMO=0: every update renumbers every parent */
  /* pair lists of gathering operations (plf_op_t::pair_list) */
  struct pair_state * pairs;              /* [nodes], by parent */
  unsigned long long * ids_version;       /* [nodes] bumped whenever a node's identifiers may have changed */
  unsigned long long ids_clock;

  int host_expm1; /* bit-exact P-matrices: expm1 from the host libm */

  /* virtual cherries: tip-tip parents of 4-state pattern-tip partitions are not
   * written to HBM; their consumers work from the tip codes, anything else materialises them first */
  int cherry_ok;                 /* this partition's kernels consume virtual cherries */
  unsigned int cherry_maxstates; /* tip alphabet size cherry_ok was decided for (0: not yet) */
  unsigned int cherry_min_sites; /* narrower alignments write every parent ($PLF_VIRTUAL_CHERRY_MIN_SITES; default 20 states 20000, 4 states: above what runs as one launch) */
  double * d_cherry_pm;          /* [clv_buffers][2][rate_cats * 16]: the P-matrices each cherry was asked with */
  /* 20 states, narrow alignments: a pattern tip that meets an inner node is ALSO kept as an expanded CLV (built on
   * first use from its codes), so that a traversal level is one inner-inner launch instead of up to three kinds */
  int aa_tip_clvs;
  double ** d_tip_expanded;      /* [tips], NULL until used */
  unsigned char * tip_expanded_valid;
  struct cherry_state * cherry;  /* [nodes] */
  struct cherry_state * cherry_saved; /* roll-back copy while an operation list is resolved */
  unsigned int cherries_pending; /* nodes whose CLV is virtual right now */
  unsigned char * scaler_zero;   /* [scale_buffers] the buffer is known to hold zeros (written by a tip-tip op only) */
  unsigned char * scaler_zero_saved;

  /* reusable host scratch for operation lists */
  plf_op_t * h_ops;
  plf_op_t * h_ops_sorted;
  unsigned int * h_level;
  unsigned int * h_level_start;
  unsigned int ops_cap;
} cuda_partition_t;

/* ------------------------------------------------------------------------ */

static void set_error(int code, const char * fmt, const char * detail)
{
  pll_errno = code;
  snprintf(pll_errmsg, sizeof(pll_errmsg), fmt, detail ? detail : "");
}

static cuda_partition_t * CP(const pll_partition_t * p)
{
  cuda_partition_t * cp = (cuda_partition_t *)p;
  if (!cp || cp->magic != PLL_CUDA_MAGIC)
  {
    set_error(PLL_ERROR_CUDA_UNSUPPORTED, "partition was not created by libpll_b200%s", NULL);
    return NULL;
  }
  return cp;
}

static int cuda_fail(cuda_partition_t * cp)
{
  set_error(PLL_ERROR_CUDA, "CUDA: %s", plf_last_error(cp->ctx));
  return PLL_FAILURE;
}

/* per-site allocations carry `states` extra pseudo-sites under ascertainment bias correction
 * (src/pll.c:524-531): CLVs, scalers, tipchars, pattern weights, sumtables */
static unsigned int sites_alloc(const pll_partition_t * p)
{
  return p->sites + (p->asc_bias_alloc ? p->states : 0);
}

/* ---- virtual cherries -------------------------------------------------------------------- */

static int tipmap_on_device(cuda_partition_t * cp);

/* doubles of one P-matrix of all rate categories */
static size_t cherry_msz(const cuda_partition_t * cp)
{
  return (size_t)cp->pub.rate_cats * cp->pub.states * cp->pub.states_padded;
}

static double * cherry_snapshot(const cuda_partition_t * cp, unsigned int node)
{
  return cp->d_cherry_pm + (size_t)(node - cp->pub.tips) * 2 * cherry_msz(cp);
}

static int is_virtual(const cuda_partition_t * cp, unsigned int node)
{
  return cp->cherry && node < cp->pub.nodes && cp->cherry[node].is_virtual;
}

/* write the CLV of a virtual cherry to its buffer: the tip-tip kernel with the P-matrices the cherry was
 * asked to be computed with.  Every reader that is not a CLV operation calls this first. */
static int ensure_real(cuda_partition_t * cp, unsigned int node)
{
  const pll_partition_t * p = &cp->pub;
  const cherry_state_t * c;
  plf_op_t op;
  if (!is_virtual(cp, node)) return 1;
  c = &cp->cherry[node];
  memset(&op, 0, sizeof(op));
  op.kind = PLF_OP_TT;
  op.parent_clv = p->clv[node];
  op.parent_scaler = c->scaler_index >= 0 ? p->scale_buffer[c->scaler_index] : NULL;
  op.left_tip = cp->d_tipchars[c->tip1];
  op.right_tip = cp->d_tipchars[c->tip2];
  op.left_matrix = cherry_snapshot(cp, node);
  op.right_matrix = op.left_matrix + cherry_msz(cp);
  op.nsites = p->sites + (p->asc_bias_alloc ? p->states : 0);
  if (!tipmap_on_device(cp) || !plf_update_partials_once(cp->ctx, &cp->shape, &op, 1, cp->d_tipmap, p->maxstates))
  {
    pll_errno = PLL_ERROR_CUDA;
    snprintf(pll_errmsg, sizeof(pll_errmsg), "CUDA: %s", plf_last_error(cp->ctx));
    return 0;
  }
  cp->cherry[node].is_virtual = 0;
  --cp->cherries_pending;
  return 1;
}

/* (re)settle whether this partition keeps tip-tip parents virtual, for the tip alphabet it has now; when
 * the answer turns to no, the cherries that are still virtual are written out first */
static int cherry_decide(cuda_partition_t * cp)
{
  unsigned int n;
  if (!cp->cherry || cp->cherry_maxstates == cp->pub.maxstates) return 1;
  cp->cherry_maxstates = cp->pub.maxstates;
  cp->cherry_ok = plf_virtual_cherries_supported(cp->ctx, &cp->shape, cp->pub.maxstates);
  if (!cp->cherry_ok && cp->cherries_pending)
    for (n = cp->pub.tips; n < cp->pub.nodes; ++n)
      if (!ensure_real(cp, n)) return 0;
  return 1;
}

/* before the codes of `tip` change: cherries that still depend on the old ones */
static int ensure_real_for_tip(cuda_partition_t * cp, unsigned int tip)
{
  unsigned int n;
  if (cp->tip_expanded_valid && tip < cp->pub.tips) cp->tip_expanded_valid[tip] = 0; /* its codes are about to change */
  if (!cp->cherry || !cp->cherries_pending) return 1;
  for (n = cp->pub.tips; n < cp->pub.nodes; ++n)
    if (cp->cherry[n].is_virtual && (cp->cherry[n].tip1 == tip || cp->cherry[n].tip2 == tip) && !ensure_real(cp, n))
      return 0;
  return 1;
}

static int env_flag(const char * name)
{
  const char * v = getenv(name);
  return v && v[0] && v[0] != '0';
}

PLL_EXPORT void * pll_aligned_alloc(size_t size, size_t alignment)
{
  void * mem = NULL;
  if (posix_memalign(&mem, alignment < sizeof(void *) ? sizeof(void *) : alignment, size ? size : alignment))
    return NULL;
  return mem;
}

PLL_EXPORT void pll_aligned_free(void * ptr) { free(ptr); }

/* ---- hardware probe ------------------------------------------------------ */

PLL_EXPORT int pll_hardware_probe(void)
{
  memset(&pll_hardware, 0, sizeof(pll_hardware));
  pll_hardware.init = 1;
#if defined(__x86_64__)
  pll_hardware.sse3_present = __builtin_cpu_supports("sse3");
  pll_hardware.avx_present = __builtin_cpu_supports("avx");
  pll_hardware.avx2_present = __builtin_cpu_supports("avx2");
  pll_hardware.popcnt_present = __builtin_cpu_supports("popcnt");
#endif
  return PLL_SUCCESS;
}

PLL_EXPORT void pll_hardware_dump(void)
{
  char err[128] = {0};
  int n = plf_device_count(err, sizeof(err));
  fprintf(stderr, "CUDA devices: %d%s%s\n", n, err[0] ? " -- " : "", err);
}

PLL_EXPORT void pll_hardware_ignore(void)
{
  memset(&pll_hardware, 0, sizeof(pll_hardware));
  pll_hardware.init = 1;
}

PLL_EXPORT int pll_cuda_device_count(void)
{
  char err[160] = {0};
  int n = plf_device_count(err, sizeof(err));
  if (n <= 0 && err[0]) set_error(PLL_ERROR_CUDA, "%s", err);
  return n;
}

PLL_EXPORT int pll_cuda_set_device(int device)
{
  if (device < 0)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "negative CUDA device index%s", NULL);
    return PLL_FAILURE;
  }
  tls_device = device;
  return PLL_SUCCESS;
}

static int pick_device(void)
{
  const char * v;
  if (tls_device >= 0) return tls_device;
  if ((v = getenv("PLL_CUDA_DEVICE")) && v[0]) return atoi(v);
  if ((v = getenv("LOCAL_RANK")) && v[0]) return atoi(v);
  return 0;
}

/* ---- partition lifecycle -------------------------------------------------- */

static void free_repeats(cuda_partition_t * cp)
{
  pll_partition_t * p = &cp->pub;
  pll_repeats_t * r = p->repeats;
  unsigned int i;
  if (!r) return;
  for (i = 0; i < p->nodes; ++i)
  {
    if (r->pernode_site_id) free(r->pernode_site_id[i]);
    if (r->pernode_id_site) free(r->pernode_id_site[i]);
    if (cp->d_site_id) plf_free(cp->ctx, cp->d_site_id[i]);
    if (cp->d_id_site) plf_free(cp->ctx, cp->d_id_site[i]);
  }
  free(r->pernode_site_id);
  free(r->pernode_id_site);
  free(r->pernode_ids);
  free(r->perscale_ids);
  free(r->pernode_allocated_clvs);
  free(r->lookup_buffer);
  free(r->toclean_buffer);
  free(r->id_site_buffer);
  free(r->charmap);
  free(r);
  p->repeats = NULL;
  free(cp->d_site_id);
  free(cp->d_id_site);
  free(cp->id_site_count);
  free(cp->ids_stale);
  if (cp->pairs)
    for (i = 0; i < p->nodes; ++i) plf_free(cp->ctx, cp->pairs[i].dev);
  free(cp->pairs);
  free(cp->ids_version);
  free(cp->rid_memo);
  plf_free(cp->ctx, cp->d_node_ids);
  plf_free(cp->ctx, cp->d_raw_ids);
  plf_free(cp->ctx, cp->d_rid_jobs);
  plf_free(cp->ctx, cp->d_lookup64);
  plf_free(cp->ctx, cp->d_rank_pool);
  plf_free(cp->ctx, cp->d_rid_scratch);
  plf_free(cp->ctx, cp->d_rid_tag);
  plf_graph_free(cp->rid_graph_exec);
  free(cp->rid_graph_jobs);
  free(cp->rid_graph_parts);
  plf_free(cp->ctx, cp->d_lookup);
  plf_free(cp->ctx, cp->d_lookup_pool);
  plf_free(cp->ctx, cp->d_keys);
  plf_free(cp->ctx, cp->d_rep_charmap);
}

static void destroy(cuda_partition_t * cp)
{
  pll_partition_t * p = &cp->pub;
  unsigned int i;
  if (cp->ctx)
  {
    plf_sync(cp->ctx);
    free_repeats(cp);
    if (p->clv)
      for (i = 0; i < p->nodes; ++i) plf_free(cp->ctx, p->clv[i]);
    if (p->scale_buffer)
      for (i = 0; i < p->scale_buffers; ++i) plf_free(cp->ctx, p->scale_buffer[i]);
    if (cp->d_tipchars)
      for (i = 0; i < p->tips; ++i) plf_free(cp->ctx, cp->d_tipchars[i]);
    for (i = 0; i < cp->n_sumtabs; ++i)
    {
      plf_free(cp->ctx, cp->sumtabs[i].dev);
      free(cp->sumtabs[i].asc_host);
    }
    plf_free(cp->ctx, cp->d_pmatrix_block);
    plf_free(cp->ctx, cp->d_cherry_pm);
    if (cp->d_tip_expanded)
      for (i = 0; i < p->tips; ++i) plf_free(cp->ctx, cp->d_tip_expanded[i]);
    free(cp->d_tip_expanded);
    free(cp->tip_expanded_valid);
    plf_free(cp->ctx, cp->d_model);
    plf_free(cp->ctx, cp->d_pattern_weights);
    plf_free(cp->ctx, cp->d_invariant);
    plf_free(cp->ctx, cp->d_tipmap);
    plf_free(cp->ctx, cp->d_persite);
    plf_free(cp->ctx, cp->d_seq);
    plf_free(cp->ctx, cp->d_map);
    plf_free(cp->ctx, cp->d_codes);
    plf_free(cp->ctx, cp->d_tip_lut);
    plf_pinned_free(cp->ctx, cp->tip_scratch);
    plf_ctx_destroy(cp->ctx);
  }
  if (p->tipchars)
    for (i = 0; i < p->tips; ++i) free(p->tipchars[i]);
  for (i = 0; i < p->rate_matrices; ++i)
  {
    if (p->eigenvecs) free(p->eigenvecs[i]);
    if (p->inv_eigenvecs) free(p->inv_eigenvecs[i]);
    if (p->eigenvals) free(p->eigenvals[i]);
    if (p->subst_params) free(p->subst_params[i]);
    if (p->frequencies) free(p->frequencies[i]);
  }
  free(p->tipchars);
  free(p->charmap);
  free(p->tipmap);
  free(p->eigenvecs);
  free(p->inv_eigenvecs);
  free(p->eigenvals);
  free(p->subst_params);
  free(p->frequencies);
  free(p->eigen_decomp_valid);
  free(p->rates);
  free(p->rate_weights);
  free(p->prop_invar);
  free(p->pattern_weights);
  free(p->invariant);
  free(p->clv);
  free(p->pmatrix);
  free(p->scale_buffer);
  free(cp->d_tipchars);
  free(cp->clv_entries);
  free(cp->scaler_entries);
  free(cp->asc_sc);
  free(cp->tipchars_stale);
  free(cp->sumtabs);
  free(cp->evicted_keys);
  free(cp->cherry);
  free(cp->cherry_saved);
  free(cp->scaler_zero);
  free(cp->scaler_zero_saved);
  free(cp->h_model);
  free(cp->h_model_sent);
  free(cp->h_ops);
  free(cp->h_ops_sorted);
  free(cp->h_level);
  free(cp->h_level_start);
  cp->magic = 0;
  free(cp);
}

static int repeats_initialize(cuda_partition_t * cp)
{
  pll_partition_t * p = &cp->pub;
  unsigned int i;
  pll_repeats_t * r = (pll_repeats_t *)calloc(1, sizeof(pll_repeats_t));
  if (!r) return PLL_FAILURE;
  p->repeats = r;
  r->enable_repeats = pll_default_enable_repeats;
  r->reallocate_repeats = pll_default_reallocate_repeats;
  r->pernode_site_id = (unsigned int **)calloc(p->nodes, sizeof(unsigned int *));
  r->pernode_id_site = (unsigned int **)calloc(p->nodes, sizeof(unsigned int *));
  r->pernode_ids = (unsigned int *)calloc(p->nodes, sizeof(unsigned int));
  r->perscale_ids = (unsigned int *)calloc(p->scale_buffers ? p->scale_buffers : 1, sizeof(unsigned int));
  r->pernode_allocated_clvs = (unsigned int *)calloc(p->nodes, sizeof(unsigned int));
  r->toclean_buffer = (unsigned int *)malloc(p->sites * sizeof(unsigned int));
  r->id_site_buffer = (unsigned int *)malloc(p->sites * sizeof(unsigned int));
  r->charmap = (char *)calloc(PLL_ASCII_SIZE, sizeof(char));
  r->bclv_buffer = NULL; /* the device kernels need no pre-multiplied child buffer */
  cp->d_site_id = (unsigned int **)calloc(p->nodes, sizeof(unsigned int *));
  cp->d_id_site = (unsigned int **)calloc(p->nodes, sizeof(unsigned int *));
  cp->id_site_count = (unsigned int *)calloc(p->nodes, sizeof(unsigned int));
  cp->ids_stale = (unsigned char *)calloc(p->nodes, 1);
  if (!r->pernode_site_id || !r->pernode_id_site || !r->pernode_ids || !r->perscale_ids ||
      !r->pernode_allocated_clvs || !r->toclean_buffer || !r->id_site_buffer || !r->charmap ||
      !cp->d_site_id || !cp->d_id_site || !cp->id_site_count || !cp->ids_stale)
    return PLL_FAILURE;
  for (i = 0; i < p->nodes; ++i)
  {
    r->pernode_site_id[i] = (unsigned int *)calloc(p->sites, sizeof(unsigned int));
    r->pernode_id_site[i] = (unsigned int *)calloc(p->sites, sizeof(unsigned int));
    cp->d_site_id[i] = (unsigned int *)plf_alloc(cp->ctx, (size_t)p->sites * sizeof(unsigned int), 1);
    /* id -> site arrays keep room for `sites` classes on the device so that the
     * identifier kernels write them in place, whatever the class count turns
     * out to be (the host mirrors are sized exactly, as in the reference) */
    cp->d_id_site[i] = (unsigned int *)plf_alloc(cp->ctx, (size_t)p->sites * sizeof(unsigned int), 0);
    if (!r->pernode_site_id[i] || !r->pernode_id_site[i] || !cp->d_site_id[i] || !cp->d_id_site[i])
      return PLL_FAILURE;
  }
  cp->d_keys = (unsigned int *)plf_alloc(cp->ctx, (size_t)p->sites * sizeof(unsigned int), 0);
  cp->d_rep_charmap = (unsigned char *)plf_alloc(cp->ctx, PLL_ASCII_SIZE, 0);
  cp->d_node_ids = (unsigned int *)plf_alloc(cp->ctx, (size_t)p->nodes * sizeof(unsigned int), 1);
  cp->pairs = (pair_state_t *)calloc(p->nodes, sizeof(pair_state_t));
  cp->ids_version = (unsigned long long *)calloc(p->nodes, sizeof(unsigned long long));
  cp->rid_memo = (rid_memo_t *)calloc(p->nodes, sizeof(rid_memo_t));
  {
    const char * v = getenv("PLL_CUDA_REPEATS_MEMO");
    cp->rid_memo_on = !(v && v[0] == '0');
  }
  cp->rid_fast = !env_flag("PLL_CUDA_REPEATS_LEVEL_SYNC");
  cp->rid_tag = 0xFFFFFFFEu;
  cp->d_rid_tag = (unsigned int *)plf_alloc(cp->ctx, sizeof(unsigned int), 0);
  {
    const char * v = getenv("PLF_GRAPH");
    cp->rid_graph_mode = !(v && v[0] == '0');
  }
  if (!cp->d_keys || !cp->d_rep_charmap || !cp->d_node_ids || !cp->pairs || !cp->ids_version || !cp->d_rid_tag ||
      !plf_upload(cp->ctx, cp->d_rid_tag, &cp->rid_tag, sizeof(unsigned int)))
    return PLL_FAILURE;
  /* the first traversal sizes every CLV and scaler to its class count (~2 allocations per node): warm the
   * pool with a quarter of the uncompressed CLV volume so that they do not each grow it through the driver */
  plf_pool_reserve(cp->ctx, (size_t)p->nodes * p->sites * p->rate_cats * p->states_padded * sizeof(double) / 4);
  return PLL_SUCCESS;
}

PLL_EXPORT pll_partition_t * pll_partition_create(unsigned int tips, unsigned int clv_buffers,
                                                  unsigned int states, unsigned int sites,
                                                  unsigned int rate_matrices, unsigned int prob_matrices,
                                                  unsigned int rate_cats, unsigned int scale_buffers,
                                                  unsigned int attributes)
{
  unsigned int i, sp;
  int narch = __builtin_popcount(attributes & PLL_ATTRIB_ARCH_MASK) + ((attributes & PLL_ATTRIB_ARCH_CUDA) ? 1 : 0);
  char err[200] = {0};
  cuda_partition_t * cp;
  pll_partition_t * p;

  /* the CUDA bit counts as an architecture (src/pll.c:438-443) */
  if (narch > 1)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "Multiple architecture flags specified.%s", NULL);
    return NULL;
  }
  if (!(attributes & PLL_ATTRIB_ARCH_CUDA))
  {
    if (!env_flag("PLL_CUDA_FORCE"))
    {
      set_error(PLL_ERROR_CUDA_UNSUPPORTED,
                "libpll_b200 implements PLL_ATTRIB_ARCH_CUDA only (set it, or PLL_CUDA_FORCE=1)%s", NULL);
      return NULL;
    }
    attributes = (attributes & ~(unsigned int)PLL_ATTRIB_ARCH_MASK) | PLL_ATTRIB_ARCH_CUDA;
  }
  /* too few sites: repeats silently off (src/pll.c:446-449) */
  if (sites < 16) attributes &= ~(unsigned int)PLL_ATTRIB_SITE_REPEATS;
  if ((attributes & (PLL_ATTRIB_AB_MASK | PLL_ATTRIB_AB_FLAG)) && (attributes & PLL_ATTRIB_SITE_REPEATS))
  {
    set_error(PLL_ERROR_CUDA_UNSUPPORTED,
              "ascertainment bias correction together with site repeats is not implemented for CUDA partitions%s", NULL);
    return NULL;
  }
  if ((attributes & PLL_ATTRIB_SITE_REPEATS) && (attributes & PLL_ATTRIB_PATTERN_TIP))
  {
    set_error(PLL_ERROR_PARAM_INVALID, "PLL_ATTRIB_PATTERN_TIP and PLL_ATTRIB_SITE_REPEATS are mutually exclusive%s",
              NULL);
    return NULL;
  }
  if (!states || !sites || !rate_cats || !rate_matrices || states > 63)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "invalid partition dimensions%s", NULL);
    return NULL;
  }

  cp = (cuda_partition_t *)calloc(1, sizeof(cuda_partition_t));
  if (!cp)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate memory for partition.%s", NULL);
    return NULL;
  }
  p = &cp->pub;
  cp->magic = PLL_CUDA_MAGIC;
  p->tips = tips;
  p->clv_buffers = clv_buffers;
  p->nodes = tips + clv_buffers;
  p->states = states;
  p->sites = sites;
  p->pattern_weight_sum = sites;
  p->rate_matrices = rate_matrices;
  p->asc_bias_alloc = (attributes & (PLL_ATTRIB_AB_MASK | PLL_ATTRIB_AB_FLAG)) > 0;
  p->asc_additional_sites = p->asc_bias_alloc ? (int)states : 0;
  p->prob_matrices = prob_matrices;
  p->rate_cats = rate_cats;
  p->scale_buffers = scale_buffers;
  p->attributes = attributes;
  p->alignment = PLL_ALIGNMENT_CUDA;
  p->states_padded = sp = (states + 3) & 0xFFFFFFFCu;
  cp->shape.states = states;
  cp->shape.states_padded = sp;
  cp->shape.rate_cats = rate_cats;
  cp->shape.per_rate_scalers = (attributes & PLL_ATTRIB_RATE_SCALERS) ? 1 : 0;
  cp->host_expm1 = !env_flag("PLL_CUDA_DEVICE_EXPM1");
  cp->sumtable_mirror = env_flag("PLL_CUDA_SUMTABLE_MIRROR");
  cp->tipchars_mirror = env_flag("PLL_CUDA_TIPCHARS_MIRROR");
  cp->tip_host_map = env_flag("PLL_CUDA_TIP_HOST_MAP");
  cp->repeats_mirror = env_flag("PLL_CUDA_REPEATS_MIRROR");
  cp->weights_dirty = 1;

  if (!plf_ctx_create(pick_device(), env_flag("PLL_CUDA_MANAGED"), &cp->ctx, err, sizeof(err)))
  {
    set_error(PLL_ERROR_CUDA, "%s", err[0] ? err : "cannot create a CUDA context");
    cp->ctx = NULL;
    destroy(cp);
    return NULL;
  }

#define NEED(x)                                                                              \
  do                                                                                         \
  {                                                                                          \
    if (!(x))                                                                                \
    {                                                                                        \
      const char * ce = plf_last_error(cp->ctx);                                             \
      set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory. %s", ce ? ce : ""); \
      destroy(cp);                                                                           \
      return NULL;                                                                           \
    }                                                                                        \
  } while (0)

  NEED(p->eigen_decomp_valid = (int *)calloc(rate_matrices, sizeof(int)));
  NEED(p->clv = (double **)calloc(p->nodes ? p->nodes : 1, sizeof(double *)));
  NEED(cp->clv_entries = (unsigned int *)calloc(p->nodes ? p->nodes : 1, sizeof(unsigned int)));
  NEED(p->scale_buffer = (unsigned int **)calloc(scale_buffers ? scale_buffers : 1, sizeof(unsigned int *)));
  NEED(cp->scaler_entries = (unsigned int *)calloc(scale_buffers ? scale_buffers : 1, sizeof(unsigned int)));
  NEED(p->pmatrix = (double **)calloc(prob_matrices ? prob_matrices : 1, sizeof(double *)));

  /* CLVs: under repeats they are sized per node later; pattern tips have none
   * (src/pll.c:556-581) */
  if (!(attributes & PLL_ATTRIB_SITE_REPEATS))
  {
    const size_t n = (size_t)sites_alloc(p) * sp * rate_cats;
    for (i = (attributes & PLL_ATTRIB_PATTERN_TIP) ? tips : 0; i < p->nodes; ++i)
    {
      NEED(p->clv[i] = (double *)plf_alloc(cp->ctx, n * sizeof(double), 1));
      cp->clv_entries[i] = sites;
    }
    for (i = 0; i < scale_buffers; ++i)
    {
      const size_t m = (size_t)sites_alloc(p) * (cp->shape.per_rate_scalers ? rate_cats : 1);
      NEED(p->scale_buffer[i] = (unsigned int *)plf_alloc(cp->ctx, m * sizeof(unsigned int) + BULK_PAD, 1));
      cp->scaler_entries[i] = (unsigned int)m;
    }
  }

  /* one contiguous P-matrix block plus the displacement the padded rows of
   * the last matrix run into (src/pll.c:597-617) */
  cp->pmatrix_doubles = (size_t)prob_matrices * states * sp * rate_cats + (size_t)(sp - states) * sp;
  NEED(cp->d_pmatrix_block = (double *)plf_alloc(cp->ctx, cp->pmatrix_doubles * sizeof(double), 1));
  for (i = 0; i < prob_matrices; ++i) p->pmatrix[i] = cp->d_pmatrix_block + (size_t)i * states * sp * rate_cats;

  NEED(p->eigenvecs = (double **)calloc(rate_matrices, sizeof(double *)));
  NEED(p->inv_eigenvecs = (double **)calloc(rate_matrices, sizeof(double *)));
  NEED(p->eigenvals = (double **)calloc(rate_matrices, sizeof(double *)));
  NEED(p->subst_params = (double **)calloc(rate_matrices, sizeof(double *)));
  NEED(p->frequencies = (double **)calloc(rate_matrices, sizeof(double *)));
  for (i = 0; i < rate_matrices; ++i)
  {
    NEED(p->eigenvecs[i] = (double *)pll_aligned_alloc((size_t)states * sp * sizeof(double), p->alignment));
    NEED(p->inv_eigenvecs[i] = (double *)pll_aligned_alloc((size_t)states * sp * sizeof(double), p->alignment));
    NEED(p->eigenvals[i] = (double *)pll_aligned_alloc(sp * sizeof(double), p->alignment));
    NEED(p->subst_params[i] =
             (double *)pll_aligned_alloc((states * (states - 1) / 2 + 1) * sizeof(double), p->alignment));
    NEED(p->frequencies[i] = (double *)pll_aligned_alloc(sp * sizeof(double), p->alignment));
    memset(p->eigenvecs[i], 0, (size_t)states * sp * sizeof(double));
    memset(p->inv_eigenvecs[i], 0, (size_t)states * sp * sizeof(double));
    memset(p->eigenvals[i], 0, sp * sizeof(double));
    memset(p->subst_params[i], 0, (states * (states - 1) / 2 + 1) * sizeof(double));
    memset(p->frequencies[i], 0, sp * sizeof(double));
  }
  NEED(p->rates = (double *)calloc(rate_cats, sizeof(double)));
  NEED(p->rate_weights = (double *)calloc(rate_cats, sizeof(double)));
  for (i = 0; i < rate_cats; ++i) p->rate_weights[i] = 1.0 / rate_cats;
  NEED(p->prop_invar = (double *)calloc(rate_matrices, sizeof(double)));
  NEED(p->pattern_weights = (unsigned int *)malloc((size_t)sites_alloc(p) * sizeof(unsigned int)));
  for (i = 0; i < sites; ++i) p->pattern_weights[i] = 1;
  for (i = sites; i < sites_alloc(p); ++i) p->pattern_weights[i] = 0; /* src/pll.c:823-824 */
  NEED(cp->d_pattern_weights = (unsigned int *)plf_alloc(cp->ctx, (size_t)sites_alloc(p) * sizeof(unsigned int), 0));

  cp->model_doubles = plf_model_doubles(rate_cats, states, sp);
  NEED(cp->h_model = (double *)calloc(cp->model_doubles, sizeof(double)));
  NEED(cp->h_model_sent = (double *)calloc(cp->model_doubles, sizeof(double)));
  NEED(cp->d_model = (double *)plf_alloc(cp->ctx, cp->model_doubles * sizeof(double), 1));
  NEED(cp->d_seq = (unsigned char *)plf_alloc(cp->ctx, sites, 0));
  NEED(cp->d_map = (unsigned long long *)plf_alloc(cp->ctx, PLL_ASCII_SIZE * sizeof(unsigned long long), 0));
  NEED(cp->d_tipmap = (unsigned long long *)plf_alloc(cp->ctx, PLL_ASCII_SIZE * sizeof(unsigned long long), 1));

  if (attributes & PLL_ATTRIB_SITE_REPEATS) NEED(repeats_initialize(cp));

  if ((attributes & PLL_ATTRIB_PATTERN_TIP) && !p->asc_bias_alloc && clv_buffers && (states == 4 || states == 20) &&
      plf_virtual_cherries_supported(cp->ctx, &cp->shape, states == 4 ? 16 : 24))
  {
    /* whether tip-tip parents stay virtual is settled at the first operation list, when the tip alphabet
     * (maxstates: table sizes of the 20-state kernels) is known: cherry_decide() */
    const char * v = getenv("PLF_VIRTUAL_CHERRY_MIN_SITES");
    /* measured (profiles/r2_notes.md, 100 taxa): from 4096 sites up virtual cherries win (90 vs 106 us, at 10k
     * sites 127 vs 168 us); below ~2500 sites a traversal is bound by its launches and the one-launch-per-level
     * kernel with every parent written is fastest (1000 sites: 62 us against 78 us) */
    /* 20 states (200 taxa, profiles/r2_aa_mid_ab.json): the consumers of virtual cherries build large half tables per
     * CTA, which pays from ~20k sites on (16k sites: 988 us against 940 us with every parent written, 32k: 1563
     * against 1694 us; at 4000 sites 543 against 362 us) */
    cp->cherry_min_sites = (v && v[0]) ? (unsigned int)strtoul(v, NULL, 10) : (states == 20 ? 20000u : 2049u);
    if (!(v && v[0]) && states == 4 && rate_cats <= 4 && (rate_cats & (rate_cats - 1)) == 0)
    {
      /* 4 states: a plain list of up to PLF_FLOW_MAX_UPDATES site-updates (default 6.5M: 100 taxa x 66k sites,
       * 1000 taxa x 6.5k sites) over at most PLF_FLOW_MAX_SITES sites (default 65536) runs as ONE launch whose
       * paths keep parents in registers (k_clv_dna_flow), every parent written.  100 taxa: 1000 sites 27 us
       * against 62 us with one launch per level, 10k sites 63 us against 127 us with virtual cherries, 60k sites
       * 251 against 275 us; equal near 80k sites (7.8M site-updates) (profiles/r2_notes.md) */
      const char * f = getenv("PLF_FLOW"), * m = getenv("PLF_FLOW_MAX_SITES"), * u = getenv("PLF_FLOW_MAX_UPDATES");
      if (!(f && f[0] == '0'))
      {
        const unsigned long long max_sites = (m && m[0]) ? strtoull(m, NULL, 10) : 65536ull;
        const unsigned long long max_updates = (u && u[0]) ? strtoull(u, NULL, 10) : 6500000ull;
        const unsigned long long ops = tips > 2 ? tips - 2 : 1;
        unsigned long long lim = max_updates / ops;
        if (lim > max_sites) lim = max_sites;
        if (lim >= 2048ull) cp->cherry_min_sites = lim >= 0xFFFFFFFEull ? 0xFFFFFFFFu : (unsigned int)lim + 1u;
      }
    }
    if (sites >= cp->cherry_min_sites)
    {
      NEED(cp->cherry = (cherry_state_t *)calloc(p->nodes, sizeof(cherry_state_t)));
      NEED(cp->cherry_saved = (cherry_state_t *)calloc(p->nodes, sizeof(cherry_state_t)));
      NEED(cp->scaler_zero = (unsigned char *)calloc(scale_buffers ? scale_buffers : 1, 1));
      NEED(cp->scaler_zero_saved = (unsigned char *)calloc(scale_buffers ? scale_buffers : 1, 1));
      NEED(cp->d_cherry_pm = (double *)plf_alloc(cp->ctx, (size_t)clv_buffers * 2 * cherry_msz(cp) * sizeof(double), 1));
    }
  }
  if ((attributes & PLL_ATTRIB_PATTERN_TIP) && states == 20 && clv_buffers)
  {
    /* 200 taxa x 250 / 1000 sites: 125 / 201 us per traversal with the tip-inner kernels (25 launches), 70 / 143 us
     * with every tip an expanded CLV (17 launches) - profiles/r2_narrow_kinds.json.  Tip + tip keeps its kernel:
     * it never scales and zeroes its scaler (src/core_partials_avx.c:942-990) */
    /* ... 4000 / 8000 sites: 362 / 555 us (tip kernels) against 320 / 547 us; from 16k sites the tip kernels win
     * (profiles/r2_aa_mid_ab.json).  The expanded copies double the memory of the tips' share: at most 2 GiB */
    const char * v = getenv("PLF_AA_TIP_CLV_MAX_SITES");
    const unsigned long lim = (v && v[0]) ? strtoul(v, NULL, 10) : 8192ul;
    const char * m = getenv("PLF_AA_MMA");
    const unsigned long long bytes = (unsigned long long)tips * sites_alloc(p) * rate_cats * p->states_padded * sizeof(double);
    cp->aa_tip_clvs = sites <= lim && !(m && m[0] == '0') && !cp->cherry && (bytes <= (2ull << 30) || (v && v[0]));
  }
#undef NEED
  return p;
}

PLL_EXPORT void pll_partition_destroy(pll_partition_t * partition)
{
  cuda_partition_t * cp;
  if (!partition) return;
  if (!(cp = CP(partition))) return;
  destroy(cp);
}

PLL_EXPORT int pll_cuda_get_device(const pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  return cp ? plf_ctx_device(cp->ctx) : -1;
}

PLL_EXPORT void * pll_cuda_get_stream(const pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  return cp ? plf_ctx_stream(cp->ctx) : NULL;
}

PLL_EXPORT int pll_cuda_synchronize(const pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  return plf_sync(cp->ctx) ? PLL_SUCCESS : cuda_fail(cp);
}

PLL_EXPORT unsigned long long pll_cuda_kernel_launches(void) { return plf_kernel_launches(); }

/* NEW (additive, debugging).  With $PLL_CUDA_GUARD=1 in the environment when the partition was created, every
 * device buffer of the partition lies between two 256-byte guard bands; this call returns how many buffers
 * had a band written to since they were allocated (0 = clean, -1 = the partition was not created in guard
 * mode).  compute-sanitizer's memcheck is not available on every GPU pool: this is the bounds check the
 * parity suite runs on small and odd shapes instead. */
PLL_EXPORT int pll_cuda_check_guards(const pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  int n;
  if (!cp) return -2;
  n = plf_check_guards(cp->ctx);
  if (n > 0) set_error(PLL_ERROR_CUDA, "CUDA: %s", plf_last_error(cp->ctx));
  return n;
}

/* NEW (debugging).  Writes one byte just behind a CLV buffer: lets a test see that guard mode notices. */
PLL_EXPORT int pll_cuda_debug_overrun(pll_partition_t * partition, unsigned int clv_index)
{
  cuda_partition_t * cp = CP(partition);
  const unsigned char b = 0;
  size_t bytes;
  if (!cp || clv_index >= partition->nodes || !partition->clv[clv_index]) return PLL_FAILURE;
  bytes = (size_t)cp->clv_entries[clv_index] * partition->states_padded * partition->rate_cats * sizeof(double);
  bytes = (bytes + 16 + 255) & ~(size_t)255; /* the allocator's rounding: the band starts here */
  return plf_upload(cp->ctx, (unsigned char *)partition->clv[clv_index] + bytes, &b, 1) ? PLL_SUCCESS : cuda_fail(cp);
}

/* ---- site repeats: bookkeeping (src/repeats.c) ------------------------------ */

PLL_EXPORT int pll_repeats_enabled(const pll_partition_t * partition)
{
  return PLL_ATTRIB_SITE_REPEATS & partition->attributes;
}

PLL_EXPORT void pll_resize_repeats_lookup(pll_partition_t * partition, unsigned int size)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp || !size || !partition->repeats) return;
  plf_free(cp->ctx, cp->d_lookup);
  partition->repeats->lookup_buffer_size = size;
  cp->d_lookup = (unsigned int *)plf_alloc(cp->ctx, (size_t)size * sizeof(unsigned int), 0);
  if (!cp->d_lookup || !plf_fill_u32(cp->ctx, cp->d_lookup, EMPTY_ELEMENT, size)) cuda_fail(cp);
}

PLL_EXPORT unsigned int pll_get_sites_number(const pll_partition_t * partition, unsigned int clv_index)
{
  unsigned int sites = (partition->attributes & PLL_ATTRIB_SITE_REPEATS) ? partition->repeats->pernode_ids[clv_index] : 0;
  sites = sites ? sites : partition->sites;
  return sites + (partition->asc_bias_alloc ? partition->states : 0); /* src/repeats.c:66-70 */
}

PLL_EXPORT unsigned int pll_get_clv_size(const pll_partition_t * partition, unsigned int clv_index)
{
  return pll_get_sites_number(partition, clv_index) * partition->states_padded * partition->rate_cats;
}

/* bring the host mirrors of a node's identifier arrays up to date */
static void sync_ids_to_host(cuda_partition_t * cp, unsigned int node)
{
  pll_repeats_t * r = cp->pub.repeats;
  if (!r || !cp->ids_stale[node]) return;
  plf_download(cp->ctx, r->pernode_site_id[node], cp->d_site_id[node], (size_t)cp->pub.sites * sizeof(unsigned int));
  if (cp->id_site_count[node] && cp->d_id_site[node])
    plf_download(cp->ctx, r->pernode_id_site[node], cp->d_id_site[node],
                 (size_t)cp->id_site_count[node] * sizeof(unsigned int));
  cp->ids_stale[node] = 0;
}

PLL_EXPORT unsigned int * pll_get_site_id(const pll_partition_t * partition, unsigned int clv_index)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp || !pll_repeats_enabled(partition) || !partition->repeats->pernode_ids[clv_index]) return NULL;
  sync_ids_to_host(cp, clv_index);
  return partition->repeats->pernode_site_id[clv_index];
}

PLL_EXPORT unsigned int * pll_get_id_site(const pll_partition_t * partition, unsigned int clv_index)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp || !pll_repeats_enabled(partition) || !partition->repeats->pernode_ids[clv_index]) return NULL;
  sync_ids_to_host(cp, clv_index);
  return partition->repeats->pernode_id_site[clv_index];
}

PLL_EXPORT unsigned int pll_default_enable_repeats(pll_partition_t * partition, unsigned int left_clv,
                                                   unsigned int right_clv)
{
  const pll_repeats_t * r = partition->repeats;
  const unsigned long long nl = r->pernode_ids[left_clv], nr = r->pernode_ids[right_clv];
  const unsigned long long pairs = nl * nr;
  if (!pairs || (unsigned long long)r->lookup_buffer_size <= pairs) return 0;
  if (nl > partition->sites / 2 || nr > partition->sites / 2) return 0;
  return 1;
}

PLL_EXPORT unsigned int pll_no_enable_repeats(pll_partition_t * partition, unsigned int left_clv,
                                              unsigned int right_clv)
{
  (void)partition;
  (void)left_clv;
  (void)right_clv;
  return 0;
}

PLL_EXPORT void pll_disable_bclv(pll_partition_t * partition) { (void)partition; }

/* NEW (additive).  Forget which children every node's identifiers were computed from: the next
 * pll_update_partials renumbers every parent of its list (measurements; a client never needs it). */
PLL_EXPORT int pll_cuda_invalidate_repeat_identifiers(pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  unsigned int i;
  if (!cp) return PLL_FAILURE;
  if (cp->rid_memo)
    for (i = 0; i < partition->nodes; ++i) cp->rid_memo[i].valid = 0;
  return PLL_SUCCESS;
}

/* resize the parent's CLV, scale buffer and id->site array to the class count
 * (src/repeats.c:256-296) */
PLL_EXPORT void pll_default_reallocate_repeats(pll_partition_t * partition, unsigned int parent, int scaler_index,
                                               unsigned int sites_to_alloc)
{
  cuda_partition_t * cp = CP(partition);
  pll_repeats_t * r;
  if (!cp) return;
  r = partition->repeats;
  if (sites_to_alloc == r->pernode_allocated_clvs[parent]) return;
  r->pernode_allocated_clvs[parent] = sites_to_alloc;
  plf_free(cp->ctx, partition->clv[parent]);
  partition->clv[parent] = (double *)plf_alloc(
      cp->ctx, (size_t)sites_to_alloc * partition->states_padded * partition->rate_cats * sizeof(double), 1);
  cp->clv_entries[parent] = sites_to_alloc;
  if (!partition->clv[parent])
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory for repeats structure. %s",
              plf_last_error(cp->ctx));
    return;
  }
  if (scaler_index != PLL_SCALE_BUFFER_NONE)
  {
    size_t n = sites_to_alloc;
    if (partition->attributes & PLL_ATTRIB_RATE_SCALERS) n *= partition->rate_cats;
    plf_free(cp->ctx, partition->scale_buffer[scaler_index]);
    partition->scale_buffer[scaler_index] = (unsigned int *)plf_alloc(cp->ctx, n * sizeof(unsigned int) + BULK_PAD, 1);
    cp->scaler_entries[scaler_index] = (unsigned int)n;
  }
  free(r->pernode_id_site[parent]);
  r->pernode_id_site[parent] = (unsigned int *)malloc((size_t)(sites_to_alloc ? sites_to_alloc : 1) * sizeof(unsigned int));
}

/* class identifiers of a tip: classes are the distinct map values, numbered
 * by first occurrence along the sequence (src/repeats.c:189-254) */
PLL_EXPORT int pll_update_repeats_tips(pll_partition_t * partition, unsigned int tip_index, const pll_state_t * map,
                                       const char * sequence)
{
  cuda_partition_t * cp = CP(partition);
  pll_repeats_t * r;
  unsigned int i, j, ids = 0;
  unsigned char next = 0;
  unsigned char * cm;
  if (!cp) return PLL_FAILURE;
  r = partition->repeats;
  if (!cp->d_lookup) pll_resize_repeats_lookup(partition, PLL_REPEATS_LOOKUP_SIZE);
  if (!cp->d_lookup) return PLL_FAILURE;

  /* dense class code per character (src/repeats.c:28-45) */
  cm = (unsigned char *)r->charmap;
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
  {
    for (j = 0; j < i; ++j)
      if (map[i] == map[j])
      {
        cm[i] = cm[j];
        break;
      }
    if (!cm[i]) cm[i] = ++next;
  }
  if (!plf_upload(cp->ctx, cp->d_seq, sequence, partition->sites) ||
      !plf_upload(cp->ctx, cp->d_rep_charmap, cm, PLL_ASCII_SIZE) ||
      !plf_tip_keys(cp->ctx, cp->d_seq, cp->d_rep_charmap, partition->sites, cp->d_keys) ||
      !plf_repeats_ids(cp->ctx, partition->sites, cp->d_keys, 0, NULL, cp->d_site_id[tip_index],
                       cp->d_id_site[tip_index], cp->d_lookup, &ids))
    return cuda_fail(cp);
  r->pernode_ids[tip_index] = ids;
  cp->id_site_count[tip_index] = ids;
  cp->ids_version[tip_index] = ++cp->ids_clock;

  free(r->pernode_id_site[tip_index]);
  r->pernode_id_site[tip_index] = (unsigned int *)malloc((size_t)(ids ? ids : 1) * sizeof(unsigned int));
  plf_free(cp->ctx, partition->clv[tip_index]);
  partition->clv[tip_index] =
      (double *)plf_alloc(cp->ctx, (size_t)ids * partition->states_padded * partition->rate_cats * sizeof(double), 1);
  cp->clv_entries[tip_index] = ids;
  r->pernode_allocated_clvs[tip_index] = ids;
  if (!r->pernode_id_site[tip_index] || !partition->clv[tip_index])
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory for repeats structure. %s",
              plf_last_error(cp->ctx));
    return PLL_FAILURE;
  }
  cp->ids_stale[tip_index] = 1;
  if (cp->repeats_mirror) sync_ids_to_host(cp, tip_index);
  return PLL_SUCCESS;
}

/* bookkeeping after the class count of `op`'s parent is known
 * (src/repeats.c:349-381): `ids` = 0 means repeats are off for this node */
static void repeats_finish_op(cuda_partition_t * cp, const pll_operation_t * op, unsigned int ids)
{
  pll_partition_t * partition = &cp->pub;
  pll_repeats_t * r = partition->repeats;
  const unsigned int parent = op->parent_clv_index;
  unsigned int sites_to_alloc = ids ? ids : partition->sites;
  r->pernode_ids[parent] = ids;
  cp->ids_version[parent] = ++cp->ids_clock;
  if (cp->rid_memo) cp->rid_memo[parent].valid = 0; /* whoever numbered the node records what from (update_repeats_fast) */
  if (op->parent_scaler_index != PLL_SCALE_BUFFER_NONE) r->perscale_ids[op->parent_scaler_index] = ids;
  if (ids) cp->ids_stale[parent] = 1;
  r->reallocate_repeats(partition, parent, op->parent_scaler_index, sites_to_alloc);
  /* no compression on this node: tell the kernels not to gather */
  if (sites_to_alloc >= partition->sites)
  {
    r->pernode_ids[parent] = 0;
    if (op->parent_scaler_index != PLL_SCALE_BUFFER_NONE) r->perscale_ids[op->parent_scaler_index] = 0;
  }
  cp->id_site_count[parent] = ids;
  if (cp->repeats_mirror) sync_ids_to_host(cp, parent);
}

/* class identifiers of the parent of `op` from its children's
 * (src/repeats.c:299-382); identifiers are computed on the device */
PLL_EXPORT void pll_update_repeats(pll_partition_t * partition, const pll_operation_t * op)
{
  cuda_partition_t * cp = CP(partition);
  pll_repeats_t * r;
  unsigned int left, right, parent, ids = 0;
  if (!cp) return;
  r = partition->repeats;
  left = op->child1_clv_index;
  right = op->child2_clv_index;
  parent = op->parent_clv_index;
  if (!cp->d_lookup) pll_resize_repeats_lookup(partition, PLL_REPEATS_LOOKUP_SIZE);
  if (!cp->d_lookup) return;

  if (r->enable_repeats(partition, left, right) &&
      !plf_repeats_ids(cp->ctx, partition->sites, cp->d_site_id[left], r->pernode_ids[left], cp->d_site_id[right],
                       cp->d_site_id[parent], cp->d_id_site[parent], cp->d_lookup, &ids))
  {
    cuda_fail(cp);
    return;
  }
  repeats_finish_op(cp, op, ids);
}

/* Identifiers of a whole operation list.  A parent's identifiers depend only
 * on its children's, so the list is walked level by level (the CLV launch
 * levels): within a level every node that keeps repeats on is numbered by the
 * SAME six kernel launches, each in its own slice of a pooled lookup table,
 * and the host synchronises once per level to learn the class counts -- which
 * it needs before the next level, because enable_repeats (a caller-installable
 * host callback) looks at the children's counts, and reallocate_repeats sizes
 * the CLV and scale buffers by them. */
static int update_repeats_levels(cuda_partition_t * cp, const pll_operation_t * ops, unsigned int count)
{
  pll_partition_t * partition = &cp->pub;
  pll_repeats_t * r = partition->repeats;
  unsigned int i, j, nlevels, njobs = 0, max_jobs;
  unsigned long long pool_used = 0;
  unsigned int * level, * order, * start;
  plf_rep_job_t * jobs;
  unsigned int * job_op, * ids;
  int nl, ok = 1;
  if (!cp->d_lookup) pll_resize_repeats_lookup(partition, PLL_REPEATS_LOOKUP_SIZE);
  if (!cp->d_lookup) return 0;
  if (cp->lookup_pool_entries < (unsigned long long)r->lookup_buffer_size * REPEATS_POOL_FACTOR)
  {
    const unsigned long long want = (unsigned long long)r->lookup_buffer_size * REPEATS_POOL_FACTOR;
    plf_free(cp->ctx, cp->d_lookup_pool);
    cp->d_lookup_pool = (unsigned int *)plf_alloc(cp->ctx, (size_t)want * sizeof(unsigned int), 0);
    cp->lookup_pool_entries = 0;
    if (!cp->d_lookup_pool || !plf_fill_u32(cp->ctx, cp->d_lookup_pool, EMPTY_ELEMENT, (size_t)want))
      return cuda_fail(cp);
    cp->lookup_pool_entries = want;
  }
  /* workspace budget: rank arrays of sites entries per job */
  max_jobs = (unsigned int)(REPEATS_BATCH_WS_BYTES / ((size_t)partition->sites * sizeof(unsigned int) + 64));
  if (max_jobs < 1) max_jobs = 1;
  if (max_jobs > 32768) max_jobs = 32768;

  level = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  order = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  start = (unsigned int *)calloc((size_t)count + 2, sizeof(unsigned int));
  jobs = (plf_rep_job_t *)malloc((size_t)(count < max_jobs ? count : max_jobs) * sizeof(plf_rep_job_t));
  job_op = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  ids = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  nl = (level && order && start && jobs && job_op && ids) ? pll_cuda_schedule_levels(ops, count, level) : -1;
  if (nl < 0)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
    ok = 0;
    goto done;
  }
  nlevels = (unsigned int)nl;
  for (i = 0; i < count; ++i) start[level[i] + 1]++;
  for (i = 0; i < nlevels; ++i) start[i + 1] += start[i];
  {
    unsigned int * cursor = (unsigned int *)malloc(((size_t)nlevels + 1) * sizeof(unsigned int));
    if (!cursor)
    {
      ok = 0;
      goto done;
    }
    memcpy(cursor, start, ((size_t)nlevels + 1) * sizeof(unsigned int));
    for (i = 0; i < count; ++i) order[cursor[level[i]]++] = i;
    free(cursor);
  }

#define FLUSH_JOBS()                                                                                          \
  do                                                                                                          \
  {                                                                                                           \
    if (njobs)                                                                                                \
    {                                                                                                         \
      if (!plf_repeats_ids_batch(cp->ctx, partition->sites, jobs, njobs, cp->d_lookup_pool, ids))             \
      {                                                                                                       \
        cuda_fail(cp);                                                                                        \
        ok = 0;                                                                                               \
        goto done;                                                                                            \
      }                                                                                                       \
      for (j = 0; j < njobs; ++j) repeats_finish_op(cp, ops + job_op[j], ids[j]);                             \
      njobs = 0;                                                                                              \
      pool_used = 0;                                                                                          \
    }                                                                                                         \
  } while (0)

  for (i = 0; i < nlevels; ++i)
  {
    unsigned int k;
    for (k = start[i]; k < start[i + 1]; ++k)
    {
      const pll_operation_t * op = ops + order[k];
      const unsigned int left = op->child1_clv_index, right = op->child2_clv_index, parent = op->parent_clv_index;
      unsigned long long need;
      if (!r->enable_repeats(partition, left, right))
      {
        repeats_finish_op(cp, op, 0);
        continue;
      }
      need = (unsigned long long)r->pernode_ids[left] * r->pernode_ids[right];
      if (!need || need > cp->lookup_pool_entries)
      {
        /* a caller-installed enable_repeats asked for more than the pool holds */
        repeats_finish_op(cp, op, 0);
        continue;
      }
      if (pool_used + need > cp->lookup_pool_entries || njobs == max_jobs) FLUSH_JOBS();
      jobs[njobs].site_id_left = cp->d_site_id[left];
      jobs[njobs].site_id_right = cp->d_site_id[right];
      jobs[njobs].site_id_parent = cp->d_site_id[parent];
      jobs[njobs].id_site_parent = cp->d_id_site[parent];
      jobs[njobs].ids_left = r->pernode_ids[left];
      jobs[njobs].lookup_offset = (unsigned int)pool_used;
      job_op[njobs] = order[k];
      pool_used += need;
      ++njobs;
    }
    FLUSH_JOBS(); /* the next level reads this level's class counts */
  }
#undef FLUSH_JOBS
done:
  free(level);
  free(order);
  free(start);
  free(jobs);
  free(job_op);
  free(ids);
  return ok;
}

/* The same without a host round trip per level, when the enable rule is the library's own
 * (pll_default_enable_repeats or pll_no_enable_repeats): the rule is three integer compares on the
 * children's class counts (src/repeats.c:100-110), which the device has as soon as the level below is
 * numbered.  All levels are queued back to back; every job gets a slice of the tagged lookup pool as large
 * as the rule can ever ask for, min(upper bound of ids(left) * ids(right), lookup_buffer_size); the host
 * reads all class counts back ONCE at the end and then sizes CLVs, scalers and the host mirrors
 * (reallocate_repeats, in operation order as the reference calls it). */
#define RID_POOL_MAX_ENTRIES ((unsigned long long)1 << 29) /* 4 GiB of 64-bit entries */
static int update_repeats_fast_run(cuda_partition_t * cp, const pll_operation_t * ops, unsigned int count)
{
  pll_partition_t * partition = &cp->pub;
  pll_repeats_t * r = partition->repeats;
  const unsigned int sites = partition->sites;
  const unsigned long long lsize = r->lookup_buffer_size ? r->lookup_buffer_size : PLL_REPEATS_LOOKUP_SIZE;
  unsigned int i, nlevels;
  unsigned int * level = NULL, * order = NULL, * start = NULL, * raw = NULL, * ub = NULL;
  plf_rid_job_t * jobs = NULL;
  unsigned long long need_pool = 0;
  unsigned int max_jobs_ws;
  int nl, ok = 0;
  if (!r->lookup_buffer_size) r->lookup_buffer_size = PLL_REPEATS_LOOKUP_SIZE;

  if (r->enable_repeats == pll_no_enable_repeats)
  {
    for (i = 0; i < count; ++i) repeats_finish_op(cp, ops + i, 0);
    return 1;
  }
  level = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  order = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  start = (unsigned int *)calloc((size_t)count + 2, sizeof(unsigned int));
  raw = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  ub = (unsigned int *)malloc((size_t)partition->nodes * sizeof(unsigned int));
  jobs = (plf_rid_job_t *)malloc((size_t)count * sizeof(plf_rid_job_t));
  nl = (level && order && start && raw && ub && jobs) ? pll_cuda_schedule_levels(ops, count, level) : -1;
  if (nl < 0)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
    goto done;
  }
  nlevels = (unsigned int)nl;
  for (i = 0; i < count; ++i) start[level[i] + 1]++;
  for (i = 0; i < nlevels; ++i) start[i + 1] += start[i];
  {
    unsigned int * cursor = (unsigned int *)malloc(((size_t)nlevels + 1) * sizeof(unsigned int));
    if (!cursor) goto done;
    memcpy(cursor, start, ((size_t)nlevels + 1) * sizeof(unsigned int));
    for (i = 0; i < count; ++i) order[cursor[level[i]]++] = i;
    free(cursor);
  }
  /* jobs in level order; upper bounds of the class counts give the lookup slices */
  for (i = 0; i < partition->nodes; ++i) ub[i] = r->pernode_ids[i];
  for (i = 0; i < count; ++i)
  {
    const pll_operation_t * op = ops + order[i];
    const unsigned int left = op->child1_clv_index, right = op->child2_clv_index, parent = op->parent_clv_index;
    unsigned long long pairs = (unsigned long long)ub[left] * ub[right];
    if (pairs > lsize) pairs = lsize;
    jobs[i].site_id_left = cp->d_site_id[left];
    jobs[i].site_id_right = cp->d_site_id[right];
    jobs[i].site_id_parent = cp->d_site_id[parent];
    jobs[i].id_site_parent = cp->d_id_site[parent];
    jobs[i].left = left;
    jobs[i].right = right;
    jobs[i].parent = parent;
    jobs[i].lookup_entries = (unsigned int)pairs;
    jobs[i].lookup_offset = 0;
    ub[parent] = pairs < sites ? (unsigned int)pairs : sites;
  }
  /* pool: the widest level, capped; levels that need more run in several parts */
  for (i = 0; i < nlevels; ++i)
  {
    unsigned long long sum = 0;
    unsigned int k;
    for (k = start[i]; k < start[i + 1]; ++k) sum += jobs[k].lookup_entries;
    if (sum > need_pool) need_pool = sum;
  }
  if (need_pool > RID_POOL_MAX_ENTRIES) need_pool = RID_POOL_MAX_ENTRIES;
  if (need_pool < lsize) need_pool = lsize;
  if (cp->lookup64_entries < need_pool)
  {
    plf_free(cp->ctx, cp->d_lookup64);
    plf_free(cp->ctx, cp->d_rank_pool);
    cp->lookup64_entries = 0;
    cp->d_lookup64 = (unsigned long long *)plf_alloc(cp->ctx, (size_t)need_pool * sizeof(unsigned long long), 0);
    cp->d_rank_pool = (unsigned int *)plf_alloc(cp->ctx, (size_t)need_pool * sizeof(unsigned int), 0);
    if (!cp->d_lookup64 || !cp->d_rank_pool ||
        !plf_fill_u32(cp->ctx, (unsigned int *)cp->d_lookup64, EMPTY_ELEMENT, (size_t)need_pool * 2))
    {
      cuda_fail(cp);
      goto done;
    }
    cp->lookup64_entries = need_pool;
    cp->rid_tag = 0xFFFFFFFEu;
    if (!plf_upload(cp->ctx, cp->d_rid_tag, &cp->rid_tag, sizeof(unsigned int)))
    {
      cuda_fail(cp);
      goto done;
    }
  }
  if (cp->rid_cap < count)
  {
    plf_free(cp->ctx, cp->d_raw_ids);
    plf_free(cp->ctx, cp->d_rid_jobs);
    cp->d_raw_ids = (unsigned int *)plf_alloc(cp->ctx, (size_t)count * sizeof(unsigned int), 0);
    cp->d_rid_jobs = (plf_rid_job_t *)plf_alloc(cp->ctx, (size_t)count * sizeof(plf_rid_job_t), 0);
    cp->rid_cap = (cp->d_raw_ids && cp->d_rid_jobs) ? count : 0;
    if (!cp->rid_cap)
    {
      cuda_fail(cp);
      goto done;
    }
  }
  /* scratch: tile counts (one per 1024 sites) per job of one part */
  max_jobs_ws = PLF_MAX_RUN_OPS;
  {
    unsigned int widest = 0;
    size_t want;
    for (i = 0; i < nlevels; ++i)
      if (start[i + 1] - start[i] > widest) widest = start[i + 1] - start[i];
    if (widest > max_jobs_ws) widest = max_jobs_ws;
    want = plf_repeats_pass_workspace(sites, widest);
    if (cp->rid_scratch_bytes < want)
    {
      plf_free(cp->ctx, cp->d_rid_scratch);
      cp->d_rid_scratch = plf_alloc(cp->ctx, want, 0);
      cp->rid_scratch_bytes = cp->d_rid_scratch ? want : 0;
      if (!cp->d_rid_scratch)
      {
        cuda_fail(cp);
        goto done;
      }
    }
  }
  /* slices within each part, then everything is queued */
  {
    unsigned int k = 0;
    for (i = 0; i < nlevels; ++i)
    {
      k = start[i];
      while (k < start[i + 1])
      {
        unsigned long long used = 0;
        unsigned int first = k;
        while (k < start[i + 1] && k - first < max_jobs_ws && used + jobs[k].lookup_entries <= cp->lookup64_entries)
        {
          jobs[k].lookup_offset = used;
          used += jobs[k].lookup_entries;
          ++k;
        }
        if (k == first) /* cannot happen: one slice never exceeds the pool */
        {
          set_error(PLL_ERROR_CUDA, "repeat identifier lookup pool too small%s", NULL);
          goto done;
        }
        /* remember the part boundaries in `level`: level[first] = end of the part that starts at first */
        level[first] = k;
      }
    }
  }
  if (!plf_upload_async(cp->ctx, cp->d_node_ids, r->pernode_ids, (size_t)partition->nodes * sizeof(unsigned int)) ||
      !plf_upload_async(cp->ctx, cp->d_rid_jobs, jobs, (size_t)count * sizeof(plf_rid_job_t)))
  {
    cuda_fail(cp);
    goto done;
  }
  /* parts of the update: [parts[k], parts[k+1]) are the jobs of one pass */
  {
    unsigned int nparts = 0, k;
    int same, replay = 0;
    for (i = 0; i < count; i = level[i]) ++nparts;
    /* tags: every pass of this update below every tag used so far */
    if (cp->rid_tag < nparts + 1)
    {
      /* exhausted (2^32 passes): start over on a clean pool */
      cp->rid_tag = 0xFFFFFFFEu;
      if (!plf_fill_u32(cp->ctx, (unsigned int *)cp->d_lookup64, EMPTY_ELEMENT, (size_t)cp->lookup64_entries * 2) ||
          !plf_upload(cp->ctx, cp->d_rid_tag, &cp->rid_tag, sizeof(unsigned int)))
      {
        cuda_fail(cp);
        goto done;
      }
    }
    /* the same update as the last one (same jobs on the same buffers)?  The second such call is captured into
     * a graph, later ones replay it: ~5 launches per tree level become one */
    same = cp->rid_graph_mode && cp->rid_graph_count == count && cp->rid_graph_nparts == nparts &&
           cp->rid_graph_pool == cp->d_lookup64 && cp->rid_graph_scratch == cp->d_rid_scratch && cp->rid_graph_jobs &&
           !memcmp(cp->rid_graph_jobs, jobs, (size_t)count * sizeof(plf_rid_job_t));
    if (same)
      for (i = 0, k = 0; i < count && same; i = level[i], ++k) same = cp->rid_graph_parts[k] == i;
    if (same && cp->rid_graph_exec)
      replay = 1;
    else if (!same)
    {
      plf_graph_free(cp->rid_graph_exec);
      cp->rid_graph_exec = NULL;
      if (cp->rid_graph_cap < count)
      {
        free(cp->rid_graph_jobs);
        free(cp->rid_graph_parts);
        cp->rid_graph_jobs = (plf_rid_job_t *)malloc((size_t)count * sizeof(plf_rid_job_t));
        cp->rid_graph_parts = (unsigned int *)malloc(((size_t)count + 1) * sizeof(unsigned int));
        cp->rid_graph_cap = (cp->rid_graph_jobs && cp->rid_graph_parts) ? count : 0;
      }
      cp->rid_graph_count = 0;
      if (cp->rid_graph_cap >= count)
      {
        memcpy(cp->rid_graph_jobs, jobs, (size_t)count * sizeof(plf_rid_job_t));
        for (i = 0, k = 0; i < count; i = level[i], ++k) cp->rid_graph_parts[k] = i;
        cp->rid_graph_count = count;
        cp->rid_graph_nparts = nparts;
        cp->rid_graph_pool = cp->d_lookup64;
        cp->rid_graph_scratch = cp->d_rid_scratch;
      }
    }
    if (replay)
    {
      if (!plf_graph_replay(cp->ctx, cp->rid_graph_exec, cp->rid_graph_launches))
      {
        cuda_fail(cp);
        goto done;
      }
    }
    else
    {
      const int capture = same && plf_capture_begin(cp->ctx);
      const unsigned long long before = pll_cuda_kernel_launches();
      int queued = plf_repeats_advance_tags(cp->ctx, cp->d_rid_tag, nparts);
      for (i = 0, k = 0; i < count && queued; i = level[i], ++k)
        queued = plf_repeats_pass(cp->ctx, sites, r->lookup_buffer_size, cp->d_rid_jobs, i, level[i] - i, cp->d_lookup64,
                                  cp->d_rank_pool, cp->d_rid_tag, nparts - 1 - k, cp->d_node_ids, cp->d_raw_ids,
                                  cp->d_rid_scratch);
      if (capture)
      {
        void * exec = NULL;
        const unsigned long long captured = pll_cuda_kernel_launches() - before;
        if (!queued || !plf_capture_end(cp->ctx, &exec))
        {
          /* capture is not available here: plain launches from now on */
          if (!queued) plf_capture_abort(cp->ctx);
          cp->rid_graph_mode = 0;
          queued = plf_repeats_advance_tags(cp->ctx, cp->d_rid_tag, nparts);
          for (i = 0, k = 0; i < count && queued; i = level[i], ++k)
            queued = plf_repeats_pass(cp->ctx, sites, r->lookup_buffer_size, cp->d_rid_jobs, i, level[i] - i,
                                      cp->d_lookup64, cp->d_rank_pool, cp->d_rid_tag, nparts - 1 - k, cp->d_node_ids,
                                      cp->d_raw_ids, cp->d_rid_scratch);
        }
        else
        {
          cp->rid_graph_exec = exec;
          cp->rid_graph_launches = captured;
          queued = plf_graph_replay(cp->ctx, exec, 0); /* the launches were counted while they were captured */
        }
      }
      if (!queued)
      {
        cuda_fail(cp);
        goto done;
      }
    }
    cp->rid_tag -= nparts;
  }
  /* the one synchronisation of the identifier update */
  if (!plf_download(cp->ctx, raw, cp->d_raw_ids, (size_t)count * sizeof(unsigned int)))
  {
    cuda_fail(cp);
    goto done;
  }
  /* bookkeeping in the order of the list, as the reference does it */
  for (i = 0; i < count; ++i) level[order[i]] = raw[i];
  for (i = 0; i < count; ++i) repeats_finish_op(cp, ops + i, level[i]);
  ok = 1;
done:
  free(level);
  free(order);
  free(start);
  free(raw);
  free(ub);
  free(jobs);
  return ok;
}

/* Front end of the identifier update of an operation list: identifiers are a pure function of the children's
 * identifiers (src/repeats.c:334-347), so a parent that was last numbered from the same two children, at the
 * identifier versions they still have, under the same enable rule and lookup size, keeps its identifiers, its
 * class count and its buffers; only the operations above a changed node are renumbered.  A tree search that
 * re-evaluates a list after a local move renumbers the path to the root, not the tree. */
static int update_repeats_fast(cuda_partition_t * cp, const pll_operation_t * ops, unsigned int count)
{
  pll_partition_t * partition = &cp->pub;
  pll_repeats_t * r = partition->repeats;
  pll_operation_t * sel;
  unsigned char * changed;
  unsigned int i, n = 0;
  int ok;
  if (!cp->rid_memo_on || !cp->rid_memo) return update_repeats_fast_run(cp, ops, count);
  sel = (pll_operation_t *)malloc((size_t)count * sizeof(pll_operation_t));
  changed = (unsigned char *)calloc(partition->nodes ? partition->nodes : 1, 1);
  if (!sel || !changed)
  {
    free(sel);
    free(changed);
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
    return 0;
  }
  for (i = 0; i < count; ++i) /* list order is dependency order (src/partials.c:253 runs it sequentially) */
  {
    const unsigned int c1 = ops[i].child1_clv_index, c2 = ops[i].child2_clv_index, par = ops[i].parent_clv_index;
    const rid_memo_t * m = &cp->rid_memo[par];
    const int keep = m->valid && m->c1 == c1 && m->c2 == c2 && m->v1 == cp->ids_version[c1] &&
                     m->v2 == cp->ids_version[c2] && !changed[c1] && !changed[c2] &&
                     m->lookup_size == r->lookup_buffer_size && m->enable == (void *)r->enable_repeats;
    if (!keep)
    {
      sel[n++] = ops[i];
      changed[par] = 1;
    }
  }
  ok = n ? update_repeats_fast_run(cp, sel, n) : 1;
  if (ok)
    for (i = 0; i < n; ++i)
    {
      rid_memo_t * m = &cp->rid_memo[sel[i].parent_clv_index];
      m->c1 = sel[i].child1_clv_index;
      m->c2 = sel[i].child2_clv_index;
      m->v1 = cp->ids_version[m->c1];
      m->v2 = cp->ids_version[m->c2];
      m->lookup_size = r->lookup_buffer_size;
      m->enable = (void *)r->enable_repeats;
      m->valid = 1;
    }
  free(sel);
  free(changed);
  return ok;
}

/* ---- tips ------------------------------------------------------------------- */

static unsigned int ceil_log2(unsigned int x)
{
  unsigned int l = 0;
  while ((1u << l) < x) ++l;
  return l;
}

/* first call: dense codes for the distinct non-zero map values, in order of
 * the first character that carries them (src/pll.c:295-422) */
static int charmap_create(cuda_partition_t * cp, const pll_state_t * usermap)
{
  pll_partition_t * p = &cp->pub;
  unsigned int i, j, k = 0;
  pll_state_t top = 0;
  p->charmap = (unsigned char *)calloc(PLL_ASCII_SIZE, sizeof(unsigned char));
  p->tipmap = (pll_state_t *)calloc(PLL_ASCII_SIZE, sizeof(pll_state_t));
  p->tipchars = (unsigned char **)calloc(p->tips ? p->tips : 1, sizeof(unsigned char *));
  cp->d_tipchars = (unsigned char **)calloc(p->tips ? p->tips : 1, sizeof(unsigned char *));
  if (!p->charmap || !p->tipmap || !p->tipchars || !cp->d_tipchars)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate charmap for tip-tip precomputation.%s", NULL);
    return PLL_FAILURE;
  }
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
  {
    if (!usermap[i]) continue;
    for (j = 0; j < i; ++j)
      if (usermap[j] == usermap[i]) break;
    if (j < i)
      p->charmap[i] = p->charmap[j];
    else
    {
      if (usermap[i] > top) top = usermap[i];
      p->charmap[i] = (unsigned char)k;
      p->tipmap[k++] = usermap[i];
    }
  }
  /* 4 states: tipchars hold the raw mask, code 0 is a fictive unused state */
  p->maxstates = (p->states == 4) ? (unsigned int)top + 1 : k;
  (void)ceil_log2; /* tip-tip tables live in shared memory: no ttlookup allocation */
  for (i = 0; i < p->tips; ++i)
  {
    p->tipchars[i] = (unsigned char *)malloc(sites_alloc(p));
    cp->d_tipchars[i] = (unsigned char *)plf_alloc(cp->ctx, (size_t)sites_alloc(p) + BULK_PAD, 1);
    if (!p->tipchars[i] || !cp->d_tipchars[i])
    {
      set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate space for storing tip characters.%s", NULL);
      return PLL_FAILURE;
    }
  }
  cp->tipmap_dirty = 1;
  return PLL_SUCCESS;
}

/* later calls with a (possibly different) map: known values keep their code,
 * new ones are appended (src/pll.c:157-286) */
static int charmap_update(cuda_partition_t * cp, const pll_state_t * map)
{
  pll_partition_t * p = &cp->pub;
  unsigned int i, j, k = 0, added = 0;
  unsigned char newmap[PLL_ASCII_SIZE];
  pll_state_t newtips[PLL_ASCII_SIZE];
  while (k < PLL_ASCII_SIZE && p->tipmap[k]) ++k;
  memset(newmap, 0, sizeof(newmap));
  memcpy(newtips, p->tipmap, sizeof(newtips));
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
  {
    if (!map[i]) continue;
    for (j = 0; j < k + added; ++j)
      if (newtips[j] == map[i]) break;
    if (j == k + added)
    {
      if (k + added + 1 >= PLL_ASCII_SIZE)
      {
        memset(p->charmap, 0, PLL_ASCII_SIZE);
        snprintf(pll_errmsg, 200, "Cannot specify 256 or more states with PLL_ATTRIB_PATTERN_TIP.");
        return PLL_FAILURE;
      }
      newtips[j] = map[i];
      ++added;
    }
    newmap[i] = (unsigned char)j;
  }
  memcpy(p->charmap, newmap, PLL_ASCII_SIZE);
  if (added)
  {
    memcpy(p->tipmap, newtips, sizeof(newtips));
    if (p->states == 4)
    {
      pll_state_t top = 0;
      for (i = 0; p->tipmap[i]; ++i)
        if (p->tipmap[i] > top) top = p->tipmap[i];
      p->maxstates = (unsigned int)top + 1;
    }
    else
      p->maxstates += added;
    cp->tipmap_dirty = 1;
  }
  return PLL_SUCCESS;
}

/* dst[i] = low byte of lut[seq[i]] (dst may be NULL); returns the OR of every entry looked up, so a flag kept
 * in bit 8 of the entries of illegal characters says whether the sequence holds one.  Eight characters per
 * step: this loop is what pll_set_tip_states costs on the host (1M sites: 0.9 ms against 3.2 ms for a
 * check pass plus a mapping pass, one character at a time). */
static unsigned int map_bytes(unsigned char * dst, const char * seq, size_t n, const unsigned short * lut)
{
  unsigned int seen = 0;
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
  {
    unsigned long long w, o;
    unsigned int c0, c1, c2, c3, c4, c5, c6, c7;
    memcpy(&w, seq + i, 8);
    c0 = lut[w & 255];
    c1 = lut[(w >> 8) & 255];
    c2 = lut[(w >> 16) & 255];
    c3 = lut[(w >> 24) & 255];
    c4 = lut[(w >> 32) & 255];
    c5 = lut[(w >> 40) & 255];
    c6 = lut[(w >> 48) & 255];
    c7 = lut[w >> 56];
    seen |= c0 | c1 | c2 | c3 | c4 | c5 | c6 | c7;
    if (dst)
    {
      o = (unsigned long long)(c0 & 255) | ((unsigned long long)(c1 & 255) << 8) |
          ((unsigned long long)(c2 & 255) << 16) | ((unsigned long long)(c3 & 255) << 24) |
          ((unsigned long long)(c4 & 255) << 32) | ((unsigned long long)(c5 & 255) << 40) |
          ((unsigned long long)(c6 & 255) << 48) | ((unsigned long long)(c7 & 255) << 56);
      memcpy(dst + i, &o, 8);
    }
  }
  for (; i < n; ++i)
  {
    const unsigned int c = lut[(unsigned char)seq[i]];
    seen |= c;
    if (dst) dst[i] = (unsigned char)c;
  }
  return seen;
}

#define ILLEGAL_CHAR 0x100u

static int check_sequence(const pll_partition_t * p, const pll_state_t * map, const char * sequence)
{
  unsigned short lut[PLL_ASCII_SIZE];
  unsigned int i;
  for (i = 0; i < PLL_ASCII_SIZE; ++i) lut[i] = map[i] ? 0 : ILLEGAL_CHAR;
  if (!(map_bytes(NULL, sequence, p->sites, lut) & ILLEGAL_CHAR)) return PLL_SUCCESS;
  for (i = 0; i < p->sites; ++i)
    if (map[(unsigned char)sequence[i]] == 0)
    {
      pll_errno = PLL_ERROR_TIPDATA_ILLEGALSTATE;
      snprintf(pll_errmsg, 200, "Illegal state code in tip \"%c\"", sequence[i]);
      return PLL_FAILURE;
    }
  return PLL_SUCCESS;
}

/* ascertainment bias: pseudo-site i of a tip CLV is the unit vector of state i, replicated over the
 * rates (src/pll.c:1002-1020, 1112-1126) */
static int upload_asc_tip_clv(cuda_partition_t * cp, unsigned int tip_index)
{
  const pll_partition_t * p = &cp->pub;
  const unsigned int st = p->states, sp = p->states_padded, R = p->rate_cats;
  const size_t n = (size_t)st * R * sp;
  double * extra = (double *)calloc(n, sizeof(double));
  unsigned int i, j;
  int ok;
  if (!extra) return 0;
  for (i = 0; i < st; ++i)
    for (j = 0; j < R; ++j) extra[((size_t)i * R + j) * sp + i] = 1.0;
  ok = plf_upload(cp->ctx, p->clv[tip_index] + (size_t)p->sites * R * sp, extra, n * sizeof(double));
  free(extra);
  return ok;
}

/* Pattern-tip codes on the device (src/pll.c:875-957): the raw characters go up as they are (1 byte per
 * site), one kernel maps them and finds the first character the map does not know, the codes are installed
 * only when there is none.  The host copy tipchars[tip] is not written here (no reference client reads it;
 * pll_cuda_host_tipchars() brings it up to date). */
static int set_pattern_tip_on_device(cuda_partition_t * cp, unsigned int tip_index, const pll_state_t * map,
                                     const char * sequence)
{
  pll_partition_t * partition = &cp->pub;
  const unsigned int sites = partition->sites;
  unsigned short lut[PLL_ASCII_SIZE];
  unsigned int i, bad = EMPTY_ELEMENT;
  unsigned int * d_flag;
  if (partition->tipchars)
    charmap_update(cp, map);
  else if (!charmap_create(cp, map))
    return PLL_FAILURE;
  if (!cp->tipchars_stale) cp->tipchars_stale = (unsigned char *)calloc(partition->tips ? partition->tips : 1, 1);
  if (!cp->d_codes) cp->d_codes = (unsigned char *)plf_alloc(cp->ctx, (size_t)sites_alloc(partition) + BULK_PAD, 1);
  if (!cp->d_tip_lut) cp->d_tip_lut = (unsigned short *)plf_alloc(cp->ctx, PLL_ASCII_SIZE * sizeof(unsigned short) + 16, 1);
  if (!cp->tipchars_stale || !cp->d_codes || !cp->d_tip_lut)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate space for storing tip characters.%s", NULL);
    return PLL_FAILURE;
  }
  d_flag = (unsigned int *)(cp->d_tip_lut + PLL_ASCII_SIZE);
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
    lut[i] = map[i] ? (unsigned short)(partition->states == 4 ? (map[i] & 255) : partition->charmap[i]) : ILLEGAL_CHAR;
  if (!plf_upload_async(cp->ctx, cp->d_tip_lut, lut, sizeof(lut)) ||
      !plf_upload_async(cp->ctx, cp->d_seq, sequence, sites) ||
      !plf_tip_map(cp->ctx, cp->d_seq, cp->d_tip_lut, sites, cp->d_codes, d_flag) ||
      !plf_download(cp->ctx, &bad, d_flag, sizeof(bad)))
    return cuda_fail(cp);
  if (bad != EMPTY_ELEMENT)
  {
    pll_errno = PLL_ERROR_TIPDATA_ILLEGALSTATE;
    snprintf(pll_errmsg, 200, "Illegal state code in tip \"%c\"", sequence[bad]);
    return PLL_FAILURE;
  }
  if (!ensure_real_for_tip(cp, tip_index)) return PLL_FAILURE;
  if (partition->asc_bias_alloc)
  {
    /* pseudo-site i: every tip shows state i (src/pll.c:897-905, 935-957) */
    unsigned char extra[64];
    if (partition->states == 4)
      for (i = 0; i < 4; ++i) extra[i] = (unsigned char)(1u << i);
    else
    {
      memset(extra, 0, sizeof(extra));
      for (i = 0; i < partition->maxstates; ++i)
      {
        const pll_state_t state = partition->tipmap[i];
        if (state && !(state & (state - 1)))
        {
          const unsigned int pos = (unsigned int)__builtin_ctzll(state);
          if (pos < partition->states) extra[pos] = (unsigned char)i;
        }
      }
    }
    if (!plf_upload_async(cp->ctx, cp->d_codes + sites, extra, partition->states)) return cuda_fail(cp);
  }
  if (!plf_copy_d2d(cp->ctx, cp->d_tipchars[tip_index], cp->d_codes, sites_alloc(partition))) return cuda_fail(cp);
  cp->tipchars_stale[tip_index] = 1;
  if (cp->tipchars_mirror && !pll_cuda_host_tipchars(partition, tip_index)) return PLL_FAILURE;
  return PLL_SUCCESS;
}

/* NEW (additive).  tipchars[tip_index] of a CUDA partition, brought up to date from the device copy. */
PLL_EXPORT const unsigned char * pll_cuda_host_tipchars(pll_partition_t * partition, unsigned int tip_index)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp || !partition->tipchars || tip_index >= partition->tips) return NULL;
  if (cp->tipchars_stale && cp->tipchars_stale[tip_index])
  {
    if (!plf_download(cp->ctx, partition->tipchars[tip_index], cp->d_tipchars[tip_index], sites_alloc(partition)))
    {
      cuda_fail(cp);
      return NULL;
    }
    cp->tipchars_stale[tip_index] = 0;
  }
  return partition->tipchars[tip_index];
}

PLL_EXPORT int pll_set_tip_states(pll_partition_t * partition, unsigned int tip_index, const pll_state_t * map,
                                  const char * sequence)
{
  cuda_partition_t * cp = CP(partition);
  unsigned int i;
  if (!cp) return PLL_FAILURE;
  if (tip_index >= partition->tips)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "tip index out of range%s", NULL);
    return PLL_FAILURE;
  }
  /* pattern tips: the legality check rides on the mapping pass (on the device, or the 4-state host loop) */
  if (!((partition->attributes & PLL_ATTRIB_PATTERN_TIP) && (partition->states == 4 || !cp->tip_host_map) &&
        !pll_repeats_enabled(partition)) &&
      !check_sequence(partition, map, sequence))
    return PLL_FAILURE;

  if (pll_repeats_enabled(partition) && !pll_update_repeats_tips(partition, tip_index, map, sequence))
    return PLL_FAILURE;

  if ((partition->attributes & PLL_ATTRIB_PATTERN_TIP) && !cp->tip_host_map)
    return set_pattern_tip_on_device(cp, tip_index, map, sequence);

  if (partition->attributes & PLL_ATTRIB_PATTERN_TIP)
  {
    /* The codes are built in a page-locked scratch buffer, sent from there without waiting (the next call waits
     * before it reuses the scratch) and copied into the host mirror while the transfer runs.  Nothing of the
     * partition is touched when the sequence holds an illegal character. */
    unsigned char * tc, * scratch;
    unsigned short lut[PLL_ASCII_SIZE];
    if (!cp->tip_scratch) cp->tip_scratch = (unsigned char *)plf_pinned_alloc(cp->ctx, (size_t)sites_alloc(partition) + 8);
    if (!cp->tip_scratch) return cuda_fail(cp);
    if (cp->tip_scratch_busy && !plf_sync(cp->ctx)) return cuda_fail(cp);
    cp->tip_scratch_busy = 0;
    scratch = cp->tip_scratch;
    if (partition->states == 4)
    {
      for (i = 0; i < PLL_ASCII_SIZE; ++i) lut[i] = map[i] ? (unsigned short)(map[i] & 255) : ILLEGAL_CHAR;
      if (map_bytes(scratch, sequence, partition->sites, lut) & ILLEGAL_CHAR)
        return check_sequence(partition, map, sequence);
    }
    if (partition->tipchars)
      charmap_update(cp, map);
    else if (!charmap_create(cp, map))
      return PLL_FAILURE;
    tc = partition->tipchars[tip_index];
    if (partition->states != 4)
    {
      for (i = 0; i < PLL_ASCII_SIZE; ++i) lut[i] = partition->charmap[i];
      map_bytes(scratch, sequence, partition->sites, lut);
    }
    if (partition->asc_bias_alloc)
    {
      /* pseudo-site i: every tip shows state i (src/pll.c:897-905, 935-957) */
      unsigned char * extra = scratch + partition->sites;
      if (partition->states == 4)
        for (i = 0; i < 4; ++i) extra[i] = (unsigned char)(1u << i);
      else
      {
        memset(extra, 0, partition->states);
        for (i = 0; i < partition->maxstates; ++i)
        {
          const pll_state_t state = partition->tipmap[i];
          if (state && !(state & (state - 1)))
          {
            const unsigned int pos = (unsigned int)__builtin_ctzll(state);
            if (pos < partition->states) extra[pos] = (unsigned char)i;
          }
        }
      }
    }
    if (!ensure_real_for_tip(cp, tip_index)) return PLL_FAILURE;
    if (!plf_upload_async(cp->ctx, cp->d_tipchars[tip_index], scratch, sites_alloc(partition))) return cuda_fail(cp);
    cp->tip_scratch_busy = 1;
    memcpy(tc, scratch, sites_alloc(partition));
    return PLL_SUCCESS;
  }

  /* tip CLV: bit j of the state mask -> entry j, replicated over the rates;
   * built on the device from the raw characters (src/pll.c:959-1024) */
  {
    const int rep = pll_repeats_enabled(partition);
    const unsigned int entries = rep ? partition->repeats->pernode_ids[tip_index] : partition->sites;
    if (!plf_upload(cp->ctx, cp->d_seq, sequence, partition->sites) ||
        !plf_upload(cp->ctx, cp->d_map, map, PLL_ASCII_SIZE * sizeof(pll_state_t)) ||
        !plf_tip_clv_from_states(cp->ctx, &cp->shape, partition->clv[tip_index], cp->d_seq, cp->d_map,
                                 rep ? cp->d_id_site[tip_index] : NULL, entries))
      return cuda_fail(cp);
    if (partition->asc_bias_alloc && !upload_asc_tip_clv(cp, tip_index)) return cuda_fail(cp);
  }
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_set_tip_clv(pll_partition_t * partition, unsigned int tip_index, const double * clv, int padding)
{
  cuda_partition_t * cp = CP(partition);
  unsigned int i, j, entries;
  const unsigned int sp = partition->states_padded, st = partition->states, R = partition->rate_cats;
  const unsigned int in_states = padding ? sp : st;
  const unsigned int * id_site = NULL;
  double * staged;
  int ok;
  if (!cp) return PLL_FAILURE;
  if (partition->attributes & PLL_ATTRIB_PATTERN_TIP)
  {
    pll_errno = PLL_ERROR_TIPDATA_ILLEGALFUNCTION;
    snprintf(pll_errmsg, 200, "Cannot use pll_set_tip_clv with PLL_ATTRIB_PATTERN_TIP.");
    return PLL_FAILURE;
  }
  entries = partition->sites;
  if (pll_repeats_enabled(partition))
  {
    entries = partition->repeats->pernode_ids[tip_index];
    sync_ids_to_host(cp, tip_index);
    id_site = partition->repeats->pernode_id_site[tip_index];
  }
  if (!partition->clv[tip_index] || cp->clv_entries[tip_index] < entries)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "tip CLV buffer is not allocated%s", NULL);
    return PLL_FAILURE;
  }
  staged = (double *)calloc((size_t)entries * R * sp, sizeof(double));
  if (!staged)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
    return PLL_FAILURE;
  }
  for (i = 0; i < entries; ++i)
  {
    const double * src = clv + (size_t)(id_site ? id_site[i] : i) * in_states;
    for (j = 0; j < R; ++j) memcpy(staged + ((size_t)i * R + j) * sp, src, st * sizeof(double));
  }
  ok = plf_upload(cp->ctx, partition->clv[tip_index], staged, (size_t)entries * R * sp * sizeof(double));
  free(staged);
  if (ok && partition->asc_bias_alloc) ok = upload_asc_tip_clv(cp, tip_index);
  return ok ? PLL_SUCCESS : cuda_fail(cp);
}

PLL_EXPORT void pll_set_pattern_weights(pll_partition_t * partition, const unsigned int * pattern_weights)
{
  cuda_partition_t * cp = CP(partition);
  unsigned int i;
  memcpy(partition->pattern_weights, pattern_weights, sizeof(unsigned int) * partition->sites);
  partition->pattern_weight_sum = 0;
  for (i = 0; i < partition->sites; ++i) partition->pattern_weight_sum += pattern_weights[i];
  if (cp) cp->weights_dirty = 1;
}

/* src/pll.c:1145-1190 */
PLL_EXPORT int pll_set_asc_bias_type(pll_partition_t * partition, int asc_bias_type)
{
  unsigned int i;
  int prop_invar = 0;
  const int asc_bias_attr = asc_bias_type & PLL_ATTRIB_AB_MASK;
  if (!partition->asc_bias_alloc)
  {
    set_error(PLL_ERROR_AB_NOSUPPORT, "Partition was not created with ascertainment bias support%s", NULL);
    return PLL_FAILURE;
  }
  for (i = 0; i < partition->rate_matrices; ++i) prop_invar |= (partition->prop_invar[i] > 0);
  if (asc_bias_type != 0 && prop_invar)
  {
    set_error(PLL_ERROR_INVAR_INCOMPAT, "Invariant sites are not compatible with asc bias correction%s", NULL);
    return PLL_FAILURE;
  }
  if (asc_bias_attr != asc_bias_type)
  {
    pll_errno = PLL_ERROR_AB_INVALIDMETHOD;
    snprintf(pll_errmsg, 200, "Illegal ascertainment bias algorithm \"%d\"", asc_bias_type);
    return PLL_FAILURE;
  }
  partition->attributes &= (unsigned int)~PLL_ATTRIB_AB_MASK;
  partition->attributes |= (unsigned int)asc_bias_attr;
  return PLL_SUCCESS;
}

/* src/pll.c:1192-1199 */
PLL_EXPORT void pll_set_asc_state_weights(pll_partition_t * partition, const unsigned int * state_weights)
{
  cuda_partition_t * cp = CP(partition);
  if (!partition->asc_bias_alloc)
  {
    set_error(PLL_ERROR_AB_NOSUPPORT, "Partition was not created with ascertainment bias support%s", NULL);
    return;
  }
  memcpy(partition->pattern_weights + partition->sites, state_weights, sizeof(unsigned int) * partition->states);
  if (cp) cp->weights_dirty = 1;
}

PLL_EXPORT void pll_fill_parent_scaler(unsigned int scaler_size, unsigned int * parent_scaler,
                                       const unsigned int * left_scaler, const unsigned int * right_scaler)
{
  unsigned int i;
  for (i = 0; i < scaler_size; ++i)
    parent_scaler[i] = (left_scaler ? left_scaler[i] : 0u) + (right_scaler ? right_scaler[i] : 0u);
}

/* ---- model parameters --------------------------------------------------------- */

PLL_EXPORT void pll_set_frequencies(pll_partition_t * partition, unsigned int params_index,
                                    const double * frequencies)
{
  unsigned int i;
  double sum = 0.;
  double * f = partition->frequencies[params_index];
  memcpy(f, frequencies, partition->states * sizeof(double));
  for (i = 0; i < partition->states; ++i) sum += f[i];
  if (fabs(sum - 1.0) > PLL_MISC_EPSILON)
    for (i = 0; i < partition->states; ++i) f[i] /= sum;
  partition->eigen_decomp_valid[params_index] = 0;
}

PLL_EXPORT void pll_set_category_rates(pll_partition_t * partition, const double * rates)
{
  memcpy(partition->rates, rates, partition->rate_cats * sizeof(double));
}

PLL_EXPORT void pll_set_category_weights(pll_partition_t * partition, const double * rate_weights)
{
  memcpy(partition->rate_weights, rate_weights, partition->rate_cats * sizeof(double));
}

PLL_EXPORT void pll_set_subst_params(pll_partition_t * partition, unsigned int params_index, const double * params)
{
  memcpy(partition->subst_params[params_index], params,
         (size_t)(partition->states * (partition->states - 1) / 2) * sizeof(double));
  partition->eigen_decomp_valid[params_index] = 0;
}

PLL_EXPORT int pll_update_eigen(pll_partition_t * partition, unsigned int params_index)
{
  if (!pll_cuda_host_eigen(partition->states, partition->states_padded, partition->subst_params[params_index],
                           partition->frequencies[params_index], partition->eigenvecs[params_index],
                           partition->inv_eigenvecs[params_index], partition->eigenvals[params_index]))
  {
    set_error(PLL_ERROR_PARAM_INVALID, "eigen-decomposition did not converge%s", NULL);
    return PLL_FAILURE;
  }
  partition->eigen_decomp_valid[params_index] = 1;
  return PLL_SUCCESS;
}

/* pack the per-rate-category model block (layout in plf_backend.h) through
 * `indices` and upload it if any byte differs from what the device holds */
static const double * model_on_device(cuda_partition_t * cp, const unsigned int * indices)
{
  const pll_partition_t * p = &cp->pub;
  const unsigned int R = p->rate_cats, st = p->states, sp = p->states_padded;
  double * m = cp->h_model;
  double * freqs = m + 3 * R;
  double * evals = freqs + (size_t)R * sp;
  double * evecs = evals + (size_t)R * sp;
  double * ievecs = evecs + (size_t)R * st * sp;
  unsigned int r;
  for (r = 0; r < R; ++r)
  {
    const unsigned int x = indices ? indices[r] : 0;
    m[r] = p->rates[r];
    m[R + r] = p->rate_weights[r];
    m[2 * R + r] = p->prop_invar[x];
    memcpy(freqs + (size_t)r * sp, p->frequencies[x], sp * sizeof(double));
    memcpy(evals + (size_t)r * sp, p->eigenvals[x], sp * sizeof(double));
    memcpy(evecs + (size_t)r * st * sp, p->eigenvecs[x], (size_t)st * sp * sizeof(double));
    memcpy(ievecs + (size_t)r * st * sp, p->inv_eigenvecs[x], (size_t)st * sp * sizeof(double));
  }
  if (!cp->model_sent_valid || memcmp(cp->h_model, cp->h_model_sent, cp->model_doubles * sizeof(double)))
  {
    if (!plf_upload_async(cp->ctx, cp->d_model, cp->h_model, cp->model_doubles * sizeof(double))) return NULL;
    memcpy(cp->h_model_sent, cp->h_model, cp->model_doubles * sizeof(double));
    cp->model_sent_valid = 1;
  }
  return cp->d_model;
}

static int weights_on_device(cuda_partition_t * cp)
{
  if (cp->weights_dirty)
  {
    if (!plf_upload(cp->ctx, cp->d_pattern_weights, cp->pub.pattern_weights,
                    (size_t)sites_alloc(&cp->pub) * sizeof(unsigned int)))
      return 0;
    cp->weights_dirty = 0;
  }
  return 1;
}

static int tipmap_on_device(cuda_partition_t * cp)
{
  if (cp->tipmap_dirty && cp->pub.tipmap)
  {
    if (!plf_upload(cp->ctx, cp->d_tipmap, cp->pub.tipmap, PLL_ASCII_SIZE * sizeof(pll_state_t))) return 0;
    cp->tipmap_dirty = 0;
    /* what a code stands for may have changed with it */
    if (cp->tip_expanded_valid) memset(cp->tip_expanded_valid, 0, cp->pub.tips);
  }
  return 1;
}

PLL_EXPORT int pll_cuda_invalidate_host_arrays(pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  cp->weights_dirty = 1;
  cp->tipmap_dirty = 1;
  cp->model_sent_valid = 0;
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_update_prob_matrices(pll_partition_t * partition, const unsigned int * params_indices,
                                        const unsigned int * matrix_indices, const double * branch_lengths,
                                        unsigned int count)
{
  cuda_partition_t * cp = CP(partition);
  const double * d_model;
  double * expd = NULL;
  unsigned int i, n, j;
  int ok;
  if (!cp) return PLL_FAILURE;
  for (n = 0; n < partition->rate_cats; ++n)
    if (!partition->eigen_decomp_valid[params_indices[n]] && !pll_update_eigen(partition, params_indices[n]))
      return PLL_FAILURE;
  for (i = 0; i < count; ++i)
    if (matrix_indices[i] >= partition->prob_matrices)
    {
      set_error(PLL_ERROR_PARAM_INVALID, "P-matrix index out of range%s", NULL);
      return PLL_FAILURE;
    }
  if (!(d_model = model_on_device(cp, params_indices))) return cuda_fail(cp);

  if (cp->host_expm1)
  {
    /* expm1 of (lambda * rate) * t [/ (1 - pinv)] with the host libm, as the
     * reference does (src/core_pmatrix.c:206-216, core_pmatrix_avx.c:99-134):
     * identical inputs to the matrix products => bit-identical P-matrices */
    const unsigned int R = partition->rate_cats, st = partition->states;
    expd = (double *)malloc((size_t)count * R * st * sizeof(double));
    if (!expd)
    {
      set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
      return PLL_FAILURE;
    }
    for (i = 0; i < count; ++i)
      for (n = 0; n < R; ++n)
      {
        const double pinv = partition->prop_invar[params_indices[n]];
        const double * ev = partition->eigenvals[params_indices[n]];
        double * e = expd + ((size_t)i * R + n) * st;
        for (j = 0; j < st; ++j)
        {
          double x = (ev[j] * partition->rates[n]) * branch_lengths[i];
          if (pinv > PLL_MISC_EPSILON) x = x / (1.0 - pinv);
          e[j] = expm1(x);
        }
      }
  }
  ok = plf_update_pmatrices(cp->ctx, &cp->shape, d_model, cp->d_pmatrix_block, matrix_indices, branch_lengths, count,
                            expd);
  free(expd);
  return ok ? PLL_SUCCESS : cuda_fail(cp);
}

/* ---- invariant sites ------------------------------------------------------------ */

static int invariant_on_device(cuda_partition_t * cp, int * d_out)
{
  pll_partition_t * p = &cp->pub;
  const int pattern = (p->attributes & PLL_ATTRIB_PATTERN_TIP) != 0;
  const size_t nb = (size_t)p->tips * sizeof(void *);
  void ** h = (void **)calloc(2 * (size_t)p->tips + 1, sizeof(void *));
  void ** d = (void **)plf_alloc(cp->ctx, 2 * nb + 8, 0);
  unsigned int i;
  int ok = 0, any_ids = 0;
  if (!h || !d) goto done;
  if (pattern)
  {
    if (!cp->d_tipchars) goto done;
    for (i = 0; i < p->tips; ++i) h[i] = cp->d_tipchars[i];
    if (!tipmap_on_device(cp)) goto done;
  }
  else
    for (i = 0; i < p->tips; ++i)
    {
      h[i] = p->clv[i];
      if (p->repeats && p->repeats->pernode_ids[i])
      {
        h[p->tips + i] = cp->d_site_id[i];
        any_ids = 1;
      }
    }
  if (!plf_upload(cp->ctx, d, h, 2 * nb)) goto done;
  ok = plf_invariant_sites(cp->ctx, &cp->shape, p->sites, p->tips, pattern ? (const unsigned char * const *)d : NULL,
                           pattern ? NULL : (const double * const *)d,
                           any_ids ? (const unsigned int * const *)(d + p->tips) : NULL, cp->d_tipmap, d_out);
done:
  if (d) plf_free(cp->ctx, d);
  free(h);
  return ok;
}

int pll_cuda_internal_pick_device(void) { return pick_device(); }

/* tip states as the parsimony kernels read them (pll_parsimony.c); same pointer-table layout as
 * invariant_on_device.  Sets pll_errno on failure. */
int pll_cuda_internal_tipsource(const pll_partition_t * partition, pll_cuda_tipsource_t * out)
{
  cuda_partition_t * cp = CP(partition);
  pll_partition_t * p;
  const size_t nb = cp ? (size_t)cp->pub.tips * sizeof(void *) : 0;
  void ** h = NULL, ** d = NULL;
  unsigned int i;
  int pattern, any_ids = 0;
  if (!cp) return PLL_FAILURE;
  p = &cp->pub;
  pattern = (p->attributes & PLL_ATTRIB_PATTERN_TIP) != 0;
  memset(out, 0, sizeof(*out));
  h = (void **)calloc(2 * (size_t)p->tips + 1, sizeof(void *));
  d = (void **)plf_alloc(cp->ctx, 2 * nb + 8, 0);
  if (!h || !d) goto fail;
  if (pattern)
  {
    if (!cp->d_tipchars) goto fail;
    for (i = 0; i < p->tips; ++i) h[i] = cp->d_tipchars[i];
    if (!tipmap_on_device(cp)) goto fail;
  }
  else
    for (i = 0; i < p->tips; ++i)
    {
      h[i] = p->clv[i];
      if (p->repeats && p->repeats->pernode_ids[i])
      {
        h[p->tips + i] = cp->d_site_id[i];
        any_ids = 1;
      }
    }
  if (!weights_on_device(cp) || !plf_upload(cp->ctx, d, h, 2 * nb) || !plf_sync(cp->ctx)) goto fail;
  free(h);
  out->ctx = cp->ctx;
  out->d_ptrs = d;
  out->tips.tips = p->tips;
  out->tips.sites = p->sites;
  out->tips.states = p->states;
  out->tips.states_padded = p->states_padded;
  out->tips.rate_cats = p->rate_cats;
  out->tips.d_tipchars = pattern ? (const unsigned char * const *)d : NULL;
  out->tips.d_tipclv = pattern ? NULL : (const double * const *)d;
  out->tips.d_tip_site_id = any_ids ? (const unsigned int * const *)(d + p->tips) : NULL;
  out->tips.d_tipmap = cp->d_tipmap;
  out->tips.d_weights = cp->d_pattern_weights;
  return PLL_SUCCESS;
fail:
  if (d) plf_free(cp->ctx, d);
  free(h);
  return cuda_fail(cp);
}

PLL_EXPORT int pll_update_invariant_sites(pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  if (!partition->invariant) partition->invariant = (int *)malloc((size_t)partition->sites * sizeof(int));
  if (!cp->d_invariant) cp->d_invariant = (int *)plf_alloc(cp->ctx, (size_t)partition->sites * sizeof(int), 0);
  if (!partition->invariant || !cp->d_invariant)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate charmap for invariant sites array.%s", NULL);
    return PLL_FAILURE;
  }
  if (!invariant_on_device(cp, cp->d_invariant) ||
      !plf_download(cp->ctx, partition->invariant, cp->d_invariant, (size_t)partition->sites * sizeof(int)))
    return cuda_fail(cp);
  return PLL_SUCCESS;
}

PLL_EXPORT unsigned int pll_count_invariant_sites(pll_partition_t * partition, unsigned int * state_inv_count)
{
  cuda_partition_t * cp = CP(partition);
  unsigned int i, total = 0;
  int * inv, * tmp = NULL;
  if (!cp) return 0;
  if (state_inv_count) memset(state_inv_count, 0, partition->states * sizeof(unsigned int));
  inv = partition->invariant;
  if (!inv)
  {
    int * d_tmp = (int *)plf_alloc(cp->ctx, (size_t)partition->sites * sizeof(int), 0);
    tmp = (int *)malloc((size_t)partition->sites * sizeof(int));
    if (!d_tmp || !tmp || !invariant_on_device(cp, d_tmp) ||
        !plf_download(cp->ctx, tmp, d_tmp, (size_t)partition->sites * sizeof(int)))
    {
      plf_free(cp->ctx, d_tmp);
      free(tmp);
      cuda_fail(cp);
      return 0;
    }
    plf_free(cp->ctx, d_tmp);
    inv = tmp;
  }
  for (i = 0; i < partition->sites; ++i)
    if (inv[i] > -1)
    {
      total += partition->pattern_weights[i];
      if (state_inv_count) state_inv_count[inv[i]]++;
    }
  free(tmp);
  return total;
}

PLL_EXPORT int pll_update_invariant_sites_proportion(pll_partition_t * partition, unsigned int params_index,
                                                     double prop_invar)
{
  if (prop_invar != 0.0 && (partition->attributes & PLL_ATTRIB_AB_MASK))
  {
    set_error(PLL_ERROR_INVAR_INCOMPAT, "Invariant sites are not compatible with asc bias correction%s", NULL);
    return PLL_FAILURE;
  }
  if (prop_invar < 0 || prop_invar >= 1)
  {
    pll_errno = PLL_ERROR_INVAR_PROPORTION;
    snprintf(pll_errmsg, 200, "Invalid proportion of invariant sites (%f)", prop_invar);
    return PLL_FAILURE;
  }
  if (params_index > partition->rate_matrices)
  {
    pll_errno = PLL_ERROR_INVAR_PARAMINDEX;
    snprintf(pll_errmsg, 200, "Invalid params index (%u)", params_index);
    return PLL_FAILURE;
  }
  if (prop_invar > 0.0 && !partition->invariant && !pll_update_invariant_sites(partition))
  {
    pll_errno = PLL_ERROR_INVAR_NONEFOUND;
    snprintf(pll_errmsg, 200, "No invariant sites found");
    return PLL_FAILURE;
  }
  partition->prop_invar[params_index] = prop_invar;
  return PLL_SUCCESS;
}

/* ---- level schedule ------------------------------------------------------------ */

/* Ops are strictly sequential in the reference (src/partials.c:253).  Here an
 * op goes into the earliest launch level that keeps every read-after-write,
 * write-after-read and write-after-write order on CLV and scaler indices. */
PLL_EXPORT int pll_cuda_schedule_levels(const pll_operation_t * operations, unsigned int count,
                                        unsigned int * level_of_op)
{
  unsigned int i, max_clv = 0, max_sc = 0, nlevels = 0;
  int * clv_w, * clv_r, * sc_w, * sc_r;
  if (!count) return 0;
  for (i = 0; i < count; ++i)
  {
    const pll_operation_t * o = operations + i;
    if (o->parent_clv_index > max_clv) max_clv = o->parent_clv_index;
    if (o->child1_clv_index > max_clv) max_clv = o->child1_clv_index;
    if (o->child2_clv_index > max_clv) max_clv = o->child2_clv_index;
    if (o->parent_scaler_index > (int)max_sc) max_sc = (unsigned int)o->parent_scaler_index;
    if (o->child1_scaler_index > (int)max_sc) max_sc = (unsigned int)o->child1_scaler_index;
    if (o->child2_scaler_index > (int)max_sc) max_sc = (unsigned int)o->child2_scaler_index;
  }
  /* level of the last write / of the latest read since that write, -1 = none */
  clv_w = (int *)malloc(((size_t)max_clv + 1) * sizeof(int));
  clv_r = (int *)malloc(((size_t)max_clv + 1) * sizeof(int));
  sc_w = (int *)malloc(((size_t)max_sc + 1) * sizeof(int));
  sc_r = (int *)malloc(((size_t)max_sc + 1) * sizeof(int));
  if (!clv_w || !clv_r || !sc_w || !sc_r)
  {
    free(clv_w);
    free(clv_r);
    free(sc_w);
    free(sc_r);
    return -1;
  }
  for (i = 0; i <= max_clv; ++i) clv_w[i] = clv_r[i] = -1;
  for (i = 0; i <= max_sc; ++i) sc_w[i] = sc_r[i] = -1;
#define AFTER(x) do { if ((x) + 1 > lv) lv = (x) + 1; } while (0)
  for (i = 0; i < count; ++i)
  {
    const pll_operation_t * o = operations + i;
    int lv = 0;
    AFTER(clv_w[o->child1_clv_index]);
    AFTER(clv_w[o->child2_clv_index]);
    if (o->child1_scaler_index >= 0) AFTER(sc_w[o->child1_scaler_index]);
    if (o->child2_scaler_index >= 0) AFTER(sc_w[o->child2_scaler_index]);
    AFTER(clv_w[o->parent_clv_index]);
    AFTER(clv_r[o->parent_clv_index]);
    if (o->parent_scaler_index >= 0)
    {
      AFTER(sc_w[o->parent_scaler_index]);
      AFTER(sc_r[o->parent_scaler_index]);
    }
    level_of_op[i] = (unsigned int)lv;
    if ((unsigned int)lv + 1 > nlevels) nlevels = (unsigned int)lv + 1;
    if (clv_r[o->child1_clv_index] < lv) clv_r[o->child1_clv_index] = lv;
    if (clv_r[o->child2_clv_index] < lv) clv_r[o->child2_clv_index] = lv;
    if (o->child1_scaler_index >= 0 && sc_r[o->child1_scaler_index] < lv) sc_r[o->child1_scaler_index] = lv;
    if (o->child2_scaler_index >= 0 && sc_r[o->child2_scaler_index] < lv) sc_r[o->child2_scaler_index] = lv;
    clv_w[o->parent_clv_index] = lv;
    clv_r[o->parent_clv_index] = -1;
    if (o->parent_scaler_index >= 0)
    {
      sc_w[o->parent_scaler_index] = lv;
      sc_r[o->parent_scaler_index] = -1;
    }
  }
#undef AFTER
  free(clv_w);
  free(clv_r);
  free(sc_w);
  free(sc_r);
  return (int)nlevels;
}

/* NEW (additive, inspection).  The kernel launches one traversal level makes for ops of the given kinds
 * (PLF_OP_* values, already sorted by kind as launch_levels does): a launch serves a run of same-kind ops,
 * blockIdx.y selecting the op, so a run is cut after PLF_MAX_RUN_OPS ops.  Returns the number of launches,
 * the longest run in *largest_run.  Pure host arithmetic: callable without a device. */
PLL_EXPORT unsigned int pll_cuda_count_launch_runs(const unsigned int * kinds, unsigned int count,
                                                   unsigned int * largest_run)
{
  plf_op_t * ops = (plf_op_t *)calloc(count ? count : 1, sizeof(plf_op_t));
  unsigned int i, runs = 0, largest = 0;
  if (!ops) return 0;
  for (i = 0; i < count; ++i) ops[i].kind = kinds[i];
  for (i = 0; i < count; ++runs)
  {
    const unsigned int j = plf_run_end(ops, i, count, NULL, NULL);
    if (j - i > largest) largest = j - i;
    i = j;
  }
  free(ops);
  if (largest_run) *largest_run = largest;
  return runs;
}

/* ---- CLV updates ---------------------------------------------------------------- */

static int reserve_ops(cuda_partition_t * cp, unsigned int count)
{
  if (count <= cp->ops_cap) return 1;
  free(cp->h_ops);
  free(cp->h_ops_sorted);
  free(cp->h_level);
  free(cp->h_level_start);
  cp->h_ops = (plf_op_t *)malloc((size_t)count * sizeof(plf_op_t));
  cp->h_ops_sorted = (plf_op_t *)malloc((size_t)count * sizeof(plf_op_t));
  cp->h_level = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  cp->h_level_start = (unsigned int *)malloc(((PLF_OP_KINDS + 1) * (size_t)count + 2) * sizeof(unsigned int));
  cp->ops_cap = (cp->h_ops && cp->h_ops_sorted && cp->h_level && cp->h_level_start) ? count : 0;
  return cp->ops_cap != 0;
}

/* the expanded CLV of a pattern tip (entry j = bit j of the state mask of its code, for every rate: what
 * pll_set_tip_states leaves in a partition without PLL_ATTRIB_PATTERN_TIP, src/pll.c:959-1024) */
static const double * expanded_tip(cuda_partition_t * cp, unsigned int tip)
{
  const pll_partition_t * p = &cp->pub;
  if (!cp->d_tip_expanded)
  {
    cp->d_tip_expanded = (double **)calloc(p->tips ? p->tips : 1, sizeof(double *));
    cp->tip_expanded_valid = (unsigned char *)calloc(p->tips ? p->tips : 1, 1);
    if (!cp->d_tip_expanded || !cp->tip_expanded_valid) return NULL;
  }
  if (!cp->d_tip_expanded[tip])
    cp->d_tip_expanded[tip] = (double *)plf_alloc(
        cp->ctx, (size_t)sites_alloc(p) * p->rate_cats * p->states_padded * sizeof(double) + BULK_PAD, 0);
  if (!cp->d_tip_expanded[tip]) return NULL;
  if (!cp->tip_expanded_valid[tip])
  {
    if (!tipmap_on_device(cp) ||
        !plf_tip_clv_from_states(cp->ctx, &cp->shape, cp->d_tip_expanded[tip], cp->d_tipchars[tip], cp->d_tipmap, NULL,
                                 sites_alloc(p)))
      return NULL;
    cp->tip_expanded_valid[tip] = 1;
  }
  return cp->d_tip_expanded[tip];
}

/* turn one pll_operation_t into device pointers + kernel variant
 * (dispatch of src/partials.c:245-291) */
static int resolve_op(cuda_partition_t * cp, const pll_operation_t * op, plf_op_t * out)
{
  const pll_partition_t * p = &cp->pub;
  const unsigned int c1 = op->child1_clv_index, c2 = op->child2_clv_index, par = op->parent_clv_index;
  unsigned int * const * sb = p->scale_buffer;
  memset(out, 0, sizeof(*out));
  /* who writes what this op reads is not known until attach_dependencies() has seen the whole list: until then
   * the op counts as ordered by its launch level only */
  out->dep[0] = PLF_DEP_ORDERED;
  out->dep[1] = out->dep[2] = out->dep[3] = PLF_DEP_NONE;
  if (par >= p->nodes || c1 >= p->nodes || c2 >= p->nodes || op->child1_matrix_index >= p->prob_matrices ||
      op->child2_matrix_index >= p->prob_matrices || op->parent_scaler_index >= (int)p->scale_buffers ||
      op->child1_scaler_index >= (int)p->scale_buffers || op->child2_scaler_index >= (int)p->scale_buffers)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "operation refers to a buffer index out of range%s", NULL);
    return 0;
  }
  out->parent_clv = p->clv[par];
  out->parent_scaler = op->parent_scaler_index >= 0 ? sb[op->parent_scaler_index] : NULL;
  out->nsites = sites_alloc(p); /* src/partials.c:33-35: the pseudo-sites are updated with the alignment */
  if (pll_repeats_enabled(p))
  {
    const pll_repeats_t * r = p->repeats;
    out->kind = PLF_OP_II;
    out->left_clv = p->clv[c1];
    out->right_clv = p->clv[c2];
    out->left_matrix = p->pmatrix[op->child1_matrix_index];
    out->right_matrix = p->pmatrix[op->child2_matrix_index];
    out->left_scaler = op->child1_scaler_index >= 0 ? sb[op->child1_scaler_index] : NULL;
    out->right_scaler = op->child2_scaler_index >= 0 ? sb[op->child2_scaler_index] : NULL;
    if (r->pernode_ids[par])
    {
      out->parent_id_site = cp->d_id_site[par];
      out->nsites = r->pernode_ids[par];
    }
    if (r->pernode_ids[c1]) out->left_site_id = cp->d_site_id[c1];
    if (r->pernode_ids[c2]) out->right_site_id = cp->d_site_id[c2];
  }
  else
  {
    const int pattern = (p->attributes & PLL_ATTRIB_PATTERN_TIP) != 0;
    const int t1 = pattern && c1 < p->tips, t2 = pattern && c2 < p->tips;
    /* a child whose CLV is virtual: computed from its two tips' codes by this op's kernel */
    const int v1 = !t1 && is_virtual(cp, c1), v2 = !t2 && is_virtual(cp, c2);
    if (t1 && t2)
    {
      out->kind = PLF_OP_TT;
      out->left_tip = cp->d_tipchars ? cp->d_tipchars[c1] : NULL;
      out->right_tip = cp->d_tipchars ? cp->d_tipchars[c2] : NULL;
      out->left_matrix = p->pmatrix[op->child1_matrix_index];
      out->right_matrix = p->pmatrix[op->child2_matrix_index];
      if (!out->left_tip || !out->right_tip) goto missing;
      if (cp->cherry_ok && par >= p->tips && p->clv[par])
      {
        /* the cherry stays virtual: only its P-matrices are snapshot and its scaler zeroed */
        out->kind = PLF_OP_TT_VIRTUAL;
        out->parent_clv = cherry_snapshot(cp, par);
        /* the scaler of a tip-tip parent is all zeros: written once, not again while nothing else wrote it */
        if (op->parent_scaler_index >= 0)
        {
          if (cp->scaler_zero[op->parent_scaler_index]) out->parent_scaler = NULL;
          cp->scaler_zero[op->parent_scaler_index] = 1;
        }
        if (!cp->cherry[par].is_virtual) ++cp->cherries_pending;
        cp->cherry[par].is_virtual = 1;
        cp->cherry[par].tip1 = c1;
        cp->cherry[par].tip2 = c2;
        cp->cherry[par].scaler_index = op->parent_scaler_index;
        return 1;
      }
    }
    else if (t1 || t2)
    {
      /* the tip is always passed as "left" (src/partials.c:68-128) */
      const unsigned int tip = t1 ? c1 : c2, inner = t1 ? c2 : c1;
      const unsigned int mt = t1 ? op->child1_matrix_index : op->child2_matrix_index;
      const unsigned int mi = t1 ? op->child2_matrix_index : op->child1_matrix_index;
      const int si = t1 ? op->child2_scaler_index : op->child1_scaler_index;
      out->left_tip = cp->d_tipchars ? cp->d_tipchars[tip] : NULL;
      out->left_matrix = p->pmatrix[mt];
      out->right_matrix = p->pmatrix[mi];
      if (!out->left_tip) goto missing;
      const double * expanded = (cp->aa_tip_clvs && !(t1 ? v2 : v1)) ? expanded_tip(cp, tip) : NULL;
      if (cp->aa_tip_clvs && !(t1 ? v2 : v1) && !expanded) cp->aa_tip_clvs = 0; /* no memory for the copies: tip kernels */
      if (expanded)
      {
        /* narrow 20-state alignment: the tip as an expanded CLV, the op as inner-inner (scales like tip-inner:
         * src/core_partials_avx2.c:343 vs :630; a tip has no scaler) */
        out->kind = PLF_OP_II;
        out->left_tip = NULL;
        out->left_clv = expanded;
        out->right_clv = p->clv[inner];
        out->right_scaler = si >= 0 ? sb[si] : NULL;
        if (!out->right_clv) goto missing;
      }
      else if (t1 ? v2 : v1)
      {
        const cherry_state_t * c = &cp->cherry[inner];
        out->kind = PLF_OP_TC;
        out->right_tip = cp->d_tipchars[c->tip1];
        out->right_tip2 = cp->d_tipchars[c->tip2];
        out->right_cm1 = cherry_snapshot(cp, inner);
        out->right_cm2 = out->right_cm1 + cherry_msz(cp);
      }
      else
      {
        out->kind = PLF_OP_TI;
        out->right_clv = p->clv[inner];
        out->right_scaler = si >= 0 ? sb[si] : NULL;
        if (!out->right_clv) goto missing;
      }
    }
    else if (v1 || v2)
    {
      /* cherries go "left"; with one of them the other child is an inner CLV on the "right"
       * (parent entry = left term * right term, commutative to the bit) */
      const int swap = !v1;
      const unsigned int cl = swap ? c2 : c1, cr = swap ? c1 : c2;
      const unsigned int ml = swap ? op->child2_matrix_index : op->child1_matrix_index;
      const unsigned int mr = swap ? op->child1_matrix_index : op->child2_matrix_index;
      const int sr = swap ? op->child1_scaler_index : op->child2_scaler_index;
      const cherry_state_t * c = &cp->cherry[cl];
      out->left_tip = cp->d_tipchars[c->tip1];
      out->left_tip2 = cp->d_tipchars[c->tip2];
      out->left_cm1 = cherry_snapshot(cp, cl);
      out->left_cm2 = out->left_cm1 + cherry_msz(cp);
      out->left_matrix = p->pmatrix[ml];
      out->right_matrix = p->pmatrix[mr];
      if (v1 && v2)
      {
        const cherry_state_t * d = &cp->cherry[cr];
        out->kind = PLF_OP_CC;
        out->right_tip = cp->d_tipchars[d->tip1];
        out->right_tip2 = cp->d_tipchars[d->tip2];
        out->right_cm1 = cherry_snapshot(cp, cr);
        out->right_cm2 = out->right_cm1 + cherry_msz(cp);
      }
      else
      {
        out->kind = PLF_OP_CI;
        out->right_clv = p->clv[cr];
        out->right_scaler = sr >= 0 ? sb[sr] : NULL;
        if (!out->right_clv) goto missing;
      }
    }
    else
    {
      out->kind = PLF_OP_II;
      out->left_clv = p->clv[c1];
      out->right_clv = p->clv[c2];
      out->left_matrix = p->pmatrix[op->child1_matrix_index];
      out->right_matrix = p->pmatrix[op->child2_matrix_index];
      out->left_scaler = op->child1_scaler_index >= 0 ? sb[op->child1_scaler_index] : NULL;
      out->right_scaler = op->child2_scaler_index >= 0 ? sb[op->child2_scaler_index] : NULL;
    }
  }
  /* this op writes the parent's buffer: whatever cherry it stood for is gone */
  if (cp->cherry && par < p->nodes && cp->cherry[par].is_virtual)
  {
    cp->cherry[par].is_virtual = 0;
    --cp->cherries_pending;
  }
  /* only the tip-tip kernels leave a scaler all zeros */
  if (cp->scaler_zero && op->parent_scaler_index >= 0) cp->scaler_zero[op->parent_scaler_index] = (out->kind == PLF_OP_TT);
  if (!out->parent_clv || (out->kind == PLF_OP_II && (!out->left_clv || !out->right_clv))) goto missing;
  return 1;
missing:
  set_error(PLL_ERROR_PARAM_INVALID, "operation refers to a CLV or tip buffer that was never set%s", NULL);
  return 0;
}

/* Pair lists for the gathering operations of a site-repeats list (4 states): entry n of the parent reads
 * entry pair[n].x / pair[n].y of its children.  A list is kept per parent node and rebuilt only when the
 * identifiers of the parent or of a child changed, or the node is fed from other children: the traversals
 * between two identifier updates (update_repeats = 0: branch-length and model optimisation) reuse it. */
static int attach_pair_lists(cuda_partition_t * cp, const pll_operation_t * ops, unsigned int count)
{
  const pll_partition_t * p = &cp->pub;
  plf_pair_job_t * jobs = NULL;
  unsigned int i, njobs = 0;
  int ok = 1;
  if (!cp->pairs || p->states != 4 || env_flag("PLL_CUDA_NO_PAIR_LISTS")) return 1;
  for (i = 0; i < count; ++i)
  {
    plf_op_t * o = cp->h_ops + i;
    const unsigned int par = ops[i].parent_clv_index, c1 = ops[i].child1_clv_index, c2 = ops[i].child2_clv_index;
    pair_state_t * ps;
    if (o->kind != PLF_OP_II || !(o->parent_id_site || o->left_site_id || o->right_site_id)) continue;
    ps = &cp->pairs[par];
    if (!(ps->valid && ps->entries == o->nsites && ps->c1 == c1 && ps->c2 == c2 && ps->vp == cp->ids_version[par] &&
          ps->v1 == cp->ids_version[c1] && ps->v2 == cp->ids_version[c2]))
    {
      if (!jobs && !(jobs = (plf_pair_job_t *)malloc((size_t)count * sizeof(plf_pair_job_t)))) return 0;
      if (ps->cap < o->nsites)
      {
        plf_free(cp->ctx, ps->dev);
        ps->dev = (unsigned int *)plf_alloc(cp->ctx, (size_t)o->nsites * 2 * sizeof(unsigned int), 0);
        ps->cap = ps->dev ? o->nsites : 0;
        if (!ps->dev)
        {
          ps->valid = 0;
          ok = 0;
          break;
        }
      }
      jobs[njobs].parent_id_site = o->parent_id_site;
      jobs[njobs].left_site_id = o->left_site_id;
      jobs[njobs].right_site_id = o->right_site_id;
      jobs[njobs].out = ps->dev;
      jobs[njobs].entries = o->nsites;
      ++njobs;
      ps->entries = o->nsites;
      ps->c1 = c1;
      ps->c2 = c2;
      ps->vp = cp->ids_version[par];
      ps->v1 = cp->ids_version[c1];
      ps->v2 = cp->ids_version[c2];
      ps->valid = 1;
    }
    o->pair_list = ps->dev;
  }
  if (ok && njobs && !plf_repeats_pairs(cp->ctx, jobs, njobs)) ok = 0;
  free(jobs);
  return ok;
}

/* Who writes what each op reads (plf_op_t.dep): lets a narrow alignment's whole traversal run as ONE kernel
 * whose work items wait for their producers' items instead of for a launch boundary.  cp->h_level[i] holds
 * the position of ops[i] in cp->h_ops_sorted.  A list that recycles buffers (an op overwrites something an
 * earlier op read or wrote) is marked PLF_DEP_ORDERED and keeps the launch levels. */
static void dependencies_of_list(const pll_operation_t * ops, unsigned int count, size_t nclv, size_t nsc,
                                 const unsigned int * position, plf_op_t * sorted)
{
  int * writer = (int *)malloc((nclv + nsc + 1) * sizeof(int));
  unsigned char * was_read = (unsigned char *)calloc(nclv + nsc + 1, 1);
  unsigned int i;
  int ordered = (!writer || !was_read);
  if (!ordered)
  {
    size_t k;
    for (k = 0; k < nclv + nsc; ++k) writer[k] = PLF_DEP_NONE;
    for (i = 0; i < count && !ordered; ++i)
    {
      const pll_operation_t * o = ops + i;
      plf_op_t * s = sorted + position[i];
      const size_t par = o->parent_clv_index, c1 = o->child1_clv_index, c2 = o->child2_clv_index;
      const int psc = o->parent_scaler_index, sc1 = o->child1_scaler_index, sc2 = o->child2_scaler_index;
      if (writer[par] != PLF_DEP_NONE || was_read[par] || par == c1 || par == c2 ||
          (psc >= 0 && (writer[nclv + psc] != PLF_DEP_NONE || was_read[nclv + psc] || psc == sc1 || psc == sc2)))
      {
        ordered = 1;
        break;
      }
      s->dep[0] = writer[c1];
      s->dep[1] = writer[c2];
      s->dep[2] = (sc1 >= 0 && writer[nclv + sc1] != s->dep[0]) ? writer[nclv + sc1] : PLF_DEP_NONE;
      s->dep[3] = (sc2 >= 0 && writer[nclv + sc2] != s->dep[1]) ? writer[nclv + sc2] : PLF_DEP_NONE;
      was_read[c1] = was_read[c2] = 1;
      if (sc1 >= 0) was_read[nclv + sc1] = 1;
      if (sc2 >= 0) was_read[nclv + sc2] = 1;
      writer[par] = (int)position[i];
      if (psc >= 0) writer[nclv + psc] = (int)position[i];
    }
  }
  if (ordered)
    for (i = 0; i < count; ++i)
    {
      plf_op_t * s = sorted + i;
      s->dep[0] = PLF_DEP_ORDERED;
      s->dep[1] = s->dep[2] = s->dep[3] = PLF_DEP_NONE;
    }
  free(writer);
  free(was_read);
}

static void attach_dependencies(cuda_partition_t * cp, const pll_operation_t * ops, unsigned int count)
{
  dependencies_of_list(ops, count, cp->pub.nodes, cp->pub.scale_buffers, cp->h_level, cp->h_ops_sorted);
}

/* NEW (additive, inspection).  How k_clv_dna_flow would run this list on a partition with `tips` pattern tips:
 * path_of_op[i] = position in the queue of the path ops[i] belongs to, carried_child_of_op[i] = 1 / 2 when the
 * CLV of child1 / child2 reaches ops[i] in registers (0: both children come from memory or are tips).  Returns the
 * number of paths; 0 when the list keeps the launch levels (it recycles a buffer, or reads a scaler another op than
 * the CLV's writer wrote).  Pure host arithmetic: callable without a device. */
PLL_EXPORT unsigned int pll_cuda_schedule_paths(const pll_operation_t * operations, unsigned int count, unsigned int tips,
                                                unsigned int path_max, unsigned int * path_of_op,
                                                int * carried_child_of_op)
{
#define FAKE_CLV(i) ((double *)(uintptr_t)(((size_t)(i) + 1) << 12))
  unsigned int i, nlevels, npaths = 0, max_clv = 0;
  int max_sc = -1, nl;
  unsigned int * level, * start, * position, * pstart, * op_of_clv;
  plf_op_t * sorted;
  struct plf_flow_op * plan;
  if (!count) return 0;
  for (i = 0; i < count; ++i)
  {
    const pll_operation_t * o = operations + i;
    if (o->parent_clv_index > max_clv) max_clv = o->parent_clv_index;
    if (o->child1_clv_index > max_clv) max_clv = o->child1_clv_index;
    if (o->child2_clv_index > max_clv) max_clv = o->child2_clv_index;
    if (o->parent_scaler_index > max_sc) max_sc = o->parent_scaler_index;
    if (o->child1_scaler_index > max_sc) max_sc = o->child1_scaler_index;
    if (o->child2_scaler_index > max_sc) max_sc = o->child2_scaler_index;
  }
  level = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  position = (unsigned int *)malloc((size_t)count * sizeof(unsigned int));
  pstart = (unsigned int *)malloc(((size_t)count + 1) * sizeof(unsigned int));
  op_of_clv = (unsigned int *)malloc(((size_t)max_clv + 1) * sizeof(unsigned int));
  sorted = (plf_op_t *)calloc(count, sizeof(plf_op_t));
  plan = (struct plf_flow_op *)malloc((size_t)count * sizeof(struct plf_flow_op));
  start = NULL;
  nl = (level && position && pstart && op_of_clv && sorted && plan) ? pll_cuda_schedule_levels(operations, count, level) : -1;
  if (nl > 0) start = (unsigned int *)calloc((size_t)nl * (PLF_OP_KINDS + 1) + 2, sizeof(unsigned int));
  if (start)
  {
    nlevels = (unsigned int)nl * (PLF_OP_KINDS + 1);
    for (i = 0; i < count; ++i)
    {
      const pll_operation_t * o = operations + i;
      const int t1 = o->child1_clv_index < tips, t2 = o->child2_clv_index < tips;
      level[i] = level[i] * (PLF_OP_KINDS + 1) + ((t1 && t2) ? PLF_OP_TT : ((t1 || t2) ? PLF_OP_TI : PLF_OP_II));
      start[level[i] + 1]++;
    }
    for (i = 0; i < nlevels; ++i) start[i + 1] += start[i];
    for (i = 0; i < count; ++i)
    {
      /* as resolve_op: a pattern tip goes to the left */
      const pll_operation_t * o = operations + i;
      const int t1 = o->child1_clv_index < tips, t2 = o->child2_clv_index < tips;
      const unsigned int first = (t2 && !t1) ? o->child2_clv_index : o->child1_clv_index;
      const unsigned int second = (t2 && !t1) ? o->child1_clv_index : o->child2_clv_index;
      plf_op_t * s;
      position[i] = start[level[i]]++;
      s = sorted + position[i];
      s->kind = level[i] % (PLF_OP_KINDS + 1);
      s->nsites = 1;
      s->parent_clv = FAKE_CLV(o->parent_clv_index);
      if (s->kind == PLF_OP_II) s->left_clv = FAKE_CLV(first); else s->left_tip = (const unsigned char *)FAKE_CLV(first);
      if (s->kind == PLF_OP_TT) s->right_tip = (const unsigned char *)FAKE_CLV(second); else s->right_clv = FAKE_CLV(second);
      s->dep[0] = s->dep[1] = s->dep[2] = s->dep[3] = PLF_DEP_NONE;
    }
    dependencies_of_list(operations, count, (size_t)max_clv + 1, (size_t)(max_sc + 1), position, sorted);
    if (sorted[0].dep[0] != PLF_DEP_ORDERED) npaths = plf_dna_flow_plan(sorted, count, path_max, plan, pstart);
    if (npaths)
    {
      unsigned int path = 0, k;
      for (i = 0; i < count; ++i) op_of_clv[operations[i].parent_clv_index] = i;
      for (k = 0; k < count; ++k)
      {
        const unsigned int clv = (unsigned int)(((uintptr_t)plan[k].parent_clv >> 12) - 1);
        const unsigned int op = op_of_clv[clv];
        const unsigned int side = plan[k].flags & 3u;
        while (k >= pstart[path + 1]) ++path;
        if (path_of_op) path_of_op[op] = path;
        if (carried_child_of_op)
        {
          carried_child_of_op[op] = 0;
          if (side)
          {
            /* the carried side holds no pointer check here: it is the child written by the previous op of the path */
            const unsigned int prev = op_of_clv[(unsigned int)(((uintptr_t)plan[k - 1].parent_clv >> 12) - 1)];
            carried_child_of_op[op] =
                operations[op].child1_clv_index == operations[prev].parent_clv_index ? 1 : 2;
          }
        }
      }
    }
  }
  free(level);
  free(start);
  free(position);
  free(pstart);
  free(op_of_clv);
  free(sorted);
  free(plan);
  return npaths;
#undef FAKE_CLV
}

static int launch_levels(cuda_partition_t * cp, const pll_operation_t * ops, unsigned int count)
{
  unsigned int i, nlevels, saved_pending = 0;
  int nl;
  if (!reserve_ops(cp, count))
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
    return 0;
  }
  if (!cherry_decide(cp)) return 0;
  /* ops are resolved in list order: whether a child is a virtual cherry is the state the list has
   * reached at that op; a list that fails to resolve leaves the state as it found it */
  if (cp->cherry)
  {
    memcpy(cp->cherry_saved, cp->cherry, (size_t)cp->pub.nodes * sizeof(cherry_state_t));
    memcpy(cp->scaler_zero_saved, cp->scaler_zero, cp->pub.scale_buffers);
    saved_pending = cp->cherries_pending;
  }
  for (i = 0; i < count; ++i)
    if (!resolve_op(cp, ops + i, cp->h_ops + i))
    {
      if (cp->cherry)
      {
        memcpy(cp->cherry, cp->cherry_saved, (size_t)cp->pub.nodes * sizeof(cherry_state_t));
        memcpy(cp->scaler_zero, cp->scaler_zero_saved, cp->pub.scale_buffers);
        cp->cherries_pending = saved_pending;
      }
      return 0;
    }
  if (pll_repeats_enabled(&cp->pub) && !attach_pair_lists(cp, ops, count))
  {
    cuda_fail(cp);
    return 0;
  }
  nl = pll_cuda_schedule_levels(ops, count, cp->h_level);
  if (nl < 0)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
    return 0;
  }
  nlevels = (unsigned int)nl;
  /* counting sort by (level, op kind), stable: each launch group is a run of
   * same-kind ops of one level */
  /* (under site repeats the gathering inner-inner ops of a level form a run of their own, next to the ones
   * whose parent and children are all uncompressed: different kernels) */
  for (i = 0; i < count; ++i)
  {
    const plf_op_t * o = cp->h_ops + i;
    const unsigned int slot =
        (o->kind == PLF_OP_II && (o->parent_id_site || o->left_site_id || o->right_site_id)) ? PLF_OP_KINDS : o->kind;
    cp->h_level[i] = cp->h_level[i] * (PLF_OP_KINDS + 1) + slot;
  }
  nlevels *= PLF_OP_KINDS + 1;
  for (i = 0; i <= nlevels; ++i) cp->h_level_start[i] = 0;
  for (i = 0; i < count; ++i) cp->h_level_start[cp->h_level[i] + 1]++;
  for (i = 0; i < nlevels; ++i) cp->h_level_start[i + 1] += cp->h_level_start[i];
  {
    unsigned int * cursor = (unsigned int *)malloc(((size_t)nlevels + 1) * sizeof(unsigned int));
    if (!cursor) return 0;
    memcpy(cursor, cp->h_level_start, ((size_t)nlevels + 1) * sizeof(unsigned int));
    for (i = 0; i < count; ++i)
    {
      const unsigned int pos = cursor[cp->h_level[i]]++;
      cp->h_ops_sorted[pos] = cp->h_ops[i];
      cp->h_level[i] = pos; /* from here on: the op's position in the sorted list */
    }
    free(cursor);
  }
  /* (under site repeats too: a list whose nodes all went without identifiers is a plain list) */
  if (count > 1 && cp->shape.states == 4) attach_dependencies(cp, ops, count);
  if (!tipmap_on_device(cp) ||
      !plf_update_partials(cp->ctx, &cp->shape, cp->h_ops_sorted, count, cp->h_level_start, nlevels, cp->d_tipmap,
                           cp->pub.maxstates))
  {
    cuda_fail(cp);
    return 0;
  }
  return 1;
}

/* does a later op overwrite a CLV that an earlier op of the list touched? */
static int reuses_buffers(const pll_operation_t * ops, unsigned int count, unsigned int nodes)
{
  unsigned char * seen = (unsigned char *)calloc(nodes ? nodes : 1, 1);
  unsigned int i;
  int reuse = 0;
  if (!seen) return 1;
  for (i = 0; i < count && !reuse; ++i)
  {
    if (ops[i].parent_clv_index < nodes && seen[ops[i].parent_clv_index]) reuse = 1;
    if (ops[i].parent_clv_index == ops[i].child1_clv_index || ops[i].parent_clv_index == ops[i].child2_clv_index)
      reuse = 1; /* identifiers are numbered in place */
    if (ops[i].parent_clv_index < nodes) seen[ops[i].parent_clv_index] = 1;
    if (ops[i].child1_clv_index < nodes) seen[ops[i].child1_clv_index] = 1;
    if (ops[i].child2_clv_index < nodes) seen[ops[i].child2_clv_index] = 1;
  }
  free(seen);
  return reuse;
}

PLL_EXPORT void pll_update_partials_rep(pll_partition_t * partition, const pll_operation_t * operations,
                                        unsigned int count, unsigned int update_repeats)
{
  cuda_partition_t * cp = CP(partition);
  unsigned int i;
  if (!cp || !count) return;
  ++cp->partials_generation;
  if (pll_repeats_enabled(partition) && update_repeats)
  {
    if (reuses_buffers(operations, count, partition->nodes))
    {
      /* identifier arrays are overwritten in place: keep the reference's
       * strict op order when the list recycles CLV indices */
      for (i = 0; i < count; ++i)
      {
        pll_update_repeats(partition, operations + i);
        if (!launch_levels(cp, operations + i, 1)) return;
      }
      return;
    }
    /* identifiers depend on the children's identifiers only, not on CLV
     * values: compute them for the whole list first (one batch of launches
     * and one host synchronisation per level), then run the CLV levels */
    if (cp->rid_fast && (partition->repeats->enable_repeats == pll_default_enable_repeats ||
                         partition->repeats->enable_repeats == pll_no_enable_repeats))
    {
      if (!update_repeats_fast(cp, operations, count)) return;
    }
    else if (!update_repeats_levels(cp, operations, count))
      return;
  }
  launch_levels(cp, operations, count);
}

PLL_EXPORT void pll_update_partials(pll_partition_t * partition, const pll_operation_t * operations,
                                    unsigned int count)
{
  pll_update_partials_rep(partition, operations, count, 1);
}

/* ---- log-likelihood ---------------------------------------------------------------- */

static double * persite_buffer(cuda_partition_t * cp)
{
  if (cp->persite_cap < cp->pub.sites)
  {
    plf_free(cp->ctx, cp->d_persite);
    cp->d_persite = (double *)plf_alloc(cp->ctx, (size_t)cp->pub.sites * sizeof(double), 0);
    cp->persite_cap = cp->d_persite ? cp->pub.sites : 0;
  }
  return cp->d_persite;
}

static int fill_edge_args(cuda_partition_t * cp, plf_lk_t * a, unsigned int parent_clv_index,
                          int parent_scaler_index, unsigned int child_clv_index, int child_scaler_index,
                          unsigned int matrix_index, const unsigned int * freqs_indices)
{
  const pll_partition_t * p = &cp->pub;
  unsigned int * const * sb = p->scale_buffer;
  memset(a, 0, sizeof(*a));
  if (parent_clv_index >= p->nodes || child_clv_index >= p->nodes || matrix_index >= p->prob_matrices ||
      parent_scaler_index >= (int)p->scale_buffers || child_scaler_index >= (int)p->scale_buffers)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "buffer index out of range%s", NULL);
    return 0;
  }
  if (!ensure_real(cp, parent_clv_index) || !ensure_real(cp, child_clv_index)) return 0;
  a->sites = p->sites;
  a->pmatrix = p->pmatrix[matrix_index];
  a->pattern_weights = cp->d_pattern_weights;
  a->invariant = cp->d_invariant;
  if ((p->attributes & PLL_ATTRIB_PATTERN_TIP) && (parent_clv_index < p->tips || child_clv_index < p->tips))
  {
    /* tip-inner: the inner node plays "parent" (src/likelihood.c:612-624) */
    const int ptip = parent_clv_index < p->tips;
    const unsigned int inner = ptip ? child_clv_index : parent_clv_index;
    const unsigned int tip = ptip ? parent_clv_index : child_clv_index;
    const int sc = ptip ? child_scaler_index : parent_scaler_index;
    if (inner < p->tips || !cp->d_tipchars)
    {
      set_error(PLL_ERROR_PARAM_INVALID, "edge log-likelihood between two pattern tips is not defined%s", NULL);
      return 0;
    }
    a->clvp = p->clv[inner];
    a->pscaler = sc >= 0 ? sb[sc] : NULL;
    a->tipchars = cp->d_tipchars[tip];
    a->tipmap = cp->d_tipmap;
    a->maxstates = p->maxstates;
    if (!tipmap_on_device(cp)) return 0;
  }
  else
  {
    a->clvp = p->clv[parent_clv_index];
    a->clvc = p->clv[child_clv_index];
    a->pscaler = parent_scaler_index >= 0 ? sb[parent_scaler_index] : NULL;
    a->cscaler = child_scaler_index >= 0 ? sb[child_scaler_index] : NULL;
    if (pll_repeats_enabled(p))
    {
      if (p->repeats->pernode_ids[parent_clv_index]) a->p_site_id = cp->d_site_id[parent_clv_index];
      if (p->repeats->pernode_ids[child_clv_index]) a->c_site_id = cp->d_site_id[child_clv_index];
    }
    if (!a->clvc)
    {
      set_error(PLL_ERROR_PARAM_INVALID, "child CLV was never set%s", NULL);
      return 0;
    }
  }
  if (!a->clvp)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "parent CLV was never set%s", NULL);
    return 0;
  }
  if (!weights_on_device(cp) || !(a->model = model_on_device(cp, freqs_indices)))
  {
    cuda_fail(cp);
    return 0;
  }
  return 1;
}

static int fill_root_args(cuda_partition_t * cp, plf_lk_t * a, unsigned int clv_index, int scaler_index,
                          const unsigned int * freqs_indices)
{
  const pll_partition_t * p = &cp->pub;
  memset(a, 0, sizeof(*a));
  if (clv_index >= p->nodes || scaler_index >= (int)p->scale_buffers || !p->clv[clv_index])
  {
    set_error(PLL_ERROR_PARAM_INVALID, "root CLV index out of range or never set%s", NULL);
    return 0;
  }
  if (!ensure_real(cp, clv_index)) return 0;
  a->sites = p->sites;
  a->clvp = p->clv[clv_index];
  a->pscaler = scaler_index >= 0 ? p->scale_buffer[scaler_index] : NULL;
  a->pattern_weights = cp->d_pattern_weights;
  a->invariant = cp->d_invariant;
  if (pll_repeats_enabled(p) && p->repeats->pernode_ids[clv_index]) a->p_site_id = cp->d_site_id[clv_index];
  if (!weights_on_device(cp) || !(a->model = model_on_device(cp, freqs_indices)))
  {
    cuda_fail(cp);
    return 0;
  }
  return 1;
}


/* ---- ascertainment bias correction ------------------------------------------------------- *
 * The `states` pseudo-sites travel through the CLV kernels with the alignment; what is left    *
 * is O(states^2 x rates) scalar work per evaluation (src/likelihood.c:24-120,190-268,342-440;  *
 * src/core_derivatives.c:851-924).  It runs here on the host from a small download of the      *
 * pseudo-site blocks, in the reference's own evaluation order.                                  */

static double asc_correction(const pll_partition_t * p, double base, unsigned int sum_w_inv)
{
  switch (p->attributes & PLL_ATTRIB_AB_MASK)
  {
    case PLL_ATTRIB_AB_LEWIS: return -(p->pattern_weight_sum * log(1 - base));
    case PLL_ATTRIB_AB_STAMATAKIS: return base;
    case PLL_ATTRIB_AB_FELSENSTEIN: return sum_w_inv * log(base);
    default:
      set_error(PLL_ERROR_AB_INVALIDMETHOD, "Illegal ascertainment bias algorithm%s", NULL);
      return -INFINITY;
  }
}

/* pseudo-site blocks of a CLV ([states][R][sp]) and of a scale buffer ([states], zero when none) */
static int asc_download(cuda_partition_t * cp, unsigned int first, unsigned int clv_index, int scaler_index,
                        double * clv_out, unsigned int * scaler_out)
{
  const pll_partition_t * p = &cp->pub;
  const size_t blk = (size_t)p->rate_cats * p->states_padded;
  if (clv_out && !plf_download(cp->ctx, clv_out, p->clv[clv_index] + (size_t)first * blk,
                               (size_t)p->states * blk * sizeof(double)))
    return 0;
  memset(scaler_out, 0, p->states * sizeof(unsigned int));
  if (scaler_index >= 0 &&
      !plf_download(cp->ctx, scaler_out, p->scale_buffer[scaler_index] + first, p->states * sizeof(unsigned int)))
    return 0;
  return 1;
}

/* child_clv_index < 0: root; child_is_tip: pattern tip (pseudo-site n shows state n) */
static double asc_loglikelihood(cuda_partition_t * cp, unsigned int parent_clv_index, int parent_scaler_index,
                                int child_clv_index, int child_scaler_index, int child_is_tip, int matrix_index,
                                const unsigned int * freqs_indices)
{
  const pll_partition_t * p = &cp->pub;
  const unsigned int st = p->states, sp = p->states_padded, R = p->rate_cats;
  const size_t blk = (size_t)R * sp;
  const int type = p->attributes & PLL_ATTRIB_AB_MASK;
  double * clvp = (double *)malloc(st * blk * sizeof(double));
  double * clvc = (double *)malloc(st * blk * sizeof(double));
  double * pm = (double *)malloc((size_t)R * st * sp * sizeof(double));
  unsigned int * psc = (unsigned int *)malloc(st * sizeof(unsigned int));
  unsigned int * csc = (unsigned int *)malloc(st * sizeof(unsigned int));
  /* the root variant of the reference locates the pseudo-sites at pll_get_sites_number() - rate_cats
   * = sites + states - rate_cats (src/likelihood.c:180), which is `sites` only when states ==
   * rate_cats.  With more rates than states it reads the last alignment sites instead: reproduced,
   * the parity contract is the reference's output.  With fewer rates than states the reference reads
   * past the end of its buffers (undefined); here the pseudo-sites themselves are used. */
  const unsigned int first = (child_clv_index < 0 && R > st && p->sites + st >= R) ? p->sites + st - R : p->sites;
  const unsigned int * w = p->pattern_weights + first;
  double logl_correction = 0, result = -INFINITY;
  unsigned int sum_w_inv = 0, n, i, j, k;
  if (!clvp || !clvc || !pm || !psc || !csc)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.%s", NULL);
    goto done;
  }
  if (!asc_download(cp, first, parent_clv_index, parent_scaler_index, clvp, psc)) goto fail;
  memset(csc, 0, st * sizeof(unsigned int));
  if (child_clv_index >= 0 && !child_is_tip &&
      !asc_download(cp, first, (unsigned int)child_clv_index, child_scaler_index, clvc, csc))
    goto fail;
  if (matrix_index >= 0 && !plf_download(cp->ctx, pm, p->pmatrix[matrix_index], (size_t)R * st * sp * sizeof(double)))
    goto fail;
  for (n = 0; n < st; ++n)
  {
    double term = 0, site_lk;
    unsigned int scale_factors;
    for (i = 0; i < R; ++i)
    {
      const double * freqs = p->frequencies[freqs_indices[i]];
      const double * cp_ = clvp + n * blk + (size_t)i * sp;
      const double * cc_ = clvc + n * blk + (size_t)i * sp;
      const double * pmat = pm + (size_t)i * st * sp;
      double term_r = 0;
      for (j = 0; j < st; ++j)
      {
        if (child_clv_index < 0)
          term_r += cp_[j] * freqs[j];                           /* likelihood.c:84-87 */
        else if (child_is_tip)
          term_r += cp_[j] * freqs[j] * pmat[(size_t)j * sp + n]; /* likelihood.c:231-235 */
        else
        {
          double termb = 0;
          for (k = 0; k < st; ++k) termb += pmat[(size_t)j * sp + k] * cc_[k];
          term_r += cp_[j] * freqs[j] * termb;                   /* likelihood.c:396-404 */
        }
      }
      term += term_r * p->rate_weights[i];
    }
    scale_factors = psc[n] + csc[n];
    sum_w_inv += w[n];
    if (type == PLL_ATTRIB_AB_STAMATAKIS)
    {
      site_lk = log(term) * w[n];
      if (scale_factors) site_lk += scale_factors * log(PLL_SCALE_THRESHOLD);
    }
    else
      site_lk = term * pow(PLL_SCALE_THRESHOLD, scale_factors);
    logl_correction += site_lk;
  }
  result = asc_correction(p, logl_correction, sum_w_inv);
  goto done;
fail:
  cuda_fail(cp);
done:
  free(clvp);
  free(clvc);
  free(pm);
  free(psc);
  free(csc);
  return result;
}

static double run_loglikelihood(cuda_partition_t * cp, plf_lk_t * a, double * persite_lnl)
{
  double logl = 0;
  if (persite_lnl && !(a->persite = persite_buffer(cp)))
  {
    cuda_fail(cp);
    return -INFINITY;
  }
  if (!plf_loglikelihood(cp->ctx, &cp->shape, a, NULL, &logl) ||
      (persite_lnl && !plf_download(cp->ctx, persite_lnl, a->persite, (size_t)cp->pub.sites * sizeof(double))))
  {
    cuda_fail(cp);
    return -INFINITY;
  }
  return logl;
}

PLL_EXPORT double pll_compute_edge_loglikelihood(pll_partition_t * partition, unsigned int parent_clv_index,
                                                 int parent_scaler_index, unsigned int child_clv_index,
                                                 int child_scaler_index, unsigned int matrix_index,
                                                 const unsigned int * freqs_indices, double * persite_lnl)
{
  cuda_partition_t * cp = CP(partition);
  plf_lk_t a;
  double logl;
  if (!cp || !fill_edge_args(cp, &a, parent_clv_index, parent_scaler_index, child_clv_index, child_scaler_index,
                             matrix_index, freqs_indices))
    return -INFINITY;
  logl = run_loglikelihood(cp, &a, persite_lnl);
  if (partition->attributes & PLL_ATTRIB_AB_MASK)
  {
    /* src/likelihood.c:325-337, 569-581: the inner node plays "parent" on a pattern-tip edge */
    if (a.tipchars)
    {
      const int ptip = parent_clv_index < partition->tips;
      logl += asc_loglikelihood(cp, ptip ? child_clv_index : parent_clv_index,
                                ptip ? child_scaler_index : parent_scaler_index, 0, -1, 1, (int)matrix_index,
                                freqs_indices);
    }
    else
      logl += asc_loglikelihood(cp, parent_clv_index, parent_scaler_index, (int)child_clv_index, child_scaler_index,
                                0, (int)matrix_index, freqs_indices);
  }
  return logl;
}

PLL_EXPORT double pll_compute_root_loglikelihood(pll_partition_t * partition, unsigned int clv_index,
                                                 int scaler_index, const unsigned int * freqs_indices,
                                                 double * persite_lnl)
{
  cuda_partition_t * cp = CP(partition);
  plf_lk_t a;
  double logl;
  if (!cp || !fill_root_args(cp, &a, clv_index, scaler_index, freqs_indices)) return -INFINITY;
  logl = run_loglikelihood(cp, &a, persite_lnl);
  if (partition->attributes & PLL_ATTRIB_AB_MASK) /* src/likelihood.c:176-188 */
    logl += asc_loglikelihood(cp, clv_index, scaler_index, -1, -1, 0, -1, freqs_indices);
  return logl;
}

PLL_EXPORT int pll_cuda_edge_loglikelihood_async(pll_partition_t * partition, unsigned int parent_clv_index,
                                                 int parent_scaler_index, unsigned int child_clv_index,
                                                 int child_scaler_index, unsigned int matrix_index,
                                                 const unsigned int * freqs_indices, double * dev_out)
{
  cuda_partition_t * cp = CP(partition);
  plf_lk_t a;
  if (cp && (partition->attributes & PLL_ATTRIB_AB_MASK))
  {
    set_error(PLL_ERROR_CUDA_UNSUPPORTED, "the asynchronous entry points do not apply the ascertainment bias correction%s", NULL);
    return PLL_FAILURE;
  }
  if (!cp || !dev_out ||
      !fill_edge_args(cp, &a, parent_clv_index, parent_scaler_index, child_clv_index, child_scaler_index,
                      matrix_index, freqs_indices))
    return PLL_FAILURE;
  return plf_loglikelihood(cp->ctx, &cp->shape, &a, dev_out, NULL) ? PLL_SUCCESS : cuda_fail(cp);
}

PLL_EXPORT int pll_cuda_root_loglikelihood_async(pll_partition_t * partition, unsigned int clv_index,
                                                 int scaler_index, const unsigned int * freqs_indices,
                                                 double * dev_out)
{
  cuda_partition_t * cp = CP(partition);
  plf_lk_t a;
  if (cp && (partition->attributes & PLL_ATTRIB_AB_MASK))
  {
    set_error(PLL_ERROR_CUDA_UNSUPPORTED, "the asynchronous entry points do not apply the ascertainment bias correction%s", NULL);
    return PLL_FAILURE;
  }
  if (!cp || !dev_out || !fill_root_args(cp, &a, clv_index, scaler_index, freqs_indices)) return PLL_FAILURE;
  return plf_loglikelihood(cp->ctx, &cp->shape, &a, dev_out, NULL) ? PLL_SUCCESS : cuda_fail(cp);
}

/* ---- ancestral states --------------------------------------------------------------------- */

/* src/likelihood.c:639-760: CLV of a virtual root placed ON `node` (identity matrix towards the
 * node, P-matrix towards `other`), then per-site posterior state probabilities.  The caller's
 * scratch buffers of the _extbuf variant are host memory and are not needed here: the temporary
 * CLV, scaler and identity matrix live in HBM for the duration of the call. */
PLL_EXPORT int pll_compute_node_ancestral_extbuf(pll_partition_t * partition, unsigned int node_clv_index,
                                                 int node_scaler_index, unsigned int other_clv_index,
                                                 int other_scaler_index, unsigned int pmatrix_index,
                                                 const unsigned int * freqs_indices, double * ancestral,
                                                 double * temp_clv, unsigned int * temp_scaler, double * ident_pmat)
{
  cuda_partition_t * cp = CP(partition);
  const pll_partition_t * p = partition;
  unsigned int st, sp, R, i, j;
  size_t clv_doubles, sc_entries, pm_doubles;
  double * d_tmp = NULL, * d_ident = NULL, * d_anc = NULL, * h_ident = NULL;
  unsigned int * d_sc = NULL;
  const double * d_model;
  plf_op_t op;
  unsigned int level_start[2] = {0, 1};
  int ok = 0;
  if (!partition || !ancestral)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "Parameter value is NULL!%s", NULL);
    return PLL_FAILURE;
  }
  if (!temp_clv || !temp_scaler || !ident_pmat)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "NULL buffer pointer%s", NULL);
    return PLL_FAILURE;
  }
  if (!cp) return PLL_FAILURE;
  if (pll_repeats_enabled(partition))
  {
    set_error(PLL_ERROR_EINVAL, "Site repeats are not compatible with ancestral state reconstruction!%s", NULL);
    return PLL_FAILURE;
  }
  if (node_clv_index >= p->nodes || other_clv_index >= p->nodes || pmatrix_index >= p->prob_matrices ||
      node_scaler_index >= (int)p->scale_buffers || other_scaler_index >= (int)p->scale_buffers ||
      !p->clv[node_clv_index])
  {
    set_error(PLL_ERROR_PARAM_INVALID, "buffer index out of range or CLV never set%s", NULL);
    return PLL_FAILURE;
  }
  if (!ensure_real(cp, node_clv_index) || !ensure_real(cp, other_clv_index)) return PLL_FAILURE;
  st = p->states;
  sp = p->states_padded;
  R = p->rate_cats;
  clv_doubles = (size_t)sites_alloc(p) * R * sp;
  sc_entries = (size_t)sites_alloc(p) * ((p->attributes & PLL_ATTRIB_RATE_SCALERS) ? R : 1);
  pm_doubles = (size_t)R * st * sp + (size_t)(sp - st) * sp;
  h_ident = (double *)calloc(pm_doubles, sizeof(double));
  d_tmp = (double *)plf_alloc(cp->ctx, clv_doubles * sizeof(double), 0);
  d_sc = (unsigned int *)plf_alloc(cp->ctx, sc_entries * sizeof(unsigned int) + BULK_PAD, 1);
  d_ident = (double *)plf_alloc(cp->ctx, pm_doubles * sizeof(double), 0);
  d_anc = (double *)plf_alloc(cp->ctx, (size_t)p->sites * st * sizeof(double), 0);
  if (!h_ident || !d_tmp || !d_sc || !d_ident || !d_anc)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate memory%s", NULL);
    goto done;
  }
  for (i = 0; i < R; ++i)
    for (j = 0; j < st; ++j) h_ident[((size_t)i * st + j) * sp + j] = 1.0;
  if (!plf_upload(cp->ctx, d_ident, h_ident, pm_doubles * sizeof(double))) goto fail;

  memset(&op, 0, sizeof(op));
  op.parent_clv = d_tmp;
  op.parent_scaler = d_sc;
  op.nsites = sites_alloc(p);
  if (other_clv_index < p->tips && (p->attributes & PLL_ATTRIB_PATTERN_TIP))
  {
    /* src/likelihood.c:692-706: the tip side takes the P-matrix, the node the identity */
    if (!cp->d_tipchars || !cp->d_tipchars[other_clv_index] || !tipmap_on_device(cp))
    {
      set_error(PLL_ERROR_PARAM_INVALID, "tip states were never set%s", NULL);
      goto done;
    }
    op.kind = PLF_OP_TI;
    op.left_tip = cp->d_tipchars[other_clv_index];
    op.left_matrix = p->pmatrix[pmatrix_index];
    op.right_clv = p->clv[node_clv_index];
    op.right_matrix = d_ident;
    op.right_scaler = node_scaler_index >= 0 ? p->scale_buffer[node_scaler_index] : NULL;
  }
  else
  {
    if (!p->clv[other_clv_index])
    {
      set_error(PLL_ERROR_PARAM_INVALID, "CLV was never set%s", NULL);
      goto done;
    }
    op.kind = PLF_OP_II;
    op.left_clv = p->clv[node_clv_index];
    op.left_matrix = d_ident;
    op.left_scaler = node_scaler_index >= 0 ? p->scale_buffer[node_scaler_index] : NULL;
    op.right_clv = p->clv[other_clv_index];
    op.right_matrix = p->pmatrix[pmatrix_index];
    op.right_scaler = other_scaler_index >= 0 ? p->scale_buffer[other_scaler_index] : NULL;
  }
  if (!(d_model = model_on_device(cp, freqs_indices)) ||
      !plf_update_partials(cp->ctx, &cp->shape, &op, 1, level_start, 1, cp->d_tipmap, p->maxstates) ||
      !plf_ancestral(cp->ctx, &cp->shape, d_tmp, d_model, p->sites, d_anc) ||
      !plf_download(cp->ctx, ancestral, d_anc, (size_t)p->sites * st * sizeof(double)))
    goto fail;
  ok = 1;
  goto done;
fail:
  cuda_fail(cp);
done:
  free(h_ident);
  if (cp)
  {
    plf_free(cp->ctx, d_tmp);
    plf_free(cp->ctx, d_sc);
    plf_free(cp->ctx, d_ident);
    plf_free(cp->ctx, d_anc);
  }
  return ok ? PLL_SUCCESS : PLL_FAILURE;
}

/* src/likelihood.c:762-823 */
PLL_EXPORT int pll_compute_node_ancestral(pll_partition_t * partition, unsigned int node_clv_index,
                                          int node_scaler_index, unsigned int other_clv_index,
                                          int other_scaler_index, unsigned int matrix_index,
                                          const unsigned int * freqs_indices, double * ancestral)
{
  double dummy_clv = 0, dummy_pmat = 0;
  unsigned int dummy_scaler = 0;
  return pll_compute_node_ancestral_extbuf(partition, node_clv_index, node_scaler_index, other_clv_index,
                                           other_scaler_index, matrix_index, freqs_indices, ancestral, &dummy_clv,
                                           &dummy_scaler, &dummy_pmat);
}

/* ---- sumtable and derivatives ---------------------------------------------------------- */

static int key_was_evicted(const cuda_partition_t * cp, const double * key)
{
  unsigned int i;
  for (i = 0; i < cp->n_evicted; ++i)
    if (cp->evicted_keys[i] == key) return 1;
  return 0;
}

static void forget_evicted(cuda_partition_t * cp, const double * key)
{
  unsigned int i;
  for (i = 0; i < cp->n_evicted; ++i)
    if (cp->evicted_keys[i] == key)
    {
      cp->evicted_keys[i] = cp->evicted_keys[--cp->n_evicted];
      return;
    }
}

/* The table of `victim` leaves the device.  Unless the host mirror is kept ($PLL_CUDA_SUMTABLE_MIRROR) the
 * caller's buffer was never written, so the key is remembered: a derivative call on it fails with a clear
 * error instead of reading those bytes (the reference writes every table into the caller's buffer,
 * src/derivatives.c:239-330, so there any number of tables can be live). */
static void remember_evicted(cuda_partition_t * cp, const double * key)
{
  if (cp->sumtable_mirror || !key || key_was_evicted(cp, key)) return;
  if (cp->n_evicted == cp->cap_evicted)
  {
    const unsigned int cap = cp->cap_evicted ? 2 * cp->cap_evicted : 64;
    const double ** grown = (const double **)realloc((void *)cp->evicted_keys, (size_t)cap * sizeof(*grown));
    if (!grown) return;
    cp->evicted_keys = grown;
    cp->cap_evicted = cap;
  }
  cp->evicted_keys[cp->n_evicted++] = key;
}

static sumtable_slot_t * sumtable_slot(cuda_partition_t * cp, const double * key, int create)
{
  const size_t need = (size_t)sites_alloc(&cp->pub) * cp->pub.rate_cats * cp->pub.states_padded;
  sumtable_slot_t * victim = NULL;
  unsigned int i;
  for (i = 0; i < cp->n_sumtabs; ++i)
    if (cp->sumtabs[i].dev && cp->sumtabs[i].key == key)
    {
      cp->sumtabs[i].stamp = ++cp->stamp;
      return &cp->sumtabs[i];
    }
  if (!create) return NULL;
  if (!cp->max_sumtabs)
  {
    const char * v = getenv("PLL_CUDA_MAX_SUMTABLES");
    cp->max_sumtabs = (v && atoi(v) > 0) ? (unsigned int)atoi(v) : DEFAULT_MAX_SUMTABLES;
  }
  for (i = 0; i < cp->n_sumtabs && !victim; ++i)
    if (!cp->sumtabs[i].dev) victim = &cp->sumtabs[i];
  if (!victim && cp->n_sumtabs < cp->max_sumtabs)
  {
    sumtable_slot_t * grown = (sumtable_slot_t *)realloc(cp->sumtabs, ((size_t)cp->n_sumtabs + 1) * sizeof(*grown));
    if (grown)
    {
      cp->sumtabs = grown;
      victim = &cp->sumtabs[cp->n_sumtabs++];
      memset(victim, 0, sizeof(*victim));
      victim->dev = (double *)plf_alloc(cp->ctx, need * sizeof(double), 0);
      victim->doubles = victim->dev ? need : 0;
      if (!victim->dev)
      {
        /* HBM is full: fall back to reusing the least recently used table's memory */
        --cp->n_sumtabs;
        victim = NULL;
      }
    }
  }
  if (!victim)
  {
    for (i = 0; i < cp->n_sumtabs; ++i)
      if (cp->sumtabs[i].dev && (!victim || cp->sumtabs[i].stamp < victim->stamp)) victim = &cp->sumtabs[i];
    if (!victim) return NULL;
    remember_evicted(cp, victim->key);
  }
  if (victim->dev && victim->doubles < need)
  {
    plf_free(cp->ctx, victim->dev);
    victim->dev = NULL;
  }
  if (!victim->dev)
  {
    victim->dev = (double *)plf_alloc(cp->ctx, need * sizeof(double), 0);
    victim->doubles = victim->dev ? need : 0;
    if (!victim->dev) return NULL;
  }
  victim->key = key;
  victim->stamp = ++cp->stamp;
  forget_evicted(cp, key);
  return victim;
}

PLL_EXPORT int pll_update_sumtable(pll_partition_t * partition, unsigned int parent_clv_index,
                                   unsigned int child_clv_index, int parent_scaler_index, int child_scaler_index,
                                   const unsigned int * params_indices, double * sumtable)
{
  cuda_partition_t * cp = CP(partition);
  const pll_partition_t * p = partition;
  plf_sumtable_t a;
  sumtable_slot_t * slot;
  unsigned int * const * sb;
  if (!cp) return PLL_FAILURE;
  sb = p->scale_buffer;
  memset(&a, 0, sizeof(a));
  if (parent_clv_index >= p->nodes || child_clv_index >= p->nodes || parent_scaler_index >= (int)p->scale_buffers ||
      child_scaler_index >= (int)p->scale_buffers || !sumtable)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "buffer index out of range%s", NULL);
    return PLL_FAILURE;
  }
  if (!ensure_real(cp, parent_clv_index) || !ensure_real(cp, child_clv_index)) return PLL_FAILURE;
  a.sites = sites_alloc(p); /* src/derivatives.c:56-58,131-133,191-193 */
  if ((p->attributes & PLL_ATTRIB_PATTERN_TIP) && (parent_clv_index < p->tips || child_clv_index < p->tips))
  {
    const int ptip = parent_clv_index < p->tips;
    const unsigned int inner = ptip ? child_clv_index : parent_clv_index;
    const unsigned int tip = ptip ? parent_clv_index : child_clv_index;
    const int sc = ptip ? child_scaler_index : parent_scaler_index;
    if (inner < p->tips)
    {
      set_error(PLL_ERROR_PARAM_INVALID, "pll_update_sumtable() was called for the tip-tip case!%s", NULL);
      return PLL_FAILURE;
    }
    /* the tip takes the pi * V^-1 side, the inner CLV the V side
     * (src/derivatives.c:24-98, src/core_derivatives.c:473) */
    a.tipchars = cp->d_tipchars ? cp->d_tipchars[tip] : NULL;
    a.tipmap = cp->d_tipmap;
    a.maxstates = p->maxstates;
    a.clvc = p->clv[inner];
    a.cscaler = sc >= 0 ? sb[sc] : NULL;
    if (!a.tipchars || !tipmap_on_device(cp))
    {
      set_error(PLL_ERROR_PARAM_INVALID, "tip states were never set%s", NULL);
      return PLL_FAILURE;
    }
  }
  else
  {
    a.clvp = p->clv[parent_clv_index];
    a.clvc = p->clv[child_clv_index];
    a.pscaler = parent_scaler_index >= 0 ? sb[parent_scaler_index] : NULL;
    a.cscaler = child_scaler_index >= 0 ? sb[child_scaler_index] : NULL;
    if (pll_repeats_enabled(p))
    {
      if (p->repeats->pernode_ids[parent_clv_index]) a.p_site_id = cp->d_site_id[parent_clv_index];
      if (p->repeats->pernode_ids[child_clv_index]) a.c_site_id = cp->d_site_id[child_clv_index];
    }
    if (!a.clvp)
    {
      set_error(PLL_ERROR_PARAM_INVALID, "parent CLV was never set%s", NULL);
      return PLL_FAILURE;
    }
  }
  if (!a.clvc)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "child CLV was never set%s", NULL);
    return PLL_FAILURE;
  }
  if (!(slot = sumtable_slot(cp, sumtable, 1)) || !(a.model = model_on_device(cp, params_indices))) return cuda_fail(cp);
  a.sumtable = slot->dev;
  if (!plf_update_sumtable(cp->ctx, &cp->shape, &a)) return cuda_fail(cp);
  if (cp->sumtable_mirror && !plf_download(cp->ctx, sumtable, slot->dev,
                                           (size_t)sites_alloc(p) * p->rate_cats * p->states_padded * sizeof(double)))
    return cuda_fail(cp);
  if (p->asc_bias_alloc)
  {
    /* the derivative calls need the pseudo-site blocks on the host (core_derivatives.c:851-924) */
    const size_t blk = (size_t)p->rate_cats * p->states_padded;
    if (!slot->asc_host) slot->asc_host = (double *)malloc(p->states * blk * sizeof(double));
    if (!slot->asc_host || !plf_download(cp->ctx, slot->asc_host, slot->dev + (size_t)p->sites * blk,
                                         p->states * blk * sizeof(double)))
      return cuda_fail(cp);
  }
  return PLL_SUCCESS;
}

static int derivative_args(cuda_partition_t * cp, plf_deriv_t * a, double branch_length,
                           const unsigned int * params_indices, const double * sumtable)
{
  const pll_partition_t * p = &cp->pub;
  sumtable_slot_t * slot = sumtable_slot(cp, sumtable, 0);
  memset(a, 0, sizeof(*a));
  if (!slot && sumtable && key_was_evicted(cp, sumtable))
  {
    pll_errno = PLL_ERROR_PARAM_INVALID;
    snprintf(pll_errmsg, sizeof(pll_errmsg),
             "this sumtable was dropped from the device (more than %u live tables): call pll_update_sumtable "
             "again or raise PLL_CUDA_MAX_SUMTABLES", cp->max_sumtabs);
    return 0;
  }
  if (!slot)
  {
    /* a table this library did not compute: the host bytes are the data */
    if (!sumtable || !(slot = sumtable_slot(cp, sumtable, 1)) ||
        !plf_upload(cp->ctx, slot->dev, sumtable,
                    (size_t)sites_alloc(p) * p->rate_cats * p->states_padded * sizeof(double)))
    {
      cuda_fail(cp);
      return 0;
    }
    if (p->asc_bias_alloc)
    {
      const size_t blk = (size_t)p->rate_cats * p->states_padded;
      if (!slot->asc_host) slot->asc_host = (double *)malloc(p->states * blk * sizeof(double));
      if (!slot->asc_host) return 0;
      memcpy(slot->asc_host, sumtable + (size_t)p->sites * blk, p->states * blk * sizeof(double));
    }
  }
  /* Stamatakis: the pseudo-sites are ordinary weighted sites of the sums (core_derivatives.c:733-741) */
  a->sites = p->sites + (((p->attributes & PLL_ATTRIB_AB_MASK) == PLL_ATTRIB_AB_STAMATAKIS) ? p->states : 0);
  a->sumtable = slot->dev;
  a->pattern_weights = cp->d_pattern_weights;
  a->invariant = cp->d_invariant;
  a->branch_length = branch_length;
  if (!weights_on_device(cp) || !(a->model = model_on_device(cp, params_indices)))
  {
    cuda_fail(cp);
    return 0;
  }
  return 1;
}

/* Lewis / Felsenstein terms of the derivatives (src/core_derivatives.c:851-924) from the host copy of
 * the pseudo-site sums; the pseudo-site scalers are cached until the next pll_update_partials */
static int asc_derivatives(cuda_partition_t * cp, int parent_scaler_index, int child_scaler_index, double branch_length,
                           const unsigned int * params_indices, const double * sumtable, double * d_f, double * dd_f)
{
  const pll_partition_t * p = &cp->pub;
  const unsigned int st = p->states, sp = p->states_padded, R = p->rate_cats;
  const int type = p->attributes & PLL_ATTRIB_AB_MASK;
  const sumtable_slot_t * slot = sumtable_slot(cp, sumtable, 0);
  double asc_Lk[3] = {0.0, 0.0, 0.0};
  unsigned int sum_w_inv = 0, n, i, j;
  double * diag;
  if (!slot || !slot->asc_host)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "sumtable was not computed for this partition%s", NULL);
    return PLL_FAILURE;
  }
  if (!cp->asc_sc_valid || cp->asc_sc_parent != parent_scaler_index || cp->asc_sc_child != child_scaler_index ||
      cp->asc_sc_generation != cp->partials_generation)
  {
    if (!cp->asc_sc) cp->asc_sc = (unsigned int *)malloc((size_t)2 * st * sizeof(unsigned int));
    if (!cp->asc_sc || !asc_download(cp, p->sites, 0, parent_scaler_index, NULL, cp->asc_sc) ||
        !asc_download(cp, p->sites, 0, child_scaler_index, NULL, cp->asc_sc + st))
      return cuda_fail(cp);
    cp->asc_sc_parent = parent_scaler_index;
    cp->asc_sc_child = child_scaler_index;
    cp->asc_sc_generation = cp->partials_generation;
    cp->asc_sc_valid = 1;
  }
  diag = (double *)malloc((size_t)R * st * 4 * sizeof(double));
  if (!diag)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate memory for diagptable%s", NULL);
    return PLL_FAILURE;
  }
  for (i = 0; i < R; ++i)
  {
    const double * ev = p->eigenvals[params_indices[i]];
    const double ki = p->rates[i] / (1.0 - p->prop_invar[params_indices[i]]);
    for (j = 0; j < st; ++j)
    {
      double * d = diag + ((size_t)i * st + j) * 4;
      d[0] = exp(ev[j] * ki * branch_length);
      d[1] = ev[j] * ki * d[0];
      d[2] = ev[j] * ki * ev[j] * ki * d[0];
      d[3] = 0;
    }
  }
  for (n = 0; n < st; ++n)
  {
    const double * sum = slot->asc_host + (size_t)n * R * sp;
    double site_lk[3] = {0, 0, 0};
    double asc_scaling;
    for (i = 0; i < R; ++i)
    {
      const double * d = diag + (size_t)i * st * 4;
      double c[3] = {0, 0, 0};
      for (j = 0; j < st; ++j)
      {
        c[0] += sum[j] * d[4 * j + 0];
        c[1] += sum[j] * d[4 * j + 1];
        c[2] += sum[j] * d[4 * j + 2];
      }
      site_lk[0] += c[0] * p->rate_weights[i];
      site_lk[1] += c[1] * p->rate_weights[i];
      site_lk[2] += c[2] * p->rate_weights[i];
      sum += sp;
    }
    asc_scaling = pow(PLL_SCALE_THRESHOLD, (double)(cp->asc_sc[n] + cp->asc_sc[st + n]));
    asc_Lk[0] += site_lk[0] * asc_scaling;
    asc_Lk[1] += site_lk[1] * asc_scaling;
    asc_Lk[2] += site_lk[2] * asc_scaling;
    sum_w_inv += p->pattern_weights[p->sites + n];
  }
  free(diag);
  if (type == PLL_ATTRIB_AB_LEWIS)
  {
    unsigned int pattern_weight_sum = 0;
    for (n = 0; n < p->sites; ++n) pattern_weight_sum += p->pattern_weights[n];
    *d_f += pattern_weight_sum * (asc_Lk[1] / (asc_Lk[0] - 1.0));
    *dd_f += pattern_weight_sum *
             (((asc_Lk[0] - 1.0) * asc_Lk[2] - asc_Lk[1] * asc_Lk[1]) / ((asc_Lk[0] - 1.0) * (asc_Lk[0] - 1.0)));
  }
  else if (type == PLL_ATTRIB_AB_FELSENSTEIN)
  {
    *d_f -= sum_w_inv * (asc_Lk[1] / asc_Lk[0]);
    *dd_f -= sum_w_inv * (((asc_Lk[2] * asc_Lk[0]) - asc_Lk[1] * asc_Lk[1]) / (asc_Lk[0] * asc_Lk[0]));
  }
  else
  {
    set_error(PLL_ERROR_AB_INVALIDMETHOD, "Illegal ascertainment bias algorithm%s", NULL);
    return PLL_FAILURE;
  }
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_compute_likelihood_derivatives(pll_partition_t * partition, int parent_scaler_index,
                                                  int child_scaler_index, double branch_length,
                                                  const unsigned int * params_indices, const double * sumtable,
                                                  double * d_f, double * dd_f)
{
  cuda_partition_t * cp = CP(partition);
  plf_deriv_t a;
  double out[2] = {0, 0};
  /* per-site ratios cancel the scaling (src/core_derivatives.c:825-848); only the ascertainment
   * bias terms need the scalers */
  if (!cp || !derivative_args(cp, &a, branch_length, params_indices, sumtable)) return PLL_FAILURE;
  if (!plf_derivatives(cp->ctx, &cp->shape, &a, NULL, out)) return cuda_fail(cp);
  *d_f = out[0];
  *dd_f = out[1];
  if ((partition->attributes & PLL_ATTRIB_AB_MASK) &&
      (partition->attributes & PLL_ATTRIB_AB_MASK) != PLL_ATTRIB_AB_STAMATAKIS)
    return asc_derivatives(cp, parent_scaler_index, child_scaler_index, branch_length, params_indices, sumtable, d_f,
                           dd_f);
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_cuda_likelihood_derivatives_async(pll_partition_t * partition, int parent_scaler_index,
                                                     int child_scaler_index, double branch_length,
                                                     const unsigned int * params_indices, const double * sumtable,
                                                     double * dev_out2)
{
  cuda_partition_t * cp = CP(partition);
  plf_deriv_t a;
  (void)parent_scaler_index;
  (void)child_scaler_index;
  if (cp && (partition->attributes & PLL_ATTRIB_AB_MASK) &&
      (partition->attributes & PLL_ATTRIB_AB_MASK) != PLL_ATTRIB_AB_STAMATAKIS)
  {
    /* the Lewis / Felsenstein terms are formed on the host per evaluation (asc_derivatives) */
    set_error(PLL_ERROR_CUDA_UNSUPPORTED, "the asynchronous entry points do not apply the ascertainment bias correction%s", NULL);
    return PLL_FAILURE;
  }
  if (!cp || !dev_out2 || !derivative_args(cp, &a, branch_length, params_indices, sumtable)) return PLL_FAILURE;
  return plf_derivatives(cp->ctx, &cp->shape, &a, dev_out2, NULL) ? PLL_SUCCESS : cuda_fail(cp);
}

/* NEW (additive).  The Newton-Raphson loop a client runs around pll_compute_likelihood_derivatives
 * (examples/newton/newton.c:67-96: evaluate d_f, dd_f at the current length; stop when |d_f| < tolerance;
 * otherwise length -= d_f / dd_f) as ONE device launch and ONE host synchronisation instead of one round
 * trip per iteration.  Steps are clamped to [min_length, max_length]; the loop also stops when a step no
 * longer changes the length.  Returns the length, the derivatives at the last evaluated length and the
 * number of evaluations.  Not available with the Lewis / Felsenstein ascertainment corrections, whose
 * terms are formed on the host per evaluation (use the per-call API there). */
PLL_EXPORT int pll_cuda_newton_branch(pll_partition_t * partition, int parent_scaler_index, int child_scaler_index,
                                      double initial_length, double min_length, double max_length, double tolerance,
                                      unsigned int max_iters, const unsigned int * params_indices,
                                      const double * sumtable, double * length, double * d_f, double * dd_f,
                                      unsigned int * iterations)
{
  cuda_partition_t * cp = CP(partition);
  plf_deriv_t a;
  double out[4] = {0, 0, 0, 0};
  (void)parent_scaler_index;
  (void)child_scaler_index;
  if (!cp) return PLL_FAILURE;
  if ((partition->attributes & PLL_ATTRIB_AB_MASK) &&
      (partition->attributes & PLL_ATTRIB_AB_MASK) != PLL_ATTRIB_AB_STAMATAKIS)
  {
    set_error(PLL_ERROR_CUDA_UNSUPPORTED, "pll_cuda_newton_branch: not with Lewis/Felsenstein ascertainment bias%s", NULL);
    return PLL_FAILURE;
  }
  if (!max_iters || !(min_length <= max_length) || !(tolerance >= 0))
  {
    set_error(PLL_ERROR_PARAM_INVALID, "pll_cuda_newton_branch: invalid bounds, tolerance or iteration limit%s", NULL);
    return PLL_FAILURE;
  }
  if (!derivative_args(cp, &a, initial_length, params_indices, sumtable)) return PLL_FAILURE;
  if ((size_t)a.sites * partition->rate_cats * partition->states_padded * sizeof(double) > NEWTON_FUSED_MAX_BYTES)
  {
    /* a table that does not stay in L2 is streamed from HBM on every evaluation: there the streaming
     * derivative kernels win and the host round trip is a small share, so the same rule is driven from here */
    double t = initial_length, d[2] = {0, 0};
    unsigned int it;
    for (it = 0; it < max_iters;)
    {
      double tn;
      a.branch_length = t;
      if (!plf_derivatives(cp->ctx, &cp->shape, &a, NULL, d)) return cuda_fail(cp);
      ++it;
      if (fabs(d[0]) < tolerance) break;
      tn = t - d[0] / d[1];
      if (tn < min_length) tn = min_length;
      if (tn > max_length) tn = max_length;
      if (!(tn == tn) || tn == t) break;
      t = tn;
    }
    out[0] = t;
    out[1] = d[0];
    out[2] = d[1];
    out[3] = (double)it;
  }
  else if (!plf_newton_branch(cp->ctx, &cp->shape, &a, initial_length, min_length, max_length, tolerance, max_iters,
                              out))
    return cuda_fail(cp);
  if (length) *length = out[0];
  if (d_f) *d_f = out[1];
  if (dd_f) *dd_f = out[2];
  if (iterations) *iterations = (unsigned int)out[3];
  return PLL_SUCCESS;
}

/* ---- site pattern compression (the step before the path) ------------------------------------ */

/* src/compress.c:171-410.  Same contract: `sequence` is overwritten with the unique columns in
 * ascending order of their encoded characters (decoded back with the lowest-ASCII representative,
 * '-' for gaps), *length becomes their number, the returned malloc'ed array holds their weights and
 * site_pattern_map[site] (optional) the pattern every original site went to.  Runs on the device
 * selected by pll_cuda_set_device() / $PLL_CUDA_DEVICE / $LOCAL_RANK. */
static unsigned int * compress_site_patterns(char ** sequence, const pll_state_t * map, int count, int * length,
                                             unsigned int * site_pattern_map)
{
  unsigned char charmap[PLL_ASCII_SIZE], inv_charmap[PLL_ASCII_SIZE];
  pll_state_t max = 0;
  plf_ctx_t * ctx = NULL;
  unsigned int * weight = NULL, * fitted;
  unsigned int compressed = 0, bad_seq = 0, bad_pos = 0;
  char err[256] = "";
  int i, rc;
  if (!count)
  {
    set_error(PLL_ERROR_MSA_EMPTY, "Number of sequences must be greater than 0.%s", NULL);
    return NULL;
  }
  if (!map)
  {
    set_error(PLL_ERROR_MSA_MAP_INVALID, "Map is undefined.%s", NULL);
    return NULL;
  }
  if (map[0])
  {
    set_error(PLL_ERROR_MSA_MAP_INVALID, "'0' cannot be used as a state.%s", NULL);
    return NULL;
  }
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
    if (map[i] > max) max = map[i];
  if (max >= PLL_ASCII_SIZE)
  {
    /* states outside the byte range: number the distinct values in order of first use
     * (src/compress.c:99-122) */
    pll_state_t seen[PLL_ASCII_SIZE];
    unsigned char k = 1;
    int j;
    memcpy(seen, map, sizeof(seen));
    memset(charmap, 0, sizeof(charmap));
    for (i = 0; i < PLL_ASCII_SIZE; ++i)
      if (seen[i])
      {
        charmap[i] = k;
        for (j = i + 1; j < PLL_ASCII_SIZE; ++j)
          if (seen[i] == seen[j])
          {
            charmap[j] = k;
            seen[j] = 0;
          }
        ++k;
      }
  }
  else
    for (i = 0; i < PLL_ASCII_SIZE; ++i) charmap[i] = (unsigned char)map[i];
  memset(inv_charmap, 0, sizeof(inv_charmap));
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
    if (map[i] && (!inv_charmap[charmap[i]] || i == '-')) inv_charmap[charmap[i]] = (unsigned char)i;

  if (*length <= 0)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "alignment length must be positive%s", NULL);
    return NULL;
  }
  weight = (unsigned int *)malloc((size_t)(*length) * sizeof(unsigned int));
  if (!weight)
  {
    set_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate space for storing site weights.%s", NULL);
    return NULL;
  }
  if (!plf_ctx_create(pick_device(), 0, &ctx, err, sizeof(err)))
  {
    set_error(PLL_ERROR_CUDA, "CUDA: %s", err);
    free(weight);
    return NULL;
  }
  rc = plf_compress_patterns(ctx, sequence, (unsigned int)count, (unsigned int)*length, charmap, inv_charmap, weight,
                             site_pattern_map, &compressed, &bad_seq, &bad_pos);
  if (rc == 1)
  {
    for (i = 0; i < count; ++i) sequence[i][compressed] = 0;
    *length = (int)compressed;
    fitted = (unsigned int *)malloc((size_t)compressed * sizeof(unsigned int));
    if (fitted)
    {
      memcpy(fitted, weight, (size_t)compressed * sizeof(unsigned int));
      free(weight);
      weight = fitted;
    }
  }
  else
  {
    if (rc == -1)
    {
      pll_errno = PLL_ERROR_TIPDATA_ILLEGALSTATE;
      snprintf(pll_errmsg, 200, "Cannot encode character %c at sequence %d position %d.", sequence[bad_seq][bad_pos],
               (int)bad_seq + 1, (int)bad_pos + 1);
    }
    else
      set_error(PLL_ERROR_CUDA, "CUDA: %s", plf_last_error(ctx));
    free(weight);
    weight = NULL;
  }
  plf_ctx_destroy(ctx);
  return weight;
}

PLL_EXPORT unsigned int * pll_compress_site_patterns(char ** sequence, const pll_state_t * map, int count, int * length)
{
  return compress_site_patterns(sequence, map, count, length, NULL);
}

PLL_EXPORT unsigned int * pll_compress_site_patterns_msa(pll_msa_t * msa, const pll_state_t * map,
                                                         unsigned int * site_pattern_map)
{
  return compress_site_patterns(msa->sequence, map, msa->count, &msa->length, site_pattern_map);
}

/* ---- explicit reads of device-resident buffers ---------------------------------------------- */

PLL_EXPORT int pll_cuda_download_clv(const pll_partition_t * partition, unsigned int clv_index, double * host_out)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  if (clv_index >= partition->nodes || !partition->clv[clv_index])
  {
    set_error(PLL_ERROR_PARAM_INVALID, "no such CLV%s", NULL);
    return PLL_FAILURE;
  }
  if (!ensure_real(cp, clv_index)) return PLL_FAILURE;
  return plf_download(cp->ctx, host_out, partition->clv[clv_index],
                      (size_t)pll_get_clv_size(partition, clv_index) * sizeof(double))
             ? PLL_SUCCESS
             : cuda_fail(cp);
}

/* NEW (additive).  1 when tip-tip parents of this partition stay virtual. */
PLL_EXPORT int pll_cuda_virtual_cherries(const pll_partition_t * partition)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp || !cp->cherry) return 0;
  /* before the first operation list: what the decision will be for the tip alphabet known so far */
  if (cp->cherry_maxstates != partition->maxstates)
    return plf_virtual_cherries_supported(cp->ctx, &cp->shape, partition->maxstates);
  return cp->cherry_ok;
}

/* NEW (additive).  Number of nodes whose CLV is virtual right now; `clv_index` < nodes asks about one node. */
PLL_EXPORT unsigned int pll_cuda_virtual_clvs(const pll_partition_t * partition, unsigned int clv_index)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return 0;
  if (clv_index < partition->nodes) return (unsigned int)is_virtual(cp, clv_index);
  return cp->cherries_pending;
}

/* NEW (additive).  Make sure partition->clv[clv_index] holds the node's values in HBM (clients that hand the
 * device pointer to their own kernels); every entry point of this library does it on its own. */
PLL_EXPORT int pll_cuda_materialize_clv(pll_partition_t * partition, unsigned int clv_index)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  if (clv_index >= partition->nodes)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "no such CLV%s", NULL);
    return PLL_FAILURE;
  }
  return ensure_real(cp, clv_index) ? PLL_SUCCESS : PLL_FAILURE;
}

PLL_EXPORT unsigned int pll_cuda_scaler_size(const pll_partition_t * partition, unsigned int scaler_index)
{
  cuda_partition_t * cp = CP(partition);
  unsigned int n;
  if (!cp || scaler_index >= partition->scale_buffers) return 0;
  n = sites_alloc(partition);
  if (pll_repeats_enabled(partition) && partition->repeats->perscale_ids[scaler_index])
    n = partition->repeats->perscale_ids[scaler_index];
  if (partition->attributes & PLL_ATTRIB_RATE_SCALERS) n *= partition->rate_cats;
  return n <= cp->scaler_entries[scaler_index] ? n : cp->scaler_entries[scaler_index];
}

PLL_EXPORT int pll_cuda_download_scaler(const pll_partition_t * partition, unsigned int scaler_index,
                                        unsigned int * host_out)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  if (scaler_index >= partition->scale_buffers || !partition->scale_buffer[scaler_index])
  {
    set_error(PLL_ERROR_PARAM_INVALID, "no such scale buffer%s", NULL);
    return PLL_FAILURE;
  }
  return plf_download(cp->ctx, host_out, partition->scale_buffer[scaler_index],
                      (size_t)pll_cuda_scaler_size(partition, scaler_index) * sizeof(unsigned int))
             ? PLL_SUCCESS
             : cuda_fail(cp);
}

static size_t pmatrix_doubles(const pll_partition_t * p)
{
  return (size_t)p->states * p->states_padded * p->rate_cats;
}

PLL_EXPORT int pll_cuda_download_pmatrix(const pll_partition_t * partition, unsigned int matrix_index,
                                         double * host_out)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  if (matrix_index >= partition->prob_matrices)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "no such P-matrix%s", NULL);
    return PLL_FAILURE;
  }
  return plf_download(cp->ctx, host_out, partition->pmatrix[matrix_index], pmatrix_doubles(partition) * sizeof(double))
             ? PLL_SUCCESS
             : cuda_fail(cp);
}

PLL_EXPORT int pll_cuda_upload_pmatrix(pll_partition_t * partition, unsigned int matrix_index, const double * host_in)
{
  cuda_partition_t * cp = CP(partition);
  if (!cp) return PLL_FAILURE;
  if (matrix_index >= partition->prob_matrices)
  {
    set_error(PLL_ERROR_PARAM_INVALID, "no such P-matrix%s", NULL);
    return PLL_FAILURE;
  }
  return plf_upload(cp->ctx, partition->pmatrix[matrix_index], host_in, pmatrix_doubles(partition) * sizeof(double))
             ? PLL_SUCCESS
             : cuda_fail(cp);
}

PLL_EXPORT int pll_cuda_download_sumtable(const pll_partition_t * partition, const double * sumtable_handle,
                                          double * host_out)
{
  cuda_partition_t * cp = CP(partition);
  sumtable_slot_t * slot;
  if (!cp) return PLL_FAILURE;
  if (!(slot = sumtable_slot(cp, sumtable_handle, 0)))
  {
    set_error(PLL_ERROR_PARAM_INVALID, "unknown sumtable handle%s", NULL);
    return PLL_FAILURE;
  }
  return plf_download(cp->ctx, host_out, slot->dev,
                      (size_t)partition->sites * partition->rate_cats * partition->states_padded * sizeof(double))
             ? PLL_SUCCESS
             : cuda_fail(cp);
}

/* ---- debug printers (format of src/output.c:26-101) ---------------------------------------------- */

PLL_EXPORT void pll_show_pmatrix(const pll_partition_t * partition, unsigned int index, unsigned int float_precision)
{
  const unsigned int st = partition->states, sp = partition->states_padded;
  double * m = (double *)malloc(pmatrix_doubles(partition) * sizeof(double));
  unsigned int i, j, k;
  if (!m || !pll_cuda_download_pmatrix(partition, index, m))
  {
    free(m);
    return;
  }
  for (k = 0; k < partition->rate_cats; ++k)
  {
    for (i = 0; i < st; ++i)
    {
      for (j = 0; j < st; ++j) printf("%+2.*f   ", float_precision, m[(size_t)k * st * sp + i * sp + j]);
      printf("\n");
    }
    printf("\n");
  }
  free(m);
}

PLL_EXPORT void pll_show_clv(const pll_partition_t * partition, unsigned int clv_index, int scaler_index,
                             unsigned int float_precision)
{
  const unsigned int st = partition->states, sp = partition->states_padded, R = partition->rate_cats;
  unsigned int s, i, j, k, t;
  double * clv;
  unsigned int * scaler = NULL;
  const unsigned int * site_id;
  if ((clv_index < partition->tips) && (partition->attributes & PLL_ATTRIB_PATTERN_TIP)) return;
  clv = (double *)malloc((size_t)pll_get_clv_size(partition, clv_index) * sizeof(double));
  if (!clv || !pll_cuda_download_clv(partition, clv_index, clv))
  {
    free(clv);
    return;
  }
  if (scaler_index != PLL_SCALE_BUFFER_NONE)
  {
    scaler = (unsigned int *)malloc((size_t)(pll_cuda_scaler_size(partition, scaler_index) + 1) * sizeof(unsigned int));
    if (!scaler || !pll_cuda_download_scaler(partition, scaler_index, scaler))
    {
      free(scaler);
      free(clv);
      return;
    }
  }
  site_id = pll_get_site_id(partition, clv_index);
  printf("[ ");
  for (s = 0; s < partition->sites; ++s)
  {
    i = site_id ? site_id[s] : s;
    printf("{");
    for (j = 0; j < R; ++j)
    {
      printf("(");
      for (k = 0; k < st; ++k)
      {
        double prob = clv[(size_t)i * R * sp + j * sp + k];
        if (scaler)
          for (t = 0; t < scaler[i]; ++t) prob *= PLL_SCALE_THRESHOLD;
        printf("%.*f%s", float_precision, prob, k + 1 < st ? "," : ")");
      }
      if (j < R - 1) printf(",");
    }
    printf("} ");
  }
  printf("]\n");
  free(scaler);
  free(clv);
}
