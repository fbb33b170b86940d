/*
 * pll_parsimony.c -- host side of Fitch parsimony on the GPU.
 *
 * Reference: src/fast_parsimony.c (pll_fastparsimony_init :532, _update_vectors :721, _edge_score :731,
 * _root_score :776), src/parsimony.c:350 (pll_parsimony_destroy), src/utree.c:762
 * (pll_utree_create_pars_buildops), src/stepwise.c:883 (pll_fastparsimony_stepwise).
 *
 * The bit vectors live in HBM (one block, node-major); node costs, the constant cost and the informative
 * flags live on the host, as the public structure promises.  Kernels are in plf_parsimony.cu; there is no
 * CPU fallback.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll_b200.h"
#include "plf_backend.h"
#include "pll_host_internal.h"

#define PARS_MAGIC 0xB200FA55u

typedef struct cuda_parsimony
{
  pll_parsimony_t pub; /* MUST be first: callers hold &pub */
  unsigned int magic;
  plf_ctx_t * ctx;     /* own context: the object outlives the partition it was made from */
  plf_pars_t * ps;
  unsigned int * d_vec; /* [nodes][states][words] */
  unsigned int nodes_count;
  unsigned int * scratch; /* host: scores of the current call */
  unsigned int scratch_cap;
  int no_levels;          /* PLF_PARS_LEVELS=0: always the one-launch chain kernel */
  int force_levels;       /* PLF_PARS_LEVELS=2: always one launch per level (tests) */
  /* weighted (Sankoff) parsimony objects made by pll_parsimony_create: sbuffer[] / anc_states[] are managed
   * allocations (the reference's clients read them on the host), mirrored as device pointer tables */
  int weighted;
  double * d_matrix;
  double ** d_sbuf_table;
  unsigned int ** d_anc_table;
  double * site_min; /* host [sites] */
} cuda_parsimony_t;

static void pars_error(int code, const char * msg)
{
  pll_errno = code;
  snprintf(pll_errmsg, sizeof(pll_errmsg), "%s", msg);
}

static cuda_parsimony_t * PP(const pll_parsimony_t * p)
{
  cuda_parsimony_t * cp = (cuda_parsimony_t *)p;
  if (!cp || cp->magic != PARS_MAGIC)
  {
    pars_error(PLL_ERROR_CUDA_UNSUPPORTED, "parsimony structure was not created by libpll_b200");
    return NULL;
  }
  return cp;
}

static int pars_cuda_fail(cuda_parsimony_t * cp)
{
  pll_errno = PLL_ERROR_CUDA;
  snprintf(pll_errmsg, sizeof(pll_errmsg), "CUDA: %s", plf_last_error(cp->ctx));
  return PLL_FAILURE;
}

static unsigned int * pars_scratch(cuda_parsimony_t * cp, unsigned int n)
{
  if (cp->scratch_cap < n)
  {
    unsigned int * s = (unsigned int *)realloc(cp->scratch, (size_t)(n + n / 2 + 16) * sizeof(unsigned int));
    if (!s)
    {
      pars_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate parsimony score scratch.");
      return NULL;
    }
    cp->scratch = s;
    cp->scratch_cap = n + n / 2 + 16;
  }
  return cp->scratch;
}

PLL_EXPORT void pll_parsimony_destroy(pll_parsimony_t * pars)
{
  cuda_parsimony_t * cp = (cuda_parsimony_t *)pars;
  if (!pars) return;
  if (cp->magic != PARS_MAGIC) return; /* not ours: nothing we can safely free */
  if (cp->ctx)
  {
    unsigned int i;
    plf_free(cp->ctx, cp->d_vec);
    if (cp->weighted)
    {
      for (i = 0; cp->pub.sbuffer && i < cp->pub.score_buffers + cp->pub.tips; ++i) plf_free(cp->ctx, cp->pub.sbuffer[i]);
      for (i = cp->pub.tips; cp->pub.anc_states && i < cp->pub.ancestral_buffers + cp->pub.tips; ++i)
        plf_free(cp->ctx, cp->pub.anc_states[i]);
      plf_free(cp->ctx, cp->d_matrix);
      plf_free(cp->ctx, cp->d_sbuf_table);
      plf_free(cp->ctx, cp->d_anc_table);
    }
    plf_pars_destroy(cp->ps);
    plf_ctx_destroy(cp->ctx);
  }
  free(cp->pub.sbuffer);
  free(cp->pub.anc_states);
  free(cp->pub.score_matrix);
  free(cp->site_min);
  free(cp->pub.packedvector);
  free(cp->pub.node_cost);
  free(cp->pub.informative);
  free(cp->scratch);
  cp->magic = 0;
  free(cp);
}

PLL_EXPORT pll_parsimony_t * pll_fastparsimony_init(const pll_partition_t * partition)
{
  pll_cuda_tipsource_t src;
  cuda_parsimony_t * cp = NULL;
  int * d_inf = NULL;
  unsigned int * d_bitpos = NULL;
  unsigned int bitcount = 0, words, i;
  size_t stride;
  char err[200] = {0};

  /* src/fast_parsimony.c:538-547 */
  if (partition && partition->states > 20 && !(partition->attributes & PLL_ATTRIB_PATTERN_TIP))
  {
    pars_error(PLL_ERROR_STEPWISE_UNSUPPORTED, "Use PLL_ATTRIB_PATTERN_TIP for more than 20 states.");
    return NULL;
  }
  if (!pll_cuda_internal_tipsource(partition, &src)) return NULL;

  cp = (cuda_parsimony_t *)calloc(1, sizeof(cuda_parsimony_t));
  if (!cp)
  {
    plf_free(src.ctx, src.d_ptrs);
    pars_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate parsimony structure.");
    return NULL;
  }
  cp->magic = PARS_MAGIC;
  {
    const char * v = getenv("PLF_PARS_LEVELS");
    cp->no_levels = v && v[0] == '0';
    cp->force_levels = v && v[0] == '2';
  }
  cp->pub.tips = partition->tips;
  cp->pub.inner_nodes = partition->tips - 1;
  cp->pub.sites = partition->sites;
  cp->pub.attributes = partition->attributes;
  cp->pub.states = partition->states;
  cp->pub.alignment = partition->alignment;
  /* one vector per tip and three (one per direction) per inner node, src/fast_parsimony.c:32 */
  cp->nodes_count = cp->pub.tips + 3 * cp->pub.inner_nodes;

  if (!plf_ctx_create(plf_ctx_device(src.ctx), 0, &cp->ctx, err, sizeof(err)))
  {
    pll_errno = PLL_ERROR_CUDA;
    snprintf(pll_errmsg, sizeof(pll_errmsg), "CUDA: %.180s", err);
    cp->ctx = NULL;
    goto fail_src;
  }
  if (!plf_pars_create(cp->ctx, &cp->ps))
  {
    pars_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate parsimony backend state.");
    goto fail_src;
  }
  d_inf = (int *)plf_alloc(cp->ctx, ((size_t)partition->sites + 1) * sizeof(int), 0);
  d_bitpos = (unsigned int *)plf_alloc(cp->ctx, ((size_t)partition->sites + 1) * sizeof(unsigned int), 0);
  cp->pub.informative = (int *)malloc(((size_t)partition->sites + 1) * sizeof(int));
  cp->pub.node_cost = (unsigned int *)calloc(cp->nodes_count, sizeof(unsigned int));
  cp->pub.packedvector = (unsigned int **)calloc(cp->nodes_count, sizeof(unsigned int *));
  if (!cp->pub.informative || !cp->pub.node_cost || !cp->pub.packedvector)
  {
    pars_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate parsimony cost array.");
    goto fail_src;
  }
  if (!d_inf || !d_bitpos) goto fail_cuda;

  /* which sites are informative, what the others cost (src/fast_parsimony.c:381-413) */
  if (!plf_pars_informative(cp->ps, &src.tips, d_inf, d_bitpos, &bitcount, &cp->pub.const_cost,
                            &cp->pub.informative_count))
    goto fail_cuda;
  /* number of 32-bit words per state row (src/fast_parsimony.c:251-258; no SIMD rounding here) */
  words = bitcount / 32 + (bitcount % 32 != 0);
  stride = (size_t)cp->pub.states * words;
  cp->d_vec = (unsigned int *)plf_alloc(cp->ctx, (stride * cp->nodes_count + 1) * sizeof(unsigned int), 1);
  if (!cp->d_vec) goto fail_cuda;
  for (i = 0; i < cp->nodes_count; ++i) cp->pub.packedvector[i] = cp->d_vec + stride * i;
  cp->pub.packedvector_count = words;
  if (!plf_pars_pack(cp->ps, &src.tips, d_bitpos, bitcount, words, cp->d_vec) ||
      !plf_download(cp->ctx, cp->pub.informative, d_inf, (size_t)partition->sites * sizeof(int)))
    goto fail_cuda;
  plf_free(cp->ctx, d_inf);
  plf_free(cp->ctx, d_bitpos);
  plf_free(src.ctx, src.d_ptrs);
  return &cp->pub;

fail_cuda:
  pars_cuda_fail(cp);
fail_src:
  if (cp->ctx)
  {
    plf_free(cp->ctx, d_inf);
    plf_free(cp->ctx, d_bitpos);
  }
  plf_free(src.ctx, src.d_ptrs);
  pll_parsimony_destroy(&cp->pub);
  return NULL;
}

static int pars_indices_ok(const cuda_parsimony_t * cp, const unsigned int * idx, size_t n)
{
  size_t i;
  for (i = 0; i < n; ++i)
    if (idx[i] >= cp->nodes_count)
    {
      pll_errno = PLL_ERROR_PARAM_INVALID;
      snprintf(pll_errmsg, sizeof(pll_errmsg), "Parsimony score index %u out of range (%u vectors).", idx[i],
               cp->nodes_count);
      return 0;
    }
  return 1;
}

/* Level of every operation of a list that must behave as if run in order: an operation goes one level above
 * the last writer of its children (read after write), above the last reader of its parent (write after read) and
 * above the last writer of its parent.  lw / lr: per vector, level of the last write / read so far (0 = none).
 * Returns the number of levels. */
static unsigned int pars_levels(const pll_pars_buildop_t * ops, unsigned int count, unsigned int * level,
                                unsigned int * lw, unsigned int * lr, unsigned int nvec)
{
  unsigned int i, nlevels = 0;
  memset(lw, 0, (size_t)nvec * sizeof(unsigned int));
  memset(lr, 0, (size_t)nvec * sizeof(unsigned int));
  for (i = 0; i < count; ++i)
  {
    const unsigned int p = ops[i].parent_score_index, a = ops[i].child1_score_index, b = ops[i].child2_score_index;
    unsigned int l = lw[a] > lw[b] ? lw[a] : lw[b];
    if (lr[p] > l) l = lr[p];
    if (lw[p] > l) l = lw[p];
    ++l;
    level[i] = l;
    lw[p] = l;
    if (lr[a] < l) lr[a] = l;
    if (lr[b] < l) lr[b] = l;
    if (l > nlevels) nlevels = l;
  }
  return nlevels;
}

/* NEW (additive): the level schedule pll_fastparsimony_update_vectors uses, for inspection and tests.
 * level_of_op[i] is 1-based; vectors = number of score indices in use (every index must be below it). */
PLL_EXPORT int pll_cuda_schedule_parsimony_levels(const pll_pars_buildop_t * ops, unsigned int count,
                                                  unsigned int vectors, unsigned int * level_of_op)
{
  unsigned int * lw, i;
  int nlevels;
  if (!count) return 0;
  for (i = 0; i < count; ++i)
    if (ops[i].parent_score_index >= vectors || ops[i].child1_score_index >= vectors ||
        ops[i].child2_score_index >= vectors)
    {
      pars_error(PLL_ERROR_PARAM_INVALID, "Parsimony score index out of range.");
      return -1;
    }
  lw = (unsigned int *)malloc((size_t)2 * vectors * sizeof(unsigned int));
  if (!lw)
  {
    pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return -1;
  }
  nlevels = (int)pars_levels(ops, count, level_of_op, lw, lw + vectors, vectors);
  free(lw);
  return nlevels;
}

PLL_EXPORT void pll_fastparsimony_update_vectors(pll_parsimony_t * parsimony, const pll_pars_buildop_t * ops,
                                                 unsigned int count)
{
  cuda_parsimony_t * cp = PP(parsimony);
  unsigned int * scores, i, nlevels = 0;
  if (!cp || !count) return;
  /* pll_pars_buildop_t is three consecutive unsigned ints: the list goes to the device as it is */
  if (!pars_indices_ok(cp, (const unsigned int *)ops, (size_t)3 * count)) return;
  /* scratch: scores [count] | level [count] | sorted ops [3 count] | position [count] | start [count + 2] |
   * last write, last read [nvec each] */
  scores = pars_scratch(cp, 7 * count + 2 + 2 * cp->nodes_count);
  if (!scores) return;
  if ((count >= 8 || cp->force_levels) && !cp->no_levels)
  {
    unsigned int * level = scores + count, * sorted = level + count, * pos = sorted + 3 * (size_t)count;
    unsigned int * start = pos + count, * lw = start + count + 2, * lr = lw + cp->nodes_count;
    nlevels = pars_levels(ops, count, level, lw, lr, cp->nodes_count);
    unsigned int launches = 0, chained = 0, in_run = 0;
    memset(start, 0, ((size_t)nlevels + 2) * sizeof(unsigned int));
    for (i = 0; i < count; ++i) start[level[i]]++; /* level l (1-based) counted in start[l] */
    /* what the level form costs: a launch per level of more than two operations and per run of narrower
     * levels (those are chained inside one launch); a launch is worth about four chained operations */
    for (i = 1; i <= nlevels; ++i)
      if (start[i] > 2)
      {
        ++launches;
        in_run = 0;
      }
      else
      {
        chained += start[i];
        if (!in_run) ++launches;
        in_run = 1;
      }
    if ((unsigned long long)launches * 4 + chained <= count || cp->force_levels)
    {
      for (i = 1; i <= nlevels; ++i) start[i] += start[i - 1];
      /* start[l] is now the end of level l = the start of level l + 1: fill every level from its end */
      for (i = count; i-- > 0;)
      {
        const unsigned int at = --start[level[i]];
        pos[i] = at;
        sorted[3 * (size_t)at] = ops[i].parent_score_index;
        sorted[3 * (size_t)at + 1] = ops[i].child1_score_index;
        sorted[3 * (size_t)at + 2] = ops[i].child2_score_index;
      }
      /* after the fill start[l] is the start of level l (1-based); start[nlevels + 1] closes the last one */
      start[nlevels + 1] = count;
      if (!plf_pars_update_levels(cp->ps, cp->d_vec, cp->pub.states, cp->pub.packedvector_count, sorted, count,
                                  start + 1, nlevels, level))
      {
        pars_cuda_fail(cp);
        return;
      }
      for (i = 0; i < count; ++i) scores[i] = level[pos[i]]; /* `level` received the scores in sorted order */
    }
    else
      nlevels = 0;
  }
  if (!nlevels && !plf_pars_update(cp->ps, cp->d_vec, cp->pub.states, cp->pub.packedvector_count,
                                   (const unsigned int *)ops, count, scores))
  {
    pars_cuda_fail(cp);
    return;
  }
  /* node costs chain through the list in its order (src/fast_parsimony.c:527-529) */
  for (i = 0; i < count; ++i)
    cp->pub.node_cost[ops[i].parent_score_index] =
        scores[i] + cp->pub.node_cost[ops[i].child1_score_index] + cp->pub.node_cost[ops[i].child2_score_index];
}

PLL_EXPORT int pll_cuda_fastparsimony_edge_scores(const pll_parsimony_t * parsimony, const unsigned int * pairs,
                                                  unsigned int n, unsigned int * scores)
{
  cuda_parsimony_t * cp = PP(parsimony);
  unsigned int i;
  if (!cp) return PLL_FAILURE;
  if (!n) return PLL_SUCCESS;
  if (!pars_indices_ok(cp, pairs, (size_t)2 * n)) return PLL_FAILURE;
  if (!plf_pars_edge_scores(cp->ps, cp->d_vec, cp->pub.states, cp->pub.packedvector_count, pairs, n, scores))
    return pars_cuda_fail(cp);
  for (i = 0; i < n; ++i)
    scores[i] += cp->pub.node_cost[pairs[2 * i]] + cp->pub.node_cost[pairs[2 * i + 1]] + cp->pub.const_cost;
  return PLL_SUCCESS;
}

PLL_EXPORT unsigned int pll_fastparsimony_edge_score(const pll_parsimony_t * parsimony, unsigned int node1_score_index,
                                                     unsigned int node2_score_index)
{
  unsigned int pair[2], score = 0;
  pair[0] = node1_score_index;
  pair[1] = node2_score_index;
  if (!pll_cuda_fastparsimony_edge_scores(parsimony, pair, 1, &score)) return ~0u;
  return score;
}

PLL_EXPORT unsigned int pll_fastparsimony_root_score(const pll_parsimony_t * parsimony, unsigned int root_index)
{
  return parsimony->node_cost[root_index] + parsimony->const_cost;
}

PLL_EXPORT int pll_cuda_download_parsimony_vector(const pll_parsimony_t * parsimony, unsigned int index,
                                                  unsigned int * dst)
{
  cuda_parsimony_t * cp = PP(parsimony);
  if (!cp) return PLL_FAILURE;
  if (!pars_indices_ok(cp, &index, 1)) return PLL_FAILURE;
  if (!plf_download(cp->ctx, dst, cp->pub.packedvector[index],
                    (size_t)cp->pub.states * cp->pub.packedvector_count * sizeof(unsigned int)))
    return pars_cuda_fail(cp);
  return PLL_SUCCESS;
}

PLL_EXPORT void pll_utree_create_pars_buildops(pll_unode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                               pll_pars_buildop_t * ops, unsigned int * ops_count)
{
  unsigned int i, n = 0;
  for (i = 0; i < trav_buffer_size; ++i)
  {
    const pll_unode_t * node = trav_buffer[i];
    if (!node->next) continue; /* tips have no operation */
    ops[n].parent_score_index = node->node_index;
    ops[n].child1_score_index = node->next->back->node_index;
    ops[n].child2_score_index = node->next->next->back->node_index;
    ++n;
  }
  *ops_count = n;
}

/* ---- weighted (Sankoff) parsimony, src/parsimony.c -------------------------------------------------------
 * pll_parsimony_create has no attribute argument: objects made by this library live on the GPU selected by
 * pll_cuda_set_device() / $PLL_CUDA_DEVICE / $LOCAL_RANK.  Score and ancestral buffers are managed allocations:
 * kernels work on them in HBM and the host may read them after any call, as the reference's clients do
 * (examples/parsimony/npr-pars.c:240-281). */

static cuda_parsimony_t * WP(const pll_parsimony_t * p)
{
  cuda_parsimony_t * cp = PP(p);
  if (cp && !cp->weighted)
  {
    pars_error(PLL_ERROR_PARAM_INVALID, "parsimony structure has no score buffers (made by pll_fastparsimony_init)");
    return NULL;
  }
  return cp;
}

PLL_EXPORT pll_parsimony_t * pll_parsimony_create(unsigned int tips, unsigned int states, unsigned int sites,
                                                  const double * score_matrix, unsigned int score_buffers,
                                                  unsigned int ancestral_buffers)
{
  cuda_parsimony_t * cp = (cuda_parsimony_t *)calloc(1, sizeof(cuda_parsimony_t));
  const size_t nbuf = (size_t)score_buffers + tips, nanc = (size_t)ancestral_buffers + tips;
  char err[200] = {0};
  unsigned int i;
  int alloc_failed = 0;
  if (!cp)
  {
    pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  cp->magic = PARS_MAGIC;
  cp->weighted = 1;
  cp->pub.tips = tips;
  cp->pub.states = states;
  cp->pub.sites = sites;
  cp->pub.score_buffers = score_buffers;
  cp->pub.ancestral_buffers = ancestral_buffers;
  if (!plf_ctx_create(pll_cuda_internal_pick_device(), 1, &cp->ctx, err, sizeof(err)))
  {
    pll_errno = PLL_ERROR_CUDA;
    snprintf(pll_errmsg, sizeof(pll_errmsg), "CUDA: %.180s", err);
    cp->ctx = NULL;
    pll_parsimony_destroy(&cp->pub);
    return NULL;
  }
  cp->pub.score_matrix = (double *)malloc((size_t)states * states * sizeof(double));
  cp->pub.sbuffer = (double **)calloc(nbuf ? nbuf : 1, sizeof(double *));
  cp->pub.anc_states = (unsigned int **)calloc(nanc ? nanc : 1, sizeof(unsigned int *));
  cp->site_min = (double *)malloc(((size_t)sites + 1) * sizeof(double));
  if (!cp->pub.score_matrix || !cp->pub.sbuffer || !cp->pub.anc_states || !cp->site_min)
  {
    pll_parsimony_destroy(&cp->pub);
    pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory for score buffers.");
    return NULL;
  }
  memcpy(cp->pub.score_matrix, score_matrix, (size_t)states * states * sizeof(double));
  cp->d_matrix = (double *)plf_alloc(cp->ctx, (size_t)states * states * sizeof(double), 0);
  cp->d_sbuf_table = (double **)plf_alloc(cp->ctx, (nbuf + 1) * sizeof(double *), 1);
  cp->d_anc_table = (unsigned int **)plf_alloc(cp->ctx, (nanc + 1) * sizeof(unsigned int *), 1);
  for (i = 0; i < nbuf; ++i)
    if (!(cp->pub.sbuffer[i] = (double *)plf_alloc(cp->ctx, ((size_t)sites * states + 1) * sizeof(double), 1)))
      alloc_failed = 1;
  for (i = tips; i < nanc && !alloc_failed; ++i)
    if (!(cp->pub.anc_states[i] = (unsigned int *)plf_alloc(cp->ctx, ((size_t)sites + 1) * sizeof(unsigned int), 1)))
      alloc_failed = 1;
  if (!cp->d_matrix || !cp->d_sbuf_table || !cp->d_anc_table || alloc_failed ||
      !plf_upload(cp->ctx, cp->d_matrix, score_matrix, (size_t)states * states * sizeof(double)) ||
      !plf_upload(cp->ctx, cp->d_sbuf_table, cp->pub.sbuffer, nbuf * sizeof(double *)) ||
      !plf_upload(cp->ctx, cp->d_anc_table, cp->pub.anc_states, nanc * sizeof(unsigned int *)))
  {
    pars_cuda_fail(cp);
    pll_parsimony_destroy(&cp->pub);
    return NULL;
  }
  return &cp->pub;
}

PLL_EXPORT int pll_set_parsimony_sequence(pll_parsimony_t * pars, unsigned int tip_index, const pll_state_t * map,
                                          const char * sequence)
{
  cuda_parsimony_t * cp = WP(pars);
  const unsigned int states = pars ? pars->states : 0;
  double inf;
  unsigned int i;
  if (!cp) return PLL_FAILURE;
  if (tip_index >= pars->tips + pars->score_buffers)
  {
    pars_error(PLL_ERROR_PARAM_INVALID, "Parsimony score buffer index out of range.");
    return PLL_FAILURE;
  }
  /* "infinity" = the largest entry of the score matrix plus one (src/parsimony.c:38-43) */
  inf = pars->score_matrix[0];
  for (i = 1; i < states * states; ++i)
    if (pars->score_matrix[i] > inf) inf = pars->score_matrix[i];
  inf++;
  for (i = 0; i < pars->sites; ++i)
    if (map[(unsigned char)sequence[i]] == 0)
    {
      pll_errno = PLL_ERROR_TIPDATA_ILLEGALSTATE;
      snprintf(pll_errmsg, 200, "Illegal state code in tip \"%c\"", sequence[i]);
      printf("%s\n", pll_errmsg); /* as the reference does (src/parsimony.c:51) */
      return PLL_FAILURE;
    }
  if (!plf_wpars_tip(cp->ctx, pars->sbuffer[tip_index], sequence, map, pars->sites, states, inf))
    return pars_cuda_fail(cp);
  return PLL_SUCCESS;
}

PLL_EXPORT double pll_parsimony_score(pll_parsimony_t * pars, unsigned int score_buffer_index)
{
  cuda_parsimony_t * cp = WP(pars);
  double sum = 0;
  unsigned int i;
  if (!cp) return 0;
  if (score_buffer_index >= pars->tips + pars->score_buffers)
  {
    pars_error(PLL_ERROR_PARAM_INVALID, "Parsimony score buffer index out of range.");
    return 0;
  }
  if (!plf_wpars_site_min(cp->ctx, pars->sbuffer[score_buffer_index], pars->states, pars->sites, cp->site_min))
  {
    pars_cuda_fail(cp);
    return 0;
  }
  /* the per-site minima are added in site order, as in the reference, so that the total has its bits */
  for (i = 0; i < pars->sites; ++i) sum += cp->site_min[i];
  return sum;
}

PLL_EXPORT double pll_parsimony_build(pll_parsimony_t * pars, const pll_pars_buildop_t * operations,
                                      unsigned int count)
{
  cuda_parsimony_t * cp = WP(pars);
  const unsigned int * idx = (const unsigned int *)operations;
  size_t i;
  if (!cp || !count) return 0;
  for (i = 0; i < (size_t)3 * count; ++i)
    if (idx[i] >= pars->tips + pars->score_buffers)
    {
      pars_error(PLL_ERROR_PARAM_INVALID, "Parsimony score buffer index out of range.");
      return 0;
    }
  if (!plf_wpars_build(cp->ctx, cp->d_sbuf_table, idx, count, pars->states, pars->sites, cp->d_matrix))
  {
    pars_cuda_fail(cp);
    return 0;
  }
  return pll_parsimony_score(pars, operations[count - 1].parent_score_index);
}

PLL_EXPORT void pll_parsimony_reconstruct(pll_parsimony_t * pars, const pll_state_t * map,
                                          const pll_pars_recop_t * operations, unsigned int count)
{
  cuda_parsimony_t * cp = WP(pars);
  unsigned int revmap[256];
  unsigned int i;
  if (!cp || !count) return;
  for (i = 0; i < count; ++i)
  {
    const pll_pars_recop_t * op = operations + i;
    if (op->node_score_index >= pars->tips + pars->score_buffers ||
        op->node_ancestral_index < pars->tips || op->node_ancestral_index >= pars->tips + pars->ancestral_buffers ||
        (i && (op->parent_score_index >= pars->tips + pars->score_buffers || op->parent_ancestral_index < pars->tips ||
               op->parent_ancestral_index >= pars->tips + pars->ancestral_buffers)))
    {
      pars_error(PLL_ERROR_PARAM_INVALID, "Parsimony reconstruction index out of range.");
      return;
    }
  }
  /* character of every one-state code; the later character wins, as in src/parsimony.c:327-334 */
  memset(revmap, 0, sizeof(revmap));
  for (i = 0; i < 256; ++i)
    if (map[i] && !(map[i] & (map[i] - 1))) revmap[__builtin_ctzll(map[i])] = i;
  /* pll_pars_recop_t is four consecutive unsigned ints */
  if (!plf_wpars_reconstruct(cp->ctx, cp->d_sbuf_table, cp->d_anc_table, (const unsigned int *)operations, count,
                             pars->states, pars->sites, revmap, map))
    pars_cuda_fail(cp);
}

/* src/rtree.c:458-520 */
PLL_EXPORT void pll_rtree_create_pars_buildops(pll_rnode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                               pll_pars_buildop_t * ops, unsigned int * ops_count)
{
  unsigned int i, n = 0;
  for (i = 0; i < trav_buffer_size; ++i)
  {
    const pll_rnode_t * node = trav_buffer[i];
    if (!node->left) continue;
    ops[n].parent_score_index = node->clv_index;
    ops[n].child1_score_index = node->left->clv_index;
    ops[n].child2_score_index = node->right->clv_index;
    ++n;
  }
  *ops_count = n;
}

PLL_EXPORT void pll_rtree_create_pars_recops(pll_rnode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                             pll_pars_recop_t * ops, unsigned int * ops_count)
{
  unsigned int i, n = 0;
  for (i = 0; i < trav_buffer_size; ++i)
  {
    const pll_rnode_t * node = trav_buffer[i];
    if (!node->left) continue;
    ops[n].node_score_index = ops[n].node_ancestral_index = node->clv_index;
    /* the root has no parent: its entries are never read */
    ops[n].parent_score_index = ops[n].parent_ancestral_index = node->parent ? node->parent->clv_index : 0;
    ++n;
  }
  *ops_count = n;
}

/* ---- randomised stepwise addition and SPR rounds (src/stepwise.c) -----------------------------------------
 *
 * Tips (or, in an SPR round, pruned subtrees) are tried on every edge of the current tree and stay on the first
 * edge of minimal parsimony length.  The reference evaluates an edge with one vector update plus one edge score
 * (two passes over the vectors and two function calls per edge and partition).  Here, per insertion:
 *   1. every directional vector that the previous insertion invalidated is recomputed by ONE launch
 *      (plf_pars_update runs a whole dependency-ordered list),
 *   2. ALL candidate edges are scored by ONE launch (plf_pars_insert_scan: the merge of the edge's two
 *      vectors never leaves registers).
 * Costs are exact integers, the edge lists are the reference's and ties go to the first edge, so trees and
 * costs are those of the reference for the same seed.
 */

typedef struct stepwise
{
  pll_parsimony_t ** list;
  cuda_parsimony_t ** pars;
  unsigned int pars_count;
  unsigned int nvec;
  unsigned char * valid;    /* per directional vector (node_index) */
  pll_pars_buildop_t * ops;
  unsigned int ops_count;
  unsigned int * pairs;     /* 2 per candidate edge */
  unsigned int * which;     /* candidate -> position in the edge list */
  unsigned int * scan;      /* per candidate */
  unsigned int * total;
} stepwise_t;

static void sw_release(stepwise_t * sw)
{
  free(sw->pars);
  free(sw->valid);
  free(sw->ops);
  free(sw->pairs);
  free(sw->which);
  free(sw->scan);
  free(sw->total);
  memset(sw, 0, sizeof(*sw));
}

/* 1 ok, 0 failure with pll_errno set */
static int sw_prepare(stepwise_t * sw, pll_parsimony_t ** list, unsigned int count)
{
  unsigned int i, tips;
  memset(sw, 0, sizeof(*sw));
  if (!list || !count || !list[0])
  {
    pars_error(PLL_ERROR_PARAM_INVALID, "Stepwise parsimony needs at least one parsimony structure.");
    return 0;
  }
  tips = list[0]->tips;
  sw->list = list;
  sw->pars_count = count;
  sw->nvec = tips + 3 * list[0]->inner_nodes;
  sw->pars = (cuda_parsimony_t **)calloc(count, sizeof(cuda_parsimony_t *));
  sw->valid = (unsigned char *)calloc(sw->nvec, 1);
  sw->ops = (pll_pars_buildop_t *)malloc((size_t)3 * tips * sizeof(pll_pars_buildop_t));
  sw->pairs = (unsigned int *)malloc((size_t)4 * tips * sizeof(unsigned int));
  sw->which = (unsigned int *)malloc((size_t)2 * tips * sizeof(unsigned int));
  sw->scan = (unsigned int *)malloc((size_t)2 * tips * sizeof(unsigned int));
  sw->total = (unsigned int *)malloc((size_t)2 * tips * sizeof(unsigned int));
  if (!sw->pars || !sw->valid || !sw->ops || !sw->pairs || !sw->which || !sw->scan || !sw->total)
  {
    sw_release(sw);
    pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return 0;
  }
  for (i = 0; i < count; ++i)
    if (!(sw->pars[i] = PP(list[i])) || list[i]->tips != tips || list[i]->inner_nodes != list[0]->inner_nodes)
    {
      if (sw->pars[i]) pars_error(PLL_ERROR_STEPWISE_STRUCT, "Parsimony structures tips/inner nodes not equal.");
      sw_release(sw);
      return 0;
    }
  return 1;
}

static pll_unode_t * sw_inner_create(unsigned int i, unsigned int tips)
{
  /* a ring of three records: clv_index tips+i, node_index tips+3i+{0,1,2} (src/stepwise.c:236-285) */
  pll_unode_t * ring[3];
  int k;
  for (k = 0; k < 3; ++k)
  {
    ring[k] = (pll_unode_t *)calloc(1, sizeof(pll_unode_t));
    if (!ring[k])
    {
      while (k--) free(ring[k]);
      return NULL;
    }
    ring[k]->clv_index = tips + i;
    ring[k]->node_index = tips + 3 * i + (unsigned int)k;
  }
  ring[0]->next = ring[1];
  ring[1]->next = ring[2];
  ring[2]->next = ring[0];
  return ring[0];
}

static void sw_link(pll_unode_t * a, pll_unode_t * b)
{
  a->back = b;
  b->back = a;
  b->pmatrix_index = a->pmatrix_index;
}

/* edge (a, a->back) becomes a--b ... c--(old a->back): the ring of b and c now sits on the edge */
static void sw_split(pll_unode_t * a, pll_unode_t * b, pll_unode_t * c)
{
  sw_link(c, a->back);
  sw_link(a, b);
}

/* takes the ring of p off the tree; returns one end of the edge that closes the gap (src/stepwise.c:338) */
static pll_unode_t * sw_prune(pll_unode_t * p)
{
  pll_unode_t * a = p->next->back, * b = p->next->next->back;
  sw_link(a, b);
  p->next->back = p->next->next->back = NULL;
  return a;
}

/* post-order list of the invalid vectors needed for the vector at `n` */
static void sw_collect(stepwise_t * sw, pll_unode_t * n)
{
  if (!n->next || sw->valid[n->node_index]) return;
  sw_collect(sw, n->next->back);
  sw_collect(sw, n->next->next->back);
  sw->valid[n->node_index] = 1;
  sw->ops[sw->ops_count].parent_score_index = n->node_index;
  sw->ops[sw->ops_count].child1_score_index = n->next->back->node_index;
  sw->ops[sw->ops_count].child2_score_index = n->next->next->back->node_index;
  sw->ops_count++;
}

/* vectors that look away from the new tip keep their content: they are the ones reached from `n` downwards */
static void sw_validate_below(stepwise_t * sw, pll_unode_t * n)
{
  if (!n->next) return;
  sw->valid[n->node_index] = 1;
  sw_validate_below(sw, n->next->back);
  sw_validate_below(sw, n->next->next->back);
}

/* Places the ring of v (v->back = the tip or pruned subtree to insert; v->next, v->next->next free) on the best
 * of the edges; src/stepwise.c:436-583.  constraint (by clv_index) restricts the edges, prune_edge is where the
 * subtree came from (SPR) or NULL (new tip: the two new edges are appended to the list).  Returns the cost of
 * the tree after the placement; *failed is set on a CUDA error. */
static unsigned int sw_insert_best(stepwise_t * sw, pll_unode_t ** edges, unsigned int edge_count, pll_unode_t * v,
                                   const unsigned int * constraint, pll_unode_t * prune_edge, int * failed)
{
  const unsigned int third = v->back->node_index;
  unsigned int e, k, n = 0, best = ~0u, min_cost = ~0u;

  /* 1. bring every directional vector of the tree, and of the subtree to insert, up to date: one launch */
  sw->ops_count = 0;
  for (e = 0; e < edge_count; ++e)
  {
    sw_collect(sw, edges[e]);
    sw_collect(sw, edges[e]->back);
  }
  sw_collect(sw, v->back);
  pll_errno = 0;
  for (k = 0; k < sw->pars_count; ++k)
  {
    pll_fastparsimony_update_vectors(sw->list[k], sw->ops, sw->ops_count);
    if (pll_errno)
    {
      *failed = 1;
      return ~0u;
    }
  }

  /* 2. score the insertion on every admissible edge: one launch per partition */
  for (e = 0; e < edge_count; ++e)
  {
    if (constraint)
    {
      const unsigned int s = constraint[v->clv_index];
      if (s && s != constraint[edges[e]->clv_index] && s != constraint[edges[e]->back->clv_index]) continue;
    }
    sw->pairs[2 * n] = edges[e]->node_index;
    sw->pairs[2 * n + 1] = edges[e]->back->node_index;
    sw->which[n] = e;
    sw->total[n] = 0;
    ++n;
  }
  if (!n)
  {
    /* no admissible edge: back to where it came from (src/stepwise.c:533-544) */
    sw->pairs[0] = prune_edge->node_index;
    sw->pairs[1] = prune_edge->back->node_index;
    sw->total[0] = 0;
  }
  for (k = 0; k < sw->pars_count; ++k)
  {
    cuda_parsimony_t * cp = sw->pars[k];
    const unsigned int m = n ? n : 1;
    if (!plf_pars_insert_scan(cp->ps, cp->d_vec, cp->pub.states, cp->pub.packedvector_count, sw->pairs, m, third,
                              sw->scan))
    {
      pars_cuda_fail(cp);
      *failed = 1;
      return ~0u;
    }
    for (e = 0; e < m; ++e)
      sw->total[e] += sw->scan[e] + cp->pub.node_cost[sw->pairs[2 * e]] + cp->pub.node_cost[sw->pairs[2 * e + 1]] +
                      cp->pub.node_cost[third] + cp->pub.const_cost;
  }
  for (e = 0; e < n; ++e)
    if (sw->total[e] < min_cost)
    {
      min_cost = sw->total[e];
      best = sw->which[e];
    }

  /* 3. place it */
  if (n)
    sw_split(edges[best], v->next, v->next->next);
  else
  {
    sw_split(prune_edge, v->next, v->next->next);
    min_cost = sw->total[0];
  }
  if (!prune_edge)
  {
    edges[edge_count] = v;
    edges[edge_count + 1] = v->next->next;
  }

  /* 4. after a new tip only the vectors that look away from it are still right; after an SPR none is trusted */
  memset(sw->valid, 0, sw->nvec);
  if (!prune_edge)
  {
    sw_validate_below(sw, v);
    sw->valid[v->node_index] = 0; /* never stored: the scan kept the merge in registers */
  }
  return min_cost;
}

/* Fisher-Yates shuffle driven by the glibc-compatible generator (src/stepwise.c:56-106); seed 0 = identity */
static unsigned int * sw_shuffled(unsigned int n, unsigned int seed)
{
  unsigned int * x = (unsigned int *)malloc(((size_t)n + 1) * sizeof(unsigned int));
  unsigned int i;
  if (!x)
  {
    pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  for (i = 0; i < n; ++i) x[i] = i;
  if (seed && n > 1)
  {
    pll_random_state * rs = pll_random_create(seed);
    if (!rs)
    {
      free(x);
      pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
      return NULL;
    }
    for (i = n; i-- > 0;)
    {
      int r;
      unsigned int j, t;
      pll_random_r(&rs->rdata, &r);
      j = (unsigned int)(((double)r / RAND_MAX) * (i + 1));
      if (j > i) j = i; /* r == RAND_MAX: the reference would index one past the range */
      t = x[i];
      x[i] = x[j];
      x[j] = t;
    }
    pll_random_destroy(rs);
  }
  return x;
}

static void sw_free_ring(pll_unode_t * n)
{
  if (!n) return;
  if (n->next)
  {
    free(n->next->next);
    free(n->next);
  }
  free(n->label);
  free(n);
}

PLL_EXPORT pll_utree_t * pll_fastparsimony_stepwise(pll_parsimony_t ** list, char * const * labels,
                                                    unsigned int * cost, unsigned int count, unsigned int seed)
{
  unsigned int tips, inner_nodes, i, edge_count;
  pll_unode_t * root = NULL, ** tipn = NULL, ** inner = NULL, ** edges = NULL;
  unsigned int * order = NULL;
  stepwise_t sw;
  pll_utree_t * tree = NULL;
  int failed = 0;

  if (!list || !count || !list[0])
  {
    pars_error(PLL_ERROR_PARAM_INVALID, "Stepwise parsimony needs at least one parsimony structure.");
    return NULL;
  }
  tips = list[0]->tips;
  inner_nodes = list[0]->inner_nodes;
  if (tips < 3)
  {
    pars_error(PLL_ERROR_STEPWISE_TIPS, "Stepwise parsimony requires at least three tips.");
    return NULL;
  }
  if (inner_nodes < tips - 2)
  {
    pars_error(PLL_ERROR_STEPWISE_UNSUPPORTED, "Stepwise parsimony currently supports only unrooted trees.");
    return NULL;
  }
  if (!sw_prepare(&sw, list, count)) return NULL;
  *cost = ~0u;

  tipn = (pll_unode_t **)calloc(tips + 1, sizeof(pll_unode_t *));
  inner = (pll_unode_t **)calloc(tips - 2, sizeof(pll_unode_t *));
  edges = (pll_unode_t **)calloc(2 * (size_t)tips - 3, sizeof(pll_unode_t *));
  order = sw_shuffled(tips, seed);
  root = sw_inner_create(tips - 3, tips);
  if (!tipn || !inner || !edges || !order || !root) failed = 1;
  for (i = 0; !failed && i + 3 < tips; ++i)
    if (!(inner[i] = sw_inner_create(i, tips))) failed = 1;
  for (i = 0; !failed && i < tips; ++i)
  {
    /* tip record i carries sequence order[i] (src/stepwise.c:975-1001) */
    tipn[i] = (pll_unode_t *)calloc(1, sizeof(pll_unode_t));
    if (tipn[i])
    {
      tipn[i]->clv_index = tipn[i]->node_index = order[i];
      tipn[i]->label = strdup(labels[order[i]]);
    }
    if (!tipn[i] || !tipn[i]->label)
      failed = 1;
    else if (i > 2)
      sw_link(inner[i - 3], tipn[i]);
  }
  if (failed)
  {
    pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    goto done;
  }

  /* the three-tip star */
  sw_link(root, tipn[0]);
  sw_link(root->next, tipn[1]);
  sw_link(root->next->next, tipn[2]);
  edges[0] = root;
  edges[1] = root->next;
  edges[2] = root->next->next;
  edge_count = 3;

  if (tips == 3)
  {
    /* src/stepwise.c:1054-1059 */
    *cost = 0;
    for (i = 0; i < count; ++i) *cost += list[i]->const_cost;
  }
  for (i = 3; i < tips && !failed; ++i)
  {
    *cost = sw_insert_best(&sw, edges, edge_count, inner[i - 3], NULL, NULL, &failed);
    edge_count += 2;
  }
  if (!failed) tree = pll_utree_wraptree(root, tips);

done:
  if (!tree)
  {
    /* every record is still reachable through the lists */
    for (i = 0; tipn && i < tips; ++i) sw_free_ring(tipn[i]);
    for (i = 0; inner && i + 3 < tips; ++i) sw_free_ring(inner[i]);
    sw_free_ring(root);
  }
  sw_release(&sw);
  free(tipn);
  free(inner);
  free(edges);
  free(order);
  return tree;
}

static int sw_cb_all(pll_unode_t * node)
{
  (void)node;
  return 1;
}

/* the edges of the tree around `root`, one record per edge, in the reference's order (src/stepwise.c:352-375):
 * the post-order node list with every tip replaced by the record it hangs on, minus the root itself */
static int sw_collect_edges(pll_unode_t * root, pll_unode_t ** edges, unsigned int * edge_count)
{
  unsigned int i;
  if (!pll_utree_traverse(root, PLL_TREE_TRAVERSE_POSTORDER, sw_cb_all, edges, edge_count)) return 0;
  for (i = 0; i < *edge_count; ++i)
    if (!edges[i]->next) edges[i] = edges[i]->back;
  (*edge_count)--;
  return 1;
}

/* src/stepwise.c:731-881: adds the taxa the tree does not have yet (those with index >= tree->tip_count in the
 * parsimony structures; labels[i] names taxon tip_count + i) by stepwise addition in shuffled order.  The tree's
 * node array is replaced; inner nodes keep their records, with clv/node indices shifted past the new tips. */
PLL_EXPORT int pll_fastparsimony_stepwise_extend(pll_utree_t * tree, pll_parsimony_t ** pars_list,
                                                 unsigned int pars_count, char * const * labels,
                                                 const unsigned int * tip_msa_idmap, unsigned int seed,
                                                 unsigned int * cost)
{
  stepwise_t sw;
  unsigned int new_tips, new_inner, old_tips, old_inner, ext, i, edge_count = 0;
  pll_unode_t ** nodes = NULL, ** edges = NULL;
  unsigned int * order = NULL;
  int failed = 0;

  if (!tree || !sw_prepare(&sw, pars_list, pars_count)) return PLL_FAILURE;
  new_tips = pars_list[0]->tips;
  new_inner = new_tips - 2;
  old_tips = tree->tip_count;
  old_inner = tree->inner_count;
  if (new_tips < old_tips || old_tips < 3)
  {
    sw_release(&sw);
    pars_error(PLL_ERROR_PARAM_INVALID, "The tree has more tips than the parsimony structures (or fewer than 3).");
    return PLL_FAILURE;
  }
  ext = new_tips - old_tips;
  nodes = (pll_unode_t **)calloc((size_t)new_tips + new_inner, sizeof(pll_unode_t *));
  edges = (pll_unode_t **)calloc(2 * (size_t)new_tips - 2, sizeof(pll_unode_t *));
  order = sw_shuffled(ext, seed);
  if (!nodes || !edges || !order) failed = 1;
  for (i = 0; !failed && i < ext; ++i)
  {
    /* new tip old_tips+order[i] hangs on new inner node old_inner+i */
    const unsigned int index = order[i] + old_tips;
    pll_unode_t * tip = (pll_unode_t *)calloc(1, sizeof(pll_unode_t));
    pll_unode_t * ring = sw_inner_create(old_inner + i, new_tips);
    nodes[old_tips + i] = tip;
    nodes[new_tips + old_inner + i] = ring;
    if (tip)
    {
      tip->clv_index = tip->node_index = index;
      tip->label = strdup(labels[index - old_tips]);
    }
    if (!tip || !ring || !tip->label)
      failed = 1;
    else
      sw_link(ring, tip);
  }
  if (failed)
  {
    for (i = 0; nodes && i < ext; ++i)
    {
      sw_free_ring(nodes[old_tips + i]);
      sw_free_ring(nodes[new_tips + old_inner + i]);
    }
    free(nodes);
    free(edges);
    free(order);
    sw_release(&sw);
    pars_error(PLL_ERROR_MEM_ALLOC, "Cannot allocate memory for nodes!");
    return PLL_FAILURE;
  }
  /* nothing below fails for lack of memory: the tree may now be changed */
  for (i = 0; i < old_tips; ++i) nodes[i] = tree->nodes[i];
  for (i = old_tips; i < old_tips + old_inner; ++i)
  {
    pll_unode_t * first = tree->nodes[i], * n = first;
    nodes[i + ext] = first;
    do
    {
      n->clv_index += ext;
      n->node_index += ext;
      n = n->next;
    } while (n != first);
  }
  if (tip_msa_idmap)
    for (i = 0; i < new_tips; ++i) nodes[i]->node_index = tip_msa_idmap[nodes[i]->node_index];

  if (!sw_collect_edges(tree->vroot, edges, &edge_count)) failed = 1;
  for (i = 0; i < ext && !failed; ++i)
  {
    *cost = sw_insert_best(&sw, edges, edge_count, nodes[new_tips + old_inner + i], NULL, NULL, &failed);
    edge_count += 2;
  }
  if (failed)
  {
    /* rings that were not placed are still only in the list */
    for (; i < ext; ++i)
    {
      pll_unode_t * ring = nodes[new_tips + old_inner + i];
      if (ring->next->back) continue;
      sw_free_ring(ring->back);
      sw_free_ring(ring);
      nodes[old_tips + i] = nodes[new_tips + old_inner + i] = NULL;
    }
  }
  free(tree->nodes);
  tree->nodes = nodes;
  tree->tip_count = new_tips;
  tree->inner_count = new_inner;
  tree->edge_count = 2 * new_tips - 3;
  tree->vroot = tree->vroot->next ? tree->vroot : tree->vroot->back;
  free(edges);
  free(order);
  sw_release(&sw);
  return failed ? PLL_FAILURE : PLL_SUCCESS;
}

/* src/stepwise.c:585-729: every subtree of the tree (three per inner node, in shuffled order) is pruned and put
 * back on the best edge of the rest; clv_index_map groups inner nodes for constrained searches (a subtree of
 * group g may only go next to group g); tip_msa_idmap renumbers tips whose order differs from the alignment's. */
PLL_EXPORT int pll_fastparsimony_stepwise_spr_round(pll_utree_t * tree, pll_parsimony_t ** pars_list,
                                                    unsigned int pars_count, const unsigned int * tip_msa_idmap,
                                                    unsigned int seed, const int * clv_index_map,
                                                    unsigned int * cost)
{
  stepwise_t sw;
  unsigned int tips, inner, node_count, subtrees, ext, i, edge_count = 0;
  pll_unode_t ** all = NULL, ** edges = NULL;
  unsigned int * constraint = NULL, * orig = NULL, * order = NULL;
  int failed = 0;

  if (!tree || !sw_prepare(&sw, pars_list, pars_count)) return PLL_FAILURE;
  tips = tree->tip_count;
  inner = tree->inner_count;
  node_count = tips + inner;
  subtrees = 3 * inner;
  if (pars_list[0]->tips < tips)
  {
    sw_release(&sw);
    pars_error(PLL_ERROR_PARAM_INVALID, "The tree has more tips than the parsimony structures.");
    return PLL_FAILURE;
  }
  ext = pars_list[0]->tips - tips;
  all = (pll_unode_t **)calloc((size_t)subtrees + 1, sizeof(pll_unode_t *));
  edges = (pll_unode_t **)calloc((size_t)tree->edge_count + 2, sizeof(pll_unode_t *));
  constraint = (unsigned int *)calloc((size_t)node_count + ext + 1, sizeof(unsigned int));
  orig = (unsigned int *)calloc((size_t)pars_list[0]->tips + 1, sizeof(unsigned int));
  order = sw_shuffled(subtrees, seed);
  if (!all || !edges || !constraint || !orig || !order)
  {
    free(all);
    free(edges);
    free(constraint);
    free(orig);
    free(order);
    sw_release(&sw);
    pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return PLL_FAILURE;
  }
  for (i = 0; i < node_count; ++i)
  {
    const unsigned int clv = tree->nodes[i]->clv_index;
    constraint[clv] = (tree->nodes[i]->next && clv_index_map) ? (unsigned int)(clv_index_map[clv] + 1) : 0;
  }
  if (tip_msa_idmap)
  {
    /* score indices follow the numbering of the parsimony structures */
    for (i = 0; i < tips; ++i)
    {
      const unsigned int old_idx = tree->nodes[i]->node_index, new_idx = tip_msa_idmap[old_idx];
      tree->nodes[i]->node_index = new_idx;
      orig[new_idx] = old_idx;
    }
    for (i = tips; i < node_count; ++i)
    {
      pll_unode_t * n = tree->nodes[i];
      n->node_index += ext;
      n->next->node_index += ext;
      n->next->next->node_index += ext;
    }
  }
  for (i = 0; i < inner; ++i)
  {
    pll_unode_t * n = tree->nodes[tips + i];
    all[3 * i] = n;
    all[3 * i + 1] = n->next;
    all[3 * i + 2] = n->next->next;
  }

  for (i = 0; i < subtrees && !failed; ++i)
  {
    pll_unode_t * v = all[order[i]], * prune_edge, * new_root;
    /* what is left must keep at least three taxa */
    if (!v->next->back->next && !v->next->next->back->next) continue;
    prune_edge = sw_prune(v);
    new_root = prune_edge->next ? prune_edge : prune_edge->back;
    if (!sw_collect_edges(new_root, edges, &edge_count))
    {
      sw_split(prune_edge, v->next, v->next->next);
      failed = 1;
      break;
    }
    {
      const unsigned int c = sw_insert_best(&sw, edges, edge_count, v, clv_index_map ? constraint : NULL, prune_edge,
                                            &failed);
      if (failed)
        sw_split(prune_edge, v->next, v->next->next); /* keep the tree whole */
      else
        *cost = c;
    }
  }

  if (tip_msa_idmap)
  {
    for (i = 0; i < tips; ++i) tree->nodes[i]->node_index = orig[tree->nodes[i]->node_index];
    for (i = tips; i < node_count; ++i)
    {
      pll_unode_t * n = tree->nodes[i];
      n->node_index -= ext;
      n->next->node_index -= ext;
      n->next->next->node_index -= ext;
    }
  }
  free(all);
  free(edges);
  free(constraint);
  free(orig);
  free(order);
  sw_release(&sw);
  return failed ? PLL_FAILURE : PLL_SUCCESS;
}
