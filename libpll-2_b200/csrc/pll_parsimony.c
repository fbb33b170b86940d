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
    plf_free(cp->ctx, cp->d_vec);
    plf_pars_destroy(cp->ps);
    plf_ctx_destroy(cp->ctx);
  }
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

PLL_EXPORT void pll_fastparsimony_update_vectors(pll_parsimony_t * parsimony, const pll_pars_buildop_t * ops,
                                                 unsigned int count)
{
  cuda_parsimony_t * cp = PP(parsimony);
  unsigned int * scores, i;
  if (!cp || !count) return;
  /* pll_pars_buildop_t is three consecutive unsigned ints: the list goes to the device as it is */
  if (!pars_indices_ok(cp, (const unsigned int *)ops, (size_t)3 * count)) return;
  scores = pars_scratch(cp, count);
  if (!scores) return;
  if (!plf_pars_update(cp->ps, cp->d_vec, cp->pub.states, cp->pub.packedvector_count, (const unsigned int *)ops, count,
                       scores))
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

/* ---- randomised stepwise addition (src/stepwise.c:883-1082) ------------------------------------------------
 *
 * Tips are added in a shuffled order; each new tip is tried on every edge of the current tree and stays on the
 * first edge of minimal parsimony length.  The reference evaluates an edge with one vector update plus one edge
 * score (two passes over the vectors and two function calls per edge and partition).  Here:
 *   1. every directional vector that the previous insertion invalidated is recomputed by ONE launch
 *      (plf_pars_update runs a whole dependency-ordered list),
 *   2. ALL candidate edges are scored by ONE launch (plf_pars_insert_scan: the merge of the edge's two
 *      vectors never leaves registers).
 * Costs are exact integers, the edge list grows in the reference's order and ties go to the first edge, so the
 * tree and its cost are those of the reference for the same seed.
 */

typedef struct stepwise
{
  cuda_parsimony_t ** pars;
  unsigned int pars_count;
  unsigned char * valid;    /* per directional vector (node_index) */
  pll_pars_buildop_t * ops;
  unsigned int ops_count;
  unsigned int * pairs;     /* 2 per candidate edge */
  unsigned int * scan;      /* per candidate edge */
  unsigned int * total;
} stepwise_t;

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

/* Fisher-Yates shuffle driven by the glibc-compatible generator (src/stepwise.c:56-106); seed 0 = identity */
static unsigned int * sw_shuffled(unsigned int n, unsigned int seed)
{
  unsigned int * x = (unsigned int *)malloc((size_t)n * sizeof(unsigned int));
  unsigned int i;
  if (!x) return NULL;
  for (i = 0; i < n; ++i) x[i] = i;
  if (seed && n > 1)
  {
    pll_random_state * rs = pll_random_create(seed);
    if (!rs)
    {
      free(x);
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

static void sw_free_nodes(pll_unode_t ** tipn, unsigned int tips, pll_unode_t ** inner, unsigned int inners,
                          pll_unode_t * root)
{
  unsigned int i;
  pll_unode_t * n;
  if (tipn)
    for (i = 0; i < tips; ++i)
      if (tipn[i])
      {
        free(tipn[i]->label);
        free(tipn[i]);
      }
  if (inner)
    for (i = 0; i < inners; ++i)
      if ((n = inner[i]))
      {
        free(n->next->next);
        free(n->next);
        free(n);
      }
  if (root)
  {
    free(root->next->next);
    free(root->next);
    free(root);
  }
}

PLL_EXPORT pll_utree_t * pll_fastparsimony_stepwise(pll_parsimony_t ** list, char * const * labels,
                                                    unsigned int * cost, unsigned int count, unsigned int seed)
{
  unsigned int tips, inner_nodes, i, k, e, edge_count, nvec;
  pll_unode_t * root = NULL, ** tipn = NULL, ** inner = NULL, ** edges = NULL;
  unsigned int * order = NULL;
  stepwise_t sw;
  pll_utree_t * tree = NULL;
  int failed = 0;

  memset(&sw, 0, sizeof(sw));
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
  for (i = 1; i < count; ++i)
    if (list[i]->tips != tips || list[i]->inner_nodes != inner_nodes)
    {
      pars_error(PLL_ERROR_STEPWISE_STRUCT, "Parsimony structures tips/inner nodes not equal.");
      return NULL;
    }
  *cost = ~0u;

  nvec = tips + 3 * inner_nodes;
  sw.pars_count = count;
  sw.pars = (cuda_parsimony_t **)calloc(count, sizeof(cuda_parsimony_t *));
  sw.valid = (unsigned char *)calloc(nvec, 1);
  sw.ops = (pll_pars_buildop_t *)malloc((size_t)3 * tips * sizeof(pll_pars_buildop_t));
  sw.pairs = (unsigned int *)malloc((size_t)4 * tips * sizeof(unsigned int));
  sw.scan = (unsigned int *)malloc((size_t)2 * tips * sizeof(unsigned int));
  sw.total = (unsigned int *)malloc((size_t)2 * tips * sizeof(unsigned int));
  tipn = (pll_unode_t **)calloc(tips + 1, sizeof(pll_unode_t *));
  inner = (pll_unode_t **)calloc(tips - 2, sizeof(pll_unode_t *));
  edges = (pll_unode_t **)calloc(2 * (size_t)tips - 3, sizeof(pll_unode_t *));
  order = sw_shuffled(tips, seed);
  root = sw_inner_create(tips - 3, tips);
  if (!sw.pars || !sw.valid || !sw.ops || !sw.pairs || !sw.scan || !sw.total || !tipn || !inner || !edges || !order ||
      !root)
    failed = 1;
  for (i = 0; !failed && i < count; ++i)
    if (!(sw.pars[i] = PP(list[i]))) failed = 2;
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
    if (failed == 1) pars_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    sw_free_nodes(tipn, tips, inner, tips - 3, root);
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

  for (i = 3; i < tips; ++i)
  {
    pll_unode_t * v = inner[i - 3]; /* v->back is the tip to insert, v->next / v->next->next are free */
    unsigned int best = 0, min_cost = ~0u;

    /* 1. bring every directional vector of the current tree up to date: one launch per partition */
    sw.ops_count = 0;
    for (e = 0; e < edge_count; ++e)
    {
      sw_collect(&sw, edges[e]);
      sw_collect(&sw, edges[e]->back);
    }
    pll_errno = 0;
    for (k = 0; k < count; ++k)
    {
      pll_fastparsimony_update_vectors(list[k], sw.ops, sw.ops_count);
      if (pll_errno) failed = 2;
    }

    /* 2. score the new tip on every edge: one launch per partition */
    for (e = 0; e < edge_count; ++e)
    {
      sw.pairs[2 * e] = edges[e]->node_index;
      sw.pairs[2 * e + 1] = edges[e]->back->node_index;
      sw.total[e] = 0;
    }
    for (k = 0; k < count && !failed; ++k)
    {
      cuda_parsimony_t * cp = sw.pars[k];
      if (!plf_pars_insert_scan(cp->ps, cp->d_vec, cp->pub.states, cp->pub.packedvector_count, sw.pairs, edge_count,
                                v->back->node_index, sw.scan))
      {
        pars_cuda_fail(cp);
        failed = 2;
        break;
      }
      for (e = 0; e < edge_count; ++e)
        sw.total[e] += sw.scan[e] + cp->pub.node_cost[sw.pairs[2 * e]] + cp->pub.node_cost[sw.pairs[2 * e + 1]] +
                       cp->pub.node_cost[v->back->node_index] + cp->pub.const_cost;
    }
    if (failed) break;
    for (e = 0; e < edge_count; ++e)
      if (sw.total[e] < min_cost)
      {
        min_cost = sw.total[e];
        best = e;
      }

    /* 3. place it; the two new edges go to the end of the list (src/stepwise.c:546-553) */
    sw_split(edges[best], v->next, v->next->next);
    edges[edge_count++] = v;
    edges[edge_count++] = v->next->next;
    *cost = min_cost;

    /* 4. only the vectors that look away from the new tip are still right */
    memset(sw.valid, 0, nvec);
    sw_validate_below(&sw, v);
    sw.valid[v->node_index] = 0; /* never computed: the scan kept the merge in registers */
  }

  if (failed)
  {
    /* the records are all linked into one graph or still in the lists: free them through the lists */
    sw_free_nodes(tipn, tips, inner, tips - 3, root);
    goto done;
  }
  tree = pll_utree_wraptree(root, tips);
  if (!tree) sw_free_nodes(tipn, tips, inner, tips - 3, root);

done:
  free(sw.pars);
  free(sw.valid);
  free(sw.ops);
  free(sw.pairs);
  free(sw.scan);
  free(sw.total);
  free(tipn);
  free(inner);
  free(edges);
  free(order);
  return tree;
}
