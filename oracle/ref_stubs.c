/* Test infrastructure, NOT product code and NOT reference code.
 *
 * The reference defines pll_utree_wraptree / pll_utree_wraptree_multi in its bison grammar file
 * (src/parse_utree.y), which cannot be generated in this image (no bison/flex).  src/utree.c refers
 * to them only from pll_utree_clone / pll_rtree_unroot and from src/stepwise.c.  The symbols below exist so that
 * oracle/_ref/libpll_ref.so loads with the reference's own traversal and stepwise-addition code inside;
 * they are our own minimal helpers, not reference code. */
#include <stdio.h>
#include <stdlib.h>

/* layouts of src/pll.h:388-411 (pointer-compatible minimal restatement for the two helpers below) */
typedef struct stub_unode
{
  char * label;
  double length;
  unsigned int node_index;
  unsigned int clv_index;
  int scaler_index;
  unsigned int pmatrix_index;
  struct stub_unode * next;
  struct stub_unode * back;
  void * data;
} stub_unode_t;

typedef struct stub_utree
{
  unsigned int tip_count;
  unsigned int inner_count;
  unsigned int edge_count;
  int binary;
  stub_unode_t ** nodes;
  stub_unode_t * vroot;
} stub_utree_t;

static void stub_collect(stub_unode_t * n, stub_unode_t ** tips, unsigned int * ntips, stub_unode_t ** inner,
                         unsigned int * ninner)
{
  if (!n->next)
  {
    tips[(*ntips)++] = n;
    return;
  }
  stub_collect(n->next->back, tips, ntips, inner, ninner);
  stub_collect(n->next->next->back, tips, ntips, inner, ninner);
  inner[(*ninner)++] = n;
}

/* Binary unrooted trees only (what pll_fastparsimony_stepwise hands over): tips first, then inner nodes, in
 * post-order from root->back then root.  The order of `nodes` is NOT the reference's; tests that use this
 * build compare topologies (Newick through our own exporter, or split sets), never node order. */
void * pll_utree_wraptree(void * vroot, unsigned int tip_count)
{
  stub_unode_t * root = (stub_unode_t *)vroot;
  stub_utree_t * t = (stub_utree_t *)calloc(1, sizeof(stub_utree_t));
  stub_unode_t ** tips = (stub_unode_t **)calloc(tip_count + 1, sizeof(void *));
  stub_unode_t ** inner = (stub_unode_t **)calloc(tip_count + 1, sizeof(void *));
  unsigned int nt = 0, ni = 0, i;
  if (root->back->next)
    stub_collect(root->back, tips, &nt, inner, &ni);
  else
    tips[nt++] = root->back;
  stub_collect(root, tips, &nt, inner, &ni);
  t->tip_count = nt;
  t->inner_count = ni;
  t->edge_count = nt + ni - 1;
  t->binary = 1;
  t->vroot = root;
  t->nodes = (stub_unode_t **)calloc(nt + ni, sizeof(void *));
  for (i = 0; i < nt; ++i) t->nodes[i] = tips[i];
  for (i = 0; i < ni; ++i) t->nodes[nt + i] = inner[i];
  free(tips);
  free(inner);
  return t;
}

void * pll_utree_wraptree_multi(void * root, unsigned int tip_count, unsigned int inner_count)
{
  (void)inner_count;
  return pll_utree_wraptree(root, tip_count);
}

/* only reached on the reference's allocation-failure paths */
void pll_utree_graph_destroy(void * root, void (*cb_destroy)(void *))
{
  (void)root;
  (void)cb_destroy;
}

/* src/parse_utree.y: frees what the wrapper above allocated and the node records of a binary tree */
void pll_utree_destroy(void * vtree, void (*cb_destroy)(void *))
{
  stub_utree_t * t = (stub_utree_t *)vtree;
  unsigned int i;
  if (!t) return;
  for (i = 0; i < t->tip_count + t->inner_count; ++i)
  {
    stub_unode_t * n = t->nodes[i];
    if (n->next)
    {
      stub_unode_t * a = n->next, * b = n->next->next;
      if (cb_destroy && a->data) cb_destroy(a->data);
      if (cb_destroy && b->data) cb_destroy(b->data);
      free(a);
      free(b);
    }
    if (cb_destroy && n->data) cb_destroy(n->data);
    free(n->label);
    free(n);
  }
  free(t->nodes);
  free(t);
}
