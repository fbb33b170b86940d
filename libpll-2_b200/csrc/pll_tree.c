/*
 * pll_tree.c -- the operation-list producers around the hot path: unrooted and
 * rooted tree structures, Newick reader/writer, traversals and the translation
 * of a traversal into pll_operation_t lists.
 *
 * Same API, struct layouts and index conventions as the reference
 * (src/pll.h:392-438, 890-1030; src/utree.c:305-463; src/rtree.c:262-387;
 * src/parse_utree.y, src/parse_rtree.y), so that clients (and the reference's
 * examples) produce the same operation lists.  The reference generates its
 * Newick reader with bison/flex; this one is a hand-written recursive-descent
 * parser over the same grammar:
 *
 *   input     := list [label] [':' number] ';'
 *   list      := '(' subtree {',' subtree} ')'
 *   subtree   := list [label] [':' number]  |  label [':' number]
 *   label     := unquoted string | number | 'quoted' | "quoted"
 *
 * Traversals are iterative (explicit stack): caterpillar trees of 10^5 taxa
 * do not overflow the C stack.
 *
 * Host code only: nothing here touches the device.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll_b200.h"

static void tree_error(int code, const char * msg)
{
  pll_errno = code;
  snprintf(pll_errmsg, 200, "%s", msg);
}

/* ---- growable string ------------------------------------------------------- */

typedef struct
{
  char * s;
  size_t len, cap;
  int failed;
} sbuf_t;

static void sb_put(sbuf_t * b, const char * s, size_t n)
{
  if (b->failed) return;
  if (b->len + n + 1 > b->cap)
  {
    size_t cap = b->cap ? b->cap : 256;
    char * p;
    while (cap < b->len + n + 1) cap *= 2;
    p = (char *)realloc(b->s, cap);
    if (!p)
    {
      b->failed = 1;
      return;
    }
    b->s = p;
    b->cap = cap;
  }
  memcpy(b->s + b->len, s, n);
  b->len += n;
  b->s[b->len] = 0;
}

static void sb_puts(sbuf_t * b, const char * s) { sb_put(b, s, strlen(s)); }

static void sb_label_length(sbuf_t * b, const char * label, double length)
{
  char num[64];
  if (label) sb_puts(b, label);
  snprintf(num, sizeof(num), ":%f", length);
  sb_puts(b, num);
}

/* ---- Newick tokens ---------------------------------------------------------- */

typedef struct
{
  const char * p;
  int line, col;
  char err[200];
} nwk_t;

static void nwk_fail(nwk_t * k, const char * what)
{
  if (!k->err[0]) snprintf(k->err, sizeof(k->err), "syntax error, %s. (line %d column %d)\n", what, k->line, k->col);
}

static void nwk_skip(nwk_t * k)
{
  for (;; ++k->p)
  {
    if (*k->p == '\n')
    {
      ++k->line;
      k->col = 0;
    }
    else if (*k->p == ' ' || *k->p == '\t' || *k->p == '\r')
      ++k->col;
    else
      return;
  }
}

static int nwk_is_delim(char c) { return c == 0 || strchr(" \t\n\r()[],:;", c) != NULL; }

/* label := quoted | run of non-delimiter characters (numbers included); NULL when none follows */
static char * nwk_label(nwk_t * k)
{
  const char * start;
  size_t n;
  char * out;
  nwk_skip(k);
  if (*k->p == '\'' || *k->p == '"')
  {
    const char q = *k->p++;
    start = k->p;
    while (*k->p && *k->p != q)
    {
      if (*k->p == '\\' && k->p[1]) ++k->p; /* escaped character stays as written */
      ++k->p;
    }
    if (!*k->p)
    {
      nwk_fail(k, "unterminated quoted label");
      return NULL;
    }
    n = (size_t)(k->p - start);
    ++k->p;
  }
  else
  {
    if (nwk_is_delim(*k->p)) return NULL;
    start = k->p;
    while (!nwk_is_delim(*k->p)) ++k->p;
    n = (size_t)(k->p - start);
  }
  k->col += (int)n;
  out = (char *)malloc(n + 1);
  if (!out)
  {
    nwk_fail(k, "out of memory");
    return NULL;
  }
  memcpy(out, start, n);
  out[n] = 0;
  return out;
}

/* [':' number]; *present tells whether a length was given */
static double nwk_length(nwk_t * k, int * present)
{
  char * end;
  double v;
  *present = 0;
  nwk_skip(k);
  if (*k->p != ':') return 0;
  ++k->p;
  nwk_skip(k);
  v = strtod(k->p, &end);
  if (end == k->p || !(isdigit((unsigned char)k->p[0]) || k->p[0] == '.' || k->p[0] == '+' || k->p[0] == '-'))
  {
    nwk_fail(k, "a branch length must follow ':'");
    return 0;
  }
  k->col += (int)(end - k->p);
  k->p = end;
  *present = 1;
  return v;
}

/* ======================================================================== *
 *  unrooted trees                                                            *
 * ======================================================================== */

static pll_unode_t * unode_new(void) { return (pll_unode_t *)calloc(1, sizeof(pll_unode_t)); }

/* frees the subtree hanging below `node` (entered through node->back of the caller) */
static void ugraph_free_from(pll_unode_t * entry, void (*cb_destroy)(void *))
{
  /* explicit stack of nodes whose whole roundabout is to be released */
  size_t cap = 64, top = 0;
  pll_unode_t ** stack = (pll_unode_t **)malloc(cap * sizeof(*stack));
  if (!stack) return;
  stack[top++] = entry;
  while (top)
  {
    pll_unode_t * node = stack[--top];
    pll_unode_t * s = node;
    char * label = node->label;
    do
    {
      pll_unode_t * next = s->next;
      if (s != node && s->back)
      {
        if (top == cap)
        {
          pll_unode_t ** grown = (pll_unode_t **)realloc(stack, 2 * cap * sizeof(*stack));
          if (!grown) break;
          stack = grown;
          cap *= 2;
        }
        stack[top++] = s->back;
      }
      if (s->data && cb_destroy) cb_destroy(s->data);
      free(s);
      s = next;
    } while (s && s != node);
    free(label); /* the members of a roundabout share one label */
  }
  free(stack);
}

PLL_EXPORT void pll_utree_graph_destroy(pll_unode_t * root, void (*cb_destroy)(void *))
{
  if (!root) return;
  /* release what hangs behind the root first, then the root's own roundabout and subtrees */
  if (root->back)
  {
    pll_unode_t * behind = root->back;
    behind->back = NULL;
    root->back = NULL;
    /* `behind` is the up-link of its own roundabout (or a tip): everything but that link is below it */
    ugraph_free_from(behind, cb_destroy);
  }
  ugraph_free_from(root, cb_destroy);
}

PLL_EXPORT void pll_utree_destroy(pll_utree_t * tree, void (*cb_destroy)(void *))
{
  unsigned int i;
  if (!tree) return;
  for (i = 0; i < tree->tip_count + tree->inner_count; ++i)
  {
    pll_unode_t * first = tree->nodes[i];
    pll_unode_t * s = first;
    if (!first) continue;
    free(first->label);
    do
    {
      pll_unode_t * next = s->next;
      if (s->data && cb_destroy) cb_destroy(s->data);
      free(s);
      s = next;
    } while (s && s != first);
  }
  free(tree->nodes);
  free(tree);
}

/* closes the chain first -> ... -> last into a ring and shares the label */
static void close_ring(pll_unode_t * first)
{
  pll_unode_t * last = first;
  while (last->next)
  {
    if (!last->next->label) last->next->label = last->label;
    last = last->next;
  }
  last->next = first;
}

static pll_unode_t * parse_usubtree(nwk_t * k, unsigned int * tips);

/* list := '(' subtree {',' subtree} ')' ; returns the chain of link nodes (one per child), not yet closed */
static pll_unode_t * parse_ulist(nwk_t * k, unsigned int * tips)
{
  pll_unode_t * head = NULL, * tail = NULL;
  nwk_skip(k);
  if (*k->p != '(')
  {
    nwk_fail(k, "expecting '('");
    return NULL;
  }
  ++k->p;
  ++k->col;
  for (;;)
  {
    pll_unode_t * child = parse_usubtree(k, tips);
    pll_unode_t * link;
    if (!child) break;
    link = unode_new();
    if (!link)
    {
      nwk_fail(k, "out of memory");
      pll_utree_graph_destroy(child, NULL);
      break;
    }
    link->back = child;
    child->back = link;
    link->length = child->length;
    if (tail)
      tail->next = link;
    else
      head = link;
    tail = link;
    nwk_skip(k);
    if (*k->p == ',')
    {
      ++k->p;
      ++k->col;
      continue;
    }
    if (*k->p == ')')
    {
      ++k->p;
      ++k->col;
      return head;
    }
    nwk_fail(k, "expecting ',' or ')'");
    break;
  }
  /* failure: release the partial chain */
  while (head)
  {
    pll_unode_t * next = head->next;
    if (head->back)
    {
      head->back->back = NULL;
      ugraph_free_from(head->back, NULL);
    }
    free(head);
    head = next;
  }
  return NULL;
}

static pll_unode_t * parse_usubtree(nwk_t * k, unsigned int * tips)
{
  pll_unode_t * node;
  int has_len;
  nwk_skip(k);
  if (*k->p == '(')
  {
    pll_unode_t * chain = parse_ulist(k, tips);
    if (!chain) return NULL;
    node = unode_new();
    if (!node)
    {
      nwk_fail(k, "out of memory");
      return NULL;
    }
    node->next = chain;
    node->label = nwk_label(k);
    node->length = nwk_length(k, &has_len);
    close_ring(node);
    if (k->err[0])
    {
      node->back = NULL;
      ugraph_free_from(node, NULL);
      return NULL;
    }
    return node;
  }
  node = unode_new();
  if (!node)
  {
    nwk_fail(k, "out of memory");
    return NULL;
  }
  node->label = nwk_label(k);
  if (!node->label)
  {
    nwk_fail(k, "expecting a label or '('");
    free(node);
    return NULL;
  }
  node->length = nwk_length(k, &has_len);
  if (k->err[0])
  {
    free(node->label);
    free(node);
    return NULL;
  }
  ++*tips;
  return node;
}

static int unode_is_rooted(const pll_unode_t * root) { return root->next && root->next->next == root; }

PLL_EXPORT int pll_utree_is_rooted(const pll_utree_t * tree) { return unode_is_rooted(tree->vroot); }

/* src/parse_utree.y:537-567: a bifurcating root becomes one edge */
PLL_EXPORT pll_unode_t * pll_utree_unroot_inplace(pll_unode_t * root)
{
  pll_unode_t * left, * right;
  double length;
  if (!unode_is_rooted(root)) return root;
  if (root->next == root)
  {
    tree_error(PLL_ERROR_NEWICK_SYNTAX, "Unifurcation detected at root");
    return NULL;
  }
  left = root->back;
  right = root->next->back;
  free(root->label);
  free(root->next);
  free(root);
  length = left->length + right->length;
  left->back = right;
  right->back = left;
  left->length = right->length = length;
  left->pmatrix_index = right->pmatrix_index =
      left->pmatrix_index < right->pmatrix_index ? left->pmatrix_index : right->pmatrix_index;
  return left->next ? left : right;
}

/* post-order walk used by index assignment, node collection and counting: at the entry node every link
 * is followed, below it every link but the one we came through (src/parse_utree.y:270-372).
 * visit(node, level, ctx) is called after the node's subtrees. */
typedef void (*uvisit_t)(pll_unode_t * node, unsigned int level, void * ctx);

typedef struct
{
  pll_unode_t * node, * cursor;
  unsigned int level;
  int started;
} uframe_t;

static int uwalk_postorder(pll_unode_t * root, uvisit_t visit, void * ctx)
{
  size_t cap = 64, top = 0;
  uframe_t * stack = (uframe_t *)malloc(cap * sizeof(*stack));
  if (!stack) return 0;
  stack[top].node = root;
  stack[top].cursor = NULL;
  stack[top].level = 0;
  stack[top].started = 0;
  ++top;
  while (top)
  {
    uframe_t * f = &stack[top - 1];
    pll_unode_t * child = NULL;
    if (f->node->next)
    {
      if (!f->started)
      {
        f->cursor = f->level ? f->node->next : f->node;
        f->started = 1;
        child = f->cursor->back;
      }
      else
      {
        f->cursor = f->cursor->next;
        if (f->cursor != f->node) child = f->cursor->back;
      }
    }
    if (child)
    {
      const unsigned int level = f->level + 1;
      if (top == cap)
      {
        uframe_t * grown = (uframe_t *)realloc(stack, 2 * cap * sizeof(*stack));
        if (!grown)
        {
          free(stack);
          return 0;
        }
        stack = grown;
        cap *= 2;
      }
      stack[top].node = child;
      stack[top].cursor = NULL;
      stack[top].level = level;
      stack[top].started = 0;
      ++top;
      continue;
    }
    visit(f->node, f->level, ctx);
    --top;
  }
  free(stack);
  return 1;
}

typedef struct
{
  unsigned int tip_clv, inner_clv, inner_node;
  int inner_scaler;
} uindex_ctx_t;

static void visit_assign(pll_unode_t * node, unsigned int level, void * vctx)
{
  uindex_ctx_t * c = (uindex_ctx_t *)vctx;
  pll_unode_t * s;
  if (!node->next)
  {
    node->node_index = node->clv_index = node->pmatrix_index = c->tip_clv++;
    node->scaler_index = PLL_SCALE_BUFFER_NONE;
    return;
  }
  s = node;
  do
  {
    s->node_index = c->inner_node++;
    s->clv_index = c->inner_clv;
    s->scaler_index = c->inner_scaler;
    /* the link towards the root owns the edge index, the others take their child's */
    s->pmatrix_index = (s == node && level > 0) ? c->inner_clv : s->back->pmatrix_index;
    s = s->next;
  } while (s != node);
  ++c->inner_clv;
  ++c->inner_scaler;
}

PLL_EXPORT void pll_utree_reset_template_indices(pll_unode_t * root, unsigned int tip_count)
{
  uindex_ctx_t c;
  c.tip_clv = 0;
  c.inner_clv = c.inner_node = tip_count;
  c.inner_scaler = 0;
  if (!root->next) root = root->back;
  uwalk_postorder(root, visit_assign, &c);
}

typedef struct
{
  pll_unode_t ** nodes;
  unsigned int tips, inner, tip_base, inner_base;
} ucollect_ctx_t;

static void visit_count(pll_unode_t * node, unsigned int level, void * vctx)
{
  ucollect_ctx_t * c = (ucollect_ctx_t *)vctx;
  (void)level;
  if (node->next)
    ++c->inner;
  else
    ++c->tips;
}

static void visit_collect(pll_unode_t * node, unsigned int level, void * vctx)
{
  ucollect_ctx_t * c = (ucollect_ctx_t *)vctx;
  (void)level;
  if (node->next)
    c->nodes[c->inner_base + c->inner++] = node;
  else
    c->nodes[c->tip_base + c->tips++] = node;
}

static pll_utree_t * utree_wrap(pll_unode_t * root, unsigned int tip_count, unsigned int inner_count, int binary)
{
  pll_utree_t * tree;
  ucollect_ctx_t c;
  if (tip_count < 3 && tip_count != 0)
  {
    pll_errno = PLL_ERROR_PARAM_INVALID;
    snprintf(pll_errmsg, 200, "Invalid tip_count value (%u).", tip_count);
    return NULL;
  }
  if (!root->next) root = root->back;
  memset(&c, 0, sizeof(c));
  if (tip_count == 0 || (!binary && inner_count == 0))
  {
    if (!root->next)
    {
      tree_error(PLL_ERROR_PARAM_INVALID, "Input tree contains no inner nodes.");
      return NULL;
    }
    if (!uwalk_postorder(root, visit_count, &c))
    {
      tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
      return NULL;
    }
    tip_count = c.tips;
    inner_count = c.inner;
    if (binary && inner_count != tip_count - 2)
    {
      tree_error(PLL_ERROR_PARAM_INVALID, "Input tree is not strictly bifurcating.");
      return NULL;
    }
  }
  else if (binary)
    inner_count = tip_count - 2;
  tree = (pll_utree_t *)malloc(sizeof(pll_utree_t));
  if (tree) tree->nodes = (pll_unode_t **)malloc(((size_t)tip_count + inner_count) * sizeof(pll_unode_t *));
  if (!tree || !tree->nodes)
  {
    free(tree);
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  memset(&c, 0, sizeof(c));
  c.nodes = tree->nodes;
  c.tip_base = 0;
  c.inner_base = tip_count;
  uwalk_postorder(root, visit_collect, &c);
  tree->tip_count = tip_count;
  tree->inner_count = inner_count;
  tree->edge_count = tip_count + inner_count - 1;
  tree->binary = (inner_count == tip_count - (unode_is_rooted(root) ? 1 : 2));
  tree->vroot = root;
  return tree;
}

PLL_EXPORT pll_utree_t * pll_utree_wraptree(pll_unode_t * root, unsigned int tip_count)
{
  return utree_wrap(root, tip_count, 0, 1);
}

PLL_EXPORT pll_utree_t * pll_utree_wraptree_multi(pll_unode_t * root, unsigned int tip_count, unsigned int inner_count)
{
  return utree_wrap(root, tip_count, inner_count, 0);
}

static pll_utree_t * utree_from_string(const char * s, int auto_unroot, int allow_rooted)
{
  nwk_t k;
  pll_unode_t * chain, * root;
  unsigned int tips = 0;
  int has_len;
  memset(&k, 0, sizeof(k));
  k.p = s;
  k.line = 1;
  chain = parse_ulist(&k, &tips);
  if (!chain)
  {
    tree_error(PLL_ERROR_NEWICK_SYNTAX, k.err[0] ? k.err : "syntax error");
    return NULL;
  }
  /* the first link of the top-level list is the (virtual) root node itself (src/parse_utree.y:188-201) */
  root = chain;
  root->label = nwk_label(&k);
  (void)nwk_length(&k, &has_len); /* a root length is ignored: the structure is unrooted */
  close_ring(root);
  nwk_skip(&k);
  if (!k.err[0] && *k.p != ';') nwk_fail(&k, "expecting ';'");
  if (k.err[0])
  {
    pll_utree_graph_destroy(root, NULL);
    tree_error(PLL_ERROR_NEWICK_SYNTAX, k.err);
    return NULL;
  }
  if (auto_unroot)
  {
    root = pll_utree_unroot_inplace(root);
    if (!root) return NULL;
  }
  if (unode_is_rooted(root) && !allow_rooted)
  {
    pll_utree_graph_destroy(root, NULL);
    tree_error(PLL_ERROR_TREE_INVALID, "Rooted tree parsed but unrooted tree is expected.");
    return NULL;
  }
  pll_utree_reset_template_indices(root, tips);
  return utree_wrap(root, 0, 0, 0);
}

static char * read_whole_file(const char * filename)
{
  FILE * fp = fopen(filename, "r");
  long size;
  char * text;
  if (!fp)
  {
    pll_errno = PLL_ERROR_FILE_OPEN;
    snprintf(pll_errmsg, 200, "Unable to open file (%s)", filename);
    return NULL;
  }
  fseek(fp, 0, SEEK_END);
  size = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  text = (char *)malloc((size_t)size + 1);
  if (!text || fread(text, 1, (size_t)size, fp) != (size_t)size)
  {
    free(text);
    fclose(fp);
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  text[size] = 0;
  fclose(fp);
  return text;
}

static pll_utree_t * utree_from_file(const char * filename, int auto_unroot, int allow_rooted)
{
  char * text = read_whole_file(filename);
  pll_utree_t * tree;
  if (!text) return NULL;
  tree = utree_from_string(text, auto_unroot, allow_rooted);
  free(text);
  return tree;
}

PLL_EXPORT pll_utree_t * pll_utree_parse_newick(const char * filename) { return utree_from_file(filename, 0, 0); }
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_rooted(const char * filename) { return utree_from_file(filename, 0, 1); }
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_unroot(const char * filename) { return utree_from_file(filename, 1, 0); }
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_string(const char * s) { return utree_from_string(s, 0, 0); }
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_string_rooted(const char * s) { return utree_from_string(s, 0, 1); }
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_string_unroot(const char * s) { return utree_from_string(s, 1, 0); }

/* ---- traversal (src/utree.c:388-451) ---------------------------------------------- *
 * The subtree behind the root edge first, then the root's own; at every node the        *
 * callback decides whether its subtree is entered (partial traversals).                 */
static int utraverse_from(pll_unode_t * start, int traversal, int (*cbtrav)(pll_unode_t *), pll_unode_t ** out,
                          unsigned int * index)
{
  size_t cap = 64, top = 0;
  struct frame
  {
    pll_unode_t * node, * cursor;
  } * stack = (struct frame *)malloc(cap * sizeof(*stack));
  if (!stack) return 0;
  if (cbtrav(start))
  {
    if (traversal == PLL_TREE_TRAVERSE_PREORDER) out[(*index)++] = start;
    stack[top].node = start;
    stack[top].cursor = start;
    ++top;
  }
  while (top)
  {
    struct frame * f = &stack[top - 1];
    pll_unode_t * child = NULL;
    if (f->node->next)
    {
      f->cursor = f->cursor->next;
      if (f->cursor && f->cursor != f->node) child = f->cursor->back;
    }
    if (child)
    {
      if (!cbtrav(child)) continue;
      if (traversal == PLL_TREE_TRAVERSE_PREORDER) out[(*index)++] = child;
      if (top == cap)
      {
        struct frame * grown = (struct frame *)realloc(stack, 2 * cap * sizeof(*stack));
        if (!grown)
        {
          free(stack);
          return 0;
        }
        stack = grown;
        cap *= 2;
      }
      stack[top].node = child;
      stack[top].cursor = child;
      ++top;
      continue;
    }
    if (traversal == PLL_TREE_TRAVERSE_POSTORDER) out[(*index)++] = f->node;
    --top;
  }
  free(stack);
  return 1;
}

PLL_EXPORT int pll_utree_traverse(pll_unode_t * root, int traversal, int (*cbtrav)(pll_unode_t *),
                                  pll_unode_t ** outbuffer, unsigned int * trav_size)
{
  *trav_size = 0;
  if (!root->next) return PLL_FAILURE;
  if (traversal != PLL_TREE_TRAVERSE_POSTORDER && traversal != PLL_TREE_TRAVERSE_PREORDER)
  {
    tree_error(PLL_ERROR_PARAM_INVALID, "Invalid traversal value.");
    return PLL_FAILURE;
  }
  if (!utraverse_from(root->back, traversal, cbtrav, outbuffer, trav_size) ||
      !utraverse_from(root, traversal, cbtrav, outbuffer, trav_size))
  {
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return PLL_FAILURE;
  }
  return PLL_SUCCESS;
}

/* src/utree.c:317-366: one CLV update per inner node of the traversal, one P-matrix per node except the
 * far end of the root edge (its matrix is the root's) */
PLL_EXPORT void pll_utree_create_operations(pll_unode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                            double * branches, unsigned int * pmatrix_indices, pll_operation_t * ops,
                                            unsigned int * matrix_count, unsigned int * ops_count)
{
  unsigned int i, n_ops = 0, n_mat = 0;
  const pll_unode_t * root_far = trav_buffer_size ? trav_buffer[trav_buffer_size - 1]->back : NULL;
  for (i = 0; i < trav_buffer_size; ++i)
  {
    const pll_unode_t * node = trav_buffer[i];
    if (node != root_far)
    {
      if (branches) branches[n_mat] = node->length;
      if (pmatrix_indices) pmatrix_indices[n_mat] = node->pmatrix_index;
      ++n_mat;
    }
    if (node->next)
    {
      const pll_unode_t * c1 = node->next->back, * c2 = node->next->next->back;
      pll_operation_t * op = ops + n_ops++;
      op->parent_clv_index = node->clv_index;
      op->parent_scaler_index = node->scaler_index;
      op->child1_clv_index = c1->clv_index;
      op->child1_scaler_index = c1->scaler_index;
      op->child1_matrix_index = c1->pmatrix_index;
      op->child2_clv_index = c2->clv_index;
      op->child2_scaler_index = c2->scaler_index;
      op->child2_matrix_index = c2->pmatrix_index;
    }
  }
  *ops_count = n_ops;
  if (matrix_count) *matrix_count = n_mat;
}

PLL_EXPORT int pll_utree_every(pll_utree_t * tree, int (*cb)(const pll_utree_t *, const pll_unode_t *))
{
  unsigned int i;
  int rc = 1;
  for (i = 0; i < tree->tip_count + tree->inner_count; ++i) rc &= cb(tree, tree->nodes[i]);
  return rc ? PLL_SUCCESS : PLL_FAILURE;
}

/* every edge is symmetric, every roundabout shares its clv / scaler index, no index is used twice */
PLL_EXPORT int pll_utree_check_integrity(const pll_utree_t * tree)
{
  unsigned int i;
  for (i = 0; i < tree->tip_count + tree->inner_count; ++i)
  {
    const pll_unode_t * first = tree->nodes[i], * s = first;
    if ((i < tree->tip_count) != (first->next == NULL))
    {
      tree_error(PLL_ERROR_TREE_INVALID, "tip / inner node in the wrong part of the node array");
      return PLL_FAILURE;
    }
    do
    {
      if (!s->back || s->back->back != s || s->back->length != s->length || s->back->pmatrix_index != s->pmatrix_index)
      {
        pll_errno = PLL_ERROR_TREE_INVALID;
        snprintf(pll_errmsg, 200, "Inconsistent edge at node with clv_index %u", s->clv_index);
        return PLL_FAILURE;
      }
      if (s->clv_index != first->clv_index || s->scaler_index != first->scaler_index)
      {
        pll_errno = PLL_ERROR_TREE_INVALID;
        snprintf(pll_errmsg, 200, "Inconsistent indices inside the node with clv_index %u", first->clv_index);
        return PLL_FAILURE;
      }
      s = s->next;
    } while (s && s != first);
  }
  return PLL_SUCCESS;
}

/* ---- Newick export (src/utree.c:160-300) ---------------------------------------------- */

static void uexport(sbuf_t * b, const pll_unode_t * node, unsigned int level, char * (*cb)(const pll_unode_t *))
{
  /* iterative: a frame per open inner node */
  size_t cap = 64, top = 0;
  struct frame
  {
    const pll_unode_t * node, * cursor;
    unsigned int level;
  } * stack = (struct frame *)malloc(cap * sizeof(*stack));
  if (!stack)
  {
    b->failed = 1;
    return;
  }
  stack[top].node = node;
  stack[top].cursor = NULL;
  stack[top].level = level;
  ++top;
  while (top && !b->failed)
  {
    struct frame * f = &stack[top - 1];
    const pll_unode_t * n = f->node;
    if (!n->next)
    {
      if (cb)
      {
        char * s = cb(n);
        sb_puts(b, s ? s : "");
        free(s);
      }
      else
        sb_label_length(b, n->label, n->length);
      --top;
      continue;
    }
    if (!f->cursor)
    {
      if (f->level > 0) sb_puts(b, "(");
      f->cursor = n->next;
    }
    else
    {
      f->cursor = f->cursor->next;
      if (f->cursor != n) sb_puts(b, ",");
    }
    if (f->cursor != n)
    {
      const unsigned int lvl = f->level + 1;
      if (top == cap)
      {
        struct frame * grown = (struct frame *)realloc(stack, 2 * cap * sizeof(*stack));
        if (!grown)
        {
          b->failed = 1;
          break;
        }
        stack = grown;
        cap *= 2;
        f = &stack[top - 1];
      }
      stack[top].node = f->cursor->back;
      stack[top].cursor = NULL;
      stack[top].level = lvl;
      ++top;
      continue;
    }
    if (f->level > 0)
    {
      sb_puts(b, ")");
      if (cb)
      {
        char * s = cb(n);
        sb_puts(b, s ? s : "");
        free(s);
      }
      else
        sb_label_length(b, n->label, n->length);
    }
    --top;
  }
  free(stack);
}

static char * utree_export(const pll_unode_t * root, int rooted, double root_brlen, char * (*cb)(const pll_unode_t *))
{
  sbuf_t b;
  memset(&b, 0, sizeof(b));
  if (!root) return NULL;
  if (!root->next) root = root->back;
  sb_puts(&b, "(");
  uexport(&b, root->back, 1, cb);
  sb_puts(&b, rooted ? ",(" : ",");
  uexport(&b, root, 0, cb);
  if (rooted)
  {
    sb_puts(&b, ")");
    sb_label_length(&b, root->label, root_brlen);
    sb_puts(&b, ");");
  }
  else
  {
    sb_puts(&b, ")");
    if (root->label) sb_puts(&b, root->label);
    sb_puts(&b, ";");
  }
  if (b.failed)
  {
    free(b.s);
    tree_error(PLL_ERROR_MEM_ALLOC, "memory allocation during newick export failed");
    return NULL;
  }
  return b.s;
}

PLL_EXPORT char * pll_utree_export_newick(const pll_unode_t * root, char * (*cb_serialize)(const pll_unode_t *))
{
  return utree_export(root, 0, 0, cb_serialize);
}

PLL_EXPORT char * pll_utree_export_newick_rooted(const pll_unode_t * root, double root_brlen)
{
  return utree_export(root, 1, root_brlen, NULL);
}

/* ---- ASCII rendering (src/utree.c:26-160) ------------------------------------------------ *
 * One spacer line and one "+---" line per node, children indented by four columns, a '|' in   *
 * every column whose branch still has children to come.  Iterative pre-order.                 */
static void show_node_info(const pll_unode_t * node, int options)
{
  if (options & PLL_UTREE_SHOW_LABEL) printf(" %s", node->label ? node->label : "(null)");
  if (options & PLL_UTREE_SHOW_BRANCH_LENGTH) printf(" %f", node->length);
  if (options & PLL_UTREE_SHOW_CLV_INDEX) printf(" %u", node->clv_index);
  if (options & PLL_UTREE_SHOW_SCALER_INDEX) printf(" %d", node->scaler_index);
  if (options & PLL_UTREE_SHOW_PMATRIX_INDEX) printf(" %u", node->pmatrix_index);
  if (options & PLL_UTREE_SHOW_DATA) printf(" %p", node->data);
  printf("\n");
}

PLL_EXPORT void pll_utree_show_ascii(const pll_unode_t * root, int options)
{
  struct item
  {
    const pll_unode_t * node;
    unsigned int depth;
    int last;
  } * stack;
  size_t cap = 64, top = 0, bars_cap = 64;
  unsigned char * is_last; /* is_last[d]: the node of the current path at depth d is its parent's last child */
  const pll_unode_t * s;
  unsigned int i;
  if (!root->next) root = root->back;
  stack = (struct item *)malloc(cap * sizeof(*stack));
  is_last = (unsigned char *)calloc(bars_cap, 1);
  if (!stack || !is_last)
  {
    free(stack);
    free(is_last);
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return;
  }
  /* children of the virtual root, pushed in reverse so that they pop in order */
  {
    size_t n = 0, k;
    s = root;
    do
    {
      ++n;
      s = s->next;
    } while (s != root);
    while (cap < n) cap *= 2;
    stack = (struct item *)realloc(stack, cap * sizeof(*stack));
    for (k = 0, s = root; k < n; ++k, s = s->next)
    {
      stack[n - 1 - k].node = s->back;
      stack[n - 1 - k].depth = 1;
      stack[n - 1 - k].last = (s->next == root);
    }
    top = n;
  }
  while (top)
  {
    const struct item it = stack[--top];
    const unsigned int d = it.depth;
    if (d >= bars_cap)
    {
      unsigned char * grown = (unsigned char *)realloc(is_last, bars_cap * 2);
      if (!grown) break;
      memset(grown + bars_cap, 0, bars_cap);
      is_last = grown;
      bars_cap *= 2;
    }
    is_last[d] = (unsigned char)it.last;
    for (i = 0; i + 1 < d; ++i) printf(is_last[i + 1] ? "    " : "|   ");
    printf("|   \n");
    for (i = 0; i + 1 < d; ++i) printf(is_last[i + 1] ? "    " : "|   ");
    printf(it.node->next ? "+---+" : "+---");
    show_node_info(it.node, options);
    if (it.node->next)
    {
      size_t n = 0, k;
      for (s = it.node->next; s != it.node; s = s->next) ++n;
      if (top + n > cap)
      {
        struct item * grown;
        while (cap < top + n) cap *= 2;
        grown = (struct item *)realloc(stack, cap * sizeof(*stack));
        if (!grown) break;
        stack = grown;
      }
      for (k = 0, s = it.node->next; k < n; ++k, s = s->next)
      {
        stack[top + n - 1 - k].node = s->back;
        stack[top + n - 1 - k].depth = d + 1;
        stack[top + n - 1 - k].last = (s->next == it.node);
      }
      top += n;
    }
  }
  free(stack);
  free(is_last);
}

/* ======================================================================== *
 *  rooted trees                                                              *
 * ======================================================================== */

static void rgraph_free(pll_rnode_t * root, void (*cb_destroy)(void *))
{
  size_t cap = 64, top = 0;
  pll_rnode_t ** stack;
  if (!root) return;
  stack = (pll_rnode_t **)malloc(cap * sizeof(*stack));
  if (!stack) return;
  stack[top++] = root;
  while (top)
  {
    pll_rnode_t * n = stack[--top];
    if (top + 2 > cap)
    {
      pll_rnode_t ** grown = (pll_rnode_t **)realloc(stack, 2 * cap * sizeof(*stack));
      if (!grown) break;
      stack = grown;
      cap *= 2;
    }
    if (n->left) stack[top++] = n->left;
    if (n->right) stack[top++] = n->right;
    if (n->data && cb_destroy) cb_destroy(n->data);
    free(n->label);
    free(n);
  }
  free(stack);
}

PLL_EXPORT void pll_rtree_graph_destroy(pll_rnode_t * root, void (*cb_destroy)(void *)) { rgraph_free(root, cb_destroy); }

PLL_EXPORT void pll_rtree_destroy(pll_rtree_t * tree, void (*cb_destroy)(void *))
{
  unsigned int i;
  if (!tree) return;
  for (i = 0; i < tree->tip_count + tree->inner_count; ++i)
  {
    pll_rnode_t * n = tree->nodes[i];
    if (n->data && cb_destroy) cb_destroy(n->data);
    free(n->label);
    free(n);
  }
  free(tree->nodes);
  free(tree);
}

/* subtree := '(' subtree ',' subtree ')' [label] [length] | label [length]  (strictly bifurcating) */
static pll_rnode_t * parse_rsubtree(nwk_t * k, unsigned int * tips)
{
  pll_rnode_t * node = (pll_rnode_t *)calloc(1, sizeof(pll_rnode_t));
  int has_len;
  if (!node)
  {
    nwk_fail(k, "out of memory");
    return NULL;
  }
  nwk_skip(k);
  if (*k->p == '(')
  {
    ++k->p;
    ++k->col;
    node->left = parse_rsubtree(k, tips);
    nwk_skip(k);
    if (node->left && *k->p == ',')
    {
      ++k->p;
      ++k->col;
      node->right = parse_rsubtree(k, tips);
      nwk_skip(k);
      if (node->right && *k->p == ')')
      {
        ++k->p;
        ++k->col;
        node->left->parent = node->right->parent = node;
        node->label = nwk_label(k);
        node->length = nwk_length(k, &has_len);
        if (!k->err[0]) return node;
      }
      else if (node->right)
        nwk_fail(k, "expecting ')': rooted trees must be bifurcating");
    }
    else if (node->left)
      nwk_fail(k, "expecting ','");
    rgraph_free(node, NULL);
    return NULL;
  }
  node->label = nwk_label(k);
  if (!node->label)
  {
    nwk_fail(k, "expecting a label or '('");
    free(node);
    return NULL;
  }
  node->length = nwk_length(k, &has_len);
  if (k->err[0])
  {
    rgraph_free(node, NULL);
    return NULL;
  }
  ++*tips;
  return node;
}

/* post-order over a rooted tree without recursion; visit(node, ctx) after both children */
static int rwalk_postorder(pll_rnode_t * root, void (*visit)(pll_rnode_t *, void *), void * ctx)
{
  size_t cap = 64, top = 0;
  struct frame
  {
    pll_rnode_t * node;
    int state;
  } * stack = (struct frame *)malloc(cap * sizeof(*stack));
  if (!stack) return 0;
  stack[top].node = root;
  stack[top].state = 0;
  ++top;
  while (top)
  {
    struct frame * f = &stack[top - 1];
    pll_rnode_t * child = NULL;
    if (f->node->left && f->state == 0) child = f->node->left;
    if (f->node->left && f->state == 1) child = f->node->right;
    if (child)
    {
      ++f->state;
      if (top == cap)
      {
        struct frame * grown = (struct frame *)realloc(stack, 2 * cap * sizeof(*stack));
        if (!grown)
        {
          free(stack);
          return 0;
        }
        stack = grown;
        cap *= 2;
      }
      stack[top].node = child;
      stack[top].state = 0;
      ++top;
      continue;
    }
    visit(f->node, ctx);
    --top;
  }
  free(stack);
  return 1;
}

typedef struct
{
  unsigned int tip_clv, inner_clv, inner_node;
  int inner_scaler;
  pll_rnode_t ** nodes;
  unsigned int tips, inner, inner_base;
} rctx_t;

static void rvisit_assign(pll_rnode_t * n, void * vctx)
{
  rctx_t * c = (rctx_t *)vctx;
  if (!n->left)
  {
    n->node_index = n->clv_index = n->pmatrix_index = c->tip_clv++;
    n->scaler_index = PLL_SCALE_BUFFER_NONE;
    return;
  }
  n->node_index = c->inner_node++;
  n->clv_index = n->pmatrix_index = c->inner_clv++;
  n->scaler_index = c->inner_scaler++;
}

/* src/parse_rtree.y:205-231 */
PLL_EXPORT void pll_rtree_reset_template_indices(pll_rnode_t * root, unsigned int tip_count)
{
  rctx_t c;
  memset(&c, 0, sizeof(c));
  c.inner_clv = c.inner_node = tip_count;
  rwalk_postorder(root, rvisit_assign, &c);
  root->pmatrix_index = 0; /* never used: the root has no branch */
}

static void rvisit_collect(pll_rnode_t * n, void * vctx)
{
  rctx_t * c = (rctx_t *)vctx;
  if (n->left)
    c->nodes[c->inner_base + c->inner++] = n;
  else
    c->nodes[c->tips++] = n;
}

static void rvisit_count(pll_rnode_t * n, void * vctx)
{
  if (!n->left) ++((rctx_t *)vctx)->tips;
}

PLL_EXPORT pll_rtree_t * pll_rtree_wraptree(pll_rnode_t * root, unsigned int tip_count)
{
  pll_rtree_t * tree;
  rctx_t c;
  if (tip_count < 2 && tip_count != 0)
  {
    pll_errno = PLL_ERROR_PARAM_INVALID;
    snprintf(pll_errmsg, 200, "Invalid tip_count value (%u).", tip_count);
    return NULL;
  }
  memset(&c, 0, sizeof(c));
  if (tip_count == 0)
  {
    rwalk_postorder(root, rvisit_count, &c);
    tip_count = c.tips;
    if (tip_count < 2)
    {
      tree_error(PLL_ERROR_PARAM_INVALID, "Input tree contains no inner nodes.");
      return NULL;
    }
  }
  tree = (pll_rtree_t *)malloc(sizeof(pll_rtree_t));
  if (tree) tree->nodes = (pll_rnode_t **)malloc((2 * (size_t)tip_count - 1) * sizeof(pll_rnode_t *));
  if (!tree || !tree->nodes)
  {
    free(tree);
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  memset(&c, 0, sizeof(c));
  c.nodes = tree->nodes;
  c.inner_base = tip_count;
  rwalk_postorder(root, rvisit_collect, &c);
  tree->tip_count = tip_count;
  tree->inner_count = tip_count - 1;
  tree->edge_count = 2 * tip_count - 2;
  tree->root = root;
  return tree;
}

PLL_EXPORT pll_rtree_t * pll_rtree_parse_newick_string(const char * s)
{
  nwk_t k;
  pll_rnode_t * root;
  unsigned int tips = 0;
  memset(&k, 0, sizeof(k));
  k.p = s;
  k.line = 1;
  root = parse_rsubtree(&k, &tips);
  if (root)
  {
    nwk_skip(&k);
    if (*k.p != ';') nwk_fail(&k, "expecting ';'");
    if (!root->left) nwk_fail(&k, "a tree needs at least two tips");
  }
  if (!root || k.err[0])
  {
    rgraph_free(root, NULL);
    tree_error(PLL_ERROR_NEWICK_SYNTAX, k.err[0] ? k.err : "syntax error");
    return NULL;
  }
  pll_rtree_reset_template_indices(root, tips);
  return pll_rtree_wraptree(root, tips);
}

PLL_EXPORT pll_rtree_t * pll_rtree_parse_newick(const char * filename)
{
  char * text = read_whole_file(filename);
  pll_rtree_t * tree;
  if (!text) return NULL;
  tree = pll_rtree_parse_newick_string(text);
  free(text);
  return tree;
}

/* src/rtree.c:306-387: tips are reported when the callback accepts them, an inner node's subtree is entered
 * only when the callback accepts the node */
PLL_EXPORT int pll_rtree_traverse(pll_rnode_t * root, int traversal, int (*cbtrav)(pll_rnode_t *),
                                  pll_rnode_t ** outbuffer, unsigned int * trav_size)
{
  size_t cap = 64, top = 0;
  struct frame
  {
    pll_rnode_t * node;
    int state;
  } * stack;
  *trav_size = 0;
  if (!root->left) return PLL_FAILURE;
  if (traversal != PLL_TREE_TRAVERSE_POSTORDER && traversal != PLL_TREE_TRAVERSE_PREORDER)
  {
    tree_error(PLL_ERROR_PARAM_INVALID, "Invalid traversal value.");
    return PLL_FAILURE;
  }
  stack = (struct frame *)malloc(cap * sizeof(*stack));
  if (!stack)
  {
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return PLL_FAILURE;
  }
  stack[top].node = root;
  stack[top].state = -1;
  ++top;
  while (top)
  {
    struct frame * f = &stack[top - 1];
    pll_rnode_t * n = f->node;
    if (f->state < 0)
    {
      if (!cbtrav(n))
      {
        --top;
        continue;
      }
      if (!n->left)
      {
        outbuffer[(*trav_size)++] = n;
        --top;
        continue;
      }
      if (traversal == PLL_TREE_TRAVERSE_PREORDER) outbuffer[(*trav_size)++] = n;
      f->state = 0;
    }
    if (f->state < 2)
    {
      pll_rnode_t * child = f->state == 0 ? n->left : n->right;
      ++f->state;
      if (top == cap)
      {
        struct frame * grown = (struct frame *)realloc(stack, 2 * cap * sizeof(*stack));
        if (!grown)
        {
          free(stack);
          tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
          return PLL_FAILURE;
        }
        stack = grown;
        cap *= 2;
      }
      stack[top].node = child;
      stack[top].state = -1;
      ++top;
      continue;
    }
    if (traversal == PLL_TREE_TRAVERSE_POSTORDER) outbuffer[(*trav_size)++] = n;
    --top;
  }
  free(stack);
  return PLL_SUCCESS;
}

/* src/rtree.c:262-304 */
PLL_EXPORT void pll_rtree_create_operations(pll_rnode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                            double * branches, unsigned int * pmatrix_indices, pll_operation_t * ops,
                                            unsigned int * matrix_count, unsigned int * ops_count)
{
  unsigned int i, n_ops = 0, n_mat = 0;
  for (i = 0; i < trav_buffer_size; ++i)
  {
    const pll_rnode_t * node = trav_buffer[i];
    if (i + 1 < trav_buffer_size) /* the last node is the root: no branch */
    {
      branches[n_mat] = node->length;
      pmatrix_indices[n_mat] = node->pmatrix_index;
      ++n_mat;
    }
    if (node->left)
    {
      pll_operation_t * op = ops + n_ops++;
      op->parent_clv_index = node->clv_index;
      op->parent_scaler_index = node->scaler_index;
      op->child1_clv_index = node->left->clv_index;
      op->child1_scaler_index = node->left->scaler_index;
      op->child1_matrix_index = node->left->pmatrix_index;
      op->child2_clv_index = node->right->clv_index;
      op->child2_scaler_index = node->right->scaler_index;
      op->child2_matrix_index = node->right->pmatrix_index;
    }
  }
  *ops_count = n_ops;
  *matrix_count = n_mat;
}

static void rexport(sbuf_t * b, const pll_rnode_t * root, char * (*cb)(const pll_rnode_t *))
{
  size_t cap = 64, top = 0;
  struct frame
  {
    const pll_rnode_t * node;
    int state;
  } * stack = (struct frame *)malloc(cap * sizeof(*stack));
  if (!stack)
  {
    b->failed = 1;
    return;
  }
  stack[top].node = root;
  stack[top].state = 0;
  ++top;
  while (top && !b->failed)
  {
    struct frame * f = &stack[top - 1];
    const pll_rnode_t * n = f->node;
    if (n->left && f->state < 2)
    {
      sb_puts(b, f->state == 0 ? "(" : ",");
      if (top == cap)
      {
        struct frame * grown = (struct frame *)realloc(stack, 2 * cap * sizeof(*stack));
        if (!grown)
        {
          b->failed = 1;
          break;
        }
        stack = grown;
        cap *= 2;
        f = &stack[top - 1];
      }
      stack[top].node = f->state == 0 ? n->left : n->right;
      stack[top].state = 0;
      ++f->state;
      ++top;
      continue;
    }
    if (n->left) sb_puts(b, ")");
    if (n != root)
    {
      if (cb)
      {
        char * s = cb(n);
        sb_puts(b, s ? s : "");
        free(s);
      }
      else
        sb_label_length(b, n->label, n->length);
    }
    --top;
  }
  free(stack);
}

/* src/rtree.c:150-258.  The terminating ';' is written only by the default serialiser of a tree with an inner
 * root, as in the reference. */
PLL_EXPORT char * pll_rtree_export_newick(const pll_rnode_t * root, char * (*cb_serialize)(const pll_rnode_t *))
{
  sbuf_t b;
  memset(&b, 0, sizeof(b));
  if (!root) return NULL;
  if (root->left && root->right) rexport(&b, root, cb_serialize);
  if (cb_serialize)
  {
    char * s = cb_serialize(root);
    sb_puts(&b, s ? s : "");
    free(s);
  }
  else
  {
    sb_label_length(&b, root->label, root->length);
    if (root->left && root->right) sb_puts(&b, ";");
  }
  if (b.failed)
  {
    free(b.s);
    tree_error(PLL_ERROR_MEM_ALLOC, "memory allocation during newick export failed");
    return NULL;
  }
  return b.s;
}

/* ---- rooted -> unrooted (src/utree.c:635-760) ------------------------------------------------ *
 * The first root child that has descendants becomes the virtual root; the other child hangs off   *
 * it through one edge whose length is the sum of the two root branches.  Indices are left zero:   *
 * callers follow up with pll_utree_reset_template_indices (as the reference's examples do).       */
static char * dup_label(const char * s)
{
  char * p;
  if (!s) return NULL;
  p = (char *)malloc(strlen(s) + 1);
  if (p) strcpy(p, s);
  return p;
}

/* converts the subtree below `rnode` into unode structures hanging off `back` */
static pll_unode_t * unroot_subtree(const pll_rnode_t * rnode, pll_unode_t * back)
{
  struct job
  {
    const pll_rnode_t * rnode;
    pll_unode_t * link; /* the node of the parent ring that will point at the converted subtree */
  } * stack;
  size_t cap = 64, top = 0;
  pll_unode_t * result = NULL;
  int failed = 0;
  stack = (struct job *)malloc(cap * sizeof(*stack));
  if (!stack) return NULL;
  stack[top].rnode = rnode;
  stack[top].link = back;
  ++top;
  while (top)
  {
    const struct job j = stack[--top];
    pll_unode_t * u = unode_new();
    if (!u)
    {
      failed = 1;
      break;
    }
    u->back = j.link;
    u->label = dup_label(j.rnode->label);
    u->length = j.link->length;
    if (j.link == back)
      result = u;
    j.link->back = u;
    if (!j.rnode->left) continue;
    u->next = unode_new();
    if (u->next) u->next->next = unode_new();
    if (!u->next || !u->next->next)
    {
      failed = 1;
      break;
    }
    u->next->next->next = u;
    u->next->length = j.rnode->left->length;
    u->next->next->length = j.rnode->right->length;
    if (top + 2 > cap)
    {
      struct job * grown = (struct job *)realloc(stack, 2 * cap * sizeof(*stack));
      if (!grown)
      {
        failed = 1;
        break;
      }
      stack = grown;
      cap *= 2;
    }
    stack[top].rnode = j.rnode->right;
    stack[top].link = u->next->next;
    ++top;
    stack[top].rnode = j.rnode->left;
    stack[top].link = u->next;
    ++top;
  }
  free(stack);
  if (failed)
  {
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL; /* the partial graph is reachable from `back` and released by the caller */
  }
  return result;
}

PLL_EXPORT pll_utree_t * pll_rtree_unroot(pll_rtree_t * tree)
{
  const pll_rnode_t * root = tree->root, * new_root, * other;
  pll_unode_t * uroot;
  if (!root->left->left && !root->right->left)
  {
    tree_error(PLL_ERROR_TREE_CONVERSION, "Tree requires at least three tips to be converted to unrooted");
    return NULL;
  }
  new_root = root->left->left ? root->left : root->right;
  other = root->left->left ? root->right : root->left;
  uroot = unode_new();
  if (uroot) uroot->next = unode_new();
  if (uroot && uroot->next) uroot->next->next = unode_new();
  if (!uroot || !uroot->next || !uroot->next->next)
  {
    if (uroot) free(uroot->next);
    free(uroot);
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  uroot->next->next->next = uroot;
  uroot->length = root->left->length + root->right->length;
  uroot->label = dup_label(new_root->label);
  uroot->next->label = uroot->next->next->label = uroot->label;
  uroot->next->length = new_root->left->length;
  uroot->next->next->length = new_root->right->length;
  if (!unroot_subtree(other, uroot) || !unroot_subtree(new_root->left, uroot->next) ||
      !unroot_subtree(new_root->right, uroot->next->next))
  {
    pll_utree_graph_destroy(uroot, NULL);
    return NULL;
  }
  return pll_utree_wraptree(uroot, 0);
}

/* ======================================================================== *
 *  copies, const iteration, rooted ASCII drawing, doubly-linked lists        *
 *  (src/utree.c:381-392,551-633; src/rtree.c:24-128; src/list.c)             *
 * ======================================================================== */

/* copy of the roundabout `entry` belongs to; returns the copy of `entry`.  Labels are duplicated once per
 * roundabout and shared by its members, as everywhere in this file; `data` pointers are copied as they are. */
static pll_unode_t * copy_roundabout(const pll_unode_t * entry)
{
  pll_unode_t * first = (pll_unode_t *)malloc(sizeof(pll_unode_t)), * tail = first;
  const pll_unode_t * s;
  if (!first) return NULL;
  *first = *entry;
  first->back = NULL;
  first->next = NULL;
  first->label = entry->label ? strdup(entry->label) : NULL;
  for (s = entry->next; s && s != entry; s = s->next)
  {
    pll_unode_t * c = (pll_unode_t *)malloc(sizeof(pll_unode_t));
    if (!c) break;
    *c = *s;
    c->back = NULL;
    c->label = first->label;
    c->next = NULL;
    tail->next = c;
    tail = c;
  }
  if (entry->next) tail->next = first;
  return first;
}

/* copies what hangs behind the members of old's roundabout other than old itself (all members when
 * `all`), attaching the copies to the corresponding members of the new roundabout */
static void copy_below(pll_unode_t * fresh, const pll_unode_t * old, int all)
{
  const pll_unode_t * s = old;
  pll_unode_t * c = fresh;
  do
  {
    if ((all || s != old) && s->back)
    {
      pll_unode_t * child = copy_roundabout(s->back);
      if (child)
      {
        c->back = child;
        child->back = c;
        copy_below(child, s->back, 0);
      }
    }
    s = s->next;
    c = c->next;
  } while (s && s != old);
}

PLL_EXPORT pll_unode_t * pll_utree_graph_clone(const pll_unode_t * root)
{
  pll_unode_t * fresh = copy_roundabout(root);
  if (!fresh)
  {
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  copy_below(fresh, root, 1);
  return fresh;
}

PLL_EXPORT pll_utree_t * pll_utree_clone(const pll_utree_t * tree)
{
  pll_unode_t * root = pll_utree_graph_clone(tree->vroot);
  if (!root) return NULL;
  return tree->binary ? pll_utree_wraptree(root, tree->tip_count)
                      : pll_utree_wraptree_multi(root, tree->tip_count, tree->inner_count);
}

PLL_EXPORT int pll_utree_every_const(const pll_utree_t * tree, int (*cb)(const pll_utree_t *, const pll_unode_t *))
{
  unsigned int i;
  int rc = 1;
  for (i = 0; i < tree->tip_count + tree->inner_count; ++i) rc &= cb(tree, tree->nodes[i]);
  return rc ? PLL_SUCCESS : PLL_FAILURE;
}

static void rshow_info(const pll_rnode_t * n, int options)
{
  if (options & PLL_UTREE_SHOW_LABEL) printf(" %s", n->label ? n->label : "(null)");
  if (options & PLL_UTREE_SHOW_BRANCH_LENGTH) printf(" %f", n->length);
  if (options & PLL_UTREE_SHOW_CLV_INDEX) printf(" %u", n->clv_index);
  if (options & PLL_UTREE_SHOW_SCALER_INDEX) printf(" %d", n->scaler_index);
  if (options & PLL_UTREE_SHOW_PMATRIX_INDEX) printf(" %u", n->pmatrix_index);
  printf("\n");
}

/* bar[d] says whether column d still carries a branch: 1 while the left child's subtree is drawn, 2 from the
 * moment the right child is reached (its own line still shows the bar, everything below it does not) */
static void rshow_subtree(const pll_rnode_t * n, unsigned int depth, int * bar, int options)
{
  unsigned int i;
  if (!n) return;
  for (i = 0; i < depth; ++i) printf(bar[i] ? "|   " : "    ");
  printf("\n");
  for (i = 0; i + 1 < depth; ++i) printf(bar[i] ? "|   " : "    ");
  printf((n->left || n->right) ? "+---+" : "+---");
  rshow_info(n, options);
  if (bar[depth - 1] == 2) bar[depth - 1] = 0;
  bar[depth] = 1;
  rshow_subtree(n->left, depth + 1, bar, options);
  bar[depth] = 2;
  rshow_subtree(n->right, depth + 1, bar, options);
}

static unsigned int rdepth(const pll_rnode_t * n, unsigned int d)
{
  unsigned int a, b;
  if (!n) return d;
  a = rdepth(n->left, d + 1);
  b = rdepth(n->right, d + 1);
  return a > b ? a : b;
}

PLL_EXPORT void pll_rtree_show_ascii(const pll_rnode_t * root, int options)
{
  int * bar = (int *)calloc((size_t)rdepth(root, 0) + 2, sizeof(int));
  if (!bar)
  {
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return;
  }
  bar[0] = bar[1] = 1;
  rshow_info(root, options);
  rshow_subtree(root->left, 1, bar, options);
  rshow_subtree(root->right, 1, bar, options);
  free(bar);
}

/* src/list.c.  A list is addressed through the pointer to its first element.  append adds at the end;
 * prepend puts the new element right after the first one, which is where the reference puts it (it then loses
 * the elements that followed; here they stay linked behind the new one); remove unlinks the first element that
 * carries `data` and keeps the rest of the list (the reference frees the wrong element when the match is not
 * the head and truncates the list: not reproduced). */
static int dlist_add(pll_dlist_t ** dlist, void * data, int at_end)
{
  pll_dlist_t * item = (pll_dlist_t *)malloc(sizeof(pll_dlist_t)), * after;
  if (!item)
  {
    tree_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return PLL_FAILURE;
  }
  item->data = data;
  item->next = item->prev = NULL;
  if (!*dlist)
  {
    *dlist = item;
    return PLL_SUCCESS;
  }
  after = *dlist;
  if (at_end)
    while (after->next) after = after->next;
  item->next = after->next;
  if (item->next) item->next->prev = item;
  after->next = item;
  item->prev = after;
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_dlist_append(pll_dlist_t ** dlist, void * data) { return dlist_add(dlist, data, 1); }

PLL_EXPORT int pll_dlist_prepend(pll_dlist_t ** dlist, void * data) { return dlist_add(dlist, data, 0); }

PLL_EXPORT int pll_dlist_remove(pll_dlist_t ** dlist, void * data)
{
  pll_dlist_t * item = *dlist;
  while (item && item->data != data) item = item->next;
  if (!item) return PLL_FAILURE;
  if (item->next) item->next->prev = item->prev;
  if (item->prev)
    item->prev->next = item->next;
  else
    *dlist = item->next;
  free(item);
  return PLL_SUCCESS;
}
