/*
 * pll_moves.c -- topological rearrangement moves on unrooted trees with rollback: the tree-search side of the
 * likelihood path (a search proposes a move, re-evaluates the affected CLVs with a partial traversal and keeps
 * or rolls back the move).  Host only: pointer surgery on pll_unode_t records.
 *
 * Reference: src/utree_moves.c (pll_utree_nni :63, pll_utree_spr :110, pll_utree_spr_safe :290,
 * pll_utree_rollback :343).  Same results, branch-length / matrix-index reports, error codes and messages.
 *
 *   NNI on the inner edge (p, p->back): the subtree behind p->next changes places with the subtree behind
 *       p->back->next (LEFT) or p->back->next->next (RIGHT); both keep their branch and P-matrix index.
 *   SPR: the ring of p (with the subtree behind p->back hanging on it) is taken out - its two other neighbours
 *       u, v are joined by one branch of the summed length, keeping u's matrix index - and put on the edge
 *       (r, r->back), which is halved; the three changed (length, matrix index) pairs are reported in the
 *       order u-v, r'-ring, r-ring.
 */
#include <stdio.h>
#include <string.h>

#include "pll_b200.h"

static int move_error(int code, const char * msg)
{
  pll_errno = code;
  snprintf(pll_errmsg, 200, "%s", msg);
  return PLL_FAILURE;
}

static void join(pll_unode_t * a, pll_unode_t * b, double length, unsigned int pmatrix_index)
{
  a->back = b;
  b->back = a;
  a->length = b->length = length;
  a->pmatrix_index = b->pmatrix_index = pmatrix_index;
}

/* is `target` one of the records of the subtree hanging behind `start` (start's own ring included)? */
static int subtree_has(const pll_unode_t * start, const pll_unode_t * target)
{
  if (!start) return 0;
  if (start == target) return 1;
  if (!start->next) return 0;
  return start->next == target || subtree_has(start->next->back, target) || start->next->next == target ||
         subtree_has(start->next->next->back, target);
}

static int report_args_ok(const double * branch_lengths, const unsigned int * matrix_indices)
{
  if ((branch_lengths == NULL) != (matrix_indices == NULL))
    return move_error(PLL_ERROR_PARAM_INVALID, "Parameters 4,5 must be both NULL or both set");
  return PLL_SUCCESS;
}

static void report(double * branch_lengths, unsigned int * matrix_indices, int k, double length, unsigned int index)
{
  if (!branch_lengths) return;
  branch_lengths[k] = length;
  matrix_indices[k] = index;
}

PLL_EXPORT int pll_utree_nni(pll_unode_t * p, int type, pll_utree_rb_t * rb)
{
  pll_unode_t * s1, * s2, * old1, * old2;
  if (type != PLL_UTREE_MOVE_NNI_LEFT && type != PLL_UTREE_MOVE_NNI_RIGHT)
    return move_error(PLL_ERROR_NNI_INVALIDMOVE, "Invalid NNI move type");
  if (!p->next || !p->back->next) return move_error(PLL_ERROR_NNI_TERMINALBRANCH, "Specified terminal branch");
  if (rb)
  {
    rb->move_type = PLL_UTREE_MOVE_NNI;
    rb->nni.p = p;
    rb->nni.nni_type = type;
  }
  s1 = p->next;
  s2 = type == PLL_UTREE_MOVE_NNI_LEFT ? p->back->next : p->back->next->next;
  /* the two ring records trade what hangs behind them; the subtrees keep their branches */
  old1 = s1->back;
  old2 = s2->back;
  join(s1, old2, old2->length, old2->pmatrix_index);
  join(s2, old1, old1->length, old1->pmatrix_index);
  return PLL_SUCCESS;
}

static int same_tree_move(const pll_unode_t * p, const pll_unode_t * r)
{
  return r == p || r == p->back || r == p->next || r == p->next->back || r == p->next->next ||
         r == p->next->next->back;
}

PLL_EXPORT int pll_utree_spr(pll_unode_t * p, pll_unode_t * r, pll_utree_rb_t * rb, double * branch_lengths,
                             unsigned int * matrix_indices)
{
  pll_unode_t * u, * v, * q, * q2;
  double half;
  if (!report_args_ok(branch_lengths, matrix_indices)) return PLL_FAILURE;
  if (!p->next) return move_error(PLL_ERROR_SPR_TERMINALBRANCH, "Prune edge must be defined by an inner node");
  if (same_tree_move(p, r)) return move_error(PLL_ERROR_SPR_NOCHANGE, "Proposed move yields the same tree");
  q = p->next;
  q2 = p->next->next;
  u = q->back;
  v = q2->back;
  if (rb)
  {
    rb->move_type = PLL_UTREE_MOVE_SPR;
    rb->spr.p = p;
    rb->spr.r = r;
    rb->spr.rb = r->back;
    rb->spr.r_len = r->length;
    rb->spr.pnb = u;
    rb->spr.pnb_len = q->length;
    rb->spr.pnnb = v;
    rb->spr.pnnb_len = q2->length;
  }
  /* close the gap the ring leaves */
  join(u, v, u->length + v->length, u->pmatrix_index);
  report(branch_lengths, matrix_indices, 0, u->length, u->pmatrix_index);
  q->back = q2->back = NULL;
  /* open the regraft edge around the ring */
  half = r->length / 2;
  join(r->back, q2, half, q2->pmatrix_index);
  report(branch_lengths, matrix_indices, 1, half, q2->pmatrix_index);
  join(r, q, half, r->pmatrix_index);
  report(branch_lengths, matrix_indices, 2, half, r->pmatrix_index);
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_utree_spr_safe(pll_unode_t * p, pll_unode_t * r, pll_utree_rb_t * rb, double * branch_lengths,
                                  unsigned int * matrix_indices)
{
  if (!p) return move_error(PLL_ERROR_PARAM_INVALID, "Node p is set to NULL");
  if (!r) return move_error(PLL_ERROR_PARAM_INVALID, "Node r is set to NULL");
  if (!p->next) return move_error(PLL_ERROR_SPR_TERMINALBRANCH, "Prune edge must be defined by an inner node");
  if (same_tree_move(p, r)) return move_error(PLL_ERROR_SPR_NOCHANGE, "Proposed move yields the same tree");
  if (subtree_has(p->back, r)) return move_error(PLL_ERROR_PARAM_INVALID, "Node r is part of the subtree to be pruned");
  return pll_utree_spr(p, r, rb, branch_lengths, matrix_indices);
}

PLL_EXPORT int pll_utree_rollback(pll_utree_rb_t * rollback, double * branch_lengths, unsigned int * matrix_indices)
{
  if (!rollback) return move_error(PLL_ERROR_PARAM_INVALID, "Provide a rollback");
  if (rollback->move_type == PLL_UTREE_MOVE_NNI) return pll_utree_nni(rollback->nni.p, rollback->nni.nni_type, NULL);
  if (rollback->move_type == PLL_UTREE_MOVE_SPR)
  {
    pll_unode_t * p = rollback->spr.p;
    if (!report_args_ok(branch_lengths, matrix_indices)) return PLL_FAILURE;
    join(rollback->spr.pnb, p->next, rollback->spr.pnb_len, rollback->spr.pnb->pmatrix_index);
    report(branch_lengths, matrix_indices, 0, rollback->spr.pnb_len, rollback->spr.pnb->pmatrix_index);
    join(rollback->spr.pnnb, p->next->next, rollback->spr.pnnb_len, p->next->next->pmatrix_index);
    report(branch_lengths, matrix_indices, 1, rollback->spr.pnnb_len, p->next->next->pmatrix_index);
    join(rollback->spr.r, rollback->spr.rb, rollback->spr.r_len, rollback->spr.r->pmatrix_index);
    report(branch_lengths, matrix_indices, 2, rollback->spr.r_len, rollback->spr.r->pmatrix_index);
    return PLL_SUCCESS;
  }
  return move_error(PLL_ERROR_PARAM_INVALID, "Invalid move type");
}
