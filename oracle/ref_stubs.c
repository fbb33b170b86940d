/* Test infrastructure, NOT product code and NOT reference code.
 *
 * The reference defines pll_utree_wraptree / pll_utree_wraptree_multi in its bison grammar file
 * (src/parse_utree.y), which cannot be generated in this image (no bison/flex).  src/utree.c refers
 * to them only from pll_utree_clone / pll_rtree_unroot, which no test calls; these two symbols exist
 * so that oracle/_ref/libpll_ref.so loads with the reference's own traversal code
 * (pll_utree_traverse, pll_utree_create_operations, pll_rtree_*) inside. */
#include <stdio.h>
#include <stdlib.h>

void * pll_utree_wraptree(void * root, unsigned int tip_count)
{
  (void)root;
  (void)tip_count;
  fprintf(stderr, "oracle/_ref: pll_utree_wraptree is not part of this build (needs bison)\n");
  abort();
}

void * pll_utree_wraptree_multi(void * root, unsigned int tip_count, unsigned int inner_count)
{
  (void)inner_count;
  return pll_utree_wraptree(root, tip_count);
}
