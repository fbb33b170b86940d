/* pll_host_internal.h -- what the other host translation units (pll_parsimony.c) may ask of pll_host.c.
 * Hidden visibility: not part of the exported ABI. */
#ifndef PLL_HOST_INTERNAL_H_
#define PLL_HOST_INTERNAL_H_

#include "pll_b200.h"
#include "plf_backend.h"

/* device views of a partition's tip states (pattern-tip codes, or tip CLVs with their site-repeat
 * identifiers), pattern weights and tip map, all current on the device when the call returns */
typedef struct pll_cuda_tipsource
{
  plf_ctx_t * ctx;      /* the partition's context */
  plf_pars_tips_t tips;
  void * d_ptrs;        /* backing store of the pointer arrays: plf_free(ctx, d_ptrs) when done */
} pll_cuda_tipsource_t;

/* device for objects that are not tied to a partition: pll_cuda_set_device() / $PLL_CUDA_DEVICE / $LOCAL_RANK */
int pll_cuda_internal_pick_device(void);

int pll_cuda_internal_tipsource(const pll_partition_t * partition, pll_cuda_tipsource_t * out);

#endif
