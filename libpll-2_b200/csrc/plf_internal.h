/* plf_internal.h -- shared by the .cu translation units only. */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "plf_backend.h"

struct plf_ws
{
  void * ptr;
  size_t bytes;
};

struct plf_guard_rec
{
  void * user;
  size_t bytes;
};

struct plf_ctx
{
  int guard;               /* $PLL_CUDA_GUARD=1: allocations between guard bands (plf_context.cu) */
  plf_guard_rec * guard_recs;
  size_t guard_count, guard_cap;
  int device;
  int managed;
  int sm_count;
  size_t smem_optin;
  size_t gen20_smem_set;
  size_t lk20_smem_set;
  int dna_occupancy[3][6]; /* resident CTAs per SM of the DNA CLV kernels [kind][log2 rates] */
  int dna_stream_occupancy[2][6];
  int dna_cherry_occupancy[3][3]; /* consumers of virtual cherries [CI, TC, CC][log2 rates] */
  int dna_cherry_items;           /* PLF_CHERRY_ITEMS: 2 (default) or 4 (site, rate) blocks per thread and tile */
  int dna_cherry_stages;          /* PLF_CHERRY_STAGES: 6 (default) or 4 ring stages */
  int dna_level_max_sites;        /* PLF_LEVEL_MAX_SITES: widest alignment whose levels run as one launch each (-1 = read on first use) */
  int dna_flow;                   /* PLF_FLOW=0: narrow alignments one launch per level instead of one per traversal (-1 = read on first use) */
  int dna_flow_max_sites;         /* PLF_FLOW_MAX_SITES: widest alignment that runs as one launch per traversal */
  unsigned long long dna_flow_max_updates; /* PLF_FLOW_MAX_UPDATES: ... and most ops x sites */
  int dna_flow_path_max;          /* PLF_FLOW_PATH_MAX: ops of a path (1 = every parent goes through memory) */
  int dna_flow_occupancy[2][3];   /* [1 or 2 blocks per thread][log2 rates] */
  int dna_cherry_bulk;            /* PLF_CHERRY_BULK=1: tip + cherry / cherry + cherry through the bulk-store kernel instead of the ring kernel */
  int dna_cherry;                 /* PLF_VIRTUAL_CHERRIES=0 writes every tip-tip parent to HBM */
  int dna_tt_bulk_occupancy[4];
  int dna_balanced_occupancy[2][6]; /* [three id arrays, pair list][log2 rates] */
  int dna_stream;          /* -1 = read PLF_DNA_STREAM / PLF_DNA_STAGES on first use */
  int dna_stages;
  size_t aa_smem_set[2];
  int aa_occupancy[2];
  int aa_fast;             /* PLF_AA_FAST=0 forces the generic 20-state kernel */
  int aa_mma;              /* PLF_AA_MMA=0: 20-state ii/ti on the bit-exact DFMA kernels instead of DMMA */
  size_t aam_smem_set[5];  /* [ii, ti, tt, stream ii, stream ti] */
  int aam_occupancy[5];
  int aam_log2r[2];
  size_t aas_smem_set[5];  /* 20-state streaming kernels [II, TI, CI, TC, CC] */
  int aas_occupancy[5];
  int aas_log2r[5];
  int aa_stream;           /* -1 = read PLF_AA_STREAM on first use; 0 keeps contiguous ops on the direct-load DMMA kernel */
  int dna_items;
  int dna_tt_bulk, dna_tt_items, dna_tt_seq, dna_balanced; /* PLF_TT_BULK / PLF_TT_ITEMS / PLF_TT_SEQ / PLF_DNA_BALANCED, read with dna_stream */
  int aa_stages;           /* PLF_AA_STAGES=6: deeper ring in the 20-state streaming kernels (A/B) */
  int aa_warps8, aa_l2pf;  /* PLF_AA_WARPS=8, PLF_AAM_L2PF=0 (A/B switches of the 20-state DMMA kernels) */
  int lka_occupancy[3][4];  /* 20-state log-likelihood kernels [mode][log2 rates] */
  size_t lka_smem_set[3][4];
  int edge_occupancy[4][4]; /* DNA edge kernels [mode][log2 rates] */
  int edge_items;           /* 0 = read PLF_EDGE_ITEMS on first use */
  int edge_fast;            /* PLF_EDGE_FAST=0 forces the generic log-likelihood / sumtable / derivative kernels */
  cudaStream_t stream;
  cudaMemPool_t pool;    /* stream-ordered allocator behind plf_alloc/plf_free (NULL in managed mode) */
  plf_ws ws_ops;      /* op descriptors of the current update_partials call   */
  plf_ws ws_once;     /* op descriptors of internal single ops (materialised cherries) */
  plf_ws ws_flow;     /* k_clv_dna_flow: queue control block + one flag per work item */
  void * ws_flow_zeroed; /* the allocation whose flags have been zeroed */
  size_t ws_flow_zeroed_bytes;
  plf_ws ws_small;    /* matrix indices, branch lengths, expm1 values          */
  plf_ws ws_partial;  /* per-block partial sums of the reductions              */
  plf_ws ws_edge;     /* synthetic op + matrices of the DNA/AA sumtable launches */
  double * d_result;  /* 4 doubles                                             */
  unsigned int * d_ticket; /* arrival counter of the fused (last block) reductions; zero between kernels */
  double * h_result;  /* pinned, 4 doubles                                     */
  /* one CUDA graph per traversal (plf_partials.cu): the last op list and its instantiated graph */
  int graph_mode;          /* -1 = read PLF_GRAPH on first use; 0 = plain launches */
  int graph_valid;
  struct plf_op * graph_ops;
  unsigned int * graph_levels;
  unsigned int graph_nops, graph_nlevels, graph_cap_ops, graph_cap_levels, graph_maxstates;
  struct plf_shape graph_shape;
  const unsigned long long * graph_tipmap;
  cudaGraphExec_t graph_exec;
  unsigned long long graph_launches;
  char err[256];
  char name[128];
};

void plf_set_error(plf_ctx * ctx, const char * fmt, ...);
void * plf_ws_reserve(plf_ctx * ctx, plf_ws * ws, size_t bytes);
void plf_count_launch(void);
void plf_count_launches(unsigned long long n);
void plf_graph_cache_destroy(plf_ctx * ctx);
struct plf_op;
int plf_launch_aa_group(plf_ctx * ctx, const struct plf_op * d_ops, unsigned int nops, unsigned int kind,
                        unsigned int rate_cats, int per_rate, unsigned int max_sites,
                        const unsigned long long * d_tipmap, unsigned int maxstates);
int plf_launch_aa_mma_group(plf_ctx * ctx, const struct plf_op * d_ops, unsigned int nops, unsigned int kind,
                            unsigned int rate_cats, int per_rate, unsigned int max_sites,
                            const unsigned long long * d_tipmap, unsigned int maxstates, int contiguous);
int plf_launch_dna_group(plf_ctx * ctx, const struct plf_op * d_ops, unsigned int nops, unsigned int kind,
                         unsigned int rate_cats, int per_rate, unsigned int max_sites, int contiguous,
                         const unsigned int * d_tile_prefix, unsigned int total_tiles, int pair_lists = 0);
unsigned int plf_dna_balanced_tiles(unsigned int nsites, unsigned int rate_cats);
int plf_launch_dna_level(plf_ctx * ctx, const struct plf_op * d_ops, unsigned int nops, unsigned int rate_cats, int per_rate,
                         unsigned int max_sites);
unsigned int plf_dna_flow_chunks(unsigned int rate_cats, unsigned int max_sites);
int plf_launch_dna_flow(plf_ctx * ctx, const struct plf_flow_op * d_fops, const unsigned int * d_path_start,
                        unsigned int npaths, unsigned int rate_cats, int per_rate, unsigned int max_sites, void * flow);
int plf_aa_virtual_cherries_supported(plf_ctx * ctx, const struct plf_shape * sh, unsigned int maxstates);
int plf_dna_virtual_cherries_supported(plf_ctx * ctx, const struct plf_shape * sh);

/* 4-state fast paths (plf_edge_dna.cu); the lk/derivative ones return -1 when the call is not eligible */
struct plf_lk;
struct plf_deriv;
struct plf_sumtable;
struct plf_shape;
int plf_loglikelihood_dna(plf_ctx * ctx, const struct plf_shape * sh, const struct plf_lk * a, double * dst, double * hdst);
int plf_loglikelihood_aa(plf_ctx * ctx, const struct plf_shape * sh, const struct plf_lk * a, unsigned int maxstates,
                         double * dst, double * hdst);
int plf_derivatives_dna(plf_ctx * ctx, const struct plf_shape * sh, const struct plf_deriv * a, double * dst, double * hdst);
int plf_sumtable_as_clv(plf_ctx * ctx, const struct plf_shape * sh, const struct plf_sumtable * a,
                        const unsigned long long * d_tipmap, unsigned int maxstates);
int plf_finish_reduction(plf_ctx * ctx, int nvals, double * h_out);

#define PLF_CHECK(ctx, call)                                                         \
  do                                                                                 \
  {                                                                                  \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess)                                                           \
    {                                                                                \
      plf_set_error((ctx), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                    __FILE__, __LINE__);                                             \
      return 0;                                                                      \
    }                                                                                \
  } while (0)
