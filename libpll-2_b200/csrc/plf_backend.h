/*
 * plf_backend.h -- thin C-ABI between the C host layer (pll_host.c) and the
 * sm_100a CUDA translation units (plf_*.cu).  Plain pointers and sizes only.
 * Internal: callers of the library use include/pll_b200.h.
 */
#ifndef PLF_BACKEND_H_
#define PLF_BACKEND_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plf_ctx plf_ctx_t; /* device, stream, workspaces */

typedef struct plf_shape
{
  unsigned int states;
  unsigned int states_padded;
  unsigned int rate_cats;
  int per_rate_scalers; /* PLL_ATTRIB_RATE_SCALERS */
} plf_shape_t;

/* One CLV update as the device sees it (pointers already resolved).
 * kind: 0 inner-inner, 1 tip-inner (tip is always "left"), 2 tip-tip.
 * The *_id arrays are NULL unless site repeats compress the node.
 *
 * Virtual cherries (4 states, contiguous CLVs): a tip-tip parent is not written to HBM; the op
 * that consumes it forms the cherry's entry from the two tip codes.
 *   3 CI  left = cherry, right = inner CLV
 *   4 TC  left = pattern tip, right = cherry
 *   5 CC  both children cherries
 *   6 TT_VIRTUAL  the cherry itself: only its two P-matrices are snapshot into
 *                 parent_clv[0 .. 2 * R * states * states_padded) -- a side buffer, not the CLV -- and
 *                 its scaler is zeroed
 * For a cherry child: {left,right}_tip / _tip2 are the codes of its two tips, {left,right}_cm1 / _cm2
 * the snapshots of its two P-matrices, {left,right}_matrix the matrix of the branch above it. */
enum { PLF_OP_II = 0, PLF_OP_TI = 1, PLF_OP_TT = 2, PLF_OP_CI = 3, PLF_OP_TC = 4, PLF_OP_CC = 5,
       PLF_OP_TT_VIRTUAL = 6, PLF_OP_KINDS = 7 };
/* most ops of one kernel launch (blockIdx.y selects the op) */
#define PLF_MAX_RUN_OPS 65535u
typedef struct plf_op
{
  double * parent_clv;
  const double * left_clv;
  const double * right_clv;
  const unsigned char * left_tip;
  const unsigned char * right_tip;
  const double * left_matrix;
  const double * right_matrix;
  unsigned int * parent_scaler;
  const unsigned int * left_scaler;
  const unsigned int * right_scaler;
  const unsigned int * parent_id_site;
  const unsigned int * left_site_id;
  const unsigned int * right_site_id;
  const unsigned int * pair_list; /* site repeats: (left entry, right entry) per parent entry, or NULL */
  const unsigned char * left_tip2;
  const unsigned char * right_tip2;
  const double * left_cm1;
  const double * left_cm2;
  const double * right_cm1;
  const double * right_cm2;
  unsigned int nsites;
  unsigned int kind;
  /* positions, in the level-sorted list this op is launched with, of the ops that write what it reads
   * (left CLV, right CLV, left scaler, right scaler); PLF_DEP_NONE when the value is older than the list.
   * dep[0] == PLF_DEP_ORDERED: the list also carries write-after-read / write-after-write order (buffers
   * recycled within the list), which only the launch levels keep */
  int dep[4];
} plf_op_t;
#define PLF_DEP_NONE (-1)
#define PLF_DEP_ORDERED (-2)

/* Model block in device memory, all doubles, per RATE CATEGORY (indices
 * already resolved through params_indices / freqs_indices):
 *   rates[R] weights[R] pinv[R] freqs[R][sp] evals[R][sp]
 *   evecs[R][st*sp] ievecs[R][st*sp]                                   */
static inline size_t plf_model_doubles(unsigned int R, unsigned int st,
                                       unsigned int sp)
{
  return (size_t)3 * R + (size_t)2 * R * sp + (size_t)2 * R * st * sp;
}

/* arguments of the log-likelihood kernels (edge when pmatrix != NULL) */
typedef struct plf_lk
{
  unsigned int sites;
  const double * clvp;
  const unsigned int * pscaler;
  const unsigned int * p_site_id; /* repeats gather or NULL */
  const double * clvc;            /* NULL for tip child / root */
  const unsigned int * cscaler;
  const unsigned int * c_site_id;
  const unsigned char * tipchars; /* child is a pattern tip */
  const unsigned long long * tipmap;
  unsigned int maxstates;         /* entries of tipmap in use */
  const double * pmatrix;         /* NULL => root log-likelihood */
  const double * model;
  const unsigned int * pattern_weights;
  const int * invariant;          /* NULL when no +I */
  double * persite;               /* device [sites] or NULL */
} plf_lk_t;

typedef struct plf_sumtable
{
  unsigned int sites;
  const double * clvp;            /* "parent" side: uses pi * Vinv */
  const unsigned int * pscaler;
  const unsigned int * p_site_id;
  const double * clvc;            /* "child" side: uses V */
  const unsigned int * cscaler;
  const unsigned int * c_site_id;
  const unsigned char * tipchars; /* parent side is a pattern tip */
  const unsigned long long * tipmap;
  unsigned int maxstates;         /* entries of tipmap in use */
  const double * model;
  double * sumtable;              /* device [sites][R][sp] */
} plf_sumtable_t;

typedef struct plf_deriv
{
  unsigned int sites;
  const double * sumtable;
  const double * model;
  const unsigned int * pattern_weights;
  const int * invariant;
  double branch_length;
} plf_deriv_t;

/* ---- context and memory ------------------------------------------------ */
int plf_device_count(char * err, size_t errlen);
int plf_ctx_create(int device, int managed, plf_ctx_t ** out, char * err,
                   size_t errlen);
void plf_ctx_destroy(plf_ctx_t * ctx);
int plf_ctx_device(const plf_ctx_t * ctx);
void * plf_ctx_stream(const plf_ctx_t * ctx);
const char * plf_last_error(const plf_ctx_t * ctx);
void * plf_alloc(plf_ctx_t * ctx, size_t bytes, int zero);
void plf_free(plf_ctx_t * ctx, void * p);
int plf_pool_reserve(plf_ctx_t * ctx, size_t bytes);
/* guard mode ($PLL_CUDA_GUARD=1): allocations whose guard bands were written to; -1 when the mode is off */
int plf_check_guards(plf_ctx_t * ctx);
int plf_upload(plf_ctx_t * ctx, void * dst, const void * src, size_t bytes);
int plf_download(plf_ctx_t * ctx, void * dst, const void * src, size_t bytes);
int plf_memset0(plf_ctx_t * ctx, void * dst, size_t bytes);
void * plf_pinned_alloc(plf_ctx_t * ctx, size_t bytes);
void plf_pinned_free(plf_ctx_t * ctx, void * p);
int plf_sync(plf_ctx_t * ctx);
unsigned long long plf_kernel_launches(void);
void plf_device_description(const plf_ctx_t * ctx, char * buf, size_t len);

/* ---- the hot path ------------------------------------------------------ */

/* P = I + Vinv diag(expm1(lambda r t / (1-pinv))) V for `count` matrices.
 * h_expd: NULL => expm1 on device; else host-computed expm1 values
 * [count][R][states] (bit-exact parity mode). */
int plf_update_pmatrices(plf_ctx_t * ctx, const plf_shape_t * sh,
                         const double * d_model, double * d_pmatrix_base,
                         const unsigned int * h_matrix_indices,
                         const double * h_branch_lengths, unsigned int count,
                         const double * h_expd);

/* CLV updates: ops grouped into `nlevels` levels (level l = ops
 * [h_level_start[l], h_level_start[l+1]) ); one kernel launch per level */
int plf_update_partials(plf_ctx_t * ctx, const plf_shape_t * sh,
                        const plf_op_t * h_ops, unsigned int nops,
                        const unsigned int * h_level_start,
                        unsigned int nlevels,
                        const unsigned long long * d_tipmap,
                        unsigned int maxstates);

/* the same without touching the one-graph-per-traversal cache (internal single ops) */
int plf_update_partials_once(plf_ctx_t * ctx, const plf_shape_t * sh,
                             const plf_op_t * h_ops, unsigned int nops,
                             const unsigned long long * d_tipmap,
                             unsigned int maxstates);
/* end of the launch run that starts at op i of a level ending at b (see plf_partials.cu) */
/* k_clv_dna_flow (plf_partials_dna.cu): one op of a path, as the kernel stages it in shared memory */
#define PLF_FLOW_PATH_MAX 8
struct plf_flow_op
{
  double * parent_clv;
  unsigned int * parent_scaler;
  const double * clv[2];          /* inner child, NULL when that side is a pattern tip */
  const unsigned char * tip[2];
  const double * matrix[2];
  const unsigned int * scaler[2]; /* child scaler to read from memory (NULL: none, or it travels in registers) */
  int dep[2];                     /* path whose flag says clv[side] / scaler[side] has been written, or PLF_DEP_NONE */
  unsigned int nsites;
  unsigned int flags;
};
/* cuts a level-sorted plain op list (dep[] filled) into paths: out_ops (nops entries) path by path, bottom to top,
 * out_start (nops + 1 entries) the first op of each path.  Returns the number of paths, 0 when the list cannot
 * run as one launch.  Host arithmetic only. */
unsigned int plf_dna_flow_plan(const plf_op_t * h_ops, unsigned int nops, unsigned int path_max,
                               struct plf_flow_op * out_ops, unsigned int * out_start);

unsigned int plf_run_end(const plf_op_t * h_ops, unsigned int i, unsigned int b,
                         unsigned int * max_sites, int * contiguous);
/* 1 when the streaming kernels that consume virtual cherries serve this shape and this many tip codes */
int plf_virtual_cherries_supported(plf_ctx_t * ctx, const plf_shape_t * sh, unsigned int maxstates);

/* results: d_out (device, may be NULL) and/or h_out (host, may be NULL; when
 * given the call synchronises the stream) */
int plf_loglikelihood(plf_ctx_t * ctx, const plf_shape_t * sh,
                      const plf_lk_t * a, double * d_out, double * h_out);
int plf_update_sumtable(plf_ctx_t * ctx, const plf_shape_t * sh,
                        const plf_sumtable_t * a);
int plf_derivatives(plf_ctx_t * ctx, const plf_shape_t * sh,
                    const plf_deriv_t * a, double * d_out2, double * h_out2);

/* Newton-Raphson on one branch in ONE cooperative launch (the loop of
 * examples/newton/newton.c:67-96): h_out4 = {length, d_f, dd_f, iterations} */
int plf_newton_branch(plf_ctx_t * ctx, const plf_shape_t * sh,
                      const plf_deriv_t * a, double t0, double tmin,
                      double tmax, double tolerance, unsigned int max_iters,
                      double * h_out4);

/* invariant-site detection (models.c:651-752): AND over tips of the per-site
 * state masks; out[site] = state index or -1 */
int plf_invariant_sites(plf_ctx_t * ctx, const plf_shape_t * sh,
                        unsigned int sites, unsigned int tips,
                        const unsigned char * const * d_tipchars_or_null,
                        const double * const * d_tipclv_or_null,
                        const unsigned int * const * d_tip_site_id_or_null,
                        const unsigned long long * d_tipmap, int * d_out);

/* site-repeat class identifiers of a parent node (repeats.c:299-382) computed
 * on device; returns the class count through *h_ids (0 = no compression).
 * d_lookup: lookup_size entries, all 0xFFFFFFFF on entry and on exit. */
int plf_repeats_ids(plf_ctx_t * ctx, unsigned int sites,
                    const unsigned int * d_site_id_left, unsigned int ids_left,
                    const unsigned int * d_site_id_right,
                    unsigned int * d_site_id_parent,
                    unsigned int * d_id_site_parent, unsigned int * d_lookup,
                    unsigned int * h_ids);

/* the same for a batch of independent parents (one traversal level): job j
 * uses the lookup entries [lookup_offset, lookup_offset + ids_left*ids_right)
 * of the pool; id_site_parent needs room for min(sites, ids_left*ids_right)
 * entries.  One host synchronisation per batch; class counts to h_ids[]. */
typedef struct plf_rep_job
{
  const unsigned int * site_id_left;
  const unsigned int * site_id_right; /* NULL: the key is the left id itself */
  unsigned int * site_id_parent;
  unsigned int * id_site_parent;
  unsigned int ids_left;
  unsigned int lookup_offset;
} plf_rep_job_t;
size_t plf_repeats_batch_workspace(unsigned int sites, unsigned int njobs);
int plf_repeats_ids_batch(plf_ctx_t * ctx, unsigned int sites,
                          const plf_rep_job_t * h_jobs, unsigned int njobs,
                          unsigned int * d_lookup_pool, unsigned int * h_ids);

/* The identifiers of a whole operation list without a host synchronisation per level, for the default
 * enable rule (pll_default_enable_repeats, src/repeats.c:100-110), which is evaluated on the device from
 * the children's class counts.  Job j numbers node `parent` from the identifiers of `left` and `right`;
 * its lookup keys live in the 64-bit entries [lookup_offset, lookup_offset + lookup_entries) of the pool
 * (d_rank_pool: as many 32-bit entries with the same indices, the class number of each key),
 * where lookup_entries >= ids(left) * ids(right) whenever the rule enables the node (the host passes
 * min(upper bound of the product, lookup_buffer_size)).  d_node_ids[node] holds the class count the
 * reference keeps in pernode_ids (0 = not compressed) for every node on entry and is updated per job;
 * d_raw_ids[first_job + j] receives the number of classes found (0 when the rule said no).  Jobs of one
 * call must be independent of each other (one traversal level, or a part of it).  Nothing is waited for. */
typedef struct plf_rid_job
{
  const unsigned int * site_id_left;
  const unsigned int * site_id_right;
  unsigned int * site_id_parent;
  unsigned int * id_site_parent;
  unsigned long long lookup_offset;
  unsigned int left, right, parent;
  unsigned int lookup_entries;
} plf_rid_job_t;
/* bytes of scratch one call needs for `njobs` jobs over `sites` sites */
size_t plf_repeats_pass_workspace(unsigned int sites, unsigned int njobs);
int plf_repeats_pass(plf_ctx_t * ctx, unsigned int sites, unsigned int lookup_buffer_size,
                     const plf_rid_job_t * d_jobs, unsigned int first_job, unsigned int njobs,
                     unsigned long long * d_lookup_pool, unsigned int * d_rank_pool,
                     const unsigned int * d_tag_base, unsigned int tag_offset,
                     unsigned int * d_node_ids, unsigned int * d_raw_ids, void * d_scratch);
/* the tag of a pass is *d_tag_base + tag_offset; one call per identifier update lowers the base by the number
 * of passes of the update (a launch on the stream: part of the update's graph) */
int plf_repeats_advance_tags(plf_ctx_t * ctx, unsigned int * d_tag_base, unsigned int n);
/* stream capture / replay of a sequence of launches queued through this backend */
int plf_capture_begin(plf_ctx_t * ctx);
int plf_capture_end(plf_ctx_t * ctx, void ** exec_out);
int plf_capture_abort(plf_ctx_t * ctx);
int plf_graph_replay(plf_ctx_t * ctx, void * exec, unsigned long long launches);
void plf_graph_free(void * exec);
/* pair list of a gathering op: out[2n] / out[2n+1] = entry of the left / right child that parent entry n
 * reads (parent_id_site, left_site_id, right_site_id may each be NULL = identity) */
typedef struct plf_pair_job
{
  const unsigned int * parent_id_site;
  const unsigned int * left_site_id;
  const unsigned int * right_site_id;
  unsigned int * out;
  unsigned int entries;
} plf_pair_job_t;
int plf_repeats_pairs(plf_ctx_t * ctx, const plf_pair_job_t * h_jobs, unsigned int njobs);

/* tip CLV from a sequence of state characters: entry n (site id_site[n] when
 * repeats compress the tip) gets bit j of map[seq[site]] replicated over rates
 * (src/pll.c:959-1024) */
int plf_tip_clv_from_states(plf_ctx_t * ctx, const plf_shape_t * sh,
                            double * d_clv, const unsigned char * d_seq,
                            const unsigned long long * d_map,
                            const unsigned int * d_id_site,
                            unsigned int entries);
/* pattern-tip codes from raw characters: d_out[s] = low byte of d_lut[d_seq[s]], *d_first_bad = first site
 * whose lut entry has bit 8 set (0xFFFFFFFF: none).  Both byte buffers carry 16 bytes of slack. */
int plf_tip_map(plf_ctx_t * ctx, const unsigned char * d_seq, const unsigned short * d_lut,
                unsigned int sites, unsigned char * d_out, unsigned int * d_first_bad);
/* keys[s] = charmap[seq[s]] (tip class codes for plf_repeats_ids with
 * d_site_id_right == NULL, where the key is the left identifier itself) */
int plf_tip_keys(plf_ctx_t * ctx, const unsigned char * d_seq,
                 const unsigned char * d_charmap, unsigned int sites,
                 unsigned int * d_keys);
/* marginal ancestral probabilities from the CLV of a virtual root placed on the node
 * (src/likelihood.c:733-760): d_out[site][state] */
int plf_ancestral(plf_ctx_t * ctx, const plf_shape_t * sh, const double * d_clv,
                  const double * d_model, unsigned int sites, double * d_out);
/* site pattern compression (src/compress.c:171-410) on `count` host strings of `n`
 * characters; see plf_compress.cu.  1 ok, 0 CUDA failure, -1 character not in the map */
int plf_compress_patterns(plf_ctx_t * ctx, char ** h_rows, unsigned int count,
                          unsigned int n, const unsigned char * h_charmap,
                          const unsigned char * h_inv_charmap,
                          unsigned int * h_weight, unsigned int * h_site_pattern,
                          unsigned int * compressed, unsigned int * bad_seq,
                          unsigned int * bad_pos);
int plf_copy_d2d(plf_ctx_t * ctx, void * dst, const void * src, size_t bytes);
int plf_fill_u32(plf_ctx_t * ctx, unsigned int * d, unsigned int value,
                 size_t n);
/* like plf_upload but does not wait: `src` must be pageable memory owned by
 * the library (the driver stages it before returning) */
int plf_upload_async(plf_ctx_t * ctx, void * dst, const void * src,
                     size_t bytes);

/* ---- Fitch parsimony on packed bit vectors (src/fast_parsimony.c) -------- */
typedef struct plf_pars plf_pars_t; /* pinned staging + small device lists */

/* where the tip states come from (all device pointers): pattern-tip codes or
 * tip CLVs, optionally compressed by site repeats */
typedef struct plf_pars_tips
{
  unsigned int tips, sites, states, states_padded, rate_cats;
  const unsigned char * const * d_tipchars;    /* [tips] or NULL */
  const double * const * d_tipclv;             /* [tips] when d_tipchars==NULL */
  const unsigned int * const * d_tip_site_id;  /* [tips] (entries may be NULL) or NULL */
  const unsigned long long * d_tipmap;         /* code -> state mask (states != 4) */
  const unsigned int * d_weights;              /* pattern weights [sites] */
} plf_pars_tips_t;

int plf_pars_create(plf_ctx_t * ctx, plf_pars_t ** out);
void plf_pars_destroy(plf_pars_t * ps);
/* informative flags [sites], bit position of every site [sites+1]; totals to
 * the host (fast_parsimony.c:381-413 and :251-254) */
int plf_pars_informative(plf_pars_t * ps, const plf_pars_tips_t * tp,
                         int * d_informative, unsigned int * d_bitpos,
                         unsigned int * h_bitcount, unsigned int * h_const_cost,
                         unsigned int * h_informative_count);
/* tip vectors [tips][states][words] (fast_parsimony.c:268-340) */
int plf_pars_pack(plf_pars_t * ps, const plf_pars_tips_t * tp,
                  const unsigned int * d_bitpos, unsigned int bitcount,
                  unsigned int words, unsigned int * d_vec);
/* h_ops: count x {parent, child1, child2}; one launch for the whole list */
int plf_pars_update(plf_pars_t * ps, unsigned int * d_vec, unsigned int states,
                    unsigned int words, const unsigned int * h_ops,
                    unsigned int count, unsigned int * h_scores);
/* the same list sorted into levels of mutually independent operations (level l =
 * entries [h_level_start[l], h_level_start[l+1])): one launch per level */
int plf_pars_update_levels(plf_pars_t * ps, unsigned int * d_vec,
                           unsigned int states, unsigned int words,
                           const unsigned int * h_ops, unsigned int count,
                           const unsigned int * h_level_start,
                           unsigned int nlevels, unsigned int * h_scores);
/* h_pairs: n x {node1, node2}; one launch for the whole batch */
int plf_pars_edge_scores(plf_pars_t * ps, const unsigned int * d_vec,
                         unsigned int states, unsigned int words,
                         const unsigned int * h_pairs, unsigned int n,
                         unsigned int * h_scores);
/* stepwise addition: mutations of merging pair e plus mutations of the merge
 * against vector `third`, for all n candidate edges in one launch */
int plf_pars_insert_scan(plf_pars_t * ps, const unsigned int * d_vec,
                         unsigned int states, unsigned int words,
                         const unsigned int * h_pairs, unsigned int n,
                         unsigned int third, unsigned int * h_scores);

/* ---- weighted (Sankoff) parsimony, src/parsimony.c: score buffers [site][state] f64 -------------- */
int plf_wpars_tip(plf_ctx_t * ctx, double * d_out, const char * h_seq,
                  const unsigned long long * h_map, unsigned int sites,
                  unsigned int states, double inf);
/* h_ops: count x {parent, child1, child2} score buffer indices into d_sbuf_table */
int plf_wpars_build(plf_ctx_t * ctx, double * const * d_sbuf_table,
                    const unsigned int * h_ops, unsigned int count,
                    unsigned int states, unsigned int sites, const double * d_matrix);
int plf_wpars_site_min(plf_ctx_t * ctx, const double * d_buf, unsigned int states,
                       unsigned int sites, double * h_out);
/* h_recops: count x {node score, node ancestral, parent score, parent ancestral} */
int plf_wpars_reconstruct(plf_ctx_t * ctx, double * const * d_sbuf_table,
                          unsigned int * const * d_anc_table,
                          const unsigned int * h_recops, unsigned int count,
                          unsigned int states, unsigned int sites,
                          const unsigned int * h_revmap,
                          const unsigned long long * h_map);

#ifdef __cplusplus
}
#endif
#endif
