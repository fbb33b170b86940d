/*
 * pll_b200.h -- public C interface of the B200-native likelihood engine.
 *
 * Drop-in boundary for the phylogenetic-likelihood hot path of libpll-2.
 * Every type and entry point below replaces the reference declaration cited
 * next to it (paths relative to the reference tree); struct field order, types
 * and function signatures are ABI and are kept identical so that a caller
 * built against the reference header can link against libpll_b200.so and flip
 * one attribute bit (PLL_ATTRIB_ARCH_CUDA).
 *
 * Pointer fields of a CUDA partition:
 *   - clv[], scale_buffer[], pmatrix[], ttlookup hold DEVICE addresses (HBM).
 *     They are dereferenceable on the host only when the partition was created
 *     with PLL_CUDA_MANAGED=1 in the environment (cudaMallocManaged).  Use
 *     pll_cuda_download_clv()/..._scaler()/..._pmatrix() to read them.
 *   - rates, rate_weights, subst_params, frequencies, prop_invar, invariant,
 *     pattern_weights, eigen*, charmap, tipmap and every field of
 *     pll_repeats_t are HOST arrays exactly as in the reference; the engine
 *     keeps device mirrors and re-uploads them when they change.
 *   - tipchars[] are host arrays too, but the codes are formed on the device:
 *     call pll_cuda_host_tipchars(partition, tip) before reading tipchars[tip].
 */
#ifndef PLL_B200_H_
#define PLL_B200_H_

/* the standard headers the reference pll.h pulls in for its clients (src/pll.h:24-41) */
#include <assert.h>
#include <ctype.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLL_EXPORT __attribute__((visibility("default")))

/* ---- constants (src/pll.h:82-137) ------------------------------------- */

#define PLL_FAILURE 0
#define PLL_SUCCESS 1
#define PLL_FALSE 0
#define PLL_TRUE 1

#define PLL_ALIGNMENT_CPU 8
#define PLL_ALIGNMENT_SSE 16
#define PLL_ALIGNMENT_AVX 32
#define PLL_ALIGNMENT_CUDA 32   /* same as AVX: identical buffer layouts */

#define PLL_ASCII_SIZE 256

/* 2^256 and 2^-256, both exactly representable (src/pll.h:96-97) */
#define PLL_SCALE_FACTOR 0x1p+256
#define PLL_SCALE_THRESHOLD 0x1p-256
#define PLL_SCALE_BUFFER_NONE (-1)
#define PLL_SCALE_RATE_MAXDIFF 4      /* src/pll.h:104 */

#define PLL_MISC_EPSILON 1e-8
#define PLL_EIGEN_MINFREQ 1e-6

/* attribute bits (src/pll.h:114-137) */
#define PLL_ATTRIB_ARCH_CPU 0
#define PLL_ATTRIB_ARCH_SSE (1 << 0)
#define PLL_ATTRIB_ARCH_AVX (1 << 1)
#define PLL_ATTRIB_ARCH_AVX2 (1 << 2)
#define PLL_ATTRIB_ARCH_AVX512 (1 << 3)
#define PLL_ATTRIB_ARCH_MASK 0xF
#define PLL_ATTRIB_PATTERN_TIP (1 << 4)
#define PLL_ATTRIB_AB_LEWIS (1 << 5)
#define PLL_ATTRIB_AB_FELSENSTEIN (2 << 5)
#define PLL_ATTRIB_AB_STAMATAKIS (3 << 5)
#define PLL_ATTRIB_AB_MASK (7 << 5)
#define PLL_ATTRIB_AB_FLAG (1 << 8)
#define PLL_ATTRIB_RATE_SCALERS (1 << 9)
#define PLL_ATTRIB_SITE_REPEATS (1 << 10)
#define PLL_REPEATS_LOOKUP_SIZE 2000000

/* NEW: the CUDA (sm_100a) architecture selector.  Bit 16 is outside the
 * reference's PLL_ATTRIB_MASK ((1<<11)-1) so it cannot collide with flags of
 * newer reference revisions; it counts as an architecture in the
 * one-architecture check of pll_partition_create (src/pll.c:438-443). */
#define PLL_ATTRIB_ARCH_CUDA (1 << 16)
#define PLL_ATTRIB_MASK (((1 << 11) - 1) | PLL_ATTRIB_ARCH_CUDA)

/* error codes used on this path (src/pll.h:156-190) */
#define PLL_ERROR_MEM_ALLOC 112
#define PLL_ERROR_PARAM_INVALID 113
#define PLL_ERROR_TIPDATA_ILLEGALSTATE 114
#define PLL_ERROR_TIPDATA_ILLEGALFUNCTION 115
#define PLL_ERROR_TREE_CONVERSION 116
#define PLL_ERROR_INVAR_INCOMPAT 117
#define PLL_ERROR_INVAR_PROPORTION 118
#define PLL_ERROR_INVAR_PARAMINDEX 119
#define PLL_ERROR_INVAR_NONEFOUND 120
#define PLL_ERROR_AB_INVALIDMETHOD 121
#define PLL_ERROR_AB_NOSUPPORT 122
#define PLL_ERROR_EINVAL 130
#define PLL_ERROR_MSA_EMPTY 131
#define PLL_ERROR_MSA_MAP_INVALID 132
#define PLL_ERROR_TREE_INVALID 133
#define PLL_ERROR_FILE_OPEN 100
#define PLL_ERROR_FILE_SEEK 101
#define PLL_ERROR_FILE_EOF 102
#define PLL_ERROR_FASTA_ILLEGALCHAR 201
#define PLL_ERROR_FASTA_UNPRINTABLECHAR 202
#define PLL_ERROR_FASTA_INVALIDHEADER 203
#define PLL_ERROR_FASTA_NONALIGNED 204
#define PLL_ERROR_NEWICK_SYNTAX 111
/* src/pll.h:163-167 */
#define PLL_ERROR_PHYLIP_SYNTAX 231
#define PLL_ERROR_PHYLIP_LONGSEQ 232
#define PLL_ERROR_PHYLIP_NONALIGNED 233
#define PLL_ERROR_PHYLIP_ILLEGALCHAR 234
#define PLL_ERROR_PHYLIP_UNPRINTABLECHAR 235
/* src/pll.h:180-183 */
#define PLL_ERROR_SPR_TERMINALBRANCH 123
#define PLL_ERROR_SPR_NOCHANGE 124
#define PLL_ERROR_NNI_INVALIDMOVE 125
#define PLL_ERROR_NNI_TERMINALBRANCH 126
/* src/pll.h:184-186 */
#define PLL_ERROR_STEPWISE_STRUCT 127
#define PLL_ERROR_STEPWISE_TIPS 128
#define PLL_ERROR_STEPWISE_UNSUPPORTED 129

/* src/pll.h:194-199 */
#define PLL_UTREE_SHOW_LABEL (1 << 0)
#define PLL_UTREE_SHOW_BRANCH_LENGTH (1 << 1)
#define PLL_UTREE_SHOW_CLV_INDEX (1 << 2)
#define PLL_UTREE_SHOW_SCALER_INDEX (1 << 3)
#define PLL_UTREE_SHOW_PMATRIX_INDEX (1 << 4)
#define PLL_UTREE_SHOW_DATA (1 << 5)

/* src/pll.h:147-148 */
#define PLL_TREE_TRAVERSE_POSTORDER 1
#define PLL_TREE_TRAVERSE_PREORDER 2
/* NEW: CUDA runtime / device failures */
#define PLL_ERROR_CUDA 900
#define PLL_ERROR_CUDA_UNSUPPORTED 901

#define PLL_GAMMA_RATES_MEAN 0
#define PLL_GAMMA_RATES_MEDIAN 1

/* ---- types (src/pll.h:222-335) ---------------------------------------- */

typedef unsigned long long pll_state_t;
typedef int pll_bool_t;

/* src/pll.h:222-239; the CPU feature words are kept for layout, a CUDA
 * partition does not consult them */
typedef struct pll_hardware_s
{
  int init;
  int altivec_present;
  int mmx_present;
  int sse_present;
  int sse2_present;
  int sse3_present;
  int ssse3_present;
  int sse41_present;
  int sse42_present;
  int popcnt_present;
  int avx_present;
  int avx2_present;
} pll_hardware_t;

struct pll_repeats;

/* src/pll.h:241-288 */
typedef struct pll_partition
{
  unsigned int tips;
  unsigned int clv_buffers;
  unsigned int nodes;
  unsigned int states;
  unsigned int sites;
  unsigned int pattern_weight_sum;
  unsigned int rate_matrices;
  unsigned int prob_matrices;
  unsigned int rate_cats;
  unsigned int scale_buffers;
  unsigned int attributes;

  size_t alignment;
  unsigned int states_padded;

  double ** clv;                 /* [node] -> DEVICE [site][rate][state_padded] */
  double ** pmatrix;             /* [matrix] -> DEVICE [rate][row][col_padded]  */
  double * rates;
  double * rate_weights;
  double ** subst_params;
  unsigned int ** scale_buffer;  /* [buffer] -> DEVICE [site] or [site][rate]   */
  double ** frequencies;
  double * prop_invar;
  int * invariant;
  unsigned int * pattern_weights;

  int * eigen_decomp_valid;
  double ** eigenvecs;
  double ** inv_eigenvecs;
  double ** eigenvals;

  unsigned int maxstates;
  unsigned char ** tipchars;     /* [tip] -> HOST [site]; device mirror inside  */
  unsigned char * charmap;
  double * ttlookup;             /* unused by the CUDA path (tables live in smem) */
  pll_state_t * tipmap;

  int asc_bias_alloc;
  int asc_additional_sites;

  struct pll_repeats * repeats;
} pll_partition_t;

/* src/pll.h:290-321 */
typedef struct pll_repeats
{
  unsigned int ** pernode_site_id;
  unsigned int ** pernode_id_site;
  unsigned int * pernode_ids;
  unsigned int * perscale_ids;
  unsigned int * pernode_allocated_clvs;
  unsigned int (*enable_repeats)(struct pll_partition * partition,
                                 unsigned int left_clv,
                                 unsigned int right_clv);
  void (*reallocate_repeats)(struct pll_partition * partition,
                             unsigned int parent,
                             int scaler_index,
                             unsigned int sites_to_alloc);
  unsigned int * lookup_buffer;
  unsigned int * toclean_buffer;
  unsigned int * id_site_buffer;
  double * bclv_buffer;
  unsigned int lookup_buffer_size;
  char * charmap;
} pll_repeats_t;

/* src/pll.h:325-335 */
typedef struct pll_operation
{
  unsigned int parent_clv_index;
  int parent_scaler_index;
  unsigned int child1_clv_index;
  unsigned int child1_matrix_index;
  int child1_scaler_index;
  unsigned int child2_clv_index;
  unsigned int child2_matrix_index;
  int child2_scaler_index;
} pll_operation_t;

/* ---- thread-local status (src/pll.h:553-555, src/pll.c:24-27) ---------- */

PLL_EXPORT extern __thread int pll_errno;
PLL_EXPORT extern __thread char pll_errmsg[200];
PLL_EXPORT extern __thread pll_hardware_t pll_hardware;

/* character -> state-mask tables (src/pll.h:557-560, values src/maps.c:26-180) */
PLL_EXPORT extern const pll_state_t pll_map_bin[256];
PLL_EXPORT extern const pll_state_t pll_map_nt[256];
PLL_EXPORT extern const pll_state_t pll_map_aa[256];

/* ---- partition lifecycle and tips (src/pll.h:638-667, src/pll.c) ------- */

/* src/pll.c:424 */
PLL_EXPORT pll_partition_t * pll_partition_create(unsigned int tips,
                                                  unsigned int clv_buffers,
                                                  unsigned int states,
                                                  unsigned int sites,
                                                  unsigned int rate_matrices,
                                                  unsigned int prob_matrices,
                                                  unsigned int rate_cats,
                                                  unsigned int scale_buffers,
                                                  unsigned int attributes);
/* src/pll.c:870 */
PLL_EXPORT void pll_partition_destroy(pll_partition_t * partition);
/* src/pll.c:1026 */
PLL_EXPORT int pll_set_tip_states(pll_partition_t * partition,
                                  unsigned int tip_index,
                                  const pll_state_t * map,
                                  const char * sequence);
/* src/pll.c:1066 */
PLL_EXPORT int pll_set_tip_clv(pll_partition_t * partition,
                               unsigned int tip_index,
                               const double * clv,
                               int padding);
/* src/pll.h:347-354 */
typedef struct pll_msa_s
{
  int count;
  int length;
  char ** sequence;
  char ** label;
} pll_msa_t;

/* src/compress.c:391,399 (pll.h:1095-1105): site pattern compression, on the device
 * (plf_compress.cu); identical output to the reference: unique columns in ascending order
 * written over `sequence`, their weights returned (caller frees), *length updated */
PLL_EXPORT unsigned int * pll_compress_site_patterns(char ** sequence,
                                                     const pll_state_t * map,
                                                     int count,
                                                     int * length);
PLL_EXPORT unsigned int * pll_compress_site_patterns_msa(pll_msa_t * msa,
                                                         const pll_state_t * map,
                                                         unsigned int * site_pattern_map);

/* src/likelihood.c:762,639 (pll.h:741-760): marginal ancestral state probabilities of a node,
 * ancestral[site][state]; the scratch buffers of the _extbuf variant must be non-NULL as in the
 * reference but are not used (the temporaries live in HBM) */
PLL_EXPORT int pll_compute_node_ancestral(pll_partition_t * partition,
                                          unsigned int node_clv_index,
                                          int node_scaler_index,
                                          unsigned int other_clv_index,
                                          int other_scaler_index,
                                          unsigned int matrix_index,
                                          const unsigned int * freqs_indices,
                                          double * ancestral);
PLL_EXPORT int pll_compute_node_ancestral_extbuf(pll_partition_t * partition,
                                                 unsigned int node_clv_index,
                                                 int node_scaler_index,
                                                 unsigned int other_clv_index,
                                                 int other_scaler_index,
                                                 unsigned int pmatrix_index,
                                                 const unsigned int * freqs_indices,
                                                 double * ancestral,
                                                 double * temp_clv,
                                                 unsigned int * temp_scaler,
                                                 double * ident_pmat);
/* src/pll.c:1131 */
PLL_EXPORT void pll_set_pattern_weights(pll_partition_t * partition,
                                        const unsigned int * pattern_weights);
/* src/pll.c:1145,1192.  Ascertainment bias correction (PLL_ATTRIB_AB_LEWIS / _FELSENSTEIN /
 * _STAMATAKIS on a partition created with PLL_ATTRIB_AB_FLAG or one of the types): the `states`
 * pseudo-sites are appended to every CLV, scaler, tipchars and sumtable buffer and go through the
 * same kernels as the alignment; the correction terms of the log-likelihood
 * (src/likelihood.c:24-120,190-268,342-440) and of the derivatives (src/core_derivatives.c:851-924)
 * are O(states^2 x rates) host arithmetic on a small download of those pseudo-site blocks.
 * Not supported together with PLL_ATTRIB_SITE_REPEATS (pll_errno 901) and not applied by the
 * pll_cuda_*_async entry points. */
PLL_EXPORT int pll_set_asc_bias_type(pll_partition_t * partition,
                                     int asc_bias_type);
PLL_EXPORT void pll_set_asc_state_weights(pll_partition_t * partition,
                                          const unsigned int * state_weights);
/* src/pll.c:1202: pure host helper on host arrays, kept for callers */
PLL_EXPORT void pll_fill_parent_scaler(unsigned int scaler_size,
                                       unsigned int * parent_scaler,
                                       const unsigned int * left_scaler,
                                       const unsigned int * right_scaler);
/* src/pll.c:134,148 */
PLL_EXPORT void * pll_aligned_alloc(size_t size, size_t alignment);
PLL_EXPORT void pll_aligned_free(void * ptr);

/* ---- model parameters and P-matrices (src/pll.h:736-776, src/models.c) -- */

PLL_EXPORT void pll_set_subst_params(pll_partition_t * partition,
                                     unsigned int params_index,
                                     const double * params);          /* models.c:485 */
PLL_EXPORT void pll_set_frequencies(pll_partition_t * partition,
                                    unsigned int params_index,
                                    const double * frequencies);      /* models.c:445 */
PLL_EXPORT void pll_set_category_rates(pll_partition_t * partition,
                                       const double * rates);         /* models.c:470 */
PLL_EXPORT void pll_set_category_weights(pll_partition_t * partition,
                                         const double * rate_weights);/* models.c:476 */
PLL_EXPORT int pll_update_eigen(pll_partition_t * partition,
                                unsigned int params_index);           /* models.c:293 */
PLL_EXPORT int pll_update_prob_matrices(pll_partition_t * partition,
                                        const unsigned int * params_indices,
                                        const unsigned int * matrix_indices,
                                        const double * branch_lengths,
                                        unsigned int count);          /* models.c:412 */
PLL_EXPORT unsigned int pll_count_invariant_sites(pll_partition_t * partition,
                                                  unsigned int * state_inv_count); /* models.c:546 */
PLL_EXPORT int pll_update_invariant_sites(pll_partition_t * partition);           /* models.c:651 */
PLL_EXPORT int pll_update_invariant_sites_proportion(pll_partition_t * partition,
                                                     unsigned int params_index,
                                                     double prop_invar);          /* models.c:495 */

/* ---- CLV updates (src/pll.h:818-828, src/partials.c:237-291) ----------- */

PLL_EXPORT void pll_update_partials(pll_partition_t * partition,
                                    const pll_operation_t * operations,
                                    unsigned int count);
PLL_EXPORT void pll_update_partials_rep(pll_partition_t * partition,
                                        const pll_operation_t * operations,
                                        unsigned int count,
                                        unsigned int update_repeats);

/* ---- log-likelihood (src/pll.h:782-797, src/likelihood.c:122,586) ------ */

PLL_EXPORT double pll_compute_root_loglikelihood(pll_partition_t * partition,
                                                 unsigned int clv_index,
                                                 int scaler_index,
                                                 const unsigned int * freqs_indices,
                                                 double * persite_lnl);
PLL_EXPORT double pll_compute_edge_loglikelihood(pll_partition_t * partition,
                                                 unsigned int parent_clv_index,
                                                 int parent_scaler_index,
                                                 unsigned int child_clv_index,
                                                 int child_scaler_index,
                                                 unsigned int matrix_index,
                                                 const unsigned int * freqs_indices,
                                                 double * persite_lnl);

/* ---- derivatives (src/pll.h:832-849, src/derivatives.c:239,333) -------- *
 * `sumtable` is the caller's host buffer in the reference.  Here it is an
 * opaque handle: the table itself lives in HBM, keyed by this pointer value,
 * and the host bytes are written only when PLL_CUDA_SUMTABLE_MIRROR=1. */
PLL_EXPORT int pll_update_sumtable(pll_partition_t * partition,
                                   unsigned int parent_clv_index,
                                   unsigned int child_clv_index,
                                   int parent_scaler_index,
                                   int child_scaler_index,
                                   const unsigned int * params_indices,
                                   double * sumtable);
PLL_EXPORT int pll_compute_likelihood_derivatives(pll_partition_t * partition,
                                                  int parent_scaler_index,
                                                  int child_scaler_index,
                                                  double branch_length,
                                                  const unsigned int * params_indices,
                                                  const double * sumtable,
                                                  double * d_f,
                                                  double * dd_f);

/* ---- site repeats (src/pll.h:685-734, src/repeats.c) ------------------- */

#define PLL_GET_ID(site_id, site) ((site_id) ? ((site_id)[(site)]) : (site))
#define PLL_GET_SITE(id_site, site) ((id_site) ? ((id_site)[(site)]) : (site))

PLL_EXPORT int pll_repeats_enabled(const pll_partition_t * partition);          /* repeats.c:46 */
PLL_EXPORT void pll_resize_repeats_lookup(pll_partition_t * partition,
                                          unsigned int size);                   /* repeats.c:51 */
PLL_EXPORT unsigned int pll_get_sites_number(const pll_partition_t * partition,
                                             unsigned int clv_index);           /* repeats.c:62 */
PLL_EXPORT unsigned int * pll_get_site_id(const pll_partition_t * partition,
                                          unsigned int clv_index);              /* repeats.c:79 */
PLL_EXPORT unsigned int * pll_get_id_site(const pll_partition_t * partition,
                                          unsigned int clv_index);              /* repeats.c:89 */
PLL_EXPORT unsigned int pll_get_clv_size(const pll_partition_t * partition,
                                         unsigned int clv_index);               /* repeats.c:72 */
PLL_EXPORT unsigned int pll_default_enable_repeats(pll_partition_t * partition,
                                                   unsigned int left_clv,
                                                   unsigned int right_clv);     /* repeats.c:100 */
PLL_EXPORT unsigned int pll_no_enable_repeats(pll_partition_t * partition,
                                              unsigned int left_clv,
                                              unsigned int right_clv);          /* repeats.c:112 */
PLL_EXPORT void pll_default_reallocate_repeats(pll_partition_t * partition,
                                               unsigned int parent,
                                               int scaler_index,
                                               unsigned int sites_to_alloc);    /* repeats.c:256 */
PLL_EXPORT int pll_update_repeats_tips(pll_partition_t * partition,
                                       unsigned int tip_index,
                                       const pll_state_t * map,
                                       const char * sequence);                  /* repeats.c:189 */
PLL_EXPORT void pll_update_repeats(pll_partition_t * partition,
                                   const pll_operation_t * op);                 /* repeats.c:299 */
PLL_EXPORT void pll_disable_bclv(pll_partition_t * partition);                  /* repeats.c:384 */

/* ---- hardware probe (src/pll.h:2694-2698, src/hardware.c) -------------- */

PLL_EXPORT int pll_hardware_probe(void);
PLL_EXPORT void pll_hardware_dump(void);
PLL_EXPORT void pll_hardware_ignore(void);

/* ---- debug printers (src/pll.h:858-866, src/output.c:26,56) ------------ */

PLL_EXPORT void pll_show_pmatrix(const pll_partition_t * partition,
                                 unsigned int index,
                                 unsigned int float_precision);
PLL_EXPORT void pll_show_clv(const pll_partition_t * partition,
                             unsigned int clv_index,
                             int scaler_index,
                             unsigned int float_precision);

/* ======================================================================= *
 *  Additive CUDA surface (nothing above changes meaning)                   *
 * ======================================================================= */

/* number of visible CUDA devices; 0 (and pll_errno set) if the runtime fails */
PLL_EXPORT int pll_cuda_device_count(void);
/* device used by partitions created afterwards on this thread.  Default:
 * $PLL_CUDA_DEVICE, else $LOCAL_RANK, else 0. */
PLL_EXPORT int pll_cuda_set_device(int device);
PLL_EXPORT int pll_cuda_get_device(const pll_partition_t * partition);
/* the CUDA stream (a cudaStream_t) every kernel and copy of this partition is
 * queued on: record events on it to time the device work, or order a
 * collective after it */
PLL_EXPORT void * pll_cuda_get_stream(const pll_partition_t * partition);
/* block until all work queued on the partition's stream has finished */
PLL_EXPORT int pll_cuda_synchronize(const pll_partition_t * partition);

/* pattern_weights / tipmap are host arrays with device mirrors refreshed by
 * their setters; after writing to those struct fields directly call this so the
 * next evaluation re-uploads them (small model arrays are compared byte-wise
 * on every call and need no notification) */
PLL_EXPORT int pll_cuda_invalidate_host_arrays(pll_partition_t * partition);

/* explicit device->host reads of device-resident buffers; sizes in elements
 * are those of the reference buffers (pll_get_clv_size() doubles, etc.) */
PLL_EXPORT int pll_cuda_download_clv(const pll_partition_t * partition,
                                     unsigned int clv_index, double * host_out);
PLL_EXPORT int pll_cuda_download_scaler(const pll_partition_t * partition,
                                        unsigned int scaler_index,
                                        unsigned int * host_out);
PLL_EXPORT int pll_cuda_download_pmatrix(const pll_partition_t * partition,
                                         unsigned int matrix_index,
                                         double * host_out);
/* host->device write of a P-matrix block (parity mode: feed the engine
 * matrices produced elsewhere) */
PLL_EXPORT int pll_cuda_upload_pmatrix(pll_partition_t * partition,
                                       unsigned int matrix_index,
                                       const double * host_in);
PLL_EXPORT int pll_cuda_download_sumtable(const pll_partition_t * partition,
                                          const double * sumtable_handle,
                                          double * host_out);
/* Virtual cherries (4 states, PLL_ATTRIB_PATTERN_TIP, 1/2/4 rate categories): the CLV
 * of a node whose two children are pattern tips (src/partials.c:100-128 -> pll_core_update_partial_tt) is
 * not written to HBM by pll_update_partials; the operation that consumes it forms the same values, to the
 * bit, from the two tip codes.  Every entry point of this library that reads such a CLV by index
 * (edge/root log-likelihood, sumtable, ancestral states, pll_cuda_download_clv, pll_show_clv) writes it
 * first; a client that passes partition->clv[i] to its own kernels calls pll_cuda_materialize_clv(i).
 * pll_cuda_virtual_cherries(): 1 when the partition works this way ($PLF_VIRTUAL_CHERRIES=0 turns it off,
 * $PLF_VIRTUAL_CHERRY_MIN_SITES sets the narrowest alignment it applies to; default 20000 sites for 20 states,
 * for 4 states one site more than a full traversal may have to run as ONE launch with all parents written:
 * min(65536, 6500000 / (tips - 2)) sites, $PLF_FLOW_MAX_SITES / $PLF_FLOW_MAX_UPDATES; such alignments are bound
 * by launch and dependency latency, not bytes).
 * pll_cuda_virtual_clvs(p, i): is node i virtual right now (i < nodes), or how many nodes are (i >= nodes). */
PLL_EXPORT int pll_cuda_virtual_cherries(const pll_partition_t * partition);
PLL_EXPORT unsigned int pll_cuda_virtual_clvs(const pll_partition_t * partition,
                                              unsigned int clv_index);
PLL_EXPORT int pll_cuda_materialize_clv(pll_partition_t * partition,
                                        unsigned int clv_index);
/* Site-repeat identifiers are a pure function of the children's identifiers: pll_update_partials renumbers a
 * parent only when a child's identifiers changed since it was last numbered from the same two children
 * ($PLL_CUDA_REPEATS_MEMO=0: always).  This call forgets that history (measurements). */
PLL_EXPORT int pll_cuda_invalidate_repeat_identifiers(pll_partition_t * partition);
/* Guard mode: with PLL_CUDA_GUARD=1 in the environment at pll_partition_create, every device buffer of the
 * partition is allocated between two 256-byte guard bands.  pll_cuda_check_guards() returns the number of
 * buffers that had a band written to (0 = clean, -1 = not in guard mode): the out-of-bounds-write check of
 * the test suite on pools where compute-sanitizer is not available.  pll_cuda_debug_overrun() writes one byte
 * just behind a CLV buffer so that a test can see the check fire. */
PLL_EXPORT int pll_cuda_check_guards(const pll_partition_t * partition);
PLL_EXPORT int pll_cuda_debug_overrun(pll_partition_t * partition, unsigned int clv_index);
/* inspection: launches of one traversal level for ops of the given kinds (sorted by kind); a launch serves at
 * most 65535 ops (gridDim.y).  Host arithmetic only. */
/* NEW (inspection, no device needed).  How a narrow 4-state traversal would run as ONE launch (k_clv_dna_flow,
 * DESIGN.md section 4): the list is cut into paths whose parents stay in registers.  path_of_op[i] = position
 * in the queue of the path ops[i] belongs to; carried_child_of_op[i] = 1 / 2 when child1 / child2 of ops[i]
 * arrives in registers.  Returns the number of paths, 0 when the list keeps one launch per level (a buffer is
 * recycled within the list).  `tips`: CLV indices below it are pattern tips. */
PLL_EXPORT unsigned int pll_cuda_schedule_paths(const pll_operation_t * operations, unsigned int count,
                                                unsigned int tips, unsigned int path_max,
                                                unsigned int * path_of_op, int * carried_child_of_op);
PLL_EXPORT unsigned int pll_cuda_count_launch_runs(const unsigned int * kinds, unsigned int count,
                                                   unsigned int * largest_run);
/* The one exchange of a site-sharded evaluation (one process per GPU of one node, each owning a contiguous
 * site slice; the reference's clients use MPI_Allreduce here) as a single-block kernel over peer memory:
 * every rank stores its values into every peer's buffer over NVLink, waits for all slots of its own buffer and
 * adds them in rank order, so all ranks hold the same sum bit for bit.  Set-up: each rank creates its group
 * and hands the 64-byte handle it gets to all other ranks (any transport: an all-gather at start-up), then
 * connects.  pll_cuda_peer_allreduce(group, pll_cuda_get_stream(partition), dev_values, n <= 4) is queued on
 * the stream the asynchronous entry points above left {logL, d_f, dd_f} on; every rank makes the same calls.
 * A rank that never arrives makes the others give up after 20 s (values become NaN,
 * pll_cuda_peer_group_check() returns 0) instead of hanging the device. */
typedef struct pll_cuda_peer_group pll_cuda_peer_group_t;
PLL_EXPORT pll_cuda_peer_group_t * pll_cuda_peer_group_create(int device, unsigned int rank, unsigned int world,
                                                              void * handle_out_64_bytes);
PLL_EXPORT int pll_cuda_peer_group_connect(pll_cuda_peer_group_t * group, const void * handles_world_x_64_bytes);
PLL_EXPORT int pll_cuda_peer_allreduce(pll_cuda_peer_group_t * group, void * stream, double * dev_values,
                                       unsigned int count);
PLL_EXPORT int pll_cuda_peer_group_check(pll_cuda_peer_group_t * group);
PLL_EXPORT void pll_cuda_peer_group_destroy(pll_cuda_peer_group_t * group);

/* tipchars[] of a PLL_ATTRIB_PATTERN_TIP partition are formed on the device from the raw characters
 * (src/pll.c:875-957); the host copy partition->tipchars[i] is written by this call, not by
 * pll_set_tip_states ($PLL_CUDA_TIPCHARS_MIRROR=1 writes it there too, as the reference does). */
PLL_EXPORT const unsigned char * pll_cuda_host_tipchars(pll_partition_t * partition, unsigned int tip_index);
/* number of elements of a scale buffer as currently allocated */
PLL_EXPORT unsigned int pll_cuda_scaler_size(const pll_partition_t * partition,
                                             unsigned int scaler_index);

/* Asynchronous variants for multi-GPU site sharding: results (logL, or
 * {d_f, dd_f}) are left in DEVICE memory at `dev_out` on the partition's
 * stream, so that a collective (NCCL all-reduce) can consume them without a
 * host round trip.  `dev_out` must be a device pointer on the partition's
 * device.  pll_cuda_synchronize() orders the stream with the host. */
PLL_EXPORT int pll_cuda_edge_loglikelihood_async(pll_partition_t * partition,
                                                 unsigned int parent_clv_index,
                                                 int parent_scaler_index,
                                                 unsigned int child_clv_index,
                                                 int child_scaler_index,
                                                 unsigned int matrix_index,
                                                 const unsigned int * freqs_indices,
                                                 double * dev_out);
PLL_EXPORT int pll_cuda_root_loglikelihood_async(pll_partition_t * partition,
                                                 unsigned int clv_index,
                                                 int scaler_index,
                                                 const unsigned int * freqs_indices,
                                                 double * dev_out);
PLL_EXPORT int pll_cuda_likelihood_derivatives_async(pll_partition_t * partition,
                                                     int parent_scaler_index,
                                                     int child_scaler_index,
                                                     double branch_length,
                                                     const unsigned int * params_indices,
                                                     const double * sumtable,
                                                     double * dev_out2);
/* NEW (additive): the whole Newton-Raphson loop on one branch (examples/newton/newton.c:67-96) in ONE
 * cooperative launch and one host synchronisation: evaluate d_f, dd_f at the current length, stop when
 * |d_f| < tolerance, else length -= d_f / dd_f clamped to [min_length, max_length]; at most max_iters
 * evaluations.  *d_f, *dd_f are the derivatives at the last evaluated length. */
PLL_EXPORT int pll_cuda_newton_branch(pll_partition_t * partition, int parent_scaler_index,
                                      int child_scaler_index, double initial_length, double min_length,
                                      double max_length, double tolerance, unsigned int max_iters,
                                      const unsigned int * params_indices, const double * sumtable,
                                      double * length, double * d_f, double * dd_f, unsigned int * iterations);

/* Level schedule of an operation list (pure host logic, no device needed):
 * writes for each op the launch level it is batched into (ops of one level are
 * mutually independent and run in one kernel launch); returns the number of
 * levels, or -1 on error.  RAW, WAR and WAW hazards on CLV and scaler indices
 * are honoured (reference semantics: strictly sequential, src/partials.c:253). */
PLL_EXPORT int pll_cuda_schedule_levels(const pll_operation_t * operations,
                                        unsigned int count,
                                        unsigned int * level_of_op);

/* counters for benchmarking: kernels launched by this library on this thread */
PLL_EXPORT unsigned long long pll_cuda_kernel_launches(void);

/* host eigendecomposition used by pll_update_eigen, exposed for tests:
 * same inputs/outputs as the reference routine (models.c:293-410) on plain
 * arrays; returns PLL_SUCCESS/PLL_FAILURE */
PLL_EXPORT int pll_cuda_host_eigen(unsigned int states, unsigned int states_padded,
                                   const double * subst_params,
                                   const double * freqs,
                                   double * eigenvecs, double * inv_eigenvecs,
                                   double * eigenvals);

/* src/gamma.c:220 (pll.h:781-785): discrete Gamma category rates (pll_gamma.c; host only) */
PLL_EXPORT int pll_compute_gamma_cats(double alpha, unsigned int categories, double * output_rates,
                                      int rates_mode);

/* ---- FASTA reader (pll_fasta.c; host only) ------------------------------------------------- */

#define PLL_LINEALLOC 2048

/* src/pll.h:358-370: same layout */
typedef struct pll_fasta
{
  FILE * fp;
  char line[PLL_LINEALLOC];
  const unsigned int * chrstatus;
  long no;
  long filesize;
  long lineno;
  long stripped_count;
  long stripped[256];
} pll_fasta_t;

/* src/maps.c:207,242 */
PLL_EXPORT extern const unsigned int pll_map_fasta[256];
PLL_EXPORT extern const unsigned int pll_map_generic[256];

/* src/fasta.c:40-417 (pll.h:864-887) */
PLL_EXPORT pll_fasta_t * pll_fasta_open(const char * filename, const unsigned int * map);
PLL_EXPORT int pll_fasta_getnext(pll_fasta_t * fd, char ** head, long * head_len, char ** seq,
                                 long * seq_len, long * seqno);
PLL_EXPORT void pll_fasta_close(pll_fasta_t * fd);
PLL_EXPORT long pll_fasta_getfilesize(const pll_fasta_t * fd);
PLL_EXPORT long pll_fasta_getfilepos(pll_fasta_t * fd);
PLL_EXPORT int pll_fasta_rewind(pll_fasta_t * fd);
PLL_EXPORT pll_msa_t * pll_fasta_load(const char * fname);
PLL_EXPORT void pll_msa_destroy(pll_msa_t * msa);

/* PHYLIP reader, src/phylip.c (handle: src/pll.h:371-384; prototypes pll.h:986-1001); pll_phylip.c */
typedef struct pll_phylip_s
{
  FILE * fp;
  char * line;
  size_t line_size;
  size_t line_maxsize;
  char buffer[PLL_LINEALLOC];
  const unsigned int * chrstatus;
  long no;
  long filesize;
  long lineno;
  long stripped_count;
  long stripped[256];
} pll_phylip_t;
PLL_EXPORT extern const unsigned int pll_map_phylip[256];
PLL_EXPORT pll_phylip_t * pll_phylip_open(const char * filename, const unsigned int * map);
PLL_EXPORT int pll_phylip_rewind(pll_phylip_t * fd);
PLL_EXPORT void pll_phylip_close(pll_phylip_t * fd);
PLL_EXPORT pll_msa_t * pll_phylip_parse_interleaved(pll_phylip_t * fd);
PLL_EXPORT pll_msa_t * pll_phylip_parse_sequential(pll_phylip_t * fd);
PLL_EXPORT pll_msa_t * pll_phylip_load(const char * fname, pll_bool_t interleaved);

/* ---- tree structures and operation-list producers (pll_tree.c; host only) ------------------ */

/* src/pll.h:388-438: same layouts */
typedef struct pll_unode_s
{
  char * label;
  double length;
  unsigned int node_index;
  unsigned int clv_index;
  int scaler_index;
  unsigned int pmatrix_index;
  struct pll_unode_s * next;
  struct pll_unode_s * back;
  void * data;
} pll_unode_t;

typedef struct pll_utree_s
{
  unsigned int tip_count;
  unsigned int inner_count;
  unsigned int edge_count;
  int binary;
  pll_unode_t ** nodes;
  pll_unode_t * vroot;
} pll_utree_t;

typedef struct pll_rnode_s
{
  char * label;
  double length;
  unsigned int node_index;
  unsigned int clv_index;
  int scaler_index;
  unsigned int pmatrix_index;
  struct pll_rnode_s * left;
  struct pll_rnode_s * right;
  struct pll_rnode_s * parent;
  void * data;
} pll_rnode_t;

typedef struct pll_rtree_s
{
  unsigned int tip_count;
  unsigned int inner_count;
  unsigned int edge_count;
  pll_rnode_t ** nodes;
  pll_rnode_t * root;
} pll_rtree_t;

/* src/parse_utree.y (pll.h:907-937): hand-written recursive-descent reader of the same grammar */
PLL_EXPORT pll_utree_t * pll_utree_parse_newick(const char * filename);
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_rooted(const char * filename);
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_unroot(const char * filename);
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_string(const char * s);
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_string_rooted(const char * s);
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_string_unroot(const char * s);
PLL_EXPORT pll_unode_t * pll_utree_unroot_inplace(pll_unode_t * root);
PLL_EXPORT void pll_utree_destroy(pll_utree_t * tree, void (*cb_destroy)(void *));
PLL_EXPORT void pll_utree_reset_template_indices(pll_unode_t * node, unsigned int tip_count);
PLL_EXPORT void pll_utree_graph_destroy(pll_unode_t * root, void (*cb_destroy)(void *));
PLL_EXPORT pll_utree_t * pll_utree_wraptree(pll_unode_t * root, unsigned int tip_count);
PLL_EXPORT pll_utree_t * pll_utree_wraptree_multi(pll_unode_t * root, unsigned int tip_count,
                                                  unsigned int inner_count);
PLL_EXPORT int pll_utree_is_rooted(const pll_utree_t * tree);
/* src/utree.c:305-463 (pll.h:943-977) */
PLL_EXPORT void pll_utree_show_ascii(const pll_unode_t * tree, int options);
PLL_EXPORT char * pll_utree_export_newick(const pll_unode_t * root,
                                          char * (*cb_serialize)(const pll_unode_t *));
PLL_EXPORT char * pll_utree_export_newick_rooted(const pll_unode_t * root, double root_brlen);
PLL_EXPORT int pll_utree_traverse(pll_unode_t * root, int traversal, int (*cbtrav)(pll_unode_t *),
                                  pll_unode_t ** outbuffer, unsigned int * trav_size);
PLL_EXPORT void pll_utree_create_operations(pll_unode_t * const * trav_buffer,
                                            unsigned int trav_buffer_size, double * branches,
                                            unsigned int * pmatrix_indices, pll_operation_t * ops,
                                            unsigned int * matrix_count, unsigned int * ops_count);
PLL_EXPORT int pll_utree_check_integrity(const pll_utree_t * root);
PLL_EXPORT pll_utree_t * pll_rtree_unroot(pll_rtree_t * tree);
PLL_EXPORT int pll_utree_every(pll_utree_t * tree, int (*cb)(const pll_utree_t *, const pll_unode_t *));
/* src/parse_rtree.y, src/rtree.c (pll.h:890-905, 1005-1030) */
PLL_EXPORT pll_rtree_t * pll_rtree_parse_newick(const char * filename);
PLL_EXPORT pll_rtree_t * pll_rtree_parse_newick_string(const char * s);
PLL_EXPORT void pll_rtree_destroy(pll_rtree_t * root, void (*cb_destroy)(void *));
PLL_EXPORT void pll_rtree_reset_template_indices(pll_rnode_t * node, unsigned int tip_count);
PLL_EXPORT void pll_rtree_graph_destroy(pll_rnode_t * root, void (*cb_destroy)(void *));
PLL_EXPORT pll_rtree_t * pll_rtree_wraptree(pll_rnode_t * root, unsigned int tip_count);
PLL_EXPORT char * pll_rtree_export_newick(const pll_rnode_t * root,
                                          char * (*cb_serialize)(const pll_rnode_t *));
PLL_EXPORT int pll_rtree_traverse(pll_rnode_t * root, int traversal, int (*cbtrav)(pll_rnode_t *),
                                  pll_rnode_t ** outbuffer, unsigned int * trav_size);
PLL_EXPORT void pll_rtree_create_operations(pll_rnode_t * const * trav_buffer,
                                            unsigned int trav_buffer_size, double * branches,
                                            unsigned int * pmatrix_indices, pll_operation_t * ops,
                                            unsigned int * matrix_count, unsigned int * ops_count);

/* src/utree.c:605-633,381-392 (pll.h:965-977); src/rtree.c:106 (pll.h:1003); src/list.c (pll.h:339-344,671-673) */
PLL_EXPORT pll_unode_t * pll_utree_graph_clone(const pll_unode_t * root);
PLL_EXPORT pll_utree_t * pll_utree_clone(const pll_utree_t * root);
PLL_EXPORT int pll_utree_every_const(const pll_utree_t * tree,
                                     int (*cb)(const pll_utree_t * tree, const pll_unode_t *));
PLL_EXPORT void pll_rtree_show_ascii(const pll_rnode_t * root, int options);
typedef struct pll_dlist
{
  struct pll_dlist * next;
  struct pll_dlist * prev;
  void * data;
} pll_dlist_t;
PLL_EXPORT int pll_dlist_append(pll_dlist_t ** dlist, void * data);
PLL_EXPORT int pll_dlist_remove(pll_dlist_t ** dlist, void * data);
PLL_EXPORT int pll_dlist_prepend(pll_dlist_t ** dlist, void * data);

/* ---- topological moves with rollback, src/utree_moves.c (pll.h:141-145, 440-464, 2505-2531); pll_moves.c ---- */
#define PLL_UTREE_MOVE_SPR 1
#define PLL_UTREE_MOVE_NNI 2
#define PLL_UTREE_MOVE_NNI_LEFT 1
#define PLL_UTREE_MOVE_NNI_RIGHT 2

typedef struct pll_utree_rb_s
{
  int move_type;
  union
  {
    struct
    {
      pll_unode_t * p;
      pll_unode_t * r;
      pll_unode_t * rb;
      pll_unode_t * pnb;
      pll_unode_t * pnnb;
      double r_len;
      double pnb_len;
      double pnnb_len;
    } spr;
    struct
    {
      pll_unode_t * p;
      int nni_type;
    } nni;
  };
} pll_utree_rb_t;

PLL_EXPORT int pll_utree_spr(pll_unode_t * p, pll_unode_t * r, pll_utree_rb_t * rb, double * branch_lengths,
                             unsigned int * matrix_indices);
PLL_EXPORT int pll_utree_spr_safe(pll_unode_t * p, pll_unode_t * r, pll_utree_rb_t * rb, double * branch_lengths,
                                  unsigned int * matrix_indices);
PLL_EXPORT int pll_utree_nni(pll_unode_t * p, int type, pll_utree_rb_t * rb);
PLL_EXPORT int pll_utree_rollback(pll_utree_rb_t * rollback, double * branch_lengths, unsigned int * matrix_indices);

/* ---- Fitch parsimony on packed bit vectors (src/fast_parsimony.c, src/stepwise.c) ----------------
 * SURVEY.md 8(f)-4.  The structure keeps the reference's layout (src/pll.h:467-492).  Differences of the
 * CUDA engine: packedvector[i] holds DEVICE addresses (row k of node i starts at packedvector[i] +
 * k*packedvector_count; read one with pll_cuda_download_parsimony_vector); node_cost[], const_cost,
 * informative[] and the counts live on the host and are current whenever a call returns.  The weighted
 * (Sankoff) members belong to objects made by pll_parsimony_create (below).  pll_fastparsimony_init reads the tip
 * states from the partition's device buffers; the parsimony object is independent of the partition afterwards. */
typedef struct pll_parsimony_s
{
  unsigned int tips;
  unsigned int inner_nodes;
  unsigned int sites;
  unsigned int states;
  unsigned int attributes;
  size_t alignment;

  unsigned int ** packedvector;
  unsigned int * node_cost;
  unsigned int packedvector_count;
  unsigned int const_cost;
  int * informative;
  unsigned int informative_count;

  unsigned int score_buffers;
  unsigned int ancestral_buffers;
  double * score_matrix;
  double ** sbuffer;
  unsigned int ** anc_states;
} pll_parsimony_t;

/* src/pll.h:495-500 */
typedef struct pll_pars_buildop_s
{
  unsigned int parent_score_index;
  unsigned int child1_score_index;
  unsigned int child2_score_index;
} pll_pars_buildop_t;

/* src/pll.h:502-508 */
typedef struct pll_pars_recop_s
{
  unsigned int node_score_index;
  unsigned int node_ancestral_index;
  unsigned int parent_score_index;
  unsigned int parent_ancestral_index;
} pll_pars_recop_t;

/* Weighted (Sankoff) parsimony, src/parsimony.c (pll.h:2535-2557).  Objects live on the GPU selected by
 * pll_cuda_set_device() / $PLL_CUDA_DEVICE / $LOCAL_RANK; sbuffer[i] ([site][state] doubles) and anc_states[i]
 * are managed allocations, so clients may read them on the host after any call, as with the reference. */
PLL_EXPORT pll_parsimony_t * pll_parsimony_create(unsigned int tips, unsigned int states, unsigned int sites,
                                                  const double * score_matrix, unsigned int score_buffers,
                                                  unsigned int ancestral_buffers);
PLL_EXPORT int pll_set_parsimony_sequence(pll_parsimony_t * pars, unsigned int tip_index, const pll_state_t * map,
                                          const char * sequence);
/* one launch for the whole operation list; returns pll_parsimony_score of the last parent */
PLL_EXPORT double pll_parsimony_build(pll_parsimony_t * pars, const pll_pars_buildop_t * operations,
                                      unsigned int count);
PLL_EXPORT double pll_parsimony_score(pll_parsimony_t * pars, unsigned int score_buffer_index);
PLL_EXPORT void pll_parsimony_reconstruct(pll_parsimony_t * pars, const pll_state_t * map,
                                          const pll_pars_recop_t * operations, unsigned int count);
/* src/rtree.c:458-520 (pll.h:1032-1040) */
PLL_EXPORT void pll_rtree_create_pars_buildops(pll_rnode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                               pll_pars_buildop_t * ops, unsigned int * ops_count);
PLL_EXPORT void pll_rtree_create_pars_recops(pll_rnode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                             pll_pars_recop_t * ops, unsigned int * ops_count);

/* src/fast_parsimony.c:532-570 (pll.h:2574) */
PLL_EXPORT pll_parsimony_t * pll_fastparsimony_init(const pll_partition_t * partition);
/* src/fast_parsimony.c:721-729 (pll.h:2576): the whole list is ONE launch */
PLL_EXPORT void pll_fastparsimony_update_vectors(pll_parsimony_t * parsimony, const pll_pars_buildop_t * ops,
                                                 unsigned int count);
/* src/fast_parsimony.c:731-773 (pll.h:2580) */
PLL_EXPORT unsigned int pll_fastparsimony_edge_score(const pll_parsimony_t * parsimony,
                                                     unsigned int node1_score_index,
                                                     unsigned int node2_score_index);
/* src/fast_parsimony.c:776-781 (pll.h:2584) */
PLL_EXPORT unsigned int pll_fastparsimony_root_score(const pll_parsimony_t * parsimony, unsigned int root_index);
/* src/parsimony.c:350-383 (pll.h:2559) */
PLL_EXPORT void pll_parsimony_destroy(pll_parsimony_t * pars);
/* src/utree.c:762-785 (pll.h:1003) */
PLL_EXPORT void pll_utree_create_pars_buildops(pll_unode_t * const * trav_buffer, unsigned int trav_buffer_size,
                                               pll_pars_buildop_t * ops, unsigned int * ops_count);
/* src/stepwise.c:883-1082 (pll.h:2587): randomised stepwise addition; every insertion evaluates ALL
 * candidate edges in one launch.  Same tree and cost as the reference for the same seed. */
PLL_EXPORT pll_utree_t * pll_fastparsimony_stepwise(pll_parsimony_t ** list, char * const * labels,
                                                    unsigned int * cost, unsigned int count, unsigned int seed);
/* src/stepwise.c:731-881 (pll.h:2599): adds the taxa missing from `tree` by stepwise addition */
PLL_EXPORT int pll_fastparsimony_stepwise_extend(pll_utree_t * tree, pll_parsimony_t ** pars_list,
                                                 unsigned int pars_count, char * const * labels,
                                                 const unsigned int * tip_msa_idmap, unsigned int seed,
                                                 unsigned int * cost);
/* src/stepwise.c:585-729 (pll.h:2591): one round of subtree pruning and regrafting, all edges per launch */
PLL_EXPORT int pll_fastparsimony_stepwise_spr_round(pll_utree_t * tree, pll_parsimony_t ** pars_list,
                                                    unsigned int pars_count, const unsigned int * tip_msa_idmap,
                                                    unsigned int seed, const int * clv_index_map,
                                                    unsigned int * cost);
/* NEW (additive): the level schedule of pll_fastparsimony_update_vectors (1-based levels; operations of a
 * level neither read nor write what another operation of the level writes); returns the level count, -1 on error */
PLL_EXPORT int pll_cuda_schedule_parsimony_levels(const pll_pars_buildop_t * ops, unsigned int count,
                                                  unsigned int vectors, unsigned int * level_of_op);
/* NEW (additive): a batch of edge scores in one launch; pairs = n x {node1, node2} score indices */
PLL_EXPORT int pll_cuda_fastparsimony_edge_scores(const pll_parsimony_t * parsimony, const unsigned int * pairs,
                                                  unsigned int n, unsigned int * scores);
/* NEW (additive): copy node `index`'s vector (states x packedvector_count words) to the host */
PLL_EXPORT int pll_cuda_download_parsimony_vector(const pll_parsimony_t * parsimony, unsigned int index,
                                                  unsigned int * dst);

/* Re-entrant random numbers, src/random.c (pll.h:534-547, 2592-2612): glibc's random_r family (additive
 * feedback generator r[i] = r[i-3] + r[i-31] for the default 128-byte state), so that seeds give the
 * reference's sequences. */
struct pll_random_data
{
  int * fptr;
  int * rptr;
  int * state;
  int rand_type;
  int rand_deg;
  int rand_sep;
  int * end_ptr;
};
typedef struct pll_random_state_s
{
  struct pll_random_data rdata;
  char * state_buf;
} pll_random_state;
PLL_EXPORT int pll_random_r(struct pll_random_data * buf, int * result);
PLL_EXPORT int pll_srandom_r(unsigned int seed, struct pll_random_data * buf);
PLL_EXPORT int pll_initstate_r(unsigned int seed, char * arg_state, size_t n, struct pll_random_data * buf);
PLL_EXPORT int pll_setstate_r(char * arg_state, struct pll_random_data * buf);
PLL_EXPORT pll_random_state * pll_random_create(unsigned int seed);
PLL_EXPORT int pll_random_getint(pll_random_state * rstate, int maxval);
PLL_EXPORT void pll_random_destroy(pll_random_state * rstate);

#ifdef __cplusplus
}
#endif

#endif /* PLL_B200_H_ */
