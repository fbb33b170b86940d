/*
 * plf_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar C restatement of the libpll-2 likelihood hot path as executed under
 * PLL_ATTRIB_ARCH_AVX2 (the baseline the CUDA engine is compared with).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it;
 * the product library never links or calls anything in oracle/.
 *
 * Parity pinning: tests/test_oracle_vs_reference.py checks every function
 * here bit-for-bit (CLVs, scalers, P-matrices, repeat ids) or to 1e-13 (logL,
 * derivatives) against oracle/_ref/libpll_ref.so -- the unmodified reference
 * compiled from /root/reference/src -- and against the golden outputs of the
 * reference's own tests committed under tests/golden/.
 */
#ifndef PLF_ORACLE_H_
#define PLF_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long long orc_state_t;

/* core_pmatrix.c:24 / core_pmatrix_avx.c:42 / core_pmatrix_avx2.c:49 */
void orc_update_pmatrix(double ** pmatrix, unsigned int states,
                        unsigned int states_padded, unsigned int rate_cats,
                        const double * rates, const double * branch_lengths,
                        const unsigned int * matrix_indices,
                        const unsigned int * params_indices,
                        const double * prop_invar, double * const * eigenvals,
                        double * const * eigenvecs,
                        double * const * inv_eigenvecs, unsigned int count);

/* core_partials_avx.c:402 (4 states), core_partials_avx2.c:630,1010 (others) */
void orc_update_partial_ii(unsigned int states, unsigned int states_padded,
                           unsigned int sites, unsigned int rate_cats,
                           double * parent_clv, unsigned int * parent_scaler,
                           const double * left_clv, const double * right_clv,
                           const double * left_matrix,
                           const double * right_matrix,
                           const unsigned int * left_scaler,
                           const unsigned int * right_scaler,
                           int per_rate_scalers);

/* core_partials_avx.c:1310 (4), core_partials_avx2.c:343 (20), :49 (others) */
void orc_update_partial_ti(unsigned int states, unsigned int states_padded,
                           unsigned int sites, unsigned int rate_cats,
                           double * parent_clv, unsigned int * parent_scaler,
                           const unsigned char * left_tipchars,
                           const double * right_clv,
                           const double * left_matrix,
                           const double * right_matrix,
                           const unsigned int * right_scaler,
                           const orc_state_t * tipmap, unsigned int maxstates,
                           int per_rate_scalers);

/* core_partials_avx.c:255+992 (4), :124+942 (20), :26+942 (others) */
void orc_update_partial_tt(unsigned int states, unsigned int states_padded,
                           unsigned int sites, unsigned int rate_cats,
                           double * parent_clv, unsigned int * parent_scaler,
                           const unsigned char * left_tipchars,
                           const unsigned char * right_tipchars,
                           const double * left_matrix,
                           const double * right_matrix,
                           const orc_state_t * tipmap, unsigned int maxstates,
                           int per_rate_scalers);

/* core_partials_avx.c:761 / core_partials_avx2.c:1682 (gathers through ids) */
void orc_update_partial_repeats(unsigned int states, unsigned int states_padded,
                                unsigned int parent_sites,
                                unsigned int rate_cats, double * parent_clv,
                                unsigned int * parent_scaler,
                                const double * left_clv,
                                const double * right_clv,
                                const double * left_matrix,
                                const double * right_matrix,
                                const unsigned int * left_scaler,
                                const unsigned int * right_scaler,
                                const unsigned int * parent_id_site,
                                const unsigned int * left_site_id,
                                const unsigned int * right_site_id,
                                int per_rate_scalers);

/* core_likelihood.c:25 (semantics), core_likelihood_avx.c:206 */
double orc_root_loglikelihood(unsigned int states, unsigned int states_padded,
                              unsigned int sites, unsigned int rate_cats,
                              const double * clv, const unsigned int * site_id,
                              const unsigned int * scaler,
                              double * const * frequencies,
                              const double * rate_weights,
                              const unsigned int * pattern_weights,
                              const double * invar_proportion,
                              const int * invar_indices,
                              const unsigned int * freqs_indices,
                              double * persite_lnl);

/* core_likelihood.c:1192, core_likelihood_avx.c:1513; *_site_id may be NULL */
double orc_edge_loglikelihood_ii(unsigned int states,
                                 unsigned int states_padded,
                                 unsigned int sites, unsigned int rate_cats,
                                 const double * clvp,
                                 const unsigned int * parent_scaler,
                                 const unsigned int * parent_site_id,
                                 const double * clvc,
                                 const unsigned int * child_scaler,
                                 const unsigned int * child_site_id,
                                 const double * pmatrix,
                                 double * const * frequencies,
                                 const double * rate_weights,
                                 const unsigned int * pattern_weights,
                                 const double * invar_proportion,
                                 const int * invar_indices,
                                 const unsigned int * freqs_indices,
                                 double * persite_lnl, int per_rate_scalers);

/* core_likelihood.c:352,581 */
double orc_edge_loglikelihood_ti(unsigned int states,
                                 unsigned int states_padded,
                                 unsigned int sites, unsigned int rate_cats,
                                 const double * clvp,
                                 const unsigned int * parent_scaler,
                                 const unsigned char * tipchars,
                                 const orc_state_t * tipmap,
                                 const double * pmatrix,
                                 double * const * frequencies,
                                 const double * rate_weights,
                                 const unsigned int * pattern_weights,
                                 const double * invar_proportion,
                                 const int * invar_indices,
                                 const unsigned int * freqs_indices,
                                 double * persite_lnl, int per_rate_scalers);

/* core_derivatives.c:321 (+repeats :25 through the id arrays) */
void orc_update_sumtable_ii(unsigned int states, unsigned int states_padded,
                            unsigned int sites, unsigned int rate_cats,
                            const double * clvp, const unsigned int * parent_site_id,
                            const double * clvc, const unsigned int * child_site_id,
                            const unsigned int * parent_scaler,
                            const unsigned int * child_scaler,
                            double * const * eigenvecs,
                            double * const * inv_eigenvecs,
                            double * const * freqs, double * sumtable,
                            int per_rate_scalers);

/* core_derivatives.c:473 */
void orc_update_sumtable_ti(unsigned int states, unsigned int states_padded,
                            unsigned int sites, unsigned int rate_cats,
                            const double * clv_inner,
                            const unsigned char * tipchars,
                            const orc_state_t * tipmap,
                            const unsigned int * inner_scaler,
                            double * const * eigenvecs,
                            double * const * inv_eigenvecs,
                            double * const * freqs, double * sumtable,
                            int per_rate_scalers);

/* core_derivatives.c:696 */
void orc_likelihood_derivatives(unsigned int states, unsigned int states_padded,
                                unsigned int sites, unsigned int rate_cats,
                                const double * rate_weights,
                                const int * invariant,
                                const unsigned int * pattern_weights,
                                double branch_length,
                                const double * prop_invar,
                                double * const * freqs, const double * rates,
                                double * const * eigenvals,
                                const double * sumtable, double * d_f,
                                double * dd_f);

/* repeats.c:299-382: class identifiers of a parent from its children's.
 * Returns the number of classes (0 = "no repeats on this node").
 * site_id_parent[sites], id_site_parent[>=classes]; lookup must hold
 * lookup_size entries all equal to 0xFFFFFFFF on entry and is restored. */
unsigned int orc_update_repeats(unsigned int sites,
                                const unsigned int * site_id_left,
                                unsigned int ids_left,
                                const unsigned int * site_id_right,
                                unsigned int ids_right,
                                unsigned int * site_id_parent,
                                unsigned int * id_site_parent,
                                unsigned int * lookup,
                                unsigned int lookup_size);

/* repeats.c:189-254: class identifiers of a tip sequence */
unsigned int orc_update_repeats_tip(unsigned int sites, const orc_state_t * map,
                                    const char * sequence,
                                    unsigned int * site_id,
                                    unsigned int * id_site);

#ifdef __cplusplus
}
#endif
#endif
