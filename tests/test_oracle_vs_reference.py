"""Pin the oracle restatement (oracle/plf_oracle.c) against the UNMODIFIED
reference (oracle/_ref/libpll_ref.so, built from /root/reference/src) run with
PLL_ATTRIB_ARCH_AVX2.  CLVs, scalers, P-matrices and repeat ids bit-for-bit;
logL / derivatives to 1e-12 relative (their summation order over sites is the
only difference).  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa

import importlib

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")


def ref_pmatrix_block(eng):
    """Copy of the reference's contiguous P-matrix block incl. displacement."""
    p = eng.part
    sp, st, rc = p.states_padded, p.states, p.rate_cats
    n = p.prob_matrices * st * sp * rc + (sp - st) * sp
    return np.ctypeslib.as_array(p.pmatrix[0], shape=(n,)).copy()


def model_arrays(eng):
    p = eng.part
    sp, st = p.states_padded, p.states
    ev = [eng.host_array("eigenvecs", i, st * sp) for i in range(p.rate_matrices)]
    iev = [eng.host_array("inv_eigenvecs", i, st * sp) for i in range(p.rate_matrices)]
    evals = [eng.host_array("eigenvals", i, sp) for i in range(p.rate_matrices)]
    freqs = [eng.host_array("frequencies", i, sp) for i in range(p.rate_matrices)]
    rates = np.ctypeslib.as_array(p.rates, shape=(p.rate_cats,)).copy()
    weights = np.ctypeslib.as_array(p.rate_weights, shape=(p.rate_cats,)).copy()
    pinv = np.ctypeslib.as_array(p.prop_invar, shape=(p.rate_matrices,)).copy()
    return ev, iev, evals, freqs, rates, weights, pinv


def oracle_pmatrix_block(orc, eng):
    p = eng.part
    sp, st, rc = p.states_padded, p.states, p.rate_cats
    ev, iev, evals, _, rates, _, pinv = model_arrays(eng)
    n = p.prob_matrices * st * sp * rc + (sp - st) * sp
    block = np.zeros(n)
    ptrs = (oa.dp * p.prob_matrices)()
    for i in range(p.prob_matrices):
        ptrs[i] = C.cast(block.ctypes.data + 8 * i * st * sp * rc, oa.dp)
    orc.orc_update_pmatrix(
        ptrs, st, sp, rc, oa.D_(rates), oa.D_(eng.branch_lengths), oa.U_(eng.matrix_indices),
        oa.U_(eng.params_indices), oa.D_(pinv), oa.ptr_array(evals), oa.ptr_array(ev), oa.ptr_array(iev),
        len(eng.matrix_indices),
    )
    return block


def run_oracle_traversal(orc, eng, block, per_rate):
    """Replay the op list with the oracle on the reference's inputs."""
    p = eng.part
    sp, st, rc, S = p.states_padded, p.states, p.rate_cats, p.sites
    tips = p.tips
    pattern_tip = bool(eng.attributes & capi.PATTERN_TIP)
    span = sp * rc
    msz = st * sp * rc
    clv, scal = {}, {}
    tipchars = {}
    tipmap = None
    if pattern_tip:
        for t in range(tips):
            tipchars[t] = eng.tipchars(t)
        tipmap = np.ctypeslib.as_array(p.tipmap, shape=(256,)).copy()
    else:
        for t in range(tips):
            clv[t] = eng.clv(t)

    def mat(i):
        return C.cast(block.ctypes.data + 8 * i * msz, oa.dp)

    for op in eng.ops:
        par, ps = op.parent_clv_index, op.parent_scaler_index
        c1, m1, s1 = op.child1_clv_index, op.child1_matrix_index, op.child1_scaler_index
        c2, m2, s2 = op.child2_clv_index, op.child2_matrix_index, op.child2_scaler_index
        out = np.zeros(S * span)
        osc = np.zeros(S * (rc if per_rate else 1), dtype=np.uint32) if ps >= 0 else None
        t1, t2 = pattern_tip and c1 < tips, pattern_tip and c2 < tips
        if t1 and t2:
            orc.orc_update_partial_tt(st, sp, S, rc, oa.D_(out), oa.U_(osc), oa.B_(tipchars[c1]),
                                      oa.B_(tipchars[c2]), mat(m1), mat(m2), oa.S_(tipmap), p.maxstates, per_rate)
        elif t1 or t2:
            if t2:
                c1, m1, s1, c2, m2, s2 = c2, m2, s2, c1, m1, s1
            orc.orc_update_partial_ti(st, sp, S, rc, oa.D_(out), oa.U_(osc), oa.B_(tipchars[c1]), oa.D_(clv[c2]),
                                      mat(m1), mat(m2), oa.U_(scal.get(s2)), oa.S_(tipmap), p.maxstates, per_rate)
        else:
            orc.orc_update_partial_ii(st, sp, S, rc, oa.D_(out), oa.U_(osc), oa.D_(clv[c1]), oa.D_(clv[c2]),
                                      mat(m1), mat(m2), oa.U_(scal.get(s1)), oa.U_(scal.get(s2)), per_rate)
        clv[par] = out
        if ps >= 0:
            scal[ps] = osc
    return clv, scal, tipchars, tipmap


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


CASES = [
    # (kind, tips, sites, tree, attrs_extra, per_rate)
    ("dna", 12, 97, "random", capi.PATTERN_TIP, False),
    ("dna", 12, 97, "random", 0, False),
    ("dna", 300, 61, "caterpillar", capi.PATTERN_TIP, False),
    ("dna", 300, 61, "caterpillar", capi.PATTERN_TIP, True),
    ("dna", 200, 33, "caterpillar", 0, True),
    ("aa", 10, 53, "random", capi.PATTERN_TIP, False),
    ("aa", 120, 21, "caterpillar", capi.PATTERN_TIP, False),
    ("aa", 120, 21, "caterpillar", capi.PATTERN_TIP, True),
    ("aa", 110, 17, "caterpillar", 0, False),
    ("g5", 10, 41, "random", capi.PATTERN_TIP, False),
    ("g5", 150, 19, "caterpillar", capi.PATTERN_TIP, False),
    ("g7", 150, 19, "caterpillar", 0, True),
]


def make_ds(kind, tips, sites, tree):
    if kind == "dna":
        return synth.dna_dataset(tips, sites, seed=11, tree_kind=tree, alpha=0.4)
    if kind == "aa":
        return synth.aa_dataset(tips, sites, seed=12, tree_kind=tree, alpha=0.4)
    return synth.generic_dataset(int(kind[1:]), tips, sites, seed=13, tree_kind=tree)


@pytest.mark.parametrize("kind,tips,sites,tree,extra,per_rate", CASES)
def test_traversal_bit_exact(reflib, oracle, kind, tips, sites, tree, extra, per_rate):
    ds = make_ds(kind, tips, sites, tree)
    attrs = capi.ARCH_AVX2 | extra | (capi.RATE_SCALERS if per_rate else 0)
    eng = harness.Engine(reflib, ds, attrs)
    eng.update_pmatrices()
    eng.update_partials()
    ref_block = ref_pmatrix_block(eng)
    my_block = oracle_pmatrix_block(oracle, eng)
    p = eng.part
    msz = p.states * p.states_padded * p.rate_cats
    for mi in eng.matrix_indices:
        a = ref_block[mi * msz:(mi + 1) * msz].reshape(p.rate_cats, p.states, p.states_padded)
        b = my_block[mi * msz:(mi + 1) * msz].reshape(p.rate_cats, p.states, p.states_padded)
        assert np.array_equal(bits(a[:, :, :p.states]), bits(b[:, :, :p.states])), f"pmatrix {mi}"
    clv, scal, tipchars, tipmap = run_oracle_traversal(oracle, eng, ref_block, per_rate)
    n_scaled = 0
    for op in eng.ops:
        ref_clv = eng.clv(op.parent_clv_index)
        assert np.array_equal(bits(ref_clv), bits(clv[op.parent_clv_index])), f"clv {op.parent_clv_index}"
        if op.parent_scaler_index >= 0:
            ref_sc = eng.scaler(op.parent_scaler_index)
            assert np.array_equal(ref_sc, scal[op.parent_scaler_index]), f"scaler {op.parent_scaler_index}"
            n_scaled += int(ref_sc.sum())
    if tree == "caterpillar":
        assert n_scaled > 0, "case was meant to trigger scaling"
    # logL on the root edge
    a, b, m = ds.tree.root_edge
    sp, st, rc = p.states_padded, p.states, p.rate_cats
    ev, iev, evals, freqs, rates, weights, pinv = model_arrays(eng)
    ref_logl, ref_ps = eng.edge_logl(persite=True)
    pw = np.ctypeslib.as_array(p.pattern_weights, shape=(p.sites,)).copy()
    ps = np.zeros(p.sites)
    mptr = C.cast(ref_block.ctypes.data + 8 * m * msz, oa.dp)
    sa, sb = ds.tree.scaler_of.get(a, -1), ds.tree.scaler_of.get(b, -1)
    if (attrs & capi.PATTERN_TIP) and b < p.tips:
        v = oracle.orc_edge_loglikelihood_ti(st, sp, p.sites, rc, oa.D_(clv[a]), oa.U_(scal.get(sa)),
                                             oa.B_(tipchars[b]), oa.S_(tipmap), mptr, oa.ptr_array(freqs),
                                             oa.D_(weights), oa.U_(pw), oa.D_(pinv), None, oa.U_(eng.params_indices),
                                             oa.D_(ps), per_rate)
    else:
        v = oracle.orc_edge_loglikelihood_ii(st, sp, p.sites, rc, oa.D_(clv[a]), oa.U_(scal.get(sa)), None,
                                             oa.D_(clv[b]), oa.U_(scal.get(sb)), None, mptr, oa.ptr_array(freqs),
                                             oa.D_(weights), oa.U_(pw), oa.D_(pinv), None, oa.U_(eng.params_indices),
                                             oa.D_(ps), per_rate)
    assert abs(v - ref_logl) <= 1e-12 * abs(ref_logl)
    np.testing.assert_allclose(ps, ref_ps, rtol=1e-12)
    if not per_rate:
        r_ref = eng.root_logl()
        r = oracle.orc_root_loglikelihood(st, sp, p.sites, rc, oa.D_(clv[a]), None, oa.U_(scal.get(sa)),
                                          oa.ptr_array(freqs), oa.D_(weights), oa.U_(pw), oa.D_(pinv), None,
                                          oa.U_(eng.params_indices), None)
        assert abs(r - r_ref) <= 1e-12 * abs(r_ref)
    # sumtable + derivatives on the root edge
    st_ref = eng.sumtable_alloc()
    eng.update_sumtable(st_ref)
    st_orc = np.zeros_like(st_ref)
    evs = [ev[i] for i in eng.params_indices]
    ievs = [iev[i] for i in eng.params_indices]
    fr = [freqs[i] for i in eng.params_indices]
    if (attrs & capi.PATTERN_TIP) and b < p.tips:
        oracle.orc_update_sumtable_ti(st, sp, p.sites, rc, oa.D_(clv[a]), oa.B_(tipchars[b]), oa.S_(tipmap),
                                      oa.U_(scal.get(sa)), oa.ptr_array(evs), oa.ptr_array(ievs), oa.ptr_array(fr),
                                      oa.D_(st_orc), per_rate)
    else:
        oracle.orc_update_sumtable_ii(st, sp, p.sites, rc, oa.D_(clv[a]), None, oa.D_(clv[b]), None,
                                      oa.U_(scal.get(sa)), oa.U_(scal.get(sb)), oa.ptr_array(evs),
                                      oa.ptr_array(ievs), oa.ptr_array(fr), oa.D_(st_orc), per_rate)
    scale = np.abs(st_ref).max()
    np.testing.assert_allclose(st_orc, st_ref, rtol=1e-9, atol=1e-13 * scale)
    for t in (0.01, 0.1, 0.7):
        d_ref = eng.derivatives(st_ref, t)
        d1, d2 = C.c_double(), C.c_double()
        pin = np.array([pinv[i] for i in eng.params_indices])
        evl = [evals[i] for i in eng.params_indices]
        oracle.orc_likelihood_derivatives(st, sp, p.sites, rc, oa.D_(weights), None, oa.U_(pw), t, oa.D_(pin),
                                          oa.ptr_array(fr), oa.D_(rates), oa.ptr_array(evl), oa.D_(st_ref),
                                          C.byref(d1), C.byref(d2))
        assert abs(d1.value - d_ref[0]) <= 1e-10 * max(1.0, abs(d_ref[0]))
        assert abs(d2.value - d_ref[1]) <= 1e-10 * max(1.0, abs(d_ref[1]))
    eng.close()
