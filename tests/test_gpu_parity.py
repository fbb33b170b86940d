"""Parity of the CUDA engine (libpll_b200.so, through its C ABI) with the
UNMODIFIED reference (oracle/_ref/libpll_ref.so run with PLL_ATTRIB_ARCH_AVX2)
on identical seeded inputs.

Contract (BASELINE.json north_star): P-matrices, CLVs, integer scalers and
site-repeat identifiers bit-exact; log-likelihood within 1e-10 relative;
derivatives within 1e-9 relative.  Needs a B200: every test is marked gpu.
"""
import ctypes as C
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")

pytestmark = pytest.mark.gpu

LOGL_RTOL = 1e-10
DERIV_RTOL = 1e-9
CLV_MMA_RTOL = 1e-12


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def make_ds(kind, tips, sites, tree, seed=11, **kw):
    if kind == "dna":
        return synth.dna_dataset(tips, sites, seed=seed, tree_kind=tree, alpha=0.4, **kw)
    if kind == "aa":
        return synth.aa_dataset(tips, sites, seed=seed + 1, tree_kind=tree, alpha=0.4)
    return synth.generic_dataset(int(kind[1:]), tips, sites, seed=seed + 2, tree_kind=tree)


def pair(reflib, cudalib, ds, extra, per_rate=False):
    flags = extra | (capi.RATE_SCALERS if per_rate else 0)
    ref = harness.Engine(reflib, ds, capi.ARCH_AVX2 | flags)
    gpu = harness.Engine(cudalib, ds, capi.ARCH_CUDA | flags)
    return ref, gpu


def assert_rel(a, b, rtol, what):
    assert abs(a - b) <= rtol * max(abs(b), 1e-300), f"{what}: {a!r} vs {b!r} rel {abs(a - b) / abs(b):.3e}"


def check_edge_and_derivatives(ref, gpu, ds, per_rate, edges=None):
    for edge in edges or [ds.tree.root_edge]:
        l_ref, ps_ref = ref.edge_logl(edge, persite=True)
        l_gpu, ps_gpu = gpu.edge_logl(edge, persite=True)
        assert_rel(l_gpu, l_ref, LOGL_RTOL, f"edge logL {edge}")
        np.testing.assert_allclose(ps_gpu, ps_ref, rtol=1e-12, atol=0)
        st_ref = ref.sumtable_alloc()
        st_gpu = gpu.sumtable_alloc()
        ref.update_sumtable(st_ref, edge)
        gpu.update_sumtable(st_gpu, edge)
        for t in (0.003, 0.1, 0.9):
            d_ref = ref.derivatives(st_ref, t, edge)
            d_gpu = gpu.derivatives(st_gpu, t, edge)
            # the derivative is a sum of per-site terms of both signs: relative
            # to the magnitude the terms reach (golden fragility)
            scale1 = max(abs(d_ref[0]), 1e-6 * ds.sites)
            scale2 = max(abs(d_ref[1]), 1e-6 * ds.sites)
            assert abs(d_gpu[0] - d_ref[0]) <= DERIV_RTOL * scale1, (t, d_gpu, d_ref)
            assert abs(d_gpu[1] - d_ref[1]) <= DERIV_RTOL * scale2, (t, d_gpu, d_ref)


CASES = [
    # kind, tips, sites, tree, attrs, per_rate
    ("dna", 12, 97, "random", capi.PATTERN_TIP, False),
    ("dna", 12, 97, "random", 0, False),
    ("dna", 40, 1501, "random", capi.PATTERN_TIP, False),
    ("dna", 300, 61, "caterpillar", capi.PATTERN_TIP, False),
    ("dna", 300, 61, "caterpillar", capi.PATTERN_TIP, True),
    ("dna", 200, 33, "caterpillar", 0, True),
    ("dna", 200, 33, "caterpillar", 0, False),
    ("aa", 10, 53, "random", capi.PATTERN_TIP, False),
    ("aa", 10, 53, "random", 0, False),
    ("aa", 120, 21, "caterpillar", capi.PATTERN_TIP, False),
    ("aa", 120, 21, "caterpillar", capi.PATTERN_TIP, True),
    ("aa", 110, 17, "caterpillar", 0, False),
    # more sites than one sweep of the persistent grid covers (8 sites per warp step)
    ("aa", 6, 25003, "random", capi.PATTERN_TIP, False),
    ("aa", 6, 25003, "random", 0, True),
    ("g5", 10, 41, "random", capi.PATTERN_TIP, False),
    ("g5", 150, 19, "caterpillar", capi.PATTERN_TIP, False),
    ("g7", 150, 19, "caterpillar", 0, True),
    ("g7", 12, 40, "random", 0, False),
]


def assert_clv_equal(a, b, exact, what):
    """bit-exact, or -- on the 20-state tensor-core (DMMA) kernels, whose
    accumulation order differs from the AVX2 lanes -- equal to a few ulp per level"""
    if exact:
        assert np.array_equal(bits(a), bits(b)), what
    else:
        np.testing.assert_allclose(b, a, rtol=CLV_MMA_RTOL, atol=0, err_msg=what)


# 20-state cases run twice: PLF_AA_MMA=0 (DFMA kernels, CLVs bit-exact) and the
# default DMMA kernels (CLVs within CLV_MMA_RTOL, integer scalers still exact)
CASES_MODES = [(c, m) for c in CASES for m in (("dfma", "dmma") if c[0] == "aa" else ("default",))]


@pytest.mark.parametrize("case,mode", CASES_MODES, ids=lambda v: v if isinstance(v, str) else "-".join(map(str, v)))
def test_traversal_parity(reflib, cudalib, case, mode, monkeypatch):
    kind, tips, sites, tree, extra, per_rate = case
    if mode != "default":
        monkeypatch.setenv("PLF_AA_MMA", "0" if mode == "dfma" else "1")
    exact = mode != "dmma"
    ds = make_ds(kind, tips, sites, tree)
    ref, gpu = pair(reflib, cudalib, ds, extra, per_rate)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    p = ref.part
    st = p.states
    for mi in ref.matrix_indices:
        a = ref.pmatrix(mi).reshape(p.rate_cats, st, p.states_padded)[:, :, :st]
        b = gpu.pmatrix(mi).reshape(p.rate_cats, st, p.states_padded)[:, :, :st]
        assert np.array_equal(bits(a), bits(b)), f"pmatrix {mi}"
    n_scaled = 0
    for op in ref.ops:
        a, b = ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index)
        assert_clv_equal(a, b, exact, f"clv {op.parent_clv_index}")
        if op.parent_scaler_index >= 0:
            sa, sb = ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index)
            assert np.array_equal(sa, sb), f"scaler {op.parent_scaler_index}"
            n_scaled += int(sa.sum())
    if tree == "caterpillar":
        assert n_scaled > 0, "case was meant to trigger scaling"
    check_edge_and_derivatives(ref, gpu, ds, per_rate)
    # a tip edge and a second inner edge
    last = ref.ops[len(ref.ops) - 1]
    extra_edges = [(last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index)]
    if not (extra & capi.PATTERN_TIP and last.child1_clv_index < ds.tree.tips and last.parent_clv_index < ds.tree.tips):
        check_edge_and_derivatives(ref, gpu, ds, per_rate, extra_edges)
    if not per_rate:  # the reference's root logL ignores per-rate scalers
        r_ref, rp_ref = ref.root_logl(persite=True)
        r_gpu, rp_gpu = gpu.root_logl(persite=True)
        assert_rel(r_gpu, r_ref, LOGL_RTOL, "root logL")
        np.testing.assert_allclose(rp_gpu, rp_ref, rtol=1e-12)
    ref.close()
    gpu.close()


@pytest.mark.parametrize("kind,extra", [("dna", capi.PATTERN_TIP), ("dna", 0), ("aa", capi.PATTERN_TIP), ("g5", 0)])
@pytest.mark.parametrize("pinv", [0.3])
def test_invariant_sites_parity(reflib, cudalib, kind, extra, pinv):
    ds = make_ds(kind, 9, 257, "random", seed=21)
    # make a good share of columns invariant
    seqs = [bytearray(s) for s in ds.seqs]
    for col in range(0, ds.sites, 3):
        for s in seqs:
            s[col] = seqs[0][col]
    ds.seqs = [bytes(s) for s in seqs]
    ds.prop_invar = pinv
    ds.pattern_weights = np.random.default_rng(3).integers(1, 6, size=ds.sites).astype(np.uint32)
    ref, gpu = pair(reflib, cudalib, ds, extra)
    inv_ref = np.ctypeslib.as_array(ref.part.invariant, shape=(ds.sites,)).copy()
    inv_gpu = np.ctypeslib.as_array(gpu.part.invariant, shape=(ds.sites,)).copy()
    assert np.array_equal(inv_ref, inv_gpu)
    assert (inv_ref >= 0).sum() > 10
    cnt_ref = np.zeros(ds.states, dtype=np.uint32)
    cnt_gpu = np.zeros(ds.states, dtype=np.uint32)
    n_ref = reflib.pll_count_invariant_sites(ref.p, cnt_ref.ctypes.data_as(capi.c_uint_p))
    n_gpu = cudalib.pll_count_invariant_sites(gpu.p, cnt_gpu.ctypes.data_as(capi.c_uint_p))
    assert n_ref == n_gpu and np.array_equal(cnt_ref, cnt_gpu)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    for mi in ref.matrix_indices[:4]:
        assert np.array_equal(bits(ref.pmatrix(mi)), bits(gpu.pmatrix(mi)))
    check_edge_and_derivatives(ref, gpu, ds, False)
    assert_rel(gpu.root_logl(), ref.root_logl(), LOGL_RTOL, "root logL +I")
    ref.close()
    gpu.close()


@pytest.mark.parametrize("cats", [1, 2, 3, 4, 8, 16])
@pytest.mark.parametrize("extra", [capi.PATTERN_TIP, 0])
def test_dna_edge_kernels_rate_counts(reflib, cudalib, cats, extra):
    """The streaming 4-state edge / sumtable / derivative kernels are specialised per
    rate count (1, 2, 4, 8); 3 and 16 take the generic kernels.  Site counts that are
    not a multiple of the 64-site tile, +I, non-unit pattern weights, scaling."""
    ds = synth.dna_dataset(90, 1003, seed=41 + cats, tree_kind="caterpillar", alpha=0.5, cats=cats)
    ds.prop_invar = 0.2
    ds.pattern_weights = np.random.default_rng(5).integers(1, 9, size=ds.sites).astype(np.uint32)
    ref, gpu = pair(reflib, cudalib, ds, extra)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    last = ref.ops[len(ref.ops) - 1]
    first = ref.ops[0]
    edges = [ds.tree.root_edge,
             (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index),
             (first.parent_clv_index, first.child1_clv_index, first.child1_matrix_index),
             (last.parent_clv_index, last.child2_clv_index, last.child2_matrix_index)]
    check_edge_and_derivatives(ref, gpu, ds, False, edges)
    r_ref, rp_ref = ref.root_logl(persite=True)
    r_gpu, rp_gpu = gpu.root_logl(persite=True)
    assert_rel(r_gpu, r_ref, LOGL_RTOL, "root logL")
    np.testing.assert_allclose(rp_gpu, rp_ref, rtol=1e-12)
    ref.close()
    gpu.close()


@pytest.mark.parametrize("kind", ["dna", "aa"])
def test_edge_fast_and_generic_kernels_agree(cudalib, monkeypatch, kind):
    """PLF_EDGE_FAST=0 forces the generic kernels: both families answer the same calls."""
    if kind == "dna":
        ds = synth.dna_dataset(30, 5000, seed=77, tree_kind="random", alpha=0.6)
    else:
        ds = synth.aa_dataset(12, 5003, seed=78, tree_kind="random", alpha=0.6)
    out = []
    for flag in ("1", "0"):
        monkeypatch.setenv("PLF_EDGE_FAST", flag)
        gpu = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
        logl = gpu.full_traversal()
        st = gpu.sumtable_alloc()
        gpu.update_sumtable(st)
        out.append((logl, gpu.root_logl(), *gpu.derivatives(st, 0.07)))
        gpu.close()
    for a, b in zip(*out):
        assert abs(a - b) <= 1e-12 * abs(b), out


@pytest.mark.parametrize("cats", [1, 2, 3, 8])
@pytest.mark.parametrize("per_rate", [False, True])
def test_aa_kernels_rate_counts(reflib, cudalib, cats, per_rate):
    """20-state DMMA kernels: the streaming variant is specialised for 1, 2, 4, 8 rate
    categories, 3 takes the direct-load DMMA kernel.  Deep caterpillar => scaling."""
    ds = synth.aa_dataset(400, 45, seed=51 + cats, tree_kind="caterpillar", alpha=0.4, cats=cats)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP, per_rate)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    n_scaled = 0
    for op in ref.ops:
        assert_clv_equal(ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index), False, f"clv {op.parent_clv_index}")
        sa, sb = ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index)
        assert np.array_equal(sa, sb), f"scaler {op.parent_scaler_index}"
        n_scaled += int(sa.sum())
    assert n_scaled > 0
    check_edge_and_derivatives(ref, gpu, ds, per_rate)
    ref.close()
    gpu.close()


@pytest.mark.parametrize("kind,extra,per_rate", [("dna", capi.PATTERN_TIP, False), ("dna", 0, False), ("dna", 0, True),
                                                  ("aa", capi.PATTERN_TIP, False), ("g7", 0, False)])
def test_node_ancestral_parity(reflib, cudalib, kind, extra, per_rate):
    """pll_compute_node_ancestral (src/likelihood.c:762): posterior state probabilities per site at an
    inner node, against an inner neighbour and against a tip neighbour; deep tree => scaled CLVs."""
    ds = make_ds(kind, 150, 97, "caterpillar", seed=81)
    ref, gpu = pair(reflib, cudalib, ds, extra, per_rate)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    t = ds.tree
    last, first = ref.ops[len(ref.ops) - 1], ref.ops[0]
    cases = [(t.root_edge[0], t.root_edge[1], t.root_edge[2]),
             (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index),
             (first.parent_clv_index, first.child1_clv_index, first.child1_matrix_index)]
    for node, other, m in cases:
        out = []
        for lib, e in ((reflib, ref), (cudalib, gpu)):
            anc = np.zeros(ds.sites * ds.states)
            rc = lib.pll_compute_node_ancestral(e.p, node, t.scaler_of.get(node, -1), other, t.scaler_of.get(other, -1),
                                                m, e.params_indices.ctypes.data_as(capi.c_uint_p),
                                                anc.ctypes.data_as(capi.c_double_p))
            assert rc == 1, lib.errmsg
            out.append(anc.reshape(ds.sites, ds.states))
        np.testing.assert_allclose(out[1].sum(axis=1), 1.0, rtol=1e-12)
        np.testing.assert_allclose(out[1], out[0], rtol=1e-10, atol=1e-300)
    ref.close()
    gpu.close()


ASC_TYPES = {"lewis": capi.AB_LEWIS, "felsenstein": capi.AB_FELSENSTEIN, "stamatakis": capi.AB_STAMATAKIS}


@pytest.mark.parametrize("kind,extra", [("dna", capi.PATTERN_TIP), ("dna", 0), ("aa", capi.PATTERN_TIP), ("aa", 0), ("g5", 0)])
def test_ascertainment_bias_parity(reflib, cudalib, kind, extra):
    """PLL_ATTRIB_AB_*: the `states` pseudo-sites go through the CLV kernels with the alignment, the
    correction terms of logL and of the derivatives follow src/likelihood.c:24-120,190-268,342-440 and
    src/core_derivatives.c:851-924 (reference test: test/src/asc-bias.c).  A deep tree so that the
    pseudo-sites carry scaling factors too."""
    ds = make_ds(kind, 120 if kind != "aa" else 160, 203, "caterpillar", seed=71)
    ref, gpu = pair(reflib, cudalib, ds, extra | capi.AB_FLAG)
    w = np.random.default_rng(9).integers(1, 6, size=ds.states).astype(np.uint32)
    last = ref.ops[len(ref.ops) - 1]
    edges = [ds.tree.root_edge, (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index),
             (ref.ops[0].parent_clv_index, ref.ops[0].child1_clv_index, ref.ops[0].child1_matrix_index)]
    for name, typ in ASC_TYPES.items():
        for lib, e in ((reflib, ref), (cudalib, gpu)):
            assert lib.pll_set_asc_bias_type(e.p, typ) == 1, lib.errmsg
            lib.pll_set_asc_state_weights(e.p, w.ctypes.data_as(capi.c_uint_p))
            e.update_pmatrices()
            e.update_partials()
        for edge in edges:
            l_ref, l_gpu = ref.edge_logl(edge), gpu.edge_logl(edge)
            assert np.isfinite(l_ref)
            assert_rel(l_gpu, l_ref, LOGL_RTOL, f"{name} edge logL {edge}")
            st_ref, st_gpu = ref.sumtable_alloc(), gpu.sumtable_alloc()
            ref.update_sumtable(st_ref, edge)
            gpu.update_sumtable(st_gpu, edge)
            for t in (0.01, 0.2):
                d_ref, d_gpu = ref.derivatives(st_ref, t, edge), gpu.derivatives(st_gpu, t, edge)
                for a, b in zip(d_gpu, d_ref):
                    assert abs(a - b) <= DERIV_RTOL * max(abs(b), 1e-6 * ds.sites), (name, edge, t, d_gpu, d_ref)
        if ds.states == ds.rate_cats:
            # src/likelihood.c:180 locates the pseudo-sites at sites + states - rate_cats: only defined
            # behaviour in the reference when the two counts agree (see asc_loglikelihood in pll_host.c)
            assert_rel(gpu.root_logl(), ref.root_logl(), LOGL_RTOL, f"{name} root logL")
        else:
            assert np.isfinite(gpu.root_logl())
    # correction off again: plain likelihood on a partition that carries the pseudo-sites
    for lib, e in ((reflib, ref), (cudalib, gpu)):
        assert lib.pll_set_asc_bias_type(e.p, 0) == 1
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL, correction off")
    # +I is incompatible (src/models.c:500-507)
    assert cudalib.pll_set_asc_bias_type(gpu.p, capi.AB_LEWIS) == 1
    assert cudalib.pll_update_invariant_sites_proportion(gpu.p, 0, 0.2) == 0 and cudalib.errno == 117
    ref.close()
    gpu.close()


REPEAT_CASES = [
    ("dna", 24, 400, "random", False, (0.002, 0.05)),
    ("dna", 64, 3000, "random", False, (0.002, 0.05)),
    ("dna", 300, 128, "caterpillar", False, (0.02, 0.22)),
    ("dna", 300, 128, "caterpillar", True, (0.02, 0.22)),
    ("aa", 40, 300, "random", False, (0.002, 0.05)),
    ("g5", 30, 200, "random", False, (0.002, 0.05)),
]


REPEAT_CASES_MODES = [(c, m) for c in REPEAT_CASES for m in (("dfma", "dmma") if c[0] == "aa" else ("default",))]


@pytest.mark.parametrize("case,mode", REPEAT_CASES_MODES,
                         ids=lambda v: v if isinstance(v, str) else "-".join(map(str, v[:5])))
def test_site_repeats_parity(reflib, cudalib, case, mode, monkeypatch):
    kind, tips, sites, tree, per_rate, brlen = case
    if mode != "default":
        monkeypatch.setenv("PLF_AA_MMA", "0" if mode == "dfma" else "1")
    exact = mode != "dmma"
    if kind == "dna":
        ds = synth.dna_dataset(tips, sites, seed=31, tree_kind=tree, alpha=0.3, brlen=brlen)
    elif kind == "aa":
        ds = synth.aa_dataset(tips, sites, seed=32, tree_kind=tree, alpha=0.3, brlen=brlen)
    else:
        ds = synth.generic_dataset(5, tips, sites, seed=33, tree_kind=tree, brlen=brlen)
    ref, gpu = pair(reflib, cudalib, ds, capi.SITE_REPEATS, per_rate)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    compressed = 0
    for node in range(ds.tree.nodes):
        ids_r, sid_r, ids_site_r = ref.repeat_ids(node)
        ids_g, sid_g, ids_site_g = gpu.repeat_ids(node)
        assert ids_r == ids_g, f"class count of node {node}"
        if ids_r:
            compressed += node >= tips
            assert np.array_equal(sid_r, sid_g), f"site_id of node {node}"
            assert np.array_equal(ids_site_r, ids_site_g), f"id_site of node {node}"
        assert ref.clv_size(node) == gpu.clv_size(node)
    assert compressed > 0, "case was meant to compress inner nodes"
    n_scaled = 0
    for op in ref.ops:
        a, b = ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index)
        assert_clv_equal(a, b, exact, f"clv {op.parent_clv_index}")
        sa, sb = ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index)
        assert np.array_equal(sa, sb), f"scaler {op.parent_scaler_index}"
        n_scaled += int(sa.sum())
    if tree == "caterpillar":
        assert n_scaled > 0
    last = ref.ops[len(ref.ops) - 1]
    edges = [ds.tree.root_edge, (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index)]
    check_edge_and_derivatives(ref, gpu, ds, per_rate, edges)
    if not per_rate:
        assert_rel(gpu.root_logl(), ref.root_logl(), LOGL_RTOL, "root logL")
    # a second traversal without recomputing identifiers (update_repeats = 0)
    bl = ds.tree.branch_lengths[ref.matrix_indices] * 1.5
    for e, lib in ((ref, reflib), (gpu, cudalib)):
        e.update_pmatrices(branch_lengths=bl)
        lib.pll_update_partials_rep(e.p, e.ops, len(e.ops), 0)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL after re-traversal")
    ref.close()
    gpu.close()


def test_unrooted_example_golden(cudalib):
    """Config 1: examples/unrooted/unrooted.c:43-124 -- 4 tips x 6 sites GTR+G4;
    expected values from the reference's own run."""
    from test_reference_examples import run_unrooted

    vals = run_unrooted(cudalib, capi.ARCH_CUDA)
    assert vals == pytest.approx([-33.387713, -34.550204, -36.830297], abs=5e-7)


def test_buffer_reuse_in_one_op_list(reflib, cudalib):
    """test/src/derivatives.c:91-98 recycles a CLV index inside one operation
    list; the level scheduler must keep the sequential meaning."""
    ds = make_ds("dna", 8, 120, "random", seed=41)
    t = ds.tree
    ops = t.ops.copy()
    # rewrite the last-but-one op's parent into a recycled CLV: append an op that
    # recomputes the first inner node from two other tips after it was consumed
    first_parent = int(ops[0][0])
    extra_op = np.array([[first_parent, int(ops[0][1]), 2, 2, -1, 3, 3, -1]], dtype=np.int64)
    ds.tree.ops = np.concatenate([ops, extra_op])
    ref, gpu = pair(reflib, cudalib, ds, 0)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    for op in ref.ops:
        assert np.array_equal(bits(ref.clv(op.parent_clv_index)), bits(gpu.clv(op.parent_clv_index)))
        assert np.array_equal(ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index))
    ref.close()
    gpu.close()


def test_device_expm1_mode_within_tolerance(reflib, cudalib, monkeypatch):
    """PLL_CUDA_DEVICE_EXPM1=1 evaluates expm1 on the GPU: P-matrices may differ
    in the last place, log-likelihood stays within the 1e-10 contract."""
    ds = make_ds("dna", 60, 500, "random", seed=51)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    a, b = ref.pmatrix(3), gpu.pmatrix(3)
    np.testing.assert_allclose(b, a, rtol=1e-14, atol=1e-17)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL, device expm1")
    ref.close()
    gpu.close()


def test_full_size_properties(cudalib):
    """Config-2-shaped run (100 taxa, pattern tips, GTR+G4) at a size the CPU
    reference would take minutes for: size-independent properties instead.
    (a) logL is additive over contiguous site slices; (b) the per-site values
    sum to the total; (c) doubling pattern weights doubles logL."""
    sites = 200_000
    ds = synth.dna_dataset(100, sites, seed=1, simulate_down_tree=False)
    full = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    total, persite = (full.full_traversal(), None)
    total2, persite = full.edge_logl(persite=True)
    assert total == total2
    assert_rel(float(np.sum(persite)), total, 1e-12, "sum of per-site logL")
    parts = 0.0
    for lo, hi in ((0, 70_016), (70_016, sites)):
        e = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP, sites_slice=slice(lo, hi))
        parts += e.full_traversal()
        e.close()
    assert_rel(parts, total, 1e-12, "site-slice additivity")
    w = np.full(sites, 2, dtype=np.uint32)
    cudalib.pll_set_pattern_weights(full.p, w.ctypes.data_as(capi.c_uint_p))
    assert_rel(full.edge_logl(), 2 * total, 1e-13, "pattern weights")
    full.close()


def test_oracle_port_agrees_with_cuda(oracle, cudalib):
    """The scalar C restatement (oracle/plf_oracle.c) replayed on the CUDA
    engine's own P-matrices gives the same CLVs and scalers bit for bit."""
    from test_oracle_vs_reference import run_oracle_traversal

    ds = make_ds("dna", 150, 77, "caterpillar", seed=61)
    gpu = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    gpu.update_pmatrices()
    gpu.update_partials()
    p = gpu.part
    msz = p.states * p.states_padded * p.rate_cats
    block = np.zeros(p.prob_matrices * msz + 64)
    for mi in gpu.matrix_indices:
        block[mi * msz:(mi + 1) * msz] = gpu.pmatrix(mi)
    clv, scal, _, _ = run_oracle_traversal(oracle, gpu, block, False)
    for op in gpu.ops:
        assert np.array_equal(bits(gpu.clv(op.parent_clv_index)), bits(clv[op.parent_clv_index]))
        assert np.array_equal(gpu.scaler(op.parent_scaler_index), scal[op.parent_scaler_index])
    gpu.close()


def test_newick_to_loglikelihood_end_to_end(reflib, cudalib):
    """The call sequence of examples/newick-fasta-unrooted: Newick string -> pll_utree_traverse ->
    pll_utree_create_operations (this library's tree layer, pll_tree.c) -> P-matrices, CLV updates and the
    edge log-likelihood at the virtual root, on the CUDA engine and on the reference with the same lists."""
    import ctypes as C
    import test_tree_cpu as tt

    own = tt.bind(C.CDLL(pkg.LIB_PATH), True)
    rng = np.random.default_rng(91)
    tips, sites = 37, 1201
    tree = own.pll_utree_parse_newick_string(tt.random_newick(rng, tips).encode())
    assert tree
    t = tree.contents
    n_nodes = t.tip_count + t.inner_count
    buf = (C.POINTER(tt.UNode) * n_nodes)()
    size = C.c_uint(0)
    assert own.pll_utree_traverse(t.vroot, 1, tt.UCB(lambda n: 1), buf, C.byref(size)) == 1
    ops = (capi.Operation * n_nodes)()
    branches = (C.c_double * n_nodes)()
    pm = (C.c_uint * n_nodes)()
    n_mat, n_ops = C.c_uint(0), C.c_uint(0)
    own.pll_utree_create_operations(buf, size.value, branches, pm, ops, C.byref(n_mat), C.byref(n_ops))
    assert n_ops.value == tips - 2 and n_mat.value == 2 * tips - 3
    seqs = synth.mutate_alignment(tips, sites, rng, synth.DNA_CODES, synth.DNA_AMBIG)
    rates = synth.gamma_rates(0.8, 4)
    params = np.zeros(4, dtype=np.uint32)
    root = t.vroot.contents
    logl = []
    for lib, arch in ((reflib, capi.ARCH_AVX2), (cudalib, capi.ARCH_CUDA)):
        p = lib.pll_partition_create(tips, tips - 2, 4, sites, 1, 2 * tips - 3, 4, tips - 2, arch | capi.PATTERN_TIP)
        assert p, lib.errmsg
        lib.pll_set_frequencies(p, 0, synth.GTR_FREQS.ctypes.data_as(capi.c_double_p))
        lib.pll_set_subst_params(p, 0, synth.GTR_RATES.ctypes.data_as(capi.c_double_p))
        lib.pll_set_category_rates(p, rates.ctypes.data_as(capi.c_double_p))
        for i in range(tips):
            label = t.nodes[i].contents.label.decode()           # tips are t0..tN-1 in the Newick string
            assert lib.pll_set_tip_states(p, t.nodes[i].contents.clv_index, lib.map("pll_map_nt"), seqs[int(label[1:])]) == 1
        assert lib.pll_update_prob_matrices(p, params.ctypes.data_as(capi.c_uint_p), pm, branches, n_mat.value) == 1
        lib.pll_update_partials(p, ops, n_ops.value)
        logl.append(lib.pll_compute_edge_loglikelihood(p, root.clv_index, root.scaler_index, root.back.contents.clv_index,
                                                       root.back.contents.scaler_index, root.pmatrix_index,
                                                       params.ctypes.data_as(capi.c_uint_p), None))
        lib.pll_partition_destroy(p)
    own.pll_utree_destroy(tree, None)
    assert np.isfinite(logl[0]) and logl[0] < 0
    assert_rel(logl[1], logl[0], LOGL_RTOL, "edge logL from a Newick tree")


EDGE_SHAPES = [
    # kind, tips, sites, cats, attrs, per_rate
    ("dna", 3, 1, 4, capi.PATTERN_TIP, False),     # one site, one operation
    ("dna", 3, 1, 1, 0, False),
    ("dna", 4, 2, 3, capi.PATTERN_TIP, True),      # rate count that is not a power of two
    ("dna", 5, 7, 5, 0, False),
    ("dna", 4, 31, 7, capi.PATTERN_TIP, False),
    ("dna", 6, 33, 2, 0, True),
    ("dna", 7, 65, 8, capi.PATTERN_TIP, False),    # one site more than a 64-site tile
    ("dna", 5, 63, 16, 0, False),
    ("dna", 4, 129, 32, capi.PATTERN_TIP, False),  # largest specialised rate count
    ("aa", 3, 1, 4, capi.PATTERN_TIP, False),
    ("aa", 4, 7, 1, 0, False),                      # fewer sites than one 8-site DMMA block
    ("aa", 5, 9, 3, capi.PATTERN_TIP, True),
    ("aa", 4, 33, 2, 0, False),
    ("aa", 6, 17, 8, capi.PATTERN_TIP, False),
    ("g2", 5, 19, 4, 0, False),                     # binary data
    ("g2", 4, 5, 3, capi.PATTERN_TIP, True),
    ("g3", 5, 21, 4, capi.PATTERN_TIP, False),
    ("g11", 4, 13, 2, 0, False),
    ("g23", 4, 9, 4, capi.PATTERN_TIP, False),      # more states than amino acids (padded to 24)
]


@pytest.mark.parametrize("kind,tips,sites,cats,extra,per_rate", EDGE_SHAPES)
def test_small_and_odd_shapes(reflib, cudalib, kind, tips, sites, cats, extra, per_rate):
    """Shapes around the tile sizes of the specialised kernels and shapes only the generic kernels take:
    single sites, single operations, rate counts 1..32 incl. non powers of two, 2 to 23 states."""
    if kind == "dna":
        ds = synth.dna_dataset(tips, sites, seed=101, cats=cats, alpha=0.5)
    elif kind == "aa":
        ds = synth.aa_dataset(tips, sites, seed=102, cats=cats, alpha=0.5)
    else:
        ds = synth.generic_dataset(int(kind[1:]), tips, sites, seed=103, cats=cats)
    ref, gpu = pair(reflib, cudalib, ds, extra, per_rate)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    for op in ref.ops:
        assert_clv_equal(ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index), kind != "aa",
                         f"clv {op.parent_clv_index}")
        assert np.array_equal(ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index))
    last = ref.ops[len(ref.ops) - 1]
    edges = [ds.tree.root_edge, (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index)]
    check_edge_and_derivatives(ref, gpu, ds, per_rate, edges)
    if not per_rate:
        assert_rel(gpu.root_logl(), ref.root_logl(), LOGL_RTOL, "root logL")
    ref.close()
    gpu.close()


def test_concurrent_partitions_from_threads(cudalib):
    """Distinct partitions used concurrently from distinct threads (the RAxML-NG pattern, SURVEY 8b
    "Threading"): every thread owns a partition, a context and a stream; results equal the serial ones."""
    import threading

    datasets = [synth.dna_dataset(20 + 3 * i, 4000 + 137 * i, seed=200 + i, alpha=0.5) for i in range(4)]
    datasets.append(synth.aa_dataset(12, 1500, seed=210, alpha=0.5))

    def evaluate(ds):
        e = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
        out = []
        for _ in range(5):
            logl = e.full_traversal()
            st = e.sumtable_alloc()
            e.update_sumtable(st)
            out.append((logl, *e.derivatives(st, 0.11)))
        e.close()
        return out

    serial = [evaluate(ds) for ds in datasets]
    results = [None] * len(datasets)
    errors = []

    def work(i):
        try:
            results[i] = evaluate(datasets[i])
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(datasets))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    assert results == serial


def test_out_of_memory_is_an_error_not_a_crash(cudalib):
    """A partition that cannot fit in HBM (100 taxa x 60M sites = 750 GB of CLVs): pll_partition_create returns
    NULL with pll_errno set, frees what it had allocated, and the next partition works."""
    p = cudalib.pll_partition_create(100, 98, 4, 60_000_000, 1, 197, 4, 98, capi.ARCH_CUDA | capi.PATTERN_TIP)
    assert not p
    assert cudalib.errno in (112, 900), (cudalib.errno, cudalib.errmsg)
    ds = synth.dna_dataset(8, 500, seed=3)
    e = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    assert np.isfinite(e.full_traversal())
    e.close()


@pytest.mark.parametrize("kind,extra", [("dna", capi.PATTERN_TIP), ("dna", capi.SITE_REPEATS), ("aa", capi.PATTERN_TIP)])
def test_repeated_traversals_replay_a_cuda_graph(reflib, cudalib, kind, extra, monkeypatch):
    """The same operation list evaluated again and again with new branch lengths (what branch-length and
    model optimisation do): from the third call on the traversal is a CUDA graph replay.  Every
    evaluation must equal the reference's, and the graph-free path (PLF_GRAPH=0)."""
    ds = make_ds(kind, 40, 777, "random", seed=131)
    ref, gpu = pair(reflib, cudalib, ds, extra)
    monkeypatch.setenv("PLF_GRAPH", "0")
    plain = harness.Engine(cudalib, ds, capi.ARCH_CUDA | extra)
    monkeypatch.delenv("PLF_GRAPH")
    rng = np.random.default_rng(7)
    launches = []
    for it in range(6):
        bl = ds.tree.branch_lengths[ref.matrix_indices] * rng.uniform(0.5, 1.5)
        vals = []
        for e in (ref, gpu, plain):
            e.update_pmatrices(branch_lengths=bl)
            before = cudalib.pll_cuda_kernel_launches()
            e.update_partials()
            if e is gpu:
                launches.append(cudalib.pll_cuda_kernel_launches() - before)
            vals.append(e.edge_logl())
        assert_rel(vals[1], vals[0], LOGL_RTOL, f"iteration {it}")
        assert vals[1] == vals[2], "graph replay and plain launches must give the same bits"
    assert len(set(launches[1:])) == 1 and launches[1] > 0, launches  # replays are counted like launches
    for op in ref.ops:
        assert_clv_equal(ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index), kind != "aa", "clv after replays")
    for e in (ref, gpu, plain):
        e.close()


# ---- the fused Newton-Raphson loop (pll_cuda_newton_branch) --------------------------------------

@pytest.mark.parametrize("kind,tips,sites,attrs", [
    ("dna", 20, 3001, capi.PATTERN_TIP),
    ("dna", 20, 3001, 0),
    ("dna", 60, 2000, capi.SITE_REPEATS),
    ("aa", 12, 800, capi.PATTERN_TIP),
    ("g5", 10, 500, capi.PATTERN_TIP),
    ("dna", 8, 40000, capi.PATTERN_TIP),
    ("dna", 5, 800000, capi.PATTERN_TIP),  # table > 96 MB: the same rule over the streaming derivative kernel
])
def test_fused_newton_matches_the_host_driven_loop(reflib, cudalib, kind, tips, sites, attrs):
    """One cooperative launch per branch against (a) the same rule driven through the reference's
    pll_compute_likelihood_derivatives and (b) through this library's own blocking call."""
    ds = make_ds(kind, tips, sites, "random", seed=21)
    ref, gpu = pair(reflib, cudalib, ds, attrs)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    last = ds.tree.ops[-1]
    edges = [ds.tree.root_edge, (int(last[0]), int(last[2]), int(last[3])), (int(last[0]), int(last[5]), int(last[6]))]
    st_ref, st_gpu = ref.sumtable_alloc(), gpu.sumtable_alloc()
    for edge in edges:
        ref.update_sumtable(st_ref, edge)
        gpu.update_sumtable(st_gpu, edge)
        for t0 in (0.01, 0.1, 0.7):
            want = ref.newton_host(st_ref, t0, edge)
            own = gpu.newton_host(st_gpu, t0, edge)
            got = gpu.newton(st_gpu, t0, edge)
            for other in (want, own):
                assert abs(got[3] - other[3]) <= 1, (got, other)
                if got[3] == other[3]:
                    assert abs(got[0] - other[0]) <= 1e-9 * max(abs(other[0]), 1e-3), (edge, t0, got, other)
                    # near the optimum d_f is a cancellation sum: it moves by dd_f * (error of the length)
                    assert abs(got[1] - other[1]) <= DERIV_RTOL * max(abs(other[1]), 1e-6 * ds.sites,
                                                                      abs(other[2] * other[0])), (got, other)
                    assert abs(got[2] - other[2]) <= DERIV_RTOL * max(abs(other[2]), 1e-6 * ds.sites), (got, other)
                else:
                    assert abs(got[0] - other[0]) <= 1e-4 * max(abs(other[0]), 1e-3), (edge, t0, got, other)
            assert 1e-8 <= got[0] <= 100.0
    # bounds and iteration limit
    t1 = gpu.newton(st_gpu, 0.1, edges[0], max_iters=1)
    d = gpu.derivatives(st_gpu, 0.1, edges[0])
    assert t1[3] == 1 and abs(t1[1] - d[0]) <= DERIV_RTOL * max(abs(d[0]), 1e-6 * ds.sites)
    lo = gpu.newton(st_gpu, 0.5, edges[0], tmin=0.4, tmax=0.6)
    assert 0.4 <= lo[0] <= 0.6
    ref.close()
    gpu.close()


def test_fused_newton_rejects_bad_arguments(cudalib):
    ds = make_ds("dna", 6, 200, "random", seed=3)
    gpu = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    gpu.update_pmatrices()
    gpu.update_partials()
    st = gpu.sumtable_alloc()
    gpu.update_sumtable(st)
    with pytest.raises(RuntimeError):
        gpu.newton(st, 0.1, max_iters=0)
    with pytest.raises(RuntimeError):
        gpu.newton(st, 0.1, tmin=1.0, tmax=0.5)
    assert gpu.newton(st, 0.1)[3] >= 1
    gpu.close()


@pytest.mark.parametrize("kind,attrs", [("dna", capi.PATTERN_TIP), ("dna", 0), ("dna", capi.SITE_REPEATS),
                                        ("aa", capi.PATTERN_TIP), ("aa", 0)])
def test_illegal_tip_character_fails_and_leaves_the_partition_intact(reflib, cudalib, kind, attrs):
    """pll_set_tip_states with a character the map does not know (src/pll.c:900-910): PLL_FAILURE,
    PLL_ERROR_TIPDATA_ILLEGALSTATE and the reference's message; the tip keeps its previous states."""
    ds = make_ds(kind, 9, 203, "random", seed=5)
    ref, gpu = pair(reflib, cudalib, ds, attrs)
    want = ref.full_traversal()
    assert_rel(gpu.full_traversal(), want, LOGL_RTOL, "logL before")
    for pos in (0, 7, 8, 100, 202):
        bad = bytearray(ds.seqs[3])
        bad[pos] = ord("!")
        bad[(pos + 50) % 203] = ord("#")
        msgs = []
        for eng in (ref, gpu):
            rc = eng.lib.pll_set_tip_states(eng.p, 3, eng.map, bytes(bad))
            assert rc == 0 and eng.lib.errno == 114
            msgs.append(eng.lib.errmsg)
        assert msgs[0] == msgs[1]
    assert_rel(gpu.full_traversal(), want, LOGL_RTOL, "logL after the failed calls")
    assert gpu.lib.pll_set_tip_states(gpu.p, 3, gpu.map, ds.seqs[3]) == 1
    assert_rel(gpu.full_traversal(), want, LOGL_RTOL, "logL after re-setting the tip")
    ref.close()
    gpu.close()


@pytest.mark.parametrize("cats,per_rate,tips,sites,tree", [
    (1, False, 40, 700, "random"), (2, False, 40, 700, "random"), (8, False, 40, 700, "random"),
    (16, True, 30, 300, "random"), (32, False, 24, 260, "random"), (3, False, 30, 400, "random"),
    (2, False, 260, 96, "caterpillar"), (8, True, 260, 96, "caterpillar"),
])
def test_site_repeats_rate_counts(reflib, cudalib, cats, per_rate, tips, sites, tree):
    """Every rate-count instantiation of the site-repeat CLV kernel (1..32 categories, the shared-memory matrix
    tables grow with it), a count that is not a power of two (generic kernel), deep trees that scale."""
    brlen = (0.002, 0.05) if tree == "random" else (0.02, 0.22)
    ds = synth.dna_dataset(tips, sites, seed=41 + cats, cats=cats, tree_kind=tree, alpha=0.3, brlen=brlen)
    ref, gpu = pair(reflib, cudalib, ds, capi.SITE_REPEATS, per_rate)
    for e in (ref, gpu):
        e.update_pmatrices()
        e.update_partials()
    n_scaled = 0
    for op in ref.ops:
        assert ref.repeat_ids(op.parent_clv_index)[0] == gpu.repeat_ids(op.parent_clv_index)[0]
        assert_clv_equal(ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index), True, f"clv {op.parent_clv_index}")
        sa, sb = ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index)
        assert np.array_equal(sa, sb), f"scaler {op.parent_scaler_index}"
        n_scaled += int(sa.sum())
    if tree == "caterpillar":
        assert n_scaled > 0
    check_edge_and_derivatives(ref, gpu, ds, per_rate)
    ref.close()
    gpu.close()
