"""Parity cases added in round 2 (all against the UNMODIFIED reference, oracle/_ref, AVX2 attribute, through the
C ABI): virtual cherries, unequal category weights, pll_set_tip_clv, +I under site repeats, BASELINE-shaped
inputs that scale, many live sumtables, and the site-sharded evaluation (two ranks) against the reference on
the same slices.  Tolerances are the north star's: integers and 4-state CLVs bit-exact, logL 1e-10, derivatives 1e-9."""
import ctypes as C
import importlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")

from test_gpu_parity import (DERIV_RTOL, LOGL_RTOL, assert_clv_equal, assert_rel, bits,  # noqa: E402
                             check_edge_and_derivatives, pair)

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def traverse(*engines):
    for e in engines:
        e.update_pmatrices()
        e.update_partials()


def compare_all_nodes(ref, gpu, exact=True):
    n_scaled = 0
    for op in ref.ops:
        assert_clv_equal(ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index), exact, f"clv {op.parent_clv_index}")
        if op.parent_scaler_index >= 0:
            sa, sb = ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index)
            assert np.array_equal(sa, sb), f"scaler {op.parent_scaler_index}"
            n_scaled += int(sa.sum())
    return n_scaled


# ---- virtual cherries ------------------------------------------------------------------------------------

def cherry_nodes(ds):
    tips = ds.tree.tips
    return [int(r[0]) for r in ds.tree.ops if int(r[2]) < tips and int(r[5]) < tips]


CHERRY_CASES = [
    # tips, sites, tree, cats, per_rate
    (12, 97, "random", 4, False),
    (40, 1501, "random", 4, False),
    (40, 1501, "random", 4, True),
    (33, 4099, "random", 2, False),
    (21, 777, "random", 1, False),
    (300, 61, "caterpillar", 4, False),
    (300, 61, "caterpillar", 4, True),
    (64, 16, "random", 4, False),
]


CHERRY_VARIANTS = {
    "default": {"PLF_LEVEL_MAX_SITES": "0", "PLF_FLOW": "0"},       # ring kernel, 256 threads x 1 block per thread
    "items2": {"PLF_CHERRY_ITEMS": "2", "PLF_LEVEL_MAX_SITES": "0", "PLF_FLOW": "0"},   # 128 threads x 2 blocks per thread
    "items4": {"PLF_CHERRY_ITEMS": "4", "PLF_LEVEL_MAX_SITES": "0", "PLF_FLOW": "0"},   # 128-site tiles
    "bulk": {"PLF_CHERRY_BULK": "1", "PLF_LEVEL_MAX_SITES": "0", "PLF_FLOW": "0"},      # write-only consumers through bulk stores
    "stages4-items4": {"PLF_CHERRY_STAGES": "4", "PLF_CHERRY_ITEMS": "4", "PLF_LEVEL_MAX_SITES": "0", "PLF_FLOW": "0"},
    "level": {"PLF_LEVEL_MAX_SITES": "1000000", "PLF_FLOW": "0"},   # one launch per traversal level (k_clv_dna_level)
    "level-written": {"PLF_LEVEL_MAX_SITES": "1000000", "PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW": "0"},
    # the whole traversal as one launch (k_clv_dna_flow): paths of up to 8 ops, of 3, and every parent through memory
    "flow-written": {"PLF_VIRTUAL_CHERRIES": "0"},
    "flow-written-3": {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW_PATH_MAX": "3"},
    "flow-written-1": {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW_PATH_MAX": "1"},
    # two (site, rate) blocks per thread (the default from 2048 sites on)
    "flow-written-u2": {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW_UNROLL": "2"},
    "flow-written-u2-5": {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW_UNROLL": "2", "PLF_FLOW_PATH_MAX": "5"},
}


@pytest.mark.parametrize("case", CHERRY_CASES, ids=lambda c: "-".join(map(str, c)))
@pytest.mark.parametrize("variant", list(CHERRY_VARIANTS))
def test_virtual_cherries_parity(reflib, cudalib, monkeypatch, case, variant):
    """Tip-tip parents are not written: consumers of every kind (tip + cherry, cherry +
    inner, cherry + cherry), scalers, edge logL before anything is materialised, then every CLV bit for bit."""
    tips, sites, tree, cats, per_rate = case
    monkeypatch.setenv("PLF_VIRTUAL_CHERRY_MIN_SITES", "0")
    for k, v in CHERRY_VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    ds = synth.dna_dataset(tips, sites, seed=100 + tips, tree_kind=tree, alpha=0.4, cats=cats)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP, per_rate)
    written = CHERRY_VARIANTS[variant].get("PLF_VIRTUAL_CHERRIES") == "0"
    assert cudalib.pll_cuda_virtual_cherries(gpu.p) == (0 if written else 1)
    cherries = [] if written else cherry_nodes(ds)
    assert written or cherries
    traverse(ref, gpu)
    assert cudalib.pll_cuda_virtual_clvs(gpu.p, 0xFFFFFFFF) == len(cherries)
    # the root edge first: nothing has been materialised, every consumer worked from tip codes
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL over virtual cherries")
    n_scaled = compare_all_nodes(ref, gpu)  # downloads materialise the cherries
    if tree == "caterpillar":
        assert n_scaled > 0
    assert cudalib.pll_cuda_virtual_clvs(gpu.p, 0xFFFFFFFF) == 0
    # three more traversals: the second identical list is captured into a graph, the third replays it
    for _ in range(3):
        traverse(gpu)
    assert cudalib.pll_cuda_virtual_clvs(gpu.p, 0xFFFFFFFF) == len(cherries)
    compare_all_nodes(ref, gpu)
    # edges that end in a cherry: log-likelihood, sumtable, derivatives materialise on their own
    traverse(gpu)
    t = ds.tree
    edges = []
    for r in t.ops:
        for child, m in ((int(r[2]), int(r[3])), (int(r[5]), int(r[6]))):
            if child in cherry_nodes(ds) and len(edges) < 3:
                edges.append((int(r[0]), child, m))
    assert edges
    # the edge is evaluated from the CLVs as the traversal left them (both libraries hold the same state)
    check_edge_and_derivatives(ref, gpu, ds, per_rate, edges)
    ref.close()
    gpu.close()


def test_virtual_cherry_survives_pmatrix_and_tip_changes(reflib, cudalib, monkeypatch):
    """A virtual cherry stands for the CLV the reference would hold: P-matrix updates and new tip states
    after the traversal must not change what a later reader sees; a partial traversal over the consumers
    alone picks the pending cherries up."""
    monkeypatch.setenv("PLF_VIRTUAL_CHERRY_MIN_SITES", "0")
    ds = synth.dna_dataset(24, 1000, seed=7, tree_kind="random", alpha=0.5)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    cherries = cherry_nodes(ds)
    traverse(ref, gpu)
    # new branch lengths everywhere, and new states for the first tip of the first cherry
    bl = ds.tree.branch_lengths[ref.matrix_indices] * 2.5
    first = [r for r in ds.tree.ops if int(r[0]) == cherries[0]][0]
    tip = int(first[2])
    other = ds.seqs[(tip + 1) % ds.tree.tips]
    for e, lib in ((ref, reflib), (gpu, cudalib)):
        e.update_pmatrices(branch_lengths=bl)
        assert lib.pll_set_tip_states(e.p, tip, e.map, other) == 1
    for n in cherries:
        assert np.array_equal(bits(ref.clv(n)), bits(gpu.clv(n))), f"cherry {n} after pmatrix / tip change"
    # consumers only (the cherries are NOT recomputed): the first cherry is real now, the others pending again
    traverse(ref, gpu)
    consumers = [k for k, r in enumerate(ds.tree.ops) if int(r[0]) not in cherries]
    ops = (capi.Operation * len(consumers))(*[ref.ops[k] for k in consumers])
    bl2 = ds.tree.branch_lengths[ref.matrix_indices] * 0.5
    for e, lib in ((ref, reflib), (gpu, cudalib)):
        e.update_pmatrices(branch_lengths=bl2)
        lib.pll_update_partials(e.p, ops, len(consumers))
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL after the consumer-only traversal")
    compare_all_nodes(ref, gpu)
    ref.close()
    gpu.close()


def test_virtual_cherries_off_switch_and_threshold(cudalib, monkeypatch):
    narrow, wide = synth.dna_dataset(10, 200, seed=3), synth.dna_dataset(10, 70000, seed=3, simulate_down_tree=False)
    ds = narrow
    gpu = harness.Engine(cudalib, narrow, capi.ARCH_CUDA | capi.PATTERN_TIP)
    assert cudalib.pll_cuda_virtual_cherries(gpu.p) == 0  # default: launch-bound widths write every parent
    gpu.close()
    gpu = harness.Engine(cudalib, wide, capi.ARCH_CUDA | capi.PATTERN_TIP)
    assert cudalib.pll_cuda_virtual_cherries(gpu.p) == 1
    gpu.close()
    monkeypatch.setenv("PLF_VIRTUAL_CHERRY_MIN_SITES", "0")
    gpu = harness.Engine(cudalib, narrow, capi.ARCH_CUDA | capi.PATTERN_TIP)
    assert cudalib.pll_cuda_virtual_cherries(gpu.p) == 1
    gpu.close()
    monkeypatch.setenv("PLF_VIRTUAL_CHERRIES", "0")
    gpu = harness.Engine(cudalib, wide, capi.ARCH_CUDA | capi.PATTERN_TIP)
    assert cudalib.pll_cuda_virtual_cherries(gpu.p) == 0
    gpu.close()
    monkeypatch.setenv("PLF_VIRTUAL_CHERRIES", "1")
    # partitions the kernels do not serve: tip CLVs instead of pattern tips, odd state counts, ascertainment bias
    for d, extra in ((ds, 0), (synth.generic_dataset(5, 8, 100, seed=4), capi.PATTERN_TIP), (ds, capi.PATTERN_TIP | capi.AB_FLAG)):
        gpu = harness.Engine(cudalib, d, capi.ARCH_CUDA | extra)
        assert cudalib.pll_cuda_virtual_cherries(gpu.p) == 0
        gpu.close()


def test_virtual_cherries_large(reflib, cudalib):
    """default switches, more sites than one sweep of the persistent grids covers"""
    ds = synth.dna_dataset(48, 150_001, seed=9, simulate_down_tree=False)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    assert cudalib.pll_cuda_virtual_cherries(gpu.p) == 1
    traverse(ref, gpu)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL")
    compare_all_nodes(ref, gpu)
    ref.close()
    gpu.close()


AA_CHERRY_CASES = [
    # tips, sites, tree, cats, per_rate
    (10, 53, "random", 4, False),
    (40, 301, "random", 4, False),
    (40, 301, "random", 4, True),
    (120, 21, "caterpillar", 4, False),
    (24, 1003, "random", 2, False),
    (24, 500, "random", 1, False),
    (20, 200, "random", 8, False),
]


@pytest.mark.parametrize("case", AA_CHERRY_CASES, ids=lambda c: "-".join(map(str, c)))
def test_virtual_cherries_parity_aa(reflib, cudalib, monkeypatch, case):
    """20 states: the consumers of a virtual cherry form the cherry's entries in registers as the A operand of
    the DMMA.  Against the reference within the DMMA tolerance (scalers exact), and against this library with
    every tip-tip parent written to HBM: the same operands, so the same bits."""
    tips, sites, tree, cats, per_rate = case
    ds = synth.aa_dataset(tips, sites, seed=200 + tips, tree_kind=tree, alpha=0.4, cats=cats)
    monkeypatch.setenv("PLF_VIRTUAL_CHERRY_MIN_SITES", "0")
    monkeypatch.setenv("PLF_VIRTUAL_CHERRIES", "0")
    monkeypatch.setenv("PLF_AA_TIP_CLV_MAX_SITES", "0")  # tips through the tip kernels, as with virtual cherries
    plain = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP | (capi.RATE_SCALERS if per_rate else 0))
    monkeypatch.setenv("PLF_VIRTUAL_CHERRIES", "1")
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP, per_rate)
    assert cudalib.pll_cuda_virtual_cherries(plain.p) == 0
    # 8 rate categories: the half tables of two cherries do not fit next to a second CTA, every parent is written
    virtual = cats <= 4
    assert cudalib.pll_cuda_virtual_cherries(gpu.p) == int(virtual)
    cherries = cherry_nodes(ds) if virtual else []
    traverse(ref, gpu, plain)
    assert cudalib.pll_cuda_virtual_clvs(gpu.p, 0xFFFFFFFF) == len(cherries)
    assert cudalib.pll_cuda_virtual_clvs(plain.p, 0xFFFFFFFF) == 0
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL over virtual cherries")
    n_scaled = compare_all_nodes(ref, gpu, exact=False)
    if tree == "caterpillar":
        assert n_scaled > 0
    for op in gpu.ops:
        assert np.array_equal(bits(plain.clv(op.parent_clv_index)), bits(gpu.clv(op.parent_clv_index))), op.parent_clv_index
        assert np.array_equal(plain.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index))
    for _ in range(3):
        traverse(gpu)
    compare_all_nodes(ref, gpu, exact=False)
    traverse(gpu)
    check_edge_and_derivatives(ref, gpu, ds, per_rate)
    for e in (ref, gpu, plain):
        e.close()


# ---- unequal category weights (LG4X-style) ------------------------------------------------------------------

WEIGHT_CASES = [
    ("dna", capi.PATTERN_TIP, False), ("dna", 0, False), ("dna", capi.SITE_REPEATS, False), ("dna", capi.PATTERN_TIP, True),
    ("aa", capi.PATTERN_TIP, False), ("aa", 0, False), ("aa", capi.SITE_REPEATS, False), ("g5", 0, False),
]


@pytest.mark.parametrize("kind,extra,per_rate", WEIGHT_CASES)
def test_unequal_category_weights_parity(reflib, cudalib, kind, extra, per_rate):
    """pll_set_category_weights with unequal weights: the reference's reductions take the weighted branch
    (core_likelihood_avx.c:1606-1684 `terma += terma_r * w_r`, core_derivatives_avx2.c:1640-1807)."""
    if kind == "dna":
        ds = synth.dna_dataset(60, 1003, seed=61, tree_kind="caterpillar", alpha=0.4, brlen=(0.01, 0.12))
    elif kind == "aa":
        ds = synth.aa_dataset(40, 301, seed=62, tree_kind="random", alpha=0.4)
    else:
        ds = synth.generic_dataset(5, 30, 200, seed=63)
    ds.cat_weights = np.array([0.1, 0.2, 0.3, 0.4])
    ds.pattern_weights = np.random.default_rng(8).integers(1, 5, size=ds.sites).astype(np.uint32)
    ref, gpu = pair(reflib, cudalib, ds, extra, per_rate)
    w = np.ctypeslib.as_array(gpu.part.rate_weights, shape=(4,))
    assert np.array_equal(w, ds.cat_weights)
    traverse(ref, gpu)
    last = ref.ops[len(ref.ops) - 1]
    edges = [ds.tree.root_edge, (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index)]
    check_edge_and_derivatives(ref, gpu, ds, per_rate, edges)
    if not per_rate:
        r_ref, rp_ref = ref.root_logl(persite=True)
        r_gpu, rp_gpu = gpu.root_logl(persite=True)
        assert_rel(r_gpu, r_ref, LOGL_RTOL, "root logL")
        np.testing.assert_allclose(rp_gpu, rp_ref, rtol=1e-12)
    ref.close()
    gpu.close()


# ---- pll_set_tip_clv (src/pll.c:1066-1129) ------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["dna", "aa", "g5"])
@pytest.mark.parametrize("padding", [0, 1])
@pytest.mark.parametrize("extra", [0, capi.SITE_REPEATS])
def test_set_tip_clv_parity(reflib, cudalib, kind, padding, extra):
    """Tip CLVs given by the caller (uncertain states): replicated over the rates, padded or not, and under
    site repeats taken at the class representatives (src/pll.c:1085-1099)."""
    if kind == "dna":
        ds = synth.dna_dataset(14, 333, seed=91, brlen=(0.002, 0.08))
    elif kind == "aa":
        ds = synth.aa_dataset(9, 120, seed=92, brlen=(0.002, 0.08))
    else:
        ds = synth.generic_dataset(5, 11, 150, seed=93, brlen=(0.002, 0.08))
    ref, gpu = pair(reflib, cudalib, ds, extra)
    st, sp = ds.states, ref.part.states_padded
    width = sp if padding else st
    rng = np.random.default_rng(4)
    # one probability vector per character: consistent with the tip's repeat classes
    per_char = {c: rng.dirichlet(np.ones(st)) for c in range(256)}
    for tip in (0, 3, ds.tree.tips - 1):
        seq = np.frombuffer(ds.seqs[tip], dtype=np.uint8)
        clv = np.zeros((ds.sites, width))
        for s, c in enumerate(seq):
            clv[s, :st] = per_char[int(c)]
        if padding:
            clv[:, st:] = 7.0  # garbage in the padding must not be copied
        flat = np.ascontiguousarray(clv.reshape(-1))
        for lib, e in ((reflib, ref), (cudalib, gpu)):
            assert lib.pll_set_tip_clv(e.p, tip, flat.ctypes.data_as(capi.c_double_p), padding) == 1, lib.errmsg
        a, b = ref.clv(tip), gpu.clv(tip)
        keep = np.tile(np.arange(sp) < st, a.size // sp)
        assert np.array_equal(bits(a[keep]), bits(b[keep])), f"tip clv {tip}"
    traverse(ref, gpu)
    compare_all_nodes(ref, gpu, exact=(kind != "aa"))
    check_edge_and_derivatives(ref, gpu, ds, False)
    ref.close()
    gpu.close()


def test_set_tip_clv_pattern_tip_rejected(cudalib):
    ds = synth.dna_dataset(6, 64, seed=1)
    gpu = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    clv = np.ones(64 * 4)
    assert cudalib.pll_set_tip_clv(gpu.p, 0, clv.ctypes.data_as(capi.c_double_p), 0) == 0
    assert "PLL_ATTRIB_PATTERN_TIP" in cudalib.errmsg
    gpu.close()


# ---- +I together with site repeats --------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["dna", "aa"])
def test_invariant_sites_with_site_repeats(reflib, cudalib, kind):
    ds = (synth.dna_dataset(20, 800, seed=23, brlen=(0.002, 0.05)) if kind == "dna"
          else synth.aa_dataset(12, 300, seed=24, brlen=(0.002, 0.05)))
    seqs = [bytearray(s) for s in ds.seqs]
    for col in range(0, ds.sites, 3):
        for s in seqs:
            s[col] = seqs[0][col]
    ds.seqs = [bytes(s) for s in seqs]
    ds.prop_invar = 0.25
    ds.pattern_weights = np.random.default_rng(3).integers(1, 6, size=ds.sites).astype(np.uint32)
    ref, gpu = pair(reflib, cudalib, ds, capi.SITE_REPEATS)
    inv_ref = np.ctypeslib.as_array(ref.part.invariant, shape=(ds.sites,)).copy()
    inv_gpu = np.ctypeslib.as_array(gpu.part.invariant, shape=(ds.sites,)).copy()
    assert np.array_equal(inv_ref, inv_gpu) and (inv_ref >= 0).sum() > 10
    traverse(ref, gpu)
    last = ref.ops[len(ref.ops) - 1]
    edges = [ds.tree.root_edge, (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index)]
    check_edge_and_derivatives(ref, gpu, ds, False, edges)
    assert_rel(gpu.root_logl(), ref.root_logl(), LOGL_RTOL, "root logL +I, repeats")
    ref.close()
    gpu.close()


# ---- BASELINE-shaped inputs that scale -----------------------------------------------

@pytest.mark.parametrize("extra", [capi.PATTERN_TIP, capi.SITE_REPEATS])
def test_dna_1000_taxa_parity(reflib, cudalib, extra):
    """DNA 1000 taxa x 20k sites (config 4's tree width): scaling events in the thousands, every scaler and
    every CLV of a sample of nodes bit-exact, identifiers exact under repeats, logL and derivatives."""
    ds = synth.dna_dataset(1000, 20_000, seed=3, alpha=0.3, brlen=(0.002, 0.05))
    ref, gpu = pair(reflib, cudalib, ds, extra)
    traverse(ref, gpu)
    n_scaled = 0
    for k, op in enumerate(ref.ops):
        sa, sb = ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index)
        assert np.array_equal(sa, sb), f"scaler {op.parent_scaler_index}"
        n_scaled += int(sa.sum())
        if k % 37 == 0 or k >= len(ref.ops) - 5:
            assert np.array_equal(bits(ref.clv(op.parent_clv_index)), bits(gpu.clv(op.parent_clv_index)))
    assert n_scaled > 0, "a 1000-taxon tree was meant to scale"
    if extra & capi.SITE_REPEATS:
        for node in range(0, ds.tree.nodes, 53):
            ids_r, sid_r, is_r = ref.repeat_ids(node)
            ids_g, sid_g, is_g = gpu.repeat_ids(node)
            assert ids_r == ids_g
            if ids_r:
                assert np.array_equal(sid_r, sid_g) and np.array_equal(is_r, is_g)
    check_edge_and_derivatives(ref, gpu, ds, False)
    ref.close()
    gpu.close()


def test_aa_lg4m_200_taxa_parity(reflib, cudalib):
    """Config 3's model and tree width: LG4M matrices and frequencies per rate category, 200 taxa x 10k sites."""
    ds, name = synth.lg4m_dataset(200, 10_000, seed=2, ref_path=pkg.REF_PATH, brlen=(0.05, 0.6))
    assert name.startswith("LG4M")
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    traverse(ref, gpu)
    n_scaled = 0
    for k, op in enumerate(ref.ops):
        sa, sb = ref.scaler(op.parent_scaler_index), gpu.scaler(op.parent_scaler_index)
        assert np.array_equal(sa, sb), f"scaler {op.parent_scaler_index}"
        n_scaled += int(sa.sum())
        if k % 17 == 0:
            assert_clv_equal(ref.clv(op.parent_clv_index), gpu.clv(op.parent_clv_index), False, f"clv {op.parent_clv_index}")
    assert n_scaled > 0, "the 200-taxon protein tree was meant to scale"
    check_edge_and_derivatives(ref, gpu, ds, False)
    ref.close()
    gpu.close()


# ---- more live sumtables than device slots ------------------------------------------------------------------

def test_many_live_sumtables(reflib, cudalib, monkeypatch):
    """The reference writes a sumtable into the caller's buffer, so any number can be live; here they stay
    in HBM keyed by the caller's pointer.  More tables than slots: every derivative call must either answer
    from the right table or fail loudly -- never from unwritten host bytes."""
    monkeypatch.setenv("PLL_CUDA_MAX_SUMTABLES", "4")
    ds = synth.dna_dataset(16, 500, seed=13)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    traverse(ref, gpu)
    edges = []
    for op in ref.ops:
        for child, m in ((op.child1_clv_index, op.child1_matrix_index), (op.child2_clv_index, op.child2_matrix_index)):
            if child >= ds.tree.tips or op.parent_clv_index >= ds.tree.tips:
                edges.append((op.parent_clv_index, child, m))
    edges = edges[:12]
    tabs_ref = [ref.sumtable_alloc() for _ in edges]
    tabs_gpu = [gpu.sumtable_alloc() for _ in edges]
    for e, tr, tg in zip(edges, tabs_ref, tabs_gpu):
        ref.update_sumtable(tr, e)
        gpu.update_sumtable(tg, e)
    answered = failed = 0
    for e, tr, tg in zip(edges, tabs_ref, tabs_gpu):
        d_ref = ref.derivatives(tr, 0.1, e)
        try:
            d_gpu = gpu.derivatives(tg, 0.1, e)
        except RuntimeError as err:
            assert "sumtable" in str(err)
            failed += 1
            continue
        answered += 1
        for a, b in zip(d_gpu, d_ref):
            assert abs(a - b) <= DERIV_RTOL * max(abs(b), 1e-6 * ds.sites), (e, d_gpu, d_ref)
    assert answered >= 4 and answered + failed == len(edges)
    # recomputing an evicted table brings it back
    gpu.update_sumtable(tabs_gpu[0], edges[0])
    d_ref, d_gpu = ref.derivatives(tabs_ref[0], 0.2, edges[0]), gpu.derivatives(tabs_gpu[0], 0.2, edges[0])
    for a, b in zip(d_gpu, d_ref):
        assert abs(a - b) <= DERIV_RTOL * max(abs(b), 1e-6 * ds.sites)
    ref.close()
    gpu.close()


def test_many_live_sumtables_default_slots(reflib, cudalib):
    """with the default slot count a dozen live tables all answer"""
    ds = synth.dna_dataset(16, 500, seed=13)
    ref, gpu = pair(reflib, cudalib, ds, 0)
    traverse(ref, gpu)
    edges = [(op.parent_clv_index, op.child1_clv_index, op.child1_matrix_index) for op in ref.ops][:12]
    tabs_ref = [ref.sumtable_alloc() for _ in edges]
    tabs_gpu = [gpu.sumtable_alloc() for _ in edges]
    for e, tr, tg in zip(edges, tabs_ref, tabs_gpu):
        ref.update_sumtable(tr, e)
        gpu.update_sumtable(tg, e)
    for e, tr, tg in zip(edges, tabs_ref, tabs_gpu):
        d_ref, d_gpu = ref.derivatives(tr, 0.1, e), gpu.derivatives(tg, 0.1, e)
        for a, b in zip(d_gpu, d_ref):
            assert abs(a - b) <= DERIV_RTOL * max(abs(b), 1e-6 * ds.sites), (e, d_gpu, d_ref)
    ref.close()
    gpu.close()


# ---- site-sharded evaluation, two ranks, against the reference on the same slices ---------------------------

def test_site_sharded_two_ranks_against_reference(reflib, cudalib, tmp_path):
    """Two ranks, each owning a contiguous site slice on the CUDA engine (two GPUs over NCCL when the box has
    them, else both ranks on cuda:0 reducing over gloo): the all-reduced {logL, d_f, dd_f} against the
    reference evaluated on the same two slices and against the reference on the whole alignment."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "sharded.npy")
    env = dict(os.environ, SHARDED_OUT=out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(REPO, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)  # [gpu reduced x3, reference slices summed x3, reference whole x3, peer-memory sum x3]
    gpu3, ref_slices, ref_whole, peer3 = got[0:3], got[3:6], got[6:9], got[9:12]
    if not np.isnan(peer3).any():
        # two GPUs: the peer-memory exchange adds the same two numbers as NCCL does
        assert np.array_equal(peer3, gpu3), (peer3, gpu3)
    assert_rel(gpu3[0], ref_slices[0], LOGL_RTOL, "reduced logL vs reference on the same slices")
    assert_rel(gpu3[0], ref_whole[0], LOGL_RTOL, "reduced logL vs reference on the whole alignment")
    for k, name in ((1, "d_f"), (2, "dd_f")):
        assert abs(gpu3[k] - ref_slices[k]) <= DERIV_RTOL * max(abs(ref_slices[k]), 1e-3), (name, got)
        assert abs(gpu3[k] - ref_whole[k]) <= DERIV_RTOL * max(abs(ref_whole[k]), 1e-3), (name, got)


# ---- pattern-tip codes formed on the device -----------------------------------------------------------------

@pytest.mark.parametrize("kind,sites", [("dna", 1), ("dna", 15), ("dna", 16), ("dna", 4099), ("aa", 333), ("g5", 77)])
@pytest.mark.parametrize("host_map", ["0", "1"])
def test_pattern_tip_codes_match_reference(reflib, cudalib, monkeypatch, kind, sites, host_map):
    """pll_set_tip_states under PLL_ATTRIB_PATTERN_TIP (src/pll.c:875-957): the codes the device forms from the
    raw characters (and the round-1 host loop, PLL_CUDA_TIP_HOST_MAP=1) against the reference's tipchars[]."""
    monkeypatch.setenv("PLL_CUDA_TIP_HOST_MAP", host_map)
    if kind == "dna":
        ds = synth.dna_dataset(7, sites, seed=5)
    elif kind == "aa":
        ds = synth.aa_dataset(7, sites, seed=6)
    else:
        ds = synth.generic_dataset(5, 7, sites, seed=7)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    for t in range(ds.tree.tips):
        assert np.array_equal(ref.tipchars(t), gpu.tipchars(t)), f"tip {t}"
    assert ref.part.maxstates == gpu.part.maxstates
    if sites > 1:
        traverse(ref, gpu)
        assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL")
    # an illegal character: same failure, same message, the tip keeps its codes
    bad = bytearray(ds.seqs[2])
    bad[sites // 2] = ord("!")
    before = gpu.tipchars(2)
    for lib, e in ((reflib, ref), (cudalib, gpu)):
        assert lib.pll_set_tip_states(e.p, 2, e.map, bytes(bad)) == 0
    assert cudalib.errno == reflib.errno and cudalib.errmsg == reflib.errmsg
    assert np.array_equal(gpu.tipchars(2), before)
    ref.close()
    gpu.close()


# ---- repeated identifier updates: captured into a graph and replayed ----------------------------------------

def test_repeated_identifier_updates_replay_a_graph(reflib, cudalib):
    """pll_update_partials with identifier update on the same list again and again (what a tree search does
    between topology moves): the second call is captured, later ones replay the graph; a tip whose states change
    alters the jobs and starts over.  Identifiers, CLVs and scalers against the reference after every stage."""
    ds = synth.dna_dataset(64, 3000, seed=31, alpha=0.3, brlen=(0.002, 0.05))
    ref, gpu = pair(reflib, cudalib, ds, capi.SITE_REPEATS)

    def check(what):
        for node in range(ds.tree.nodes):
            ids_r, sid_r, is_r = ref.repeat_ids(node)
            ids_g, sid_g, is_g = gpu.repeat_ids(node)
            assert ids_r == ids_g, f"{what}: class count of node {node}"
            if ids_r:
                assert np.array_equal(sid_r, sid_g) and np.array_equal(is_r, is_g), f"{what}: identifiers of node {node}"
        compare_all_nodes(ref, gpu)
        assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, what)

    traverse(ref)
    for k in range(3):  # nothing changes between these: parents keep their identifiers
        traverse(gpu)
        check(f"traversal {k}")
    for k in range(4):  # every parent renumbered: plain launches, graph capture, two replays
        assert cudalib.pll_cuda_invalidate_repeat_identifiers(gpu.p) == 1
        traverse(gpu)
        check(f"renumbered traversal {k}")
    other = ds.seqs[5]
    for lib, e in ((reflib, ref), (cudalib, gpu)):
        assert lib.pll_set_tip_states(e.p, 9, e.map, other) == 1
    traverse(ref)
    for k in range(3):  # the first renumbers the path from tip 9 to the root only
        traverse(gpu)
        check(f"after the tip change, traversal {k}")
    # the per-op entry point on a node in the middle of the tree, with OTHER children than the list's: the list's
    # next update must notice that the node (and everything above it) no longer holds its identifiers
    mid = ref.ops[len(ref.ops) // 2]
    odd = capi.Operation(mid.parent_clv_index, mid.parent_scaler_index, 0, 0, -1, 1, 1, -1)
    cudalib.pll_update_repeats(gpu.p, C.byref(odd))
    traverse(gpu)
    check("after pll_update_repeats on a node of the list")
    ref.close()
    gpu.close()


# ---- the per-kind ring / bulk kernels on the small shapes that the level kernel now takes by default ---------

from test_gpu_parity import CASES as R1_CASES  # noqa: E402


@pytest.mark.parametrize("case", [c for c in R1_CASES if c[0] == "dna" and c[4] & capi.PATTERN_TIP],
                         ids=lambda c: "-".join(map(str, c)))
@pytest.mark.parametrize("cherries", ["0", "1", "level"], ids=["written", "virtual", "per-level"])
def test_per_kind_kernels_small_shapes(reflib, cudalib, monkeypatch, case, cherries):
    """Alignments up to 2048 sites run as one launch per traversal by default (k_clv_dna_flow); the streaming
    per-kind kernels (ring copies, bulk stores: PLF_LEVEL_MAX_SITES=0) and the one-launch-per-level kernel
    (PLF_FLOW=0) keep their small-shape coverage here."""
    kind, tips, sites, tree, extra, per_rate = case
    if cherries == "level":
        monkeypatch.setenv("PLF_FLOW", "0")
        cherries = "0"
    else:
        monkeypatch.setenv("PLF_LEVEL_MAX_SITES", "0")
        monkeypatch.setenv("PLF_FLOW", "0")
    monkeypatch.setenv("PLF_VIRTUAL_CHERRIES", cherries)
    monkeypatch.setenv("PLF_VIRTUAL_CHERRY_MIN_SITES", "0")
    ds = synth.dna_dataset(tips, sites, seed=11, tree_kind=tree, alpha=0.4)
    ref, gpu = pair(reflib, cudalib, ds, extra, per_rate)
    traverse(ref, gpu)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL")
    n_scaled = compare_all_nodes(ref, gpu)
    if tree == "caterpillar":
        assert n_scaled > 0
    for _ in range(3):
        traverse(gpu)
    compare_all_nodes(ref, gpu)
    ref.close()
    gpu.close()


# ---- the whole traversal as one launch: dependencies between work items -----------------------------------

def test_flow_kernel_lists(reflib, cudalib, monkeypatch):
    """k_clv_dna_flow orders work items by who writes what an op reads (plf_op_t.dep).  Lists that put that to the
    test: the full traversal in its given order and reordered by level (subtrees interleaved); partial lists whose
    children are older than the list; a list that recycles a CLV buffer (keeps the launch levels); repeated replays."""
    ds = synth.dna_dataset(40, 777, seed=21, tree_kind="random", alpha=0.4)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP, False)
    traverse(ref, gpu)
    l0 = cudalib.pll_cuda_kernel_launches()
    gpu.update_partials()
    assert cudalib.pll_cuda_kernel_launches() - l0 == 1, "a narrow plain traversal is one launch"
    compare_all_nodes(ref, gpu)
    ops = list(gpu.ops)
    n = len(ops)
    # a valid reordering: stable sort by level keeps producers before consumers but interleaves subtrees
    level = {}
    for op in ops:
        level[op.parent_clv_index] = 1 + max(level.get(op.child1_clv_index, 0), level.get(op.child2_clv_index, 0))
    order = sorted(range(n), key=lambda i: (level[ops[i].parent_clv_index], -i))
    arr = (capi.Operation * n)(*[ops[i] for i in order])
    for e in (ref, gpu):
        e.update_pmatrices()
        e.lib.pll_update_partials(e.p, arr, n)
    compare_all_nodes(ref, gpu)
    # partial lists: the upper half of the list reads CLVs the lower half left behind; new branch lengths first
    bl = ds.tree.branch_lengths * 1.3
    for e in (ref, gpu):
        e.update_pmatrices(branch_lengths=bl[e.matrix_indices])
        half = (capi.Operation * (n - n // 2))(*ops[n // 2:])
        for _ in range(4):  # plain, captured, replayed twice
            e.lib.pll_update_partials(e.p, half, n - n // 2)
    compare_all_nodes(ref, gpu)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL after the partial lists")
    # a list that recycles a buffer: the last op again, its result then overwritten by an op with other children
    a, b = ops[-1], ops[0]
    recycled = capi.Operation(a.parent_clv_index, a.parent_scaler_index, b.child1_clv_index, b.child1_matrix_index,
                              b.child1_scaler_index, b.child2_clv_index, b.child2_matrix_index, b.child2_scaler_index)
    lst = (capi.Operation * 3)(a, recycled, a)
    for e in (ref, gpu):
        for _ in range(3):
            e.lib.pll_update_partials(e.p, lst, 3)
    compare_all_nodes(ref, gpu)
    ref.close()
    gpu.close()


@pytest.mark.parametrize("sites", [1, 31, 32, 33, 64, 65, 513, 2048])
@pytest.mark.parametrize("cats", [1, 2, 4])
@pytest.mark.parametrize("path_max", ["8", "2", "u2"])
def test_flow_kernel_chunk_edges(reflib, cudalib, monkeypatch, sites, cats, path_max):
    """work-item boundaries of k_clv_dna_flow (32, 64 or 128 sites per sweep for 4, 2, 1 rate categories), per-rate
    scalers on a caterpillar that scales, tip CLVs instead of pattern tips for the odd widths"""
    if path_max.startswith("u"):
        monkeypatch.setenv("PLF_FLOW_UNROLL", path_max[1:])
    else:
        monkeypatch.setenv("PLF_FLOW_PATH_MAX", path_max)
    attrs = capi.PATTERN_TIP if sites % 2 else 0
    ds = synth.dna_dataset(400, sites, seed=sites + cats, tree_kind="caterpillar", alpha=0.3, cats=cats)
    ref, gpu = pair(reflib, cudalib, ds, attrs, cats == 4)
    for _ in range(3):
        traverse(gpu)
    traverse(ref)
    l0 = cudalib.pll_cuda_kernel_launches()
    gpu.update_partials()
    assert cudalib.pll_cuda_kernel_launches() - l0 == 1
    n_scaled = compare_all_nodes(ref, gpu)
    assert n_scaled > 0 or sites < 3
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL")
    ref.close()
    gpu.close()


@pytest.mark.parametrize("per_rate", [False, True], ids=["per-site", "per-rate"])
def test_plain_lists_under_site_repeats_run_as_one_launch(reflib, cudalib, per_rate):
    """PLL_ATTRIB_SITE_REPEATS on columns that do not repeat: above the tips the default rule (src/repeats.c:100-111)
    leaves the nodes without identifiers, so a list of upper operations is a plain list and takes k_clv_dna_flow -
    with the dependencies between its operations, not without them."""
    ds = synth.dna_dataset(300, 45, seed=77, alpha=0.6, simulate_down_tree=False)
    ref, gpu = pair(reflib, cudalib, ds, capi.SITE_REPEATS, per_rate)
    traverse(ref, gpu)
    ids = gpu.part.repeats.contents.pernode_ids
    tips = ds.tree.tips
    upper = [op for op in gpu.ops if op.child1_clv_index >= tips and op.child2_clv_index >= tips and
             not ids[op.child1_clv_index] and not ids[op.child2_clv_index] and not ids[op.parent_clv_index]]
    assert len(upper) >= 8, "the data is meant to leave the upper nodes without identifiers"
    compare_all_nodes(ref, gpu)
    n = len(upper)
    arr = (capi.Operation * n)(*upper)
    levels = np.zeros(n, dtype=np.uint32)
    assert cudalib.pll_cuda_schedule_levels(arr, n, levels.ctypes.data_as(capi.c_uint_p)) >= 3, "operations that depend on each other"
    bl = ds.tree.branch_lengths * 0.7
    for e in (ref, gpu):
        e.update_pmatrices(branch_lengths=bl[e.matrix_indices])
    ref.lib.pll_update_partials(ref.p, arr, n)
    for _ in range(4):  # plain launch, capture, two replays
        l0 = cudalib.pll_cuda_kernel_launches()
        cudalib.pll_update_partials(gpu.p, arr, n)
        assert cudalib.pll_cuda_kernel_launches() - l0 == 1, "a plain list is one launch"
    compare_all_nodes(ref, gpu)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL")
    ref.close()
    gpu.close()


# ---- 20 states, narrow alignments: pattern tips that meet inner nodes as expanded CLVs -------------------------

@pytest.mark.parametrize("per_rate", [False, True], ids=["per-site", "per-rate"])
def test_aa_narrow_tips_as_expanded_clvs(reflib, cudalib, monkeypatch, per_rate):
    """Up to 2048 sites a 20-state pattern tip under a tip + inner operation is read as an expanded CLV (one
    inner-inner launch per level instead of up to three kinds; tip + tip keeps its kernel).  Same values within the
    tensor-core tolerance, same scalers, also after a tip's sequence changed; PLF_AA_TIP_CLV_MAX_SITES=0 turns it off."""
    ds = synth.aa_dataset(80, 333, seed=31, alpha=0.4)
    ref, gpu = pair(reflib, cudalib, ds, capi.PATTERN_TIP, per_rate)
    monkeypatch.setenv("PLF_AA_TIP_CLV_MAX_SITES", "0")
    old = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP | (capi.RATE_SCALERS if per_rate else 0))
    traverse(ref, gpu, old)
    l0 = cudalib.pll_cuda_kernel_launches()
    gpu.update_partials()
    l1 = cudalib.pll_cuda_kernel_launches()
    old.update_partials()
    l2 = cudalib.pll_cuda_kernel_launches()
    assert l1 - l0 < l2 - l1, "fewer launches per traversal"
    compare_all_nodes(ref, gpu, exact=False)
    compare_all_nodes(ref, old, exact=False)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL")
    # another sequence for two tips (one of a cherry, one under a tip + inner operation)
    for tip in (0, ds.tree.tips - 1):
        seq = ds.seqs[(tip + 5) % ds.tree.tips]
        for e in (ref, gpu):
            assert e.lib.pll_set_tip_states(e.p, tip, e.map, seq) == 1
    for _ in range(3):
        traverse(gpu)
    traverse(ref)
    n_scaled = compare_all_nodes(ref, gpu, exact=False)
    assert_rel(gpu.edge_logl(), ref.edge_logl(), LOGL_RTOL, "edge logL after new tip states")
    assert n_scaled >= 0
    for e in (ref, gpu, old):
        e.close()
