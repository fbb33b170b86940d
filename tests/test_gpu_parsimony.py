"""Fitch parsimony on the GPU (libpll-2_b200/csrc/plf_parsimony.cu + pll_parsimony.c) against
the UNMODIFIED reference (src/fast_parsimony.c, src/stepwise.c in oracle/_ref) on identical seeded inputs.

Everything is integer work: informative flags, constant cost, packed tip vectors, every updated vector, node
costs, edge and root scores, and the stepwise-addition tree and its cost must be IDENTICAL.  The reference runs
with PLL_ATTRIB_ARCH_CPU so that its vectors are not padded to a SIMD width (the padding words are all ones and
never change a score); padded AVX2 vectors are compared on their common prefix in one case.
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


def make_ds(kind, tips, sites, seed, weights=False):
    if kind == "dna":
        ds = synth.dna_dataset(tips, sites, seed=seed, alpha=0.5, weights=weights, brlen=(0.05, 0.4))
    elif kind == "aa":
        ds = synth.aa_dataset(tips, sites, seed=seed, alpha=0.5, brlen=(0.05, 0.4))
        if weights:
            ds.pattern_weights = np.random.default_rng(seed).integers(1, 5, size=sites).astype(np.uint32)
    else:
        ds = synth.generic_dataset(int(kind[1:]), tips, sites, seed=seed, brlen=(0.05, 0.4))
    return ds


class Pars:
    """pll_fastparsimony_init on an Engine's partition, with readers that work for both libraries"""

    def __init__(self, lib, eng):
        self.lib, self.eng = lib, eng
        self.p = lib.pll_fastparsimony_init(eng.p)
        assert self.p, f"pll_fastparsimony_init failed: {lib.errno} {lib.errmsg}"
        self.c = self.p.contents
        self.nodes = self.c.tips + 3 * self.c.inner_nodes

    def vector(self, i):
        n = self.c.states * self.c.packedvector_count
        if self.lib.is_cuda:
            out = np.empty(n, dtype=np.uint32)
            assert self.lib.pll_cuda_download_parsimony_vector(self.p, i, out.ctypes.data_as(capi.c_uint_p)) == 1
            return out.reshape(self.c.states, -1)
        return np.ctypeslib.as_array(self.c.packedvector[i], shape=(n,)).copy().reshape(self.c.states, -1)

    def informative(self):
        return np.ctypeslib.as_array(self.c.informative, shape=(self.c.sites,)).copy()

    def costs(self):
        return np.ctypeslib.as_array(self.c.node_cost, shape=(self.nodes,)).copy()

    def update(self, triples):
        ops = (capi.ParsBuildOp * len(triples))(*[capi.ParsBuildOp(*[int(x) for x in t]) for t in triples])
        self.lib.pll_fastparsimony_update_vectors(self.p, ops, len(triples))

    def close(self):
        if self.p:
            self.lib.pll_parsimony_destroy(self.p)
            self.p = None


def pars_pair(reflib, cudalib, ds, flags, ref_arch=capi.ARCH_CPU):
    ref_eng = harness.Engine(reflib, ds, ref_arch | flags)
    gpu_eng = harness.Engine(cudalib, ds, capi.ARCH_CUDA | flags)
    return Pars(reflib, ref_eng), Pars(cudalib, gpu_eng)


def tree_triples(ds):
    return [(int(r[0]), int(r[2]), int(r[5])) for r in ds.tree.ops]


CASES = [
    # kind, tips, sites, attrs, weights
    ("dna", 8, 1, capi.PATTERN_TIP, False),
    ("dna", 8, 31, capi.PATTERN_TIP, False),
    ("dna", 8, 32, capi.PATTERN_TIP, False),
    ("dna", 9, 33, capi.PATTERN_TIP, True),
    ("dna", 30, 2501, capi.PATTERN_TIP, False),
    ("dna", 30, 2501, capi.PATTERN_TIP, True),
    ("dna", 30, 2501, 0, True),
    ("dna", 30, 2501, capi.SITE_REPEATS, False),
    ("dna", 150, 4099, capi.PATTERN_TIP, False),
    ("aa", 25, 1201, capi.PATTERN_TIP, False),
    ("aa", 25, 1201, capi.PATTERN_TIP, True),
    ("aa", 25, 601, 0, False),  # > 8 states from tip CLVs: the O(tips^2) informative kernel
    ("aa", 25, 601, capi.SITE_REPEATS, False),
    ("g5", 20, 777, capi.PATTERN_TIP, False),
    ("g7", 20, 777, 0, False),
]


@pytest.mark.parametrize("kind,tips,sites,attrs,weights", CASES)
def test_parsimony_matches_reference(reflib, cudalib, kind, tips, sites, attrs, weights):
    ds = make_ds(kind, tips, sites, seed=tips + sites, weights=weights)
    ref, gpu = pars_pair(reflib, cudalib, ds, attrs)
    try:
        for f in ("tips", "inner_nodes", "sites", "states", "packedvector_count", "const_cost", "informative_count"):
            assert getattr(gpu.c, f) == getattr(ref.c, f), f
        np.testing.assert_array_equal(gpu.informative(), ref.informative())
        for t in range(tips):
            np.testing.assert_array_equal(gpu.vector(t), ref.vector(t), err_msg=f"tip vector {t}")
        triples = tree_triples(ds)
        ref.update(triples)
        gpu.update(triples)
        np.testing.assert_array_equal(gpu.costs(), ref.costs())
        for parent, _, _ in triples:
            np.testing.assert_array_equal(gpu.vector(parent), ref.vector(parent), err_msg=f"vector {parent}")
        a, b, _ = ds.tree.root_edge
        assert cudalib.pll_fastparsimony_edge_score(gpu.p, a, b) == reflib.pll_fastparsimony_edge_score(ref.p, a, b)
        assert cudalib.pll_fastparsimony_root_score(gpu.p, a) == reflib.pll_fastparsimony_root_score(ref.p, a)
        # arbitrary pairs, singly and as one batch
        rng = np.random.default_rng(3)
        pairs = rng.integers(0, ds.tree.nodes, size=(40, 2)).astype(np.uint32)
        want = [reflib.pll_fastparsimony_edge_score(ref.p, int(x), int(y)) for x, y in pairs]
        got = np.zeros(len(pairs), dtype=np.uint32)
        assert cudalib.pll_cuda_fastparsimony_edge_scores(
            gpu.p, pairs.ctypes.data_as(capi.c_uint_p), len(pairs), got.ctypes.data_as(capi.c_uint_p)) == 1
        assert got.tolist() == want
        assert [cudalib.pll_fastparsimony_edge_score(gpu.p, int(x), int(y)) for x, y in pairs[:5]] == want[:5]
    finally:
        ref.close()
        gpu.close()


def test_partial_update_lists_and_recycled_vectors(reflib, cudalib):
    """ops applied in several calls, and a list that overwrites a vector it read earlier (strict list order)"""
    ds = make_ds("dna", 40, 3001, seed=5)
    ref, gpu = pars_pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    try:
        triples = tree_triples(ds)
        for lo in range(0, len(triples), 7):
            ref.update(triples[lo:lo + 7])
            gpu.update(triples[lo:lo + 7])
        np.testing.assert_array_equal(gpu.costs(), ref.costs())
        top = ds.tree.nodes  # first unused directional vector
        chain = [(top, 0, 1), (top + 1, top, 2), (top, top + 1, 3), (top + 1, top, top + 1), (top + 2, top + 1, top)]
        ref.update(chain)
        gpu.update(chain)
        np.testing.assert_array_equal(gpu.costs(), ref.costs())
        for v in (top, top + 1, top + 2):
            np.testing.assert_array_equal(gpu.vector(v), ref.vector(v))
    finally:
        ref.close()
        gpu.close()


@pytest.mark.parametrize("mode", ["0", "2"])
def test_level_scheduled_and_chained_updates_agree_with_the_reference(reflib, cudalib, monkeypatch, mode):
    """PLF_PARS_LEVELS=2 forces one launch per level, =0 the one-launch chain.  Random operation lists over a
    small pool of vectors: children that are rewritten later, parents that were read before, in-place updates -
    every read-after-write, write-after-read and write-after-write order of the sequential list must hold."""
    monkeypatch.setenv("PLF_PARS_LEVELS", mode)
    ds = make_ds("dna", 12, 1999, seed=9)
    ref, gpu = pars_pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    rng = np.random.default_rng(int(mode) + 1)
    try:
        pool = list(range(12, 24))  # twelve writable vectors on top of the twelve tips
        init = [(v, v - 12, (v - 11) % 12) for v in pool]  # the reference's inner vectors start uninitialised
        ref.update(init)
        gpu.update(init)
        for count in (1, 3, 8, 60, 400):
            triples = []
            for _ in range(count):
                p_ = int(rng.choice(pool))
                a, b = (int(x) for x in rng.integers(0, 24, size=2))
                triples.append((p_, a, b))
            ref.update(triples)
            gpu.update(triples)
            np.testing.assert_array_equal(gpu.costs(), ref.costs(), err_msg=f"count {count}")
            for v in pool:
                np.testing.assert_array_equal(gpu.vector(v), ref.vector(v), err_msg=f"vector {v} after {count} ops")
        triples = tree_triples(ds)
        ref.update(triples)
        gpu.update(triples)
        np.testing.assert_array_equal(gpu.costs(), ref.costs())
    finally:
        ref.close()
        gpu.close()


def test_simd_padded_reference_vectors_agree_on_the_common_prefix(reflib, cudalib):
    ds = make_ds("dna", 20, 1000, seed=8)
    ref, gpu = pars_pair(reflib, cudalib, ds, capi.PATTERN_TIP, ref_arch=capi.ARCH_AVX2)
    try:
        w = gpu.c.packedvector_count
        assert ref.c.packedvector_count >= w
        triples = tree_triples(ds)
        ref.update(triples)
        gpu.update(triples)
        np.testing.assert_array_equal(gpu.costs(), ref.costs())
        for v in list(range(20)) + [t[0] for t in triples]:
            rv = ref.vector(v)
            np.testing.assert_array_equal(gpu.vector(v), rv[:, :w])
            assert (rv[:, w:] == 0xFFFFFFFF).all()
    finally:
        ref.close()
        gpu.close()


def test_parsimony_outlives_its_partition_and_rejects_bad_indices(cudalib):
    ds = make_ds("dna", 10, 500, seed=2)
    eng = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    p = Pars(cudalib, eng)
    eng.close()
    triples = tree_triples(ds)
    p.update(triples)
    a, b, _ = ds.tree.root_edge
    s1 = cudalib.pll_fastparsimony_edge_score(p.p, a, b)
    assert s1 == cudalib.pll_fastparsimony_root_score(p.p, a) or s1 > 0
    before = p.costs()
    p.update([(p.nodes, 0, 1)])
    assert cudalib.errno == 113  # PLL_ERROR_PARAM_INVALID, nothing launched
    np.testing.assert_array_equal(p.costs(), before)
    p.close()


def test_more_than_twenty_states_need_pattern_tips(cudalib):
    ds = synth.generic_dataset(22, 6, 50, seed=4)
    eng = harness.Engine(cudalib, ds, capi.ARCH_CUDA)
    assert not cudalib.pll_fastparsimony_init(eng.p)
    assert cudalib.errno == 129  # PLL_ERROR_STEPWISE_UNSUPPORTED (src/fast_parsimony.c:538-547)
    eng.close()


# ---- weighted (Sankoff) parsimony --------------------------------------------------------------------------

def weighted_pair(reflib, cudalib, ds, matrix):
    out = []
    t = ds.tree
    maps = []
    for lib in (reflib, cudalib):
        if ds.map_name.startswith("custom"):
            m = synth.custom_map(ds.states)
            mp = m.ctypes.data_as(C.POINTER(capi.pll_state_t))
            maps.append((m, mp))
        else:
            maps.append((None, lib.map(ds.map_name)))
        p = lib.pll_parsimony_create(t.tips, ds.states, ds.sites, matrix.ctypes.data_as(capi.c_double_p), t.inner, t.inner)
        assert p, (lib.errno, lib.errmsg)
        for i, seq in enumerate(ds.seqs):
            assert lib.pll_set_parsimony_sequence(p, i, maps[-1][1], seq) == 1
        out.append(p)
    return out, maps


@pytest.mark.parametrize("kind,tips,sites,integer", [("dna", 6, 1, True), ("dna", 25, 2003, True), ("dna", 25, 2003, False),
                                                   ("aa", 18, 700, False), ("g7", 12, 333, True)])
def test_weighted_parsimony_matches_reference(reflib, cudalib, kind, tips, sites, integer):
    """pll_parsimony_create / _build / _score / _reconstruct (src/parsimony.c): score buffers, totals and
    ancestral states identical to the reference; the buffers are read on the host through the struct."""
    ds = make_ds(kind, tips, sites, seed=tips + sites)
    st = ds.states
    rng = np.random.default_rng(sites)
    matrix = rng.integers(1, 5, size=(st, st)).astype(np.float64) if integer else rng.uniform(0.3, 3.0, size=(st, st))
    np.fill_diagonal(matrix, 0.0)
    matrix = np.ascontiguousarray(matrix)
    (ref, gpu), maps = weighted_pair(reflib, cudalib, ds, matrix)
    try:
        t = ds.tree
        triples = tree_triples(ds)
        ops = (capi.ParsBuildOp * len(triples))(*[capi.ParsBuildOp(*x) for x in triples])
        s_ref = reflib.pll_parsimony_build(ref, ops, len(triples))
        s_gpu = cudalib.pll_parsimony_build(gpu, ops, len(triples))
        assert s_gpu == s_ref
        n = ds.sites * st
        for node in range(t.nodes):
            a = np.ctypeslib.as_array(ref.contents.sbuffer[node], shape=(n,))
            b = np.ctypeslib.as_array(gpu.contents.sbuffer[node], shape=(n,))  # managed memory: host-readable
            np.testing.assert_array_equal(a.view(np.uint64), b.view(np.uint64), err_msg=f"score buffer {node}")
        for node in (t.tips, t.nodes - 1):
            assert cudalib.pll_parsimony_score(gpu, node) == reflib.pll_parsimony_score(ref, node)
        # reconstruction in pre-order: reverse of the post-order list; the first entry is the subtree root
        parent_of = {}
        for p_, c1, c2 in triples:
            parent_of[c1] = p_
            parent_of[c2] = p_
        top = triples[-1][0]  # a second subtree root (other end of the root edge) is hung under the first one
        rec = [(p_, p_, parent_of.get(p_, top), parent_of.get(p_, top)) for p_, _, _ in reversed(triples)]
        recops = (capi.ParsRecOp * len(rec))(*[capi.ParsRecOp(*x) for x in rec])
        reflib.pll_parsimony_reconstruct(ref, maps[0][1], recops, len(rec))
        cudalib.pll_parsimony_reconstruct(gpu, maps[1][1], recops, len(rec))
        for node in range(t.tips, t.nodes):
            a = np.ctypeslib.as_array(ref.contents.anc_states[node], shape=(ds.sites,))
            b = np.ctypeslib.as_array(gpu.contents.anc_states[node], shape=(ds.sites,))
            np.testing.assert_array_equal(a, b, err_msg=f"ancestral states {node}")
    finally:
        reflib.pll_parsimony_destroy(ref)
        cudalib.pll_parsimony_destroy(gpu)


def test_weighted_parsimony_rejects_illegal_characters_and_wrong_objects(cudalib, capfd):
    m = np.ones((4, 4))
    p = cudalib.pll_parsimony_create(3, 4, 5, np.ascontiguousarray(m).ctypes.data_as(capi.c_double_p), 1, 1)
    assert cudalib.pll_set_parsimony_sequence(p, 0, cudalib.map("pll_map_nt"), b"AC!GT") == 0
    assert cudalib.errno == 114  # PLL_ERROR_TIPDATA_ILLEGALSTATE
    assert "Illegal state code" in capfd.readouterr().out  # the reference prints the message too
    assert cudalib.pll_set_parsimony_sequence(p, 0, cudalib.map("pll_map_nt"), b"ACNGT") == 1
    ops = (capi.ParsBuildOp * 1)(capi.ParsBuildOp(9, 0, 1))
    assert cudalib.pll_parsimony_build(p, ops, 1) == 0 and cudalib.errno == 113
    cudalib.pll_fastparsimony_update_vectors(p, ops, 1)  # a score-buffer object has no bit vectors: no crash
    cudalib.pll_parsimony_destroy(p)


# ---- stepwise addition ------------------------------------------------------------------------------------

def splits(tree_p):
    """the tree as a set of tip-label bipartitions (independent of node order and rooting)"""
    t = tree_p.contents
    all_labels = set()
    out = set()

    def below(n):
        if not n.next:
            lab = n.label.decode()
            all_labels.add(lab)
            return frozenset([lab])
        s = below(n.next.contents.back.contents) | below(n.next.contents.next.contents.back.contents)
        out.add(s)
        return s

    root = t.vroot.contents
    a = below(root)
    b = below(root.back.contents)
    assert not (a & b)
    everything = frozenset(all_labels)
    canon = set()
    for s in out:
        if 1 < len(s) < len(everything) - 1:
            canon.add(min(s, everything - s, key=lambda x: (len(x), sorted(x))))
    return canon, everything


def run_stepwise(lib, pars_list, labels, seed):
    arr = (capi.ParsimonyP * len(pars_list))(*[p.p for p in pars_list])
    lab = (C.c_char_p * len(labels))(*[x.encode() for x in labels])
    cost = C.c_uint(0)
    tree = lib.pll_fastparsimony_stepwise(arr, lab, C.byref(cost), len(pars_list), seed)
    assert tree, f"stepwise failed: {lib.errno} {lib.errmsg}"
    return tree, cost.value


@pytest.mark.parametrize("kind,tips,sites,attrs,seed", [
    ("dna", 3, 200, capi.PATTERN_TIP, 1),
    ("dna", 4, 200, capi.PATTERN_TIP, 0),
    ("dna", 25, 1500, capi.PATTERN_TIP, 0),
    ("dna", 25, 1500, capi.PATTERN_TIP, 1),
    ("dna", 25, 1500, 0, 42),
    ("dna", 60, 3000, capi.PATTERN_TIP, 7),
    ("dna", 60, 40, capi.PATTERN_TIP, 3),  # few sites: many ties, the first minimal edge must win
    ("aa", 30, 800, capi.PATTERN_TIP, 12345),
    ("g5", 18, 600, capi.PATTERN_TIP, 99),
])
def test_stepwise_addition_builds_the_reference_tree(reflib, cudalib, kind, tips, sites, attrs, seed):
    ds = make_ds(kind, tips, sites, seed=tips * 3 + sites)
    ref, gpu = pars_pair(reflib, cudalib, ds, attrs)
    labels = [f"taxon{i:03d}" for i in range(tips)]
    try:
        t_ref, c_ref = run_stepwise(reflib, [ref], labels, seed)
        t_gpu, c_gpu = run_stepwise(cudalib, [gpu], labels, seed)
        assert c_gpu == c_ref
        s_ref, l_ref = splits(t_ref)
        s_gpu, l_gpu = splits(t_gpu)
        assert l_gpu == l_ref == frozenset(labels)
        assert s_gpu == s_ref
        g = t_gpu.contents
        assert (g.tip_count, g.inner_count, g.edge_count, g.binary) == (tips, tips - 2, 2 * tips - 3, 1)
    finally:
        ref.close()
        gpu.close()


def test_no_informative_site_at_all(reflib, cudalib):
    """constant and singleton columns only: zero-length bit vectors, every placement costs the same and the first
    edge wins; counts, costs and the stepwise tree still equal the reference's"""
    tips, sites = 9, 64
    ds = make_ds("dna", tips, sites, seed=77)
    col = np.frombuffer(b"ACGT" * (sites // 4), dtype=np.uint8).copy()
    seqs = [col.copy() for _ in range(tips)]
    seqs[3][5] = ord("T") if seqs[3][5] != ord("T") else ord("A")  # one singleton
    ds.seqs = [bytes(x) for x in seqs]
    ref, gpu = pars_pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    labels = [f"x{i}" for i in range(tips)]
    try:
        assert gpu.c.informative_count == ref.c.informative_count == 0
        assert gpu.c.packedvector_count == ref.c.packedvector_count == 0
        assert gpu.c.const_cost == ref.c.const_cost == 1
        triples = tree_triples(ds)
        ref.update(triples)
        gpu.update(triples)
        np.testing.assert_array_equal(gpu.costs(), ref.costs())
        a, b, _ = ds.tree.root_edge
        assert cudalib.pll_fastparsimony_edge_score(gpu.p, a, b) == reflib.pll_fastparsimony_edge_score(ref.p, a, b) == 1
        t_ref, c_ref = run_stepwise(reflib, [ref], labels, 4)
        t_gpu, c_gpu = run_stepwise(cudalib, [gpu], labels, 4)
        assert c_gpu == c_ref == 1
        assert splits(t_gpu) == splits(t_ref)
    finally:
        ref.close()
        gpu.close()


def test_stepwise_over_two_partitions(reflib, cudalib):
    tips = 20
    ds1 = make_ds("dna", tips, 900, seed=31)
    ds2 = make_ds("dna", tips, 400, seed=32, weights=True)
    r1, g1 = pars_pair(reflib, cudalib, ds1, capi.PATTERN_TIP)
    r2, g2 = pars_pair(reflib, cudalib, ds2, capi.PATTERN_TIP)
    labels = [f"t{i}" for i in range(tips)]
    try:
        t_ref, c_ref = run_stepwise(reflib, [r1, r2], labels, 5)
        t_gpu, c_gpu = run_stepwise(cudalib, [g1, g2], labels, 5)
        assert c_gpu == c_ref
        assert splits(t_gpu) == splits(t_ref)
    finally:
        for p in (r1, r2, g1, g2):
            p.close()


# ---- extending a tree and SPR rounds: both libraries work on trees parsed by THIS library's Newick reader
# (identical struct layouts), so node arrays, indices and hence the shuffled orders are the same on both sides

def parse_tree(cudalib, newick):
    f = cudalib.lib.pll_utree_parse_newick_string
    f.restype, f.argtypes = C.POINTER(capi.UTree), [C.c_char_p]
    t = f(newick.encode())
    assert t, cudalib.errmsg
    return t


def export(cudalib, tree_p):
    f = cudalib.lib.pll_utree_export_newick
    f.restype, f.argtypes = C.c_void_p, [C.POINTER(capi.UNode), C.c_void_p]
    t = tree_p.contents
    raw = f(t.nodes[t.tip_count + t.inner_count - 1], None)
    text = C.string_at(raw).decode()
    return text


def idmap_for(tree_p, n_pars_tips):
    """tip node_index -> taxon number (labels are t<number>); identity beyond the tree's tips"""
    t = tree_p.contents
    m = np.arange(n_pars_tips, dtype=np.uint32)
    for i in range(t.tip_count):
        n = t.nodes[i].contents
        m[n.node_index] = int(n.label.decode()[1:])
    return m


@pytest.mark.parametrize("tips,start,sites,seed", [(12, 4, 600, 0), (30, 11, 1500, 3), (30, 29, 900, 8), (16, 16, 300, 1)])
def test_stepwise_extend_matches_reference(reflib, cudalib, tips, start, sites, seed):
    import test_tree_cpu as tt

    ds = make_ds("dna", tips, sites, seed=tips + start)
    ref, gpu = pars_pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    # a random tree over taxa 0..start-1, whatever order they appear in
    newick = tt.random_newick(np.random.default_rng(seed + 50), start)
    labels = (C.c_char_p * (tips - start + 1))(*[f"t{i}".encode() for i in range(start, tips)], None)
    out = []
    try:
        for lib, pars in ((reflib, ref), (cudalib, gpu)):
            tree = parse_tree(cudalib, newick)
            idmap = idmap_for(tree, tips)
            arr = (capi.ParsimonyP * 1)(pars.p)
            cost = C.c_uint(0)
            rc = lib.pll_fastparsimony_stepwise_extend(tree, arr, 1, labels, idmap.ctypes.data_as(capi.c_uint_p), seed,
                                                       C.byref(cost))
            assert rc == 1, (lib.errno, lib.errmsg)
            t = tree.contents
            assert (t.tip_count, t.inner_count, t.edge_count) == (tips, tips - 2, 2 * tips - 3)
            out.append((cost.value if tips > start else None, splits(tree), export(cudalib, tree)))
        assert out[0][0] == out[1][0]
        assert out[0][1] == out[1][1]
        assert out[0][2] == out[1][2]  # same records in the same places: identical Newick text
        assert out[1][1][1] == frozenset(f"t{i}" for i in range(tips))
    finally:
        ref.close()
        gpu.close()


@pytest.mark.parametrize("tips,sites,seed,constrained", [(10, 400, 1, False), (28, 1200, 5, False), (28, 1200, 0, False),
                                                        (28, 300, 9, True), (45, 60, 4, False)])
def test_spr_round_matches_reference(reflib, cudalib, tips, sites, seed, constrained):
    import test_tree_cpu as tt

    ds = make_ds("dna", tips, sites, seed=tips + sites)
    ref, gpu = pars_pair(reflib, cudalib, ds, capi.PATTERN_TIP)
    newick = tt.random_newick(np.random.default_rng(seed + 70), tips)  # a random (poor) tree: many SPRs improve it
    rng = np.random.default_rng(seed)
    # constraint groups per clv index (two groups split the inner nodes; a subtree only moves within its group)
    cmap = (rng.integers(0, 2, size=2 * tips) if constrained else np.zeros(2 * tips)).astype(np.int32)
    out = []
    try:
        for lib, pars in ((reflib, ref), (cudalib, gpu)):
            tree = parse_tree(cudalib, newick)
            idmap = idmap_for(tree, tips)
            arr = (capi.ParsimonyP * 1)(pars.p)
            cost = C.c_uint(0)
            costs = []
            for rnd in range(2):
                rc = lib.pll_fastparsimony_stepwise_spr_round(tree, arr, 1, idmap.ctypes.data_as(capi.c_uint_p),
                                                              seed + rnd, cmap.ctypes.data_as(C.POINTER(C.c_int)),
                                                              C.byref(cost))
                assert rc == 1, (lib.errno, lib.errmsg)
                costs.append(cost.value)
            out.append((costs, splits(tree), export(cudalib, tree)))
        assert out[0][0] == out[1][0]
        assert out[0][1] == out[1][1]
        assert out[0][2] == out[1][2]
        assert out[1][0][1] <= out[1][0][0]
    finally:
        ref.close()
        gpu.close()


def test_stepwise_error_paths(cudalib):
    ds = make_ds("dna", 5, 100, seed=1)
    eng = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    p = Pars(cudalib, eng)
    p.c.tips = 2
    arr = (capi.ParsimonyP * 1)(p.p)
    lab = (C.c_char_p * 5)(*[b"a", b"b", b"c", b"d", b"e"])
    cost = C.c_uint(0)
    assert not cudalib.pll_fastparsimony_stepwise(arr, lab, C.byref(cost), 1, 1)
    assert cudalib.errno == 128  # PLL_ERROR_STEPWISE_TIPS
    p.c.tips = 5
    p.close()
    eng.close()
