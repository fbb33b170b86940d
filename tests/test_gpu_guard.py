"""Out-of-bounds-write check of the kernels without compute-sanitizer (closed on this GPU pool): with
PLL_CUDA_GUARD=1 every device buffer of a partition is allocated between two 256-byte guard bands and
pll_cuda_check_guards() counts the buffers whose bands were written to.  Small and odd shapes (sites around the
16-byte copy granule and the 64-site tile, 1 .. 8 rate categories, per-rate scalers, pattern tips with virtual
cherries, tip CLVs, site repeats with identifier updates and pair lists, 4 / 5 / 20 states) through every entry
point of the path; the log-likelihood is compared with the reference on the way."""
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")

from test_gpu_parity import LOGL_RTOL, assert_rel  # noqa: E402

pytestmark = pytest.mark.gpu


def exercise(lib, eng, ds, repeats):
    """every call of the path that launches kernels on the partition's buffers"""
    t = ds.tree
    for _ in range(3):  # plain launches, graph capture, graph replay
        eng.update_pmatrices()
        eng.update_partials()
    logl, _ = eng.edge_logl(persite=True)
    eng.root_logl(persite=True)
    last = eng.ops[len(eng.ops) - 1]
    edges = [t.root_edge, (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index),
             (eng.ops[0].parent_clv_index, eng.ops[0].child1_clv_index, eng.ops[0].child1_matrix_index)]
    pattern = bool(eng.attributes & capi.PATTERN_TIP)
    for edge in edges:
        if pattern and edge[0] < t.tips and edge[1] < t.tips:
            continue
        eng.edge_logl(edge)
        st = eng.sumtable_alloc()
        eng.update_sumtable(st, edge)
        eng.derivatives(st, 0.1, edge)
        if lib.is_cuda:
            eng.newton(st, 0.1, edge)
    if not repeats:
        anc = np.zeros(eng.sites * ds.states)
        node, other, m = t.root_edge
        assert lib.pll_compute_node_ancestral(eng.p, node, t.scaler_of.get(node, -1), other, t.scaler_of.get(other, -1), m,
                                              eng.params_indices.ctypes.data_as(capi.c_uint_p),
                                              anc.ctypes.data_as(capi.c_double_p)) == 1
    else:
        lib.pll_update_partials_rep(eng.p, eng.ops, len(eng.ops), 0)
    assert lib.pll_set_tip_states(eng.p, 1, eng.map, ds.seqs[1]) == 1
    eng.update_partials()
    for op in eng.ops:
        eng.clv(op.parent_clv_index)
        eng.scaler(op.parent_scaler_index)
    return logl


SHAPES = [
    # kind, tips, sites, cats, attrs, per_rate
    ("dna", 9, 1, 4, capi.PATTERN_TIP, False),
    ("dna", 9, 15, 4, capi.PATTERN_TIP, False),
    ("dna", 9, 17, 4, capi.PATTERN_TIP, True),
    ("dna", 12, 63, 2, capi.PATTERN_TIP, False),
    ("dna", 12, 65, 1, capi.PATTERN_TIP, False),
    ("dna", 12, 97, 8, capi.PATTERN_TIP, False),
    ("dna", 30, 1501, 4, capi.PATTERN_TIP, False),
    ("dna", 12, 97, 4, 0, False),
    ("dna", 12, 65, 3, 0, True),
    ("dna", 24, 400, 4, capi.SITE_REPEATS, False),
    ("dna", 24, 333, 4, capi.SITE_REPEATS, True),
    ("dna", 20, 17, 2, capi.SITE_REPEATS, False),
    ("aa", 9, 15, 4, capi.PATTERN_TIP, False),
    ("aa", 10, 53, 4, capi.PATTERN_TIP, True),
    ("aa", 10, 65, 2, capi.PATTERN_TIP, False),
    ("aa", 10, 53, 1, capi.PATTERN_TIP, False),
    ("aa", 10, 53, 4, 0, False),
    ("aa", 16, 200, 4, capi.SITE_REPEATS, False),
    ("g5", 10, 41, 4, capi.PATTERN_TIP, False),
    ("g5", 10, 41, 4, capi.SITE_REPEATS, False),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "-".join(map(str, s)))
@pytest.mark.parametrize("launches", ["per-kind", "per-level", "per-traversal"])
def test_guard_bands_stay_intact(reflib, cudalib, monkeypatch, shape, launches):
    kind, tips, sites, cats, attrs, per_rate = shape
    if launches != "per-kind" and not (kind == "dna" and cats in (1, 2, 4) and not attrs & capi.SITE_REPEATS):
        pytest.skip("k_clv_dna_level / k_clv_dna_flow serve contiguous 4-state CLVs with 1, 2 or 4 rate categories")
    monkeypatch.setenv("PLL_CUDA_GUARD", "1")
    monkeypatch.setenv("PLF_LEVEL_MAX_SITES", "0" if launches == "per-kind" else "1000000")
    monkeypatch.setenv("PLF_FLOW", "1" if launches == "per-traversal" else "0")
    if kind == "dna":
        ds = synth.dna_dataset(tips, sites, seed=300 + sites, cats=cats, brlen=(0.002, 0.08))
    elif kind == "aa":
        ds = synth.aa_dataset(tips, sites, seed=301 + sites, cats=cats, brlen=(0.002, 0.08))
    else:
        ds = synth.generic_dataset(5, tips, sites, seed=302, cats=cats, brlen=(0.002, 0.08))
    flags = attrs | (capi.RATE_SCALERS if per_rate else 0)
    gpu = harness.Engine(cudalib, ds, capi.ARCH_CUDA | flags)
    assert cudalib.pll_cuda_check_guards(gpu.p) == 0, "guard mode is on and nothing has run yet"
    logl = exercise(cudalib, gpu, ds, bool(attrs & capi.SITE_REPEATS))
    assert cudalib.pll_cuda_check_guards(gpu.p) == 0, cudalib.errmsg
    ref = harness.Engine(reflib, ds, capi.ARCH_AVX2 | flags)
    ref.update_pmatrices()
    ref.update_partials()
    assert_rel(logl, ref.edge_logl(), LOGL_RTOL, "edge logL")
    ref.close()
    gpu.close()


def test_guard_mode_sees_an_overrun(cudalib, monkeypatch):
    ds = synth.dna_dataset(8, 100, seed=5)
    plain = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    assert cudalib.pll_cuda_check_guards(plain.p) == -1  # not created in guard mode
    plain.close()
    monkeypatch.setenv("PLL_CUDA_GUARD", "1")
    gpu = harness.Engine(cudalib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    gpu.full_traversal()
    assert cudalib.pll_cuda_check_guards(gpu.p) == 0
    assert cudalib.pll_cuda_debug_overrun(gpu.p, ds.tree.tips + 2) == 1
    assert cudalib.pll_cuda_check_guards(gpu.p) == 1
    assert "guard band" in cudalib.errmsg
    gpu.close()
