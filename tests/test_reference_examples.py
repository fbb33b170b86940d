"""Config 1 of BASELINE.json: the reference's shipped example
(/root/reference/examples/unrooted/unrooted.c:43-250) replayed call for call
through the ctypes binding.  The CPU test pins the replay against the
reference's own output (-33.387713 / -34.550204 /
-36.830297); the GPU test (tests/test_gpu_parity.py) runs the same replay on
the CUDA engine."""
import ctypes as C
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")

GOLDEN = [-33.387713, -34.550204, -36.830297]


def run_unrooted(lib, arch):
    dp, up = capi.c_double_p, capi.c_uint_p
    p = lib.pll_partition_create(4, 2, 4, 6, 1, 5, 4, 2, arch)
    assert p, lib.errmsg
    bl = np.array([0.2, 0.4, 0.3, 0.5, 0.6])
    freqs = np.array([0.17, 0.19, 0.25, 0.39])
    mi = np.arange(5, dtype=np.uint32)
    subst = np.ones(6)
    rates = np.ascontiguousarray(synth.gamma_rates(1.0, 4))
    lib.pll_set_frequencies(p, 0, freqs.ctypes.data_as(dp))
    lib.pll_set_subst_params(p, 0, subst.ctypes.data_as(dp))
    lib.pll_set_category_rates(p, rates.ctypes.data_as(dp))
    nt = lib.map("pll_map_nt")
    for i, s in enumerate([b"WAAAAB", b"CACACD", b"AGGACA", b"CGTAGT"]):
        assert lib.pll_set_tip_states(p, i, nt, s) == 1
    pi = np.zeros(4, dtype=np.uint32)
    ops = (capi.Operation * 2)(capi.Operation(4, 0, 0, 0, -1, 1, 1, -1), capi.Operation(5, 1, 2, 2, -1, 3, 3, -1))

    def evaluate():
        assert lib.pll_update_prob_matrices(p, pi.ctypes.data_as(up), mi.ctypes.data_as(up), bl.ctypes.data_as(dp), 5) == 1
        lib.pll_update_partials(p, ops, 2)
        return lib.pll_compute_edge_loglikelihood(p, 4, 0, 5, 1, 4, pi.ctypes.data_as(up), None)

    out = [evaluate()]
    assert lib.pll_update_invariant_sites(p) == 1
    assert lib.pll_update_invariant_sites_proportion(p, 0, 0.5) == 1
    out.append(evaluate())
    assert lib.pll_update_invariant_sites_proportion(p, 0, 0.75) == 1
    out.append(evaluate())
    lib.pll_partition_destroy(p)
    return out


@pytest.mark.parametrize("arch", [capi.ARCH_AVX, capi.ARCH_AVX2, capi.ARCH_AVX2 | capi.PATTERN_TIP, capi.ARCH_CPU])
def test_unrooted_example_on_reference(reflib, arch):
    assert run_unrooted(reflib, arch) == pytest.approx(GOLDEN, abs=5e-7)
