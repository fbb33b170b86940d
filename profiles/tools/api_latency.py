#!/usr/bin/env python
"""Host-side cost of one evaluation on a narrow alignment, call by call (100 taxa x 1000 sites, pattern tips):
for each public call the time until it returns (enqueue) and the time until the device has finished it
(call + pll_cuda_synchronize), medians of 200."""
import importlib
import json
import os
import statistics
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402


def med(fn, sync, reps=200):
    ret, done = [], []
    for i in range(reps + 10):
        sync()
        t0 = time.perf_counter()
        fn()
        t1 = time.perf_counter()
        sync()
        t2 = time.perf_counter()
        if i >= 10:
            ret.append(t1 - t0)
            done.append(t2 - t0)
    return {"returns_us": round(1e6 * statistics.median(ret), 2), "finished_us": round(1e6 * statistics.median(done), 2)}


def main():
    lib = pkg.load()
    out = {}
    for sites in (1000, 10000):
        ds = bench.make_dataset("dna", 100, sites, 1, 0)
        eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
        eng.full_traversal()
        sync = lambda: lib.pll_cuda_synchronize(eng.p)
        st = eng.sumtable_alloc()
        t = float(ds.tree.branch_lengths[ds.tree.root_edge[2]])
        one = eng.matrix_indices[:1].copy()
        row = {
            "pll_update_prob_matrices_all": med(eng.update_pmatrices, sync),
            "pll_update_prob_matrices_one": med(lambda: eng.update_pmatrices(one, eng.branch_lengths[:1]), sync),
            "pll_update_partials_full": med(eng.update_partials, sync),
            "pll_update_partials_last_3_ops": med(lambda: lib.pll_update_partials(eng.p, eng.ops_tail3, 3), sync)
            if hasattr(eng, "ops_tail3") else None,
            "pll_compute_edge_loglikelihood": med(eng.edge_logl, sync),
            "pll_update_sumtable": med(lambda: eng.update_sumtable(st), sync),
            "pll_compute_likelihood_derivatives": med(lambda: eng.derivatives(st, t), sync),
            "synchronize_only": med(lambda: None, sync),
            "full_evaluation": med(eng.full_traversal, sync),
        }
        out[str(sites)] = row
        eng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
