#!/usr/bin/env python
"""Where a branch-length optimisation spends its time: pll_update_sumtable alone, pll_cuda_newton_branch with the
iteration limit at 1 .. 8 (tolerance 0: the slope is the cost of one iteration inside the one-launch loop), one
blocking pll_compute_likelihood_derivatives, and pll_update_prob_matrices of one branch.  Wall-clock per call
(every call blocks), medians over `reps` calls.

  python profiles/tools/newton_breakdown.py repeats|dna [--sites N]"""
import argparse
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
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402


def med(fn, reps=30):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["repeats", "dna"])
    ap.add_argument("--sites", type=int, default=0)
    args = ap.parse_args()
    lib = pkg.load()
    if args.config == "repeats":
        sites = args.sites or 100_000
        ds = synth.dna_dataset(1000, sites, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
        attrs = capi.SITE_REPEATS
    else:
        sites = args.sites or 1_000_000
        ds = bench.make_dataset("dna", 100, sites, 1, 0)
        attrs = capi.PATTERN_TIP
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | attrs)
    eng.full_traversal()
    st = eng.sumtable_alloc()
    edge = ds.tree.root_edge
    t0 = 1.5 * float(ds.tree.branch_lengths[edge[2]])
    out = {"config": args.config, "sites": sites}

    def sumtable_blocking():
        eng.update_sumtable(st, edge)
        lib.pll_cuda_synchronize(eng.p)

    out["update_sumtable_us"] = med(sumtable_blocking)
    out["derivatives_blocking_us"] = med(lambda: eng.derivatives(st, t0, edge))
    out["synchronize_only_us"] = med(lambda: lib.pll_cuda_synchronize(eng.p))
    for iters in (1, 2, 3, 4, 6, 8):
        r = eng.newton(st, t0, edge, tol=0.0, max_iters=iters)
        out[f"newton_limit_{iters}_us"] = med(lambda: eng.newton(st, t0, edge, tol=0.0, max_iters=iters))
        out[f"newton_limit_{iters}_evaluations"] = r[3]
    r = eng.newton(st, t0, edge)
    out["newton_default_us"] = med(lambda: eng.newton(st, t0, edge))
    out["newton_default_evaluations"] = r[3]
    out["update_all_pmatrices_us"] = med(eng.update_pmatrices)
    print(json.dumps(out, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
