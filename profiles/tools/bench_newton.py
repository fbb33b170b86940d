#!/usr/bin/env python
"""Newton branch-length sweep of config 4 (DNA 1000 taxa x 100k sites, site repeats): per branch one
pll_update_sumtable + Newton-Raphson to |d_f| < 1e-5, host-driven (one blocking
pll_compute_likelihood_derivatives per evaluation) vs fused (pll_cuda_newton_branch, one cooperative launch).
  python profiles/tools/bench_newton.py [--tips 1000] [--sites 100000] [--branches 400] [--plain]"""
import argparse
import importlib
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
harness = importlib.import_module("libpll-2_b200.harness")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tips", type=int, default=1000)
    ap.add_argument("--sites", type=int, default=100_000)
    ap.add_argument("--branches", type=int, default=400)
    ap.add_argument("--plain", action="store_true", help="no site repeats (pattern tips)")
    args = ap.parse_args()
    lib = pkg.load()
    synth = importlib.import_module("libpll-2_b200.synth")
    ds = synth.dna_dataset(args.tips, args.sites, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | (capi.PATTERN_TIP if args.plain else capi.SITE_REPEATS))
    eng.update_pmatrices()
    eng.update_partials()
    st = eng.sumtable_alloc()
    # The operation list is rooted at ds.tree.root_edge: that is the edge whose two CLVs are the two halves of
    # the tree, so its likelihood is a real function of the branch length.  `branches` Newton runs from starting
    # points spread over 0.3x .. 3x the simulated length (each run recomputes the sumtable, as a sweep would).
    edge = ds.tree.root_edge
    true = float(ds.tree.branch_lengths[edge[2]])
    starts = [true * (0.3 + 2.7 * k / max(args.branches - 1, 1)) for k in range(args.branches)]
    edges = [edge] * args.branches
    out = {"tips": args.tips, "sites": args.sites, "branches": len(edges), "repeats": not args.plain,
           "simulated_length": true,
           "threads": os.environ.get("PLF_NEWTON_THREADS", "default"), "bps": os.environ.get("PLF_NEWTON_BPS", "default")}
    for key, fn in (("host_loop", eng.newton_host), ("fused", eng.newton)):
        for warm in (True, False):
            t0 = time.perf_counter()
            evals, lens = 0, 0.0
            for t_start in starts[:40] if warm else starts:
                eng.update_sumtable(st, edge)
                r = fn(st, t_start, edge)
                evals += r[3]
                lens += r[0]
            dt = time.perf_counter() - t0
        out[key + "_ms_per_branch"] = 1e3 * dt / len(edges)
        out[key + "_evaluations_per_branch"] = evals / len(edges)
        out[key + "_mean_length"] = lens / len(edges)
    t0 = time.perf_counter()
    for edge in edges:
        eng.update_sumtable(st, edge)
    lib.pll_cuda_synchronize(eng.p)
    out["sumtable_only_ms_per_branch"] = 1e3 * (time.perf_counter() - t0) / len(edges)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
