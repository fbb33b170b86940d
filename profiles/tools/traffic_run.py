#!/usr/bin/env python
"""The CLV traversals of one bench configuration and nothing else, to be run under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
so that DRAM bytes per traversal can be divided out (profiles/tools/traffic_summary.py).

  python profiles/tools/traffic_run.py dna|aa|repeats|repeats_ids [--sites N] [--reps K]

dna          config 2 / 5 shape (100 taxa, GTR+G4, pattern tips), default 1M sites
narrow       the same tree on 1000 sites: the whole traversal is one launch of k_clv_dna_flow
aa           config 3 (LG4M, 200 taxa x 100k sites), the same input as bench.py's sub-record
repeats      config 4 (1000 taxa x 100k, SITE_REPEATS): K traversals with the identifiers kept
repeats_ids  config 4: K traversals with identifier update
Prints {"traversals": K, ...}; the kernels of the set-up traversal are launched BEFORE the line
"TRAFFIC-RUN-START" is printed and K identical traversals follow, so a capture of the last
K x launches_per_traversal kernels is the measurement."""
import argparse
import importlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["dna", "narrow", "aa", "repeats", "repeats_ids"])
    ap.add_argument("--sites", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    lib = pkg.load()
    if args.config in ("dna", "narrow"):
        sites = args.sites or (1_000_000 if args.config == "dna" else 1000)
        ds = bench.make_dataset("dna", 100, sites, 1, 0)
        attrs = capi.PATTERN_TIP
    elif args.config == "aa":
        sites = args.sites or 100_000
        ds, _ = synth.lg4m_dataset(200, sites, seed=2, ref_path=pkg.REF_PATH)
        attrs = capi.PATTERN_TIP
    else:
        sites = args.sites or 100_000
        ds = synth.dna_dataset(1000, sites, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
        attrs = capi.SITE_REPEATS
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | attrs)
    eng.update_pmatrices()
    launches0 = lib.pll_cuda_kernel_launches()
    eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    first = lib.pll_cuda_kernel_launches() - launches0
    n_ops = len(ds.tree.ops)
    # PLF_GRAPH=0 in the environment keeps every traversal on plain launches (ncu sees graph kernels too,
    # but the launch count per traversal is easier to read this way)
    print("TRAFFIC-RUN-START", flush=True)
    launches0 = lib.pll_cuda_kernel_launches()
    for _ in range(args.reps):
        if args.config == "repeats":
            lib.pll_update_partials_rep(eng.p, eng.ops, n_ops, 0)
        else:
            if args.config == "repeats_ids":
                lib.pll_cuda_invalidate_repeat_identifiers(eng.p)  # renumber every parent, every time
            eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    per = (lib.pll_cuda_kernel_launches() - launches0) // args.reps
    logl = eng.edge_logl()
    print(json.dumps({"config": args.config, "sites": sites, "traversals": args.reps, "launches_per_traversal": int(per),
                      "launches_first_traversal": int(first), "logl": logl}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
