#!/usr/bin/env python
"""Wall time of pll_compress_site_patterns (host strings in, compressed host strings out) on the device
against the reference's one-core multikey quicksort (oracle/_ref), same alignment.
  python profiles/tools/bench_compress.py [taxa] [sites] [distinct_columns]"""
import importlib
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
pkg = importlib.import_module("libpll-2_b200")
import test_gpu_compress as t  # noqa: E402


def main():
    taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    sites = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    distinct = int(sys.argv[3]) if len(sys.argv) > 3 else sites // 2
    seqs = t.random_alignment(taxa, sites, b"ACGT-N", 5, distinct)
    out = {"taxa": taxa, "sites": sites, "distinct_pool": distinct}
    gpu = pkg.load()
    t.run(gpu, seqs[:4], "pll_map_nt", True)  # context + module load
    t0 = time.perf_counter()
    g = t.run(gpu, seqs, "pll_map_nt", True)
    out["gpu_wall_s"] = time.perf_counter() - t0
    out["patterns"] = g[0]
    if os.path.exists(pkg.REF_PATH):
        ref = pkg.capi.PllLibrary(pkg.REF_PATH, cuda=False)
        t0 = time.perf_counter()
        r = t.run(ref, seqs, "pll_map_nt", True)
        out["reference_1core_wall_s"] = time.perf_counter() - t0
        out["identical"] = bool(r[0] == g[0] and r[1] == g[1] and (r[2] == g[2]).all() and (r[3] == g[3]).all())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
