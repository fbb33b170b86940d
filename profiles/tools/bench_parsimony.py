#!/usr/bin/env python
"""Fitch parsimony timings on one GPU next to the reference (oracle/_ref, AVX2, one thread) on the same input.

  python profiles/tools/bench_parsimony.py [--tips 100] [--sites 1000000] [--stepwise-tips 100]

traversal  one pll_fastparsimony_update_vectors over a full post-order list (tips-2 ops): ONE launch
edges      all 2*tips-3 edge scores: one launch (pll_cuda_fastparsimony_edge_scores) vs 2*tips-3 reference calls
stepwise   pll_fastparsimony_stepwise, seed 1
Wall-clock host time around the blocking calls (results are on the host when they return).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402


def wall(fn, reps=5, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps * 1e3, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tips", type=int, default=100)
    ap.add_argument("--sites", type=int, default=1_000_000)
    ap.add_argument("--no-stepwise", action="store_true")
    args = ap.parse_args()
    ds = bench.make_dataset("dna", args.tips, args.sites, 1, 0)
    gpu = pkg.load()
    ref = capi.PllLibrary(pkg.REF_PATH, cuda=False)
    out = {"tips": args.tips, "sites": args.sites}
    triples = [(int(r[0]), int(r[2]), int(r[5])) for r in ds.tree.ops]
    ops = (capi.ParsBuildOp * len(triples))(*[capi.ParsBuildOp(*t) for t in triples])
    edges = [(int(r[0]), int(r[2])) for r in ds.tree.ops] + [(int(r[0]), int(r[5])) for r in ds.tree.ops]
    edges.append(tuple(int(x) for x in ds.tree.root_edge[:2]))
    pairs = np.asarray(edges, dtype=np.uint32)
    labels = (C.c_char_p * args.tips)(*[f"t{i}".encode() for i in range(args.tips)])
    for name, lib, attrs in (("gpu", gpu, capi.ARCH_CUDA), ("ref_avx2_1thread", ref, capi.ARCH_AVX2)):
        eng = harness.Engine(lib, ds, attrs | capi.PATTERN_TIP)
        t0 = time.perf_counter()
        p = lib.pll_fastparsimony_init(eng.p)
        r = {"init_ms": (time.perf_counter() - t0) * 1e3, "words": p.contents.packedvector_count,
             "informative": p.contents.informative_count}
        r["traversal_ms"], _ = wall(lambda: lib.pll_fastparsimony_update_vectors(p, ops, len(triples)))
        a, b = edges[-1]
        r["score"] = lib.pll_fastparsimony_edge_score(p, a, b)
        r["edge_score_ms"], _ = wall(lambda: lib.pll_fastparsimony_edge_score(p, a, b), reps=20)
        if lib.is_cuda:
            got = np.zeros(len(pairs), dtype=np.uint32)
            r["all_edges_ms"], _ = wall(lambda: lib.pll_cuda_fastparsimony_edge_scores(
                p, pairs.ctypes.data_as(capi.c_uint_p), len(pairs), got.ctypes.data_as(capi.c_uint_p)))
            r["all_edges_sum"] = int(got.astype(np.uint64).sum())
            # bytes: a traversal reads 2 and writes 1 vector per op; an edge score reads 2
            vec = 4 * ds.states * r["words"]
            r["traversal_GBs"] = 3 * len(triples) * vec / r["traversal_ms"] / 1e6
            r["all_edges_GBs"] = 2 * len(pairs) * vec / r["all_edges_ms"] / 1e6
        else:
            r["all_edges_ms"], s = wall(lambda: sum(lib.pll_fastparsimony_edge_score(p, int(x), int(y)) for x, y in pairs),
                                        reps=2)
            r["all_edges_sum"] = int(s)
        if not args.no_stepwise:
            arr = (capi.ParsimonyP * 1)(p)
            cost = C.c_uint(0)
            t0 = time.perf_counter()
            tree = lib.pll_fastparsimony_stepwise(arr, labels, C.byref(cost), 1, 1)
            r["stepwise_ms"] = (time.perf_counter() - t0) * 1e3
            r["stepwise_cost"] = cost.value if tree else None
        lib.pll_parsimony_destroy(p)
        eng.close()
        out[name] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
