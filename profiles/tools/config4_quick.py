#!/usr/bin/env python
"""Config 4 (DNA 1000 taxa x 100k sites, site repeats) traversal times only: identifiers kept / updated.
Environment switches (PLF_PAIRS_CTAS, PLL_CUDA_NO_PAIR_LISTS, PLF_GRAPH, ...) apply."""
import importlib
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402


def main():
    lib = pkg.load()
    ds = synth.dna_dataset(1000, 100_000, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.SITE_REPEATS)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p))
    n_ops = len(ds.tree.ops)
    eng.update_pmatrices()
    eng.update_partials()
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith(("PLF_", "PLL_CUDA"))}}
    for rep in range(3):
        out[f"no_id_ms_{rep}"] = bench.device_timed(torch, ext, lambda: lib.pll_update_partials_rep(eng.p, eng.ops, n_ops, 0), reps=20)
    for _ in range(3):
        lib.pll_cuda_invalidate_repeat_identifiers(eng.p)
        eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    t0 = time.perf_counter()
    for _ in range(10):
        lib.pll_cuda_invalidate_repeat_identifiers(eng.p)
        eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    out["with_id_ms"] = 1e2 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for _ in range(10):
        eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    out["with_id_nothing_changed_ms"] = 1e2 * (time.perf_counter() - t0)
    # one tip re-set: only the path above it is renumbered
    lib.pll_set_tip_states(eng.p, 17, eng.map, ds.seqs[17])
    lib.pll_cuda_synchronize(eng.p)
    t0 = time.perf_counter()
    eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    out["with_id_one_tip_changed_ms"] = 1e3 * (time.perf_counter() - t0)
    out["logl"] = eng.edge_logl()
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
