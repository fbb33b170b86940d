#!/usr/bin/env python
"""Narrow alignments beyond 4-state pattern tips: traversal time (graph replay, CUDA events on the partition's
stream) and kernel launches per traversal for protein, tip-CLV and site-repeat partitions."""
import importlib
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402


def run(lib, ds, attrs):
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | attrs)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p))
    eng.update_pmatrices()
    us = 1e3 * bench.device_timed(torch, ext, eng.update_partials, reps=200, warm=5)
    l0 = lib.pll_cuda_kernel_launches()
    eng.update_partials()
    per = lib.pll_cuda_kernel_launches() - l0
    logl = eng.edge_logl()
    eng.close()
    return {"traversal_us": round(us, 2), "launches": int(per), "ops": len(ds.tree.ops), "logl": logl}


def main():
    lib = pkg.load()
    out = {}
    for sites in (250, 1000, 4000):
        ds = synth.aa_dataset(200, sites, seed=2)
        out[f"aa_200x{sites}_pattern_tip"] = run(lib, ds, capi.PATTERN_TIP)
    for sites in (250, 1000):
        ds = synth.aa_dataset(200, sites, seed=2)
        out[f"aa_200x{sites}_tip_clvs"] = run(lib, ds, 0)
    for sites in (4000, 8000):
        ds = synth.dna_dataset(1000, sites, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
        for name, env in (("ring", {"PLF_FLOW": "0", "PLF_VIRTUAL_CHERRY_MIN_SITES": "0"}),
                          ("flow", {"PLF_FLOW": "1", "PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW_MAX_UPDATES": "1000000000"})):
            os.environ.update(env)
            out[f"dna_1000x{sites}_pattern_tip_{name}"] = run(lib, ds, capi.PATTERN_TIP)
            for k in env:
                os.environ.pop(k)
    for sites in (1000, 4000):
        ds = bench.make_dataset("dna", 100, sites, 1, 0)
        out[f"dna_100x{sites}_tip_clvs"] = run(lib, ds, 0)
        ds = synth.dna_dataset(1000, sites, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
        out[f"dna_1000x{sites}_repeats"] = run(lib, ds, capi.SITE_REPEATS)
        out[f"dna_1000x{sites}_pattern_tip"] = run(lib, ds, capi.PATTERN_TIP)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
