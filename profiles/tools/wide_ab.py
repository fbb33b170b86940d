#!/usr/bin/env python
"""Wide alignments: the ring kernels (every parent written / virtual cherries) against the one-launch path kernel
with one and two (site, rate) blocks per thread (every parent written, carried children never re-read).  100 taxa,
CUDA events on the partition's stream, graph replay."""
import importlib
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402

FLOW = {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW": "1", "PLF_FLOW_MAX_SITES": "100000000", "PLF_FLOW_MAX_UPDATES": "100000000000"}
VARIANTS = (
    ("ring_written", {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW": "0"}),
    ("ring_virtual", {"PLF_VIRTUAL_CHERRIES": "1", "PLF_VIRTUAL_CHERRY_MIN_SITES": "0", "PLF_FLOW": "0"}),
    ("flow_u1", dict(FLOW, PLF_FLOW_UNROLL="1")),
    ("flow_u2", dict(FLOW, PLF_FLOW_UNROLL="2")),
)
# (profiles/r2_wide_ab.json also holds "flow_wide": a shape of the kernel with 8 sweeps per item and all loads of a path
# issued first, measured in session 30 and removed again - see profiles/r2_notes.md)
KEYS = sorted({k for _, env in VARIANTS for k in env})


def main():
    lib = pkg.load()
    out = {}
    sizes = [int(x) for x in sys.argv[1:]] or [30000, 100000, 300000, 1000000]
    for sites in sizes:
        ds = bench.make_dataset("dna", 100, sites, 1, 0)
        row = {}
        for name, env in VARIANTS:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
            ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p))
            eng.update_pmatrices()
            row[name + "_us"] = round(1e3 * bench.device_timed(torch, ext, eng.update_partials, reps=20, warm=3), 2)
            row[name + "_logl"] = eng.edge_logl()
            eng.close()
        out[str(sites)] = row
        print(sites, {k: v for k, v in row.items() if k.endswith("_us")}, file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
