#!/usr/bin/env python
"""Where do virtual cherries start to pay?  Traversal time of the config-2 tree (100 taxa, GTR+G4, pattern tips)
over a range of alignment widths with every tip-tip parent written (PLF_VIRTUAL_CHERRIES=0) and with virtual
cherries forced on (PLF_VIRTUAL_CHERRY_MIN_SITES=0); CUDA events on the partition's stream, graph replay."""
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


def main():
    lib = pkg.load()
    out = {}
    for sites in (250, 1000, 4096, 10000, 30000, 60000, 100000, 300000):
        ds = bench.make_dataset("dna", 100, sites, 1, 0)
        row = {}
        for name, env in (("written", {"PLF_VIRTUAL_CHERRIES": "0", "PLF_LEVEL_MAX_SITES": "0", "PLF_FLOW": "0"}),
                          ("virtual", {"PLF_VIRTUAL_CHERRIES": "1", "PLF_VIRTUAL_CHERRY_MIN_SITES": "0", "PLF_LEVEL_MAX_SITES": "0", "PLF_FLOW": "0"}),
                          ("level", {"PLF_VIRTUAL_CHERRIES": "1", "PLF_VIRTUAL_CHERRY_MIN_SITES": "0", "PLF_LEVEL_MAX_SITES": "100000000", "PLF_FLOW": "0"}),
                          ("level_written", {"PLF_VIRTUAL_CHERRIES": "0", "PLF_LEVEL_MAX_SITES": "100000000", "PLF_FLOW": "0"}),
                          ("flow_u1", {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW_MAX_SITES": "100000000", "PLF_FLOW_MAX_UPDATES": "100000000000", "PLF_FLOW": "1", "PLF_FLOW_UNROLL": "1"}),
                          ("flow_u2", {"PLF_VIRTUAL_CHERRIES": "0", "PLF_FLOW_MAX_SITES": "100000000", "PLF_FLOW_MAX_UPDATES": "100000000000", "PLF_FLOW": "1", "PLF_FLOW_UNROLL": "2"})):
            for k in ("PLF_VIRTUAL_CHERRIES", "PLF_VIRTUAL_CHERRY_MIN_SITES", "PLF_LEVEL_MAX_SITES", "PLF_FLOW",
                      "PLF_FLOW_MAX_SITES", "PLF_FLOW_MAX_UPDATES", "PLF_FLOW_UNROLL", "PLF_FLOW_PATH_MAX"):
                os.environ.pop(k, None)
            os.environ.update(env)
            eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
            ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p))
            eng.update_pmatrices()
            row[name + "_us"] = round(1e3 * bench.device_timed(torch, ext, eng.update_partials, reps=200, warm=5), 2)
            row[name + "_logl"] = eng.edge_logl()
            eng.close()
        out[str(sites)] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
