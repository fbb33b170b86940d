#!/usr/bin/env python
"""20 states between narrow and wide: pattern tips through the tip kernels with virtual cherries (the default above
2048 sites) against tips under tip + inner operations read as expanded CLVs (the default up to 2048 sites).
200 taxa, LG4M-shaped model, traversal replayed as one graph."""
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

VARIANTS = (("virtual_cherries", {"PLF_AA_TIP_CLV_MAX_SITES": "0", "PLF_VIRTUAL_CHERRY_MIN_SITES": "0"}),
            ("tip_kernels_written", {"PLF_AA_TIP_CLV_MAX_SITES": "0", "PLF_VIRTUAL_CHERRIES": "0"}),
            ("expanded_tips", {"PLF_AA_TIP_CLV_MAX_SITES": "100000000", "PLF_VIRTUAL_CHERRIES": "0"}))
KEYS = sorted({k for _, e in VARIANTS for k in e})


def main():
    lib = pkg.load()
    out = {}
    for sites in (1000, 2048, 4000, 8000, 16000, 32000):
        ds = synth.aa_dataset(200, sites, seed=2)
        row = {}
        for name, env in VARIANTS:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
            ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p))
            eng.update_pmatrices()
            row[name + "_us"] = round(1e3 * bench.device_timed(torch, ext, eng.update_partials, reps=50, warm=5), 2)
            row[name + "_logl"] = eng.edge_logl()
            eng.close()
        out[str(sites)] = row
        print(sites, {k: v for k, v in row.items() if k.endswith("_us")}, file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
