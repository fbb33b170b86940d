#!/usr/bin/env python
"""Config 3 (protein LG4M, 200 taxa x 100k sites, pattern tips) traversal time only.
Environment switches (PLF_AA_STAGES, PLF_AA_WARPS, PLF_VIRTUAL_CHERRIES, ...) apply."""
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


def main():
    lib = pkg.load()
    ds, _ = synth.lg4m_dataset(200, 100_000, seed=2, ref_path=pkg.REF_PATH)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p))
    eng.update_pmatrices()
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith(("PLF_", "PLL_CUDA"))}}
    for rep in range(3):
        out[f"traversal_ms_{rep}"] = bench.device_timed(torch, ext, eng.update_partials, reps=10)
    out["logl"] = eng.edge_logl()
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
