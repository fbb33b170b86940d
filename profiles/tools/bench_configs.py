#!/usr/bin/env python
"""Timings of the non-headline BASELINE.json configs on one GPU (device time,
CUDA events on the partition's stream) with achieved GB/s per API call.

  python profiles/tools/bench_configs.py dna|aa|repeats [--sites N] [--tips T] [--ref]

dna     config 2 shape: every call of the path timed separately
aa      config 3: protein, 4 rate matrices (LG4M-style), 200 taxa x 100k sites
repeats config 4: DNA 1000 taxa x 100k repeat-heavy sites, traversal + Newton steps on every branch
--ref   also time the reference (oracle/_ref, AVX2, one thread) on a 1/10 sample
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")
import bench  # noqa: E402


def timed(eng, ext, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ext)
    for _ in range(reps):
        fn()
    e1.record(ext)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["dna", "aa", "repeats"])
    ap.add_argument("--sites", type=int, default=0)
    ap.add_argument("--tips", type=int, default=0)
    ap.add_argument("--per-rate", action="store_true")
    ap.add_argument("--ref", action="store_true")
    args = ap.parse_args()
    lib = pkg.load()
    out = {"config": args.config}
    if args.config == "dna":
        tips, sites = args.tips or 100, args.sites or 1_000_000
        ds = bench.make_dataset("dna", tips, sites, 1, 0)
        attrs = capi.PATTERN_TIP
    elif args.config == "aa":
        tips, sites = args.tips or 200, args.sites or 100_000
        ds = bench.make_dataset("aa", tips, sites, 2, 0)
        attrs = capi.PATTERN_TIP
    else:
        tips, sites = args.tips or 1000, args.sites or 100_000
        ds = synth.dna_dataset(tips, sites, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
        attrs = capi.SITE_REPEATS
    if args.per_rate:
        attrs |= capi.RATE_SCALERS
    t0 = time.perf_counter()
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | attrs)
    lib.pll_cuda_synchronize(eng.p)
    out["setup_s"] = time.perf_counter() - t0
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p))
    n_ops = len(ds.tree.ops)
    b = bench.BYTES_PER_SITE[ds.states]
    out.update(tips=tips, sites=sites, states=ds.states, ops=n_ops)

    eng.update_pmatrices()
    if attrs & capi.SITE_REPEATS:
        t0 = time.perf_counter()
        eng.update_partials()  # computes identifiers + sizes buffers
        lib.pll_cuda_synchronize(eng.p)
        out["first_traversal_with_ids_ms"] = 1e3 * (time.perf_counter() - t0)
        ids = [int(eng.part.repeats.contents.pernode_ids[n]) or sites for n in range(tips, ds.tree.nodes)]
        out["class_ratio_mean"] = float(np.mean(ids)) / sites
        out["logl"] = eng.edge_logl()
        ms = timed(eng, ext, lambda: lib.pll_update_partials_rep(eng.p, eng.ops, n_ops, 0))
        out["traversal_no_id_update_ms"] = ms
        clv_bytes = sum(i * b["ii"] for i in ids)  # per class, gathers
        out["traversal_gbs_algorithmic"] = clv_bytes / ms / 1e6
        t0 = time.perf_counter()
        for _ in range(3):
            eng.update_partials()
        lib.pll_cuda_synchronize(eng.p)
        out["traversal_with_id_update_ms_wall"] = 1e3 * (time.perf_counter() - t0) / 3
    else:
        ms = timed(eng, ext, eng.update_partials)
        clv_bytes, edge_bytes = bench.traversal_bytes(ds, sites)
        out["traversal_ms"] = ms
        out["traversal_gbs"] = clv_bytes / ms / 1e6
        out["site_updates_per_s"] = n_ops * sites / ms * 1e3
        out["logl"] = eng.edge_logl()
    ms = timed(eng, ext, eng.update_pmatrices)
    out["pmatrices_ms"] = ms
    # edge logL, sumtable, derivatives on the root edge (inner-inner) and on a tip edge
    last = eng.ops[n_ops - 1]
    edges = {"root_edge": ds.tree.root_edge,
             "tip_or_child_edge": (last.parent_clv_index, last.child1_clv_index, last.child1_matrix_index)}
    for name, edge in edges.items():
        tip = (attrs & capi.PATTERN_TIP) and (edge[0] < tips or edge[1] < tips)
        ms = timed(eng, ext, lambda: eng.edge_logl(edge), reps=10)
        eb = b["edge_ti" if tip else "edge_ii"] * sites
        out[f"{name}_logl_ms"] = ms
        out[f"{name}_logl_gbs"] = eb / ms / 1e6
        st = eng.sumtable_alloc()
        ms = timed(eng, ext, lambda: eng.update_sumtable(st, edge), reps=10)
        blk = 8 * ds.rate_cats * eng.part.states_padded
        sb = (3 * blk if not tip else 2 * blk + 1) * sites
        out[f"{name}_sumtable_ms"] = ms
        out[f"{name}_sumtable_gbs"] = sb / ms / 1e6
        ms = timed(eng, ext, lambda: eng.derivatives(st, 0.1, edge), reps=20)
        out[f"{name}_derivatives_ms"] = ms
        out[f"{name}_derivatives_gbs"] = (blk + 4) * sites / ms / 1e6
        # the same two reductions queued without host synchronisation (device time per call)
        import ctypes as C
        dev = torch.zeros(2, dtype=torch.float64, device="cuda")
        t = ds.tree
        pidx = eng.params_indices.ctypes.data_as(capi.c_uint_p)
        a, b_, m = edge
        ms = timed(eng, ext, lambda: lib.pll_cuda_edge_loglikelihood_async(
            eng.p, a, t.scaler_of.get(a, -1), b_, t.scaler_of.get(b_, -1), m, pidx, C.c_void_p(dev.data_ptr())), reps=50)
        out[f"{name}_logl_async_ms"] = ms
        out[f"{name}_logl_async_gbs"] = eb / ms / 1e6
        ms = timed(eng, ext, lambda: lib.pll_cuda_likelihood_derivatives_async(
            eng.p, t.scaler_of.get(a, -1), t.scaler_of.get(b_, -1), 0.1, pidx, st.ctypes.data_as(capi.c_double_p),
            C.c_void_p(dev.data_ptr())), reps=50)
        out[f"{name}_derivatives_async_ms"] = ms
        out[f"{name}_derivatives_async_gbs"] = (blk + 4) * sites / ms / 1e6
    if args.config == "repeats":
        # Newton-style sweep: every branch, sumtable + 4 derivative evaluations (examples/newton/newton.c:31-96)
        st = eng.sumtable_alloc()
        t0 = time.perf_counter()
        nb = 0
        for op in list(eng.ops)[:200]:
            for child, m in ((op.child1_clv_index, op.child1_matrix_index), (op.child2_clv_index, op.child2_matrix_index)):
                edge = (op.parent_clv_index, child, m)
                eng.update_sumtable(st, edge)
                for it in range(4):
                    eng.derivatives(st, 0.05 * (it + 1), edge)
                nb += 1
        out["newton_branches"] = nb
        out["newton_ms_per_branch_wall"] = 1e3 * (time.perf_counter() - t0) / nb
        # the same sweep with the whole Newton loop on the device (pll_cuda_newton_branch): sumtable + one launch
        for fused in (False, True):
            t0 = time.perf_counter()
            nb = evals = 0
            for op in list(eng.ops)[:200]:
                for child, m in ((op.child1_clv_index, op.child1_matrix_index), (op.child2_clv_index, op.child2_matrix_index)):
                    edge = (op.parent_clv_index, child, m)
                    eng.update_sumtable(st, edge)
                    r = (eng.newton if fused else eng.newton_host)(st, 1.5 * float(ds.tree.branch_lengths[m]), edge)
                    evals += r[3]
                    nb += 1
            key = "newton_fused" if fused else "newton_host_loop"
            out[key + "_ms_per_branch_wall"] = 1e3 * (time.perf_counter() - t0) / nb
            out[key + "_evaluations_per_branch"] = evals / nb
    eng.close()
    if args.ref and os.path.exists(pkg.REF_PATH):
        ref = capi.PllLibrary(pkg.REF_PATH, cuda=False)
        s10 = max(sites // 10, 16)
        refeng = harness.Engine(ref, ds, capi.ARCH_AVX2 | attrs, sites_slice=slice(0, s10))
        refeng.update_pmatrices()
        refeng.update_partials()
        t0 = time.perf_counter()
        refeng.update_partials()
        out["ref_traversal_ms_1core_tenth_sample"] = 1e3 * (time.perf_counter() - t0)
        out["ref_logl_sample"] = refeng.edge_logl()
        refeng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
