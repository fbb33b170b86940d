#!/usr/bin/env python
"""bench.py -- full-traversal likelihood throughput of the B200 engine.

Workload (BASELINE.json configs[4], the north-star target): DNA, 100 taxa x
10,000,000 sites, GTR+G4, PLL_ATTRIB_PATTERN_TIP, STRONG scaling: the alignment
is cut into N contiguous site slices (libpll-2_b200/sharding.py), rank g owns
slice g and all of its CLVs; one GPU holds all 10M sites (about 130 GB).

One "step" = what a client does per likelihood evaluation and Newton start on
the virtual-root edge:
    pll_update_prob_matrices (all branches)
  + pll_update_partials      (the whole post-order operation list, level-batched)
  + edge log-likelihood      (root edge)
  + pll_update_sumtable + first/second derivatives on the same edge
  + ONE all-reduce of {logL, d_f, dd_f} (3 doubles, NCCL) when N > 1.

  python bench.py [--gpus N] [--steps K] [--warmup W]       # this engine
  python bench.py --impl reference ...                      # the reference's AVX2 CPU path on the host cores

At N = 1 the line also carries `configs`: BASELINE configs 3 (protein LG4M
200 x 100k) and 4 (site repeats 1000 x 100k + Newton), each with roofline,
cpu_baseline and a parity block against the reference on the same input, and
`parity`: this engine against the reference on a 1M-site sample of the workload.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")
sharding = importlib.import_module("libpll-2_b200.sharding")

METRIC = ("CLV site-updates/sec (full traversal: P-matrices + all CLV ops + edge logL + sumtable/derivatives on the "
          "root edge)")
UNIT = "site-updates/s"
TOTAL_SITES_DEFAULT = 10_000_000
SAMPLE_SITES = 1_000_000  # bounded sample of the workload the CPU arm runs per step

# algorithmic bytes per site-update (R = 4 rate categories,
# per-site scalers): reads + writes of CLVs, scalers and tip characters
BYTES_PER_SITE = {
    4: {"ii": 396, "ti": 265, "tt": 134, "edge_ii": 268, "edge_ti": 137, "sumtable_ii": 384, "sumtable_ti": 257,
        "deriv": 132},
    20: {"ii": 1932, "ti": 1289, "tt": 646, "edge_ii": 1292, "edge_ti": 649, "sumtable_ii": 1920, "sumtable_ti": 1281,
         "deriv": 644},
}


def make_dataset(kind: str, tips: int, sites: int, seed: int, rank: int):
    """Same tree and model on every rank (seed), rank-specific columns."""
    rng = np.random.default_rng(seed)
    tree = synth.random_tree(tips, rng)
    col_rng = np.random.default_rng([seed, 1 + rank])
    rates = synth.gamma_rates(0.7, 4)
    if kind == "dna":
        seqs = synth.mutate_alignment(tips, sites, col_rng, synth.DNA_CODES, synth.DNA_AMBIG)
        return synth.Dataset(tree, 4, sites, 4, [synth.GTR_RATES.copy()], [synth.GTR_FREQS.copy()], rates, None,
                             np.zeros(4, dtype=np.uint32), seqs, "pll_map_nt")
    models = [synth.random_aa_model(rng) for _ in range(4)]
    seqs = synth.mutate_alignment(tips, sites, col_rng, synth.AA_CODES, synth.AA_AMBIG)
    return synth.Dataset(tree, 20, sites, 4, [m[0] for m in models], [m[1] for m in models], rates, None,
                         np.arange(4, dtype=np.uint32), seqs, "pll_map_aa")


def op_kinds(ds, pattern_tips=True):
    tips = ds.tree.tips if pattern_tips else 0
    out = []
    for r in ds.tree.ops:
        t1, t2 = int(r[2]) < tips, int(r[5]) < tips
        out.append("tt" if (t1 and t2) else "ti" if (t1 or t2) else "ii")
    return out


def traversal_bytes(ds, sites: int) -> tuple[int, int]:
    """(CLV-update bytes, edge-logL bytes) one full traversal moves, from the op list."""
    b = BYTES_PER_SITE[ds.states]
    total = sum(b[k] for k in op_kinds(ds))
    a, c, _ = ds.tree.root_edge
    tips = ds.tree.tips
    edge = b["edge_ti"] if (a < tips or c < tips) else b["edge_ii"]
    return total * sites, edge * sites


def newton_bytes(ds, sites: int) -> int:
    """sumtable + one derivative evaluation on the root edge"""
    b = BYTES_PER_SITE[ds.states]
    a, c, _ = ds.tree.root_edge
    tips = ds.tree.tips
    return (b["sumtable_ti"] if (a < tips or c < tips) else b["sumtable_ii"]) * sites + b["deriv"] * sites


def fused_traversal_bytes(ds, sites: int) -> int:
    """Bytes the CLV launches move when tip-tip parents stay virtual: a cherry is
    never written nor read back, its consumer reads the two tip codes instead."""
    b = BYTES_PER_SITE[ds.states]
    tips = ds.tree.tips
    kinds = op_kinds(ds)
    cherry = {int(r[0]) for r, k in zip(ds.tree.ops, kinds) if k == "tt"}
    blk = b["tt"] - 2 - 4  # one CLV entry: rates x padded states x 8 B
    total = 0
    for r, k in zip(ds.tree.ops, kinds):
        if k == "tt":
            continue
        cost = blk + 4  # parent CLV + scaler
        for child in (int(r[2]), int(r[5])):
            cost += 1 if child < tips else 2 if child in cherry else blk + 4
        total += cost
    return total * sites


def captured_traffic(key: str):
    """DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum over the CLV launches of one
    traversal) from the committed ncu capture of this workload (profiles/r2_traffic.json, else r1), or None.
    A capture of the same tree and model at another site count is scaled linearly (every byte of a traversal
    is per site) and says so in the source string."""
    prefix, _, sites = key.rpartition("x")
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            doc = json.load(open(os.path.join(REPO, "profiles", name)))
        except Exception:
            continue
        if key in doc:
            return doc[key]["traffic_bytes_per_step"], name
        for other, rec in doc.items():
            if isinstance(rec, dict) and other.startswith(prefix + "x") and sites.isdigit() and other[len(prefix) + 1:].isdigit():
                scale = int(sites) / int(other[len(prefix) + 1:])
                return rec["traffic_bytes_per_step"] * scale, f"{name}: capture of {other} scaled by {scale:g} (bytes are per site)"
    return None, None


def measured_peak_gbs() -> tuple[float, str]:
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
#  the reference's CPU path: one thread per host core, each owning a partition #
#  over a contiguous site slice (the RAxML-NG scheme)     #
# --------------------------------------------------------------------------- #

def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def eval_step(eng, deriv_length: float):
    """One step through the public (blocking) C API: (logL, d_f, dd_f)."""
    logl = eng.full_traversal()
    if getattr(eng, "_st", None) is None:
        eng._st = eng.sumtable_alloc()
    eng.update_sumtable(eng._st)
    d1, d2 = eng.derivatives(eng._st, deriv_length)
    return logl, d1, d2


def cpu_reference_eval(ds, attrs: int, lo: int, hi: int, threads: int, steps: int, warmup: int, step_fn=None):
    """The unmodified reference (oracle/_ref/libpll_ref.so, AVX2) on columns [lo, hi) of `ds`, the columns
    split into `threads` contiguous slices, one partition and one host thread per slice, barrier-timed.
    Returns (seconds per step, per-step sums over the threads of step_fn's values) or None."""
    if not os.path.exists(pkg.REF_PATH):
        return None
    ref = capi.PllLibrary(pkg.REF_PATH, cuda=False)
    n = hi - lo
    threads = max(1, min(threads, n // 64 or 1))
    per = -(-n // threads)
    bounds = [(lo + i * per, min(lo + (i + 1) * per, hi)) for i in range(threads)]
    bounds = [b for b in bounds if b[1] > b[0]]
    threads = len(bounds)
    engines = [None] * threads
    t_len = float(ds.tree.branch_lengths[ds.tree.root_edge[2]])
    step_fn = step_fn or (lambda e: eval_step(e, t_len))

    def setup(i):
        engines[i] = harness.Engine(ref, ds, capi.ARCH_AVX2 | attrs, sites_slice=slice(*bounds[i]))

    ts = [threading.Thread(target=setup, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    barrier = threading.Barrier(threads + 1)
    vals = [[None] * (warmup + steps) for _ in range(threads)]

    def work(i):
        for s in range(warmup + steps):
            barrier.wait()
            vals[i][s] = step_fn(engines[i])
            barrier.wait()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    times = []
    for s in range(warmup + steps):
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        times.append(time.perf_counter() - t0)
    [t.join() for t in ts]
    for e in engines:
        e.close()
    last = np.sum(np.array([v[-1] for v in vals], dtype=np.float64), axis=0)
    return sum(times[warmup:]) / steps, last, threads


def cpu_port_run(kind: str, tips: int, sites: int, seed: int):
    """Fallback when the reference build did not travel: the scalar C port
    (oracle/plf_oracle.c) replayed on one core with host-computed P-matrices."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_api
    from test_oracle_vs_reference import oracle_pmatrix_block, run_oracle_traversal

    orc = oracle_api.load(pkg.ORACLE_PATH)
    lib = pkg.load()
    ds = make_dataset(kind, tips, sites, seed, 0)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    eng.update_pmatrices()  # fills the host eigen arrays the port needs
    t0 = time.perf_counter()
    block = oracle_pmatrix_block(orc, eng)
    run_oracle_traversal(orc, eng, block, False)
    dt = time.perf_counter() - t0
    eng.close()
    return len(ds.tree.ops) * sites / dt, 1e3 * dt, float("nan"), "port"


def rel_err(a: float, b: float) -> float:
    return abs(a - b) / max(abs(b), 1e-300)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.cpu_threads or host_threads()
    sample = min(args.cpu_sample_sites or SAMPLE_SITES, args.total_sites)
    ds = make_dataset(args.kind, args.tips, sample, args.seed, 0)
    res = cpu_reference_eval(ds, capi.PATTERN_TIP, 0, sample, threads, args.steps, args.warmup)
    n_ops = len(ds.tree.ops)
    if res is None:
        value, ms, logl, knd = cpu_port_run(args.kind, args.tips, 20000, args.seed)
        threads, text = 1, f"{args.tips} taxa x 20000 sites, scalar port, 1 step"
        vals = [logl, float("nan"), float("nan")]
    else:
        sec, vals, threads = res
        value, ms, knd = n_ops * sample / sec, 1e3 * sec, "reference"
        text = (f"{args.tips} taxa x {sample} sites per step (a {sample}-site sample of the {args.total_sites}-site "
                f"workload, {-(-sample // threads)} per thread), AVX2+PATTERN_TIP, one partition per thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": knd, "sample": text},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "logl": float(vals[0]), "d_f": float(vals[1]), "dd_f": float(vals[2]),
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    def short(n):
        return f"{n // 1_000_000}M" if n % 1_000_000 == 0 else str(n)
    if args.kind == "dna":
        name = (f"synthetic DNA {args.tips} taxa x {short(args.total_sites)} sites GTR+G4, pattern-tip on, full "
                f"traversal + edge logL + sumtable/derivatives, site-sharded over {world} GPU(s)")
    else:
        name = (f"synthetic protein {args.tips} taxa x {short(args.total_sites)} sites, 4 rate matrices, pattern-tip on, "
                f"full traversal + edge logL + sumtable/derivatives, site-sharded over {world} GPU(s)")
    return {"workload": name, "taxa": args.tips, "sites_total": args.total_sites,
            "sites_per_gpu": -(-args.total_sites // world) if args.scaling == "strong" else args.sites,
            "states": 4 if args.kind == "dna" else 20, "rate_cats": 4, "attributes": "ARCH_CUDA|PATTERN_TIP",
            "sharding": (f"contiguous site slices x{world}, one all-reduce of {{logL, d_f, dd_f}} per step (peer-memory "
                         "kernel over NVLink; NCCL with --nccl-allreduce)" if world > 1 else "single GPU"),
            "l2": "inputs larger than L2: each step streams all CLVs of the slice (>= 12 GB per 1M sites) vs 126 MB L2"}


# --------------------------------------------------------------------------- #
#  this engine                                                                  #
# --------------------------------------------------------------------------- #

def device_timed(torch, ext, fn, reps, warm=2):
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


def gpu_eval(lib, eng, deriv_length):
    return eval_step(eng, deriv_length)


def parity_block(gpu_vals, ref_vals, what: str):
    return {"against": what, "logl": float(gpu_vals[0]), "logl_reference": float(ref_vals[0]),
            "logl_rel_err": rel_err(gpu_vals[0], ref_vals[0]),
            "d_f_rel_err": rel_err(gpu_vals[1], ref_vals[1]), "dd_f_rel_err": rel_err(gpu_vals[2], ref_vals[2]),
            "tolerance": {"logl": 1e-10, "derivatives": 1e-9}}


def config3_record(lib, torch, local, threads):
    """BASELINE config 3: protein, LG4M per-rate matrices, 200 taxa x 100k sites, pattern tips."""
    tips, sites = 200, 100_000
    ds, model_name = synth.lg4m_dataset(tips, sites, seed=2, ref_path=pkg.REF_PATH)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=torch.device("cuda", local))
    t_len = float(ds.tree.branch_lengths[ds.tree.root_edge[2]])
    eng.update_pmatrices()
    launches0 = lib.pll_cuda_kernel_launches()
    ms_trav = device_timed(torch, ext, eng.update_partials, reps=10)
    launches = (lib.pll_cuda_kernel_launches() - launches0) // 12
    gpu_vals = gpu_eval(lib, eng, t_len)
    t0 = time.perf_counter()
    for _ in range(5):
        eval_step(eng, t_len)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / 5
    eng.close()
    n_ops = len(ds.tree.ops)
    clv_bytes, _ = traversal_bytes(ds, sites)
    peak, peak_src = measured_peak_gbs()
    traffic, tsrc = captured_traffic(f"aa_lg4m_{tips}x{sites}")
    rec = {"workload": f"synthetic protein {tips} taxa x {sites} sites, {model_name}, G4, pattern-tip on",
           "metric": METRIC, "value": n_ops * sites / (e2e_ms * 1e-3), "unit": UNIT,
           "traversal_ms": ms_trav, "traversal_site_updates_per_s": n_ops * sites / (ms_trav * 1e-3),
           "e2e_ms_per_step": e2e_ms, "launches_per_traversal": int(launches), "logl": float(gpu_vals[0]),
           "roofline": {"bound": "hbm", "kernel": "k_clv_aa_mma_stream<ii|ti> + k_clv_aa_tt (all CLV launches of a traversal)",
                        "achieved": clv_bytes / ms_trav / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": clv_bytes / ms_trav / 1e6 / peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_step": clv_bytes, "traffic": traffic, "traffic_source": tsrc,
                        "achieved_dram": (traffic / ms_trav / 1e6) if traffic else None,
                        "frac_dram": (traffic / ms_trav / 1e6 / peak) if traffic else None}}
    res = cpu_reference_eval(ds, capi.PATTERN_TIP, 0, sites, threads, 3, 1)
    if res is not None:
        sec, ref_vals, used = res
        rec["cpu_baseline"] = {"value": n_ops * sites / sec, "unit": UNIT, "cores": used, "kind": "reference",
                               "ms_per_step": 1e3 * sec,
                               "sample": f"the whole {tips} x {sites} input, 3 timed steps, AVX2+PATTERN_TIP, one partition per thread"}
        rec["parity"] = parity_block(gpu_vals, ref_vals, "reference (oracle/_ref, AVX2) on the same input, site-split")
    # "also run once with per-rate scalers" (PLL_ATTRIB_RATE_SCALERS)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP | capi.RATE_SCALERS)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=torch.device("cuda", local))
    eng.update_pmatrices()
    ms_rate = device_timed(torch, ext, eng.update_partials, reps=10)
    vals_rate = gpu_eval(lib, eng, t_len)
    eng.close()
    rec["per_rate_scalers"] = {"traversal_ms": ms_rate, "logl": float(vals_rate[0])}
    res = cpu_reference_eval(ds, capi.PATTERN_TIP | capi.RATE_SCALERS, 0, sites, threads, 1, 0)
    if res is not None:
        rec["per_rate_scalers"]["cpu_ms_per_step"] = 1e3 * res[0]
        rec["per_rate_scalers"]["parity"] = parity_block(vals_rate, res[1], "reference (oracle/_ref, AVX2|RATE_SCALERS) on the same input")
    return rec


def config4_record(lib, torch, local, threads):
    """BASELINE config 4: DNA 1000 taxa x 100k repeat-heavy sites, PLL_ATTRIB_SITE_REPEATS, traversal with and
    without identifier update, Newton-Raphson on branches (examples/newton/newton.c:31-96)."""
    tips, sites = 1000, 100_000
    ds = synth.dna_dataset(tips, sites, seed=3, alpha=0.3, brlen=(0.002, 0.05), simulate_down_tree=True)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.SITE_REPEATS)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=torch.device("cuda", local))
    t_len = float(ds.tree.branch_lengths[ds.tree.root_edge[2]])
    n_ops = len(ds.tree.ops)
    eng.update_pmatrices()
    t0 = time.perf_counter()
    eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    first_ms = 1e3 * (time.perf_counter() - t0)
    rep = eng.part.repeats.contents
    ids = [int(rep.pernode_ids[n]) or sites for n in range(tips, ds.tree.nodes)]
    ms_noid = device_timed(torch, ext, lambda: lib.pll_update_partials_rep(eng.p, eng.ops, n_ops, 0), reps=10)
    # identifier update of every parent of the list (the library renumbers a parent only when a child's identifiers
    # changed: pll_cuda_invalidate_repeat_identifiers forgets that history before every traversal) ...
    for _ in range(3):
        lib.pll_cuda_invalidate_repeat_identifiers(eng.p)
        eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    t0 = time.perf_counter()
    for _ in range(5):
        lib.pll_cuda_invalidate_repeat_identifiers(eng.p)
        eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    ms_id = 1e3 * (time.perf_counter() - t0) / 5
    # ... and of none: the same list again with unchanged tips (nothing to renumber)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.update_partials()
    lib.pll_cuda_synchronize(eng.p)
    ms_id_unchanged = 1e3 * (time.perf_counter() - t0) / 5
    gpu_vals = gpu_eval(lib, eng, t_len)
    t0 = time.perf_counter()
    for _ in range(5):
        eval_step(eng, t_len)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / 5
    # Newton sweep over the branches below the first 100 operations: sumtable + Newton-Raphson to |d_f| < 1e-5
    branches = []
    for op in list(eng.ops)[:100]:
        branches.append((op.parent_clv_index, op.child1_clv_index, op.child1_matrix_index))
        branches.append((op.parent_clv_index, op.child2_clv_index, op.child2_matrix_index))

    def sweep(e, fused):
        st = e._st if getattr(e, "_st", None) is not None else e.sumtable_alloc()
        e._st = st
        evals, acc = 0, 0.0
        for edge in branches:
            e.update_sumtable(st, edge)
            r = (e.newton if fused else e.newton_host)(st, 1.5 * float(ds.tree.branch_lengths[edge[2]]), edge)
            evals += r[3]
            acc += r[0]
        return acc, evals

    sweep(eng, True)
    t0 = time.perf_counter()
    len_sum_gpu, evals_gpu = sweep(eng, True)
    newton_ms = 1e3 * (time.perf_counter() - t0) / len(branches)
    t0 = time.perf_counter()
    _, evals_host = sweep(eng, False)
    newton_host_ms = 1e3 * (time.perf_counter() - t0) / len(branches)
    eng.close()
    peak, peak_src = measured_peak_gbs()
    alg_bytes = sum(i * BYTES_PER_SITE[4]["ii"] for i in ids)
    traffic, tsrc = captured_traffic(f"repeats_{tips}x{sites}")
    rec = {"workload": f"synthetic DNA {tips} taxa x {sites} sites GTR+G4, PLL_ATTRIB_SITE_REPEATS, traversal + Newton",
           "metric": METRIC, "value": n_ops * sites / (e2e_ms * 1e-3), "unit": UNIT,
           "class_ratio_mean": float(np.mean(ids)) / sites, "first_traversal_ms": first_ms,
           "traversal_no_id_update_ms": ms_noid, "traversal_with_id_update_ms": ms_id,
           "traversal_with_id_update_nothing_changed_ms": ms_id_unchanged,
           "traversal_site_updates_per_s": n_ops * sites / (ms_id * 1e-3), "e2e_ms_per_step": e2e_ms,
           "newton": {"branches": len(branches), "ms_per_branch_fused": newton_ms,
                      "evaluations_per_branch_fused": evals_gpu / len(branches),
                      "ms_per_branch_host_loop": newton_host_ms,
                      "evaluations_per_branch_host_loop": evals_host / len(branches)},
           "logl": float(gpu_vals[0]),
           "roofline": {"bound": "hbm", "kernel": "k_clv_dna_rep (all CLV launches of a traversal, identifiers kept)",
                        "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": traffic, "traffic_source": tsrc,
                        "achieved": (traffic / ms_noid / 1e6) if traffic else alg_bytes / ms_noid / 1e6,
                        "frac": ((traffic if traffic else alg_bytes) / ms_noid / 1e6 / peak),
                        "basis": "ncu DRAM bytes of the traversal / live time" if traffic else
                                 "class counts x 396 B (no ncu capture of this shape committed)",
                        "class_count_bytes_per_step": alg_bytes}}

    def ref_step(e):
        vals = eval_step(e, t_len)
        return vals

    res = cpu_reference_eval(ds, capi.SITE_REPEATS, 0, sites, threads, 3, 1, ref_step)
    if res is not None:
        sec, ref_vals, used = res
        rec["cpu_baseline"] = {"value": n_ops * sites / sec, "unit": UNIT, "cores": used, "kind": "reference",
                               "ms_per_step": 1e3 * sec,
                               "sample": f"the whole {tips} x {sites} input, 3 timed steps, AVX2+SITE_REPEATS, one partition per thread"}
        rec["parity"] = parity_block(gpu_vals, ref_vals, "reference (oracle/_ref, AVX2) on the same input, site-split")
    return rec


def narrow_record(lib, torch, local):
    """A narrow alignment (100 taxa x 1000 sites, the config-2 model): a traversal is bound by launch and dependency
    latency, not by bytes."""
    out = {}
    for sites in (1000, 10000):
        ds = make_dataset("dna", 100, sites, 1, 0)
        eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
        ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=torch.device("cuda", local))
        eng.update_pmatrices()
        us = 1e3 * device_timed(torch, ext, eng.update_partials, reps=200, warm=5)
        t0 = time.perf_counter()
        for _ in range(200):
            eng.full_traversal()
        e2e_us = 1e6 * (time.perf_counter() - t0) / 200
        eng.close()
        out[f"{sites}_sites"] = {"traversal_us": us, "site_updates_per_s": len(ds.tree.ops) * sites / (us * 1e-6),
                                 "full_evaluation_blocking_api_us": e2e_us}
    # the protein counterpart (proteins are a few hundred sites long): 200 taxa, LG4M-shaped model
    aa = {}
    for sites in (250, 1000):
        ds = synth.aa_dataset(200, sites, seed=2)
        eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
        ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=torch.device("cuda", local))
        eng.update_pmatrices()
        us = 1e3 * device_timed(torch, ext, eng.update_partials, reps=100, warm=5)
        eng.close()
        aa[f"{sites}_sites"] = {"traversal_us": us, "site_updates_per_s": len(ds.tree.ops) * sites / (us * 1e-6)}
    return {"workload": "synthetic DNA 100 taxa x 1000 / 10000 sites GTR+G4, pattern-tip on (latency-bound: the whole "
                        "traversal is one launch, k_clv_dna_flow)", **out,
            "protein_200_taxa": {"workload": "synthetic protein 200 taxa x 250 / 1000 sites, 4 rate matrices, G4, pattern-tip on", **aa}}


def run_b200_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    lib = pkg.load()
    lib.pll_cuda_set_device(local)

    # one GPU must hold its slice: CLVs (inner nodes) + scalers + tip codes + sumtable
    free_b, _ = torch.cuda.mem_get_info(dev)
    per_site = (args.tips - 2) * (128 + 4) + args.tips + 128 + 16 if args.kind == "dna" else \
               (args.tips - 2) * (640 + 4) + args.tips + 640 + 16
    fit = int(0.96 * free_b / per_site)
    note = None
    if args.sites > fit:
        note = (f"slice of {args.sites} sites does not fit in {free_b >> 30} GiB of free HBM: reduced to {fit // 32 * 32}")
        args.sites = fit // 32 * 32
        if args.scaling == "strong":
            args.total_sites = args.sites * world

    t_gen = time.perf_counter()
    ds = make_dataset(args.kind, args.tips, args.sites, args.seed, rank)
    gen_s = time.perf_counter() - t_gen
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=dev)
    result = torch.zeros(4, dtype=torch.float64, device=dev)
    a, b, m = ds.tree.root_edge
    sa, sb = ds.tree.scaler_of.get(a, -1), ds.tree.scaler_of.get(b, -1)
    pidx = eng.params_indices.ctypes.data_as(capi.c_uint_p)
    t_len = float(ds.tree.branch_lengths[m])
    st = eng.sumtable_alloc() if args.sites <= 2_000_000 else np.zeros(8)  # a handle: the table lives in HBM
    st_p = st.ctypes.data_as(capi.c_double_p)
    n_ops = len(ds.tree.ops)

    def barrier_sync():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # the exchange of a step: one single-block kernel over peer memory (pll_cuda_peer_allreduce); NCCL when the
    # peer buffers cannot be mapped (or with --nccl-allreduce)
    peer = None
    stream_ptr = C.c_void_p(lib.pll_cuda_get_stream(eng.p))
    if dist and not args.nccl_allreduce:
        handle = C.create_string_buffer(64)
        group = lib.pll_cuda_peer_group_create(local, rank, world, handle)
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw if group else b"")
        ok = bool(group) and all(len(h) == 64 for h in handles) and \
            lib.pll_cuda_peer_group_connect(group, b"".join(handles)) == 1
        agreed = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        if int(agreed.item()) == 1:
            peer = group
        elif group:
            lib.pll_cuda_peer_group_destroy(group)

    def exchange(buf3):
        if peer:
            rc = lib.pll_cuda_peer_allreduce(peer, stream_ptr, C.c_void_p(buf3.data_ptr()), 3)
            assert rc == 1
        else:
            with torch.cuda.stream(ext):
                dist.all_reduce(buf3)

    def step_device(ev=None):
        """everything queued on the partition's stream, results left on the device"""
        eng.update_pmatrices()
        if ev:
            ev[0].record(ext)
        eng.update_partials()
        if ev:
            ev[1].record(ext)
        rc = lib.pll_cuda_edge_loglikelihood_async(eng.p, a, sa, b, sb, m, pidx, C.c_void_p(result.data_ptr()))
        assert rc == 1, lib.errmsg
        rc = lib.pll_update_sumtable(eng.p, a, b, sa, sb, pidx, st_p)
        assert rc == 1, lib.errmsg
        rc = lib.pll_cuda_likelihood_derivatives_async(eng.p, sa, sb, t_len, pidx, st_p,
                                                       C.c_void_p(result.data_ptr() + 8))
        assert rc == 1, lib.errmsg
        if ev:
            ev[2].record(ext)
        if dist:
            exchange(result[:3])  # the one exchange of an evaluation: {logL, d_f, dd_f}
        if ev:
            ev[3].record(ext)

    def step_e2e():
        """the call sequence a libpll-2 client makes: host arguments in, host doubles out"""
        eng.update_pmatrices()
        eng.update_partials()
        v = eng.edge_logl()
        rc = lib.pll_update_sumtable(eng.p, a, b, sa, sb, pidx, st_p)
        assert rc == 1, lib.errmsg
        d1, d2 = C.c_double(), C.c_double()
        rc = lib.pll_compute_likelihood_derivatives(eng.p, sa, sb, t_len, pidx, st_p, C.byref(d1), C.byref(d2))
        assert rc == 1, lib.errmsg
        vals = [v, d1.value, d2.value]
        if dist:
            t = torch.tensor(vals, dtype=torch.float64).pin_memory().to(dev, non_blocking=True)
            dist.all_reduce(t)
            vals = t.tolist()
        return vals

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    barrier_sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = lib.pll_cuda_kernel_launches()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier_sync()
    start.record(ext)
    for k in range(args.steps):
        step_device(evs[k])
    stop.record(ext)
    barrier_sync()
    launches = lib.pll_cuda_kernel_launches() - launches0
    clk = clocks.stop() if rank == 0 else None
    ms_total = start.elapsed_time(stop)
    ms_partials = sum(e[0].elapsed_time(e[1]) for e in evs)
    ms_newton = sum(e[1].elapsed_time(e[2]) for e in evs)
    ms_allreduce = sum(e[2].elapsed_time(e[3]) for e in evs)
    dev_vals = [float(x) for x in result[:3].tolist()]

    # the two ways of summing three doubles over the ranks, alone on the stream (50 back to back)
    collective = None
    if dist:
        collective = {"kind": "peer-memory one-shot all-reduce (k_peer_allreduce over NVLink stores)" if peer else "NCCL all-reduce"}
        scratch = torch.ones(4, dtype=torch.float64, device=dev)

        def timed_exchange(fn):
            for _ in range(5):
                fn()
            barrier_sync()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(ext)
            for _ in range(50):
                fn()
            c1.record(ext)
            barrier_sync()
            return c0.elapsed_time(c1) / 50

        def nccl_once():
            with torch.cuda.stream(ext):
                dist.all_reduce(scratch[:3])

        collective["nccl_us"] = 1e3 * timed_exchange(nccl_once)
        if peer:
            collective["peer_us"] = 1e3 * timed_exchange(
                lambda: lib.pll_cuda_peer_allreduce(peer, stream_ptr, C.c_void_p(scratch.data_ptr()), 3))
            collective["peer_ok"] = lib.pll_cuda_peer_group_check(peer) == 1

    # Newton-Raphson on the root edge over all ranks (examples/newton/newton.c:67-96 with the sums reduced over the
    # site slices): per evaluation the derivative kernel of the slice, the exchange of {d_f, dd_f}, 16 bytes to the
    # host for the step decision.  Every rank takes the same step (the exchange returns the same bits everywhere).
    newton = None
    if not args.no_newton:
        nd = torch.zeros(4, dtype=torch.float64, device=dev)
        rc = lib.pll_update_sumtable(eng.p, a, b, sa, sb, pidx, st_p)
        assert rc == 1, lib.errmsg

        def newton_run(t0, tol=1e-7, max_iters=32):
            t, evals = t0, 0
            while evals < max_iters:
                rc = lib.pll_cuda_likelihood_derivatives_async(eng.p, sa, sb, t, pidx, st_p, C.c_void_p(nd.data_ptr()))
                assert rc == 1, lib.errmsg
                if dist:
                    if peer:
                        assert lib.pll_cuda_peer_allreduce(peer, stream_ptr, C.c_void_p(nd.data_ptr()), 2) == 1
                    else:
                        with torch.cuda.stream(ext):
                            dist.all_reduce(nd[:2])
                lib.pll_cuda_synchronize(eng.p)
                evals += 1
                d1, d2 = nd[:2].tolist()
                if abs(d1) < tol:
                    break
                tn = min(max(t - d1 / d2, 1e-8), 100.0)
                if tn != tn or tn == t:
                    break
                t = tn
            return t, evals

        newton_run(1.5 * t_len)
        barrier_sync()
        t0 = time.perf_counter()
        total_evals = 0
        for k in range(10):
            length, ev = newton_run((0.5 + 0.2 * k) * t_len)
            total_evals += ev
        barrier_sync()
        newton = {"runs": 10, "evaluations": total_evals, "us_per_evaluation": 1e6 * (time.perf_counter() - t0) / total_evals,
                  "branch_length": length, "sites_per_gpu": args.sites,
                  "note": "host-driven Newton-Raphson on the root edge: derivative kernel over the slice + exchange of "
                          "{d_f, dd_f} over the ranks + 16 bytes to the host, per evaluation (wall clock); on this "
                          "synthetic alignment (columns are not evolved along the tree) the length runs to its upper bound"}

    # end to end through the public C API with host buffers and host results
    for _ in range(2):
        step_e2e()
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_vals = step_e2e()
    barrier_sync()
    e2e_s = time.perf_counter() - t0

    # the same with the alignment itself re-sent every step (pll_set_tip_states for every tip: host
    # characters -> state codes -> HBM); a libpll-2 client does this once per analysis
    cold_steps = min(args.steps, 3)
    eng.set_tips()
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(cold_steps):
        eng.set_tips()
        step_e2e()
    barrier_sync()
    cold_s = time.perf_counter() - t0

    times = torch.tensor([ms_total, ms_partials, e2e_s * 1e3, cold_s * 1e3, ms_newton, ms_allreduce],
                         dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_partials, e2e_ms, cold_ms, ms_newton, ms_allreduce = [float(x) for x in times.tolist()]

    total_sites = args.total_sites if args.scaling == "strong" else args.sites * world
    updates_per_step = n_ops * total_sites
    value = updates_per_step * args.steps / (ms_total * 1e-3)
    clv_bytes, edge_bytes = traversal_bytes(ds, args.sites)
    # tip-tip parents stay virtual on the 4-state path when the library says so (pll_cuda_virtual_cherries)
    virtual = hasattr(lib, "pll_cuda_virtual_cherries") and lib.pll_cuda_virtual_cherries(eng.p) == 1
    fused_bytes = fused_traversal_bytes(ds, args.sites) if virtual else clv_bytes
    nwt_bytes = newton_bytes(ds, args.sites)
    peak, peak_src = measured_peak_gbs()
    achieved = clv_bytes * args.steps / (ms_partials * 1e-3) / 1e9
    n_levels = lib.pll_cuda_schedule_levels(eng.ops, n_ops, np.zeros(n_ops, dtype=np.uint32).ctypes.data_as(capi.c_uint_p))
    # per-step host inputs of this path: matrix indices, branch lengths, expm1 values, op descriptors
    h2d = len(eng.matrix_indices) * (4 + 8 + 8 * ds.rate_cats * ds.states) + n_ops * 96
    traffic, tsrc = captured_traffic(f"{args.kind}_{args.tips}x{args.sites}")
    kinds = op_kinds(ds)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "traversals_per_s": args.steps / (ms_total * 1e-3),
        "logl": dev_vals[0], "d_f": dev_vals[1], "dd_f": dev_vals[2], "logl_e2e": e2e_vals[0],
        "clocks": clk,
        "step_breakdown_ms": {"clv_updates": ms_partials / args.steps, "edge_logl_sumtable_derivatives": ms_newton / args.steps,
                              "allreduce_3_doubles": ms_allreduce / args.steps,
                              "pmatrices_and_rest": max(0.0, ms_total - ms_partials - ms_newton - ms_allreduce) / args.steps,
                              "note": "each entry is the max over ranks; the all-reduce entry includes waiting for the slowest rank"},
        "collective": collective, "newton_root_edge": newton,
        "e2e": {"value": updates_per_step * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 24, "ms_per_step": e2e_ms / args.steps,
                "note": "pll_update_prob_matrices + pll_update_partials + pll_compute_edge_loglikelihood + "
                        "pll_update_sumtable + pll_compute_likelihood_derivatives with host arguments (branch lengths, "
                        "expm1 values, operation list) copied in and the host doubles copied back every step; the "
                        "alignment stays resident between evaluations as in the reference",
                "with_alignment_upload": {
                    "value": updates_per_step * cold_steps / (cold_ms * 1e-3), "unit": UNIT, "steps": cold_steps,
                    "ms_per_step": cold_ms / cold_steps, "h2d_bytes_per_step": h2d + args.tips * args.sites,
                    "note": "additionally pll_set_tip_states for every tip inside the timed region"}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm",
                     "kernel": ("k_clv_dna_stream<*> + k_clv_dna_tt_bulk (all CLV launches of the step)" if ds.states == 4
                                else "k_clv_aa_mma_stream<ii|ti> + k_clv_aa_tt (all CLV launches of the step)"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic, "traffic_source": tsrc,
                     "algorithmic_bytes_per_step": clv_bytes,
                     "algorithmic_basis": "SURVEY.md 8(d): every operation reads its children and writes its parent "
                                          f"({kinds.count('tt')} tt x 134 + {kinds.count('ti')} ti x 265 + "
                                          f"{kinds.count('ii')} ii x 396 B per site); `achieved` can exceed `peak` when "
                                          "tip-tip parents are kept virtual and never touch HBM",
                     "virtual_cherries": bool(virtual), "moved_bytes_per_step": fused_bytes,
                     "achieved_moved": fused_bytes * args.steps / (ms_partials * 1e-3) / 1e9,
                     "frac_moved": fused_bytes * args.steps / (ms_partials * 1e-3) / 1e9 / peak,
                     "launches_per_step": int(launches // args.steps), "levels": int(n_levels),
                     "ms_per_step_in_kernel": ms_partials / args.steps,
                     "whole_step_gbs": (clv_bytes + edge_bytes + nwt_bytes) * args.steps / (ms_total * 1e-3) / 1e9},
        "setup": {"dataset_generation_s": gen_s, "note": note},
    }
    # nothing of torch's may outlive the partition's stream it was used on
    del evs, start, stop
    torch.cuda.synchronize()
    eng.close()
    del eng
    if peer:
        lib.pll_cuda_peer_group_destroy(peer)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = args.cpu_threads or host_threads()
        sample = min(args.cpu_sample_sites or SAMPLE_SITES, args.sites)
        res = cpu_reference_eval(ds, capi.PATTERN_TIP, 0, sample, threads, 10, 1)
        if res is not None:
            sec, ref_vals, used = res
            line["cpu_baseline"] = {
                "value": n_ops * sample / sec, "unit": UNIT, "cores": used, "kind": "reference", "ms_per_step": 1e3 * sec,
                "sample": (f"columns [0, {sample}) of the same alignment ({-(-sample // used)} per thread), 10 timed steps, "
                           "AVX2+PATTERN_TIP, one partition per thread")}
            # parity on the same sample: this engine on columns [0, sample) against the reference
            small = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP, sites_slice=slice(0, sample))
            gpu_vals = eval_step(small, t_len)
            small.close()
            line["parity"] = parity_block(gpu_vals, ref_vals,
                                          f"reference (oracle/_ref, AVX2) on columns [0, {sample}) of the workload")
        else:
            v, ms, _, knd = cpu_port_run(args.kind, args.tips, 20000, args.seed)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": knd, "ms_per_step": ms,
                                    "sample": f"{args.tips} taxa x 20000 sites, scalar port, 1 traversal"}
    del ds
    if rank == 0 and world == 1 and not args.no_configs:
        threads = args.cpu_threads or host_threads()
        line["configs"] = {}
        try:
            line["configs"]["dna_100_taxa_narrow"] = narrow_record(lib, torch, local)
        except Exception as e:
            line["configs"]["dna_100_taxa_narrow"] = {"error": f"{type(e).__name__}: {e}"}
        for name, fn in (("aa_lg4m_200x100k", config3_record), ("repeats_1000x100k_newton", config4_record)):
            try:
                line["configs"][name] = fn(lib, torch, local, threads)
            except Exception as e:  # a failing sub-record must not take the headline down with it
                line["configs"][name] = {"error": f"{type(e).__name__}: {e}"}
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--kind", choices=["dna", "aa"], default="dna")
    ap.add_argument("--tips", type=int, default=100)
    ap.add_argument("--total-sites", type=int, default=TOTAL_SITES_DEFAULT,
                    help="strong scaling (BASELINE config 5, the default): this many sites split into contiguous "
                         "slices over the GPUs; one GPU holds them all")
    ap.add_argument("--sites", type=int, default=0,
                    help="weak scaling instead: this many sites per GPU (e.g. 1000000 for BASELINE config 2 per GPU)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-sample-sites", type=int, default=0,
                    help=f"sites of the workload the CPU arm evaluates per step (default {SAMPLE_SITES})")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 3 / config 4 sub-records (N = 1 only)")
    ap.add_argument("--no-newton", action="store_true", help="skip the Newton-Raphson timing on the root edge")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="sum {logL, d_f, dd_f} over the ranks with NCCL instead of the peer-memory kernel")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.sites:
        args.scaling = "weak"
        args.total_sites = args.sites * world
    else:
        # contiguous site slices, boundaries at multiples of 32 sites (libpll-2_b200/sharding.py)
        lo, hi = sharding.shard_bounds(args.total_sites, world, int(os.environ.get("RANK", "0")))
        args.sites = hi - lo
        args.scaling = "strong"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
