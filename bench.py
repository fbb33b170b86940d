#!/usr/bin/env python
"""bench.py -- full-traversal likelihood throughput of the B200 engine.

One "step" = pll_update_prob_matrices (all branches) + pll_update_partials
(the whole post-order operation list, level-batched) + one edge
log-likelihood, on a synthetic alignment already resident in HBM.

Workload (BASELINE.json configs[1]): DNA, 100 taxa x 1,000,000 sites per GPU,
GTR+G4, PLL_ATTRIB_PATTERN_TIP.  With N GPUs every rank owns a contiguous
1M-site slice of an N x 1M-site alignment and all of its CLVs (weak scaling);
the only exchange is one NCCL all-reduce of the log-likelihood scalar.

  python bench.py [--gpus N] [--steps K] [--warmup W]       # this engine
  python bench.py --impl reference ...                      # the reference's AVX2 CPU path on the host cores

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

METRIC = "CLV site-updates/sec (full traversal: P-matrices + all CLV ops + edge logL)"
UNIT = "site-updates/s"

# algorithmic bytes per site-update (R = 4 rate categories,
# per-site scalers): reads + writes of CLVs, scalers and tip characters
BYTES_PER_SITE = {
    4: {"ii": 396, "ti": 265, "tt": 134, "edge_ii": 268, "edge_ti": 137},
    20: {"ii": 1932, "ti": 1289, "tt": 646, "edge_ii": 1292, "edge_ti": 649},
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


def traversal_bytes(ds, sites: int) -> tuple[int, int]:
    """(CLV-update bytes, edge-logL bytes) one full traversal moves, from the op list."""
    b = BYTES_PER_SITE[ds.states]
    tips = ds.tree.tips
    total = 0
    for r in ds.tree.ops:
        t1, t2 = int(r[2]) < tips, int(r[5]) < tips
        total += b["tt"] if (t1 and t2) else b["ti"] if (t1 or t2) else b["ii"]
    a, c, _ = ds.tree.root_edge
    edge = b["edge_ti"] if (a < tips or c < tips) else b["edge_ii"]
    return total * sites, edge * sites


def captured_traffic(kind: str, tips: int, sites: int):
    """DRAM bytes per step of the CLV launches from the committed ncu capture of this
    workload (profiles/r1_traffic.json), or None when the shape was not captured."""
    try:
        doc = json.load(open(os.path.join(REPO, "profiles", "r1_traffic.json")))
        return doc[f"{kind}_{tips}x{sites}"]["traffic_bytes_per_step"]
    except Exception:
        return None


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
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def cpu_reference_run(kind: str, tips: int, sites_per_thread: int, threads: int, steps: int, warmup: int, seed: int):
    """Returns (site-updates/s, ms per step, logL, kind) for the unmodified
    reference (oracle/_ref/libpll_ref.so, AVX2 + PATTERN_TIP)."""
    if not os.path.exists(pkg.REF_PATH):
        return None
    ref = capi.PllLibrary(pkg.REF_PATH, cuda=False)
    ds = make_dataset(kind, tips, sites_per_thread * threads, seed, 0)
    engines = [None] * threads

    def setup(i):
        sl = slice(i * sites_per_thread, (i + 1) * sites_per_thread)
        engines[i] = harness.Engine(ref, ds, capi.ARCH_AVX2 | capi.PATTERN_TIP, sites_slice=sl)

    ts = [threading.Thread(target=setup, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    barrier = threading.Barrier(threads + 1)
    logls = np.zeros((threads, warmup + steps))

    def work(i):
        for s in range(warmup + steps):
            barrier.wait()
            logls[i, s] = engines[i].full_traversal()
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
    total = sum(times[warmup:])
    updates = len(ds.tree.ops) * sites_per_thread * threads * steps
    return updates / total, 1e3 * total / steps, float(logls[:, -1].sum()), "reference"


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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.cpu_threads or host_threads()
    spt = args.cpu_sites_per_thread or -(-args.sites // threads)
    res = cpu_reference_run(args.kind, args.tips, spt, threads, args.steps, args.warmup, args.seed)
    sample = f"{args.tips} taxa x {spt * threads} sites ({spt} per thread) per step, AVX2+PATTERN_TIP"
    if res is None:
        value, ms, logl, knd = cpu_port_run(args.kind, args.tips, 20000, args.seed)
        threads, sample = 1, f"{args.tips} taxa x 20000 sites, scalar port, 1 step"
    else:
        value, ms, logl, knd = res
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": knd, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "logl": logl,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    per = f"{args.sites // 1_000_000}M" if args.sites % 1_000_000 == 0 else str(args.sites)
    name = (f"synthetic DNA {args.tips} taxa x {per} sites per GPU GTR+G4, pattern-tip on, full traversal + edge logL"
            if args.kind == "dna" else
            f"synthetic protein {args.tips} taxa x {per} sites per GPU, LG4M-style 4 matrices, pattern-tip on, "
            "full traversal + edge logL")
    return {"workload": name, "taxa": args.tips, "sites_per_gpu": args.sites,
            "sites_total": args.total_sites if getattr(args, "total_sites", 0) else args.sites * world,
            "states": 4 if args.kind == "dna" else 20, "rate_cats": 4, "attributes": "ARCH_CUDA|PATTERN_TIP",
            "sharding": f"contiguous site slices x{world}, one NCCL all-reduce of logL" if world > 1 else "single GPU",
            "l2": "inputs larger than L2: each step streams all CLVs (>= 12 GB per GPU at 1M sites) vs 126 MB L2"}


# --------------------------------------------------------------------------- #
#  this engine                                                                  #
# --------------------------------------------------------------------------- #

def run_b200_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.load()
    lib.pll_cuda_set_device(local)

    ds = make_dataset(args.kind, args.tips, args.sites, args.seed, rank)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP)
    ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=torch.device("cuda", local))
    result = torch.zeros(2, dtype=torch.float64, device=f"cuda:{local}")
    a, b, m = ds.tree.root_edge
    sa, sb = ds.tree.scaler_of.get(a, -1), ds.tree.scaler_of.get(b, -1)
    pidx = eng.params_indices.ctypes.data_as(capi.c_uint_p)

    def barrier_sync():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        """everything queued on the partition's stream, result left on the device"""
        eng.update_pmatrices()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        eng.update_partials()
        e1.record(ext)
        rc = lib.pll_cuda_edge_loglikelihood_async(eng.p, a, sa, b, sb, m, pidx, C.c_void_p(result.data_ptr()))
        assert rc == 1, lib.errmsg
        if dist:
            with torch.cuda.stream(ext):
                dist.all_reduce(result[:1])
        return e0, e1

    def step_e2e():
        """the call sequence a libpll-2 client makes: host arguments in, host double out"""
        v = eng.full_traversal()
        if dist:
            t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t)
            v = float(t.item())
        return v

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier_sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = lib.pll_cuda_kernel_launches()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    barrier_sync()
    start.record(ext)
    pairs = [step_device() for _ in range(args.steps)]
    stop.record(ext)
    barrier_sync()
    launches = lib.pll_cuda_kernel_launches() - launches0
    clk = clocks.stop() if rank == 0 else None
    ms_total = start.elapsed_time(stop)
    ms_partials = sum(e0.elapsed_time(e1) for e0, e1 in pairs)
    logl_device = float(result[0].item())

    # end to end through the public C API with host buffers and a host result
    for _ in range(2):
        step_e2e()
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        logl_e2e = step_e2e()
    barrier_sync()
    e2e_s = time.perf_counter() - t0

    # the same with the alignment itself re-sent every step (pll_set_tip_states for every tip: host
    # characters -> state codes on the host -> HBM); a libpll-2 client does this once per analysis
    cold_steps = min(args.steps, 3)
    eng.set_tips()
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(cold_steps):
        eng.set_tips()
        step_e2e()
    barrier_sync()
    cold_s = time.perf_counter() - t0

    times = torch.tensor([ms_total, ms_partials, e2e_s * 1e3, cold_s * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_partials, e2e_ms, cold_ms = [float(x) for x in times.tolist()]

    n_ops = len(ds.tree.ops)
    updates_per_step = n_ops * (args.total_sites if args.total_sites else args.sites * world)
    value = updates_per_step * args.steps / (ms_total * 1e-3)
    clv_bytes, edge_bytes = traversal_bytes(ds, args.sites)
    peak, peak_src = measured_peak_gbs()
    achieved = clv_bytes * args.steps / (ms_partials * 1e-3) / 1e9
    n_levels = lib.pll_cuda_schedule_levels(eng.ops, n_ops, np.zeros(n_ops, dtype=np.uint32).ctypes.data_as(capi.c_uint_p))
    # per-step host inputs of this path: matrix indices, branch lengths, expm1 values, op descriptors
    h2d = len(eng.matrix_indices) * (4 + 8 + 8 * ds.rate_cats * ds.states) + n_ops * 96
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "traversals_per_s": args.steps / (ms_total * 1e-3), "logl": logl_device, "logl_e2e": logl_e2e,
        "clocks": clk,
        "e2e": {"value": updates_per_step * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms / args.steps,
                "note": "pll_update_prob_matrices + pll_update_partials + pll_compute_edge_loglikelihood with host "
                        "arguments (branch lengths, expm1 values, operation list) copied in and the host double "
                        "copied back every step; the alignment stays resident between evaluations as in the reference",
                "with_alignment_upload": {
                    "value": updates_per_step * cold_steps / (cold_ms * 1e-3), "unit": UNIT, "steps": cold_steps,
                    "ms_per_step": cold_ms / cold_steps, "h2d_bytes_per_step": h2d + args.tips * args.sites,
                    "note": "additionally pll_set_tip_states for every tip inside the timed region (host char -> "
                            "state code mapping on one host core, then H2D)"}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": ("k_clv_dna_stream<ii|ti> + k_clv_dna_tt_bulk (all CLV launches of the step)" if ds.states == 4
                                else "k_clv_aa_mma_stream<ii|ti> + k_clv_aa_tt (all CLV launches of the step)"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": captured_traffic(args.kind, args.tips, args.sites),
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the CLV launches of one step "
                                       "(profiles/r1_traffic.json); same unit as algorithmic_bytes_per_step",
                     "algorithmic_bytes_per_step": clv_bytes, "launches_per_step": int(n_levels),
                     "ms_per_step_in_kernel": ms_partials / args.steps,
                     "whole_step_gbs": (clv_bytes + edge_bytes) * args.steps / (ms_total * 1e-3) / 1e9},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = args.cpu_threads or host_threads()
        spt = args.cpu_sites_per_thread or -(-args.sites // threads)
        res = cpu_reference_run(args.kind, args.tips, spt, threads, 10, 1, args.seed)
        if res is not None:
            v, ms, _, knd = res
            sample = (f"{args.tips} taxa x {spt * threads} sites ({spt} per thread), 10 timed traversals, "
                      f"AVX2+PATTERN_TIP, one partition per thread")
        else:
            v, ms, _, knd = cpu_port_run(args.kind, args.tips, 20000, args.seed)
            threads, sample = 1, f"{args.tips} taxa x 20000 sites, scalar port, 1 traversal"
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": knd, "sample": sample,
                                "ms_per_step": ms}
    eng.close()
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
    ap.add_argument("--sites", type=int, default=1_000_000, help="sites per GPU (weak scaling)")
    ap.add_argument("--total-sites", type=int, default=0,
                    help="strong scaling (BASELINE config 5): this many sites split into contiguous slices over the GPUs")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-sites-per-thread", type=int, default=0,
                    help="site slice per host thread of the reference arm (default: --sites split over the host threads, "
                         "i.e. the same alignment as one GPU's share)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.scaling = "weak"
    if args.total_sites:
        # contiguous site slices, boundaries at multiples of 32 sites (libpll-2_b200/sharding.py)
        sharding = importlib.import_module("libpll-2_b200.sharding")
        world = int(os.environ.get("WORLD_SIZE", "1"))
        lo, hi = sharding.shard_bounds(args.total_sites, world, int(os.environ.get("RANK", "0")))
        args.sites = hi - lo
        args.scaling = "strong"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
