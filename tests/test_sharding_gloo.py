"""The N > 1 path on CPU (world_size 2, gloo): contiguous site slices per rank, one all-reduce of the
log-likelihood and of the derivative pair.  The per-rank evaluation runs on the reference build
(oracle/_ref, the checker) because there is no GPU here; what is under test is the sharding arithmetic
and the collective that bench.py and a multi-GPU client use: shard_bounds() covers every site once,
per-slice results sum to the single-partition result."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("libpll-2_b200")
sharding = importlib.import_module("libpll-2_b200.sharding")


def test_shard_bounds_cover_every_site_once():
    for sites in (1, 31, 32, 33, 1000, 1_000_003, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            b = sharding.all_bounds(sites, world)
            assert b[0][0] == 0 and b[-1][1] == sites
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and lo <= hi
            assert all(lo % sharding.ALIGN == 0 for lo, hi in b if lo < sites)
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b if hi > lo) <= max(
                sharding.ALIGN * world, sites % sharding.ALIGN + sharding.ALIGN * world)


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, REPO)
    p = importlib.import_module("libpll-2_b200")
    capi = p.capi
    synth = importlib.import_module("libpll-2_b200.synth")
    harness = importlib.import_module("libpll-2_b200.harness")
    sh = importlib.import_module("libpll-2_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ref = capi.PllLibrary(p.REF_PATH, cuda=False)
    ds = synth.dna_dataset(40, 3001, seed=5, alpha=0.6)
    ds.pattern_weights = np.random.default_rng(1).integers(1, 4, size=ds.sites).astype(np.uint32)
    lo, hi = sh.shard_bounds(ds.sites, world, rank)
    eng = harness.Engine(ref, ds, capi.ARCH_AVX2 | capi.PATTERN_TIP, sites_slice=slice(lo, hi))
    logl = eng.full_traversal()
    st = eng.sumtable_alloc()
    eng.update_sumtable(st)
    d1, d2 = eng.derivatives(st, 0.13)
    t = torch.tensor([logl, d1, d2], dtype=torch.float64)
    dist.all_reduce(t)  # the one collective of an evaluation
    if rank == 0:
        full = harness.Engine(ref, ds, capi.ARCH_AVX2 | capi.PATTERN_TIP)
        fl = full.full_traversal()
        fst = full.sumtable_alloc()
        full.update_sumtable(fst)
        f1, f2 = full.derivatives(fst, 0.13)
        np.save(out_path, np.array([t[0].item(), t[1].item(), t[2].item(), fl, f1, f2]))
    dist.barrier()
    dist.destroy_process_group()


def test_site_sharded_evaluation_world_size_2(tmp_path, reflib):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "result.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    assert abs(got[0] - got[3]) <= 1e-10 * abs(got[3]), got
    assert abs(got[1] - got[4]) <= 1e-9 * max(abs(got[4]), 1e-3), got
    assert abs(got[2] - got[5]) <= 1e-9 * max(abs(got[5]), 1e-3), got
