"""Worker of test_site_sharded_two_ranks_against_reference (launched with torch.distributed.run, 2 ranks).

Each rank owns a contiguous site slice (libpll-2_b200/sharding.py) on the CUDA engine and leaves
{logL, d_f, dd_f} of its slice in device memory through the asynchronous entry points; ONE all-reduce sums
them (NCCL over the partition's stream when the box has a GPU per rank, else both ranks share cuda:0 and
the three doubles are reduced over gloo).  Rank 0 then evaluates the reference (oracle/_ref, AVX2) on the
same two slices and on the whole alignment and writes all nine numbers to $SHARDED_OUT."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")
sharding = importlib.import_module("libpll-2_b200.sharding")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    per_rank_gpu = torch.cuda.device_count() >= world
    device = local if per_rank_gpu else 0
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    if per_rank_gpu:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    lib = pkg.load()
    lib.pll_cuda_set_device(device)
    ds = synth.dna_dataset(60, 20_011, seed=5, alpha=0.5)
    ds.pattern_weights = np.random.default_rng(1).integers(1, 4, size=ds.sites).astype(np.uint32)
    lo, hi = sharding.shard_bounds(ds.sites, world, rank)
    eng = harness.Engine(lib, ds, capi.ARCH_CUDA | capi.PATTERN_TIP, sites_slice=slice(lo, hi))
    a, b, m = ds.tree.root_edge
    sa, sb = ds.tree.scaler_of.get(a, -1), ds.tree.scaler_of.get(b, -1)
    pidx = eng.params_indices.ctypes.data_as(capi.c_uint_p)
    t_len = 0.13
    result = torch.zeros(3, dtype=torch.float64, device=dev)
    st = eng.sumtable_alloc()
    eng.update_pmatrices()
    eng.update_partials()
    assert lib.pll_cuda_edge_loglikelihood_async(eng.p, a, sa, b, sb, m, pidx, C.c_void_p(result.data_ptr())) == 1
    eng.update_sumtable(st)
    assert lib.pll_cuda_likelihood_derivatives_async(eng.p, sa, sb, t_len, pidx, st.ctypes.data_as(capi.c_double_p),
                                                     C.c_void_p(result.data_ptr() + 8)) == 1
    peer_sum = None
    if per_rank_gpu:
        # the same three doubles through the peer-memory kernel (pll_cuda_peer_allreduce) first, on a copy
        handle = C.create_string_buffer(64)
        group = lib.pll_cuda_peer_group_create(device, rank, world, handle)
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw if group else b"")
        assert group and all(len(h) == 64 for h in handles), "peer group set-up failed"
        assert lib.pll_cuda_peer_group_connect(group, b"".join(handles)) == 1
        lib.pll_cuda_synchronize(eng.p)
        copy = result.clone()
        for _ in range(3):  # three exchanges: both slot sets and the reuse of the first
            again = result.clone()
            assert lib.pll_cuda_peer_allreduce(group, C.c_void_p(lib.pll_cuda_get_stream(eng.p)),
                                               C.c_void_p(again.data_ptr()), 3) == 1
            lib.pll_cuda_synchronize(eng.p)
            copy = again
        assert lib.pll_cuda_peer_group_check(group) == 1
        peer_sum = copy.cpu()
        dist.barrier()
        lib.pll_cuda_peer_group_destroy(group)
        ext = torch.cuda.ExternalStream(lib.pll_cuda_get_stream(eng.p), device=dev)
        with torch.cuda.stream(ext):
            dist.all_reduce(result)
        lib.pll_cuda_synchronize(eng.p)
        reduced = result.cpu()
    else:
        lib.pll_cuda_synchronize(eng.p)
        reduced = result.cpu()
        dist.all_reduce(reduced)
    eng.close()
    if rank == 0:
        ref = capi.PllLibrary(pkg.REF_PATH, cuda=False)
        parts = np.zeros(3)
        for r in range(world):
            rlo, rhi = sharding.shard_bounds(ds.sites, world, r)
            e = harness.Engine(ref, ds, capi.ARCH_AVX2 | capi.PATTERN_TIP, sites_slice=slice(rlo, rhi))
            logl = e.full_traversal()
            rst = e.sumtable_alloc()
            e.update_sumtable(rst)
            parts += np.array([logl, *e.derivatives(rst, t_len)])
            e.close()
        e = harness.Engine(ref, ds, capi.ARCH_AVX2 | capi.PATTERN_TIP)
        logl = e.full_traversal()
        rst = e.sumtable_alloc()
        e.update_sumtable(rst)
        whole = np.array([logl, *e.derivatives(rst, t_len)])
        e.close()
        extra = peer_sum.numpy() if peer_sum is not None else np.full(3, np.nan)
        np.save(os.environ["SHARDED_OUT"], np.concatenate([reduced.numpy(), parts, whole, extra]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
