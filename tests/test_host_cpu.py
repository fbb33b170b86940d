"""Host-side logic that needs no GPU: the C-ABI library loads and exports every
symbol include/pll_b200.h declares, fails loudly without a device, the host
eigendecomposition matches the reference bit for bit, and the level scheduler
honours RAW/WAR/WAW hazards."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")
harness = importlib.import_module("libpll-2_b200.harness")

HEADER = os.path.join(pkg.REPO_DIR, "include", "pll_b200.h")


@pytest.fixture(scope="module")
def lib():
    return pkg.load()


def test_exports_every_declared_symbol(lib):
    names = capi.exported_symbols_declared_in_header(HEADER)
    assert len(names) > 60
    missing = [n for n in names if not hasattr(lib.lib, n)]
    assert not missing


def test_no_oracle_or_reference_linked(lib):
    """The product library must not depend on anything under oracle/."""
    import subprocess

    out = subprocess.run(["ldd", lib.path], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libpll_ref" not in out and "plf_oracle" not in out


def test_create_fails_loudly_without_device_or_cuda_bit(lib):
    import torch

    p = lib.pll_partition_create(4, 2, 4, 100, 1, 6, 4, 2, capi.ARCH_AVX2)
    assert not p and lib.errno == 901
    p = lib.pll_partition_create(4, 2, 4, 100, 1, 6, 4, 2, capi.ARCH_CUDA | capi.ARCH_AVX)
    assert not p and "Multiple architecture" in lib.errmsg
    if not torch.cuda.is_available():
        p = lib.pll_partition_create(4, 2, 4, 100, 1, 6, 4, 2, capi.ARCH_CUDA)
        assert not p and lib.errno == 900  # no CPU fallback


def _ref_eigen(reflib, ds, index=0):
    p = reflib.pll_partition_create(4, 2, ds.states, 16, len(ds.subst_params), 6, ds.rate_cats, 2, capi.ARCH_AVX2)
    part = p.contents
    f = np.ascontiguousarray(ds.freqs[index])
    s = np.ascontiguousarray(ds.subst_params[index])
    reflib.pll_set_frequencies(p, index, f.ctypes.data_as(capi.c_double_p))
    reflib.pll_set_subst_params(p, index, s.ctypes.data_as(capi.c_double_p))
    assert reflib.pll_update_eigen(p, index) == 1
    st, sp = part.states, part.states_padded
    out = (
        np.ctypeslib.as_array(part.eigenvecs[index], shape=(st * sp,)).copy(),
        np.ctypeslib.as_array(part.inv_eigenvecs[index], shape=(st * sp,)).copy(),
        np.ctypeslib.as_array(part.eigenvals[index], shape=(sp,)).copy(),
        np.ctypeslib.as_array(part.frequencies[index], shape=(sp,)).copy(),
        sp,
    )
    reflib.pll_partition_destroy(p)
    return out


@pytest.mark.parametrize("kind", ["dna", "aa", "g5", "g7", "dna_zero_freq"])
def test_host_eigen_bit_exact(lib, reflib, kind):
    if kind == "dna":
        ds = synth.dna_dataset(4, 16, seed=1, simulate_down_tree=False)
    elif kind == "dna_zero_freq":
        ds = synth.dna_dataset(4, 16, seed=1, simulate_down_tree=False)
        ds.freqs[0] = np.array([0.5, 0.0, 0.3, 0.2])
    elif kind == "aa":
        ds = synth.aa_dataset(4, 16, seed=2, simulate_down_tree=False)
    else:
        ds = synth.generic_dataset(int(kind[1:]), 4, 16, seed=3)
    for index in range(len(ds.subst_params)):
        ev, iev, evals, freqs, sp = _ref_eigen(reflib, ds, index)
        st = ds.states
        my_ev, my_iev, my_evals = np.zeros(st * sp), np.zeros(st * sp), np.zeros(sp)
        s = np.ascontiguousarray(ds.subst_params[index])
        rc = lib.pll_cuda_host_eigen(
            st, sp, s.ctypes.data_as(capi.c_double_p), freqs.ctypes.data_as(capi.c_double_p),
            my_ev.ctypes.data_as(capi.c_double_p), my_iev.ctypes.data_as(capi.c_double_p),
            my_evals.ctypes.data_as(capi.c_double_p),
        )
        assert rc == 1
        assert np.array_equal(my_evals.view(np.uint64), evals.view(np.uint64))
        assert np.array_equal(my_ev.view(np.uint64), ev.view(np.uint64))
        assert np.array_equal(my_iev.view(np.uint64), iev.view(np.uint64))


def _levels(lib, rows):
    ops = (capi.Operation * len(rows))()
    for i, r in enumerate(rows):
        ops[i] = capi.Operation(*r)
    lv = np.zeros(len(rows), dtype=np.uint32)
    n = lib.pll_cuda_schedule_levels(ops, len(rows), lv.ctypes.data_as(capi.c_uint_p))
    return n, lv.tolist()


def test_schedule_levels_tree(lib):
    # ((0,1)4,(2,3)5)6 : two independent cherries then their join
    rows = [(4, 0, 0, 0, -1, 1, 1, -1), (5, 1, 2, 2, -1, 3, 3, -1), (6, 2, 4, 4, 0, 5, 5, 1)]
    assert _levels(lib, rows) == (2, [0, 0, 1])


def test_schedule_levels_buffer_reuse(lib):
    # op2 overwrites CLV 4 that op1 reads (WAR), op3 reads the new CLV 4 (RAW);
    # op4 rewrites scaler 0 that op3 read (WAR on a scaler)
    rows = [
        (4, 0, 0, 0, -1, 1, 1, -1),
        (5, 1, 4, 4, 0, 2, 2, -1),
        (4, 2, 2, 2, -1, 3, 3, -1),
        (6, 3, 4, 4, 2, 5, 5, 1),
        (7, 0, 0, 0, -1, 3, 3, -1),
    ]
    n, lv = _levels(lib, rows)
    assert lv[1] > lv[0] and lv[2] > lv[1] and lv[3] > lv[2]
    assert lv[4] > lv[1]  # scaler 0 was read by op 1
    assert n == max(lv) + 1


def test_schedule_levels_random_tree_matches_depth(lib):
    ds = synth.dna_dataset(64, 16, seed=5, simulate_down_tree=False)
    rows = [tuple(int(x) for x in r) for r in ds.tree.ops]
    n, lv = _levels(lib, rows)
    depth = {}
    for r, l in zip(rows, lv):
        d = 1 + max(depth.get(r[2], -1), depth.get(r[5], -1))
        depth[r[0]] = d
        assert l == d
    assert n == max(depth.values()) + 1


@pytest.mark.parametrize("mode", [0, 1])
def test_gamma_category_rates_match_reference(lib, reflib, mode):
    """pll_compute_gamma_cats (src/gamma.c:220): same published algorithms, constants and termination
    rules => the same bits, over the alpha range clients use and 1..16 categories."""
    import ctypes as C

    for dll in (lib.lib, reflib.lib):
        dll.pll_compute_gamma_cats.restype = C.c_int
        dll.pll_compute_gamma_cats.argtypes = [C.c_double, C.c_uint, C.POINTER(C.c_double), C.c_int]
    rng = np.random.default_rng(17)
    alphas = [0.02, 0.05, 0.1, 0.3, 0.5, 0.7, 1.0, 1.5, 2.0, 5.0, 10.0, 50.0, 99.0] + list(rng.uniform(0.02, 20, 40))
    for alpha in alphas:
        for cats in (1, 2, 3, 4, 5, 8, 16):
            a = (C.c_double * cats)()
            b = (C.c_double * cats)()
            assert reflib.lib.pll_compute_gamma_cats(alpha, cats, a, mode) == 1
            assert lib.lib.pll_compute_gamma_cats(alpha, cats, b, mode) == 1
            assert list(a) == list(b), (alpha, cats, list(a), list(b))
            assert abs(sum(b) / cats - 1.0) < 1e-5
    out = (C.c_double * 4)()
    assert lib.lib.pll_compute_gamma_cats(0.001, 4, out, 0) == 0 and lib.errno == 113
    assert lib.lib.pll_compute_gamma_cats(1.0, 4, out, 7) == 0


def test_parsimony_level_schedule_keeps_every_order_of_the_sequential_list():
    """pll_cuda_schedule_parsimony_levels (host logic of pll_fastparsimony_update_vectors): for random lists over
    a small pool of vectors every read-after-write, write-after-read and write-after-write pair is separated by
    levels, and no operation sits higher than its conflicts demand."""
    import ctypes as C
    import importlib

    import numpy as np

    pkg = importlib.import_module("libpll-2_b200")
    capi = pkg.capi
    dll = C.CDLL(pkg.LIB_PATH)
    f = dll.pll_cuda_schedule_parsimony_levels
    f.restype, f.argtypes = C.c_int, [C.POINTER(capi.ParsBuildOp), C.c_uint, C.c_uint, capi.c_uint_p]
    rng = np.random.default_rng(3)
    for count, vectors in ((1, 3), (7, 5), (60, 12), (300, 40), (300, 400)):
        trip = [(int(rng.integers(0, vectors)), int(rng.integers(0, vectors)), int(rng.integers(0, vectors)))
                for _ in range(count)]
        ops = (capi.ParsBuildOp * count)(*[capi.ParsBuildOp(*t) for t in trip])
        level = np.zeros(count, dtype=np.uint32)
        n = f(ops, count, vectors, level.ctypes.data_as(capi.c_uint_p))
        assert n == level.max() and level.min() >= 1
        for j, (pj, aj, bj) in enumerate(trip):
            need = 1
            for i in range(j):
                pi, ai, bi = trip[i]
                conflict = pi in (aj, bj) or pj in (ai, bi) or pi == pj  # RAW, WAR, WAW
                if conflict:
                    assert level[i] < level[j], (i, j, trip[i], trip[j])
                    need = max(need, int(level[i]) + 1)
            assert level[j] == need, (j, level[j], need)
    bad = (capi.ParsBuildOp * 1)(capi.ParsBuildOp(9, 0, 1))
    assert f(bad, 1, 3, np.zeros(1, dtype=np.uint32).ctypes.data_as(capi.c_uint_p)) == -1


def test_launch_runs_are_cut_at_the_grid_limit(lib):
    """A level with more same-kind operations than gridDim.y allows (65535; a balanced tree of ~130k taxa) is
    served by several launches instead of one invalid one; kinds never mix within a launch."""
    def runs(kinds):
        k = np.asarray(kinds, dtype=np.uint32)
        largest = C.c_uint(0)
        n = lib.pll_cuda_count_launch_runs(k.ctypes.data_as(capi.c_uint_p), len(k), C.byref(largest))
        return n, largest.value

    assert runs([]) == (0, 0)
    assert runs([2] * 10) == (1, 10)
    assert runs([0] * 3 + [1] * 4 + [2] * 5) == (3, 5)
    assert runs([2] * 65535) == (1, 65535)
    assert runs([2] * 65536) == (2, 65535)
    assert runs([2] * 70000 + [3] * 5) == (3, 65535)
    assert runs([1] * 200000) == (4, 65535)


def test_peer_group_fails_cleanly_without_a_device(lib):
    """pll_cuda_peer_* (the exchange of a site-sharded evaluation): bad arguments and a missing device give
    NULL / 0, never a crash."""
    handle = C.create_string_buffer(64)
    assert not lib.pll_cuda_peer_group_create(0, 0, 0, handle)      # world of 0
    assert not lib.pll_cuda_peer_group_create(0, 3, 2, handle)      # rank outside the world
    assert not lib.pll_cuda_peer_group_create(0, 0, 2, None)        # nowhere to put the handle
    import torch
    if not torch.cuda.is_available():
        assert not lib.pll_cuda_peer_group_create(0, 0, 2, handle)  # no device
    assert lib.pll_cuda_peer_allreduce(None, None, None, 3) == 0
    assert lib.pll_cuda_peer_group_check(None) == 0
    lib.pll_cuda_peer_group_destroy(None)


# ---- k_clv_dna_flow: the plan of a one-launch traversal (host arithmetic) ------------------------------------

def _paths(lib, rows, tips, path_max=8):
    n = len(rows)
    ops = (capi.Operation * n)(*[capi.Operation(*r) for r in rows])
    path = (C.c_uint * n)()
    carried = (C.c_int * n)()
    npaths = lib.pll_cuda_schedule_paths(ops, n, tips, path_max, path, carried)
    return npaths, list(path), list(carried)


def _check_plan(rows, npaths, path, carried, path_max):
    """every op in exactly one path; a path is a chain (each op carries the previous op's parent, nobody else
    reads it); whatever else an op reads was written by an EARLIER path; no path is longer than path_max"""
    writer = {r[0]: i for i, r in enumerate(rows)}
    readers = {}
    for i, r in enumerate(rows):
        for c in (r[2], r[5]):
            readers.setdefault(c, []).append(i)
    members = {}
    for i, p in enumerate(path):
        assert 0 <= p < npaths
        members.setdefault(p, []).append(i)
    assert sorted(members) == list(range(npaths))
    for p, ops in members.items():
        assert len(ops) <= path_max
        carrying = [i for i in ops if carried[i]]
        assert len(carrying) == len(ops) - 1, "all but the first op of a path carry a child"
        for i in carrying:
            child = rows[i][2] if carried[i] == 1 else rows[i][5]
            assert child in writer and path[writer[child]] == p, "the carried child is written inside the path"
            assert readers[child] == [i], "nobody else reads a carried CLV"
    for i, r in enumerate(rows):
        for side, c in ((1, r[2]), (2, r[5])):
            if c in writer and carried[i] != side:
                assert path[writer[c]] < path[i], "a child that comes from memory was written by an earlier path"


@pytest.mark.parametrize("tips,kind,path_max", [(64, "random", 8), (64, "random", 3), (64, "random", 1),
                                                (50, "caterpillar", 8), (50, "caterpillar", 2), (300, "random", 8)])
def test_flow_plan_invariants(lib, tips, kind, path_max):
    ds = synth.dna_dataset(tips, 8, seed=tips, tree_kind=kind, simulate_down_tree=False)
    rows = [tuple(int(x) for x in r) for r in ds.tree.ops]
    for t in (tips, 0):  # pattern tips, and tip CLVs (every op inner-inner)
        npaths, path, carried = _paths(lib, rows, t, path_max)
        assert npaths > 0
        _check_plan(rows, npaths, path, carried, path_max)
        if path_max == 1:
            assert npaths == len(rows) and not any(carried)
    if kind == "caterpillar" and path_max == 8:
        npaths, _, _ = _paths(lib, rows, tips, 8)
        assert npaths == -(-len(rows) // 8), "a caterpillar is one chain, cut every 8 ops"


def test_flow_plan_partial_and_shared_children(lib):
    # ((0,1)4,(2,3)5)6: the join carries one cherry, the other one comes from memory
    rows = [(4, 0, 0, 0, -1, 1, 1, -1), (5, 1, 2, 2, -1, 3, 3, -1), (6, 2, 4, 4, 0, 5, 5, 1)]
    npaths, path, carried = _paths(lib, rows, 4)
    assert npaths == 2 and carried[2] in (1, 2) and carried[:2] == [0, 0]
    _check_plan(rows, npaths, path, carried, 8)
    # a partial list: CLV 5 is older than the list
    npaths, path, carried = _paths(lib, [rows[0], rows[2]], 4)
    assert npaths == 1 and carried == [0, 1]
    # CLV 4 read by two ops: it cannot travel in registers
    rows2 = rows + [(7, 3, 4, 4, 0, 0, 0, -1)]
    npaths, path, carried = _paths(lib, rows2, 4)
    assert carried[2] != 1 and carried[3] == 0
    _check_plan(rows2, npaths, path, carried, 8)
    # both children the same CLV
    rows3 = [rows[0], (6, 2, 4, 4, 0, 4, 5, 0)]
    npaths, path, carried = _paths(lib, rows3, 4)
    assert npaths == 2 and carried == [0, 0]


def test_flow_plan_refuses_lists_that_recycle_buffers(lib):
    # op 2 overwrites CLV 4 that op 1 read: only launch levels keep that order
    rows = [(4, 0, 0, 0, -1, 1, 1, -1), (5, 1, 4, 4, 0, 2, 2, -1), (4, 2, 2, 2, -1, 3, 3, -1)]
    assert _paths(lib, rows, 4)[0] == 0
    # a scaler written by another op than the CLV's writer
    rows = [(4, 0, 0, 0, -1, 1, 1, -1), (5, 1, 2, 2, -1, 3, 3, -1), (6, 2, 4, 4, 1, 5, 5, 0)]
    assert _paths(lib, rows, 4)[0] == 0


def test_flow_plan_random_lists(lib):
    """random trees, random sub-lists (children older than the list), random path limits, some lists with a recycled
    buffer: the plan either keeps its invariants or refuses the list"""
    rng = np.random.default_rng(2024)
    refused = planned = 0
    for trial in range(120):
        tips = int(rng.integers(4, 90))
        ds = synth.dna_dataset(tips, 4, seed=1000 + trial, tree_kind="caterpillar" if trial % 7 == 0 else "random",
                               simulate_down_tree=False)
        rows = [tuple(int(x) for x in r) for r in ds.tree.ops]
        keep = rng.random(len(rows)) < rng.choice([1.0, 0.8, 0.5])
        rows = [r for r, k in zip(rows, keep) if k] or rows[:1]
        recycle = trial % 5 == 0 and len(rows) > 2
        if recycle:  # an op that overwrites the parent of an earlier op which a later op has read
            i = int(rng.integers(0, len(rows) - 1))
            rows = rows + [rows[i]]
        path_max = int(rng.integers(1, 9))
        pattern_tips = tips if trial % 2 else 0
        npaths, path, carried = _paths(lib, rows, pattern_tips, path_max)
        if recycle:
            assert npaths == 0
            refused += 1
            continue
        assert npaths > 0
        _check_plan(rows, npaths, path, carried, path_max)
        planned += 1
    assert refused and planned


def test_flow_queue_never_deadlocks(lib):
    """A model of k_clv_dna_flow's queue: W persistent workers claim (path, chunk) items in queue order, each holding
    the item it works on plus one claimed ahead; an item finishes only when the items that write what it reads have
    finished.  For random trees and any worker count the queue drains: an item's producers sit earlier in the queue,
    so the unfinished item with the smallest position is always some worker's CURRENT item and never waits."""
    rng = np.random.default_rng(7)
    for trial in range(40):
        tips = int(rng.integers(5, 70))
        ds = synth.dna_dataset(tips, 4, seed=3000 + trial, tree_kind="caterpillar" if trial % 6 == 0 else "random",
                               simulate_down_tree=False)
        rows = [tuple(int(x) for x in r) for r in ds.tree.ops]
        path_max = int(rng.integers(1, 9))
        npaths, path, carried = _paths(lib, rows, tips, path_max)
        writer = {r[0]: i for i, r in enumerate(rows)}
        needs = [set() for _ in range(npaths)]
        for i, r in enumerate(rows):
            for side, c in ((1, r[2]), (2, r[5])):
                if c in writer and carried[i] != side:
                    needs[path[i]].add(path[writer[c]])
        assert all(q < p for p in range(npaths) for q in needs[p])
        chunks = int(rng.integers(1, 4))
        items = [(p, c) for p in range(npaths) for c in range(chunks)]  # queue order: path-major
        for workers in (1, 2, 3, 7, 64):
            done, nxt = set(), 0
            slots = []  # per worker: [current, claimed ahead]
            for _ in range(workers):
                cur = items[nxt] if nxt < len(items) else None
                nxt += cur is not None
                ahead = items[nxt] if cur is not None and nxt < len(items) else None
                nxt += ahead is not None
                slots.append([cur, ahead])
            for _ in range(4 * len(items) + 8):
                progressed = False
                for s in slots:
                    cur = s[0]
                    if cur is None:
                        continue
                    if all((q, cur[1]) in done for q in needs[cur[0]]):
                        done.add(cur)
                        s[0] = s[1]
                        s[1] = items[nxt] if s[0] is not None and nxt < len(items) else None
                        nxt += s[1] is not None
                        progressed = True
                if len(done) == len(items):
                    break
                assert progressed, f"deadlock: trial {trial}, {workers} workers, {len(done)} of {len(items)} items done"
            assert len(done) == len(items)
