"""Tree rearrangement moves (pll_moves.c: NNI, SPR, safe SPR, rollback; host code) against the UNMODIFIED
reference's src/utree_moves.c (oracle/_ref).  Two copies of the same parsed tree receive the same random
sequence of moves, one through each library: return codes, pll_errno, reported (branch length, matrix index)
triples, rollback records and the exported Newick text must be identical after every step, and rolling
everything back must restore the starting tree."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
UNode, UTree = capi.UNode, capi.UTree
UP = C.POINTER(UNode)


class SprRb(C.Structure):
    _fields_ = [("p", UP), ("r", UP), ("rb", UP), ("pnb", UP), ("pnnb", UP), ("r_len", C.c_double),
                ("pnb_len", C.c_double), ("pnnb_len", C.c_double)]


class NniRb(C.Structure):
    _fields_ = [("p", UP), ("nni_type", C.c_int)]


class RbUnion(C.Union):
    _fields_ = [("spr", SprRb), ("nni", NniRb)]


class Rollback(C.Structure):
    _anonymous_ = ("u",)
    _fields_ = [("move_type", C.c_int), ("u", RbUnion)]


def bind(path):
    dll = C.CDLL(path)
    for name in ("pll_utree_spr", "pll_utree_spr_safe"):
        f = getattr(dll, name)
        f.restype, f.argtypes = C.c_int, [UP, UP, C.POINTER(Rollback), C.POINTER(C.c_double), C.POINTER(C.c_uint)]
    dll.pll_utree_nni.restype, dll.pll_utree_nni.argtypes = C.c_int, [UP, C.c_int, C.POINTER(Rollback)]
    dll.pll_utree_rollback.restype = C.c_int
    dll.pll_utree_rollback.argtypes = [C.POINTER(Rollback), C.POINTER(C.c_double), C.POINTER(C.c_uint)]
    return dll


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(pkg.REF_PATH):
        pytest.skip("oracle/_ref/libpll_ref.so not built (needs /root/reference)")
    ref, own = bind(pkg.REF_PATH), bind(pkg.LIB_PATH)
    own.pll_utree_parse_newick_string.restype, own.pll_utree_parse_newick_string.argtypes = C.POINTER(UTree), [C.c_char_p]
    own.pll_utree_export_newick.restype, own.pll_utree_export_newick.argtypes = C.c_void_p, [UP, C.c_void_p]
    own.pll_utree_destroy.argtypes = [C.POINTER(UTree), C.c_void_p]
    return ref, own


def errno_of(dll):
    return C.c_int.in_dll(dll, "pll_errno").value


def records(tree):
    """every record of the tree in a fixed order: tips, then the three records of every inner node"""
    t = tree.contents
    out = [t.nodes[i] for i in range(t.tip_count)]
    for i in range(t.tip_count, t.tip_count + t.inner_count):
        n = t.nodes[i]
        out += [n, n.contents.next, n.contents.next.contents.next]
    return out


def newick(own, tree):
    t = tree.contents
    raw = own.pll_utree_export_newick(t.nodes[t.tip_count + t.inner_count - 1], None)
    return C.string_at(raw)


def rb_tuple(rb, recs):
    where = {C.addressof(r.contents): i for i, r in enumerate(recs)}
    idx = lambda p: where.get(C.addressof(p.contents)) if p else None
    if rb.move_type == 1:
        s = rb.spr
        return (1, idx(s.p), idx(s.r), idx(s.rb), idx(s.pnb), idx(s.pnnb), s.r_len, s.pnb_len, s.pnnb_len)
    return (rb.move_type, idx(rb.nni.p), rb.nni.nni_type)


@pytest.mark.parametrize("tips,seed", [(4, 1), (7, 2), (25, 3), (120, 4)])
def test_random_moves_and_rollbacks_match_reference(libs, tips, seed):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_tree_cpu as tt

    ref, own = libs
    rng = np.random.default_rng(seed)
    text = tt.random_newick(rng, tips).encode()
    trees = [own.pll_utree_parse_newick_string(text) for _ in range(2)]
    recs = [records(t) for t in trees]
    start = newick(own, trees[0])
    history = [[], []]
    outcomes = {"ok": 0, "fail": 0}
    for step in range(60):
        kind = rng.integers(0, 3)
        a, b = (int(x) for x in rng.integers(0, len(recs[0]), size=2))
        nni_type = int(rng.integers(0, 4))
        with_report = bool(rng.integers(0, 2))
        res = []
        for side, dll in enumerate((ref, own)):
            rb = Rollback()
            bl, mi = (C.c_double * 3)(), (C.c_uint * 3)()
            C.c_int.in_dll(dll, "pll_errno").value = 0
            if kind == 0:
                rc = dll.pll_utree_nni(recs[side][a], nni_type, C.byref(rb))
            else:
                # plain SPR does not check that r is outside the pruned subtree: only the safe variant gets
                # arbitrary pairs
                rc = dll.pll_utree_spr_safe(recs[side][a], recs[side][b], C.byref(rb), bl if with_report else None,
                                            mi if with_report else None)
            res.append((rc, errno_of(dll) if rc != 1 else 0, list(bl), list(mi),
                        rb_tuple(rb, recs[side]) if rc == 1 else None, newick(own, trees[side])))
            if rc == 1:
                history[side].append(rb)
        assert res[0] == res[1], (step, kind, a, b)
        outcomes["ok" if res[1][0] == 1 else "fail"] += 1
    assert outcomes["ok"] > 5
    # mismatched report arguments, missing rollback
    for dll in (ref, own):
        assert dll.pll_utree_rollback(None, None, None) == 0 and errno_of(dll) == 113
    # undo everything, newest first
    for side, dll in enumerate((ref, own)):
        for rb in reversed(history[side]):
            bl, mi = (C.c_double * 3)(), (C.c_uint * 3)()
            assert dll.pll_utree_rollback(C.byref(rb), bl, mi) == 1
    assert newick(own, trees[0]) == newick(own, trees[1]) == start
    for t in trees:
        own.pll_utree_destroy(t, None)


def test_unchecked_spr_matches_reference_on_valid_moves(libs):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_tree_cpu as tt

    ref, own = libs
    rng = np.random.default_rng(9)
    text = tt.random_newick(rng, 40).encode()
    trees = [own.pll_utree_parse_newick_string(text) for _ in range(2)]
    recs = [records(t) for t in trees]
    done = 0
    for _ in range(200):
        a, b = (int(x) for x in rng.integers(0, len(recs[0]), size=2))
        probe = Rollback()
        # validity is established on our copy with the safe variant, then undone
        if own.pll_utree_spr_safe(recs[1][a], recs[1][b], C.byref(probe), None, None) != 1:
            continue
        assert own.pll_utree_rollback(C.byref(probe), None, None) == 1
        out = []
        for side, dll in enumerate((ref, own)):
            bl, mi = (C.c_double * 3)(), (C.c_uint * 3)()
            assert dll.pll_utree_spr(recs[side][a], recs[side][b], None, bl, mi) == 1
            out.append((list(bl), list(mi), newick(own, trees[side])))
        assert out[0] == out[1]
        done += 1
    assert done > 20
    for t in trees:
        own.pll_utree_destroy(t, None)
