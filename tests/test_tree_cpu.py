"""Tree layer (libpll-2_b200/csrc/pll_tree.c: Newick reader/writer, traversals, operation lists; host
code, no GPU needed).

* golden: the string cases of the reference's own test (test/src/newick-tree.c:17-23 ->
  test/out/newick-tree.out, copied to tests/golden/): tree type, virtual root, counts and the three
  Newick exports (the custom one prints every clv_index, which pins the index template);
* differential: trees parsed by this library are handed to the UNMODIFIED reference's
  pll_utree_traverse / pll_utree_create_operations / pll_utree_export_newick / pll_rtree_* (oracle/_ref,
  same struct layouts) and to this library's: traversal order, operation lists, branch and matrix index
  lists and Newick strings must be identical."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
Operation = capi.Operation
libc = C.CDLL(None)
libc.malloc.restype = C.c_void_p
libc.free.argtypes = [C.c_void_p]


class UNode(C.Structure):
    pass


UNode._fields_ = [("label", C.c_char_p), ("length", C.c_double), ("node_index", C.c_uint), ("clv_index", C.c_uint),
                  ("scaler_index", C.c_int), ("pmatrix_index", C.c_uint), ("next", C.POINTER(UNode)),
                  ("back", C.POINTER(UNode)), ("data", C.c_void_p)]


class UTree(C.Structure):
    _fields_ = [("tip_count", C.c_uint), ("inner_count", C.c_uint), ("edge_count", C.c_uint), ("binary", C.c_int),
                ("nodes", C.POINTER(C.POINTER(UNode))), ("vroot", C.POINTER(UNode))]


class RNode(C.Structure):
    pass


RNode._fields_ = [("label", C.c_char_p), ("length", C.c_double), ("node_index", C.c_uint), ("clv_index", C.c_uint),
                  ("scaler_index", C.c_int), ("pmatrix_index", C.c_uint), ("left", C.POINTER(RNode)),
                  ("right", C.POINTER(RNode)), ("parent", C.POINTER(RNode)), ("data", C.c_void_p)]


class RTree(C.Structure):
    _fields_ = [("tip_count", C.c_uint), ("inner_count", C.c_uint), ("edge_count", C.c_uint),
                ("nodes", C.POINTER(C.POINTER(RNode))), ("root", C.POINTER(RNode))]


UCB = C.CFUNCTYPE(C.c_int, C.POINTER(UNode))
RCB = C.CFUNCTYPE(C.c_int, C.POINTER(RNode))
USER = C.CFUNCTYPE(C.c_void_p, C.POINTER(UNode))


def bind(dll, own):
    f = dll.pll_utree_traverse
    f.restype, f.argtypes = C.c_int, [C.POINTER(UNode), C.c_int, UCB, C.POINTER(C.POINTER(UNode)), C.POINTER(C.c_uint)]
    f = dll.pll_utree_create_operations
    f.restype, f.argtypes = None, [C.POINTER(C.POINTER(UNode)), C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_uint),
                                   C.POINTER(Operation), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    f = dll.pll_utree_export_newick
    f.restype, f.argtypes = C.c_void_p, [C.POINTER(UNode), C.c_void_p]
    f = dll.pll_utree_export_newick_rooted
    f.restype, f.argtypes = C.c_void_p, [C.POINTER(UNode), C.c_double]
    f = dll.pll_rtree_traverse
    f.restype, f.argtypes = C.c_int, [C.POINTER(RNode), C.c_int, RCB, C.POINTER(C.POINTER(RNode)), C.POINTER(C.c_uint)]
    f = dll.pll_rtree_create_operations
    f.restype, f.argtypes = None, [C.POINTER(C.POINTER(RNode)), C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_uint),
                                   C.POINTER(Operation), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    f = dll.pll_rtree_export_newick
    f.restype, f.argtypes = C.c_void_p, [C.POINTER(RNode), C.c_void_p]
    if own:
        for name in ("pll_utree_parse_newick_string", "pll_utree_parse_newick_string_rooted",
                     "pll_utree_parse_newick_string_unroot", "pll_utree_parse_newick"):
            f = getattr(dll, name)
            f.restype, f.argtypes = C.POINTER(UTree), [C.c_char_p]
        f = dll.pll_rtree_parse_newick_string
        f.restype, f.argtypes = C.POINTER(RTree), [C.c_char_p]
        dll.pll_utree_destroy.argtypes = [C.POINTER(UTree), C.c_void_p]
        dll.pll_rtree_destroy.argtypes = [C.POINTER(RTree), C.c_void_p]
        dll.pll_utree_is_rooted.argtypes = [C.POINTER(UTree)]
        dll.pll_utree_check_integrity.argtypes = [C.POINTER(UTree)]
        dll.pll_utree_show_ascii.argtypes = [C.POINTER(UNode), C.c_int]
    return dll


@pytest.fixture(scope="module")
def own():
    return bind(C.CDLL(pkg.LIB_PATH), True)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(pkg.REF_PATH):
        pytest.skip("oracle/_ref/libpll_ref.so not built (needs /root/reference)")
    return bind(C.CDLL(pkg.REF_PATH), False)


def take_string(ptr):
    s = C.string_at(ptr).decode()
    libc.free(ptr)
    return s


@USER
def print_cb(node):
    n = node.contents
    s = ("%s[%u]:%.4f" % ((n.label or b"").decode(), n.clv_index, n.length)).encode()
    p = libc.malloc(len(s) + 1)
    C.memmove(p, s, len(s) + 1)
    return p


def capture_stdout(fn):
    """what a C function printf()s"""
    import tempfile

    libc.fflush(None)
    saved = os.dup(1)
    with tempfile.TemporaryFile() as tmp:
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
            libc.fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read().decode()


GOLDEN_STRINGS = [
    "(A,B,(C,D));",
    "((A,B),C,(D,E));",
    "(A:0.1,B:0.2,(C:0.3,D:0.4):0.5)0:0;",
    "((A:0.1,B:0.2,C:0.3):1,(D:0.4,(E:0.5,F):0.1):0.6);",
    "((taxon1:0.100,2:2)100,  (taxon3:0.2,\n(t4:0.5,t5:0.3)95:0.22):0.6);",
]


def test_newick_golden_string_cases(own):
    text = open(os.path.join(os.path.dirname(__file__), "golden", "newick-tree.out")).read()
    blocks = text.split("*** TREE # ")[1:]
    for i, newick in enumerate(GOLDEN_STRINGS):
        blk = [b for b in blocks if b.startswith(f"S{i + 1}\n")][0]
        tree = own.pll_utree_parse_newick_string(newick.encode())
        note = "NOTE: tree was automatically unrooted!" in blk
        if not tree:
            assert note, f"case {i + 1} should parse as unrooted"
            pll_errno = C.c_int.in_dll(own, "pll_errno").value
            assert pll_errno == 133  # PLL_ERROR_TREE_INVALID: rooted tree where an unrooted one is expected
            rooted = own.pll_utree_parse_newick_string_rooted(newick.encode())
            assert rooted and own.pll_utree_is_rooted(rooted)
            own.pll_utree_destroy(rooted, None)
            tree = own.pll_utree_parse_newick_string_unroot(newick.encode())
        else:
            assert not note
        assert tree and not own.pll_utree_is_rooted(tree)
        t = tree.contents
        assert own.pll_utree_check_integrity(tree) == 1
        kind = "BINARY" if t.binary else "MULTIFURCATING"
        assert f"Tree: {kind}, virtual root at clv_id: {t.vroot.contents.clv_index}\n" in blk
        assert f"Number of tips/inner nodes/edges in tree: {t.tip_count} / {t.inner_count} / {t.edge_count}\n" in blk
        want = dict(re.findall(r"Newick export \((\w+)\): (.*)\n", blk))
        assert take_string(own.pll_utree_export_newick(t.vroot, None)) == want["default"]
        assert take_string(own.pll_utree_export_newick(t.vroot, C.cast(print_cb, C.c_void_p))) == want["custom"]
        assert take_string(own.pll_utree_export_newick_rooted(t.vroot, 6.13)) == want["rooted"]
        # the ASCII rendering: everything between the counts line and the first export line
        art = blk.split("edges in tree:")[1].split("\n", 1)[1].split("Newick export (default)")[0]
        assert capture_stdout(lambda: own.pll_utree_show_ascii(t.vroot, 1 | 2 | 4)) == art
        own.pll_utree_destroy(tree, None)


def random_newick(rng, tips, rooted=False, multifurcate=0.0):
    names = [f"t{i}" for i in range(tips)]
    nodes = [f"{n}:{rng.uniform(0.001, 0.5):.6f}" for n in names]
    while len(nodes) > (2 if rooted else 3):
        k = 3 if (rng.random() < multifurcate and len(nodes) > 4) else 2
        idx = sorted(rng.choice(len(nodes), size=k, replace=False), reverse=True)
        kids = [nodes.pop(i) for i in idx]
        label = f"n{len(nodes)}" if rng.random() < 0.3 else ""
        nodes.append("(" + ",".join(kids) + f"){label}:{rng.uniform(0.001, 0.5):.6f}")
    return "(" + ",".join(nodes) + ");"


def run_utree(dll, vroot, n_nodes, traversal, cb):
    buf = (C.POINTER(UNode) * n_nodes)()
    size = C.c_uint(0)
    rc = dll.pll_utree_traverse(vroot, traversal, cb, buf, C.byref(size))
    order = [C.addressof(buf[i].contents) for i in range(size.value)]
    ops = (Operation * max(size.value, 1))()
    branches = (C.c_double * max(size.value, 1))()
    pm = (C.c_uint * max(size.value, 1))()
    n_mat, n_ops = C.c_uint(0), C.c_uint(0)
    if size.value:
        dll.pll_utree_create_operations(buf, size.value, branches, pm, ops, C.byref(n_mat), C.byref(n_ops))
    return (rc, order, bytes(ops)[:n_ops.value * C.sizeof(Operation)], list(branches)[:n_mat.value],
            list(pm)[:n_mat.value])


@pytest.mark.parametrize("tips,multi,seed", [(4, 0.0, 1), (9, 0.0, 2), (57, 0.0, 3), (400, 0.0, 4), (60, 0.5, 5)])
def test_utree_traversal_and_operations_match_reference(own, ref, tips, multi, seed):
    rng = np.random.default_rng(seed)
    newick = random_newick(rng, tips, multifurcate=multi)
    tree = own.pll_utree_parse_newick_string(newick.encode())
    assert tree, C.c_char_p.in_dll(own, "pll_errmsg")
    t = tree.contents
    assert t.tip_count == tips and own.pll_utree_check_integrity(tree) == 1
    n_nodes = t.tip_count + t.inner_count
    full = UCB(lambda node: 1)
    partial = UCB(lambda node: 1 if (node.contents.clv_index * 2654435761) % 7 else 0)
    starts = [t.vroot] + [t.nodes[i] for i in rng.choice(np.arange(t.tip_count, n_nodes), size=min(4, t.inner_count), replace=False)]
    for start in starts:
        for traversal in (1, 2):
            for cb in (full, partial):
                a = run_utree(ref, start, n_nodes, traversal, cb)
                b = run_utree(own, start, n_nodes, traversal, cb)
                if multi:  # operations are defined for bifurcations only: compare the traversal
                    assert a[:2] == b[:2]
                else:
                    assert a == b
        assert take_string(ref.pll_utree_export_newick(start, None)) == take_string(own.pll_utree_export_newick(start, None))
        assert (take_string(ref.pll_utree_export_newick_rooted(start, 0.25)) ==
                take_string(own.pll_utree_export_newick_rooted(start, 0.25)))
    # the traversal of a full tree visits every node once and yields tips - 2 operations
    rc, order, ops, branches, pm = run_utree(own, t.vroot, n_nodes, 1, full)
    assert rc == 1 and len(order) == n_nodes and len(set(order)) == n_nodes
    if not multi:
        assert len(ops) == (tips - 2) * C.sizeof(Operation) and len(branches) == 2 * tips - 3
    # round trip: exporting and re-reading gives the same string
    s = take_string(own.pll_utree_export_newick(t.vroot, None))
    again = own.pll_utree_parse_newick_string(s.encode())
    assert take_string(own.pll_utree_export_newick(again.contents.vroot, None)) == s
    own.pll_utree_destroy(again, None)
    own.pll_utree_destroy(tree, None)


@pytest.mark.parametrize("tips,seed", [(2, 1), (5, 2), (64, 3), (333, 4)])
def test_rtree_traversal_and_operations_match_reference(own, ref, tips, seed):
    rng = np.random.default_rng(seed)
    newick = random_newick(rng, tips, rooted=True)
    tree = own.pll_rtree_parse_newick_string(newick.encode())
    assert tree, C.c_char_p.in_dll(own, "pll_errmsg")
    t = tree.contents
    assert t.tip_count == tips and t.inner_count == tips - 1 and t.edge_count == 2 * tips - 2
    n_nodes = 2 * tips - 1
    full = RCB(lambda node: 1)
    partial = RCB(lambda node: 1 if (node.contents.clv_index * 2654435761) % 5 else 0)

    def run(dll, traversal, cb):
        buf = (C.POINTER(RNode) * n_nodes)()
        size = C.c_uint(0)
        rc = dll.pll_rtree_traverse(t.root, traversal, cb, buf, C.byref(size))
        order = [C.addressof(buf[i].contents) for i in range(size.value)]
        ops = (Operation * n_nodes)()
        branches = (C.c_double * n_nodes)()
        pm = (C.c_uint * n_nodes)()
        n_mat, n_ops = C.c_uint(0), C.c_uint(0)
        if size.value and cb is full:
            dll.pll_rtree_create_operations(buf, size.value, branches, pm, ops, C.byref(n_mat), C.byref(n_ops))
        return rc, order, bytes(ops)[:n_ops.value * C.sizeof(Operation)], list(branches)[:n_mat.value], list(pm)[:n_mat.value]

    for traversal in (1, 2):
        for cb in (full, partial):
            assert run(ref, traversal, cb) == run(own, traversal, cb)
    assert take_string(ref.pll_rtree_export_newick(t.root, None)) == take_string(own.pll_rtree_export_newick(t.root, None))
    rc, order, ops, branches, pm = run(own, 1, full)
    assert len(order) == n_nodes and len(ops) == (tips - 1) * C.sizeof(Operation) and len(branches) == n_nodes - 1
    # index template (src/parse_rtree.y:167-231): tips 0..n-1 in reading order, inner nodes in post-order
    assert sorted(t.nodes[i].contents.clv_index for i in range(n_nodes)) == list(range(n_nodes))
    assert all(t.nodes[i].contents.clv_index == i for i in range(n_nodes))
    own.pll_rtree_destroy(tree, None)


def test_newick_syntax_errors(own):
    for bad in ("(A,B,(C,D);", "(A,B,(C,D)));", "(A,B,,C);", "A;", "(A:x,B,C);", "(A,B,C)"):
        assert not own.pll_utree_parse_newick_string(bad.encode()), bad
        assert C.c_int.in_dll(own, "pll_errno").value == 111
    assert not own.pll_rtree_parse_newick_string(b"(A,B,C);")
    assert not own.pll_utree_parse_newick(b"/nonexistent/file.tree")
    assert C.c_int.in_dll(own, "pll_errno").value == 100


def test_rtree_unroot(own):
    own.pll_rtree_unroot.restype, own.pll_rtree_unroot.argtypes = C.POINTER(UTree), [C.POINTER(RTree)]
    own.pll_utree_reset_template_indices.argtypes = [C.POINTER(UNode), C.c_uint]
    cases = {
        "((A:1,B:2)ab:3,(C:4,D:5)cd:6);": "((C:4.000000,D:5.000000)cd:9.000000,A:1.000000,B:2.000000)ab;",
        "(A:1,((B:2,C:3)x:4,D:5)y:6);": "(A:7.000000,(B:2.000000,C:3.000000)x:4.000000,D:5.000000)y;",
    }
    for rooted, unrooted in cases.items():
        rt = own.pll_rtree_parse_newick_string(rooted.encode())
        ut = own.pll_rtree_unroot(rt)
        assert ut
        u = ut.contents
        own.pll_utree_reset_template_indices(u.vroot, u.tip_count)
        assert own.pll_utree_check_integrity(ut) == 1 and not own.pll_utree_is_rooted(ut)
        assert (u.tip_count, u.inner_count, u.binary) == (4, 2, 1)
        assert take_string(own.pll_utree_export_newick(u.vroot, None)) == unrooted
        own.pll_utree_destroy(ut, None)
        own.pll_rtree_destroy(rt, None)
    rt = own.pll_rtree_parse_newick_string(b"(A:1,B:2);")
    assert not own.pll_rtree_unroot(rt) and C.c_int.in_dll(own, "pll_errno").value == 116
    own.pll_rtree_destroy(rt, None)


# ---- copies, rooted drawing, lists (appended helpers of pll_tree.c) --------------------------------------

def test_utree_clone_is_an_independent_identical_tree(own, ref):
    rng = np.random.default_rng(12)
    for tips, multi in ((5, 0.0), (40, 0.0), (30, 0.5)):
        newick = random_newick(rng, tips, multifurcate=multi)
        tree = own.pll_utree_parse_newick_string(newick.encode())
        own.pll_utree_clone.restype, own.pll_utree_clone.argtypes = C.POINTER(UTree), [C.POINTER(UTree)]
        copy = own.pll_utree_clone(tree)
        assert copy
        a, b = tree.contents, copy.contents
        assert (a.tip_count, a.inner_count, a.edge_count, a.binary) == (b.tip_count, b.inner_count, b.edge_count, b.binary)
        n = a.tip_count + a.inner_count
        seen = set()
        for i in range(n):
            x, y = a.nodes[i].contents, b.nodes[i].contents
            assert C.addressof(x) != C.addressof(y)
            assert (x.label, x.length, x.node_index, x.clv_index, x.scaler_index, x.pmatrix_index) == \
                   (y.label, y.length, y.node_index, y.clv_index, y.scaler_index, y.pmatrix_index)
            seen.add(C.addressof(y))
        text = lambda t: C.string_at(own.pll_utree_export_newick(t.contents.nodes[n - 1], None))
        assert text(tree) == text(copy)
        # the reference's own clone of the same tree exports the same text
        ref.pll_utree_graph_clone.restype, ref.pll_utree_graph_clone.argtypes = C.POINTER(UNode), [C.POINTER(UNode)]
        rcopy = ref.pll_utree_graph_clone(a.nodes[n - 1])
        assert C.string_at(ref.pll_utree_export_newick(rcopy, None)) == text(tree)
        own.pll_utree_check_integrity.restype, own.pll_utree_check_integrity.argtypes = C.c_int, [C.POINTER(UTree)]
        assert own.pll_utree_check_integrity(copy) == 1
        own.pll_utree_destroy(tree, None)
        assert text(copy)  # still readable after the original is gone
        own.pll_utree_destroy(copy, None)


def test_rtree_show_ascii_matches_reference(own, ref, capfd):
    rng = np.random.default_rng(4)
    libc.fflush.argtypes = [C.c_void_p]
    for tips in (2, 3, 9, 31):
        newick = random_newick(rng, tips, rooted=True)
        tree = own.pll_rtree_parse_newick_string(newick.encode())
        outs = []
        for dll in (ref, own):
            dll.pll_rtree_show_ascii.restype, dll.pll_rtree_show_ascii.argtypes = None, [C.POINTER(RNode), C.c_int]
            for options in (1, 1 | 2 | 4, 31):
                dll.pll_rtree_show_ascii(tree.contents.root, options)
            libc.fflush(None)
            outs.append(capfd.readouterr().out)
        assert outs[0] == outs[1] and outs[1].count("+---") >= tips
        own.pll_rtree_destroy(tree, None)


def test_dlist_append_prepend_remove(own):
    class DList(C.Structure):
        pass

    DList._fields_ = [("next", C.POINTER(DList)), ("prev", C.POINTER(DList)), ("data", C.c_void_p)]
    for name in ("pll_dlist_append", "pll_dlist_prepend", "pll_dlist_remove"):
        f = getattr(own, name)
        f.restype, f.argtypes = C.c_int, [C.POINTER(C.POINTER(DList)), C.c_void_p]

    def items(head):
        out, prev = [], None
        while head:
            assert (C.addressof(head.contents.prev.contents) if head.contents.prev else None) == prev
            out.append(head.contents.data)
            prev = C.addressof(head.contents)
            head = head.contents.next
        return out

    head = C.POINTER(DList)()
    for v in (1, 2, 3):
        assert own.pll_dlist_append(C.byref(head), v) == 1
    assert items(head) == [1, 2, 3]
    assert own.pll_dlist_prepend(C.byref(head), 9) == 1  # right after the first element, as in the reference
    assert items(head) == [1, 9, 2, 3]
    assert own.pll_dlist_remove(C.byref(head), 2) == 1
    assert items(head) == [1, 9, 3]
    assert own.pll_dlist_remove(C.byref(head), 1) == 1
    assert items(head) == [9, 3]
    assert own.pll_dlist_remove(C.byref(head), 77) == 0
    assert own.pll_dlist_remove(C.byref(head), 3) == 1 and own.pll_dlist_remove(C.byref(head), 9) == 1
    assert not head
