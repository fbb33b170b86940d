"""Site pattern compression on the device (libpll_b200.so: plf_compress.cu) against the UNMODIFIED
reference (src/compress.c:171-410 in oracle/_ref/libpll_ref.so): unique columns, their order, the
weights and the site -> pattern map must be identical bit for bit.  Needs a B200."""
import ctypes as C
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi

pytestmark = pytest.mark.gpu


class Msa(C.Structure):
    _fields_ = [("count", C.c_int), ("length", C.c_int), ("sequence", C.POINTER(C.c_char_p)), ("label", C.POINTER(C.c_char_p))]


def random_alignment(count, length, alphabet, seed, distinct_columns):
    """columns drawn from a pool of `distinct_columns` random columns, so that many repeat"""
    rng = np.random.default_rng(seed)
    pool = rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=(distinct_columns, count))
    cols = pool[rng.integers(0, distinct_columns, size=length)]
    return [bytes(cols[:, t]) for t in range(count)]


def run(lib, seqs, map_name, with_map):
    count, length = len(seqs), len(seqs[0])
    bufs = [C.create_string_buffer(s, length + 1) for s in seqs]
    arr = (C.c_char_p * count)(*[C.cast(b, C.c_char_p) for b in bufs])
    m = lib.map(map_name)
    site_map = np.full(length, 0xFFFFFFFF, dtype=np.uint32)
    if with_map:
        msa = Msa(count, length, C.cast(arr, C.POINTER(C.c_char_p)), None)
        w = lib.pll_compress_site_patterns_msa(C.byref(msa), m, site_map.ctypes.data_as(capi.c_uint_p))
        n = msa.length
    else:
        ln = C.c_int(length)
        w = lib.pll_compress_site_patterns(arr, m, count, C.byref(ln))
        n = ln.value
    if not w:
        return None
    weights = np.ctypeslib.as_array(w, shape=(n,)).copy()
    lib.pll_aligned_free  # weights come from malloc: freed by the C library's free
    C.CDLL(None).free(w)
    return n, [b.raw[:n + 1] for b in bufs], weights, site_map


@pytest.mark.parametrize("map_name,alphabet,count,length,distinct", [
    ("pll_map_nt", b"ACGTRYKMSWBDHVN-acgt", 12, 5000, 700),
    ("pll_map_nt", b"ACGT-", 4, 20000, 150),
    ("pll_map_nt", b"ACGTN", 100, 40000, 30000),
    ("pll_map_aa", b"ARNDCQEGHILKMFPSTWYVBZX-*", 25, 9000, 4000),
    ("pll_map_nt", b"ACGT", 3, 17, 1000),
])
@pytest.mark.parametrize("with_map", [False, True])
def test_compress_site_patterns_parity(reflib, cudalib, map_name, alphabet, count, length, distinct, with_map):
    seqs = random_alignment(count, length, alphabet, seed=count * 7 + length, distinct_columns=distinct)
    a = run(reflib, seqs, map_name, with_map)
    b = run(cudalib, seqs, map_name, with_map)
    assert a is not None and b is not None, cudalib.errmsg
    assert a[0] == b[0], "compressed length"
    assert a[0] <= length
    assert a[1] == b[1], "unique columns (order and characters, terminating zero)"
    assert np.array_equal(a[2], b[2]), "weights"
    assert int(b[2].sum()) == length
    if with_map:
        assert np.array_equal(a[3], b[3]), "site -> pattern map"


def test_compress_rejects_unmapped_character(reflib, cudalib):
    seqs = [b"ACGTACGT", b"ACGTAC!T", b"ACGTACGT"]
    assert run(cudalib, seqs, "pll_map_nt", False) is None
    assert cudalib.errno == 114 and "sequence 2 position 7" in cudalib.errmsg
    assert run(reflib, seqs, "pll_map_nt", False) is None
    assert reflib.errno == 114 and "sequence 2 position 7" in reflib.errmsg


# ---- the reference's own golden file (test/src/compress-patterns.c -> test/out/compress-patterns.out) ----

def odd7_map():
    m = (capi.pll_state_t * 256)()
    for ch in "*-?":
        m[ord(ch)] = 0x3F
    for k, v in enumerate([0x01, 0x02, 0x04, 0x08, 0x0C, 0x10, 0x20]):
        m[ord("A") + k] = v
        m[ord("a") + k] = v
    return m


def golden_blocks():
    import os
    import re

    text = open(os.path.join(os.path.dirname(__file__), "golden", "compress-patterns.out")).read()
    for blk in text.split("* TEST: ")[1:]:
        head = blk.splitlines()[0]
        dt = re.search(r"DATATYPE = (\w+)", head).group(1)
        backmap = "BACKMAP = YES" in head
        orig = re.search(r"ORIGINAL MSA \((\d+)\):\n((?:.+\n)+)", blk).group(2).split()
        comp = re.search(r"COMPRESSED MSA \((\d+)\):\n((?:.+\n)+)", blk).group(2).split()
        weights = [int(x) for x in re.search(r"PATTERN WEIGHTS: ([\d ]+)", blk).group(1).split()]
        smap = re.search(r"SITE-TO-PATTERN MAP: ([\d ]+)", blk)
        yield dt, backmap, orig, comp, weights, ([int(x) for x in smap.group(1).split()] if smap else None)


def run_golden(lib, dt, backmap, orig):
    seqs = [s.encode() for s in orig]
    count, length = len(seqs), len(seqs[0])
    bufs = [C.create_string_buffer(s, length + 1) for s in seqs]
    arr = (C.c_char_p * count)(*[C.cast(b, C.c_char_p) for b in bufs])
    m = {"DNA": lambda: lib.map("pll_map_nt"), "AA": lambda: lib.map("pll_map_aa"), "ODD7": odd7_map}[dt]()
    site_map = np.zeros(length, dtype=np.uint32)
    msa = Msa(count, length, C.cast(arr, C.POINTER(C.c_char_p)), None)
    w = lib.pll_compress_site_patterns_msa(C.byref(msa), m, site_map.ctypes.data_as(capi.c_uint_p) if backmap else None)
    assert w, lib.errmsg
    n = msa.length
    return [b.raw[:n].decode() for b in bufs], list(np.ctypeslib.as_array(w, shape=(n,))), list(site_map)


def test_compress_reference_golden_file(cudalib):
    n = 0
    for dt, backmap, orig, comp, weights, smap in golden_blocks():
        got_comp, got_w, got_map = run_golden(cudalib, dt, backmap, orig)
        assert got_comp == comp and got_w == weights, (dt, backmap)
        if smap is not None:
            assert got_map == smap, (dt, backmap)
        n += 1
    assert n == 6
