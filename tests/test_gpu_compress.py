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
