"""PHYLIP reader (pll_phylip.c) against the UNMODIFIED reference's src/phylip.c (oracle/_ref) on the same files:
sequential and interleaved layouts, CRLF line ends, blank lines, labels glued to data by a tab, long lines,
stripped-character statistics, rewind, pll_phylip_load, and every syntax error the reference reports (same
pll_errno and message)."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")


class Phylip(C.Structure):
    _fields_ = [("fp", C.c_void_p), ("line", C.c_void_p), ("line_size", C.c_size_t), ("line_maxsize", C.c_size_t),
                ("buffer", C.c_char * 2048), ("chrstatus", C.POINTER(C.c_uint)), ("no", C.c_long),
                ("filesize", C.c_long), ("lineno", C.c_long), ("stripped_count", C.c_long), ("stripped", C.c_long * 256)]


class Msa(C.Structure):
    _fields_ = [("count", C.c_int), ("length", C.c_int), ("sequence", C.POINTER(C.c_char_p)), ("label", C.POINTER(C.c_char_p))]


def bind(path):
    dll = C.CDLL(path)
    dll.pll_phylip_open.restype, dll.pll_phylip_open.argtypes = C.POINTER(Phylip), [C.c_char_p, C.c_void_p]
    dll.pll_phylip_close.argtypes = [C.POINTER(Phylip)]
    dll.pll_phylip_rewind.restype, dll.pll_phylip_rewind.argtypes = C.c_int, [C.POINTER(Phylip)]
    for f in (dll.pll_phylip_parse_sequential, dll.pll_phylip_parse_interleaved):
        f.restype, f.argtypes = C.POINTER(Msa), [C.POINTER(Phylip)]
    dll.pll_phylip_load.restype, dll.pll_phylip_load.argtypes = C.POINTER(Msa), [C.c_char_p, C.c_int]
    dll.pll_msa_destroy.argtypes = [C.POINTER(Msa)]
    return dll


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(pkg.REF_PATH):
        pytest.skip("oracle/_ref/libpll_ref.so not built (needs /root/reference)")
    return bind(pkg.REF_PATH), bind(pkg.LIB_PATH)


def status(dll):
    return C.c_int.in_dll(dll, "pll_errno").value, (C.c_char * 200).in_dll(dll, "pll_errmsg").value


def clear(dll):
    C.c_int.in_dll(dll, "pll_errno").value = 0
    (C.c_char * 200).in_dll(dll, "pll_errmsg").value = b""


def msa_tuple(dll, msa):
    if not msa:
        return ("failed",) + status(dll)
    m = msa.contents
    out = (m.count, m.length, [m.label[i] for i in range(m.count)], [m.sequence[i] for i in range(m.count)])
    dll.pll_msa_destroy(msa)
    return out


def parse(dll, path, interleaved, mapname="pll_map_phylip", twice=False):
    clear(dll)
    fd = dll.pll_phylip_open(str(path).encode(), C.addressof((C.c_uint * 256).in_dll(dll, mapname)))
    if not fd:
        return ("open failed", status(dll)[0])
    fn = dll.pll_phylip_parse_interleaved if interleaved else dll.pll_phylip_parse_sequential
    res = [msa_tuple(dll, fn(fd))]
    res.append((fd.contents.stripped_count, list(fd.contents.stripped), fd.contents.filesize, fd.contents.lineno))
    if twice:
        assert dll.pll_phylip_rewind(fd) == 1
        res.append(msa_tuple(dll, fn(fd)))
        res.append(fd.contents.stripped_count)
    dll.pll_phylip_close(fd)
    return res


def write_seq(path, labels, seqs, width=60, eol="\n", sep="  ", blank_between=False):
    with open(path, "w", newline="") as f:
        f.write(f" {len(seqs)} {len(seqs[0])}{eol}")
        for lab, s in zip(labels, seqs):
            chunks = [s[k:k + width] for k in range(0, len(s), width)]
            f.write(lab + sep + chunks[0] + eol)
            for c in chunks[1:]:
                f.write(c + eol)
            if blank_between:
                f.write(eol)


def write_int(path, labels, seqs, width=50, eol="\n", gaps=True):
    n = len(seqs[0])
    with open(path, "w", newline="") as f:
        f.write(f"{len(seqs)} {n}{eol}")
        for k in range(0, n, width):
            for lab, s in zip(labels, seqs):
                chunk = s[k:k + width]
                if gaps:
                    chunk = " ".join(chunk[j:j + 10] for j in range(0, len(chunk), 10))
                f.write((lab.ljust(12) if k == 0 else "") + chunk + eol)
            f.write(eol)


def random_alignment(rng, taxa, sites):
    alpha = np.frombuffer(b"ACGTNRY-?acgt", dtype=np.uint8)
    return ["".join(map(chr, alpha[rng.integers(0, len(alpha), size=sites)])) for _ in range(taxa)]


@pytest.mark.parametrize("taxa,sites,width,eol", [(5, 37, 60, "\n"), (7, 301, 60, "\n"), (4, 5000, 5000, "\n"),
                                                  (6, 250, 40, "\r\n"), (3, 1, 10, "\n")])
def test_sequential_files_parse_like_the_reference(libs, tmp_path, taxa, sites, width, eol):
    ref, own = libs
    rng = np.random.default_rng(taxa * sites)
    seqs = random_alignment(rng, taxa, sites)
    labels = [f"taxon_{i}" for i in range(taxa)]
    path = tmp_path / "seq.phy"
    write_seq(path, labels, seqs, width=width, eol=eol, blank_between=(taxa == 7))
    a, b = parse(ref, path, False), parse(own, path, False, twice=True)
    assert a == b[:2]
    assert b[2] == b[0] and b[3] == b[1][0]  # after a rewind (the reference's rewind crashes once EOF was reached)
    assert b[0][2] == [x.encode() for x in labels] and b[0][3] == [x.encode() for x in seqs]


@pytest.mark.parametrize("taxa,sites,width,eol,gaps", [(5, 37, 50, "\n", True), (8, 733, 50, "\n", True),
                                                       (4, 120, 60, "\r\n", False), (3, 4100, 4100, "\n", False)])
def test_interleaved_files_parse_like_the_reference(libs, tmp_path, taxa, sites, width, eol, gaps):
    ref, own = libs
    rng = np.random.default_rng(taxa + sites)
    seqs = random_alignment(rng, taxa, sites)
    labels = [f"sp{i}" for i in range(taxa)]
    path = tmp_path / "int.phy"
    write_int(path, labels, seqs, width=width, eol=eol, gaps=gaps)
    a, b = parse(ref, path, True), parse(own, path, True, twice=True)
    assert a == b[:2]
    assert b[2] == b[0] and b[3] == b[1][0]
    assert b[0][3] == [x.encode() for x in seqs]
    # blanks inside the data are class 0 of pll_map_phylip: stripped and counted
    assert b[1][0] > 0


def test_load_uses_the_generic_map(libs, tmp_path):
    ref, own = libs
    path = tmp_path / "g.phy"
    path.write_text("3 6\nA  01{}!*\nB  ab#$%^\nC  ......\n")
    for inter in (0, 1):
        for dll in (ref, own):
            clear(dll)
        assert msa_tuple(ref, ref.pll_phylip_load(str(path).encode(), inter)) == \
               msa_tuple(own, own.pll_phylip_load(str(path).encode(), inter))
    assert msa_tuple(own, own.pll_phylip_load(str(path).encode(), 0))[3] == [b"01{}!*", b"ab#$%^", b"......"]


BAD = {
    "no_header_numbers": "taxa sites\nA ACGT\n",
    "one_number": "3\nA ACGT\n",
    "zero_taxa": "0 4\n",
    "header_options": "2 4 I\nA ACGT\nB ACGT\n",
    "too_few_sequences": "3 4\nA ACGT\nB ACGT\n",
    "too_many_sequences": "2 4\nA ACGT\nB ACGT\nC ACGT\n",
    "sequence_too_long": "2 4\nA ACGTA\nB ACGT\n",
    "sequence_too_short": "2 4\nA ACGT\nB ACG\n",
    "illegal_character": "2 4\nA AC\x01T\nB ACGT\n",
    "ragged_block": "2 8\nA ACGT\nB ACG\n\nACGT\nACGTA\n",
    "incomplete_last_block": "3 8\nA ACGT\nB ACGT\nC ACGT\n\nACGT\nACGT\n",
    "short_total": "2 8\nA ACGT\nB ACGT\n\nAC\nAC\n",
    "empty_after_header": "2 4\n",
    "label_only": "2 4\nA\n",
}


@pytest.mark.parametrize("name", sorted(BAD))
@pytest.mark.parametrize("interleaved", [False, True])
def test_malformed_files_fail_like_the_reference(libs, tmp_path, name, interleaved):
    ref, own = libs
    path = tmp_path / (name + ".phy")
    with open(path, "w", newline="") as f:
        f.write(BAD[name])
    a, b = parse(ref, path, interleaved), parse(own, path, interleaved)
    if name == "zero_taxa" or (name == "short_total" and interleaved):
        # the reference leaves pll_errno untouched (and accepts a zero count, then finds no sequences);
        # here both are PLL_ERROR_PHYLIP_SYNTAX
        assert b[0][0] == "failed" and b[0][1] == 231
        return
    assert a[0] == b[0], (a[0], b[0])
    assert a[1] == b[1]


def test_missing_and_empty_files(libs, tmp_path):
    ref, own = libs
    assert parse(ref, tmp_path / "nope.phy", False) == parse(own, tmp_path / "nope.phy", False) == ("open failed", 100)
    empty = tmp_path / "empty.phy"
    empty.write_text("")
    assert parse(own, empty, False)[0] == "open failed"
