"""FASTA reader (pll_fasta.c) against the UNMODIFIED reference's src/fasta.c (oracle/_ref) on the same files:
records, header/sequence lengths, sequence numbers, stripped-character statistics, pll_fasta_load, errors."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]


class Fasta(C.Structure):
    _fields_ = [("fp", C.c_void_p), ("line", C.c_char * 2048), ("chrstatus", C.POINTER(C.c_uint)), ("no", C.c_long),
                ("filesize", C.c_long), ("lineno", C.c_long), ("stripped_count", C.c_long), ("stripped", C.c_long * 256)]


class Msa(C.Structure):
    _fields_ = [("count", C.c_int), ("length", C.c_int), ("sequence", C.POINTER(C.c_char_p)), ("label", C.POINTER(C.c_char_p))]


def bind(path):
    dll = C.CDLL(path)
    dll.pll_fasta_open.restype, dll.pll_fasta_open.argtypes = C.POINTER(Fasta), [C.c_char_p, C.c_void_p]
    dll.pll_fasta_getnext.restype = C.c_int
    dll.pll_fasta_getnext.argtypes = [C.POINTER(Fasta), C.POINTER(C.c_void_p), C.POINTER(C.c_long), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_long), C.POINTER(C.c_long)]
    dll.pll_fasta_close.argtypes = [C.POINTER(Fasta)]
    dll.pll_fasta_rewind.argtypes = [C.POINTER(Fasta)]
    dll.pll_fasta_getfilesize.restype, dll.pll_fasta_getfilesize.argtypes = C.c_long, [C.POINTER(Fasta)]
    dll.pll_fasta_load.restype, dll.pll_fasta_load.argtypes = C.POINTER(Msa), [C.c_char_p]
    dll.pll_msa_destroy.argtypes = [C.POINTER(Msa)]
    return dll


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(pkg.REF_PATH):
        pytest.skip("oracle/_ref/libpll_ref.so not built (needs /root/reference)")
    return bind(pkg.REF_PATH), bind(pkg.LIB_PATH)


def errno_of(dll):
    return C.c_int.in_dll(dll, "pll_errno").value


def read_all(dll, path, mapname):
    fd = dll.pll_fasta_open(path.encode(), C.addressof((C.c_uint * 256).in_dll(dll, mapname)))
    if not fd:
        return ("open failed", errno_of(dll))
    out = []
    for rounds in range(2):  # second round after a rewind
        recs = []
        while True:
            head, seq = C.c_void_p(), C.c_void_p()
            hl, sl, no = C.c_long(), C.c_long(), C.c_long()
            if not dll.pll_fasta_getnext(fd, C.byref(head), C.byref(hl), C.byref(seq), C.byref(sl), C.byref(no)):
                recs.append(("stop", errno_of(dll), C.c_char_p.in_dll(dll, "pll_errmsg").value if False else None))
                break
            recs.append((C.string_at(head), hl.value, C.string_at(seq), sl.value, no.value))
            libc.free(head)
            libc.free(seq)
        f = fd.contents
        out.append((recs, f.stripped_count, list(f.stripped), f.lineno, dll.pll_fasta_getfilesize(fd)))
        if recs[-1][1] != 102:  # not EOF: an error, the handle stays where it is
            break
        assert dll.pll_fasta_rewind(fd) == 1
    dll.pll_fasta_close(fd)
    return out


def write(tmp_path, name, text):
    p = tmp_path / name
    p.write_bytes(text)
    return str(p)


def test_fasta_records_match_reference(libs, tmp_path):
    ref, own = libs
    rng = np.random.default_rng(3)
    long_seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgt-?.", dtype=np.uint8), size=7000))
    files = {
        "plain.fa": b">t1 first\nACGT-ACGT\nTTGA\n>t2\nacgtNNNN\nAC-T\n>t3 desc here\nACGTACGTACGT\n",
        "crlf.fa": b">t1 x\r\nACGT\r\nAC\r\n>t2\r\nGGTTAA\r\n",
        "stripped.fa": b">a\nAC GT*12#3\n>b\nAC!GT@@\n\n\n>c\n\tAC\x0bGT\n",
        "long.fa": b">" + b"h" * 3000 + b"\n" + long_seq + b"\n>second\n" + long_seq[:100] + b"\n",
        "noeol.fa": b">x\nACGT\n>y\nAC",
        "badheader.fa": b"ACGT\n>x\nAC\n",
        "fatal.fa": b">x\nAC\x01GT\n",
    }
    for name, text in files.items():
        path = write(tmp_path, name, text)
        for mapname in ("pll_map_fasta", "pll_map_generic"):
            assert read_all(ref, path, mapname) == read_all(own, path, mapname), (name, mapname)
    # bytes >= 0x80 index the reference's table with a negative (signed char) subscript: undefined there;
    # here they are looked up as unsigned, i.e. what the tables say (generic: 0xff is fatal)
    path = write(tmp_path, "fatal8bit.fa", b">x\nAC\xffGT\n")
    assert read_all(own, path, "pll_map_generic")[0][0] == [("stop", 202, None)]
    assert read_all(own, path, "pll_map_fasta")[0][0][0][2] == b"ACGT"
    assert read_all(own, str(tmp_path / "missing.fa"), "pll_map_fasta") == ("open failed", 100)
    assert read_all(own, write(tmp_path, "empty.fa", b""), "pll_map_fasta") == read_all(ref, str(tmp_path / "empty.fa"), "pll_map_fasta")


def test_fasta_load_matches_reference(libs, tmp_path):
    ref, own = libs
    good = write(tmp_path, "aln.fa", b">s1\nACGTAC\nGT\n>s2 label\nAC-TACGT\n>s3\nNNNNNNNN\n")
    ragged = write(tmp_path, "ragged.fa", b">s1\nACGT\n>s2\nACG\n")
    for dll in (ref, own):
        msa = dll.pll_fasta_load(good.encode())
        assert msa
        m = msa.contents
        assert (m.count, m.length) == (3, 8)
        assert [m.sequence[i] for i in range(3)] == [b"ACGTACGT", b"AC-TACGT", b"NNNNNNNN"]
        assert [m.label[i] for i in range(3)] == [b"s1", b"s2 label", b"s3"]
        dll.pll_msa_destroy(msa)
        assert not dll.pll_fasta_load(ragged.encode())
        assert errno_of(dll) == 204
