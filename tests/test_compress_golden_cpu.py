"""Pins the checker of tests/test_gpu_compress.py: the reference build (oracle/_ref) reproduces the
reference's own golden file test/out/compress-patterns.out (copied to tests/golden/), and the parser
of that file finds all six blocks."""
import test_gpu_compress as t


def test_reference_build_reproduces_compress_golden(reflib):
    n = 0
    for dt, backmap, orig, comp, weights, smap in t.golden_blocks():
        got_comp, got_w, got_map = t.run_golden(reflib, dt, backmap, orig)
        assert got_comp == comp and got_w == weights, (dt, backmap)
        if smap is not None:
            assert got_map == smap
        n += 1
    assert n == 6
