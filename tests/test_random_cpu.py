"""pll_random_* (libpll-2_b200/csrc/pll_random.c, host code): same sequences as the UNMODIFIED reference's
src/random.c for every state-buffer size class, across setstate switches, and through the
create/getint convenience layer (which stepwise addition uses to shuffle taxa)."""
import ctypes as C
import importlib

import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi


@pytest.fixture(scope="module")
def own():
    return capi.PllLibrary(pkg.LIB_PATH, cuda=False)


def draw(lib, buf, n):
    out = []
    r = C.c_int()
    for _ in range(n):
        assert lib.pll_random_r(C.byref(buf), C.byref(r)) == 0
        out.append(r.value)
    return out


@pytest.mark.parametrize("nbytes", [8, 32, 64, 128, 256])
@pytest.mark.parametrize("seed", [0, 1, 42, 0xFFFFFFFF])
def test_initstate_sequences_match_reference(own, reflib, nbytes, seed):
    res = []
    for lib in (own, reflib):
        state = C.create_string_buffer(nbytes)
        buf = capi.RandomData()
        assert lib.pll_initstate_r(seed, state, nbytes, C.byref(buf)) == 0
        seq = draw(lib, buf, 500)
        assert lib.pll_srandom_r(seed + 7, C.byref(buf)) == 0
        seq += draw(lib, buf, 100)
        res.append((seq, state.raw))
    assert res[0][0] == res[1][0]
    assert res[0][1] == res[1][1]


def test_too_small_state_is_rejected(own, reflib):
    for lib in (own, reflib):
        state = C.create_string_buffer(4)
        buf = capi.RandomData()
        assert lib.pll_initstate_r(1, state, 4, C.byref(buf)) == -1


def test_setstate_switches_match_reference(own, reflib):
    res = []
    for lib in (own, reflib):
        a, b = C.create_string_buffer(128), C.create_string_buffer(32)
        buf = capi.RandomData()
        lib.pll_initstate_r(5, a, 128, C.byref(buf))
        seq = draw(lib, buf, 17)
        lib.pll_initstate_r(9, b, 32, C.byref(buf))
        seq += draw(lib, buf, 11)
        assert lib.pll_setstate_r(a, C.byref(buf)) == 0
        seq += draw(lib, buf, 40)
        assert lib.pll_setstate_r(b, C.byref(buf)) == 0
        seq += draw(lib, buf, 40)
        res.append((seq, a.raw, b.raw))
    assert res[0] == res[1]


@pytest.mark.parametrize("seed", [1, 12345])
def test_getint_matches_reference(own, reflib, seed):
    res = []
    for lib in (own, reflib):
        rs = lib.pll_random_create(seed)
        res.append([lib.pll_random_getint(rs, m) for m in (2, 3, 10, 1000, 2 ** 31 - 1) * 40])
        lib.pll_random_destroy(rs)
    assert res[0] == res[1]
