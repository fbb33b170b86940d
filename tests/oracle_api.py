"""ctypes prototypes of oracle/libplf_oracle.so (TEST INFRASTRUCTURE)."""
import ctypes as C

import numpy as np

dp = C.POINTER(C.c_double)
up = C.POINTER(C.c_uint)
ubp = C.POINTER(C.c_ubyte)
ip = C.POINTER(C.c_int)
sp_ = C.POINTER(C.c_ulonglong)
dpp = C.POINTER(dp)
U, I, D = C.c_uint, C.c_int, C.c_double

PROTOS = {
    "orc_update_pmatrix": (None, [dpp, U, U, U, dp, dp, up, up, dp, dpp, dpp, dpp, U]),
    "orc_update_partial_ii": (None, [U, U, U, U, dp, up, dp, dp, dp, dp, up, up, I]),
    "orc_update_partial_ti": (None, [U, U, U, U, dp, up, ubp, dp, dp, dp, up, sp_, U, I]),
    "orc_update_partial_tt": (None, [U, U, U, U, dp, up, ubp, ubp, dp, dp, sp_, U, I]),
    "orc_update_partial_repeats": (None, [U, U, U, U, dp, up, dp, dp, dp, dp, up, up, up, up, up, I]),
    "orc_root_loglikelihood": (D, [U, U, U, U, dp, up, up, dpp, dp, up, dp, ip, up, dp]),
    "orc_edge_loglikelihood_ii": (D, [U, U, U, U, dp, up, up, dp, up, up, dp, dpp, dp, up, dp, ip, up, dp, I]),
    "orc_edge_loglikelihood_ti": (D, [U, U, U, U, dp, up, ubp, sp_, dp, dpp, dp, up, dp, ip, up, dp, I]),
    "orc_update_sumtable_ii": (None, [U, U, U, U, dp, up, dp, up, up, up, dpp, dpp, dpp, dp, I]),
    "orc_update_sumtable_ti": (None, [U, U, U, U, dp, ubp, sp_, up, dpp, dpp, dpp, dp, I]),
    "orc_likelihood_derivatives": (None, [U, U, U, U, dp, ip, up, D, dp, dpp, dp, dpp, dp, dp, dp]),
    "orc_update_repeats": (U, [U, up, U, up, U, up, up, up, U]),
    "orc_update_repeats_tip": (U, [U, sp_, C.c_char_p, up, up]),
}


def load(path):
    lib = C.CDLL(path)
    for name, (res, args) in PROTOS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def D_(a):
    return None if a is None else a.ctypes.data_as(dp)


def U_(a):
    return None if a is None else a.ctypes.data_as(up)


def B_(a):
    return None if a is None else a.ctypes.data_as(ubp)


def I_(a):
    return None if a is None else a.ctypes.data_as(ip)


def S_(a):
    return None if a is None else a.ctypes.data_as(sp_)


def ptr_array(arrays):
    """double** from a list of numpy arrays (kept alive by the caller)."""
    arr = (dp * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data_as(dp)
    return arr
