"""ctypes view of the libpll-2 C API (include/pll_b200.h).

The same binding works for any shared library that exports the reference's
partition API: our CUDA engine (``libpll_b200.so``) and, in tests, the
unmodified reference built as ``oracle/_ref/libpll_ref.so``.  That symmetry is
the drop-in property: parity tests drive both libraries with identical calls.

Struct layouts follow ``/root/reference/src/pll.h:241-335``.
"""
from __future__ import annotations

import ctypes as C
import os

c_uint_p = C.POINTER(C.c_uint)
c_double_p = C.POINTER(C.c_double)
c_ubyte_p = C.POINTER(C.c_ubyte)
pll_state_t = C.c_ulonglong

# attribute bits (include/pll_b200.h)
ARCH_CPU = 0
ARCH_SSE = 1 << 0
ARCH_AVX = 1 << 1
ARCH_AVX2 = 1 << 2
PATTERN_TIP = 1 << 4
AB_LEWIS = 1 << 5
AB_FELSENSTEIN = 2 << 5
AB_STAMATAKIS = 3 << 5
AB_MASK = 7 << 5
AB_FLAG = 1 << 8
RATE_SCALERS = 1 << 9
SITE_REPEATS = 1 << 10
ARCH_CUDA = 1 << 16
SCALE_BUFFER_NONE = -1


class Repeats(C.Structure):
    pass


class Partition(C.Structure):
    _fields_ = [
        ("tips", C.c_uint),
        ("clv_buffers", C.c_uint),
        ("nodes", C.c_uint),
        ("states", C.c_uint),
        ("sites", C.c_uint),
        ("pattern_weight_sum", C.c_uint),
        ("rate_matrices", C.c_uint),
        ("prob_matrices", C.c_uint),
        ("rate_cats", C.c_uint),
        ("scale_buffers", C.c_uint),
        ("attributes", C.c_uint),
        ("alignment", C.c_size_t),
        ("states_padded", C.c_uint),
        ("clv", C.POINTER(c_double_p)),
        ("pmatrix", C.POINTER(c_double_p)),
        ("rates", c_double_p),
        ("rate_weights", c_double_p),
        ("subst_params", C.POINTER(c_double_p)),
        ("scale_buffer", C.POINTER(c_uint_p)),
        ("frequencies", C.POINTER(c_double_p)),
        ("prop_invar", c_double_p),
        ("invariant", C.POINTER(C.c_int)),
        ("pattern_weights", c_uint_p),
        ("eigen_decomp_valid", C.POINTER(C.c_int)),
        ("eigenvecs", C.POINTER(c_double_p)),
        ("inv_eigenvecs", C.POINTER(c_double_p)),
        ("eigenvals", C.POINTER(c_double_p)),
        ("maxstates", C.c_uint),
        ("tipchars", C.POINTER(c_ubyte_p)),
        ("charmap", c_ubyte_p),
        ("ttlookup", c_double_p),
        ("tipmap", C.POINTER(pll_state_t)),
        ("asc_bias_alloc", C.c_int),
        ("asc_additional_sites", C.c_int),
        ("repeats", C.POINTER(Repeats)),
    ]


Repeats._fields_ = [
    ("pernode_site_id", C.POINTER(c_uint_p)),
    ("pernode_id_site", C.POINTER(c_uint_p)),
    ("pernode_ids", c_uint_p),
    ("perscale_ids", c_uint_p),
    ("pernode_allocated_clvs", c_uint_p),
    ("enable_repeats", C.c_void_p),
    ("reallocate_repeats", C.c_void_p),
    ("lookup_buffer", c_uint_p),
    ("toclean_buffer", c_uint_p),
    ("id_site_buffer", c_uint_p),
    ("bclv_buffer", c_double_p),
    ("lookup_buffer_size", C.c_uint),
    ("charmap", C.c_char_p),
]


class Operation(C.Structure):
    _fields_ = [
        ("parent_clv_index", C.c_uint),
        ("parent_scaler_index", C.c_int),
        ("child1_clv_index", C.c_uint),
        ("child1_matrix_index", C.c_uint),
        ("child1_scaler_index", C.c_int),
        ("child2_clv_index", C.c_uint),
        ("child2_matrix_index", C.c_uint),
        ("child2_scaler_index", C.c_int),
    ]


PartitionP = C.POINTER(Partition)


class Parsimony(C.Structure):
    """pll_parsimony_t (src/pll.h:467-492)"""
    _fields_ = [
        ("tips", C.c_uint),
        ("inner_nodes", C.c_uint),
        ("sites", C.c_uint),
        ("states", C.c_uint),
        ("attributes", C.c_uint),
        ("alignment", C.c_size_t),
        ("packedvector", C.POINTER(c_uint_p)),
        ("node_cost", c_uint_p),
        ("packedvector_count", C.c_uint),
        ("const_cost", C.c_uint),
        ("informative", C.POINTER(C.c_int)),
        ("informative_count", C.c_uint),
        ("score_buffers", C.c_uint),
        ("ancestral_buffers", C.c_uint),
        ("score_matrix", c_double_p),
        ("sbuffer", C.POINTER(c_double_p)),
        ("anc_states", C.POINTER(c_uint_p)),
    ]


class ParsBuildOp(C.Structure):
    _fields_ = [("parent_score_index", C.c_uint), ("child1_score_index", C.c_uint), ("child2_score_index", C.c_uint)]


class UNode(C.Structure):
    """pll_unode_t (src/pll.h:388-400)"""


UNode._fields_ = [("label", C.c_char_p), ("length", C.c_double), ("node_index", C.c_uint), ("clv_index", C.c_uint),
                  ("scaler_index", C.c_int), ("pmatrix_index", C.c_uint), ("next", C.POINTER(UNode)),
                  ("back", C.POINTER(UNode)), ("data", C.c_void_p)]


class UTree(C.Structure):
    _fields_ = [("tip_count", C.c_uint), ("inner_count", C.c_uint), ("edge_count", C.c_uint), ("binary", C.c_int),
                ("nodes", C.POINTER(C.POINTER(UNode))), ("vroot", C.POINTER(UNode))]


class RandomData(C.Structure):
    """struct pll_random_data (src/pll.h:534-543)"""
    _fields_ = [("fptr", C.POINTER(C.c_int)), ("rptr", C.POINTER(C.c_int)), ("state", C.POINTER(C.c_int)),
                ("rand_type", C.c_int), ("rand_deg", C.c_int), ("rand_sep", C.c_int), ("end_ptr", C.POINTER(C.c_int))]


ParsimonyP = C.POINTER(Parsimony)


class ParsRecOp(C.Structure):
    _fields_ = [("node_score_index", C.c_uint), ("node_ancestral_index", C.c_uint), ("parent_score_index", C.c_uint),
                ("parent_ancestral_index", C.c_uint)]


# Fitch parsimony and the random_r family: the same names in the reference build and in ours
_PARS_PROTOS = {
    "pll_parsimony_create": (ParsimonyP, [C.c_uint, C.c_uint, C.c_uint, c_double_p, C.c_uint, C.c_uint]),
    "pll_set_parsimony_sequence": (C.c_int, [ParsimonyP, C.c_uint, C.POINTER(pll_state_t), C.c_char_p]),
    "pll_parsimony_build": (C.c_double, [ParsimonyP, C.POINTER(ParsBuildOp), C.c_uint]),
    "pll_parsimony_score": (C.c_double, [ParsimonyP, C.c_uint]),
    "pll_parsimony_reconstruct": (None, [ParsimonyP, C.POINTER(pll_state_t), C.POINTER(ParsRecOp), C.c_uint]),
    "pll_fastparsimony_init": (ParsimonyP, [PartitionP]),
    "pll_fastparsimony_update_vectors": (None, [ParsimonyP, C.POINTER(ParsBuildOp), C.c_uint]),
    "pll_fastparsimony_edge_score": (C.c_uint, [ParsimonyP, C.c_uint, C.c_uint]),
    "pll_fastparsimony_root_score": (C.c_uint, [ParsimonyP, C.c_uint]),
    "pll_parsimony_destroy": (None, [ParsimonyP]),
    "pll_fastparsimony_stepwise": (
        C.POINTER(UTree), [C.POINTER(ParsimonyP), C.POINTER(C.c_char_p), c_uint_p, C.c_uint, C.c_uint]),
    "pll_fastparsimony_stepwise_extend": (
        C.c_int, [C.POINTER(UTree), C.POINTER(ParsimonyP), C.c_uint, C.POINTER(C.c_char_p), c_uint_p, C.c_uint, c_uint_p]),
    "pll_fastparsimony_stepwise_spr_round": (
        C.c_int, [C.POINTER(UTree), C.POINTER(ParsimonyP), C.c_uint, c_uint_p, C.c_uint, C.POINTER(C.c_int), c_uint_p]),
    "pll_utree_create_pars_buildops": (
        None, [C.POINTER(C.POINTER(UNode)), C.c_uint, C.POINTER(ParsBuildOp), c_uint_p]),
    "pll_random_r": (C.c_int, [C.POINTER(RandomData), C.POINTER(C.c_int)]),
    "pll_srandom_r": (C.c_int, [C.c_uint, C.POINTER(RandomData)]),
    "pll_initstate_r": (C.c_int, [C.c_uint, C.c_char_p, C.c_size_t, C.POINTER(RandomData)]),
    "pll_setstate_r": (C.c_int, [C.c_char_p, C.POINTER(RandomData)]),
    "pll_random_create": (C.c_void_p, [C.c_uint]),
    "pll_random_getint": (C.c_int, [C.c_void_p, C.c_int]),
    "pll_random_destroy": (None, [C.c_void_p]),
}

_PROTOS = {
    "pll_partition_create": (PartitionP, [C.c_uint] * 9),
    "pll_partition_destroy": (None, [PartitionP]),
    "pll_set_tip_states": (C.c_int, [PartitionP, C.c_uint, C.POINTER(pll_state_t), C.c_char_p]),
    "pll_set_tip_clv": (C.c_int, [PartitionP, C.c_uint, c_double_p, C.c_int]),
    "pll_set_pattern_weights": (None, [PartitionP, c_uint_p]),
    "pll_set_subst_params": (None, [PartitionP, C.c_uint, c_double_p]),
    "pll_set_frequencies": (None, [PartitionP, C.c_uint, c_double_p]),
    "pll_set_category_rates": (None, [PartitionP, c_double_p]),
    "pll_set_category_weights": (None, [PartitionP, c_double_p]),
    "pll_update_eigen": (C.c_int, [PartitionP, C.c_uint]),
    "pll_update_prob_matrices": (C.c_int, [PartitionP, c_uint_p, c_uint_p, c_double_p, C.c_uint]),
    "pll_update_invariant_sites": (C.c_int, [PartitionP]),
    "pll_update_invariant_sites_proportion": (C.c_int, [PartitionP, C.c_uint, C.c_double]),
    "pll_count_invariant_sites": (C.c_uint, [PartitionP, c_uint_p]),
    "pll_update_partials": (None, [PartitionP, C.POINTER(Operation), C.c_uint]),
    "pll_update_partials_rep": (None, [PartitionP, C.POINTER(Operation), C.c_uint, C.c_uint]),
    "pll_update_repeats": (None, [PartitionP, C.POINTER(Operation)]),
    "pll_compute_root_loglikelihood": (C.c_double, [PartitionP, C.c_uint, C.c_int, c_uint_p, c_double_p]),
    "pll_compute_edge_loglikelihood": (
        C.c_double,
        [PartitionP, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, c_uint_p, c_double_p],
    ),
    "pll_update_sumtable": (C.c_int, [PartitionP, C.c_uint, C.c_uint, C.c_int, C.c_int, c_uint_p, c_double_p]),
    "pll_compute_likelihood_derivatives": (
        C.c_int,
        [PartitionP, C.c_int, C.c_int, C.c_double, c_uint_p, c_double_p, c_double_p, c_double_p],
    ),
    "pll_compute_node_ancestral": (
        C.c_int, [PartitionP, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, c_uint_p, c_double_p]),
    "pll_compress_site_patterns": (c_uint_p, [C.POINTER(C.c_char_p), C.POINTER(pll_state_t), C.c_int, C.POINTER(C.c_int)]),
    "pll_compress_site_patterns_msa": (c_uint_p, [C.c_void_p, C.POINTER(pll_state_t), c_uint_p]),
    "pll_set_asc_bias_type": (C.c_int, [PartitionP, C.c_int]),
    "pll_set_asc_state_weights": (None, [PartitionP, c_uint_p]),
    "pll_repeats_enabled": (C.c_int, [PartitionP]),
    "pll_get_sites_number": (C.c_uint, [PartitionP, C.c_uint]),
    "pll_get_clv_size": (C.c_uint, [PartitionP, C.c_uint]),
    "pll_get_site_id": (c_uint_p, [PartitionP, C.c_uint]),
    "pll_get_id_site": (c_uint_p, [PartitionP, C.c_uint]),
    "pll_resize_repeats_lookup": (None, [PartitionP, C.c_uint]),
    "pll_disable_bclv": (None, [PartitionP]),
    "pll_aligned_alloc": (C.c_void_p, [C.c_size_t, C.c_size_t]),
    "pll_aligned_free": (None, [C.c_void_p]),
    "pll_hardware_probe": (C.c_int, []),
}

# additive CUDA surface; absent from the reference library
_CUDA_PROTOS = {
    "pll_cuda_device_count": (C.c_int, []),
    "pll_cuda_set_device": (C.c_int, [C.c_int]),
    "pll_cuda_get_device": (C.c_int, [PartitionP]),
    "pll_cuda_get_stream": (C.c_void_p, [PartitionP]),
    "pll_cuda_synchronize": (C.c_int, [PartitionP]),
    "pll_cuda_download_clv": (C.c_int, [PartitionP, C.c_uint, c_double_p]),
    "pll_cuda_download_scaler": (C.c_int, [PartitionP, C.c_uint, c_uint_p]),
    "pll_cuda_download_pmatrix": (C.c_int, [PartitionP, C.c_uint, c_double_p]),
    "pll_cuda_upload_pmatrix": (C.c_int, [PartitionP, C.c_uint, c_double_p]),
    "pll_cuda_download_sumtable": (C.c_int, [PartitionP, c_double_p, c_double_p]),
    "pll_cuda_scaler_size": (C.c_uint, [PartitionP, C.c_uint]),
    "pll_cuda_count_launch_runs": (C.c_uint, [c_uint_p, C.c_uint, c_uint_p]),
    "pll_cuda_schedule_paths": (C.c_uint, [C.POINTER(Operation), C.c_uint, C.c_uint, C.c_uint, c_uint_p, C.POINTER(C.c_int)]),
    "pll_cuda_invalidate_repeat_identifiers": (C.c_int, [PartitionP]),
    "pll_cuda_check_guards": (C.c_int, [PartitionP]),
    "pll_cuda_debug_overrun": (C.c_int, [PartitionP, C.c_uint]),
    "pll_cuda_host_tipchars": (c_ubyte_p, [PartitionP, C.c_uint]),
    "pll_cuda_peer_group_create": (C.c_void_p, [C.c_int, C.c_uint, C.c_uint, C.c_void_p]),
    "pll_cuda_peer_group_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pll_cuda_peer_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]),
    "pll_cuda_peer_group_check": (C.c_int, [C.c_void_p]),
    "pll_cuda_peer_group_destroy": (None, [C.c_void_p]),
    "pll_cuda_virtual_cherries": (C.c_int, [PartitionP]),
    "pll_cuda_virtual_clvs": (C.c_uint, [PartitionP, C.c_uint]),
    "pll_cuda_materialize_clv": (C.c_int, [PartitionP, C.c_uint]),
    "pll_cuda_edge_loglikelihood_async": (
        C.c_int,
        [PartitionP, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, c_uint_p, C.c_void_p],
    ),
    "pll_cuda_root_loglikelihood_async": (C.c_int, [PartitionP, C.c_uint, C.c_int, c_uint_p, C.c_void_p]),
    "pll_cuda_likelihood_derivatives_async": (
        C.c_int,
        [PartitionP, C.c_int, C.c_int, C.c_double, c_uint_p, c_double_p, C.c_void_p],
    ),
    "pll_cuda_newton_branch": (
        C.c_int,
        [PartitionP, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint, c_uint_p, c_double_p,
         c_double_p, c_double_p, c_double_p, c_uint_p],
    ),
    "pll_cuda_invalidate_host_arrays": (C.c_int, [PartitionP]),
    "pll_cuda_schedule_levels": (C.c_int, [C.POINTER(Operation), C.c_uint, c_uint_p]),
    "pll_cuda_kernel_launches": (C.c_ulonglong, []),
    "pll_cuda_fastparsimony_edge_scores": (C.c_int, [ParsimonyP, c_uint_p, C.c_uint, c_uint_p]),
    "pll_cuda_schedule_parsimony_levels": (C.c_int, [C.POINTER(ParsBuildOp), C.c_uint, C.c_uint, c_uint_p]),
    "pll_cuda_download_parsimony_vector": (C.c_int, [ParsimonyP, C.c_uint, c_uint_p]),
    "pll_cuda_host_eigen": (
        C.c_int,
        [C.c_uint, C.c_uint, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p],
    ),
}


class PllLibrary:
    """A loaded libpll-compatible shared library with typed entry points."""

    def __init__(self, path: str, cuda: bool | None = None):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: the native library is mandatory (there is no "
                "CPU or Python fallback); build it with __graft_entry__.build()"
            )
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL if False else C.DEFAULT_MODE)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(self.lib, name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)
        for name, (res, args) in _PARS_PROTOS.items():
            if hasattr(self.lib, name):
                fn = getattr(self.lib, name)
                fn.restype, fn.argtypes = res, args
                setattr(self, name, fn)
        self.is_cuda = hasattr(self.lib, "pll_cuda_device_count") if cuda is None else cuda
        if self.is_cuda:
            for name, (res, args) in _CUDA_PROTOS.items():
                if not hasattr(self.lib, name):
                    continue  # an older build of the library ($PLL_B200_LIB)
                fn = getattr(self.lib, name)
                fn.restype, fn.argtypes = res, args
                setattr(self, name, fn)

    # -- thread-local status ------------------------------------------------
    @property
    def errno(self) -> int:
        return C.c_int.in_dll(self.lib, "pll_errno").value

    @property
    def errmsg(self) -> str:
        return (C.c_char * 200).in_dll(self.lib, "pll_errmsg").value.decode(errors="replace")

    def map(self, name: str):
        """One of pll_map_nt / pll_map_aa / pll_map_bin as a ctypes array."""
        return (pll_state_t * 256).in_dll(self.lib, name)


def exported_symbols_declared_in_header(header_path: str) -> list[str]:
    """Function and object names marked PLL_EXPORT in include/pll_b200.h."""
    import re

    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#define[^\n]*", "", text)
    names = []
    for m in re.finditer(r"PLL_EXPORT\s+[^;{]*?\b(pll_[A-Za-z0-9_]+)\s*(\(|\[|;)", text):
        names.append(m.group(1))
    return sorted(set(names))
