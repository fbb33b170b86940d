"""Drive a libpll-compatible library (ours or the reference) over a Dataset.

The call sequence is the reference's own (``examples/unrooted/unrooted.c``,
``examples/newton/newton.c``): create partition, set model, set tips, update
P-matrices, update partials over the operation list, evaluate an edge.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import Operation, PllLibrary
from .synth import Dataset, custom_map


def _dp(a: np.ndarray):
    return a.ctypes.data_as(capi.c_double_p)


def _up(a: np.ndarray):
    return a.ctypes.data_as(capi.c_uint_p)


class Engine:
    def __init__(self, lib: PllLibrary, ds: Dataset, attributes: int, sites_slice: slice | None = None):
        self.lib, self.ds, self.attributes = lib, ds, attributes
        t = ds.tree
        sl = sites_slice or slice(0, ds.sites)
        self.site_lo, self.site_hi = sl.start, sl.stop
        self.sites = sl.stop - sl.start
        self.rate_matrices = len(ds.subst_params)
        self.n_scalers = t.inner
        self.p = lib.pll_partition_create(
            t.tips, t.inner, ds.states, self.sites, self.rate_matrices, t.nodes, ds.rate_cats,
            self.n_scalers, attributes,
        )
        if not self.p:
            raise RuntimeError(f"pll_partition_create failed: errno={lib.errno} {lib.errmsg}")
        self.part = self.p.contents
        self.params_indices = np.ascontiguousarray(ds.params_indices, dtype=np.uint32)
        for i in range(self.rate_matrices):
            f = np.ascontiguousarray(ds.freqs[i], dtype=np.float64)
            s = np.ascontiguousarray(ds.subst_params[i], dtype=np.float64)
            lib.pll_set_frequencies(self.p, i, _dp(f))
            lib.pll_set_subst_params(self.p, i, _dp(s))
        r = np.ascontiguousarray(ds.cat_rates, dtype=np.float64)
        lib.pll_set_category_rates(self.p, _dp(r))
        if ds.cat_weights is not None:
            w = np.ascontiguousarray(ds.cat_weights, dtype=np.float64)
            lib.pll_set_category_weights(self.p, _dp(w))
        if ds.map_name.startswith("custom"):
            self._map_np = custom_map(ds.states)
            self.map = self._map_np.ctypes.data_as(C.POINTER(capi.pll_state_t))
        else:
            self.map = lib.map(ds.map_name)
        self.set_tips()
        if ds.pattern_weights is not None:
            pw = np.ascontiguousarray(ds.pattern_weights[self.site_lo:self.site_hi], dtype=np.uint32)
            lib.pll_set_pattern_weights(self.p, _up(pw))
        if ds.prop_invar > 0:
            for i in range(self.rate_matrices):
                if lib.pll_update_invariant_sites_proportion(self.p, i, ds.prop_invar) != 1:
                    raise RuntimeError(f"invariant sites: {lib.errno} {lib.errmsg}")
        self.ops = (Operation * len(t.ops))()
        for k, row in enumerate(t.ops):
            self.ops[k] = Operation(*[int(x) for x in row])
        self.matrix_indices = t.matrix_indices()
        self.branch_lengths = np.ascontiguousarray(t.branch_lengths[self.matrix_indices], dtype=np.float64)

    def set_tips(self):
        """pll_set_tip_states for every tip: host characters -> state codes (or tip CLVs) in HBM"""
        lo, hi = self.site_lo, self.site_hi
        whole = lo == 0 and hi == self.ds.sites
        for tip, seq in enumerate(self.ds.seqs):
            rc = self.lib.pll_set_tip_states(self.p, tip, self.map, seq if whole else seq[lo:hi])
            if rc != 1:
                raise RuntimeError(f"pll_set_tip_states failed: {self.lib.errno} {self.lib.errmsg}")

    # -- the hot path -------------------------------------------------------
    def update_pmatrices(self, matrix_indices=None, branch_lengths=None):
        mi = self.matrix_indices if matrix_indices is None else np.ascontiguousarray(matrix_indices, dtype=np.uint32)
        bl = self.branch_lengths if branch_lengths is None else np.ascontiguousarray(branch_lengths, dtype=np.float64)
        rc = self.lib.pll_update_prob_matrices(self.p, _up(self.params_indices), _up(mi), _dp(bl), len(mi))
        if rc != 1:
            raise RuntimeError(f"pll_update_prob_matrices failed: {self.lib.errno} {self.lib.errmsg}")

    def update_partials(self, count=None):
        self.lib.pll_update_partials(self.p, self.ops, len(self.ops) if count is None else count)

    def edge_logl(self, edge=None, persite=False):
        a, b, m = edge or self.ds.tree.root_edge
        t = self.ds.tree
        ps = np.empty(self.sites, dtype=np.float64) if persite else None
        v = self.lib.pll_compute_edge_loglikelihood(
            self.p, a, t.scaler_of.get(a, -1), b, t.scaler_of.get(b, -1), m, _up(self.params_indices),
            _dp(ps) if persite else None,
        )
        return (v, ps) if persite else v

    def root_logl(self, node=None, persite=False):
        t = self.ds.tree
        node = t.root_edge[0] if node is None else node
        ps = np.empty(self.sites, dtype=np.float64) if persite else None
        v = self.lib.pll_compute_root_loglikelihood(
            self.p, node, t.scaler_of.get(node, -1), _up(self.params_indices), _dp(ps) if persite else None
        )
        return (v, ps) if persite else v

    def full_traversal(self):
        self.update_pmatrices()
        self.update_partials()
        return self.edge_logl()

    def sumtable_alloc(self):
        # 64-byte aligned: the reference's AVX kernels use aligned stores
        # ascertainment bias: `states` pseudo-sites follow the alignment (src/derivatives.c:131-133)
        asc = self.ds.states if self.part.asc_bias_alloc else 0
        n = (self.sites + asc) * self.ds.rate_cats * self.part.states_padded
        raw = np.zeros(n + 8, dtype=np.float64)
        off = (-raw.ctypes.data % 64) // 8
        return raw[off:off + n]

    def update_sumtable(self, sumtable: np.ndarray, edge=None):
        a, b, _ = edge or self.ds.tree.root_edge
        t = self.ds.tree
        rc = self.lib.pll_update_sumtable(
            self.p, a, b, t.scaler_of.get(a, -1), t.scaler_of.get(b, -1), _up(self.params_indices), _dp(sumtable)
        )
        if rc != 1:
            raise RuntimeError(f"pll_update_sumtable failed: {self.lib.errno} {self.lib.errmsg}")

    def derivatives(self, sumtable: np.ndarray, branch_length: float, edge=None):
        a, b, _ = edge or self.ds.tree.root_edge
        t = self.ds.tree
        d1, d2 = C.c_double(), C.c_double()
        rc = self.lib.pll_compute_likelihood_derivatives(
            self.p, t.scaler_of.get(a, -1), t.scaler_of.get(b, -1), branch_length, _up(self.params_indices),
            _dp(sumtable), C.byref(d1), C.byref(d2),
        )
        if rc != 1:
            raise RuntimeError(f"derivatives failed: {self.lib.errno} {self.lib.errmsg}")
        return d1.value, d2.value

    def newton(self, sumtable: np.ndarray, t0: float, edge=None, tmin=1e-8, tmax=100.0, tol=1e-5, max_iters=32):
        """Newton-Raphson on one branch: (length, d_f, dd_f, evaluations).  One fused launch on the CUDA
        engine (pll_cuda_newton_branch); the same rule driven from the host through
        pll_compute_likelihood_derivatives otherwise (examples/newton/newton.c:67-96 plus clamping)."""
        a, b, _ = edge or self.ds.tree.root_edge
        t = self.ds.tree
        if self.lib.is_cuda:
            ln, d1, d2, it = C.c_double(), C.c_double(), C.c_double(), C.c_uint()
            rc = self.lib.pll_cuda_newton_branch(
                self.p, t.scaler_of.get(a, -1), t.scaler_of.get(b, -1), t0, tmin, tmax, tol, max_iters,
                _up(self.params_indices), _dp(sumtable), C.byref(ln), C.byref(d1), C.byref(d2), C.byref(it))
            if rc != 1:
                raise RuntimeError(f"pll_cuda_newton_branch failed: {self.lib.errno} {self.lib.errmsg}")
            return ln.value, d1.value, d2.value, it.value
        return self.newton_host(sumtable, t0, edge, tmin, tmax, tol, max_iters)

    def newton_host(self, sumtable, t0, edge=None, tmin=1e-8, tmax=100.0, tol=1e-5, max_iters=32):
        ln, d1, d2, it = t0, 0.0, 0.0, 0
        for it in range(1, max_iters + 1):
            d1, d2 = self.derivatives(sumtable, ln, edge)
            if abs(d1) < tol:
                break
            new = min(max(ln - d1 / d2, tmin), tmax)
            if new != new or new == ln:
                break
            ln = new
        return ln, d1, d2, it

    # -- buffer readers (host pointers for the reference, downloads for CUDA) -
    def clv_size(self, idx: int) -> int:
        return int(self.lib.pll_get_clv_size(self.p, idx))

    def clv(self, idx: int) -> np.ndarray:
        n = self.clv_size(idx)
        if self.lib.is_cuda:
            out = np.empty(n, dtype=np.float64)
            if self.lib.pll_cuda_download_clv(self.p, idx, _dp(out)) != 1:
                raise RuntimeError(f"download_clv: {self.lib.errmsg}")
            return out
        return np.ctypeslib.as_array(self.part.clv[idx], shape=(n,)).copy()

    def scaler_size(self, idx: int) -> int:
        if self.lib.is_cuda:
            return int(self.lib.pll_cuda_scaler_size(self.p, idx))
        n = self.sites
        if self.attributes & capi.SITE_REPEATS:
            ids = self.part.repeats.contents.perscale_ids[idx]
            n = ids if ids else self.sites
        return n * (self.ds.rate_cats if self.attributes & capi.RATE_SCALERS else 1)

    def scaler(self, idx: int) -> np.ndarray:
        n = self.scaler_size(idx)
        if self.lib.is_cuda:
            out = np.empty(n, dtype=np.uint32)
            if self.lib.pll_cuda_download_scaler(self.p, idx, _up(out)) != 1:
                raise RuntimeError(f"download_scaler: {self.lib.errmsg}")
            return out
        return np.ctypeslib.as_array(self.part.scale_buffer[idx], shape=(n,)).copy()

    def pmatrix(self, idx: int) -> np.ndarray:
        n = self.ds.states * self.part.states_padded * self.ds.rate_cats
        if self.lib.is_cuda:
            out = np.empty(n, dtype=np.float64)
            if self.lib.pll_cuda_download_pmatrix(self.p, idx, _dp(out)) != 1:
                raise RuntimeError(f"download_pmatrix: {self.lib.errmsg}")
            return out
        return np.ctypeslib.as_array(self.part.pmatrix[idx], shape=(n,)).copy()

    def tipchars(self, tip: int) -> np.ndarray:
        """Pattern-tip codes of one tip (host copy; the CUDA engine forms them on the device and brings the
        host array up to date on request)."""
        if self.lib.is_cuda:
            ptr = self.lib.pll_cuda_host_tipchars(self.p, tip)
            if not ptr:
                raise RuntimeError(f"host_tipchars: {self.lib.errmsg}")
        else:
            ptr = self.part.tipchars[tip]
        return np.ctypeslib.as_array(ptr, shape=(self.sites,)).copy()

    def host_array(self, field: str, idx: int, n: int) -> np.ndarray:
        """Host-canonical per-matrix arrays: eigenvecs, inv_eigenvecs, eigenvals, frequencies."""
        return np.ctypeslib.as_array(getattr(self.part, field)[idx], shape=(n,)).copy()

    def repeat_ids(self, node: int):
        """(class count, site->class, class->first site) through the public
        accessors; count 0 means the node is not compressed."""
        rep = self.part.repeats.contents
        ids = int(rep.pernode_ids[node])
        if not ids:
            return 0, None, None
        sp = self.lib.pll_get_site_id(self.p, node)
        ip = self.lib.pll_get_id_site(self.p, node)
        site_id = np.ctypeslib.as_array(sp, shape=(self.sites,)).copy()
        id_site = np.ctypeslib.as_array(ip, shape=(ids,)).copy()
        return ids, site_id, id_site

    def close(self):
        if self.p:
            self.lib.pll_partition_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
