"""Synthetic trees, models and alignments for parity tests and bench.py.

Everything here is input generation: a random joining
topology expressed directly as a ``pll_operation_t`` list (the reference's
tests hard-code their operation lists the same way, e.g.
``/root/reference/test/src/derivatives.c:91-98``), branch lengths, a GTR or
amino-acid model, mean-discretised gamma rates and sequences evolved down the
tree with gaps and ambiguity codes sprinkled in so that tip codes beyond the
four/twenty plain states are exercised.
"""
from __future__ import annotations

import dataclasses

import numpy as np

DNA_CODES = b"ACGT"
DNA_AMBIG = b"MRWSYKVHDBN-"  # IUPAC ambiguity codes + gap
AA_CODES = b"ARNDCQEGHILKMFPSTWYV"
AA_AMBIG = b"BZJX-"


@dataclasses.dataclass
class Tree:
    tips: int
    ops: np.ndarray  # (n_ops, 8) int64 rows in pll_operation_t field order
    branch_lengths: np.ndarray  # indexed by matrix index (= child node index)
    root_edge: tuple[int, int, int]  # (parent clv, child clv, matrix index)
    scaler_of: dict[int, int]

    @property
    def inner(self) -> int:
        return self.tips - 2

    @property
    def nodes(self) -> int:
        return 2 * self.tips - 2

    def matrix_indices(self) -> np.ndarray:
        used = sorted({int(r[3]) for r in self.ops} | {int(r[6]) for r in self.ops} | {self.root_edge[2]})
        return np.asarray(used, dtype=np.uint32)


def random_tree(tips: int, rng: np.random.Generator, brlen=(0.02, 0.22), scalers: bool = True) -> Tree:
    """Random joining topology; node i's branch to its parent uses matrix i."""
    assert tips >= 3
    pool = list(range(tips))
    ops = []
    scaler_of: dict[int, int] = {}
    nxt = tips
    while len(pool) > 2:
        i, j = rng.choice(len(pool), size=2, replace=False)
        a, b = pool[i], pool[j]
        for k in sorted((int(i), int(j)), reverse=True):
            pool.pop(k)
        parent = nxt
        nxt += 1
        ps = parent - tips if scalers else -1
        scaler_of[parent] = ps
        ops.append(
            (parent, ps, a, a, scaler_of.get(a, -1), b, b, scaler_of.get(b, -1))
        )
        pool.append(parent)
    a, b = pool
    if a < tips:  # keep an inner node on the "parent" side of the root edge
        a, b = b, a
    n_nodes = 2 * tips - 2
    bl = rng.uniform(brlen[0], brlen[1], size=n_nodes)
    return Tree(tips, np.asarray(ops, dtype=np.int64), bl, (a, b, b), scaler_of)


def caterpillar_tree(tips: int, rng: np.random.Generator, brlen=(0.02, 0.22)) -> Tree:
    """Maximally deep (ladder) topology: triggers CLV scaling with few taxa."""
    ops = []
    scaler_of: dict[int, int] = {}
    cur = 0
    nxt = tips
    for t in range(1, tips - 1):
        parent = nxt
        nxt += 1
        scaler_of[parent] = parent - tips
        ops.append((parent, parent - tips, cur, cur, scaler_of.get(cur, -1), t, t, -1))
        cur = parent
    n_nodes = 2 * tips - 2
    bl = rng.uniform(brlen[0], brlen[1], size=n_nodes)
    return Tree(tips, np.asarray(ops, dtype=np.int64), bl, (cur, tips - 1, tips - 1), scaler_of)


def gamma_rates(alpha: float, cats: int) -> np.ndarray:
    """Mean-discretised gamma rates (Yang 1994), mean rate 1."""
    from scipy.stats import gamma as G

    if cats == 1:
        return np.ones(1)
    cuts = G.ppf(np.arange(1, cats) / cats, a=alpha, scale=1.0 / alpha)
    cdf1 = np.concatenate([[0.0], G.cdf(cuts, a=alpha + 1, scale=1.0 / alpha), [1.0]])
    return np.diff(cdf1) * cats


def gtr_q(rates6: np.ndarray, freqs: np.ndarray) -> np.ndarray:
    n = len(freqs)
    q = np.zeros((n, n))
    k = 0
    for i in range(n):
        for j in range(i + 1, n):
            q[i, j] = rates6[k] * freqs[j]
            q[j, i] = rates6[k] * freqs[i]
            k += 1
    np.fill_diagonal(q, -q.sum(axis=1))
    q /= -(freqs * np.diag(q)).sum()
    return q


def _expm(q: np.ndarray, t: float) -> np.ndarray:
    from scipy.linalg import expm

    p = expm(q * t)
    p = np.clip(p, 0, None)
    return p / p.sum(axis=1, keepdims=True)


def simulate(
    tree: Tree,
    q_per_cat: list[np.ndarray],
    freqs: np.ndarray,
    cat_rates: np.ndarray,
    sites: int,
    rng: np.random.Generator,
    alphabet: bytes,
    ambig: bytes,
    gap_frac: float = 0.01,
    ambig_frac: float = 0.005,
) -> list[bytes]:
    """Evolve ``sites`` columns down ``tree``; returns one byte string per tip."""
    n = len(alphabet)
    cats = len(cat_rates)
    site_cat = rng.integers(0, cats, size=sites)
    state = {}
    a, b, m = tree.root_edge
    root_state = rng.choice(n, size=sites, p=freqs / freqs.sum())
    state[a] = root_state

    def evolve(parent_state, node):
        out = np.empty(sites, dtype=np.int64)
        t = tree.branch_lengths[node]
        for c in range(cats):
            sel = np.nonzero(site_cat == c)[0]
            if sel.size == 0:
                continue
            p = _expm(q_per_cat[c % len(q_per_cat)], t * cat_rates[c])
            cum = np.cumsum(p, axis=1)
            u = rng.random(sel.size)
            out[sel] = (u[:, None] > cum[parent_state[sel]]).sum(axis=1).clip(0, n - 1)
        return out

    state[b] = evolve(root_state, b)
    children = {int(r[0]): (int(r[2]), int(r[5])) for r in tree.ops}
    stack = [a, b]
    while stack:
        node = stack.pop()
        if node in children:
            for ch in children[node]:
                state[ch] = evolve(state[node], ch)
                stack.append(ch)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    amb = np.frombuffer(ambig, dtype=np.uint8)
    seqs = []
    for tip in range(tree.tips):
        s = alpha[state[tip]].copy()
        u = rng.random(sites)
        g = u < gap_frac
        s[g] = ord("-")
        am = (u >= gap_frac) & (u < gap_frac + ambig_frac)
        s[am] = amb[rng.integers(0, len(amb), size=int(am.sum()))]
        seqs.append(s.tobytes())
    return seqs


def mutate_alignment(tips: int, sites: int, rng: np.random.Generator, alphabet: bytes, ambig: bytes,
                     sub_frac: float = 0.15, gap_frac: float = 0.01, ambig_frac: float = 0.005) -> list[bytes]:
    """Cheap alignment for large benchmarks: one base sequence, per-tip random
    substitutions (the survey's timing harness)."""
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    amb = np.frombuffer(ambig, dtype=np.uint8)
    base = rng.integers(0, len(alpha), size=sites, dtype=np.uint8)
    seqs = []
    for _ in range(tips):
        s = base.copy()
        u = rng.random(sites, dtype=np.float32)
        sub = u < sub_frac
        s[sub] = rng.integers(0, len(alpha), size=int(sub.sum()), dtype=np.uint8)
        out = alpha[s]
        g = (u >= sub_frac) & (u < sub_frac + gap_frac)
        out[g] = ord("-")
        am = (u >= sub_frac + gap_frac) & (u < sub_frac + gap_frac + ambig_frac)
        out[am] = amb[rng.integers(0, len(amb), size=int(am.sum()))]
        seqs.append(out.tobytes())
    return seqs


@dataclasses.dataclass
class Dataset:
    """One synthetic problem instance, library-independent."""

    tree: Tree
    states: int
    sites: int
    rate_cats: int
    subst_params: list[np.ndarray]  # one per rate matrix
    freqs: list[np.ndarray]
    cat_rates: np.ndarray
    cat_weights: np.ndarray | None
    params_indices: np.ndarray  # per rate category
    seqs: list[bytes]
    map_name: str
    pattern_weights: np.ndarray | None = None
    prop_invar: float = 0.0


GTR_RATES = np.array([1.2, 3.1, 0.8, 0.9, 3.4, 1.0])
GTR_FREQS = np.array([0.21, 0.29, 0.27, 0.23])


def dna_dataset(tips: int, sites: int, seed: int, alpha: float = 0.7, cats: int = 4,
                brlen=(0.02, 0.22), tree_kind: str = "random", simulate_down_tree: bool = True,
                weights: bool = False, prop_invar: float = 0.0) -> Dataset:
    rng = np.random.default_rng(seed)
    tree = (caterpillar_tree if tree_kind == "caterpillar" else random_tree)(tips, rng, brlen)
    rates = gamma_rates(alpha, cats)
    if simulate_down_tree:
        q = gtr_q(GTR_RATES, GTR_FREQS)
        seqs = simulate(tree, [q], GTR_FREQS, rates, sites, rng, DNA_CODES, DNA_AMBIG)
    else:
        seqs = mutate_alignment(tips, sites, rng, DNA_CODES, DNA_AMBIG)
    pw = rng.integers(1, 5, size=sites).astype(np.uint32) if weights else None
    return Dataset(tree, 4, sites, cats, [GTR_RATES.copy()], [GTR_FREQS.copy()], rates, None,
                   np.zeros(cats, dtype=np.uint32), seqs, "pll_map_nt", pw, prop_invar)


def random_aa_model(rng: np.random.Generator) -> tuple[np.ndarray, np.ndarray]:
    r = rng.gamma(1.0, 1.0, size=190) + 0.01
    r /= r[-1]
    f = rng.dirichlet(np.full(20, 5.0))
    return r, f


def aa_dataset(tips: int, sites: int, seed: int, alpha: float = 0.7, cats: int = 4,
               rate_matrices: int = 4, brlen=(0.02, 0.22), tree_kind: str = "random",
               simulate_down_tree: bool = True) -> Dataset:
    """LG4M-style protein problem: one rate matrix + frequency set per category."""
    rng = np.random.default_rng(seed)
    tree = (caterpillar_tree if tree_kind == "caterpillar" else random_tree)(tips, rng, brlen)
    rates = gamma_rates(alpha, cats)
    models = [random_aa_model(rng) for _ in range(rate_matrices)]
    if simulate_down_tree:
        qs = [gtr_q(m[0], m[1]) for m in models]
        seqs = simulate(tree, qs, models[0][1], rates, sites, rng, AA_CODES, AA_AMBIG)
    else:
        seqs = mutate_alignment(tips, sites, rng, AA_CODES, AA_AMBIG)
    pidx = (np.arange(cats) % rate_matrices).astype(np.uint32)
    return Dataset(tree, 20, sites, cats, [m[0] for m in models], [m[1] for m in models], rates, None,
                   pidx, seqs, "pll_map_aa")


def lg4m_tables(ref_path: str):
    """(rates[4][190], freqs[4][20]) of the LG4M model (Le, Dang & Gascuel 2012) as the reference build
    exports them (``pll_aa_rates_lg4m`` / ``pll_aa_freqs_lg4m``, /root/reference/src/pll.h:596,629; set by
    ``examples/lg4/lg4.c:298-301``), read out of the checker library's data segment: input data only,
    nothing of the reference's code runs.  None when that library did not travel."""
    import ctypes as C
    import os

    if not os.path.exists(ref_path):
        return None
    lib = C.CDLL(ref_path)
    try:
        r = np.array((C.c_double * 190 * 4).in_dll(lib, "pll_aa_rates_lg4m"), dtype=np.float64)
        f = np.array((C.c_double * 20 * 4).in_dll(lib, "pll_aa_freqs_lg4m"), dtype=np.float64)
    except ValueError:
        return None
    return r.copy(), f.copy()


def lg4m_dataset(tips: int, sites: int, seed: int, ref_path: str, alpha: float = 1.0, brlen=(0.02, 0.22),
                 simulate_down_tree: bool = True, tree_kind: str = "random") -> tuple[Dataset, str]:
    """BASELINE config 3: protein, 4 rate categories each with its own LG4M matrix and frequencies
    (params_indices {0,1,2,3}, examples/lg4/lg4.c:286-310).  Falls back to random LG4M-style matrices
    when the tables are not available; the second value says which."""
    rng = np.random.default_rng(seed)
    tree = (caterpillar_tree if tree_kind == "caterpillar" else random_tree)(tips, rng, brlen)
    rates = gamma_rates(alpha, 4)
    tab = lg4m_tables(ref_path)
    if tab is None:
        models = [random_aa_model(rng) for _ in range(4)]
        name = "random LG4M-style matrices (oracle/_ref not present)"
    else:
        models = [(tab[0][i], tab[1][i]) for i in range(4)]
        name = "LG4M (pll_aa_rates_lg4m / pll_aa_freqs_lg4m)"
    if simulate_down_tree:
        qs = [gtr_q(m[0], m[1]) for m in models]
        seqs = simulate(tree, qs, models[0][1], rates, sites, rng, AA_CODES, AA_AMBIG)
    else:
        seqs = mutate_alignment(tips, sites, rng, AA_CODES, AA_AMBIG)
    return Dataset(tree, 20, sites, 4, [m[0] for m in models], [m[1] for m in models], rates, None,
                   np.arange(4, dtype=np.uint32), seqs, "pll_map_aa"), name


def generic_dataset(states: int, tips: int, sites: int, seed: int, cats: int = 4,
                    tree_kind: str = "random", brlen=(0.02, 0.22)) -> Dataset:
    """Odd state counts (5, 7, ...) as in test/src/00012 and derivatives-oddstates."""
    rng = np.random.default_rng(seed)
    tree = (caterpillar_tree if tree_kind == "caterpillar" else random_tree)(tips, rng, brlen)
    rates = gamma_rates(0.8, cats)
    npar = states * (states - 1) // 2
    r = rng.gamma(1.0, 1.0, size=npar) + 0.05
    r /= r[-1]
    f = rng.dirichlet(np.full(states, 5.0))
    alphabet = bytes(range(ord("a"), ord("a") + states))
    q = gtr_q(r, f)
    seqs = simulate(tree, [q], f, rates, sites, rng, alphabet, b"-", gap_frac=0.02, ambig_frac=0.0)
    return Dataset(tree, states, sites, cats, [r], [f], rates, None, np.zeros(cats, dtype=np.uint32),
                   seqs, f"custom{states}")


def custom_map(states: int) -> np.ndarray:
    """State map for generic_dataset: 'a'.. -> one bit each, '-' -> all bits."""
    m = np.zeros(256, dtype=np.uint64)
    for i in range(states):
        m[ord("a") + i] = 1 << i
    m[ord("-")] = (1 << states) - 1
    return m
