"""The reference's own example programs, UNMODIFIED, as clients of this library (the drop-in check).

tests/examples/Makefile compiles them from the sources where they lie under /root/reference/examples twice:
against include/pll_b200.h + libpll_b200.so (tests/examples/_bin/<name>) and, where no Newick parser is
needed, against the reference build (tests/examples/_bin/ref/<name>).  The programs request
PLL_ATTRIB_ARCH_AVX / SSE; PLL_CUDA_FORCE=1 redirects them to the CUDA engine.

* self-contained examples (unrooted, rooted, rooted-tacg, newton, heterotachy): the two binaries must
  print the same text -- P-matrices, CLVs (pll_show_*), log-likelihoods, Newton iterates;
* file-driven examples (newick-fasta-unrooted/-rooted, partial-traversal, load-utree, newick-export): run on a
  generated tree + alignment; the printed log-likelihood is checked against the reference library driven
  through the ctypes binding with the operation list of this library's tree layer."""
import ctypes as C
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

pkg = importlib.import_module("libpll-2_b200")
capi = pkg.capi
synth = importlib.import_module("libpll-2_b200.synth")

pytestmark = pytest.mark.gpu

BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "examples", "_bin")
NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?")


def run(path, *args, force_cuda=False):
    env = dict(os.environ)
    if force_cuda:
        env["PLL_CUDA_FORCE"] = "1"
    r = subprocess.run([path, *args], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (path, r.returncode, r.stdout[-500:], r.stderr[-500:])
    return r.stdout


def same_text(a, b, rtol):
    la, lb = a.splitlines(), b.splitlines()
    assert len(la) == len(lb)
    for x, y in zip(la, lb):
        if x == y:
            continue
        assert NUM.sub("#", x) == NUM.sub("#", y), (x, y)
        for u, v in zip(NUM.findall(x), NUM.findall(y)):
            assert abs(float(u) - float(v)) <= rtol * max(abs(float(v)), 1e-6) + 1e-12, (x, y)


@pytest.mark.parametrize("name", ["unrooted", "rooted", "rooted-tacg", "newton", "heterotachy"])
def test_self_contained_examples_print_the_same(cudalib, name):
    ours, ref = os.path.join(BIN, name), os.path.join(BIN, "ref", name)
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("example binaries not built (tests/examples/Makefile needs /root/reference)")
    out_ref = run(ref)
    out_gpu = run(ours, force_cuda=True)
    assert "Log-L" in out_gpu
    same_text(out_gpu, out_ref, 1e-9)


def write_inputs(tmp_path, tips, sites, rooted):
    import test_tree_cpu as tt

    rng = np.random.default_rng(5)
    newick = tt.random_newick(rng, tips, rooted=rooted)
    seqs = synth.mutate_alignment(tips, sites, rng, synth.DNA_CODES, synth.DNA_AMBIG)
    tree, fasta = tmp_path / "tree.nwk", tmp_path / "aln.fa"
    tree.write_text(newick + "\n")
    with open(fasta, "w") as f:
        for i in rng.permutation(tips):  # file order differs from tree order: the examples match by label
            s = seqs[i].decode()
            f.write(f">t{i}\n" + "\n".join(s[k:k + 70] for k in range(0, len(s), 70)) + "\n")
    return newick, seqs, str(tree), str(fasta)


def expected_unrooted_logl(reflib, newick, seqs):
    """examples/newick-fasta-unrooted/newick-fasta-unrooted.c:216-375 through the reference library"""
    import test_tree_cpu as tt

    own = tt.bind(C.CDLL(pkg.LIB_PATH), True)
    tree = own.pll_utree_parse_newick_string(newick.encode())
    t = tree.contents
    tips, n_nodes = t.tip_count, t.tip_count + t.inner_count
    root = t.nodes[n_nodes - 1]
    buf = (C.POINTER(tt.UNode) * n_nodes)()
    size = C.c_uint(0)
    assert own.pll_utree_traverse(root, 1, tt.UCB(lambda n: 1), buf, C.byref(size)) == 1
    ops = (capi.Operation * n_nodes)()
    branches, pm = (C.c_double * n_nodes)(), (C.c_uint * n_nodes)()
    n_mat, n_ops = C.c_uint(0), C.c_uint(0)
    own.pll_utree_create_operations(buf, size.value, branches, pm, ops, C.byref(n_mat), C.byref(n_ops))
    rates = (C.c_double * 4)()
    own.pll_compute_gamma_cats.argtypes = [C.c_double, C.c_uint, C.POINTER(C.c_double), C.c_int]
    assert own.pll_compute_gamma_cats(1.0, 4, rates, 0) == 1
    p = reflib.pll_partition_create(tips, tips - 2, 4, len(seqs[0]), 1, 2 * tips - 3, 4, tips - 2, capi.ARCH_AVX)
    freqs = np.array([0.17, 0.19, 0.25, 0.39])
    reflib.pll_set_frequencies(p, 0, freqs.ctypes.data_as(capi.c_double_p))
    reflib.pll_set_subst_params(p, 0, np.ones(6).ctypes.data_as(capi.c_double_p))
    reflib.pll_set_category_rates(p, rates)
    for i in range(tips):
        node = t.nodes[i].contents
        assert reflib.pll_set_tip_states(p, node.clv_index, reflib.map("pll_map_nt"), seqs[int(node.label[1:])]) == 1
    params = np.zeros(4, dtype=np.uint32)
    reflib.pll_update_prob_matrices(p, params.ctypes.data_as(capi.c_uint_p), pm, branches, n_mat.value)
    reflib.pll_update_partials(p, ops, n_ops.value)
    r = root.contents
    logl = reflib.pll_compute_edge_loglikelihood(p, r.clv_index, r.scaler_index, r.back.contents.clv_index,
                                                 r.back.contents.scaler_index, r.pmatrix_index,
                                                 params.ctypes.data_as(capi.c_uint_p), None)
    reflib.pll_partition_destroy(p)
    own.pll_utree_destroy(tree, None)
    return logl


def test_newick_fasta_unrooted_example(cudalib, reflib, tmp_path):
    exe = os.path.join(BIN, "newick-fasta-unrooted")
    if not os.path.exists(exe):
        pytest.skip("example binaries not built")
    newick, seqs, tree, fasta = write_inputs(tmp_path, 23, 400, rooted=False)
    out = run(exe, tree, fasta, force_cuda=True)
    got = float(re.search(r"Log-L: (-?[\d.]+)", out).group(1))
    want = expected_unrooted_logl(reflib, newick, seqs)
    assert abs(got - want) <= 5e-7 * abs(want) + 1e-6, (got, want)  # printed with 6 decimals
    assert f"Number of tip/leaf nodes in tree: 23" in out


def test_newick_phylip_unrooted_matches_the_reference_likelihood(cudalib, reflib, tmp_path):
    """examples/newick-phylip-unrooted: the same pipeline fed from an interleaved PHYLIP file (pll_phylip.c)"""
    exe = os.path.join(BIN, "newick-phylip-unrooted")
    if not os.path.exists(exe):
        pytest.skip("example binaries not built")
    newick, seqs, tree, _ = write_inputs(tmp_path, 23, 400, rooted=False)
    phy = tmp_path / "aln.phy"
    order = np.random.default_rng(9).permutation(23)
    with open(phy, "w") as f:
        f.write(f"23 {len(seqs[0])}\n")
        for k in range(0, len(seqs[0]), 60):
            for i in order:
                f.write((f"t{i}".ljust(10) if k == 0 else "") + seqs[i][k:k + 60].decode() + "\n")
            f.write("\n")
    out = run(exe, tree, str(phy), force_cuda=True)
    got = float(re.search(r"Log-L: (-?[\d.]+)", out).group(1))
    want = expected_unrooted_logl(reflib, newick, seqs)
    assert abs(got - want) <= 5e-7 * abs(want) + 1e-6, (got, want)


@pytest.mark.parametrize("name,rooted,nargs", [("partial-traversal", False, 2), ("newick-fasta-rooted", True, 2),
                                               ("load-utree", False, 1), ("newick-export", False, 1)])
def test_file_driven_examples_run(cudalib, tmp_path, name, rooted, nargs):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("example binaries not built")
    _, _, tree, fasta = write_inputs(tmp_path, 17, 300, rooted=rooted)
    out = run(exe, *([tree, fasta][:nargs]), force_cuda=True)
    assert out.strip()
    for m in re.findall(r"Log-L[^:]*: (-?[\d.]+(?:[eE][-+]?\d+)?|-?inf|nan)", out):
        assert np.isfinite(float(m)) and float(m) < 0, out[-400:]


def test_weighted_parsimony_example(cudalib, reflib, tmp_path):
    """examples/parsimony/npr-pars.c (rooted Newick + interleaved PHYLIP -> Sankoff parsimony with unit costs ->
    score buffers and reconstructed ancestral sequences read on the host through the struct), unmodified, on
    the GPU.  The score and the root's reconstructed sequence are checked against the reference library driven
    through the same operation lists."""
    import test_tree_cpu as tt

    exe = os.path.join(BIN, "parsimony")
    if not os.path.exists(exe):
        pytest.skip("example binaries not built")
    rng = np.random.default_rng(17)
    tips, sites = 14, 90
    newick = tt.random_newick(rng, tips, rooted=True)
    seqs = synth.mutate_alignment(tips, sites, rng, synth.AA_CODES, synth.AA_AMBIG)
    tree, phy = tmp_path / "tree.nwk", tmp_path / "aln.phy"
    tree.write_text(newick + "\n")
    with open(phy, "w") as f:
        f.write(f"{tips} {sites}\n")
        for i in rng.permutation(tips):
            f.write(f"t{i}".ljust(10) + seqs[i].decode() + "\n")
    out = run(exe, str(tree), str(phy), force_cuda=True)
    got = float(re.search(r"Minimum parsimony score: ([\d.]+)", out).group(1))

    own = tt.bind(C.CDLL(pkg.LIB_PATH), True)
    rt = own.pll_rtree_parse_newick_string(newick.encode())
    t = rt.contents
    n_nodes = t.tip_count + t.inner_count
    buf = (C.POINTER(tt.RNode) * n_nodes)()
    size = C.c_uint(0)
    assert own.pll_rtree_traverse(t.root, 1, tt.RCB(lambda n: 1), buf, C.byref(size)) == 1
    ops = (capi.ParsBuildOp * n_nodes)()
    n_ops = C.c_uint(0)
    f = own.pll_rtree_create_pars_buildops
    f.restype, f.argtypes = None, [C.POINTER(C.POINTER(tt.RNode)), C.c_uint, C.POINTER(capi.ParsBuildOp), capi.c_uint_p]
    f(buf, size.value, ops, C.byref(n_ops))
    matrix = np.ones((20, 20)) - np.eye(20)
    p = reflib.pll_parsimony_create(tips, 20, sites, np.ascontiguousarray(matrix).ctypes.data_as(capi.c_double_p),
                                    tips - 1, tips - 1)
    for i in range(size.value):
        node = buf[i].contents
        if not node.left:
            assert reflib.pll_set_parsimony_sequence(p, node.clv_index, reflib.map("pll_map_aa"),
                                                     seqs[int(node.label[1:])]) == 1
    want = reflib.pll_parsimony_build(p, ops, n_ops.value)
    assert got == want
    # post-order: the root's score buffer is the last "label : ..." line of the first block
    root_clv = ops[n_ops.value - 1].parent_score_index
    sb = np.ctypeslib.as_array(p.contents.sbuffer[root_clv], shape=(sites * 20,))
    block = out.split("Reconstruction:")[0].strip().splitlines()
    printed = [float(x) for x in block[-1].split(":", 1)[1].replace("+", " ").split()]
    assert printed == [float(f"{v:.0f}") for v in sb]
    reflib.pll_parsimony_destroy(p)
    own.pll_rtree_destroy(rt, None)
    assert len(out.split("Reconstruction:")[1].strip().splitlines()) == tips - 1


def newick_splits(newick):
    """tip-label bipartitions of a Newick string (labels are plain words, lengths ignored)"""
    text = re.sub(r":[-+0-9.eE]+", "", newick.strip().rstrip(";"))
    stack, out, everything = [[]], [], set()
    for tok in re.findall(r"[(),]|[^(),]+", text):
        if tok == "(":
            stack.append([])
        elif tok == ")":
            group = frozenset().union(*stack.pop())
            out.append(group)
            stack[-1].append(group)
        elif tok != ",":
            everything.add(tok)
            stack[-1].append(frozenset([tok]))
    everything = frozenset(everything)
    canon = {min(s, everything - s, key=lambda x: (len(x), sorted(x))) for s in out if 1 < len(s) < len(everything) - 1}
    return canon, everything


@pytest.mark.parametrize("attrib,states,seed", [("tpcpu", 4, 1), ("cpu", 4, 7), ("tpcpu", 20, 3)])
def test_stepwise_example_builds_the_same_tree(cudalib, tmp_path, attrib, states, seed):
    """examples/stepwise/stepwise.c (FASTA -> pattern compression -> partition -> pll_fastparsimony_init ->
    pll_fastparsimony_stepwise -> Newick), unmodified, linked against this library and against the reference
    build: same parsimony score and the same tree (the printed Newick strings are rooted at a random inner node
    of the library's node array, so they are compared as sets of bipartitions)."""
    ours, ref = os.path.join(BIN, "stepwise"), os.path.join(BIN, "ref", "stepwise")
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("example binaries not built (tests/examples/Makefile needs /root/reference)")
    rng = np.random.default_rng(states + seed)
    tips, sites = 24, 700
    if states == 4:
        ds = synth.dna_dataset(tips, sites, seed=seed, brlen=(0.05, 0.3))
    else:
        ds = synth.aa_dataset(tips, sites, seed=seed, brlen=(0.05, 0.3))
    fasta = tmp_path / "aln.fa"
    with open(fasta, "w") as f:
        for i in rng.permutation(tips):
            f.write(f">taxon{i}\n{ds.seqs[i].decode()}\n")
    out_ref = run(ref, str(fasta), str(seed), attrib, str(states))
    out_gpu = run(ours, str(fasta), str(seed), attrib, str(states), force_cuda=True)
    score = lambda o: [x for x in o.splitlines() if x.startswith("Score:")]
    assert score(out_gpu) == score(out_ref) and score(out_gpu)
    tree = lambda o: [x for x in o.splitlines() if x.startswith("(")][0]
    assert newick_splits(tree(out_gpu)) == newick_splits(tree(out_ref))
    assert [x for x in out_gpu.splitlines() if x.startswith("Number of")] == \
           [x for x in out_ref.splitlines() if x.startswith("Number of")]


# ---- the reference's own test programs against their golden outputs ------------------------------

REF_TESTS = ["pmatrix", "alpha-cats", "hky", "derivatives", "derivatives-oddstates", "00010_NMDU_lkcalc",
             "00012_NMOU_lkcalc", "00020_NMDR_lkcalc", "00022_NMOR_lkcalc", "00030_NMDU_gamma", "00032_NMOU_gamma",
             "compress-patterns"]


def same_as_golden(out, golden, name):
    """the reference's runner does an exact diff; here numbers may differ in their last printed digit"""
    la, lb = out.splitlines(), golden.splitlines()
    assert len(la) == len(lb), (name, len(la), len(lb))
    bad = 0
    for x, y in zip(la, lb):
        if x == y:
            continue
        assert NUM.sub("#", x) == NUM.sub("#", y), (name, x, y)
        for u, v in zip(NUM.findall(x), NUM.findall(y)):
            digits = len(v.split(".")[1].split("e")[0].split("E")[0]) if "." in v else 0
            ulp = 10.0 ** (-digits) * (10.0 ** int(re.split("[eE]", v)[1]) if re.search("[eE]", v) else 1.0)
            assert abs(float(u) - float(v)) <= 1.5 * ulp + 1e-9 * abs(float(v)), (name, x, y)
        bad += 1
    return bad


@pytest.mark.parametrize("name", REF_TESTS)
@pytest.mark.parametrize("attrs", ["", "tv", "sr"])
def test_reference_test_programs_match_their_golden_outputs(cudalib, name, attrs):
    """test/src/<name>.c (+ common.c) of the reference, compiled unmodified against this library and run with
    the runner's attribute words ("tv" = pattern tips, "sr" = site repeats; the arch words are moot under
    PLL_CUDA_FORCE=1), compared with test/out/<name>.out (tests/golden/)."""
    exe = os.path.join(BIN, "reftests", name)
    golden_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if not os.path.exists(exe):
        pytest.skip("reference test programs not built (tests/examples/Makefile needs /root/reference)")
    env = dict(os.environ, PLL_CUDA_FORCE="1")
    if name == "pmatrix":
        # test/src/pmatrix.c:47-56 reads partition->pmatrix[i] on the host: device-resident buffers become
        # host-dereferenceable with managed allocations (include/pll_b200.h, "PLL_CUDA_MANAGED")
        env["PLL_CUDA_MANAGED"] = "1"
    r = subprocess.run([exe] + attrs.split(), capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, (r.returncode, r.stdout[-400:], r.stderr[-400:])
    if r.stdout == open(os.path.join(golden_dir, "skip.out")).read():
        pytest.skip("the program skips this attribute set")
    same_as_golden(r.stdout, open(os.path.join(golden_dir, name + ".out")).read(), name)
