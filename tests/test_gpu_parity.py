"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libtst.so), against the CPU
oracle on the same seeded inputs.  Bar: bit-exact (bins, folds, every node of every tree, integer
votes, s(x) as float32 bits, IBP masks, work counters)."""
import numpy as np
import pytest

from oracle import hist_rf as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import tst

    e = tst.Engine(0)
    yield e
    e.close()


def _torch():
    import torch

    return torch


def oracle_table(X, n_bins=256):
    edges, ne = H.find_edges(X, n_bins)
    return edges, ne, H.bin_table(X, edges, ne)


def compare_tree(gn, gc, on, oc, tag):
    assert len(gn) == len(on), f"{tag}: node count {len(gn)} != oracle {len(on)}"
    for name in on.dtype.names:
        bad = np.flatnonzero(gn[name] != on[name])
        assert bad.size == 0, f"{tag}: field {name} differs at node {bad[0]}: gpu {gn[bad[0]]} oracle {on[bad[0]]}"
    bad = np.argwhere(gc != oc)
    assert bad.size == 0, f"{tag}: counts differ at {bad[0]}: gpu {gc[bad[0][0]]} oracle {oc[bad[0][0]]}"


def run_both(eng, bins, y, fold, k, T, C, seed=0, keep=False, group=None, n_groups=1, **pk):
    """Fit on GPU and oracle with identical inputs; returns (gpu votes, oracle votes, gpu stats, oracle stats, forest)."""
    import tst

    torch = _torch()
    d = bins.shape[0]
    eng.set_bins(bins)
    y_dev = eng.to_dev(y, torch.uint8)
    fold_dev = eng.to_dev(fold, torch.uint8)
    g_dev = None if group is None else eng.to_dev(group, torch.int32)
    prm = tst.make_params(d, keep_forest=keep, **pk)
    votes = eng.new_votes(C)
    eng.stats(reset=True)
    eng.fit_oof(y_dev, fold_dev, k, T, C, prm, seed=seed, group_dev=g_dev, n_groups=n_groups, votes=votes)
    gst = eng.stats()
    oprm = H.make_params(d, **pk)
    oimp = np.zeros(d + 1, np.uint64)
    ov, ost, forest = H.fit_oof(bins, y, fold, k, T, C, oprm, seed=seed, keep=keep, group=group, n_groups=n_groups, importances=oimp)
    # feature importances accumulated on the device while the trees grew: the integer sums equal the oracle's in every test
    gimp, gcnt = eng.importance_sums()
    assert gcnt == int(oimp[d]), f"trees with a split: gpu {gcnt} oracle {int(oimp[d])}"
    assert np.array_equal(gimp, oimp[:d]), f"importance sums differ: gpu {gimp} oracle {oimp[:d]}"
    return votes.cpu().numpy().view(np.uint32), ov, gst, ost, forest


def check_stats(gst, ost):
    assert gst["row_visits_big"] + gst["row_visits_mid"] + gst["row_visits_small"] + gst["row_visits_sub"] + gst["row_visits_wide"] == ost["row_visits"]
    for key in ("nodes", "leaves", "searched", "fallback", "inbag_rows", "trees", "oof_steps", "max_depth"):
        assert gst[key] == ost[key], f"stat {key}: gpu {gst[key]} oracle {ost[key]}"


# --------------------------------------------------------------------------------------------- binning / folds
def test_bin_map_bit_exact(eng, vsi_small):
    X = vsi_small["X"].copy()
    X[3, 2] = np.nan
    X[11, 0] = np.inf
    X[12, 0] = -np.inf
    edges, ne, ref = oracle_table(X)
    eng.bin(X)
    assert np.array_equal(eng.edges, edges) and np.array_equal(eng.n_edges, ne)
    got = eng.get_bins()
    assert got.dtype == np.uint8 and np.array_equal(got, ref)
    assert got[2, 3] == 255


def test_bin_map_ragged_shapes(eng):
    rng = np.random.default_rng(5)
    for n, d in ((1, 1), (127, 3), (129, 17), (1000, 33)):
        X = rng.standard_normal((n, d)).astype(np.float32)
        X[:, 0] = np.round(X[:, 0])  # low cardinality column
        edges, ne, ref = oracle_table(X)
        eng.bin(X)
        assert np.array_equal(eng.get_bins(), ref), (n, d)


@pytest.mark.parametrize("n_bins", [256, 16, 3])
def test_device_edge_selection_bit_exact(eng, n_bins):
    """tst_bin_auto (radix sort + integer-rank quantiles on the GPU) against the oracle's edge rule."""
    rng = np.random.default_rng(n_bins)
    n = 5003
    cols = [rng.standard_normal(n), np.round(rng.standard_normal(n) * 3), rng.poisson(3.0, n).astype(float),
            np.full(n, 2.5), rng.integers(0, 300, n).astype(float), -np.abs(rng.standard_normal(n)) * 1e-30,
            np.where(rng.random(n) < 0.3, np.nan, rng.standard_normal(n)), np.full(n, np.nan),
            np.where(rng.random(n) < 0.5, -0.0, 0.0), rng.standard_normal(n) * 1e30]
    X = np.stack(cols, axis=1).astype(np.float32)
    X[7, 0] = np.inf
    X[8, 0] = -np.inf
    edges, ne = H.find_edges(X, n_bins)
    eng.bin(X, n_bins=n_bins)
    assert np.array_equal(eng.n_edges, ne)
    assert np.array_equal(eng.edges, edges)
    assert np.array_equal(eng.get_bins(), H.bin_table(X, edges, ne))
    eng.bin(X, n_bins=n_bins, host_edges=True)  # the host rule gives the same table
    assert np.array_equal(eng.edges, edges) and np.array_equal(eng.get_bins(), H.bin_table(X, edges, ne))
    for nn in (1, 2, 31, 257):
        Xs = X[:nn].copy()
        e2, n2 = H.find_edges(Xs, n_bins)
        eng.bin(Xs, n_bins=n_bins)
        assert np.array_equal(eng.n_edges, n2) and np.array_equal(eng.edges, e2), nn


def test_set_bins_roundtrip(eng, vsi_small):
    eng.set_bins(vsi_small["bins"])
    assert np.array_equal(eng.get_bins(), vsi_small["bins"])


@pytest.mark.parametrize("k,n_groups", [(5, 1), (3, 4), (10, 1)])
def test_fold_assignment(eng, vsi_small, k, n_groups):
    torch = _torch()
    y = vsi_small["y"]
    n = len(y)
    rng = np.random.default_rng(k)
    active = (rng.random(n) < 0.85).astype(np.uint8)
    group = rng.integers(0, n_groups, n).astype(np.int32) if n_groups > 1 else None
    eng.set_bins(vsi_small["bins"])
    got = eng.assign_folds(eng.to_dev(y, torch.uint8), eng.to_dev(active, torch.uint8), k, 4,
                           None if group is None else eng.to_dev(group, torch.int32), n_groups).cpu().numpy()
    ref = H.assign_folds(y, active, k, 4, group, n_groups)
    assert np.array_equal(got, ref)


# --------------------------------------------------------------------------------------------- trees
def test_trees_bit_exact_vsi_small(eng, vsi_small):
    v = vsi_small
    T = 8
    gv, ov, gst, ost, forest = run_both(eng, v["bins"], v["y"], v["fold"], v["k"], T, v["C"], seed=3, keep=True)
    for part in range(v["k"]):
        for t in range(T):
            gn, gc = eng.export_tree(part, t, v["C"])
            ot = forest.tree(part, t)
            compare_tree(gn, gc, ot.nodes, ot.counts, f"tree({part},{t})")
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)


def test_votes_and_s_bit_exact_vsi_config(eng):
    """BASELINE.json configs[1]: 5k x 32, 5-fold x 100 trees."""
    from tst import synth

    torch = _torch()
    X, y = synth.vsi_table()
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4)
    gv, ov, gst, ost, _ = run_both(eng, bins, y, fold, 5, 100, 4, seed=0)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    s, cor, tot = eng.finalize(eng.to_dev(gv.view(np.int32), torch.int32), eng.to_dev(y, torch.uint8),
                               eng.to_dev(fold, torch.uint8), 5, 100, 4)
    os_, ocor, otot = H.finalize(ov, y, fold, 5, 100)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    assert np.array_equal(cor.cpu().numpy(), ocor.astype(np.int32)) and np.array_equal(tot.cpu().numpy(), otot.astype(np.int32))


def _edge_case_tables():
    rng = np.random.default_rng(11)
    out = {}
    # low-cardinality only: many nodes where all m sampled features are constant -> fallback draws
    n = 600
    out["lowcard"] = (rng.integers(0, 3, (6, n)).astype(np.uint8), rng.integers(0, 3, n).astype(np.uint8), 3)
    # duplicate rows with conflicting labels -> mixed leaves
    b = rng.integers(0, 4, (4, 40)).astype(np.uint8)
    out["dups"] = (np.tile(b, (1, 10)), rng.integers(0, 2, 400).astype(np.uint8), 2)
    # single class
    out["oneclass"] = (rng.integers(0, 255, (5, 300)).astype(np.uint8), np.zeros(300, np.uint8), 1)
    # constant columns + one informative column + NaN bin
    b = np.zeros((8, 500), np.uint8)
    b[3] = rng.integers(0, 256, 500)
    out["const"] = (b, (b[3] > 100).astype(np.uint8) + (b[3] > 200).astype(np.uint8), 3)
    # many classes, wide table
    out["manyclass"] = (rng.integers(0, 64, (40, 900)).astype(np.uint8), rng.integers(0, 13, 900).astype(np.uint8), 13)
    # the class limit (32): the warp-per-node tier runs with 2 warps per CTA (its unpacked histograms are 33 KB each)
    out["maxclass"] = (rng.integers(0, 200, (12, 2500)).astype(np.uint8), rng.integers(0, 32, 2500).astype(np.uint8), 32)
    # tiny
    out["tiny"] = (rng.integers(0, 5, (2, 7)).astype(np.uint8), rng.integers(0, 2, 7).astype(np.uint8), 2)
    return out


@pytest.fixture(params=["auto", "deep", "plain"])
def paths(request, eng):
    """The throughput variants that are chosen automatically only for deep batches (>= 2^24 (tree, row) pairs) -- size-class queues
    of the warp-per-node tier, the light queue of the CTA tier, several draw positions per pass in the warp-per-subtree kernel --
    forced on for the small tables of these tests as well ("deep"), and all forced off ("plain": the one-position warp-per-subtree
    kernel for criterion 2, which "auto" replaces by the multi-position one)."""
    knobs = {"deep": dict(SMALL_MULTI=1, BIG_LITE=1, MID_CLASSES=3),
             "plain": dict(SMALL_MULTI=0, BIG_LITE=0, MID_CLASSES=0)}.get(request.param, {})  # plain: also what criterion 2 would not pick by itself
    for kk, v in knobs.items():
        eng.set_option(kk, v)
    yield request.param
    for kk in knobs:
        eng.set_option(kk, -1)


@pytest.mark.parametrize("name", ["lowcard", "dups", "oneclass", "const", "manyclass", "maxclass", "tiny"])
@pytest.mark.parametrize("pk", [dict(), dict(max_depth=3), dict(min_samples_leaf=3, min_samples_split=8), dict(max_features=None)])
def test_edge_cases_bit_exact(eng, paths, name, pk):
    bins, y, C = _edge_case_tables()[name]
    n = bins.shape[1]
    k = 3
    fold = H.assign_folds(y, np.ones(n, np.uint8), k, C)
    T = 5
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, k, T, C, seed=9, keep=True, **pk)
    for part in range(k):
        for t in range(T):
            gn, gc = eng.export_tree(part, t, C)
            ot = forest.tree(part, t)
            compare_tree(gn, gc, ot.nodes, ot.counts, f"{name} tree({part},{t})")
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)


@pytest.mark.parametrize("seed", range(24))
def test_random_configurations_bit_exact(eng, paths, seed):
    """Randomised shapes / cardinalities / parameters: every tier (CTA, warp-per-node, warp-per-subtree, packed and
    generic) is hit with node sizes straddling the 32- and 256-row tier boundaries."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([37, 64, 65, 300, 513, 1400, 3000]))
    d = int(rng.integers(1, 40))
    C = int(rng.choice([1, 2, 3, 4, 5, 9, 17]))
    card = int(rng.choice([2, 3, 7, 40, 256]))
    k = int(rng.integers(1, 5))
    bins = rng.integers(0, card, (d, n)).astype(np.uint8)
    if rng.random() < 0.3:
        bins[rng.integers(0, d)] = 255 * (rng.random(n) < 0.2)  # a NaN-bin column
    y = rng.integers(0, C, n).astype(np.uint8)
    if rng.random() < 0.5 and d > 1:  # make the labels learnable so trees are not all noise
        y = ((bins[0].astype(int) * C) // card).clip(0, C - 1).astype(np.uint8)
        flip = rng.random(n) < 0.15
        y[flip] = rng.integers(0, C, int(flip.sum()))
    active = (rng.random(n) < 0.9).astype(np.uint8)
    fold = H.assign_folds(y, active, k, C)
    pk = {}
    if rng.random() < 0.3:
        pk["max_depth"] = int(rng.integers(1, 12))
    if rng.random() < 0.3:
        pk["min_samples_leaf"] = int(rng.integers(1, 6)); pk["min_samples_split"] = int(rng.integers(2, 12))
    if rng.random() < 0.3:
        pk["max_features"] = int(rng.integers(1, d + 1))
    T = 3
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, k, T, C, seed=seed, keep=True, **pk)
    assert np.array_equal(gv, ov), (n, d, C, card, k, pk)
    check_stats(gst, ost)
    part, t = int(rng.integers(0, k)), int(rng.integers(0, T))
    gn, gc = eng.export_tree(part, t, C)
    ot = forest.tree(part, t)
    compare_tree(gn, gc, ot.nodes, ot.counts, f"seed {seed} tree({part},{t})")


@pytest.mark.parametrize("n,d,C,k", [(3000, 16, 4, 3), (5001, 9, 3, 2), (2100, 40, 7, 4)])
def test_dense_root_kernels_bit_exact(eng, request, n, d, C, k):
    """k_root_hist / k_root_split (TMA-streamed column-major histogram of the root) forced on small tables."""
    eng.set_option("ROOT_DENSE", 2)
    eng.set_option("WIDE_MAX", 0)  # roots of <= 1024 rows would go to the four-warps-per-node tier instead of dense root + k_grow_big
    request.addfinalizer(lambda: (eng.set_option("ROOT_DENSE", -1), eng.set_option("WIDE_MAX", -1)))
    rng = np.random.default_rng(n)
    bins = rng.integers(0, 256, (d, n)).astype(np.uint8)
    y = ((bins[1].astype(int) * C) // 256).astype(np.uint8)
    flip = rng.random(n) < 0.2
    y[flip] = rng.integers(0, C, int(flip.sum()))
    fold = H.assign_folds(y, np.ones(n, np.uint8), k, C)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, k, 4, C, seed=5, keep=True)
    assert gst["launches_root"] > 0
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    gn, gc = eng.export_tree(k - 1, 2, C)
    ot = forest.tree(k - 1, 2)
    compare_tree(gn, gc, ot.nodes, ot.counts, "dense-root tree")
    # a pure table: the dense root path must produce the one-leaf tree
    y0 = np.zeros(n, np.uint8)
    gv, ov, gst, ost, _ = run_both(eng, bins, y0, fold, k, 2, 1)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)


def test_partition_stash_option_bit_exact(eng, request):
    """TST_STASH=1: k_grow_big keeps the sampled bins of every row for its partition pass (off by default)."""
    from tst import synth

    eng.set_option("STASH", 1)
    request.addfinalizer(lambda: eng.set_option("STASH", 0))
    X, y = synth.vsi_table(30000, 24, 4, seed=8)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 2, 4)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, 2, 3, 4, seed=4, keep=True)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    gn, gc = eng.export_tree(1, 2, 4)
    ot = forest.tree(1, 2)
    compare_tree(gn, gc, ot.nodes, ot.counts, "stash tree")


@pytest.mark.parametrize("d,C,mf", [(100, 3, "sqrt"), (200, 4, "sqrt"), (130, 6, 40), (70, 2, None)])
def test_wide_tables_bit_exact(eng, d, C, mf):
    """d > 64: rows wider than the shared row tile (byte-gather path), more than 8 sampled features per node
    (several histogram batches), non-power-of-two row strides."""
    rng = np.random.default_rng(d)
    n = 1500
    bins = rng.integers(0, 256, (d, n)).astype(np.uint8)
    y = ((bins[3].astype(int) + bins[d - 1]) % C).astype(np.uint8)
    fold = H.assign_folds(y, np.ones(n, np.uint8), 3, C)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, 3, 3, C, seed=2, keep=True, max_features=mf)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    gn, gc = eng.export_tree(1, 1, C)
    ot = forest.tree(1, 1)
    compare_tree(gn, gc, ot.nodes, ot.counts, f"wide d={d}")


@pytest.mark.parametrize("n,d,mf", [(6000, 20, "sqrt"), (20000, 64, "sqrt"), (4000, 9, None)])
def test_two_class_packed_cta_tier_bit_exact(eng, request, n, d, mf):
    """Binary labels, 256-valued bins, non-root nodes of > 256 rows that weigh < 2^16: k_grow_big's packed 16-bit
    histogram (640 words per feature) is larger than the unpacked 2-class layout (544) -- the region must hold it."""
    eng.set_option("WIDE_MAX", 0)  # keep the 257 .. 1024-row nodes on k_grow_big
    request.addfinalizer(lambda: eng.set_option("WIDE_MAX", -1))
    rng = np.random.default_rng(n + d)
    bins = rng.integers(0, 256, (d, n)).astype(np.uint8)
    y = ((bins[0].astype(int) + bins[d - 1].astype(int)) > 255).astype(np.uint8)
    flip = rng.random(n) < 0.25
    y[flip] ^= 1
    fold = H.assign_folds(y, np.ones(n, np.uint8), 3, 2)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, 3, 6, 2, seed=9, keep=True, max_features=mf)
    assert gst["row_visits_big"] > 0
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    for part, t in ((0, 0), (2, 5)):
        gn, gc = eng.export_tree(part, t, 2)
        ot = forest.tree(part, t)
        compare_tree(gn, gc, ot.nodes, ot.counts, f"C=2 tree({part},{t})")


@pytest.mark.parametrize("n,d,C,mf,pk", [(9000, 12, 4, 3, {}), (7000, 40, 3, None, {}), (12000, 9, 2, 1, {}), (5000, 20, 4, 12, dict(min_samples_leaf=7, min_samples_split=20)),
                                         (8000, 16, 4, "sqrt", dict(max_depth=4)), (2300, 6, 4, 2, {})])
def test_wide_tier_bit_exact(eng, n, d, C, mf, pk):
    """k_grow_wide (four warps per node, 257..1024 rows; rows staged once through a tile when a row is <= 64 bytes): constant, duplicated and two-valued columns force its fallback draw
    (warp 0 continues the draw alone), m = 1 / 3 / 12 / all features exercise the warp-to-position mapping and its tie-breaks,
    leaf / split minima and a depth limit its stop rules; tree roots of <= 1024 rows enter the tier unchecked."""
    rng = np.random.default_rng(n + d)
    bins = rng.integers(0, 256, (d, n)).astype(np.uint8)
    bins[1] = 9                        # constant
    bins[2] = bins[3]                  # duplicated: exact gain ties between two features
    bins[4] = (bins[4] > 128) * 7      # two-valued
    if d > 8:
        bins[6:9] = 3                  # more constants: the first m positions often hold no valid split
    y = ((bins[0].astype(int) * C) // 256).astype(np.uint8)
    flip = rng.random(n) < 0.35
    y[flip] = rng.integers(0, C, int(flip.sum()))
    fold = H.assign_folds(y, np.ones(n, np.uint8), 3, C)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, 3, 5, C, seed=11, keep=True, max_features=mf, **pk)
    assert gst["row_visits_wide"] > 0 and gst["launches_grow_wide"] > 0
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    if isinstance(mf, int) and mf <= 3:
        assert ost["fallback"] > 0  # the draw had to go beyond the first m positions somewhere
    for part, t in ((0, 1), (2, 4)):
        gn, gc = eng.export_tree(part, t, C)
        ot = forest.tree(part, t)
        compare_tree(gn, gc, ot.nodes, ot.counts, f"wide tier tree({part},{t})")


@pytest.mark.parametrize("sub_rows", [2048, 1024, 0])
@pytest.mark.parametrize("shape", [(20000, 24, 4, 3, "sqrt"), (5000, 32, 4, 5, "sqrt"), (3100, 9, 3, 2, None), (2500, 64, 2, 2, 20), (900, 5, 4, 3, 1)])
def test_subtree_tier_bit_exact(eng, request, sub_rows, shape):
    """k_grow_sub (one CTA finishes a subtree of <= 2048 / 1024 rows from shared memory) against the oracle, and the
    level-synchronous warp tiers it replaces (SUB_ROWS=0) on the same tables: votes, counters, whole trees."""
    from tst import synth

    n, d, C, k, mf = shape
    eng.set_option("SUB_ROWS", sub_rows)
    request.addfinalizer(lambda: eng.set_option("SUB_ROWS", -1))
    X, y = synth.vsi_table(n, d, C, seed=n + d)
    X[:, 0] = np.round(X[:, 0] * 2)  # a low-cardinality column: many ties
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(y, np.ones(n, np.uint8), k, C)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, k, 5, C, seed=sub_rows + n, keep=True, max_features=mf)
    if sub_rows and d <= 64:
        assert gst["row_visits_sub"] > 0 and gst["row_visits_mid"] == 0 and gst["launches_grow_sub"] > 0
    else:
        assert gst["row_visits_sub"] == 0
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    for part, t in ((0, 0), (k - 1, 4)):
        gn, gc = eng.export_tree(part, t, C)
        ot = forest.tree(part, t)
        compare_tree(gn, gc, ot.nodes, ot.counts, f"sub_rows={sub_rows} tree({part},{t})")


@pytest.mark.parametrize("pk", [dict(max_depth=6), dict(min_samples_leaf=5, min_samples_split=12), dict(max_features=None)])
def test_subtree_tier_growth_parameters(eng, request, pk):
    """Stop rules and the fallback draw inside the CTA-per-subtree kernel: depth limit, leaf / split minima on weighted counts,
    constant and duplicated columns (features that cannot split)."""
    eng.set_option("SUB_ROWS", 2048)
    request.addfinalizer(lambda: eng.set_option("SUB_ROWS", -1))
    rng = np.random.default_rng(17)
    n, d, C = 6000, 12, 4
    bins = rng.integers(0, 200, (d, n)).astype(np.uint8)
    bins[3] = 7                       # constant column
    bins[4] = bins[5]                 # duplicated column
    bins[6] = (bins[6] > 100) * 3     # two-valued column
    y = ((bins[0].astype(int) // 50) % C).astype(np.uint8)
    flip = rng.random(n) < 0.3
    y[flip] = rng.integers(0, C, int(flip.sum()))
    fold = H.assign_folds(y, np.ones(n, np.uint8), 3, C)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, 3, 4, C, seed=2, keep=True, **pk)
    assert gst["row_visits_sub"] > 0
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    gn, gc = eng.export_tree(1, 3, C)
    ot = forest.tree(1, 3)
    compare_tree(gn, gc, ot.nodes, ot.counts, f"params {pk}")


@pytest.mark.parametrize("seed", range(16))
def test_random_large_configurations_bit_exact(eng, seed):
    """Randomised tables large enough for every level-synchronous tier at once (CTA per node, four warps per node, warp per
    node, warp per subtree, thread per subtree) with random knobs: blind-level group sizes, tier boundaries, batch sizes --
    the queue / counter plumbing between the tiers is what this test is after."""
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.choice([2600, 5000, 9000, 20000, 45000]))
    d = int(rng.choice([3, 8, 16, 33, 64]))
    C = int(rng.choice([2, 3, 4]))
    card = int(rng.choice([3, 12, 256]))
    k = int(rng.integers(1, 4))
    bins = rng.integers(0, card, (d, n)).astype(np.uint8)
    y = ((bins[0].astype(int) * C) // card).clip(0, C - 1).astype(np.uint8)
    flip = rng.random(n) < 0.3
    y[flip] = rng.integers(0, C, int(flip.sum()))
    fold = H.assign_folds(y, np.ones(n, np.uint8), k, C)
    pk = {}
    if rng.random() < 0.4:
        pk["max_depth"] = int(rng.integers(2, 14))
    if rng.random() < 0.3:
        pk["min_samples_leaf"] = int(rng.integers(1, 9)); pk["min_samples_split"] = int(rng.integers(2, 30))
    if rng.random() < 0.4:
        pk["max_features"] = int(rng.integers(1, d + 1))
    knobs = dict(BLIND_LEVELS=int(rng.choice([1, 2, 4, 8])), THIN_NODES=int(rng.choice([0, 64, 2048, 100000])),
                 WIDE_MAX=int(rng.choice([0, 600, 1024])), MID_MAX=int(rng.choice([64, 256])), TINY_MAX=int(rng.choice([0, 3, 6])), MID_CLASSES=int(rng.integers(0, 4)), BIG_LITE=int(rng.integers(0, 2)), SMALL_MULTI=int(rng.integers(0, 2)))
    for kk, v in knobs.items():
        eng.set_option(kk, v)
    try:
        T = int(rng.integers(2, 6))
        gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, k, T, C, seed=seed, keep=True, **pk)
    finally:
        for kk, v in dict(BLIND_LEVELS=4, THIN_NODES=2048, WIDE_MAX=-1, MID_MAX=-1, TINY_MAX=-1, MID_CLASSES=-1, BIG_LITE=-1, SMALL_MULTI=-1).items():
            eng.set_option(kk, v)
    assert np.array_equal(gv, ov), (n, d, C, card, k, pk, knobs)
    check_stats(gst, ost)
    part, t = int(rng.integers(0, k)), int(rng.integers(0, T))
    gn, gc = eng.export_tree(part, t, C)
    ot = forest.tree(part, t)
    compare_tree(gn, gc, ot.nodes, ot.counts, f"large seed {seed} {knobs} tree({part},{t})")


def test_fewer_rows_than_folds(eng):
    rng = np.random.default_rng(2)
    bins = rng.integers(0, 9, (3, 4)).astype(np.uint8)
    y = np.array([0, 1, 0, 1], np.uint8)
    fold = H.assign_folds(y, np.ones(4, np.uint8), 5, 2)  # fold 4 is empty, folds hold one row each
    gv, ov, gst, ost, _ = run_both(eng, bins, y, fold, 5, 3, 2, seed=1)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)


@pytest.mark.parametrize("kind,T", [("mmmu", 20), ("cvbench", 12), ("videomme", 10)])
def test_grouped_forests_one_launch(eng, kind, T):
    """BASELINE.json configs[2]: per-question-type forests, every group's k x T forests in the same launches --
    MMMU-shaped (30 groups x 30 rows x 16), CV-Bench-shaped (4 x 660 x 16), VideoMME-shaped (12 x 225 x 24)."""
    from tst import synth

    X, y, group, n_groups = synth.grouped_table(kind)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4, group, n_groups)
    gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, 5, T, 4, seed=4, group=group, n_groups=n_groups, keep=True)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    assert gst["trees"] == n_groups * 5 * T
    part = (n_groups - 1) * 5 + 2  # (last group, fold 2)
    gn, gc = eng.export_tree(part, T - 1, 4)
    ot = forest.tree(part, T - 1)
    compare_tree(gn, gc, ot.nodes, ot.counts, f"{kind} tree({part},{T - 1})")


def test_run_oof_single_call(eng, vsi_small):
    """tst_run_oof (folds + forests + s(x) in one C call on device buffers) == the oracle, with a partly inactive table."""
    import tst

    torch = _torch()
    v = vsi_small
    rng = np.random.default_rng(3)
    active = (rng.random(len(v["y"])) < 0.8).astype(np.uint8)
    eng.set_bins(v["bins"])
    fold, votes, s = eng.run_oof(eng.to_dev(v["y"], torch.uint8), eng.to_dev(active, torch.uint8), 4, 9, 4, tst.make_params(16), seed=8)
    ofold = H.assign_folds(v["y"], active, 4, 4)
    ov, _, _ = H.fit_oof(v["bins"], v["y"], ofold, 4, 9, 4, H.make_params(16), seed=8)
    os_, _, _ = H.finalize(ov, v["y"], ofold, 4, 9)
    assert np.array_equal(fold.cpu().numpy(), ofold) and np.array_equal(votes.cpu().numpy().view(np.uint32), ov)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    assert (s.cpu().numpy()[active == 0] == -1.0).all()  # pruned rows are never scored


def test_grouped_public_api_matches_per_group_runs(eng):
    """run_tst(group=...) == one run per question type: s(x) of a row only depends on its own group's forests."""
    import tst
    from tst import synth

    X, y, group, n_groups = synth.grouped_table("cvbench")
    res = tst.run_tst(X, y, k=5, n_trees=8, seed=2, group=group, engine=eng)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, 4, group, n_groups)
    ov, _, _ = H.fit_oof(bins, y, fold, 5, 8, 4, H.make_params(16), seed=2, group=group, n_groups=n_groups)
    os_, cor, tot = H.finalize(ov, y, fold, 5, 8, group, n_groups)
    assert np.array_equal(res.fold_id, fold) and np.array_equal(res.oof_votes, ov)
    assert np.array_equal(res.s.view(np.uint32), os_.view(np.uint32)) and res.fold_accuracy.shape == (n_groups * 5,)
    assert np.allclose(res.fold_accuracy, cor / tot)


def test_batching_does_not_change_results(eng, vsi_small):
    import tst

    torch = _torch()
    v = vsi_small
    eng.set_bins(v["bins"])
    y_dev, fold_dev = eng.to_dev(v["y"], torch.uint8), eng.to_dev(v["fold"], torch.uint8)
    outs = []
    for tif in (0, 7, 1):
        prm = tst.make_params(v["bins"].shape[0], trees_in_flight=tif)
        votes = eng.new_votes(v["C"])
        eng.fit_oof(y_dev, fold_dev, v["k"], 6, v["C"], prm, seed=5, votes=votes)
        outs.append(votes.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_rank_sharding_sums_to_single_rank(vsi_small):
    """world in {1, 2, 4}: ranks emulated sequentially on one GPU; integer votes make the sum exact."""
    import tst

    torch = _torch()
    v = vsi_small
    T = 6
    ref = None
    for world in (1, 2, 4):
        parts = []
        for rank in range(world):
            e = tst.Engine(0, rank, world)
            e.set_bins(v["bins"])
            votes = e.new_votes(v["C"])
            prm = tst.make_params(v["bins"].shape[0])
            e.fit_oof(e.to_dev(v["y"], torch.uint8), e.to_dev(v["fold"], torch.uint8), v["k"], T, v["C"], prm, seed=2, votes=votes)
            parts.append(votes)
            if rank == world - 1:
                gathered = torch.stack(parts).contiguous()
                out = torch.empty_like(votes)
                tst._lib.check(tst._lib.lib().tst_reduce_votes(e._h, e._ptr(gathered), world, votes.numel(), e._ptr(out), e._stream()))
                total = out.cpu().numpy()
            e.close()
        if ref is None:
            ref = total
        assert np.array_equal(total, ref), world
    ov, _, _ = H.fit_oof(v["bins"], v["y"], v["fold"], v["k"], T, v["C"], H.make_params(v["bins"].shape[0]), seed=2)
    assert np.array_equal(ref.view(np.uint32), ov)


def test_mid_size_table_bit_exact(eng):
    """60k x 64 scale-up-shaped table: exercises deep level-wise growth and many warp subtrees."""
    from tst import synth

    X, y = synth.scaleup_table(60_000, 64, 4, seed=1)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 10, 4)
    gv, ov, gst, ost, _ = run_both(eng, bins, y, fold, 10, 2, 4, seed=0)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)


# --------------------------------------------------------------------------------------------- IBP / API
def test_ibp_bit_exact(eng, vsi_small):
    import tst

    torch = _torch()
    v = vsi_small
    n = len(v["y"])
    T, rounds, tau, batch = 10, 4, 0.45, 60
    eng.set_bins(v["bins"])
    active = torch.ones(n, dtype=torch.uint8, device=eng.dev)
    prm = tst.make_params(v["bins"].shape[0])
    removed, s = eng.ibp(eng.to_dev(v["y"], torch.uint8), active, v["k"], T, v["C"], prm, 7, rounds, tau, batch)
    oact, orem, os_ = H.ibp(v["bins"], v["y"], v["k"], T, v["C"], H.make_params(v["bins"].shape[0]), seed=7, rounds=rounds,
                            tau=tau, batch=batch)
    assert removed.tolist() == orem.tolist()
    assert np.array_equal(active.cpu().numpy(), oact)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), os_.view(np.uint32))


def test_ibp_select_ties_and_threshold(eng):
    import tst

    torch = _torch()
    rng = np.random.default_rng(0)
    n, C, k = 3000, 3, 5
    eng.set_bins(rng.integers(0, 4, (2, n)).astype(np.uint8))
    y = rng.integers(0, C, n).astype(np.uint8)
    fold = rng.integers(0, k, n).astype(np.uint8)
    votes = (rng.integers(0, 40, (n, C)) * 65536).astype(np.uint32)  # heavy ties
    for vmin, batch in ((0, 100), (20 * 65536, 57), (39 * 65536, 5000), (41 * 65536, 10), (0, 1)):
        act0 = (rng.random(n) < 0.9).astype(np.uint8)
        a_dev = eng.to_dev(act0, torch.uint8)
        took = eng.ibp_select(eng.to_dev(votes.view(np.int32), torch.int32), eng.to_dev(y, torch.uint8),
                              eng.to_dev(fold, torch.uint8), k, C, vmin, batch, a_dev)
        ref = act0.copy()
        otook = H.ibp_select(votes, y, fold, k, vmin, batch, ref)
        assert took == otook and np.array_equal(a_dev.cpu().numpy(), ref), (vmin, batch)


def test_public_api_run_tst_and_forest(vsi_small):
    import tst

    v = vsi_small
    res = tst.run_tst(v["X"], v["y"], k=5, n_trees=12, seed=1)
    ov, _, _ = H.fit_oof(v["bins"], v["y"], v["fold"], 5, 12, 4, H.make_params(16), seed=1)
    assert np.array_equal(res.fold_id, v["fold"]) and np.array_equal(res.oof_votes, ov)
    os_, cor, tot = H.finalize(ov, v["y"], v["fold"], 5, 12)
    assert np.array_equal(res.s.view(np.uint32), os_.view(np.uint32))
    assert np.allclose(res.fold_accuracy, cor / tot)
    # sklearn-like classifier
    clf = tst.TsTForest(n_estimators=10, random_state=3).fit(v["X"], v["y"])
    proba = clf.predict_proba(v["X"][:200])
    assert proba.shape == (200, 4) and np.allclose(proba.sum(1), 1.0)
    fold = np.full(len(v["y"]), H.FOLD_TRAIN_ONLY, np.uint8)
    _, _, forest = H.fit_oof(v["bins"], v["y"], fold, 1, 10, 4, H.make_params(16), seed=3, keep=True)
    assert np.array_equal(clf.predict_votes(v["X"][:200]), forest.predict_votes(v["bins"][:, :200].copy()))
    gn, gc = clf.export_tree(4)
    ot = forest.tree(0, 4)
    compare_tree(gn, gc, ot.nodes, ot.counts, "forest tree 4")
    assert (clf.predict(v["X"]) == v["y"]).mean() > 0.9  # fully grown trees memorise the training set
    imp = clf.feature_importances_
    assert imp.shape == (16,) and abs(imp.sum() - 1.0) < 1e-9 and (imp >= 0).all()
    assert np.allclose(imp, clf.feature_importances_from_trees(), atol=1e-6)  # device accumulation == host recomputation from the trees
    oimp = np.zeros(17, np.uint64)
    H.fit_oof(v["bins"], v["y"], fold, 1, 10, 4, H.make_params(16), seed=3, importances=oimp)
    assert np.allclose(imp, H.importances_from_sums(oimp), atol=1e-12)
    # the out-of-fold forests of run_tst carry importances too (no kept forest needed)
    oimp = np.zeros(17, np.uint64)
    H.fit_oof(v["bins"], v["y"], v["fold"], 5, 12, 4, H.make_params(16), seed=1, importances=oimp)
    assert np.allclose(res.feature_importances, H.importances_from_sums(oimp), atol=1e-12)
    sk = clf.sklearn_tree(0)
    assert sk["threshold"].shape == sk["feature"].shape and np.isfinite(sk["threshold"]).all()
    # per-feature attribution of the blind prediction: telescopes to the forest's probability of the true class
    prior, contrib = clf.explain(v["X"][:64], v["y"][:64])
    p_true = proba[np.arange(64), np.searchsorted(clf.classes_, v["y"][:64])]
    assert contrib.shape == (64, 16) and np.allclose(prior + contrib.sum(1), p_true, atol=1e-3)


# --------------------------------------------------------------------------------------------- squared-error criterion
@pytest.mark.parametrize("n,d,L,k,mf,pk", [(3000, 10, 16, 3, 1.0, {}), (5000, 20, 32, 2, "sqrt", {}), (700, 6, 8, 4, None, {}),
                                            (20000, 12, 32, 2, 4, {}), (2500, 9, 4, 3, 3, {}), (4000, 8, 2, 2, "sqrt", dict(max_depth=6)),
                                            (1500, 14, 24, 3, 0.5, dict(min_samples_leaf=4, min_samples_split=10))])
def test_squared_error_criterion_bit_exact(eng, n, d, L, k, mf, pk):
    """Numeric answers: labels are ordinal levels, splits maximise S_L^2/N_L + S_R^2/N_R (every tier's generic kernel;
    L <= 4 checks that the packed Gini kernels are NOT used for this criterion)."""
    from tst import quantise_target, synth

    X, y = synth.numeric_table(n, d, seed=n)
    lev, lo, step = quantise_target(y, L)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(lev, np.ones(n, np.uint8), k, L)
    gv, ov, gst, ost, forest = run_both(eng, bins, lev, fold, k, 4, L, seed=3, keep=True, criterion="squared_error_levels", max_features=mf, **pk)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    for part, t in ((0, 0), (k - 1, 3)):
        gn, gc = eng.export_tree(part, t, L)
        ot = forest.tree(part, t)
        compare_tree(gn, gc, ot.nodes, ot.counts, f"regression tree({part},{t})")


def test_squared_error_dense_root_bit_exact(eng, request):
    eng.set_option("ROOT_DENSE", 2)
    request.addfinalizer(lambda: eng.set_option("ROOT_DENSE", -1))
    from tst import quantise_target, synth

    X, y = synth.numeric_table(4000, 16, seed=9)
    lev, lo, step = quantise_target(y, 20)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(lev, np.ones(len(lev), np.uint8), 3, 20)
    gv, ov, gst, ost, forest = run_both(eng, bins, lev, fold, 3, 4, 20, seed=1, keep=True, criterion="squared_error_levels", max_features=6)
    assert gst["launches_root"] > 0
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    gn, gc = eng.export_tree(1, 2, 20)
    ot = forest.tree(1, 2)
    compare_tree(gn, gc, ot.nodes, ot.counts, "dense-root regression tree")


def test_regression_golden_and_public_api(eng):
    import os

    import tst

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spec_regression.npz"))
    gv, ov, gst, ost, forest = run_both(eng, g["bins"], g["levels"], g["fold"], 3, 5, 16, seed=21, keep=True, criterion="squared_error_levels", max_features=1.0)
    assert np.array_equal(gv, g["votes"]) and np.array_equal(ov, g["votes"])
    gn, gc = eng.export_tree(1, 2, 16)
    compare_tree(gn, gc, g["tree_nodes"], g["tree_counts"], "golden regression tree")
    # host-to-host call
    res = tst.run_tst_regression(g["X"], g["y"], k=3, n_trees=5, seed=21, n_levels=16, engine=eng, criterion="squared_error_levels")
    assert np.array_equal(res.fold_id, g["fold"]) and np.array_equal(res.oof_votes, g["votes"]) and np.array_equal(res.levels, g["levels"])
    assert np.allclose(res.y_pred, tst.votes_to_value(g["votes"], float(g["lo"]), float(g["step"])))
    assert ((res.s >= 0) & (res.s <= 1)).all()
    # sklearn-like regressor: fully grown trees reproduce the training targets up to the quantisation step
    reg = tst.TsTForest(n_estimators=20, criterion="squared_error_levels", n_levels=32, max_features=1.0, random_state=2).fit(g["X"], g["y"])
    err = np.abs(reg.predict(g["X"]) - g["y"])
    assert np.median(err) < 2.0 * (g["y"].max() - g["y"].min()) / 32


# --------------------------------------------------------------------------------------------- continuous squared error (criterion 2)
def run_both_reg(eng, bins, yq, lev, fold, k, T, seed=0, keep=False, **pk):
    """criterion "squared_error" on GPU and oracle: votes [n, 2], stats, importances, forest."""
    import tst

    torch = _torch()
    d = bins.shape[0]
    eng.set_bins(bins)
    eng.set_target(yq)
    y_dev, fold_dev = eng.to_dev(lev, torch.uint8), eng.to_dev(fold, torch.uint8)
    prm = tst.make_params(d, keep_forest=keep, criterion="squared_error", **pk)
    votes = eng.new_votes(2)
    eng.stats(reset=True)
    eng.fit_oof(y_dev, fold_dev, k, T, 2, prm, seed=seed, votes=votes)
    gst = eng.stats()
    H.set_target(yq)
    oimp = np.zeros(d + 1, np.uint64)
    ov, ost, forest = H.fit_oof(bins, lev, fold, k, T, 2, H.make_params(d, criterion="squared_error", **pk), seed=seed, keep=keep, importances=oimp)
    gimp, gcnt = eng.importance_sums()
    assert gcnt == int(oimp[d]) and np.array_equal(gimp, oimp[:d]), "importance sums differ"
    return votes.cpu().numpy().view(np.uint32), ov, gst, ost, forest


@pytest.mark.parametrize("n,d,k,mf,ties,pk", [(3000, 10, 3, 1.0, False, {}), (5000, 32, 5, 1.0, False, {}), (700, 6, 4, None, True, {}),
                                              (40000, 12, 2, 4, False, {}), (2500, 9, 3, 3, True, {}), (4000, 8, 2, "sqrt", False, dict(max_depth=6)),
                                              (1500, 14, 3, 0.5, True, dict(min_samples_leaf=4, min_samples_split=10)), (3000, 70, 2, 12, False, {})])
def test_continuous_squared_error_bit_exact(eng, paths, n, d, k, mf, ties, pk):
    """Numeric answers on a continuous target (16-bit fixed point): (weight, sum y) histograms in every tier -- CTA per node,
    warp per node, warp per subtree, thread per subtree -- against the oracle: votes, counters, importances, whole trees.
    `ties`: integer-valued targets (counts), so many multi-row nodes are constant and must become leaves unsearched."""
    import tst
    from tst import synth

    X, y = synth.numeric_table(n, d, seed=n)
    if ties:
        y = np.round(y / y.max() * 9)
    yq, lo, span = tst.quantise_target16(y)
    lev = np.minimum((yq.astype(np.uint32) * 16) >> 16, 15).astype(np.uint8)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(lev, np.ones(n, np.uint8), k, 16)
    gv, ov, gst, ost, forest = run_both_reg(eng, bins, yq, lev, fold, k, 4, seed=3, keep=True, max_features=mf, **pk)
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    for part, t in ((0, 0), (k - 1, 3)):
        gn, gc = eng.export_tree(part, t, 2)
        ot = forest.tree(part, t)
        compare_tree(gn, gc, ot.nodes, ot.counts, f"continuous regression tree({part},{t})")


@pytest.mark.parametrize("n,d,C,mf,pk", [(30000, 16, 4, "sqrt", {}), (9000, 12, 2, 3, {}), (20000, 40, 3, 12, {}), (12000, 9, 4, None, dict(min_samples_leaf=9, min_samples_split=40)),
                                         (15000, 20, 4, 1, dict(max_depth=7)), (8000, 64, 4, "sqrt", {})])
def test_cta_tier_light_queue_bit_exact(eng, request, n, d, C, mf, pk):
    """BIG_LITE: CTA-tier nodes that weigh < 2^16 leave emit_split through their own queue and run the packed-Gini-only instantiation
    of k_grow_big (more than 8 features per node in batches of 8, fallback draws on constant columns, depth / leaf-size limits);
    heavier nodes and pre-split roots keep the generic kernel.  WIDE_MAX = 0 sends every node above 256 rows there."""
    rng = np.random.default_rng(n + d)
    bins = rng.integers(0, 256, (d, n)).astype(np.uint8)
    bins[1] = 7                      # constant column: fallback draws
    bins[2] = bins[2] & 1            # two-valued column
    y = ((bins[0].astype(int) + bins[3]) * C // 511).clip(0, C - 1).astype(np.uint8)
    flip = rng.random(n) < 0.3
    y[flip] = rng.integers(0, C, int(flip.sum()))
    k = 2
    fold = H.assign_folds(y, np.ones(n, np.uint8), k, C)
    for wide in (-1, 0):
        eng.set_option("BIG_LITE", 1)
        eng.set_option("WIDE_MAX", wide)
        request.addfinalizer(lambda: (eng.set_option("BIG_LITE", -1), eng.set_option("WIDE_MAX", -1)))
        gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, k, 3, C, seed=9, keep=True, max_features=mf, **pk)
        assert np.array_equal(gv, ov), (wide,)
        check_stats(gst, ost)
        assert gst["row_visits_big"] > 0
        gn, gc = eng.export_tree(1, 2, C)
        ot = forest.tree(1, 2)
        compare_tree(gn, gc, ot.nodes, ot.counts, f"light queue wide={wide}")


@pytest.mark.parametrize("classes", [0, 1, 2, 3])
@pytest.mark.parametrize("crit", ["gini", "squared_error"])
def test_mid_tier_size_classes_bit_exact(eng, request, classes, crit):
    """The warp-per-node tier routes 33..64-row and 65..128-row nodes to their own queues / kernel instantiations (MID_CLASSES
    bit 0 / bit 1); every combination must give the trees of the single-queue tier (= the oracle's), for both criteria."""
    import tst
    from tst import synth

    eng.set_option("MID_CLASSES", classes)
    request.addfinalizer(lambda: eng.set_option("MID_CLASSES", -1))
    n, d, k = 12000, 16, 3
    if crit == "gini":
        rng = np.random.default_rng(77)
        bins = rng.integers(0, 256, (d, n)).astype(np.uint8)
        y = ((bins[0].astype(int) + bins[1]) * 4 // 511).clip(0, 3).astype(np.uint8)
        flip = rng.random(n) < 0.35
        y[flip] = rng.integers(0, 4, int(flip.sum()))
        fold = H.assign_folds(y, np.ones(n, np.uint8), k, 4)
        gv, ov, gst, ost, forest = run_both(eng, bins, y, fold, k, 4, 4, seed=5, keep=True)
        C = 4
    else:
        X, yv = synth.numeric_table(n, d, seed=11)
        yq, lo, span = tst.quantise_target16(yv)
        lev = np.minimum((yq.astype(np.uint32) * 16) >> 16, 15).astype(np.uint8)
        edges, ne, bins = oracle_table(X)
        fold = H.assign_folds(lev, np.ones(n, np.uint8), k, 16)
        gv, ov, gst, ost, forest = run_both_reg(eng, bins, yq, lev, fold, k, 4, seed=5, keep=True, max_features=0.5)
        C = 2
    assert np.array_equal(gv, ov)
    check_stats(gst, ost)
    assert gst["row_visits_mid"] > 0
    gn, gc = eng.export_tree(1, 2, C)
    ot = forest.tree(1, 2)
    compare_tree(gn, gc, ot.nodes, ot.counts, f"size classes {classes} {crit}")


def test_continuous_squared_error_public_api(eng):
    """run_tst_regression / TsTForest(criterion="squared_error") host to host; out-of-fold R^2 against the oracle's votes."""
    import tst
    from tst import synth

    X, y = synth.numeric_table(3000, 16, seed=4)
    res = tst.run_tst_regression(X, y, k=4, n_trees=30, seed=5, engine=eng)
    yq, lo, span = tst.quantise_target16(y)
    lev = np.minimum((yq.astype(np.uint32) * 32) >> 16, 31).astype(np.uint8)
    edges, ne, bins = oracle_table(X)
    fold = H.assign_folds(lev, np.ones(len(y), np.uint8), 4, 32)
    H.set_target(yq)
    ov, _, _ = H.fit_oof(bins, lev, fold, 4, 30, 2, H.make_params(16, max_features=1.0, criterion="squared_error"), seed=5)
    assert np.array_equal(res.fold_id, fold) and np.array_equal(res.oof_votes, ov)
    assert np.allclose(res.y_pred, tst.votes16_to_value(ov, 30, lo, span))
    r2 = 1.0 - ((res.y_pred - y) ** 2).sum() / ((y - y.mean()) ** 2).sum()
    assert 0.3 < r2 < 0.9 and ((res.s >= 0) & (res.s <= 1)).all() and abs(res.feature_importances.sum() - 1.0) < 1e-9
    # sklearn-like regressor: fully grown trees reproduce their training targets (bootstrap: ~63 % of the rows per tree)
    reg = tst.TsTForest(n_estimators=25, criterion="squared_error", max_features=1.0, random_state=2).fit(X, y)
    pred = reg.predict(X)
    assert 1.0 - ((pred - y) ** 2).sum() / ((y - y.mean()) ** 2).sum() > 0.85
    gn, gc = reg.export_tree(3)
    S = gc[:, 0].astype(np.uint64) | (gc[:, 1].astype(np.uint64) << np.uint64(32))
    assert S[0] == S[gn["left"][0]] + S[gn["right"][0]] and gn["n_weighted"][0] == 3000


# --------------------------------------------------------------------------------------------- full size (BASELINE.json configs[4])
@pytest.fixture(scope="module")
def full_size():
    """The headline table (BASELINE.json configs[4]): 1M x 64, 4 classes, 10 folds, binned and folded on the device."""
    import tst
    from tst import synth

    torch = _torch()
    n, d, C, k, T = 1_000_000, 64, 4, 10, 1000
    X, y = synth.scaleup_table(n, d, C, seed=0)
    e = tst.Engine(0)
    e.bin(X)
    assert e.n_edges.min() == 254  # continuous columns: all 255 bins used
    y_dev = e.to_dev(y, torch.uint8)
    active = torch.ones(n, dtype=torch.uint8, device=e.dev)
    fold = e.assign_folds(y_dev, active, k, C)
    yield dict(eng=e, y=y, y_dev=y_dev, fold=fold, n=n, d=d, C=C, k=k, T=T)
    e.close()


def test_full_size_bit_exact_vs_oracle(full_size):
    """1M x 64, 10 folds, trees t in [0, 2) of every fold (20 trees of ~295k nodes, root weight 900k > 2^16): GPU votes,
    every work counter and one exported tree against the CPU oracle (OpenMP; ~1 s per tree per core)."""
    import tst

    fs = full_size
    eng, d, C, k, T = fs["eng"], fs["d"], fs["C"], fs["k"], fs["T"]
    bins = eng.get_bins()
    fold_h = fs["fold"].cpu().numpy()
    assert np.array_equal(fold_h, H.assign_folds(fs["y"], np.ones(fs["n"], np.uint8), k, C))
    ov, ost, forest = H.fit_oof(bins, fs["y"], fold_h, k, T, C, H.make_params(d), seed=3, t0=0, t1=2, keep=True)
    votes = eng.new_votes(C)
    eng.stats(reset=True)
    eng.fit_oof(fs["y_dev"], fs["fold"], k, T, C, tst.make_params(d), seed=3, t0=0, t1=2, votes=votes)
    gst = eng.stats()
    assert np.array_equal(votes.cpu().numpy().view(np.uint32), ov)
    check_stats(gst, ost)
    assert gst["max_depth"] == ost["max_depth"] and gst["trees"] == 20
    # one whole tree, node by node (fold 7, tree 1), from a kept fit of the same (seed, gid)
    e2 = tst.Engine(0)
    try:
        e2.set_bins(bins)
        e2.fit_oof(fs["y_dev"], fs["fold"], k, T, C, tst.make_params(d, keep_forest=True, trees_in_flight=4), seed=3, t0=1, t1=2)
        gn, gc = e2.export_tree(7, 1, C)
    finally:
        e2.close()
    ot = forest.tree(7, 1)
    compare_tree(gn, gc, ot.nodes, ot.counts, "1M x 64 tree (7, 1)")
    assert int(ot.nodes["n_weighted"][0]) == 900_000


def test_full_size_properties(full_size):
    """1M x 64, 10 folds, beyond the oracle comparison above: the size-independent properties of the domain --
    determinism, independence from batching / tree ranges / rank count (integer votes), Q0.16 vote
    mass per held-out row, binary-tree counters, fully grown trees reproducing their in-bag labels."""
    import tst

    torch = _torch()
    fs = full_size
    eng, y_dev, fold, n, d, C, k, T = fs["eng"], fs["y_dev"], fs["fold"], fs["n"], fs["d"], fs["C"], fs["k"], fs["T"]
    sizes = torch.bincount(fold.long(), minlength=k)
    assert int(sizes.max() - sizes.min()) <= 1
    prm = tst.make_params(d)

    def fit(t0, t1, tif=0, rank=0, world=1, e=eng):
        votes = e.new_votes(C)
        p = tst.make_params(d, trees_in_flight=tif)
        e.stats(reset=True)
        e.fit_oof(y_dev, fold, k, T, C, p, seed=3, t0=t0, t1=t1, votes=votes)
        return votes, e.stats()

    v_all, st = fit(0, 2)
    assert st["trees"] == 20 and st["nodes"] == 2 * st["leaves"] - st["trees"]
    assert st["row_visits_big"] > st["row_visits_mid"] > st["row_visits_small"] > 0 and st["launches_root"] > 0
    # every held-out row received 2 trees x 65536 of vote mass (floor rounding loses < C units per tree)
    mass = v_all.sum(1)
    assert int(mass.max()) <= 2 * 65536 and int(mass.min()) > 2 * 65536 - 2 * C
    # determinism and batching independence
    v_again, _ = fit(0, 2)
    v_b3, _ = fit(0, 2, tif=3)
    assert torch.equal(v_all, v_again) and torch.equal(v_all, v_b3)
    # tree ranges compose
    v0, _ = fit(0, 1)
    v1, _ = fit(1, 2)
    assert torch.equal(v_all, v0 + v1)
    # two emulated ranks sum to the single-rank result
    parts = []
    for r in range(2):
        e2 = tst.Engine(0, r, 2)
        e2.set_bins(eng.get_bins())
        vv = e2.new_votes(C)
        e2.fit_oof(y_dev, fold, k, T, C, prm, seed=3, t0=0, t1=2, votes=vv)
        parts.append(vv)
        e2.close()
    assert torch.equal(v_all, parts[0] + parts[1])
    # blind accuracy is far above chance on this learnable table and s(x) is a probability
    s, cor, tot = eng.finalize(v_all, y_dev, fold, k, 2, C)
    acc = cor.double().sum() / tot.double().sum()
    assert 0.4 < float(acc) < 0.9 and float(s.min()) >= 0 and float(s.max()) <= 1
