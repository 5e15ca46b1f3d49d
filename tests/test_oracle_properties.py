"""Properties of the CPU oracle itself (edge cases the domain has: empty/ragged inputs, single class,
duplicates, tiny folds), plus hypothesis-driven invariants on small random tables."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import hist_rf as H


def fit(bins, y, k, T, C, seed=0, keep=False, nthreads=0, **pk):
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), k, C)
    votes, stt, forest = H.fit_oof(bins, y, fold, k, T, C, H.make_params(bins.shape[0], **pk), seed=seed, keep=keep, nthreads=nthreads)
    return fold, votes, stt, forest


def check_tree_invariants(tree, C):
    n, c = tree.nodes, tree.counts
    assert n["depth"][0] == 0 and n["begin"][0] == 0
    order = np.lexsort((n["begin"], n["depth"]))
    assert np.array_equal(order, np.arange(len(n)))  # canonical breadth-first order
    for i in range(len(n)):
        assert c[i].sum() == n["n_weighted"][i]
        if n["feature"][i] >= 0:
            l, r = n["left"][i], n["right"][i]
            assert n["begin"][l] == n["begin"][i] and n["end"][l] == n["begin"][r] and n["end"][r] == n["end"][i]
            assert n["depth"][l] == n["depth"][i] + 1 and np.array_equal(c[l] + c[r], c[i])
            assert n["end"][l] > n["begin"][l] and n["end"][r] > n["begin"][r]
        else:
            assert n["left"][i] == -1 and n["right"][i] == -1


def test_thread_count_does_not_change_results(vsi_small):
    v = vsi_small
    a = H.fit_oof(v["bins"], v["y"], v["fold"], 5, 6, 4, H.make_params(16), seed=1, nthreads=1)[0]
    b = H.fit_oof(v["bins"], v["y"], v["fold"], 5, 6, 4, H.make_params(16), seed=1, nthreads=8)[0]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("world", [2, 4])
def test_rank_shards_sum_to_the_whole(vsi_small, world):
    v = vsi_small
    whole = H.fit_oof(v["bins"], v["y"], v["fold"], 5, 6, 4, H.make_params(16), seed=2)[0]
    parts = sum(H.fit_oof(v["bins"], v["y"], v["fold"], 5, 6, 4, H.make_params(16), seed=2, rank=r, world=world)[0].astype(np.uint64)
                for r in range(world))
    assert np.array_equal(parts, whole.astype(np.uint64))


def test_tree_ranges_compose_to_the_whole(vsi_small):
    v = vsi_small
    a = H.fit_oof(v["bins"], v["y"], v["fold"], 5, 8, 4, H.make_params(16), seed=3)[0]
    b = H.fit_oof(v["bins"], v["y"], v["fold"], 5, 8, 4, H.make_params(16), seed=3, t0=0, t1=3)[0]
    c = H.fit_oof(v["bins"], v["y"], v["fold"], 5, 8, 4, H.make_params(16), seed=3, t0=3, t1=8)[0]
    assert np.array_equal(a, b + c)


def test_full_depth_trees_memorise_distinct_rows():
    rng = np.random.default_rng(0)
    bins = rng.integers(0, 256, (6, 500)).astype(np.uint8)  # all rows distinct with overwhelming probability
    y = rng.integers(0, 3, 500).astype(np.uint8)
    fold = np.full(500, H.FOLD_TRAIN_ONLY, np.uint8)
    _, stt, forest = H.fit_oof(bins, y, fold, 1, 5, 3, H.make_params(6), seed=0, keep=True)
    votes = forest.predict_votes(bins)
    for t in range(5):
        check_tree_invariants(forest.tree(0, t), 3)
    # every in-bag row reaches a pure leaf of its own class; 5 trees x 63 % in-bag -> the true class leads
    assert (np.argmax(votes, 1) == y).mean() > 0.9
    assert stt["leaves"] == stt["nodes"] - stt["searched"] + (stt["searched"] - (stt["nodes"] - stt["leaves"]))


def test_votes_are_q16_and_bounded(vsi_small):
    v = vsi_small
    T = 7
    votes = H.fit_oof(v["bins"], v["y"], v["fold"], 5, T, 4, H.make_params(16), seed=4)[0]
    tot = votes.sum(1)
    assert (tot <= T * 65536).all() and (tot > T * 65536 - T * 4).all()  # floor rounding loses < 1 unit per class per tree
    s, cor, total = H.finalize(votes, v["y"], v["fold"], 5, T)
    assert ((s >= 0) & (s <= 1)).all() and total.sum() == len(v["y"])


def test_edge_cases():
    rng = np.random.default_rng(1)
    # single class: every tree is one pure leaf, s == 1
    bins = rng.integers(0, 9, (3, 50)).astype(np.uint8)
    y = np.zeros(50, np.uint8)
    fold, votes, stt, _ = fit(bins, y, 5, 3, 1)
    assert stt["nodes"] == stt["trees"] == 15 and (votes[:, 0] == 3 * 65536).all()
    # identical rows, conflicting labels: no split possible -> mixed root leaves, soft votes
    bins = np.zeros((2, 40), np.uint8)
    y = (np.arange(40) % 2).astype(np.uint8)
    fold, votes, stt, forest = fit(bins, y, 4, 2, 2, keep=True)
    assert stt["nodes"] == stt["trees"] and stt["searched"] == stt["trees"] and stt["fallback"] == stt["trees"]  # m = 1 of d = 2
    assert (votes.sum(1) >= 2 * 65536 - 2).all() and (votes > 0).all()
    # fewer rows than folds: empty folds are skipped, one-row training sets give single-leaf trees
    bins = rng.integers(0, 5, (2, 3)).astype(np.uint8)
    y = np.array([0, 1, 0], np.uint8)
    fold, votes, stt, _ = fit(bins, y, 5, 2, 2)
    assert stt["trees"] == 10 and votes.sum() == 3 * 2 * 65536
    # one row in the table: its training set is empty, the tree is an empty leaf and casts no vote
    fold, votes, stt, _ = fit(np.zeros((1, 1), np.uint8), np.zeros(1, np.uint8), 2, 2, 1)
    assert votes.sum() == 0 and stt["inbag_rows"] == 2  # only the other fold's two trees see the row
    # max_depth and min_samples_leaf are honoured
    bins = rng.integers(0, 200, (5, 400)).astype(np.uint8)
    y = rng.integers(0, 4, 400).astype(np.uint8)
    fold, votes, stt, forest = fit(bins, y, 2, 3, 4, keep=True, max_depth=3, min_samples_leaf=5)
    assert stt["max_depth"] <= 3
    for t in range(3):
        tr = forest.tree(0, t)
        check_tree_invariants(tr, 4)
        assert (tr.nodes["n_weighted"][tr.nodes["feature"] < 0] >= 5).all()


def test_low_cardinality_forces_fallback_draws():
    rng = np.random.default_rng(3)
    bins = np.zeros((16, 300), np.uint8)
    bins[5] = rng.integers(0, 2, 300)  # only one of 16 features carries information; m = 4
    y = (bins[5] ^ (rng.random(300) < 0.1)).astype(np.uint8)
    fold, votes, stt, forest = fit(bins, y, 3, 4, 2, keep=True)
    assert stt["fallback"] > 0
    for t in range(4):
        tr = forest.tree(1, t)
        assert set(tr.nodes["feature"][tr.nodes["feature"] >= 0]) == {5}


def test_ibp_invariants(vsi_small):
    v = vsi_small
    act, removed, s = H.ibp(v["bins"], v["y"], 5, 8, 4, H.make_params(16), seed=5, rounds=4, tau=0.5, batch=37)
    assert (removed <= 37).all() and act.sum() == len(act) - removed.sum()
    gone = act == 0
    # a removed row was scored >= tau in the round that removed it, and is unscored (-1) afterwards
    first_unscored = np.argmax(s < 0, axis=0)
    for i in np.flatnonzero(gone)[:50]:
        r = first_unscored[i] - 1 if (s[:, i] < 0).any() else len(removed) - 1
        assert s[r, i] >= 0.5
    act2, removed2, _ = H.ibp(v["bins"], v["y"], 5, 8, 4, H.make_params(16), seed=5, rounds=4, tau=2.0, batch=37)
    assert removed2.sum() == 0 and act2.all()  # nothing reaches tau: stops immediately


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 120), d=st.integers(1, 9), C=st.integers(1, 5), card=st.integers(1, 40), k=st.integers(1, 4),
       seed=st.integers(0, 2**31 - 1), msl=st.integers(1, 3))
def test_random_small_tables(n, d, C, card, k, seed, msl):
    rng = np.random.default_rng(seed)
    bins = rng.integers(0, card, (d, n)).astype(np.uint8)
    y = rng.integers(0, C, n).astype(np.uint8)
    fold = H.assign_folds(y, np.ones(n, np.uint8), k, C)
    prm = H.make_params(d, min_samples_leaf=msl)
    votes, stt, forest = H.fit_oof(bins, y, fold, k, 2, C, prm, seed=seed, keep=True)
    assert stt["nodes"] == 2 * stt["leaves"] - stt["trees"]  # binary trees
    for p in range(k):
        for t in range(2):
            check_tree_invariants(forest.tree(p, t), C)
    if k > 1:
        assert (votes.sum(1) <= 2 * 65536).all()


# --------------------------------------------------------------------------------------------- fp32 gain estimate (tst_common.cuh)
def _f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32)


def _estimate_gini(SL, NL, SR, NR):
    """tst_common.cuh: gain_estimate restated in IEEE float32 (numpy), the division exact-then-rounded."""
    SLf, NLf, SRf, NRf = _f32(SL), _f32(NL), _f32(SR), _f32(NR)
    t = (SRf * NLf).astype(np.float32)
    num = _f32(SLf.astype(np.float64) * NRf.astype(np.float64) + t.astype(np.float64))  # fmaf: one rounding
    den = (NLf * NRf).astype(np.float32)
    return (num / den).astype(np.float32)


def _estimate_mse(SL, NL, SR, NR):
    SLf, NLf, SRf, NRf = _f32(SL), _f32(NL), _f32(SR), _f32(NR)
    a = (SLf / NLf).astype(np.float32)
    b = (SRf / NRf).astype(np.float32)
    t = (SRf * b).astype(np.float32)
    return _f32(SLf.astype(np.float64) * a.astype(np.float64) + t.astype(np.float64))


@pytest.mark.parametrize("scale", [6, 200, 60000, 5_000_000, 2_000_000_000])
def test_gain_estimate_error_is_far_inside_the_margin(scale):
    """The CUDA split search evaluates the exact fp64 gain only for candidates whose fp32 estimate is within
    TST_EST_MARGIN = 1 - 2^-18 of the best estimate.  That is only sound if the estimate's relative error is well
    below 2^-19; the arithmetic bound is 2^-21 (+ 2 ulp = 2^-22 for __fdividef's reciprocal, CUDA C Programming
    Guide, intrinsic error table).  Checked here on random class-count configurations of every node size."""
    rng = np.random.default_rng(scale)
    C = 4
    L = rng.integers(0, scale, (20000, C)).astype(object)
    R = rng.integers(0, scale, (20000, C)).astype(object)
    ok = (L.sum(1) > 0) & (R.sum(1) > 0)
    L, R = L[ok], R[ok]
    NL, NR = L.sum(1), R.sum(1)
    SL, SR = (L * L).sum(1), (R * R).sum(1)
    exact = np.array([float(a) / float(b) + float(c) / float(d) for a, b, c, d in zip(SL, NL, SR, NR)])
    est = _estimate_gini(SL.astype(np.float64), NL.astype(np.float64), SR.astype(np.float64), NR.astype(np.float64)).astype(np.float64)
    rel = np.abs(est - exact) / exact
    assert rel.max() < 2.0 ** -21, rel.max()
    # squared-error proxy on ordinal levels (S = sum_c c * count_c)
    lev = np.arange(C, dtype=object)
    S_L, S_R = (L * lev).sum(1), (R * lev).sum(1)
    keep = (S_L + S_R) > 0
    exact = np.array([float(a) ** 2 / float(b) + float(c) ** 2 / float(d) for a, b, c, d in zip(S_L[keep], NL[keep], S_R[keep], NR[keep])])
    est = _estimate_mse(S_L[keep].astype(np.float64), NL[keep].astype(np.float64), S_R[keep].astype(np.float64), NR[keep].astype(np.float64)).astype(np.float64)
    rel = np.abs(est - exact) / exact
    assert rel.max() < 2.0 ** -21, rel.max()
