"""Host-side feature extraction and interpretability helpers; the GPU is not needed: the
planted-bias check runs the extracted table through the CPU oracle."""
import numpy as np

from oracle import hist_rf as H


def make_records(n=600, seed=0, bias=0.75):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        if i % 3 == 0:  # numeric-option question: the largest number is right with probability `bias`
            vals = rng.choice(np.arange(1, 40), 4, replace=False)
            ans = int(np.argmax(vals)) if rng.random() < bias else int(rng.integers(0, 4))
            recs.append(dict(question=f"How many chairs are in the room shown at minute {rng.integers(1, 9)}?",
                             options=[str(v) for v in vals], answer=ans, question_type="object_count"))
        elif i % 3 == 1:  # text options: the longest option is right with probability `bias`
            opts = ["left", "right", "behind the sofa", "up"]
            rng.shuffle(opts)
            ans = int(np.argmax([len(o) for o in opts])) if rng.random() < bias else int(rng.integers(0, 4))
            recs.append(dict(question="Where is the lamp relative to the table?", options=opts, answer="ABCD"[ans],
                             question_type="relative_direction"))
        else:  # open numeric answer
            recs.append(dict(question=f"What is the distance between the door and the window in meters, not counting the {rng.integers(2, 5)} steps?",
                             options=None, answer=float(rng.gamma(2.0, 2.0)), question_type="abs_distance"))
    return recs


def test_extract_features_shapes_and_determinism():
    from tst import extract_features

    recs = make_records()
    X, y, names, group, classes = extract_features(recs)
    X2, y2, names2, group2, _ = extract_features(recs)
    assert X.shape == (600, len(names)) and X.dtype == np.float32 and y.dtype == np.uint8
    assert np.array_equal(np.nan_to_num(X, nan=-7), np.nan_to_num(X2, nan=-7)) and np.array_equal(y, y2) and names == names2
    assert set(group.tolist()) == {0, 1, 2} and len(set(names)) == len(names)
    assert "q_tokens" in names and "opt0_number_rank" in names and "longest_option" in names
    assert np.isnan(X[2, names.index("opt0_chars")])  # open question: option features are missing -> NaN bin
    assert max(classes) <= 3


def test_planted_non_visual_bias_is_found():
    """A blind forest on the extracted features beats chance by a wide margin on the biased question types."""
    from tst import extract_features

    X, y, names, group, classes = extract_features(make_records(900, seed=3))
    C = len(classes)
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    fold = H.assign_folds(y, np.ones(len(y), np.uint8), 5, C, group, 3)
    votes, _, _ = H.fit_oof(bins, y, fold, 5, 30, C, H.make_params(X.shape[1]), seed=0, group=group, n_groups=3)
    s, cor, tot = H.finalize(votes, y, fold, 5, 30, group, 3)
    acc = (cor.reshape(3, 5).sum(1) / tot.reshape(3, 5).sum(1))
    types = sorted({"object_count", "relative_direction", "abs_distance"})
    assert acc[types.index("object_count")] > 0.6 and acc[types.index("relative_direction")] > 0.6  # chance = 0.25
    assert s[group == types.index("object_count")].mean() > 0.5


def test_gini_importances_and_sklearn_arrays():
    from tst import gini_importances, to_sklearn_arrays

    rng = np.random.default_rng(0)
    bins = rng.integers(0, 200, (6, 800)).astype(np.uint8)
    y = (bins[2] > 100).astype(np.uint8)  # only feature 2 matters
    fold = np.full(800, H.FOLD_TRAIN_ONLY, np.uint8)
    _, _, forest = H.fit_oof(bins, y, fold, 1, 8, 2, H.make_params(6, max_features=None), seed=1, keep=True)
    imp = np.mean([gini_importances(forest.tree(0, t).nodes, forest.tree(0, t).counts, 6) for t in range(8)], axis=0)
    assert abs(imp.sum() - 1.0) < 1e-9 and imp.argmax() == 2 and imp[2] > 0.95
    tr = forest.tree(0, 0)
    sk = to_sklearn_arrays(tr.nodes, tr.counts)
    leaves = sk["children_left"] == -1
    assert (sk["feature"][leaves] == -2).all() and np.allclose(sk["value"].sum(-1), 1.0)
    assert sk["n_node_samples"][0] == tr.nodes["end"][0] and (sk["impurity"][leaves] == 0).all()


def test_path_contributions_telescope_to_the_leaf_probability():
    """Per-feature attribution: prior + sum of contributions == leaf P(class), and the mass
    lands on the columns that carry the label."""
    from oracle import hist_rf as H
    from tst import binning, path_contributions

    rng = np.random.default_rng(0)
    n, d = 1500, 8
    X = rng.standard_normal((n, d)).astype(np.float32)
    y = ((X[:, 2] > 0.2).astype(np.uint8) + 2 * (X[:, 5] > -0.1).astype(np.uint8)).astype(np.uint8)  # 4 classes from columns 2 and 5
    edges, ne = H.find_edges(X)
    bins = H.bin_table(X, edges, ne)
    assert np.array_equal(binning.bin_rows(X, edges, ne), bins.T)  # host bin map == oracle bin map
    fold = np.full(n, H.FOLD_TRAIN_ONLY, np.uint8)
    _, _, forest = H.fit_oof(bins, y, fold, 1, 6, 4, H.make_params(d), seed=3, keep=True)
    rows = bins.T[:300]
    total = np.zeros(d)
    for t in range(6):
        tr = forest.tree(0, t)
        prior, contrib = path_contributions(tr.nodes, tr.counts, rows, y[:300], d)
        # leaf probability of the true class, by walking the tree independently
        leaf = np.zeros(300)
        for i in range(300):
            nd = 0
            while tr.nodes["feature"][nd] >= 0:
                nd = tr.nodes["left"][nd] if rows[i, tr.nodes["feature"][nd]] <= tr.nodes["bin"][nd] else tr.nodes["right"][nd]
            leaf[i] = tr.counts[nd, y[i]] / tr.counts[nd].sum()
        assert np.allclose(prior + contrib.sum(1), leaf, atol=1e-12)
        total += np.abs(contrib).sum(0)
    assert set(np.argsort(total)[-2:]) == {2, 5}


def test_benchmark_record_adapters():
    """Row layouts of the four benchmarks named at REF/README.md:14-15 -> records -> feature table."""
    import pytest
    from tst import extract_features, records_from

    vsi = [dict(id=1, dataset="arkitscenes", scene_name="s1", question_type="object_rel_direction_easy",
                question="If I am standing by the sofa and facing the TV, is the lamp to my left or right?", options=["A. left", "B. right"], ground_truth="B"),
           dict(id=2, dataset="scannet", scene_name="s2", question_type="object_counting", question="How many chairs are in this room?", options=None, ground_truth="4"),
           dict(id=3, dataset="scannet", scene_name="s2", question_type="room_size_estimation", question="What is the size of this room (in square meters)?", options=None, ground_truth=23.5)]
    r = records_from("VSI-Bench", vsi)
    assert r[0]["options"] == ["left", "right"] and r[0]["answer"] == 1 and r[1]["options"] is None and r[1]["answer"] == 4.0 and r[2]["answer"] == 23.5
    cvb = [dict(idx=0, type="2D", task="Count", question="How many cups are in the image?", choices=["2", "3", "4", "5"], answer="(C)", source="ADE20K"),
           dict(idx=1, type="3D", task="Depth", question="Which object is closer to the camera, the chair or the desk?", choices=["chair", "desk"], answer="(A)", source="Omni3D")]
    r = records_from("cvbench", cvb)
    assert r[0]["answer"] == 2 and r[0]["question_type"] == "Count" and r[1]["answer"] == 0
    mmmu = [dict(id="validation_Math_1", question="What is the value of x in <image 1>?", options="['1', '2', '3', '4']", answer="D", question_type="multiple-choice", subfield="Algebra"),
            dict(id="validation_Physics_2", question="Compute the force in newtons.", options="[]", answer="12.5", question_type="open", subfield="Mechanics"),
            dict(id="validation_Art_3", question="Name the painter.", options="[]", answer="Monet", question_type="open", subfield="Art History")]
    r = records_from("MMMU", mmmu)
    assert len(r) == 2 and r[0]["answer"] == 3 and r[0]["question_type"] == "Algebra" and r[1]["answer"] == 12.5  # the free-text open question is skipped
    vmme = [dict(video_id="001", duration="short", domain="Knowledge", sub_category="Humanity & History", question_id="001-1", task_type="Counting Problem",
                 question="How many people appear in the video?", options=["A. 1.", "B. 2.", "C. 3.", "D. 4."], answer="C")]
    r = records_from("Video-MME", vmme)
    assert r[0]["answer"] == 2 and r[0]["options"][1] == "2." and r[0]["question_type"] == "Counting Problem/short"
    X, y, names, group, classes = extract_features(records_from("vsibench", vsi) + records_from("cvbench", cvb) + records_from("videomme", vmme))
    assert X.shape[0] == 6 and len(names) == X.shape[1] and group.max() >= 3
    with pytest.raises(KeyError, match="ground_truth"):
        records_from("vsibench", [dict(question="q", question_type="t")])
    with pytest.raises(ValueError, match="unknown benchmark"):
        records_from("imagenet", [])
