"""Hand-crafted NON-VISUAL features from benchmark question/answer records.

REF/README.md:35: the RF diagnostic operates "on hand-crafted features" of the test set's textual side.
The reference ships neither the feature lists nor data, so this is a generic, interpretable extractor:
everything a blind test-taker could compute from the question text, the answer options and the
question type -- never from the image/video and never from the identity of the correct answer.

    X, y, names, group, classes = extract_features(records)
    res = tst.run_tst(X, y, k=5, n_trees=100, group=group)     # per-question-type forests
    res.s                                                       # bias score s(x) per record

A record is a dict with
    question : str
    options  : list[str] (multiple choice) or None (open numeric answer)
    answer   : index into options | option letter | option text   (MC)   or a number (numeric)
    question_type : str (optional; becomes the group id)
Numeric answers are discretised into `numeric_bins` quantile classes so the classifier path applies.
"""
from __future__ import annotations

import re
import zlib

import numpy as np

_NUM = re.compile(r"[-+]?\d+(?:\.\d+)?")
_TOK = re.compile(r"[a-z]+|\d+")


def _numbers(text: str):
    return [float(x) for x in _NUM.findall(text or "")]


def _tokens(text: str):
    return _TOK.findall((text or "").lower())


def _answer_index(rec) -> int:
    opts, ans = rec["options"], rec["answer"]
    if isinstance(ans, (int, np.integer)):
        return int(ans)
    a = str(ans).strip()
    if len(a) == 1 and a.upper() in "ABCDEFGHIJ" and len(opts) > "ABCDEFGHIJ".index(a.upper()):
        return "ABCDEFGHIJ".index(a.upper())
    low = [str(o).strip().lower() for o in opts]
    if a.lower() in low:
        return low.index(a.lower())
    raise ValueError(f"cannot resolve answer {ans!r} among {opts!r}")


def extract_features(records, max_options: int = 5, vocab_size: int = 16, numeric_bins: int = 4):
    """-> (X float32 [n, d], y uint8 [n], names list[str], group int32 [n], classes list)."""
    n = len(records)
    is_mc = [r.get("options") is not None for r in records]
    names, cols = [], []

    def add(name, values):
        names.append(name)
        cols.append(np.asarray(values, dtype=np.float64))

    # ---- question-level features
    toks = [_tokens(r["question"]) for r in records]
    nums = [_numbers(r["question"]) for r in records]
    add("q_chars", [len(r["question"]) for r in records])
    add("q_tokens", [len(t) for t in toks])
    add("q_numbers", [len(v) for v in nums])
    add("q_max_number", [max(v) if v else np.nan for v in nums])
    add("q_min_number", [min(v) if v else np.nan for v in nums])
    add("q_has_negation", [int(any(w in ("not", "no", "never", "except") for w in t)) for t in toks])
    add("q_is_how_many", [int(t[:2] == ["how", "many"]) for t in toks])
    add("q_first_token_hash", [zlib.crc32(t[0].encode()) % 97 if t else -1 for t in toks])
    # most frequent content tokens over the test set (the "vocabulary" a text-only model could exploit)
    freq = {}
    for t in toks:
        for w in set(t):
            freq[w] = freq.get(w, 0) + 1
    vocab = [w for w, _ in sorted(freq.items(), key=lambda kv: (-kv[1], kv[0])) if len(w) > 2][:vocab_size]
    for w in vocab:
        add(f"q_has[{w}]", [int(w in t) for t in toks])
    # ---- option-level features (multiple choice)
    n_opt = [len(r["options"]) if m else 0 for r, m in zip(records, is_mc)]
    add("n_options", n_opt)
    for j in range(max_options):
        texts = [str(r["options"][j]) if m and j < len(r["options"]) else None for r, m in zip(records, is_mc)]
        vals = [(_numbers(t) or [np.nan])[0] if t is not None else np.nan for t in texts]
        add(f"opt{j}_chars", [len(t) if t is not None else np.nan for t in texts])
        add(f"opt{j}_number", vals)
        rank = []
        for r, m, v in zip(records, is_mc, vals):
            if not m or v != v:
                rank.append(np.nan)
                continue
            others = [(_numbers(str(o)) or [np.nan])[0] for o in r["options"]]
            rank.append(float(sum(1 for o in others if o == o and o < v)))
        add(f"opt{j}_number_rank", rank)
        add(f"opt{j}_overlap_with_question", [len(set(_tokens(t)) & set(q)) if t is not None else np.nan for t, q in zip(texts, toks)])
    longest = [int(np.argmax([len(str(o)) for o in r["options"]])) if m else -1 for r, m in zip(records, is_mc)]
    add("longest_option", longest)
    # ---- labels and groups
    types = sorted({r.get("question_type", "") for r in records})
    group = np.array([types.index(r.get("question_type", "")) for r in records], dtype=np.int32)
    add("question_type", group)
    y = np.zeros(n, dtype=np.int64)
    numeric = [float(r["answer"]) for r, m in zip(records, is_mc) if not m]
    if numeric:
        qs = np.quantile(numeric, np.linspace(0, 1, numeric_bins + 1)[1:-1])
    for i, (r, m) in enumerate(zip(records, is_mc)):
        y[i] = _answer_index(r) if m else int(np.searchsorted(qs, float(r["answer"]), side="right"))
    classes = sorted(set(y.tolist()))
    X = np.stack(cols, axis=1).astype(np.float32)
    return X, y.astype(np.uint8), names, group, classes


# ------------------------------------------------------------------------------------------------- benchmark record schemas
# Adapters from the row layouts of the four benchmarks the paper diagnoses (REF/README.md:35) to the record dict above.  The
# field names follow the public Hugging Face datasets named at REF/README.md:14-15 AS REMEMBERED -- the datasets are not
# reachable offline, so every adapter only touches the fields listed in its docstring and raises a KeyError that names the
# missing field; nothing visual (image / video paths) is ever read.
_LETTERS = "ABCDEFGHIJ"
_OPT_PREFIX = re.compile(r"^\s*\(?([A-Ja-j])[\).:]\s*")


def _strip_option_prefix(o) -> str:
    return _OPT_PREFIX.sub("", str(o)).strip()


def _letter_index(ans) -> int | None:
    a = str(ans).strip().strip("()").strip(".").upper()
    return _LETTERS.index(a) if len(a) == 1 and a in _LETTERS else None


def _need(row, *names):
    for nm in names:
        if nm not in row:
            raise KeyError(f"record has no field {nm!r} (fields: {sorted(row)})")
    return [row[nm] for nm in names]


def from_vsibench(row) -> dict:
    """VSI-Bench (nyu-visionx/VSI-Bench): question, question_type, options (list of 'A. ...' or None for numeric answers),
    ground_truth (option letter, or a number for the numeric-answer types: object counting, distances, room size)."""
    q, qt, gt = _need(row, "question", "question_type", "ground_truth")
    opts = row.get("options")
    if opts:
        opts = [_strip_option_prefix(o) for o in opts]
        idx = _letter_index(gt)
        return dict(question=q, options=opts, answer=idx if idx is not None else str(gt), question_type=qt)
    return dict(question=q, options=None, answer=float(gt), question_type=qt)


def from_cvbench(row) -> dict:
    """CV-Bench (nyu-visionx/CV-Bench): question, choices (list), answer ('(A)'), task (Count / Relation / Depth / Distance)."""
    q, ch, ans, task = _need(row, "question", "choices", "answer", "task")
    idx = _letter_index(ans)
    return dict(question=q, options=[str(c) for c in ch], answer=idx if idx is not None else str(ans), question_type=str(task))


def from_mmmu(row) -> dict:
    """MMMU (MMMU/MMMU): question, options (a list or its string repr), answer (letter; free text for open questions),
    question_type ('multiple-choice' / 'open'), subfield.  Open questions without a numeric answer are skipped (None)."""
    import ast

    q, opts, ans = _need(row, "question", "options", "answer")
    if isinstance(opts, str):
        opts = ast.literal_eval(opts) if opts.strip() else []
    qt = str(row.get("subfield", row.get("question_type", "")))
    if opts:
        idx = _letter_index(ans)
        return dict(question=q, options=[str(o) for o in opts], answer=idx if idx is not None else str(ans), question_type=qt)
    try:
        return dict(question=q, options=None, answer=float(str(ans).replace(",", "")), question_type=qt)
    except ValueError:
        return None


def from_videomme(row) -> dict:
    """Video-MME (lmms-lab/Video-MME): question, options (list of 'A. ...'), answer (letter), task_type, duration (short / medium /
    long: a non-visual property of the item, kept in the question type)."""
    q, opts, ans, task = _need(row, "question", "options", "answer", "task_type")
    qt = f"{task}/{row['duration']}" if "duration" in row else str(task)
    idx = _letter_index(ans)
    return dict(question=q, options=[_strip_option_prefix(o) for o in opts], answer=idx if idx is not None else str(ans), question_type=qt)


ADAPTERS = {"vsibench": from_vsibench, "cvbench": from_cvbench, "mmmu": from_mmmu, "videomme": from_videomme}


def records_from(benchmark: str, rows) -> list:
    """rows of one of the four benchmarks (dicts, e.g. a datasets.Dataset iterated) -> records for extract_features."""
    try:
        f = ADAPTERS[benchmark.lower().replace("-", "").replace("_", "")]
    except KeyError:
        raise ValueError(f"unknown benchmark {benchmark!r}; known: {sorted(ADAPTERS)}") from None
    return [r for r in (f(row) for row in rows) if r is not None]


# ------------------------------------------------------------------------------------------------- interpretability
def gini_importances(nodes: np.ndarray, counts: np.ndarray, n_features: int) -> np.ndarray:
    """Mean decrease in impurity of one exported tree (SK/tree/_tree.pyx:1244-1280), on bootstrap-weighted counts."""
    imp = np.zeros(n_features, dtype=np.float64)
    c = counts.astype(np.float64)
    N = c.sum(1)
    gini = 1.0 - np.divide((c * c).sum(1), N * N, out=np.ones_like(N), where=N > 0)
    for i in np.flatnonzero(nodes["feature"] >= 0):
        l, r = nodes["left"][i], nodes["right"][i]
        imp[nodes["feature"][i]] += N[i] * gini[i] - N[l] * gini[l] - N[r] * gini[r]
    if N[0] > 0:
        imp /= N[0]
    s = imp.sum()
    return imp / s if s > 0 else imp


def to_sklearn_arrays(nodes: np.ndarray, counts: np.ndarray, edges: np.ndarray | None = None, n_edges: np.ndarray | None = None) -> dict:
    """The field set of sklearn's Tree (SK/tree/_tree.pxd:15-25) from a canonical export; thresholds are
    the raw-space bin edges when `edges` [d, 255] is given, else the bin index.  A split at bin >= n_edges[feature]
    separates every finite value from the NaN bin: its raw threshold is +inf (x <= inf goes left, NaN goes right)."""
    internal = nodes["feature"] >= 0
    thr = nodes["bin"].astype(np.float64)
    if edges is not None:
        f, b = np.maximum(nodes["feature"], 0), np.maximum(nodes["bin"], 0)
        ne = np.full(edges.shape[0], edges.shape[1], dtype=np.int64) if n_edges is None else np.asarray(n_edges, dtype=np.int64)
        thr = np.where(b < ne[f], edges[f, np.minimum(b, edges.shape[1] - 1)], np.inf)
    c = counts.astype(np.float64)
    N = c.sum(1)
    return dict(children_left=nodes["left"].copy(), children_right=nodes["right"].copy(),
                feature=np.where(internal, nodes["feature"], -2), threshold=np.where(internal, thr, -2.0),
                value=np.divide(c, N[:, None], out=np.zeros_like(c), where=N[:, None] > 0)[:, None, :],
                n_node_samples=(nodes["end"] - nodes["begin"]).astype(np.int64), weighted_n_node_samples=N,
                impurity=1.0 - np.divide((c * c).sum(1), N * N, out=np.ones_like(N), where=N > 0))


def path_contributions(nodes: np.ndarray, counts: np.ndarray, bins_rows: np.ndarray, cls: np.ndarray, n_features: int) -> tuple:
    """Per-feature attribution of one tree's blind prediction ("interpretable auditing"):
    walking the decision path of a row, the change of P(cls) from a node to the child it enters is credited to the
    node's split feature.  prior + contributions.sum(1) == P(cls) in the reached leaf, exactly (telescoping sum).

    nodes / counts: a canonical export (tst_export_tree layout); bins_rows [n, d] uint8; cls [n].
    -> (prior [n], contributions [n, n_features])
    """
    c = counts.astype(np.float64)
    N = np.maximum(c.sum(1), 1.0)
    bins_rows = np.asarray(bins_rows)
    cls = np.asarray(cls, dtype=np.int64)
    n = bins_rows.shape[0]
    prior = np.full(n, 0.0)
    contrib = np.zeros((n, n_features), dtype=np.float64)
    feat, thr, left, right = nodes["feature"], nodes["bin"], nodes["left"], nodes["right"]
    cur = np.zeros(n, dtype=np.int64)
    prior[:] = c[0, cls] / N[0]
    active = feat[cur] >= 0
    rows = np.arange(n)
    while active.any():
        i = rows[active]
        nd = cur[i]
        f = feat[nd]
        nxt = np.where(bins_rows[i, f] <= thr[nd], left[nd], right[nd])
        contrib[i, f] += c[nxt, cls[i]] / N[nxt] - c[nd, cls[i]] / N[nd]
        cur[i] = nxt
        active = feat[cur] >= 0
    return prior, contrib
