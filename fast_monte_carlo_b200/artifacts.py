"""Artifact compiler: reference model files -> flat structure-of-arrays forests.

The reference loads its models through xgboost / scikit-learn objects
(fast_monte_carlo_cfb.py:641-668).  Here every model is parsed *without*
xgboost (it is not installed) into a `Forest`: per-node SoA arrays plus the
feature recipe (one-hot groups + the 17 numerics), which the C-ABI library
uploads to the GPU (`fmc_load_forest`, include/fmc.h).

Formats handled
  * XGBoost JSON boosters (pass_stage1_complete_vs_not.json, run_fumble.json,
    optional pass_stage2_notcomplete.json)          -- FMC:641-642
  * play_model.xgb: pickle(XGBClassifier) whose booster handle is UBJSON
    
  * scikit-learn `Pipeline[ColumnTransformer, GradientBoostingRegressor]`
    quantile models (*_yards_q{10,50,90}.joblib)    -- FMC:658-668
  * ColumnTransformer preprocessors for the one-hot category lists
    (pass_stage1_preprocessor.joblib ...)            -- FMC:651-654
  * scaler.pkl / coach_label_encoder.pkl for play_model.xgb.

Nothing in this module evaluates a tree; it only re-lays data out.
"""
from __future__ import annotations

import io
import json
import os
import pickle
import struct
import sys
import types
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# 17 numerics, identical order in every trainer (train_pass_outcome_stage1.py:12-17, FMC:676-682)
NUM_FEATURES = [
    "down", "distance", "yardsToGoal", "is_red_zone", "score_diff", "seconds_remaining",
    "offenseTimeouts", "defenseTimeouts",
    "sp_rating_off", "sp_offense_rating_off", "sp_defense_rating_def", "sp_rating_def",
    "goal_to_go", "fourth_and_short", "fg_range", "half", "two_minute",
]
N_NUM = len(NUM_FEATURES)

# play_model.xgb numerics (first 12 columns)
PLAY_FEATURES = [
    "down", "distance", "yardsToGoal", "is_red_zone", "score_diff", "seconds_remaining",
    "offenseTimeouts", "defenseTimeouts",
    "sp_rating_off", "sp_offense_rating_off", "sp_defense_rating_def", "sp_rating_def",
]
PLAY_CLASSES = ["field_goal", "pass", "punt", "run", "timeout"]

KIND_XGB = 0   # f32 accumulate, go left iff x < thr, optional missing handling
KIND_SKL = 1   # f64 accumulate, go left iff f32(x) <= thr

LINK_IDENTITY = 0
LINK_SIGMOID = 1
LINK_SOFTMAX = 2

# model ids shared with include/fmc.h
MODEL_IDS = {
    "pass_stage1": 0,
    "pass_stage2": 1,
    "pass_yards": 2,   # q10,q50,q90 as three "classes" of one forest
    "run_yards": 3,
    "sack_yards": 4,
    "play_model": 5,
    "play_binary": 5,  # play_model.json (16 features, FMC:319-337): shares the play-model slot with play_model.xgb
    "run_fumble": 6,
}


@dataclass
class OneHotGroup:
    """One categorical input column expanded by OneHotEncoder(handle_unknown='ignore')."""
    name: str            # passer_name / target_name / rusher_name / coach
    base: int            # first column of the group
    categories: List[str]

    def column_of(self, value: Optional[str]) -> int:
        """Absolute column that is 1 for `value`, or -1 when it is not a category."""
        if value is None:
            return -1
        try:
            return self.base + self.categories.index(str(value))
        except ValueError:
            return -1


@dataclass
class Forest:
    name: str
    kind: int                       # KIND_XGB | KIND_SKL
    link: int
    n_outputs: int                  # classes (xgb multi) or quantiles (skl family)
    n_features: int                 # width of the model's input row
    num_base: int                   # column of NUM_FEATURES[0] (numerics are contiguous)
    n_num: int                      # how many numerics follow num_base
    zero_is_missing: bool           # CSR-fed boosters: exact zeros take the default branch
    base_margin: np.ndarray         # f64[n_outputs]  (xgb: margin offset; skl: init constant)
    scale: float                    # skl learning rate (leaf values are stored unscaled); 1 for xgb
    groups: List[OneHotGroup]
    # nodes (SoA, all trees concatenated; child indices are absolute)
    feat: np.ndarray                # i32[n_nodes]  (-1 for leaves)
    thr: np.ndarray                 # f32[n_nodes]
    left: np.ndarray                # i32[n_nodes]  (-1 for leaves)
    right: np.ndarray               # i32[n_nodes]
    default_left: np.ndarray        # u8[n_nodes]
    value: np.ndarray               # f64[n_nodes]  leaf value (xgb: f32-exact; skl: raw tree value)
    tree_root: np.ndarray           # i32[n_trees]
    tree_out: np.ndarray            # i32[n_trees]  output slot each tree adds into
    cover: Optional[np.ndarray] = None   # f64[n_nodes] hessian / weighted samples (analysis only)
    best_iteration: Optional[int] = None
    # play_model.xgb only: StandardScaler over these columns (index -> mean, scale)
    scaler_cols: Optional[np.ndarray] = None
    scaler_mean: Optional[np.ndarray] = None
    scaler_scale: Optional[np.ndarray] = None
    extra: Dict[str, object] = field(default_factory=dict)

    @property
    def n_trees(self) -> int:
        return int(self.tree_root.shape[0])

    @property
    def n_nodes(self) -> int:
        return int(self.feat.shape[0])

    def group(self, name: str) -> Optional[OneHotGroup]:
        for g in self.groups:
            if g.name == name:
                return g
        return None

    def rounds(self) -> int:
        """Boosting rounds (trees per output)."""
        return self.n_trees // max(1, self.n_outputs)


# ----------------------------------------------------------------------------------------------
# UBJSON (the subset XGBoost writes)
# ----------------------------------------------------------------------------------------------
_UBJ_INT = {b"i": ">b", b"U": ">B", b"I": ">h", b"l": ">i", b"L": ">q"}
_UBJ_NP = {b"i": ">i1", b"U": ">u1", b"I": ">i2", b"l": ">i4", b"L": ">i8", b"d": ">f4", b"D": ">f8"}


class _UBJReader:
    def __init__(self, buf: bytes):
        self.b = memoryview(bytes(buf))
        self.p = 0

    def _take(self, n: int) -> bytes:
        out = self.b[self.p:self.p + n].tobytes()
        if len(out) != n:
            raise ValueError("UBJSON: truncated input")
        self.p += n
        return out

    def _int(self, marker: bytes) -> int:
        fmt = _UBJ_INT[marker]
        return struct.unpack(fmt, self._take(struct.calcsize(fmt)))[0]

    def _length(self) -> int:
        return self._int(self._take(1))

    def _string(self) -> str:
        n = self._length()
        return self._take(n).decode("utf-8")

    def value(self, marker: Optional[bytes] = None):
        m = marker if marker is not None else self._take(1)
        while m == b"N":
            m = self._take(1)
        if m == b"Z":
            return None
        if m == b"T":
            return True
        if m == b"F":
            return False
        if m in _UBJ_INT:
            return self._int(m)
        if m == b"d":
            return struct.unpack(">f", self._take(4))[0]
        if m == b"D":
            return struct.unpack(">d", self._take(8))[0]
        if m == b"C":
            return self._take(1).decode("latin-1")
        if m in (b"S", b"H"):
            return self._string()
        if m == b"[":
            return self._array()
        if m == b"{":
            return self._object()
        raise ValueError(f"UBJSON: unknown marker {m!r} at {self.p}")

    def _container_header(self) -> Tuple[Optional[bytes], Optional[int]]:
        typ = None
        cnt = None
        if self.b[self.p:self.p + 1].tobytes() == b"$":
            self.p += 1
            typ = self._take(1)
        if self.b[self.p:self.p + 1].tobytes() == b"#":
            self.p += 1
            cnt = self._length()
        return typ, cnt

    def _array(self):
        typ, cnt = self._container_header()
        if typ is not None:
            if cnt is None:
                raise ValueError("UBJSON: typed array without count")
            if typ in _UBJ_NP:
                dt = np.dtype(_UBJ_NP[typ])
                raw = self._take(cnt * dt.itemsize)
                return np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("="))
            return [self.value(typ) for _ in range(cnt)]
        out = []
        if cnt is not None:
            for _ in range(cnt):
                out.append(self.value())
            return out
        while True:
            m = self._take(1)
            if m == b"]":
                return out
            out.append(self.value(m))

    def _object(self):
        typ, cnt = self._container_header()
        out = {}
        if cnt is not None:
            for _ in range(cnt):
                k = self._string()
                out[k] = self.value(typ)
            return out
        while True:
            if self.b[self.p:self.p + 1].tobytes() == b"}":
                self.p += 1
                return out
            k = self._string()
            out[k] = self.value(typ)


def parse_ubjson(buf: bytes):
    return _UBJReader(buf).value()


# ----------------------------------------------------------------------------------------------
# XGBoost boosters
# ----------------------------------------------------------------------------------------------
def _logit(p: float) -> float:
    return float(np.log(p / (1.0 - p)))


def _forest_from_xgb_learner(name: str, learner: dict, *, zero_is_missing: bool,
                             groups: List[OneHotGroup], num_base: int, n_num: int,
                             categorical_code: Optional[int] = None) -> Forest:
    """Re-lay one XGBoost `learner` dict (JSON or UBJSON, same keys) as a Forest.

    Semantics restated in SURVEY Appendix D.1: node i is a leaf iff left_children[i] == -1, leaf
    value = split_conditions[i]; margin = base_margin + sum of leaves, tree t adds to class
    tree_info[t]; base_margin = logit(base_score) for binary:logistic, base_score otherwise.
    """
    lmp = learner["learner_model_param"]
    objective = learner["objective"]["name"]
    n_class = max(1, int(lmp.get("num_class", "0")))
    n_features = int(lmp["num_feature"])
    base_score = float(np.float32(float(lmp["base_score"])))
    model = learner["gradient_booster"]["model"]
    trees = model["trees"]
    tree_info = np.asarray(model["tree_info"], dtype=np.int32)

    if objective == "binary:logistic":
        # xgboost keeps the margin in f32: ProbToMargin = -logf(1/p - 1)
        p = np.float32(base_score)
        base = float(-np.log(np.float32(1.0) / p - np.float32(1.0), dtype=np.float32))
        link = LINK_SIGMOID
    elif objective in ("multi:softprob", "multi:softmax"):
        base = base_score
        link = LINK_SOFTMAX
    else:
        base = base_score
        link = LINK_IDENTITY

    feat, thr, left, right, dl, val, cov, roots = [], [], [], [], [], [], [], []
    off = 0
    for t in trees:
        lc = np.asarray(t["left_children"], dtype=np.int64)
        rc = np.asarray(t["right_children"], dtype=np.int64)
        si = np.asarray(t["split_indices"], dtype=np.int64)
        sc = np.asarray(t["split_conditions"], dtype=np.float64)
        d = np.asarray(t["default_left"], dtype=np.uint8)
        st = np.asarray(t.get("split_type", np.zeros_like(lc)), dtype=np.int64)
        sh = np.asarray(t.get("sum_hessian", np.zeros_like(sc)), dtype=np.float64)
        n = lc.shape[0]
        leaf = lc < 0
        if np.any(st != 0):
            # Categorical split (XGBoost common/categorical.h `Decision`): a row goes RIGHT iff its category is
            # in the node's set.  Supported only where the row's category code is a known constant
            # (`categorical_code`): the node becomes a numeric test on a column that always holds 0.0,
            # threshold -1 (0 < -1 false -> right) when the code is in the set, +1 (-> left) otherwise.
            if categorical_code is None:
                raise NotImplementedError(f"{name}: categorical splits need a constant category code (SURVEY 8f row 3)")
            cats = np.asarray(t.get("categories", []), dtype=np.int64)
            cnodes = np.asarray(t.get("categories_nodes", []), dtype=np.int64)
            cseg = np.asarray(t.get("categories_segments", []), dtype=np.int64)
            csz = np.asarray(t.get("categories_sizes", []), dtype=np.int64)
            sc = sc.copy()
            for j, node in enumerate(cnodes):
                members = set(int(x) for x in cats[cseg[j]:cseg[j] + csz[j]])
                sc[node] = -1.0 if int(categorical_code) in members else 1.0
            if np.any((st != 0) & ~np.isin(np.arange(n), cnodes) & ~leaf):
                raise ValueError(f"{name}: categorical node without a category set")
        if not np.all(rc[~leaf] == lc[~leaf] + 1):
            raise ValueError(f"{name}: right child is not left+1")
        roots.append(off)
        feat.append(np.where(leaf, -1, si).astype(np.int32))
        thr.append(np.where(leaf, 0.0, sc).astype(np.float32))
        left.append(np.where(leaf, -1, lc + off).astype(np.int32))
        right.append(np.where(leaf, -1, rc + off).astype(np.int32))
        dl.append(np.where(leaf, 0, d).astype(np.uint8))
        val.append(np.where(leaf, sc.astype(np.float32).astype(np.float64), 0.0))
        cov.append(sh)
        off += n
    best_it = None
    attrs = learner.get("attributes", {}) or {}
    if "best_iteration" in attrs:
        best_it = int(attrs["best_iteration"])
    return Forest(
        name=name, kind=KIND_XGB, link=link, n_outputs=n_class, n_features=n_features,
        num_base=num_base, n_num=n_num, zero_is_missing=zero_is_missing,
        base_margin=np.full(n_class, base, dtype=np.float64), scale=1.0, groups=groups,
        feat=np.concatenate(feat), thr=np.concatenate(thr), left=np.concatenate(left),
        right=np.concatenate(right), default_left=np.concatenate(dl), value=np.concatenate(val),
        tree_root=np.asarray(roots, dtype=np.int32), tree_out=tree_info.copy(),
        cover=np.concatenate(cov), best_iteration=best_it,
    )


def load_xgb_json(path: str, name: str, groups: List[OneHotGroup], *, zero_is_missing: bool = True) -> Forest:
    with open(path, "r") as f:
        learner = json.load(f)["learner"]
    n_features = int(learner["learner_model_param"]["num_feature"])
    num_base = n_features - N_NUM
    return _forest_from_xgb_learner(name, learner, zero_is_missing=zero_is_missing,
                                    groups=groups, num_base=num_base, n_num=N_NUM)


class _XgbStub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.state = s


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("xgboost"):
            return type(name, (_XgbStub,), {})
        return super().find_class(module, name)


def load_play_model_xgb(path: str, scaler_path: Optional[str], coach_encoder_path: Optional[str]) -> Forest:
    """play_model.xgb = pickle(XGBClassifier) -> UBJSON booster."""
    with open(path, "rb") as f:
        obj = _StubUnpickler(f).load()
    handle = obj.state["_Booster"].state["handle"]
    doc = parse_ubjson(bytes(handle))
    learner = doc["Model"]["learner"] if "Model" in doc else doc["learner"]
    names = list(learner.get("feature_names") or [])
    coach_names = [n[len("coach_"):] for n in names[len(PLAY_FEATURES):]]
    if names and names[:len(PLAY_FEATURES)] != PLAY_FEATURES:
        raise ValueError("play_model.xgb: unexpected leading feature names")
    groups = [OneHotGroup("coach", len(PLAY_FEATURES), coach_names)]
    fo = _forest_from_xgb_learner("play_model", learner, zero_is_missing=False, groups=groups,
                                  num_base=0, n_num=len(PLAY_FEATURES))
    if scaler_path and os.path.exists(scaler_path):
        _install_sklearn_shims()
        import joblib
        sc = joblib.load(scaler_path)
        cols = [c for c in PLAY_FEATURES if c != "is_red_zone"]
        fo.scaler_cols = np.asarray([PLAY_FEATURES.index(c) for c in cols], dtype=np.int32)
        fo.scaler_mean = np.asarray(sc.mean_, dtype=np.float64)
        fo.scaler_scale = np.asarray(sc.scale_, dtype=np.float64)
        if fo.scaler_mean.shape[0] != len(cols):
            raise ValueError("scaler.pkl does not cover the 11 expected columns")
    if coach_encoder_path and os.path.exists(coach_encoder_path):
        _install_sklearn_shims()
        import joblib
        le = joblib.load(coach_encoder_path)
        if [str(c) for c in le.classes_] != coach_names:
            raise ValueError("coach_label_encoder.pkl does not match play_model.xgb coach columns")
    fo.extra["classes"] = list(PLAY_CLASSES)
    return fo


def load_play_model_json(path: str, features_path: str, label_encoder_path: str,
                         calibration_path: Optional[str] = None) -> Forest:
    """play_model.json: the binary PASS/RUN booster of train_run_pass.py (multi:softprob over the classes of
    label_encoder.pkl, 16 features of features.pkl incl. the categorical `head_coach`), loaded at FMC:319-337 and
    evaluated by `play_call_pass_prob_binary` (FMC:407-427).

    The forest is re-laid in the canonical numeric order (NUM_FEATURES, columns 0..16) with `head_coach` as column
    17.  The reference feeds `pd.Categorical([coach])` -- ONE category, so the category code the booster sees is 0
    for every team (FMC:310-313); numeric splits on that column therefore see 0.0 and categorical splits fold on
    code 0.  DataFrame-fed DMatrix: zeros are values, not missing."""
    _install_sklearn_shims()
    import joblib
    with open(path, "r") as f:
        learner = json.load(f)["learner"]
    names = [str(c) for c in joblib.load(features_path)]
    classes = [str(c) for c in joblib.load(label_encoder_path).classes_]
    temp = 1.0
    if calibration_path and os.path.exists(calibration_path):
        with open(calibration_path, "r") as f:
            temp = float(json.load(f).get("temperature", 1.0))
    return play_binary_from_learner(learner, names, classes, temp)


def play_binary_from_learner(learner: dict, names: List[str], classes: List[str], temperature: float = 1.0) -> Forest:
    """The `learner` dict of play_model.json + the contents of features.pkl / label_encoder.pkl -> Forest."""
    if int(learner["learner_model_param"]["num_feature"]) != len(names):
        raise ValueError("play_model.json: num_feature does not match features.pkl")
    if "pass" not in classes:
        raise ValueError("label_encoder.pkl has no 'pass' class")
    remap = np.zeros(len(names), dtype=np.int64)
    for c, nm in enumerate(names):
        if nm in NUM_FEATURES:
            remap[c] = NUM_FEATURES.index(nm)
        elif nm == "head_coach":
            remap[c] = N_NUM
        else:
            raise ValueError(f"play_model.json: unknown feature {nm!r}")
    fo = _forest_from_xgb_learner("play_binary", learner, zero_is_missing=False,
                                  groups=[OneHotGroup("head_coach", N_NUM, ["<category code 0>"])],
                                  num_base=0, n_num=N_NUM, categorical_code=0)
    internal = fo.left >= 0
    fo.feat = np.where(internal, remap[np.where(internal, fo.feat, 0)], -1).astype(np.int32)
    fo.n_features = N_NUM + 1
    if fo.n_outputs != len(classes):
        raise ValueError("play_model.json: num_class does not match label_encoder.pkl")
    fo.extra.update(classes=list(classes), pass_class=classes.index("pass"), temperature=float(temperature),
                    features=list(names))
    return fo


# ----------------------------------------------------------------------------------------------
# scikit-learn artifacts
# ----------------------------------------------------------------------------------------------
_SHIMS_DONE = False


def _install_sklearn_shims() -> None:
    """Make pickles written by scikit-learn 1.5.2 loadable under 1.9."""
    global _SHIMS_DONE
    if _SHIMS_DONE:
        return
    from collections import UserList
    import sklearn.compose._column_transformer as ct
    if not hasattr(ct, "_RemainderColsList"):
        class _RemainderColsList(UserList):
            def __init__(self, columns=None, *, future_dtype=None, warning_was_emitted=False,
                         warning_enabled=True):
                super().__init__(columns if columns is not None else [])
                self.future_dtype = future_dtype
                self.warning_was_emitted = warning_was_emitted
                self.warning_enabled = warning_enabled
        ct._RemainderColsList = _RemainderColsList
    if "_loss" not in sys.modules:
        import sklearn._loss._loss as real
        mod = types.ModuleType("_loss")
        mod.__dict__.update({k: v for k, v in real.__dict__.items() if not k.startswith("__")})

        def __pyx_unpickle_CyPinballLoss(cls, checksum, state):
            return real.CyPinballLoss(quantile=float(state[0]))
        mod.__pyx_unpickle_CyPinballLoss = __pyx_unpickle_CyPinballLoss
        sys.modules["_loss"] = mod
    _SHIMS_DONE = True


def load_sklearn_object(path: str):
    _install_sklearn_shims()
    import joblib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return joblib.load(path)


def _column_transformer_groups(ct) -> Tuple[List[OneHotGroup], int]:
    """One-hot layout of ColumnTransformer[('cat', OneHotEncoder), ('num','passthrough')].

    Output columns = categories of each cat column in order, then the numerics
    (train_pass_outcome_stage1.py:46-56).
    """
    enc = ct.named_transformers_["cat"]
    cat_cols = None
    for nm, _tr, cols in ct.transformers_:
        if nm == "cat":
            cat_cols = list(cols)
    groups = []
    base = 0
    for col, cats in zip(cat_cols, enc.categories_):
        cats = [str(c) for c in cats]
        groups.append(OneHotGroup(str(col), base, cats))
        base += len(cats)
    return groups, base


def f64_to_f32_floor(x: np.ndarray) -> np.ndarray:
    """Largest float32 <= x, so that `f32(v) <= x` is the same predicate as `v32 <= result`."""
    x = np.asarray(x, dtype=np.float64)
    y = x.astype(np.float32)
    too_big = y.astype(np.float64) > x
    y[too_big] = np.nextafter(y[too_big], np.float32(-np.inf))
    return y


def load_gbr_family(paths: Sequence[str], name: str) -> Forest:
    """Three quantile pipelines (q10,q50,q90) -> one Forest with n_outputs=3.

    sklearn semantics: pred = init + lr * sum(value[leaf]) accumulated in
    f64 stage by stage; X is cast to f32 and goes left iff x <= threshold.
    """
    feat, thr, left, right, val, cov, roots, outs, init = [], [], [], [], [], [], [], [], []
    groups = None
    num_base = None
    lr = None
    off = 0
    for q, p in enumerate(paths):
        pipe = load_sklearn_object(p)
        pre = pipe.named_steps["pre"]
        gb = pipe.named_steps["gb"]
        g, nb = _column_transformer_groups(pre)
        if groups is None:
            groups, num_base = g, nb
        elif [(x.name, x.base, x.categories) for x in g] != [(x.name, x.base, x.categories) for x in groups]:
            raise ValueError(f"{name}: quantile pipelines disagree on the one-hot layout")
        if lr is None:
            lr = float(gb.learning_rate)
        elif lr != float(gb.learning_rate):
            raise ValueError(f"{name}: learning rates differ")
        init.append(float(np.asarray(gb.init_.constant_).reshape(-1)[0]))
        for est in gb.estimators_[:, 0]:
            t = est.tree_
            n = t.node_count
            lc = t.children_left.astype(np.int64)
            rc = t.children_right.astype(np.int64)
            leaf = lc < 0
            roots.append(off)
            outs.append(q)
            feat.append(np.where(leaf, -1, t.feature).astype(np.int32))
            thr.append(np.where(leaf, 0.0, f64_to_f32_floor(t.threshold)).astype(np.float32))
            left.append(np.where(leaf, -1, lc + off).astype(np.int32))
            right.append(np.where(leaf, -1, rc + off).astype(np.int32))
            val.append(np.where(leaf, t.value.reshape(n), 0.0).astype(np.float64))
            cov.append(t.weighted_n_node_samples.astype(np.float64))
            off += n
    n_features = num_base + N_NUM
    return Forest(
        name=name, kind=KIND_SKL, link=LINK_IDENTITY, n_outputs=len(paths), n_features=n_features,
        num_base=num_base, n_num=N_NUM, zero_is_missing=False,
        base_margin=np.asarray(init, dtype=np.float64), scale=lr, groups=groups,
        feat=np.concatenate(feat), thr=np.concatenate(thr), left=np.concatenate(left),
        right=np.concatenate(right), default_left=np.zeros(off, dtype=np.uint8),
        value=np.concatenate(val), tree_root=np.asarray(roots, dtype=np.int32),
        tree_out=np.asarray(outs, dtype=np.int32), cover=np.concatenate(cov),
    )


def preprocessor_groups(path: str) -> List[OneHotGroup]:
    ct = load_sklearn_object(path)
    g, _ = _column_transformer_groups(ct)
    return g


# ----------------------------------------------------------------------------------------------
# Whole model directory
# ----------------------------------------------------------------------------------------------
@dataclass
class ModelSet:
    forests: Dict[str, Forest]
    source: str = ""

    def __getitem__(self, k: str) -> Forest:
        return self.forests[k]

    def __contains__(self, k: str) -> bool:
        return k in self.forests


def compile_reference_dir(d: str, *, stage2_json: Optional[str] = None) -> ModelSet:
    """Read every model the hot path uses from a directory laid out like the reference repo."""
    j = lambda *a: os.path.join(d, *a)
    forests: Dict[str, Forest] = {}
    forests["pass_stage1"] = load_xgb_json(
        j("pass_stage1_complete_vs_not.json"), "pass_stage1",
        preprocessor_groups(j("pass_stage1_preprocessor.joblib")))
    s2 = stage2_json or j("pass_stage2_notcomplete.json")
    if os.path.exists(s2):
        forests["pass_stage2"] = load_xgb_json(
            s2, "pass_stage2", preprocessor_groups(j("pass_stage2_preprocessor.joblib")))
    for fam, prefix in (("pass_yards", "pass_yards"), ("run_yards", "run_yards"), ("sack_yards", "sack_yards")):
        forests[fam] = load_gbr_family([j(f"{prefix}_q{q}.joblib") for q in (10, 50, 90)], fam)
    if os.path.exists(j("play_model.xgb")):
        forests["play_model"] = load_play_model_xgb(j("play_model.xgb"), j("scaler.pkl"),
                                                    j("coach_label_encoder.pkl"))
    if os.path.exists(j("play_model.json")) and os.path.exists(j("features.pkl")) and os.path.exists(j("label_encoder.pkl")):
        forests["play_binary"] = load_play_model_json(j("play_model.json"), j("features.pkl"), j("label_encoder.pkl"),
                                                      j("calibration.json"))
    if os.path.exists(j("run_fumble.json")):
        forests["run_fumble"] = load_xgb_json(
            j("run_fumble.json"), "run_fumble",
            preprocessor_groups(j("run_fumble_preprocessor.joblib")))
    for f in forests.values():
        check_forest(f)
    return ModelSet(forests, source=d)


def check_forest(f: Forest) -> None:
    """Structural self-checks (SURVEY 8c-iv / gate G1)."""
    n = f.n_nodes
    leaf = f.left < 0
    assert np.all((f.feat >= 0) == ~leaf), f"{f.name}: leaf/feature mismatch"
    assert np.all(f.feat[~leaf] < f.n_features), f"{f.name}: split index out of range"
    assert np.all(f.left[~leaf] < n) and np.all(f.right[~leaf] < n)
    assert f.tree_root.shape == f.tree_out.shape
    assert np.all(f.tree_out >= 0) and np.all(f.tree_out < f.n_outputs)
    if f.kind == KIND_XGB:
        assert np.all(f.right[~leaf] == f.left[~leaf] + 1), f"{f.name}: right != left+1"
    used = f.feat[~leaf]
    for g in f.groups:
        assert g.base + len(g.categories) <= f.n_features
    assert f.num_base + f.n_num <= f.n_features or f.name == "play_model"
    # every split feature is either inside a one-hot group or one of the numerics
    in_group = np.zeros_like(used, dtype=bool)
    for g in f.groups:
        in_group |= (used >= g.base) & (used < g.base + len(g.categories))
    in_num = (used >= f.num_base) & (used < f.num_base + f.n_num)
    assert np.all(in_group | in_num), f"{f.name}: split on a column outside the recipe"


# ----------------------------------------------------------------------------------------------
# Blob (single file, our own format) so that run time never needs the reference directory
# ----------------------------------------------------------------------------------------------
_BLOB_MAGIC = b"FMCF0001"
_ARRAYS = ("base_margin", "feat", "thr", "left", "right", "default_left", "value", "tree_root", "tree_out",
           "cover", "scaler_cols", "scaler_mean", "scaler_scale")


def save_modelset(ms: ModelSet, path: str) -> None:
    """npz container: per-forest arrays + a JSON header with the feature recipe."""
    arrays = {}
    header = {}
    for k, f in ms.forests.items():
        header[k] = dict(
            name=f.name, kind=f.kind, link=f.link, n_outputs=f.n_outputs, n_features=f.n_features,
            num_base=f.num_base, n_num=f.n_num, zero_is_missing=bool(f.zero_is_missing), scale=f.scale,
            best_iteration=f.best_iteration,
            groups=[dict(name=g.name, base=g.base, categories=g.categories) for g in f.groups],
            extra=f.extra,
        )
        for a in _ARRAYS:
            v = getattr(f, a)
            if v is not None:
                if a == "cover":
                    v = v.astype(np.float32)
                arrays[f"{k}/{a}"] = v
    arrays["__header__"] = np.frombuffer(json.dumps(header).encode("utf-8"), dtype=np.uint8)
    with open(path, "wb") as fh:
        np.savez_compressed(fh, **arrays)


def load_modelset(path: str) -> ModelSet:
    z = np.load(path, allow_pickle=False)
    header = json.loads(bytes(z["__header__"]).decode("utf-8"))
    forests = {}
    for k, h in header.items():
        kw = {}
        for a in _ARRAYS:
            key = f"{k}/{a}"
            kw[a] = z[key] if key in z.files else None
        if kw["cover"] is not None:
            kw["cover"] = kw["cover"].astype(np.float64)
        forests[k] = Forest(
            name=h["name"], kind=h["kind"], link=h["link"], n_outputs=h["n_outputs"],
            n_features=h["n_features"], num_base=h["num_base"], n_num=h["n_num"],
            zero_is_missing=h["zero_is_missing"], scale=h["scale"],
            groups=[OneHotGroup(g["name"], g["base"], g["categories"]) for g in h["groups"]],
            best_iteration=h.get("best_iteration"), extra=h.get("extra") or {}, **kw)
    return ModelSet(forests, source=path)


def default_model_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "models_2025.npz")


def load_default_models(reference_dir: Optional[str] = None) -> ModelSet:
    """Models for the shipped artifacts: an explicit reference-layout directory if given, else the
    compiled blob committed with this package."""
    if reference_dir:
        return compile_reference_dir(reference_dir)
    env = os.environ.get("FMC_MODEL_DIR")
    if env:
        return compile_reference_dir(env)
    p = default_model_path()
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} is missing; run `python -m fast_monte_carlo_b200.artifacts <reference_dir>` to compile it")
    return load_modelset(p)


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else default_model_path()
    os.makedirs(os.path.dirname(out), exist_ok=True)
    ms = compile_reference_dir(src)
    save_modelset(ms, out)
    for k, f in ms.forests.items():
        print(f"{k:12s} kind={f.kind} outs={f.n_outputs} trees={f.n_trees} nodes={f.n_nodes} "
              f"features={f.n_features} num_base={f.num_base}")
    print("wrote", out, os.path.getsize(out), "bytes")
