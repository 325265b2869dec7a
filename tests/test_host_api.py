"""Host-side logic of the reference-facing API, exercised on CPU by putting the CPU oracle into the
``KernelShap._explainer`` slot (tests may use the oracle; the product never does)."""
import json
import logging

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

from conftest import make_problem
from distributedkernelshap_b200.data import DenseData, convert_to_data, kmeans, sample
from distributedkernelshap_b200.explainers import distributed, kernel_shap
from distributedkernelshap_b200.explainers.interface import Explanation, NumpyEncoder
from distributedkernelshap_b200.explainers.kernel_shap import KernelShap, rank_by_importance, sum_categories
from distributedkernelshap_b200.explainers.utils import Bunch, batch, batch_slices, get_filename, load_data, load_model
from distributedkernelshap_b200.predictors import LinearModelSpec, LinearSoftmaxClassifier, extract_linear_spec
from oracle.shap_kernel_oracle import DenseData as OracleDenseData
from oracle.shap_kernel_oracle import KernelExplainerWrapperOracle


class OracleBackedWrapper(KernelExplainerWrapperOracle):
    """Stands in for the CUDA engine: same ctor shape, oracle arithmetic."""

    def __init__(self, predictor, data, device=None, **kwargs):
        if isinstance(data, DenseData):
            data = OracleDenseData(data.data, data.group_names, data.groups, data.weights)
        super().__init__(predictor, data, **kwargs)
        self.device = device


@pytest.fixture
def cpu_backend(monkeypatch):
    monkeypatch.setattr(kernel_shap, "KernelExplainerWrapper", OracleBackedWrapper)
    monkeypatch.setattr(distributed.parallel, "visible_gpus", lambda: 2)


def test_batch_split_rule():
    X = np.arange(2560 * 2).reshape(2560, 2)
    assert [len(b) for b in batch(X, batch_size=10)] == [10] * 256
    sizes = [len(b) for b in batch(X, batch_size=7)]
    assert sizes == [7] * 365 + [5] and sum(sizes) == 2560
    assert [len(b) for b in batch(X, n_batches=3)] == [854, 853, 853]
    for kw in (dict(batch_size=7), dict(n_batches=3), dict(batch_size=10), dict(n_batches=5000)):
        ref = np.array_split(X, [7 * i for i in range(1, 366)]) if kw.get("batch_size") == 7 else None
        got = batch(X, **kw)
        np.testing.assert_array_equal(np.concatenate(got), X)
        if ref is not None:
            assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    assert len(batch(sparse.csr_matrix(np.eye(6)), batch_size=4)) == 2
    assert batch_slices(10, None, 4) == [slice(0, 3), slice(3, 6), slice(6, 8), slice(8, 10)]
    assert get_filename(4, 10, serve=False) == "results/ray_workers_4_bsize_10_actorfr_1.0.pkl"
    assert get_filename(4, 10) == "results/ray_replicas_4_maxbatch_10_actorfr_1.0.pkl"
    b = Bunch(a=1)
    b.c = 2
    assert b.a == 1 and b["c"] == 2


def test_sum_categories_and_rank_by_importance():
    vals = np.arange(24, dtype=float).reshape(2, 12)
    out = sum_categories(vals, [2, 7], [3, 4])
    assert out.shape == (2, 7)
    np.testing.assert_array_equal(out[0], [0, 1, 2 + 3 + 4, 5, 6, 7 + 8 + 9 + 10, 11])
    out = sum_categories(vals, [0, 3], [3, 9])
    np.testing.assert_array_equal(out[0], [0 + 1 + 2, sum(range(3, 12))])
    inter = np.ones((2, 5, 5))
    assert sum_categories(inter, [1], [3]).shape == (2, 3, 3) and sum_categories(inter, [1], [3])[0, 1, 1] == 9
    with pytest.raises(ValueError):
        sum_categories(vals, [0], [3, 4])
    with pytest.raises(ValueError):
        sum_categories(vals, None, [3])
    with pytest.raises(ValueError):
        sum_categories(vals, [0], [13])
    with pytest.raises(ValueError):
        sum_categories(vals[0], [0], [3])
    sv = [np.array([[1.0, -3.0, 0.5], [1.0, 1.0, -0.5]]), np.array([[-1.0, 3.0, 2.5], [-1.0, -1.0, 0.5]])]
    imp = rank_by_importance(sv, ["a", "b", "c"])
    assert imp["0"]["names"] == ["b", "a", "c"] and imp["1"]["names"] == ["b", "c", "a"]
    np.testing.assert_allclose(imp["aggregated"]["ranked_effect"], [4.0, 2.0, 2.0])
    assert rank_by_importance(sv, ["a"])["0"]["names"][0] == "feature_1"   # wrong name count -> default names


def test_explanation_json_roundtrip():
    e = Explanation(meta={"name": "KernelShap", "params": {"k": np.int64(3)}},
                    data={"shap_values": [np.ones((2, 2))], "x": np.float32(1.5), "flag": np.bool_(True)})
    js = e.to_json()
    back = Explanation.from_json(js)
    assert back.shap_values == [[[1.0, 1.0], [1.0, 1.0]]] and back.meta["params"]["k"] == 3 and back.data["x"] == 1.5
    assert json.loads(json.dumps({"a": np.arange(3)}, cls=NumpyEncoder)) == {"a": [0, 1, 2]}
    with pytest.warns(DeprecationWarning):
        assert e["meta"]["name"] == "KernelShap"


def test_predictor_extraction():
    prob = make_problem()
    clf = prob["clf"]
    spec = extract_linear_spec(clf.predict_proba)
    assert spec.activation == "binary_logistic" and spec.kappa == 2.0 and spec.n_outputs == 2
    np.testing.assert_allclose(spec(prob["X"]), clf.predict_proba(prob["X"]))
    z = clf.decision_function(prob["X"])
    np.testing.assert_allclose(clf.predict_proba(prob["X"])[:, 1], 1 / (1 + np.exp(-2 * z)))   # softmax([-z, z])
    dspec = extract_linear_spec(clf.decision_function)
    assert dspec.activation == "identity" and dspec.scalar_out
    np.testing.assert_allclose(dspec(prob["X"]), z)
    with pytest.raises(TypeError):
        extract_linear_spec(lambda X: X)
    with pytest.raises(TypeError):
        extract_linear_spec(clf.predict)
    from sklearn.linear_model import LinearRegression, LogisticRegression
    Xs = np.random.default_rng(0).standard_normal((60, 4))
    ys = (Xs[:, 0] + 0.3 * Xs[:, 1] > 0).astype(int)
    sk = LogisticRegression().fit(Xs, ys)
    spec = extract_linear_spec(sk.predict_proba)
    np.testing.assert_allclose(spec(Xs), sk.predict_proba(Xs), rtol=1e-10)
    lin = LinearRegression().fit(Xs, Xs @ np.arange(4.0))
    np.testing.assert_allclose(extract_linear_spec(lin.predict)(Xs), lin.predict(Xs), rtol=1e-10)
    assert isinstance(extract_linear_spec(LinearModelSpec(np.ones((1, 3)), [0.0], "identity")), LinearModelSpec)
    assert LinearSoftmaxClassifier(np.ones((3, 4)), np.zeros(3)).predict_proba(Xs).shape == (60, 3)


def test_dense_data_and_summaries():
    d = DenseData(np.ones((5, 4)), ["a", "b"], [[0, 1], [2, 3]], [1, 1, 2, 2, 4])
    assert d.groups_size == 2 and abs(d.weights.sum() - 1) < 1e-15 and not d.transposed
    assert DenseData(np.ones((4, 5)), ["a", "b"], [[0, 1], [2, 3]]).transposed
    with pytest.raises(AssertionError):
        DenseData(np.ones((5, 4)), ["a"], [[0, 1, 2]])
    assert convert_to_data(pd.DataFrame(np.ones((3, 2)), columns=["u", "v"])).group_names == ["u", "v"]
    assert convert_to_data(np.ones(4)).data.shape == (1, 4)
    X = np.random.default_rng(0).standard_normal((50, 3))
    assert sample(X, 10).shape == (10, 3) and sample(X, 100) is X
    km = kmeans(X, 4)
    assert km.data.shape == (4, 3) and abs(km.weights.sum() - 1) < 1e-12
    assert all(km.data[i, j] in X[:, j] for i in range(4) for j in range(3))


def test_kernelshap_fit_explain_sequential(cpu_backend):
    """The reference's call sequence (benchmarks/ray_pool.py:34-37, :73) with sparse grouped background."""
    prob = make_problem(seed=31, n=6, N=10, widths=(1, 1, 3, 2, 1))
    ks = KernelShap(prob["clf"].predict_proba, link="logit", feature_names=prob["group_names"], seed=0)
    assert not ks.distribute and ks.meta["name"] == "KernelShap" and ks.meta["task"] == "classification"
    with pytest.raises(TypeError):
        ks.explain(prob["X"])                                  # not fitted
    ks.fit(sparse.csr_matrix(prob["bg"]), group_names=prob["group_names"], groups=prob["groups"])
    assert ks.use_groups and isinstance(ks.background_data, DenseData) and ks.meta["params"]["groups"] == prob["groups"]
    exp = ks.explain(sparse.csr_matrix(prob["X"]), silent=True, nsamples=100, l1_reg=False)
    assert len(exp.shap_values) == 2 and exp.shap_values[0].shape == (6, 5)
    assert exp.data["raw"]["instances"].shape == (6, 8) and list(exp.feature_names) == prob["group_names"]
    np.testing.assert_allclose(exp.shap_values[1].sum(1) + exp.expected_value[1], exp.raw["raw_prediction"][:, 1], atol=1e-9)
    assert set(exp.raw["importances"]) == {"0", "1", "aggregated"}
    assert exp.raw["prediction"].shape == (6,)
    json.loads(exp.to_json())


def test_kernelshap_input_checks_warn_and_disable(cpu_backend, caplog):
    prob = make_problem(seed=32, n=3, N=10, widths=(1, 1, 3, 2, 1))
    f = prob["clf"].predict_proba
    caplog.set_level(logging.WARNING)
    ks = KernelShap(f, link="logit").fit(prob["bg"], groups=prob["groups"])      # groups, no names -> auto names
    assert ks.use_groups and ks.feature_names == [f"group_{i}" for i in range(5)]
    ks = KernelShap(f, link="logit").fit(prob["bg"], group_names=["a", "b"], groups=prob["groups"])
    assert not ks.use_groups and "does not match the number of groups" in caplog.text
    ks = KernelShap(f, link="logit").fit(prob["bg"], group_names=["a"] * 3, groups=[[0, 1], [2, 3], [4]])
    assert not ks.use_groups and "did not match the number of features" in caplog.text
    ks = KernelShap(f, link="logit").fit(prob["bg"], group_names=prob["group_names"], groups=prob["groups"],
                                         weights=np.ones(4))
    assert ks.ignore_weights and ks.meta["params"]["weights"] is None
    ks = KernelShap(f, link="logit").fit(prob["bg"], group_names=prob["group_names"], groups=prob["groups"],
                                         weights=np.arange(1, 11.0))
    assert not ks.ignore_weights and abs(ks.background_data.weights[9] - 10 / 55) < 1e-15
    ks = KernelShap(f, link="logit").fit(pd.DataFrame(prob["bg"]), group_names=prob["group_names"], groups=prob["groups"])
    assert isinstance(ks.background_data, DenseData)

    class ArrayLike:                       # passes the shape checks, is not a supported container
        shape, ndim = (10, 8), 2
    with pytest.raises(TypeError):
        KernelShap(f).fit(ArrayLike())
    big = np.repeat(prob["bg"], 31, axis=0)
    caplog.clear()
    ks = KernelShap(f, link="logit", seed=0).fit(big, summarise_background=True, n_background_samples=20,
                                                 group_names=prob["group_names"], groups=prob["groups"])
    assert ks.summarise_background and ks.background_data.data.shape[0] == 20
    ks = KernelShap(f, link="logit", seed=0).fit(big, summarise_background="auto")
    assert ks.background_data.data.shape[0] == 300 and not ks.use_groups
    exp = ks.explain(prob["X"], silent=True, nsamples=60, l1_reg=False, summarise_result=True,
                     cat_vars_start_idx=[2, 5], cat_vars_enc_dim=[3, 2])
    assert exp.shap_values[0].shape == (3, 5) and exp.meta["params"]["summarise_result"] is True


def test_distributed_explainer_pool_matches_sequential(cpu_backend):
    prob = make_problem(seed=33, n=23, N=8, widths=(1, 1, 2, 1, 1, 1))
    kw = dict(link="logit", feature_names=prob["group_names"], seed=0)
    fitkw = dict(group_names=prob["group_names"], groups=prob["groups"])
    seq = KernelShap(prob["clf"].predict_proba, **kw).fit(prob["bg"], **fitkw)
    want = seq.explain(prob["X"], silent=True, nsamples=62, l1_reg=False).shap_values       # full enumeration: RNG-free
    for opts in ({"n_cpus": 1, "batch_size": 10}, {"n_cpus": 2, "batch_size": 5}, {"n_cpus": 2, "batch_size": None},
                 {"n_cpus": 4, "batch_size": 1, "actor_cpu_fraction": 0.5}):
        ks = KernelShap(prob["clf"].predict_proba, distributed_opts=opts, **kw).fit(prob["bg"], **fitkw)
        assert ks.distribute and isinstance(ks._explainer, distributed.DistributedExplainer)
        assert len(ks._explainer.pool) == min(int(opts["n_cpus"] // opts.get("actor_cpu_fraction", 1.0)), 2)
        assert ks._explainer.vector_out is True                   # __getattr__ proxy to worker 0
        got = ks.explain(prob["X"], silent=True, nsamples=62, l1_reg=False).shap_values
        for g, w in zip(got, want):
            np.testing.assert_allclose(g, w, atol=1e-12)
        with pytest.raises(TypeError):
            ks.explain(pd.DataFrame(prob["X"]))
    assert list(distributed.invert_permutation([2, 0, 1])) == [1, 2, 0]
    merged = distributed.kernel_shap_postprocess_fn([[np.ones((2, 3)), np.zeros((2, 3))], [np.ones((1, 3)), np.zeros((1, 3))]])
    assert merged[0].shape == (3, 3) and distributed.kernel_shap_postprocess_fn([np.ones((2, 3)), np.ones((1, 3))]).shape == (3, 3)


def test_distributed_explainer_sends_global_row_offsets(monkeypatch):
    """Every mini-batch travels with the index of its first row, so that plans drawn on the device per instance depend
    on the row and not on how the rows were split over workers."""
    seen = []

    class Recording(OracleBackedWrapper):
        def get_explanation(self, X, **kwargs):
            seen.append((int(kwargs.pop("row_offset")), len(X[1])))
            return super().get_explanation(X, **kwargs)

    monkeypatch.setattr(kernel_shap, "KernelExplainerWrapper", Recording)
    monkeypatch.setattr(distributed.parallel, "visible_gpus", lambda: 2)
    prob = make_problem(seed=34, n=23, N=8, widths=(1, 1, 2, 1, 1, 1))
    ks = KernelShap(prob["clf"].predict_proba, link="logit", seed=0, distributed_opts={"n_cpus": 2, "batch_size": 5})
    ks.fit(prob["bg"], group_names=prob["group_names"], groups=prob["groups"])
    ks.explain(prob["X"], silent=True, nsamples=62, l1_reg=False)
    assert sorted(seen) == [(0, 5), (5, 5), (10, 5), (15, 5), (20, 3)]


def test_serving_wrappers(cpu_backend):
    from distributedkernelshap_b200.explainers.wrappers import BatchKernelShapModel, KernelShapModel
    prob = make_problem(seed=34, n=5, N=8, widths=(1, 1, 2, 1))

    class Req:
        def __init__(self, arr):
            self.json = {"array": arr.tolist()}
    ckw = dict(link="logit", feature_names=prob["group_names"], seed=0)
    fkw = dict(group_names=prob["group_names"], groups=prob["groups"])
    single = KernelShapModel(prob["clf"], prob["bg"], ckw, fkw)
    one = json.loads(single(Req(prob["X"][0:1])))
    batched = BatchKernelShapModel(prob["clf"], prob["bg"], ckw, fkw)
    outs = [json.loads(s) for s in batched([Req(prob["X"][i:i + 1]) for i in range(5)])]
    assert len(outs) == 5
    np.testing.assert_allclose(outs[0]["data"]["shap_values"], one["data"]["shap_values"], atol=1e-12)
    assert np.asarray(outs[3]["data"]["raw"]["instances"]).shape == (1, 5)


def test_loaders_fall_back_to_synthetic_adult(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    data = load_data()
    clf = load_model("assets/predictor.pkl")
    assert data["all"]["X"]["processed"]["test"].shape == (2560, 49) and data["background"]["X"]["preprocessed"].shape == (100, 49)
    assert len(data["all"]["groups"]) == 12 and sum(len(g) for g in data["all"]["groups"]) == 49
    assert clf.predict_proba(data["all"]["X"]["processed"]["test"].toarray()[:3]).shape == (3, 2)


def test_device_summary_host_side_helpers():
    """category_segments / importances_from_device: the host half of the device-side build_explanation post-processing."""
    from distributedkernelshap_b200.explainers.kernel_shap import category_segments, importances_from_device
    np.testing.assert_array_equal(category_segments(12, [2, 7], [3, 4]), [0, 1, 2, 5, 6, 7, 11, 12])
    np.testing.assert_array_equal(category_segments(5, [0], [5]), [0, 5])
    vals = np.arange(24, dtype=float).reshape(2, 12)
    seg = category_segments(12, [2, 7], [3, 4])
    np.testing.assert_array_equal(np.add.reduceat(vals, seg[:-1], axis=1), sum_categories(vals, [2, 7], [3, 4]))
    sv = [np.array([[1.0, -3.0, 0.5], [1.0, 1.0, -0.5]]), np.array([[-1.0, 3.0, 2.5], [-1.0, -1.0, 0.5]])]
    per = [np.abs(v).mean(0) for v in sv]
    mean_abs = np.stack(per + [per[0] + per[1]])
    order = np.stack([np.argsort(m, kind="stable")[::-1] for m in mean_abs]).astype(np.int32)
    got = importances_from_device({"mean_abs": mean_abs, "order": order}, ["a", "b", "c"])
    want = rank_by_importance(sv, ["a", "b", "c"])
    for key in want:
        assert got[key]["names"] == want[key]["names"]
        np.testing.assert_allclose(got[key]["ranked_effect"], want[key]["ranked_effect"])
    assert importances_from_device({"mean_abs": mean_abs, "order": order}, ["a"])["0"]["names"][0].startswith("feature_")


def test_bench_sizes_the_reference_pool_by_the_cpu_quota():
    import bench
    n = bench.usable_cores()
    assert 1 <= n <= (bench.os.cpu_count() or 1)


def test_partial_varying_sets_are_refused_beyond_64_groups():
    """Multi-word coalition rows run on the shared-plan path only (every group must vary): the engine refuses the batch
    before building plans for sizes no kernel evaluates; up to 64 groups the general kernels take partial sets."""
    from distributedkernelshap_b200.engine import refuse_partial_sets_beyond_64_groups as refuse
    hist = np.zeros(13, dtype=np.int32)
    hist[[3, 12]] = 4, 9
    refuse(12, hist)
    hist = np.zeros(201, dtype=np.int32)
    hist[200] = 7
    refuse(200, hist)
    hist[[0, 198]] = 1, 2
    with pytest.raises(NotImplementedError, match="3 instance.*more than 64 groups.*unsupported"):
        refuse(200, hist)
