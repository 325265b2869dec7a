"""GPU tests of what sits around the hot path at the API level: build_explanation's post-processing on the device
(reference kernel_shap.py:36-109, :112-207, :952-956) and the serving back-ends (reference explainers/wrappers.py:12-88)."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _adult(n=48):
    from distributedkernelshap_b200.datasets import adult_like
    return adult_like(n_explain=n)


def test_device_summary_matches_numpy_ranking_and_argmax():
    from distributedkernelshap_b200.explainers.kernel_shap import KernelShap, rank_by_importance
    d = _adult()
    ks = KernelShap(d["predictor"].predict_proba, link="logit", feature_names=d["group_names"], seed=0)
    ks.fit(d["data"]["background"]["X"]["preprocessed"], group_names=d["group_names"], groups=d["groups"])
    exp = ks.explain(d["X_explain"], silent=True, nsamples=2048, l1_reg=False)
    summary = ks._explainer.summarise(48)
    assert summary is not None                                   # the device path was taken
    want = rank_by_importance(exp.shap_values, feature_names=d["group_names"])
    got = exp.raw["importances"]
    assert set(got) == {"0", "1", "aggregated"}
    for key in want:
        np.testing.assert_allclose(got[key]["ranked_effect"], want[key]["ranked_effect"], rtol=1e-9, atol=1e-11)
        assert got[key]["names"] == want[key]["names"]
    p = d["predictor"].predict_proba(d["X_explain"])
    np.testing.assert_array_equal(exp.raw["prediction"], p.argmax(1))
    # another call in between invalidates the resident result: the host path takes over, same numbers
    ks._explainer.varying(d["X_explain"][:3])
    assert ks._explainer.summarise(48) is None


def test_device_sum_categories():
    """Ungrouped columns + summarise_result: the one-hot blocks are added up on the device like ``sum_categories``."""
    from distributedkernelshap_b200.datasets import ADULT_ONEHOT_WIDTHS
    from distributedkernelshap_b200.explainers.kernel_shap import KernelShap, sum_categories
    d = _adult(20)
    starts, start = [], 4
    for wd in ADULT_ONEHOT_WIDTHS:
        starts.append(start)
        start += wd
    ks = KernelShap(d["predictor"].predict_proba, link="logit", seed=0)
    ks.fit(d["background"])
    exp = ks.explain(d["X_explain"], silent=True, nsamples=500, l1_reg=False, summarise_result=True,
                     cat_vars_start_idx=starts, cat_vars_enc_dim=list(ADULT_ONEHOT_WIDTHS))
    raw = ks._explainer.shap_values(d["X_explain"], nsamples=500, l1_reg=False)
    for c in range(2):
        want = sum_categories(raw[c], starts, list(ADULT_ONEHOT_WIDTHS))
        assert exp.shap_values[c].shape == (20, 12)
        np.testing.assert_allclose(exp.shap_values[c], want, rtol=0, atol=1e-12)
    agg = exp.raw["importances"]["aggregated"]["ranked_effect"]
    assert np.all(np.diff(agg) <= 0)


def test_serving_backends_on_the_gpu():
    """KernelShapModel answers one request, BatchKernelShapModel a list of them with ONE engine call: same explanations."""
    from distributedkernelshap_b200.explainers.wrappers import BatchKernelShapModel, KernelShapModel
    d = _adult(9)

    class Req:
        def __init__(self, arr):
            self.json = {"array": arr.tolist()}
    ckw = dict(link="logit", feature_names=d["group_names"], seed=0)
    fkw = dict(group_names=d["group_names"], groups=d["groups"])
    single = KernelShapModel(d["predictor"], d["background"], ckw, fkw)
    batched = BatchKernelShapModel(d["predictor"], d["background"], ckw, fkw)
    launches0 = batched.explainer._explainer.kernel_launches()
    outs = [json.loads(s) for s in batched([Req(d["X_explain"][i:i + 1]) for i in range(9)])]
    per_batch = batched.explainer._explainer.kernel_launches() - launches0
    assert len(outs) == 9
    launches1 = single.explainer._explainer.kernel_launches()
    ones = [json.loads(single(Req(d["X_explain"][i:i + 1]))) for i in range(9)]
    per_single = (single.explainer._explainer.kernel_launches() - launches1) / 9
    for a, b in zip(outs, ones):
        np.testing.assert_allclose(a["data"]["shap_values"], b["data"]["shap_values"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(a["data"]["raw"]["raw_prediction"], b["data"]["raw"]["raw_prediction"], rtol=1e-12)
    assert per_batch < 3 * per_single                             # the batch was coalesced, not looped over
    fx = d["predictor"].predict_proba(d["X_explain"])
    sv = np.asarray([o["data"]["shap_values"][1][0] for o in outs])
    ev = outs[0]["data"]["expected_value"][1]
    np.testing.assert_allclose(sv.sum(1) + ev, np.log(fx[:, 1] / fx[:, 0]), rtol=1e-7, atol=1e-7)


def test_pool_of_gpus_matches_one_gpu_and_is_repeatable():
    """Single process, several GPUs (the ActorPool pattern without ray): mini-batches handed to whichever worker is free.
    With a seed every worker evaluates the same keyed shared plans, so the pooled result equals the sequential one and does
    not depend on thread scheduling (sampled M: nsamples 2048 of 4094).  Needs two visible GPUs."""
    from distributedkernelshap_b200 import parallel
    from distributedkernelshap_b200.explainers.kernel_shap import KernelShap
    if parallel.visible_gpus() < 2:
        pytest.skip("needs two GPUs")
    d = _adult(300)
    args = dict(link="logit", feature_names=d["group_names"], seed=5)
    fkw = dict(group_names=d["group_names"], groups=d["groups"])
    one = KernelShap(d["predictor"].predict_proba, **args)
    one.fit(d["background"], **fkw)
    want = one.explain(d["X_explain"], silent=True, nsamples=2048, l1_reg=False).shap_values
    for trial in range(2):
        pool = KernelShap(d["predictor"].predict_proba, distributed_opts={"n_cpus": 2, "batch_size": 37}, **args)
        pool.fit(d["background"], **fkw)
        assert len(pool._explainer.pool) == 2
        got = pool.explain(d["X_explain"], silent=True, nsamples=2048, l1_reg=False).shap_values
        np.testing.assert_array_equal(got[1], want[1])
    few = KernelShap(d["predictor"].predict_proba, distributed_opts={"n_cpus": 2, "batch_size": None}, **args)
    few.fit(d["background"], **fkw)
    np.testing.assert_array_equal(few.explain(d["X_explain"][:1], silent=True, nsamples=2048, l1_reg=False).shap_values[1],
                                  want[1][:1])                      # fewer rows than workers: no empty mini-batch


def test_zero_rows():
    from distributedkernelshap_b200.explainers.kernel_shap import KernelExplainerWrapper
    d = _adult(4)
    from distributedkernelshap_b200.data import DenseData
    eng = KernelExplainerWrapper(d["predictor"].predict_proba, DenseData(d["background"], d["group_names"], d["groups"]),
                                 link="logit", seed=0)
    sv = eng.shap_values(d["X_explain"][:0], nsamples=2048, l1_reg=False)
    assert len(sv) == 2 and sv[0].shape == (0, 12)
