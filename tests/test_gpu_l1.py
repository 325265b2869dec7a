"""l1 feature selection on the device (csrc/dks_l1.cuh) against the oracle's restatement of upstream ``solve`` with
scikit-learn 0.23.2 semantics (oracle/sklearn_lars_restated.py): the engine draws its shared plan, the oracle is fed the
same plan and the same ``l1_reg``; the selected features (the non-zero pattern) must be identical and phi within 1e-5."""
import numpy as np
import pytest

from conftest import make_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _pair(prob, seed):
    from distributedkernelshap_b200.data import DenseData
    from distributedkernelshap_b200.engine import GpuKernelExplainer
    from oracle.shap_kernel_oracle import DenseData as ODenseData, KernelExplainerOracle
    eng = GpuKernelExplainer(prob["clf"].predict_proba, DenseData(prob["bg"], prob["group_names"], prob["groups"]),
                             link="logit", seed=seed)
    orc = KernelExplainerOracle(prob["clf"].predict_proba, ODenseData(prob["bg"], prob["group_names"], prob["groups"]),
                                link="logit")
    return eng, orc


def _check(eng, orc, X, nsamples, l1_reg, G):
    got = eng.shap_values(X, nsamples=nsamples, l1_reg=l1_reg)
    plan = eng.shared_plan(G, nsamples)
    selected = []
    for i in range(X.shape[0]):
        want = orc.explain(X[i:i + 1], plan=(plan.dense(), plan.weights), nsamples=nsamples, l1_reg=l1_reg)
        np.testing.assert_array_equal(got[1][i] != 0, want[:, 1] != 0, err_msg=f"instance {i}: other features selected")
        for c in range(2):
            assert rel_err(got[c][i], want[:, c]) < TOL, (l1_reg, i, c)
        selected.append(int(np.count_nonzero(want[:, 1])))
    return selected, got


@pytest.mark.parametrize("l1_reg", ["auto", "aic", "bic", "num_features(5)", "num_features(1)"])
def test_l1_selection_small_problem(l1_reg):
    prob = make_problem(seed=70, n=12, N=20, widths=(1,) * 16)
    eng, orc = _pair(prob, seed=3)
    selected, got = _check(eng, orc, prob["X"], 300, l1_reg, 16)
    if l1_reg.startswith("num_features"):
        assert set(selected) == {int(l1_reg[13:-1])}
    else:
        assert min(selected) < 16                          # a real selection took place somewhere
    fx = prob["clf"].predict_proba(prob["X"])
    np.testing.assert_allclose(got[1].sum(1), np.log(fx[:, 1] / fx[:, 0]) - eng.expected_value[1], rtol=1e-8, atol=1e-8)
    plain = eng.shap_values(prob["X"], nsamples=300, l1_reg=False)          # and the mode switches back
    assert np.count_nonzero(plain[1]) == plain[1].size


def test_l1_auto_only_triggers_under_20_percent():
    """nsamples = 2048 of 4094 (Adult): 'auto' does not select; the result equals l1_reg=False bit for bit."""
    prob = make_problem(seed=71, n=6, N=10, widths=(1,) * 12)
    eng, _ = _pair(prob, seed=5)
    a = eng.shap_values(prob["X"], nsamples=2048, l1_reg="auto")
    b = eng.shap_values(prob["X"], nsamples=2048, l1_reg=False)
    np.testing.assert_array_equal(a[1], b[1])


def test_l1_selection_config2_shape_and_two_word_rows():
    """64 ungrouped features (BASELINE configs[2]: the reference default l1_reg='auto' selects features there) at a reduced
    background, and 80 features (two-word coalition rows)."""
    from distributedkernelshap_b200.datasets import dense_tabular
    from distributedkernelshap_b200.engine import GpuKernelExplainer
    from oracle.shap_kernel_oracle import KernelExplainerOracle
    for G, N, ns, n in [(64, 100, 1200, 5), (80, 40, 700, 3)]:
        d = dense_tabular(n=n, n_features=G, n_background=N, seed=G)
        eng = GpuKernelExplainer(d["predictor"].predict_proba, d["background"], link="logit", seed=2)
        orc = KernelExplainerOracle(d["predictor"].predict_proba, d["background"], link="logit")
        selected, _ = _check(eng, orc, d["X_explain"], ns, "auto", G)
        assert max(selected) < G
        _check(eng, orc, d["X_explain"], ns, "num_features(10)", G)


def test_l1_through_the_kernelshap_api_default_kwargs():
    """The reference benchmark passes only silent=True (ray_pool.py:73): l1_reg='auto' is what runs."""
    from distributedkernelshap_b200.datasets import dense_tabular
    from distributedkernelshap_b200.explainers.kernel_shap import KernelShap
    from oracle.shap_kernel_oracle import KernelExplainerOracle
    d = dense_tabular(n=4, n_features=20, n_background=30, seed=9)
    ks = KernelShap(d["predictor"].predict_proba, link="logit", seed=4)
    ks.fit(d["background"])
    exp = ks.explain(d["X_explain"], silent=True)                       # nsamples 'auto' = 2088 of 2^20 - 2: selection
    eng = ks._explainer
    plan = eng.shared_plan(20, "auto")
    orc = KernelExplainerOracle(d["predictor"].predict_proba, d["background"], link="logit")
    for i in range(4):
        want = orc.explain(d["X_explain"][i:i + 1], plan=(plan.dense(), plan.weights))
        np.testing.assert_array_equal(exp.shap_values[1][i] != 0, want[:, 1] != 0)
        assert rel_err(exp.shap_values[1][i], want[:, 1]) < TOL
