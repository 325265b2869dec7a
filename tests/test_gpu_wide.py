"""More than 128 groups (sixteen 64-bit words per coalition row): the shared-plan coalition kernel instantiated for
sixteen-word rows + the float64 projection solve of csrc/dks_wide.cuh, against the CPU oracle fed the same plan.
Tolerance as everywhere: 1e-5 relative to the largest |phi| of the instance; additivity to 1e-8.  The full configs[3]
singleton shape (1024 groups, 256 background rows, 8192 coalitions) is checked against a committed oracle fixture in
test_gpu_baseline_shapes.py."""
import numpy as np
import pytest

from conftest import make_problem as _make_problem, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


def make_problem(**kw):
    """conftest.make_problem with the coefficients scaled for hundreds of columns (scores of unit order: with the default
    N(0, 0.7) per column a 160-column score saturates the logistic head to exactly 1.0 in float64 and the logit link of the
    oracle divides by zero)."""
    from distributedkernelshap_b200.predictors import LinearSoftmaxClassifier
    prob = _make_problem(**kw)
    D = prob["bg"].shape[1]
    clf = prob["clf"]
    prob["clf"] = LinearSoftmaxClassifier(clf.coef_ * (2.0 / np.sqrt(D)), clf.intercept_, multi_class="multinomial")
    return prob


def _oracle(prob, link):
    from oracle.shap_kernel_oracle import DenseData, KernelExplainerOracle
    return KernelExplainerOracle(prob["clf"].predict_proba, DenseData(prob["bg"], prob["group_names"], prob["groups"]),
                                 link=link)


def _engine(prob, link, **kw):
    from distributedkernelshap_b200.data import DenseData
    from distributedkernelshap_b200.engine import GpuKernelExplainer
    return GpuKernelExplainer(prob["clf"].predict_proba, DenseData(prob["bg"], prob["group_names"], prob["groups"]),
                              link=link, **kw)


# (groups, background rows, coalitions): 160 -> 40 nibble tables (two full words + half a word), 200 -> 50 tables;
# 40 / 130 / 256 background rows -> one chunk with an 8-column tail, two chunks (128 + 2), two full chunks; 129 groups is
# the first shape past the two-word rows, with a single nibble in its third word
SHAPES = [(160, 40, 1200), (200, 130, 900), (129, 256, 700), (320, 16, 2000)]


@pytest.mark.parametrize("G,N,S", SHAPES)
@pytest.mark.parametrize("link", ["logit", "identity"])
def test_wide_shared_plan_matches_oracle(G, N, S, link):
    prob = make_problem(seed=G + N, n=9, N=N, widths=(1,) * G)
    eng = _engine(prob, link, seed=5)
    got = eng.shap_values(prob["X"], nsamples=S, l1_reg=False)
    M, _ = eng.varying(prob["X"])
    assert (M == G).all()
    plan = eng.shared_plan(G, S)
    assert plan.zbits.shape == (S, 16)
    orc = _oracle(prob, link)
    want = np.stack([orc.explain(prob["X"][i:i + 1], plan=(plan.dense(), plan.weights), nsamples=S, l1_reg=False)
                     for i in range(len(prob["X"]))])                      # [n, G, 2]
    for c in range(2):
        assert got[c].shape == (len(prob["X"]), G)
        err = rel_err(got[c], want[:, :, c])
        assert err < TOL, (G, N, S, link, c, err)
    np.testing.assert_allclose(got[0], -got[1], rtol=0, atol=1e-12)
    # additivity: sum_k phi_k = link(f(x)) - link(E f)
    fx = prob["clf"].predict_proba(prob["X"])
    lf = np.log(fx / (1 - fx)) if link == "logit" else fx
    for c in range(2):
        np.testing.assert_allclose(got[c].sum(axis=1), lf[:, c] - eng.expected_value[c], rtol=1e-8, atol=1e-7)


def test_wide_grouped_columns_and_repeat_calls():
    """Groups of several columns each (150 groups over 330 columns), a second call on other rows (plan and projection are
    reused), and a call through the device-resident entry point (the CUDA-graph path replays the same launches)."""
    widths = tuple(1 + (k % 3 == 0) + (k % 7 == 0) * 2 for k in range(150))
    prob = make_problem(seed=77, n=24, N=33, widths=widths)
    eng = _engine(prob, "logit", seed=2)
    S = 1000
    first = eng.shap_values(prob["X"][:10], nsamples=S, l1_reg=False)
    second = eng.shap_values(prob["X"][10:], nsamples=S, l1_reg=False)
    plan = eng.shared_plan(150, S)
    orc = _oracle(prob, "logit")
    for rows, got in ((range(0, 10), first), (range(10, 24), second)):
        want = np.stack([orc.explain(prob["X"][i:i + 1], plan=(plan.dense(), plan.weights), nsamples=S, l1_reg=False)
                         for i in rows])
        for c in range(2):
            assert rel_err(got[c], want[:, :, c]) < TOL
    again = eng.shap_values(prob["X"][:10], nsamples=S, l1_reg=False)
    for c in range(2):
        np.testing.assert_array_equal(again[c], first[c])                 # fixed summation order: bit-reproducible

    import torch
    dev = torch.device("cuda", eng.device)
    X = torch.from_numpy(prob["X"]).to(dev)
    phi = torch.empty((2, 24, 150), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(3):                                                 # the third call replays the captured graph
            eng.explain_device(X.data_ptr(), 24, phi.data_ptr(), nsamples=S)
            eng.check_status()
    host = phi.cpu().numpy()
    for c in range(2):
        np.testing.assert_array_equal(host[c][:10], first[c])
        np.testing.assert_array_equal(host[c][10:], second[c])


def test_wide_refusals():
    """What the sixteen-word path does not cover raises; nothing is solved another way."""
    prob = make_problem(seed=8, n=6, N=20, widths=(1,) * 140)
    eng = _engine(prob, "logit", seed=1)
    with pytest.raises(NotImplementedError, match="128 groups"):
        eng.shap_values(prob["X"], nsamples=600)                           # default l1_reg='auto' would select features
    with pytest.raises(NotImplementedError, match="128 groups"):
        eng.shap_values(prob["X"], nsamples=600, l1_reg="num_features(10)")
    # an instance whose varying set is partial has no kernel beyond 64 groups: reported, not computed
    prob2 = make_problem(seed=8, n=6, N=20, widths=(1,) * 140, constant_groups=(3,))
    eng2 = _engine(prob2, "logit", seed=1)
    from distributedkernelshap_b200._cabi import DksError
    with pytest.raises((DksError, NotImplementedError, RuntimeError)):
        eng2.shap_values(prob2["X"], nsamples=600, l1_reg=False)
    # per-instance plans drawn on the device stop at 64 groups
    eng3 = _engine(prob, "logit", seed=1, plan_mode="per_instance")
    with pytest.raises((DksError, NotImplementedError, RuntimeError)):
        eng3.shap_values(prob["X"], nsamples=600, l1_reg=False)


def test_kernel_shap_api_with_200_ungrouped_features():
    """KernelShap.fit / explain (kernel_shap.py:581-621: an ungrouped array, one group per column) on 200 columns: shap
    values, expected value and the importance ranking computed on the device."""
    from distributedkernelshap_b200.explainers.kernel_shap import KernelShap
    prob = make_problem(seed=21, n=12, N=50, widths=(1,) * 200)
    ks = KernelShap(prob["clf"].predict_proba, link="logit", feature_names=prob["group_names"], seed=0)
    ks.fit(prob["bg"])
    exp = ks.explain(prob["X"], nsamples=1500, l1_reg=False, silent=True)
    sv = exp.shap_values
    eng = ks._explainer
    plan = eng.shared_plan(200, 1500)
    orc = _oracle(prob, "logit")
    want = np.stack([orc.explain(prob["X"][i:i + 1], plan=(plan.dense(), plan.weights), nsamples=1500, l1_reg=False)
                     for i in range(4)])
    for c in range(2):
        assert rel_err(sv[c][:4], want[:, :, c]) < TOL
    ranked = exp.raw["importances"]["aggregated"]["ranked_effect"]
    agg = np.abs(sv[0]).mean(axis=0) + np.abs(sv[1]).mean(axis=0)
    np.testing.assert_allclose(np.asarray(ranked), np.sort(agg)[::-1], rtol=1e-9)
