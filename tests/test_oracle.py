"""Pins the CPU oracle (NumPy restatement of shap 0.35.0 KernelExplainer) with analytic known answers.
The reference ships no tests or golden vectors, so these replace them."""
import numpy as np
import pytest

from conftest import make_problem
from oracle.shap_kernel_oracle import (DenseData, KernelExplainerOracle, KernelExplainerWrapperOracle, build_plan,
                                       effective_nsamples, exact_shapley, shapley_size_weights)


def _oracle(prob, link="logit", predict="predict_proba", **kw):
    args = (prob["groups"],) + ((prob["weights"],) if prob["weights"] is not None else ())
    return KernelExplainerOracle(getattr(prob["clf"], predict), DenseData(prob["bg"], prob["group_names"], *args),
                                 link=link, **kw)


@pytest.mark.parametrize("link", ["logit", "identity"])
def test_full_enumeration_equals_exact_shapley_values(link):
    """With every coalition enumerated, KernelSHAP's constrained WLS solution IS the Shapley value of the set function
    v(T) = link(E_bg[f(x_T, bg_~T)]) - link(fnull).  Brute-force subset formula vs oracle."""
    prob = make_problem(seed=21, n=3, N=7, widths=(1, 2, 1, 1, 3, 1), weights=True)
    orc = _oracle(prob, link)
    groups = prob["groups"]
    wb = prob["weights"] / prob["weights"].sum()
    f = prob["clf"].predict_proba
    lf = orc.link.f
    for i in range(3):
        x = prob["X"][i]

        def value(mask):
            rows = prob["bg"].copy()
            for k, on in enumerate(mask):
                if on:
                    rows[:, groups[k]] = x[groups[k]]
            ey = (f(rows) * wb[:, None]).sum(0)
            return lf(ey) - lf(orc.fnull)
        want = exact_shapley(value, len(groups))
        got = orc.explain(x[None], nsamples=10 ** 6, l1_reg=False)
        np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)


def test_affine_model_closed_form():
    prob = make_problem(seed=22, n=5, N=9, widths=(1, 2, 1, 3, 1, 1, 2), weights=True)
    orc = _oracle(prob, "identity", "decision_function")
    assert not orc.vector_out and isinstance(orc.expected_value, float)
    np.random.seed(0)
    got = orc.shap_values(prob["X"], nsamples=60, l1_reg=False)
    wb = prob["weights"] / prob["weights"].sum()
    coef = prob["clf"].coef_[0]
    for g, cols in enumerate(prob["groups"]):
        closed = ((prob["X"][:, cols] - (wb[:, None] * prob["bg"][:, cols]).sum(0)) * coef[cols]).sum(1)
        np.testing.assert_allclose(got[:, g], closed, rtol=1e-9, atol=1e-12)


def test_additivity_antisymmetry_and_expected_value():
    prob = make_problem(seed=23, n=6, N=11, widths=(1,) * 5 + (2, 3))
    orc = _oracle(prob)
    np.random.seed(1)
    sv = orc.shap_values(prob["X"], nsamples=80, l1_reg=False)
    fx = prob["clf"].predict_proba(prob["X"])
    for c in range(2):
        np.testing.assert_allclose(sv[c].sum(1), np.log(fx[:, c] / (1 - fx[:, c])) - orc.expected_value[c], atol=1e-10)
    np.testing.assert_allclose(sv[0], -sv[1], atol=1e-10)
    fnull = prob["clf"].predict_proba(prob["bg"]).mean(0)
    np.testing.assert_allclose(orc.expected_value, np.log(fnull / (1 - fnull)), rtol=1e-12)


def test_plan_invariants_adult():
    """Numbers computed from the upstream rule for M = 12."""
    np.random.seed(0)
    Z, w, info = build_plan(12, 2048)
    assert info["nfixed"] == 24 + 132 + 440 == 596 and info["num_full_subsets"] == 3
    assert Z.shape == (2048, 12) and abs(w.sum() - 1) < 1e-12
    assert abs(info["weight_left"] - 0.29290) < 1e-5 and abs(w[596:].sum() - info["weight_left"]) < 1e-12
    sizes = Z.sum(1)
    assert set(sizes[:24]) == {1, 11} and set(sizes[24:156]) == {2, 10} and set(sizes[156:596]) == {3, 9}
    assert set(sizes[596:]) <= {4, 5, 6, 7, 8}
    np.testing.assert_array_equal(Z[0:596:2] + Z[1:596:2], 1)          # complements adjacent
    assert len({tuple(r) for r in Z}) == 2048                          # sampled rows are de-duplicated
    np.random.seed(0)
    Z2, _, info2 = build_plan(12, effective_nsamples(12, "auto")[0])
    assert Z2.shape[0] == 2072 and info2["nfixed"] == 596
    assert effective_nsamples(12, 10 ** 6) == (4094, 4094) and effective_nsamples(40, "auto") == (2128, 2 ** 30)
    wv, nss, nps = shapley_size_weights(12)
    assert (nss, nps) == (6, 5) and abs(wv.sum() - 1) < 1e-15


@pytest.mark.parametrize("M,S,R", [(12, 300, 250), (9, 300, 300)])
def test_sampled_estimate_is_centred_on_the_exact_shapley_values(M, S, R):
    """Statistical pin of what the analytic tests cannot reach -- the SAMPLED part of the plan (subset-size
    distribution, duplicate folding, complement rows, the rescaling of the sampled weights to the mass the enumerated
    sizes left): over R independent seeds the mean of the sampled KernelSHAP estimate must agree with the exact Shapley
    values (brute-force subset formula) within the CLT bound.  Wrong rescaling moves the mean by > 8 standard errors,
    dropping the duplicate counts or the halving of paired sizes by 3-4 in the M = 9 regime where most draws repeat."""
    prob = make_problem(seed=41, n=1, N=8, widths=(1,) * M)
    orc = _oracle(prob)
    x = prob["X"][0]

    def value(mask):
        rows = prob["bg"].copy()
        rows[:, mask.astype(bool)] = x[mask.astype(bool)]
        ey = prob["clf"].predict_proba(rows).mean(0)
        return orc.link.f(ey) - orc.link.f(orc.fnull)
    exact = exact_shapley(value, M)[:, 1]
    est = np.zeros((R, M))
    for r in range(R):
        np.random.seed(1000 + r)
        est[r] = orc.explain(x[None], nsamples=S, l1_reg=False)[:, 1]
    se = est.std(0, ddof=1) / np.sqrt(R)
    z = (est.mean(0) - exact) / se
    assert se.max() < 0.01 * np.abs(exact).max()        # the bound below is a tight one
    assert np.abs(z).max() < 4.0, z


def test_full_plan_is_rng_free():
    a = build_plan(6, 62)[0:2]
    np.random.seed(5)
    b = build_plan(6, 62)[0:2]
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert len({tuple(r) for r in a[0]}) == 62


def test_degenerate_M_and_nonvarying_groups():
    prob = make_problem(seed=24, n=4, N=6, widths=(1, 2, 1, 3, 1), constant_groups=(1, 3))
    prob["bg"][:] = prob["bg"][0]
    prob["X"][1] = prob["bg"][0]
    prob["X"][2] = prob["bg"][0]
    prob["X"][2, 0] += 1.0
    orc = _oracle(prob)
    sv = orc.shap_values(prob["X"], nsamples=50, l1_reg=False)
    assert np.all(sv[1][1] == 0)
    fx = prob["clf"].predict_proba(prob["X"][2:3])[0]
    assert sv[1][2, 0] == pytest.approx(np.log(fx[1] / (1 - fx[1])) - orc.expected_value[1], abs=1e-12)
    assert np.all(sv[1][2, 1:] == 0)
    assert np.all(sv[1][:, [1, 3]] == 0)
    assert list(orc.varying_groups(prob["X"][0:1])) == [0, 2, 4]


def test_wrapper_seeding_and_batch_tuple():
    """KernelExplainerWrapper semantics (kernel_shap.py:225-254): seeding in the ctor makes runs reproducible; tuples
    carry the batch index through."""
    prob = make_problem(seed=25, n=4, N=6, widths=(1,) * 9)
    dd = DenseData(prob["bg"], prob["group_names"], prob["groups"])
    a = KernelExplainerWrapperOracle(prob["clf"].predict_proba, dd, link="logit", seed=3)
    ra = a.get_explanation(prob["X"], nsamples=100, l1_reg=False, silent=True)
    b = KernelExplainerWrapperOracle(prob["clf"].predict_proba, dd, link="logit", seed=3)
    idx, rb = b.get_explanation((7, prob["X"]), nsamples=100, l1_reg=False)
    assert idx == 7
    np.testing.assert_array_equal(ra[1], rb[1])
    assert b.return_attribute("vector_out") is True


def test_faithful_and_vectorised_run_agree_and_l1_branch_runs():
    prob = make_problem(seed=26, n=2, N=5, widths=(1,) * 8)
    np.random.seed(0)
    a = _oracle(prob, faithful_run=True).shap_values(prob["X"], nsamples=40, l1_reg=False)
    np.random.seed(0)
    b = _oracle(prob, faithful_run=False).shap_values(prob["X"], nsamples=40, l1_reg=False)
    np.testing.assert_allclose(a[1], b[1], atol=1e-12)
    np.random.seed(0)
    c = _oracle(prob).shap_values(prob["X"], nsamples=40, l1_reg="num_features(3)")
    assert (np.count_nonzero(c[1], axis=1) <= 4).all()           # 3 selected + the eliminated feature


def test_golden_vectors():
    """tests/golden/*.npz were generated by tests/golden/make_golden.py (oracle + brute-force Shapley values)."""
    import glob
    import os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
    assert files, "golden fixtures missing"
    from distributedkernelshap_b200.predictors import LinearSoftmaxClassifier
    for path in files:
        g = np.load(path, allow_pickle=True)
        groups = [list(map(int, x)) for x in g["groups"]]
        clf = LinearSoftmaxClassifier(g["coef"], g["intercept"], multi_class=str(g["multi_class"]))
        dd = DenseData(g["bg"], [f"g{i}" for i in range(len(groups))], groups, g["weights"])
        orc = KernelExplainerOracle(clf.predict_proba, dd, link=str(g["link"]))
        n = g["X"].shape[0]
        for i in range(n):
            plan = None if g["full"] else (g["Z"][i], g["w"][i])
            phi = orc.explain(g["X"][i:i + 1], plan=plan, nsamples=int(g["nsamples"]), l1_reg=False)
            np.testing.assert_allclose(phi, g["phi"][i], rtol=1e-10, atol=1e-12)
            if g["full"]:
                np.testing.assert_allclose(phi, g["phi_exact"][i], rtol=1e-8, atol=1e-10)
