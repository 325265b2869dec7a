"""Pins the oracle's restatement of scikit-learn 0.23.2's feature selection (oracle/sklearn_lars_restated.py), the part
of upstream ``solve`` behind ``l1_reg='auto' | 'aic' | 'bic' | 'num_features(k)'``.  The LARS iteration is checked against
the scikit-learn installed here (the iteration is unchanged since 0.23 except for a rounding step, reproduced with
``round_corr=True``); the 0.23.2 information criterion is checked against an independent composition."""
import warnings

import numpy as np
import pytest

from oracle.shap_kernel_oracle import build_plan, effective_nsamples
from oracle.sklearn_lars_restated import lars_path_gram, lasso_lars_ic, preprocess, select_features


def augmented_problem(M, nsamples, seed, sparsity=0.4, noise=0.05):
    """The regression upstream ``solve`` hands to scikit-learn: (mask_aug, eyAdj_aug) of a KernelSHAP plan."""
    rng = np.random.default_rng(seed)
    S, _ = effective_nsamples(M, nsamples)
    np.random.seed(seed)
    Z, w, _ = build_plan(M, S)
    Z = Z.astype(float)
    s = Z.sum(1)
    beta = rng.normal(0, 1, M) * (rng.random(M) < sparsity)
    y = Z @ beta + noise * rng.standard_normal(S)
    delta = beta.sum()
    sq = np.sqrt(np.hstack((w * (M - s), w * s)))
    return (sq * np.vstack((Z, Z - 1)).T).T, np.hstack((y, y - delta)) * sq


@pytest.mark.parametrize("M,nsamples,seed", [(8, 100, 0), (16, 400, 1), (30, 1000, 2), (64, 1000, 3), (12, 300, 4)])
def test_lars_iteration_matches_installed_sklearn(M, nsamples, seed):
    from sklearn.linear_model import lars_path
    X, y = augmented_problem(M, nsamples, seed)
    Xn, yc, *_ = preprocess(X, y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for A, b in ((Xn, yc), (X, y)):
            for method in ("lasso", "lar"):
                a0, act0, c0 = lars_path(A, b, method=method, Gram="auto")
                a1, act1, c1 = lars_path_gram(A.T @ A, A.T @ b, b.size, method=method, round_corr=True)
                assert list(act0) == list(act1)
                np.testing.assert_allclose(a1, a0, rtol=0, atol=1e-10)
                np.testing.assert_allclose(c1, c0, rtol=0, atol=1e-10)
                _, act2, c2 = lars_path_gram(A.T @ A, A.T @ b, b.size, method=method)       # 0.23.2: no rounding step
                assert list(act2) == list(act0)
                np.testing.assert_allclose(c2, c0, rtol=0, atol=1e-8)
        for r in (1, 3, 5):
            _, act0, _ = lars_path(X, y, max_iter=r)
            assert list(select_features(f"num_features({r})", X, y)) == list(act0)


def test_lasso_path_with_drops_matches_installed_sklearn():
    """Correlated regressors make coefficients cross zero (variables leave the active set): the Cholesky down-date and
    the covariance re-computation of the lasso variant."""
    from sklearn.linear_model import lars_path
    rng = np.random.default_rng(7)
    dropped = 0
    for trial in range(6):
        n, p = 60, 12
        base = rng.standard_normal((n, 4))
        X = base @ rng.standard_normal((4, p)) + 0.3 * rng.standard_normal((n, p))
        y = X @ (rng.standard_normal(p) * (rng.random(p) < 0.5)) + 0.5 * rng.standard_normal(n)
        Xn, yc, *_ = preprocess(X, y)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a0, act0, c0 = lars_path(Xn, yc, method="lasso", Gram="auto")
        a1, act1, c1 = lars_path_gram(Xn.T @ Xn, Xn.T @ yc, n, method="lasso", round_corr=True)
        assert c0.shape == c1.shape and list(act0) == list(act1)
        np.testing.assert_allclose(c1, c0, rtol=0, atol=1e-9)
        dropped += int(c0.shape[1] > p + 1)
    assert dropped > 0, "no trial exercised a drop"


@pytest.mark.parametrize("criterion", ["aic", "bic"])
def test_information_criterion_of_0_23_2(criterion):
    """LassoLarsIC of scikit-learn 0.23.2 = lasso path on centred, unit-norm columns + n * MSE / var(y) + K * df.
    Independent composition: the installed lars_path for the path, the criterion written out here."""
    from sklearn.linear_model import lars_path
    for M, ns, seed in [(16, 400, 5), (30, 800, 6), (64, 1000, 8)]:
        X, y = augmented_problem(M, ns, seed)
        coef, info = lasso_lars_ic(X, y, criterion)
        Xc = X - X.mean(0)
        scale = np.sqrt((Xc ** 2).sum(0))
        Xn, yc = Xc / scale, y - y.mean()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _, _, path = lars_path(Xn, yc, method="lasso", Gram="auto", alpha_min=0.0, max_iter=500)
        n = y.size
        K = 2.0 if criterion == "aic" else np.log(n)
        crit = [n * np.mean((yc - Xn @ path[:, k]) ** 2) / (np.var(yc) + np.finfo(float).eps)
                + K * np.sum(np.abs(path[:, k]) > np.finfo(float).eps) for k in range(path.shape[1])]
        best = int(np.argmin(crit))
        assert best == info["n_best"]
        np.testing.assert_allclose(coef, path[:, best] / scale, rtol=0, atol=1e-9)
        assert 0 < np.count_nonzero(coef) < M                       # a real selection took place


def test_oracle_l1_branches_select_and_stay_additive():
    from conftest import make_problem
    from oracle.shap_kernel_oracle import DenseData, KernelExplainerOracle
    prob = make_problem(seed=8, n=3, N=6, widths=(1,) * 16)
    orc = KernelExplainerOracle(prob["clf"].predict_proba, DenseData(prob["bg"], prob["group_names"], prob["groups"]),
                                link="logit")
    fx = prob["clf"].predict_proba(prob["X"])
    for l1 in ("auto", "aic", "bic", "num_features(4)"):
        np.random.seed(0)
        sv = orc.shap_values(prob["X"], nsamples=200, l1_reg=l1)
        np.testing.assert_allclose(sv[1].sum(1), np.log(fx[:, 1] / fx[:, 0]) - orc.expected_value[1], atol=1e-9)
        assert (np.count_nonzero(sv[1], axis=1) < 16).any() or l1 == "auto"
    np.random.seed(0)
    assert (np.count_nonzero(orc.shap_values(prob["X"], nsamples=200, l1_reg="num_features(4)")[1], axis=1) == 4).all()
