"""Feature-selection step of ``shap.KernelExplainer.solve`` (the ``l1_reg`` branch) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

shap 0.35.0 hands the augmented, weighted regression problem to scikit-learn: ``LassoLarsIC(criterion=...)`` for
``l1_reg in ('auto', 'aic', 'bic')`` and ``lars_path(..., max_iter=r)`` for ``'num_features(r)'`` (SURVEY.md App. A.7;
reference call path explainers/kernel_shap.py:836-845, :880 -> upstream ``solve``).  The reference pins
**scikit-learn 0.23.2** (poetry.lock:455-458).  The scikit-learn in this image is 1.9: its ``LassoLarsIC`` no longer
normalises the columns (``normalize=True`` was the 0.23 default and is gone) and uses another information criterion
(noise variance from an OLS fit, log-likelihood form), so calling it would NOT reproduce the reference's selection.
This module therefore restates the published 0.23.2 algorithm in NumPy:

    sklearn 0.23.2 symbol                                        here
    -----------------------------------------------------------  ----------------------------
    linear_model._base._preprocess_data(normalize=True)          preprocess
    linear_model._least_angle._lars_path_solver (Gram branch)    lars_path_gram
    linear_model._least_angle.LassoLarsIC.fit                    lasso_lars_ic
    shap ... solve(): the three l1 sub-branches                  select_features

PINNED BY: ``tests/test_oracle_l1.py`` compares ``lars_path_gram`` with the installed ``sklearn.linear_model.lars_path``
(alphas, active set, coefficient path; the LARS iteration itself has not changed between 0.23 and 1.9 apart from a
rounding of the equiangular correlations added later, which ``round_corr`` reproduces), and ``lasso_lars_ic`` with a
composition of the installed ``lars_path`` on the normalised data and the 0.23.2 criterion written out independently.
"""
import numpy as np

TINY32 = np.finfo(np.float32).tiny
EQ_TOL = np.finfo(np.float32).eps
EPS64 = np.finfo(np.float64).eps


def min_pos(x):
    """sklearn.utils.arrayfuncs.min_pos: smallest strictly positive entry, or the largest float when there is none."""
    x = np.asarray(x, dtype=np.float64)
    pos = x[x > 0]
    return float(pos.min()) if pos.size else float(np.finfo(np.float64).max)


def preprocess(X, y):
    """``_preprocess_data(X, y, fit_intercept=True, normalize=True)``: centre the columns and y, scale every centred column
    to unit Euclidean norm (a zero norm is replaced by 1).  Returns ``(Xn, yc, X_offset, y_offset, X_scale)``."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    X_offset = X.mean(axis=0)
    Xc = X - X_offset
    X_scale = np.sqrt((Xc ** 2).sum(axis=0))
    X_scale[X_scale == 0.0] = 1.0
    return Xc / X_scale, y - y.mean(), X_offset, y.mean(), X_scale


def _cholesky_delete(L, go_out):
    """arrayfuncs.cholesky_delete: remove row/column ``go_out`` from the lower Cholesky factor L (n x n) in place with Givens
    rotations; the leading (n-1) x (n-1) block holds the result."""
    n = L.shape[0]
    for i in range(go_out, n - 1):
        L[i, :] = L[i + 1, :]
    for i in range(go_out, n - 1):
        a, b = L[i, i], L[i, i + 1]
        r = np.hypot(a, b)
        c, s = a / r, b / r
        L[i, i] = r
        L[i, i + 1] = 0.0
        for rr in range(i + 1, n - 1):
            t0, t1 = L[rr, i], L[rr, i + 1]
            L[rr, i] = c * t0 + s * t1
            L[rr, i + 1] = -s * t0 + c * t1


def lars_path_gram(Gram, Xy, n_samples, max_iter=500, method="lasso", alpha_min=0.0, eps=EPS64, round_corr=False):
    """Least-angle regression / lasso path from the Gram matrix ``X^T X`` and ``X^T y`` (the branch ``lars_path`` takes when
    ``n_samples > n_features``, always the case for KernelSHAP's augmented system).  Returns ``(alphas, active, coefs)``
    with ``coefs`` of shape ``(n_features, n_steps + 1)`` like scikit-learn.  ``round_corr``: round the equiangular
    correlations to 15 decimals (present in scikit-learn >= 0.24, absent from 0.23.2)."""
    Gram = np.array(Gram, dtype=np.float64)
    Cov = np.array(Xy, dtype=np.float64)
    Gram_copy, Cov_copy = Gram.copy(), Cov.copy()
    n_features = Cov.shape[0]
    max_features = min(max_iter, n_features)
    coefs = np.zeros((max_features + 1, n_features))
    alphas = np.zeros(max_features + 1)
    n_iter, n_active = 0, 0
    active, indices = [], np.arange(n_features)
    sign_active = np.zeros(max_features)
    L = np.zeros((max_features, max_features))
    drop = False
    lo = 0                                   # Cov[lo:] are the correlations of the inactive variables (Cov = Cov[1:] upstream)

    while True:
        inactive = Cov[lo:]
        if inactive.size:
            C_idx = int(np.argmax(np.abs(inactive)))
            C_ = inactive[C_idx]
            C = abs(C_)
        else:
            C_idx, C_, C = 0, 0.0, 0.0
        alphas[n_iter] = C / n_samples
        if alphas[n_iter] <= alpha_min + EQ_TOL:
            if abs(alphas[n_iter] - alpha_min) > EQ_TOL:
                if n_iter > 0:
                    ss = (alphas[n_iter - 1] - alpha_min) / (alphas[n_iter - 1] - alphas[n_iter])
                    coefs[n_iter] = coefs[n_iter - 1] + ss * (coefs[n_iter] - coefs[n_iter - 1])
                alphas[n_iter] = alpha_min
            break
        if n_iter >= max_iter or n_active >= n_features:
            break
        if not drop:
            # the most correlated variable joins the active set: grow the Cholesky factor of its Gram block by one row
            sign_active[n_active] = np.sign(C_)
            m, n = n_active, C_idx + n_active
            Cov[lo + C_idx], Cov[lo] = Cov[lo], Cov[lo + C_idx]
            indices[n], indices[m] = indices[m], indices[n]
            lo += 1
            Gram[[m, n]] = Gram[[n, m]]
            Gram[:, [m, n]] = Gram[:, [n, m]]
            c = Gram[n_active, n_active]
            L[n_active, :n_active] = Gram[n_active, :n_active]
            for r in range(n_active):                    # forward substitution L w = Gram[active, new]
                L[n_active, r] = (L[n_active, r] - L[r, :r] @ L[n_active, :r]) / L[r, r]
            v = L[n_active, :n_active] @ L[n_active, :n_active]
            diag = max(np.sqrt(abs(c - v)), eps)
            L[n_active, n_active] = diag
            if diag < 1e-7:                              # degenerate regressor: "dropped for good"
                lo -= 1
                Cov[lo] = 0.0
                Cov[lo + C_idx], Cov[lo] = Cov[lo], Cov[lo + C_idx]
                continue
            active.append(int(indices[n_active]))
            n_active += 1
        if method == "lasso" and n_iter > 0 and alphas[n_iter - 1] < alphas[n_iter]:
            break                                         # alpha increasing: numerical error dominates, bail out
        # equiangular direction: (L L^T) ls = sign
        Lk = L[:n_active, :n_active]
        sgn = sign_active[:n_active]
        tmp = np.zeros(n_active)
        for r in range(n_active):
            tmp[r] = (sgn[r] - Lk[r, :r] @ tmp[:r]) / Lk[r, r]
        least_squares = np.zeros(n_active)
        for r in range(n_active - 1, -1, -1):
            least_squares[r] = (tmp[r] - Lk[r + 1:, r] @ least_squares[r + 1:]) / Lk[r, r]
        if least_squares.size == 1 and least_squares[0] == 0:
            least_squares[...] = 1
            AA = 1.0
        else:
            AA = 1.0 / np.sqrt(np.sum(least_squares * sgn))
            if not np.isfinite(AA):
                raise FloatingPointError("lars: Cholesky factor too ill-conditioned (upstream regularises here)")
            least_squares = least_squares * AA
        corr_eq_dir = Gram[:n_active, n_active:].T @ least_squares
        if round_corr:
            corr_eq_dir = np.around(corr_eq_dir, decimals=np.finfo(np.float64).precision)
        inactive = Cov[lo:]
        g1 = min_pos((C - inactive) / (AA - corr_eq_dir + TINY32))
        g2 = min_pos((C + inactive) / (AA + corr_eq_dir + TINY32))
        gamma_ = min(g1, g2, C / AA)
        drop = False
        z = -coefs[n_iter][active] / (least_squares + TINY32)
        z_pos = min_pos(z)
        idx = []
        if z_pos < gamma_:
            idx = list(np.where(z == z_pos)[0][::-1])
            sign_active[idx] = -sign_active[idx]
            if method == "lasso":
                gamma_ = z_pos
            drop = True
        n_iter += 1
        if n_iter >= coefs.shape[0]:
            add = 2 * max(1, max_features - n_active)
            coefs = np.vstack([coefs, np.zeros((add, n_features))])
            alphas = np.concatenate([alphas, np.zeros(add)])
        coefs[n_iter][active] = coefs[n_iter - 1][active] + gamma_ * least_squares
        Cov[lo:] -= gamma_ * corr_eq_dir
        if drop and method == "lasso":
            for ii in idx:
                _cholesky_delete(L[:n_active, :n_active], ii)
            n_active -= 1
            drop_idx = [active.pop(ii) for ii in idx]
            for ii in idx:
                for i in range(ii, n_active):
                    indices[i], indices[i + 1] = indices[i + 1], indices[i]
                    Gram[[i, i + 1]] = Gram[[i + 1, i]]
                    Gram[:, [i, i + 1]] = Gram[:, [i + 1, i]]
            temp = Cov_copy[drop_idx] - Gram_copy[drop_idx] @ coefs[n_iter]
            lo -= len(drop_idx)
            Cov[lo:lo + len(drop_idx)] = temp
            sign_active = np.append(np.delete(sign_active, idx), 0.0)
    return alphas[:n_iter + 1], active, coefs[:n_iter + 1].T


def lasso_lars_ic(X, y, criterion="aic", max_iter=500):
    """``LassoLarsIC(criterion).fit(X, y).coef_`` of scikit-learn 0.23.2 (fit_intercept=True, normalize=True,
    precompute='auto'): lasso path on the centred, normalised data, information criterion per path step
    ``n * MSE / (var(y) + eps) + K * df`` with ``df`` = number of coefficients above machine epsilon and ``K = 2`` (aic)
    or ``log(n)`` (bic); the coefficients of the arg-min step, rescaled to the original columns."""
    Xn, yc, _, _, X_scale = preprocess(X, y)
    n_samples = Xn.shape[0]
    alphas, _, coef_path = lars_path_gram(Xn.T @ Xn, Xn.T @ yc, n_samples, max_iter=max_iter, method="lasso")
    K = 2.0 if criterion == "aic" else np.log(n_samples)
    R = yc[:, None] - Xn @ coef_path
    mse = np.mean(R ** 2, axis=0)
    sigma2 = np.var(yc)
    df = np.array([int(np.sum(np.abs(coef_path[:, k]) > EPS64)) for k in range(coef_path.shape[1])])
    crit = n_samples * mse / (sigma2 + EPS64) + K * df
    n_best = int(np.argmin(crit))
    return coef_path[:, n_best] / X_scale, dict(alphas=alphas, criterion=crit, n_best=n_best, coef_path=coef_path)


def select_features(l1_reg, mask_aug, eyAdj_aug):
    """``nonzero_inds`` of upstream ``solve`` for the string-valued ``l1_reg`` settings."""
    if isinstance(l1_reg, str) and l1_reg.startswith("num_features("):
        r = int(l1_reg[len("num_features("):-1])
        X = np.asarray(mask_aug, dtype=np.float64)
        y = np.asarray(eyAdj_aug, dtype=np.float64)
        # lars_path(X, y, max_iter=r): method='lar', no centring / normalisation, Gram='auto'
        _, active, _ = lars_path_gram(X.T @ X, X.T @ y, y.size, max_iter=r, method="lar")
        return np.asarray(active, dtype=int)
    if l1_reg in ("auto", "aic", "bic"):
        coef, _ = lasso_lars_ic(mask_aug, eyAdj_aug, "aic" if l1_reg == "auto" else l1_reg)
        return np.nonzero(coef)[0]
    raise ValueError(f"unsupported l1_reg {l1_reg!r}")
