"""Coalition plans: which subsets of the M varying feature groups are evaluated, and with what kernel weight.

Host-side product code (the north star keeps host code in Python).  It follows the enumeration + sampling rule
of ``shap.KernelExplainer.explain`` (shap==0.35.0; reached from the reference at explainers/kernel_shap.py:250/253,
see SURVEY.md App. A.4 steps 5-9) and emits each coalition as 64-bit words (bit k = k-th varying group
present; one word up to 64 groups, two up to 128, sixteen up to 1024), the layout ``dks_set_shared_plan`` /
``dks_explain_*`` consume.

The random part draws from a legacy ``numpy.random.RandomState`` -- or the global ``numpy.random`` module, the
stream the reference seeds (kernel_shap.py:228, :744) -- with exactly the calls upstream makes (one vectorised
``choice`` then one ``permutation(M)`` per draw), so a plan built here from a given stream state is the plan
shap would have built from it.
"""
from math import comb

import numpy as np

MAX_GROUPS = 1024         # coalition rows: one 64-bit word up to 64 varying groups, two up to 128, sixteen up to 1024


def resolve_nsamples(M, nsamples="auto"):
    """Rows an instance with ``M`` varying groups evaluates, and the size of the coalition space.

    'auto' = 2M + 2048; with M <= 30 the request is capped at 2**M - 2 (all proper, non-empty subsets)."""
    if nsamples in ("auto", None, 0):
        nsamples = 2 * M + 2 ** 11
    max_samples = 2 ** 30
    if M <= 30:
        max_samples = 2 ** M - 2
        nsamples = min(nsamples, max_samples)
    return int(nsamples), int(max_samples)


def size_weights(M):
    """Shapley-kernel mass of subset sizes 1..ceil((M-1)/2); sizes with a distinct complement size count twice."""
    n_sizes = int(np.ceil((M - 1) / 2.0))
    n_paired = int(np.floor((M - 1) / 2.0))
    wv = np.array([(M - 1.0) / (s * (M - s)) for s in range(1, n_sizes + 1)])
    wv[:n_paired] *= 2
    wv /= wv.sum()
    return wv, n_sizes, n_paired


def _combination_bits(M, size):
    """All size-``size`` subsets of range(M) as bit words, in itertools.combinations (lexicographic) order."""
    from itertools import combinations
    out = np.empty(comb(M, size), dtype=np.uint64)
    for r, inds in enumerate(combinations(range(M), size)):
        word = 0
        for k in inds:
            word |= 1 << k
        out[r] = word
    return out


def mask_words(M):
    """64-bit words per coalition row: the engine's kernels exist for rows of 1, 2 and 16 words (``dks_plan_words``)."""
    return 1 if M <= 64 else 2 if M <= 128 else 16


class CoalitionPlan:
    """``zbits`` uint64[S] (M <= 64) or uint64[S, W] (little-endian words; W = 2 for 64 < M <= 128, 16 above), ``weights``
    float64[S] in upstream row order, plus bookkeeping."""

    def __init__(self, M, zbits, weights, nfixed, num_full_subsets, weight_left):
        self.M = M
        self.zbits = zbits
        self.weights = weights
        self.nfixed = nfixed
        self.num_full_subsets = num_full_subsets
        self.weight_left = weight_left

    @property
    def S(self):
        return len(self.zbits)

    def dense(self):
        """[S, M] 0/1 matrix (upstream's ``maskMatrix``)."""
        words = self.zbits.reshape(len(self.zbits), -1)
        k = np.arange(self.M)
        return ((words[:, k // 64] >> (k % 64).astype(np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8)


def build_plan(M, nsamples="auto", rng=None):
    """Plan for ``M`` varying groups.  ``rng``: RandomState-like (``choice``/``permutation``) or None for the global
    ``numpy.random`` stream.  ``nsamples`` is the *request*; it is resolved with ``resolve_nsamples``."""
    if not 2 <= M <= MAX_GROUPS:
        raise ValueError(f"plans need 2 <= M <= {MAX_GROUPS} (got {M})")
    if rng is None:
        rng = np.random
    if M > 64:
        return _build_plan_wide(M, nsamples, rng)
    S, _ = resolve_nsamples(M, nsamples)
    full_mask = (1 << M) - 1
    wv, n_sizes, n_paired = size_weights(M)

    zbits = np.zeros(S, dtype=np.uint64)
    weights = np.zeros(S, dtype=np.float64)
    filled = 0

    # --- subset sizes that fit entirely in the budget are enumerated, each followed by its complement ------
    n_full = 0
    budget = S
    rem = wv.copy()
    for size in range(1, n_sizes + 1):
        paired = size <= n_paired
        n_sub = float(comb(M, size)) * (2 if paired else 1)
        if budget * rem[size - 1] / n_sub < 1.0 - 1e-8:
            break
        n_full += 1
        budget -= n_sub
        if rem[size - 1] < 1.0:
            rem /= (1 - rem[size - 1])
        w_row = wv[size - 1] / comb(M, size)
        words = _combination_bits(M, size)
        if paired:
            w_row /= 2.0
            block = np.empty(2 * len(words), dtype=np.uint64)
            block[0::2] = words
            block[1::2] = words ^ np.uint64(full_mask)
        else:
            block = words
        zbits[filled:filled + len(block)] = block
        weights[filled:filled + len(block)] = w_row
        filled += len(block)

    nfixed = filled
    weight_left = 0.0
    # --- the rest of the budget is sampled; repeated draws add to the weight of the first occurrence --------
    if n_full != n_sizes:
        left = S - filled
        p = wv.copy()
        p[:n_paired] /= 2  # a paired size yields two rows per draw
        p = p[n_full:]
        p /= p.sum()
        picks = rng.choice(len(p), 4 * left, p=p)
        first_row = {}
        pos = 0
        while left > 0 and pos < len(picks):
            size = int(picks[pos]) + n_full + 1
            pos += 1
            members = rng.permutation(M)[:size]
            word = 0
            for k in members:
                word |= 1 << int(k)
            row = first_row.get(word)
            fresh = row is None
            if fresh:
                first_row[word] = filled
                zbits[filled] = word
                weights[filled] = 1.0
                filled += 1
                left -= 1
            else:
                weights[row] += 1.0
            if left > 0 and size <= n_paired:
                if fresh:
                    zbits[filled] = word ^ full_mask
                    weights[filled] = 1.0
                    filled += 1
                    left -= 1
                else:
                    weights[row + 1] += 1.0  # the complement sits right after its original
        weight_left = float(wv[n_full:].sum())
        weights[nfixed:] *= weight_left / weights[nfixed:].sum()

    return CoalitionPlan(M, zbits, weights, nfixed, n_full, weight_left)


def _build_plan_wide(M, nsamples, rng):
    """Same rule for M > 64, with Python integers as masks (``mask_words(M)`` 64-bit words per row on the way out)."""
    from itertools import combinations
    S, _ = resolve_nsamples(M, nsamples)
    full_mask = (1 << M) - 1
    wv, n_sizes, n_paired = size_weights(M)
    rows, weights = [], []
    n_full, budget, rem = 0, S, wv.copy()
    for size in range(1, n_sizes + 1):
        paired = size <= n_paired
        n_sub = float(comb(M, size)) * (2 if paired else 1)
        if budget * rem[size - 1] / n_sub < 1.0 - 1e-8:
            break
        n_full += 1
        budget -= n_sub
        if rem[size - 1] < 1.0:
            rem /= (1 - rem[size - 1])
        w_row = wv[size - 1] / comb(M, size) / (2.0 if paired else 1.0)
        for inds in combinations(range(M), size):
            word = 0
            for k in inds:
                word |= 1 << k
            rows.append(word)
            weights.append(w_row)
            if paired:
                rows.append(word ^ full_mask)
                weights.append(w_row)
    nfixed = len(rows)
    weight_left = 0.0
    if n_full != n_sizes:
        left = S - nfixed
        p = wv.copy()
        p[:n_paired] /= 2
        p = p[n_full:]
        p /= p.sum()
        picks = rng.choice(len(p), 4 * left, p=p)
        first_row, pos = {}, 0
        while left > 0 and pos < len(picks):
            size = int(picks[pos]) + n_full + 1
            pos += 1
            word = 0
            for k in rng.permutation(M)[:size]:
                word |= 1 << int(k)
            row = first_row.get(word)
            fresh = row is None
            if fresh:
                first_row[word] = len(rows)
                rows.append(word)
                weights.append(1.0)
                left -= 1
            else:
                weights[row] += 1.0
            if left > 0 and size <= n_paired:
                if fresh:
                    rows.append(word ^ full_mask)
                    weights.append(1.0)
                    left -= 1
                else:
                    weights[row + 1] += 1.0
        weight_left = float(wv[n_full:].sum())
        wts = np.array(weights)
        wts[nfixed:] *= weight_left / wts[nfixed:].sum()
        weights = list(wts)
    W = mask_words(M)
    zbits = np.zeros((S, W), dtype=np.uint64)
    wout = np.zeros(S)
    lo64 = (1 << 64) - 1
    for r, word in enumerate(rows):
        for q in range((M + 63) // 64):
            zbits[r, q] = (word >> (64 * q)) & lo64
    wout[:len(weights)] = weights
    return CoalitionPlan(M, zbits, wout, nfixed, n_full, weight_left)


def sampling_info(plan):
    """What the device-side sampler (csrc/dks_sampler.cuh, ``dks_set_plan_sampling``) needs to continue a plan past its
    enumerated prefix: ``(nfixed, n_full, n_paired, cdf float64[ncdf], weight_left)``.  ``cdf`` is the cumulative
    distribution of the subset sizes left to sample -- upstream's ``remaining_weight_vector`` after halving the paired
    sizes and renormalising -- and is empty when the plan is fully enumerated."""
    wv, n_sizes, n_paired = size_weights(plan.M)
    n_full = plan.num_full_subsets
    if n_full == n_sizes:
        return plan.nfixed, n_full, n_paired, np.zeros(0), 0.0
    p = wv.copy()
    p[:n_paired] /= 2
    p = p[n_full:]
    p /= p.sum()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    return plan.nfixed, n_full, n_paired, np.ascontiguousarray(cdf), float(wv[n_full:].sum())


def pack_dense_plan(Z):
    """[S, M] 0/1 matrix -> uint64[S] bit words, or uint64[S, mask_words(M)] for M > 64 (for feeding externally built
    plans to the engine and for comparing plans in tests)."""
    Z = np.asarray(Z)
    S, M = Z.shape
    if M > MAX_GROUPS:
        raise ValueError(f"at most {MAX_GROUPS} varying groups per coalition row")
    words = np.zeros((S, mask_words(M)), dtype=np.uint64)
    for q in range((M + 63) // 64):
        blk = Z[:, 64 * q:64 * q + 64].astype(np.uint64)
        k = np.arange(blk.shape[1], dtype=np.uint64)
        words[:, q] = (blk << k[None, :]).sum(axis=1, dtype=np.uint64)
    return words[:, 0] if M <= 64 else words


def projection(plan):
    """Projection form of upstream's constrained WLS for a shared plan (``KernelExplainer.solve`` without the l1 branch):
    with the last group eliminated, ``E = Z[:, :M-1] - Z[:, M-1:]``, ``A = E^T diag(w) E`` and
    ``beta = inv(A) E^T diag(w) (y - z_L delta) = P y - delta d``.  Returns ``(PT [S, M-1], d [M-1])`` in float64 with
    ``PT = P^T`` (one row per coalition) and ``d = P z_L``.  Plans of up to 128 groups get their factorisation on the
    device (``dks_set_shared_plan``); wider plans are factored here once -- ``np.linalg.inv`` like upstream -- and uploaded
    with ``dks_set_plan_projection``."""
    Z = plan.dense().astype(np.float64)
    w = np.asarray(plan.weights, dtype=np.float64)
    zl = Z[:, -1]
    E = Z[:, :-1] - zl[:, None]
    EW = E * w[:, None]
    A = E.T @ EW
    P = np.linalg.inv(A) @ EW.T
    return np.ascontiguousarray(P.T), np.ascontiguousarray(P @ zl)


def l1_tables(plan):
    """Per-plan tables of the l1 feature-selection step (``l1_reg='auto' | 'aic' | 'bic' | 'num_features(k)'`` of upstream
    ``solve``, which regresses on the AUGMENTED system ``[sqrt(w (M - |z|)) z ; sqrt(w |z|) (z - 1)]``).  Everything that
    depends on the plan only -- the Gram matrix of the augmented columns, raw and centred / normalised the way
    scikit-learn 0.23.2's ``LassoLarsIC`` does it, the column sums, the weighted Gram of the plain rows for the final
    restricted WLS, and the per-row weights -- is formed here once in float64 (host Python, like the plan itself); the
    engine then needs only per-instance moment vectors of y (``dks_set_l1_tables`` / csrc/dks_l1.cuh).

    Returns a dict of C-contiguous float64 arrays: ``gram_raw`` [M, M], ``gram_norm`` [M, M], ``colsum`` [M],
    ``scale`` [M], ``bz`` [M] (sum_s b_s z_sk), ``gram_w`` [M, M] (sum_s w_s z_sk z_sl), ``wz`` [M] (sum_s w_s z_sk),
    ``a`` / ``b`` / ``sqa`` / ``sqb`` [S], and the scalars ``sum_b``, ``sum_sqb``, ``sum_w``, ``n_aug``."""
    Z = plan.dense().astype(np.float64)
    w = np.asarray(plan.weights, dtype=np.float64)
    M = plan.M
    sz = Z.sum(axis=1)
    a = w * (M - sz)
    b = w * sz
    sqa, sqb = np.sqrt(a), np.sqrt(b)
    Zc = 1.0 - Z
    gram_raw = Z.T @ (a[:, None] * Z) + Zc.T @ (b[:, None] * Zc)
    colsum = Z.T @ sqa - Zc.T @ sqb
    n_aug = 2 * len(w)
    gram_c = gram_raw - np.outer(colsum, colsum) / n_aug
    scale = np.sqrt(np.maximum(np.diag(gram_c), 0.0))
    scale[scale == 0.0] = 1.0
    gram_norm = gram_c / np.outer(scale, scale)
    out = dict(gram_raw=gram_raw, gram_norm=gram_norm, colsum=colsum, scale=scale, bz=Z.T @ b,
               gram_w=Z.T @ (w[:, None] * Z), wz=Z.T @ w, a=a, b=b, sqa=sqa, sqb=sqb)
    out = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in out.items()}
    out.update(sum_b=float(b.sum()), sum_sqb=float(sqb.sum()), sum_w=float(w.sum()), n_aug=int(n_aug))
    return out
