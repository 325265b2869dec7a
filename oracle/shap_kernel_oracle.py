"""CPU oracle for the KernelSHAP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product (``distributedkernelshap_b200``) never does.

What it restates
----------------
The reference (alexcoca/DistributedKernelShap @ 04c96d4) does not contain the arithmetic of its hot
path: ``explainers/kernel_shap.py:217`` subclasses ``shap.KernelExplainer`` and
``explainers/kernel_shap.py:250,253`` call ``super().shap_values``.  The algorithm lives in the
third-party dependency **shap == 0.35.0** (pinned at ``poetry.lock:483-486``; constraint
``pyproject.toml:16``), file ``shap/explainers/kernel.py`` (+ ``shap/common.py``), which is absent
from ``/root/reference`` and not installable offline.  This module restates that published
algorithm in NumPy (float64 end to end, legacy global ``np.random`` MT19937 stream exactly as the
reference seeds it at ``explainers/kernel_shap.py:228`` and ``:744``), function by function:

    upstream symbol (shap 0.35.0)                 here
    --------------------------------------------  ----------------------------------
    shap.common.IdentityLink / LogitLink          IdentityLink / LogitLink / convert_to_link
    shap.common.DenseData                         DenseData
    KernelExplainer.__init__                      KernelExplainerOracle.__init__
    KernelExplainer.shap_values (dense 2-D path)  KernelExplainerOracle.shap_values
    KernelExplainer.explain                       KernelExplainerOracle.explain (+ build_plan)
    KernelExplainer.varying_groups / not_equal    KernelExplainerOracle.varying_groups
    KernelExplainer.allocate / addsample          KernelExplainerOracle._masked_batch
    KernelExplainer.run                           KernelExplainerOracle._run
    KernelExplainer.solve                         KernelExplainerOracle._solve
    explainers.kernel_shap.KernelExplainerWrapper KernelExplainerWrapperOracle (kernel_shap.py:217-261)

PARITY UNPINNED: the reference ships no tests, golden vectors or saved explanations (SURVEY.md §4,
§8c) and shap cannot be imported here, so this restatement cannot be checked against outputs of the
reference itself.  It is pinned instead by analytic known answers in ``tests/test_oracle.py``:
exact Shapley values by brute-force subset enumeration, the affine-model closed form, additivity,
two-class antisymmetry, plan invariants (counts / weight sums computed from the upstream rule).
"""

import copy
import itertools
import logging

import numpy as np
from scipy.special import binom

log = logging.getLogger(__name__)


# --------------------------------------------------------------------------------------------------
# shap.common: links and DenseData  (reference call sites: kernel_shap.py:15, :594, :616, :646, :949)
# --------------------------------------------------------------------------------------------------

class IdentityLink:
    def __str__(self):
        return "identity"

    @staticmethod
    def f(x):
        return x

    @staticmethod
    def finv(x):
        return x


class LogitLink:
    def __str__(self):
        return "logit"

    @staticmethod
    def f(x):
        return np.log(x / (1 - x))

    @staticmethod
    def finv(x):
        return 1 / (1 + np.exp(-x))


def convert_to_link(val):
    """shap.common.convert_to_link: accept a link object or the strings 'identity' / 'logit'."""
    if isinstance(val, (IdentityLink, LogitLink)):
        return val
    if val == "identity":
        return IdentityLink()
    if val == "logit":
        return LogitLink()
    raise ValueError("Passed link object must be a subclass of iml.Link")


class DenseData:
    """shap.common.DenseData(data, group_names, groups=None, weights=None).

    ``groups`` defaults to one singleton group per column, ``weights`` to ones and is normalised to
    sum 1.  When the group sizes add up to ``data.shape[0]`` rather than ``data.shape[1]`` upstream
    treats the matrix as transposed (the reference warns about it at kernel_shap.py:443-449).
    """

    def __init__(self, data, group_names, *args):
        data = np.asarray(data)
        self.groups = args[0] if len(args) > 0 and args[0] is not None \
            else [np.array([i]) for i in range(len(group_names))]
        length = sum(len(g) for g in self.groups)
        num_samples = data.shape[0]
        t = False
        if length != data.shape[1]:
            t = True
            num_samples = data.shape[1]
        valid = (not t and length == data.shape[1]) or (t and length == data.shape[0])
        assert valid, "# of names must match data matrix!"
        self.weights = args[1] if len(args) > 1 and args[1] is not None else np.ones(num_samples)
        self.weights = np.asarray(self.weights, dtype=np.float64)
        self.weights = self.weights / np.sum(self.weights)
        wl = len(self.weights)
        valid = (not t and wl == data.shape[0]) or (t and wl == data.shape[1])
        assert valid, "# weights must match data matrix!"
        self.transposed = t
        self.group_names = group_names
        self.data = data
        self.groups_size = len(self.groups)


def convert_to_data(val):
    """shap.common.convert_to_data for the dense inputs the reference hands over."""
    if isinstance(val, DenseData):
        return val
    val = np.asarray(val)
    if val.ndim == 1:
        val = val.reshape(1, -1)
    return DenseData(val, [str(i) for i in range(val.shape[1])])


# --------------------------------------------------------------------------------------------------
# The coalition plan (KernelExplainer.explain, the part between allocate() and run())
# --------------------------------------------------------------------------------------------------

def shapley_size_weights(M):
    """Normalised Shapley-kernel mass per subset size 1..ceil((M-1)/2) with paired sizes doubled."""
    num_subset_sizes = int(np.ceil((M - 1) / 2.0))
    num_paired_subset_sizes = int(np.floor((M - 1) / 2.0))
    weight_vector = np.array([(M - 1.0) / (i * (M - i)) for i in range(1, num_subset_sizes + 1)])
    weight_vector[:num_paired_subset_sizes] *= 2
    weight_vector /= np.sum(weight_vector)
    return weight_vector, num_subset_sizes, num_paired_subset_sizes


def effective_nsamples(M, nsamples="auto"):
    """nsamples resolution of upstream ``explain``: 'auto' = 2M + 2**11, capped at 2**M - 2 for M <= 30."""
    if nsamples == "auto" or nsamples is None:
        nsamples = 2 * M + 2 ** 11
    max_samples = 2 ** 30
    if M <= 30:
        max_samples = 2 ** M - 2
        if nsamples > max_samples:
            nsamples = max_samples
    return int(nsamples), int(max_samples)


def build_plan(M, nsamples, rng=None):
    """Coalition matrix ``Z`` (S x M, 0/1) and kernel weights ``w`` (S) in upstream row order.

    Restates the enumeration + sampling section of ``KernelExplainer.explain`` (shap 0.35.0).  ``rng``
    is anything exposing ``choice`` and ``permutation`` -- the ``np.random`` module itself by default,
    i.e. the global legacy stream the reference seeds.  ``nsamples`` must already be capped
    (``effective_nsamples``).  Returns ``(Z uint8[S,M], w float64[S], info dict)``.
    """
    if rng is None:
        rng = np.random
    S = int(nsamples)
    Z = np.zeros((S, M), dtype=np.uint8)
    w = np.zeros(S, dtype=np.float64)
    added = 0

    weight_vector, num_subset_sizes, num_paired_subset_sizes = shapley_size_weights(M)

    # sizes that fit completely in the budget are enumerated (with their complements)
    num_full_subsets = 0
    num_samples_left = S
    mask = np.zeros(M)
    remaining_weight_vector = copy.copy(weight_vector)
    for subset_size in range(1, num_subset_sizes + 1):
        nsubsets = binom(M, subset_size)
        if subset_size <= num_paired_subset_sizes:
            nsubsets *= 2
        if num_samples_left * remaining_weight_vector[subset_size - 1] / nsubsets >= 1.0 - 1e-8:
            num_full_subsets += 1
            num_samples_left -= nsubsets
            if remaining_weight_vector[subset_size - 1] < 1.0:
                remaining_weight_vector /= (1 - remaining_weight_vector[subset_size - 1])
            wt = weight_vector[subset_size - 1] / binom(M, subset_size)
            if subset_size <= num_paired_subset_sizes:
                wt /= 2.0
            for inds in itertools.combinations(range(M), subset_size):
                mask[:] = 0.0
                mask[np.array(inds, dtype="int64")] = 1.0
                Z[added] = mask
                w[added] = wt
                added += 1
                if subset_size <= num_paired_subset_sizes:
                    mask[:] = np.abs(mask - 1)
                    Z[added] = mask
                    w[added] = wt
                    added += 1
        else:
            break

    # what is left of the budget is drawn at random, duplicates folded into the weight
    nfixed_samples = added
    samples_left = S - added
    weight_left = 0.0
    if num_full_subsets != num_subset_sizes:
        remaining_weight_vector = copy.copy(weight_vector)
        remaining_weight_vector[:num_paired_subset_sizes] /= 2  # two samples are drawn per pick below
        remaining_weight_vector = remaining_weight_vector[num_full_subsets:]
        remaining_weight_vector /= np.sum(remaining_weight_vector)
        ind_set = rng.choice(len(remaining_weight_vector), 4 * samples_left, p=remaining_weight_vector)
        ind_set_pos = 0
        used_masks = {}
        while samples_left > 0 and ind_set_pos < len(ind_set):
            mask.fill(0.0)
            ind = ind_set[ind_set_pos]
            ind_set_pos += 1
            subset_size = ind + num_full_subsets + 1
            mask[rng.permutation(M)[:subset_size]] = 1.0

            mask_tuple = tuple(mask)
            new_sample = False
            if mask_tuple not in used_masks:
                new_sample = True
                used_masks[mask_tuple] = added
                samples_left -= 1
                Z[added] = mask
                w[added] = 1.0
                added += 1
            else:
                w[used_masks[mask_tuple]] += 1.0

            if samples_left > 0 and subset_size <= num_paired_subset_sizes:
                mask[:] = np.abs(mask - 1)
                if new_sample:
                    samples_left -= 1
                    Z[added] = mask
                    w[added] = 1.0
                    added += 1
                else:
                    # the complement was stored right after its original
                    w[used_masks[mask_tuple] + 1] += 1.0

        weight_left = np.sum(weight_vector[num_full_subsets:])
        w[nfixed_samples:] *= weight_left / w[nfixed_samples:].sum()

    info = dict(M=M, nsamples=S, nfixed=nfixed_samples, nadded=added, num_full_subsets=num_full_subsets,
                num_subset_sizes=num_subset_sizes, num_paired_subset_sizes=num_paired_subset_sizes,
                weight_left=float(weight_left))
    return Z, w, info


# --------------------------------------------------------------------------------------------------
# KernelExplainer
# --------------------------------------------------------------------------------------------------

class KernelExplainerOracle:
    """NumPy restatement of ``shap.KernelExplainer`` (0.35.0) for dense data.

    Parameters mirror upstream: ``model`` is a callable ``f(X) -> [rows] | [rows, C]``, ``data`` the
    background (array or ``DenseData``), ``link`` 'identity' | 'logit'.

    Extras for testing (no upstream counterpart):
      * ``explain(..., plan=(Z, w))`` evaluates a caller-supplied coalition plan instead of drawing one
        (lets the CUDA path and the oracle consume *identical inputs*);
      * ``record_plans=True`` keeps each instance's ``(varyingInds, Z, w)`` in ``self.plans``;
      * ``faithful_run=True`` keeps upstream's interpreted ``S x N`` reduction loop in ``run()`` (what the
        reference actually pays for); ``False`` uses one matrix product with the same result up to
        float64 summation order.
    """

    def __init__(self, model, data, link="identity", faithful_run=False, record_plans=False, rng=None,
                 chunk_rows=None):
        self.chunk_rows = chunk_rows    # coalitions per masked-batch chunk (None: the whole S*N x D batch at once)
        self.link = convert_to_link(link)
        self.model = model
        self.data = convert_to_data(data)
        assert not self.data.transposed, "transposed DenseData is not handled by the oracle"
        self.faithful_run = faithful_run
        self.record_plans = record_plans
        self.plans = []
        self.rng = rng

        self.N = self.data.data.shape[0]
        self.P = self.data.data.shape[1]
        self.linkfv = np.vectorize(self.link.f)
        self.nsamplesAdded = 0
        self.nsamplesRun = 0

        model_null = np.asarray(self.model(self.data.data))
        self.fnull = np.sum((model_null.T * self.data.weights).T, 0)
        self.expected_value = self.linkfv(self.fnull)

        self.vector_out = True
        if len(self.fnull.shape) == 0:
            self.vector_out = False
            self.fnull = np.array([self.fnull])
            self.D = 1
            self.expected_value = float(self.expected_value)
        else:
            self.D = self.fnull.shape[0]

    # ---- KernelExplainer.shap_values (dense branch) -------------------------------------------------
    def shap_values(self, X, **kwargs):
        X = np.asarray(X)
        assert X.ndim in (1, 2), "Instance must have 1 or 2 dimensions!"
        if X.ndim == 1:
            explanation = self.explain(X.reshape((1, X.shape[0])), **kwargs)
            s = explanation.shape
            if len(s) == 2:
                return [explanation[:, i] for i in range(s[1])]
            return explanation

        explanations = [self.explain(X[i:i + 1, :], **kwargs) for i in range(X.shape[0])]
        s = explanations[0].shape
        if len(s) == 2:
            outs = [np.zeros((X.shape[0], s[0])) for _ in range(s[1])]
            for i in range(X.shape[0]):
                for j in range(s[1]):
                    outs[j][i] = explanations[i][:, j]
            return outs
        out = np.zeros((X.shape[0], s[0]))
        for i in range(X.shape[0]):
            out[i] = explanations[i]
        return out

    # ---- KernelExplainer.varying_groups / not_equal ------------------------------------------------
    def varying_groups(self, x):
        varying = np.zeros(self.data.groups_size)
        for i in range(self.data.groups_size):
            inds = np.asarray(self.data.groups[i])
            x_group = x[0, inds]
            # np.isclose(x, data) with defaults rtol=1e-5, atol=1e-8 and NaN == NaN
            mism = ~np.isclose(x_group[None, :], self.data.data[:, inds], equal_nan=True)
            varying[i] = np.sum(mism) > 0
        return np.nonzero(varying)[0]

    # ---- KernelExplainer.explain --------------------------------------------------------------------
    def explain(self, x, plan=None, **kwargs):
        x = np.asarray(x, dtype=np.float64).reshape(1, -1)
        assert x.shape[1] == self.P

        self.varyingInds = self.varying_groups(x)
        self.varyingFeatureGroups = [np.asarray(self.data.groups[i]) for i in self.varyingInds]
        self.M = len(self.varyingFeatureGroups)

        model_out = np.asarray(self.model(x))
        self.fx = model_out[0]
        if not self.vector_out:
            self.fx = np.array([self.fx])

        G = self.data.groups_size
        if self.M == 0:
            phi = np.zeros((G, self.D))
        elif self.M == 1:
            phi = np.zeros((G, self.D))
            diff = self.link.f(self.fx) - self.link.f(self.fnull)
            for d in range(self.D):
                phi[self.varyingInds[0], d] = diff[d]
        else:
            self.l1_reg = kwargs.get("l1_reg", "auto")
            self.nsamples, self.max_samples = effective_nsamples(self.M, kwargs.get("nsamples", "auto"))

            if plan is None:
                Z, w, _ = build_plan(self.M, self.nsamples, self.rng)
            else:
                Z, w = plan
                Z = np.asarray(Z)
                w = np.asarray(w, dtype=np.float64)
                assert Z.shape == (self.nsamples, self.M) and w.shape == (self.nsamples,), \
                    "plan must be [nsamples, M] / [nsamples] for this instance"
            self.maskMatrix = Z.astype(np.float64)
            self.kernelWeights = w
            if self.record_plans:
                self.plans.append((self.varyingInds.copy(), Z.copy(), w.copy()))

            if self.chunk_rows:
                self._run_chunked(x, Z, int(self.chunk_rows))
            else:
                self._run(self._masked_batch(x, Z))

            phi = np.zeros((G, self.D))
            for d in range(self.D):
                vphi = self._solve(self.nsamples / self.max_samples, d)
                phi[self.varyingInds, d] = vphi

        if self.record_plans and self.M < 2:
            self.plans.append((self.varyingInds.copy(), None, None))
        if not self.vector_out:
            phi = np.squeeze(phi, axis=1)
        return phi

    # ---- allocate() + addsample(): the mask/impute stage ------------------------------------------
    def _masked_batch(self, x, Z):
        S = Z.shape[0]
        synth_data = np.tile(self.data.data, (S, 1))
        for s in range(S):
            offset = s * self.N
            for j in range(self.M):
                if Z[s, j] == 1.0:
                    grp = self.varyingFeatureGroups[j]
                    synth_data[offset:offset + self.N, grp] = x[0, grp]
        return synth_data

    # ---- run(): predict + background reduction -----------------------------------------------------
    def _run(self, synth_data):
        S = synth_data.shape[0] // self.N
        modelOut = np.asarray(self.model(synth_data))
        y = np.reshape(modelOut, (S * self.N, self.D))
        if self.faithful_run:
            ey = np.zeros((S, self.D))
            for i in range(S):
                eyVal = np.zeros(self.D)
                for j in range(self.N):
                    eyVal += y[i * self.N + j, :] * self.data.weights[j]
                ey[i, :] = eyVal
        else:
            ey = np.einsum("sjd,j->sd", y.reshape(S, self.N, self.D), self.data.weights)
        self.y = y
        self.ey = ey

    def _run_chunked(self, x, Z, chunk):
        """allocate/addsample/run over `chunk` coalitions at a time: the same rows, model calls and weighted background
        means as ``_run(_masked_batch(x, Z))`` without holding the whole S*N x D batch (17 GB for BASELINE configs[3])."""
        S = Z.shape[0]
        ey = np.zeros((S, self.D))
        for s0 in range(0, S, chunk):
            Zc = Z[s0:s0 + chunk]
            c = Zc.shape[0]
            synth = np.tile(self.data.data, (c, 1)).reshape(c, self.N, self.P)
            for j in range(self.M):
                grp = self.varyingFeatureGroups[j]
                rows = np.nonzero(Zc[:, j] == 1)[0]
                if len(rows):
                    synth[np.ix_(rows, np.arange(self.N), grp)] = x[0, grp]
            y = np.reshape(np.asarray(self.model(synth.reshape(c * self.N, self.P))), (c, self.N, self.D))
            ey[s0:s0 + c] = np.einsum("sjd,j->sd", y, self.data.weights)
        self.y = None
        self.ey = ey

    # ---- solve(): constrained weighted least squares ------------------------------------------------
    def _solve(self, fraction_evaluated, dim):
        eyAdj = self.linkfv(self.ey[:, dim]) - self.link.f(self.fnull[dim])
        s = np.sum(self.maskMatrix, 1)

        nonzero_inds = np.arange(self.M)
        if (self.l1_reg not in ["auto", False, 0]) or (fraction_evaluated < 0.2 and self.l1_reg == "auto"):
            w_aug = np.hstack((self.kernelWeights * (self.M - s), self.kernelWeights * s))
            w_sqrt_aug = np.sqrt(w_aug)
            eyAdj_aug = np.hstack((eyAdj, eyAdj - (self.link.f(self.fx[dim]) - self.link.f(self.fnull[dim]))))
            eyAdj_aug *= w_sqrt_aug
            mask_aug = np.transpose(w_sqrt_aug * np.transpose(np.vstack((self.maskMatrix, self.maskMatrix - 1))))
            if isinstance(self.l1_reg, str):
                # 'num_features(r)' -> lars_path(max_iter=r); 'auto' / 'aic' / 'bic' -> LassoLarsIC: scikit-learn 0.23.2
                # semantics (the reference's pin), restated in sklearn_lars_restated.py -- the scikit-learn installed
                # here (1.9) normalises and scores differently and would select other features
                from .sklearn_lars_restated import select_features
                nonzero_inds = select_features(self.l1_reg, mask_aug, eyAdj_aug)
            else:
                from sklearn.linear_model import Lasso          # a float: fixed regularisation strength
                nonzero_inds = np.nonzero(Lasso(alpha=self.l1_reg).fit(mask_aug, eyAdj_aug).coef_)[0]
            self.last_nonzero_inds = np.asarray(nonzero_inds)

        if len(nonzero_inds) == 0:
            return np.zeros(self.M)

        delta = self.link.f(self.fx[dim]) - self.link.f(self.fnull[dim])
        eyAdj2 = eyAdj - self.maskMatrix[:, nonzero_inds[-1]] * delta
        etmp = np.transpose(np.transpose(self.maskMatrix[:, nonzero_inds[:-1]]) - self.maskMatrix[:, nonzero_inds[-1]])
        tmp = np.transpose(np.transpose(etmp) * np.transpose(self.kernelWeights))
        etmp_dot = np.dot(np.transpose(tmp), etmp)
        tmp2 = np.linalg.inv(etmp_dot)
        wsol = np.dot(tmp2, np.dot(np.transpose(tmp), eyAdj2))
        phi = np.zeros(self.M)
        phi[nonzero_inds[:-1]] = wsol
        phi[nonzero_inds[-1]] = delta - sum(wsol)
        for i in range(self.M):
            if np.abs(phi[i]) < 1e-10:
                phi[i] = 0
        return phi


class KernelExplainerWrapperOracle(KernelExplainerOracle):
    """Restates ``explainers.kernel_shap.KernelExplainerWrapper`` (kernel_shap.py:217-261) over the oracle:
    seeds the global legacy RNG in the constructor when ``seed`` is given, accepts ``(batch_idx, batch)``
    tuples, exposes ``return_attribute``."""

    def __init__(self, *args, **kwargs):
        if "seed" in kwargs:
            np.random.seed(kwargs.pop("seed"))
        super().__init__(*args, **kwargs)

    def get_explanation(self, X, **kwargs):
        kwargs.pop("silent", None)
        if isinstance(X, tuple):
            batch_idx, batch = X
            return batch_idx, self.shap_values(batch, **kwargs)
        return self.shap_values(X, **kwargs)

    def return_attribute(self, name):
        return self.__getattribute__(name)


# --------------------------------------------------------------------------------------------------
# Independent known answer: exact Shapley values by subset enumeration (definition, not KernelSHAP)
# --------------------------------------------------------------------------------------------------

def exact_shapley(value_fn, M):
    """Shapley values of the set function ``value_fn(mask: uint8[M]) -> float | [C]`` by the subset formula
    phi_k = sum_{T subseteq M\\{k}} |T|!(M-|T|-1)!/M! * (v(T u {k}) - v(T)).  O(2^M): tests only."""
    from math import factorial
    masks = [np.array([(t >> k) & 1 for k in range(M)], dtype=np.uint8) for t in range(2 ** M)]
    vals = [np.asarray(value_fn(m), dtype=np.float64) for m in masks]
    phi = np.zeros((M,) + vals[0].shape)
    for k in range(M):
        for t in range(2 ** M):
            if (t >> k) & 1:
                continue
            size = bin(t).count("1")
            coef = factorial(size) * factorial(M - size - 1) / factorial(M)
            phi[k] += coef * (vals[t | (1 << k)] - vals[t])
    return phi
