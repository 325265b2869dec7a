"""The product's host-side plan builder (distributedkernelshap_b200/plan.py) against the oracle's: same stream state
in, bit-identical coalition rows and weights out."""
import numpy as np
import pytest

from distributedkernelshap_b200.plan import build_plan, pack_dense_plan, resolve_nsamples
from oracle import shap_kernel_oracle as orc


@pytest.mark.parametrize("M,nsamples", [(2, "auto"), (3, "auto"), (5, "auto"), (7, 50), (12, 2048), (12, "auto"),
                                        (12, 4094), (13, 300), (20, 2048), (31, 1000), (64, 4096), (40, 100)])
def test_plan_matches_oracle_bit_for_bit(M, nsamples):
    S, max_s = resolve_nsamples(M, nsamples)
    assert (S, max_s) == orc.effective_nsamples(M, nsamples)
    np.random.seed(123)
    Z, w, info = orc.build_plan(M, S)
    state_after_oracle = np.random.get_state()[1].copy()
    np.random.seed(123)
    plan = build_plan(M, nsamples)
    np.testing.assert_array_equal(np.random.get_state()[1], state_after_oracle)   # consumed the stream identically
    np.testing.assert_array_equal(plan.dense(), Z)
    np.testing.assert_array_equal(plan.weights, w)
    assert plan.nfixed == info["nfixed"] and plan.num_full_subsets == info["num_full_subsets"]
    np.testing.assert_array_equal(pack_dense_plan(Z), plan.zbits)


def test_explicit_random_state_leaves_global_stream_alone():
    np.random.seed(9)
    before = np.random.get_state()[1].copy()
    a = build_plan(12, 500, rng=np.random.RandomState(4))
    b = build_plan(12, 500, rng=np.random.RandomState(4))
    np.testing.assert_array_equal(np.random.get_state()[1], before)
    np.testing.assert_array_equal(a.zbits, b.zbits)


def test_plan_rejects_unsupported_sizes():
    with pytest.raises(ValueError):
        build_plan(1)
    with pytest.raises(ValueError):
        build_plan(1025)


def test_sampling_info_and_philox_stream_adapter():
    """sampling_info() describes the sampled part of a plan, and the Philox stream adapter used to check the device
    sampler drives the product's and the oracle's plan builders to the same plan."""
    from oracle.shap_kernel_oracle import build_plan as oracle_build_plan
    from sampler_twin import PhiloxPlanStream, philox4x32_10
    from distributedkernelshap_b200.plan import build_plan, pack_dense_plan, sampling_info
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors)
    assert philox4x32_10((0, 0), (0, 0, 0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff, 0xffffffff), (0xffffffff,) * 4) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    for M, ns in [(12, 300), (6, 40), (9, 120), (20, "auto")]:
        plan = build_plan(M, ns, rng=PhiloxPlanStream(123, 7))
        Z, w, info = oracle_build_plan(M, plan.S, rng=PhiloxPlanStream(123, 7))
        np.testing.assert_array_equal(plan.zbits, pack_dense_plan(Z))
        np.testing.assert_allclose(plan.weights, w, rtol=1e-14)
        nfixed, n_full, n_paired, cdf, weight_left = sampling_info(plan)
        assert nfixed == info["nfixed"] and n_full == info["num_full_subsets"]
        assert n_paired == info["num_paired_subset_sizes"]
        assert len(cdf) == info["num_subset_sizes"] - n_full and cdf[-1] == 1.0 and np.all(np.diff(cdf) > 0)
        assert abs(weight_left - info["weight_left"]) < 1e-15
    full = build_plan(5, 1000)
    assert len(sampling_info(full)[3]) == 0


def test_wide_plans_two_words_match_the_oracle():
    from oracle.shap_kernel_oracle import build_plan as oracle_build_plan
    from distributedkernelshap_b200.plan import build_plan, pack_dense_plan
    for M, ns in [(65, 400), (100, 700), (128, "auto")]:
        np.random.seed(4)
        plan = build_plan(M, ns)
        np.random.seed(4)
        Z, w, info = oracle_build_plan(M, plan.S)
        assert plan.zbits.shape == (plan.S, 2) and plan.nfixed == info["nfixed"]
        np.testing.assert_array_equal(plan.zbits, pack_dense_plan(Z))
        np.testing.assert_array_equal(plan.dense(), Z)
        np.testing.assert_allclose(plan.weights, w, rtol=1e-14)
    with pytest.raises(ValueError):
        build_plan(1025, 100)


def test_plans_of_more_than_128_groups_and_their_projection():
    """Sixteen-word rows (129..1024 groups): the product's builder reproduces the oracle's plan bit for bit, and
    plan.projection() is the solve of upstream's constrained WLS: beta = P y - delta d equals inv(E^T W E) E^T W (y - z_L delta)
    and the oracle's own _solve on the same (Z, w, y)."""
    from distributedkernelshap_b200.plan import build_plan, mask_words, pack_dense_plan, projection
    from oracle.shap_kernel_oracle import build_plan as oracle_build_plan
    for M, ns in [(129, 600), (200, 1000), (1024, 4096)]:
        np.random.seed(11)
        plan = build_plan(M, ns)
        np.random.seed(11)
        Z, w, info = oracle_build_plan(M, plan.S)
        assert mask_words(M) == 16 and plan.zbits.shape == (plan.S, 16)
        np.testing.assert_array_equal(plan.dense(), Z)
        np.testing.assert_array_equal(plan.zbits, pack_dense_plan(Z))
        np.testing.assert_array_equal(plan.weights, w)
        assert not plan.zbits[:, (M + 63) // 64:].any()              # unused words stay zero
    np.random.seed(3)
    plan = build_plan(200, 1500)
    PT, d = projection(plan)
    assert PT.shape == (plan.S, 199) and d.shape == (199,)
    Zf = plan.dense().astype(np.float64)
    rng = np.random.default_rng(0)
    y, delta = rng.standard_normal(plan.S), 0.7
    E = Zf[:, :-1] - Zf[:, -1:]
    A = E.T @ (E * plan.weights[:, None])
    want = np.linalg.inv(A) @ (E.T @ (plan.weights * (y - Zf[:, -1] * delta)))
    got = PT.T @ y - delta * d
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * np.abs(want).max())


def test_device_plan_fixture_is_reproduced_by_the_twin():
    """tests/golden/plans/device_plans_philox.npz pins the Philox stream adapter + sequential loop (what the GPU sampler must
    reproduce) independently of the code that generated it."""
    import os
    from oracle.shap_kernel_oracle import build_plan as oracle_build_plan
    from sampler_twin import PhiloxPlanStream
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "plans", "device_plans_philox.npz"))
    k = 0
    while f"case{k}" in fx:
        M, S, seed, row = (int(v) for v in fx[f"case{k}"])
        Z, w, _ = oracle_build_plan(M, S, rng=PhiloxPlanStream(seed, row))
        np.testing.assert_array_equal(pack_dense_plan(Z), fx[f"zbits{k}"])
        np.testing.assert_allclose(w, fx[f"w{k}"], rtol=1e-15)
        plan = build_plan(M, S, rng=PhiloxPlanStream(seed, row))        # the product's builder on the same stream
        np.testing.assert_array_equal(plan.zbits, fx[f"zbits{k}"])
        k += 1
    assert k == 5


def test_l1_tables_reproduce_the_augmented_system():
    """plan.l1_tables: the per-plan Gram matrices / column statistics of upstream's augmented regression, and the moment
    identities the device kernels rely on (csrc/dks_l1.cuh), against the explicit 2S x M system."""
    from distributedkernelshap_b200.plan import l1_tables
    from oracle.sklearn_lars_restated import preprocess
    for M, S, seed in [(16, 400, 1), (70, 600, 2)]:
        np.random.seed(seed)
        plan = build_plan(M, S)
        t = l1_tables(plan)
        Z = plan.dense().astype(float)
        w = plan.weights
        rng = np.random.default_rng(seed)
        y, delta = rng.standard_normal(plan.S), 0.7
        s = Z.sum(1)
        sq = np.sqrt(np.hstack((w * (M - s), w * s)))
        X = (sq * np.vstack((Z, Z - 1)).T).T
        ya = np.hstack((y, y - delta)) * sq
        Xn, yc, _, _, xs = preprocess(X, ya)
        np.testing.assert_allclose(t["gram_raw"], X.T @ X, atol=1e-12)
        np.testing.assert_allclose(t["gram_norm"], Xn.T @ Xn, atol=1e-12)
        np.testing.assert_allclose(t["scale"], xs, atol=1e-13)
        c, u = Z.T @ (w * y), Z.T @ (t["b"] * y)
        T1, Qw, R = (t["b"] * y).sum(), (w * y * y).sum(), ((t["sqa"] + t["sqb"]) * y).sum()
        xty = (M * c - u) - ((T1 - u) - delta * (t["sum_b"] - t["bz"]))
        np.testing.assert_allclose(xty, X.T @ ya, atol=1e-12)
        ybar = (R - delta * t["sum_sqb"]) / t["n_aug"]
        np.testing.assert_allclose((xty - t["colsum"] * ybar) / t["scale"], Xn.T @ yc, atol=1e-12)
        yy = M * Qw - 2 * delta * T1 + delta ** 2 * t["sum_b"] - t["n_aug"] * ybar ** 2
        assert abs(yy - yc @ yc) < 1e-11
        # the restricted WLS of the selected features comes from the w-weighted Gram of the plain rows
        sel = [1, 4, 7, M - 1]
        L = sel[-1]
        E = Z[:, sel[:-1]] - Z[:, [L]]
        A = E.T @ (w[:, None] * E)
        gw = t["gram_w"]
        A2 = np.array([[gw[a, b] - gw[a, L] - gw[b, L] + gw[L, L] for b in sel[:-1]] for a in sel[:-1]])
        np.testing.assert_allclose(A2, A, atol=1e-13)
        rhs = E.T @ (w * (y - Z[:, L] * delta))
        rhs2 = np.array([(c[a] - c[L]) - delta * (gw[a, L] - gw[L, L]) for a in sel[:-1]])
        np.testing.assert_allclose(rhs2, rhs, atol=1e-13)
