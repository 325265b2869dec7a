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
        build_plan(65)
