"""Per-instance coalition plans drawn on the device (plan_mode='per_instance').

The device sampler must reproduce, bit for bit, the plan upstream's sequential sampling loop builds when its random
numbers come from the same Philox stream: tests/sampler_twin.py wraps that stream in the RandomState interface and the
oracle's restatement of the loop (oracle.build_plan) does the rest.  phi computed from device-drawn plans is then checked
against the oracle fed those very plans."""
import numpy as np
import pytest

from conftest import make_problem, rel_err
from sampler_twin import PhiloxPlanStream
from test_gpu_parity import _engine, _oracle

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _fixture_plans():
    """(M, S, seed, row) -> (zbits, w) from tests/golden/plans/device_plans_philox.npz."""
    import os
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "plans", "device_plans_philox.npz"))
    out, k = {}, 0
    while f"case{k}" in fx:
        out[tuple(int(v) for v in fx[f"case{k}"])] = (fx[f"zbits{k}"], fx[f"w{k}"])
        k += 1
    return out


def _expected_plan(M, nsamples, seed, row):
    from oracle.shap_kernel_oracle import build_plan
    from distributedkernelshap_b200.plan import pack_dense_plan, resolve_nsamples
    S, _ = resolve_nsamples(M, nsamples)
    Z, w, _ = build_plan(M, S, rng=PhiloxPlanStream(seed, row))
    return pack_dense_plan(Z), w


@pytest.mark.parametrize("widths,nsamples", [
    ((1, 1, 3, 2, 1, 2, 1, 1, 2, 1, 1, 1), 300),     # M = 12: sizes 1 enumerated, the rest sampled, few duplicates
    ((1, 2, 1, 3, 1, 1), 40),                          # M = 6: 62 coalitions in all -> duplicates on most draws
    ((1,) * 9, 120),                                   # M = 9 (odd: every size is paired)
    ((1,) * 20, "auto"),                               # M = 20, 2088 rows
    ((1,) * 40, 1500),                                 # M = 40: subsets of up to 20 members (several Philox blocks per draw)
    ((1,) * 64, 4096),                                 # M = 64 (all 64 bits of the word in use), 3968 sampled rows
])
def test_device_plans_equal_the_sequential_loop_on_the_same_stream(widths, nsamples):
    prob = make_problem(seed=21, n=12 if len(widths) < 40 else 4, N=10, widths=widths)
    eng = _engine(prob, kernel="simt", seed=77, plan_mode="per_instance")
    orc = _oracle(prob)
    got = eng.shap_values(prob["X"], nsamples=nsamples, l1_reg=False)
    zb, w = eng.instance_plans()
    Ms, _ = eng.varying(prob["X"])
    assert zb.shape[0] == prob["X"].shape[0]
    for i, M in enumerate(Ms):
        want_z, want_w = _expected_plan(int(M), nsamples, 77, i)
        S = len(want_w)
        np.testing.assert_array_equal(zb[i, :S], want_z, err_msg=f"instance {i}")
        np.testing.assert_allclose(w[i, :S], want_w, rtol=1e-13, atol=0)
        assert np.all(w[i, S:] == 0)
        committed = _fixture_plans().get((int(M), S, 77, i))          # the same plan as a committed golden array
        if committed is not None:
            np.testing.assert_array_equal(zb[i, :S], committed[0])
            np.testing.assert_allclose(w[i, :S], committed[1], rtol=1e-13, atol=0)
        # the regression prepared with the plan (normal matrix from popcounts of the bit-transposed rows, factored in
        # the sampler kernel) must give the oracle's phi for that plan
        k = np.arange(int(M), dtype=np.uint64)
        Z = ((want_z[:, None] >> k[None, :]) & np.uint64(1)).astype(np.uint8)
        phi = orc.explain(prob["X"][i:i + 1], plan=(Z, want_w), nsamples=nsamples, l1_reg=False)
        for c in range(2):
            assert rel_err(got[c][i], phi[:, c]) < TOL


@pytest.mark.parametrize("kernel", ["simt", "tcgen05", "auto"])
def test_phi_from_device_drawn_plans_matches_oracle(kernel):
    prob = make_problem(seed=22, n=20, N=16, widths=(1, 1, 1, 1, 3, 2, 1, 2, 1, 4, 1, 1), constant_groups=(3,))
    from distributedkernelshap_b200.plan import resolve_nsamples
    orc = _oracle(prob)
    eng = _engine(prob, kernel=kernel, seed=5, plan_mode="per_instance")
    got = eng.shap_values(prob["X"], nsamples=400, l1_reg=False)
    zb, w = eng.instance_plans()
    Ms, _ = eng.varying(prob["X"])
    for i in range(prob["X"].shape[0]):
        M = int(Ms[i])
        S = resolve_nsamples(M, 400)[0]
        k = np.arange(M, dtype=np.uint64)
        Z = ((zb[i, :S, None] >> k[None, :]) & np.uint64(1)).astype(np.uint8)
        phi = orc.explain(prob["X"][i:i + 1], plan=(Z, w[i, :S]), nsamples=400, l1_reg=False)
        for c in range(2):
            assert rel_err(got[c][i], phi[:, c]) < TOL
    # plans differ between instances (that is the point of the mode)
    assert not np.array_equal(zb[0], zb[1])


def test_plans_depend_on_the_global_row_only():
    prob = make_problem(seed=23, n=16, N=12, widths=(1,) * 11)
    eng = _engine(prob, seed=9, plan_mode="per_instance")
    full = eng.shap_values(prob["X"], nsamples=200, l1_reg=False)
    part = eng.shap_values(prob["X"][5:11], nsamples=200, l1_reg=False, row_offset=5)
    for c in range(2):
        np.testing.assert_array_equal(part[c], full[c][5:11])
    other = _engine(prob, seed=10, plan_mode="per_instance").shap_values(prob["X"], nsamples=200, l1_reg=False)
    assert not np.allclose(other[0], full[0], rtol=1e-9, atol=0)


def test_per_instance_estimates_agree_with_exact_shapley_on_average():
    """Statistical sanity: with a fresh plan per instance the sampled estimate scatters around the exact value."""
    prob = make_problem(seed=24, n=8, N=8, widths=(1,) * 12)
    eng = _engine(prob, seed=3, plan_mode="per_instance")
    got = eng.shap_values(prob["X"], nsamples=1000, l1_reg=False)
    exact = eng.shap_values(prob["X"], nsamples=10000, l1_reg=False)     # full enumeration (2^12 - 2 rows)
    err = np.abs(got[1] - exact[1]).max() / np.abs(exact[1]).max()
    assert err < 0.05, err
