"""Parity at the FULL BASELINE.json shapes against committed oracle fixtures (tests/golden/baseline, generator next to
them): all 2560 Adult-shaped instances with a fresh plan per instance, configs[2] (S=4096, N=512, M=64), configs[3]
grouped (1024 columns, N=256, S=8192), configs[4] (M=128, N=512, S=4096, shared plan).  The plans are regenerated with
the oracle's build_plan from the fixture's seeds and their SHA-256 is checked, so the engine is fed exactly the plans the
oracle evaluated.  Two criteria, both stated here:
  * max-norm: |got - want| <= 1e-5 * max_k |want_k| per instance (the bar used since round 1);
  * element-wise: |got - want| <= 1e-5 |want| + ATOL per component, ATOL = 5e-7 in link units (the float32 sigmoid /
    accumulate stage leaves an absolute error of ~1e-7 on every phi, DESIGN.md §5.2: a component of size 1e-3 cannot be
    good to 1e-8), and the fraction of components that also meet the stricter ATOL = 1e-9 is reported."""
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import elementwise_excess, rel_err

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden", "baseline"))
TOL = 1e-5
ATOL_ELEM = 5e-7


def _load(name):
    import make_golden_baseline as gen
    g = np.load(os.path.join(HERE, "golden", "baseline", gen.FIXTURES[name] + ".npz"))
    d, nsamples, _ = gen.problem(name)
    assert gen.data_sha(d) == str(g["data_sha256"]), "datasets.py no longer generates the inputs the fixture was made from"
    assert int(g["nsamples"]) == nsamples
    return gen, g, d, nsamples


def _plans(gen, g, d, nsamples, shared):
    """Per-instance (Z, w) exactly as the generator drew them; checks the SHA-256 stored with the fixture."""
    Ms = g["M"]
    h = hashlib.sha256()
    plans = []
    cache = None
    for i, M in enumerate(Ms):
        if M < 2:
            plans.append(None)
            h.update(b"\0" * 32)
            continue
        if shared and cache is not None:
            Z, w = cache
        else:
            Z, w = gen.instance_plan(int(M), nsamples, i, shared)
            cache = (Z, w)
        hh = hashlib.sha256(gen.pack_bits(Z).tobytes())
        hh.update(w.tobytes())
        h.update(hh.digest())
        plans.append((Z, w))
    assert h.hexdigest() == str(g["plans_sha256"]), "regenerated plans differ from the ones the oracle evaluated"
    return plans


def _engine(d, **kw):
    from distributedkernelshap_b200.data import DenseData
    from distributedkernelshap_b200.engine import GpuKernelExplainer
    return GpuKernelExplainer(d["predictor"].predict_proba, DenseData(d["background"], d["group_names"], d["groups"]),
                              link="logit", **kw)


def _check(got, want, label):
    worst = 0.0
    for c in range(2):
        worst = max(worst, rel_err(got[c], want[:, :, c]))
    frac, mx = elementwise_excess(np.stack(got, axis=-1), want, rtol=TOL, atol=ATOL_ELEM)
    strict_frac, _ = elementwise_excess(np.stack(got, axis=-1), want, rtol=TOL, atol=1e-9)
    print(f"[{label}] max-norm rel err {worst:.2e}; element-wise (1e-5|phi| + {ATOL_ELEM:g}) worst ratio {mx:.3f}, "
          f"violations {frac:.2e}; components outside the strict 1e-5|phi| + 1e-9: {strict_frac:.3%}")
    assert worst < TOL, (label, worst)
    assert frac == 0.0, (label, frac, mx)


@pytest.mark.parametrize("kernel", ["auto", "simt"])
def test_adult_all_2560_instances_fresh_plan_per_instance(kernel):
    gen, g, d, nsamples = _load("adult")
    plans = _plans(gen, g, d, nsamples, False)
    eng = _engine(d, kernel=kernel)
    M, _ = eng.varying(d["X_explain"])
    np.testing.assert_array_equal(M, g["M"])
    got = eng.shap_values(d["X_explain"], nsamples=nsamples, l1_reg=False, plans=plans)
    np.testing.assert_allclose(eng.expected_value, g["expected_value"], rtol=1e-12)
    _check(got, g["phi"], f"adult/{kernel}")
    np.testing.assert_allclose(got[0], -got[1], rtol=0, atol=1e-12)


def test_config2_full_shape_per_instance_plans():
    gen, g, d, nsamples = _load("cfg2")
    plans = _plans(gen, g, d, nsamples, False)
    eng = _engine(d)
    got = eng.shap_values(d["X_explain"], nsamples=nsamples, l1_reg=False, plans=plans)
    _check(got, g["phi"], "cfg2 per-instance plans")


def test_config3_grouped_full_shape_per_instance_plans():
    gen, g, d, nsamples = _load("cfg3")
    plans = _plans(gen, g, d, nsamples, False)
    eng = _engine(d)
    got = eng.shap_values(d["X_explain"], nsamples=nsamples, l1_reg=False, plans=plans)
    _check(got, g["phi"], "cfg3 grouped per-instance plans")


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_full_shape_shared_plan_path(name):
    """The shared-plan fast path at full S and N (4 background chunks of 128, the M = 64 / 128 solves): the engine is
    given the fixture's plan (cfg4) or draws its own (cfg2, cfg3: the oracle is then run here on 3 instances)."""
    gen, g, d, nsamples = _load(name)
    if name == "cfg4":
        np.random.seed(int(g["plan_seed"]))                 # unseeded engine: its M = 128 plan comes from this state
        eng = _engine(d)
        got = eng.shap_values(d["X_explain"], nsamples=nsamples, l1_reg=False)
        Z, w = gen.instance_plan(128, nsamples, 0, True)
        np.random.seed(int(g["plan_seed"]))
        from distributedkernelshap_b200.plan import build_plan
        p = build_plan(128, nsamples)
        np.testing.assert_array_equal(p.dense(), Z)         # product plan builder == oracle's, bit for bit
        _check(got, g["phi"], "cfg4 shared plan")
        return
    from oracle.shap_kernel_oracle import DenseData, KernelExplainerOracle
    eng = _engine(d, seed=7)
    got = eng.shap_values(d["X_explain"], nsamples=nsamples, l1_reg=False)
    Mg = len(d["groups"])
    plan = eng.shared_plan(Mg, nsamples)
    orc = KernelExplainerOracle(d["predictor"].predict_proba, DenseData(d["background"], d["group_names"], d["groups"]),
                                link="logit", chunk_rows=512 if name == "cfg3" else None)
    rows = [i for i in range(len(g["M"])) if g["M"][i] == Mg][:3]
    want = np.stack([orc.explain(d["X_explain"][i:i + 1], plan=(plan.dense(), plan.weights), nsamples=nsamples,
                                 l1_reg=False) for i in rows])
    _check([got[c][rows] for c in range(2)], want, f"{name} shared plan")


def test_config3_singleton_reading_1024_groups_shared_plan():
    """configs[3] read as 1024 singleton groups (M = D = 1024, S = 8192, N = 256): sixteen-word coalition rows through the
    shared-plan coalition kernel, two background chunks of 128, and the float64 projection solve with the 1023 x 1023
    normal matrix factored on the host (csrc/dks_wide.cuh).  The unseeded engine draws its M = 1024 plan from the state the
    fixture's generator seeded, i.e. the very plan the oracle evaluated."""
    gen, g, d, nsamples = _load("cfg3s")
    assert (g["M"] == 1024).all()
    np.random.seed(int(g["plan_seed"]))
    eng = _engine(d)
    got = eng.shap_values(d["X_explain"], nsamples=nsamples, l1_reg=False)
    np.random.seed(int(g["plan_seed"]))
    from distributedkernelshap_b200.plan import build_plan
    p = build_plan(1024, nsamples)
    Z, w = gen.instance_plan(1024, nsamples, 0, True)
    np.testing.assert_array_equal(p.dense(), Z)             # product plan builder == oracle's, bit for bit
    np.testing.assert_array_equal(p.weights, w)
    np.testing.assert_allclose(eng.expected_value, g["expected_value"], rtol=1e-12)
    _check(got, g["phi"], "cfg3 singleton shared plan")
    np.testing.assert_allclose(got[0], -got[1], rtol=0, atol=1e-12)
