"""Generates the golden fixtures in this directory.

The reference ships no golden vectors and its arithmetic (shap==0.35.0) cannot be imported offline, so the fixtures
hold (a) outputs of the oracle restatement on small seeded problems, including the per-instance sampled plans it
drew from the legacy MT19937 stream, and (b) for the fully-enumerated cases, exact Shapley values from the
brute-force subset formula (an answer that does not depend on the restatement).  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from conftest import make_problem  # noqa: E402
from oracle.shap_kernel_oracle import DenseData, KernelExplainerOracle, exact_shapley  # noqa: E402


def make(name, seed, widths, n, N, link, kappa, nsamples, full, weights):
    prob = make_problem(seed=seed, n=n, N=N, widths=widths, kappa=kappa, weights=weights)
    wts = prob["weights"] if prob["weights"] is not None else np.ones(N)
    dd = DenseData(prob["bg"], prob["group_names"], prob["groups"], wts)
    orc = KernelExplainerOracle(prob["clf"].predict_proba, dd, link=link, record_plans=True)
    np.random.seed(seed)
    phis = np.stack([orc.explain(prob["X"][i:i + 1], nsamples=nsamples, l1_reg=False) for i in range(n)])
    Z = np.stack([p[1] for p in orc.plans])
    w = np.stack([p[2] for p in orc.plans])
    exact = np.zeros_like(phis)
    if full:
        wb = wts / wts.sum()
        for i in range(n):
            x = prob["X"][i]

            def value(mask):
                rows = prob["bg"].copy()
                for k, on in enumerate(mask):
                    if on:
                        rows[:, prob["groups"][k]] = x[prob["groups"][k]]
                ey = (prob["clf"].predict_proba(rows) * wb[:, None]).sum(0)
                return orc.link.f(ey) - orc.link.f(orc.fnull)
            exact[i] = exact_shapley(value, len(prob["groups"]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), X=prob["X"], bg=prob["bg"], weights=wts,
                        groups=np.array([np.array(g) for g in prob["groups"]], dtype=object),
                        coef=prob["clf"].coef_, intercept=prob["clf"].intercept_, multi_class=prob["clf"].multi_class,
                        link=link, nsamples=nsamples, full=full, Z=Z, w=w, phi=phis, phi_exact=exact,
                        expected_value=orc.expected_value)


def make_device_plan_fixture(name, cases):
    """Plans of the device-side sampler (csrc/dks_sampler.cuh) as the sequential loop builds them from the Philox stream
    (tests/sampler_twin.py): the GPU test compares the kernel's output with these committed arrays as well."""
    from oracle.shap_kernel_oracle import build_plan, effective_nsamples
    from sampler_twin import PhiloxPlanStream
    from distributedkernelshap_b200.plan import pack_dense_plan
    out = {}
    for k, (M, nsamples, seed, row) in enumerate(cases):
        S, _ = effective_nsamples(M, nsamples)
        Z, w, _ = build_plan(M, S, rng=PhiloxPlanStream(seed, row))
        out[f"case{k}"] = np.array([M, S, seed, row], dtype=np.int64)
        out[f"zbits{k}"] = pack_dense_plan(Z)
        out[f"w{k}"] = w
    np.savez_compressed(os.path.join(HERE, "plans", name + ".npz"), **out)


if __name__ == "__main__":
    make_device_plan_fixture("device_plans_philox", [(12, 300, 77, 0), (12, 300, 77, 5), (6, 40, 77, 3), (9, 120, 77, 11),
                                                     (20, 2088, 77, 2)])
    make("full_logit_m7", 101, (1, 2, 1, 1, 3, 1, 2), n=6, N=9, link="logit", kappa=2.0, nsamples=126, full=True, weights=True)
    make("full_identity_m6", 102, (1, 1, 2, 1, 4, 1), n=5, N=8, link="identity", kappa=1.0, nsamples=62, full=True, weights=False)
    make("sampled_logit_m12", 103, (1, 1, 1, 1, 3, 2, 1, 2, 1, 4, 1, 1), n=8, N=20, link="logit", kappa=2.0, nsamples=400,
         full=False, weights=False)
    make("sampled_logit_m16", 104, (1,) * 10 + (2, 3, 1, 2, 1, 1), n=6, N=14, link="logit", kappa=1.0, nsamples=600,
         full=False, weights=True)
    print("golden fixtures written to", HERE)
