"""Shipped defaults of the 129..1024-group path (wide_gemm 2, wide_acache 1) on three of the shapes of tests/test_gpu_wide.py
(partial tiles, a 2-column second background chunk, two full chunks) against the oracle fed the same plan.  NumPy only."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import test_gpu_wide as T  # noqa: E402

for G, N, S in [(200, 130, 900), (129, 256, 700), (160, 40, 1200)]:
    prob = T.make_problem(seed=G + N, n=5, N=N, widths=(1,) * G)
    eng = T._engine(prob, "logit", seed=5)
    got = eng.shap_values(prob["X"], nsamples=S, l1_reg=False)
    plan = eng.shared_plan(G, S)
    orc = T._oracle(prob, "logit")
    want = np.stack([orc.explain(prob["X"][i:i + 1], plan=(plan.dense(), plan.weights), nsamples=S, l1_reg=False)
                     for i in range(5)])
    err = max(T.rel_err(got[c], want[:, :, c]) for c in range(2))
    eng.set_option("wide_gemm", 1)
    eng.set_option("wide_acache", 0)
    old = eng.shap_values(prob["X"], nsamples=S, l1_reg=False)
    print(f"G={G} N={N} S={S}: rel err vs oracle {err:.2e}; bitwise equal to the first versions: "
          f"{all(np.array_equal(old[c], got[c]) for c in range(2))}", flush=True)
    eng.close()
