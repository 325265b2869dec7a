"""The two tuning options of the 129..1024-group path ('wide_gemm' 1|2, 'wide_acache' 0|1) on the configs[3] singleton
fixture (tests/golden/baseline): correctness against the oracle fixture for every combination, bitwise agreement between
the combinations, and the device time of a 2048-instance call (CUDA events around the coalition + solve stage).  NumPy +
ctypes only."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden", "baseline"))

import make_golden_baseline as gen  # noqa: E402
from distributedkernelshap_b200.data import DenseData  # noqa: E402
from distributedkernelshap_b200.engine import GpuKernelExplainer  # noqa: E402

t0 = time.time()
g = np.load(os.path.join(REPO, "tests", "golden", "baseline", gen.FIXTURES["cfg3s"] + ".npz"))
d, nsamples, _ = gen.problem("cfg3s")
np.random.seed(int(g["plan_seed"]))
eng = GpuKernelExplainer(d["predictor"].predict_proba, DenseData(d["background"], d["group_names"], d["groups"]), link="logit")
X4 = np.ascontiguousarray(d["X_explain"])
Xbig = np.ascontiguousarray(np.tile(X4, (512, 1)))
want = g["phi"]
first = None
for gemm, acache in [(1, 0), (2, 0), (1, 1), (2, 1)]:
    eng.set_option("wide_gemm", gemm)
    eng.set_option("wide_acache", acache)
    got = eng.shap_values(X4, nsamples=nsamples, l1_reg=False)
    err = max(float((np.abs(got[c] - want[:, :, c]).max(axis=1) / np.abs(want[:, :, c]).max(axis=1)).max()) for c in range(2))
    if first is None:
        first = got
    same = all(np.array_equal(got[c], first[c]) for c in range(2))
    eng.shap_values(Xbig, nsamples=nsamples, l1_reg=False)
    big = eng.shap_values(Xbig, nsamples=nsamples, l1_reg=False)
    tm = eng.last_timings_ms()
    rep = all(np.array_equal(big[c][:4], got[c]) and np.array_equal(big[c][-4:], got[c]) for c in range(2))
    print(f"wide_gemm={gemm} wide_acache={acache}: max-norm rel err vs oracle fixture {err:.2e}, bitwise equal to the first "
          f"combination {same}, 2048-row call consistent {rep}, coalition+solve {tm['coalitions']:.3f} ms, prepare {tm['prepare']:.3f} ms",
          flush=True)
print(f"total {time.time() - t0:.1f} s")
