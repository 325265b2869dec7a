"""One shared-plan explain call at the configs[3] singleton shape (1024 one-hot columns = 1024 groups, bg=256,
nsamples=8192) through the host API, for profiling the sixteen-word coalition kernel and the wide-solve kernels.  NumPy +
ctypes only (no torch import: the profile sessions are short)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributedkernelshap_b200.data import DenseData
from distributedkernelshap_b200.datasets import wide_onehot
from distributedkernelshap_b200.engine import GpuKernelExplainer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
t0 = time.time()
d = wide_onehot(n, 64, 16, 256, seed=0, singleton_groups=True)
eng = GpuKernelExplainer(d["predictor"].predict_proba, DenseData(d["background"], d["group_names"], d["groups"]),
                         link="logit", seed=0)
X = np.ascontiguousarray(d["X_explain"])
eng.shap_values(X[:64], nsamples=8192, l1_reg=False)            # plan, Dm table and projection built and uploaded
t1 = time.time()
sv = eng.shap_values(X, nsamples=8192, l1_reg=False)
print("setup %.1f s, call %.3f s, timings %s, launches %d, additivity residual %.2e" % (
    t1 - t0, time.time() - t1, eng.last_timings_ms(), eng.kernel_launches(),
    float(np.abs(sv[1].sum(1) - (eng.link_predictions()[:, 1] - eng.expected_value[1])).max())))
