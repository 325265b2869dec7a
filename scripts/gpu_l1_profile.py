"""One l1_reg='auto' call on the BASELINE configs[2] shape (64 features, bg=512, nsamples=4096) for profiling the l1 kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from distributedkernelshap_b200.datasets import dense_tabular
from distributedkernelshap_b200.engine import GpuKernelExplainer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = dense_tabular(n, 64, 512, seed=0)
eng = GpuKernelExplainer(d["predictor"].predict_proba, d["background"], link="logit", seed=0)
eng.shap_values(d["X_explain"][:64], nsamples=4096)
sv = eng.shap_values(d["X_explain"], nsamples=4096)
print("features selected (mean):", float(np.count_nonzero(sv[1], axis=1).mean()), "timings", eng.last_timings_ms())
