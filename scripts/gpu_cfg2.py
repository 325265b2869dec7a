"""BASELINE.json configs[2] (64 features, bg=512, nsamples=4096) and configs[3] in its grouped reading (64 one-hot
variables x 16 levels = 1024 columns, bg=256, nsamples=8192) at a bounded number of instances: instances/s through the
engine (host X in, host phi out) and device-only.  Not the headline bench line -- coverage data points.
usage: gpu_cfg2.py <instances> [shared|per_instance] [cfg2|cfg3|cfg4]"""
import json
import sys
import time

import numpy as np
import torch

from distributedkernelshap_b200.datasets import dense_tabular, wide_onehot
from distributedkernelshap_b200.engine import GpuKernelExplainer
from distributedkernelshap_b200.data import DenseData

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = sys.argv[2] if len(sys.argv) > 2 else "shared"
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
NS = 8192 if cfg == "cfg3" else 4096
wl = {"cfg2": lambda: dense_tabular(n, 64, 512, seed=0), "cfg3": lambda: wide_onehot(n, 64, 16, 256, seed=0),
      "cfg4": lambda: dense_tabular(n, 128, 512, seed=0)}[cfg]()
dd = DenseData(wl["background"], wl["group_names"], wl["groups"])
eng = GpuKernelExplainer(wl["predictor"].predict_proba, dd, link="logit", seed=0, plan_mode=mode)
X = wl["X_explain"]
t0 = time.perf_counter()
sv = eng.shap_values(X[:256], nsamples=NS, l1_reg=False)
t_first = time.perf_counter() - t0
t0 = time.perf_counter()
sv = eng.shap_values(X, nsamples=NS, l1_reg=False)
t_host = time.perf_counter() - t0
X_dev = torch.from_numpy(X).cuda()
phi = torch.empty((2, n, len(wl["groups"])), dtype=torch.float64, device="cuda")
eng.set_stream(torch.cuda.current_stream().cuda_stream)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
eng.explain_device(X_dev.data_ptr(), n, phi.data_ptr(), nsamples=NS)
ev1.record()
torch.cuda.synchronize()
eng.check_status()
fx = wl["predictor"].predict_proba(X[:64])
add = np.abs(sv[1][:64].sum(1) - (np.log(fx[:, 1] / fx[:, 0]) - eng.expected_value[1])).max()
print(json.dumps({"config": {"cfg2": "cfg2: D=M=64, bg=512, nsamples=4096",
                             "cfg3": "cfg3 grouped: D=1024 (64 one-hot variables x 16 levels), bg=256, nsamples=8192",
                             "cfg4": "cfg4 (one GPU's share): D=M=128, bg=512, nsamples=4096"}[cfg], "plan_mode": mode, "instances": n, "kernel": eng.kernel,
                  "first_call_s": t_first, "host_api_inst_per_s": n / t_host,
                  "device_inst_per_s": n / (ev0.elapsed_time(ev1) / 1e3), "device_ms": ev0.elapsed_time(ev1),
                  "additivity_max_abs_err": float(add), "timings_ms": eng.last_timings_ms()}))
