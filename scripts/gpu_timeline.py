"""Prints the producer/epilogue hand-off latencies of CTA 0 (debug kernel variant) for the bench workload."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributedkernelshap_b200.datasets import adult_like
from distributedkernelshap_b200.data import DenseData
from distributedkernelshap_b200.engine import GpuKernelExplainer

d = adult_like(n_explain=2560)
np.random.seed(0)
eng = GpuKernelExplainer(d["predictor"].predict_proba, DenseData(d["background"], d["group_names"], d["groups"]), link="logit")
tl = eng.debug_timeline(d["X_explain"], nsamples=2048)
names = ["A ready", "acc free", "MMA issued", "epi waits", "acc full", "acc drained"]
g = np.arange(40, 104)
print("tile period (issue->issue):", np.diff(tl[2][40:104]).mean())
print("A ready -> acc free (producer waits for epilogue):", (tl[1][g] - tl[0][g]).mean())
print("acc free -> MMA issued:", (tl[2][g] - tl[1][g]).mean())
print("MMA issued -> acc full seen by epilogue:", (tl[4][g] - tl[2][g]).mean())
print("epilogue waits -> acc full (epilogue idle):", (tl[4][g] - tl[3][g]).mean())
print("acc full -> drained (epilogue busy):", (tl[5][g] - tl[4][g]).mean())
print("drained(g) -> acc free seen for g+4:", (tl[1][g + 4] - tl[5][g]).mean())
print("drained(g) -> epilogue waits (g+4) [y/log tail]:", (tl[3][g + 4] - tl[5][g]).mean())
for k in range(48, 60):
    print(k, " ".join(f"{names[e]}={tl[e][k]:.0f}" for e in range(6)))
