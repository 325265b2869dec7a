"""Host-side end-to-end breakdown for the bench workload (2560 Adult-shaped instances)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributedkernelshap_b200 import _cabi
from distributedkernelshap_b200.datasets import adult_like
from distributedkernelshap_b200.explainers.kernel_shap import KernelShap

d = adult_like(n_explain=2560)
ks = KernelShap(d["predictor"].predict_proba, link="logit", feature_names=d["group_names"], seed=0)
ks.fit(d["data"]["background"]["X"]["preprocessed"], group_names=d["group_names"], groups=d["groups"])
eng = ks._explainer
X = np.ascontiguousarray(d["X_explain"])
eng.get_explanation(X, nsamples=2048, l1_reg=False)
lib = eng.lib
Xp = torch.from_numpy(X).pin_memory(); php = torch.empty((2, 2560, 12), dtype=torch.float64).pin_memory()
phn = np.zeros((2, 2560, 12))

def timeit(fn, reps=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6

print("C call, pinned X -> pinned phi      : %.1f us" % timeit(lambda: lib.dks_explain_host(eng._ctx, C.c_void_p(Xp.data_ptr()), 2560, C.c_void_p(php.data_ptr()), None, None, 0)))
print("C call, pinned X -> pageable phi    : %.1f us" % timeit(lambda: lib.dks_explain_host(eng._ctx, C.c_void_p(Xp.data_ptr()), 2560, _cabi.ptr(phn), None, None, 0)))
print("C call, pageable X -> pageable phi  : %.1f us" % timeit(lambda: lib.dks_explain_host(eng._ctx, _cabi.ptr(X), 2560, _cabi.ptr(phn), None, None, 0)))
print("engine.get_explanation(pinned numpy): %.1f us" % timeit(lambda: eng.get_explanation(Xp.numpy(), nsamples=2048, l1_reg=False, silent=True)))
print("engine.get_explanation(numpy)       : %.1f us" % timeit(lambda: eng.get_explanation(X, nsamples=2048, l1_reg=False, silent=True)))
print("KernelShap.explain (full API)       : %.1f us" % timeit(lambda: ks.explain(X, nsamples=2048, l1_reg=False, silent=True), reps=50))
print("device timings of last call (ms):", eng.last_timings_ms())

import cProfile, pstats, io
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    ks.explain(X, nsamples=2048, l1_reg=False, silent=True)
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22)
print(st.getvalue()[:4500])
