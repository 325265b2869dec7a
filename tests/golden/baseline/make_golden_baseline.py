"""Oracle outputs at the FULL BASELINE.json shapes (VERDICT r1 task 1): fixtures under tests/golden/baseline/.

    python tests/golden/baseline/make_golden_baseline.py [adult] [cfg2] [cfg3] [cfg4] [cfg3s]   (default: all; 8 processes)

adult   configs[1]: all 2560 Adult-shaped instances (the bench workload), D=49, 12 groups, bg=100, nsamples=2048,
        a FRESH coalition plan per instance (what shap does): instance i's plan is the one the oracle's build_plan draws
        after ``np.random.seed(PLAN_SEED + i)``.
cfg2    configs[2] shape: 64 ungrouped features, bg=512, nsamples=4096, 8 instances, per-instance plans (same seeding).
cfg3    configs[3], grouped reading: 64 one-hot variables x 16 levels = 1024 columns, bg=256, nsamples=8192, 8 instances,
        per-instance plans; the oracle's masked batch (17 GB in float64) is evaluated 512 coalitions at a time.
cfg4    configs[4] shape: 128 ungrouped features (two-word coalition rows), bg=512, nsamples=4096, 8 instances, ONE plan
        shared by the 8 instances (drawn after ``np.random.seed(PLAN_SEED)``): the engine evaluates two-word rows on
        its shared-plan path.

cfg3s   configs[3], the other reading: every one of the 1024 one-hot columns its own group (M = D = 1024, sixteen-word
        coalition rows, a 1023 x 1023 normal matrix), bg=256, nsamples=8192, uniform level probabilities so that all 1024
        groups vary, 4 instances, ONE shared plan (seeded like cfg4); the masked batch goes through 256 coalitions at a
        time (about two minutes per instance).

The plans are not stored (Adult alone would be 84 MB): the GPU test regenerates them with the oracle's build_plan from the
same seeds and checks their SHA-256 against the one stored here, then feeds them to the engine.  Inputs come from
distributedkernelshap_b200.datasets (seeded); their SHA-256 is stored too.  The oracle is the plain restatement
(vectorised ``run``: same arithmetic as the interpreted loop up to float64 summation order).
"""
import hashlib
import multiprocessing as mp
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, REPO)

PLAN_SEED = 20260921
SHARED_PLAN = ("cfg4", "cfg3s")         # one plan for all instances of the fixture (the engine's shared-plan path)
FIXTURES = {"adult": "adult_2560_s2048", "cfg2": "cfg2_64feat_bg512_s4096", "cfg3": "cfg3_grouped_1024col_bg256_s8192",
            "cfg4": "cfg4_128feat_bg512_s4096", "cfg3s": "cfg3_singleton_1024groups_bg256_s8192"}


def problem(name):
    from distributedkernelshap_b200 import datasets
    if name == "adult":
        d = datasets.adult_like(n_explain=2560, n_background=100, seed=0)
        return d, 2048, None
    if name == "cfg2":
        return datasets.dense_tabular(n=8, n_features=64, n_background=512, seed=0), 4096, None
    if name == "cfg3":
        return datasets.wide_onehot(n=8, n_blocks=64, block_width=16, n_background=256, seed=3), 8192, 512
    if name == "cfg4":
        return datasets.dense_tabular(n=8, n_features=128, n_background=512, seed=4), 4096, None
    if name == "cfg3s":
        return datasets.wide_onehot(n=4, n_blocks=64, block_width=16, n_background=256, seed=3, singleton_groups=True), 8192, 256
    raise ValueError(name)


def data_sha(d):
    h = hashlib.sha256()
    for a in (d["X_explain"], d["background"], d["predictor"].coef_, d["predictor"].intercept_):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def pack_bits(Z):
    """uint64 words [S, W] of a 0/1 matrix [S, M] (bit k of word k // 64 = column k), little-endian like the engine."""
    S, M = Z.shape
    W = (M + 63) // 64
    out = np.zeros((S, W), dtype=np.uint64)
    for k in range(M):
        out[:, k // 64] |= Z[:, k].astype(np.uint64) << np.uint64(k % 64)
    return out


def instance_plan(M, nsamples, i, shared):
    from oracle.shap_kernel_oracle import build_plan, effective_nsamples
    S, _ = effective_nsamples(M, nsamples)
    np.random.seed(PLAN_SEED if shared else PLAN_SEED + i)
    Z, w, _ = build_plan(M, S)
    return Z, w


def _work(args):
    name, lo, hi = args
    from oracle.shap_kernel_oracle import DenseData, KernelExplainerOracle
    d, nsamples, chunk = problem(name)
    orc = KernelExplainerOracle(d["predictor"].predict_proba, DenseData(d["background"], d["group_names"], d["groups"]),
                                link="logit", chunk_rows=chunk)
    G = len(d["groups"])
    phi = np.zeros((hi - lo, G, 2))
    Ms = np.zeros(hi - lo, dtype=np.int32)
    digests = []
    for i in range(lo, hi):
        x = d["X_explain"][i:i + 1]
        M = len(orc.varying_groups(x))
        Ms[i - lo] = M
        if M >= 2:
            Z, w = instance_plan(M, nsamples, i, name in SHARED_PLAN)
            phi[i - lo] = orc.explain(x, plan=(Z, w), nsamples=nsamples, l1_reg=False)
            h = hashlib.sha256(pack_bits(Z).tobytes())
            h.update(w.tobytes())
            digests.append(h.digest())
        else:
            phi[i - lo] = orc.explain(x, nsamples=nsamples, l1_reg=False)
            digests.append(b"\0" * 32)
    return lo, phi, Ms, digests, np.asarray(orc.expected_value)


def make(name, procs=8):
    d, nsamples, _ = problem(name)
    n = d["X_explain"].shape[0]
    step = max(1, min(40, n // procs))
    jobs = [(name, lo, min(n, lo + step)) for lo in range(0, n, step)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        parts = pool.map(_work, jobs, chunksize=1)
    parts.sort(key=lambda p: p[0])
    phi = np.concatenate([p[1] for p in parts])
    Ms = np.concatenate([p[2] for p in parts])
    h = hashlib.sha256()
    for p in parts:
        for dg in p[3]:
            h.update(dg)
    out = os.path.join(HERE, FIXTURES[name] + ".npz")
    np.savez_compressed(out, phi=phi, M=Ms, expected_value=parts[0][4], nsamples=nsamples, plan_seed=PLAN_SEED,
                        plans_sha256=h.hexdigest(), data_sha256=data_sha(d), shared_plan=(name in SHARED_PLAN))
    print(f"{name}: {n} instances in {time.time() - t0:.0f} s -> {out} ({os.path.getsize(out) / 1e3:.0f} kB)")


if __name__ == "__main__":
    names = sys.argv[1:] or ["adult", "cfg2", "cfg3", "cfg4", "cfg3s"]
    for nm in names:
        make(nm)
