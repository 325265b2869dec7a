"""Worker for test_distributed_gloo.py: run under torchrun with the gloo backend (CPU).  Each rank explains its row
block with the ORACLE in the explainer slot (no GPU here) and the blocks are all-gathered; rank 0 checks the result
against a sequential run."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import torch.distributed as dist  # noqa: E402

from conftest import make_problem  # noqa: E402
from distributedkernelshap_b200 import parallel  # noqa: E402
from distributedkernelshap_b200.explainers import kernel_shap  # noqa: E402
from test_host_api import OracleBackedWrapper  # noqa: E402


def main():
    n = int(sys.argv[1])
    parallel.init_from_env(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert parallel.is_distributed() and parallel.world_size() == world
    kernel_shap.KernelExplainerWrapper = OracleBackedWrapper
    prob = make_problem(seed=41, n=n, N=8, widths=(1, 1, 2, 1, 1, 1))
    kw = dict(link="logit", feature_names=prob["group_names"], seed=0)
    fitkw = dict(group_names=prob["group_names"], groups=prob["groups"])
    ks = kernel_shap.KernelShap(prob["clf"].predict_proba, distributed_opts={"n_cpus": world, "batch_size": 3}, **kw)
    ks.fit(prob["bg"], **fitkw)
    assert ks._explainer.spmd and len(ks._explainer.pool) == 1
    got = ks.explain(prob["X"], silent=True, nsamples=62, l1_reg=False).shap_values       # full enumeration (M = 6)
    bounds = parallel.shard_bounds(n, world)
    assert bounds[rank][1] - bounds[rank][0] in (n // world, n // world + 1)
    # uneven all-gather primitive on its own
    lo, hi = bounds[rank]
    local = np.arange(2 * n * 3, dtype=np.float64).reshape(2, n, 3)[:, lo:hi]
    full = parallel.allgather_rows(local, [b[1] - b[0] for b in bounds])
    np.testing.assert_array_equal(full, np.arange(2 * n * 3, dtype=np.float64).reshape(2, n, 3))
    if rank == 0:
        seq = kernel_shap.KernelShap(prob["clf"].predict_proba, **kw).fit(prob["bg"], **fitkw)
        want = seq.explain(prob["X"], silent=True, nsamples=62, l1_reg=False).shap_values
        for g, w in zip(got, want):
            assert g.shape == w.shape == (n, 6)
            np.testing.assert_allclose(g, w, atol=1e-12)
        print("GLOO_OK", n, world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
