"""GPU counterpart of the reference's ``benchmarks/ray_pool.py`` (same CLI: -b/--batch, -w/--workers, -benchmark,
-n/--nruns; same result files ``results/ray_workers_{w}_bsize_{b}_actorfr_1.0.pkl`` holding ``t_elapsed``).

Workers are GPUs (one CUDA context each) instead of ray CPU actors; ``-w -1`` runs without a DistributedExplainer.
Data/model: the reference's pickles when present under data/ and assets/, else the Adult-shaped synthetic stand-in.
Pass ``--nsamples`` to override shap's default (2 * 12 + 2048 = 2072, what the reference benchmark runs).

    python benchmarks/gpu_pool.py -w 1 -b 10
    python -m torch.distributed.run --nproc-per-node 8 benchmarks/gpu_pool.py -w 8 -b 320
"""
import argparse
import logging
import os
import pickle
import sys
from timeit import default_timer as timer

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from explainers.kernel_shap import KernelShap  # noqa: E402
from explainers.utils import get_filename, load_data, load_model  # noqa: E402

logging.basicConfig(level=logging.INFO)


def fit_kernel_shap_explainer(clf, data, distributed_opts=None):
    group_names, groups = data['all']['group_names'], data['all']['groups']
    explainer = KernelShap(clf.predict_proba, link='logit', feature_names=group_names, distributed_opts=distributed_opts, seed=0)
    explainer.fit(data['background']['X']['preprocessed'], group_names=group_names, groups=groups)
    return explainer


def run_explainer(explainer, X_explain, distributed_opts, nruns, explain_kwargs):
    os.makedirs('./results', exist_ok=True)
    result = {'t_elapsed': []}
    for run in range(nruns):
        t_start = timer()
        explainer.explain(X_explain, silent=True, **explain_kwargs)
        t_elapsed = timer() - t_start
        logging.info(f"run {run}: {t_elapsed:.6f} s ({X_explain.shape[0] / t_elapsed:.1f} instances/s)")
        result['t_elapsed'].append(t_elapsed)
        with open(get_filename(distributed_opts['n_cpus'], distributed_opts['batch_size'], serve=False), 'wb') as f:
            pickle.dump(result, f)


def main(args):
    from distributedkernelshap_b200 import parallel
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        parallel.init_from_env()
    nruns = args.nruns if args.benchmark else 1
    data = load_data()
    predictor = load_model('assets/predictor.pkl')
    X_explain = data['all']['X']['processed']['test'].toarray()
    explain_kwargs = {'l1_reg': False}
    if args.nsamples:
        explain_kwargs['nsamples'] = args.nsamples
    if args.workers == -1:
        opts = {'batch_size': None, 'n_cpus': None, 'actor_cpu_fraction': 1.0}
        run_explainer(fit_kernel_shap_explainer(predictor, data, opts), X_explain, opts, nruns, explain_kwargs)
        return
    workers_range = range(1, args.workers + 1) if args.benchmark == 1 else range(args.workers, args.workers + 1)
    for workers in workers_range:
        for batch_size in [int(b) for b in args.batch]:
            opts = {'batch_size': batch_size, 'n_cpus': workers, 'actor_cpu_fraction': 1.0}
            run_explainer(fit_kernel_shap_explainer(predictor, data, opts), X_explain, opts, nruns, explain_kwargs)


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument("-b", "--batch", nargs='+', required=True, help="Mini-batch sizes sent to a worker at a time.")
    parser.add_argument("-w", "--workers", default=-1, type=int, help="Number of GPU workers; -1 = no DistributedExplainer.")
    parser.add_argument("-benchmark", default=0, type=int, help="1: sweep workers in range(1, workers + 1).")
    parser.add_argument("-n", "--nruns", default=5, type=int, help="Runs per setting in benchmark mode.")
    parser.add_argument("--nsamples", default=0, type=int, help="Override shap's default nsamples (0 = 'auto').")
    main(parser.parse_args())
