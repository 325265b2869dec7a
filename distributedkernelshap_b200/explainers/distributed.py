"""Data-parallel explanation over GPUs (reference: explainers/distributed.py).

The reference's only parallelism strategy is data parallelism over instances: a ``ray.util.ActorPool`` of
single-CPU actors, each holding a full explainer replica, fed mini-batches through the plasma store and
re-ordered afterwards (distributed.py:85-179).  Same class name, constructor and ``get_explanation`` signature
here, but the workers are CUDA contexts:

* inside a ``torch.distributed`` job (``torchrun``, one process per GPU): every rank is one worker, takes the
  contiguous row block ``np.array_split`` would give it (utils.py:120), explains it on its GPU, and the shap
  values are exchanged with a single all-gather (NCCL on GPUs, gloo in CPU tests) -- the only collective;
* in a single process: ``n_actors`` engines on the visible GPUs, driven by one host thread each (the C-ABI call
  releases the GIL), consuming mini-batches as they become free -- the ActorPool pattern without ray.

Every worker seeds the same legacy NumPy stream (kernel_shap.py:779), so with shared coalition plans the result is
independent of the number of workers.
"""
import logging
import os
from concurrent.futures import ThreadPoolExecutor, as_completed
from functools import partial
from typing import Any, Callable, Dict, List, Optional, Union

import numpy as np
from scipy import sparse

from distributedkernelshap_b200 import parallel
from distributedkernelshap_b200.explainers.utils import batch, batch_slices

logger = logging.getLogger(__name__)


def kernel_shap_target_fn(actor: Any, instances: tuple, kwargs: Optional[Dict] = None) -> Callable:
    """What a free worker runs on a ``(batch_index, batch)`` tuple (distributed.py:11-34, minus ``.remote``)."""
    if kwargs is None:
        kwargs = {}
    return actor.get_explanation(instances, **kwargs)


def kernel_shap_postprocess_fn(ordered_result: List[Union[np.ndarray, List[np.ndarray]]]) \
        -> List[Union[np.ndarray, List[np.ndarray]]]:
    """Concatenates per-batch results: arrays for scalar-output predictors, one array per class otherwise."""
    if isinstance(ordered_result[0], np.ndarray):
        return np.concatenate(ordered_result, axis=0)
    n_classes = len(ordered_result[0])
    return [np.concatenate([res[c] for res in ordered_result], axis=0) for c in range(n_classes)]


def invert_permutation(p: list):
    """``s`` with ``s[p[i]] = i`` for a permutation ``p`` of ``0..len(p)-1``."""
    p = np.asarray(p)
    s = np.empty_like(p)
    s[p] = np.arange(len(p))
    return s


class DistributedExplainer:
    """Orchestrates the explanation of a batch of instances over several GPU workers."""

    def __init__(self, distributed_opts, explainer_type, init_args, init_kwargs):
        self.n_jobs = distributed_opts['n_cpus']
        self.n_actors = int(distributed_opts['n_cpus'] // distributed_opts['actor_cpu_fraction'])
        self.actor_cpu_frac = distributed_opts['actor_cpu_fraction']
        self.batch_size = distributed_opts['batch_size']
        self.algorithm = distributed_opts['algorithm']
        self.target_fn = globals()[f"{distributed_opts['algorithm']}_target_fn"]
        try:
            self.post_process_fcn = globals()[f"{distributed_opts['algorithm']}_postprocess_fn"]
        except KeyError:
            self.post_process_fcn = None

        self.explainer = explainer_type
        self.explainer_args = init_args
        self.explainer_kwargs = init_kwargs

        self.spmd = parallel.is_distributed()
        self.pool = self.create_parallel_pool()

    def __getattr__(self, item):
        """State shared by all workers (``expected_value``, ``vector_out``, ...) is read from the first one."""
        if item in ("pool", "explainer", "spmd"):
            raise AttributeError(item)
        return self.pool[0].return_attribute(item)

    def create_parallel_pool(self):
        """One explainer replica per worker.  Under torchrun: this rank's GPU only.  Otherwise ``n_actors`` replicas
        spread round-robin over the visible GPUs."""
        if self.spmd:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            return [self.explainer(*self.explainer_args, device=device, **self.explainer_kwargs)]
        n_devices = max(parallel.visible_gpus(), 1)
        n_workers = max(1, min(self.n_actors, n_devices))
        if self.n_actors > n_devices:
            logger.info("%d workers requested but %d GPU(s) visible: using %d", self.n_actors, n_devices, n_workers)
        return [self.explainer(*self.explainer_args, device=k % n_devices, **self.explainer_kwargs)
                for k in range(n_workers)]

    def get_explanation(self, X: np.ndarray, **kwargs) -> np.ndarray:
        """Explains the rows of ``X`` in parallel; ``kwargs`` go to the explainer's ``shap_values``.  Every mini-batch
        is sent with the index of its first row (``row_offset``) so that device-drawn per-instance plans depend on the
        row, not on how the rows were split over workers."""
        kwargs = dict(kwargs or {})
        if self.spmd:
            return self._get_explanation_spmd(X, kwargs)

        if sparse.issparse(X):
            X = X.toarray()
        slices = [sl for sl in batch_slices(X.shape[0], self.batch_size, self.n_jobs) if sl.stop > sl.start]
        items = [((idx, X[sl]), sl.start) for idx, sl in enumerate(slices)]      # (fewer rows than workers: no empty batches)

        def call(actor, item, start):
            return self.target_fn(actor, item, kwargs={**kwargs, "row_offset": start})

        if len(self.pool) == 1:
            unordered = [call(self.pool[0], item, start) for item, start in items]
            return self.order_result(unordered)

        # ActorPool.map_unordered: a free worker takes the next mini-batch
        import queue
        free = queue.Queue()
        for actor in self.pool:
            free.put(actor)

        def run(item, start):
            actor = free.get()
            try:
                return call(actor, item, start)
            finally:
                free.put(actor)

        with ThreadPoolExecutor(max_workers=len(self.pool)) as ex:
            futures = [ex.submit(run, item, start) for item, start in items]
            unordered = [f.result() for f in as_completed(futures)]
        return self.order_result(unordered)

    def _get_explanation_spmd(self, X, kwargs):
        """One process per GPU: explain this rank's row block, then all-gather the shap values."""
        rank, world = parallel.rank(), parallel.world_size()
        if sparse.issparse(X):
            X = X.toarray()
        n = X.shape[0]
        blocks = batch_slices(n, None, world)          # np.array_split rule
        counts = [b.stop - b.start for b in blocks]
        mine = X[blocks[rank]]
        engine = self.pool[0]
        # GPUs + NCCL: the block stays on the device between the solve and the collective (one D2H of the gathered rows)
        if (self.batch_size is None and "plans" not in kwargs and hasattr(engine, "explain_block_to_device")
                and parallel._dist().get_backend() == "nccl"):
            local_dev = engine.explain_block_to_device(mine, row_offset=blocks[rank].start, **kwargs)
            gathered = parallel.allgather_rows_device(local_dev, counts)
            if not engine.return_attribute("vector_out"):
                return gathered[0]
            return [gathered[c] for c in range(gathered.shape[0])]
        if mine.shape[0] > 0:
            local_slices = batch_slices(mine.shape[0], self.batch_size, 1)
            results = [self.target_fn(self.pool[0], (idx, mine[sl]),
                                      kwargs={**kwargs, "row_offset": blocks[rank].start + sl.start})
                       for idx, sl in enumerate(local_slices)]
            local = self.order_result(results)
        else:
            local = None
        vector_out = self.pool[0].return_attribute("vector_out")
        G = self.pool[0].return_attribute("data").groups_size
        if local is None:
            stacked = np.zeros((self.pool[0].return_attribute("D"), 0, G))
        else:
            stacked = np.stack(local, axis=0) if isinstance(local, list) else local[None]
        gathered = parallel.allgather_rows(stacked, counts)          # [C, n, G] on every rank
        if not vector_out:
            return gathered[0]
        return [gathered[c] for c in range(gathered.shape[0])]

    def order_result(self, unordered_result: List[tuple]) -> np.ndarray:
        """Re-orders ``(batch_index, result)`` tuples by batch index and concatenates them."""
        result_order, results = list(zip(*[(idx, res) for idx, res in unordered_result]))
        orig_order = invert_permutation(list(result_order))
        ordered_result = [results[idx] for idx in orig_order]
        if self.post_process_fcn is not None:
            return self.post_process_fcn(ordered_result)
        return ordered_result
