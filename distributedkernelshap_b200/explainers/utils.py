"""Helpers of the reference's explainers/utils.py that the hot path touches: ``batch`` (row sharding rule,
utils.py:89-121), ``get_filename`` (result naming, utils.py:67-86), ``Bunch``, ``methdispatch``; plus loaders that
read the reference's pickles when present and otherwise produce the Adult-shaped synthetic stand-in
(there is no network here; utils.py:124-188 downloads)."""
import os
import pickle
from functools import singledispatch, update_wrapper
from typing import Callable, List

import numpy as np
from scipy import sparse

EXPLANATIONS_SET_LOCAL = 'data/adult_processed.pkl'
BACKGROUND_SET_LOCAL = 'data/adult_background.pkl'
MODEL_LOCAL = 'assets/predictor.pkl'


class Bunch(dict):
    """Dictionary whose keys can also be read and written as attributes (the container the reference's data pickles hold)."""

    def __init__(self, **fields):
        dict.__init__(self, fields)

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __dir__(self):
        return list(self.keys())


def methdispatch(func: Callable):
    """``functools.singledispatch`` for methods: the implementation is chosen by the type of the first argument after
    ``self``."""
    registry = singledispatch(func)

    def dispatching_method(self, arg, *rest, **kwargs):
        return registry.dispatch(type(arg))(self, arg, *rest, **kwargs)

    dispatching_method.register = registry.register
    return update_wrapper(dispatching_method, registry)


def get_filename(workers: int, batch_size: int, cpu_fraction: float = 1.0, serve: bool = True):
    """Result-file name under ``results/`` for an experiment configuration (same scheme as the reference)."""
    stem = "ray_replicas_{}_maxbatch_{}" if serve else "ray_workers_{}_bsize_{}"
    return ("results/" + stem + "_actorfr_{}.pkl").format(workers, batch_size, cpu_fraction)


def batch_slices(n_records: int, batch_size: int = None, n_batches: int = 4) -> List[slice]:
    """Row ranges of the mini-batches ``batch`` produces (``np.array_split`` rule): with ``batch_size`` every batch
    but the last has that many rows; otherwise ``n_records % n_batches`` batches get one extra row."""
    if batch_size:
        n_batches = n_records // batch_size + (1 if n_records % batch_size else 0)
        bounds = [min(batch_size * i, n_records) for i in range(n_batches + 1)]
    else:
        base, extra = divmod(n_records, n_batches)
        sizes = [base + 1] * extra + [base] * (n_batches - extra)
        bounds = [0] + list(np.cumsum(sizes))
    return [slice(int(bounds[i]), int(bounds[i + 1])) for i in range(len(bounds) - 1)]


def batch(X: np.ndarray, batch_size: int = None, n_batches: int = 4) -> List[np.ndarray]:
    """Splits the input into mini-batches (sparse inputs are densified first, as in the reference)."""
    if isinstance(X, sparse.spmatrix) or sparse.issparse(X):
        X = X.toarray()
    return [X[s] for s in batch_slices(X.shape[0], batch_size, n_batches)]


def load_model(path: str = MODEL_LOCAL):
    """The reference's pickled scikit-learn model if it is on disk, else the synthetic Adult-shaped classifier."""
    if os.path.exists(path):
        with open(path, "rb") as f:
            return pickle.load(f)
    from distributedkernelshap_b200.datasets import adult_like
    return adult_like()["predictor"]


def load_data():
    """The reference's Adult pickles if present (``data/adult_*.pkl``), else the synthetic stand-in with the same
    dictionary structure (``data['all']['groups']``, ``data['background']['X']['preprocessed']``, ...)."""
    sources = {"background": BACKGROUND_SET_LOCAL, "all": EXPLANATIONS_SET_LOCAL}
    if all(os.path.exists(path) for path in sources.values()):
        loaded = {}
        for key, path in sources.items():
            with open(path, "rb") as handle:
                loaded[key] = pickle.load(handle)
        return loaded
    from distributedkernelshap_b200.datasets import adult_like
    return adult_like()["data"]
