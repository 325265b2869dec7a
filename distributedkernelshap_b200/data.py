"""Host-side data containers the reference obtains from ``shap.common`` (not vendored in the reference):
``DenseData`` (kernel_shap.py:594, :616, :646, :665) and the link objects behind ``convert_to_link``
(kernel_shap.py:15, :949).  Same constructor shapes and field names, so ``KernelShap`` code reads alike."""
import numpy as np


class Data:
    """Marker base class (``shap.common.Data``); ``KernelShap._check_inputs`` tests isinstance against it."""


class DenseData(Data):
    """Background matrix with optional feature groups and per-row weights.

    ``DenseData(data, group_names, groups=None, weights=None)``: ``groups`` defaults to one group per column,
    ``weights`` to uniform; weights are normalised to sum to one.  If the group sizes add up to the number of
    rows rather than columns the matrix is taken as transposed (``KernelShap`` warns about this case at
    kernel_shap.py:443-449)."""

    def __init__(self, data, group_names, *args):
        data = np.asarray(data)
        groups = args[0] if len(args) > 0 and args[0] is not None else None
        weights = args[1] if len(args) > 1 and args[1] is not None else None
        self.groups = [np.asarray(g, dtype=np.int64) for g in groups] if groups is not None \
            else [np.array([i]) for i in range(len(group_names))]
        covered = sum(len(g) for g in self.groups)
        self.transposed = covered != data.shape[1]
        n_rows = data.shape[1] if self.transposed else data.shape[0]
        if covered != (data.shape[0] if self.transposed else data.shape[1]):
            raise AssertionError("# of names must match data matrix!")
        self.weights = np.ones(n_rows) if weights is None else np.asarray(weights, dtype=np.float64)
        if len(self.weights) != n_rows:
            raise AssertionError("# weights must match data matrix!")
        self.weights = self.weights / np.sum(self.weights)
        self.group_names = list(group_names)
        self.data = data
        self.groups_size = len(self.groups)


class DenseDataWithIndex(DenseData):
    """``shap.common.DenseDataWithIndex`` (kernel_shap.py:638-644): keeps a DataFrame index alongside the data."""

    def __init__(self, data, group_names, index, index_name, *args):
        DenseData.__init__(self, data, group_names, *args)
        self.index_value = index
        self.index_name = index_name


def convert_to_data(val):
    """Arrays / DataFrames / Series -> ``DenseData`` with singleton groups (``shap.common.convert_to_data``)."""
    if isinstance(val, Data):
        return val
    try:
        import pandas as pd
        if isinstance(val, pd.DataFrame):
            return DenseData(val.values, list(val.columns))
        if isinstance(val, pd.Series):
            return DenseData(val.values.reshape(1, len(val)), list(val.index))
    except ImportError:  # pragma: no cover
        pass
    try:
        from scipy import sparse
        if sparse.issparse(val):
            val = val.toarray()
    except ImportError:  # pragma: no cover
        pass
    arr = np.asarray(val)
    if arr.ndim == 1:
        arr = arr.reshape(1, -1)
    if arr.ndim != 2:
        raise TypeError("Unknown type passed as data object: " + str(type(val)))
    return DenseData(arr, [str(i) for i in range(arr.shape[1])])


class IdentityLink:
    def __str__(self):
        return "identity"

    @staticmethod
    def f(x):
        return x

    @staticmethod
    def finv(x):
        return x


class LogitLink:
    def __str__(self):
        return "logit"

    @staticmethod
    def f(x):
        return np.log(x / (1 - x))

    @staticmethod
    def finv(x):
        return 1 / (1 + np.exp(-x))


def convert_to_link(val):
    if isinstance(val, (IdentityLink, LogitLink)):
        return val
    if val == "identity":
        return IdentityLink()
    if val == "logit":
        return LogitLink()
    raise ValueError("Passed link object must be 'identity' or 'logit'")


def sample(X, nsamples=100, random_state=0):
    """``shap.sample`` (used by kernel_shap.py:535): ``nsamples`` rows drawn with sklearn's ``resample`` (with
    replacement, fixed random state); the data itself when it already has at most ``nsamples`` rows."""
    if nsamples >= X.shape[0]:
        return X
    from sklearn.utils import resample
    return resample(X, n_samples=nsamples, random_state=random_state)


def kmeans(X, k, round_values=True):
    """``shap.kmeans`` (used by kernel_shap.py:542): k-means centroids, each coordinate snapped to the nearest value
    seen in the data, weighted by cluster size, as a ``DenseData``."""
    from sklearn.cluster import KMeans
    group_names = [str(i) for i in range(X.shape[1])]
    if hasattr(X, "columns"):
        group_names = list(X.columns)
        X = X.values
    km = KMeans(n_clusters=k, random_state=0, n_init=10).fit(X)
    centers = km.cluster_centers_.copy()
    if round_values:
        for i in range(k):
            for j in range(X.shape[1]):
                centers[i, j] = X[np.argmin(np.abs(X[:, j] - centers[i, j])), j]
    return DenseData(centers, group_names, None, 1.0 * np.bincount(km.labels_, minlength=k))
