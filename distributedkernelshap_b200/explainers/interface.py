"""Result containers and explainer base classes (reference: explainers/interface.py).

Same names and behaviour as the reference (`Explainer`, `FitMixin`, `Explanation`, `NumpyEncoder`, the default
meta/data dictionaries at interface.py:14-37) without the attrs/prettyprinter dependencies, and without
``np.float_`` (interface.py:159), which NumPy 2 removed."""
import abc
import copy
import json
import logging
import pprint
import warnings
from typing import Any

import numpy as np

logger = logging.getLogger(__name__)


def _meta_template(kinds=(), scopes=(), with_task=False):
    out = {"name": None, "type": list(kinds)}
    if with_task:
        out["task"] = None
    out.update(explanations=list(scopes), params={})
    return out


# what an explainer / an explanation starts from (keys as in the reference, interface.py:14-37)
DEFAULT_META = _meta_template()
DEFAULT_META_KERNEL_SHAP = _meta_template(kinds=("blackbox",), scopes=("local", "global"), with_task=True)
DEFAULT_DATA_KERNEL_SHAP = dict(
    shap_values=[], expected_value=[], link="identity", categorical_names={}, feature_names=[],
    raw=dict(raw_prediction=None, prediction=None, instances=None, importances={}),
)


class _KeysAsAttributes:
    """Mixin: every key of the given dictionaries becomes an attribute (meta keys win over data keys)."""

    def _expose(self, *dicts):
        for d in reversed(dicts):
            for key, value in d.items():
                setattr(self, key, value)


class Explainer(_KeysAsAttributes, abc.ABC):
    """Base class for explainer algorithms: carries a ``meta`` dict whose keys are also exposed as attributes."""

    def __init__(self, meta: dict = None):
        self.meta = meta if meta is not None else copy.deepcopy(DEFAULT_META)
        self.meta["name"] = type(self).__name__
        self._expose(self.meta)

    def __repr__(self):
        return "{}(meta={})".format(type(self).__name__, pprint.pformat(self.meta))

    @abc.abstractmethod
    def explain(self, X: Any) -> "Explanation":
        """Explain the rows of ``X``."""


class FitMixin(abc.ABC):
    @abc.abstractmethod
    def fit(self, X: Any) -> "Explainer":
        """Fit the explainer on background data."""


class Explanation(_KeysAsAttributes):
    """Explanation returned by explainers: ``meta`` and ``data`` dicts, keys exposed as attributes."""

    def __init__(self, meta: dict, data: dict):
        self.meta, self.data = meta, data
        self._expose(meta, data)

    def __repr__(self):
        return "Explanation(meta={}, data={})".format(pprint.pformat(self.meta), pprint.pformat(self.data))

    def to_json(self) -> str:
        """The explanation (meta + data) as a JSON string; NumPy scalars and arrays become plain numbers and lists."""
        return json.dumps(dict(meta=self.meta, data=self.data), cls=NumpyEncoder)

    @classmethod
    def from_json(cls, jsonrepr) -> "Explanation":
        """Inverse of ``to_json`` (arrays come back as lists)."""
        parsed = json.loads(jsonrepr)
        missing = [k for k in ("meta", "data") if k not in parsed]
        if missing:
            logger.error("Invalid explanation representation: no %s", missing)
            raise KeyError(missing[0])
        return cls(meta=parsed["meta"], data=parsed["data"])

    def __getitem__(self, item):
        """Dictionary-style access, deprecated in the reference and kept the same way here."""
        warnings.warn("Explanation objects are not dictionaries: read '{}' as an attribute; item access will be removed "
                      "in a future version.".format(item), DeprecationWarning, stacklevel=2)
        return getattr(self, item)


class NumpyEncoder(json.JSONEncoder):
    """JSON encoder that understands NumPy scalars and arrays."""

    _CASTS = ((np.bool_, bool), (np.integer, int), (np.floating, float), (np.ndarray, np.ndarray.tolist))

    def default(self, obj):
        for numpy_type, cast in self._CASTS:
            if isinstance(obj, numpy_type):
                return cast(obj)
        return super().default(obj)
