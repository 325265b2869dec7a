"""Result containers and explainer base classes (reference: explainers/interface.py).

Same names and behaviour as the reference (`Explainer`, `FitMixin`, `Explanation`, `NumpyEncoder`, the default
meta/data dictionaries at interface.py:14-37) without the attrs/prettyprinter dependencies, and without
``np.float_`` (interface.py:159), which NumPy 2 removed."""
import abc
import copy
import json
import logging
import pprint
from collections import ChainMap
from typing import Any

import numpy as np

logger = logging.getLogger(__name__)

DEFAULT_META_KERNEL_SHAP = {
    "name": None,
    "type": ["blackbox"],
    "task": None,
    "explanations": ["local", "global"],
    "params": {},
}  # type: dict

DEFAULT_DATA_KERNEL_SHAP = {
    "shap_values": [],
    "expected_value": [],
    "link": "identity",
    "categorical_names": {},
    "feature_names": [],
    "raw": {
        "raw_prediction": None,
        "prediction": None,
        "instances": None,
        "importances": {},
    },
}  # type: dict

DEFAULT_META = {
    "name": None,
    "type": [],
    "explanations": [],
    "params": {},
}  # type: dict


class Explainer(abc.ABC):
    """Base class for explainer algorithms: carries a ``meta`` dict whose keys are also exposed as attributes."""

    def __init__(self, meta: dict = None):
        self.meta = copy.deepcopy(DEFAULT_META) if meta is None else meta
        self.meta["name"] = self.__class__.__name__
        for key, value in self.meta.items():
            setattr(self, key, value)

    def __repr__(self):
        return f"{self.__class__.__name__}(meta={pprint.pformat(self.meta)})"

    @abc.abstractmethod
    def explain(self, X: Any) -> "Explanation":
        pass


class FitMixin(abc.ABC):
    @abc.abstractmethod
    def fit(self, X: Any) -> "Explainer":
        pass


class Explanation:
    """Explanation returned by explainers: ``meta`` and ``data`` dicts, keys exposed as attributes."""

    def __init__(self, meta: dict, data: dict):
        self.meta = meta
        self.data = data
        for key, value in ChainMap(self.meta, self.data).items():
            setattr(self, key, value)

    def __repr__(self):
        return f"Explanation(meta={pprint.pformat(self.meta)}, data={pprint.pformat(self.data)})"

    def to_json(self) -> str:
        """Serialize the explanation data and metadata into a json format."""
        return json.dumps({"meta": self.meta, "data": self.data}, cls=NumpyEncoder)

    @classmethod
    def from_json(cls, jsonrepr) -> "Explanation":
        """Create an Explanation from its json representation."""
        dictrepr = json.loads(jsonrepr)
        try:
            meta = dictrepr["meta"]
            data = dictrepr["data"]
        except KeyError:
            logger.exception("Invalid explanation representation")
            raise
        return cls(meta=meta, data=data)

    def __getitem__(self, item):
        """Deprecated dictionary-style access, kept because the reference keeps it."""
        import warnings
        msg = "The Explanation object is not a dictionary anymore and accessing elements should " \
              "be done via attribute access. Accessing via item will stop working in a future version."
        warnings.warn(msg, DeprecationWarning, stacklevel=2)
        return getattr(self, item)


class NumpyEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.bool_):
            return bool(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return json.JSONEncoder.default(self, obj)
