"""Serving back-ends (reference: explainers/wrappers.py).  Same classes and call signatures; ``ray.serve`` is not
required: a back-end is just a callable object holding a fitted ``KernelShap``.  The batch back-end coalesces the
requests of a batch into ONE engine call (the reference loops over them, wrappers.py:83-86)."""
import logging
from typing import Any, Dict, List

import numpy as np

from distributedkernelshap_b200.explainers.kernel_shap import KernelShap

try:  # pragma: no cover - ray is optional
    from ray import serve
    accept_batch = serve.accept_batch
except Exception:  # ray missing (or a ray version without accept_batch)
    def accept_batch(fn):
        fn._serve_accept_batch = True
        return fn


class KernelShapModel:
    """Backend class for serving explanations."""

    def __init__(self, predictor, background_data: np.ndarray, constructor_kwargs: Dict[str, Any],
                 fit_kwargs: Dict[str, Any]):
        """``predictor``: model to be explained (its ``predict_proba`` is used when present, else ``predict``);
        ``constructor_kwargs`` / ``fit_kwargs``: forwarded to ``KernelShap`` / ``KernelShap.fit``."""
        if not hasattr(predictor, "predict_proba"):
            logging.warning("Predictor does not have predict_proba attribute, defaulting to predict")
            predict_fcn = predictor.predict
        else:
            predict_fcn = predictor.predict_proba
        self.explainer = KernelShap(predict_fcn, **constructor_kwargs)
        self.explainer.fit(background_data, **fit_kwargs)

    def __call__(self, flask_request) -> str:
        """Explains the instance in the ``array`` field of a json request; returns the explanation as json."""
        instance = np.array(flask_request.json["array"])
        explanations = self.explainer.explain(instance, silent=True)
        return explanations.to_json()


class BatchKernelShapModel(KernelShapModel):
    """Extends KernelShapModel to batches of requests."""

    @accept_batch
    def __call__(self, flask_requests: List) -> List[str]:
        """One json explanation per request.  All instances are explained by a single GPU call."""
        instances = [np.atleast_2d(np.array(request.json["array"])) for request in flask_requests]
        if not instances:
            return []
        rows = [inst.shape[0] for inst in instances]
        stacked = np.concatenate(instances, axis=0)
        explainer = self.explainer
        shap_values = explainer._explainer.get_explanation(stacked, silent=True)
        if isinstance(shap_values, np.ndarray):
            shap_values = [shap_values]
        expected_value = explainer._explainer.expected_value
        if isinstance(expected_value, float):
            expected_value = [expected_value]
        out, start = [], 0
        for inst, n_rows in zip(instances, rows):
            sl = slice(start, start + n_rows)
            start += n_rows
            explanation = explainer.build_explanation(inst, [sv[sl] for sv in shap_values], expected_value)
            out.append(explanation.to_json())
        return out
