"""Drop-in alias: the reference's ``explainers.kernel_shap`` served by the B200 engine."""
from distributedkernelshap_b200.explainers.kernel_shap import *  # noqa: F401,F403
from distributedkernelshap_b200.explainers import kernel_shap as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("__")]
