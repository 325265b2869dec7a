"""Drop-in alias: the reference's ``explainers.wrappers`` served by the B200 engine."""
from distributedkernelshap_b200.explainers.wrappers import *  # noqa: F401,F403
from distributedkernelshap_b200.explainers import wrappers as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("__")]
