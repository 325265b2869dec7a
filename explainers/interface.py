"""Drop-in alias: the reference's ``explainers.interface`` served by the B200 engine."""
from distributedkernelshap_b200.explainers.interface import *  # noqa: F401,F403
from distributedkernelshap_b200.explainers import interface as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("__")]
