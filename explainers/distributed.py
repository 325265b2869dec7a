"""Drop-in alias: the reference's ``explainers.distributed`` served by the B200 engine."""
from distributedkernelshap_b200.explainers.distributed import *  # noqa: F401,F403
from distributedkernelshap_b200.explainers import distributed as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("__")]
