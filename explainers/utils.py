"""Drop-in alias: the reference's ``explainers.utils`` served by the B200 engine."""
from distributedkernelshap_b200.explainers.utils import *  # noqa: F401,F403
from distributedkernelshap_b200.explainers import utils as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("__")]
