"""Drop-in alias of the reference's ``explainers`` package (same module names) over distributedkernelshap_b200."""
