"""Mirror of the reference's ``explainers`` package (same module and symbol names) over the CUDA engine."""
