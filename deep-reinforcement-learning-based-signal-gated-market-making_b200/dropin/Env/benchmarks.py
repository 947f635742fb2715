"""Drop-in for the reference's Env/benchmarks.py: same module path, same names."""
from sgmm_b200.benchmarks import FOICPolicy, GLFTPolicy  # noqa: F401
