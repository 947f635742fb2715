"""Drop-in for the reference's Env/drl_engine.py: same module path, same names."""
from sgmm_b200.engine import DRLEngine, evaluate_individual  # noqa: F401
