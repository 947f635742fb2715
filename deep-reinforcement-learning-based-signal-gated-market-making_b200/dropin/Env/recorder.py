"""Drop-in for the reference's Env/recorder.py: same module path, same names."""
from sgmm_b200.recorder import StrategyRecorder  # noqa: F401
