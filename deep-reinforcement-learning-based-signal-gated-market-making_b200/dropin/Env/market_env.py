"""Drop-in for the reference's Env/market_env.py: same module path, same names."""
from sgmm_b200.env import FTPEnv  # noqa: F401
