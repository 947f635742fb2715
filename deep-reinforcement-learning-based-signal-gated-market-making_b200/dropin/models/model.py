"""Drop-in for the policy classes of the reference's models/model.py (GateUnits stays the
reference's own: SGU training is host-side and out of scope)."""
from sgmm_b200.policy import AdversaryPolicy, NeuroEvolution, TradingPolicy  # noqa: F401
