"""Importable alias of the product package.

The package directory is named after the reference repository
(``deep-reinforcement-learning-based-signal-gated-market-making_b200/``), which is not a valid
Python identifier; this alias extends its ``__path__`` to that directory so that
``import sgmm_b200`` / ``from sgmm_b200.engine import DRLEngine`` resolve to it.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "deep-reinforcement-learning-based-signal-gated-market-making_b200")
__path__.insert(0, _REAL)
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
