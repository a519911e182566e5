"""Drop-in for reference 2D-Burgers-SWAG/nn/bayesNN.py -- re-exports the libarpde-backed implementation."""
import os as _os, sys as _sys
_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from arpde_b200.nn.bayesnn import BayesNN  # noqa: E402,F401
