"""Drop-in for reference 1D-KS-SWAG/nn/ksFiniteDifference.py -- re-exports the libarpde-backed implementation."""
import os as _os, sys as _sys
_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from arpde_b200.nn.integrators import KSIntegrate, GradFilter1D, LaplaceFilter1D, FourthOrderFilter1D  # noqa: E402,F401
