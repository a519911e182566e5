"""Import shim: the product package lives in the directory ``ar-pde-cnn_b200/`` (not a valid Python
identifier), so ``import arpde_b200`` maps onto it by extending ``__path__``."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'ar-pde-cnn_b200')
__path__.insert(0, _PKG_DIR)
PACKAGE_DIR = _PKG_DIR

from . import _lib            # noqa: E402  (fails loudly if libarpde.so is missing)
