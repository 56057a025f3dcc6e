"""Alias for the package whose directory name is not a Python identifier.

``import ljmd_b200 as lj`` == importlib.import_module(
    "molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200").
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

PACKAGE_NAME = "molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200"
_pkg = importlib.import_module(PACKAGE_NAME)
sys.modules[__name__] = _pkg
