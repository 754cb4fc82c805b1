"""Importable alias of the ``probabilistic-multiplanar-unet_b200/`` package directory (a hyphen
is not a valid Python identifier).  ``import pmu_b200`` executes that directory's __init__ with
this package's __path__ pointing at it, so ``pmu_b200.ops``, ``pmu_b200.model`` ... resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "probabilistic-multiplanar-unet_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
