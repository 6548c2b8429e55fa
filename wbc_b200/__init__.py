"""Importable alias of the package directory ``mech5845m-wbc-for-legged-manipulator_b200`` (hyphens are
not valid in a Python module name).  ``import wbc_b200`` executes that package's ``__init__`` with
``__path__`` pointing at it, so ``wbc_b200.robot_model`` etc. resolve to the real files."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "mech5845m-wbc-for-legged-manipulator_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
