"""Import alias: the package directory is named after the reference repository
(`semi-supervised-adaptive-distillation_b200`), which is not a valid Python identifier.  This
module makes it importable as `sad_b200` by pointing its search path at that directory."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "semi-supervised-adaptive-distillation_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
