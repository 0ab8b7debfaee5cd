"""Importable alias for the package that lives in `l-step_b200/` (a hyphen is not a valid
module name). `import lstep_b200` executes `l-step_b200/__init__.py` with `__path__` pointing
at that directory, so `lstep_b200.sampler` etc. resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "l-step_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
