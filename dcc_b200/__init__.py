"""Import alias: the product package lives in `dynamic-coverage-control_b200/` (a directory name Python
cannot import); this shim exposes it as `dcc_b200` by pointing the package search path there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dynamic-coverage-control_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
