"""Importable alias of the ``object-tracking_b200/`` package directory (the hyphen in the task's
directory name is not a valid Python identifier).  ``import object_tracking_b200.engine`` etc. resolve
to the files under ``object-tracking_b200/``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "object-tracking_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
