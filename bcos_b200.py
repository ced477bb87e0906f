"""Import alias: `import bcos_b200` -> the package that lives in `b-cosification_b200/`.

The package directory name is fixed by the project layout and is not a valid Python
identifier, so this module turns itself into that package (sets `__path__` and executes
the package's `__init__.py` in its own namespace).
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "b-cosification_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
