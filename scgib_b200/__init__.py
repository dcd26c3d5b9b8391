"""Import alias: the product package lives in ``s-cgib_b200/`` (not a valid Python identifier),
so ``import scgib_b200`` resolves its sub-modules there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "s-cgib_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
