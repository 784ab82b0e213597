"""Import alias for the product package.

The package directory is ``temporal-span-proposal-network-vidvrd_b200/`` (the name the
build contract fixes); that is not a valid Python identifier, so ``tspn_b200`` extends its
``__path__`` to that directory and executes its ``__init__`` — ``import tspn_b200`` and
``from tspn_b200.model import BaseModel`` resolve to the files there.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "temporal-span-proposal-network-vidvrd_b200")
__path__.insert(0, _PKG_DIR)
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
