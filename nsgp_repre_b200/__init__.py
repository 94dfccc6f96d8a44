"""Import alias for the package directory ``nsgp-repre_b200/`` (a hyphen cannot be
written in an ``import`` statement).  ``import nsgp_repre_b200`` is the one
canonical module name; the code lives next door in ``nsgp-repre_b200/``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "nsgp-repre_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"), globals())
del _os, _f
