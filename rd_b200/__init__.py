"""Importable alias for the product package, which lives in `representation-disentanglement_b200/`
(a hyphenated directory name cannot be imported directly).  `import rd_b200` executes that
package's __init__ with __path__ pointing at the real directory, so `rd_b200.model`,
`rd_b200.lib`, ... resolve to the files there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "representation-disentanglement_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
