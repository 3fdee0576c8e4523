"""Importable alias for the package directory ``qc-quantum-linear-systems_b200/``.

The product package lives in the directory the build contract names
(``qc-quantum-linear-systems_b200``), which is not a valid Python identifier.  This
alias points ``__path__`` at that directory and executes its ``__init__`` so that every
module exists exactly once, as ``qls_b200.<module>``.
"""
import os as _os

_real = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "qc-quantum-linear-systems_b200",
)
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r", encoding="utf-8") as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _fh
