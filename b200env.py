"""Import alias: the package directory name (``sac-gat-her_transportationrl_b200``) contains a
hyphen, so it cannot appear in an ``import`` statement.  ``import b200env`` exposes the same
top-level names."""
import importlib as _importlib

_pkg = _importlib.import_module("sac-gat-her_transportationrl_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
_native = _pkg._native
package = _pkg
__all__ = list(_pkg.__all__) + ["package"]
