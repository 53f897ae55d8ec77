"""Import alias for the package directory `moog.github.io_b200/`.

The product package lives in a directory whose name (it carries the upstream
repo name) is not a valid Python identifier, so this stub makes it importable
as `moog_b200`: the sub-modules are resolved from that directory, and the
MOOG-compatible `moog` API package that sits inside it is appended to
`sys.path` (appended, so an installed upstream `moog` -- e.g. in parity runs
against the reference -- still wins).
"""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
_PKG_DIR = _os.path.join(_ROOT, 'moog.github.io_b200')
__path__.insert(0, _PKG_DIR)
if _PKG_DIR not in _sys.path:
    _sys.path.append(_PKG_DIR)

with open(_os.path.join(_PKG_DIR, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, '__init__.py'), 'exec'))
