"""Stand-in `matplotlib` package exposing only `path`, `transforms`, `pyplot`.

TEST INFRASTRUCTURE ONLY: lets the unmodified reference import in the build
container (matplotlib is not installed, no wheel, no network).  See path.py.
"""
__version__ = '0.0-moog-oracle-shim'
from . import path, transforms  # noqa: F401,E402
