"""Import shim: the product package lives in ``instant-nsr-pl_b200/`` (not an importable name);
``import nsr_b200`` resolves its submodules there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'instant-nsr-pl_b200')]

from .lib import lib, NsrError, library_path  # noqa: E402,F401
