"""Opt-in stand-in for `librosa` (see shim_stubs/_mwd_stub.py): defers to the real package when it is
installed, otherwise installs an inert stub and warns on stderr."""
import os
import sys
from unittest.mock import MagicMock

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
try:
    import _mwd_stub
finally:
    sys.path.pop(0)

_REAL = _mwd_stub.activate(__name__, ('feature', 'core', 'display'))


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    return MagicMock(name='librosa.' + name)
