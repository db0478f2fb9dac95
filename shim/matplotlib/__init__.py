"""Inert stand-in so the reference's utils/ (which imports this package, absent from the image)
can be imported by the unchanged driver scripts.  Nothing on the EM hot path uses it."""
import sys
import types
from unittest.mock import MagicMock


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return MagicMock(name='%s.%s' % (self.__name__, name))


def _install(name):
    mod = _Stub(name)
    mod.__path__ = []
    sys.modules[name] = mod
    return mod


for _sub in ('pyplot', 'colors', 'cm'):
    _install(__name__ + '.' + _sub)


def use(*a, **k):
    return None


def __getattr__(name):
    return MagicMock(name='matplotlib.' + name)
