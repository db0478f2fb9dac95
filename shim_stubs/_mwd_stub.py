"""Conditional stand-ins for packages the reference's utils/ imports (nltk, librosa, matplotlib).

This directory is OPT-IN: put it on PYTHONPATH only to run the reference's unchanged driver scripts on
a machine that lacks those packages.  Each stub package first looks for the REAL package on the rest of
sys.path and, if it exists, loads that one in its place; only when it is missing does it install an
inert MagicMock module -- and says so on stderr, because evaluation / plotting code running on a stub
produces no meaningful output.  Nothing on the EM hot path uses these packages."""
import importlib.machinery
import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))


class StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return MagicMock(name='%s.%s' % (self.__name__, name))


def _real_spec(name):
    paths = [p for p in sys.path if os.path.abspath(p or '.') != _HERE]
    return importlib.machinery.PathFinder.find_spec(name, paths)


def activate(name, submodules=(), attrs=None, sub_attrs=None):
    """Called from shim_stubs/<name>/__init__.py.  Returns True if the real package took over.
    ``attrs``: plain attributes of the stub package; ``sub_attrs``: {submodule: {attribute: value}}."""
    spec = _real_spec(name)
    if spec is not None and spec.loader is not None:
        real = importlib.util.module_from_spec(spec)
        sys.modules[name] = real            # replaces the stub package being imported
        spec.loader.exec_module(real)
        return True
    sys.stderr.write('[mwd_b200 shim_stubs] %s is not installed: using an inert stub (evaluation / plotting '
                     'through it is a no-op)\n' % name)
    pkg = sys.modules[name]
    for sub in submodules:
        mod = StubModule(name + '.' + sub)
        mod.__path__ = []
        for k, v in ((sub_attrs or {}).get(sub) or {}).items():
            setattr(mod, k, v)
        sys.modules[name + '.' + sub] = mod
        setattr(pkg, sub.split('.')[0], sys.modules[name + '.' + sub.split('.')[0]])
    for k, v in (attrs or {}).items():
        setattr(pkg, k, v)
    return False
