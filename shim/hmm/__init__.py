"""Shadows the reference's hmm/ namespace package (see shim/README.md); modules not mirrored here
fall through to the reference's own hmm/ directory found on sys.path."""
import os
import sys

for _p in list(sys.path):
    _d = os.path.join(_p or '.', 'hmm')
    if os.path.isdir(_d) and os.path.abspath(_d) != os.path.dirname(os.path.abspath(__file__)) and _d not in __path__:
        __path__.append(_d)
