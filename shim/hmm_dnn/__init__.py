"""Shadows the reference's hmm_dnn/ namespace package (see shim/README.md).

Modules present in this directory are the CUDA-backed mirrors.  Every OTHER hmm_dnn module the
drivers import (image_phone_hmm_dnn_word_discoverer, image_audio_*_word_discoverer, ...) falls
through to the reference's own file: the reference's hmm_dnn/ directories found on sys.path are
appended to this package's search path, after this directory."""
import os
import sys

for _p in list(sys.path):
    _d = os.path.join(_p or '.', 'hmm_dnn')
    if os.path.isdir(_d) and os.path.abspath(_d) != os.path.dirname(os.path.abspath(__file__)) and _d not in __path__:
        __path__.append(_d)
