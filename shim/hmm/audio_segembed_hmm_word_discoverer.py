# Shadows the reference's hmm/audio_segembed_hmm_word_discoverer.py (see shim/README.md)
from multimodalworddiscovery_b200.hmm.audio_segembed_hmm_word_discoverer import *  # noqa: F401,F403
from multimodalworddiscovery_b200.hmm.audio_segembed_hmm_word_discoverer import np, math, json, time, NULL, DEBUG  # noqa: F401
