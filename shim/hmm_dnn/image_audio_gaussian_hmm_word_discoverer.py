# Shadows the reference's hmm_dnn/image_audio_gaussian_hmm_word_discoverer.py (see shim/README.md)
from multimodalworddiscovery_b200.hmm_dnn.image_audio_gaussian_hmm_word_discoverer import *  # noqa: F401,F403
from multimodalworddiscovery_b200.hmm_dnn.image_audio_gaussian_hmm_word_discoverer import np, math, json, time, logsumexp, random, deepcopy, KMeans, NULL, DEBUG, EPS  # noqa: F401
