"""B200-native EM hot path for MultimodalWordDiscovery's HMM / HMM-DNN word discoverers.

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/mwd_b200.h)
  _lib.py          ctypes binding of libmwd_b200.so (fails loudly when the library is missing)
  corpus.py        host-side packing: CSR offsets, (n, T) length buckets, rank sharding
  engine.py        device state + one-EM-iteration driver over the C ABI
  hmm_dnn/, hmm/   host-side mirrors of the reference class API (same module and class names)
"""
__version__ = '0.1.0'
