"""CPU oracle for the EM hot path of MultimodalWordDiscovery's HMM / HMM-DNN word discoverers.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multimodalworddiscovery_b200/`` (the product) may
import this package.  The only permitted callers are ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as the checker
or the timed CPU baseline, never as the thing shipped.

Each function restates, in plain NumPy float64, the algorithm of the reference file:line it
cites.  Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the unmodified reference classes imported from
``/root/reference`` in the build container; the vectors are committed under ``tests/golden/``
together with the generating script ``tests/golden/make_golden.py``.
"""
