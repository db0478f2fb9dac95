"""The native printAlignment writer (csrc/align_json.cu) is byte-identical to what the
reference does: per-pair dicts + ``json.dump(aligns, f, indent=4, sort_keys=True)`` and the
``'%d '`` text file (hmm_dnn/image_phone_hmm_word_discoverer.py:620-648).  Host-only: no GPU."""
import ctypes as C
import json
import math
import os
import struct

import numpy as np
import pytest

from multimodalworddiscovery_b200 import _lib
from multimodalworddiscovery_b200.hmm_dnn._ik_base import write_alignment_files


def _reference_dump(prefix, alis, ics, aps, cas=None, cps=None, cls=None, is_phoneme=True):
    aligns = []
    with open(prefix + '.txt', 'w') as f:
        for i in range(len(alis)):
            n = len(ics[i])
            info = {'index': i, 'image_concepts': [int(c) for c in ics[i]],
                    'alignment': [int(a) for a in alis[i]],
                    'align_probs': np.asarray(aps[i]).reshape(-1, n).tolist(), 'is_phoneme': is_phoneme}
            if cas is not None:
                info['concept_alignment'] = [int(c) for c in cas[i]]
            if cps is not None:
                info['concept_probs'] = np.asarray(cps[i]).tolist()
            if cls is not None:
                info['cluster_probs'] = np.asarray(cls[i]).tolist()
            aligns.append(info)
            for a in alis[i]:
                f.write('%d ' % a)
            f.write('\n\n')
    with open(prefix + '.json', 'w') as f:
        json.dump(aligns, f, indent=4, sort_keys=True)


def _corpus(rng, N, K):
    alis, ics, aps, cas, cps = [], [], [], [], []
    for _ in range(N):
        n, T = int(rng.integers(1, 7)), int(rng.integers(1, 30))
        alis.append(rng.integers(0, n, T).astype(np.int32))
        ics.append(rng.integers(0, K, n).astype(np.int32))
        scale = 10.0 ** rng.integers(-60, 3, (T, n))           # exercise both repr notations
        aps.append((rng.random((T, n)) * scale).ravel())
        cas.append(rng.integers(0, K, T).astype(np.int32))
        cps.append(rng.random((n, K)) * 10.0 ** rng.integers(-20, 1, (n, K)))
    if N:
        aps[0][:3] = [float('nan'), float('inf'), -float('inf')][:len(aps[0][:3])]
    return alis, ics, aps, cas, cps


@pytest.mark.parametrize('variant', ['linear', 'gaussian', 'two-layer', 'audio', 'empty'])
def test_writer_is_byte_identical_to_json_dump(variant, tmp_path):
    rng = np.random.default_rng(3)
    K = 7
    alis, ics, aps, cas, cps = _corpus(rng, 0 if variant == 'empty' else 23, K)
    kw_ref, kw = {}, {}
    if variant in ('linear', 'gaussian', 'empty'):
        kw_ref['cas'], kw['concept_alignment'] = cas, cas
    if variant == 'gaussian':
        kw_ref['cps'], kw['concept_probs'] = cps, cps
    if variant == 'two-layer':
        kw_ref['cls'], kw['cluster_probs'] = cps, cps
    ref, out = str(tmp_path / 'ref'), str(tmp_path / 'out')
    _reference_dump(ref, alis, ics, aps, is_phoneme=(variant != 'audio'), **kw_ref)
    write_alignment_files(out, alis, ics, aps, n_concepts=K, is_phoneme=(variant != 'audio'), **kw)
    for ext in ('.txt', '.json'):
        with open(ref + ext, 'rb') as f1, open(out + ext, 'rb') as f2:
            assert f1.read() == f2.read(), ext


def test_float_repr_matches_python():
    lib = _lib.load()
    buf = C.create_string_buffer(64)
    rng = np.random.default_rng(0)
    vals = [0.0, -0.0, 1.0, 0.1, 1e16, 1e15, 9999999999999998.0, 1e-4, 1e-5, 9.999e-5, 1e22, 5e-324,
            1.7976931348623157e308, 2 / 3, 123456789012345678.0]
    vals += [struct.unpack('d', struct.pack('Q', int(b)))[0] for b in rng.integers(0, 2 ** 63, 20000)]
    vals += list(rng.random(20000) * 10.0 ** rng.integers(-30, 30, 20000))
    for v in vals:
        if math.isnan(v) or math.isinf(v):
            continue
        n = lib.mwd_format_float_repr(float(v), buf, 64)
        assert buf.value.decode() == repr(float(v)) and n == len(repr(float(v)))
