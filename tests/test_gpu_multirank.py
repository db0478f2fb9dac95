"""Class-level multi-rank path on real GPUs (needs >= 2 devices: run with `gpurun --gpus 2`): the word-discoverer
classes under torch.distributed shard the corpus, all-gather the packed counts, and printAlignment -- a
collective: decode on every rank, gather to rank 0 -- must terminate and write the same files as a single
process.  Regression test for the Gaussian class, whose `concept_probs` gather used to run on rank 0 only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason='needs 2 GPUs (NCCL refuses two ranks on one device)')
@pytest.mark.parametrize('case', ['mixed_gaussian', 'mixed_linear'])
def test_two_rank_print_alignment_matches_single_process(case, tmp_path):
    one, two = str(tmp_path / 'one'), str(tmp_path / 'two')
    worker = os.path.join(HERE, 'multirank_worker.py')
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
    r1 = subprocess.run([sys.executable, worker, case, one], env=env, capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stderr[-2000:]
    r2 = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                         '--master-addr', '127.0.0.1', '--master-port', '29631', worker, case, two],
                        env=env, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr[-3000:]
    a, b = json.load(open(one + '.json')), json.load(open(two + '.json'))
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x['alignment'] == y['alignment'] and x['image_concepts'] == y['image_concepts']
        assert x['concept_alignment'] == y['concept_alignment']
        np.testing.assert_allclose(np.array(x['align_probs']), np.array(y['align_probs']), rtol=1e-9)
        if 'concept_probs' in x:
            np.testing.assert_allclose(np.array(x['concept_probs']), np.array(y['concept_probs']), rtol=1e-9, atol=1e-300)
    t1, t2 = np.load(one + '_tables.npz'), np.load(two + '_tables.npz')
    for k in t1.files:
        np.testing.assert_allclose(t1[k], t2[k], rtol=1e-10, atol=1e-300)
