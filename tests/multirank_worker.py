"""Worker of tests/test_gpu_multirank.py: builds a golden case's class (gaussian or linear), trains two EM
iterations and writes the alignment files -- as ONE process, or as a rank of a torchrun job (NCCL, one GPU per
rank; the class shards the corpus itself).  Usage: multirank_worker.py <case> <out_prefix>"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist
    from helpers import load_ik, make_model
    case, prefix = sys.argv[1], sys.argv[2]
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1:
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    g = load_ik(case)
    tmp = prefix + '_rank%d_tmp' % rank
    os.makedirs(tmp, exist_ok=True)
    m = make_model(tmp, g)
    with contextlib.redirect_stdout(io.StringIO()):
        m.trainUsingEM(2, printStatus=True)
        m.printAlignment(prefix)                     # collective: every rank calls it, rank 0 writes
    if rank == 0:
        key = 'mus' if g['kind'] == 'gaussian' else 'W'
        np.savez(prefix + '_tables.npz', obs=m.obs, post=getattr(m, key),
                 **{'init_%d' % k: v for k, v in m.init.items()}, **{'trans_%d' % k: v for k, v in m.trans.items()})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
