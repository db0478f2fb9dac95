#!/usr/bin/env python
"""Host->device staging probe: which kind of pinned host memory feeds the GPUs fastest on this box?

Round 1's end-to-end scaling stopped at 2.77x on 8 GPUs because the host->device stream moved 145 GB/s in aggregate
(52 GB/s on one GPU).  The 8-GPU box is a single-NUMA-node VM (nvidia-smi topo: every GPU 'CPU Affinity 0-31,
NUMA 0'), so placement cannot be the lever; this probe measures the staging memory itself:

  pinned     torch pin_memory (cudaHostAlloc, default flags)
  wc         cudaHostAlloc(cudaHostAllocWriteCombined): not snooped in the CPU caches during the DMA read
  thp        2 MB-aligned anonymous memory + madvise(MADV_HUGEPAGE), touched, then cudaHostRegister (fewer IOMMU pages)
  pinned x2  the default buffer copied as two halves on two streams

each solo (rank 0 copies, the others idle) and with every rank copying at once.  Run:
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/scripts/h2d_probe.py
"""
import ctypes
import json
import mmap
import os

import numpy as np
import torch
import torch.distributed as dist

NBYTES = int(os.environ.get('H2D_PROBE_BYTES', 1306 << 20))
REPS = 5


def cudart():
    for name in ('libcudart.so.12', 'libcudart.so'):
        try:
            return ctypes.CDLL(name)
        except OSError:
            pass
    raise SystemExit('libcudart not found')


def as_tensor(ptr, n):
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(ptr)))


def main():
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    rt = cudart()
    dst = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    bufs = {}
    bufs['pinned'] = torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True)
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NBYTES), ctypes.c_uint(0x04))
    if rc == 0:
        bufs['wc'] = as_tensor(p.value, NBYTES)
    try:
        mm = mmap.mmap(-1, NBYTES + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
        aligned = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
        libc = ctypes.CDLL('libc.so.6', use_errno=True)
        adv = libc.madvise(ctypes.c_void_p(aligned), ctypes.c_size_t(NBYTES & ~((2 << 20) - 1)), 14)   # MADV_HUGEPAGE
        t = as_tensor(aligned, NBYTES)
        t.fill_(1)
        rc = rt.cudaHostRegister(ctypes.c_void_p(aligned), ctypes.c_size_t(NBYTES), ctypes.c_uint(0))
        if rc == 0:
            bufs['thp' if adv == 0 else 'registered (madvise failed)'] = t
    except Exception as e:          # informational probe
        if rank == 0:
            print('thp setup failed:', e)
    for t in bufs.values():
        t.fill_(3)
    s2 = torch.cuda.Stream()
    half = NBYTES // 2

    def copy(name):
        if name == 'pinned x2':
            src = bufs['pinned']
            s2.wait_stream(torch.cuda.current_stream())
            dst[:half].copy_(src[:half], non_blocking=True)
            with torch.cuda.stream(s2):
                dst[half:].copy_(src[half:], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
        else:
            dst.copy_(bufs[name], non_blocking=True)

    def timed(name, active):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        if active:
            copy(name)
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if active:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(REPS):
                copy(name)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / REPS
        if world > 1:
            dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    out = {'bytes': NBYTES, 'world': world, 'is_pinned': {k: bool(v.is_pinned()) for k, v in bufs.items()}}
    for name in list(bufs) + ['pinned x2']:
        solo = timed(name, rank == 0)
        allr = timed(name, True)
        out[name] = {'solo_gbs': NBYTES / solo / 1e6, 'all_ranks_gbs_per_rank': NBYTES / allr / 1e6,
                     'all_ranks_gbs_aggregate': world * NBYTES / allr / 1e6}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
