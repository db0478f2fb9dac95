#!/usr/bin/env python
"""Benchmark of the EM hot path: caption-pairs/s per EM iteration of the image-phone HMM word
discoverer (BASELINE.json metric), on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full EM iteration (image posterior, forward/backward + expected counts, concept
posteriors, count reduction [+ all-reduce], posterior gradient, M-step) over the whole synthetic
corpus: `--pairs` MSCOCO-shaped caption-image pairs (default 1 000 000, BASELINE.json configs[4]),
sorted into length buckets and dealt over the ranks (strong scaling: total work fixed).

  value : pairs/s with the corpus resident in HBM (device timed, CUDA events, max over ranks)
  e2e   : pairs/s through the host-facing API with every step copying the shard + parameters
          from pinned host memory and reading log-likelihood + updated parameter tables back
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md "Measurement"

`--impl reference` times the reference algorithm's CPU implementation (the NumPy oracle port of
the reference classes -- the reference itself is Python and is not present on the GPU box) on all
host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_CONCEPTS, P_PHONES, D_FEAT = 65, 49, 512
T_MEAN, T_STD = 50.0, 10.0
# empirical region-count distribution of the Flickr30k captions (n mean 3.05, max 8)
FLICKR_N_PMF = [0.12, 0.28, 0.27, 0.17, 0.09, 0.04, 0.02, 0.01]


def apply_variant(variant):
    """Workload shapes: coco5 / coco10 = C1/C5 (K=65, P=49, T~N(50,10)); flickr = C3
    (K=100 concepts, P=69 phones, T~N(49,13), n ~ Flickr30k's empirical 1..8)."""
    global K_CONCEPTS, P_PHONES, T_MEAN, T_STD
    if variant == 'flickr':
        K_CONCEPTS, P_PHONES, T_MEAN, T_STD = 100, 69, 49.0, 13.0
    else:
        K_CONCEPTS, P_PHONES, T_MEAN, T_STD = 65, 49, 50.0, 10.0
METRIC = 'em_caption_pairs_per_sec'
UNIT = 'pairs/s'


# ----------------------------------------------------------------------------------------------
# synthetic MSCOCO-shaped workload (C1/C5)
# ----------------------------------------------------------------------------------------------
def region_counts(n_pairs, variant, gen, torch, dev):
    if variant == 'coco5':
        return torch.full((n_pairs,), 5, dtype=torch.int64, device=dev)
    if variant == 'flickr':
        pmf = torch.tensor(FLICKR_N_PMF, dtype=torch.float64, device=dev)
        return torch.multinomial(pmf, n_pairs, replacement=True, generator=gen) + 1
    return torch.randint(1, 11, (n_pairs,), generator=gen, device=dev)


def make_shard(torch, dev, n_pairs_global, rank, world, variant, seed=20261018):
    """Generate the rank's shard directly on the device, sorted by (n, T)."""
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    N = n_pairs_global
    # lengths for the WHOLE corpus (cheap), then the round-robin shard of the sorted order
    T = torch.clamp(torch.round(T_MEAN + T_STD * torch.randn(N, generator=gen, device=dev)), 15, 125).long()
    n = region_counts(N, variant, gen, torch, dev)
    key = n * 1000 + T
    order = torch.argsort(key, stable=True)
    mine = order[rank::world]
    n_m, T_m = n[mine], T[mine]
    lens = sorted(int(v) for v in torch.unique(n).tolist())
    region_off = torch.zeros(len(mine) + 1, dtype=torch.int32, device=dev)
    region_off[1:] = torch.cumsum(n_m, 0)
    phone_off = torch.zeros(len(mine) + 1, dtype=torch.int32, device=dev)
    phone_off[1:] = torch.cumsum(T_m, 0)
    R, Tt = int(region_off[-1]), int(phone_off[-1])
    gen2 = torch.Generator(device=dev)
    gen2.manual_seed(seed + 1 + rank)
    cgen = torch.Generator(device=dev)
    cgen.manual_seed(seed + 7)
    centroids = 10.0 * torch.randn(K_CONCEPTS, D_FEAT, generator=cgen, device=dev)
    concept = torch.randint(0, K_CONCEPTS, (R,), generator=gen2, device=dev)
    feats = torch.empty((R, D_FEAT), dtype=torch.float32, device=dev)
    chunk = 1 << 20
    for lo in range(0, R, chunk):
        hi = min(R, lo + chunk)
        feats[lo:hi] = centroids[concept[lo:hi]] + torch.randn(hi - lo, D_FEAT, generator=gen2, device=dev)
    pw = 1.0 / torch.arange(1, P_PHONES + 1, device=dev, dtype=torch.float64) ** 1.2
    phones = torch.multinomial(pw / pw.sum(), Tt, replacement=True, generator=gen2).to(torch.int32)
    W = 0.01 * torch.randn(K_CONCEPTS, D_FEAT + 1, generator=cgen, device=dev, dtype=torch.float64)
    # gaussian class (configs[1]): visual anchors = the cluster centroids + noise
    mus = centroids.double() + 0.5 * torch.randn(K_CONCEPTS, D_FEAT, generator=cgen, device=dev, dtype=torch.float64)
    return dict(region_off=region_off, phone_off=phone_off, feats=feats, phones=phones, lens=lens,
                W=W, mus=mus, n_local=len(mine))


def flops_per_pair(T_mean, n_mean, n3_mean):
    """Algorithmic float64 flop per pair: posterior GEMM + recursion and counts
    + restricted concept chains + gradient GEMM."""
    K, D = K_CONCEPTS, D_FEAT
    return 2 * n_mean * K * (D + 1) + 30 * T_mean * n_mean * K + 2 * T_mean * K * n3_mean + 2 * n_mean * K * (D + 1)


def bytes_per_pair(T_mean, n_mean):
    """SURVEY 8(d) algorithmic bytes per pair: phones + features + conceptCounts + concept
    alignment argmax + log-likelihood."""
    return 4 * T_mean + 4 * n_mean * D_FEAT + 4 * n_mean * K_CONCEPTS + 4 * T_mean + 8


# ----------------------------------------------------------------------------------------------
# CPU arm: the NumPy oracle port on all host cores
# ----------------------------------------------------------------------------------------------
def cpu_sample_numpy(n_pairs, variant, seed=20261018):
    """Bounded sample of the same workload distribution, generated with NumPy (CPU arm only)."""
    rng = np.random.default_rng(seed)
    crng = np.random.default_rng(20261018 + 7)
    centroids = 10.0 * crng.standard_normal((K_CONCEPTS, D_FEAT))
    W = 0.01 * crng.standard_normal((K_CONCEPTS, D_FEAT + 1))
    pw = 1.0 / np.arange(1, P_PHONES + 1) ** 1.2
    pw /= pw.sum()
    feats, phones = [], []
    for _ in range(n_pairs):
        T = int(np.clip(round(rng.normal(T_MEAN, T_STD)), 15, 125))
        if variant == 'coco5':
            n = 5
        elif variant == 'flickr':
            n = int(rng.choice(len(FLICKR_N_PMF), p=FLICKR_N_PMF)) + 1
        else:
            n = int(rng.integers(1, 11))
        v = centroids[rng.integers(0, K_CONCEPTS, n)] + rng.standard_normal((n, D_FEAT))
        feats.append(v.astype(np.float32).astype(np.float64))
        phones.append(rng.choice(P_PHONES, size=T, p=pw))
    return feats, phones, W


def _cpu_proc(idx, n_pairs, variant, n_rounds, barrier, out_q):
    """One CPU worker: builds ITS slice of the sample (untimed), then runs `n_rounds` E-steps of
    the NumPy oracle over it, each round released by the shared barrier."""
    from oracle import image_phone_hmm as orc
    apply_variant(variant)
    feats, phones, W = cpu_sample_numpy(n_pairs, variant, seed=20261018 + 1000 + idx)
    params = orc.initial_params(feats, K_CONCEPTS, P_PHONES, 'linear', W=W, lr=0.1)
    params['toeplitz'] = variant != 'coco5'
    for _ in range(n_rounds):
        barrier.wait()
        orc.em_iteration(feats, phones, params, 'linear')
        barrier.wait()
    out_q.put(idx)


def run_cpu_arm(n_sample, variant, steps, warmup, cores):
    """Times `steps` EM iterations of the oracle over an `n_sample`-pair sample with `cores`
    single-BLAS-thread processes.  Returns (pairs/s, seconds per step, pairs actually run)."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    per = max(1, n_sample // cores)
    saved = {k: os.environ.get(k) for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS')}
    for k in saved:
        os.environ[k] = '1'
    barrier = ctx.Barrier(cores + 1)
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_proc, args=(i, per, variant, warmup + steps, barrier, q), daemon=True)
             for i in range(cores)]
    for p in procs:
        p.start()
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    times = []
    for s in range(warmup + steps):
        barrier.wait()            # all workers ready -> start
        t0 = time.perf_counter()
        barrier.wait()            # all workers done
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    for p in procs:
        p.join(timeout=30)
    total = per * cores
    return total * len(times) / sum(times), float(np.mean(times)), total


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.cpu_pairs or 2048 * cores
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    pps, sec, n_sample = run_cpu_arm(n_sample, args.variant, steps, warmup, cores)
    sample = '%d pairs of the %s workload per step, %d processes x 1 BLAS thread' % (n_sample, args.variant, cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': pps, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, n_sample),
        'cpu_baseline': {'value': pps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': pps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args, n_pairs):
    return {'workload': 'image-phone HMM EM iteration, synthetic %s shape (%s): %d pairs, '
                        'T~clip(N(%g,%g),15,125), K=%d concepts, P=%d phones, D=%d res34-like features'
                        % ('Flickr30k' if args.variant == 'flickr' else 'MSCOCO', args.variant, n_pairs, T_MEAN, T_STD,
                           K_CONCEPTS, P_PHONES, D_FEAT),
            'pairs': n_pairs, 'variant': args.variant, 'class': 'ImagePhoneGaussianHMMWordDiscoverer' if args.model == 'gaussian' else 'ImagePhoneHMMWordDiscoverer',
            'l2_policy': ('inputs (%.2f GB features) exceed the 126 MB L2' if n_pairs * 3 * D_FEAT * 4 > 126e6 else
                          'inputs (%.2f GB features) fit in the 126 MB L2 and are NOT flushed (non-default size)')
                         % (n_pairs * (3 if args.variant == 'flickr' else 5) * D_FEAT * 4 / 1e9),
            'parallelism': 'pairs sharded over %d GPU(s), one packed fp64 count all-reduce per iteration' % args.gpus}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix='mwd_clocks_', suffix='.csv')
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                   f[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out['sm_mhz'] = float(np.median(sm))
            out['sm_max_mhz'] = float(max(mx))
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from multimodalworddiscovery_b200.corpus import pack_sorted_arrays
    from multimodalworddiscovery_b200.engine import IKEngine

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit('--gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d ...)'
                         % (args.gpus, args.gpus))

    # CPU baseline first (rank 0, N=1 only), before this process touches CUDA
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_sample = args.cpu_pairs or 2048 * cores
        pps, sec, n_sample = run_cpu_arm(n_sample, args.variant, 1, 0, cores)
        cpu_baseline = {'value': pps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': '%d pairs of the %s workload, one EM iteration of the NumPy oracle, '
                                  '%d processes x 1 BLAS thread (%.1f s)' % (n_sample, args.variant, cores, sec)}

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    sh = make_shard(torch, dev, args.pairs, rank, world, args.variant)
    # pinned host copy of the shard (the e2e leg copies from here every step)
    host = {}
    for k in ('region_off', 'phone_off', 'feats', 'phones'):
        h = torch.empty(sh[k].shape, dtype=sh[k].dtype, pin_memory=True)
        h.copy_(sh[k])
        host[k] = h
    torch.cuda.synchronize()
    pk = pack_sorted_arrays(host['region_off'].numpy(), host['phone_off'].numpy(), host['feats'].numpy(),
                            host['phones'].numpy(), lens=sh['lens'], n_pairs_global=args.pairs)
    del sh['feats'], sh['phones']
    torch.cuda.empty_cache()
    gaussian = args.model == 'gaussian'
    width = float(D_FEAT) if gaussian else 1.0       # RBF width of the order of |v - mu|^2 (unit-variance noise)
    eng = IKEngine(pk, K_CONCEPTS, P_PHONES, gaussian=gaussian, device=dev, keep_concept_counts_a=False,
                   mixed_precision=args.mixed)
    if args.chunks <= 0:
        # enough chunks to overlap the PCIe copy with the kernels, not so many that a small corpus
        # drowns in launches (1 M pairs: 10.4 GB -> 16 chunks; MSCOCO-2k: 21 MB -> 1 chunk)
        shard_bytes = sum(host[k].numel() * host[k].element_size() for k in host)
        args.chunks = int(min(16, max(1, shard_bytes // (640 << 20))))

    # initializeModel(): uniform init/trans/obs, injected W  (parameter snapshot restored every step)
    init = {m: np.ones(m) / m for m in pk.lens}
    trans = {m: np.ones((m, m)) / m for m in pk.lens}
    obs = np.ones((K_CONCEPTS, P_PHONES)) / P_PHONES
    eng.set_params(init, trans, obs, (sh['mus'] if gaussian else sh['W']).cpu().numpy())
    snap = [t.clone() for t in (eng.init_t, eng.trans_t, eng.obsT, eng.post)]
    h_params = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in snap]
    h_out = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in snap]
    h_ll = torch.empty((1,), dtype=torch.float64, pin_memory=True)
    lr, mom = 0.1, 0.0

    def restore():
        for dst, src in zip((eng.init_t, eng.trans_t, eng.obsT, eng.post), snap):
            dst.copy_(src)

    def step_resident(timers=None):
        restore()
        return eng.em_iteration(lr, mom, width, with_cA=False, timers=timers)

    def step_e2e():
        # host -> device: the shard (streamed in chunks that overlap the kernels) and the
        # parameters; device -> host: LL + updated tables
        for dst, src in zip((eng.init_t, eng.trans_t, eng.obsT, eng.post), h_params):
            dst.copy_(src, non_blocking=True)
        ll = eng.em_iteration_streamed(host, lr, mom, width, n_chunks=args.chunks)
        h_ll.copy_(ll.reshape(1), non_blocking=True)
        for dst, src in zip(h_out, (eng.init_t, eng.trans_t, eng.obsT, eng.post)):
            dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(h_ll[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- resident leg ------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    timers = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ll = None
    for _ in range(args.steps):
        ll = step_resident(timers)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clock_info = clocks.stop() if clocks else None
    ms_step = ms_total / args.steps
    value = args.pairs / (ms_step * 1e-3)
    avg_ll = float(ll) / args.pairs
    kern_ms = {}
    for name, a, b in timers:
        kern_ms[name] = kern_ms.get(name, 0.0) + a.elapsed_time(b)
    kern_ms = {k: v / args.steps for k, v in kern_ms.items()}

    # ---- e2e leg -----------------------------------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    h2d = sum(host[k].numel() * host[k].element_size() for k in host) + sum(
        t.numel() * t.element_size() for t in h_params)
    d2h = 8 + sum(t.numel() * t.element_size() for t in h_out)
    if world > 1:
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        h2d, d2h = int(tot[0]), int(tot[1])

    # ---- align leg ("plus align-pairs/s separately"): batched Viterbi + cluster of the whole
    # shard under the current parameters (posterior GEMM + K6), alignments written to HBM
    restore()
    for _ in range(2):
        eng.decode(floor_norm=gaussian, want_probs=False, width=width)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.decode(floor_norm=gaussian, want_probs=False, width=width)
    e1.record()
    barrier()
    ms_align = max_over_ranks(e0.elapsed_time(e1)) / args.steps

    if rank == 0:
        T_mean = pk.n_phones_total / max(pk.n_pairs, 1)
        n_mean = pk.n_regions / max(pk.n_pairs, 1)
        bpp = bytes_per_pair(T_mean, n_mean)
        dom = max(kern_ms, key=kern_ms.get)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        achieved = pk.n_pairs * bpp / (kern_ms[dom] * 1e-3) / 1e9
        n_arr = np.diff(pk.region_off).astype(np.float64)
        fpp = flops_per_pair(T_mean, n_mean, float(np.mean(n_arr ** 3)))
        fp64_peak = 37.0          # TFLOP/s, measured: profiles/r01_fp64_peak_microbench.txt
        traffic = None
        try:                      # DRAM bytes of the dominant kernel, ncu --set full at this workload
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')))
            if tr.get('pairs_per_launch') == pk.n_pairs and tr.get('kernel') == dom:
                traffic = tr['dram_bytes_per_launch']
        except (OSError, ValueError):
            pass
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, args.pairs),
            'avg_log_likelihood': avg_ll,
            'e2e': {'value': args.pairs / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e},
            'gpu_launches': (eng.kernel_launches_per_iteration() + eng.kernel_launches_per_iteration(args.chunks))
                            * args.steps * world,
            'kernel_ms_per_step': kern_ms,
            'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650',
                         'algorithmic_bytes_per_pair': bpp, 'pairs_per_launch': pk.n_pairs,
                         'note': 'path is FP64-pipe bound (SURVEY 8d), see DESIGN.md'},
            'fp64_roofline': {'bound': 'fp64 pipe (DFMA/DMMA)', 'achieved': args.pairs * fpp / (ms_step * 1e-3) / 1e12 / world,
                              'peak': fp64_peak, 'unit': 'TFLOP/s per GPU', 'frac': args.pairs * fpp / (ms_step * 1e-3) / 1e12 / world / fp64_peak,
                              'algorithmic_flop_per_pair': fpp,
                              'peak_source': 'profiles/r01_fp64_peak_microbench.txt (measured DFMA = DMMA = 37.0)'},
            'gemm_roofline': {'bound': 'fp64 tensor path (DMMA m8n8k4)', 'kernel': 'posterior',
                              'achieved': 2.0 * pk.n_regions * (D_FEAT + 1) * K_CONCEPTS / (kern_ms['posterior'] * 1e-3) / 1e12,
                              'peak': fp64_peak, 'unit': 'TFLOP/s',
                              'frac': 2.0 * pk.n_regions * (D_FEAT + 1) * K_CONCEPTS / (kern_ms['posterior'] * 1e-3) / 1e12 / fp64_peak,
                              'note': 'regions x concepts emission GEMM + row softmax; ncu: profiles/r01_ncu_full_summary_final.txt'},
            'align': {'value': args.pairs / (ms_align * 1e-3), 'unit': 'pairs/s', 'ms_per_pass': ms_align,
                      'what': 'align + cluster of every pair (posterior GEMM + Viterbi kernel), resident'},
            'cpu_baseline': cpu_baseline,
            'clocks': clock_info,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--pairs', type=int, default=1000000)
    ap.add_argument('--variant', default='coco5', choices=['coco5', 'coco10', 'flickr'])
    ap.add_argument('--model', default='linear', choices=['linear', 'gaussian'],
                    help="image posterior: linear softmax (default, BASELINE configs[0]/[4]) or RBF (configs[1]); "
                         "the CPU arm always times the linear class")
    ap.add_argument('--mixed', default='float64',
                    help="precision of the floor-free parts: 'float64' (reference arithmetic) | 'mixed' | subset like 'concept+posterior'")
    ap.add_argument('--cpu-pairs', type=int, default=0, help='CPU-arm sample size (default 2048 x cores)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--chunks', type=int, default=0,
                    help='chunks of the streamed (e2e) iteration (0 = one per ~640 MB of shard, at most 16)')
    args = ap.parse_args()
    apply_variant(args.variant)
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
