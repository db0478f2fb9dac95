#!/usr/bin/env python
"""Benchmark of the EM hot path: caption-pairs/s per EM iteration of the image-phone HMM word
discoverer (BASELINE.json metric), on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1|c2|c3|c4|c5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--config selects one of BASELINE.json's configs (default c5 = configs[4], the one the metric is quoted on):
  c1 MSCOCO-2k shape, linear class (2 000 pairs)      c2 the same with the Gaussian (RBF) class
  c3 Flickr30k shape (30 000 pairs, K=100, P=69)      c4 segment-embedding HMM (6 610 utterances, 120-d)
  c5 1 000 000 MSCOCO-shaped pairs, sharded over the ranks

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
host cores over a bounded sample of the SAME corpus (every m-th pair of the sorted order, so the
region-count / caption-length mix is the GPU arm's).  The port is ~3.9x FASTER per process than the
unmodified reference class (profiles/r02_reference_vs_port.json, measured in the build container).
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
    if os.environ.get('MWD_BENCH_CONCEPTS'):          # experiment knob: another concept count on the same corpus shape
        K_CONCEPTS = int(os.environ['MWD_BENCH_CONCEPTS'])


METRIC = 'em_caption_pairs_per_sec'
UNIT = 'pairs/s'
# BASELINE.json configs[0..4]
CONFIGS = {
    'c1': dict(pairs=2000, variant='coco5', model='linear'),
    'c2': dict(pairs=2000, variant='coco5', model='gaussian'),
    'c3': dict(pairs=30000, variant='flickr', model='linear'),
    'c4': dict(pairs=6610, variant='flickr', model='linear'),      # segment-embedding HMM: see c4_arm
    'c5': dict(pairs=1000000, variant='coco5', model='linear'),
}
ORIG_AFFINITY = sorted(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else None


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
    """Algorithmic flop per pair: posterior GEMM + recursion and counts (25 T n K) + restricted concept
    chains + gradient GEMM."""
    K, D = K_CONCEPTS, D_FEAT
    return 2 * n_mean * K * (D + 1) + 25 * T_mean * n_mean * K + 2 * T_mean * K * n3_mean + 2 * n_mean * K * (D + 1)


def bytes_per_pair(T_mean, n_mean):
    """SURVEY 8(d) algorithmic bytes per pair: phones + features + conceptCounts + concept
    alignment argmax + log-likelihood."""
    return 4 * T_mean + 4 * n_mean * D_FEAT + 4 * n_mean * K_CONCEPTS + 4 * T_mean + 8


# ----------------------------------------------------------------------------------------------
# CPU arm: the NumPy oracle port on all host cores, over a sample of the GPU arm's own corpus
# ----------------------------------------------------------------------------------------------
PORT_VS_REFERENCE = 3.9      # port pairs/s / unmodified-reference pairs/s, profiles/r02_reference_vs_port.json


def cpu_sample_numpy(n_pairs, variant, seed=20261018):
    """Fallback sample of the same workload DISTRIBUTION, generated with NumPy (used only when the box has
    no CUDA device to regenerate the GPU arm's corpus; also by tools/probe_reference_speed.py)."""
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


def sample_of_shard(region_off, phone_off, feats, phones, n_sample):
    """Every m-th pair of the (n, T)-sorted shard: same region-count / caption-length mix as the whole corpus."""
    n_pairs = len(region_off) - 1
    idx = np.unique(np.linspace(0, n_pairs - 1, min(n_sample, n_pairs)).astype(np.int64))
    fl = [np.asarray(feats[region_off[i]:region_off[i + 1]], dtype=np.float64) for i in idx]
    pl = [np.asarray(phones[phone_off[i]:phone_off[i + 1]], dtype=np.int64) for i in idx]
    return fl, pl


def _cpu_proc(idx, path, variant, kind, n_rounds, barrier, out_q):
    """One CPU worker: loads ITS slice of the sample (untimed), then runs `n_rounds` EM iterations of
    the NumPy oracle over it, each round released by the shared barrier."""
    from oracle import image_phone_hmm as orc
    apply_variant(variant)
    z = {k: v for k, v in np.load(path).items()}     # NpzFile re-reads an array on every [] access
    ro, po, F, X = z['region_off'], z['phone_off'], z['feats'], z['phones']
    feats = [F[ro[i]:ro[i + 1]] for i in range(len(ro) - 1)]
    phones = [X[po[i]:po[i + 1]] for i in range(len(po) - 1)]
    if kind == 'gaussian':
        params = orc.initial_params(feats, K_CONCEPTS, P_PHONES, 'gaussian', mus=z['post'], width=float(D_FEAT), lr=0.1)
    else:
        params = orc.initial_params(feats, K_CONCEPTS, P_PHONES, 'linear', W=z['post'], lr=0.1)
    params['toeplitz'] = bool(z['toeplitz'])
    for _ in range(n_rounds):
        barrier.wait()
        orc.em_iteration(feats, phones, params, kind)
        barrier.wait()
    out_q.put(idx)


def run_cpu_arm(feats, phones, post, toeplitz, variant, kind, steps, warmup, cores):
    """Times `steps` EM iterations of the oracle over the given pairs with `cores` single-BLAS-thread
    processes.  Returns (pairs/s, seconds per step, pairs actually run)."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    cores = max(1, min(cores, len(feats)))
    tmp = tempfile.mkdtemp(prefix='mwd_cpu_arm_')
    paths, total = [], 0
    for w in range(cores):
        fl, pl = feats[w::cores], phones[w::cores]
        total += len(fl)
        ro = np.concatenate([[0], np.cumsum([len(v) for v in fl])]).astype(np.int64)
        po = np.concatenate([[0], np.cumsum([len(x) for x in pl])]).astype(np.int64)
        path = os.path.join(tmp, 'w%d.npz' % w)
        np.savez(path, region_off=ro, phone_off=po, feats=np.concatenate(fl), phones=np.concatenate(pl), post=post,
                 toeplitz=np.array(bool(toeplitz)))
        paths.append(path)
    saved = {k: os.environ.get(k) for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS')}
    for k in saved:
        os.environ[k] = '1'
    barrier = ctx.Barrier(cores + 1)
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_proc, args=(i, paths[i], variant, kind, warmup + steps, barrier, q), daemon=True)
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
    for path in paths:
        os.unlink(path)
    os.rmdir(tmp)
    return total * len(times) / sum(times), float(np.mean(times)), total


def cpu_sample_for(args, host=None, post=None):
    """(feats, phones, posterior parameter, how) of the CPU arm: a strided sample of the GPU arm's corpus when one
    is at hand (or can be regenerated on this box's GPU), else NumPy draws from the same distribution."""
    cores = os.cpu_count() or 1
    n_sample = args.cpu_pairs or min(args.pairs, 2048 * cores)
    if host is None:
        try:
            import torch
            if torch.cuda.is_available():
                dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
                sh = make_shard(torch, dev, args.pairs, 0, 1, args.variant)
                host = {k: sh[k].cpu().numpy() for k in ('region_off', 'phone_off', 'feats', 'phones')}
                post = (sh['mus'] if args.model == 'gaussian' else sh['W']).cpu().numpy()
                del sh
                torch.cuda.empty_cache()
        except ImportError:
            host = None
    if host is not None:
        fl, pl = sample_of_shard(host['region_off'], host['phone_off'], host['feats'], host['phones'], n_sample)
        return fl, pl, post, 'every m-th pair of the GPU arm\'s corpus (same generator, same n / T mix)'
    fl, pl, W = cpu_sample_numpy(n_sample, args.variant)
    if args.model == 'gaussian':
        raise SystemExit('the NumPy fallback sample has no RBF anchors: run --impl reference --model gaussian on a CUDA box')
    return fl, pl, W, 'NumPy draws from the same distribution (no CUDA device to regenerate the corpus)'


def cpu_line(args, pps, sec, n_run, cores, how):
    return {'value': pps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d pairs, %s; one EM iteration of the NumPy oracle port, %d processes x 1 BLAS thread, %.1f s '
                      'per step; the port runs %.1fx the pairs/s of the unmodified reference class per process '
                      '(profiles/r02_reference_vs_port.json)' % (n_run, how, cores, sec, PORT_VS_REFERENCE)}


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.config == 'c4':
        return c4_reference_arm(args)
    args.mixed = 'float64 (NumPy)'
    cores = os.cpu_count() or 1
    fl, pl, post, how = cpu_sample_for(args)
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    toeplitz = args.variant != 'coco5'
    pps, sec, n_run = run_cpu_arm(fl, pl, post, toeplitz, args.variant, args.model, steps, warmup, cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': pps, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, args.pairs),
        'cpu_baseline': cpu_line(args, pps, sec, n_run, min(cores, n_run), how),
        'e2e': {'value': pps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args, n_pairs):
    feat_gb = n_pairs * (3 if args.variant == 'flickr' else 5) * D_FEAT * 4 / 1e9
    return {'workload': '%s: image-phone HMM EM iteration, synthetic %s shape (%s): %d pairs, '
                        'T~clip(N(%g,%g),15,125), K=%d concepts, P=%d phones, D=%d res34-like features'
                        % (args.config, 'Flickr30k' if args.variant == 'flickr' else 'MSCOCO', args.variant, n_pairs,
                           T_MEAN, T_STD, K_CONCEPTS, P_PHONES, D_FEAT),
            'pairs': n_pairs, 'variant': args.variant,
            'class': 'ImagePhoneGaussianHMMWordDiscoverer' if args.model == 'gaussian' else 'ImagePhoneHMMWordDiscoverer',
            'precision': args.mixed,
            'l2_policy': ('inputs (%.2f GB features) exceed the 126 MB L2' % feat_gb) if feat_gb > 0.126 else
                         ('inputs (%.3f GB features) fit in L2: a 256 MB buffer is overwritten between timed iterations' % feat_gb),
            'parallelism': 'pairs sharded over %d GPU(s), one packed fp64 count all-gather per iteration' % args.gpus}


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
def _rel_to_scale(a, b):
    """max |a - b| relative to the table's scale max |b| (entries of W / counts pass through zero)."""
    import torch
    s = float(b.abs().max())
    return float((a - b).abs().max()) / s if s > 0 else 0.0


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
    affinity = pin_to_gpu_numa_node(local_rank)

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    sh = make_shard(torch, dev, args.pairs, rank, world, args.variant)
    # pinned host copy of the shard (the e2e leg copies from here every step); allocated after the process was
    # bound to the GPU's NUMA node, so first touch places the pages next to the GPU's PCIe root
    host = {}
    for k in ('region_off', 'phone_off', 'feats', 'phones'):
        h = torch.empty(sh[k].shape, dtype=sh[k].dtype, pin_memory=True)
        h.copy_(sh[k])
        host[k] = h
    torch.cuda.synchronize()
    host_np = {k: v.numpy() for k, v in host.items()}
    pk = pack_sorted_arrays(host_np['region_off'], host_np['phone_off'], host_np['feats'], host_np['phones'],
                            lens=sh['lens'], n_pairs_global=args.pairs)
    del sh['feats'], sh['phones']
    torch.cuda.empty_cache()
    gaussian = args.model == 'gaussian'
    post0 = (sh['mus'] if gaussian else sh['W']).cpu().numpy()

    # CPU baseline (rank 0, N=1 only): the NumPy port on a strided sample of this very corpus, GPU idle
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        fl, pl, post, how = cpu_sample_for(args, host_np, post0)
        bound = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, ORIG_AFFINITY)            # the CPU arm gets every core, not the GPU's NUMA node only
        pps, sec, n_run = run_cpu_arm(fl, pl, post, args.variant != 'coco5', args.variant, args.model, 1, 0, cores)
        os.sched_setaffinity(0, bound)
        cpu_baseline = cpu_line(args, pps, sec, n_run, min(cores, n_run), how)
        del fl, pl

    width = float(D_FEAT) if gaussian else 1.0       # RBF width of the order of |v - mu|^2 (unit-variance noise)
    eng = IKEngine(pk, K_CONCEPTS, P_PHONES, gaussian=gaussian, device=dev, keep_concept_counts_a=False,
                   mixed_precision=args.mixed)
    shard_bytes = sum(host[k].numel() * host[k].element_size() for k in host)
    if args.chunks <= 0:
        # enough chunks to overlap the PCIe copy with the kernels -- the kernels of the LAST chunk are the exposed
        # tail -- not so many that a small corpus drowns in launches (1 M pairs on one GPU: 10.4 GB -> 32
        # chunks; an eighth of it: 1.3 GB -> 10 chunks; MSCOCO-2k: 21 MB -> 1 chunk)
        args.chunks = int(min(32, max(1, shard_bytes // (128 << 20))))
    in_l2 = shard_bytes * world < 2 * 126e6
    flush_buf = torch.empty((256 << 20,), dtype=torch.uint8, device=dev) if in_l2 else None
    use_graph = world == 1 and pk.n_pairs <= IKEngine.GRAPH_MAX_PAIRS and os.environ.get('MWD_GRAPH', '1') != '0'

    # initializeModel(): uniform init/trans/obs, injected W  (parameter snapshot restored every step)
    init = {m: np.ones(m) / m for m in pk.lens}
    trans = {m: np.ones((m, m)) / m for m in pk.lens}
    obs = np.ones((K_CONCEPTS, P_PHONES)) / P_PHONES
    eng.set_params(init, trans, obs, post0)
    ptensors = (eng.init_t, eng.trans_t, eng.obsT, eng.post)
    snap = [t.clone() for t in ptensors]
    h_params = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in snap]
    h_out = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in snap]
    h_ll = torch.empty((1,), dtype=torch.float64, pin_memory=True)
    lr, mom = 0.1, 0.0

    def restore():
        for dst, src in zip(ptensors, snap):
            dst.copy_(src)

    def step_resident(timers=None, graph=None):
        restore()
        if (use_graph if graph is None else graph) and timers is None:
            return eng.em_iteration_graph(lr, mom, width)
        return eng.em_iteration(lr, mom, width, with_cA=False, timers=timers)

    def step_e2e():
        # host -> device: the shard (streamed in chunks that overlap the kernels) and the
        # parameters; device -> host: LL + updated tables
        for dst, src in zip(ptensors, h_params):
            dst.copy_(src, non_blocking=True)
        ll = eng.em_iteration_streamed(host, lr, mom, width, n_chunks=args.chunks)
        h_ll.copy_(ll.reshape(1), non_blocking=True)
        for dst, src in zip(h_out, ptensors):
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

    def timed_loop(fn, steps):
        """K steps bracketed by barrier + synchronize; with an L2-resident corpus every step is timed on its own and
        a 256 MB buffer is overwritten in between (outside the timed events)."""
        barrier()
        if flush_buf is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = fn()
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)) / steps, out
        evs = []
        for _ in range(steps):
            flush_buf.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)) / steps, out

    # ---- resident leg ------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    timers = []
    collect = None if use_graph else timers       # a graph replay has no per-kernel events: separate pass below
    ms_step, ll = timed_loop(lambda: step_resident(collect), args.steps)
    clock_info = clocks.stop() if clocks else None
    if use_graph:
        for _ in range(2):
            step_resident(graph=False)
        n_t = min(args.steps, 3)
        for _ in range(n_t):
            step_resident(timers, graph=False)
        torch.cuda.synchronize()
    else:
        n_t = args.steps
    value = args.pairs / (ms_step * 1e-3)
    avg_ll = float(ll) / args.pairs
    kern_ms = {}
    for name, a, b in timers:
        kern_ms[name] = kern_ms.get(name, 0.0) + a.elapsed_time(b)
    kern_ms = {k: v / n_t for k, v in kern_ms.items()}

    # ---- float64 path beside it + parity of the mixed path against it, same parameters, same run ---------------
    f64_info, parity = None, None
    if eng.mixed:
        bits = eng.mixed
        def one_iteration(spec):
            restore()
            eng.set_mixed(spec)
            eng.estep(width, with_cA=False)
            red = eng.reduced.clone()                 # [counts | gradient] as the E-step leaves them
            eng.allreduce()
            ll_ = eng._keep_ll().clone()
            eng.mstep(lr, mom, width)
            return [red] + [t.clone() for t in ptensors], ll_
        # the comparison starts from the parameters AFTER one float64 iteration: from the uniform start every concept
        # explains a caption equally well, conceptCounts == pz and the gradient is pure rounding noise (~1e-13)
        one_iteration(0)
        snap0, snap = snap, [t.clone() for t in ptensors]
        got, _ = one_iteration(args.mixed)
        ref, _ = one_iteration(0)
        snap = snap0
        ll64 = None
        cl = eng.counts_len
        pe = K_CONCEPTS * P_PHONES
        parity = {
            'what': 'second EM iteration (from the parameters one float64 iteration leaves), %s path vs float64 path: max '
                    '|diff| relative to each table\'s largest entry; north-star tolerance 1e-5' % args.mixed,
            'log_likelihood_rel': abs(float(got[0][cl - 1]) - float(ref[0][cl - 1])) / abs(float(ref[0][cl - 1])),
            'phone_counts': _rel_to_scale(got[0][:pe], ref[0][:pe]),
            'init_trans_counts': _rel_to_scale(got[0][pe:cl - 1], ref[0][pe:cl - 1]),
            # the gradient enters the model as lr * grad / N: its error is quoted relative to the parameter it moves
            'posterior_gradient_step': float((got[0][cl:] - ref[0][cl:]).abs().max()) * lr / args.pairs / float(ref[4].abs().max()),
            'obs_after_mstep': _rel_to_scale(got[3], ref[3]),
            'posterior_param_after_mstep': _rel_to_scale(got[4], ref[4]),
        }
        parity['max'] = max(v for k, v in parity.items() if k != 'what')
        t64 = []
        ms64, _ = timed_loop(lambda: step_resident(t64, graph=False), max(2, min(args.steps, 3)))
        k64 = {}
        for name, a, b in t64:
            k64[name] = k64.get(name, 0.0) + a.elapsed_time(b)
        f64_info = {'ms_per_step': ms64, 'value': args.pairs / (ms64 * 1e-3), 'unit': UNIT,
                    'kernel_ms_per_step': {k: v / max(2, min(args.steps, 3)) for k, v in k64.items()},
                    'what': 'the same iteration with every kernel in float64 (precision float64, the class default)'}
        eng.set_mixed(bits)

    # ---- e2e leg -----------------------------------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e, _ = timed_loop(step_e2e, args.steps)
    h2d = shard_bytes + sum(t.numel() * t.element_size() for t in h_params)
    d2h = 8 + sum(t.numel() * t.element_size() for t in h_out)
    h2d_local = h2d
    if world > 1:
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        h2d, d2h = int(tot[0]), int(tot[1])
        per_rank = [None] * world              # where each rank's staging lives (diagnoses the host->device limiter)
        dist.all_gather_object(per_rank, affinity)
        affinity = per_rank

    # ---- align leg ("plus align-pairs/s separately"): batched Viterbi + cluster of the whole
    # shard under the current parameters (posterior GEMM + K6), alignments written to HBM
    restore()
    for _ in range(2):
        eng.decode(floor_norm=gaussian, want_probs=False, width=width)
    align_evs = []

    def align_pass():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.decode(floor_norm=gaussian, want_probs=False, width=width)
        e1.record()
        align_evs.append((e0, e1))
        return out

    ms_align, _ = timed_loop(align_pass, args.steps)
    ms_align_each = [a.elapsed_time(b) for a, b in align_evs]

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
        try:                      # DRAM bytes of the dominant kernel per pair (ncu --set full), scaled to this launch
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
            if tr.get('variant') == args.variant and tr.get('n_concepts') == K_CONCEPTS and args.mixed == 'mixed':
                traffic = tr['kernels'][dom]['dram_bytes_per_pair'] * pk.n_pairs
        except (OSError, ValueError, KeyError):
            pass
        gemm_flop = 2.0 * pk.n_regions * (D_FEAT + 1) * K_CONCEPTS
        if eng._tc_posterior:
            npad = (K_CONCEPTS + 15) & ~15
            tf32_peak = float(peaks.get('bf16_tflops', 2250.0)) / 2.0
            issued = 3.0 * 2.0 * pk.n_regions * D_FEAT * npad
            gemm = {'bound': 'tensor (tcgen05.mma kind::tf32, 3 split passes, N padded to %d)' % npad, 'kernel': 'posterior',
                    'achieved': issued / (kern_ms['posterior'] * 1e-3) / 1e12, 'peak': tf32_peak, 'unit': 'TFLOP/s',
                    'frac': issued / (kern_ms['posterior'] * 1e-3) / 1e12 / tf32_peak,
                    'useful_tflops': gemm_flop / (kern_ms['posterior'] * 1e-3) / 1e12,
                    'hbm_gbs': (pk.n_regions * (4.0 * D_FEAT + 8.0 * K_CONCEPTS)) / (kern_ms['posterior'] * 1e-3) / 1e9,
                    'hbm_frac': (pk.n_regions * (4.0 * D_FEAT + 8.0 * K_CONCEPTS)) / (kern_ms['posterior'] * 1e-3) / 1e9 / peak,
                    'peak_source': 'MEASURED_PEAKS.json bf16_tflops / 2 (TF32 issues at half the bf16 rate)',
                    'note': 'regions x concepts emission GEMM + row softmax; the kernel is bound by HBM (feature read + '
                            'float64 posterior write) and shared-memory operand traffic, see DESIGN.md'}
        else:
            gemm = {'bound': 'fp64 tensor path (DMMA m8n8k4)', 'kernel': 'posterior',
                    'achieved': gemm_flop / (kern_ms['posterior'] * 1e-3) / 1e12, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                    'frac': gemm_flop / (kern_ms['posterior'] * 1e-3) / 1e12 / fp64_peak,
                    'note': 'regions x concepts emission GEMM + row softmax'}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None,
            'dtype': 'f64' if not eng.mixed else 'f64 (counts, statistics, softmax, M-step) + scaled f32 forward/backward lattice '
                     '+ f32 concept chains (f32-pair clamped emission) + split-tf32 tensor-core GEMMs with fp32 TMEM accumulation '
                     '(precision %s)' % args.mixed,
            'data': 'synthetic',
            'config': workload_config(args, args.pairs),
            'avg_log_likelihood': avg_ll,
            'e2e': {'value': args.pairs / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e,
                    'h2d_gbs_per_rank': h2d_local / (ms_e2e * 1e-3) / 1e9, 'cpu_affinity': affinity},
            'gpu_launches': ((1 if use_graph else eng.kernel_launches_per_iteration())
                             + eng.kernel_launches_per_iteration(args.chunks)) * args.steps * world,
            'launch_mode': 'one CUDA-graph replay per iteration (%d kernels inside)' % eng.kernel_launches_per_iteration()
                           if use_graph else 'kernel by kernel',
            'kernel_ms_per_step': kern_ms,
            'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650',
                         'algorithmic_bytes_per_pair': bpp, 'pairs_per_launch': pk.n_pairs,
                         'note': 'the path is bound by the FP32 / FP64 issue rate, not HBM (SURVEY 8d), see DESIGN.md'},
            'fp64_roofline': {'bound': 'fp64 pipe (DFMA/DMMA)', 'achieved': args.pairs * fpp / (ms_step * 1e-3) / 1e12 / world,
                              'peak': fp64_peak, 'unit': 'TFLOP/s per GPU', 'frac': args.pairs * fpp / (ms_step * 1e-3) / 1e12 / world / fp64_peak,
                              'algorithmic_flop_per_pair': fpp,
                              'peak_source': 'profiles/r01_fp64_peak_microbench.txt (measured DFMA = DMMA = 37.0)',
                              'note': 'whole-iteration algorithmic flop (SURVEY 8d, 25 T n K recursion term) over the step time; with '
                                      'the GEMMs on the tensor cores part of this flop no longer runs on the FP64 pipe'},
            'gemm_roofline': gemm,
            'parity_vs_float64': parity,
            'float64_path': f64_info,
            'align': {'value': args.pairs / (ms_align * 1e-3), 'unit': 'pairs/s', 'ms_per_pass': ms_align, 'ms_each_pass': ms_align_each,
                      'what': 'align + cluster of every pair (posterior GEMM + Viterbi kernel), resident'},
            'cpu_baseline': cpu_baseline,
            'clocks': clock_info,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def pin_to_gpu_numa_node(local_rank):
    """Bind this rank to the CPUs of its GPU's NUMA node BEFORE any pinned buffer is allocated (first touch then
    places the staging pages on the memory next to the GPU's PCIe root complex).  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = (n_cpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinityWithinScope(h, words, pynvml.NVML_AFFINITY_SCOPE_NODE)
        cpus = sorted(i for i in range(n_cpu) if (mask[i // 64] >> (i % 64)) & 1)
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return 'NUMA-local: %d of %d CPUs (%d..%d)' % (len(cpus), len(allowed), cpus[0], cpus[-1])
        return 'all %d CPUs (single NUMA node or no NVML mask)' % len(allowed)
    except Exception as e:                      # pragma: no cover - informational only
        return 'unchanged (%s)' % type(e).__name__


# ----------------------------------------------------------------------------------------------
# config c4: segment-embedding HMM (SegEmbedHMMWordDiscoverer's acoustic model)
# ----------------------------------------------------------------------------------------------
C4_VT, C4_D, C4_M = 1422, 120, 1


def c4_corpus(n_utts, seed=20261018 + 4):
    """S ~ clip(N(26,10),4,98) segments of 120-d embeddings ~ N(mu_word, 0.02 I); states = NULL + n
    concepts (n ~ empirical Flickr 1..8, words Zipf(1.0) over a 1 422-word vocabulary)."""
    rng = np.random.default_rng(seed)
    ns = rng.choice(8, size=n_utts, p=np.array(FLICKR_N_PMF)) + 1
    Ss = np.clip(np.round(rng.normal(26, 10, n_utts)), 4, 98).astype(np.int64)
    pw = 1.0 / np.arange(1, C4_VT)
    cent = rng.standard_normal((C4_VT, C4_D))
    tgt, embs = [], []
    for n, S in zip(ns, Ss):
        e = np.concatenate([[0], 1 + rng.choice(C4_VT - 1, size=n, p=pw / pw.sum())])
        st = rng.integers(0, len(e), S)
        embs.append((cent[e[st]] + np.sqrt(0.02) * rng.standard_normal((S, C4_D))).astype(np.float32))
        tgt.append(e)
    means = cent[:, None, :] + 0.1 * rng.standard_normal((C4_VT, C4_M, C4_D))
    return tgt, embs, means


def c4_params(tgt, means):
    lens = sorted({len(e) for e in tgt})
    return dict(init={m: np.log(1. / m) * np.ones(m) for m in lens},
                trans={m: np.log(1. / m) * np.ones((m, m)) for m in lens},
                lprior=np.zeros((C4_VT, C4_M)), means=means.copy(), var=0.02 * np.ones((C4_VT, C4_M, C4_D)))


def c4_cpu(tgt, embs, means, n_sample, steps):
    """The NumPy restatement (oracle/segembed_hmm.py, pinned to the reference's gaussian / gmmProb / embed) on a
    prefix sample, one process (the reference class is single-threaded Python)."""
    from oracle import plain_hmm as ph
    from oracle import segembed_hmm as sh
    tg, em = tgt[:n_sample], [x.astype(np.float64) for x in embs[:n_sample]]
    p = c4_params(tg, means)
    lens = sorted(p['init'])
    times = []
    for _ in range(steps):
        acc = ph.LogAccumulators(lens, C4_VT, 1)
        t0 = time.perf_counter()
        sh.em_iteration(em, tg, p, acc)
        times.append(time.perf_counter() - t0)
    return len(tg) / float(np.mean(times)), float(np.mean(times)), len(tg)


def c4_config(n_utts, n_seg):
    return {'workload': 'c4: segment-embedding HMM EM iteration (SegEmbedHMMWordDiscoverer acoustic model), %d synthetic '
                        'utterance-image pairs, %d segments of %d-d embeddings, states = NULL + n concepts, vocabulary %d, '
                        '%d Gaussian per word' % (n_utts, n_seg, C4_D, C4_VT, C4_M),
            'pairs': n_utts, 'class': 'SegEmbedHMMWordDiscoverer',
            'l2_policy': 'inputs (%.1f MB embeddings) fit in L2: a 256 MB buffer is overwritten between timed iterations'
                         % (n_seg * C4_D * 4 / 1e6),
            'parallelism': 'single GPU (the corpus is 6 610 utterances)'}


def c4_reference_arm(args):
    n_utts = args.pairs
    tgt, embs, means = c4_corpus(n_utts)
    steps = max(1, min(args.steps, 2))
    pps, sec, n_run = c4_cpu(tgt, embs, means, args.cpu_pairs or 2048, steps)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': pps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': 0, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': c4_config(n_utts, int(sum(len(x) for x in embs))),
        'cpu_baseline': {'value': pps, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': 'first %d utterances of the same corpus, one EM iteration of oracle/segembed_hmm.py, '
                                   '1 process (%.1f s per step)' % (n_run, sec)},
        'e2e': {'value': pps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


def c4_arm(args):
    import torch
    from multimodalworddiscovery_b200.engine_hmm import SegmentHMMEngine
    if args.gpus != 1:
        raise SystemExit('config c4 is a 6 610-utterance corpus: single GPU only')
    n_utts = args.pairs
    tgt, embs, means = c4_corpus(n_utts)
    n_seg = int(sum(len(x) for x in embs))
    cpu_baseline = None
    if not args.no_cpu_baseline:
        pps, sec, n_run = c4_cpu(tgt, embs, means, args.cpu_pairs or 2048, 1)
        cpu_baseline = {'value': pps, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                        'sample': 'first %d utterances of the same corpus, one EM iteration of oracle/segembed_hmm.py '
                                  '(pinned to the reference\'s gaussian / gmmProb / embed), 1 process (%.1f s)' % (n_run, sec)}
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    eng = SegmentHMMEngine(tgt, embs, C4_VT, C4_M, device=dev)
    p0 = c4_params(tgt, means)
    host_emb = torch.empty(eng.emb.shape, dtype=eng.emb.dtype, pin_memory=True).copy_(eng.emb)
    snap = None
    flush_buf = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)

    def reset():
        eng.set_chain_params(p0['init'], p0['trans'])
        eng.set_emission_params(p0['lprior'], p0['means'], p0['var'])
        eng.reset_accumulators()

    reset()
    snap = [t.clone() for t in (eng.init_t, eng.trans_t, eng.lprior, eng.means, eng.var)]

    def restore():
        for d, s_ in zip((eng.init_t, eng.trans_t, eng.lprior, eng.means, eng.var), snap):
            d.copy_(s_)
        eng.reset_accumulators()

    def step():
        restore()
        return eng.em_iteration()

    h_ll = torch.empty((1,), dtype=torch.float64, pin_memory=True)

    def step_e2e():
        eng.emb.copy_(host_emb, non_blocking=True)
        ll = step()
        h_ll.copy_(ll.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps):
        evs = []
        torch.cuda.synchronize()
        for _ in range(steps):
            flush_buf.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / steps, out

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(0)
    ms_step, ll = timed(step, args.steps)
    clock_info = clocks.stop()
    ms_emis, _ = timed(eng.emission, args.steps)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    slots = int(eng.pk.n_slots)
    # 4 bytes x 120 dims per segment in, one float64 log-emission (+ responsibility) per (segment, state) out
    alg_bytes = n_seg * C4_D * 4 + slots * 8 * (1 + C4_M)
    print(json.dumps({
        'metric': METRIC, 'value': n_utts / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': c4_config(n_utts, n_seg),
        'avg_log_likelihood': float(ll) / n_utts,
        'e2e': {'value': n_utts / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(host_emb.numel() * 4),
                'd2h_bytes_per_step': 8, 'ms_per_step': ms_e2e},
        'gpu_launches': 0,
        'kernel_ms_per_step': {'gauss_emission': ms_emis, 'whole_iteration': ms_step},
        'roofline': {'bound': 'hbm', 'kernel': 'gauss_emission', 'achieved': alg_bytes / (ms_emis * 1e-3) / 1e9, 'peak': peak,
                     'unit': 'GB/s', 'frac': alg_bytes / (ms_emis * 1e-3) / 1e9 / peak, 'traffic': None,
                     'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650',
                     'algorithmic_bytes_per_launch': alg_bytes,
                     'note': 'a 6 610-utterance corpus (3 MB of embeddings) is launch / latency bound, not bandwidth bound'},
        'cpu_baseline': cpu_baseline, 'clocks': clock_info}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c5', choices=sorted(CONFIGS),
                    help='BASELINE.json configs[0..4] as c1..c5 (default c5, the one the metric is quoted on)')
    ap.add_argument('--pairs', type=int, default=None)
    ap.add_argument('--variant', default=None, choices=['coco5', 'coco10', 'flickr'])
    ap.add_argument('--model', default=None, choices=['linear', 'gaussian'],
                    help='image posterior: linear softmax or RBF (default: the config\'s)')
    ap.add_argument('--mixed', default='mixed',
                    help="'mixed' (default: tcgen05 split-TF32 tensor-core GEMMs + scaled-float32 forward/backward lattice + "
                         "float32 concept chains, validated at 1e-5 against float64 in the same run) | 'float64' (reference "
                         "arithmetic everywhere) | subset like 'posterior+grad+recursion'")
    ap.add_argument('--cpu-pairs', type=int, default=0, help='CPU-arm sample size (default 2048 x cores)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--chunks', type=int, default=0,
                    help='chunks of the streamed (e2e) iteration (0 = one per ~128 MB of shard, at most 32)')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    for k, v in cfg.items():
        if getattr(args, k) is None:
            setattr(args, k, v)
    apply_variant(args.variant)
    if args.impl == 'reference':
        reference_arm(args)
    elif args.config == 'c4':
        c4_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
