"""world_size-2 gloo test (CPU) of the multi-rank host logic: sharding by round-robin over the
(n, T)-sorted order + the fixed-order count reduction reproduce the single-rank counts.  The
per-shard counts come from the oracle (test infrastructure) -- no GPU needed."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_hmm, load_ik, oracle_params_from_golden


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import image_phone_hmm as orc
    from oracle import plain_hmm as ph
    from multimodalworddiscovery_b200.corpus import pack_pairs
    from multimodalworddiscovery_b200.dist import fixed_order_allreduce
    from multimodalworddiscovery_b200.engine_hmm import PackedSentences
    # --- (i,k) model: sum of shard counts == global counts
    g = load_ik('mixed_linear')
    p = oracle_params_from_golden(g)
    pk = pack_pairs(g['feats_list'], g['phones_list'], rank=rank, world=world)
    feats = [g['feats_list'][i] for i in pk.order]
    phones = [g['phones_list'][i] for i in pk.order]
    _, info = orc.em_iteration(feats, phones, p, 'linear')
    lens = sorted(p['init'])
    buf = np.concatenate([info['phoneC'].ravel()] + [info['initC'][m] for m in lens]
                         + [info['transC'][m].ravel() for m in lens]
                         + [[info['avg_ll'] * len(feats)], (info['grad'] * len(feats)).ravel()])
    t = torch.from_numpy(buf.copy())
    fixed_order_allreduce(t)
    np.save(os.path.join(out_dir, 'ik_%d.npy' % rank), t.numpy())
    np.save(os.path.join(out_dir, 'ik_order_%d.npy' % rank), pk.order)
    # --- log-domain plain HMM: log-sum-exp over ranks, plain sum for the LL entry
    h = load_hmm('synth_log')
    lens_h = [int(m) for m in h['lens']]
    pkh = PackedSentences(h['tgt_list'], h['src_list'], h['Vf'], rank=rank, world=world)
    obs = ph.log_initial_obs(h['tgt_list'], h['src_list'], h['Vt'], h['Vf'])
    acc = ph.LogAccumulators(lens_h, h['Vt'], h['Vf'])
    ll = 0.0
    for ex in pkh.order:
        e, f = h['tgt_list'][ex], h['src_list'][ex]
        n = len(e)
        r = ph.log_estep_pair(e, f, obs, np.log(1. / n) * np.ones(n), np.log(1. / n) * np.ones((n, n)))
        acc.init[n] = np.logaddexp(acc.init[n], r['init'])
        ll += r['ll']
    bufh = np.concatenate([acc.init[m] for m in lens_h] + [[ll]])
    th = torch.from_numpy(bufh.copy())
    fixed_order_allreduce(th, log_domain=True, ll_index=len(bufh) - 1)
    np.save(os.path.join(out_dir, 'hmm_%d.npy' % rank), th.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_counts_match_single_rank(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import image_phone_hmm as orc
    from oracle import plain_hmm as ph
    r0, r1 = np.load(str(tmp_path / 'ik_0.npy')), np.load(str(tmp_path / 'ik_1.npy'))
    assert np.array_equal(r0, r1)                          # bitwise identical on every rank
    o0, o1 = np.load(str(tmp_path / 'ik_order_0.npy')), np.load(str(tmp_path / 'ik_order_1.npy'))
    g = load_ik('mixed_linear')
    N = len(g['feats_list'])
    assert sorted(np.concatenate([o0, o1]).tolist()) == list(range(N))
    p = oracle_params_from_golden(g)
    _, info = orc.em_iteration(g['feats_list'], g['phones_list'], p, 'linear')
    lens = sorted(p['init'])
    ref = np.concatenate([info['phoneC'].ravel()] + [info['initC'][m] for m in lens]
                         + [info['transC'][m].ravel() for m in lens]
                         + [[info['avg_ll'] * N], (info['grad'] * N).ravel()])
    np.testing.assert_allclose(r0, ref, rtol=1e-10, atol=1e-300)
    # log-domain
    h0, h1 = np.load(str(tmp_path / 'hmm_0.npy')), np.load(str(tmp_path / 'hmm_1.npy'))
    assert np.array_equal(h0, h1)
    h = load_hmm('synth_log')
    lens_h = [int(m) for m in h['lens']]
    obs = ph.log_initial_obs(h['tgt_list'], h['src_list'], h['Vt'], h['Vf'])
    acc = ph.LogAccumulators(lens_h, h['Vt'], h['Vf'])
    ll = 0.0
    for e, f in zip(h['tgt_list'], h['src_list']):
        n = len(e)
        r = ph.log_estep_pair(e, f, obs, np.log(1. / n) * np.ones(n), np.log(1. / n) * np.ones((n, n)))
        acc.init[n] = np.logaddexp(acc.init[n], r['init'])
        ll += r['ll']
    refh = np.concatenate([acc.init[m] for m in lens_h] + [[ll]])
    np.testing.assert_allclose(h0, refh, rtol=1e-10)
