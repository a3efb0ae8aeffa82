"""world_size-2 gloo tests (CPU) of the sharding / gradient all-reduce plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tc_gan_b200 import dist as sd
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        n = 11
        mine = sd.shard_indices(n)
        assert mine == list(range(rank, n, world_size))
        # every rank contributes gradients of its own networks; the sum must equal the serial sum
        rs = np.random.RandomState(0)
        per_net = rs.randn(n, 3, 2, 2)
        loc = per_net[mine].sum(axis=0)
        dJ, dD, dS = sd.allreduce_generator_grads(*(torch.tensor(loc[i]) for i in range(3)))
        tot = per_net.sum(axis=0)
        np.testing.assert_allclose(dJ.numpy(), tot[0], atol=1e-12)
        np.testing.assert_allclose(dD.numpy(), tot[1], atol=1e-12)
        np.testing.assert_allclose(dS.numpy(), tot[2], atol=1e-12)
        # rejection bookkeeping in global generator order
        ok_global = np.array([1, 0, 1, 1, 0, 1, 1, 1, 0, 1, 1, 1], dtype=bool)     # 12 networks, 6 per rank
        ok_local = torch.tensor(ok_global[rank::world_size])
        keep, n_kept, rej = sd.first_successes(ok_local, num=5)
        want_global = np.zeros(12, bool)
        want_global[[0, 2, 3, 5, 6]] = True
        assert n_kept == 5 and rej == 2
        np.testing.assert_array_equal(keep.numpy(), want_global[rank::world_size])
        keep, n_kept, rej = sd.first_successes(ok_local, num=20)
        assert n_kept == int(ok_global.sum()) and rej == int((~ok_global).sum())
        out.put((rank, 'ok'))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert got == [(0, 'ok'), (1, 'ok')]


def test_single_process_paths():
    from tc_gan_b200 import dist as sd
    assert sd.shard_indices(5) == [0, 1, 2, 3, 4]
    assert sd.shard_indices(7, rank=1, world_size=3) == [1, 4]
    a = torch.arange(4.).reshape(2, 2)
    dJ, dD, dS = sd.allreduce_generator_grads(a, a + 1, a + 2)
    assert torch.equal(dJ, a) and torch.equal(dS, a + 2)
    keep, n, rej = sd.first_successes(torch.tensor([True, False, True, True]), num=2)
    assert keep.tolist() == [True, False, True, False] and n == 2 and rej == 1
