"""Multi-rank host logic on CPU: world_size-2 gloo.  Replica sharding, the parallel-tempering gather / swap / scatter
plumbing and the observable reduction.  The swap decisions themselves are the checker's (oracle) here -- the CUDA swap
kernel is tested against the same checker in tests/test_gpu_ensembles.py."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _oracle_swap(t_all, e_all, seed, rnd):
    from oracle import capi

    t = t_all.numpy()
    e = e_all.numpy()
    order = np.argsort(t, kind="stable")
    acc = torch.zeros(len(t), dtype=torch.int8)
    for k in range(rnd & 1, len(t) - 1, 2):
        ra, rb = order[k], order[k + 1]
        u = capi.uniform_pair(seed, rnd, k, 3)[1]
        if capi.lib().qo_pt_swap_accept(e[ra], e[rb], t[ra], t[rb], u):
            t[ra], t[rb] = t[rb], t[ra]
            acc[ra] = 1
    return acc


def _worker(rank, world, port, n_global, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from quansino_b200.mc.tempering import ReplicaExchange, reduce_observables, shard

        temps = 300.0 * 4.0 ** (np.arange(n_global) / (n_global - 1))
        rng = np.random.default_rng(0)
        energies = 2.0 + 0.004 * temps + rng.normal(scale=0.3, size=n_global)
        off, cnt = shard(n_global, world, rank)
        counts = [shard(n_global, world, r)[1] for r in range(world)]
        t_loc = torch.from_numpy(temps[off : off + cnt].copy())
        e_loc = torch.from_numpy(energies[off : off + cnt].copy())
        ex = ReplicaExchange(e_loc, t_loc, counts, rank, seed=77, swap=_oracle_swap)
        for _ in range(4):
            ex.exchange()
        tot = reduce_observables(torch.tensor([float(cnt), float(e_loc.sum())], dtype=torch.float64))
        out[rank] = (off, t_loc.numpy().copy(), ex.n_accepted, tot.numpy().copy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_tempering_plumbing_world_size_2():
    n_global, world = 11, 2  # uneven shards on purpose
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_global, out), nprocs=world, join=True)
    # single-process reference of the same four rounds
    temps = 300.0 * 4.0 ** (np.arange(n_global) / (n_global - 1))
    rng = np.random.default_rng(0)
    energies = 2.0 + 0.004 * temps + rng.normal(scale=0.3, size=n_global)
    t_all = torch.from_numpy(temps.copy())
    n_acc = 0
    for rnd in range(4):
        n_acc += int(_oracle_swap(t_all, torch.from_numpy(energies), 77, rnd).sum())
    got = np.concatenate([out[r][1] for r in range(world)])
    assert np.array_equal(got, t_all.numpy())  # every rank applied the same permutation to its shard
    assert out[0][2] == out[1][2] == n_acc and n_acc > 0
    assert np.array_equal(np.sort(got), np.sort(temps))
    for r in range(world):
        assert out[r][3][0] == n_global and abs(out[r][3][1] - energies.sum()) < 1e-12
