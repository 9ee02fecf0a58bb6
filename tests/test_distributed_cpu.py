"""World-size-2 gloo tests of the host-side multi-GPU logic (no GPU needed): slot partitioning,
the in-place all-gather of per-replica energies, identical swap decisions on every rank
(replica_exchange.jl:158-178 semantics, checked against the oracle), and the histogram all-reduce of
parallel multicanonical (parallel_multicanonical.jl:38-52)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import mcx_b200 as m
    from mcx_b200.rng import exchange_u
    n, seed, rounds = 8, 77, 30
    backend = m.GPUBackend()
    assert backend.rank == rank and backend.size == world and backend.is_root == (rank == 0)
    first, count = backend.slots(n)
    assert (first, count) == (rank * n // world, n // world)
    betas = m.set_betas(n, 0.3, 0.6, "uniform")
    pt = m.ParallelTempering(betas, seed=seed, backend=backend)
    assert [a.rng.chain for a in pt.replica.algs] == list(range(n))
    rng = np.random.default_rng(1234)          # same stream on both ranks: the "true" energies
    hist = torch.zeros(16, dtype=torch.int64)
    for rd in range(rounds):
        energies = rng.integers(-2000, -1000, n).astype(np.float64)
        x = torch.zeros(n, dtype=torch.float64)
        x[first:first + count] = torch.from_numpy(energies[first:first + count])   # each rank knows only its slots
        backend.all_gather_inplace(x, first, count)
        assert np.array_equal(x.numpy(), energies)
        pt.update_(x.numpy())                  # every rank evaluates the same decisions
        hist[rd % 16] += rank + 1
    backend.all_reduce_sum(hist)
    res = {"indices": pt.indices.copy(), "steps": pt.steps.copy(), "accepted": pt.accepted.copy(),
           "betas": np.array([a.ensemble.beta for a in pt.replica.algs]), "hist": hist.numpy().copy(),
           "u0": exchange_u(seed, 0, 0)}
    np.save(os.path.join(out, "rank%d.npy" % rank), res, allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replica_exchange_and_histogram_merge(tmp_path, oracle):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(str(tmp_path), "rank%d.npy" % r), allow_pickle=True).item() for r in range(world)]
    for k in ("indices", "steps", "accepted", "betas", "hist"):
        assert np.array_equal(res[0][k], res[1][k]), k
    # oracle replay of the same 30 rounds
    import mcx_b200 as m
    n, seed, rounds = 8, 77, 30
    betas = np.array(m.set_betas(n, 0.3, 0.6, "uniform"))
    idx = np.arange(1, n + 1, dtype=np.int64)
    steps = np.zeros(n - 1, dtype=np.int64)
    acc = np.zeros(n - 1, dtype=np.int64)
    rng = np.random.default_rng(1234)
    stage = 0
    for rd in range(rounds):
        xs = rng.integers(-2000, -1000, n).astype(np.float64)
        us = [oracle.lib().mcxo_exchange_u(seed, r, rd) for r in range(n)]
        stage = oracle.rx_update(stage, idx, steps, acc, betas, xs, us)
    assert np.array_equal(res[0]["indices"], idx) and np.array_equal(res[0]["steps"], steps)
    assert np.array_equal(res[0]["accepted"], acc) and np.array_equal(res[0]["betas"], betas)
    assert acc.sum() > 0
    assert res[0]["u0"] == oracle.lib().mcxo_exchange_u(seed, 0, 0)
    expect = np.zeros(16, dtype=np.int64)
    for rd in range(rounds):
        expect[rd % 16] += 3
    assert np.array_equal(res[0]["hist"], expect)


def test_partition_slots():
    import mcx_b200 as m
    assert [m.partition_slots(256, 8, r) for r in (0, 3, 7)] == [(0, 32), (96, 32), (224, 32)]
    with pytest.raises(ValueError):
        m.partition_slots(10, 4, 0)
