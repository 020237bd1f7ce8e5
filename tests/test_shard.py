"""The N > 1 path on CPU: world-size-2 (and 3) gloo runs of the partition + exchange logic of
locointerpolate_b200/shard.py, with the oracle standing in for the GPU evaluator (test infrastructure only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import protoschema
from conftest import GOLDEN
from locointerpolate_b200 import shard


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 7, 8, 9, 1000, 1001):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pb, q_np, truth_np, ret):
    from oracle import oracleapi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = oracleapi.Oracle(protoschema.read_binary(pb))

        def evaluate(q):
            out, leaf, status, _ = orc.eval(q.numpy())
            return torch.from_numpy(out), torch.from_numpy(leaf), torch.from_numpy(status)

        def check(q, truth):
            err, split, _ = orc.error_check(q.numpy(), truth.numpy(), 0.05)
            return torch.from_numpy(err), torch.from_numpy(split)

        q, truth = torch.from_numpy(q_np), torch.from_numpy(truth_np)
        out, leaf, status = shard.sharded_eval(evaluate, q)
        err, split = shard.sharded_error_check(check, q, truth)
        if rank == 0:
            ret["out"], ret["leaf"], ret["status"] = out.numpy(), leaf.numpy(), status.numpy()
            ret["err"], ret["split"] = err.numpy(), split.numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Q", [(2, 1001), (3, 100), (2, 1)])
def test_sharded_eval_and_error_reduce_gloo(world, Q):
    from oracle import oracleapi
    pb = os.path.join(GOLDEN, "cub2d_adapt6.pb")
    z = np.load(os.path.join(GOLDEN, "cub2d_adapt6.npz"))
    q, truth = z["q"][:Q].copy(), z["true_values"][:Q].copy()
    if Q > 300:
        truth[300] = np.nan  # NaN must win the max-reduction and force a split
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), pb, q, truth, ret), nprocs=world, join=True)
        got = dict(ret)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    o_out, o_leaf, o_status, _ = orc.eval(q)
    assert np.array_equal(got["leaf"], o_leaf) and np.array_equal(got["status"], o_status)
    assert np.array_equal(got["out"], o_out, equal_nan=True)  # the same evaluator on the same rows: bit-identical
    o_err, o_split, _ = orc.error_check(q, truth, 0.05)
    assert np.array_equal(got["err"], o_err, equal_nan=True) and np.array_equal(got["split"], o_split)
    if Q > 300:
        assert np.isnan(got["err"]).sum() == 1 and got["split"][np.isnan(got["err"])].all()


def _exchange_worker(rank, world, port, n_batches, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = shard.RotatingExchange(n, torch.float64, "cpu")
        assert ex.transport == "p2p"
        ex.warm()

        def shard_of(b, r):   # what rank r "evaluates" for batch b
            return torch.arange(n, dtype=torch.float64) + 1000.0 * b + 10.0 * r

        ok = True
        for epoch in range(2):   # drain() ends an epoch; the exchange is reusable (warm-up, then the timed region)
            checked = 0
            for b in range(n_batches):
                out = ex.out_buffer(b)
                out.copy_(shard_of(b + 100 * epoch, rank))
                ex.submit(b)
                done = b - (world - 1)   # every shard of this batch has been shipped by now (skew: rank r ships batch b at step b + r)
                if done >= 0 and ex.consumer(done) == rank:
                    g = ex.gathered(done)
                    ok &= all(torch.equal(g[r], shard_of(done + 100 * epoch, r)) for r in range(world))
                    checked += 1
            ex.drain()
            for b in range(max(0, n_batches - (world - 1)), n_batches):
                if ex.consumer(b) == rank and b >= n_batches - 2 * world:
                    g = ex.gathered(b)
                    ok &= all(torch.equal(g[r], shard_of(b + 100 * epoch, r)) for r in range(world))
                    checked += 1
            ret[(rank, epoch)] = (bool(ok), checked)
        ex.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_batches", [(2, 7), (3, 8), (3, 2), (4, 9)])
def test_rotating_exchange_schedule(world, n_batches):
    """RotatingExchange (the bench's exchange step for scalar results): batch b is gathered on rank b % world, rank r ships
    batch b at step b + r so that every step is a permutation; point-to-point transport over gloo stands in for the
    copy-engine pushes."""
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_exchange_worker, args=(world, port, n_batches, 257, ret), nprocs=world, join=True)
        res = dict(ret)
    assert len(res) == 2 * world
    total = 0
    for (rank, epoch), (ok, checked) in res.items():
        assert ok, (rank, epoch)
        total += checked
    assert total == 2 * n_batches   # every batch was checked on its consumer, in both epochs
