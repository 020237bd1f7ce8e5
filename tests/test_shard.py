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
