"""The N>1 path of bench.py on CPU: two gloo ranks shard a batch, 'solve' their block with the
oracle, and gather the results; the gathered batch must equal the single-rank answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, outdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from socp_b200 import sharding
    import scenarios as S
    from backends import OracleBackend
    ora = OracleBackend()
    lo, hi = sharding.shard_bounds(total, rank, world)
    assert hi - lo == max(0, min(-(-total // world), total - rank * -(-total // world)))
    # each rank solves its own block of double-integrator problems with different targets
    xs, infos, nfevs = [], [], []
    for k in range(lo, hi):
        spec = S.di_problem()
        spec["Xb"][1][1] = 15.0 + k
        r = ora.solve(spec)
        xs.append(r["x"]); infos.append(r["info"]); nfevs.append(r["nfev"])
    x = torch.tensor(np.array(xs).reshape(-1, 13)); info = torch.tensor(infos, dtype=torch.int32); nfev = torch.tensor(nfevs, dtype=torch.int32)
    gx, gi, gn = sharding.gather_results(x, info, nfev, total=total)
    tot = sharding.reduce_sum([float((info == 1).sum()), float(nfev.sum())], "cpu")
    tmax = sharding.reduce_max(float(rank + 1), "cpu")
    assert tmax == world
    if rank == 0:
        np.save(os.path.join(outdir, "gx.npy"), gx.numpy())
        np.save(os.path.join(outdir, "gi.npy"), gi.numpy())
        np.save(os.path.join(outdir, "tot.npy"), np.array(tot))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total,world", [(6, 2), (5, 2), (1, 2)])     # equal, ragged and empty last shard
def test_two_rank_shard_and_gather(tmp_path, oracle_lib, total, world):
    import scenarios as S
    from backends import OracleBackend
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    gx = np.load(tmp_path / "gx.npy")
    gi = np.load(tmp_path / "gi.npy")
    tot = np.load(tmp_path / "tot.npy")
    assert gx.shape == (total, 13) and gi.shape == (total,)
    ora = OracleBackend()
    for k in range(total):
        spec = S.di_problem()
        spec["Xb"][1][1] = 15.0 + k
        r = ora.solve(spec)
        assert np.array_equal(gx[k], r["x"]) and gi[k] == r["info"]
    assert tot[0] == np.sum(gi == 1)


def test_shard_bounds_cover_exactly():
    from socp_b200.sharding import shard_bounds
    for total in (0, 1, 7, 100000, 10 ** 6 + 3):
        for world in (1, 2, 4, 8):
            blocks = [shard_bounds(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert all(hi >= lo for lo, hi in blocks)
