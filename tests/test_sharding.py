"""CPU, world_size 2 over gloo: the N>1 plumbing -- contiguous batch slices, no data-path
collective, barrier + max-over-ranks timing."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition(nttb200):
    from importlib import import_module
    sh = import_module("ntt-based-polynomial-multiplier-fpga_b200.sharding")
    for batch in (0, 1, 7, 8, 65536, 2**20 + 3):
        for world in (1, 2, 3, 4, 8):
            cuts = [sh.shard_bounds(batch, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == batch
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_bounds(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from importlib import import_module
    sh = import_module("ntt-based-polynomial-multiplier-fpga_b200.sharding")
    from oracle import loader
    dist = sh.init_distributed("gloo")
    O = loader.Oracle()
    n, qq, batch = 64, 257, 10
    a, b = O.random((batch, n), qq, 1), O.random((batch, n), qq, 2)   # same synthetic batch on all ranks
    lo, hi = sh.shard_bounds(batch, world, rank)
    sh.barrier()
    c = O.product(n, qq, a[lo:hi], b[lo:hi], 10)                       # stand-in for the GPU call
    t = sh.max_over_ranks(float(rank + 1))
    done = sh.sum_over_ranks(float(hi - lo))
    every = sh.gather_over_ranks(float(10 * rank + 1))
    q.put((rank, lo, hi, c.tolist(), t, done, every))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import loader
    O = loader.Oracle()
    a, b = O.random((10, 64), 257, 1), O.random((10, 64), 257, 2)
    want = O.product(64, 257, a, b, 10)
    got = np.concatenate([np.array(r[3], dtype=np.int32) for r in res])
    assert (got == want).all()                       # slices tile the batch, no exchange needed
    assert [r[1:3] for r in res] == [(0, 5), (5, 10)]
    assert all(r[4] == 2.0 for r in res)             # max over ranks
    assert all(r[5] == 10.0 for r in res)            # units processed by all ranks
    assert all(r[6] == [1.0, 11.0] for r in res)     # every rank's value, in rank order


def test_c_shard_bounds_match_the_python_partition(nttb200):
    sh = __import__("importlib").import_module("ntt-based-polynomial-multiplier-fpga_b200.sharding")
    for batch in (0, 1, 7, 65536, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            got = [nttb200.shard_bounds(batch, world, r) for r in range(world)]
            assert got == [sh.shard_bounds(batch, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == batch
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
