"""The N>1 path on CPU: instance sharding + the final gather over torch.distributed with the gloo backend, world_size 2
(and 3, to cover a short last block). The data path itself has no collective."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from spice21_b200.shard import gather_instances, shard_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n, rank, world)
    ids = np.arange(lo, hi)
    x = np.stack([ids * 1.5, -ids.astype(np.float64)], axis=1)          # stand-in for per-instance solutions
    iters = (ids % 7).astype(np.int32)                                   # and iteration counts
    fx, fi = gather_instances(x, n), gather_instances(iters, n)
    ok = np.array_equal(fx[:, 0], np.arange(n) * 1.5) and np.array_equal(fi, (np.arange(n) % 7).astype(np.int32))
    q.put((rank, bool(ok), int(fi.sum())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8192), (2, 101), (3, 100)])
def test_shard_and_gather_gloo(world, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert len({s for _, _, s in res}) == 1  # every rank sees the same gathered totals


def _agg_worker(rank, world, port, fails_on, q):
    sys.path.insert(0, ROOT)
    import torch
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    D = bench.Dist(backend="gloo", device="cpu")
    n = 64
    # strong-scaling shard of a 200-instance problem, the device-side gather of equal blocks, and the reductions
    lo, hi = bench.shard(200, rank, world, "strong")
    wlo, whi = bench.shard(200, rank, world, "weak")
    block = torch.full((n,), float(rank), dtype=torch.float64)
    full = D.gather_device(block)
    times = D.max_f([1.0 + rank, 0.5 + rank, 2.0 - rank])
    maybe = D.max_f([None if rank == fails_on else 5.0 + rank])  # a value one rank could not measure is dropped everywhere
    (tot,) = D.sum_i([n * (10 + rank)])
    D.barrier()
    q.put((rank, times, tot, maybe[0], full.tolist(), (lo, hi), (wlo, whi)))
    D.close()


@pytest.mark.parametrize("world,fails_on", [(2, -1), (2, 1), (3, 0)])
def test_bench_collectives_are_symmetric_gloo(world, fails_on):
    """bench.Dist holds every collective of the multi-GPU measurement; every rank must come out of each with the same
    totals — also when one rank could not measure a secondary configuration, which then is dropped everywhere instead of
    leaving the other ranks waiting in a collective (that hang cost a whole 8-GPU run in round 1)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + world * 13 + fails_on) % 2000
    procs = [ctx.Process(target=_agg_worker, args=(r, world, port, fails_on, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    want_times = [1.0 + world - 1, 0.5 + world - 1, 2.0]
    want_tot = 64 * sum(10 + r for r in range(world))
    want_full = [float(r) for r in range(world) for _ in range(64)]
    per = -(-200 // world)
    for rank, times, tot, maybe, full, (lo, hi), (wlo, whi) in res:
        assert times == want_times and tot == want_tot and full == want_full
        assert maybe == (None if fails_on >= 0 else 5.0 + world - 1)
        assert (lo, hi) == (min(200, rank * per), min(200, (rank + 1) * per)) and (wlo, whi) == (rank * 200, (rank + 1) * 200)


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from spice21_b200.shard import shard_bounds
    for n in (1, 7, 8192, 100000):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[k][1] == b[k + 1][0] for k in range(world - 1))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: one JSON line on rank 0 with the contract keys; other ranks print nothing."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
