"""N > 1 host logic of bench.py on CPU: contiguous layer sharding across ranks (world_size 2, gloo) and the
max-over-ranks / sum-over-ranks reductions.  Layers are independent chains, so the shards' chains concatenated per level
must equal the chain of the whole image (checked with the oracle -- this is a test, not the product path)."""
import hashlib
import os
import socket

import numpy as np
import pytest

import bench
from floor_b200.image_types import IMAGE_TYPE as T

M = T.FLAG_MIPMAPPED | T.READ_WRITE


def test_shard_layers_partitions_exactly():
    for total, mult in [(2048, 1), (64, 1), (7, 1), (384, 6), (6, 6), (30, 6)]:
        for world in (1, 2, 3, 4, 8):
            got = [bench.shard_layers(total, world, r, mult) for r in range(world)]
            assert got[0][0] == 0 and sum(n for _, n in got) == total
            for (lo, n), (lo2, _) in zip(got, got[1:]):
                assert lo + n == lo2 and lo % mult == 0 and n % mult == 0


def test_shard_round_robin_partitions_exactly():
    """batches of independent textures: texture i on rank i % N -- every texture exactly once, shares differ by at most one"""
    for total in (512, 7, 1, 64):
        for world in (1, 2, 3, 4, 8):
            shares = [bench.shard_round_robin(total, world, r) for r in range(world)]
            assert sorted(i for sh in shares for i in sh) == list(range(total))
            assert max(len(sh) for sh in shares) - min(len(sh) for sh in shares) <= 1
            assert all(i % world == r for r, sh in enumerate(shares) for i in sh)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, dim, t, cid, out):
    import torch
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    is_cube = bool(t & T.FLAG_CUBE)
    lo, n = bench.shard_layers(dim[2], world, rank)
    sdim = (dim[0], dim[1], n)
    layer_id0 = lo * (6 if is_cube else 1)
    l0 = oracle.fill_synthetic(sdim, t, cid, layer_id0=layer_id0)
    chain = oracle.generate_mip_map_chain(l0, sdim, t, threads=2)
    # per-level digests of this rank's shard, gathered on every rank
    levels = oracle.mip_level_count(sdim, t)
    mine = [hashlib.sha256(chain[oracle.level_offset(sdim, t, l): oracle.level_offset(sdim, t, l) + oracle.level_size(sdim, t, l)].tobytes()).hexdigest() for l in range(levels)]
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, lo, n, mine, chain if rank != 0 else None))
    # the reductions bench.py uses: MAX of a time, SUM of bytes
    tt = torch.tensor([float(rank + 1), float(chain.size)], dtype=torch.float64)
    mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        full_l0 = oracle.fill_synthetic(dim, t, cid)
        full = oracle.generate_mip_map_chain(full_l0, dim, t, threads=2)
        ok = float(mx[0]) == world and float(sm[1]) == full.size
        shards = sorted(gathered, key=lambda g: g[1])
        for l in range(levels):
            off, size = oracle.level_offset(dim, t, l), oracle.level_size(dim, t, l)
            parts = []
            for r, slo, sn, _, ch in shards:
                c = chain if r == 0 else ch
                sd = (dim[0], dim[1], sn)
                parts.append(c[oracle.level_offset(sd, t, l): oracle.level_offset(sd, t, l) + oracle.level_size(sd, t, l)])
            ok = ok and np.array_equal(np.concatenate(parts), full[off: off + size])
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dim,t,cid", [((64, 64, 6), T.IMAGE_2D_ARRAY | T.RGBA8 | M, 3), ((32, 32, 2), T.IMAGE_CUBE_ARRAY | T.RGBA32F | M, 4)])
def test_sharded_chain_equals_whole_chain_gloo(oracle_mod, dim, t, cid):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dim, t, cid, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
