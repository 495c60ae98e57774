"""The N > 1 path on CPU: world_size 2, gloo.  Each rank steps its shard of independent
worlds (with the CPU checker standing in for the device) and the statistics are
all-reduced; the union must equal the single-process run world for world (worlds are
independent, so this is exact)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ballbox_b200 import binding as bb, sharding

WORLDS_PER_RANK, STEPS = 3, 25


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _step_worlds(ids):
    from tests import oracles
    lib = oracles.load_all()["oracle"]
    body_steps = contacts = 0
    hsum = 0
    per_world = {}
    for gid in ids:
        w = lib.build_scene(bb.SCENE_RLWORLD, 0, sharding.world_seed(gid))
        for _ in range(STEPS):
            w.step(1 / 60)
        h = w.state_hash()
        per_world[gid] = h
        hsum = (hsum + h) & ((1 << 64) - 1)
        body_steps += (w.n_spheres + w.n_boxes) * STEPS
        contacts += len(w.constraints())
    return body_steps, contacts, hsum, per_world


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    ids = sharding.rank_world_ids(rank, world_size, WORLDS_PER_RANK)
    bs, ct, hs, per_world = _step_worlds(ids)
    # checksums are summed in two 32-bit halves; carry is re-applied by the caller
    total = sharding.reduce_stats(sharding.pack_stats(bs, ct, len(ids), hs), dist)
    gathered = [None] * world_size
    dist.all_gather_object(gathered, per_world)
    if rank == 0:
        out.put((total, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_world_ids_partition():
    seen = []
    for r in range(4):
        seen += sharding.rank_world_ids(r, 4, 5)
    assert seen == list(range(20))
    with pytest.raises(ValueError):
        sharding.rank_world_ids(4, 4, 5)


def test_two_ranks_equal_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    total, gathered = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bs, ct, hs, per_world = _step_worlds(list(range(2 * WORLDS_PER_RANK)))
    merged = {}
    for g in gathered:
        merged.update(g)
    assert merged == per_world                                   # every world bit-identical to the 1-process run
    assert total["body_steps"] == bs and total["contacts"] == ct and total["worlds"] == 2 * WORLDS_PER_RANK
    # checksum of checksums: the two 32-bit halves were summed separately; recombine modulo 2^64
    combined = (int(total["hash_lo"]) + (int(total["hash_hi"]) << 32)) & ((1 << 64) - 1)
    assert combined == hs
