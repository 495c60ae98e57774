"""The N > 1 path on CPU: world_size 2, gloo.  Each rank builds its shard of ONE batch of
independent worlds through ballbox_b200.sharding (the code bench.py --gpus N runs, with the
CPU checker standing in for the device), steps it, and the statistics are all-reduced; the
union must equal the single-process run world for world and the order-independent checksum
must not depend on the number of ranks (worlds are independent, so this is exact)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ballbox_b200 import binding as bb, sharding

TOTAL_WORLDS, STEPS = 5, 25          # ragged on purpose: ranks own 3 + 2 worlds


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _step_worlds(ids):
    from tests import oracles
    lib = oracles.load_all()["oracle"]
    shard = sharding.WorldShard(lib, ids, bb.SCENE_RLWORLD)
    shard.step(1 / 60, STEPS)
    per_world = shard.world_hashes()
    out = (shard.n_bodies() * STEPS, shard.n_contacts(), sharding.combine_hashes(per_world.values()), per_world)
    shard.destroy()
    return out


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    ids = sharding.strong_world_ids(rank, world_size, TOTAL_WORLDS)
    bs, ct, hs, per_world = _step_worlds(ids)
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


@pytest.mark.parametrize("total,n", [(4096, 1), (4096, 2), (4096, 4), (4096, 8), (10, 4), (3, 8), (0, 2)])
def test_strong_partition_covers_every_world_once(total, n):
    seen = []
    for r in range(n):
        ids = sharding.strong_world_ids(r, n, total)
        assert all(sharding.owner_of(w, n, total) == r for w in ids)
        seen += ids
    assert seen == list(range(total))
    if total == 4096:
        assert all(len(sharding.strong_world_ids(r, n, total)) == 4096 // n for r in range(n))


def test_checksum_words_survive_a_float_sum():
    import random
    rnd = random.Random(7)
    hashes = [rnd.getrandbits(64) for _ in range(64)]
    parts = [hashes[i::8] for i in range(8)]                      # 8 "ranks"
    vecs = [sharding.pack_stats(1, 2, len(p), sharding.combine_hashes(p)) for p in parts]
    summed = [sum(col) for col in zip(*vecs)]
    st = dict(zip(sharding.STAT_KEYS, summed))
    assert sharding.unpack_hash(st) == sharding.combine_hashes(hashes)
    assert st["worlds"] == 64


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
    bs, ct, hs, per_world = _step_worlds(list(range(TOTAL_WORLDS)))
    merged = {}
    for g in gathered:
        merged.update(g)
    assert merged == per_world                                   # every world bit-identical to the 1-process run
    assert total["body_steps"] == bs and total["contacts"] == ct and total["worlds"] == TOTAL_WORLDS
    assert sharding.unpack_hash(total) == hs                      # checksum of checksums does not depend on N
