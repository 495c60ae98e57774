"""Host-side logic of the multi-GPU path: which worlds a rank owns, how a rank's
shard is built and stepped, and the one collective of the design (a sum of
statistics after the last step).

Independent worlds never interact (the reference has no global state, SURVEY.md
section 8e), so the step path has no collective; `reduce_stats` is the only place
`torch.distributed` is used, with NCCL on GPUs and gloo in the CPU tests.

Two partitions of a batch of worlds:
  * strong scaling (BASELINE.json config 5, "4096 worlds batched across 1/2/4/8
    B200"): ONE batch of `total` worlds, world w lives on rank w // ceil(total / N)
    (`strong_world_ids`); the work per GPU shrinks as N grows.
  * weak scaling: every rank owns `worlds_per_rank` worlds (`rank_world_ids`).
Worlds are seeded by their GLOBAL id, so a world's trajectory does not depend on N:
the order-independent checksum of all worlds' state hashes must be identical for
every N (checked on CPU in tests/test_dist_cpu.py, on GPUs from bench.py's lines)."""
from __future__ import annotations

import ctypes as C

BASE_SEED = 7000
MASK64 = (1 << 64) - 1


def rank_world_ids(rank: int, world_size: int, worlds_per_rank: int):
    """Global ids of the worlds stepped by `rank` (contiguous block, weak scaling)."""
    if not (0 <= rank < world_size) or worlds_per_rank < 0:
        raise ValueError("bad rank / world_size / worlds_per_rank")
    lo = rank * worlds_per_rank
    return list(range(lo, lo + worlds_per_rank))


def strong_world_ids(rank: int, world_size: int, total_worlds: int):
    """Global ids owned by `rank` when ONE batch of `total_worlds` is partitioned over
    `world_size` ranks: world w -> rank w // ceil(total / N).  Ragged totals leave the
    last ranks with fewer (possibly zero) worlds; the union is exactly range(total)."""
    if not (0 <= rank < world_size) or total_worlds < 0:
        raise ValueError("bad rank / world_size / total_worlds")
    per = -(-total_worlds // world_size) if total_worlds else 0
    lo = min(rank * per, total_worlds)
    return list(range(lo, min(lo + per, total_worlds)))


def owner_of(world_id: int, world_size: int, total_worlds: int) -> int:
    per = -(-total_worlds // world_size)
    return world_id // per


def world_seed(global_world_id: int) -> int:
    return BASE_SEED + global_world_id


def combine_hashes(hashes) -> int:
    """Order-independent 64-bit checksum of checksums: the sum modulo 2^64."""
    s = 0
    for h in hashes:
        s = (s + int(h)) & MASK64
    return s


STAT_KEYS = ("body_steps", "contacts", "worlds", "hash_w0", "hash_w1", "hash_w2", "hash_w3")


def pack_stats(body_steps: int, contacts: int, worlds: int, hash_sum: int):
    """Statistics vector; the 64-bit checksum of checksums travels as four 16-bit words so that
    it survives a float64 SUM all-reduce exactly for any number of ranks below 2^37."""
    h = hash_sum & MASK64
    return [float(body_steps), float(contacts), float(worlds)] + [float((h >> (16 * i)) & 0xffff) for i in range(4)]


def unpack_hash(stats: dict) -> int:
    """Recombine the four summed words modulo 2^64 (carries propagate through the shifts)."""
    return sum(int(stats["hash_w%d" % i]) << (16 * i) for i in range(4)) & MASK64


def reduce_stats(vec, dist=None, device="cpu"):
    """Sum the statistics vector over ranks (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(zip(STAT_KEYS, vec))
    import torch
    t = torch.tensor(vec, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dict(zip(STAT_KEYS, t.tolist()))


class WorldShard:
    """The worlds one rank owns.  With the product library they form ONE PhysicsWorldBatch (device
    resident, stepped by PhysicsStepBatch); with a CPU checker (tests only) they are plain worlds
    stepped one after the other.  bench.py and the distributed tests both go through this class, so
    they exercise the same partition / seeding / checksum code."""

    def __init__(self, lib, world_ids, scene_kind, device=None, stream=None):
        from . import binding as bb
        self.lib, self.ids, self.bb = lib, list(world_ids), bb
        self.batch, self.worlds = None, []
        if not self.ids:
            return
        if not lib.product:
            self.worlds = [lib.build_scene(scene_kind, 0, world_seed(gid)) for gid in self.ids]
            return
        d = lib.dll
        s, b, c = lib.scene_capacity(scene_kind, 0)
        self.batch = d.CreatePhysicsWorldBatch(len(self.ids), s, b, c)
        if not self.batch:
            raise RuntimeError("CreatePhysicsWorldBatch failed")
        for slot, gid in enumerate(self.ids):
            d.BbxSceneBuild(d.BatchWorld(self.batch, slot), scene_kind, 0, world_seed(gid))
        if device is not None:
            d.BatchSetDevice(self.batch, device)
        if stream is not None:
            d.BatchSetStream(self.batch, C.c_void_p(stream))
        self._chk(d.BatchUpload(self.batch))

    def _chk(self, rc=0):
        e = self.lib.dll.BatchGetLastError(self.batch)
        if e:
            raise RuntimeError(e.decode())
        if rc:
            raise RuntimeError("batch call failed with code %d" % rc)

    def step(self, dt, n, sync=True):
        """n steps of every world of the shard (queued on the batch's stream; `sync` waits and checks)."""
        if n <= 0:
            return
        for w in self.worlds:
            for _ in range(n):
                w.step(dt)
        if self.batch:
            rc = self.lib.dll.PhysicsStepBatch(self.batch, dt, n)
            if sync:
                self.lib.dll.BatchSynchronize(self.batch)
            self._chk(rc)

    def synchronize(self):
        if self.batch:
            self.lib.dll.BatchSynchronize(self.batch); self._chk()

    def stats(self):
        st = self.bb.BallboxStats()
        if self.batch:
            self.lib.dll.BatchGetStats(self.batch, C.byref(st))
        return st

    def n_bodies(self):
        if self.batch:
            st = self.stats()
            return st.spheres + st.boxes
        return sum(w.n_spheres + w.n_boxes for w in self.worlds)

    def n_contacts(self):
        if self.batch:
            return self.stats().contacts
        return sum(len(w.constraints()) for w in self.worlds)

    def world_hashes(self):
        """{global id: FNV-1a of the world's full mutable state} (after a body download on the product)."""
        if self.batch:
            d = self.lib.dll
            self._chk(d.BatchDownload(self.batch, self.bb.DL_BODIES))
            return {gid: int(d.BbxStateHash(d.BatchWorld(self.batch, slot))) for slot, gid in enumerate(self.ids)}
        return {gid: w.state_hash() for gid, w in zip(self.ids, self.worlds)}

    def destroy(self):
        if self.batch:
            self.lib.dll.DestroyPhysicsWorldBatch(self.batch)
            self.batch = None
        for w in self.worlds:
            w.destroy()
        self.worlds = []
