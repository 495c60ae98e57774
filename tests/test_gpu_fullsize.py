"""Parity at BASELINE.json's full sizes AND at the state bench.py times.

The scenes are settled on the GPU (resident stepping), downloaded, and the public host
arrays are copied into a world of the CPU checker built from the same seed (same
capacities, masses, statics).  Both then take two more steps -- the product through the
drop-in call PhysicsStep(world, dt) with host arrays, the checker through the reference's
own stage functions behind the CPU grid (oracle/ref_shim.c RefPhysicsStepPruned, proven
identical to the all-pairs loops in tests/test_oracle.py) -- and everything is compared
bit for bit: body state and the full 104-byte constraint arrays in array order
(reference ballbox.h:1787-1817, :119-144).
"""
import numpy as np
import pytest

from ballbox_b200 import binding as bb
from tests import parity

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0

STATE_S = ("pos", "vel", "angvel", "sleeping", "timer", "force", "torque")
STATE_B = ("pos", "rot", "vel", "angvel", "sleeping", "timer", "force", "torque")


def copy_public_state(src, dst):
    """Host arrays of `src` -> host arrays of `dst` (same scene, same counts)."""
    assert src.n_spheres == dst.n_spheres and src.n_boxes == dst.n_boxes
    sa, da = src.sphere_arrays(), dst.sphere_arrays()
    for k in STATE_S:
        da[k][...] = sa[k]
    sa, da = src.box_arrays(), dst.box_arrays()
    for k in STATE_B:
        da[k][...] = sa[k]


def settle_then_compare(product, truth, kind, n, settle, steps=2):
    g = product.build_scene(kind, n, 20261019)
    g.set_sync_mode(bb.SYNC_RESIDENT)
    done = 0
    while done < settle:                         # in slices, so that a hang is bounded
        k = min(100, settle - done)
        g.step_n(DT, k); g.synchronize(); done += k
    g.download(bb.DL_BODIES)
    st = g.stats()
    assert st.overflow == 0
    c = truth.build_scene(kind, n, 20261019)
    copy_public_state(g, c)
    assert not parity.compare_worlds(g, c, constraints=False)
    g.set_sync_mode(bb.SYNC_COMPAT)              # from here on: the drop-in call with host arrays
    for s in range(steps):
        g.step(DT); c.step(DT, pruned=True)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"scene {kind} n={n}, step {s} after {settle} settle steps: {mm[:4]}"
    return g, c, g.stats()


def test_c3_full_size_settled_pile_against_pruned_reference(product, truth):
    """BASELINE config 3 (500k spheres + 500k boxes) in the state bench.py measures: settled 1500 steps."""
    g, c, st = settle_then_compare(product, truth, bb.SCENE_MIXED, 1_000_000, 1500)
    assert st.contacts > 5_000_000, st.contacts
    assert st.levels >= 40, st.levels
    assert len(g.constraints()) == st.contacts
    g.destroy(); c.destroy()


def test_c4_full_size_settled_sdf_terrain_against_pruned_reference(product, truth):
    """BASELINE config 4 at its stated size (200k bodies on the wavy SDF), settled 400 steps."""
    g, c, st = settle_then_compare(product, truth, bb.SCENE_SDF, 200_000, 400)
    assert st.contacts > 600_000, st.contacts
    g.destroy(); c.destroy()


def test_c2_full_size_settled_balls_against_pruned_reference(product, truth):
    """BASELINE config 2 (100k spheres in six walls), settled 300 steps."""
    g, c, st = settle_then_compare(product, truth, bb.SCENE_BALLS, 100_000, 300)
    assert st.contacts > 150_000, st.contacts
    g.destroy(); c.destroy()


def test_c5_full_size_batch_sampled_worlds_against_reference(product, truth):
    """BASELINE config 5: 4096 worlds x 256 bodies settled 300 steps on the GPU; 64 sampled worlds (first, last,
    every 65th) are copied into reference worlds; both take 2 steps; state and constraint arrays must be equal."""
    d = product.dll
    n_worlds = 4096
    s, b, cc = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, cc)
    assert batch
    for i in range(n_worlds):
        d.BbxSceneBuild(d.BatchWorld(batch, i), bb.SCENE_RLWORLD, 0, 7000 + i)
    for _ in range(3):
        assert d.PhysicsStepBatch(batch, DT, 100) == 0, d.BatchGetLastError(batch)
        assert d.BatchSynchronize(batch) == 0
    assert d.BatchDownload(batch, bb.DL_BODIES) == 0, d.BatchGetLastError(batch)
    sample = sorted(set(list(range(0, n_worlds, 65)) + [n_worlds - 1]))
    assert len(sample) >= 64 and sample[-1] == n_worlds - 1
    views = {i: bb.World(product, d.BatchWorld(batch, i), owned=False) for i in sample}
    refs = {}
    for i in sample:
        r = truth.build_scene(bb.SCENE_RLWORLD, 0, 7000 + i)
        copy_public_state(views[i], r)
        refs[i] = r
    for step in range(2):
        assert d.PhysicsStepBatch(batch, DT, 1) == 0, d.BatchGetLastError(batch)
        assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
        for i in sample:
            refs[i].step(DT)
            mm = parity.compare_worlds(views[i], refs[i])
            assert not mm, f"world {i} step {step}: {mm[:3]}"
    st = bb.BallboxStats(); d.BatchGetStats(batch, bb.C.byref(st))
    assert st.overflow == 0 and st.contacts > 3_000_000, st.contacts
    d.DestroyPhysicsWorldBatch(batch)


@pytest.mark.parametrize("total,ranks", [(64, 2), (50, 4)])
def test_partitioned_batch_is_world_for_world_identical_to_the_single_batch(product, total, ranks):
    """SURVEY 8e on hardware: ONE batch of `total` worlds stepped as a whole, and the same worlds stepped
    as the `ranks` shards of ballbox_b200.sharding (what bench.py --gpus N builds per GPU; here one after
    the other on cuda:0).  Every world's state hash and the order-independent checksum must agree."""
    from ballbox_b200 import sharding
    whole = sharding.WorldShard(product, sharding.strong_world_ids(0, 1, total), bb.SCENE_RLWORLD, device=0)
    whole.step(DT, 120)
    ref = whole.world_hashes()
    whole.destroy()
    merged = {}
    for r in range(ranks):
        sh = sharding.WorldShard(product, sharding.strong_world_ids(r, ranks, total), bb.SCENE_RLWORLD, device=0)
        sh.step(DT, 120)
        merged.update(sh.world_hashes())
        sh.destroy()
    assert merged == ref
    assert sharding.combine_hashes(merged.values()) == sharding.combine_hashes(ref.values())
