"""The public stage functions (reference ballbox.h:217-229) as GPU-backed drop-ins: calling them one
by one in PhysicsStep's order (:1797-1816) must give what the reference gives after every single
stage, including when the user edits the contact list between stages."""
import numpy as np
import pytest

from ballbox_b200 import binding as bb
from tests import parity

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


def ref_with_stages(checkers):
    ref = checkers.get("ref")
    if ref is None or not ref.has_stages:
        pytest.skip("oracle/_ref (which exports the reference's stage functions) is not available")
    return ref


def same_lists(g, c, what):
    a, b = (g.collisions(), c.collisions()) if what == "collisions" else (g.constraints(), c.constraints())
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)}"
    m = parity.first_mismatch(a, b)
    assert m is None, f"{what}: {m}"


def staged_step(worlds, edit=None, sweeps=None, check=None):
    """one PhysicsStep spelled out in stages on every world of `worlds`"""
    for name in ("ApplyForces", "IntegrateVelocities", "CollectCollisions"):
        for w in worlds:
            w.w.dt = DT
            w.stage(name)
        if check:
            check(name)
    if edit:
        for w in worlds:
            edit(w)
    if sweeps is None:
        for w in worlds:
            w.stage("SolveConstraints")
        if check:
            check("SolveConstraints")
    else:
        for name in ["GenerateContactConstraints", "PreSolveConstraints"] + ["SolveVelocityConstraints"] * sweeps:
            for w in worlds:
                w.stage(name)
            if check:
                check(name)
    for name in ("IntegratePositions", "UpdateSleepStates", "CleanupPhysics"):
        for w in worlds:
            w.stage(name)
        if check:
            check(name)


@pytest.mark.parametrize("kind,n", [(bb.SCENE_MIXED, 300), (bb.SCENE_SDF, 200), (bb.SCENE_BALLS, 400)])
def test_every_stage_matches_the_reference(product, checkers, kind, n):
    ref = ref_with_stages(checkers)
    g, c = product.build_scene(kind, n, 61), ref.build_scene(kind, n, 61)
    for w, lib in ((g, product), (c, ref)):
        w.w.settings.solver_iterations = 4
        lib.dll.AddSphereForce(w.ptr, 2, bb.Vec3(3, 1, 0))

    def check(stage):
        mm = parity.compare_states(g.snapshot(), c.snapshot())
        assert not mm, f"after {stage}: {mm[:3]}"
        if stage == "CollectCollisions":
            same_lists(g, c, "collisions")
        if stage in ("PreSolveConstraints", "SolveVelocityConstraints", "SolveConstraints"):
            same_lists(g, c, "constraints")
    for step in range(12):
        staged_step([g, c], sweeps=4 if step % 2 else None, check=check)
    assert len(g.constraints()) > 100


def test_staged_equals_physics_step(product):
    a, b = product.build_scene(bb.SCENE_MIXED, 400, 9), product.build_scene(bb.SCENE_MIXED, 400, 9)
    for _ in range(15):
        a.step(DT)
        staged_step([b])
    assert not parity.compare_worlds(a, b)


def test_contact_list_edited_between_stages(product, checkers):
    """the solver stages read the HOST list in array order: reverse it and drop every 5th contact"""
    ref = ref_with_stages(checkers)
    g, c = product.build_scene(bb.SCENE_MIXED, 250, 5), ref.build_scene(bb.SCENE_MIXED, 250, 5)
    for w in (g, c):
        w.w.settings.solver_iterations = 3

    def edit(w):
        cc = w.w.collisions.contents
        arr = w.collisions()
        keep = arr[::-1].copy()
        keep = keep[np.arange(len(keep)) % 5 != 4]
        arr[:len(keep)] = keep
        cc.count = len(keep)

    def check(stage):
        mm = parity.compare_states(g.snapshot(), c.snapshot())
        assert not mm, f"after {stage}: {mm[:3]}"
        if stage in ("PreSolveConstraints", "SolveVelocityConstraints", "SolveConstraints"):
            same_lists(g, c, "constraints")
    for step in range(8):
        staged_step([g, c], edit=edit, sweeps=3 if step % 2 else None, check=check)


def test_collect_truncates_at_capacity_like_the_reference(product, checkers):
    ref = ref_with_stages(checkers)
    g = product.build_scene(bb.SCENE_BALLS, 300, 3, max_collisions=40)
    c = ref.build_scene(bb.SCENE_BALLS, 300, 3, max_collisions=40)
    for w in (g, c):
        w.w.dt = DT
        w.stage("CollectCollisions")
    assert len(c.collisions()) == 40
    same_lists(g, c, "collisions")
