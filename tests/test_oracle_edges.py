"""Edge cases of the oracle pinning: the C restatement (oracle/bbx_oracle.c) against the
unmodified reference compiled in oracle/_ref, bit for bit, under settings and scenes the standard scene builders do
not reach: every PhysicsSettings field moved off its default, zero iterations, no friction, degenerate geometry
(coincident spheres, a sphere centred in a box, face-aligned boxes), static-only and empty worlds, a contact list that
overflows max_collisions (the reference truncates silently, ballbox.h:1988 / :2020 / :2052), and the dt guards of
PhysicsStep (:1791).  No GPU; skipped where oracle/_ref is not built (the restatement is then pinned by the golden
vectors alone)."""
import pytest

from ballbox_b200 import binding as bb
from tests import parity

DT = 1.0 / 60.0


def both(checkers):
    if "ref" not in checkers:
        pytest.skip("oracle/_ref not built here")
    return checkers["ref"], checkers["oracle"]


def lockstep(a, b, steps, dt=DT, what=""):
    for s in range(steps):
        a.step(dt); b.step(dt)
        mm = parity.compare_worlds(a, b)
        assert not mm, f"{what} step {s}: {mm[:3]}"


SETTINGS = [
    dict(solver_iterations=0),
    dict(solver_iterations=1),
    dict(solver_iterations=3, friction=0.0),
    dict(friction=0.9, restitution=0.0),
    dict(restitution=1.0, velocity_threshold=0.0),
    dict(baumgarte_bias=0.05, allowed_penetration=0.0),
    dict(baumgarte_bias=0.4, allowed_penetration=0.05),
    dict(linear_damping=0.5, angular_damping=0.9),
    dict(sleep_linear_threshold=5.0, sleep_angular_threshold=5.0, sleep_time_required=0.05),
    dict(manifold_max_contacts=1),
    dict(manifold_max_contacts=2, manifold_contact_tolerance=0.2, manifold_penetration_tolerance=0.2),
    dict(collision_epsilon=1e-2, vector_normalize_epsilon=1e-3, quaternion_epsilon=1e-3),
]


@pytest.mark.parametrize("over", SETTINGS, ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
def test_restatement_follows_every_setting(checkers, over):
    ref, ora = both(checkers)
    worlds = []
    for lib in (ref, ora):
        w = lib.build_scene(bb.SCENE_MIXED, 160, 77)
        for k, v in over.items():
            setattr(w.w.settings, k, v)
        worlds.append(w)
    lockstep(worlds[0], worlds[1], 45, what=str(over))
    assert len(worlds[0].constraints()) > 0 or over.get("sleep_time_required")


@pytest.mark.parametrize("gravity", [(0.0, 0.0, 0.0), (3.0, -9.8, -2.0), (0.0, 9.8, 0.0)])
def test_gravity_variants_on_the_sdf_scene(checkers, gravity):
    ref, ora = both(checkers)
    a, b = ref.build_scene(bb.SCENE_SDF, 120, 5), ora.build_scene(bb.SCENE_SDF, 120, 5)
    a.set_gravity(gravity); b.set_gravity(gravity)
    lockstep(a, b, 50, what=str(gravity))


def _degenerate(lib):
    w = lib.create_world(16, 16, 512)
    w.set_gravity((0, -9.8, 0))
    d = lib.dll
    # two coincident spheres (zero-length centre distance), a third touching them exactly at r1 + r2
    d.AddSphere(w.ptr, bb.Vec3(0, 1, 0), 0.5)
    d.AddSphere(w.ptr, bb.Vec3(0, 1, 0), 0.5)
    d.AddSphere(w.ptr, bb.Vec3(1.0, 1, 0), 0.5)
    # a sphere whose centre is inside a box, and one exactly on a box face
    d.AddBox(w.ptr, bb.Vec3(4, 1, 0), bb.Vec3(0.5, 0.5, 0.5), d.QuatIdentity())
    d.AddSphere(w.ptr, bb.Vec3(4, 1, 0), 0.3)
    d.AddSphere(w.ptr, bb.Vec3(4.5, 1.0, 0), 0.25)
    # face-aligned boxes (SAT ties between parallel axes), one stacked with zero gap, one rotated by 90 degrees
    d.AddBox(w.ptr, bb.Vec3(8, 0.5, 0), bb.Vec3(0.5, 0.5, 0.5), d.QuatIdentity())
    d.AddBox(w.ptr, bb.Vec3(8, 1.5, 0), bb.Vec3(0.5, 0.5, 0.5), d.QuatIdentity())
    d.AddBox(w.ptr, bb.Vec3(8.9, 0.5, 0), bb.Vec3(0.5, 0.5, 0.5), bb.Quat(*w.quat_axis_angle((0, 1, 0), 1.5707963)))
    # a static floor box and a static sphere
    d.AddBox(w.ptr, bb.Vec3(4, -0.5, 0), bb.Vec3(20, 0.5, 20), d.QuatIdentity())
    d.SetBoxStatic(w.ptr, 4, True)
    d.AddSphere(w.ptr, bb.Vec3(12, 0.5, 0), 0.5)
    d.SetSphereStatic(w.ptr, 5, True)
    d.AddSphere(w.ptr, bb.Vec3(12, 1.4, 0.05), 0.5)
    return w


def test_degenerate_geometry(checkers):
    ref, ora = both(checkers)
    a, b = _degenerate(ref), _degenerate(ora)
    lockstep(a, b, 120, what="degenerate")
    assert len(a.constraints()) > 0


def test_static_only_and_empty_worlds(checkers):
    ref, ora = both(checkers)
    for build in ("empty", "static"):
        ws = []
        for lib in (ref, ora):
            w = lib.create_world(4, 4, 64)
            if build == "static":
                d = lib.dll
                d.AddSphere(w.ptr, bb.Vec3(0, 0, 0), 1.0); d.SetSphereStatic(w.ptr, 0, True)
                d.AddBox(w.ptr, bb.Vec3(0.5, 0, 0), bb.Vec3(1, 1, 1), d.QuatIdentity()); d.SetBoxStatic(w.ptr, 0, True)
            ws.append(w)
        lockstep(ws[0], ws[1], 5, what=build)


def test_contact_list_overflow_truncates_like_the_reference(checkers):
    """max_collisions far below what the pile produces: the reference stops appending (:1988, :2020, :2052) and solves the
    truncated list; the restatement must keep exactly the same contacts in the same order."""
    ref, ora = both(checkers)
    a = ref.build_scene(bb.SCENE_MIXED, 300, 13, max_collisions=96)
    b = ora.build_scene(bb.SCENE_MIXED, 300, 13, max_collisions=96)
    hit_cap = False
    for s in range(60):
        a.step(DT); b.step(DT)
        mm = parity.compare_worlds(a, b)
        assert not mm, f"step {s}: {mm[:3]}"
        hit_cap = hit_cap or len(a.constraints()) == 96
    assert hit_cap, "the scene never filled its 96-contact list"


@pytest.mark.parametrize("dt", [0.0, -0.01, 5e-7, 1e-6, 1.0 / 240.0, 0.05])
def test_dt_guards_and_unusual_steps(checkers, dt):
    """dt <= 0 and dt < 1e-6 return before anything happens (:1791); other values step normally."""
    ref, ora = both(checkers)
    a, b = ref.build_scene(bb.SCENE_MIXED, 120, 3), ora.build_scene(bb.SCENE_MIXED, 120, 3)
    lockstep(a, b, 12, dt=dt, what=f"dt={dt}")


def test_user_forces_and_direct_state_edits(checkers):
    """AddForce / AddTorque on both body kinds, direct writes into the public velocity arrays, mass and static changes
    between steps (what a user of the header does between PhysicsStep calls)."""
    ref, ora = both(checkers)
    ws = [lib.build_scene(bb.SCENE_MIXED, 140, 21) for lib in (ref, ora)]
    for s in range(40):
        for lib, w in zip((ref, ora), ws):
            d = lib.dll
            if s % 3 == 0:
                d.AddSphereForce(w.ptr, 1, bb.Vec3(30.0, 5.0, -4.0)); d.AddSphereTorque(w.ptr, 2, bb.Vec3(0.5, 0.0, 1.5))
                d.AddBoxForce(w.ptr, 1, bb.Vec3(-20.0, 40.0, 3.0)); d.AddBoxTorque(w.ptr, 3, bb.Vec3(2.0, -1.0, 0.25))
            if s == 10:
                w.sphere_arrays()["vel"][4] = (1.5, 6.0, -2.0)
                w.box_arrays()["angvel"][2] = (0.0, 9.0, 0.0)
            if s == 15:
                d.SetSphereMass(w.ptr, 6, 25.0); d.SetBoxMass(w.ptr, 5, 0.2)
            if s == 20:
                d.SetSphereStatic(w.ptr, 7, True); d.SetBoxStatic(w.ptr, 6, True)
            w.step(DT)
        mm = parity.compare_worlds(ws[0], ws[1])
        assert not mm, f"step {s}: {mm[:3]}"
