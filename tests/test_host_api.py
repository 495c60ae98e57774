"""Host-side logic of the drop-in library against the reference: world creation,
body management, settings, error convention.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from ballbox_b200 import binding as bb


def libs(product, checkers):
    return [product] + list(checkers.values())


def build_by_api(lib):
    w = lib.create_world(4, 3, 64)
    d = lib.dll
    assert d.AddSphere(w.ptr, bb.Vec3(1, 2, 3), 0.5) == 0
    assert d.AddSphere(w.ptr, bb.Vec3(-1, 0, 3), 1.5) == 1
    d.SetSphereMass(w.ptr, 1, 2.5)
    d.SetSphereMass(w.ptr, 0, -1.0)          # ignored (ballbox.h:391)
    d.SetSphereMass(w.ptr, 9, 2.0)           # out of range: silently ignored
    q = d.QuatFromAxisAngle(bb.Vec3(1, 2, 3), 0.7)
    assert d.AddBox(w.ptr, bb.Vec3(0, 0, 0), bb.Vec3(0.5, 1.0, 1.5), q) == 0
    assert d.AddBox(w.ptr, bb.Vec3(0, -3, 0), bb.Vec3(5, 1, 5), d.QuatIdentity()) == 1
    d.SetBoxMass(w.ptr, 0, 3.0)
    d.AddBoxForce(w.ptr, 0, bb.Vec3(1, 0, 0)); d.AddBoxTorque(w.ptr, 0, bb.Vec3(0, 2, 0))
    d.AddSphereForce(w.ptr, 1, bb.Vec3(0, 1, 0)); d.AddSphereTorque(w.ptr, 1, bb.Vec3(0, 0, 4))
    d.SetBoxStatic(w.ptr, 1, True)           # zeroes v, w, F, T (ballbox.h:779-784)
    d.SetSphereStatic(w.ptr, 0, True)
    return w


def test_body_management_matches_reference(product, checkers):
    worlds = [build_by_api(l) for l in libs(product, checkers)]
    a = worlds[0]
    for b in worlds[1:]:
        for k in a.sphere_arrays():
            assert np.array_equal(a.sphere_arrays()[k].view(np.uint8), b.sphere_arrays()[k].view(np.uint8)), k
        for k in a.box_arrays():
            assert np.array_equal(a.box_arrays()[k].view(np.uint8), b.box_arrays()[k].view(np.uint8)), k


def test_capacity_and_error_convention(product):
    w = product.create_world(1, 1, 8)
    d = product.dll
    assert d.AddSphere(w.ptr, bb.Vec3(0, 0, 0), 1.0) == 0
    assert d.AddSphere(w.ptr, bb.Vec3(0, 0, 0), 1.0) == -1          # full -> -1 (ballbox.h:366)
    assert d.AddBox(w.ptr, bb.Vec3(0, 0, 0), bb.Vec3(1, 1, 1), d.QuatIdentity()) == 0
    assert d.AddBox(w.ptr, bb.Vec3(0, 0, 0), bb.Vec3(1, 1, 1), d.QuatIdentity()) == -1
    assert d.AddSphere(None, bb.Vec3(0, 0, 0), 1.0) == -1           # NULL world -> -1
    d.PhysicsStep(None, 0.01)                                       # NULL world: silent no-op
    d.SetSphereStatic(w.ptr, 5, True)                               # out of range: silent
    assert product.dll.BallboxGetLastError(w.ptr) is None


def test_defaults_match_reference(product, checkers):
    def settings(lib):
        w = lib.create_world(1, 1, 1)
        return bytes(w.w.settings), (w.w.gravity.x, w.w.gravity.y, w.w.gravity.z), w.w.dt, w.w.usesSDF
    ref = settings(product)
    for l in checkers.values():
        assert settings(l) == ref
    s = product.create_world(1, 1, 1).w.settings
    assert (s.solver_iterations, s.manifold_max_contacts) == (25, 8)          # real defaults (ballbox.h:2493, :2508)
    assert abs(s.linear_damping - 0.98) < 1e-7 and s.restitution == 0.0


def test_setters(product, checkers):
    for lib in libs(product, checkers):
        w = lib.create_world(1, 1, 1)
        d = lib.dll
        d.SetDamping(w.ptr, 0.9, 0.8); d.SetRestitution(w.ptr, 0.25); d.SetCorrectionParameters(w.ptr, 0.1, 0.02)
        d.SetSolverIterations(w.ptr, 7); d.SetManifoldSettings(w.ptr, 0.02, -0.1, 99)
        s = w.w.settings
        assert s.solver_iterations == 7 and s.manifold_max_contacts == 8      # clamped (ballbox.h:2543)
        assert abs(s.linear_damping - 0.9) < 1e-7 and abs(s.restitution - 0.25) < 1e-7


def test_scene_builder_is_backend_independent(product, checkers):
    for kind, n in ((bb.SCENE_BALLS, 50), (bb.SCENE_MIXED, 64), (bb.SCENE_SDF, 40), (bb.SCENE_RLWORLD, 0)):
        ws = [l.build_scene(kind, n, 77) for l in libs(product, checkers)]
        hs = {w.state_hash() for w in ws}
        assert len(hs) == 1, (kind, hs)
        assert len({(w.n_spheres, w.n_boxes) for w in ws}) == 1


def test_batch_worlds_are_slices_of_one_allocation(product):
    d = product.dll
    b = d.CreatePhysicsWorldBatch(3, 4, 2, 16)
    assert b and d.BatchWorldCount(b) == 3
    ws = [bb.World(product, d.BatchWorld(b, i), owned=False) for i in range(3)]
    for i, w in enumerate(ws):
        assert d.AddSphere(w.ptr, bb.Vec3(i, 0, 0), 1.0) == 0
    p0 = C.addressof(ws[0].w.spheres.contents.positions.contents)
    p1 = C.addressof(ws[1].w.spheres.contents.positions.contents)
    assert p1 - p0 == 4 * 12                                              # stride = max_spheres * sizeof(Vec3)
    assert d.BatchWorld(b, 3) is None or not d.BatchWorld(b, 3)
    d.DestroyPhysicsWorldBatch(b)
