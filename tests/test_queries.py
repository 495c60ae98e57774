"""Standalone queries and helpers of the public API (reference ballbox.h:231-239, :262-297) against the
compiled reference on the same inputs.  No GPU: the product's queries are host functions over the SAME
__host__ __device__ geometry code the narrowphase kernels run (bbx_geom.cuh), so these tests also pin the
device geometry on the CPU.  Bit-exact: every float is compared through its bytes."""
import ctypes as C

import numpy as np
import pytest

from ballbox_b200 import binding as bb

V, Q, K = bb.Vec3, bb.Quat, bb.CollisionContact


class Manifold(C.Structure):            # ballbox.h:105-116
    _fields_ = [("points", V * 8), ("penetrations", C.c_float * 8), ("normal", V), ("point_count", C.c_int),
                ("bodyA_type", C.c_int), ("bodyA_index", C.c_int), ("bodyB_type", C.c_int), ("bodyB_index", C.c_int)]


def sigs(d):
    d.CheckSphereCollisionWithData.argtypes = [V, C.c_float, V, C.c_float, C.POINTER(K)]; d.CheckSphereCollisionWithData.restype = C.c_bool
    d.CheckSphereCollision.argtypes = [V, C.c_float, V, C.c_float]; d.CheckSphereCollision.restype = C.c_bool
    d.CheckSphereBoxCollisionWithData.argtypes = [V, C.c_float, V, V, Q, C.POINTER(K)]; d.CheckSphereBoxCollisionWithData.restype = C.c_bool
    d.CheckSphereBoxCollision.argtypes = [V, C.c_float, V, V, Q]; d.CheckSphereBoxCollision.restype = C.c_bool
    d.CheckBoxCollisionWithData.argtypes = [V, V, Q, V, V, Q, C.POINTER(K)]; d.CheckBoxCollisionWithData.restype = C.c_bool
    d.CheckBoxCollision.argtypes = [V, V, Q, V, V, Q]; d.CheckBoxCollision.restype = C.c_bool
    d.CheckSphereSystemCollisions.argtypes = [C.c_void_p, C.c_int]; d.CheckSphereSystemCollisions.restype = C.c_bool
    d.CheckBoxSystemCollisions.argtypes = [C.c_void_p, C.c_int]; d.CheckBoxSystemCollisions.restype = C.c_bool
    d.GenerateBoxBoxManifold.argtypes = [bb.WP, C.c_int, C.c_int, C.POINTER(Manifold)]; d.GenerateBoxBoxManifold.restype = None
    d.SelectOptimalContactPoints.argtypes = [C.POINTER(V), C.POINTER(C.c_float), C.c_int, C.POINTER(Manifold), V, C.c_int]
    d.SelectOptimalContactPoints.restype = None
    d.AddManifoldToConstraints.argtypes = [bb.WP, C.POINTER(Manifold)]; d.AddManifoldToConstraints.restype = None
    d.SDF_EstimateNormal.argtypes = [bb.WP, V]; d.SDF_EstimateNormal.restype = V
    d.CheckBoxSDFCollision.argtypes = [bb.WP, V, V, Q, C.POINTER(K)]; d.CheckBoxSDFCollision.restype = C.c_bool
    d.GenerateBoxSDFContacts.argtypes = [bb.WP, C.c_int]; d.GenerateBoxSDFContacts.restype = None
    d.GetBodyVelocityAtPoint.argtypes = [bb.WP, C.c_int, C.c_int, V]; d.GetBodyVelocityAtPoint.restype = V
    d.CalculateEffectiveMass.argtypes = [bb.WP, C.POINTER(K)]; d.CalculateEffectiveMass.restype = C.c_float
    d.ApplyImpulse.argtypes = [bb.WP, C.c_int, C.c_int, V, V]; d.ApplyImpulse.restype = None
    d.SolvePositionConstraints.argtypes = [bb.WP]; d.SolvePositionConstraints.restype = None
    return d


@pytest.fixture(scope="module")
def pair(product, checkers):
    if "ref" not in checkers:
        pytest.skip("the compiled reference (oracle/_ref) is not available")
    return sigs(product.dll), sigs(checkers["ref"].dll), product, checkers["ref"]


def raw(x):
    return bytes(memoryview(x))


def rnd_quat(r, unit=True):
    q = r.normal(size=4)
    if unit:
        q /= np.linalg.norm(q)
    return Q(*[float(np.float32(v)) for v in q])


def vec(a):
    return V(*[float(np.float32(v)) for v in a])


def contact_bits(k):
    return raw(k.contact_point) + raw(k.contact_normal) + np.float32(k.penetration).tobytes()


def test_sphere_sphere_and_sphere_box(pair):
    p, t, _, _ = pair
    r = np.random.default_rng(1)
    hits = 0
    for i in range(3000):
        a, b = r.uniform(-2, 2, 3), r.uniform(-2, 2, 3)
        if i % 50 == 0:
            b = a.copy()                                   # coincident centres: normal (1,0,0)
        ra, rb = float(np.float32(r.uniform(0.2, 1.5))), float(np.float32(r.uniform(0.2, 1.5)))
        ka, kb = K(), K()
        ha, hb = p.CheckSphereCollisionWithData(vec(a), ra, vec(b), rb, ka), t.CheckSphereCollisionWithData(vec(a), ra, vec(b), rb, kb)
        assert ha == hb == p.CheckSphereCollision(vec(a), ra, vec(b), rb)
        if ha:
            hits += 1
            assert contact_bits(ka) == contact_bits(kb), i
    assert hits > 500
    hits = 0
    for i in range(3000):
        sp, bp = r.uniform(-2, 2, 3), r.uniform(-2, 2, 3)
        if i % 20 == 0:
            sp = bp + r.uniform(-0.2, 0.2, 3)             # sphere centre inside the box
        sr = float(np.float32(r.uniform(0.2, 1.2)))
        half = vec(r.uniform(0.2, 1.5, 3)); q = rnd_quat(r, unit=(i % 7 != 0))
        ka, kb = K(), K()
        ha = p.CheckSphereBoxCollisionWithData(vec(sp), sr, vec(bp), half, q, ka)
        hb = t.CheckSphereBoxCollisionWithData(vec(sp), sr, vec(bp), half, q, kb)
        assert ha == hb == p.CheckSphereBoxCollision(vec(sp), sr, vec(bp), half, q)
        if ha:
            hits += 1
            assert contact_bits(ka) == contact_bits(kb), i
    assert hits > 500


def test_box_box_sat(pair):
    p, t, _, _ = pair
    r = np.random.default_rng(2)
    hits = 0
    for i in range(4000):
        pa, pb = r.uniform(-1.2, 1.2, 3), r.uniform(-1.2, 1.2, 3)
        ha_, hb_ = vec(r.uniform(0.2, 1.0, 3)), vec(r.uniform(0.2, 1.0, 3))
        qa = rnd_quat(r) if i % 5 else Q(0, 0, 0, 1)       # axis-aligned boxes: degenerate cross axes are dropped
        qb = rnd_quat(r) if i % 3 else Q(0, 0, 0, 1)
        ka, kb = K(), K()
        ha = p.CheckBoxCollisionWithData(vec(pa), ha_, qa, vec(pb), hb_, qb, ka)
        hb = t.CheckBoxCollisionWithData(vec(pa), ha_, qa, vec(pb), hb_, qb, kb)
        assert ha == hb == p.CheckBoxCollision(vec(pa), ha_, qa, vec(pb), hb_, qb), i
        if ha:
            hits += 1
            assert contact_bits(ka) == contact_bits(kb), i
    assert hits > 1000


def box_world(lib, r, n=36, sdf=None):
    w = lib.create_world(8, n, 4096)
    for i in range(n):
        lib.dll.AddBox(w.ptr, vec(r.uniform(-1.6, 1.6, 3)), vec(r.uniform(0.25, 0.8, 3)), rnd_quat(r))
    for i in range(8):
        lib.dll.AddSphere(w.ptr, vec(r.uniform(-1.5, 1.5, 3)), float(np.float32(r.uniform(0.2, 0.7))))
    if sdf is not None:
        lib.dll.UseSDF(w.ptr, C.cast(getattr(lib.dll, sdf), C.c_void_p))
    return w


def test_manifolds_and_system_queries(pair):
    p, t, plib, tlib = pair
    wa, wb = box_world(plib, np.random.default_rng(3)), box_world(tlib, np.random.default_rng(3))
    n = wa.n_boxes
    sizes = {}
    for i in range(n):
        assert p.CheckBoxSystemCollisions(C.cast(wa.w.boxes, C.c_void_p), i) == t.CheckBoxSystemCollisions(C.cast(wb.w.boxes, C.c_void_p), i)
        for j in range(n):
            ma, mb = Manifold(), Manifold()
            p.GenerateBoxBoxManifold(wa.ptr, i, j, ma); t.GenerateBoxBoxManifold(wb.ptr, i, j, mb)
            if i == j:
                continue
            assert ma.point_count == mb.point_count, (i, j)
            k = ma.point_count
            sizes[k] = sizes.get(k, 0) + 1
            if k:
                assert raw(ma.normal) == raw(mb.normal) and raw(ma.points)[:12 * k] == raw(mb.points)[:12 * k], (i, j)
                assert raw(ma.penetrations)[:4 * k] == raw(mb.penetrations)[:4 * k], (i, j)
                assert (ma.bodyA_index, ma.bodyB_index) == (mb.bodyA_index, mb.bodyB_index) == (i, j)
    assert sum(v for k, v in sizes.items() if k >= 3) > 50, sizes
    for i in range(wa.n_spheres):
        assert p.CheckSphereSystemCollisions(C.cast(wa.w.spheres, C.c_void_p), i) == t.CheckSphereSystemCollisions(C.cast(wb.w.spheres, C.c_void_p), i)
    # a manifold becomes constraint records
    ma, mb = Manifold(), Manifold()
    done = 0
    for i in range(n):
        for j in range(i + 1, n):
            p.GenerateBoxBoxManifold(wa.ptr, i, j, ma); t.GenerateBoxBoxManifold(wb.ptr, i, j, mb)
            if ma.point_count:
                p.AddManifoldToConstraints(wa.ptr, ma); t.AddManifoldToConstraints(wb.ptr, mb); done += 1
    ca, cb = wa.constraints(), wb.constraints()
    assert len(ca) == len(cb) > done
    for f in ("a_type", "a_index", "b_type", "b_index", "point", "normal", "penetration", "restitution", "friction"):
        assert ca[f].tobytes() == cb[f].tobytes(), f
    # ... and the position pass moves the same bodies by the same amounts
    p.SolvePositionConstraints(wa.ptr); t.SolvePositionConstraints(wb.ptr)
    assert wa.box_arrays()["pos"].tobytes() == wb.box_arrays()["pos"].tobytes()


def test_select_optimal_contact_points(pair):
    p, t, _, _ = pair
    r = np.random.default_rng(4)
    for trial in range(400):
        n = int(r.integers(1, 25)); mx = int(r.integers(1, 9))
        pts = (V * n)(*[vec(r.uniform(-1, 1, 3)) for _ in range(n)])
        pens = (C.c_float * n)(*[float(np.float32(x)) for x in r.uniform(-0.2, 0.5, n)])
        if trial % 9 == 0:                                  # ties in depth and in distance
            for k in range(1, n):
                pens[k] = pens[0]
        if mx < n <= 8 or n <= mx or n <= 2:               # inside the struct's 8 slots the reference is well defined
            ma, mb = Manifold(), Manifold()
            p.SelectOptimalContactPoints(pts, pens, n, ma, V(0, 1, 0), mx); t.SelectOptimalContactPoints(pts, pens, n, mb, V(0, 1, 0), mx)
            if n <= 8 or mx < n:
                assert ma.point_count == mb.point_count, (n, mx)
                k = ma.point_count
                assert raw(ma.points)[:12 * k] == raw(mb.points)[:12 * k] and raw(ma.penetrations)[:4 * k] == raw(mb.penetrations)[:4 * k], (n, mx)
        else:
            ma, mb = Manifold(), Manifold()
            p.SelectOptimalContactPoints(pts, pens, n, ma, V(0, 1, 0), mx); t.SelectOptimalContactPoints(pts, pens, n, mb, V(0, 1, 0), mx)
            assert ma.point_count == mb.point_count == min(mx, n)
            k = ma.point_count
            assert raw(ma.points)[:12 * k] == raw(mb.points)[:12 * k]


def test_sdf_helpers(pair):
    p, t, plib, tlib = pair
    for sdf in ("BbxHostSdfWavyDet", "BbxHostSdfPlane0"):
        wa, wb = box_world(plib, np.random.default_rng(5), sdf=sdf), box_world(tlib, np.random.default_rng(5), sdf=sdf)
        r = np.random.default_rng(6)
        for _ in range(300):
            x = vec(r.uniform(-3, 3, 3))
            assert raw(p.SDF_EstimateNormal(wa.ptr, x)) == raw(t.SDF_EstimateNormal(wb.ptr, x))
        hits = 0
        for i in range(wa.n_boxes):
            bx = wa.box_arrays()
            pos, half, rot = V(*map(float, bx["pos"][i])), V(*map(float, bx["size"][i])), Q(*map(float, bx["rot"][i]))
            ka, kb = K(), K()
            ha, hb = p.CheckBoxSDFCollision(wa.ptr, pos, half, rot, ka), t.CheckBoxSDFCollision(wb.ptr, pos, half, rot, kb)
            assert ha == hb
            if ha:
                hits += 1
                assert contact_bits(ka) == contact_bits(kb)
            p.GenerateBoxSDFContacts(wa.ptr, i); t.GenerateBoxSDFContacts(wb.ptr, i)
        ca, cb = wa.collisions(), wb.collisions()
        assert len(ca) == len(cb) > 10 and hits > 3
        assert ca.tobytes() == cb.tobytes()


def test_legacy_impulse_helpers(pair):
    p, t, plib, tlib = pair
    wa, wb = box_world(plib, np.random.default_rng(7)), box_world(tlib, np.random.default_rng(7))
    for lib, w in ((plib, wa), (tlib, wb)):
        lib.dll.SetBoxStatic(w.ptr, 3, True); lib.dll.SetSphereStatic(w.ptr, 1, True)
        lib.dll.SetBoxMass(w.ptr, 5, 2.5); lib.dll.SetSphereMass(w.ptr, 2, 0.7)
    r = np.random.default_rng(8)
    for trial in range(400):
        ty = int(r.integers(0, 3)); idx = int(r.integers(0, 8)) if ty == 0 else int(r.integers(0, wa.n_boxes)) if ty == 1 else 0
        imp, at = vec(r.uniform(-2, 2, 3)), vec(r.uniform(-2, 2, 3))
        p.ApplyImpulse(wa.ptr, ty, idx, imp, at); t.ApplyImpulse(wb.ptr, ty, idx, imp, at)
        assert raw(p.GetBodyVelocityAtPoint(wa.ptr, ty, idx, at)) == raw(t.GetBodyVelocityAtPoint(wb.ptr, ty, idx, at))
        k = K()
        k.bodyA_type = int(r.integers(0, 2)); k.bodyA_index = int(r.integers(0, 8))
        k.bodyB_type = int(r.integers(0, 3)); k.bodyB_index = int(r.integers(0, 8)) if k.bodyB_type < 2 else 0
        k.contact_point = at; n = r.normal(size=3); k.contact_normal = vec(n / np.linalg.norm(n))
        assert np.float32(p.CalculateEffectiveMass(wa.ptr, k)).tobytes() == np.float32(t.CalculateEffectiveMass(wb.ptr, k)).tobytes()
    for name in ("vel", "angvel", "sleeping", "timer"):
        assert wa.sphere_arrays()[name].tobytes() == wb.sphere_arrays()[name].tobytes(), name
        assert wa.box_arrays()[name].tobytes() == wb.box_arrays()[name].tobytes(), name
