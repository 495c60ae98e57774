"""Joints and springs on the GPU (extension, include/ballbox_ext.h) against their sequential definition in
oracle/bbx_oracle.c: bit-exact, every step, through the public calls.  The reference has no joints (README.md:25), so the
checker here is always the restatement, never oracle/_ref."""
import numpy as np
import pytest

from ballbox_b200 import binding as bb
from tests import joint_scenes as js, parity

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


@pytest.fixture(scope="module")
def oracle(checkers):
    return checkers["oracle"]


def pair(product, oracle, scene, **kw):
    g = scene(product.create_world(*js.capacity(scene.__name__)), **kw)
    c = scene(oracle.create_world(*js.capacity(scene.__name__)), **kw)
    return g, c


def lockstep(g, c, steps, what=""):
    for s in range(steps):
        g.step(DT); c.step(DT)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"{what} step {s}: {mm[:4]}"


@pytest.mark.parametrize("seed", [1, 2])
def test_mobile_in_lockstep(product, oracle, seed):
    """pendulums on rods and ball joints, a hinged chain, springs sharing a body, a net of rods: 300 steps"""
    g, c = pair(product, oracle, js.mobile, seed=seed)
    assert product.dll.BallboxJointCount(g.ptr) == oracle.dll.BallboxJointCount(c.ptr) > 30
    lockstep(g, c, 300, "mobile")
    sn = g.snapshot()
    assert np.isfinite(sn["s_pos"]).all() and float(np.abs(sn["s_vel"]).max()) > 0.0


def test_chain_and_springs_on_a_heap_in_lockstep(product, oracle):
    """joint rows and contact rows act on the same bodies, iteration by iteration"""
    g, c = pair(product, oracle, js.chain_on_pile, seed=4)
    lockstep(g, c, 260, "heap")
    assert len(g.constraints()) > 50
    assert g.stats().levels > 3


def test_resident_stepping_with_joints(product, oracle):
    g, c = pair(product, oracle, js.chain_on_pile, seed=9)
    g.set_sync_mode(bb.SYNC_RESIDENT)
    g.step_n(DT, 150); g.download()
    for _ in range(150):
        c.step(DT)
    assert not parity.compare_worlds(g, c)
    # a joint added to a resident world takes its anchors from the CURRENT positions (the bodies are downloaded first)
    sn = c.snapshot()
    p = tuple(float(x) for x in sn["s_pos"][0])
    for w in (g, c):
        w.add_distance_joint((js.S, 0), (js.W, 0), p, (p[0], p[1] + 1.0, p[2]))
    g.step_n(DT, 40); g.download()
    for _ in range(40):
        c.step(DT)
    assert not parity.compare_worlds(g, c)


@pytest.mark.parametrize("iterations", [0, 1, 3])
def test_iteration_counts(product, oracle, iterations):
    g, c = pair(product, oracle, js.chain_on_pile, seed=2, n_loose=20)
    for w in (g, c):
        w.lib.dll.SetSolverIterations(w.ptr, iterations)
    lockstep(g, c, 60, f"iterations={iterations}")


def test_warm_starting_and_other_schedules_with_joints(product, oracle):
    """warm-started contacts + joints; a coloured / flow request falls back to the exact schedule when joints exist"""
    g, c = pair(product, oracle, js.chain_on_pile, seed=6, n_loose=30)
    g.set_warm_starting(0.8); c.set_warm_starting(0.8)
    lockstep(g, c, 80, "warm")
    for sched in (bb.SOLVER_COLOURED, bb.SOLVER_EXACT_FLOW):
        product.dll.BallboxSetSolverSchedule(g.ptr, sched)
        lockstep(g, c, 10, f"schedule {sched}")


def test_clear_joints_gives_the_plain_world_back(product, oracle, truth):
    g, c = pair(product, oracle, js.mobile, seed=5)
    lockstep(g, c, 30, "before")
    g.clear_joints(); c.clear_joints()
    assert product.dll.BallboxJointCount(g.ptr) == 0 and product.dll.BallboxSpringCount(g.ptr) == 0
    lockstep(g, c, 60, "after")


def test_batch_of_worlds_with_joints(product, oracle):
    """joints of different worlds of a batch: components never cross worlds; worlds get different joint sets"""
    d = product.dll
    n_worlds = 24
    s, b, k = js.capacity("pile")
    batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, k)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(n_worlds)]
    singles = []
    for i, v in enumerate(views):
        kw = dict(seed=100 + i, n_loose=10 + i, links=2 + i % 5)
        if i % 6 == 5:
            kw["links"] = 1                                                              # a world with one rod and one spring only
        js.chain_on_pile(v, **kw)
        singles.append(js.chain_on_pile(oracle.create_world(s, b, k), **kw))
    assert d.PhysicsStepBatch(batch, DT, 90) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    for w in singles:
        for _ in range(90):
            w.step(DT)
    for i, (v, w) in enumerate(zip(views, singles)):
        mm = parity.compare_worlds(v, w)
        assert not mm, f"world {i}: {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)


def test_many_components_and_deep_joint_levels(product, oracle):
    """600 pendulums (600 components over the sweep's blocks) and one 120-link chain (240 joint levels in one block)"""
    def scene(w):
        w.set_gravity((0.0, -9.81, 0.0))
        for k in range(600):
            x, z = 1.5 * (k % 30), 1.5 * (k // 30)
            i = w.add_sphere((x + 1.0, 6.0, z), 0.2)
            w.add_distance_joint((js.S, i), (js.W, 0), (x + 1.0, 6.0, z), (x, 6.0, z))
        a = w.add_box((-5.0, 8.0, -5.0), (0.2, 0.2, 0.2)); w.lib.dll.SetBoxStatic(w.ptr, a, True)
        prev = a
        for k in range(120):
            b = w.add_box((-5.0 + 0.3 * (k + 1), 8.0, -5.0), (0.12, 0.05, 0.05))
            w.add_hinge_joint((js.B, prev), (js.B, b), (-5.0 + 0.3 * k + 0.15, 8.0, -5.0), (0, 0, 1), 0.05)
            prev = b
        return w
    g, c = scene(product.create_world(700, 200, 20000)), scene(oracle.create_world(700, 200, 20000))
    lockstep(g, c, 40, "many")


def test_contact_graph_deeper_than_the_replay_tables_is_an_error_with_joints(product):
    """a column of 300 touching spheres is a 299-level dependency chain; with a joint in the world the one-iteration
    replays cannot hold it (ballbox_ext.h): the step fails loudly instead of skipping sweeps"""
    def column():
        w = product.create_world(400, 4, 4000)
        w.set_gravity((0.0, -9.81, 0.0))
        f = w.add_box((0.0, -0.5, 0.0), (5.0, 0.5, 5.0)); product.dll.SetBoxStatic(w.ptr, f, True)
        for k in range(300):
            w.add_sphere((0.0, 0.19 + 0.38 * k, 0.0), 0.2)
        return w
    w = column()
    for _ in range(3):
        w.step(DT)                                                    # without joints the deep graph is fine
    assert w.stats().levels > 256
    w = column()
    w.add_distance_joint((js.S, 299), (js.W, 0), (0.0, 0.19 + 0.38 * 299, 0.0), (0.0, 0.19 + 0.38 * 299 + 1.0, 0.0))
    with pytest.raises(bb.BallboxError, match="deeper"):
        for _ in range(3):
            w.step(DT)


def _log(w, log):
    for i in range(0, w.n_spheres, 2):
        w.set_sphere_callback(i, lambda s, t, o, c: log.append(("s", s, t, o, c.contents.bodyA_index, c.contents.bodyB_index, c.contents.penetration)))
    for i in range(0, w.n_boxes, 3):
        w.set_box_callback(i, lambda s, t, o, c: log.append(("b", s, t, o, c.contents.bodyA_index, c.contents.bodyB_index, c.contents.penetration)))


@pytest.mark.parametrize("mode", [bb.SYNC_COMPAT, bb.SYNC_COMPAT_BODIES, bb.SYNC_RESIDENT])
def test_callbacks_with_joints_in_every_sync_mode(product, oracle, mode):
    """the one-iteration-per-launch solve of a jointed world leaves callbacks, events and the constraint export alone"""
    g, c = pair(product, oracle, js.chain_on_pile, seed=12, n_loose=30)
    gl, cl = [], []
    _log(g, gl); _log(c, cl)
    g.set_sync_mode(mode)
    if mode == bb.SYNC_RESIDENT:
        g.step_n(DT, 70); g.download()
    else:
        for _ in range(70):
            g.step(DT)
        if mode == bb.SYNC_COMPAT_BODIES:
            g.download(bb.DL_CONSTRAINTS)
    for _ in range(70):
        c.step(DT)
    assert len(cl) > 100 and gl == cl
    assert not parity.compare_worlds(g, c)


def test_batch_reset_with_joints(product, oracle):
    """BatchSnapshot / BatchResetWorlds on jointed worlds: joints are kept (their anchors are body-frame), the reset world
    equals a fresh one"""
    d = product.dll
    d.BatchSnapshot.argtypes = [bb.C.c_void_p]
    d.BatchResetWorlds.argtypes = [bb.C.c_void_p, bb.C.POINTER(bb.C.c_int), bb.C.c_int]
    n_worlds = 16
    s, b, k = js.capacity("pile")
    batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, k)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(n_worlds)]
    kws = [dict(seed=300 + i, n_loose=12, links=3 + i % 3) for i in range(n_worlds)]
    refs = []
    for v, kw in zip(views, kws):
        js.chain_on_pile(v, **kw)
        refs.append(js.chain_on_pile(oracle.create_world(s, b, k), **kw))
    assert d.BatchUpload(batch) == 0
    assert d.BatchSnapshot(batch) == 0
    assert d.PhysicsStepBatch(batch, DT, 45) == 0, d.BatchGetLastError(batch)
    for r in refs:
        for _ in range(45):
            r.step(DT)
    ids = (bb.C.c_int * 3)(2, 7, 15)
    assert d.BatchResetWorlds(batch, ids, 3) == 0
    for i in (2, 7, 15):
        refs[i] = js.chain_on_pile(oracle.create_world(s, b, k), **kws[i])
    assert d.PhysicsStepBatch(batch, DT, 30) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    for r in refs:
        for _ in range(30):
            r.step(DT)
    for i, (v, r) in enumerate(zip(views, refs)):
        mm = parity.compare_worlds(v, r)
        assert not mm, f"world {i}: {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)
