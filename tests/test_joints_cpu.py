"""Joints and springs (extension, include/ballbox_ext.h): the SEQUENTIAL definition in oracle/bbx_oracle.c behaves like
joints, and the product library's host side (no GPU needed) keeps the same lists.  The reference has no joints
(README.md:25 lists them as planned): nothing pins this definition but itself, the GPU path is compared with it bit for
bit in tests/test_gpu_joints.py."""
import numpy as np
import pytest

from ballbox_b200 import binding as bb
from tests import joint_scenes as js

DT = 1.0 / 60.0


@pytest.fixture(scope="module")
def oracle(checkers):
    return checkers["oracle"]


def test_rod_and_ball_joint_hold_their_anchors(oracle):
    w = oracle.create_world(8, 8, 64)
    w.set_gravity((0.0, -9.81, 0.0))
    s = w.add_sphere((1.0, 5.0, 0.0), 0.2)
    w.add_distance_joint((js.S, s), (js.W, 0), (1.0, 5.0, 0.0), (0.0, 5.0, 0.0))
    b = w.add_box((4.0, 5.0, 0.0), (0.2, 0.2, 0.2))
    w.add_ball_joint((js.B, b), (js.W, 0), (3.0, 5.0, 0.0))
    lowest = 9.0
    for _ in range(240):
        w.step(DT)
        sn = w.snapshot()
        assert abs(np.linalg.norm(sn["s_pos"][s] - np.array([0, 5, 0])) - 1.0) < 0.02        # the rod keeps its length
        assert abs(np.linalg.norm(sn["b_pos"][b] - np.array([3, 5, 0])) - 1.0) < 0.02        # the box swings about its anchor
        lowest = min(lowest, sn["s_pos"][s][1])
    assert lowest < 4.1                                                                       # and it did swing


def test_spring_settles_at_its_loaded_length(oracle):
    w = oracle.create_world(4, 4, 16)
    w.set_gravity((0.0, -9.81, 0.0))
    s = w.add_sphere((0.0, 4.0, 0.0), 0.2)
    w.add_spring((js.S, s), (js.W, 0), (0.0, 4.0, 0.0), (0.0, 6.0, 0.0), 1.5, 50.0, 2.0)
    for _ in range(900):
        w.step(DT)
    y = w.snapshot()["s_pos"][s][1]
    assert abs((6.0 - y) - (1.5 + 9.81 / 50.0)) < 0.02                                       # rest + m g / k, m = 1


def test_hinged_chain_hangs_from_its_static_end(oracle):
    w = oracle.create_world(4, 16, 64)
    w.set_gravity((0.0, -9.81, 0.0))
    a = w.add_box((0.0, 5.0, 0.0), (0.2, 0.2, 0.2)); oracle.dll.SetBoxStatic(w.ptr, a, True)
    prev, links = a, []
    for k in range(4):
        b = w.add_box((0.6 * (k + 1), 5.0, 0.0), (0.25, 0.1, 0.1)); links.append(b)
        assert w.add_hinge_joint((js.B, prev), (js.B, b), (0.6 * k + 0.3, 5.0, 0.0), (0, 0, 1), 0.2) == 2 * k
        prev = b
    assert oracle.dll.BallboxJointCount(w.ptr) == 8
    for _ in range(900):
        w.step(DT)
    sn = w.snapshot()
    assert np.allclose(sn["b_pos"][links[-1]], [0.3, 5.0 - 2.1, 0.0], atol=0.08)             # straight down from the first hinge
    assert abs(sn["b_pos"][links[-1]][2]) < 0.02                                             # the hinge axis keeps it in plane


def test_joint_arguments_and_clear(oracle, product):
    for lib in (oracle, product):                          # the product's host side needs no GPU for this
        w = lib.create_world(4, 4, 16)
        s = w.add_sphere((0, 1, 0), 0.2); b = w.add_box((1, 1, 0), (0.2, 0.2, 0.2))
        d = lib.dll
        V = bb.Vec3
        assert d.BallboxAddBallJoint(w.ptr, 0, 7, 1, b, V(0, 1, 0)) == -1                    # no such sphere
        assert d.BallboxAddBallJoint(w.ptr, -1, 0, 1, b, V(0, 1, 0)) == -1                   # the world cannot be end A
        assert d.BallboxAddBallJoint(w.ptr, 3, 0, 1, b, V(0, 1, 0)) == -1                    # no such type
        assert d.BallboxAddSpring(w.ptr, 0, s, 1, 9, V(0, 1, 0), V(1, 1, 0), 1.0, 10.0, 0.0) == -1
        assert d.BallboxAddBallJoint(w.ptr, 0, s, 1, b, V(0.5, 1, 0)) == 0
        assert d.BallboxAddDistanceJoint(w.ptr, 0, s, -1, 0, V(0, 1, 0), V(0, 3, 0)) == 1
        assert d.BallboxAddSpring(w.ptr, 0, s, 1, b, V(0, 1, 0), V(1, 1, 0), 1.0, 10.0, 0.0) == 0
        assert d.BallboxJointCount(w.ptr) == 2 and d.BallboxSpringCount(w.ptr) == 1
        d.BallboxClearJoints(w.ptr)
        assert d.BallboxJointCount(w.ptr) == 0 and d.BallboxSpringCount(w.ptr) == 0
        w.destroy()


def test_world_without_joints_is_untouched_by_the_extension(oracle, checkers):
    """the joint code paths must not change a world that has none: the restatement still equals the compiled reference"""
    ref = checkers.get("ref")
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    a, b = oracle.build_scene(bb.SCENE_MIXED, 120, 5), ref.build_scene(bb.SCENE_MIXED, 120, 5)
    for _ in range(60):
        a.step(DT); b.step(DT)
    from tests import parity
    assert not parity.compare_worlds(a, b)


def test_joint_scenes_are_deterministic_and_lively(oracle):
    runs = []
    for _ in range(2):
        w = js.mobile(oracle.create_world(*js.capacity("mobile")), seed=3)
        for _ in range(120):
            w.step(DT)
        runs.append(w.snapshot())
    for f in ("s_pos", "b_pos", "b_rot", "s_vel"):
        assert np.array_equal(runs[0][f], runs[1][f])
    assert np.isfinite(runs[0]["s_pos"]).all() and np.isfinite(runs[0]["b_pos"]).all()
    w = js.chain_on_pile(oracle.create_world(*js.capacity("pile")), seed=1)
    for _ in range(150):
        w.step(DT)
    sn = w.snapshot()
    assert np.isfinite(sn["b_pos"]).all() and len(w.constraints()) > 50                      # the chain lies on the heap
