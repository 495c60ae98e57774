"""Pins the CPU checkers: the C restatement (oracle/bbx_oracle.c) against the golden
vectors generated from the unmodified reference, against the compiled reference
itself where it is available, and the pruned drivers against the full all-pairs
loops.  No GPU."""
import pytest

from ballbox_b200 import binding as bb
from tests import golden_tools, parity


@pytest.mark.parametrize("case", list(golden_tools.CASES))
def test_restatement_reproduces_reference_golden(checkers, case):
    golden_tools.check_case(checkers["oracle"], case)


@pytest.mark.parametrize("case", ["readme", "mixed_300"])
def test_compiled_reference_reproduces_golden(checkers, case):
    if "ref" not in checkers:
        pytest.skip("oracle/_ref not built here")
    golden_tools.check_case(checkers["ref"], case)


def test_readme_golden_trace_of_survey(checkers):
    """The numbers SURVEY.md section 8c quotes for config 1."""
    lib = checkers.get("ref") or checkers["oracle"]
    w = lib.build_scene(bb.SCENE_README, 0, 1)
    hits = []
    w.set_sphere_callback(0, lambda s, t, o, c: hits.append((step, t, o, c.contents.penetration)))
    for step in range(1000):
        if step < 30:
            lib.dll.AddSphereForce(w.ptr, 0, bb.Vec3(2.0, 0, 0))
        w.step(1 / 60)
        if step == 0:
            p = w.sphere_arrays()["pos"][0]
            assert abs(p[0] - 0.000544444483) < 1e-12 and abs(p[1] - 4.9973321) < 1e-6
    assert len(hits) == 933 and hits[0][0] == 67 and abs(hits[0][3] - 0.089353) < 1e-6
    p = w.sphere_arrays()["pos"][0]
    assert abs(p[0] - 0.656878471) < 1e-7 and abs(p[1] - 0.949999928) < 1e-7


@pytest.mark.parametrize("kind,n,steps", [(bb.SCENE_BALLS, 600, 40), (bb.SCENE_MIXED, 500, 40), (bb.SCENE_SDF, 300, 50)])
def test_restatement_equals_reference_every_step(checkers, kind, n, steps):
    if "ref" not in checkers:
        pytest.skip("oracle/_ref not built here")
    a, b = checkers["ref"].build_scene(kind, n, 9), checkers["oracle"].build_scene(kind, n, 9)
    for s in range(steps):
        a.step(1 / 60); b.step(1 / 60)
        mm = parity.compare_worlds(a, b)
        assert not mm, f"step {s}: {mm[:3]}"


@pytest.mark.parametrize("which", ["oracle", "ref"])
@pytest.mark.parametrize("kind,n", [(bb.SCENE_BALLS, 2500), (bb.SCENE_MIXED, 1200), (bb.SCENE_SDF, 800)])
def test_pruned_driver_is_identical_to_all_pairs(checkers, which, kind, n):
    if which not in checkers:
        pytest.skip(which + " not built here")
    lib = checkers[which]
    a, b = lib.build_scene(kind, n, 31), lib.build_scene(kind, n, 31)
    for s in range(25):
        a.step(1 / 60); b.step(1 / 60, pruned=True)
        mm = parity.compare_worlds(a, b)
        assert not mm, f"step {s}: {mm[:3]}"


def test_callback_order_is_the_reference_order(checkers):
    """sphere-sphere, box-box (once per pair), sphere-box, sphere-SDF; box-SDF fires nothing."""
    logs = []
    for lib in checkers.values():
        w = lib.create_world(8, 8, 256)
        w.set_gravity((0, -9.8, 0))
        d = lib.dll
        for p in ((0, 0.4, 0), (0.6, 0.4, 0), (0.3, 0.9, 0.1)):
            d.AddSphere(w.ptr, bb.Vec3(*p), 0.5)
        for p in ((0, 0.3, 0.2), (0.5, 0.35, 0.2), (3, 0.2, 3)):
            d.AddBox(w.ptr, bb.Vec3(*p), bb.Vec3(0.4, 0.4, 0.4), d.QuatIdentity())
        w.use_sdf(bb.SDF_PLANE_Y)
        log = []
        for i in range(3):
            w.set_sphere_callback(i, lambda s, t, o, c, tag="s": log.append((tag, s, t, o, round(c.contents.penetration, 6))))
            w.set_box_callback(i, lambda s, t, o, c, tag="b": log.append((tag, s, t, o, round(c.contents.penetration, 6))))
        w.step(1 / 60)
        assert log, "scene must produce contacts"
        kinds = [(t, o_t) for (t, _, o_t, _, _) in log]
        phases = [0 if k == ("s", 0) else 1 if k == ("b", 1) else 2 if k in (("s", 1), ("b", 0)) else 3 for k in kinds]
        assert phases == sorted(phases)
        logs.append(log)
    assert all(l == logs[0] for l in logs)


def test_warm_starting_definition(checkers):
    """Warm starting is an extension (ballbox_ext.h), defined sequentially in the restatement only.  Factor 0 is the
    reference's behaviour bit for bit; factor 1 with TWO sweeps leaves a sphere pile less interpenetrated than the
    reference's eight cold sweeps, which is what the feature is for."""
    import numpy as np
    o = checkers["oracle"]

    def run(factor, iters):
        w = o.build_scene(bb.SCENE_BALLS, 400, 9)
        w.w.settings.solver_iterations = iters
        if factor is not None:
            w.set_warm_starting(factor)
        pens = []
        for s in range(200):
            w.step(1 / 60, pruned=True)
            if s >= 150:
                k = w.constraints()
                ss = k[(k["a_type"] == 0) & (k["b_type"] == 0)]
                pens.append(float(ss["penetration"].mean()))
        return w, float(np.mean(pens))

    cold8, pen_cold8 = run(None, 8)
    zero8, _ = run(0.0, 8)
    assert not parity.compare_worlds(cold8, zero8)
    _, pen_warm2 = run(1.0, 2)
    assert pen_warm2 < 0.8 * pen_cold8, (pen_warm2, pen_cold8)
