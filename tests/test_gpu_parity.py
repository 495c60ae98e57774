"""Parity tests proper: the CUDA path, called through the C-ABI library exactly as a
reference user would (PhysicsStep(world, dt) with host arrays), against the CPU
checker on identical scenes.  Bar (BASELINE.json north_star): bit-exact contact
identities; normals, penetrations and post-solve state within 1e-4 relative over one
step.  With the exact (level-scheduled) solver everything is in fact bit-identical,
so most tests assert equality of the bits; test_tolerance_as_stated spells out the
stated tolerance."""
import numpy as np
import pytest

from ballbox_b200 import binding as bb
from tests import golden_tools, parity

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0
REL_TOL = 1e-4          # north_star: 1e-4 relative over 1 step


def lockstep(product, truth, kind, n, seed, steps, pruned=False, tweak=None, every=1):
    g, c = product.build_scene(kind, n, seed), truth.build_scene(kind, n, seed)
    if tweak:
        tweak(product, g); tweak(truth, c)
    for s in range(steps):
        g.step(DT); c.step(DT, pruned=pruned)
        if s % every == 0 or s == steps - 1:
            mm = parity.compare_worlds(g, c)
            assert not mm, f"scene {kind} n={n} step {s}: {mm[:4]}"
    return g, c


@pytest.mark.parametrize("case", list(golden_tools.CASES))
def test_golden_vectors_on_gpu(product, case):
    """includes config 1 (README quick start): 933 callbacks replayed in reference order."""
    golden_tools.check_case(product, case)


@pytest.mark.parametrize("kind,n,seed,steps", [
    (bb.SCENE_BALLS, 800, 3, 150), (bb.SCENE_MIXED, 600, 4, 200), (bb.SCENE_SDF, 500, 5, 200),
    (bb.SCENE_RLWORLD, 0, 6, 200), (bb.SCENE_MIXED, 64, 7, 300)])
def test_bit_exact_every_step(product, truth, kind, n, seed, steps):
    lockstep(product, truth, kind, n, seed, steps)


def test_tolerance_as_stated(product, truth):
    g, c = product.build_scene(bb.SCENE_MIXED, 2000, 12), truth.build_scene(bb.SCENE_MIXED, 2000, 12)
    for _ in range(30):
        g.step(DT); c.step(DT, pruned=True)
    kg, kc = g.constraints(), c.constraints()
    assert len(kg) == len(kc) > 500
    for f in ("a_type", "a_index", "b_type", "b_index"):            # contact-pair identities: exact
        assert np.array_equal(kg[f], kc[f])
    assert parity.max_rel_err(kg["normal"], kc["normal"]) <= REL_TOL
    assert parity.max_rel_err(kg["penetration"], kc["penetration"]) <= REL_TOL
    sg, sc = g.snapshot(), c.snapshot()
    for f in ("s_pos", "s_vel", "b_pos", "b_vel", "b_rot", "b_angvel"):
        assert parity.max_rel_err(sg[f], sc[f]) <= REL_TOL, f


def test_drift_over_1000_steps_is_zero_with_exact_schedule(product, truth):
    g, c = product.build_scene(bb.SCENE_MIXED, 250, 21), truth.build_scene(bb.SCENE_MIXED, 250, 21)
    g.set_sync_mode(bb.SYNC_RESIDENT)
    g.step_n(DT, 1000); g.download()
    for _ in range(1000):
        c.step(DT)
    assert not parity.compare_worlds(g, c)


@pytest.mark.parametrize("kind,n,steps", [(bb.SCENE_BALLS, 100_000, 3), (bb.SCENE_MIXED, 200_000, 3), (bb.SCENE_SDF, 100_000, 3)])
def test_large_scenes_against_pruned_reference(product, truth, kind, n, steps):
    """BASELINE configs 2-4 at (or near) full size; the checker runs behind its CPU grid."""
    lockstep(product, truth, kind, n, 20261019, steps, pruned=True)


def test_resident_equals_compat_and_is_deterministic(product):
    a, b, c = (product.build_scene(bb.SCENE_MIXED, 3000, 8) for _ in range(3))
    b.set_sync_mode(bb.SYNC_RESIDENT); c.set_sync_mode(bb.SYNC_RESIDENT)
    for _ in range(40):
        a.step(DT)
    b.step_n(DT, 40); b.download()
    for _ in range(4):
        c.step_n(DT, 10)
    c.download()
    assert not parity.compare_worlds(a, b)
    assert not parity.compare_worlds(a, c)
    assert a.state_hash() == b.state_hash() == c.state_hash()


def test_batch_equals_independent_worlds(product, truth):
    d = product.dll
    n_worlds = 12
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, c)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(n_worlds)]
    singles = []
    for i, v in enumerate(views):
        d.BbxSceneBuild(v.ptr, bb.SCENE_RLWORLD, 0, 900 + i)
        singles.append(truth.build_scene(bb.SCENE_RLWORLD, 0, 900 + i))
    assert d.PhysicsStepBatch(batch, DT, 60) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    for w in singles:
        for _ in range(60):
            w.step(DT)
    for i, (v, w) in enumerate(zip(views, singles)):
        mm = parity.compare_worlds(v, w)
        assert not mm, f"world {i}: {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)


# ---------------------------------------------------------------------------- edge cases
def test_empty_world_and_bad_dt(product):
    w = product.create_world(4, 4, 16)
    w.step(DT)
    w.step(0.0); w.step(-1.0); w.step(1e-7)            # ignored (ballbox.h:1791)
    assert len(w.constraints()) == 0
    w.add_sphere((0, 0, 0), 1.0)
    w.set_gravity((0, -9.8, 0))
    w.step(0.0)
    assert w.sphere_arrays()["pos"][0][1] == 0.0


def test_static_only_world_emits_static_contacts(product, truth):
    def build(lib):
        w = lib.create_world(4, 4, 64)
        for p in ((0, 0, 0), (1.5, 0.2, 0), (0.5, 1.2, 0.3)):
            i = w.add_box(p, (1, 1, 1)); lib.dll.SetBoxStatic(w.ptr, i, True)
        i = w.add_sphere((0.2, 0.1, 0.9), 0.7); lib.dll.SetSphereStatic(w.ptr, i, True)
        return w
    g, c = build(product), build(truth)
    for _ in range(3):
        g.step(DT); c.step(DT)
        assert not parity.compare_worlds(g, c)
    assert len(g.constraints()) > 0


def _tweaks():
    def zero_iter(lib, w): w.w.settings.solver_iterations = 0
    def one_iter(lib, w): w.w.settings.solver_iterations = 1
    def no_friction(lib, w): w.w.settings.friction = 0.0
    def defaults25(lib, w): lib.dll.ResetPhysicsSettingsToDefaults(w.ptr)
    def manifold4(lib, w): lib.dll.SetManifoldSettings(w.ptr, 0.02, -0.01, 4)
    def manifold1(lib, w): lib.dll.SetManifoldSettings(w.ptr, 0.01, -0.5, 1)
    def bouncy(lib, w): lib.dll.SetRestitution(w.ptr, 0.9); w.w.settings.velocity_threshold = 1.0
    def no_sleep(lib, w): w.w.settings.sleep_time_required = 1e9
    def heavy(lib, w):
        for i in range(0, w.n_spheres, 3): lib.dll.SetSphereMass(w.ptr, i, 5.0)
        for i in range(6, w.n_boxes, 4): lib.dll.SetBoxMass(w.ptr, i, 0.25)
    return [zero_iter, one_iter, no_friction, defaults25, manifold4, manifold1, bouncy, no_sleep, heavy]


@pytest.mark.parametrize("tweak", _tweaks(), ids=lambda f: f.__name__)
def test_settings_variants(product, truth, tweak):
    lockstep(product, truth, bb.SCENE_MIXED, 400, 15, 60, tweak=tweak)


def test_user_forces_torques_and_direct_velocity_writes(product, truth):
    g, c = product.build_scene(bb.SCENE_MIXED, 300, 5), truth.build_scene(bb.SCENE_MIXED, 300, 5)
    for s in range(50):
        for lib, w in ((product, g), (truth, c)):
            d = lib.dll
            if s % 3 == 0:
                d.AddSphereForce(w.ptr, 3, bb.Vec3(5, 1, -2)); d.AddSphereTorque(w.ptr, 4, bb.Vec3(0.1, 0.2, 0.3))
                d.AddBoxForce(w.ptr, 9, bb.Vec3(-3, 8, 0.5)); d.AddBoxTorque(w.ptr, 10, bb.Vec3(1, -1, 0.5))
            if s == 20:
                w.sphere_arrays()["vel"][7] = (1.0, 2.0, 3.0)          # "direct velocity control" (README.md:154)
                w.box_arrays()["pos"][8] += np.float32(0.25)
                w.w.gravity.y = -3.0
        g.step(DT); c.step(DT)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"step {s}: {mm[:3]}"


def test_odd_geometry(product, truth):
    """non-unit static rotation, sphere centre inside a box, a large dynamic box, a far-away body."""
    def build(lib):
        w = lib.create_world(64, 64, 4096)
        w.set_gravity((0, -9.8, 0))
        d = lib.dll
        g = w.add_box((0, -1, 0), (8, 1, 8)); d.SetBoxStatic(w.ptr, g, True)
        t = w.add_box((3, 0.4, 3), (1, 0.5, 1), (0.1, 0.3, 0.0, 1.2)); d.SetBoxStatic(w.ptr, t, True)     # |q| != 1
        w.add_box((0, 1.5, 0), (2.5, 0.5, 2.5))                                                          # big dynamic
        w.add_sphere((0.1, 1.4, 0.1), 0.3)                                                               # inside the big box
        w.add_sphere((500.0, 3.0, -700.0), 0.5)                                                          # far away
        for i in range(20):
            w.add_sphere((-3 + 0.35 * i, 2.5 + 0.1 * (i % 3), -2 + 0.2 * i), 0.3 + 0.02 * (i % 5))
            w.add_box((-3 + 0.3 * i, 4.0, 2 - 0.2 * i), (0.2 + 0.03 * (i % 4), 0.3, 0.25), w.quat_axis_angle((1, i + 1, 2), 0.3 * i))
        return w
    g, c = build(product), build(truth)
    for s in range(120):
        g.step(DT); c.step(DT)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"step {s}: {mm[:3]}"


@pytest.mark.parametrize("extra", ["far_1e6", "far_1e30", "nan_sphere", "inf_sphere", "huge_radius"])
def test_extreme_inputs_match_the_reference(product, truth, extra):
    """Bodies far outside the pile (the grid must not blow up), non-finite sphere positions (every comparison with a
    NaN is false: the reference never reports a contact for them) and a sphere that contains the whole scene."""
    add = {"far_1e6": lambda w: w.add_sphere((1e6, 3.0, -2e6), 0.5), "far_1e30": lambda w: w.add_sphere((1e30, 3.0, 0.0), 0.5),
           "nan_sphere": lambda w: w.add_sphere((float("nan"), 3.0, 0.0), 0.5), "inf_sphere": lambda w: w.add_sphere((float("inf"), 3.0, 0.0), 0.5),
           "huge_radius": lambda w: w.add_sphere((0.0, 50.0, 0.0), 1e6)}[extra]

    def build(lib):
        w = lib.create_world(64, 64, 4096)
        w.set_gravity((0, -9.8, 0))
        g = w.add_box((0, -1, 0), (8, 1, 8)); lib.dll.SetBoxStatic(w.ptr, g, True)
        for i in range(20):
            w.add_sphere((-3 + 0.35 * i, 1.5 + 0.1 * (i % 3), -2 + 0.2 * i), 0.3)
            w.add_box((-3 + 0.3 * i, 3.0, 2 - 0.2 * i), (0.2, 0.3, 0.25), w.quat_axis_angle((1, i + 1, 2), 0.3 * i))
        add(w)
        return w
    g, c = build(product), build(truth)
    for s in range(40):
        g.step(DT); c.step(DT)
        sa, sb = g.snapshot(), c.snapshot()
        for f in ("s_pos", "s_vel", "s_angvel", "b_pos", "b_vel", "b_rot", "b_angvel"):
            assert np.array_equal(sa[f], sb[f], equal_nan=True), (s, f)
        assert len(g.constraints()) == len(c.constraints()), s


def _sdf_scene(lib, fn_ptr):
    """Spheres and rotated boxes dropped onto a host-function SDF (no UseSDFDevice)."""
    w = lib.create_world(64, 32, 4096)
    w.set_gravity((0.0, -9.8, 0.0))
    for i in range(40):
        w.add_sphere((-4.0 + 0.9 * (i % 10), 1.5 + 0.8 * (i // 10), -1.5 + 1.1 * (i % 4)), 0.3 + 0.02 * (i % 5))
    for i in range(20):
        w.add_box((-4.2 + 0.85 * (i % 10), 5.0 + 0.9 * (i // 10), 0.4 * (i % 3) - 0.4), (0.25, 0.2 + 0.02 * (i % 4), 0.3),
                  w.quat_axis_angle((1, 2 + i, 3), 0.37 * i))
    lib.dll.UseSDF(w.ptr, fn_ptr)
    return w


def test_host_sdf_is_baked_plane(product, truth):
    """UseSDF(world, host_fn) without a device SDF: the function is sampled once on a grid and interpolated
    on the GPU (ballbox_ext.h).  f(p) = p.y is linear, trilinear interpolation reproduces it up to rounding,
    so the run stays within the north star's 1e-4 of the reference that calls the function itself."""
    g = _sdf_scene(product, bb.C.cast(product.dll.BbxHostSdfPlane0, bb.C.c_void_p))
    c = _sdf_scene(truth, bb.C.cast(truth.dll.BbxHostSdfPlane0, bb.C.c_void_p))
    assert product.dll.BallboxSdfInUse(g.ptr) == 5            # BBX_SDF_BAKED
    for s in range(45):
        g.step(DT); c.step(DT)
    assert len(g.constraints()) == len(c.constraints()) > 20
    a, b = g.snapshot(), c.snapshot()
    for f in ("s_pos", "b_pos", "s_vel", "b_vel"):
        err = np.abs(a[f].astype(np.float64) - b[f]).max()
        assert err < 1e-4 * max(1.0, np.abs(b[f]).max()), (f, err)


def test_host_sdf_python_callback_is_baked_wavy(product, truth):
    """An arbitrary user function (here a Python callback) on the README terrain: baked over an explicit region
    at 0.1 spacing.  Accuracy is the interpolation error (h^2/8 |f''| ~ 1.3e-3 here), so the comparison with
    the reference is a drift bound, not bit parity; nothing tunnels through the terrain."""
    import math
    f = lambda p: p.y - math.sin(p.x) * math.sin(p.z)
    fn_g, fn_c = bb.SDFFunction(f), bb.SDFFunction(f)
    g = _sdf_scene(product, bb.C.cast(fn_g, bb.C.c_void_p))
    c = _sdf_scene(truth, bb.C.cast(fn_c, bb.C.c_void_p))
    product.dll.BallboxSetSDFBakeRegion(g.ptr, bb.Vec3(-8, -2, -6), bb.Vec3(8, 10, 6), 161)
    assert product.dll.BallboxBakeSDF(g.ptr) == 0, product.dll.BallboxGetLastError(g.ptr)
    for s in range(40):
        g.step(DT); c.step(DT)
    a, b = g.snapshot(), c.snapshot()
    assert np.abs(a["s_pos"].astype(np.float64) - b["s_pos"]).max() < 0.05
    assert np.abs(a["b_pos"].astype(np.float64) - b["b_pos"]).max() < 0.08
    n_s = g.n_spheres
    sp = g.sphere_arrays()
    for i in range(n_s):
        x, y, z = (float(v) for v in sp["pos"][i])
        assert y - math.sin(x) * math.sin(z) > -0.15, f"sphere {i} sank into the terrain"
    assert len(g.constraints()) > 5 and abs(len(g.constraints()) - len(c.constraints())) <= 4


def test_baked_sdf_gradient_channels_sharpen_normals(product):
    """Three worlds reach the same state with the exact device SDF; two of them then switch to a bake of the same
    function (host pointer only) and take one more step.  Penetrations agree with the exact ones within the
    interpolation bound; with gradient channels (default) the contact normals are an order of magnitude closer than
    with differences of the interpolated distance (tools/sdf_bake_report.py, profiles/r01_sdf_bake.md)."""
    worlds = [product.build_scene(bb.SCENE_SDF, 600, 77) for _ in range(3)]
    for w in worlds:
        for _ in range(90):
            w.step(DT)
    sa, ba = worlds[0].sphere_arrays(), worlds[0].box_arrays()
    ext = np.concatenate([sa["radius"], np.linalg.norm(ba["size"], axis=1)])[:, None]
    pos = np.concatenate([sa["pos"], ba["pos"]])
    lo, hi = (pos - ext).min(0) - 1.0, (pos + ext).max(0) + 1.0
    h = float((hi - lo).max()) / 192
    for w, grad in ((worlds[1], 0), (worlds[2], 1)):
        product.dll.UseSDFDevice(w.ptr, 0, None)
        product.dll.BallboxSetSDFBakeGradient(w.ptr, grad)
        product.dll.BallboxSetSDFBakeRegion(w.ptr, bb.Vec3(*map(float, lo)), bb.Vec3(*map(float, hi)), 193)
    lists = []
    for w in worlds:
        w.step(DT)
        k = w.constraints().copy()
        lists.append(k[k["b_type"] == 2])
    assert product.dll.BallboxSdfInUse(worlds[0].ptr) == bb.SDF_WAVY_DET and product.dll.BallboxSdfInUse(worlds[2].ptr) == 5
    exact = {}
    for r in lists[0]:
        exact.setdefault((int(r["a_type"]), int(r["a_index"])), []).append(r)
    worst = []
    for baked in lists[1:]:
        ang, dpen = [], []
        for r in baked:
            cands = exact.get((int(r["a_type"]), int(r["a_index"])), [])
            d = [float(np.abs(c["point"].astype(np.float64) - r["point"]).max()) for c in cands]
            if not d or min(d) > 0.05:
                continue
            c = cands[int(np.argmin(d))]
            dpen.append(abs(float(c["penetration"]) - float(r["penetration"])))
            ang.append(np.degrees(np.arccos(np.clip(np.dot(c["normal"].astype(np.float64), r["normal"]), -1, 1))))
        assert len(ang) > 100 and max(dpen) <= 2.5 * h * h / 8, (len(ang), max(dpen), h)
        worst.append(max(ang))
    assert worst[1] < 0.2 * worst[0] and worst[1] < 1.0, worst


def test_contact_overflow_is_reported(product):
    w = product.build_scene(bb.SCENE_BALLS, 3000, 2, max_collisions=64)
    with pytest.raises(bb.BallboxError, match="overflow|max_collisions"):
        for _ in range(5):
            w.step(DT)


def test_callbacks_all_phases_match_reference(product, truth):
    logs = []
    for lib in (product, truth):
        w = lib.build_scene(bb.SCENE_SDF, 120, 33)
        log = []
        for i in range(0, w.n_spheres, 2):
            w.set_sphere_callback(i, lambda s, t, o, c: log.append(("s", s, t, o, c.contents.penetration, c.contents.contact_point.y)))
        for i in range(0, w.n_boxes, 2):
            w.set_box_callback(i, lambda s, t, o, c: log.append(("b", s, t, o, c.contents.penetration, c.contents.contact_point.y)))
        for _ in range(60):
            w.step(DT)
        logs.append(log)
    assert len(logs[0]) > 50 and logs[0] == logs[1]


def _log_callbacks(lib, w, log, every=2):
    for i in range(0, w.n_spheres, every):
        w.set_sphere_callback(i, lambda s, t, o, c: log.append(("s", s, t, o, c.contents.bodyA_index, c.contents.bodyB_index,
                                                                c.contents.penetration, c.contents.contact_point.y)))
    for i in range(0, w.n_boxes, every):
        w.set_box_callback(i, lambda s, t, o, c: log.append(("b", s, t, o, c.contents.bodyA_index, c.contents.bodyB_index,
                                                             c.contents.penetration, c.contents.contact_point.y)))


def test_resident_callbacks_use_device_event_buffer(product, truth):
    """PhysicsStepN on a resident world: events are collected on the device and replayed
    afterwards in (step, emission) order -- the concatenated log equals the reference's."""
    ref_log, gpu_log = [], []
    c = truth.build_scene(bb.SCENE_SDF, 150, 44)
    _log_callbacks(truth, c, ref_log)
    for _ in range(45):
        c.step(DT)
    g = product.build_scene(bb.SCENE_SDF, 150, 44)
    g.set_sync_mode(bb.SYNC_RESIDENT)
    g.upload()
    _log_callbacks(product, g, gpu_log)            # registered AFTER the upload: flags are refreshed lazily
    g.step_n(DT, 20); g.step_n(DT, 25)             # chunks of 8 inside
    assert len(ref_log) > 100
    assert gpu_log == ref_log
    g.download()
    assert not parity.compare_worlds(g, c, constraints=False)


def test_batch_callbacks(product, truth):
    d = product.dll
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(3, s, b, c)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(3)]
    logs = [[] for _ in range(3)]
    ref_logs = [[] for _ in range(3)]
    singles = []
    for i, v in enumerate(views):
        d.BbxSceneBuild(v.ptr, bb.SCENE_RLWORLD, 0, 300 + i)
        _log_callbacks(product, v, logs[i], every=5)
        w = truth.build_scene(bb.SCENE_RLWORLD, 0, 300 + i)
        _log_callbacks(truth, w, ref_logs[i], every=5)
        singles.append(w)
    assert d.PhysicsStepBatch(batch, DT, 30) == 0, d.BatchGetLastError(batch)
    for w in singles:
        for _ in range(30):
            w.step(DT)
    for i in range(3):
        assert len(ref_logs[i]) > 10 and logs[i] == ref_logs[i], f"world {i}"
    d.DestroyPhysicsWorldBatch(batch)


def test_zero_copy_batch_interface(product, truth):
    """Device views alias the library's state; forces written on the device act exactly like
    AddSphereForce/AddBoxTorque on the host; per-world resets restore the uploaded state."""
    import torch
    from ballbox_b200.rl import BatchEnv
    W = 6
    env = BatchEnv(product, W, seed_base=500)
    refs = [truth.build_scene(bb.SCENE_RLWORLD, 0, 500 + i) for i in range(W)]
    sf = torch.zeros(W, env.cap_spheres, 3, device="cuda"); bt = torch.zeros(W, env.cap_boxes, 3, device="cuda")
    for s in range(30):
        sf.zero_(); bt.zero_()
        sf[:, 5, 0] = 3.0 + s * 0.1; sf[:, 7, 1] = 12.0
        bt[:, 9, 2] = 0.5
        env.add_wrench(sphere_force=sf, box_torque=bt)
        env.step(DT, 1)
        for w in refs:
            truth.dll.AddSphereForce(w.ptr, 5, bb.Vec3(3.0 + s * 0.1, 0, 0)); truth.dll.AddSphereForce(w.ptr, 7, bb.Vec3(0, 12.0, 0))
            truth.dll.AddBoxTorque(w.ptr, 9, bb.Vec3(0, 0, 0.5))
            w.step(DT)
    # observations straight from HBM
    obs = env.sphere_pos[:, :, :3].cpu().numpy()
    for i, w in enumerate(refs):
        n = w.n_spheres
        assert np.array_equal(obs[i, :n].view(np.uint32), w.sphere_arrays()["pos"].view(np.uint32)), f"world {i}"
    env.download()
    for i, w in enumerate(refs):
        assert not parity.compare_worlds(env.worlds[i], w, constraints=False), f"world {i}"
    # episode reset of worlds 1 and 4 on the device
    env.reset([1, 4])
    env.step(DT, 10)
    env.download()
    fresh = {i: truth.build_scene(bb.SCENE_RLWORLD, 0, 500 + i) for i in (1, 4)}
    for i in range(W):
        w = fresh.get(i, refs[i])
        for _ in range(10):
            w.step(DT)
        assert not parity.compare_worlds(env.worlds[i], w, constraints=False), f"after reset, world {i}"
    env.close()


def test_coloured_schedule_is_deterministic_and_only_changes_the_solve(product, truth):
    """BBX_SOLVER_COLOURED reorders the Gauss-Seidel sweep: detection (contact identities, normals,
    penetrations) is unaffected on the first step, post-solve state differs from the reference
    (profiles/r01_drift.md quantifies it), and two runs give identical bits."""
    def run():
        w = product.build_scene(bb.SCENE_MIXED, 1500, 77)
        product.dll.BallboxSetSolverSchedule(w.ptr, bb.SOLVER_COLOURED)
        w.step(DT)
        first = w.constraints().copy()
        for _ in range(19):
            w.step(DT)
        return w, first
    a, ka = run()
    b, kb = run()
    assert a.state_hash() == b.state_hash()
    c = truth.build_scene(bb.SCENE_MIXED, 1500, 77)
    c.step(DT, pruned=True)
    kc = c.constraints()
    for f in ("a_type", "a_index", "b_type", "b_index", "point", "normal", "penetration", "normal_mass", "position_bias"):
        assert parity.first_mismatch(ka[f], kc[f]) is None, f
    assert parity.first_mismatch(ka["normal_impulse"], kc["normal_impulse"]) is not None   # the sweep order differs
    sa = a.snapshot()
    for _ in range(19):
        c.step(DT, pruned=True)
    sc = c.snapshot()
    err = np.abs(sa["s_pos"].astype(np.float64) - sc["s_pos"]).max()
    assert 0 < err < 1.0


def test_dense_cluster_exercises_overflow_paths(product, truth):
    """120 spheres and 40 boxes dropped into one small region: > 40 neighbours per body (pair-buffer
    overflow path), > 32 solver nodes per body (long per-body lists)."""
    def build(lib):
        w = lib.create_world(160, 64, 60000)
        w.set_gravity((0, -9.8, 0))
        w.w.settings.solver_iterations = 6
        g = w.add_box((0, -1, 0), (6, 1, 6)); lib.dll.SetBoxStatic(w.ptr, g, True)
        s = 12345
        def u():
            nonlocal s
            s = (s * 1664525 + 1013904223) & 0xffffffff
            return (s >> 8) / 16777216.0
        for i in range(120):
            w.add_sphere((u() * 1.2 - 0.6, 0.6 + u() * 1.2, u() * 1.2 - 0.6), 0.45 + 0.1 * u())
        for i in range(40):
            w.add_box((u() * 1.2 - 0.6, 0.6 + u() * 1.2, u() * 1.2 - 0.6), (0.3 + 0.2 * u(), 0.3, 0.35),
                      w.quat_axis_angle((u() + 0.1, u(), u()), 3.0 * u()))
        return w
    g, c = build(product), build(truth)
    for s in range(25):
        g.step(DT); c.step(DT)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"step {s}: {mm[:3]}"
        if s == 0:
            k = g.constraints()
            per_body = np.bincount(k["a_index"][k["a_type"] == 0], minlength=120)
            assert per_body.max() > 40 and len(k) > 5000


def test_boxes_sunk_into_sdf_produce_long_runs(product, truth):
    """boxes entirely below the SDF surface: all 20 probes fire (the duplicated face-centre probes of
    ballbox.h:2454-2458 included), one solver node of 20 constraints per box."""
    def build(lib):
        w = lib.create_world(8, 16, 2048)
        w.set_gravity((0, -9.8, 0))
        for i in range(10):
            w.add_box((i * 1.7 - 8, -1.5 - 0.1 * i, 0.3 * i), (0.4, 0.3 + 0.02 * i, 0.5), w.quat_axis_angle((1, 0.3 * i, 0.2), 0.4 * i))
        w.add_sphere((0, -3, 0), 0.5); w.add_sphere((2, 0.2, 1), 0.5)
        w.use_sdf(bb.SDF_PLANE_Y)
        return w
    g, c = build(product), build(truth)
    for s in range(20):
        g.step(DT); c.step(DT)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"step {s}: {mm[:3]}"
        if s == 0:
            k = g.constraints()
            assert ((k["a_type"] == 1) & (k["b_type"] == 2)).sum() == 200      # 10 boxes x 20 probes


def test_ragged_batch(product, truth):
    """worlds of a batch with different (including zero) body counts"""
    d = product.dll
    batch = d.CreatePhysicsWorldBatch(5, 40, 24, 2048)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(5)]
    singles = [truth.create_world(40, 24, 2048) for _ in range(5)]
    sizes = [0, 7, 40, 1, 23]
    for wi, n in enumerate(sizes):
        for lib, w in ((product, views[wi]), (truth, singles[wi])):
            w.set_gravity((0, -9.8, 0))
            w.w.settings.solver_iterations = 8
            if n:
                g = w.add_box((0, -1, 0), (5, 1, 5)); lib.dll.SetBoxStatic(w.ptr, g, True)
            for i in range(n):
                w.add_sphere(((i % 5) * 0.8 - 1.6, 0.4 + (i // 5) * 0.9, (i % 3) * 0.7), 0.45)
                if i < 23:
                    w.add_box(((i % 4) * 0.9 - 1.5, 5.0 + i * 0.3, (i % 2) * 0.8), (0.3, 0.25, 0.35), w.quat_axis_angle((1, i, 2), 0.2 * i))
    # gravity/settings come from world 0 of the batch
    assert d.PhysicsStepBatch(batch, DT, 50) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    for wi in range(5):
        for _ in range(50):
            singles[wi].step(DT)
        mm = parity.compare_worlds(views[wi], singles[wi])
        assert not mm, f"world {wi} ({sizes[wi]} bodies): {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)


def test_world_block_solver_path(product, truth):
    """>= 16 small worlds take the block-per-world solve kernel (body state in shared memory,
    __syncthreads between levels); every world must still equal the same world stepped alone."""
    d = product.dll
    n_worlds = 72
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, c)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(n_worlds)]
    for i, v in enumerate(views):
        d.BbxSceneBuild(v.ptr, bb.SCENE_RLWORLD, 0, 4000 + i)
    log = []
    views[5].set_sphere_callback(3, lambda s_, t, o, k: log.append((s_, t, o)))
    assert d.PhysicsStepBatch(batch, DT, 40) == 0, d.BatchGetLastError(batch)
    assert d.PhysicsStepBatch(batch, DT, 35) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    st = bb.BallboxStats(); d.BatchGetStats(batch, bb.C.byref(st))
    assert st.overflow == 0 and st.levels > 3
    for i in (0, 5, 17, 40, 71):
        w = truth.build_scene(bb.SCENE_RLWORLD, 0, 4000 + i)
        for _ in range(75):
            w.step(DT)
        mm = parity.compare_worlds(views[i], w)
        assert not mm, f"world {i}: {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)


@pytest.mark.parametrize("env", [{"BBX_GRP_SMEM": "0"}, {"BBX_GRP_SMEM": "1"}, {"BBX_GRP_SMEM": "1", "BBX_GRP_PAIR_CAP": "256"},
                                 {"BBX_GRP_SMEM": "2", "BBX_GRP_WPB": "2"}, {"BBX_GRP_SMEM": "0", "BBX_GRP_WPB": "3"}])
def test_group_solver_forms_agree_with_the_reference(product, truth, env, monkeypatch):
    """k_solve_groups has two forms: body state in the packed global array (any group), and -- for small groups -- body
    state, visiting order and body slots in shared memory (BBX_GRP_SMEM = largest group in worlds, read when the batch's
    device state is created).  BBX_GRP_PAIR_CAP forces the hand-over from the second to the first after iteration 0 (a group
    with more constraints than its shared-memory tables hold), BBX_GRP_WPB groups of several worlds on a batch that would
    not need them.  Ragged batch of 37 worlds (the last group of the 2- and 3-worlds-per-group runs is short), 70 steps, sampled worlds compared bit for bit with the same world stepped alone by
    the reference, all worlds by hash across the forms."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    d = product.dll
    n_worlds = 37
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, c)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(n_worlds)]
    for i, v in enumerate(views):
        d.BbxSceneBuild(v.ptr, bb.SCENE_RLWORLD, 0, 5100 + i)
    assert d.PhysicsStepBatch(batch, DT, 45) == 0, d.BatchGetLastError(batch)
    assert d.PhysicsStepBatch(batch, DT, 25) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    st = bb.BallboxStats(); d.BatchGetStats(batch, bb.C.byref(st))
    assert st.overflow == 0 and st.levels > 3
    for i in (0, 18, 36):
        w = truth.build_scene(bb.SCENE_RLWORLD, 0, 5100 + i)
        for _ in range(70):
            w.step(DT)
        mm = parity.compare_worlds(views[i], w)
        assert not mm, f"{env}: world {i}: {mm[:3]}"
    hashes = tuple(v.state_hash() for v in views)
    seen = _GROUP_FORM_HASHES.setdefault("h", hashes)
    assert hashes == seen, f"{env}: world hashes differ from the first form run"
    d.DestroyPhysicsWorldBatch(batch)


_GROUP_FORM_HASHES = {}


@pytest.mark.parametrize("kind,n,steps", [(bb.SCENE_MIXED, 3000, 40), (bb.SCENE_SDF, 1500, 40), (bb.SCENE_BALLS, 5000, 30)])
def test_flow_schedule_is_bit_exact(product, truth, kind, n, steps):
    """BBX_SOLVER_EXACT_FLOW: the barrier-free solver (every constraint waits for its own two predecessors
    through versioned body records, iterations overlap) keeps every body's constraints in the reference's
    order, so state AND the full constraint arrays equal the reference's bit for bit, every step."""
    g = product.build_scene(kind, n, 909)
    product.dll.BallboxSetSolverSchedule(g.ptr, bb.SOLVER_EXACT_FLOW)
    c = truth.build_scene(kind, n, 909)
    for s in range(steps):
        g.step(DT); c.step(DT, pruned=True)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"step {s}: {mm[:3]}"
    st = g.stats()
    assert st.overflow == 0 and st.levels >= 3          # "levels" = dependency depth in constraints for this schedule


def test_flow_schedule_batch_and_single_iteration(product, truth):
    """The flow solver on a batch of worlds (chains never cross worlds) and with solver_iterations = 1
    (the final visit of a body happens in iteration 0)."""
    d = product.dll
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(20, s, b, c)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(20)]
    for i, v in enumerate(views):
        d.BbxSceneBuild(v.ptr, bb.SCENE_RLWORLD, 0, 6000 + i)
        d.BallboxSetSolverSchedule(v.ptr, bb.SOLVER_EXACT_FLOW)
    assert d.PhysicsStepBatch(batch, DT, 50) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    for i in (0, 7, 19):
        w = truth.build_scene(bb.SCENE_RLWORLD, 0, 6000 + i)
        for _ in range(50):
            w.step(DT)
        mm = parity.compare_worlds(views[i], w)
        assert not mm, f"world {i}: {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)
    g = product.build_scene(bb.SCENE_MIXED, 800, 31); t = truth.build_scene(bb.SCENE_MIXED, 800, 31)
    d.BallboxSetSolverSchedule(g.ptr, bb.SOLVER_EXACT_FLOW)
    g.w.settings.solver_iterations = 1; t.w.settings.solver_iterations = 1
    for s_ in range(25):
        g.step(DT); t.step(DT)
        mm = parity.compare_worlds(g, t)
        assert not mm, f"step {s_}: {mm[:3]}"


def test_compat_bodies_mode_matches_full_compat(product, truth):
    """BBX_SYNC_COMPAT_BODIES: the drop-in call without the per-step download of the constraint list.  Bodies,
    callbacks and (on request) the constraint list are the ones of the full mode, i.e. of the reference."""
    logs = []
    worlds = []
    for lib, mode in ((product, bb.SYNC_COMPAT_BODIES), (truth, None)):
        w = lib.build_scene(bb.SCENE_MIXED, 500, 71)
        if mode is not None:
            w.set_sync_mode(mode)
        log = []
        for i in range(0, 40, 3):
            w.set_sphere_callback(i, lambda s, t, o, c, log=log: log.append(("s", s, t, o, round(c.contents.penetration, 7))))
            w.set_box_callback(i, lambda s, t, o, c, log=log: log.append(("b", s, t, o, round(c.contents.penetration, 7))))
        logs.append(log); worlds.append(w)
    g, c = worlds
    for s in range(30):
        g.step(DT); c.step(DT)
        assert not parity.compare_worlds(g, c, constraints=False), f"step {s}"
        assert g.w.constraints.contents.count == 0
    assert logs[0] == logs[1] and len(logs[0]) > 10
    g.download(bb.DL_CONSTRAINTS)
    assert not parity.compare_worlds(g, c)


def test_full_size_two_exact_schedules_agree(product):
    """BASELINE config 3 at full size (500k spheres + 500k boxes): the reference cannot be run at this size (~9 h
    per step), so parity is checked through a size-independent property: the level-scheduled solver and the
    barrier-free flow solver are two independent implementations of the reference's per-body constraint order
    (different graph construction, different node granularity, different synchronisation), and they must agree
    on every bit of the state and on the contact count after every block of steps."""
    a = product.build_scene(bb.SCENE_MIXED, 1_000_000, 20261019)
    b = product.build_scene(bb.SCENE_MIXED, 1_000_000, 20261019)
    product.dll.BallboxSetSolverSchedule(b.ptr, bb.SOLVER_EXACT_FLOW)
    a.set_sync_mode(bb.SYNC_RESIDENT); b.set_sync_mode(bb.SYNC_RESIDENT)
    for block in range(4):
        a.step_n(DT, 60); b.step_n(DT, 60)
        a.download(bb.DL_BODIES); b.download(bb.DL_BODIES)
        sa, sb = a.stats(), b.stats()
        assert sa.overflow == 0 and sb.overflow == 0
        assert sa.contacts == sb.contacts and sa.solver_contacts == sb.solver_contacts, block
        assert a.state_hash() == b.state_hash(), f"after {60 * (block + 1)} steps"
    assert a.stats().contacts > 500_000


def test_full_size_batch_two_exact_schedules_agree(product):
    """BASELINE config 5 at full size (4096 worlds x 256 bodies): the block-per-group solver and the flow solver (two
    independent implementations of the reference's per-body constraint order) must produce identical worlds."""
    d = product.dll
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    hashes = []
    for mode in ("groups", "flow"):
        batch = d.CreatePhysicsWorldBatch(4096, s, b, c)
        views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in (0, 1, 777, 2048, 4095)]
        for i in range(4096):
            w = d.BatchWorld(batch, i)
            d.BbxSceneBuild(w, bb.SCENE_RLWORLD, 0, 7000 + i)
            if mode == "flow":
                d.BallboxSetSolverSchedule(w, bb.SOLVER_EXACT_FLOW)
        assert d.PhysicsStepBatch(batch, DT, 80) == 0, d.BatchGetLastError(batch)
        assert d.BatchDownload(batch, bb.DL_BODIES) == 0, d.BatchGetLastError(batch)
        st = bb.BallboxStats(); d.BatchGetStats(batch, bb.C.byref(st))
        assert st.overflow == 0
        hashes.append((st.contacts, tuple(v.state_hash() for v in views)))
        d.DestroyPhysicsWorldBatch(batch)
    assert hashes[0] == hashes[1] and hashes[0][0] > 1_000_000


@pytest.mark.parametrize("kind,n,steps", [(bb.SCENE_MIXED, 2500, 40), (bb.SCENE_SDF, 1200, 30)])
def test_in_kernel_replay_is_bit_exact(product, truth, monkeypatch, kind, n, steps):
    """BBX_REPLAY=0: iterations 1..I-1 inside k_solve_levels (the round-1 replay loop, still the path of schedules deeper
    than the staged replay's tables); the default staged replay kernel is what every other test runs."""
    monkeypatch.setenv("BBX_REPLAY", "0")
    lockstep(product, truth, kind, n, 31, steps, every=3)


@pytest.mark.parametrize("kind,n,steps", [(bb.SCENE_MIXED, 2500, 60), (bb.SCENE_SDF, 1200, 50), (bb.SCENE_BALLS, 4000, 40)])
def test_staged_replay_on_thin_schedules_is_bit_exact(product, truth, monkeypatch, kind, n, steps):
    """BBX_REPLAY=2 forces k_replay_staged where the default would replay inside k_solve_levels (thin schedules): small
    scenes with empty classes, ragged chunks and levels that most warps skip."""
    monkeypatch.setenv("BBX_REPLAY", "2")
    lockstep(product, truth, kind, n, 77, steps, every=3)


def test_schedule_deeper_than_the_staged_replay_tables(product, truth):
    """A column of 700 touching spheres on a static box: contact (i, i+1) shares a body with (i+1, i+2), so the exact
    schedule is ~700 levels deep, more than RS_MAXLEV = 256: k_solve_levels must run the replay itself."""
    def build(lib):
        w = lib.create_world(800, 4, 4000)
        w.set_gravity((0.0, -9.8, 0.0))
        w.w.settings.solver_iterations = 8
        g = w.add_box((0.0, -1.0, 0.0), (5.0, 1.0, 5.0))
        lib.dll.SetBoxStatic(w.ptr, g, True)
        for i in range(700):
            w.add_sphere((0.0, 0.45 + 0.9 * i, 0.0), 0.5)
        return w
    g, c = build(product), build(truth)
    for s in range(6):
        g.step(DT); c.step(DT)
        mm = parity.compare_worlds(g, c)
        assert not mm, f"step {s}: {mm[:3]}"
    assert g.stats().levels > 256, g.stats().levels


@pytest.mark.parametrize("kind,n,steps,factor,schedule", [
    (bb.SCENE_MIXED, 1500, 60, 1.0, bb.SOLVER_EXACT), (bb.SCENE_SDF, 800, 60, 0.85, bb.SOLVER_EXACT),
    (bb.SCENE_BALLS, 3000, 40, 1.0, bb.SOLVER_EXACT), (bb.SCENE_MIXED, 1000, 50, 0.9, bb.SOLVER_EXACT_FLOW)])
def test_warm_starting_is_bit_exact_with_its_sequential_definition(product, checkers, kind, n, steps, factor, schedule):
    """Extension (ballbox_ext.h BallboxSetWarmStarting): matched contacts start from the previous step's impulses and
    a pre-pass applies them.  Checked against the sequential definition in oracle/bbx_oracle.c, state and constraint
    list, every few steps.  (The reference itself has no warm starting: parity for this feature is pinned by the
    restatement only.)"""
    o = checkers["oracle"]
    g, c = product.build_scene(kind, n, 17), o.build_scene(kind, n, 17)
    g.set_warm_starting(factor); c.set_warm_starting(factor)
    product.dll.BallboxSetSolverSchedule(g.ptr, schedule)
    for s in range(steps):
        g.step(DT); c.step(DT, pruned=True)
        if s % 4 == 0 or s == steps - 1:
            mm = parity.compare_worlds(g, c)
            assert not mm, f"step {s}: {mm[:4]}"
    cold = product.build_scene(kind, n, 17)
    for s in range(steps):
        cold.step(DT)
    assert parity.compare_worlds(g, cold, constraints=False), "warm starting changed nothing"


def test_warm_starting_batch_and_world_reset(product, checkers):
    """Batched worlds (block-per-group solver, per-contact graph) with warm starting equal independent oracle worlds;
    resetting a world forgets its impulses: after the reset it equals a fresh oracle world started from the snapshot state."""
    o = checkers["oracle"]
    d = product.dll
    n_worlds = 20
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, c)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(n_worlds)]
    refs = []
    for i, v in enumerate(views):
        d.BbxSceneBuild(v.ptr, bb.SCENE_RLWORLD, 0, 900 + i)
        r = o.build_scene(bb.SCENE_RLWORLD, 0, 900 + i)
        r.set_warm_starting(1.0)
        refs.append(r)
    views[0].set_warm_starting(1.0)                                 # the batch steps with one setting
    assert d.BatchUpload(batch) == 0
    assert d.BatchSnapshot(batch) == 0                              # snapshot = initial state
    assert d.PhysicsStepBatch(batch, DT, 40) == 0, d.BatchGetLastError(batch)
    for r in refs:
        for _ in range(40):
            r.step(DT)
    ids = (bb.C.c_int * 2)(3, 11)
    assert d.BatchResetWorlds(batch, ids, 2) == 0
    for i in (3, 11):                                               # same initial state, no remembered impulses
        refs[i] = o.build_scene(bb.SCENE_RLWORLD, 0, 900 + i)
        refs[i].set_warm_starting(1.0)
    assert d.PhysicsStepBatch(batch, DT, 25) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    for i, (v, r) in enumerate(zip(views, refs)):
        for _ in range(25):
            r.step(DT)
        mm = parity.compare_worlds(v, r)
        assert not mm, f"world {i}: {mm[:4]}"
    d.DestroyPhysicsWorldBatch(batch)


def test_overflow_is_reported_by_a_bodies_only_download(product):
    """Resident stepping never looks at the host; a step that dropped contacts (max_collisions too small) must
    still surface at the next download, also when only the bodies are fetched."""
    w = product.build_scene(bb.SCENE_BALLS, 3000, 2, max_collisions=64)
    w.set_sync_mode(bb.SYNC_RESIDENT)
    w.step_n(DT, 5)
    with pytest.raises(bb.BallboxError, match="overflow|max_collisions"):
        w.download(bb.DL_BODIES)


@pytest.mark.parametrize("schedule", [bb.SOLVER_EXACT, bb.SOLVER_EXACT_FLOW])
def test_batch_edge_settings(product, truth, schedule):
    """Batched worlds with unusual settings: sphere-only worlds (no box arrays in use), no friction (the tangent
    sweep is skipped, :1462), a single solver iteration, zero iterations, restitution 1 and a manifold limit of 2."""
    d = product.dll
    cases = [dict(kind=bb.SCENE_BALLS, n=150, friction=0.0, iterations=3, restitution=0.3, mmax=8),
             dict(kind=bb.SCENE_RLWORLD, n=0, friction=0.5, iterations=1, restitution=1.0, mmax=2),
             dict(kind=bb.SCENE_RLWORLD, n=0, friction=0.2, iterations=0, restitution=0.0, mmax=8)]
    for case in cases:
        n_worlds = 18
        s, b, c = product.scene_capacity(case["kind"], case["n"])
        batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, c)
        views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(n_worlds)]
        refs = []
        for i, v in enumerate(views):
            d.BbxSceneBuild(v.ptr, case["kind"], case["n"], 300 + i)
            d.BallboxSetSolverSchedule(v.ptr, schedule)
            r = truth.build_scene(case["kind"], case["n"], 300 + i)
            for w in (v, r):
                st = w.w.settings
                st.friction = case["friction"]; st.solver_iterations = case["iterations"]
                st.restitution = case["restitution"]; st.manifold_max_contacts = case["mmax"]
            refs.append(r)
        assert d.PhysicsStepBatch(batch, DT, 45) == 0, d.BatchGetLastError(batch)
        assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
        for i in (0, 9, 17):
            for _ in range(45):
                refs[i].step(DT)
            mm = parity.compare_worlds(views[i], refs[i])
            assert not mm, f"{case}: world {i}: {mm[:3]}"
        d.DestroyPhysicsWorldBatch(batch)


def _resting_scene(lib):
    """two spheres and a box at rest in free space without gravity, next to a static box: they fall asleep after 0.5 s
    (a body with a contact never sleeps in the reference: every impulse application wakes it, ballbox.h:1373-1376)"""
    w = lib.create_world(8, 8, 256)
    w.set_gravity((0.0, 0.0, 0.0))
    g = w.add_box((0.0, -1.0, 0.0), (10.0, 1.0, 10.0)); lib.dll.SetBoxStatic(w.ptr, g, True)
    w.add_sphere((0.0, 0.6, 0.0), 0.5); w.add_sphere((3.0, 0.55, 1.0), 0.5)
    w.add_box((-3.0, 0.45, 0.0), (0.4, 0.4, 0.4))
    return w


def test_resident_force_wakes_a_sleeping_body(product, truth):
    """Add*Force/Torque on a body that went to sleep while the state is RESIDENT (ballbox.h:420-423: the call wakes the body,
    also for a zero force).  The host flags are stale then, so the wake has to reach the device."""
    g, c = _resting_scene(product), _resting_scene(truth)
    g.set_sync_mode(bb.SYNC_RESIDENT)
    g.step_n(DT, 100)
    for _ in range(100):
        c.step(DT)
    assert c.sphere_arrays()["sleeping"][:2].all() and c.box_arrays()["sleeping"][1], "scene did not fall asleep"
    for lib, w in ((product, g), (truth, c)):                            # no download: the host flags of g still say "awake"
        lib.dll.AddSphereForce(w.ptr, 0, bb.Vec3(40.0, -20.0, 0.0))      # a push (towards the static box)
        lib.dll.AddSphereTorque(w.ptr, 1, bb.Vec3(0.0, 0.0, 0.0))        # zero torque: still wakes (:434-437)
        lib.dll.AddBoxForce(w.ptr, 1, bb.Vec3(0.0, -30.0, 5.0))
    g.step_n(DT, 30)
    for _ in range(30):
        c.step(DT)
    g.download(bb.DL_BODIES)
    mm = parity.compare_worlds(g, c, constraints=False)
    assert not mm, mm[:3]
    assert abs(float(c.sphere_arrays()["pos"][0][0])) > 1e-3, "the push did not move the reference sphere"


def test_batch_force_wakes_a_sleeping_body(product, truth):
    """the same through PhysicsStepBatch (always resident)"""
    d = product.dll
    batch = d.CreatePhysicsWorldBatch(3, 8, 8, 256)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(3)]
    refs = [_resting_scene(truth) for _ in range(3)]
    for v in views:
        v.set_gravity((0.0, 0.0, 0.0))
        gb = v.add_box((0.0, -1.0, 0.0), (10.0, 1.0, 10.0)); d.SetBoxStatic(v.ptr, gb, True)
        v.add_sphere((0.0, 0.6, 0.0), 0.5); v.add_sphere((3.0, 0.55, 1.0), 0.5)
        v.add_box((-3.0, 0.45, 0.0), (0.4, 0.4, 0.4))
    assert d.PhysicsStepBatch(batch, DT, 100) == 0, d.BatchGetLastError(batch)
    for r in refs:
        for _ in range(100):
            r.step(DT)
        assert r.sphere_arrays()["sleeping"][:2].all()
    d.AddSphereForce(views[1].ptr, 0, bb.Vec3(40.0, -20.0, 0.0)); truth.dll.AddSphereForce(refs[1].ptr, 0, bb.Vec3(40.0, -20.0, 0.0))
    d.AddBoxTorque(views[2].ptr, 1, bb.Vec3(0.0, 9.0, 0.0)); truth.dll.AddBoxTorque(refs[2].ptr, 1, bb.Vec3(0.0, 9.0, 0.0))
    assert d.PhysicsStepBatch(batch, DT, 25) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES) == 0, d.BatchGetLastError(batch)
    for i, r in enumerate(refs):
        for _ in range(25):
            r.step(DT)
        mm = parity.compare_worlds(views[i], r, constraints=False)
        assert not mm, f"world {i}: {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)


def test_resident_edits_do_not_snap_back_to_stale_host_state(product, truth):
    """AddSphere / AddBox / Set*Mass / Set*Static / BallboxSetSyncMode on a RESIDENT world whose host arrays are stale:
    the device state must survive the re-upload the edit causes (the library downloads first)."""
    g, c = product.build_scene(bb.SCENE_MIXED, 200, 77), truth.build_scene(bb.SCENE_MIXED, 200, 77, )
    g.set_sync_mode(bb.SYNC_RESIDENT)
    g.step_n(DT, 40)
    for _ in range(40):
        c.step(DT)
    for lib, w in ((product, g), (truth, c)):                      # no download in between: the host arrays of g are 40 steps old
        lib.dll.SetSphereMass(w.ptr, 5, 3.5)
        lib.dll.SetBoxMass(w.ptr, 9, 0.25)
        lib.dll.SetSphereStatic(w.ptr, 11, True)
        lib.dll.SetBoxStatic(w.ptr, 12, True)
    g.step_n(DT, 20)
    for _ in range(20):
        c.step(DT)
    g.download(bb.DL_BODIES)
    mm = parity.compare_worlds(g, c, constraints=False)
    assert not mm, f"after the mass / static edits: {mm[:3]}"
    g.step_n(DT, 10)
    for _ in range(10):
        c.step(DT)
    g.set_sync_mode(bb.SYNC_COMPAT)                                # leaving RESIDENT with stale host arrays
    for _ in range(5):
        g.step(DT); c.step(DT)
    mm = parity.compare_worlds(g, c)
    assert not mm, f"after the mode change: {mm[:3]}"


def test_ragged_last_group_overflow_is_an_error_not_a_stray_write(product):
    """More worlds than the group solver has blocks, not a multiple of the worlds per block: the last group holds fewer
    worlds, so its slice of the queue / record arrays is shorter.  Its world produces more contacts than max_collisions:
    that has to end in the overflow error (it used to write past the allocations), and a batch whose last world
    stays within its capacity must step normally."""
    d = product.dll
    for dense in (False, True):
        n_worlds = 6 * 148 + 1                                     # one more than the solver's resident blocks on B200: 2 worlds per block, last block 1
        batch = d.CreatePhysicsWorldBatch(n_worlds, 12, 2, 24)
        assert batch
        for i in range(n_worlds):
            v = bb.World(product, d.BatchWorld(batch, i), owned=False)
            v.set_gravity((0.0, -9.8, 0.0))
            gb = v.add_box((0.0, -1.0, 0.0), (4.0, 1.0, 4.0)); d.SetBoxStatic(v.ptr, gb, True)
            v.add_sphere((0.0, 0.45, 0.0), 0.5)
            if i == n_worlds - 1 and dense:
                for j in range(10):                                # 11 spheres in one spot: 55 sphere pairs + 11 ground contacts > 24
                    v.add_sphere((0.01 * j, 0.45, 0.01 * j), 0.5)
        rc = d.PhysicsStepBatch(batch, DT, 3)
        d.BatchSynchronize(batch)
        rc2 = d.BatchDownload(batch, bb.DL_BODIES)
        err = d.BatchGetLastError(batch)
        if dense:
            assert (rc or rc2) and err and (b"overflow" in err.lower() or b"exceeds" in err.lower()), (rc, rc2, err)      # BBX_ERR_OVERFLOW
        else:
            assert rc == 0 and rc2 == 0 and not err, err
        d.DestroyPhysicsWorldBatch(batch)


def test_deferred_terminal_nodes_keep_the_trajectory(product, truth):
    """Thin schedules hold heavy terminal nodes (the probes of a box on the SDF) back for one extra last level.  This must
    not change a bit: an 8k-body SDF scene is stepped in lockstep with the reference while boxes land on the terrain,
    compared every 10 steps, and the mechanism must have been in use."""
    g, c = product.build_scene(bb.SCENE_SDF, 8000, 4711), truth.build_scene(bb.SCENE_SDF, 8000, 4711)
    deferred = 0
    for s in range(160):
        g.step(DT); c.step(DT, pruned=True)
        if s % 10 == 9:
            mm = parity.compare_worlds(g, c)
            assert not mm, f"step {s}: {mm[:3]}"
            g.stats()
            deferred = max(deferred, product.dll.BallboxDebugDeferred(g.ptr))
    assert deferred > 0, "no heavy terminal node was deferred: the test does not cover the mechanism"


def test_resident_stepping_is_deterministic_at_scale(product):
    """Two worlds from the same seed, resident, 200k bodies on the SDF while the contact graph grows from nothing to
    240k contacts: equal state hashes every 50 steps (a grid-wide race in the level kernel once showed up only here)."""
    a, b = product.build_scene(bb.SCENE_SDF, 200_000, 20261019), product.build_scene(bb.SCENE_SDF, 200_000, 20261019)
    for w in (a, b):
        w.set_sync_mode(bb.SYNC_RESIDENT)
    for block in range(4):
        a.step_n(DT, 50); b.step_n(DT, 50)
        a.download(bb.DL_BODIES); b.download(bb.DL_BODIES)
        assert a.state_hash() == b.state_hash(), f"after {50 * (block + 1)} steps"
    assert a.stats().contacts == b.stats().contacts > 200_000


def test_mid_size_pile_trajectory_in_lockstep(product, truth):
    """20k mixed bodies falling into a pile, in lockstep with the reference (pruned driver) for 120 steps, compared every
    20: the schedule grows from a few levels to dozens, with fat levels, manifold nodes and the staged replay's packing."""
    g, c = product.build_scene(bb.SCENE_MIXED, 20_000, 99), truth.build_scene(bb.SCENE_MIXED, 20_000, 99)
    for s in range(120):
        g.step(DT); c.step(DT, pruned=True)
        if s % 20 == 19:
            mm = parity.compare_worlds(g, c)
            assert not mm, f"step {s}: {mm[:3]}"
    assert g.stats().levels >= 10


@pytest.mark.parametrize("mode", ["0", "1"])
def test_both_cell_walks_find_the_same_pairs(product, truth, monkeypatch, mode):
    """BBX_FIND_PAIRS=0 (one lane per body) and =1 (candidates of a warp flattened) on a single world and on a batch; the
    default picks by batch size, so each environment setting forces the other kernel somewhere."""
    monkeypatch.setenv("BBX_FIND_PAIRS", mode)
    lockstep(product, truth, bb.SCENE_MIXED, 3000, 5, 30, every=3)
    d = product.dll
    s, b, c = product.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(20, s, b, c)
    views = [bb.World(product, d.BatchWorld(batch, i), owned=False) for i in range(20)]
    for i, v in enumerate(views):
        d.BbxSceneBuild(v.ptr, bb.SCENE_RLWORLD, 0, 900 + i)
    assert d.PhysicsStepBatch(batch, DT, 40) == 0, d.BatchGetLastError(batch)
    assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0, d.BatchGetLastError(batch)
    for i in (0, 7, 19):
        w = truth.build_scene(bb.SCENE_RLWORLD, 0, 900 + i)
        for _ in range(40):
            w.step(DT)
        mm = parity.compare_worlds(views[i], w)
        assert not mm, f"world {i}: {mm[:3]}"
    d.DestroyPhysicsWorldBatch(batch)


def _schedule(product, w, cap=2_000_000):
    """(level, constraints, body A, body B) of every node of the last solve, in visiting order."""
    import ctypes as C
    buf = (C.c_int * (4 * cap))()
    product.dll.BallboxDebugSchedule.argtypes = [bb.WP, C.POINTER(C.c_int), C.c_int]
    n = product.dll.BallboxDebugSchedule(w.ptr, buf, cap)
    assert 0 < n <= cap, n
    return np.frombuffer(buf, dtype=np.int32)[:4 * n].reshape(n, 4).copy()


def _assert_levels_are_independent_sets(sched):
    """No dynamic body twice in one level / colour: the property that lets a level be swept without atomics."""
    lv = np.concatenate([sched[:, 0], sched[:, 0]]).astype(np.int64)
    body = np.concatenate([sched[:, 2], sched[:, 3]]).astype(np.int64)
    keep = body >= 0
    key = lv[keep] * (1 << 32) + body[keep]
    assert len(np.unique(key)) == len(key)


@pytest.mark.gpu
def test_levels_of_the_exact_schedule_are_independent_sets(product):
    w = product.build_scene(bb.SCENE_MIXED, 30000, 5)
    w.set_sync_mode(bb.SYNC_RESIDENT)
    w.step_n(DT, 150); w.synchronize()
    st = w.stats()
    s = _schedule(product, w)
    assert s[:, 1].sum() == st.solver_contacts and s[:, 0].max() + 1 == st.levels
    _assert_levels_are_independent_sets(s)


@pytest.mark.gpu
def test_coloured_schedule_replays_colour_major(product, truth, monkeypatch):
    """BBX_SOLVER_COLOURED with the staged replay: iteration 0 sweeps the hash levels and takes greedy colours, iterations
    1.. replay colour classes.  Far fewer classes than exact levels, every class an independent set, every constraint
    scheduled exactly once, identical bits from two runs and from a different grid size; detection untouched."""
    def run(blocks=None, steps=60):
        if blocks: monkeypatch.setenv("BBX_SOLVE_BLOCKS", str(blocks))
        else: monkeypatch.delenv("BBX_SOLVE_BLOCKS", raising=False)
        w = product.build_scene(bb.SCENE_MIXED, 30000, 5)
        w.set_sync_mode(bb.SYNC_RESIDENT)
        w.step_n(DT, 150); w.synchronize()                      # settle with the exact schedule
        exact_levels = w.stats().levels
        product.dll.BallboxSetSolverSchedule(w.ptr, bb.SOLVER_COLOURED)
        w.step_n(DT, steps); w.synchronize()
        w.download(bb.DL_BODIES | bb.DL_CONSTRAINTS)
        return w, exact_levels
    a, exact_levels = run()
    st = a.stats()
    assert st.overflow == 0
    assert 4 <= st.levels <= 24 and st.levels < exact_levels, (st.levels, exact_levels)
    s = _schedule(product, a)
    assert s[:, 1].sum() == st.solver_contacts and s[:, 0].max() + 1 == st.levels
    _assert_levels_are_independent_sets(s)
    b, _ = run()
    assert a.state_hash() == b.state_hash()
    c, _ = run(blocks=96)
    assert a.state_hash() == c.state_hash()
    # detection does not depend on the schedule: one coloured step from a state the checker shares
    g = product.build_scene(bb.SCENE_MIXED, 4000, 9)
    product.dll.BallboxSetSolverSchedule(g.ptr, bb.SOLVER_COLOURED)
    t = truth.build_scene(bb.SCENE_MIXED, 4000, 9)
    g.step(DT); t.step(DT, pruned=True)
    kg, kt = g.constraints(), t.constraints()
    for f in ("a_type", "a_index", "b_type", "b_index", "point", "normal", "penetration", "normal_mass", "position_bias"):
        assert parity.first_mismatch(kg[f], kt[f]) is None, f


@pytest.mark.gpu
def test_coloured_schedule_falls_back_when_colours_run_out(product, monkeypatch):
    """More than 64 colours cannot be represented: the step then replays the hash levels (identity re-bucketing).  Forced
    here by BBX_COLOUR_LIMIT; the result must equal the coloured schedule with the in-kernel replay (BBX_REPLAY=0),
    which sweeps the same hash levels in every iteration."""
    def run(env):
        for k in ("BBX_COLOUR_LIMIT", "BBX_REPLAY"): monkeypatch.delenv(k, raising=False)
        for k, v in env.items(): monkeypatch.setenv(k, v)
        w = product.build_scene(bb.SCENE_MIXED, 6000, 11)
        product.dll.BallboxSetSolverSchedule(w.ptr, bb.SOLVER_COLOURED)
        w.set_sync_mode(bb.SYNC_RESIDENT)
        w.step_n(DT, 80); w.synchronize()
        w.download(bb.DL_BODIES | bb.DL_CONSTRAINTS)
        return w
    a = run({"BBX_COLOUR_LIMIT": "3", "BBX_REPLAY": "2"})
    b = run({"BBX_REPLAY": "0"})
    assert a.stats().contacts > 5000
    assert a.state_hash() == b.state_hash()
    assert not parity.compare_worlds(a, b, constraints=True)
