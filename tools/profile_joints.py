"""Cost of the joints / springs extension at scale: N spheres on rods (N components), N/2 hinged box pairs and N/4
springs over a floor, resident stepping, stage timing with and without the joints.
Usage: python tools/profile_joints.py [N] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
lib = bbx.load(); d = lib.dll
side = int(N ** 0.5) + 1


def scene(joints):
    w = lib.create_world(N + N // 4 + 8, N + 8, 8 * N)
    w.set_gravity((0.0, -9.81, 0.0))
    f = w.add_box((side, -0.5, side), (2.0 * side, 0.5, 2.0 * side)); d.SetBoxStatic(w.ptr, f, True)
    t0 = time.time()
    for k in range(N):
        x, z = 2.0 * (k % side), 2.0 * (k // side)
        i = w.add_sphere((x + 0.8, 3.0, z), 0.2)
        if joints: w.add_distance_joint((0, i), (-1, 0), (x + 0.8, 3.0, z), (x, 3.0, z))
    for k in range(N // 2):
        x, z = 2.0 * (k % side) + 0.3, 2.0 * (k // side) + 1.0
        a = w.add_box((x, 1.0, z), (0.25, 0.1, 0.1)); b = w.add_box((x + 0.5, 1.0, z), (0.25, 0.1, 0.1))
        if joints: w.add_hinge_joint((1, a), (1, b), (x + 0.25, 1.0, z), (0, 0, 1), 0.08)
    for k in range(N // 4):
        x, z = 2.0 * (k % side) + 1.2, 2.0 * (k // side) + 0.5
        i = w.add_sphere((x, 2.0, z), 0.15)
        if joints: w.add_spring((0, i), (-1, 0), (x, 2.0, z), (x, 4.0, z), 1.5, 40.0, 1.0)
    print(f"scene built in {time.time() - t0:.1f} s: joints {d.BallboxJointCount(w.ptr)} springs {d.BallboxSpringCount(w.ptr)}")
    return w


for joints in (False, True):
    w = scene(joints)
    w.set_sync_mode(bb.SYNC_RESIDENT)
    w.step_n(1 / 60, 60); w.synchronize()
    w.set_stage_timing(True)
    w.stage_timing()
    t0 = time.time()
    w.step_n(1 / 60, steps); w.synchronize()
    wall = (time.time() - t0) * 1e3 / steps
    st = w.stats()
    tm, ns = w.stage_timing()
    print("joints" if joints else "plain ", {k: round(v / max(ns, 1), 4) for k, v in tm.items()}, f"wall {wall:.3f} ms/step; contacts {st.contacts} levels {st.levels} launches/step {st.kernel_launches / st.steps:.1f}")
    w.destroy()
