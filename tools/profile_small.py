"""Launch-bound regime: wall-clock per resident step of small scenes. Usage: python tools/profile_small.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb
lib = bbx.load(); d = lib.dll


def timed_world(kind, n, steps=2000):
    w = lib.build_scene(kind, n, 5)
    w.set_sync_mode(bb.SYNC_RESIDENT); w.upload()
    w.step_n(1 / 60, 200); w.synchronize()
    t0 = time.perf_counter()
    w.step_n(1 / 60, steps); w.synchronize()
    dt = (time.perf_counter() - t0) / steps
    st = w.stats()
    print(f"scene {kind} n={n}: {dt * 1e6:8.1f} us/step  contacts {st.contacts} levels {st.levels}")


def timed_batch(worlds, steps=1000):
    s, b, c = lib.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = d.CreatePhysicsWorldBatch(worlds, s, b, c)
    for i in range(worlds):
        d.BbxSceneBuild(d.BatchWorld(batch, i), bb.SCENE_RLWORLD, 0, 7000 + i)
    d.BatchUpload(batch)
    d.PhysicsStepBatch(batch, 1 / 60, 200); d.BatchSynchronize(batch)
    t0 = time.perf_counter()
    d.PhysicsStepBatch(batch, 1 / 60, steps); d.BatchSynchronize(batch)
    dt = (time.perf_counter() - t0) / steps
    print(f"batch {worlds} x 256: {dt * 1e6:8.1f} us/step  {worlds * 256 / dt:.3e} body-steps/s")
    d.DestroyPhysicsWorldBatch(batch)


if len(sys.argv) > 1 and sys.argv[1] == "worlds":
    for n in (1000, 4000, 10000, 30000, 100000):
        timed_world(bb.SCENE_MIXED, n, 500)
else:
    timed_world(bb.SCENE_README, 0)
    timed_world(bb.SCENE_BALLS, 1000)
    timed_world(bb.SCENE_MIXED, 1000)
    timed_world(bb.SCENE_MIXED, 10000)
    for W in (1, 16, 128):
        timed_batch(W)
