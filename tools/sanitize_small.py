"""Small scenes through every solver path, for compute-sanitizer (memcheck): python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb
lib = bbx.load(); d = lib.dll
DT = 1 / 60
for sched in (bb.SOLVER_EXACT, bb.SOLVER_EXACT_FLOW, bb.SOLVER_COLOURED):
    for kind, n in ((bb.SCENE_MIXED, 400), (bb.SCENE_SDF, 300), (bb.SCENE_BALLS, 500)):
        w = lib.build_scene(kind, n, 5)
        d.BallboxSetSolverSchedule(w.ptr, sched)
        for _ in range(6):
            w.step(DT)
        print("schedule", sched, "scene", kind, "contacts", w.stats().contacts, "hash", hex(w.state_hash()), flush=True)
        w.destroy()
s, b, c = lib.scene_capacity(bb.SCENE_RLWORLD, 0)
batch = d.CreatePhysicsWorldBatch(20, s, b, c)
for i in range(20):
    d.BbxSceneBuild(d.BatchWorld(batch, i), bb.SCENE_RLWORLD, 0, 100 + i)
assert d.PhysicsStepBatch(batch, DT, 12) == 0
assert d.BatchDownload(batch, bb.DL_BODIES | bb.DL_CONSTRAINTS) == 0
print("batch ok", flush=True)
d.DestroyPhysicsWorldBatch(batch)
