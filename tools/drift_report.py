"""Drift of the coloured (reordered) solver schedule against the exact schedule, on the GPU.
Usage: python tools/drift_report.py [bodies] -> markdown on stdout (profiles/r02_drift.md; round 1: r01_drift.md)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
lib = bbx.load()
DT = 1 / 60


def make(schedule, settle):
    w = lib.build_scene(bb.SCENE_MIXED, n, 4711)
    w.set_sync_mode(bb.SYNC_RESIDENT)
    w.upload()
    w.step_n(DT, settle)                       # settle with the EXACT schedule: both start from the same state
    lib.dll.BallboxSetSolverSchedule(w.ptr, schedule)
    return w


def stats(a, b, key, mask=None):
    d = np.linalg.norm(a[key].astype(np.float64) - b[key].astype(np.float64), axis=1)
    if mask is not None:
        d = d[mask]
    return d.mean(), np.percentile(d, 99), d.max()


print(f"# Drift of BBX_SOLVER_COLOURED vs BBX_SOLVER_EXACT (config-3 shape, {n} bodies, settled 600 steps, GPU)\n")
print("Both worlds are settled with the exact schedule, then stepped with different schedules. Differences are"
      " Euclidean norms per body (m, m/s); the exact schedule itself is bit-identical to the reference.\n")
print("| steps | levels exact | levels coloured | pos mean | pos p99 | pos max | vel mean | vel p99 | vel max |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
e, c = make(bb.SOLVER_EXACT, 600), make(bb.SOLVER_COLOURED, 600)
done = 0
for target in (1, 10, 100, 1000):
    e.step_n(DT, target - done); c.step_n(DT, target - done); done = target
    e.download(bb.DL_BODIES); c.download(bb.DL_BODIES)
    se, sc = e.snapshot(), c.snapshot()
    pos = np.concatenate([se["s_pos"], se["b_pos"]]), np.concatenate([sc["s_pos"], sc["b_pos"]])
    vel = np.concatenate([se["s_vel"], se["b_vel"]]), np.concatenate([sc["s_vel"], sc["b_vel"]])
    dp = np.linalg.norm(pos[0].astype(np.float64) - pos[1], axis=1); dv = np.linalg.norm(vel[0].astype(np.float64) - vel[1], axis=1)
    print(f"| {target} | {e.stats().levels} | {c.stats().levels} | {dp.mean():.3e} | {np.percentile(dp, 99):.3e} | {dp.max():.3e} | "
          f"{dv.mean():.3e} | {np.percentile(dv, 99):.3e} | {dv.max():.3e} |")
# aggregate state after the last target: the pile is chaotic (bodies end elsewhere) but statistically the same
for name, w in (("exact", e), ("coloured", c)):
    sn = w.snapshot()
    v = np.concatenate([sn["s_vel"], sn["b_vel"]]).astype(np.float64); y = np.concatenate([sn["s_pos"], sn["b_pos"]])[:, 1].astype(np.float64)
    print(f"\n{name} after {done} steps: contacts {w.stats().contacts}, mean speed {np.linalg.norm(v, axis=1).mean():.4e} m/s, "
          f"mean height {y.mean():.4f} m, sleeping {int(sn['s_sleeping'].sum() + sn['b_sleeping'].sum()) if 's_sleeping' in sn else -1}")
# timing of both schedules on this scene
for name, w in (("exact", e), ("coloured", c)):
    w.set_stage_timing(True); w.stage_timing(); w.step_n(DT, 20); t, k = w.stage_timing()
    print(f"\n{name}: solve kernel {t['solve_kernel'] / k:.3f} ms/step, whole step {sum(t.values()) / k:.3f} ms/step, levels {w.stats().levels}")
# determinism of the coloured schedule
c2 = make(bb.SOLVER_COLOURED, 600); c3 = make(bb.SOLVER_COLOURED, 600)
c2.step_n(DT, 50); c3.step_n(DT, 50); c2.download(bb.DL_BODIES); c3.download(bb.DL_BODIES)
print(f"\ncoloured schedule deterministic across runs: {c2.state_hash() == c3.state_hash()}")
