"""Warm starting at BASELINE config 3 (extension, ballbox_ext.h): cost per step and what it buys.
Every row settles the same scene with the reference's settings (8 cold sweeps), switches to the row's setting, runs
200 more steps and reports the step time and the mean penetration of the sphere contacts (spheres are all dynamic).
Usage: python tools/warm_report.py [bodies] > profiles/rNN_warm_start.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
lib = bbx.load()
print(f"# Warm starting, config 3 ({n} bodies), 600 settle steps (cold, 8 sweeps) + 200 steps with the row's setting\n")
print("| warm factor | sweeps | ms/step | solve kernel ms | presolve+graph ms | mean sphere-contact penetration | max | kinetic energy |")
print("|---|---|---|---|---|---|---|---|")
for factor, iters in ((0.0, 8), (0.0, 4), (0.0, 2), (1.0, 8), (1.0, 4), (1.0, 2), (1.0, 1), (0.9, 4)):
    w = lib.build_scene(bb.SCENE_MIXED, n, 20261019)
    w.set_sync_mode(bb.SYNC_RESIDENT); w.upload()
    w.step_n(1 / 60, 600); w.synchronize()
    w.w.settings.solver_iterations = iters
    w.set_warm_starting(factor)
    w.step_n(1 / 60, 190); w.synchronize()
    w.set_stage_timing(True); w.stage_timing()
    w.step_n(1 / 60, 10); w.synchronize()
    tm, k = w.stage_timing()
    w.download(bb.DL_BODIES | bb.DL_CONSTRAINTS)
    c = w.constraints()
    sc = c[c["a_type"] == 0]
    snap = w.snapshot()
    ke = float((snap["s_vel"].astype(np.float64) ** 2).sum() + (snap["b_vel"].astype(np.float64) ** 2).sum())
    print(f"| {factor} | {iters} | {sum(tm.values()) / k:.2f} | {tm['solve_kernel'] / k:.2f} | {tm['presolve_graph'] / k:.2f} | "
          f"{float(sc['penetration'].mean()):.5f} | {float(sc['penetration'].max()):.3f} | {ke:.3e} |", flush=True)
    w.destroy()
