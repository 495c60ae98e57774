"""Profiling driver: build the C3 scene, settle, then run a few steps inside a
cudaProfilerStart/Stop window (use with `ncu --profile-from-start off`)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb

ap = argparse.ArgumentParser()
ap.add_argument("--bodies", type=int, default=1_000_000)
ap.add_argument("--settle", type=int, default=1500)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--scene", type=int, default=bb.SCENE_MIXED)
ap.add_argument("--friction", type=float, default=None, help="override friction for the timed steps (timing experiment)")
ap.add_argument("--iterations", type=int, default=None)
ap.add_argument("--coloured", action="store_true", help="timed steps with BBX_SOLVER_COLOURED (settled with the exact schedule)")
a = ap.parse_args()
lib = bbx.load()
w = lib.build_scene(a.scene, a.bodies, 20261019)
w.set_sync_mode(bb.SYNC_RESIDENT)
w.upload()
done = 0
t0 = time.time()
while done < a.settle:
    k = min(100, a.settle - done); w.step_n(1 / 60, k); w.synchronize(); done += k
st = w.stats()
print(f"settled {a.settle} steps in {time.time()-t0:.1f}s: contacts={st.contacts} solver={st.solver_contacts} touched={st.touched_bodies} "
      f"levels={st.levels} pairs={st.candidate_pairs} sb={st.reserved[0]} bb={st.reserved[1]} ovf={st.overflow}", flush=True)
if a.friction is not None: w.w.settings.friction = a.friction
if a.iterations is not None: w.w.settings.solver_iterations = a.iterations
if a.coloured:
    lib.dll.BallboxSetSolverSchedule(w.ptr, bb.SOLVER_COLOURED)
    w.step_n(1 / 60, 3); w.synchronize()          # (first coloured step allocates the second buffer set)
w.set_stage_timing(True); w.stage_timing()
torch.cuda.synchronize()
torch.cuda.profiler.start()
w.step_n(1 / 60, a.steps); w.synchronize()
torch.cuda.profiler.stop()
tm, n = w.stage_timing()
rms, rn = w.replay_timing()
print({k: round(v / max(n, 1), 4) for k, v in tm.items()}, "ms/step over", n, "| k_replay_staged alone: %.4f ms" % (rms / max(rn, 1)), flush=True)

import ctypes as C
import numpy as np
L = st.levels if False else w.stats().levels
cnt = (C.c_int * (8 * 4096))(); ns = (C.c_ulonglong * 4097)()
lib.dll.BallboxDebugLevels.argtypes = [bb.WP, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong), C.c_int]
lib.dll.BallboxDebugLevels(w.ptr, cnt, ns, 4096)
fd = np.frombuffer(ns, dtype=np.uint64)[4088:4096].astype(np.float64)
if os.environ.get("BBX_TRACE_IT0"):
    NW = 2 * 148 * 8
    tr = (C.c_ulonglong * (8 * NW))()
    lib.dll.BallboxDebugWarpTrace.argtypes = [bb.WP, C.POINTER(C.c_ulonglong), C.c_int]
    lib.dll.BallboxDebugWarpTrace(w.ptr, tr, 2 * NW)
    t = np.frombuffer(tr, dtype=np.uint64).reshape(NW, 8).astype(np.int64)
    ok = (t[:, 1] > 0) & (t[:, 4] > 0)
    t0 = t[ok, 0].min()
    names = ["node+links", "record+bodies", "sweep+stores", "in-degree atomics", "rest of level (other rounds, flushes)", "grid barrier"]
    print(f"iteration 0, level {int(os.environ['BBX_TRACE_LEVEL']) - 1}: {ok.sum()} warps with a first chunk; rounds {int(t[ok, 6].max())}")
    for k, nm in enumerate(names):
        d = (t[ok, k + 1] - t[ok, k]) / 1e3
        print(f"  {nm}: mean {d.mean():.2f} us  p90 {np.percentile(d, 90):.2f}  max {d.max():.2f}")
    print(f"  level, first warp start to last barrier exit: {(t[ok, 7].max() - t0) / 1e3:.2f} us")
elif os.environ.get("BBX_TRACE_LEVEL"):
    NW = 2 * 148 * 8
    tr = (C.c_ulonglong * (4 * NW))()
    lib.dll.BallboxDebugWarpTrace.argtypes = [bb.WP, C.POINTER(C.c_ulonglong), C.c_int]
    lib.dll.BallboxDebugWarpTrace(w.ptr, tr, NW)
    t = np.frombuffer(tr, dtype=np.uint64).reshape(NW, 4).astype(np.int64)
    t0 = t[:, 0].min()
    work = t[:, 2] > 0
    print(f"warp trace of level {int(os.environ['BBX_TRACE_LEVEL']) - 1}, iteration 1: {work.sum()} of {NW} warps with work; "
          f"level start spread {(t[:, 0].max() - t0) / 1e3:.2f} us; last warp ends at {(t[work, 2].max() - t0) / 1e3:.2f} us")
    busy = (t[:, 2] - t[:, 0])[work] / 1e3; wait = (t[:, 1] - t[:, 0])[work] / 1e3
    print(f"  busy us: mean {busy.mean():.2f} p50 {np.median(busy):.2f} p90 {np.percentile(busy, 90):.2f} max {busy.max():.2f} | first data wait us: mean {wait.mean():.2f} max {wait.max():.2f}")
    m = ((t[:, 3] >> 16) & 0xffff)[work]; nc = (t[:, 3] & 0xffff)[work]; cls = (t[:, 3] >> 32)[work]
    order = np.argsort(-busy)[:6]
    widx = np.nonzero(work)[0]
    print("  slowest warps (warp index, class, chunks, constraints, first data wait us, busy us):",
          [(int(widx[o]), int(cls[o]), int(nc[o]), int(m[o]), round(float(wait[o]), 2), round(float(busy[o]), 2)) for o in order])
    for kk in sorted(set(cls.tolist())):                      # per node class: what a 32-node chunk costs (calibration of rs_cost in bbx_replay.cu)
        sel = cls == kk
        per = (busy[sel] - wait[sel]) / np.maximum(nc[sel], 1)
        print(f"  class {kk}: warps {sel.sum()} chunks/warp {nc[sel].mean():.2f} (busy - first wait) per chunk us: mean {per.mean():.2f} p90 {np.percentile(per, 90):.2f} | busy mean {busy[sel].mean():.2f} max {busy[sel].max():.2f}")
    for mm in sorted(set(m.tolist())):
        sel = m == mm
        print(f"  head node m={mm}: warps {sel.sum()} chunks/warp {nc[sel].mean():.2f} busy mean {busy[sel].mean():.2f} max {busy[sel].max():.2f} end mean {((t[work, 2][sel] - t0) / 1e3).mean():.2f}")
ns0 = np.frombuffer(ns, dtype=np.uint64)[2048:2048 + L].astype(np.int64)
if L > 1 and ns0.min() > 0:
    d0 = np.diff(ns0) / 1e3
    print("iteration 0 level time us (levels 1..):", np.round(d0[:40], 1).tolist(), "sum ms:", round(d0.sum() / 1e3, 4))
cnt = np.frombuffer(cnt, dtype=np.int32).reshape(-1, 8)[:L]; ns = np.frombuffer(ns, dtype=np.uint64)[:L + 1].astype(np.int64)
dt_us = np.diff(ns)[1:] / 1e3
print("levels", L, "nodes/level (first 12):", cnt.sum(1)[:12].tolist())
print("class totals:", cnt.sum(0).tolist())
print("level time us (levels 1..):", np.round(dt_us[:40], 1).tolist(), "| levels 41..:", np.round(dt_us[40:], 1).tolist())
print("sum level times (iteration 1) ms:", dt_us.sum() / 1e3, "tail levels (<1000 nodes):", int((cnt.sum(1) < 1000).sum()))
for l in range(0, L, 4):
    print(l, cnt[l].tolist(), "us=%.1f" % (dt_us[l - 1] if l >= 1 and l - 1 < len(dt_us) else -1))
