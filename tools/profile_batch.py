"""Stage timing of the batched-worlds path (config 5). Usage: python tools/profile_batch.py [worlds] [settle] [steps]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
iters = int(sys.argv[4]) if len(sys.argv) > 4 else None
lib = bbx.load(); d = lib.dll
s, b, c = lib.scene_capacity(bb.SCENE_RLWORLD, 0)
batch = d.CreatePhysicsWorldBatch(W, s, b, c)
for i in range(W):
    d.BbxSceneBuild(d.BatchWorld(batch, i), bb.SCENE_RLWORLD, 0, 7000 + i)
d.BatchUpload(batch)
done = 0
while done < settle:
    k = min(50, settle - done); d.PhysicsStepBatch(batch, 1 / 60, k); d.BatchSynchronize(batch); done += k
if iters is not None:
    d.SetSolverIterations(d.BatchWorld(batch, 0), iters)      # the batch steps with world 0's settings
d.BatchSetStageTiming(batch, 1)
out = (C.c_float * 6)(); n = C.c_int()
d.BatchGetStageTiming(batch, out, C.byref(n))
try:
    import torch; torch.cuda.profiler.start()
except Exception:
    torch = None
d.PhysicsStepBatch(batch, 1 / 60, steps); d.BatchSynchronize(batch)
if torch: torch.cuda.profiler.stop()
d.BatchGetStageTiming(batch, out, C.byref(n))
st = bb.BallboxStats(); d.BatchGetStats(batch, C.byref(st))
e = d.BatchGetLastError(batch)
print("error:", e)
names = bb.World.STAGES
print({k: round(out[i] / max(n.value, 1), 4) for i, k in enumerate(names)}, "total", round(sum(out) / max(n.value, 1), 3), "ms/step over", n.value)
print("contacts", st.contacts, "solver", st.solver_contacts, "levels", st.levels, "pairs", st.candidate_pairs, "ovf", st.overflow)
d.BatchDownload(batch, bb.DL_BODIES)
print("checksum %016x" % (sum(int(d.BbxStateHash(d.BatchWorld(batch, i))) for i in range(W)) & (2 ** 64 - 1)), "BBX_GRP_SMEM", os.environ.get("BBX_GRP_SMEM"))
if os.environ.get("BBX_SOLVER_FLOW"):
    import numpy as np
    cnt = (C.c_int * (8 * 4096))(); ns = (C.c_ulonglong * 4097)()
    d.BallboxDebugLevels.argtypes = [bb.WP, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong), C.c_int]
    d.BallboxDebugLevels(d.BatchWorld(batch, 0), cnt, ns, 4096)
    ud = int(np.frombuffer(ns, dtype=np.uint64)[4080])
    its = d.BatchWorld(batch, 0).contents.settings.solver_iterations
    print(f"flow solver: dependency depth of ALL iterations together = {ud} constraints; iteration 0 alone {st.levels}; {its} iterations one after the other {st.levels * its}")
if os.environ.get("BBX_TRACE_LEVEL") == "-1":
    # per-block phase stamps of k_solve_groups (first 256 groups): where the slowest block of a small batch spends its time
    import numpy as np
    cnt = (C.c_int * (8 * 4096))(); ns = (C.c_ulonglong * 4097)()
    d.BallboxDebugLevels.argtypes = [bb.WP, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong), C.c_int]
    d.BallboxDebugLevels(d.BatchWorld(batch, 0), cnt, ns, 4096)
    a = np.frombuffer(ns, dtype=np.uint64)[:2048].reshape(256, 8).astype(np.int64)[:min(256, W)]
    a = a[a[:, 4] > 0]
    slow = int(np.argmax(a[:, 4]))
    for k, v, sc in (("iteration 0 (us)", a[:, 1] - a[:, 0], 1e-3), ("iterations 1.. (us)", a[:, 4] - a[:, 1], 1e-3), ("kernel (us)", a[:, 4] - a[:, 0], 1e-3),
                     ("levels", a[:, 5], 1), ("constraints", a[:, 7], 1)):
        print(f"  group {k:20s} mean {v.mean() * sc:9.2f}  max {v.max() * sc:9.2f}  last block to finish {v[slow] * sc:9.2f}")
