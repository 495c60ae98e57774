"""Where the drop-in call's time goes at 1M bodies: python tools/profile_e2e.py [settle]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
lib = bbx.load()
w = lib.build_scene(bb.SCENE_MIXED, 1_000_000, 20261019)
w.set_sync_mode(bb.SYNC_RESIDENT); w.upload()
done = 0
while done < settle:
    k = min(100, settle - done); w.step_n(1 / 60, k); w.synchronize(); done += k
def timed(f, n=3):
    f(); w.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    w.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
w.download(bb.DL_BODIES)          # host arrays = settled state, so that the timed uploads do not reset the scene
print("upload              %.2f ms" % timed(lambda: w.upload()))
print("step (resident)     %.2f ms" % timed(lambda: w.step_n(1 / 60, 1)))
print("download bodies     %.2f ms" % timed(lambda: w.download(bb.DL_BODIES)))
print("download constraints %.2f ms (export + sort + gather + %d x 104 B D2H)" % (timed(lambda: w.download(bb.DL_CONSTRAINTS)), len(w.constraints())))
w.set_sync_mode(bb.SYNC_COMPAT)
print("PhysicsStep compat  %.2f ms" % timed(lambda: w.step(1 / 60)))
