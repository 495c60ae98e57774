"""Diagnostic: step identical scenes on the GPU library and on the CPU checker and
report the first divergence.  Usage: python tools/gpu_check.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb
from tests import oracles, parity

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lib = bbx.load()
chk = oracles.load_all()
truth = chk.get("ref") or chk["oracle"]
print("product:", lib.backend_name(), "devices:", lib.dll.BallboxDeviceCount(), "| checker:", truth.backend_name(), flush=True)
bad = 0
for kind, n, pruned in ((bb.SCENE_README, 0, False), (bb.SCENE_BALLS, 300, False), (bb.SCENE_MIXED, 200, False),
                        (bb.SCENE_SDF, 200, False), (bb.SCENE_RLWORLD, 0, False), (bb.SCENE_MIXED, 3000, True),
                        (bb.SCENE_BALLS, 20000, True)):
    g = lib.build_scene(kind, n, 11); c = truth.build_scene(kind, n, 11)
    t0 = time.time(); ok = True
    for s in range(steps):
        if kind == bb.SCENE_README and s < 30:
            lib.dll.AddSphereForce(g.ptr, 0, bb.Vec3(2, 0, 0)); truth.dll.AddSphereForce(c.ptr, 0, bb.Vec3(2, 0, 0))
        g.step(1 / 60); c.step(1 / 60, pruned=pruned)
        mm = parity.compare_worlds(g, c)
        if mm:
            print(f"scene {kind} n={n}: MISMATCH at step {s}:"); [print("   ", m) for m in mm[:6]]
            ok = False; bad += 1
            break
    st = g.stats()
    print(f"scene {kind} n={n}: {'bit-exact' if ok else 'DIFF'} over {s + 1} steps, contacts={st.contacts} levels={st.levels} "
          f"cand={st.candidate_pairs} ovf={st.overflow} launches={st.kernel_launches} ({time.time() - t0:.1f}s)", flush=True)
sys.exit(1 if bad else 0)
