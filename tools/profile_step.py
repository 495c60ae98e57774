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
w.set_stage_timing(True); w.stage_timing()
torch.cuda.synchronize()
torch.cuda.profiler.start()
w.step_n(1 / 60, a.steps); w.synchronize()
torch.cuda.profiler.stop()
tm, n = w.stage_timing()
print({k: round(v / max(n, 1), 4) for k, v in tm.items()}, "ms/step over", n, flush=True)
