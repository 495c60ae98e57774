"""Run-to-run determinism of resident stepping: two worlds built from the same seed are stepped side by side and their
state hashes compared every `every` steps.  Usage: python tools/determinism_check.py [scene] [bodies] [steps] [every]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb
scene = int(sys.argv[1]) if len(sys.argv) > 1 else bb.SCENE_SDF
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
every = int(sys.argv[4]) if len(sys.argv) > 4 else 20
lib = bbx.load()
a, b = lib.build_scene(scene, n, 20261019), lib.build_scene(scene, n, 20261019)
for w in (a, b):
    w.set_sync_mode(bb.SYNC_RESIDENT); w.upload()
done = 0
while done < steps:
    a.step_n(1 / 60, every); b.step_n(1 / 60, every); done += every
    a.download(bb.DL_BODIES); b.download(bb.DL_BODIES)
    ha, hb = a.state_hash(), b.state_hash()
    sa, sb = a.stats(), b.stats()
    print(f"step {done}: hash {'EQUAL' if ha == hb else 'DIFFERENT'} contacts {sa.contacts} / {sb.contacts} levels {sa.levels} / {sb.levels}", flush=True)
    if ha != hb:
        break
