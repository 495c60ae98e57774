import sys, numpy as np
sys.path.insert(0, "/root/repo")
import ballbox_b200 as bbx
from ballbox_b200 import binding as bb
from tests import oracles, parity
lib = bbx.load(); ref = oracles.load_all()
truth = ref.get("ref") or ref["oracle"]
def build(L, extra):
    w = L.create_world(64, 64, 4096); w.set_gravity((0, -9.8, 0)); d = L.dll
    g = w.add_box((0, -1, 0), (8, 1, 8)); d.SetBoxStatic(w.ptr, g, True)
    for i in range(20):
        w.add_sphere((-3 + 0.35 * i, 1.5 + 0.1 * (i % 3), -2 + 0.2 * i), 0.3)
        w.add_box((-3 + 0.3 * i, 3.0, 2 - 0.2 * i), (0.2, 0.3, 0.25), w.quat_axis_angle((1, i + 1, 2), 0.3 * i))
    extra(w)
    return w
cases = {
  "far 1e6": lambda w: w.add_sphere((1e6, 3.0, -2e6), 0.5),
  "far 1e12": lambda w: w.add_sphere((1e12, 3.0, 0), 0.5),
  "far 1e30": lambda w: w.add_sphere((1e30, 3.0, 0), 0.5),
  "nan pos": lambda w: w.add_sphere((float("nan"), 3.0, 0), 0.5),
  "inf pos": lambda w: w.add_sphere((float("inf"), 3.0, 0), 0.5),
  "huge radius": lambda w: w.add_sphere((0, 50.0, 0), 1e6),
  "nan box": lambda w: w.add_box((float("nan"), 3.0, 0), (0.3, 0.3, 0.3)),
}
for name, ex in cases.items():
    try:
        g, c = build(lib, ex), build(truth, ex)
        bad = None
        for s in range(40):
            g.step(1/60); c.step(1/60)
            sa, sb = g.snapshot(), c.snapshot()
            for f in ("s_pos", "s_vel", "b_pos", "b_vel", "b_rot"):
                if not np.array_equal(sa[f], sb[f], equal_nan=True): bad = (s, f); break
            if bad or len(g.constraints()) != len(c.constraints()): bad = bad or (s, "count"); break
        print(name, "OK" if not bad else f"MISMATCH {bad}", "contacts", len(g.constraints()))
    except Exception as e:
        print(name, "EXC", str(e)[:150])
