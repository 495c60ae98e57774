"""Generate the golden vectors under tests/golden/ from oracle/_ref (the UNMODIFIED
reference header compiled by oracle/Makefile).  Run in the build container, where
/root/reference exists:  python tests/golden/make_golden.py
The .npz files are committed; tests never read /root/reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ballbox_b200 import binding as bb          # noqa: E402
from tests import oracles                        # noqa: E402

CASES = {   # name: (scene kind, n, seed, steps)
    "readme": (bb.SCENE_README, 0, 1, 1000),
    "balls_400": (bb.SCENE_BALLS, 400, 101, 60),
    "mixed_300": (bb.SCENE_MIXED, 300, 102, 60),
    "sdf_200": (bb.SCENE_SDF, 200, 103, 80),
    "rlworld": (bb.SCENE_RLWORLD, 0, 104, 60),
}


def run_case(lib, kind, n, seed, steps, on_step=None):
    w = lib.build_scene(kind, n, seed)
    events = []
    if kind == bb.SCENE_README:
        def cb(selfi, otype, oidx, c):
            k = c.contents
            events.append((cur[0], selfi, otype, oidx, k.penetration, k.contact_normal.x, k.contact_normal.y, k.contact_normal.z,
                           k.contact_point.x, k.contact_point.y, k.contact_point.z))
        w.set_sphere_callback(0, cb)
    cur = [0]
    hashes, counts = [], []
    for s in range(steps):
        cur[0] = s
        if kind == bb.SCENE_README and s < 30:
            lib.dll.AddSphereForce(w.ptr, 0, bb.Vec3(2.0, 0.0, 0.0))
        w.step(1.0 / 60.0)
        hashes.append(w.state_hash()); counts.append(len(w.constraints()))
        if on_step:
            on_step(s, w)
    out = {k: v for k, v in w.snapshot().items()}
    out["hashes"] = np.array(hashes, dtype=np.uint64)
    out["contact_counts"] = np.array(counts, dtype=np.int32)
    out["constraints"] = w.constraints().copy().view(np.uint8)
    out["events"] = np.array(events, dtype=np.float64).reshape(-1, 11)
    return out


if __name__ == "__main__":
    ref = oracles.load_all().get("ref")
    if ref is None:
        raise SystemExit("oracle/_ref is not built (needs /root/reference): cannot regenerate golden vectors")
    for name, (kind, n, seed, steps) in CASES.items():
        out = run_case(ref, kind, n, seed, steps)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "steps", steps, "final contacts", out["contact_counts"][-1], "events", len(out["events"]))
