"""Schedule statistics of config-5 worlds (256 bodies, every constraint its own node as in k_solve_groups), on the CPU
checker: numbers quoted in profiles/r02_notes.md ("Small batches once more", experiment 1).

  python tools/analysis/batch_schedule_stats.py [worlds] [settle_steps]

Per world (reference stepped `settle_steps` steps): Kahn levels L of one iteration, the skew period P that ASAP levels
allow (time(c, k) = level(c) + k * P is valid iff P > level(last constraint of b) - level(first constraint of b) for every
body b), the exact depth of all 8 iterations taken as one graph, and the exact depth of iterations 1..7 alone (what an
unrolled replay after the discovery iteration would run: one "last touch time" per body, walked in level order).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "analysis"))
import schedule_stats as ss                      # noqa: E402
from ballbox_b200 import binding as bb, sharding  # noqa: E402
from tests import oracles                         # noqa: E402


def main():
    n_worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    settle = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    chk = oracles.load_all()
    lib = chk.get("ref") or chk["oracle"]
    rows = []
    t0 = time.time()
    for gid in range(n_worlds):
        w = lib.build_scene(bb.SCENE_RLWORLD, 0, sharding.world_seed(gid))
        for _ in range(settle):
            w.step(1 / 60.)
        c = w.constraints()
        ns = w.n_spheres
        a = np.where(c["a_index"] >= 0, c["a_type"].astype(np.int64) * ns + c["a_index"], -1).astype(np.int64)
        b = np.where(c["b_index"] >= 0, c["b_type"].astype(np.int64) * ns + c["b_index"], -1).astype(np.int64)
        dyn = np.concatenate([w.sphere_arrays()["static"] == 0, w.box_arrays()["static"] == 0])
        pa, pb, first, last, da, db = ss.preds(a, b, dyn)
        keep = da | db
        lev = ss.asap_levels(pa, pb)
        L = int(lev[keep].max()) + 1 if keep.any() else 0
        used = first >= 0
        P = int((lev[last[used]] - lev[first[used]]).max()) + 1 if used.any() else 1
        depth8 = ss.unrolled_depth(a, b, pa, pb, last, da, db, lev, 8)[-1]
        touch = np.zeros(len(dyn), np.int64)
        tmax = 0
        for _ in range(7):
            for p in np.argsort(lev, kind="stable"):
                if not keep[p]:
                    continue
                t = max(touch[a[p]] if da[p] else 0, touch[b[p]] if db[p] else 0) + 1
                if da[p]:
                    touch[a[p]] = t
                if db[p]:
                    touch[b[p]] = t
                tmax = max(tmax, t)
        rows.append((gid, int(keep.sum()), L, P, depth8, tmax))
        w.destroy()
    r = np.array(rows)
    print(f"{n_worlds} worlds, {settle} settle steps, {time.time() - t0:.0f} s")
    print(f"constraints per world: mean {r[:, 1].mean():.0f} max {r[:, 1].max()}")
    print(f"levels per iteration: mean {r[:, 2].mean():.1f} max {r[:, 2].max()};  ASAP skew period: mean {r[:, 3].mean():.1f} max {r[:, 3].max()}")
    print(f"8 iterations: level replay 8 L mean {8 * r[:, 2].mean():.0f} max {8 * r[:, 2].max()};  exact depth mean {r[:, 4].mean():.0f} max {r[:, 4].max()}")
    print(f"iterations 1..7: 7 L mean {7 * r[:, 2].mean():.0f} max {7 * r[:, 2].max()};  unrolled mean {r[:, 5].mean():.0f} max {r[:, 5].max()}")
    print("deepest worlds (id, constraints, L, P, depth of 8, unrolled 1..7):")
    for row in r[np.argsort(-r[:, 2])][:8]:
        print("  ", row.tolist())


if __name__ == "__main__":
    main()
