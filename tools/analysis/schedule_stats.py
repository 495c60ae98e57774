"""Schedule statistics of the exact solver order on a settled scene (analysis tool; numbers quoted in profiles/r02_notes.md).

  python tools/analysis/schedule_stats.py [bodies] [settle_steps] [compare_steps]     (GPU box: product library)
  python tools/analysis/schedule_stats.py /tmp/graph.npz                               (a saved contact list: static part only)

For consecutive settled steps it downloads the constraint list (body ids in emission order) and computes, vectorised:
  * the solver's nodes (runs of contacts of one body pair) and their Kahn levels (what the product replays per iteration);
  * the exact depth of the dependency graph of ALL iterations together (node (n, k+1) waits for the last node of each of
    its bodies in iteration k): the lower bound on dependent levels of ANY bit-exact schedule, and the period P of the
    skewed schedule time(n, k) = level(n) + k * P that ASAP levels allow;
  * how much of a kept level assignment survives from one step to the next: nodes added /
    removed, kept levels that are still locally valid, ASAP levels that changed.
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")


def nodes_of(a, b):
    new = np.ones(len(a), bool)
    new[1:] = (a[1:] != a[:-1]) | (b[1:] != b[:-1])
    return a[new], b[new]


def preds(na, nb, dyn):
    """pa/pb: previous node on body A / B (-1: none); first/last node per body (-1: untouched)."""
    N = len(na)
    idx = np.arange(N, dtype=np.int64)
    da = (na >= 0) & dyn[np.maximum(na, 0)]
    db = (nb >= 0) & dyn[np.maximum(nb, 0)]
    body = np.concatenate([na[da], nb[db]])
    node = np.concatenate([idx[da], idx[db]])
    side = np.concatenate([np.zeros(da.sum(), np.int8), np.ones(db.sum(), np.int8)])
    o = np.lexsort((node, body))
    body, node, side = body[o], node[o], side[o]
    same = np.zeros(len(body), bool)
    same[1:] = body[1:] == body[:-1]
    prev = np.where(same, np.concatenate([[-1], node[:-1]]), -1)
    pa = np.full(N, -1, np.int64); pb = np.full(N, -1, np.int64)
    pa[node[side == 0]] = prev[side == 0]
    pb[node[side == 1]] = prev[side == 1]
    nbody = len(dyn)
    first = np.full(nbody, -1, np.int64); last = np.full(nbody, -1, np.int64)
    first[body[~same]] = node[~same]
    is_last = np.ones(len(body), bool)
    is_last[:-1] = body[1:] != body[:-1]
    last[body[is_last]] = node[is_last]
    return pa, pb, first, last, da, db


def asap_levels(pa, pb):
    lev = np.zeros(len(pa), np.int64)
    while True:
        new = np.maximum(np.where(pa >= 0, lev[np.maximum(pa, 0)] + 1, 0), np.where(pb >= 0, lev[np.maximum(pb, 0)] + 1, 0))
        if np.array_equal(new, lev):
            return lev
        lev = new


def unrolled_depth(na, nb, pa, pb, last, da, db, lev, iters):
    L = int(lev.max()) + 1
    order = np.argsort(lev, kind="stable")
    bounds = np.searchsorted(lev[order], np.arange(L + 1))
    lastA = np.where(da, last[np.maximum(na, 0)], -1)
    lastB = np.where(db, last[np.maximum(nb, 0)], -1)
    prev = lev.copy()
    depth = [L]
    for k in range(1, iters):
        cur = np.zeros(len(na), np.int64)
        for l in range(L):
            i = order[bounds[l]:bounds[l + 1]]
            ta = np.where(pa[i] >= 0, cur[np.maximum(pa[i], 0)], np.where(lastA[i] >= 0, prev[np.maximum(lastA[i], 0)], -1))
            tb = np.where(pb[i] >= 0, cur[np.maximum(pb[i], 0)], np.where(lastB[i] >= 0, prev[np.maximum(lastB[i], 0)], -1))
            cur[i] = np.maximum(ta, tb) + 1
        depth.append(int(cur.max()) + 1)
        prev = cur
    return depth


def static_stats(a, b, dyn, iters=8):
    na, nb = nodes_of(a, b)
    pa, pb, first, last, da, db = preds(na, nb, dyn)
    lev = asap_levels(pa, pb)
    L = int(lev.max()) + 1
    used = first >= 0
    span = lev[last[used]] - lev[first[used]]
    depth = unrolled_depth(na, nb, pa, pb, last, da, db, lev, iters)
    out = dict(contacts=int(len(a)), nodes=int(len(na)), levels=L, dependent_levels_per_step=L * iters,
               exact_unrolled_depth=depth, asap_period=int(span.max()) + 1,
               span_percentiles_50_90_99_999=[float(x) for x in np.percentile(span, [50, 90, 99, 99.9])],
               nodes_per_level_head=[int(x) for x in np.bincount(lev)[:8]])
    return out, (na, nb, pa, pb, lev)


def keys_of(na, nb):
    return (na + 1) * (1 << 32) + (nb + 1)


def persistence(old, new):
    (ona, onb, _, _, olev), (na, nb, pa, pb, lev) = old, new
    ok, nk = keys_of(ona, onb), keys_of(na, nb)
    o = np.argsort(ok)
    oks = ok[o]
    dup = int((oks[1:] == oks[:-1]).sum())
    pos = np.searchsorted(oks, nk)
    pos = np.minimum(pos, len(oks) - 1)
    found = oks[pos] == nk
    kept = np.where(found, olev[o[pos]], -1)                       # last step's level, -1 for a new node
    ka = np.where(pa >= 0, kept[np.maximum(pa, 0)], -2)            # -2: no predecessor on that side
    kb = np.where(pb >= 0, kept[np.maximum(pb, 0)], -2)
    valid = found & (ka != -1) & (kb != -1) & (kept > ka) & (kept > kb)
    removed = int(len(ok) - found.sum()) if dup == 0 else None
    return dict(nodes=int(len(nk)), added=int((~found).sum()), removed=removed, duplicate_keys=dup,
                kept_level_invalid=int((~valid).sum()), kept_level_invalid_frac=float((~valid).mean()),
                asap_level_changed=int((found & (kept != lev)).sum()),
                levels_old=int(olev.max()) + 1, levels_new=int(lev.max()) + 1,
                whole_assignment_valid=bool(valid.all()))


def ids_of(w, bb):
    c = w.constraints()
    ns = w.n_spheres
    a = np.where(c["a_index"] >= 0, c["a_type"].astype(np.int64) * ns + c["a_index"], -1)
    b = np.where(c["b_index"] >= 0, c["b_type"].astype(np.int64) * ns + c["b_index"], -1)
    return a.astype(np.int64), b.astype(np.int64)


def main():
    if len(sys.argv) > 1 and sys.argv[1].endswith(".npz"):
        z = np.load(sys.argv[1])
        print(json.dumps(static_stats(z["a"], z["b"], z["dyn"])[0]))
        return
    from ballbox_b200 import binding as bb
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    settle = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    import os
    cpu = os.environ.get("BBX_ANALYSIS_CPU") == "1"              # small scenes without a GPU: the CPU checker steps
    if cpu:
        from tests import oracles
        lib = oracles.load_all()["oracle"]
    else:
        import ballbox_b200
        lib = ballbox_b200.load()
    w = lib.build_scene(bb.SCENE_MIXED, n, 20261019)

    def advance(k):
        if cpu:
            for _ in range(k):
                w.step(1.0 / 60.0, pruned=True)
        else:
            w.step_n(1.0 / 60.0, k)
            w.download(bb.DL_BODIES | bb.DL_CONSTRAINTS)
    if not cpu:
        w.set_sync_mode(bb.SYNC_RESIDENT)
    advance(settle)
    dyn = None
    res = dict(bodies=n, settle=settle, steps=[])
    old = None
    for s in range(steps):
        advance(1)
        if dyn is None:
            sa, ba = w.sphere_arrays(), w.box_arrays()
            dyn = np.concatenate([sa["static"] == 0, ba["static"] == 0])
        a, b = ids_of(w, bb)
        t = time.time()
        if s == 0:
            st, cur = static_stats(a, b, dyn)
            res["static"] = st
        else:
            na, nb = nodes_of(a, b)
            pa, pb, *_ = preds(na, nb, dyn)
            cur = (na, nb, pa, pb, asap_levels(pa, pb))
            res["steps"].append(persistence(old, cur))
        old = cur
        print("step %d analysed in %.1f s" % (s, time.time() - t), file=sys.stderr)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
