"""How deep is the dependency graph of the 8 Gauss-Seidel iterations taken TOGETHER?

Builds a config-3-like scene with the CPU checker (analysis tool, not a product path), takes the constraint list of a settled
step, forms the solver's nodes (runs of contacts of one body pair) and computes
  * the Kahn levels of one iteration (what the product replays 8 times),
  * the exact depth of the graph unrolled over all iterations (node (n, k+1) waits for the last node of each of its bodies in
    iteration k when it is the first of that body),
  * the period P of a skewed schedule time(n, k) = lambda(n) + k * P for lambda = ASAP and ALAP levels.
"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ballbox_b200 import binding as bb


def main():
    if len(sys.argv) > 1 and sys.argv[1].endswith(".npz"):
        z = np.load(sys.argv[1])
        return analyse(z["a"], z["b"], 8, None, z["dyn"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    settle = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    iters = 8
    from tests import oracles
    lib = oracles.load_all()['oracle']
    w = lib.build_scene(bb.SCENE_MIXED, n, 20261019)
    t = time.time()
    for s in range(settle):
        w.step(1.0 / 60.0, pruned=True)
    print("settled %d steps in %.1f s" % (settle, time.time() - t))
    c = w.constraints()
    ns = w.n_spheres
    a = np.where(c["a_index"] >= 0, c["a_type"] * ns + c["a_index"], -1).astype(np.int64)
    b = np.where(c["b_index"] >= 0, c["b_type"] * ns + c["b_index"], -1).astype(np.int64)
    sa, ba = w.sphere_arrays(), w.box_arrays()
    dyn = np.concatenate([sa['static'][:ns] == 0, ba['static'][:w.n_boxes] == 0])
    np.savez('/tmp/graph_%d_%d.npz' % (n, settle), a=a, b=b, dyn=dyn)
    analyse(a, b, iters, w, dyn)


def analyse(a, b, iters, w=None, dyn=None):
    C = len(a)
    # nodes: runs of consecutive contacts with the same (a, b)
    new = np.ones(C, bool)
    new[1:] = (a[1:] != a[:-1]) | (b[1:] != b[:-1])
    na, nb = a[new], b[new]
    N = len(na)
    print("contacts %d nodes %d" % (C, N))
    nbody = int(max(na.max(), nb.max())) + 1
    if dyn is None:
        dyn = np.ones(nbody, bool)
    last = np.full(nbody, -1, np.int64)     # last node seen on the body
    pa = np.full(N, -1, np.int64); pb = np.full(N, -1, np.int64)
    first = np.full(nbody, -1, np.int64)
    for i in range(N):
        x, y = na[i], nb[i]
        if x >= 0 and dyn[x]:
            pa[i] = last[x]; last[x] = i
            if first[x] < 0: first[x] = i
        if y >= 0 and dyn[y]:
            pb[i] = last[y]; last[y] = i
            if first[y] < 0: first[y] = i
    lev = np.zeros(N, np.int64)
    for i in range(N):
        l = 0
        if pa[i] >= 0: l = max(l, lev[pa[i]] + 1)
        if pb[i] >= 0: l = max(l, lev[pb[i]] + 1)
        lev[i] = l
    L = int(lev.max()) + 1
    print("levels of one iteration: %d  -> %d dependent levels per step" % (L, L * iters))
    # exact unrolled depth
    prev = lev.copy()
    depth = [int(prev.max()) + 1]
    for k in range(1, iters):
        cur = np.zeros(N, np.int64)
        for i in range(N):
            l = 0
            x, y = na[i], nb[i]
            if x >= 0 and dyn[x]:
                l = max(l, (cur[pa[i]] if pa[i] >= 0 else prev[last[x]]) + 1)
            if y >= 0 and dyn[y]:
                l = max(l, (cur[pb[i]] if pb[i] >= 0 else prev[last[y]]) + 1)
            cur[i] = l
        depth.append(int(cur.max()) + 1)
        prev = cur
    print("exact unrolled depth after iteration k:", depth)
    used = first >= 0
    span = lev[last[used]] - lev[first[used]]
    print("ASAP: P = %d (span percentiles 50/90/99/99.9: %s)" % (span.max() + 1, np.percentile(span, [50, 90, 99, 99.9])))
    # ALAP levels with the same depth
    sa = np.full(N, -1, np.int64); sb = np.full(N, -1, np.int64)
    for i in range(N):
        if pa[i] >= 0: sa[pa[i]] = i if na[pa[i]] == na[i] else sa[pa[i]]
    # successors: simplest is a reverse sweep
    alap = np.full(N, L - 1, np.int64)
    for i in range(N - 1, -1, -1):
        if pa[i] >= 0: alap[pa[i]] = min(alap[pa[i]], alap[i] - 1)
        if pb[i] >= 0: alap[pb[i]] = min(alap[pb[i]], alap[i] - 1)
    span2 = alap[last[used]] - alap[first[used]]
    print("ALAP: P = %d (span percentiles: %s)" % (span2.max() + 1, np.percentile(span2, [50, 90, 99, 99.9])))
    # mixed: first nodes ALAP, everything else as early as possible after that?  lower bound on P: nodes per body
    cnt = np.zeros(nbody, np.int64)
    np.add.at(cnt, na[(na >= 0)], 1); np.add.at(cnt, nb[(nb >= 0)], 1)
    print("max nodes on one dynamic body: %d" % cnt[dyn[:len(cnt)]].max())
    return lev


if __name__ == "__main__":
    main()
