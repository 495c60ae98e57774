"""Scenes with joints and springs (extension, include/ballbox_ext.h), built through the public calls so that the same
function sets up a product world and a checker world (oracle/bbx_oracle.c holds the sequential definition)."""
import numpy as np

S, B, W = 0, 1, -1          # body types of the joint calls: sphere, box, the world


def mobile(w, seed=0, n_pendulums=6, chain_links=5, with_floor=True):
    """Pendulums (sphere on a rod, box on a ball joint), a hinged chain hanging from a static box, a sprung sphere, a
    net of spheres tied by rods, all above a static floor box that some of them reach."""
    rng = np.random.default_rng(seed)
    w.set_gravity((0.0, -9.81, 0.0))
    if with_floor:
        f = w.add_box((0.0, -0.5, 0.0), (30.0, 0.5, 30.0))
        w.lib.dll.SetBoxStatic(w.ptr, f, True)
    for k in range(n_pendulums):
        x = -10.0 + 2.5 * k
        L = 1.0 + 0.5 * rng.random()
        if k % 2 == 0:
            i = w.add_sphere((x + L, 4.0, 0.3 * k), 0.25)
            w.add_distance_joint((S, i), (W, 0), (x + L, 4.0, 0.3 * k), (x, 4.0, 0.3 * k))
        else:
            i = w.add_box((x + L, 4.0, 0.0), (0.3, 0.2, 0.25), w.quat_axis_angle((0.3, 1.0, 0.2), 0.4 * k))
            w.add_ball_joint((B, i), (W, 0), (x, 4.0, 0.0))
    # hinged chain from a static box; long enough to drape over the floor
    a = w.add_box((0.0, 3.0, 5.0), (0.2, 0.2, 0.2))
    w.lib.dll.SetBoxStatic(w.ptr, a, True)
    prev = a
    for k in range(chain_links):
        b = w.add_box((0.6 * (k + 1), 3.0, 5.0), (0.25, 0.08, 0.12))
        w.add_hinge_joint((B, prev), (B, b), (0.6 * k + 0.3, 3.0, 5.0), (0.0, 0.0, 1.0), 0.15)
        prev = b
    tip = w.add_sphere((0.6 * chain_links + 0.55, 3.0, 5.0), 0.3)            # a ball on the end of the chain
    w.add_ball_joint((B, prev), (S, tip), (0.6 * chain_links + 0.55, 3.0, 5.0))
    # springs: to the world, and between two dynamic bodies (one of them shared by two springs)
    s0 = w.add_sphere((-4.0, 3.0, -5.0), 0.3)
    s1 = w.add_sphere((-4.0, 1.8, -5.0), 0.25)
    bx = w.add_box((-2.5, 1.8, -5.0), (0.3, 0.3, 0.3))
    w.add_spring((S, s0), (W, 0), (-4.0, 3.0, -5.0), (-4.0, 5.0, -5.0), 1.5, 60.0, 0.8)
    w.add_spring((S, s0), (S, s1), (-4.0, 3.0, -5.0), (-4.0, 1.8, -5.0), 1.0, 40.0, 0.5)
    w.add_spring((B, bx), (S, s1), (-2.3, 2.0, -5.0), (-4.0, 1.8, -5.0), 1.2, 30.0, 0.2)
    # a 4 x 4 net of spheres tied by rods, two corners pinned to the world: every inner sphere carries four joints
    n = 4
    ids = [[w.add_sphere((6.0 + 0.7 * i, 3.0, -6.0 + 0.7 * j), 0.2) for j in range(n)] for i in range(n)]
    P = lambda i, j: (6.0 + 0.7 * i, 3.0, -6.0 + 0.7 * j)
    for i in range(n):
        for j in range(n):
            if i + 1 < n: w.add_distance_joint((S, ids[i][j]), (S, ids[i + 1][j]), P(i, j), P(i + 1, j))
            if j + 1 < n: w.add_distance_joint((S, ids[i][j]), (S, ids[i][j + 1]), P(i, j), P(i, j + 1))
    w.add_ball_joint((S, ids[0][0]), (W, 0), P(0, 0))
    w.add_ball_joint((S, ids[n - 1][0]), (W, 0), P(n - 1, 0))
    return w


def chain_on_pile(w, seed=0, n_loose=60, links=8):
    """A free chain of boxes (ball joints) and a sprung pair dropped on a heap of loose spheres and boxes inside walls:
    contacts and joints act on the same bodies."""
    rng = np.random.default_rng(seed)
    w.set_gravity((0.0, -9.81, 0.0))
    for (p, h) in (((0, -0.5, 0), (6, 0.5, 6)), ((-6.5, 3, 0), (0.5, 4, 6)), ((6.5, 3, 0), (0.5, 4, 6)), ((0, 3, -6.5), (6, 4, 0.5)), ((0, 3, 6.5), (6, 4, 0.5))):
        f = w.add_box(tuple(float(x) for x in p), tuple(float(x) for x in h))
        w.lib.dll.SetBoxStatic(w.ptr, f, True)
    for k in range(n_loose):
        p = (float(rng.uniform(-4, 4)), float(rng.uniform(0.4, 3.0)), float(rng.uniform(-4, 4)))
        if k % 2: w.add_sphere(p, float(rng.uniform(0.2, 0.4)))
        else: w.add_box(p, tuple(float(x) for x in rng.uniform(0.15, 0.4, 3)), w.quat_axis_angle((1.0, 0.5, 0.2), float(rng.uniform(0, 3))))
    prev = None
    for k in range(links):
        b = w.add_box((-2.0 + 0.55 * k, 4.5, 0.0), (0.25, 0.1, 0.15))
        if prev is not None:
            w.add_ball_joint((B, prev), (B, b), (-2.0 + 0.55 * k - 0.275, 4.5, 0.0))
        prev = b
    s0 = w.add_sphere((0.0, 5.5, 2.0), 0.3); s1 = w.add_sphere((1.2, 5.5, 2.0), 0.3)
    w.add_spring((S, s0), (S, s1), (0.0, 5.5, 2.0), (1.2, 5.5, 2.0), 1.0, 80.0, 1.0)
    w.add_distance_joint((S, s1), (B, prev), (1.2, 5.5, 2.0), (-2.0 + 0.55 * (links - 1), 4.5, 0.0))
    return w


def capacity(scene, **kw):
    """(max_spheres, max_boxes, max_collisions) large enough for the scene"""
    return 256, 256, 8192
