"""The drop-in claim, exercised by gcc: a plain C program written against the reference's README
quick start (reference README.md:37-87; callback at file scope, #define BALLBOX_IMPLEMENTATION kept)
compiles against include/ and links with -lballbox_b200 with no other change.
On the GPU box it runs, and its callbacks / state hashes / contact counts must equal
tests/golden/readme.npz, which was generated from the unmodified reference header."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "readme_quickstart.c")
LIBDIR = os.path.join(ROOT, "ballbox_b200")


def build(tmp_path, src=SRC):
    import ballbox_b200
    ballbox_b200.load()                                   # makes sure libballbox_b200.so exists (builds it if missing)
    exe = str(tmp_path / os.path.splitext(os.path.basename(src))[0])
    cmd = ["gcc", "-std=gnu11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", LIBDIR,
           "-lballbox_b200", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def test_readme_quickstart_compiles_and_links_against_the_drop_in(tmp_path):
    exe = build(tmp_path)
    nm = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    for sym in ("CreatePhysicsWorld", "AddSphere", "AddBox", "SetBoxStatic", "SetSphereCollisionCallback", "AddSphereForce",
                "PhysicsStep", "DestroyPhysicsWorld", "QuatIdentity"):
        assert sym in nm, f"{sym} is not taken from the library"


def f32(hexbits):
    return struct.unpack("<f", struct.pack("<I", int(hexbits, 16)))[0]


@pytest.mark.gpu
def test_readme_quickstart_c_program_matches_the_reference_golden_trace(tmp_path):
    exe = build(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    events, hashes, counts = [], [], []
    for line in p.stdout.splitlines():
        f = line.split()
        if f[0] == "E":
            events.append([int(f[1]), int(f[2]), int(f[3]), int(f[4])] + [f32(x) for x in f[5:12]])
        elif f[0] == "H":
            hashes.append(int(f[2], 16)); counts.append(int(f[3]))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "readme.npz"))
    ev = np.array(events, dtype=np.float64).reshape(-1, 11)
    assert ev.shape == gold["events"].shape == (933, 11)
    assert ev[0, 0] == 67                                  # first contact of the falling sphere
    assert np.array_equal(ev, gold["events"])
    assert np.array_equal(np.array(hashes, dtype=np.uint64), gold["hashes"])
    assert np.array_equal(np.array(counts, dtype=np.int32), gold["contact_counts"])


JOINTS_SRC = os.path.join(ROOT, "tests", "c", "joints_bridge.c")
PLANKS = 7


def bridge(w):
    """tests/c/joints_bridge.c, call for call, through the Python binding (for the sequential definition in oracle/)"""
    f = lambda *v: tuple(float(np.float32(x)) for x in v)
    w.set_gravity(f(0, -9.8, 0))
    left = w.add_box(f(-0.5, 3, 0), f(0.5, 0.5, 1)); right = w.add_box(f(np.float32(0.5) + np.float32(0.8) * PLANKS + np.float32(0.5), 3, 0), f(0.5, 0.5, 1))
    w.lib.dll.SetBoxStatic(w.ptr, left, True); w.lib.dll.SetBoxStatic(w.ptr, right, True)
    prev, mid = left, -1
    for k in range(PLANKS):
        plank = w.add_box(f(np.float32(0.4) + np.float32(0.8) * k, 3.3, 0), f(0.38, 0.06, 0.8))
        w.add_hinge_joint((1, prev), (1, plank), f(np.float32(0.8) * k, 3.3, 0), f(0, 0, 1), float(np.float32(0.6)))
        if k == PLANKS // 2:
            mid = plank
        prev = plank
    w.add_hinge_joint((1, prev), (1, right), f(np.float32(0.8) * PLANKS, 3.3, 0), f(0, 0, 1), float(np.float32(0.6)))
    x = np.float32(0.4) + np.float32(0.8) * (PLANKS // 2)
    bob = w.add_sphere(f(x, 2.0, 0), float(np.float32(0.25)))
    w.add_distance_joint((0, bob), (1, mid), f(x, 2.0, 0), f(x, 3.24, 0))
    sprung = w.add_sphere(f(-3, 3, 0), float(np.float32(0.3)))
    w.add_spring((0, sprung), (-1, 0), f(-3, 3, 0), f(-3, 5, 0), 1.5, 40.0, 0.5)
    w.add_sphere(f(2.1, 5.0, 0.1), float(np.float32(0.35)))
    return w


def test_joints_c_program_compiles_against_the_extension_header(tmp_path):
    exe = build(tmp_path, JOINTS_SRC)
    nm = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    for sym in ("BallboxAddHingeJoint", "BallboxAddDistanceJoint", "BallboxAddSpring", "BallboxJointCount", "PhysicsStep"):
        assert sym in nm, f"{sym} is not taken from the library"


@pytest.mark.gpu
def test_joints_c_program_matches_the_sequential_definition(tmp_path, checkers):
    exe = build(tmp_path, JOINTS_SRC)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.returncode, p.stderr[-2000:])
    got = [(int(f[2], 16), int(f[3])) for f in (line.split() for line in p.stdout.splitlines()) if f and f[0] == "H"]
    oracle = checkers["oracle"]
    w = bridge(oracle.create_world(16, 16, 512))
    want = []
    for _ in range(240):
        w.step(1.0 / 60.0)
        want.append((w.state_hash(), len(w.constraints())))
    assert len(got) == 240
    assert got == want
    assert max(c for _, c in got) > 0                      # the loose sphere did land on the bridge
