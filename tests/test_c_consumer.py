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


def build(tmp_path):
    import ballbox_b200
    ballbox_b200.load()                                   # makes sure libballbox_b200.so exists (builds it if missing)
    exe = str(tmp_path / "readme_quickstart")
    cmd = ["gcc", "-std=gnu11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR,
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
