"""Loading of the CPU checkers (TEST INFRASTRUCTURE).

`ref`    = oracle/_ref/libballbox_ref.so, the unmodified reference header compiled
           by oracle/Makefile where /root/reference is present (prebuilt file
           travels to the GPU box; it is never read from /root/reference there).
`oracle` = oracle/libbbx_oracle.so, our plain-C restatement (always buildable).
"""
import os
import subprocess

from ballbox_b200.binding import BallboxLib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ODIR, "_ref", "libballbox_ref.so")
ORACLE_SO = os.path.join(ODIR, "libbbx_oracle.so")


def load_all():
    out = {}
    if os.path.exists(os.path.join(ODIR, "bbx_oracle.c")):
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ODIR, "bbx_oracle.c")):
            subprocess.run(["make", "-s", "-C", ODIR, "all"], check=True)
        out["oracle"] = BallboxLib(ORACLE_SO, product=False)
    if not os.path.exists(REF_SO) and os.path.exists("/root/reference/ballbox.h"):
        subprocess.run(["make", "-s", "-C", ODIR, "ref"], check=True)
    if os.path.exists(REF_SO):
        out["ref"] = BallboxLib(REF_SO, product=False)
    return out
