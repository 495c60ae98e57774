"""Bit-level comparison helpers shared by the parity tests."""
import numpy as np

BODY_FIELDS = ("s_pos", "s_vel", "s_angvel", "s_sleeping", "s_timer", "s_force", "s_torque",
               "b_pos", "b_rot", "b_vel", "b_angvel", "b_sleeping", "b_timer", "b_force", "b_torque")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)


def first_mismatch(a, b):
    """None when bit-identical, else a short description of the first difference."""
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    if a.dtype.names:
        for name in a.dtype.names:
            m = first_mismatch(a[name], b[name])
            if m:
                return f"{name}: {m}"
        return None
    if a.size == 0:
        return None
    ab = bits(a).reshape(a.shape[0], -1)
    bb_ = bits(b).reshape(b.shape[0], -1)
    if np.array_equal(ab, bb_):
        return None
    rows = np.nonzero((ab != bb_).any(axis=1))[0]
    i = int(rows[0])
    return f"{len(rows)} rows differ, first row {i}: {a[i]!r} vs {b[i]!r}"


def compare_states(sa, sb):
    """sa, sb: World.snapshot() dicts.  Returns list of mismatch strings."""
    out = []
    for f in BODY_FIELDS:
        m = first_mismatch(sa[f], sb[f])
        if m:
            out.append(f"{f}: {m}")
    return out


def compare_worlds(wa, wb, constraints=True):
    out = compare_states(wa.snapshot(), wb.snapshot())
    if constraints:
        ca, cb = wa.constraints(), wb.constraints()
        if len(ca) != len(cb):
            out.append(f"constraint count {len(ca)} vs {len(cb)}")
        else:
            m = first_mismatch(ca, cb)
            if m:
                out.append("constraints: " + m)
    return out


def max_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(1e-6, np.maximum(np.abs(a), np.abs(b)))))
