"""Shared code of the golden-vector tests: replays a case of tests/golden/make_golden.py
on any backend and compares with the committed .npz bit for bit."""
import importlib.util
import os

import numpy as np

from tests import parity

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.CASES


def check_case(lib, name):
    kind, n, seed, steps = CASES[name]
    gold = np.load(os.path.join(HERE, name + ".npz"))
    got = make_golden.run_case(lib, kind, n, seed, steps)
    bad = np.nonzero(got["hashes"] != gold["hashes"])[0]
    assert len(bad) == 0, f"{name}: state hash differs first at step {bad[0]}"
    assert np.array_equal(got["contact_counts"], gold["contact_counts"]), f"{name}: contact counts differ"
    for f in parity.BODY_FIELDS:
        m = parity.first_mismatch(got[f], gold[f])
        assert m is None, f"{name}: final {f}: {m}"
    assert np.array_equal(got["constraints"], gold["constraints"]), f"{name}: final constraint array differs"
    assert got["events"].shape == gold["events"].shape, f"{name}: {len(got['events'])} callbacks vs {len(gold['events'])}"
    assert np.array_equal(got["events"], gold["events"]), f"{name}: callback sequence differs"
