"""Shared loader for tests/golden/reference_outputs.npz (real reference output, see tests/golden/make_golden.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES  # noqa: E402

_NPZ = None


def load(name):
    """-> (mask uint8, res, reference sdf f32, (ref max, ref min))"""
    global _NPZ
    if _NPZ is None:
        _NPZ = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))
    mask, res = CASES[name]()
    ext = _NPZ[name + "__extrema"]
    assert float(_NPZ[name + "__res"][0]) == float(res)
    return mask, res, _NPZ[name + "__sdf"], (float(ext[0]), float(ext[1]))


NAMES = list(CASES.keys())


def check_against_reference(name, got_sdf, got_extrema):
    """Parity protocol P0/P2/P3/P4 of SURVEY.md 8c against REAL reference output.
    Returns (n_differ, max_rel_dev) of the voxels where the reference's propagation over-estimated."""
    mask, res, ref, (rmx, rmn) = load(name)
    assert got_sdf.shape == ref.shape
    assert np.array_equal(np.signbit(got_sdf), mask != 0), "P0: sign != occupancy"
    assert np.array_equal(np.signbit(ref), mask != 0)
    differ = got_sdf != ref
    agree = ~differ
    # P2: where the reference found the true nearest voxel, results are identical (hence within 1e-5 relative)
    assert np.allclose(got_sdf[agree], ref[agree], rtol=1e-5, atol=0.0)
    # P3: elsewhere the reference over-estimates, never under-estimates
    assert (np.abs(ref[differ]) > np.abs(got_sdf[differ])).all(), "P3: reference below the exact distance?"
    # P4: extrema are the extrema of the field; the reference's can only be farther out
    gmx, gmn = got_extrema
    assert gmx == float(got_sdf.astype(np.float64).max()) or np.isclose(gmx, got_sdf.max(), rtol=1e-6)
    assert rmx >= gmx - 1e-12 and rmn <= gmn + 1e-12
    rel = np.abs(ref[differ] - got_sdf[differ]) / np.abs(ref[differ]) if differ.any() else np.zeros(1)
    return int(differ.sum()), float(rel.max())
