"""tests/golden/make_golden.py -- generates tests/golden/reference_outputs.npz by RUNNING THE REFERENCE ITSELF
(oracle/_ref = /root/reference/include/sdf_tools/sdf_generation.hpp:19-271 compiled unmodified, see
oracle/build_ref.sh) on the fixture geometries of the reference's own tests and on small copies of the BASELINE
configs.  Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The .npz holds, per case, the reference's f32 SDF and (max, min), so the GPU box can check against real reference
output without the reference tree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from sdf_tools_b200 import synth        # noqa: E402

CASES = {
    "single_cell": lambda: synth.fixture_single_cell(),
    "demo_blocks": lambda: synth.fixture_demo_blocks(),
    "tutorial_octant": lambda: synth.fixture_tutorial_octant(),
    "walls_cross": lambda: synth.fixture_walls_cross(),
    "c1_64": lambda: (synth.make_mask("c1"), synth.CONFIGS["c1"]["res"]),
    "c2_48": lambda: (synth.make_mask("c2", (48, 48, 48)), 0.05),
    "c3_64": lambda: (synth.make_mask("c3", (64, 64, 64)), 0.01),
    "c5_64": lambda: (synth.make_mask("c5", (64, 64, 64)), 0.02),
}


def main():
    out = {}
    for name, fn in CASES.items():
        mask, res = fn()
        sdf, mx, mn, ms = O.ref_build_sdf(mask, res)
        out[name + "__sdf"] = sdf
        out[name + "__extrema"] = np.array([mx, mn])
        out[name + "__res"] = np.array([res])
        print(f"{name}: {mask.shape} reference {ms:.0f} ms  max {mx:.6f} min {mn:.6f}")
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_outputs.npz"), **out)


if __name__ == "__main__":
    main()
