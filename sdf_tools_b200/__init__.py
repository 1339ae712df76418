"""sdf_tools_b200 -- B200 (sm_100a) build of sdf_tools' signed-distance-field hot path.

The product is libsdf_b200.so (hand-written CUDA behind the C ABI of include/sdf_b200.h).  This package holds its
sources (csrc/), the in-tree build (build.py), the ctypes binding (_capi.py), a small host-side API (api.py)
and the synthetic workloads of SURVEY.md section 8d (synth.py).
"""
from .api import Context, SdfB200Error, build_sdf  # noqa: F401
