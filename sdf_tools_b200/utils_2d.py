"""sdf_tools_b200/utils_2d.py -- 2-D convenience wrapper with the call signature of the reference's
``sdf_tools.utils_2d.compute_sdf_and_gradient`` (src/sdf_tools/utils_2d.py:6-55), implemented on bulk arrays:
one SetOccupancy call instead of a Python SetValue loop, one memcpy instead of a list copy, one GPU gradient pass
instead of a Python callback per voxel."""
from __future__ import annotations

import numpy as np

from . import pysdf_tools


def _gradient(sdf_grid) -> np.ndarray:
    """(nx, ny, nz, 3) float64 gradient field of a pysdf_tools.SignedDistanceField with the reference's edge rule
    (SignedDistanceField.GetGradient(.., True), sdf.hpp:432-526): central differences inside, one-sided at the
    faces, 0 along axes of length 1 -- computed by the GPU kernel behind GetFullGradientArray, bit-identical to the
    reference's per-voxel GetFullGradient loop (utils_2d.py:44-47)."""
    return sdf_grid.GetFullGradientArray(True)


def compute_sdf_and_gradient(grid_world, sdf_resolution, sdf_origin, frame="world"):
    """grid_world: 2-D array indexed [y, x] of 0/1; returns (sdf [y, x] float32, gradient [y, x, 2] float32)."""
    grid_world = np.asarray(grid_world)
    y_height, x_width = grid_world.shape
    origin = pysdf_tools.Isometry3d([[1.0, 0.0, 0.0, float(sdf_origin[0])], [0.0, 1.0, 0.0, float(sdf_origin[1])],
                                     [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    oob = pysdf_tools.COLLISION_CELL(-10000)
    grid = pysdf_tools.CollisionMapGrid(origin, frame, sdf_resolution, x_width, y_height, 1, oob)
    grid.SetOccupancy(np.ascontiguousarray((grid_world == 1).T[:, :, None], dtype=np.float32))
    sdf_grid, _extrema = grid.ExtractSignedDistanceField(oob.occupancy, False, False)
    sdf_xy = sdf_grid.numpy()[:, :, 0]                                   # (x, y)
    grad_xy = _gradient(sdf_grid)[:, :, 0, 0:2]                          # (x, y, 2): d/dx, d/dy ("remove the z gradient")
    return sdf_xy.T.astype(np.float32), np.transpose(grad_xy, (1, 0, 2)).astype(np.float32)
