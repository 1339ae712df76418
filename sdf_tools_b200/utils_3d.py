"""sdf_tools_b200/utils_3d.py -- 3-D convenience wrappers with the call signatures of the reference's
``sdf_tools.utils_3d`` (src/sdf_tools/utils_3d.py:5-108), implemented on bulk arrays."""
from __future__ import annotations

import numpy as np

from . import pysdf_tools
from .utils_2d import _gradient


def _grid_from_env(env, res, origin_point):
    env = np.asarray(env)
    origin = pysdf_tools.Isometry3d([[1.0, 0.0, 0.0, float(origin_point[0])], [0.0, 1.0, 0.0, float(origin_point[1])],
                                     [0.0, 0.0, 1.0, float(origin_point[2])], [0.0, 0.0, 0.0, 1.0]])
    oob = pysdf_tools.COLLISION_CELL(-10000)
    y_shape, x_shape, z_shape = env.shape                                  # "Yes, it goes y,x,z" (utils_3d.py:22)
    grid = pysdf_tools.CollisionMapGrid(origin, "world", res, x_shape, y_shape, z_shape, oob)
    grid.SetOccupancy(np.ascontiguousarray(np.transpose(env == 1, (1, 0, 2)), dtype=np.float32))
    return grid, oob


def compute_sdf(env, res, origin_point):
    """env [y, x, z] of 0/1 -> pysdf_tools.SignedDistanceField."""
    grid, oob = _grid_from_env(env, res, origin_point)
    return grid.ExtractSignedDistanceField(oob.occupancy, False, False)[0]


def compute_sdf_and_gradient(env, res, origin_point):
    """env [y, x, z] of 0/1 -> (sdf [y, x, z] float32, gradient [y, x, z, 3] float32 with components d/dx, d/dy, d/dz)."""
    sdf = compute_sdf(env, res, origin_point)
    xyz = sdf.numpy()
    grad = _gradient(sdf)
    return np.transpose(xyz, (1, 0, 2)).astype(np.float32), np.transpose(grad, (1, 0, 2, 3)).astype(np.float32)


def get_gradient(sdf, dtype=np.float64):
    """Full gradient field (nx, ny, nz, 3) of a pysdf_tools.SignedDistanceField, edge gradients enabled."""
    return _gradient(sdf).astype(dtype)
