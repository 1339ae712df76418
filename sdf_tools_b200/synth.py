"""sdf_tools_b200/synth.py -- the synthetic occupancy grids of BASELINE.json / SURVEY.md section 8d.

All generators are counter-based (splitmix64 of (seed, index)), so a grid is a pure function of its name and
shape and every rank / language reproduces it bit-for-bit.  Masks are uint8 (1 = filled) in the reference's
memory order (nx, ny, nz), z contiguous.

  c1  64^3           iid Bernoulli p = 0.30, seed 1                       (CPU-reference-runnable plumbing case)
  c2  256^3          solid sphere, centre (128,128,128), r = 0.3*256, on cell centres
  c3  1024^3         "point-cloud voxelisation": 2e6 points on the faces of 64 random boxes + ground plane z=8,
                     voxelised with floor(p / res) as scripts/3d_sdf_demo_rviz.py:22-29 does  (~0.15 % fill)
  c4  2048x2048x1024 same generator, 8e6 points, seed 4
  c5  512^3          octree-style occupancy, leaves of edge 4..32, 30 % fill, seed 5; 1 % of cells "unknown" (0.5)
plus the small fixture geometries of the reference's tests.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser of (seed + (idx+1)*golden) -> uint64, vectorised, wrap-around arithmetic."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, idx: np.ndarray) -> np.ndarray:
    """float64 in [0,1) from the top 53 bits."""
    return (splitmix64(seed, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def iid(shape, p: float, seed: int) -> np.ndarray:
    n = int(np.prod(shape))
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 24
    for s in range(0, n, step):
        idx = np.arange(s, min(n, s + step), dtype=np.uint64)
        out[s:s + len(idx)] = uniform01(seed, idx) < p
    return out.reshape(shape)


def sphere(shape, radius_frac: float = 0.3) -> np.ndarray:
    nx, ny, nz = shape
    r = radius_frac * nx
    cx, cy, cz = nx / 2.0, ny / 2.0, nz / 2.0
    x = (np.arange(nx) + 0.5 - cx)[:, None, None]
    y = (np.arange(ny) + 0.5 - cy)[None, :, None]
    z = (np.arange(nz) + 0.5 - cz)[None, None, :]
    return ((x * x + y * y + z * z) <= r * r).astype(np.uint8)


def pointcloud(shape, n_points: int, seed: int, n_boxes: int = 64, ground_z: int = 8, zrange=None) -> np.ndarray:
    """Voxelised synthetic point cloud: points uniform (by area) on box faces and a ground plane.
    zrange = (z0, z1): return only that z slab, shape (nx, ny, z1 - z0), without building the whole grid."""
    nx, ny, nz = shape
    z0, z1 = (0, nz) if zrange is None else zrange
    dims = np.array([nx, ny, nz], dtype=np.float64)
    scale = min(nx, ny, nz) / 1024.0
    b = np.arange(n_boxes, dtype=np.uint64)
    edge = np.stack([32 + uniform01(seed + 101 + k, b) * (256 - 32) for k in range(3)], axis=1) * scale
    edge = np.minimum(edge, dims - 2)
    org = np.stack([uniform01(seed + 201 + k, b) for k in range(3)], axis=1) * (dims - edge - 1)
    # face areas: 2 faces per axis pair
    fa = np.stack([edge[:, 1] * edge[:, 2], edge[:, 0] * edge[:, 2], edge[:, 0] * edge[:, 1]], axis=1)  # normal x,y,z
    areas = np.concatenate([np.repeat(fa.reshape(-1), 2), [float(nx) * ny]])      # 6 per box + ground
    cdf = np.cumsum(areas) / areas.sum()
    mask = np.zeros((nx, ny, z1 - z0), dtype=np.uint8)
    step = 1 << 22
    for s in range(0, n_points, step):
        idx = np.arange(s, min(n_points, s + step), dtype=np.uint64)
        surf = np.searchsorted(cdf, uniform01(seed, idx), side="right").clip(0, len(areas) - 1)
        u = uniform01(seed + 1, idx)
        v = uniform01(seed + 2, idx)
        pts = np.empty((len(idx), 3), dtype=np.float64)
        ground = surf == len(areas) - 1
        bi = np.minimum(surf // 6, n_boxes - 1)
        axis = (surf % 6) // 2
        side = (surf % 6) % 2
        for ax in range(3):
            o1, o2 = [a for a in range(3) if a != ax]
            sel = (~ground) & (axis == ax)
            bsel = bi[sel]
            pts[sel, ax] = org[bsel, ax] + side[sel] * edge[bsel, ax]
            pts[sel, o1] = org[bsel, o1] + u[sel] * edge[bsel, o1]
            pts[sel, o2] = org[bsel, o2] + v[sel] * edge[bsel, o2]
        pts[ground, 0] = u[ground] * nx
        pts[ground, 1] = v[ground] * ny
        pts[ground, 2] = min(ground_z, nz - 1) + 0.5
        ijk = np.floor(pts).astype(np.int64)
        np.clip(ijk, 0, np.array([nx - 1, ny - 1, nz - 1]), out=ijk)
        keep = (ijk[:, 2] >= z0) & (ijk[:, 2] < z1)
        ijk = ijk[keep]
        mask[ijk[:, 0], ijk[:, 1], ijk[:, 2] - z0] = 1
    return mask


def octree(shape, fill: float = 0.30, seed: int = 5, min_edge: int = 4, max_edge: int = 32):
    """Octree-style occupancy: recursively split to leaves of edge min_edge..max_edge; a leaf is filled with
    probability `fill`.  Returns float32 occupancy in {0, 1} with 1 % of cells set to 0.5 ("unknown")."""
    nx, ny, nz = shape
    occ = np.zeros(shape, dtype=np.float32)
    top = max_edge
    while top * 2 <= min(shape):
        top *= 2
    # start from a coarse lattice of `max_edge` cells; each cell either becomes a leaf or splits (p=0.5) down to min_edge
    def rec(x0, y0, z0, e, key):
        u = float(uniform01(seed, np.array([key], dtype=np.uint64))[0])
        if e > min_edge and (e > max_edge or u < 0.55):
            h = e // 2
            k = 0
            for dx in (0, h):
                for dy in (0, h):
                    for dz in (0, h):
                        rec(x0 + dx, y0 + dy, z0 + dz, h, key * 8 + k + 1)
                        k += 1
        else:
            f = float(uniform01(seed + 7, np.array([key], dtype=np.uint64))[0])
            if f < fill:
                occ[x0:x0 + e, y0:y0 + e, z0:z0 + e] = 1.0
    e0 = max_edge
    key = 1
    for x0 in range(0, nx, e0):
        for y0 in range(0, ny, e0):
            for z0 in range(0, nz, e0):
                rec(x0, y0, z0, e0, key)
                key += 1
    n = occ.size
    flat = occ.reshape(-1)
    step = 1 << 24
    for s in range(0, n, step):
        idx = np.arange(s, min(n, s + step), dtype=np.uint64)
        unk = uniform01(seed + 13, idx) < 0.01
        flat[s:s + len(idx)][unk] = 0.5
    return occ


# ------------------------------------------------------------------------------------------------- registry
CONFIGS = {
    "c1": dict(shape=(64, 64, 64), res=0.25, desc="64^3 iid Bernoulli p=0.30 seed 1"),
    "c2": dict(shape=(256, 256, 256), res=0.05, desc="256^3 solid sphere r=0.3N"),
    "c3": dict(shape=(1024, 1024, 1024), res=0.01, desc="1024^3 point-cloud voxelisation, 2e6 points, seed 3"),
    "c4": dict(shape=(2048, 2048, 1024), res=0.01, desc="2048x2048x1024 point-cloud voxelisation, 8e6 points, seed 4"),
    "c5": dict(shape=(512, 512, 512), res=0.02, desc="512^3 octree-style occupancy, 30% fill, seed 5"),
}


def make_mask(name: str, shape=None, zrange=None) -> np.ndarray:
    """uint8 mask of a named config (optionally at another shape, e.g. scaled-down copies for tests).
    zrange = (z0, z1) returns only that z slab (nx, ny, z1 - z0) -- what one rank of a slab build holds."""
    cfg = CONFIGS[name]
    shape = tuple(shape or cfg["shape"])
    if name == "c3":
        return pointcloud(shape, int(2e6 * (np.prod(shape) / 1024.0 ** 3) ** (2.0 / 3.0)), 3, zrange=zrange)
    if name == "c4":
        return pointcloud(shape, int(8e6 * (np.prod(shape) / (2048.0 * 2048 * 1024)) ** (2.0 / 3.0)), 4, zrange=zrange)
    if name == "c1":
        full = iid(shape, 0.30, 1)
    elif name == "c2":
        full = sphere(shape, 0.3)
    elif name == "c5":
        full = (octree(shape, 0.30, 5) > 0.5).astype(np.uint8)
    else:
        raise KeyError(name)
    if zrange is not None:
        full = np.ascontiguousarray(full[:, :, zrange[0]:zrange[1]])
    return full


# --------------------------------------------------------------------- fixture geometries of the reference's tests
def fixture_single_cell():
    """test/test_bindings.py:12-33: 20 x 40 x 1 grid (x_width=20, y_height=40), one filled cell at x=3, y=1."""
    m = np.zeros((20, 40, 1), dtype=np.uint8)
    m[3, 1, 0] = 1
    return m, 0.05


def fixture_demo_blocks():
    """test/demo_bindings.py:9-16: 100 x 100 x 1, three 10x10 blocks (grid_world[y, x])."""
    gw = np.zeros((100, 100), dtype=np.uint8)
    gw[0:10, 0:10] = 1
    gw[90:100, 90:100] = 1
    gw[20:30, 40:50] = 1
    return np.ascontiguousarray(gw.T)[:, :, None], 0.01


def fixture_tutorial_octant():
    """src/sdf_tools_tutorial.cpp:42-62: 40^3 map, one octant filled."""
    m = np.zeros((40, 40, 40), dtype=np.uint8)
    m[0:20, 0:20, 0:20] = 1
    return m, 0.25


def fixture_walls_cross():
    """src/compute_convex_segments_test.cpp:13-41: 100 x 100 x 50 map, outer walls + a cross."""
    m = np.zeros((100, 100, 50), dtype=np.uint8)
    m[0, :, :] = 1; m[-1, :, :] = 1; m[:, 0, :] = 1; m[:, -1, :] = 1
    m[45:55, 10:90, :] = 1
    m[10:90, 45:55, :] = 1
    return m, 1.0
