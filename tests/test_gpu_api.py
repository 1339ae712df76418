"""GPU tests of the reference-facing faces: the C++ headers (include/sdf_tools/*.hpp) through the pybind11 module
`pysdf_tools`, and the Python helpers.  They restate the reference's only known-answer test
(/root/reference/test/test_bindings.py:12-33) and drive the API exactly the way the reference's own utilities do
(per-voxel SetValue loop, GetRawData, GetFullGradient with a Python callback: src/sdf_tools/utils_2d.py:26-52)."""
import numpy as np
import pytest

from oracle import oracle as O
from sdf_tools_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from sdf_tools_b200 import pysdf_tools
    return pysdf_tools


def _reference_style_2d(P, grid_world, res, origin):
    """The call sequence of the reference's utils_2d.compute_sdf_and_gradient, statement by statement."""
    y_height, x_width = grid_world.shape
    tf = P.Isometry3d([[1.0, 0.0, 0.0, origin[0]], [0.0, 1.0, 0.0, origin[1]], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    oob_value = P.COLLISION_CELL(-10000)
    occupied_value = P.COLLISION_CELL(1)
    grid = P.CollisionMapGrid(tf, "world", res, x_width, y_height, 1, oob_value)
    for x in range(grid.GetNumXCells()):
        for y in range(grid.GetNumYCells()):
            if grid_world[y, x] == 1:
                grid.SetValue(x, y, 0, occupied_value)
    sdf_result = grid.ExtractSignedDistanceField(oob_value.occupancy, False, False)
    sdf = sdf_result[0]
    np_sdf = np.array(sdf.GetRawData()).reshape(sdf.GetNumXCells(), sdf.GetNumYCells()).T.astype(np.float32)

    def gradient_function(x, y, z, enable_edge_gradients=False):
        return sdf.GetGradient(x, y, z, enable_edge_gradients)

    g = sdf.GetFullGradient(gradient_function, True)
    np_g = np.array(g.GetRawData()).reshape(g.GetNumXCells(), g.GetNumYCells(), 3)[:, :, 0:2]
    return np_sdf, np.transpose(np_g, [1, 0, 2]).astype(np.float32), sdf_result[1]


def test_reference_known_answer(P):
    """test/test_bindings.py:12-33, value for value."""
    res, x_width, y_height = 0.05, 20, 40
    grid_world = np.zeros([y_height, x_width], dtype=np.uint8)
    grid_world[1, 3] = 1
    origin = [0 - x_width / 2, 0 - y_height / 2]
    for sdf, grad in (_reference_style_2d(P, grid_world, res, origin)[:2],
                      __import__("sdf_tools_b200.utils_2d", fromlist=["x"]).compute_sdf_and_gradient(grid_world, res, origin)):
        assert sdf[1, 3] == pytest.approx(-res, abs=1e-7)
        assert sdf[2, 3] == pytest.approx(res, abs=1e-7)
        assert sdf[0, 3] == pytest.approx(res, abs=1e-7)
        assert sdf[1, 2] == pytest.approx(res, abs=1e-7)
        assert sdf[1, 4] == pytest.approx(res, abs=1e-7)
        assert sdf[3, 6] > 3 * res
        assert sdf.shape == (y_height, x_width)
        assert grad.shape == (y_height, x_width, 2)
        np.testing.assert_allclose(grad[1, 4], [1.5, 0], rtol=1e-6, atol=1e-6)


def test_bulk_and_per_voxel_paths_agree(P):
    from sdf_tools_b200 import utils_2d
    rng = np.random.default_rng(4)
    gw = (rng.random((23, 31)) < 0.1).astype(np.uint8)
    a_sdf, a_grad, (mx, mn) = _reference_style_2d(P, gw, 0.02, [-1.0, 2.0])
    b_sdf, b_grad = utils_2d.compute_sdf_and_gradient(gw, 0.02, [-1.0, 2.0])
    assert np.array_equal(a_sdf, b_sdf)
    np.testing.assert_allclose(a_grad, b_grad, rtol=1e-6, atol=1e-7)
    want, wmx, wmn = O.build_sdf(np.ascontiguousarray(gw.T)[:, :, None], 0.02)
    assert np.array_equal(a_sdf, want[:, :, 0].T) and (mx, mn) == (wmx, wmn)


def test_utils_3d_against_oracle(P):
    from sdf_tools_b200 import utils_3d
    env = (synth.iid((14, 11, 9), 0.15, 3) != 0).astype(np.uint8)        # env[y, x, z]
    sdf, grad = utils_3d.compute_sdf_and_gradient(env, 0.1, [0.0, 0.0, 0.0])
    want, _, _ = O.build_sdf(np.ascontiguousarray(np.transpose(env, (1, 0, 2))), 0.1)
    assert np.array_equal(sdf, np.transpose(want, (1, 0, 2)))
    assert grad.shape == env.shape + (3,)
    obj = utils_3d.compute_sdf(env, 0.1, [0.0, 0.0, 0.0])
    assert obj.GetNumXCells() == 11 and obj.GetNumYCells() == 14 and obj.GetNumZCells() == 9
    assert obj.GetValueByIndex(0, 0, 0)[1] and not obj.GetValueByIndex(99, 0, 0)[1]


def test_unknown_is_filled_and_virtual_border(P):
    """collision_map.hpp:686-709 (unknown cells) and sdf_generation.hpp:287-419 (virtual border)."""
    occ = synth.octree((24, 20, 16), 0.3, 5)
    tf = P.Isometry3d([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    grid = P.CollisionMapGrid(tf, "w", 0.5, 24, 20, 16, P.COLLISION_CELL(0.0))
    grid.SetOccupancy(occ)
    for unknown in (False, True):
        mask = ((occ > 0.5) | (unknown & (occ == 0.5))).astype(np.uint8)
        sdf, (mx, mn) = grid.ExtractSignedDistanceField(7.0, unknown, False)
        want, wmx, wmn = O.build_sdf(mask, 0.5)
        assert np.array_equal(sdf.numpy(), want) and (mx, mn) == (wmx, wmn)
    # virtual border: two padded builds + per-voxel select, restated with the oracle
    mask = (occ > 0.5).astype(np.uint8)
    sdf, (mx, mn) = grid.ExtractSignedDistanceField(7.0, False, True)
    pad_filled = np.pad(mask, 1, constant_values=1)
    pad_free = np.pad(mask, 1, constant_values=0)
    a, amx, _ = O.build_sdf(pad_filled, 0.5)
    b, _, bmn = O.build_sdf(pad_free, 0.5)
    a, b = a[1:-1, 1:-1, 1:-1], b[1:-1, 1:-1, 1:-1]
    want = np.where(a >= 0.0, a, np.where(b <= -0.0, b, np.float32(0.0)))
    assert np.array_equal(sdf.numpy(), want)
    assert (mx, mn) == (amx, bmn)


def test_cpp_header_errors(P):
    """Non-uniform resolution cannot be expressed through the Python constructor; out-of-range axes surface the
    C ABI's error as RuntimeError through the C++ wrapper."""
    tf = P.Isometry3d([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    grid = P.CollisionMapGrid(tf, "w", 0.5, 2049, 1, 1, P.COLLISION_CELL(0.0))
    with pytest.raises(RuntimeError):
        grid.ExtractSignedDistanceField(0.0, False, False)


def test_build_distance_field_compat(P):
    """sdf_generation::BuildDistanceField wrapper (reference sdf_generation.hpp:95-207): distance_square to the seeds."""
    mask = synth.iid((20, 17, 13), 0.03, 8)
    seeds = [tuple(int(v) for v in p) for p in np.argwhere(mask != 0)]
    d2 = P.BuildDistanceFieldD2(20, 17, 13, seeds)
    want = O.edt_d2(mask).astype(np.float64)
    assert np.array_equal(d2, want)
    with pytest.raises(RuntimeError):
        P.BuildDistanceFieldD2(4, 4, 4, [(9, 0, 0)])          # out-of-bounds seed (:133)


def _gradient_restated(sdf: np.ndarray, res: float, edges: bool) -> np.ndarray:
    """GetGridAlignedGradient (reference sdf.hpp:432-526) line by line in numpy, with its arithmetic types: interior
    voxels subtract the float neighbours IN FLOAT and multiply by the double 1/(2 res) (:445-453); face voxels widen to
    double first and multiply by 1/((hi-lo) res) (:486-518); no gradient (NaN here) on faces without edge gradients."""
    nx, ny, nz = sdf.shape
    out = np.full(sdf.shape + (3,), np.nan, np.float64)
    inv2 = 1.0 / (2.0 * res)
    X, Y, Z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    interior = (X > 0) & (Y > 0) & (Z > 0) & (X < nx - 1) & (Y < ny - 1) & (Z < nz - 1)
    for a, (I, n) in enumerate(((X, nx), (Y, ny), (Z, nz))):
        lo, hi = np.maximum(I - 1, 0), np.minimum(I + 1, n - 1)
        idx_lo, idx_hi = [X, Y, Z], [X, Y, Z]
        idx_lo[a], idx_hi[a] = lo, hi
        vlo, vhi = sdf[tuple(idx_lo)], sdf[tuple(idx_hi)]                      # float32
        with np.errstate(invalid="ignore", divide="ignore"):
            g_int = (vhi - vlo).astype(np.float64) * inv2                       # float subtraction, then double
            inc = (hi - lo).astype(np.float64) * res
            g_edge = np.where(inc > 0.0, (vhi.astype(np.float64) - vlo.astype(np.float64)) * (1.0 / np.where(inc > 0, inc, 1.0)), 0.0)
        out[..., a] = np.where(interior, g_int, g_edge if edges else np.nan)
    return out


@pytest.mark.parametrize("shape", [(9, 7, 5), (20, 40, 1), (1, 1, 6), (3, 3, 3), (33, 17, 65)])
def test_device_gradient_is_bit_identical(P, shape):
    """the GPU gradient field == GetGridAlignedGradient of every voxel, bit for bit."""
    mask = synth.iid(shape, 0.1, 31)
    mask.reshape(-1)[mask.size // 2] = 1                  # never an all-free grid (its SDF is +inf everywhere)
    tf = P.Isometry3d([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    grid = P.CollisionMapGrid(tf, "w", 0.037, *shape, P.COLLISION_CELL(0.0))
    grid.SetOccupancy(mask.astype(np.float32))
    sdf = grid.ExtractSignedDistanceField(0.0, False, False)[0]
    field = sdf.numpy()
    def same(a, b):                                  # bit-identical, any NaN equal to any NaN (payloads are not specified)
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        na, nb = np.isnan(a), np.isnan(b)
        return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))

    for edges in (True, False):
        got = sdf.GetFullGradientArray(edges)
        want = _gradient_restated(field, sdf.GetResolution(), edges)
        assert got.shape == shape + (3,)
        assert same(got, want)
        # and against the C++ per-voxel path (host loop with the reference's float / double order)
        rng = np.random.default_rng(1)
        for _ in range(60):
            x, y, z = (int(rng.integers(0, n)) for n in shape)
            g = sdf.GetGradient(x, y, z, edges)
            if len(g) == 0:
                assert np.isnan(got[x, y, z]).all()
            else:
                assert same(g, got[x, y, z])
