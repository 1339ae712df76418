"""The C++ face (include/sdf_tools/*.hpp, the reference's signatures) exercised by a COMPILED C++ caller on the GPU:
CollisionMapGrid::ExtractSignedDistanceField, the GPU gradient field, the batched TaggedObjectCollisionMapGrid SDFs
(MakeObjectSDFs, ExtractFreeAndNamedObjectsSignedDistanceField, ExtractSignedDistanceField(objects_to_use)) and the
multi-device context selected with SDF_B200_DEVICES -- every result compared with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from sdf_tools_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CPP = r'''
#include <sdf_tools/collision_map.hpp>
#include <sdf_tools/tagged_object_collision_map.hpp>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace sdf_tools;
static void dump(const char* dir, const char* name, const void* p, size_t bytes) {
    std::string f = std::string(dir) + "/" + name; FILE* fp = std::fopen(f.c_str(), "wb");
    if (!fp || std::fwrite(p, 1, bytes, fp) != bytes) std::abort();
    std::fclose(fp);
}
int main(int argc, char** argv) {
    if (argc != 5) return 2;
    const char* dir = argv[1];
    const int64_t nx = std::atoll(argv[2]), ny = std::atoll(argv[3]), nz = std::atoll(argv[4]);
    const size_t N = (size_t)(nx * ny * nz);
    std::vector<float> occ(N); std::vector<uint32_t> ids(N);
    { std::string f = std::string(dir) + "/occ.bin"; FILE* fp = std::fopen(f.c_str(), "rb"); if (!fp || std::fread(occ.data(), 4, N, fp) != N) return 3; std::fclose(fp); }
    { std::string f = std::string(dir) + "/ids.bin"; FILE* fp = std::fopen(f.c_str(), "rb"); if (!fp || std::fread(ids.data(), 4, N, fp) != N) return 3; std::fclose(fp); }
    const double res = 0.04;
    // ---- CollisionMapGrid (collision_map.hpp:680-712) + gradient field
    CollisionMapGrid g(Eigen::Isometry3d::Identity(), "world", res, nx, ny, nz, COLLISION_CELL(0.0f));
    for (size_t i = 0; i < N; ++i) g.GetMutableRawData()[i] = COLLISION_CELL(occ[i]);
    auto r = g.ExtractSignedDistanceField(7.0f, true, false);
    dump(dir, "cmap_sdf.bin", r.first.GetImmutableRawData().data(), N * 4);
    double ext[2] = {r.second.first, r.second.second};
    dump(dir, "cmap_ext.bin", ext, sizeof(ext));
    const std::vector<double> grad = sdf_generation::GetFullGradientArray(r.first, true);
    dump(dir, "cmap_grad.bin", grad.data(), grad.size() * 8);
    // per-voxel C++ gradient at a few voxels must be the same doubles
    for (int64_t t = 0; t < 50; ++t) {
        const int64_t x = (t * 7) % nx, y = (t * 5) % ny, z = (t * 3) % nz;
        const std::vector<double> gv = r.first.GetGridAlignedGradient(x, y, z, true);
        const size_t q = (size_t)((x * ny + y) * nz + z) * 3;
        for (int a = 0; a < 3; ++a) if (!(gv[a] == grad[q + a]) && !(gv[a] != gv[a] && grad[q + a] != grad[q + a])) { std::printf("gradient mismatch\n"); return 4; }
    }
    // ---- TaggedObjectCollisionMapGrid (tagged_object_collision_map.hpp:730-915)
    TaggedObjectCollisionMapGrid tg(Eigen::Isometry3d::Identity(), "world", res, nx, ny, nz, TAGGED_OBJECT_COLLISION_CELL(0.0f, 0u));
    for (size_t i = 0; i < N; ++i) tg.GetMutableRawData()[i] = TAGGED_OBJECT_COLLISION_CELL(occ[i], ids[i]);
    const auto objs = tg.MakeObjectSDFs(std::vector<uint32_t>{1u, 2u, 3u, 9u}, false, false);
    for (const auto& kv : objs) { char nm[64]; std::snprintf(nm, sizeof(nm), "obj_%u.bin", kv.first); dump(dir, nm, kv.second.GetImmutableRawData().data(), N * 4); }
    const auto all = tg.MakeAllObjectSDFs(true, false);
    std::printf("all objects: %zu\n", all.size());
    auto fn = tg.ExtractFreeAndNamedObjectsSignedDistanceField(1.0f, false);
    dump(dir, "free_named.bin", fn.first.GetImmutableRawData().data(), N * 4);
    double e2[2] = {fn.second.first, fn.second.second};
    dump(dir, "free_named_ext.bin", e2, sizeof(e2));
    auto any = tg.ExtractSignedDistanceField(1.0f, std::vector<uint32_t>{}, true, false);
    dump(dir, "any.bin", any.first.GetImmutableRawData().data(), N * 4);
    auto two = tg.ExtractSignedDistanceField(1.0f, std::vector<uint32_t>{1u, 3u}, false, false);
    dump(dir, "two.bin", two.first.GetImmutableRawData().data(), N * 4);
    // ---- virtual border through the predicate overload: the predicate runs twice per voxel (sdf_generation.hpp:381-382)
    size_t calls = 0;
    const std::function<bool(const VoxelGrid::GRID_INDEX&)> counted = [&](const VoxelGrid::GRID_INDEX& i) {
        ++calls;
        return g.GetImmutable(i).first.occupancy > 0.5f;
    };
    auto vb = sdf_generation::ExtractSignedDistanceField(g, counted, 3.0f, "world", true);
    dump(dir, "border.bin", vb.first.GetImmutableRawData().data(), N * 4);
    double e3[2] = {vb.second.first, vb.second.second};
    dump(dir, "border_ext.bin", e3, sizeof(e3));
    std::printf("predicate calls %zu of %zu voxels\n", calls, N);
    std::printf("devices %d\n", sdf_b200_num_devices(sdf_generation::detail::Context()));
    return 0;
}
'''


def _build(tmp_path):
    src = tmp_path / "face.cpp"
    src.write_text(CPP)
    exe = tmp_path / "face"
    libdir = os.path.join(ROOT, "sdf_tools_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-Wshadow", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lsdf_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def _ndev():
    from sdf_tools_b200 import _capi
    return _capi.load().sdf_b200_device_count()


@pytest.mark.parametrize("devices", ["0", "all"])
def test_compiled_cpp_caller_against_oracle(tmp_path, devices):
    if devices == "all" and _ndev() < 2:
        pytest.skip("needs >= 2 GPUs")
    shape = (48, 40, 64)
    rng = np.random.default_rng(3)
    occ = synth.octree(shape, 0.3, 5)                                   # {0, 0.5, 1}
    ids = np.zeros(shape, np.uint32)
    blocks = rng.integers(0, 5, size=(6, 5, 8)).astype(np.uint32)       # object ids 0..4 in 8^3 blocks
    ids[...] = np.kron(blocks, np.ones((8, 8, 8), np.uint32))
    occ.astype(np.float32).tofile(tmp_path / "occ.bin")
    ids.tofile(tmp_path / "ids.bin")
    exe = _build(tmp_path)
    env = dict(os.environ)
    if devices == "all":
        env["SDF_B200_DEVICES"] = ",".join(str(i) for i in range(min(_ndev(), 8)))
    out = subprocess.run([str(exe), str(tmp_path), *map(str, shape)], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    if devices == "all":
        assert f"devices {min(_ndev(), 8)}" in out.stdout
    res = 0.04
    load = lambda n, dt=np.float32: np.fromfile(tmp_path / n, dtype=dt)
    filled = occ > 0.5
    filled_u = filled | (occ == 0.5)
    # CollisionMapGrid, unknown_is_filled = true
    want, wmx, wmn = O.build_sdf(filled_u.astype(np.uint8), res)
    got = load("cmap_sdf.bin").reshape(shape)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert tuple(load("cmap_ext.bin", np.float64)) == (wmx, wmn)
    grad = load("cmap_grad.bin", np.float64).reshape(shape + (3,))
    from test_gpu_api import _gradient_restated
    wg = _gradient_restated(want, res, True)
    assert np.array_equal(np.isnan(grad), np.isnan(wg)) and np.array_equal(grad[~np.isnan(grad)], wg[~np.isnan(wg)])
    # MakeObjectSDFs: one reference build per object id
    for k in (1, 2, 3, 9):
        w, _, _ = O.build_sdf((filled & (ids == k)).astype(np.uint8), res)
        assert np.array_equal(load(f"obj_{k}.bin").reshape(shape).view(np.uint32), w.view(np.uint32)), k
    assert f"all objects: {len(set(np.unique(ids)) - {0})}" in out.stdout
    # free + named objects (:730-811)
    a, amx, _ = O.build_sdf(filled.astype(np.uint8), res)
    b, _, bmn = O.build_sdf((filled & (ids > 0)).astype(np.uint8), res)
    wfn = np.where(a >= 0.0, a, np.where(b <= -0.0, b, np.float32(0.0)))
    assert np.array_equal(load("free_named.bin").reshape(shape), wfn)
    assert tuple(load("free_named_ext.bin", np.float64)) == (amx, bmn)
    # virtual border (sdf_generation.hpp:287-419), restated with two padded oracle builds
    pa, pamx, _ = O.build_sdf(np.pad(filled.astype(np.uint8), 1, constant_values=1), res)
    pb, _, pbmn = O.build_sdf(np.pad(filled.astype(np.uint8), 1, constant_values=0), res)
    pa, pb = pa[1:-1, 1:-1, 1:-1], pb[1:-1, 1:-1, 1:-1]
    assert np.array_equal(load("border.bin").reshape(shape), np.where(pa >= 0.0, pa, np.where(pb <= -0.0, pb, np.float32(0.0))))
    assert tuple(load("border_ext.bin", np.float64)) == (pamx, pbmn)
    assert f"predicate calls {2 * occ.size} of {occ.size} voxels" in out.stdout
    w, _, _ = O.build_sdf(filled_u.astype(np.uint8), res)
    assert np.array_equal(load("any.bin").reshape(shape).view(np.uint32), w.view(np.uint32))
    w, _, _ = O.build_sdf((filled & ((ids == 1) | (ids == 3))).astype(np.uint8), res)
    assert np.array_equal(load("two.bin").reshape(shape).view(np.uint32), w.view(np.uint32))


def test_batched_object_sdfs_python(ctx):
    """SURVEY.md 8f-3 through the C ABI: 128^3 labelled grid, 11 objects (two passes over the cells: 8 + 3 queries)."""
    from sdf_tools_b200.api import TAGGED_CELL_DTYPE, QUERY_ANY, QUERY_NAMED
    shape = (128, 128, 128)
    rng = np.random.default_rng(8)
    occ = synth.octree(shape, 0.3, 5)
    ids = np.kron(rng.integers(0, 12, size=(8, 8, 8)), np.ones((16, 16, 16), np.int64)).astype(np.uint32)
    cells = np.zeros(shape, dtype=TAGGED_CELL_DTYPE)
    cells["occupancy"], cells["object_id"], cells["component"] = occ, ids, 77
    queries = list(range(1, 12)) + [(0, QUERY_ANY), (0, QUERY_NAMED), 0]
    for unknown in (False, True):
        filled = (occ > 0.5) | (unknown & (occ == 0.5))
        sdfs, ext = ctx.build_object_sdfs(cells, 0.05, queries, unknown)
        assert sdfs.shape == (len(queries),) + shape
        for k, q in enumerate(queries):
            m = filled & (ids == q) if not isinstance(q, tuple) else (filled if q[1] == QUERY_ANY else filled & (ids > 0))
            want, wmx, wmn = O.build_sdf(m.astype(np.uint8), 0.05)
            assert np.array_equal(sdfs[k].view(np.uint32), want.view(np.uint32)), (k, q)
            assert ext[k] == (wmx, wmn)


def test_virtual_border_on_device(ctx):
    """SURVEY.md 8f-2 through the C ABI: shapes with one-cell axes (no shell there), a second mask for the second sweep,
    and the cost of the whole call against a plain build of the same grid."""
    import time
    rng = np.random.default_rng(12)
    for shape in [(24, 20, 16), (20, 40, 1), (1, 1, 9), (5, 1, 7)]:
        m = (rng.random(shape) < 0.2).astype(np.uint8)
        m2 = (rng.random(shape) < 0.2).astype(np.uint8)
        for second in (None, m2):
            got, (mx, mn) = ctx.build_with_border(m, 0.5, second)
            pad = [(1, 1) if n > 1 else (0, 0) for n in shape]
            a, amx, _ = O.build_sdf(np.pad(m, pad, constant_values=1), 0.5)
            b, _, bmn = O.build_sdf(np.pad(m if second is None else second, pad, constant_values=0), 0.5)
            crop = tuple(slice(p[0], a.shape[i] - p[1]) for i, p in enumerate(pad))
            a, b = a[crop], b[crop]
            assert np.array_equal(got, np.where(a >= 0.0, a, np.where(b <= -0.0, b, np.float32(0.0)))), shape
            assert (mx, mn) == (amx, bmn)
    mask = synth.make_mask("c2")                                       # 256^3: whole call vs a plain host build
    ctx.build_with_border(mask, 0.05); ctx.build_from_mask(mask, 0.05)
    t0 = time.perf_counter(); ctx.build_from_mask(mask, 0.05); t_plain = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.build_with_border(mask, 0.05); t_border = time.perf_counter() - t0
    print(f"256^3: plain {t_plain * 1e3:.2f} ms, with border {t_border * 1e3:.2f} ms, ratio {t_border / t_plain:.2f}")
    assert t_border < 3.0 * t_plain
