// sdf_tools_b200/csrc/pybind_module.cpp -- the Python face `pysdf_tools` for the SDF-build subset.
// Same class / method names and argument meaning as the reference module (src/sdf_tools/bindings.cpp:15-106) so that
// callers such as utils_2d.py / utils_3d.py keep working, plus bulk numpy ingest / egress: the
// reference's per-voxel Python SetValue loop (utils_3d.py:62-67) and list copy (bindings.cpp:37) cost far more than
// the GPU build itself.  Serialisation / message helpers are out of scope and not bound.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <array>
#include <cstring>

#include <sdf_tools/collision_map.hpp>

namespace py = pybind11;
using namespace sdf_tools;
using VoxelGridVecd = VoxelGrid::VoxelGrid<std::vector<double>>;

static Eigen::Isometry3d IsometryFromRows(const std::vector<std::vector<double>>& rows) {
    if (rows.size() != 4) throw std::invalid_argument("Isometry3d expects a 4x4 matrix");
    Eigen::Matrix4d m;
    for (int r = 0; r < 4; ++r) {
        if (rows[r].size() != 4) throw std::invalid_argument("Isometry3d expects a 4x4 matrix");
        for (int c = 0; c < 4; ++c) m(r, c) = rows[r][c];
    }
    return Eigen::Isometry3d(m);
}

PYBIND11_MODULE(pysdf_tools, m) {
    m.doc() = "sdf_tools SDF-build subset on B200 (libsdf_b200.so); API of the reference's pysdf_tools";

    py::class_<COLLISION_CELL>(m, "COLLISION_CELL")
        .def(py::init<float>())
        .def(py::init<float, uint32_t>())
        .def_readwrite("occupancy", &COLLISION_CELL::occupancy)
        .def_readwrite("component", &COLLISION_CELL::component);

    py::class_<Eigen::Isometry3d>(m, "Isometry3d")
        .def(py::init(&IsometryFromRows))
        .def("translation", [](const Eigen::Isometry3d& t) {
            const Eigen::Vector3d v = t.translation();
            return std::vector<double>{v(0), v(1), v(2)};
        })
        .def("matrix", [](const Eigen::Isometry3d& t) {
            std::vector<std::vector<double>> rows(4, std::vector<double>(4));
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) rows[r][c] = t.matrix()(r, c);
            return rows;
        });

    py::class_<SignedDistanceField>(m, "SignedDistanceField")
        .def(py::init<>())
        .def("GetRawData", &SignedDistanceField::GetImmutableRawData, "flat list, x-major / z-fastest")
        .def("numpy", [](const SignedDistanceField& s) {       // (nx, ny, nz) float32 copy in one memcpy
            py::array_t<float> a({s.GetNumXCells(), s.GetNumYCells(), s.GetNumZCells()});
            const auto& d = s.GetImmutableRawData();
            if (!d.empty()) std::memcpy(a.mutable_data(), d.data(), d.size() * sizeof(float));
            return a;
        })
        .def("GetFullGradient", &SignedDistanceField::GetFullGradient)
        .def("GetFullGradientArray",                            // (nx, ny, nz, 3) float64, one GPU pass
             [](const SignedDistanceField& s, bool edges) {
                 py::array_t<double> a({s.GetNumXCells(), s.GetNumYCells(), s.GetNumZCells(), (int64_t)3});
                 if (a.size() > 0) {
                     sdf_b200_ctx* ctx = sdf_generation::detail::Context();
                     sdf_generation::detail::Check(
                         sdf_b200_gradient_host(ctx, s.GetImmutableRawData().data(), s.GetNumXCells(), s.GetNumYCells(),
                                                s.GetNumZCells(), s.GetResolution(), edges ? 1 : 0, 1, a.mutable_data()), ctx);
                 }
                 return a;
             },
             "grid-aligned gradient of every voxel, bit-identical to GetGridAlignedGradient per voxel",
             py::arg("enable_edge_gradients") = true)
        .def("GetResolution", &SignedDistanceField::GetResolution)
        .def("GetFrame", &SignedDistanceField::GetFrame)
        .def("GetGradient",
             [](const SignedDistanceField& s, int64_t x, int64_t y, int64_t z, bool e) { return s.GetGradient(x, y, z, e); },
             "get the gradient based on index", py::arg("x_index"), py::arg("y_index"), py::arg("z_index"),
             py::arg("enable_edge_gradients") = false)
        .def("GetOriginTransform", [](const SignedDistanceField& s) { return s.GetOriginTransform(); })
        .def("GetValueByCoordinates",
             [](const SignedDistanceField& s, double x, double y, double z) { const auto q = s.GetImmutable(x, y, z); return std::make_pair(q.first, q.second); },
             py::arg("x"), py::arg("y"), py::arg("z"))
        .def("GetValueByIndex",
             [](const SignedDistanceField& s, int64_t x, int64_t y, int64_t z) { const auto q = s.GetImmutable(x, y, z); return std::make_pair(q.first, q.second); },
             py::arg("x_index"), py::arg("y_index"), py::arg("z_index"))
        .def("GetNumXCells", &SignedDistanceField::GetNumXCells)
        .def("GetNumYCells", &SignedDistanceField::GetNumYCells)
        .def("GetNumZCells", &SignedDistanceField::GetNumZCells);

    py::class_<CollisionMapGrid>(m, "CollisionMapGrid")
        .def(py::init<Eigen::Isometry3d const&, std::string, double, int64_t, int64_t, int64_t, COLLISION_CELL const&>())
        .def("SetValue", [](CollisionMapGrid& g, int64_t x, int64_t y, int64_t z, const COLLISION_CELL& c) { return g.SetValue(x, y, z, c); })
        .def("SetValueByCoordinates", [](CollisionMapGrid& g, double x, double y, double z, const COLLISION_CELL& c) { return g.SetValue(x, y, z, c); })
        .def("SetOccupancy",                                    // bulk ingest: (nx, ny, nz) array of occupancies
             [](CollisionMapGrid& g, py::array_t<float, py::array::c_style | py::array::forcecast> occ) {
                 if (occ.ndim() != 3 || occ.shape(0) != g.GetNumXCells() || occ.shape(1) != g.GetNumYCells() ||
                     occ.shape(2) != g.GetNumZCells())
                     throw std::invalid_argument("SetOccupancy: array must have shape (nx, ny, nz)");
                 auto& d = g.GetMutableRawData();
                 const float* p = occ.data();
                 for (size_t i = 0; i < d.size(); ++i) { d[i].occupancy = p[i]; d[i].component = 0u; }
             })
        .def("GetRawData", &CollisionMapGrid::GetImmutableRawData, "flat list of COLLISION_CELL")
        .def("GetValueByCoordinates",
             [](const CollisionMapGrid& g, double x, double y, double z) { const auto q = g.GetImmutable(x, y, z); return std::make_pair(q.first, q.second); },
             py::arg("x"), py::arg("y"), py::arg("z"))
        .def("GetValueByIndex",
             [](const CollisionMapGrid& g, int64_t x, int64_t y, int64_t z) { const auto q = g.GetImmutable(x, y, z); return std::make_pair(q.first, q.second); },
             py::arg("x_index"), py::arg("y_index"), py::arg("z_index"))
        .def("GetNumXCells", &CollisionMapGrid::GetNumXCells)
        .def("GetNumYCells", &CollisionMapGrid::GetNumYCells)
        .def("GetNumZCells", &CollisionMapGrid::GetNumZCells)
        .def("ExtractSignedDistanceField", &CollisionMapGrid::ExtractSignedDistanceField, py::arg("oob_value"),
             py::arg("unknown_is_filled"), py::arg("add_virtual_border"));

    // sdf_generation::BuildDistanceField (compat wrapper): seeds as (x, y, z) triples -> flat distance_square array
    m.def("BuildDistanceFieldD2", [](int64_t nx, int64_t ny, int64_t nz, const std::vector<std::array<int64_t, 3>>& seeds) {
        std::vector<VoxelGrid::GRID_INDEX> pts;
        pts.reserve(seeds.size());
        for (const auto& s : seeds) pts.emplace_back(s[0], s[1], s[2]);
        const sdf_generation::DistanceField f = sdf_generation::BuildDistanceField(Eigen::Isometry3d::Identity(), 1.0, nx, ny, nz, pts);
        py::array_t<double> out({nx, ny, nz});
        double* o = out.mutable_data();
        const auto& d = f.GetImmutableRawData();
        for (size_t i = 0; i < d.size(); ++i) o[i] = d[i].distance_square;
        return out;
    });

    py::class_<VoxelGridVecd>(m, "VoxelGrid")
        .def(py::init<>())
        .def("GetRawData", &VoxelGridVecd::GetImmutableRawData)
        .def("GetNumXCells", &VoxelGridVecd::GetNumXCells)
        .def("GetNumYCells", &VoxelGridVecd::GetNumYCells)
        .def("GetNumZCells", &VoxelGridVecd::GetNumZCells)
        .def("GetValueByCoordinates",
             [](const VoxelGridVecd& g, double x, double y, double z) { const auto q = g.GetImmutable(x, y, z); return std::make_pair(q.first, q.second); },
             py::arg("x"), py::arg("y"), py::arg("z"))
        .def("GetValueByIndex",
             [](const VoxelGridVecd& g, int64_t x, int64_t y, int64_t z) { const auto q = g.GetImmutable(x, y, z); return std::make_pair(q.first, q.second); },
             py::arg("x_index"), py::arg("y_index"), py::arg("z_index"));
}
