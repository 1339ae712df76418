// oracle/ref_shim.hpp -- TEST INFRASTRUCTURE ONLY.
// The smallest container surface that lets the reference's own hot-path text
// (/root/reference/include/sdf_tools/sdf_generation.hpp:19-271, sed-extracted at build time into
// oracle/_ref/, never committed) compile unmodified.  The real containers come from the un-vendored
// catkin package arc_utilities (no version pin: package.xml:21,35); the surface used on the hot path is
//   VoxelGrid<T>(origin, res, nx, ny, nz, default)           sdf_generation.hpp:105
//   GetNum{X,Y,Z}Cells()                                      :107
//   GetMutable(GRID_INDEX) / GetMutable(x,y,z) -> pair<T&,bool>   :117, :172
//   GetImmutable(x,y,z) -> pair<const T&,bool>                :254
//   SignedDistanceField(origin, frame, res, nx, ny, nz, oob), GetResolution(), SetValue()   :245, :254, :265
// Layout idx = x*ny*nz + y*nz + z (utils_3d.py:71-74, sdf.cpp:241-245, test_voxel_grid.cpp:20-31).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include <functional>
#include <limits>
#include <stdexcept>
#include <utility>

namespace Eigen { struct Isometry3d {}; }

namespace VoxelGrid {
struct GRID_INDEX {
    int64_t x, y, z;
    GRID_INDEX() : x(-1), y(-1), z(-1) {}
    GRID_INDEX(int64_t in_x, int64_t in_y, int64_t in_z) : x(in_x), y(in_y), z(in_z) {}
};

template <typename T>
class VoxelGrid {
protected:
    std::vector<T> data_;
    double res_;
    int64_t nx_, ny_, nz_;
    T default_, oob_;
public:
    VoxelGrid(const Eigen::Isometry3d&, double res, int64_t nx, int64_t ny, int64_t nz, const T& dflt)
        : data_((size_t)(nx * ny * nz), dflt), res_(res), nx_(nx), ny_(ny), nz_(nz), default_(dflt), oob_(dflt) {}
    int64_t GetNumXCells() const { return nx_; }
    int64_t GetNumYCells() const { return ny_; }
    int64_t GetNumZCells() const { return nz_; }
    bool InBounds(int64_t x, int64_t y, int64_t z) const {
        return x >= 0 && y >= 0 && z >= 0 && x < nx_ && y < ny_ && z < nz_;
    }
    std::pair<T&, bool> GetMutable(int64_t x, int64_t y, int64_t z) {
        if (InBounds(x, y, z)) return std::pair<T&, bool>(data_[(size_t)((x * ny_ + y) * nz_ + z)], true);
        return std::pair<T&, bool>(oob_, false);
    }
    std::pair<T&, bool> GetMutable(const GRID_INDEX& i) { return GetMutable(i.x, i.y, i.z); }
    std::pair<const T&, bool> GetImmutable(int64_t x, int64_t y, int64_t z) const {
        if (InBounds(x, y, z)) return std::pair<const T&, bool>(data_[(size_t)((x * ny_ + y) * nz_ + z)], true);
        return std::pair<const T&, bool>(oob_, false);
    }
    const std::vector<T>& GetImmutableRawData() const { return data_; }
};
}  // namespace VoxelGrid

namespace sdf_tools {
class SignedDistanceField : public VoxelGrid::VoxelGrid<float> {
    std::string frame_;
public:
    SignedDistanceField(const Eigen::Isometry3d& o, const std::string& frame, double res,
                        int64_t nx, int64_t ny, int64_t nz, float oob)
        : VoxelGrid::VoxelGrid<float>(o, res, nx, ny, nz, oob), frame_(frame) {}
    double GetResolution() const { return res_; }
    bool SetValue(int64_t x, int64_t y, int64_t z, const float& v) {
        auto q = GetMutable(x, y, z);
        if (q.second) { q.first = v; return true; }
        return false;
    }
};
}  // namespace sdf_tools
