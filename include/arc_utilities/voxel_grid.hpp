// include/arc_utilities/voxel_grid.hpp -- dense voxel grid container with the contract sdf_tools relies on.
//
// sdf_tools takes this container from the un-vendored catkin package arc_utilities (package.xml:21,35); it is not
// part of /root/reference.  This is an independent implementation of the surface the SDF-build path and its
// callers use, pinned by the reference's own expectations:
//   storage order idx = x*ny*nz + y*nz + z, z fastest            utils_3d.py:71-74, sdf.cpp:241-245, test_voxel_grid.cpp:20-31
//   default origin centred on the grid, cell centres at +-(k+0.5)*res   test_voxel_grid.cpp:66-141
//   GetImmutable / GetMutable return (reference, in_bounds)       sdf_generation.hpp:117-118,172-173,254
//   out-of-bounds queries yield the OOB value and `false`         collision_map.hpp:300-312
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../sdf_tools_b200/eigen_lite.hpp"

namespace VoxelGrid {

struct GRID_INDEX {
    int64_t x, y, z;
    GRID_INDEX() : x(-1), y(-1), z(-1) {}
    GRID_INDEX(int64_t in_x, int64_t in_y, int64_t in_z) : x(in_x), y(in_y), z(in_z) {}
    bool operator==(const GRID_INDEX& o) const { return x == o.x && y == o.y && z == o.z; }
};

template <typename T, typename BackingStore = std::vector<T>>
class VoxelGrid {
protected:
    Eigen::Isometry3d origin_transform_, inverse_origin_transform_;
    BackingStore data_;
    double cell_x_size_ = 0, cell_y_size_ = 0, cell_z_size_ = 0;
    double x_size_ = 0, y_size_ = 0, z_size_ = 0;
    int64_t stride1_ = 0, stride2_ = 0;
    int64_t num_x_cells_ = 0, num_y_cells_ = 0, num_z_cells_ = 0;
    T default_value_{}, oob_value_{};
    bool initialized_ = false;

    void Init(const Eigen::Isometry3d& origin, double cx, double cy, double cz, int64_t nx, int64_t ny, int64_t nz,
              const T& dflt, const T& oob) {
        if (nx < 0 || ny < 0 || nz < 0 || !(cx > 0) || !(cy > 0) || !(cz > 0))
            throw std::invalid_argument("VoxelGrid: sizes must be positive");
        origin_transform_ = origin;
        inverse_origin_transform_ = origin.inverse();
        cell_x_size_ = cx; cell_y_size_ = cy; cell_z_size_ = cz;
        num_x_cells_ = nx; num_y_cells_ = ny; num_z_cells_ = nz;
        x_size_ = cx * (double)nx; y_size_ = cy * (double)ny; z_size_ = cz * (double)nz;
        stride1_ = ny * nz; stride2_ = nz;
        default_value_ = dflt; oob_value_ = oob;
        data_.assign((size_t)(nx * ny * nz), dflt);
        initialized_ = true;
    }
    static Eigen::Isometry3d Centred(double xs, double ys, double zs) {
        Eigen::Isometry3d o = Eigen::Isometry3d::Identity();
#if SDF_B200_HAVE_EIGEN
        o.translation() = Eigen::Vector3d(-xs * 0.5, -ys * 0.5, -zs * 0.5);
#else
        o.setTranslation(-xs * 0.5, -ys * 0.5, -zs * 0.5);
#endif
        return o;
    }
    static int64_t Cells(double size, double cell) { return (int64_t)std::ceil(size / cell); }

public:
    VoxelGrid() = default;
    virtual ~VoxelGrid() = default;
    // the user-declared destructor would otherwise suppress the implicit moves, and "moving" a grid (e.g. the
    // SignedDistanceField returned by ExtractSignedDistanceField) would copy its whole backing store
    VoxelGrid(const VoxelGrid&) = default;
    VoxelGrid& operator=(const VoxelGrid&) = default;
    VoxelGrid(VoxelGrid&&) = default;
    VoxelGrid& operator=(VoxelGrid&&) = default;
    // cell-count constructors
    VoxelGrid(const Eigen::Isometry3d& origin, double cell, int64_t nx, int64_t ny, int64_t nz, const T& dflt)
    { Init(origin, cell, cell, cell, nx, ny, nz, dflt, dflt); }
    VoxelGrid(const Eigen::Isometry3d& origin, double cell, int64_t nx, int64_t ny, int64_t nz, const T& dflt, const T& oob)
    { Init(origin, cell, cell, cell, nx, ny, nz, dflt, oob); }
    VoxelGrid(double cell, int64_t nx, int64_t ny, int64_t nz, const T& dflt)
    { Init(Centred(cell * nx, cell * ny, cell * nz), cell, cell, cell, nx, ny, nz, dflt, dflt); }
    VoxelGrid(double cell, int64_t nx, int64_t ny, int64_t nz, const T& dflt, const T& oob)
    { Init(Centred(cell * nx, cell * ny, cell * nz), cell, cell, cell, nx, ny, nz, dflt, oob); }
    // metric-size constructors (cells = ceil(size / cell))
    VoxelGrid(const Eigen::Isometry3d& origin, double cell, double xs, double ys, double zs, const T& dflt)
    { Init(origin, cell, cell, cell, Cells(xs, cell), Cells(ys, cell), Cells(zs, cell), dflt, dflt); }
    VoxelGrid(const Eigen::Isometry3d& origin, double cell, double xs, double ys, double zs, const T& dflt, const T& oob)
    { Init(origin, cell, cell, cell, Cells(xs, cell), Cells(ys, cell), Cells(zs, cell), dflt, oob); }
    VoxelGrid(double cell, double xs, double ys, double zs, const T& dflt) {
        const int64_t nx = Cells(xs, cell), ny = Cells(ys, cell), nz = Cells(zs, cell);
        Init(Centred(cell * nx, cell * ny, cell * nz), cell, cell, cell, nx, ny, nz, dflt, dflt);
    }
    VoxelGrid(double cell, double xs, double ys, double zs, const T& dflt, const T& oob) {
        const int64_t nx = Cells(xs, cell), ny = Cells(ys, cell), nz = Cells(zs, cell);
        Init(Centred(cell * nx, cell * ny, cell * nz), cell, cell, cell, nx, ny, nz, dflt, oob);
    }
    // non-uniform cells (lets callers hit the "Grid must have uniform resolution" check, sdf_generation.hpp:277-281)
    VoxelGrid(const Eigen::Isometry3d& origin, double cx, double cy, double cz, int64_t nx, int64_t ny, int64_t nz,
              const T& dflt, const T& oob) { Init(origin, cx, cy, cz, nx, ny, nz, dflt, oob); }

    bool IsInitialized() const { return initialized_; }
    int64_t GetNumXCells() const { return num_x_cells_; }
    int64_t GetNumYCells() const { return num_y_cells_; }
    int64_t GetNumZCells() const { return num_z_cells_; }
    double GetXSize() const { return x_size_; }
    double GetYSize() const { return y_size_; }
    double GetZSize() const { return z_size_; }
    Eigen::Vector3d GetCellSizes() const { return Eigen::Vector3d(cell_x_size_, cell_y_size_, cell_z_size_); }
    const Eigen::Isometry3d& GetOriginTransform() const { return origin_transform_; }
    const Eigen::Isometry3d& GetInverseOriginTransform() const { return inverse_origin_transform_; }
    const T& GetDefaultValue() const { return default_value_; }
    const T& GetOOBValue() const { return oob_value_; }
    const BackingStore& GetImmutableRawData() const { return data_; }
    BackingStore& GetMutableRawData() { return data_; }
    bool SetRawData(const BackingStore& d) {
        if ((int64_t)d.size() != num_x_cells_ * num_y_cells_ * num_z_cells_) return false;
        data_ = d;
        return true;
    }

    bool IndexInBounds(int64_t x, int64_t y, int64_t z) const {
        return x >= 0 && y >= 0 && z >= 0 && x < num_x_cells_ && y < num_y_cells_ && z < num_z_cells_;
    }
    bool IndexInBounds(const GRID_INDEX& i) const { return IndexInBounds(i.x, i.y, i.z); }
    int64_t GetDataIndex(int64_t x, int64_t y, int64_t z) const { return x * stride1_ + y * stride2_ + z; }
    int64_t GetDataIndex(const GRID_INDEX& i) const { return GetDataIndex(i.x, i.y, i.z); }

    // ---- location <-> index --------------------------------------------------------------------------------
    GRID_INDEX LocationToGridIndex4d(const Eigen::Vector4d& location) const {
        const Eigen::Vector4d p = inverse_origin_transform_ * location;
        return GRID_INDEX((int64_t)std::floor(p(0) / cell_x_size_), (int64_t)std::floor(p(1) / cell_y_size_),
                          (int64_t)std::floor(p(2) / cell_z_size_));
    }
    GRID_INDEX LocationToGridIndex3d(const Eigen::Vector3d& l) const {
        return LocationToGridIndex4d(Eigen::Vector4d(l(0), l(1), l(2), 1.0));
    }
    GRID_INDEX LocationToGridIndex(double x, double y, double z) const {
        return LocationToGridIndex4d(Eigen::Vector4d(x, y, z, 1.0));
    }
    Eigen::Vector4d GridIndexToLocation(int64_t x, int64_t y, int64_t z) const {
        return origin_transform_ * Eigen::Vector4d(cell_x_size_ * ((double)x + 0.5), cell_y_size_ * ((double)y + 0.5),
                                                   cell_z_size_ * ((double)z + 0.5), 1.0);
    }
    Eigen::Vector4d GridIndexToLocation(const GRID_INDEX& i) const { return GridIndexToLocation(i.x, i.y, i.z); }
    bool LocationInBounds(double x, double y, double z) const { return IndexInBounds(LocationToGridIndex(x, y, z)); }

    // ---- access -------------------------------------------------------------------------------------------
    virtual std::pair<const T&, bool> GetImmutable(int64_t x, int64_t y, int64_t z) const {
        if (IndexInBounds(x, y, z)) return std::pair<const T&, bool>(data_[(size_t)GetDataIndex(x, y, z)], true);
        return std::pair<const T&, bool>(oob_value_, false);
    }
    virtual std::pair<const T&, bool> GetImmutable(const GRID_INDEX& i) const { return GetImmutable(i.x, i.y, i.z); }
    virtual std::pair<const T&, bool> GetImmutable(double x, double y, double z) const {
        return GetImmutable(LocationToGridIndex(x, y, z));
    }
    virtual std::pair<T&, bool> GetMutable(int64_t x, int64_t y, int64_t z) {
        if (IndexInBounds(x, y, z)) return std::pair<T&, bool>(data_[(size_t)GetDataIndex(x, y, z)], true);
        return std::pair<T&, bool>(oob_value_, false);
    }
    virtual std::pair<T&, bool> GetMutable(const GRID_INDEX& i) { return GetMutable(i.x, i.y, i.z); }
    virtual std::pair<T&, bool> GetMutable(double x, double y, double z) { return GetMutable(LocationToGridIndex(x, y, z)); }
    virtual bool SetValue(int64_t x, int64_t y, int64_t z, const T& v) {
        if (!IndexInBounds(x, y, z)) return false;
        data_[(size_t)GetDataIndex(x, y, z)] = v;
        return true;
    }
    virtual bool SetValue(const GRID_INDEX& i, const T& v) { return SetValue(i.x, i.y, i.z, v); }
    virtual bool SetValue(double x, double y, double z, const T& v) { return SetValue(LocationToGridIndex(x, y, z), v); }
};

}  // namespace VoxelGrid
