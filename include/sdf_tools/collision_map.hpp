// include/sdf_tools/collision_map.hpp -- sdf_tools::COLLISION_CELL and CollisionMapGrid, the INPUT type of the
// SDF-build hot path (reference include/sdf_tools/collision_map.hpp:20-32 cell, :34 class, :164-270 ctors,
// :680-712 ExtractSignedDistanceField).  Build subset only: connected components, topology, resampling,
// serialisation and RViz export are out of scope and are not provided.
#pragma once
#include <cstdint>
#include <string>
#include <utility>

#include "../arc_utilities/voxel_grid.hpp"
#include "sdf.hpp"
#include "sdf_generation.hpp"

namespace sdf_tools {

struct COLLISION_CELL {
    float occupancy;
    uint32_t component;
    COLLISION_CELL() : occupancy(0.0f), component(0u) {}
    explicit COLLISION_CELL(const float in_occupancy) : occupancy(in_occupancy), component(0u) {}
    COLLISION_CELL(const float in_occupancy, const uint32_t in_component) : occupancy(in_occupancy), component(in_component) {}
};
static_assert(sizeof(COLLISION_CELL) == 8, "COLLISION_CELL must stay {float, uint32}: the GPU packs it as 8 bytes");

class CollisionMapGrid : public VoxelGrid::VoxelGrid<COLLISION_CELL> {
protected:
    uint32_t number_of_components_ = 0u;
    std::string frame_;
    bool components_valid_ = false;
    using Base = ::VoxelGrid::VoxelGrid<COLLISION_CELL>;

public:
    CollisionMapGrid() : Base(), frame_("") {}
    // cell-count constructors (the Python binding exposes the first one, bindings.cpp:68)
    CollisionMapGrid(const Eigen::Isometry3d& origin, const std::string& frame, double resolution, int64_t x_cells,
                     int64_t y_cells, int64_t z_cells, const COLLISION_CELL& oob_default)
        : Base(origin, resolution, x_cells, y_cells, z_cells, oob_default), frame_(frame) {}
    CollisionMapGrid(const Eigen::Isometry3d& origin, const std::string& frame, double resolution, int64_t x_cells,
                     int64_t y_cells, int64_t z_cells, const COLLISION_CELL& dflt, const COLLISION_CELL& oob)
        : Base(origin, resolution, x_cells, y_cells, z_cells, dflt, oob), frame_(frame) {}
    CollisionMapGrid(const std::string& frame, double resolution, int64_t x_cells, int64_t y_cells, int64_t z_cells,
                     const COLLISION_CELL& oob_default)
        : Base(resolution, x_cells, y_cells, z_cells, oob_default), frame_(frame) {}
    CollisionMapGrid(const std::string& frame, double resolution, int64_t x_cells, int64_t y_cells, int64_t z_cells,
                     const COLLISION_CELL& dflt, const COLLISION_CELL& oob)
        : Base(resolution, x_cells, y_cells, z_cells, dflt, oob), frame_(frame) {}
    // metric-size constructors
    CollisionMapGrid(const Eigen::Isometry3d& origin, const std::string& frame, double resolution, double x_size,
                     double y_size, double z_size, const COLLISION_CELL& oob_default)
        : Base(origin, resolution, x_size, y_size, z_size, oob_default), frame_(frame) {}
    CollisionMapGrid(const Eigen::Isometry3d& origin, const std::string& frame, double resolution, double x_size,
                     double y_size, double z_size, const COLLISION_CELL& dflt, const COLLISION_CELL& oob)
        : Base(origin, resolution, x_size, y_size, z_size, dflt, oob), frame_(frame) {}
    CollisionMapGrid(const std::string& frame, double resolution, double x_size, double y_size, double z_size,
                     const COLLISION_CELL& oob_default)
        : Base(resolution, x_size, y_size, z_size, oob_default), frame_(frame) {}
    CollisionMapGrid(const std::string& frame, double resolution, double x_size, double y_size, double z_size,
                     const COLLISION_CELL& dflt, const COLLISION_CELL& oob)
        : Base(resolution, x_size, y_size, z_size, dflt, oob), frame_(frame) {}

    double GetResolution() const { return GetCellSizes().x(); }
    std::string GetFrame() const { return frame_; }
    void SetFrame(const std::string& f) { frame_ = f; }
    bool AreComponentsValid() const { return components_valid_; }

    // every mutation invalidates the connected components (reference collision_map.hpp:327,343,395,412)
    using Base::GetMutable;
    using Base::SetValue;
    std::pair<COLLISION_CELL&, bool> GetMutable(int64_t x, int64_t y, int64_t z) override {
        if (IndexInBounds(x, y, z)) { components_valid_ = false; return std::pair<COLLISION_CELL&, bool>(data_[(size_t)GetDataIndex(x, y, z)], true); }
        return std::pair<COLLISION_CELL&, bool>(oob_value_, false);
    }
    std::pair<COLLISION_CELL&, bool> GetMutable(const GRID_INDEX& i) override { return GetMutable(i.x, i.y, i.z); }
    bool SetValue(int64_t x, int64_t y, int64_t z, const COLLISION_CELL& v) override {
        if (!IndexInBounds(x, y, z)) return false;
        components_valid_ = false;
        data_[(size_t)GetDataIndex(x, y, z)] = v;
        return true;
    }
    bool SetValue(const GRID_INDEX& i, const COLLISION_CELL& v) override { return SetValue(i.x, i.y, i.z, v); }

    // The hot path.  Same signature and meaning as the reference (collision_map.hpp:680-712): filled iff
    // occupancy > 0.5, or == 0.5 when unknown_is_filled.  Without a virtual border the 8-byte cells go to the GPU as
    // they are and the rule is applied there (no std::function call per voxel).
    std::pair<SignedDistanceField, std::pair<double, double>> ExtractSignedDistanceField(const float oob_value,
                                                                                         const bool unknown_is_filled,
                                                                                         const bool add_virtual_border) const {
        const Eigen::Vector3d cs = GetCellSizes();
        if ((cs.x() != cs.y()) || (cs.x() != cs.z())) throw std::invalid_argument("Grid must have uniform resolution");
        if (!add_virtual_border) {
            SignedDistanceField sdf(GetOriginTransform(), frame_, cs.x(), num_x_cells_, num_y_cells_, num_z_cells_, oob_value);
            double mx = 0.0, mn = 0.0;
            sdf_b200_ctx* ctx = sdf_generation::detail::Context();
            sdf_generation::detail::Check(
                sdf_b200_build_from_cells(ctx, data_.data(), unknown_is_filled ? 1 : 0, num_x_cells_, num_y_cells_,
                                          num_z_cells_, cs.x(), sdf.GetMutableRawData().data(), &mx, &mn), ctx);
            return std::make_pair(std::move(sdf), std::make_pair(mx, mn));     // downloaded in place, moved out
        }
        const std::function<bool(const GRID_INDEX&)> is_filled_fn = [&](const GRID_INDEX& index) {
            const auto query = GetImmutable(index);
            if (!query.second) throw std::runtime_error("index out of grid bounds");
            return (query.first.occupancy > 0.5f) || (unknown_is_filled && (query.first.occupancy == 0.5f));
        };
        return sdf_generation::ExtractSignedDistanceField(*this, is_filled_fn, oob_value, frame_, add_virtual_border);
    }
};

}  // namespace sdf_tools
