// include/sdf_tools/sdf.hpp -- sdf_tools::SignedDistanceField, the OUTPUT type of the SDF-build hot path
// (reference: include/sdf_tools/sdf.hpp:20-81 type/ctors, :93-127 frame/lock, :163-331 lock-guarded mutation,
// :337-358 GetFullGradient, :402-526 grid-aligned gradients).  Build subset only: the queries that run on a
// finished SDF (EstimateDistance, projection, extrema maps, serialisation, RViz export) are out of scope
//  and are not provided.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../arc_utilities/voxel_grid.hpp"

namespace sdf_tools {
using VoxelGrid::GRID_INDEX;

class SignedDistanceField : public VoxelGrid::VoxelGrid<float> {
protected:
    std::string frame_;
    bool locked_ = false;
    using Base = ::VoxelGrid::VoxelGrid<float>;

public:
    typedef std::shared_ptr<SignedDistanceField> Ptr;
    typedef std::shared_ptr<const SignedDistanceField> ConstPtr;

    SignedDistanceField() : Base(), frame_("") {}
    SignedDistanceField(const Eigen::Isometry3d& origin, const std::string& frame, double resolution,
                        int64_t x_cells, int64_t y_cells, int64_t z_cells, float oob_value)
        : Base(origin, resolution, x_cells, y_cells, z_cells, oob_value), frame_(frame) {}
    SignedDistanceField(const std::string& frame, double resolution, int64_t x_cells, int64_t y_cells, int64_t z_cells,
                        float oob_value)
        : Base(resolution, x_cells, y_cells, z_cells, oob_value), frame_(frame) {}
    SignedDistanceField(const Eigen::Isometry3d& origin, const std::string& frame, double resolution,
                        double x_size, double y_size, double z_size, float oob_value)
        : Base(origin, resolution, x_size, y_size, z_size, oob_value), frame_(frame) {}
    SignedDistanceField(const std::string& frame, double resolution, double x_size, double y_size, double z_size,
                        float oob_value)
        : Base(resolution, x_size, y_size, z_size, oob_value), frame_(frame) {}

    double GetResolution() const { return GetCellSizes().x(); }
    std::string GetFrame() const { return frame_; }
    void SetFrame(const std::string& f) { frame_ = f; }
    bool IsLocked() const { return locked_; }
    void Lock() { locked_ = true; }
    void Unlock() { locked_ = false; }

    // Mutation is refused while locked (reference sdf.hpp:173,188,239,255).
    using Base::GetMutable;
    using Base::SetValue;
    std::pair<float&, bool> GetMutable(int64_t x, int64_t y, int64_t z) override {
        if (IndexInBounds(x, y, z) && !locked_) return std::pair<float&, bool>(data_[(size_t)GetDataIndex(x, y, z)], true);
        return std::pair<float&, bool>(oob_value_, false);
    }
    std::pair<float&, bool> GetMutable(const GRID_INDEX& i) override { return GetMutable(i.x, i.y, i.z); }
    bool SetValue(int64_t x, int64_t y, int64_t z, const float& v) override {
        if (!IndexInBounds(x, y, z) || locked_) return false;
        data_[(size_t)GetDataIndex(x, y, z)] = v;
        return true;
    }
    bool SetValue(const GRID_INDEX& i, const float& v) override { return SetValue(i.x, i.y, i.z, v); }

    // Adopt a finished field (nx*ny*nz floats in storage order) without a per-voxel SetValue loop.
    bool AdoptData(std::vector<float>&& values) {
        if ((int64_t)values.size() != num_x_cells_ * num_y_cells_ * num_z_cells_ || locked_) return false;
        data_ = std::move(values);
        return true;
    }

    // ---- gradients (what the Python callers invoke right after the build: utils_3d.py:77-90) ----------------
    // Central differences in grid axes; one-sided at the faces when enable_edge_gradients (reference sdf.hpp:432-526).
    std::vector<double> GetGridAlignedGradient(int64_t x, int64_t y, int64_t z, bool enable_edge_gradients = false) const {
        if (!IndexInBounds(x, y, z)) return std::vector<double>();
        const bool interior = x > 0 && y > 0 && z > 0 && x < num_x_cells_ - 1 && y < num_y_cells_ - 1 && z < num_z_cells_ - 1;
        if (!interior && !enable_edge_gradients) return std::vector<double>();
        const double res = GetResolution();
        auto at = [&](int64_t a, int64_t b, int64_t c) { return data_[(size_t)GetDataIndex(a, b, c)]; };
        // Bit-identical to the reference: in the interior the two neighbours are subtracted IN FLOAT and the difference
        // is scaled in double (sdf.hpp:445-453); on the faces both values are widened to double first (:486-518).
        auto axis = [&](int64_t lo, int64_t hi, float vlo, float vhi) {
            const double inc = (double)(hi - lo) * res;
            return inc > 0.0 ? ((double)vhi - (double)vlo) * (1.0 / inc) : 0.0;
        };
        if (interior) {
            const double inv = 1.0 / (2.0 * res);
            return std::vector<double>{(double)(at(x + 1, y, z) - at(x - 1, y, z)) * inv,
                                       (double)(at(x, y + 1, z) - at(x, y - 1, z)) * inv,
                                       (double)(at(x, y, z + 1) - at(x, y, z - 1)) * inv};
        }
        const int64_t lx = std::max<int64_t>(0, x - 1), hx = std::min(num_x_cells_ - 1, x + 1);
        const int64_t ly = std::max<int64_t>(0, y - 1), hy = std::min(num_y_cells_ - 1, y + 1);
        const int64_t lz = std::max<int64_t>(0, z - 1), hz = std::min(num_z_cells_ - 1, z + 1);
        return std::vector<double>{axis(lx, hx, at(lx, y, z), at(hx, y, z)), axis(ly, hy, at(x, ly, z), at(x, hy, z)),
                                   axis(lz, hz, at(x, y, lz), at(x, y, hz))};
    }
    // Gradient in the frame of the grid's origin transform (rotation applied, reference sdf.hpp:402-430).
    std::vector<double> GetGradient(int64_t x, int64_t y, int64_t z, bool enable_edge_gradients = false) const {
        const std::vector<double> g = GetGridAlignedGradient(x, y, z, enable_edge_gradients);
        if (g.size() != 3) return std::vector<double>();
#if SDF_B200_HAVE_EIGEN
        const Eigen::Vector3d r = origin_transform_.rotation() * Eigen::Vector3d(g[0], g[1], g[2]);
#else
        const Eigen::Vector3d r = origin_transform_.rotate(Eigen::Vector3d(g[0], g[1], g[2]));
#endif
        return std::vector<double>{r.x(), r.y(), r.z()};
    }
    std::vector<double> GetGradient(const GRID_INDEX& i, bool enable_edge_gradients = false) const {
        return GetGradient(i.x, i.y, i.z, enable_edge_gradients);
    }

    using GradientFunction = std::function<std::vector<double>(int64_t, int64_t, int64_t, bool)>;
    // One gradient_function call per voxel in x->y->z order (reference sdf.hpp:341-358).
    ::VoxelGrid::VoxelGrid<std::vector<double>> GetFullGradient(const GradientFunction& gradient_function,
                                                              bool enable_edge_gradients = false) const {
        ::VoxelGrid::VoxelGrid<std::vector<double>> grid(origin_transform_, GetResolution(), num_x_cells_, num_y_cells_,
                                                       num_z_cells_, std::vector<double>(3, (double)oob_value_));
        for (int64_t x = 0; x < num_x_cells_; ++x)
            for (int64_t y = 0; y < num_y_cells_; ++y)
                for (int64_t z = 0; z < num_z_cells_; ++z)
                    grid.SetValue(x, y, z, gradient_function(x, y, z, enable_edge_gradients));
        return grid;
    }
};

}  // namespace sdf_tools
