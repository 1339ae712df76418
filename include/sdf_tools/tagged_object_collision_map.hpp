// include/sdf_tools/tagged_object_collision_map.hpp -- the SDF-producing part of sdf_tools::TaggedObjectCollisionMapGrid
// (reference include/sdf_tools/tagged_object_collision_map.hpp: cell :22-43, ExtractFreeAndNamedObjectsSignedDistanceField
// :730-811, ExtractSignedDistanceField(objects_to_use) :813-856, MakeObjectSDFs :875-897, MakeAllObjectSDFs :899-915),
// executed on a B200 through sdf_b200_build_object_sdfs: the 16-byte cells are uploaded once and the per-object builds
// are batched.  Topology, convex segmentation, serialisation and RViz export of the reference class
// are out of scope and not provided.
#pragma once
#include <cstdint>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "sdf_generation.hpp"

namespace sdf_tools {

struct TAGGED_OBJECT_COLLISION_CELL {
    float occupancy;
    uint32_t component;
    uint32_t object_id;
    uint32_t convex_segment;
    TAGGED_OBJECT_COLLISION_CELL() : occupancy(0.0f), component(0u), object_id(0u), convex_segment(0u) {}
    TAGGED_OBJECT_COLLISION_CELL(float in_occupancy, uint32_t in_object_id)
        : occupancy(in_occupancy), component(0u), object_id(in_object_id), convex_segment(0u) {}
    TAGGED_OBJECT_COLLISION_CELL(float in_occupancy, uint32_t in_object_id, uint32_t in_component, uint32_t in_convex_segment)
        : occupancy(in_occupancy), component(in_component), object_id(in_object_id), convex_segment(in_convex_segment) {}
};
static_assert(sizeof(TAGGED_OBJECT_COLLISION_CELL) == 16, "the device kernels read 16-byte tagged cells");

class TaggedObjectCollisionMapGrid : public VoxelGrid::VoxelGrid<TAGGED_OBJECT_COLLISION_CELL> {
protected:
    using Base = ::VoxelGrid::VoxelGrid<TAGGED_OBJECT_COLLISION_CELL>;
    std::string frame_;

    using Result = std::pair<SignedDistanceField, std::pair<double, double>>;
    // One batched call: query k -> (SignedDistanceField, (max, min)).  The fields are downloaded into the objects' own storage.
    std::vector<Result> RunQueries(const std::vector<sdf_b200_object_query>& queries, const float oob_value,
                                   const bool unknown_is_filled) const {
        const Eigen::Vector3d cs = GetCellSizes();
        if ((cs.x() != cs.y()) || (cs.x() != cs.z())) throw std::invalid_argument("Grid must have uniform resolution");
        const int64_t N = num_x_cells_ * num_y_cells_ * num_z_cells_;
        std::vector<float> all((size_t)N * queries.size());
        std::vector<double> mx(queries.size()), mn(queries.size());
        sdf_b200_ctx* ctx = sdf_generation::detail::Context();
        sdf_generation::detail::Check(
            sdf_b200_build_object_sdfs(ctx, data_.data(), num_x_cells_, num_y_cells_, num_z_cells_, cs.x(), unknown_is_filled ? 1 : 0,
                                       queries.data(), (int)queries.size(), all.data(), mx.data(), mn.data()), ctx);
        std::vector<Result> out;
        out.reserve(queries.size());
        for (size_t k = 0; k < queries.size(); ++k) {
            SignedDistanceField sdf(GetOriginTransform(), frame_, cs.x(), num_x_cells_, num_y_cells_, num_z_cells_, oob_value);
            sdf.AdoptData(std::vector<float>(all.begin() + (size_t)N * k, all.begin() + (size_t)N * (k + 1)));
            out.emplace_back(std::move(sdf), std::make_pair(mx[k], mn[k]));
        }
        return out;
    }

public:
    TaggedObjectCollisionMapGrid() : Base(), frame_("") {}
    TaggedObjectCollisionMapGrid(const Eigen::Isometry3d& origin, const std::string& frame, double resolution, int64_t x_cells,
                                 int64_t y_cells, int64_t z_cells, const TAGGED_OBJECT_COLLISION_CELL& default_value)
        : Base(origin, resolution, x_cells, y_cells, z_cells, default_value), frame_(frame) {}
    TaggedObjectCollisionMapGrid(const Eigen::Isometry3d& origin, const std::string& frame, double resolution, int64_t x_cells,
                                 int64_t y_cells, int64_t z_cells, const TAGGED_OBJECT_COLLISION_CELL& default_value,
                                 const TAGGED_OBJECT_COLLISION_CELL& oob_value)
        : Base(origin, resolution, x_cells, y_cells, z_cells, default_value, oob_value), frame_(frame) {}

    double GetResolution() const { return GetCellSizes().x(); }
    std::string GetFrame() const { return frame_; }
    void SetFrame(const std::string& f) { frame_ = f; }

    // reference :813-856
    Result ExtractSignedDistanceField(const float oob_value, const std::vector<uint32_t>& objects_to_use,
                                      const bool unknown_is_filled, const bool add_virtual_border) const {
        if (!add_virtual_border && objects_to_use.size() <= 1) {
            sdf_b200_object_query q;
            q.object_id = objects_to_use.empty() ? 0u : objects_to_use[0];
            q.kind = objects_to_use.empty() ? 1u : 0u;
            return std::move(RunQueries({q}, oob_value, unknown_is_filled)[0]);
        }
        const std::set<uint32_t> use(objects_to_use.begin(), objects_to_use.end());
        const std::function<bool(const GRID_INDEX&)> is_filled_fn = [&](const GRID_INDEX& index) {
            const auto query = GetImmutable(index);
            if (!query.second) throw std::runtime_error("index out of grid bounds");
            if (use.empty() || use.count(query.first.object_id))
                return (query.first.occupancy > 0.5f) || (unknown_is_filled && (query.first.occupancy == 0.5f));
            return false;
        };
        return sdf_generation::ExtractSignedDistanceField(*this, is_filled_fn, oob_value, frame_, add_virtual_border);
    }

    // reference :730-811: the free-space field where it is >= 0, else the named-objects field where it is <= -0, else 0
    Result ExtractFreeAndNamedObjectsSignedDistanceField(const float oob_value, const bool unknown_is_filled) const {
        std::vector<Result> r = RunQueries({sdf_b200_object_query{0u, 1u}, sdf_b200_object_query{0u, 2u}}, oob_value, unknown_is_filled);
        SignedDistanceField combined = std::move(r[0].first);
        std::vector<float>& c = combined.GetMutableRawData();
        const std::vector<float>& named = r[1].first.GetImmutableRawData();
        for (size_t i = 0; i < c.size(); ++i) {
            const float free_v = c[i], named_v = named[i];
            c[i] = (free_v >= 0.0f) ? free_v : ((named_v <= -0.0f) ? named_v : 0.0f);
        }
        return std::make_pair(std::move(combined), std::make_pair(r[0].second.first, r[1].second.second));
    }

    // reference :875-897 (one ExtractSignedDistanceField per id there; one batched call here)
    std::map<uint32_t, SignedDistanceField> MakeObjectSDFs(const std::vector<uint32_t>& object_ids, const bool unknown_is_filled,
                                                            const bool add_virtual_border) const {
        std::map<uint32_t, SignedDistanceField> per_object_sdfs;
        const float oob = std::numeric_limits<float>::infinity();
        if (add_virtual_border) {
            for (const uint32_t id : object_ids)
                per_object_sdfs[id] = ExtractSignedDistanceField(oob, std::vector<uint32_t>{id}, unknown_is_filled, true).first;
            return per_object_sdfs;
        }
        std::vector<sdf_b200_object_query> queries;
        for (const uint32_t id : object_ids) queries.push_back(sdf_b200_object_query{id, 0u});
        if (queries.empty()) return per_object_sdfs;
        std::vector<Result> r = RunQueries(queries, oob, unknown_is_filled);
        for (size_t k = 0; k < object_ids.size(); ++k) per_object_sdfs[object_ids[k]] = std::move(r[k].first);
        return per_object_sdfs;
    }

    // reference :899-915
    std::map<uint32_t, SignedDistanceField> MakeAllObjectSDFs(const bool unknown_is_filled, const bool add_virtual_border) const {
        std::set<uint32_t> ids;
        for (const TAGGED_OBJECT_COLLISION_CELL& cell : data_)
            if (cell.object_id > 0u) ids.insert(cell.object_id);
        return MakeObjectSDFs(std::vector<uint32_t>(ids.begin(), ids.end()), unknown_is_filled, add_virtual_border);
    }
};

}  // namespace sdf_tools
