// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the reference's OWN text (sdf_generation.hpp:19-271), which build_ref.sh extracts
// into oracle/_ref/ref_sdf_generation.inc at build time.  Nothing of the reference is stored in this repo.
#include "ref_shim.hpp"
#include <chrono>
#include <cstring>

namespace sdf_generation {
#include "_ref/ref_sdf_generation.inc"
}

namespace {
struct Dummy {};
}

extern "C" {

// Whole path: classify + 2x BuildDistanceField + combine (sdf_generation.hpp:209-271).
// mask: 1 = filled, reference order.  Returns wall milliseconds of the ExtractSignedDistanceField call, <0 on error.
double ref_build_sdf(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, double res, float oob,
                     float* out_sdf, double* out_max, double* out_min)
{
    try {
        const std::function<bool(const VoxelGrid::GRID_INDEX&)> fn = [&](const VoxelGrid::GRID_INDEX& i) {
            return mask[(size_t)((i.x * ny + i.y) * nz + i.z)] != 0;
        };
        const auto t0 = std::chrono::steady_clock::now();
        auto result = sdf_generation::ExtractSignedDistanceField<Dummy>(Eigen::Isometry3d(), res, nx, ny, nz, fn, oob, "ref");
        const auto t1 = std::chrono::steady_clock::now();
        const std::vector<float>& d = result.first.GetImmutableRawData();
        if (out_sdf) std::memcpy(out_sdf, d.data(), d.size() * sizeof(float));
        if (out_max) *out_max = result.second.first;
        if (out_min) *out_min = result.second.second;
        return std::chrono::duration<double, std::milli>(t1 - t0).count();
    } catch (...) { return -1.0; }
}

// One BuildDistanceField (sdf_generation.hpp:95-207): seeds = voxels where (mask!=0) == want_filled, in x->y->z scan
// order (as :221-240 builds them), or reversed when reverse_seeds!=0 (to expose the order dependence).
// d2_out receives distance_square (f64, +inf where unreached).
double ref_build_d2(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int want_filled, int reverse_seeds,
                    double* d2_out)
{
    try {
        std::vector<VoxelGrid::GRID_INDEX> pts;
        for (int64_t x = 0; x < nx; x++) for (int64_t y = 0; y < ny; y++) for (int64_t z = 0; z < nz; z++)
            if ((mask[(size_t)((x * ny + y) * nz + z)] != 0) == (want_filled != 0)) pts.emplace_back(x, y, z);
        if (reverse_seeds) { std::vector<VoxelGrid::GRID_INDEX> r(pts.rbegin(), pts.rend()); pts.swap(r); }
        const auto t0 = std::chrono::steady_clock::now();
        const sdf_generation::DistanceField df = sdf_generation::BuildDistanceField(Eigen::Isometry3d(), 1.0, nx, ny, nz, pts);
        const auto t1 = std::chrono::steady_clock::now();
        const std::vector<sdf_generation::bucket_cell>& d = df.GetImmutableRawData();
        for (size_t i = 0; i < d.size(); i++) d2_out[i] = d[i].distance_square;
        return std::chrono::duration<double, std::milli>(t1 - t0).count();
    } catch (...) { return -1.0; }
}

int ref_sizeof_bucket_cell() { return (int)sizeof(sdf_generation::bucket_cell); }

}  // extern "C"
