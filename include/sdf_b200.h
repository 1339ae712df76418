/*
 * include/sdf_b200.h -- C ABI of libsdf_b200.so, the B200 (sm_100a) build of sdf_tools' signed-distance-field
 * hot path.  Plain pointers and sizes only; no C++ / torch / Eigen types cross this boundary.
 *
 * The reference has no FFI layer: the hot path is header-only C++ reached through
 *   sdf_tools::CollisionMapGrid::ExtractSignedDistanceField        include/sdf_tools/collision_map.hpp:680-712
 *     -> sdf_generation::ExtractSignedDistanceField (grid overload) include/sdf_tools/sdf_generation.hpp:273-286
 *     -> sdf_generation::ExtractSignedDistanceField<T> (dims)        include/sdf_tools/sdf_generation.hpp:209-271
 *          = classify (:219-240) + 2 x BuildDistanceField (:95-207, called at :242-243) + combine (:246-268)
 * and, from Python, through pysdf_tools.CollisionMapGrid.ExtractSignedDistanceField (src/sdf_tools/bindings.cpp:81).
 * Each entry point below names the reference code it stands in for.  The C++ face that mirrors the reference
 * signatures (include/sdf_tools/ headers) and the Python face (sdf_tools_b200.pysdf_tools) are thin callers
 * of this ABI; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - memory order of every grid: idx = x*(ny*nz) + y*nz + z, z contiguous (VoxelGrid: utils_3d.py:71-74,
 *     sdf.cpp:241-245).  mask: 1 byte per voxel, non-zero = FILLED.
 *   - result: f32 signed distance, negative inside filled voxels, exactly the value the reference stores at
 *     sdf_generation.hpp:265  (float)(sqrt(d2_filled)*res - sqrt(d2_free)*res), with d2 the exact squared EDT.
 *   - all functions return SDF_B200_OK (0) or a negative error code; sdf_b200_last_error() gives the text.
 *     No exceptions cross the ABI.  A context is not thread-safe (the reference is single-threaded too);
 *     use one context per host thread.  The caller owns every buffer it passes in.
 *   - there is NO CPU fallback: without a CUDA device sdf_b200_create() fails.
 */
#ifndef SDF_B200_H
#define SDF_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDF_B200_OK 0
#define SDF_B200_ERR_CUDA (-1)        /* a CUDA runtime call failed (text in last_error)            */
#define SDF_B200_ERR_ARG (-2)         /* null pointer / negative size / bad argument                 */
#define SDF_B200_ERR_UNSUPPORTED (-3) /* an axis longer than SDF_B200_MAX_AXIS                       */
#define SDF_B200_ERR_NOMEM (-4)       /* device or pinned-host allocation failed                     */

#define SDF_B200_MAX_AXIS 2048        /* 64 bands of 32 rows per line in the envelope kernel        */
#define SDF_B200_INF 0x40000000       /* |signed d2| sentinel: no voxel of the other polarity exists */

typedef struct sdf_b200_ctx sdf_b200_ctx;

/* Library / device ------------------------------------------------------------------------------------- */
const char* sdf_b200_version(void);
int sdf_b200_device_count(void);

/* One context = one device + one stream + its scratch.  `device` is a CUDA ordinal. */
int sdf_b200_create(sdf_b200_ctx** out, int device);
void sdf_b200_destroy(sdf_b200_ctx* ctx);
const char* sdf_b200_last_error(const sdf_b200_ctx* ctx); /* ctx may be NULL: error of the last create() */

/* One context over SEVERAL devices of this process (one stream per device).  The host entry points
 * (sdf_b200_build_from_mask / _from_cells) then cut the grid into z slabs, one per device, exchange the per-line slab
 * summaries between the devices (stores into peer memory over NVLink from the summary kernel itself; NCCL all-gather
 * when SDF_B200_EXCHANGE=nccl or without peer access; NCCL is bound at run time with dlopen("libnccl.so.2")) and run
 * the three passes on every device concurrently -- same result, same extrema as on one device.  This is what lets
 * CollisionMapGrid::ExtractSignedDistanceField (collision_map.hpp:680-712) scale without Python.  The device-buffer
 * entry points of a multi-device context act on devices[0] only.  ndev == 1 is the same as sdf_b200_create. */
int sdf_b200_create_multi(sdf_b200_ctx** out, const int* devices, int ndev);
int sdf_b200_num_devices(const sdf_b200_ctx* ctx);          /* 1 for ordinary contexts */
const char* sdf_b200_exchange_kind(const sdf_b200_ctx* ctx); /* "none" | "p2p" | "nccl" | "copy" */

/* Stream for all launches and copies: use_own_stream != 0 selects the context's own non-blocking stream (the
 * default after create); otherwise `cuda_stream` is used as given (e.g. torch's current stream; NULL = the CUDA
 * default stream). */
int sdf_b200_set_stream(sdf_b200_ctx* ctx, void* cuda_stream, int use_own_stream);
int sdf_b200_synchronize(sdf_b200_ctx* ctx);

/* Kernels of the two strided envelope passes (the per-axis loops of BuildDistanceField, sdf_generation.hpp:95-207; no
 * counterpart in the reference -- results are identical in every mode):
 *   0 auto  : per pass, the line kernel when its launch has several hundred thousand lines of >= 512 voxels whose hull
 *             entries fit 32 bits, else the banded kernel (default; the environment variable
 *             SDF_B200_HULL=band|line|mixed|auto changes the default of contexts that never call this);
 *   1 band  : in-place banded kernel, hull stacks in shared memory, no scratch;
 *   2 line  : one thread per line for both passes, hull stacks in a device scratch (8-12 extra bytes per voxel);
 *   3 mixed : line-per-thread y pass, banded x pass (4-12 extra bytes per voxel).
 * Modes that need scratch fall back to band when it cannot be allocated. */
int sdf_b200_set_envelope_kernel(sdf_b200_ctx* ctx, int mode);

/* Pinned host buffers for the host-side entry points (optional; any host pointer works, pinned is faster). */
int sdf_b200_host_alloc(void** out, size_t bytes);
int sdf_b200_host_free(void* p);

/* Whole path, HOST buffers ----------------------------------------------------------------------------- */
/* Replaces sdf_generation::ExtractSignedDistanceField<T> (sdf_generation.hpp:209-271) once the caller's
 * is_filled_fn has been evaluated into `mask` (x->y->z order, :221-227).  out_sdf: nx*ny*nz floats.
 * out_max / out_min: the extrema pair returned at :269 (max first).  Copies H2D, runs, copies D2H, syncs. */
int sdf_b200_build_from_mask(sdf_b200_ctx* ctx, const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz,
                             double resolution, float* out_sdf, double* out_max, double* out_min);

/* Same, straight from CollisionMapGrid storage: cells = nx*ny*nz x {float occupancy; uint32 component}
 * (collision_map.hpp:20-32).  Applies the is_filled rule of collision_map.hpp:686-709 on the device:
 * occupancy > 0.5, or == 0.5 when unknown_is_filled. */
int sdf_b200_build_from_cells(sdf_b200_ctx* ctx, const void* cells, int unknown_is_filled, int64_t nx, int64_t ny,
                              int64_t nz, double resolution, float* out_sdf, double* out_max, double* out_min);

/* Whole path, DEVICE buffers (asynchronous on the context's stream) ---------------------------------- */
/* d_mask: nx*ny*nz bytes; d_sdf: nx*ny*nz floats (also used as the in-place int32 work field: no other
 * scratch is needed).  Extrema are left on the device; fetch them with sdf_b200_get_extrema(). */
int sdf_b200_build_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz,
                          double resolution, float* d_sdf);
/* Synchronises the stream and returns the (max, min) of the last build on this context (:246-264, :269). */
int sdf_b200_get_extrema(sdf_b200_ctx* ctx, double* out_max, double* out_min);

/* The is_filled rule alone (collision_map.hpp:686-709): d_cells -> d_mask, n cells, asynchronous. */
int sdf_b200_pack_cells_device(sdf_b200_ctx* ctx, const void* d_cells, int64_t n, int unknown_is_filled,
                               uint8_t* d_mask);

/* The two BuildDistanceField results (sdf_generation.hpp:242-243) as ONE signed int32 field, before the
 * sqrt/resolution combine:  +d2(nearest filled) at free voxels, -d2(nearest free) at filled voxels,
 * +-SDF_B200_INF where the other polarity does not exist.  For bit-exact integer parity tests. */
int sdf_b200_signed_d2_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz,
                              int32_t* d_out);

/* Same with HOST buffers (synchronous): out = nx*ny*nz int32.  Backs the BuildDistanceField compatibility wrapper of
 * include/sdf_tools/sdf_generation.hpp: with seeds as the FILLED voxels, max(out, 0) is distance_square (:95-207). */
int sdf_b200_signed_d2_from_mask(sdf_b200_ctx* ctx, const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int32_t* out);

/* Slab builds: the grid is cut along z into `nranks` slabs, one context (GPU) each -------------------------- */
/* No reference counterpart (the reference is single-process); this is how the one BuildDistanceField pass that
 * crosses the cut (the binary pass along z) gets its context.  Protocol per build, on every rank:
 *   1. sdf_b200_slab_summary_device : d_summary[nx*ny] = 4-byte per-line summary of the local slab
 *   2. all-gather the summaries over the ranks (NCCL / torch.distributed; layout [rank][nx*ny])
 *   3. sdf_b200_slab_build_device   : resolves this rank's carries and runs the three passes on the local slab
 *   4. all-reduce(max) the two integers of sdf_b200_get_extrema_raw and convert with sdf_b200_extrema_from_raw.
 * slab_widths: host array, thickness (cells along z) of every slab ("rank" = slab index; a process may own several
 * slabs and build them one after the other -- keep_extrema != 0 then accumulates the per-polarity maxima onto those of
 * the previous build on this context instead of restarting them).  d_mask / d_sdf: local slab (nx, ny, nz_local),
 * z contiguous. */
int sdf_b200_slab_summary_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz_local,
                                 uint32_t* d_summary);
int sdf_b200_slab_build_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz_local,
                               double resolution, const uint32_t* d_all_summaries, const int32_t* slab_widths,
                               int nranks, int rank, int keep_extrema, float* d_sdf);
/* Steps 1 + 2 in ONE kernel, for ranks that can store to each other's memory (cudaDeviceEnablePeerAccess in one
 * process; CUDA IPC or symmetric memory between processes): the summary word of every line of the local slab is stored
 * straight into slot `rank` of the gather buffer of every rank -- peer_gather[k] (k < nranks, host array of device
 * pointers valid on this device) is rank k's buffer of nranks * nx * ny words.  The caller orders the launch before the
 * peers' sdf_b200_slab_build_device (events, or a barrier on the ranks' signal pads). */
int sdf_b200_slab_summary_scatter_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz_local,
                                         void* const* peer_gather, int nranks, int rank);
/* Per-polarity maxima of the exact squared distance of the last build (0 = that polarity does not occur). */
int sdf_b200_get_extrema_raw(sdf_b200_ctx* ctx, int32_t* out_max_free_d2, int32_t* out_max_filled_d2);
/* (max, min) of sdf_generation.hpp:269 from (all-reduced) raw maxima. */
int sdf_b200_extrema_from_raw(int32_t max_free_d2, int32_t max_filled_d2, double resolution, double* out_max,
                              double* out_min);

/* add_virtual_border (reference sdf_generation.hpp:287-419) ------------------------------------- */
/* Every axis longer than one cell grows by a one-cell shell; the free-space values come from a build whose shell is
 * FILLED, the inside values from a build whose shell is FREE, selected per real voxel as
 * a >= 0 ? a : (b <= -0 ? b : 0) (:381-409); extrema = (max of the first build, min of the second) (:412-413).  All of it
 * on the device: one upload of the inner mask(s), two builds on the padded grid, one download of nx*ny*nz floats.
 * mask_free / mask_filled: the is_filled predicate as evaluated during the reference's first / second extraction (it
 * calls the predicate once per voxel in each, :381-382); the same pointer twice for a pure predicate. */
int sdf_b200_build_with_border(sdf_b200_ctx* ctx, const uint8_t* mask_free, const uint8_t* mask_filled, int64_t nx, int64_t ny,
                               int64_t nz, double resolution, float* out_sdf, double* out_max, double* out_min);

/* Batched object SDFs --------------------------------------------------------------------------- */
/* TaggedObjectCollisionMapGrid callers (reference tagged_object_collision_map.hpp): MakeObjectSDFs (:875-915) runs one
 * full ExtractSignedDistanceField(objects_to_use = {id}) (:813-856) per object id, ExtractFreeAndNamedObjects... (:730-811)
 * two more.  Here the tagged cells -- nx*ny*nz x {float occupancy; u32 component; u32 object_id; u32 convex_segment},
 * 16 bytes (:22-43) -- are uploaded once, one pass over them answers up to 8 queries (one mask bit each) and the builds
 * run back to back while the previous result downloads.  Query k fills out_sdfs[k*N .. (k+1)*N) (N = nx*ny*nz) and
 * out_max[k] / out_min[k] with exactly what the reference's call for that predicate returns (exact EDT):
 *   kind 0  filled cells whose object_id == object_id        (objects_to_use = {object_id})
 *   kind 1  every filled cell                                  (objects_to_use empty; the "free" field of :733-753)
 *   kind 2  filled cells of named objects, object_id > 0       (the "named objects" field of :755-778)
 * filled = occupancy > 0.5, or == 0.5 when unknown_is_filled. */
typedef struct { uint32_t object_id; uint32_t kind; } sdf_b200_object_query;
int sdf_b200_build_object_sdfs(sdf_b200_ctx* ctx, const void* tagged_cells, int64_t nx, int64_t ny, int64_t nz,
                               double resolution, int unknown_is_filled, const sdf_b200_object_query* queries, int nqueries,
                               float* out_sdfs, double* out_max, double* out_min);

/* Gradient field ------------------------------------------------------------------------------ */
/* What every Python caller of the reference runs right after the build (utils_3d.py:77-90): SignedDistanceField::
 * GetFullGradient (sdf.hpp:341-358) over GetGridAlignedGradient (sdf.hpp:432-526) -- central differences, one-sided
 * at the faces when enable_edge_gradients, 0 along axes of one cell -- as one kernel.  d_grad / out_grad: 3 values per
 * voxel (d/dx, d/dy, d/dz; voxel order as the field, component fastest), f64 (out_f64 != 0: bit-identical to the
 * reference's doubles) or f32 (the f64 value rounded).  Voxels for which the reference returns an empty vector (faces
 * with enable_edge_gradients == 0) get NaN.  The rotation by the grid's origin transform that GetGradient applies
 * (sdf.hpp:402-430) is left to the caller (identity for axis-aligned grids, as in utils_2d / utils_3d).
 * _host: sdf == NULL takes the field that the last host build on this context left on the device (same grid). */
int sdf_b200_gradient_device(sdf_b200_ctx* ctx, const float* d_sdf, int64_t nx, int64_t ny, int64_t nz, double resolution,
                             int enable_edge_gradients, int out_f64, void* d_grad);
int sdf_b200_gradient_host(sdf_b200_ctx* ctx, const float* sdf, int64_t nx, int64_t ny, int64_t nz, double resolution,
                           int enable_edge_gradients, int out_f64, void* out_grad);

/* Device time of the last build on this context in milliseconds (CUDA events around the launch sequence);
 * per-kernel times in out_ms[0..3] = {pass 1 (z, binary), pass 2 (y), pass 3 (x + combine), total}. */
int sdf_b200_last_timing(sdf_b200_ctx* ctx, float* out_ms4);
/* Number of kernels the last build launched (for bench.py's gpu_launches). */
int sdf_b200_last_launches(const sdf_b200_ctx* ctx);
/* Envelope kernels the last build on this context used: 1 band, 2 line, 3 mixed (sdf_b200_set_envelope_kernel). */
int sdf_b200_last_envelope_kernel(const sdf_b200_ctx* ctx);

/* Diagnostics: when enabled, thread 0 of the first 1024 CTAs of each envelope pass stores clock64() at its phase
 * boundaries (start, loaded, built, barrier, merged, queried) into a device buffer; `out` (2*1024*8 values, pass 2
 * then pass 3) receives them.  enable = 0 frees the buffer.  Not for production use. */
int sdf_b200_debug_phase_clocks(sdf_b200_ctx* ctx, int enable, long long* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* SDF_B200_H */
