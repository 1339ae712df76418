// sdf_tools_b200/csrc/edt_kernels.cuh -- launchers of the sm_100a kernels (see edt_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SDF_MAX_PEERS 32   // ranks of one slab build whose gather buffers a single launch can store to

namespace sdfb200 {

// K0: COLLISION_CELL[n] (8 B each) -> uint8 mask, rule of collision_map.hpp:686-709
cudaError_t launch_pack_cells(const void* d_cells, int64_t n, int unknown_is_filled, uint8_t* d_mask, cudaStream_t st);

// K1: binary pass along the contiguous z axis.  mask[nlines][nz] -> signed int32 d^2 along z
//     d_carries: NULL, or per line {distance from z=0 / z=nz-1 to the nearest other-polarity voxel in the slabs
//     below / above} (int2, 40000 = none) for slab builds.
//     mask_bits: which bits of a mask byte mean FILLED (0xff: any non-zero byte; 1 << q: query q of launch_pack_tagged)
cudaError_t launch_binary_z(const uint8_t* d_mask, int32_t* d_out, int nz, int64_t nlines, const void* d_carries,
                            cudaStream_t st, unsigned mask_bits = 0xffu);

// TAGGED_OBJECT_COLLISION_CELL[n] (16 B each) -> one byte per voxel, bit q = "filled and matches query q" (nq <= 8;
// kinds: 0 object_id == ids[q], 1 any cell, 2 object_id > 0): tagged_object_collision_map.hpp:733-778, :813-856
cudaError_t launch_pack_tagged(const void* d_cells16, int64_t n, int unknown_is_filled, const uint32_t* ids,
                               const uint32_t* kinds, int nq, uint8_t* d_mask, cudaStream_t st);

// Slab (multi-GPU) support for the binary pass: per-line 4-byte summary of one slab, and the resolution of one
// rank's carries from the all-gathered summaries of every rank (d_all[r*nlines + line], d_widths[r]).
cudaError_t launch_slab_summary(const uint8_t* d_mask, int nz, int64_t nlines, uint32_t* d_summary, cudaStream_t st);
// summary + exchange in one kernel: peer_gather[k] = gather buffer of rank k (device pointers valid on this device)
cudaError_t launch_slab_summary_scatter(const uint8_t* d_mask, int nz, int64_t nlines, void* const* peer_gather, int nranks,
                                        int rank, cudaStream_t st);
cudaError_t launch_slab_carry(const uint32_t* d_all, const int* d_widths, int nranks, int rank, int64_t nlines,
                              void* d_carries, cudaStream_t st);

// K2/K3: in-place envelope pass along a strided axis: nouter blocks of n rows x ncols contiguous columns.
//        final_pass: also apply (float)(sqrt((double)d2)*res) with the polarity sign and reduce the
//        per-polarity maxima of d2 into d_ext[0..1].
cudaError_t launch_band_strided(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                int64_t outer_stride, bool final_pass, double res, int* d_ext, long long* d_dbg,
                                const float* d_lut, int lut_n, cudaStream_t st);

// K2/K3, line-per-thread variant (line_kernels.cu): OUT OF PLACE.  Line (outer, col), col < ncols contiguous:
//        input position y at in[outer*outer_stride + col + y*line_stride], result at out[outer*out_outer_stride + col +
//        y*out_stride] (so a pass can also re-order the axes); hull stacks in d_stack: line_stack_entries(n, lines)
//        entries of line_entry_bytes(entry_format) bytes, entry_format = line_entry_format(n, largest height the pass can
//        meet) (dead after the launch: every launch may use the same scratch); one polarity bit per voxel in d_pol
//        (nouter x ceil(n/32) x ncols words).
int line_entry_format(int64_t n, int64_t hmax);     // 0: {pos:10,h:22}  1: {pos:11,h:21}  2: 64-bit entries
int line_entry_bytes(int entry_format);
inline int64_t line_stack_entries(int64_t n, int64_t lines) { return ((lines + 31) / 32) * 32 * ((n + 15) / 16 * 16); }
cudaError_t launch_line_hull(const int32_t* in, void* out, void* d_stack, uint32_t* d_pol, int n, int64_t line_stride, int64_t ncols,
                             int64_t nouter, int64_t outer_stride, int64_t out_stride, int64_t out_outer_stride, int entry_format,
                             bool final_pass, double res, int* d_ext, const float* d_lut, int lut_n, cudaStream_t st);

// d_lut[m] = (float)(sqrt((double)m) * res), m < n: the final pass's exact distance table for one resolution
cudaError_t launch_sqrt_lut(float* d_lut, int n, double res, cudaStream_t st);   // d_dbg: optional 2 x 1024 x 8 phase clocks (diagnostics), or NULL

// add_virtual_border (reference sdf_generation.hpp:287-419): pad the inner mask with a one-cell shell on every axis longer than
// one cell (ox/oy/oz = 1 or 0), and the per-voxel select between the two padded fields
cudaError_t launch_border_pad(const uint8_t* d_inner, int64_t nx, int64_t ny, int64_t nz, int ox, int oy, int oz, int shell,
                              uint8_t* d_padded, cudaStream_t st);
cudaError_t launch_border_select(const float* d_a, const float* d_b, int64_t nx, int64_t ny, int64_t nz, int ox, int oy, int oz,
                                 float* d_out, cudaStream_t st);

// Grid-aligned gradient field of an f32 SDF (reference sdf.hpp:341-358 + :432-526): d_grad = 3 values per voxel, f64 or f32
cudaError_t launch_gradient(const float* d_sdf, int64_t nx, int64_t ny, int64_t nz, double res, int edges, int out_f64,
                            void* d_grad, cudaStream_t st);

}  // namespace sdfb200
