// sdf_tools_b200/csrc/edt_kernels.cuh -- launchers of the sm_100a kernels (see edt_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdfb200 {

// K0: COLLISION_CELL[n] (8 B each) -> uint8 mask, rule of collision_map.hpp:686-709
cudaError_t launch_pack_cells(const void* d_cells, int64_t n, int unknown_is_filled, uint8_t* d_mask, cudaStream_t st);

// K1: binary pass along the contiguous z axis.  mask[nlines][nz] -> signed int32 d^2 along z
//     d_carries: NULL, or per line {distance from z=0 / z=nz-1 to the nearest other-polarity voxel in the slabs
//     below / above} (int2, 40000 = none) for slab builds.
cudaError_t launch_binary_z(const uint8_t* d_mask, int32_t* d_out, int nz, int64_t nlines, const void* d_carries,
                            cudaStream_t st);

// Slab (multi-GPU) support for the binary pass: per-line 4-byte summary of one slab, and the resolution of one
// rank's carries from the all-gathered summaries of every rank (d_all[r*nlines + line], d_widths[r]).
cudaError_t launch_slab_summary(const uint8_t* d_mask, int nz, int64_t nlines, uint32_t* d_summary, cudaStream_t st);
cudaError_t launch_slab_carry(const uint32_t* d_all, const int* d_widths, int nranks, int rank, int64_t nlines,
                              void* d_carries, cudaStream_t st);

// K2/K3: in-place envelope pass along a strided axis: nouter blocks of n rows x ncols contiguous columns.
//        final_pass: also apply (float)(sqrt((double)d2)*res) with the polarity sign and reduce the
//        per-polarity maxima of d2 into d_ext[0..1].
cudaError_t launch_band_strided(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                int64_t outer_stride, bool final_pass, double res, int* d_ext, long long* d_dbg,
                                const float* d_lut, int lut_n, cudaStream_t st);

// d_lut[m] = (float)(sqrt((double)m) * res), m < n: the final pass's exact distance table for one resolution
cudaError_t launch_sqrt_lut(float* d_lut, int n, double res, cudaStream_t st);   // d_dbg: optional 2 x 1024 x 8 phase clocks (diagnostics), or NULL

// out[i] = the final pass's distance value for signed d2 = m0 + i (diagnostics / tests)
cudaError_t launch_sqrt_check(int32_t m0, int32_t count, double res, float* d_out, cudaStream_t st);

}  // namespace sdfb200
