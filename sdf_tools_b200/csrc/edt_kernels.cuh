// sdf_tools_b200/csrc/edt_kernels.cuh -- launchers of the sm_100a kernels (see edt_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdfb200 {

// K0: COLLISION_CELL[n] (8 B each) -> uint8 mask, rule of collision_map.hpp:686-709
cudaError_t launch_pack_cells(const void* d_cells, int64_t n, int unknown_is_filled, uint8_t* d_mask, cudaStream_t st);

// K1: binary pass along x.  mask[nx][ncols] -> signed int32 d^2 along x (ncols = ny*nz)
cudaError_t launch_binary_x(const uint8_t* d_mask, int32_t* d_out, int nx, int64_t ncols, cudaStream_t st);

// K2: in-place envelope pass along a strided axis: nouter blocks of n rows x ncols contiguous columns
cudaError_t launch_hull_strided(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                int64_t outer_stride, cudaStream_t st);

// K3: in-place envelope pass along the contiguous axis (nlines lines of n).  final_pass: also apply
//     (float)(sqrt((double)d2)*res) with the polarity sign and reduce the per-polarity maxima into d_ext[0..1].
cudaError_t launch_hull_contig(int32_t* g, int n, int64_t nlines, bool final_pass, double res, int* d_ext,
                               cudaStream_t st);

}  // namespace sdfb200
