// sdf_tools_b200/csrc/edt_kernels.cu -- sm_100a kernels of the signed-distance-field build and their launchers.
//
// Replaces, per call of sdf_generation::ExtractSignedDistanceField<T>
// (/root/reference/include/sdf_tools/sdf_generation.hpp:209-271):
//   K0 pack_cells      is_filled rule                         collision_map.hpp:686-709 (+ classify loop :219-240)
//   K1 binary_x        first third of BuildDistanceField x2   sdf_generation.hpp:95-207, on the binary mask
//   K2 hull_strided    middle third (y axis, strided lines)   same
//   K3 hull_contig     last third (z axis, contiguous lines)  same, + the combine loop :246-268 fused in
// All three passes work IN PLACE on one 4-byte-per-voxel field (signed int32 d^2, finally f32 distance), so a
// build touches  1 B (mask read) + 4 B + 4+4 B + 4+4 B = 21 B per voxel of HBM traffic and needs no scratch.
//
// Grid layout: idx = x*(ny*nz) + y*nz + z.  Pass order x -> y -> z: the binary pass runs along the slowest
// axis (the one slabs are cut along for multi-GPU builds, where only per-column carries cross the cut), the
// two envelope passes are local.
#include "edt_kernels.cuh"
#include "line_hull.cuh"

#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>

namespace sdfb200 {

// =============================================================================================== K0 pack
// COLLISION_CELL {float occupancy; uint32 component} -> 1 byte, filled iff occupancy > 0.5 or (== 0.5 && unknown)
__global__ void __launch_bounds__(256) pack_cells_kernel(const float2* __restrict__ cells, int64_t n,
                                                         int unknown_is_filled, uint8_t* __restrict__ mask)
{
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n && ((reinterpret_cast<uintptr_t>(mask) & 3) == 0)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(cells + i4));
        const float4 b = __ldg(reinterpret_cast<const float4*>(cells + i4 + 2));
        const float o[4] = {a.x, a.z, b.x, b.z};
        uchar4 r;
        unsigned char* rp = reinterpret_cast<unsigned char*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) rp[k] = (o[k] > 0.5f) || (unknown_is_filled && o[k] == 0.5f);
        *reinterpret_cast<uchar4*>(mask + i4) = r;
    } else {
        for (int64_t i = i4; i < n && i < i4 + 4; ++i) {
            const float o = cells[i].x;
            mask[i] = (o > 0.5f) || (unknown_is_filled && o == 0.5f);
        }
    }
}

// ============================================================================================ K1 binary x
// One thread per (y,z) column.  The column's occupancy bits are gathered into registers (coalesced byte rows),
// then  s(x) = +-min(x-a+1, e-x)^2  with [a,e) the run of equal polarity around x: the squared distance to the
// nearest voxel of the OTHER polarity along x (SDF_INF if the run touches both grid faces).
template <int NW>
__global__ void __launch_bounds__(128) binary_x_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ out,
                                                       int nx, int64_t ncols)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const uint8_t* mp = mask + c;
    uint32_t bits[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        uint32_t b = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const int x = w * 32 + j;
            if (x < nx) b |= (uint32_t)(__ldg(mp + (int64_t)x * ncols) != 0) << j;
        }
        bits[w] = b;
    }
    uint32_t pc[NW];
    int na[NW];
    {
        uint32_t carry = bits[0] & 1u;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t b = bits[w];
            uint32_t p = b ^ ((b << 1) | carry);
            carry = b >> 31;
            const int valid = nx - w * 32;
            if (valid < 32) p &= (valid > 0) ? ((1u << valid) - 1u) : 0u;
            pc[w] = p;
        }
        int nxt = nx;
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) {
            na[w] = nxt;
            if (pc[w]) nxt = w * 32 + __ffs((int)pc[w]) - 1;
        }
    }
    int a = -1;
    int32_t* op = out + c;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t p = pc[w], b = bits[w];
        const int nw = na[w];
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const int x = w * 32 + j;
            if (x < nx) {
                if ((p >> j) & 1u) a = x;
                const uint32_t rest = (j < 31) ? (p >> (j + 1)) : 0u;
                const int e = rest ? (x + __ffs((int)rest)) : nw;
                int32_t best = SDF_INF;
                if (a >= 1) { const int d = x - a + 1; best = d * d; }
                if (e < nx) { const int d = e - x; const int32_t z = d * d; best = z < best ? z : best; }
                op[(int64_t)x * ncols] = ((b >> j) & 1u) ? -best : best;
            }
        }
    }
}

// ========================================================================================= K2 hull strided
// Lines run along an axis whose elements are `line_stride` apart; neighbouring lines (columns) are contiguous.
// A CTA stages W columns x n rows in shared memory (each global row segment is W*4 contiguous bytes), one warp
// then owns one column (see line_hull.cuh), and the tile is written back in place.
template <int LOGC, typename MaskT, int W>
__global__ void __launch_bounds__(W * 32) hull_strided_kernel(int32_t* __restrict__ g, int n, int64_t line_stride,
                                                              int64_t ncols, int64_t outer_stride)
{
    using LH = LineHull<LOGC, MaskT>;
    using L = Line<LOGC>;
    constexpr int CHUNK = L::CHUNK;
    constexpr int RPI = 32 / W;                 // rows covered by one warp-wide access
    constexpr int LP = L::WORDS + RPI;          // line pitch (words): staggers the banks of neighbouring columns
    constexpr int U = 8;                        // loads in flight per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* vals = reinterpret_cast<int32_t*>(smem_raw);
    MaskT* masks = reinterpret_cast<MaskT*>(smem_raw + (((size_t)W * LP * 4 + 15) & ~(size_t)15));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c0 = (int64_t)blockIdx.x * W;
    int32_t* gt = g + (int64_t)blockIdx.y * outer_stride + c0;
    const int cl = lane & (W - 1), rl = lane / W;
    const bool col_ok = (c0 + cl) < ncols;
    int32_t* vcol = vals + cl * LP;

    for (int r0 = warp * RPI + rl; r0 < n; r0 += W * RPI * U) {
        int32_t tmp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + u * W * RPI;
            tmp[u] = (r < n && col_ok) ? gt[(int64_t)r * line_stride + cl] : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + u * W * RPI;
            if (r < n && col_ok) vcol[L::idx(r)] = tmp[u];
        }
    }
    __syncthreads();

    if (c0 + warp < ncols) {
        int32_t* myval = vals + warp * LP;
        MaskT* Mw = masks + warp * 64;
        MaskT* PCw = Mw + 32;
        int32_t out[CHUNK];
        const MaskT pw = LH::lane_build(myval, Mw, PCw, n, lane);
        __syncwarp();
#pragma unroll 1
        for (int lv = 0; lv < 5; ++lv) {
            LH::lane_merge(myval, Mw, PCw, n, lane, lv);
            __syncwarp();
        }
        LH::lane_query(myval, Mw, PCw, n, lane, pw, out);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
            const int q = (lane << LOGC) + j;
            if (q < n) myval[L::idx(q)] = out[j];
        }
    }
    __syncthreads();

    for (int r0 = warp * RPI + rl; r0 < n; r0 += W * RPI * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + u * W * RPI;
            if (r < n && col_ok) gt[(int64_t)r * line_stride + cl] = vcol[L::idx(r)];
        }
    }
}

// ========================================================================================== K3 hull contig
// Lines are contiguous (the z axis).  One warp per line, WPB warps per CTA, each with a private line buffer.
// FINAL fuses the reference's combine loop (sdf_generation.hpp:246-268): the signed d^2 becomes
// (float)(sqrt((double)d2) * res) with the sign of the polarity, and the per-polarity maxima are reduced for
// the (max, min) pair returned at :269.
__device__ __forceinline__ float sdf_from_signed_d2(int32_t s, double res)
{
    const int32_t m = s < 0 ? -s : s;
    const float f = (m >= SDF_INF) ? __int_as_float(0x7f800000) : (float)(sqrt((double)m) * res);
    return s < 0 ? -f : f;
}

template <int LOGC, typename MaskT, bool FINAL, int WPB>
__global__ void __launch_bounds__(WPB * 32, (LOGC >= 6) ? 2 : 3) hull_contig_kernel(int32_t* __restrict__ g, int n, int64_t nlines,
                                                               double res, int* __restrict__ ext)
{
    using LH = LineHull<LOGC, MaskT>;
    using L = Line<LOGC>;
    constexpr int CHUNK = L::CHUNK;
    constexpr int U = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* vals = reinterpret_cast<int32_t*>(smem_raw);
    MaskT* masks = reinterpret_cast<MaskT*>(smem_raw + (size_t)WPB * L::WORDS * 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* myval = vals + warp * L::WORDS;
    MaskT* Mw = masks + warp * 64;
    MaskT* PCw = Mw + 32;
    int32_t maxpos = 0, maxneg = 0;

    for (int64_t line = (int64_t)blockIdx.x * WPB + warp; line < nlines; line += (int64_t)gridDim.x * WPB) {
        int32_t* gl = g + line * n;
        for (int r0 = lane; r0 < n; r0 += 32 * U) {
            int32_t tmp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int r = r0 + u * 32; tmp[u] = (r < n) ? gl[r] : 0; }
#pragma unroll
            for (int u = 0; u < U; ++u) { const int r = r0 + u * 32; if (r < n) myval[L::idx(r)] = tmp[u]; }
        }
        __syncwarp();
        int32_t out[CHUNK];
        const MaskT pw = LH::lane_build(myval, Mw, PCw, n, lane);
        __syncwarp();
#pragma unroll 1
        for (int lv = 0; lv < 5; ++lv) {
            LH::lane_merge(myval, Mw, PCw, n, lane, lv);
            __syncwarp();
        }
        LH::lane_query(myval, Mw, PCw, n, lane, pw, out);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
            const int q = (lane << LOGC) + j;
            if (q < n) {
                const int32_t s = out[j];
                if (FINAL) {
                    if (s > maxpos) maxpos = s;
                    if (-s > maxneg) maxneg = -s;
                    myval[L::idx(q)] = __float_as_int(sdf_from_signed_d2(s, res));
                } else {
                    myval[L::idx(q)] = s;
                }
            }
        }
        __syncwarp();
        for (int r0 = lane; r0 < n; r0 += 32 * U) {
#pragma unroll
            for (int u = 0; u < U; ++u) { const int r = r0 + u * 32; if (r < n) gl[r] = myval[L::idx(r)]; }
        }
        __syncwarp();
    }
    if (FINAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            maxpos = max(maxpos, __shfl_xor_sync(0xffffffffu, maxpos, o));
            maxneg = max(maxneg, __shfl_xor_sync(0xffffffffu, maxneg, o));
        }
        if (lane == 0) {
            if (maxpos > 0) atomicMax(ext + 0, maxpos);
            if (maxneg > 0) atomicMax(ext + 1, maxneg);
        }
    }
}

// ================================================================================================ launchers
static inline int logc_for(int n) { int lc = 0; while ((32 << lc) < n) ++lc; return lc; }

cudaError_t launch_pack_cells(const void* d_cells, int64_t n, int unknown_is_filled, uint8_t* d_mask, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int64_t threads = (n + 3) / 4;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    pack_cells_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float2*>(d_cells), n, unknown_is_filled, d_mask);
    return cudaGetLastError();
}

template <int NW>
static cudaError_t launch_binary_x_nw(const uint8_t* m, int32_t* out, int nx, int64_t ncols, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((ncols + 127) / 128);
    binary_x_kernel<NW><<<blocks, 128, 0, st>>>(m, out, nx, ncols);
    return cudaGetLastError();
}

cudaError_t launch_binary_x(const uint8_t* d_mask, int32_t* d_out, int nx, int64_t ncols, cudaStream_t st)
{
    const int nw = (nx + 31) / 32;
    if (nw <= 1) return launch_binary_x_nw<1>(d_mask, d_out, nx, ncols, st);
    if (nw <= 2) return launch_binary_x_nw<2>(d_mask, d_out, nx, ncols, st);
    if (nw <= 4) return launch_binary_x_nw<4>(d_mask, d_out, nx, ncols, st);
    if (nw <= 8) return launch_binary_x_nw<8>(d_mask, d_out, nx, ncols, st);
    if (nw <= 16) return launch_binary_x_nw<16>(d_mask, d_out, nx, ncols, st);
    if (nw <= 32) return launch_binary_x_nw<32>(d_mask, d_out, nx, ncols, st);
    if (nw <= 64) return launch_binary_x_nw<64>(d_mask, d_out, nx, ncols, st);
    return cudaErrorInvalidValue;
}

template <int LOGC>
static cudaError_t launch_strided_lc(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                     int64_t outer_stride, cudaStream_t st)
{
    using MaskT = typename std::conditional<(LOGC >= 6), uint64_t, uint32_t>::type;
    constexpr int W = (LOGC >= 5) ? 16 : 32;
    constexpr int LP = Line<LOGC>::WORDS + 32 / W;
    const size_t smem = (((size_t)W * LP * 4 + 15) & ~(size_t)15) + (size_t)W * 64 * sizeof(MaskT);
    auto kern = hull_strided_kernel<LOGC, MaskT, W>;
    if (smem > 48 * 1024) {   // per-device attribute; cheap enough to set on every launch
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)((ncols + W - 1) / W), (unsigned)nouter);
    kern<<<grid, W * 32, smem, st>>>(g, n, line_stride, ncols, outer_stride);
    return cudaGetLastError();
}

cudaError_t launch_hull_strided(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                int64_t outer_stride, cudaStream_t st)
{
    switch (logc_for(n)) {
        case 0: return launch_strided_lc<0>(g, n, line_stride, ncols, nouter, outer_stride, st);
        case 1: return launch_strided_lc<1>(g, n, line_stride, ncols, nouter, outer_stride, st);
        case 2: return launch_strided_lc<2>(g, n, line_stride, ncols, nouter, outer_stride, st);
        case 3: return launch_strided_lc<3>(g, n, line_stride, ncols, nouter, outer_stride, st);
        case 4: return launch_strided_lc<4>(g, n, line_stride, ncols, nouter, outer_stride, st);
        case 5: return launch_strided_lc<5>(g, n, line_stride, ncols, nouter, outer_stride, st);
        case 6: return launch_strided_lc<6>(g, n, line_stride, ncols, nouter, outer_stride, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int LOGC, bool FINAL>
static cudaError_t launch_contig_lc(int32_t* g, int n, int64_t nlines, double res, int* ext, cudaStream_t st)
{
    using MaskT = typename std::conditional<(LOGC >= 6), uint64_t, uint32_t>::type;
    constexpr int WPB = 8;
    const size_t smem = (size_t)WPB * Line<LOGC>::WORDS * 4 + (size_t)WPB * 64 * sizeof(MaskT);
    auto kern = hull_contig_kernel<LOGC, MaskT, FINAL, WPB>;
    if (smem > 48 * 1024) {   // per-device attribute; cheap enough to set on every launch
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int64_t blocks = (nlines + WPB - 1) / WPB;
    kern<<<(unsigned)blocks, WPB * 32, smem, st>>>(g, n, nlines, res, ext);
    return cudaGetLastError();
}

template <bool FINAL>
static cudaError_t launch_contig_f(int32_t* g, int n, int64_t nlines, double res, int* ext, cudaStream_t st)
{
    switch (logc_for(n)) {
        case 0: return launch_contig_lc<0, FINAL>(g, n, nlines, res, ext, st);
        case 1: return launch_contig_lc<1, FINAL>(g, n, nlines, res, ext, st);
        case 2: return launch_contig_lc<2, FINAL>(g, n, nlines, res, ext, st);
        case 3: return launch_contig_lc<3, FINAL>(g, n, nlines, res, ext, st);
        case 4: return launch_contig_lc<4, FINAL>(g, n, nlines, res, ext, st);
        case 5: return launch_contig_lc<5, FINAL>(g, n, nlines, res, ext, st);
        case 6: return launch_contig_lc<6, FINAL>(g, n, nlines, res, ext, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_hull_contig(int32_t* g, int n, int64_t nlines, bool final_pass, double res, int* d_ext,
                               cudaStream_t st)
{
    return final_pass ? launch_contig_f<true>(g, n, nlines, res, d_ext, st)
                      : launch_contig_f<false>(g, n, nlines, res, d_ext, st);
}

}  // namespace sdfb200
