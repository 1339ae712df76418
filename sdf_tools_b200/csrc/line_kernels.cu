// sdf_tools_b200/csrc/line_kernels.cu -- the strided envelope passes (y, then x with the fused distance epilogue) with one
// THREAD per line and the hull stacks in an L2-resident global scratch field: see line_hull.cuh for the algorithm and
// the layout.  Replaces the per-axis loops of sdf_generation::BuildDistanceField
// (/root/reference/include/sdf_tools/sdf_generation.hpp:95-207) and, in the last pass, the combine loop (:246-268).
//
// Threads are numbered along the contiguous direction (lid = outer * ncols + col), so a warp reads and writes 128-byte
// rows of the field.  Out of place (in -> out: the axes may be re-ordered on the way), every voxel is read once and
// written once; the stack traffic stays in L1 / L2 except for the long-lived part of a hull.
#include "edt_kernels.cuh"
#include "line_hull.cuh"

namespace sdfb200 {

// Resident threads per SM the register allocation must allow: 1536 (48 warps, 40 registers, a few spilled words) measured
// best at 1024^3 -- 1280 (unconstrained, 48-56 registers): y / x pass 4.55 / 8.10 ms; 1536: 4.36 / 7.77; 1792-2048 (32
// registers, heavy spills): 5.04 / 9.65.  The kernel is latency bound on its L2-resident stacks, so warps beat registers.
#ifndef LH_THREADS_PER_SM
#define LH_THREADS_PER_SM 1536
#endif

template <int FMT, bool FINAL, int TPB>
__global__ void __launch_bounds__(TPB, LH_THREADS_PER_SM / TPB) line_hull_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                        typename LineEntry<FMT>::type* __restrict__ stk, uint32_t* __restrict__ pol,
                                                        int n, int64_t line_stride, int64_t ncols, int64_t nlines, int64_t outer_stride,
                                                        int64_t out_stride, int64_t out_outer_stride,
                                                        double res, int* __restrict__ ext, const float* __restrict__ lut, int lut_n)
{
    const int64_t lid = (int64_t)blockIdx.x * TPB + threadIdx.x;
    const bool ok = lid < nlines;
    int32_t maxpos = 0, maxneg = 0;
    if (ok) {
        const int64_t outer = lid / ncols, col = lid - outer * ncols;
        const int64_t base = outer * outer_stride + col;
        const int nw = (n + 31) >> 5;
        int32_t* dst = out + outer * out_outer_stride + col;
        // final pass: the table load of one position is issued a whole position before its result is stored
        float fpend = 0.f;
        bool negpend = false, first = true;
        // the warp's block of the stack scratch, then this lane's slot 0 inside it (line_hull.cuh, "Stack layout")
        const int lane = threadIdx.x & 31;
        typename LineEntry<FMT>::type* my_stk = stk + (lid >> 5) * lh_block_entries(n) + lh_lane_offset(lane);
        line_hull_pass<FMT>(in + base, line_stride, n, my_stk, pol + outer * (int64_t)nw * ncols + col, ncols,
            [&](int32_t m, bool neg) {
                if (FINAL) {
                    if (neg) maxneg = max(maxneg, m); else maxpos = max(maxpos, m);
                    if (!first) { __stcs(reinterpret_cast<float*>(dst), negpend ? -fpend : fpend); dst += out_stride; }
                    first = false;
                    // (float)(sqrt((double)d2) * res), the value the combine loop stores (sdf_generation.hpp:254-265)
                    if (m < lut_n) fpend = __ldg(lut + m);
                    else fpend = (m >= SDF_INF) ? __int_as_float(0x7f800000) : (float)(sqrt((double)m) * res);
                    negpend = neg;
                } else {
                    __stcs(dst, neg ? -m : m);
                    dst += out_stride;
                }
            });
        if (FINAL && !first) __stcs(reinterpret_cast<float*>(dst), negpend ? -fpend : fpend);
    }
    if (FINAL) {
        maxpos = __reduce_max_sync(0xffffffffu, maxpos);
        maxneg = __reduce_max_sync(0xffffffffu, maxneg);
        if ((threadIdx.x & 31) == 0) {
            if (maxpos > 0) atomicMax(ext + 0, maxpos);
            if (maxneg > 0) atomicMax(ext + 1, maxneg);
        }
    }
}

template <int FMT, bool FINAL>
static cudaError_t launch_line_cfg(const int32_t* in, void* out, void* stk, uint32_t* pol, int n, int64_t line_stride, int64_t ncols,
                                   int64_t nouter, int64_t outer_stride, int64_t out_stride, int64_t out_outer_stride, double res, int* ext, const float* lut,
                                   int lut_n, cudaStream_t st)
{
    constexpr int TPB = 128;
    const int64_t nlines = ncols * nouter;
    if (nlines <= 0 || n <= 0) return cudaSuccess;
    const int64_t blocks = (nlines + TPB - 1) / TPB;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    line_hull_kernel<FMT, FINAL, TPB><<<(unsigned)blocks, TPB, 0, st>>>(in, (int32_t*)out, (typename LineEntry<FMT>::type*)stk, pol, n, line_stride, ncols,
                                                                      nlines, outer_stride, out_stride, out_outer_stride, res, ext, lut, lut_n);
    return cudaGetLastError();
}

cudaError_t launch_line_hull(const int32_t* in, void* out, void* d_stack, uint32_t* d_pol, int n, int64_t line_stride, int64_t ncols,
                             int64_t nouter, int64_t outer_stride, int64_t out_stride, int64_t out_outer_stride, int entry_format,
                             bool final_pass, double res, int* d_ext, const float* d_lut, int lut_n, cudaStream_t st)
{
    if ((entry_format == LH_FMT_P10 && n > 1024) || (entry_format == LH_FMT_P11 && n > 2048) || entry_format < 0 || entry_format > LH_FMT_WIDE)
        return cudaErrorInvalidValue;
#define LH_ARGS in, out, d_stack, d_pol, n, line_stride, ncols, nouter, outer_stride, out_stride, out_outer_stride, res, d_ext, d_lut, lut_n, st
    if (entry_format == LH_FMT_WIDE) return final_pass ? launch_line_cfg<LH_FMT_WIDE, true>(LH_ARGS) : launch_line_cfg<LH_FMT_WIDE, false>(LH_ARGS);
    if (entry_format == LH_FMT_P11) return final_pass ? launch_line_cfg<LH_FMT_P11, true>(LH_ARGS) : launch_line_cfg<LH_FMT_P11, false>(LH_ARGS);
    return final_pass ? launch_line_cfg<LH_FMT_P10, true>(LH_ARGS) : launch_line_cfg<LH_FMT_P10, false>(LH_ARGS);
#undef LH_ARGS
}

int line_entry_format(int64_t n, int64_t hmax) { return lh_entry_format(n, hmax); }
int line_entry_bytes(int entry_format) { return entry_format == LH_FMT_WIDE ? 8 : 4; }

}  // namespace sdfb200
