// sdf_tools_b200/csrc/line_kernels.cu -- the strided envelope passes (y, then x with the fused distance epilogue) with one
// THREAD per line and the hull stacks in an L2-resident global scratch field: see line_hull.cuh for the algorithm and
// the layout.  Replaces the per-axis loops of sdf_generation::BuildDistanceField
// (/root/reference/include/sdf_tools/sdf_generation.hpp:95-207) and, in the last pass, the combine loop (:246-268).
//
// Threads are numbered along the contiguous direction (lid = outer * ncols + col), so a warp reads, pushes and writes
// 128-byte rows.  Out of place: in -> out (the stack must not be clobbered by the results), every voxel is read once and
// written once; stack traffic stays in L2 except for the long-lived part of a hull.
#include "edt_kernels.cuh"
#include "line_hull.cuh"
#include <stdlib.h>

namespace sdfb200 {

// Resident threads per SM the register allocation must allow: 1536 (48 warps, 40 registers, a few spilled words) measured
// best at 1024^3 -- 1280 (unconstrained, 48-56 registers): y / x pass 4.55 / 8.10 ms; 1536: 4.36 / 7.77; 1792-2048 (32
// registers, heavy spills): 5.04 / 9.65.  The kernel is latency bound on its L2-resident stacks, so warps beat registers.
#ifndef LH_THREADS_PER_SM
#define LH_THREADS_PER_SM 1536
#endif

template <bool WIDE, bool FINAL, int TPB>
__global__ void __launch_bounds__(TPB, LH_THREADS_PER_SM / TPB) line_hull_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                        typename LineEntry<WIDE>::type* __restrict__ stk, uint32_t* __restrict__ pol,
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
        typename LineEntry<WIDE>::type* my_stk = stk + (lid >> 5) * lh_block_entries(n) + lh_lane_offset(lane);
        line_hull_pass<WIDE>(in + base, line_stride, n, my_stk, pol + outer * (int64_t)nw * ncols + col, ncols,
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

static int line_knob(const char* name, int dflt)
{
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

template <bool WIDE, bool FINAL, int TPB>
static cudaError_t launch_line_tpb(const int32_t* in, void* out, void* stk, uint32_t* pol, int n, int64_t line_stride, int64_t ncols,
                                   int64_t nouter, int64_t outer_stride, int64_t out_stride, int64_t out_outer_stride, double res, int* ext, const float* lut,
                                   int lut_n, cudaStream_t st)
{
    const int64_t nlines = ncols * nouter;
    const int64_t blocks = (nlines + TPB - 1) / TPB;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    auto kern = line_hull_kernel<WIDE, FINAL, TPB>;
    static int smem = -1;          // SDF_B200_LINE_SMEM: unused dynamic shared memory per block, to cap the resident warps (experiments)
    if (smem < 0) smem = line_knob("SDF_B200_LINE_SMEM", 0);
    if (smem > 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    static int carve = -2;         // SDF_B200_LINE_CARVE: preferred shared-memory carve-out in percent (experiments)
    if (carve == -2) carve = line_knob("SDF_B200_LINE_CARVE", -1);
    if (carve >= 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)blocks, TPB, smem, st>>>(in, (int32_t*)out, (typename LineEntry<WIDE>::type*)stk, pol, n, line_stride, ncols, nlines,
                                             outer_stride, out_stride, out_outer_stride, res, ext, lut, lut_n);
    return cudaGetLastError();
}

template <bool WIDE, bool FINAL>
static cudaError_t launch_line_cfg(const int32_t* in, void* out, void* stk, uint32_t* pol, int n, int64_t line_stride, int64_t ncols,
                                   int64_t nouter, int64_t outer_stride, int64_t out_stride, int64_t out_outer_stride, double res, int* ext, const float* lut,
                                   int lut_n, cudaStream_t st)
{
    if (ncols * nouter <= 0 || n <= 0) return cudaSuccess;
    static int tpb = -1;
    if (tpb < 0) tpb = line_knob("SDF_B200_LINE_TPB", 128);
    if (tpb == 64) return launch_line_tpb<WIDE, FINAL, 64>(in, out, stk, pol, n, line_stride, ncols, nouter, outer_stride, out_stride, out_outer_stride, res, ext, lut, lut_n, st);
    if (tpb == 256) return launch_line_tpb<WIDE, FINAL, 256>(in, out, stk, pol, n, line_stride, ncols, nouter, outer_stride, out_stride, out_outer_stride, res, ext, lut, lut_n, st);
    return launch_line_tpb<WIDE, FINAL, 128>(in, out, stk, pol, n, line_stride, ncols, nouter, outer_stride, out_stride, out_outer_stride, res, ext, lut, lut_n, st);
}

cudaError_t launch_line_hull(const int32_t* in, void* out, void* d_stack, uint32_t* d_pol, int n, int64_t line_stride, int64_t ncols,
                             int64_t nouter, int64_t outer_stride, int64_t out_stride, int64_t out_outer_stride, bool wide_entries,
                             bool final_pass, double res, int* d_ext, const float* d_lut, int lut_n, cudaStream_t st)
{
    if (!wide_entries && n > 1024) return cudaErrorInvalidValue;
#define LH_ARGS in, out, d_stack, d_pol, n, line_stride, ncols, nouter, outer_stride, out_stride, out_outer_stride, res, d_ext, d_lut, lut_n, st
    if (wide_entries) return final_pass ? launch_line_cfg<true, true>(LH_ARGS) : launch_line_cfg<true, false>(LH_ARGS);
    return final_pass ? launch_line_cfg<false, true>(LH_ARGS) : launch_line_cfg<false, false>(LH_ARGS);
#undef LH_ARGS
}

}  // namespace sdfb200
