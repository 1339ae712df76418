// sdf_tools_b200/csrc/capi.cu -- the C ABI of libsdf_b200.so (include/sdf_b200.h) over the kernel launchers.
// Host orchestration of sdf_generation::ExtractSignedDistanceField<T> (sdf_generation.hpp:209-271):
//   mask -> K1 (z, binary) -> K2 (y) -> K3 (x + combine) -> f32 field + (max, min).
#include "../../include/sdf_b200.h"
#include "edt_kernels.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits>
#include <vector>

struct sdf_b200_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int* d_ext = nullptr;              // [0] max +d2 over free voxels, [1] max d2 over filled voxels, [2] sticky: a tile copy timed out
    void* d_carries = nullptr; size_t carries_bytes = 0;   // slab builds: int2 per z line
    int* d_widths = nullptr;
    float* d_lut = nullptr;            // (float)(sqrt((double)m)*res) for m < LUT_N at resolution lut_res
    double lut_res = -1.0;
    long long* d_dbg = nullptr;        // optional phase clocks of the band kernels: [pass 2|3][1024 CTAs][8]
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;  // host entry points: H2D / D2H chunks overlap the kernels of the other chunks
    cudaEvent_t chunk_ev[17] = {};       // [0..7] upload chunks, [8..15] download chunks, [16] stream hand-over
    cudaEvent_t pack_ev[2] = {};         // cells staging halves: the pack kernel that read them has finished
    bool timed = false;
    int launches = 0;
    void* d_in = nullptr;  size_t in_bytes = 0;     // staging for the host entry points
    void* d_mask = nullptr; size_t mask_bytes = 0;
    void* d_out = nullptr; size_t out_bytes = 0;
    void* d_grad = nullptr; size_t grad_bytes = 0;  // gradient staging of the host entry point
    void* d_out2 = nullptr; size_t out2_bytes = 0;  // second result buffer of the batched object builds
    int* d_ext_objs = nullptr; size_t ext_objs_n = 0;
    int last_hull = 1;                              // envelope kernels the last build resolved to (HULL_BAND / LINE / MIXED)
    int hull_mode = -1;                             // -1: SDF_B200_HULL default; else HULL_AUTO / BAND / LINE / MIXED (sdf_b200_set_envelope_kernel)
    // scratch of the line-per-thread envelope passes: the field between the y and x passes, the hull stacks, polarity bits
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    void* d_stack = nullptr; size_t stack_bytes = 0;
    void* d_pol = nullptr; size_t pol_bytes = 0;
    int64_t last_dims[3] = {0, 0, 0};               // grid of the last HOST build whose field is still in d_out
    double res = 1.0;
    bool built = false;
    char err[512] = {0};
    // ---- multi-device contexts (sdf_b200_create_multi): this context is the leader and peers[0]
    std::vector<sdf_b200_ctx*> peers;      // empty for an ordinary context
    bool p2p = false;                      // every pair of devices can store to each other (NVLink / NVSwitch)
    void* d_gather = nullptr; size_t gather_bytes = 0;     // [slab][nx*ny] summaries of every slab of the current build
    cudaEvent_t ev_summary = nullptr;      // this device's summaries have reached every peer
    void* nccl_comm = nullptr;             // ncclComm_t of this device when the exchange goes through NCCL
    int raw_ext[2] = {0, 0};               // extrema of the last multi-device build (already reduced over the devices)
};

static char g_create_err[512] = {0};
static void (*nccl_destroy_comm)(void*) = nullptr;     // set when a multi-device context created NCCL communicators

static int fail(sdf_b200_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess)
{
    char* dst = ctx ? ctx->err : g_create_err;
    if (e != cudaSuccess) snprintf(dst, 512, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else snprintf(dst, 512, "%s", what);
    return code;
}

#define CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(ctx, SDF_B200_ERR_CUDA, what, _e); } while (0)

// ext[2] != 0: an envelope kernel gave up waiting for its TMA tile copy (the result of that build is garbage)
static int check_tile_flag(sdf_b200_ctx* ctx, sdf_b200_ctx* owner, int flag)
{
    if (!flag) return SDF_B200_OK;
    cudaSetDevice(owner->device);
    cudaMemsetAsync(owner->d_ext + 2, 0, sizeof(int), owner->stream);
    return fail(ctx, SDF_B200_ERR_CUDA, "a TMA tile copy of the envelope pass timed out (result discarded); SDF_B200_TMA=0 selects plain loads");
}

static int check_dims(sdf_b200_ctx* ctx, int64_t nx, int64_t ny, int64_t nz)
{
    if (nx < 0 || ny < 0 || nz < 0) return fail(ctx, SDF_B200_ERR_ARG, "negative grid size");
    if (nx > SDF_B200_MAX_AXIS || ny > SDF_B200_MAX_AXIS || nz > SDF_B200_MAX_AXIS)
        return fail(ctx, SDF_B200_ERR_UNSUPPORTED, "grid axis longer than SDF_B200_MAX_AXIS (2048)");
    return SDF_B200_OK;
}

static int grow(sdf_b200_ctx* ctx, void** p, size_t* have, size_t need)
{
    if (need <= *have) return SDF_B200_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *have = 0; }
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, SDF_B200_ERR_NOMEM, "cudaMalloc(staging)", e); }
    *have = need;
    return SDF_B200_OK;
}

extern "C" {

const char* sdf_b200_version(void) { return "sdf_b200 0.1 (sm_100a)"; }

int sdf_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sdf_b200_create(sdf_b200_ctx** out, int device)
{
    sdf_b200_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, SDF_B200_ERR_ARG, "sdf_b200_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(nullptr, SDF_B200_ERR_CUDA, "sdf_b200_create: no CUDA device (this library has no CPU fallback)", e);
    }
    if (device < 0 || device >= n) return fail(nullptr, SDF_B200_ERR_ARG, "sdf_b200_create: bad device ordinal");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, SDF_B200_ERR_CUDA, "cudaSetDevice", e);
    ctx = new sdf_b200_ctx();
    ctx->device = device;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->d_ext, 4 * sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(ctx->d_ext, 0, 4 * sizeof(int))) != cudaSuccess) {
        fail(nullptr, SDF_B200_ERR_CUDA, "sdf_b200_create: stream/alloc", e);
        delete ctx;
        return SDF_B200_ERR_CUDA;
    }
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 17 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->pack_ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        fail(nullptr, SDF_B200_ERR_CUDA, "sdf_b200_create: events / copy stream", e);
        sdf_b200_destroy(ctx);
        return SDF_B200_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return SDF_B200_OK;
}

void sdf_b200_destroy(sdf_b200_ctx* ctx)
{
    if (!ctx) return;
    if (ctx->peers.size() > 1) {                     // a multi-device leader owns its peers (peers[0] is the leader)
        std::vector<sdf_b200_ctx*> ps;
        ps.swap(ctx->peers);
        for (size_t i = 0; i < ps.size(); ++i) {
            if (ps[i]->nccl_comm && nccl_destroy_comm) nccl_destroy_comm(ps[i]->nccl_comm);
            ps[i]->nccl_comm = nullptr;
            if (ps[i] != ctx) sdf_b200_destroy(ps[i]);
        }
    }
    cudaSetDevice(ctx->device);
    if (ctx->ev_summary) cudaEventDestroy(ctx->ev_summary);
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 17; ++i) if (ctx->chunk_ev[i]) cudaEventDestroy(ctx->chunk_ev[i]);
    for (int i = 0; i < 2; ++i) if (ctx->pack_ev[i]) cudaEventDestroy(ctx->pack_ev[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->d_ext) cudaFree(ctx->d_ext);
    if (ctx->d_dbg) cudaFree(ctx->d_dbg);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->d_carries) cudaFree(ctx->d_carries);
    if (ctx->d_widths) cudaFree(ctx->d_widths);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_mask) cudaFree(ctx->d_mask);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_grad) cudaFree(ctx->d_grad);
    if (ctx->d_out2) cudaFree(ctx->d_out2);
    if (ctx->d_ext_objs) cudaFree(ctx->d_ext_objs);
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    if (ctx->d_stack) cudaFree(ctx->d_stack);
    if (ctx->d_pol) cudaFree(ctx->d_pol);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* sdf_b200_last_error(const sdf_b200_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int sdf_b200_set_stream(sdf_b200_ctx* ctx, void* cuda_stream, int use_own_stream)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    ctx->stream = use_own_stream ? ctx->own_stream : (cudaStream_t)cuda_stream;   // NULL = CUDA default stream
    return SDF_B200_OK;
}

int sdf_b200_synchronize(sdf_b200_ctx* ctx)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    return SDF_B200_OK;
}

int sdf_b200_host_alloc(void** out, size_t bytes)
{
    if (!out) return SDF_B200_ERR_ARG;
    *out = nullptr;
    if (bytes == 0) return SDF_B200_OK;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, SDF_B200_ERR_NOMEM, "cudaHostAlloc", e); }
    return SDF_B200_OK;
}

int sdf_b200_host_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return SDF_B200_ERR_CUDA; }
    return SDF_B200_OK;
}

// Which envelope kernels (sdf_b200_set_envelope_kernel; SDF_B200_HULL=auto|band|line|mixed sets the default of new contexts):
//   band  : both strided passes with the in-place banded kernel (stacks in shared memory, no scratch);
//   line  : both passes with the line-per-thread kernel (hull stacks in an L2-resident scratch field; 8-12 extra bytes
//           per voxel); the y pass leaves its result x/y-swapped so that both passes read rows that are nz apart;
//   mixed : line-per-thread y pass, banded x pass;
//   auto  : the line kernel for a pass whose launch has enough lines of >= 512 voxels with 32-bit stack entries -- 2^19
//           for a y pass of up to 1024 voxels per line, 2^20 for longer y lines and for the x pass (measured: with fewer
//           or shorter lines the line kernel cannot fill the GPU with independent work, with 64-bit entries its stack
//           traffic doubles, and deeper hulls need more lines to win; the banded kernel is faster there) -- i.e. line,
//           mixed or band.
// A plan that needs scratch falls back to band when the scratch cannot be allocated.
enum { HULL_AUTO = 0, HULL_BAND = 1, HULL_LINE = 2, HULL_MIXED = 3 };

static int hull_default()
{
    static int knob = -1;
    if (knob < 0) {
        const char* e = getenv("SDF_B200_HULL");
        knob = HULL_AUTO;
        if (e && e[0] == 'b') knob = HULL_BAND;
        else if (e && e[0] == 'l') knob = HULL_LINE;
        else if (e && e[0] == 'm') knob = HULL_MIXED;
    }
    return knob;
}

struct LinePlan {
    int mode = HULL_BAND;                  // resolved: HULL_BAND, HULL_LINE or HULL_MIXED
    int fmt_y = 0, fmt_x = 0;              // stack entry formats of the two passes (line_entry_format)
};

// Decide the kernels for a grid and size the scratch.  nz_total: the whole z extent the binary pass saw (slab builds).
static LinePlan plan_line_passes(sdf_b200_ctx* ctx, int64_t nx, int64_t ny, int64_t nz, int64_t nz_total, int64_t x_per_launch = 0)
{
    LinePlan P;
    int mode = ctx->hull_mode < 0 ? hull_default() : ctx->hull_mode;
    const int64_t hz = (nz_total - 1) * (nz_total - 1), hy = (ny - 1) * (ny - 1);     // largest heights the y / x pass can meet: hz, hz + hy
    P.fmt_y = sdfb200::line_entry_format(ny, hz);
    P.fmt_x = sdfb200::line_entry_format(nx, hz + hy);
    if (mode == HULL_AUTO) {
        const int64_t lines_y = (x_per_launch > 0 ? x_per_launch : nx) * nz, lines_x = ny * nz;
        const bool y_ok = lines_y >= (ny <= 1024 ? (1 << 19) : (1 << 20)) && ny >= 512 && sdfb200::line_entry_bytes(P.fmt_y) == 4;
        const bool x_ok = x_per_launch == 0 && lines_x >= (1 << 20) && nx >= 512 && sdfb200::line_entry_bytes(P.fmt_x) == 4;
        mode = (y_ok && x_ok) ? HULL_LINE : (y_ok ? HULL_MIXED : HULL_BAND);
    }
    if (ny <= 1 && mode == HULL_MIXED) mode = HULL_BAND;        // no y pass
    if (mode == HULL_BAND) return P;
    const int64_t N = nx * ny * nz;
    const int64_t polw_y = nx * ((ny + 31) / 32) * nz, polw_x = ((nx + 31) / 32) * ny * nz;
    const int64_t polw = mode == HULL_LINE ? (polw_y > polw_x ? polw_y : polw_x) : polw_y;
    size_t stack_need = (size_t)sdfb200::line_stack_entries(ny, nx * nz) * sdfb200::line_entry_bytes(P.fmt_y);
    if (mode == HULL_LINE) {
        const size_t sx = (size_t)sdfb200::line_stack_entries(nx, ny * nz) * sdfb200::line_entry_bytes(P.fmt_x);
        stack_need = sx > stack_need ? sx : stack_need;
    }
    char keep[512]; memcpy(keep, ctx->err, sizeof(keep));
    if (grow(ctx, &ctx->d_tmp, &ctx->tmp_bytes, (size_t)N * 4) || grow(ctx, &ctx->d_stack, &ctx->stack_bytes, stack_need) ||
        grow(ctx, &ctx->d_pol, &ctx->pol_bytes, (size_t)polw * 4)) {
        memcpy(ctx->err, keep, sizeof(keep));          // not an error: the in-place kernel needs no scratch
        return P;
    }
    P.mode = mode;
    return P;
}

// mask -> signed int32 d^2 (final_pass = false) or f32 distance (final_pass = true), in d_field
static int run_passes(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz, double res,
                      int32_t* d_field, bool final_pass, const void* d_carries = nullptr, bool keep_extrema = false,
                      unsigned mask_bits = 0xffu, int64_t nz_total = 0)
{
    cudaStream_t st = ctx->stream;
    ctx->launches = 0;
    ctx->timed = false;
    if (!keep_extrema) CK(cudaMemsetAsync(ctx->d_ext, 0, 2 * sizeof(int), st), "memset extrema");
    if (nx * ny * nz == 0) { ctx->built = true; ctx->res = res; return SDF_B200_OK; }
    constexpr int LUT_N = 1 << 20;                 // 4 MB: stays resident in the 126 MB L2
    if (final_pass && (ctx->d_lut == nullptr || ctx->lut_res != res)) {
        if (!ctx->d_lut) CK(cudaMalloc(&ctx->d_lut, (size_t)LUT_N * sizeof(float)), "cudaMalloc(lut)");
        CK(sdfb200::launch_sqrt_lut(ctx->d_lut, LUT_N, res, st), "launch sqrt_lut");
        ctx->lut_res = res;
        ctx->launches++;
    }
    const LinePlan LP = plan_line_passes(ctx, nx, ny, nz, nz_total > 0 ? nz_total : nz);
    ctx->last_hull = LP.mode;
    int32_t* tmp = (int32_t*)ctx->d_tmp;
    CK(cudaEventRecord(ctx->ev[0], st), "event");
    // z pass: into the buffer the next pass reads
    int32_t* after_z = d_field;
    if (LP.mode == HULL_MIXED || (LP.mode == HULL_LINE && ny <= 1)) after_z = tmp;
    CK(sdfb200::launch_binary_z(d_mask, after_z, (int)nz, nx * ny, d_carries, st, mask_bits), "launch binary_z");
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[1], st), "event");
    if (ny > 1) {
        if (LP.mode == HULL_LINE)            // field[x][y][z] -> tmp[y][x][z]
            CK(sdfb200::launch_line_hull(d_field, tmp, ctx->d_stack, (uint32_t*)ctx->d_pol, (int)ny, nz, nz, nx, ny * nz, nx * nz, nz, LP.fmt_y,
                                         false, res, ctx->d_ext, nullptr, 0, st), "launch line_hull(y)");
        else if (LP.mode == HULL_MIXED)      // tmp[x][y][z] -> field[x][y][z]
            CK(sdfb200::launch_line_hull(tmp, d_field, ctx->d_stack, (uint32_t*)ctx->d_pol, (int)ny, nz, nz, nx, ny * nz, nz, ny * nz, LP.fmt_y,
                                         false, res, ctx->d_ext, nullptr, 0, st), "launch line_hull(y)");
        else
            CK(sdfb200::launch_band_strided(d_field, (int)ny, nz, nz, nx, ny * nz, false, res, ctx->d_ext, ctx->d_dbg, nullptr, 0, st),
               "launch band_strided(y)");
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev[2], st), "event");
    if (LP.mode == HULL_LINE)                // tmp[y][x][z] -> field[x][y][z]
        CK(sdfb200::launch_line_hull(tmp, d_field, ctx->d_stack, (uint32_t*)ctx->d_pol, (int)nx, nz, nz, ny, nx * nz, ny * nz, nz, LP.fmt_x,
                                     final_pass, res, ctx->d_ext, final_pass ? ctx->d_lut : nullptr, final_pass ? LUT_N : 0, st),
           "launch line_hull(x)");
    else
        CK(sdfb200::launch_band_strided(d_field, (int)nx, ny * nz, ny * nz, 1, 0, final_pass, res, ctx->d_ext,
                                        ctx->d_dbg ? ctx->d_dbg + 1024 * 8 : nullptr, final_pass ? ctx->d_lut : nullptr, final_pass ? LUT_N : 0, st),
           "launch band_strided(x)");
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[3], st), "event");
    ctx->timed = true;
    ctx->built = final_pass;
    ctx->res = res;
    return SDF_B200_OK;
}

int sdf_b200_set_envelope_kernel(sdf_b200_ctx* ctx, int mode)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    if (mode < HULL_AUTO || mode > HULL_MIXED) return fail(ctx, SDF_B200_ERR_ARG, "envelope kernel: 0 auto, 1 band, 2 line, 3 mixed");
    ctx->hull_mode = mode;
    for (sdf_b200_ctx* p : ctx->peers) p->hull_mode = mode;
    return SDF_B200_OK;
}

int sdf_b200_build_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz,
                          double resolution, float* d_sdf)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz);
    if (rc) return rc;
    if (nx * ny * nz > 0 && (!d_mask || !d_sdf)) return fail(ctx, SDF_B200_ERR_ARG, "null device pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    return run_passes(ctx, d_mask, nx, ny, nz, resolution, reinterpret_cast<int32_t*>(d_sdf), true);
}

int sdf_b200_signed_d2_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz,
                              int32_t* d_out)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz);
    if (rc) return rc;
    if (nx * ny * nz > 0 && (!d_mask || !d_out)) return fail(ctx, SDF_B200_ERR_ARG, "null device pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    return run_passes(ctx, d_mask, nx, ny, nz, 1.0, d_out, false);
}

int sdf_b200_slab_summary_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz_local,
                                 uint32_t* d_summary)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz_local);
    if (rc) return rc;
    if (nx * ny * nz_local == 0) return SDF_B200_OK;
    if (!d_mask || !d_summary) return fail(ctx, SDF_B200_ERR_ARG, "null device pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(sdfb200::launch_slab_summary(d_mask, (int)nz_local, nx * ny, d_summary, ctx->stream), "launch slab_summary");
    return SDF_B200_OK;
}

int sdf_b200_slab_build_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz_local,
                               double resolution, const uint32_t* d_all_summaries, const int32_t* slab_widths,
                               int nranks, int rank, int keep_extrema, float* d_sdf)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz_local);
    if (rc) return rc;
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks || !slab_widths)
        return fail(ctx, SDF_B200_ERR_ARG, "bad rank / nranks / slab_widths");
    {   // the carries make the binary pass see the WHOLE z extent: its squared distances must stay inside the 24-bit
        // heights of the envelope passes, exactly as for an unsharded grid
        int64_t total_nz = 0;
        for (int r = 0; r < nranks; ++r) {
            if (slab_widths[r] <= 0) return fail(ctx, SDF_B200_ERR_ARG, "empty slab (every rank must own at least one z plane)");
            total_nz += slab_widths[r];
        }
        if (total_nz > SDF_B200_MAX_AXIS)
            return fail(ctx, SDF_B200_ERR_UNSUPPORTED, "slabs add up to a z axis longer than SDF_B200_MAX_AXIS (2048)");
        if (slab_widths[rank] != nz_local) return fail(ctx, SDF_B200_ERR_ARG, "slab_widths[rank] != nz_local");
    }
    const int64_t nlines = nx * ny;
    if (nlines * nz_local > 0 && (!d_mask || !d_sdf || !d_all_summaries)) return fail(ctx, SDF_B200_ERR_ARG, "null device pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    const void* carries = nullptr;
    if (nlines * nz_local > 0) {
        if ((rc = grow(ctx, &ctx->d_carries, &ctx->carries_bytes, (size_t)nlines * 8))) return rc;
        if (!ctx->d_widths) CK(cudaMalloc(&ctx->d_widths, 64 * sizeof(int)), "cudaMalloc(widths)");
        CK(cudaMemcpyAsync(ctx->d_widths, slab_widths, (size_t)nranks * sizeof(int), cudaMemcpyHostToDevice, st), "H2D widths");
        CK(sdfb200::launch_slab_carry(d_all_summaries, ctx->d_widths, nranks, rank, nlines, ctx->d_carries, st),
           "launch slab_carry");
        carries = ctx->d_carries;
    }
    int64_t nz_total = 0;
    for (int r = 0; r < nranks; ++r) nz_total += slab_widths[r];
    rc = run_passes(ctx, d_mask, nx, ny, nz_local, resolution, reinterpret_cast<int32_t*>(d_sdf), true, carries,
                    keep_extrema != 0, 0xffu, nz_total);
    if (rc == SDF_B200_OK && carries) ctx->launches++;
    return rc;
}

int sdf_b200_signed_d2_from_mask(sdf_b200_ctx* ctx, const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int32_t* out)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz);
    if (rc) return rc;
    const int64_t N = nx * ny * nz;
    if (N == 0) return SDF_B200_OK;
    if (!mask || !out) return fail(ctx, SDF_B200_ERR_ARG, "null host pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if ((rc = grow(ctx, &ctx->d_mask, &ctx->mask_bytes, (size_t)N))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)N * 4))) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->d_mask, mask, (size_t)N, cudaMemcpyHostToDevice, st), "H2D mask");
    rc = run_passes(ctx, (const uint8_t*)ctx->d_mask, nx, ny, nz, 1.0, (int32_t*)ctx->d_out, false);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->d_out, (size_t)N * 4, cudaMemcpyDeviceToHost, st), "D2H d2");
    CK(cudaStreamSynchronize(st), "sync");
    return SDF_B200_OK;
}

int sdf_b200_get_extrema_raw(sdf_b200_ctx* ctx, int32_t* out_max_free_d2, int32_t* out_max_filled_d2)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    if (!ctx->built) return fail(ctx, SDF_B200_ERR_ARG, "sdf_b200_get_extrema_raw: no finished build on this context");
    if (ctx->peers.size() > 1) {
        if (out_max_free_d2) *out_max_free_d2 = ctx->raw_ext[0];
        if (out_max_filled_d2) *out_max_filled_d2 = ctx->raw_ext[1];
        return SDF_B200_OK;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int h[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(h, ctx->d_ext, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream), "copy extrema");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    if (int rc = check_tile_flag(ctx, ctx, h[2])) return rc;
    if (out_max_free_d2) *out_max_free_d2 = h[0];
    if (out_max_filled_d2) *out_max_filled_d2 = h[1];
    return SDF_B200_OK;
}

int sdf_b200_extrema_from_raw(int32_t max_free_d2, int32_t max_filled_d2, double resolution, double* out_max, double* out_min)
{
    const double inf = std::numeric_limits<double>::infinity();
    auto dist = [&](int m) { return (m >= SDF_B200_INF) ? inf : sqrt((double)m) * resolution; };
    if (out_max) *out_max = max_free_d2 > 0 ? dist(max_free_d2) : -inf;
    if (out_min) *out_min = max_filled_d2 > 0 ? -dist(max_filled_d2) : inf;
    return SDF_B200_OK;
}

int sdf_b200_get_extrema(sdf_b200_ctx* ctx, double* out_max, double* out_min)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    if (!ctx->built) return fail(ctx, SDF_B200_ERR_ARG, "sdf_b200_get_extrema: no finished build on this context");
    if (ctx->peers.size() > 1) return sdf_b200_extrema_from_raw(ctx->raw_ext[0], ctx->raw_ext[1], ctx->res, out_max, out_min);
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int h[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(h, ctx->d_ext, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream), "copy extrema");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    if (int rc = check_tile_flag(ctx, ctx, h[2])) return rc;
    const double inf = std::numeric_limits<double>::infinity();
    // distance = sqrt(d2_filled)*res - sqrt(d2_free)*res, one term being 0 (sdf_generation.hpp:254-256);
    // sqrt is monotone, so the extrema of the f64 value sit at the per-polarity maxima of d2.
    auto dist = [&](int m) { return (m >= SDF_B200_INF) ? inf : sqrt((double)m) * ctx->res; };
    const bool any_free = h[0] > 0, any_filled = h[1] > 0;
    const double mx = any_free ? dist(h[0]) : -inf;                 // no free voxel: every value is -inf (or grid empty)
    const double mn = any_filled ? -dist(h[1]) : inf;               // no filled voxel: every value is +inf (or grid empty)
    if (out_max) *out_max = mx;
    if (out_min) *out_min = mn;
    return SDF_B200_OK;
}

int sdf_b200_pack_cells_device(sdf_b200_ctx* ctx, const void* d_cells, int64_t n, int unknown_is_filled,
                               uint8_t* d_mask)
{
    if (!ctx || n < 0) return SDF_B200_ERR_ARG;
    if (n > 0 && (!d_cells || !d_mask)) return fail(ctx, SDF_B200_ERR_ARG, "null device pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(sdfb200::launch_pack_cells(d_cells, n, unknown_is_filled, d_mask, ctx->stream), "launch pack_cells");
    return SDF_B200_OK;
}

}  // extern "C"

// ================================================================================= multi-device builds (one process)
// The grid is cut into z slabs, one per device (include/sdf_b200.h, "Slab builds").  Step 2 of the protocol, the
// exchange of the 4-byte per-line summaries, is either
//   p2p   (default on NVLink / NVSwitch machines) the summary kernel stores each word straight into every peer's gather
//         buffer (slab_summary_scatter_kernel) and an event per device orders the peers' builds behind it, or
//   nccl  one ncclAllGather per device inside a group (SDF_B200_EXCHANGE=nccl, or no peer access), or
//   copy  plain cudaMemcpyPeerAsync of the summaries (no peer access and no NCCL).
// NCCL is bound at run time (dlopen "libnccl.so.2"): a link-time dependency would make this library load the system
// NCCL before PyTorch's bundled, newer one in processes that import both, which breaks `import torch`.
namespace {
typedef int ncclResult_t;
typedef void* ncclComm_t;
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (api.lib) {
            api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.lib, "ncclCommInitAll");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
            api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
            api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.AllGather;
        }
    }
    return api;
}
constexpr int NCCL_UINT32 = 3;     // ncclUint32 (nccl.h: ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3)

// balanced z ranges, cut at multiples of 16 cells where possible (16-byte mask loads), as distributed.slab_ranges
void slab_ranges(int64_t nz, int n, std::vector<int64_t>& z0, std::vector<int32_t>& w)
{
    const int64_t align = 16, units = nz / align;
    w.assign(n, 0);
    if (units >= n) {
        for (int r = 0; r < n; ++r) w[r] = (int32_t)((units / n + (r < units % n ? 1 : 0)) * align);
        int64_t sum = 0; for (int r = 0; r < n; ++r) sum += w[r];
        w[n - 1] += (int32_t)(nz - sum);
    } else {
        for (int r = 0; r < n; ++r) w[r] = (int32_t)(nz / n + (r < nz % n ? 1 : 0));
    }
    z0.assign(n, 0);
    for (int r = 1; r < n; ++r) z0[r] = z0[r - 1] + w[r - 1];
}
}  // namespace

// Host buffers -> z slabs on the devices of a multi-device context -> host buffer.  Result and extrema are identical to
// the single-device build: only the binary pass crosses a slab cut, and it gets its context through the carries.
static int build_multi_host(sdf_b200_ctx* lead, const void* h_in, bool is_cells, int unknown_is_filled, int64_t nx,
                            int64_t ny, int64_t nz, double res, float* out_sdf, double* out_max, double* out_min)
{
    sdf_b200_ctx* ctx = lead;                                   // for CK / fail
    const int64_t nlines = nx * ny;
    int n = (int)lead->peers.size();
    if ((int64_t)n > nz) n = (int)nz;                            // never an empty slab
    std::vector<int64_t> z0; std::vector<int32_t> w;
    slab_ranges(nz, n, z0, w);
    int rc;
    // ---- upload (strided: a z slab is nlines rows of w[r] bytes in the host array) and summaries
    std::vector<void*> gather((size_t)n);
    for (int r = 0; r < n; ++r) {
        sdf_b200_ctx* p = lead->peers[r];
        CK(cudaSetDevice(p->device), "cudaSetDevice");
        if ((rc = grow(lead, &p->d_mask, &p->mask_bytes, (size_t)(nlines * w[r])))) return rc;
        if ((rc = grow(lead, &p->d_out, &p->out_bytes, (size_t)(nlines * w[r]) * 4))) return rc;
        if ((rc = grow(lead, &p->d_gather, &p->gather_bytes, (size_t)n * nlines * 4))) return rc;
        if ((rc = grow(lead, &p->d_carries, &p->carries_bytes, (size_t)nlines * 8))) return rc;
        if (!p->d_widths) CK(cudaMalloc(&p->d_widths, 64 * sizeof(int)), "cudaMalloc(widths)");
        gather[r] = p->d_gather;
    }
    const bool want_nccl = [] { const char* e = getenv("SDF_B200_EXCHANGE"); return e && e[0] == 'n'; }();
    const bool use_nccl = n > 1 && (want_nccl || !lead->p2p) && nccl().ok && lead->nccl_comm;
    for (int r = 0; r < n; ++r) {
        sdf_b200_ctx* p = lead->peers[r];
        CK(cudaSetDevice(p->device), "cudaSetDevice");
        cudaStream_t st = p->stream;
        if (is_cells) {
            if ((rc = grow(lead, &p->d_in, &p->in_bytes, (size_t)(nlines * w[r]) * 8))) return rc;
            CK(cudaMemcpy2DAsync(p->d_in, (size_t)w[r] * 8, (const char*)h_in + z0[r] * 8, (size_t)nz * 8, (size_t)w[r] * 8,
                                 (size_t)nlines, cudaMemcpyHostToDevice, st), "H2D cells slab");
            CK(sdfb200::launch_pack_cells(p->d_in, nlines * w[r], unknown_is_filled, (uint8_t*)p->d_mask, st), "launch pack_cells");
        } else {
            CK(cudaMemcpy2DAsync(p->d_mask, (size_t)w[r], (const uint8_t*)h_in + z0[r], (size_t)nz, (size_t)w[r],
                                 (size_t)nlines, cudaMemcpyHostToDevice, st), "H2D mask slab");
        }
        if (n == 1) continue;
        if (lead->p2p && !use_nccl) {
            CK(sdfb200::launch_slab_summary_scatter((const uint8_t*)p->d_mask, w[r], nlines, gather.data(), n, r, st),
               "launch slab_summary_scatter");
        } else {
            uint32_t* mine = (uint32_t*)p->d_gather + (size_t)r * nlines;
            CK(sdfb200::launch_slab_summary((const uint8_t*)p->d_mask, w[r], nlines, mine, st), "launch slab_summary");
            if (!use_nccl)
                for (int k = 0; k < n; ++k)
                    if (k != r)
                        CK(cudaMemcpyPeerAsync((uint32_t*)lead->peers[k]->d_gather + (size_t)r * nlines, lead->peers[k]->device,
                                               mine, p->device, (size_t)nlines * 4, st), "peer copy of the summaries");
        }
        if (!use_nccl) CK(cudaEventRecord(p->ev_summary, st), "event");
    }
    if (use_nccl) {
        if (nccl().GroupStart() != 0) return fail(lead, SDF_B200_ERR_CUDA, "ncclGroupStart failed");
        for (int r = 0; r < n; ++r) {
            sdf_b200_ctx* p = lead->peers[r];
            cudaSetDevice(p->device);
            const ncclResult_t e = nccl().AllGather((uint32_t*)p->d_gather + (size_t)r * nlines, p->d_gather, (size_t)nlines,
                                                    NCCL_UINT32, (ncclComm_t)p->nccl_comm, p->stream);
            if (e != 0) { nccl().GroupEnd(); return fail(lead, SDF_B200_ERR_CUDA, nccl().GetErrorString ? nccl().GetErrorString(e) : "ncclAllGather failed"); }
        }
        if (nccl().GroupEnd() != 0) return fail(lead, SDF_B200_ERR_CUDA, "ncclGroupEnd failed");
    }
    // ---- per device: wait for every peer's summaries, resolve the carries, three passes, download
    for (int r = 0; r < n; ++r) {
        sdf_b200_ctx* p = lead->peers[r];
        CK(cudaSetDevice(p->device), "cudaSetDevice");
        cudaStream_t st = p->stream;
        const void* carries = nullptr;
        if (n > 1) {
            if (!use_nccl)
                for (int k = 0; k < n; ++k)
                    if (k != r) CK(cudaStreamWaitEvent(st, lead->peers[k]->ev_summary, 0), "wait for peer summaries");
            CK(cudaMemcpyAsync(p->d_widths, w.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st), "H2D widths");
            CK(sdfb200::launch_slab_carry((const uint32_t*)p->d_gather, p->d_widths, n, r, nlines, p->d_carries, st),
               "launch slab_carry");
            carries = p->d_carries;
        }
        rc = run_passes(p, (const uint8_t*)p->d_mask, nx, ny, w[r], res, (int32_t*)p->d_out, true, carries, false, 0xffu, nz);
        if (rc) { if (p != lead) snprintf(lead->err, sizeof(lead->err), "%s", p->err); return rc; }
        CK(cudaMemcpy2DAsync(out_sdf + z0[r], (size_t)nz * 4, p->d_out, (size_t)w[r] * 4, (size_t)w[r] * 4, (size_t)nlines,
                             cudaMemcpyDeviceToHost, st), "D2H sdf slab");
    }
    int raw[2] = {0, 0}, launches = 0;
    for (int r = 0; r < n; ++r) {
        sdf_b200_ctx* p = lead->peers[r];
        CK(cudaSetDevice(p->device), "cudaSetDevice");
        int h[3] = {0, 0, 0};
        CK(cudaMemcpyAsync(h, p->d_ext, sizeof(h), cudaMemcpyDeviceToHost, p->stream), "copy extrema");
        CK(cudaStreamSynchronize(p->stream), "sync");
        if (int frc = check_tile_flag(lead, p, h[2])) return frc;
        raw[0] = h[0] > raw[0] ? h[0] : raw[0];
        raw[1] = h[1] > raw[1] ? h[1] : raw[1];
        launches += p->launches + (n > 1 ? 2 : 0) + (is_cells ? 1 : 0);
    }
    CK(cudaSetDevice(lead->device), "cudaSetDevice");
    lead->raw_ext[0] = raw[0]; lead->raw_ext[1] = raw[1];
    lead->launches = launches;
    lead->built = true;
    lead->res = res;
    return sdf_b200_extrema_from_raw(raw[0], raw[1], res, out_max, out_min);
}

extern "C" {

static int build_host_common(sdf_b200_ctx* ctx, const void* h_in, size_t in_bytes, bool is_cells, int unknown_is_filled,
                             int64_t nx, int64_t ny, int64_t nz, double res, float* out_sdf, double* out_max,
                             double* out_min)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz);
    if (rc) return rc;
    const int64_t N = nx * ny * nz;
    if (N > 0 && (!h_in || !out_sdf)) return fail(ctx, SDF_B200_ERR_ARG, "null host pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    if (N > 0) {
        if ((rc = grow(ctx, &ctx->d_mask, &ctx->mask_bytes, (size_t)N))) return rc;
        if ((rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)N * 4))) return rc;
        if (is_cells) {
            if ((rc = grow(ctx, &ctx->d_in, &ctx->in_bytes, in_bytes))) return rc;
            CK(cudaMemcpyAsync(ctx->d_in, h_in, in_bytes, cudaMemcpyHostToDevice, st), "H2D cells");
            CK(sdfb200::launch_pack_cells(ctx->d_in, N, unknown_is_filled, (uint8_t*)ctx->d_mask, st), "launch pack_cells");
        } else {
            CK(cudaMemcpyAsync(ctx->d_mask, h_in, in_bytes, cudaMemcpyHostToDevice, st), "H2D mask");
        }
    }
    rc = run_passes(ctx, (const uint8_t*)ctx->d_mask, nx, ny, nz, res, (int32_t*)ctx->d_out, true);
    if (rc) return rc;
    ctx->last_dims[0] = nx; ctx->last_dims[1] = ny; ctx->last_dims[2] = nz;
    if (is_cells && N > 0) ctx->launches++;
    if (N > 0) CK(cudaMemcpyAsync(out_sdf, ctx->d_out, (size_t)N * 4, cudaMemcpyDeviceToHost, st), "D2H sdf");
    return sdf_b200_get_extrema(ctx, out_max, out_min);
}

// Large grids from host masks: the same three passes, software-pipelined against the PCIe copies.
//   x-chunks :  H2D(mask chunk i) on the copy stream  ->  K1 + K2 of chunk i on the compute stream (both are local to
//               an x range), so the upload of chunk i+1 hides behind the kernels of chunk i;
//   column-chunks of the final x pass (a y range, all z): each finishes every x of its columns, so its result leaves
//               with one strided D2H (nx rows of ychunk*nz floats) while the next chunk is still computing.
// Result and extrema are identical to the unpipelined path: chunks only partition independent lines.
static int build_host_pipelined(sdf_b200_ctx* ctx, const void* h_in, bool is_cells, int unknown_is_filled, int64_t nx,
                                int64_t ny, int64_t nz, double res, float* out_sdf, double* out_max, double* out_min)
{
    constexpr int NCH = 8;
    constexpr int LUT_N = 1 << 20;
    const int64_t N = nx * ny * nz, plane = ny * nz;
    const uint8_t* h_mask = (const uint8_t*)h_in;
    int rc;
    if ((rc = grow(ctx, &ctx->d_mask, &ctx->mask_bytes, (size_t)N))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)N * 4))) return rc;
    // cells (8 B per voxel): two x-chunks of staging, used alternately, so the upload of chunk i+1 overlaps the pack of chunk i
    const int64_t chunk_planes = (nx + NCH - 1) / NCH;
    if (is_cells && (rc = grow(ctx, &ctx->d_in, &ctx->in_bytes, (size_t)(2 * chunk_planes * plane) * 8))) return rc;
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    uint8_t* d_mask = (uint8_t*)ctx->d_mask;
    int32_t* d_field = (int32_t*)ctx->d_out;
    const LinePlan LP = plan_line_passes(ctx, nx, ny, nz, nz, (nx + NCH - 1) / NCH);
    ctx->last_hull = LP.mode;
    int32_t* tmp = (int32_t*)ctx->d_tmp;
    const int64_t nwy = (ny + 31) / 32, nwx = (nx + 31) / 32;
    ctx->launches = 0;
    ctx->timed = false;
    CK(cudaMemsetAsync(ctx->d_ext, 0, 2 * sizeof(int), st), "memset extrema");
    if (ctx->d_lut == nullptr || ctx->lut_res != res) {
        if (!ctx->d_lut) CK(cudaMalloc(&ctx->d_lut, (size_t)LUT_N * sizeof(float)), "cudaMalloc(lut)");
        CK(sdfb200::launch_sqrt_lut(ctx->d_lut, LUT_N, res, st), "launch sqrt_lut");
        ctx->lut_res = res;
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->chunk_ev[16], st), "event");          // copies must not overtake earlier work on `st`
    CK(cudaStreamWaitEvent(cs, ctx->chunk_ev[16], 0), "wait");
    for (int i = 0; i < NCH; ++i) {                                 // ---- upload + z / y passes, chunked along x
        const int64_t x0 = nx * i / NCH, x1 = nx * (i + 1) / NCH;
        if (x1 == x0) continue;
        const int64_t off = x0 * plane, cnt = (x1 - x0) * plane;
        if (is_cells) {
            char* stage = (char*)ctx->d_in + (size_t)(i & 1) * (size_t)(chunk_planes * plane) * 8;
            if (i >= 2) CK(cudaStreamWaitEvent(cs, ctx->pack_ev[i & 1], 0), "wait for the staging half");   // its previous pack is done
            CK(cudaMemcpyAsync(stage, (const char*)h_in + (size_t)off * 8, (size_t)cnt * 8, cudaMemcpyHostToDevice, cs), "H2D cells chunk");
            CK(cudaEventRecord(ctx->chunk_ev[i], cs), "event");
            CK(cudaStreamWaitEvent(st, ctx->chunk_ev[i], 0), "wait");
            CK(sdfb200::launch_pack_cells(stage, cnt, unknown_is_filled, d_mask + off, st), "launch pack_cells");
            CK(cudaEventRecord(ctx->pack_ev[i & 1], st), "event");
            ctx->launches++;
        } else {
            CK(cudaMemcpyAsync(d_mask + off, h_mask + off, (size_t)cnt, cudaMemcpyHostToDevice, cs), "H2D mask chunk");
            CK(cudaEventRecord(ctx->chunk_ev[i], cs), "event");
            CK(cudaStreamWaitEvent(st, ctx->chunk_ev[i], 0), "wait");
        }
        // (ny >= 64 here: there is always a y pass)
        CK(sdfb200::launch_binary_z(d_mask + off, (LP.mode == HULL_MIXED ? tmp : d_field) + off, (int)nz, (x1 - x0) * ny, nullptr, st),
           "launch binary_z");
        ctx->launches++;
        void* stack_y = ctx->d_stack;                               // dead after each launch: every chunk reuses it
        uint32_t* pol_y = (uint32_t*)ctx->d_pol + x0 * nwy * nz;
        if (LP.mode == HULL_LINE)            // x-chunk of the y pass: field[x0..x1) -> tmp[y][x0..x1)[z]
            CK(sdfb200::launch_line_hull(d_field + off, tmp + x0 * nz, stack_y, pol_y, (int)ny, nz, nz, x1 - x0, plane, nx * nz, nz, LP.fmt_y,
                                         false, res, ctx->d_ext, nullptr, 0, st), "launch line_hull(y)");
        else if (LP.mode == HULL_MIXED)      // tmp[x0..x1) -> field[x0..x1)
            CK(sdfb200::launch_line_hull(tmp + off, d_field + off, stack_y, pol_y, (int)ny, nz, nz, x1 - x0, plane, nz, plane, LP.fmt_y,
                                         false, res, ctx->d_ext, nullptr, 0, st), "launch line_hull(y)");
        else
            CK(sdfb200::launch_band_strided(d_field + off, (int)ny, nz, nz, x1 - x0, plane, false, res, ctx->d_ext, nullptr,
                                            nullptr, 0, st), "launch band_strided(y)");
        ctx->launches++;
    }
    for (int j = 0; j < NCH; ++j) {                                 // ---- final x pass + download, chunked along y
        const int64_t y0 = ny * j / NCH, y1 = ny * (j + 1) / NCH;
        if (y1 == y0) continue;
        const int64_t col0 = y0 * nz, ncols = (y1 - y0) * nz;
        if (LP.mode == HULL_LINE)            // y-chunk of the x pass: tmp[y0..y1) -> field[.][y0..y1)[z]
            CK(sdfb200::launch_line_hull(tmp + y0 * nx * nz, d_field + col0, ctx->d_stack,
                                         (uint32_t*)ctx->d_pol + y0 * nwx * nz, (int)nx, nz, nz, y1 - y0, nx * nz, plane, nz, LP.fmt_x, true,
                                         res, ctx->d_ext, ctx->d_lut, LUT_N, st), "launch line_hull(x)");
        else
            CK(sdfb200::launch_band_strided(d_field + col0, (int)nx, plane, ncols, 1, 0, true, res, ctx->d_ext, nullptr,
                                            ctx->d_lut, LUT_N, st), "launch band_strided(x)");
        ctx->launches++;
        CK(cudaEventRecord(ctx->chunk_ev[8 + j], st), "event");
        CK(cudaStreamWaitEvent(cs, ctx->chunk_ev[8 + j], 0), "wait");
        CK(cudaMemcpy2DAsync(out_sdf + col0, (size_t)plane * 4, d_field + col0, (size_t)plane * 4, (size_t)ncols * 4,
                             (size_t)nx, cudaMemcpyDeviceToHost, cs), "D2H sdf chunk");
    }
    CK(cudaStreamSynchronize(cs), "sync copies");
    ctx->built = true;
    ctx->res = res;
    ctx->last_dims[0] = nx; ctx->last_dims[1] = ny; ctx->last_dims[2] = nz;
    return sdf_b200_get_extrema(ctx, out_max, out_min);
}

int sdf_b200_build_from_mask(sdf_b200_ctx* ctx, const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz,
                             double resolution, float* out_sdf, double* out_max, double* out_min)
{
    if (ctx && ctx->peers.size() > 1 && mask && out_sdf && check_dims(ctx, nx, ny, nz) == SDF_B200_OK && nx * ny * nz > 0)
        return build_multi_host(ctx, mask, false, 0, nx, ny, nz, resolution, out_sdf, out_max, out_min);
    if (ctx && mask && out_sdf && check_dims(ctx, nx, ny, nz) == SDF_B200_OK && nx >= 64 && ny >= 64 &&
        nx * ny * nz >= ((int64_t)1 << 25)) {           // big enough for the copies to be worth overlapping
        CK(cudaSetDevice(ctx->device), "cudaSetDevice");
        return build_host_pipelined(ctx, mask, false, 0, nx, ny, nz, resolution, out_sdf, out_max, out_min);
    }
    return build_host_common(ctx, mask, (size_t)(nx * ny * nz), false, 0, nx, ny, nz, resolution, out_sdf, out_max, out_min);
}

int sdf_b200_build_from_cells(sdf_b200_ctx* ctx, const void* cells, int unknown_is_filled, int64_t nx, int64_t ny,
                              int64_t nz, double resolution, float* out_sdf, double* out_max, double* out_min)
{
    if (ctx && ctx->peers.size() > 1 && cells && out_sdf && check_dims(ctx, nx, ny, nz) == SDF_B200_OK && nx * ny * nz > 0)
        return build_multi_host(ctx, cells, true, unknown_is_filled, nx, ny, nz, resolution, out_sdf, out_max, out_min);
    if (ctx && cells && out_sdf && check_dims(ctx, nx, ny, nz) == SDF_B200_OK && nx >= 64 && ny >= 64 &&
        nx * ny * nz >= ((int64_t)1 << 25)) {           // big enough for the copies to be worth overlapping
        CK(cudaSetDevice(ctx->device), "cudaSetDevice");
        return build_host_pipelined(ctx, cells, true, unknown_is_filled, nx, ny, nz, resolution, out_sdf, out_max, out_min);
    }
    return build_host_common(ctx, cells, (size_t)(nx * ny * nz) * 8, true, unknown_is_filled, nx, ny, nz, resolution,
                             out_sdf, out_max, out_min);
}

int sdf_b200_last_timing(sdf_b200_ctx* ctx, float* out_ms4)
{
    if (!ctx || !out_ms4) return SDF_B200_ERR_ARG;
    if (!ctx->timed) return fail(ctx, SDF_B200_ERR_ARG, "sdf_b200_last_timing: nothing timed");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaEventSynchronize(ctx->ev[3]), "event sync");
    for (int i = 0; i < 3; ++i) CK(cudaEventElapsedTime(&out_ms4[i], ctx->ev[i], ctx->ev[i + 1]), "elapsed");
    CK(cudaEventElapsedTime(&out_ms4[3], ctx->ev[0], ctx->ev[3]), "elapsed");
    return SDF_B200_OK;
}

int sdf_b200_last_launches(const sdf_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int sdf_b200_last_envelope_kernel(const sdf_b200_ctx* ctx) { return ctx ? ctx->last_hull : 0; }

int sdf_b200_debug_phase_clocks(sdf_b200_ctx* ctx, int enable, long long* out, int n)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t bytes = (size_t)2 * 1024 * 8 * sizeof(long long);
    if (enable && !ctx->d_dbg) {
        CK(cudaMalloc(&ctx->d_dbg, bytes), "cudaMalloc(debug clocks)");
        CK(cudaMemset(ctx->d_dbg, 0, bytes), "memset");
    }
    if (out && n > 0 && ctx->d_dbg) {
        CK(cudaStreamSynchronize(ctx->stream), "sync");
        const size_t want = (size_t)n * sizeof(long long);
        CK(cudaMemcpy(out, ctx->d_dbg, want < bytes ? want : bytes, cudaMemcpyDeviceToHost), "copy debug clocks");
    }
    if (!enable && ctx->d_dbg) { cudaFree(ctx->d_dbg); ctx->d_dbg = nullptr; }
    return SDF_B200_OK;
}

int sdf_b200_slab_summary_scatter_device(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz_local,
                                         void* const* peer_gather, int nranks, int rank)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz_local);
    if (rc) return rc;
    if (nranks < 1 || nranks > SDF_MAX_PEERS || rank < 0 || rank >= nranks || !peer_gather)
        return fail(ctx, SDF_B200_ERR_ARG, "bad rank / nranks / peer_gather");
    if (nx * ny * nz_local == 0) return SDF_B200_OK;
    if (!d_mask) return fail(ctx, SDF_B200_ERR_ARG, "null device pointer");
    for (int k = 0; k < nranks; ++k) if (!peer_gather[k]) return fail(ctx, SDF_B200_ERR_ARG, "null peer gather buffer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(sdfb200::launch_slab_summary_scatter(d_mask, (int)nz_local, nx * ny, peer_gather, nranks, rank, ctx->stream),
       "launch slab_summary_scatter");
    return SDF_B200_OK;
}

int sdf_b200_create_multi(sdf_b200_ctx** out, const int* devices, int ndev)
{
    if (!out) return fail(nullptr, SDF_B200_ERR_ARG, "sdf_b200_create_multi: out is NULL");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > SDF_MAX_PEERS) return fail(nullptr, SDF_B200_ERR_ARG, "sdf_b200_create_multi: bad device list");
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(nullptr, SDF_B200_ERR_ARG, "sdf_b200_create_multi: a device is listed twice");
    std::vector<sdf_b200_ctx*> ps;
    for (int i = 0; i < ndev; ++i) {
        sdf_b200_ctx* p = nullptr;
        const int rc = sdf_b200_create(&p, devices[i]);
        if (rc != SDF_B200_OK) { for (sdf_b200_ctx* q : ps) sdf_b200_destroy(q); return rc; }
        ps.push_back(p);
    }
    if (ndev == 1) { *out = ps[0]; return SDF_B200_OK; }
    // peer access between every pair (NVLink / NVSwitch): lets the summary kernel store into the peers' gather buffers
    bool p2p = true;
    for (int i = 0; i < ndev && p2p; ++i)
        for (int j = 0; j < ndev && p2p; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) { cudaGetLastError(); p2p = false; }
        }
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < ndev; ++i) {
        if ((e = cudaSetDevice(devices[i])) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&ps[i]->ev_summary, cudaEventDisableTiming)) != cudaSuccess) break;
        if (p2p)
            for (int j = 0; j < ndev; ++j) {
                if (i == j) continue;
                const cudaError_t pe = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) p2p = false;
                cudaGetLastError();
            }
    }
    if (e != cudaSuccess) {
        fail(nullptr, SDF_B200_ERR_CUDA, "sdf_b200_create_multi: events", e);
        for (sdf_b200_ctx* q : ps) sdf_b200_destroy(q);
        return SDF_B200_ERR_CUDA;
    }
    const char* ex = getenv("SDF_B200_EXCHANGE");
    if ((!p2p || (ex && ex[0] == 'n')) && nccl().ok) {          // NCCL communicators only when they will be used
        std::vector<ncclComm_t> comms((size_t)ndev, nullptr);
        if (nccl().CommInitAll(comms.data(), ndev, devices) == 0) {
            for (int i = 0; i < ndev; ++i) ps[i]->nccl_comm = comms[i];
            nccl_destroy_comm = [](void* c) { nccl().CommDestroy((ncclComm_t)c); };
        }
    }
    cudaSetDevice(devices[0]);
    ps[0]->peers = ps;
    ps[0]->p2p = p2p;
    *out = ps[0];
    return SDF_B200_OK;
}

int sdf_b200_num_devices(const sdf_b200_ctx* ctx) { return !ctx ? 0 : (ctx->peers.empty() ? 1 : (int)ctx->peers.size()); }

const char* sdf_b200_exchange_kind(const sdf_b200_ctx* ctx)
{
    if (!ctx || ctx->peers.size() < 2) return "none";
    const char* ex = getenv("SDF_B200_EXCHANGE");
    if (ctx->nccl_comm && (!ctx->p2p || (ex && ex[0] == 'n'))) return "nccl";
    return ctx->p2p ? "p2p" : "copy";
}

int sdf_b200_gradient_device(sdf_b200_ctx* ctx, const float* d_sdf, int64_t nx, int64_t ny, int64_t nz, double resolution,
                             int enable_edge_gradients, int out_f64, void* d_grad)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    if (nx < 0 || ny < 0 || nz < 0 || !(resolution > 0.0)) return fail(ctx, SDF_B200_ERR_ARG, "bad grid size / resolution");
    if (nx * ny * nz == 0) return SDF_B200_OK;
    if (!d_sdf || !d_grad) return fail(ctx, SDF_B200_ERR_ARG, "null device pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(sdfb200::launch_gradient(d_sdf, nx, ny, nz, resolution, enable_edge_gradients, out_f64, d_grad, ctx->stream),
       "launch gradient");
    return SDF_B200_OK;
}

int sdf_b200_gradient_host(sdf_b200_ctx* ctx, const float* sdf, int64_t nx, int64_t ny, int64_t nz, double resolution,
                           int enable_edge_gradients, int out_f64, void* out_grad)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    if (nx < 0 || ny < 0 || nz < 0 || !(resolution > 0.0)) return fail(ctx, SDF_B200_ERR_ARG, "bad grid size / resolution");
    const int64_t N = nx * ny * nz;
    if (N == 0) return SDF_B200_OK;
    if (!out_grad) return fail(ctx, SDF_B200_ERR_ARG, "null host pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    int rc;
    if (sdf) {
        if ((rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)N * 4))) return rc;
        CK(cudaMemcpyAsync(ctx->d_out, sdf, (size_t)N * 4, cudaMemcpyHostToDevice, st), "H2D sdf");
        ctx->last_dims[0] = nx; ctx->last_dims[1] = ny; ctx->last_dims[2] = nz;
    } else if (!ctx->d_out || ctx->last_dims[0] != nx || ctx->last_dims[1] != ny || ctx->last_dims[2] != nz ||
               ctx->peers.size() > 1) {
        return fail(ctx, SDF_B200_ERR_ARG, "sdf == NULL needs the field of a host build of the same grid on this (single-device) context");
    }
    const size_t esz = out_f64 ? 8 : 4;
    if ((rc = grow(ctx, &ctx->d_grad, &ctx->grad_bytes, (size_t)N * 3 * esz))) return rc;
    CK(sdfb200::launch_gradient((const float*)ctx->d_out, nx, ny, nz, resolution, enable_edge_gradients, out_f64, ctx->d_grad, st),
       "launch gradient");
    CK(cudaMemcpyAsync(out_grad, ctx->d_grad, (size_t)N * 3 * esz, cudaMemcpyDeviceToHost, st), "D2H gradient");
    CK(cudaStreamSynchronize(st), "sync");
    return SDF_B200_OK;
}

// Batched object SDFs: the 16-byte tagged cells go to the device ONCE; every pass over them answers up to 8 queries
// (one mask bit each); the builds then run back to back on the compute stream while the copy stream downloads the
// previous result from the other of two result buffers.
int sdf_b200_build_object_sdfs(sdf_b200_ctx* ctx, const void* tagged_cells, int64_t nx, int64_t ny, int64_t nz,
                               double resolution, int unknown_is_filled, const sdf_b200_object_query* queries, int nqueries,
                               float* out_sdfs, double* out_max, double* out_min)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    int rc = check_dims(ctx, nx, ny, nz);
    if (rc) return rc;
    if (nqueries < 0 || (nqueries > 0 && !queries)) return fail(ctx, SDF_B200_ERR_ARG, "bad query list");
    for (int k = 0; k < nqueries; ++k) if (queries[k].kind > 2u) return fail(ctx, SDF_B200_ERR_ARG, "unknown query kind");
    const int64_t N = nx * ny * nz;
    if (nqueries == 0) return SDF_B200_OK;
    if (N > 0 && (!tagged_cells || !out_sdfs)) return fail(ctx, SDF_B200_ERR_ARG, "null host pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    const double inf = std::numeric_limits<double>::infinity();
    if (N == 0) {
        for (int k = 0; k < nqueries; ++k) { if (out_max) out_max[k] = -inf; if (out_min) out_min[k] = inf; }
        return SDF_B200_OK;
    }
    if ((rc = grow(ctx, &ctx->d_in, &ctx->in_bytes, (size_t)N * 16))) return rc;
    if ((rc = grow(ctx, &ctx->d_mask, &ctx->mask_bytes, (size_t)N))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)N * 4))) return rc;
    if ((rc = grow(ctx, &ctx->d_out2, &ctx->out2_bytes, (size_t)N * 4))) return rc;
    if (ctx->ext_objs_n < (size_t)nqueries) {
        if (ctx->d_ext_objs) cudaFree(ctx->d_ext_objs);
        ctx->d_ext_objs = nullptr; ctx->ext_objs_n = 0;
        CK(cudaMalloc(&ctx->d_ext_objs, (size_t)nqueries * 2 * sizeof(int)), "cudaMalloc(object extrema)");
        ctx->ext_objs_n = (size_t)nqueries;
    }
    CK(cudaMemcpyAsync(ctx->d_in, tagged_cells, (size_t)N * 16, cudaMemcpyHostToDevice, st), "H2D tagged cells");
    int launches = 0;
    for (int k0 = 0; k0 < nqueries; k0 += 8) {
        const int nq = nqueries - k0 < 8 ? nqueries - k0 : 8;
        uint32_t ids[8], kinds[8];
        for (int q = 0; q < nq; ++q) { ids[q] = queries[k0 + q].object_id; kinds[q] = queries[k0 + q].kind; }
        CK(sdfb200::launch_pack_tagged(ctx->d_in, N, unknown_is_filled, ids, kinds, nq, (uint8_t*)ctx->d_mask, st), "launch pack_tagged");
        ++launches;
        for (int q = 0; q < nq; ++q) {
            const int k = k0 + q;
            void* buf = (k & 1) ? ctx->d_out2 : ctx->d_out;
            if (k >= 2) CK(cudaStreamWaitEvent(st, ctx->chunk_ev[8 + (k & 1)], 0), "wait for the download of this buffer");
            rc = run_passes(ctx, (const uint8_t*)ctx->d_mask, nx, ny, nz, resolution, (int32_t*)buf, true, nullptr, false, 1u << q);
            if (rc) return rc;
            launches += ctx->launches;
            CK(cudaMemcpyAsync(ctx->d_ext_objs + 2 * k, ctx->d_ext, 2 * sizeof(int), cudaMemcpyDeviceToDevice, st), "keep extrema");
            CK(cudaEventRecord(ctx->chunk_ev[k & 1], st), "event");
            CK(cudaStreamWaitEvent(cs, ctx->chunk_ev[k & 1], 0), "wait");
            CK(cudaMemcpyAsync(out_sdfs + (size_t)k * N, buf, (size_t)N * 4, cudaMemcpyDeviceToHost, cs), "D2H object sdf");
            CK(cudaEventRecord(ctx->chunk_ev[8 + (k & 1)], cs), "event");
        }
    }
    std::vector<int> raw((size_t)nqueries * 2);
    CK(cudaMemcpyAsync(raw.data(), ctx->d_ext_objs, raw.size() * sizeof(int), cudaMemcpyDeviceToHost, st), "copy extrema");
    int tile_flag = 0;
    CK(cudaMemcpyAsync(&tile_flag, ctx->d_ext + 2, sizeof(int), cudaMemcpyDeviceToHost, st), "copy flag");
    CK(cudaStreamSynchronize(st), "sync");
    if ((rc = check_tile_flag(ctx, ctx, tile_flag))) { cudaStreamSynchronize(cs); return rc; }
    CK(cudaStreamSynchronize(cs), "sync copies");
    for (int k = 0; k < nqueries; ++k)
        sdf_b200_extrema_from_raw(raw[2 * k], raw[2 * k + 1], resolution, out_max ? out_max + k : nullptr, out_min ? out_min + k : nullptr);
    ctx->launches = launches;
    ctx->last_dims[0] = ctx->last_dims[1] = ctx->last_dims[2] = 0;      // d_out no longer holds "the" field of a host build
    return SDF_B200_OK;
}

// add_virtual_border on the device: one upload of the inner mask(s), two builds on the padded grid, fused select, one
// download of the real grid.  mask_free / mask_filled: the caller's predicate evaluated for the free-space build and for
// the inside build (the reference calls it once per voxel in EACH of its two extractions, sdf_generation.hpp:381-382);
// pass the same pointer twice for a pure predicate.
int sdf_b200_build_with_border(sdf_b200_ctx* ctx, const uint8_t* mask_free, const uint8_t* mask_filled, int64_t nx, int64_t ny,
                               int64_t nz, double resolution, float* out_sdf, double* out_max, double* out_min)
{
    if (!ctx) return SDF_B200_ERR_ARG;
    const int ox = nx > 1 ? 1 : 0, oy = ny > 1 ? 1 : 0, oz = nz > 1 ? 1 : 0;
    const int64_t bx = nx + 2 * ox, by = ny + 2 * oy, bz = nz + 2 * oz;
    int rc = check_dims(ctx, bx, by, bz);
    if (rc) return rc;
    const int64_t N = nx * ny * nz, NB = bx * by * bz;
    const double inf = std::numeric_limits<double>::infinity();
    if (N == 0) { if (out_max) *out_max = -inf; if (out_min) *out_min = inf; return SDF_B200_OK; }
    if (!mask_free || !mask_filled || !out_sdf) return fail(ctx, SDF_B200_ERR_ARG, "null host pointer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    const bool same = mask_free == mask_filled;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->in_bytes, (size_t)N * (same ? 1 : 2)))) return rc;       // inner mask(s)
    if ((rc = grow(ctx, &ctx->d_mask, &ctx->mask_bytes, (size_t)NB))) return rc;                   // padded mask
    if ((rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)NB * 4))) return rc;                 // padded free-space field
    if ((rc = grow(ctx, &ctx->d_out2, &ctx->out2_bytes, (size_t)NB * 4))) return rc;               // padded inside field
    if ((rc = grow(ctx, &ctx->d_grad, &ctx->grad_bytes, (size_t)N * 4))) return rc;                // combined field (staging)
    uint8_t* d_inner = (uint8_t*)ctx->d_in;
    CK(cudaMemcpyAsync(d_inner, mask_free, (size_t)N, cudaMemcpyHostToDevice, st), "H2D mask");
    if (!same) CK(cudaMemcpyAsync(d_inner + N, mask_filled, (size_t)N, cudaMemcpyHostToDevice, st), "H2D mask");
    int launches = 0, raw_free[2] = {0, 0}, raw_filled[2] = {0, 0};
    // free-space build: shell FILLED; its max is the combined max (:412-413)
    CK(sdfb200::launch_border_pad(d_inner, nx, ny, nz, ox, oy, oz, 1, (uint8_t*)ctx->d_mask, st), "launch border_pad");
    if ((rc = run_passes(ctx, (const uint8_t*)ctx->d_mask, bx, by, bz, resolution, (int32_t*)ctx->d_out, true))) return rc;
    launches += ctx->launches + 1;
    CK(cudaMemcpyAsync(raw_free, ctx->d_ext, sizeof(raw_free), cudaMemcpyDeviceToHost, st), "copy extrema");
    // inside build: shell FREE; its min is the combined min
    CK(sdfb200::launch_border_pad(same ? d_inner : d_inner + N, nx, ny, nz, ox, oy, oz, 0, (uint8_t*)ctx->d_mask, st), "launch border_pad");
    if ((rc = run_passes(ctx, (const uint8_t*)ctx->d_mask, bx, by, bz, resolution, (int32_t*)ctx->d_out2, true))) return rc;
    launches += ctx->launches + 1;
    CK(cudaMemcpyAsync(raw_filled, ctx->d_ext, sizeof(raw_filled), cudaMemcpyDeviceToHost, st), "copy extrema");
    CK(sdfb200::launch_border_select((const float*)ctx->d_out, (const float*)ctx->d_out2, nx, ny, nz, ox, oy, oz, (float*)ctx->d_grad, st),
       "launch border_select");
    ++launches;
    CK(cudaMemcpyAsync(out_sdf, ctx->d_grad, (size_t)N * 4, cudaMemcpyDeviceToHost, st), "D2H sdf");
    int tile_flag = 0;
    CK(cudaMemcpyAsync(&tile_flag, ctx->d_ext + 2, sizeof(int), cudaMemcpyDeviceToHost, st), "copy flag");
    CK(cudaStreamSynchronize(st), "sync");
    if ((rc = check_tile_flag(ctx, ctx, tile_flag))) return rc;
    ctx->launches = launches;
    ctx->last_dims[0] = ctx->last_dims[1] = ctx->last_dims[2] = 0;
    double mx = 0, mn = 0, dummy = 0;
    sdf_b200_extrema_from_raw(raw_free[0], raw_free[1], resolution, &mx, &dummy);
    sdf_b200_extrema_from_raw(raw_filled[0], raw_filled[1], resolution, &dummy, &mn);
    if (out_max) *out_max = mx;
    if (out_min) *out_min = mn;
    return SDF_B200_OK;
}

}  // extern "C"
