// sdf_tools_b200/csrc/capi.cu -- the C ABI of libsdf_b200.so (include/sdf_b200.h) over the kernel launchers.
// Host orchestration of sdf_generation::ExtractSignedDistanceField<T> (sdf_generation.hpp:209-271):
//   mask -> K1 (z, binary) -> K2 (y) -> K3 (x + combine) -> f32 field + (max, min).
#include "../../include/sdf_b200.h"
#include "edt_kernels.cuh"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <limits>

struct sdf_b200_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int* d_ext = nullptr;              // [0] max +d2 over free voxels, [1] max d2 over filled voxels
    void* d_carries = nullptr; size_t carries_bytes = 0;   // slab builds: int2 per z line
    int* d_widths = nullptr;
    float* d_lut = nullptr;            // (float)(sqrt((double)m)*res) for m < LUT_N at resolution lut_res
    double lut_res = -1.0;
    long long* d_dbg = nullptr;        // optional phase clocks of the band kernels: [pass 2|3][1024 CTAs][8]
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;  // host entry points: H2D / D2H chunks overlap the kernels of the other chunks
    cudaEvent_t chunk_ev[17] = {};       // [0..7] upload chunks, [8..15] download chunks, [16] stream hand-over
    bool timed = false;
    int launches = 0;
    void* d_in = nullptr;  size_t in_bytes = 0;     // staging for the host entry points
    void* d_mask = nullptr; size_t mask_bytes = 0;
    void* d_out = nullptr; size_t out_bytes = 0;
    double res = 1.0;
    bool built = false;
    char err[512] = {0};
};

static char g_create_err[512] = {0};

static int fail(sdf_b200_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess)
{
    char* dst = ctx ? ctx->err : g_create_err;
    if (e != cudaSuccess) snprintf(dst, 512, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else snprintf(dst, 512, "%s", what);
    return code;
}

#define CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(ctx, SDF_B200_ERR_CUDA, what, _e); } while (0)

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
        (e = cudaMalloc(&ctx->d_ext, 2 * sizeof(int))) != cudaSuccess) {
        fail(nullptr, SDF_B200_ERR_CUDA, "sdf_b200_create: stream/alloc", e);
        delete ctx;
        return SDF_B200_ERR_CUDA;
    }
    for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 17; ++i) cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming);
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return SDF_B200_OK;
}

void sdf_b200_destroy(sdf_b200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 17; ++i) if (ctx->chunk_ev[i]) cudaEventDestroy(ctx->chunk_ev[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->d_ext) cudaFree(ctx->d_ext);
    if (ctx->d_dbg) cudaFree(ctx->d_dbg);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->d_carries) cudaFree(ctx->d_carries);
    if (ctx->d_widths) cudaFree(ctx->d_widths);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_mask) cudaFree(ctx->d_mask);
    if (ctx->d_out) cudaFree(ctx->d_out);
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

// mask -> signed int32 d^2 (final_pass = false) or f32 distance (final_pass = true), in place in d_field
static int run_passes(sdf_b200_ctx* ctx, const uint8_t* d_mask, int64_t nx, int64_t ny, int64_t nz, double res,
                      int32_t* d_field, bool final_pass, const void* d_carries = nullptr, bool keep_extrema = false)
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
    CK(cudaEventRecord(ctx->ev[0], st), "event");
    CK(sdfb200::launch_binary_z(d_mask, d_field, (int)nz, nx * ny, d_carries, st), "launch binary_z");
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[1], st), "event");
    if (ny > 1) {
        CK(sdfb200::launch_band_strided(d_field, (int)ny, nz, nz, nx, ny * nz, false, res, ctx->d_ext, ctx->d_dbg, nullptr, 0, st),
           "launch band_strided(y)");
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev[2], st), "event");
    CK(sdfb200::launch_band_strided(d_field, (int)nx, ny * nz, ny * nz, 1, 0, final_pass, res, ctx->d_ext, ctx->d_dbg ? ctx->d_dbg + 1024 * 8 : nullptr,
                                    final_pass ? ctx->d_lut : nullptr, final_pass ? LUT_N : 0, st),
       "launch band_strided(x)");
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[3], st), "event");
    ctx->timed = true;
    ctx->built = final_pass;
    ctx->res = res;
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
    rc = run_passes(ctx, d_mask, nx, ny, nz_local, resolution, reinterpret_cast<int32_t*>(d_sdf), true, carries,
                    keep_extrema != 0);
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
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, ctx->d_ext, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream), "copy extrema");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
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
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, ctx->d_ext, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream), "copy extrema");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
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
static int build_host_pipelined(sdf_b200_ctx* ctx, const uint8_t* h_mask, int64_t nx, int64_t ny, int64_t nz, double res,
                                float* out_sdf, double* out_max, double* out_min)
{
    constexpr int NCH = 8;
    constexpr int LUT_N = 1 << 20;
    const int64_t N = nx * ny * nz, plane = ny * nz;
    int rc;
    if ((rc = grow(ctx, &ctx->d_mask, &ctx->mask_bytes, (size_t)N))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)N * 4))) return rc;
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    uint8_t* d_mask = (uint8_t*)ctx->d_mask;
    int32_t* d_field = (int32_t*)ctx->d_out;
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
        CK(cudaMemcpyAsync(d_mask + off, h_mask + off, (size_t)cnt, cudaMemcpyHostToDevice, cs), "H2D mask chunk");
        CK(cudaEventRecord(ctx->chunk_ev[i], cs), "event");
        CK(cudaStreamWaitEvent(st, ctx->chunk_ev[i], 0), "wait");
        CK(sdfb200::launch_binary_z(d_mask + off, d_field + off, (int)nz, (x1 - x0) * ny, nullptr, st), "launch binary_z");
        ctx->launches++;
        if (ny > 1) {
            CK(sdfb200::launch_band_strided(d_field + off, (int)ny, nz, nz, x1 - x0, plane, false, res, ctx->d_ext, nullptr,
                                            nullptr, 0, st), "launch band_strided(y)");
            ctx->launches++;
        }
    }
    for (int j = 0; j < NCH; ++j) {                                 // ---- final x pass + download, chunked along y
        const int64_t y0 = ny * j / NCH, y1 = ny * (j + 1) / NCH;
        if (y1 == y0) continue;
        const int64_t col0 = y0 * nz, ncols = (y1 - y0) * nz;
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
    return sdf_b200_get_extrema(ctx, out_max, out_min);
}

int sdf_b200_build_from_mask(sdf_b200_ctx* ctx, const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz,
                             double resolution, float* out_sdf, double* out_max, double* out_min)
{
    if (ctx && mask && out_sdf && check_dims(ctx, nx, ny, nz) == SDF_B200_OK && nx >= 64 && ny >= 64 &&
        nx * ny * nz >= ((int64_t)1 << 25)) {           // big enough for the copies to be worth overlapping
        CK(cudaSetDevice(ctx->device), "cudaSetDevice");
        return build_host_pipelined(ctx, mask, nx, ny, nz, resolution, out_sdf, out_max, out_min);
    }
    return build_host_common(ctx, mask, (size_t)(nx * ny * nz), false, 0, nx, ny, nz, resolution, out_sdf, out_max, out_min);
}

int sdf_b200_build_from_cells(sdf_b200_ctx* ctx, const void* cells, int unknown_is_filled, int64_t nx, int64_t ny,
                              int64_t nz, double resolution, float* out_sdf, double* out_max, double* out_min)
{
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

int sdf_b200_debug_distance_values(sdf_b200_ctx* ctx, int32_t m0, int32_t count, double resolution, float* out)
{
    if (!ctx || count < 0 || (count > 0 && !out)) return SDF_B200_ERR_ARG;
    if (count == 0) return SDF_B200_OK;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int rc = grow(ctx, &ctx->d_out, &ctx->out_bytes, (size_t)count * 4);
    if (rc) return rc;
    CK(sdfb200::launch_sqrt_check(m0, count, resolution, (float*)ctx->d_out, ctx->stream), "launch sqrt_check");
    CK(cudaMemcpyAsync(out, ctx->d_out, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream), "D2H values");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    return SDF_B200_OK;
}

}  // extern "C"
