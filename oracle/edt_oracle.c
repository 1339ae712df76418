/*
 * oracle/edt_oracle.c  --  TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * A plain-C restatement of what the reference's hot path is *meant* to compute:
 *   sdf_generation::ExtractSignedDistanceField<T>   /root/reference/include/sdf_tools/sdf_generation.hpp:209-271
 *     = classify voxels (:219-240) + BuildDistanceField(filled) (:242) + BuildDistanceField(free) (:243)
 *       + per-voxel combine  sqrt(d2_filled)*res - sqrt(d2_free)*res  with running max/min (:246-268).
 *
 * BuildDistanceField (:95-207) is a bucket-queue vector propagation that only *approximates* the Euclidean
 * distance transform (it over-estimates d^2 at 0-1 % of voxels and its errors depend on the seed order,
 * SURVEY.md finding 1).  This oracle computes the EXACT squared EDT with the separable lower-envelope
 * method (Felzenszwalb & Huttenlocher, "Distance Transforms of Sampled Functions", 2012; integer form after
 * Meijster et al. 2000) -- the quantity the propagation converges to wherever it is right.  The real
 * reference text, compiled unmodified, lives in oracle/_ref (see oracle/build_ref.sh) and is the second
 * checker; tests/ pin this file against it (equal on dense grids, reference >= exact elsewhere),
 * against scipy.ndimage.distance_transform_edt, against an O(N^2) brute force and against the reference's
 * own known-answer test (test/test_bindings.py:12-33).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (libsdf_b200.so) never links or calls it.
 *
 * Memory order everywhere: idx = x*(ny*nz) + y*nz + z  (z contiguous), as in the reference's VoxelGrid
 * (utils_3d.py:71-74, sdf.cpp:241-245).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_INF ((int32_t)0x40000000)

/* 1-D transform  out[p] = min_q (p-q)^2 + f[q]  over a strided line; f[q] == ORACLE_INF means "no site".
 * Lower hull of the lifted points (q, q^2+f[q]) with exact integer cross-multiplication, then a monotone walk. */
static void line_envelope(const int32_t* f, int64_t stride, int64_t n, int32_t* out,
                          int32_t* v /*n*/, int64_t* Y /*n*/)
{
    int64_t k = -1;
    for (int64_t q = 0; q < n; q++) {
        const int32_t fq = f[q * stride];
        if (fq >= ORACLE_INF) continue;
        const int64_t Yq = q * q + (int64_t)fq;
        while (k >= 1) {
            /* top is on or above the chord second->new  <=>  it never wins */
            const int64_t lhs = (Y[k] - Y[k - 1]) * (q - v[k]);
            const int64_t rhs = (Yq - Y[k]) * ((int64_t)v[k] - v[k - 1]);
            if (lhs >= rhs) k--; else break;
        }
        k++; v[k] = (int32_t)q; Y[k] = Yq;
    }
    if (k < 0) { for (int64_t p = 0; p < n; p++) out[p * stride] = ORACLE_INF; return; }
    int64_t j = 0;
    for (int64_t p = 0; p < n; p++) {
        for (;;) {
            const int64_t cur = Y[j] - 2 * p * v[j];
            if (j < k && Y[j + 1] - 2 * p * v[j + 1] <= cur) j++; else break;
        }
        out[p * stride] = (int32_t)(Y[j] - 2 * p * v[j] + p * p);
    }
}

/* Exact squared EDT: d2[i] = min over voxels j with site[j]!=0 of |i-j|^2 ; ORACLE_INF when there is no site.
 * Mirrors the *contract* of BuildDistanceField (sdf_generation.hpp:95-100: seeds -> distance_square per voxel). */
int oracle_edt_d2(const uint8_t* site, int64_t nx, int64_t ny, int64_t nz, int32_t* d2, int nthreads)
{
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    const int64_t N = nx * ny * nz;
    int64_t nmax = nx; if (ny > nmax) nmax = ny; if (nz > nmax) nmax = nz;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) d2[i] = site[i] ? 0 : ORACLE_INF;

    #pragma omp parallel
    {
        int32_t* v = (int32_t*)malloc(sizeof(int32_t) * (size_t)nmax);
        int64_t* Y = (int64_t*)malloc(sizeof(int64_t) * (size_t)nmax);
        int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)nmax);
        /* z lines (contiguous) */
        #pragma omp for schedule(static)
        for (int64_t l = 0; l < nx * ny; l++) {
            int32_t* line = d2 + l * nz;
            memcpy(tmp, line, sizeof(int32_t) * (size_t)nz);
            line_envelope(tmp, 1, nz, line, v, Y);
        }
        /* y lines */
        #pragma omp for schedule(static)
        for (int64_t l = 0; l < nx * nz; l++) {
            const int64_t x = l / nz, z = l % nz;
            int32_t* line = d2 + x * ny * nz + z;
            for (int64_t y = 0; y < ny; y++) tmp[y] = line[y * nz];
            line_envelope(tmp, 1, ny, tmp, v, Y);       /* in place is fine: out[p] written after f consumed into Y */
            for (int64_t y = 0; y < ny; y++) line[y * nz] = tmp[y];
        }
        /* x lines */
        #pragma omp for schedule(static)
        for (int64_t l = 0; l < ny * nz; l++) {
            int32_t* line = d2 + l;
            for (int64_t x = 0; x < nx; x++) tmp[x] = line[x * ny * nz];
            line_envelope(tmp, 1, nx, tmp, v, Y);
            for (int64_t x = 0; x < nx; x++) line[x * ny * nz] = tmp[x];
        }
        free(v); free(Y); free(tmp);
    }
    return 0;
}

/* Signed fused field:  +d2(nearest filled) at free voxels, -d2(nearest free) at filled voxels
 * (exactly one of the two distances of sdf_generation.hpp:254-255 is non-zero at every voxel). */
int oracle_signed_d2(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int32_t* s, int nthreads)
{
    const int64_t N = nx * ny * nz;
    if (N <= 0) return 0;
    uint8_t* inv = (uint8_t*)malloc((size_t)N);
    int32_t* b = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    if (!inv || !b) { free(inv); free(b); return -1; }
    for (int64_t i = 0; i < N; i++) inv[i] = mask[i] ? 0 : 1;
    oracle_edt_d2(mask, nx, ny, nz, s, nthreads);   /* A: to nearest filled  (BuildDistanceField(filled), :242) */
    oracle_edt_d2(inv, nx, ny, nz, b, nthreads);    /* B: to nearest free    (BuildDistanceField(free),   :243) */
    for (int64_t i = 0; i < N; i++) if (mask[i]) s[i] = -b[i];
    free(inv); free(b);
    return 0;
}

/* Full build, following sdf_generation.hpp:209-271 line by line with the exact EDT in place of the propagation.
 * distance = sqrt(d2_filled)*res - sqrt(d2_free)*res in f64 (:254-256), stored as f32 (:265),
 * max/min tracked on the f64 value starting from -inf/+inf (:246-247, :257-264). */
int oracle_build_sdf(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, double res,
                     float* sdf, double* out_max, double* out_min, int nthreads)
{
    const int64_t N = nx * ny * nz;
    double mx = -INFINITY, mn = INFINITY;
    if (N > 0) {
        uint8_t* inv = (uint8_t*)malloc((size_t)N);
        int32_t* a = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
        int32_t* b = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
        if (!inv || !a || !b) { free(inv); free(a); free(b); return -1; }
        for (int64_t i = 0; i < N; i++) inv[i] = mask[i] ? 0 : 1;
        oracle_edt_d2(mask, nx, ny, nz, a, nthreads);
        oracle_edt_d2(inv, nx, ny, nz, b, nthreads);
        for (int64_t i = 0; i < N; i++) {
            const double d2a = (a[i] >= ORACLE_INF) ? INFINITY : (double)a[i];
            const double d2b = (b[i] >= ORACLE_INF) ? INFINITY : (double)b[i];
            const double distance1 = sqrt(d2a) * res;
            const double distance2 = sqrt(d2b) * res;
            const double distance = distance1 - distance2;
            if (distance > mx) mx = distance;
            if (distance < mn) mn = distance;
            sdf[i] = (float)distance;
        }
        free(inv); free(a); free(b);
    }
    if (out_max) *out_max = mx;
    if (out_min) *out_min = mn;
    return 0;
}

/* is_filled rule of CollisionMapGrid::ExtractSignedDistanceField (collision_map.hpp:686-709):
 * occupancy > 0.5, or == 0.5 when unknown_is_filled.  cells = {float occupancy; uint32 component} (8 bytes). */
int oracle_pack_cells(const void* cells, int64_t n, int unknown_is_filled, uint8_t* mask)
{
    const float* c = (const float*)cells;
    for (int64_t i = 0; i < n; i++) {
        const float occ = c[2 * i];
        mask[i] = (occ > 0.5f) || (unknown_is_filled && occ == 0.5f);
    }
    return 0;
}

int32_t oracle_inf(void) { return ORACLE_INF; }
