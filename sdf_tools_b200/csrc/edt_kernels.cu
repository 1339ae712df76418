// sdf_tools_b200/csrc/edt_kernels.cu -- sm_100a kernels of the signed-distance-field build and their launchers.
//
// Replaces, per call of sdf_generation::ExtractSignedDistanceField<T>
// (/root/reference/include/sdf_tools/sdf_generation.hpp:209-271):
//   K0 pack_cells      is_filled rule                         collision_map.hpp:686-709 (+ classify loop :219-240)
//   K1 binary_z        first third of BuildDistanceField x2   sdf_generation.hpp:95-207, on the binary mask (z lines)
//   K2 band_strided    middle third (y axis)                  same
//   K3 band_strided    last third (x axis)                    same, + the combine loop :246-268 fused in
// All three passes work IN PLACE on one 4-byte-per-voxel field (signed int32 d^2, finally f32 distance), so a
// build touches  1 B (mask read) + 4 B + 4+4 B + 4+4 B = 21 B per voxel of HBM traffic and needs no scratch.
//
// Grid layout: idx = x*(ny*nz) + y*nz + z.  Pass order z -> y -> x: the cheap binary pass takes the contiguous
// axis (and is the axis slabs are cut along for multi-GPU builds: only per-line carries cross the cut); the two
// envelope passes then both run along strided axes, where 32 adjacent lines are 32 adjacent addresses.
#include "edt_kernels.cuh"
#include "band_hull.cuh"

#include <cuda.h>            // CUtensorMap + enums only; cuTensorMapEncodeTiled is fetched at run time (no -lcuda)
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
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

// TAGGED_OBJECT_COLLISION_CELL {float occupancy; u32 component; u32 object_id; u32 convex_segment} (16 B, reference
// tagged_object_collision_map.hpp:22-43) -> one byte per voxel whose bit q answers query q (up to 8 per pass):
//   kind 0: filled and object_id == id   (ExtractSignedDistanceField(objects_to_use = {id}), :813-856; MakeObjectSDFs :875-915)
//   kind 1: filled                        (objects_to_use empty; the "free" field of :733-753)
//   kind 2: filled and object_id > 0      (the "named objects" field of :755-778)
// with filled = occupancy > 0.5 || (unknown_is_filled && occupancy == 0.5).  The binary pass then reads bit q through
// its selector, so one read of the 16-byte cells serves 8 object SDFs.
struct TaggedQueries { uint32_t id[8]; uint32_t kind[8]; int n; };
__global__ void __launch_bounds__(256) pack_tagged_kernel(const uint4* __restrict__ cells, int64_t n, int unknown_is_filled,
                                                          TaggedQueries q, uint8_t* __restrict__ mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 c = __ldg(cells + i);
    const float occ = __uint_as_float(c.x);
    const bool filled = (occ > 0.5f) || (unknown_is_filled && occ == 0.5f);
    uint32_t bits = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool hit = q.kind[k] == 0 ? (c.z == q.id[k]) : (q.kind[k] == 1 ? true : c.z > 0u);
        bits |= (uint32_t)(k < q.n && filled && hit) << k;
    }
    mask[i] = (uint8_t)bits;
}

// ============================================================================================ K1 binary z
// Binary pass along the contiguous z axis: s(z) = +-min(z-a+1, e-z)^2 with [a,e) the run of equal polarity around z
// (squared distance to the nearest voxel of the other polarity on the line; SDF_INF if there is none).
// A group of LPL lanes owns one line (nz <= 32*LPL), each lane one 32-position word: the lane loads its own 32
// occupancy bytes (2 x 16 B, or byte-wise when unaligned), packs them into a bit word, run boundaries come from bit
// operations plus two group-level carries (ballot + shuffle), every lane then walks its 32 positions with two
// counters, and the warp's 32x32 block is transposed through a padded shared tile so rows leave as 128-byte stores.
// carries (slab builds): per line {distance from z=0 / z=nz-1 to the nearest other-polarity voxel in the slabs below /
// above}, FARB = none.
__device__ __forceinline__ uint32_t pack_nonzero_bytes16(uint4 x, uint32_t sel)
{
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
    uint32_t bits = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t t = __vcmpne4(xs[q] & sel, 0u) & 0x01010101u;    // 1 per byte with a selected bit set
        bits |= (((t * 0x01020408u) >> 24) & 0xFu) << (4 * q);
    }
    return bits;
}

template <int LPL, int WPL>   // lanes per line; words per lane (WPL = 2 only with LPL = 32, for nz in (1024, 2048])
__global__ void __launch_bounds__(256) binary_z_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ out,
                                                       int nz, int64_t nlines, const int2* __restrict__ carries, uint32_t sel)
{   // sel: which bits of a mask byte mean FILLED, replicated in all four bytes (0xffffffff: any; 0x01010101 << k: bit k
    // of a multi-object mask, see pack_tagged_kernel)
    static_assert(WPL == 1 || LPL == 32, "two words per lane only for full-warp lines");
    __shared__ int32_t tile_all[8][32 * 33];
    constexpr int LNW = 32 / LPL;               // lines per warp
    constexpr int FARB = 40000;                 // "no zero site on this side"
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane & (LPL - 1);             // word index inside the line (first block)
    const int gbase = lane - g;                 // first lane of this line's group
    const uint32_t gmask = (LPL == 32) ? 0xffffffffu : (((1u << LPL) - 1u) << gbase);
    const int64_t line = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wib) * LNW + (lane / LPL);
    const bool line_ok = line < nlines;
    int32_t* tile = tile_all[wib];
    const uint8_t* mp = mask + (line_ok ? line : 0) * nz;
    const int2 cy = (carries && line_ok) ? __ldg(carries + line) : make_int2(FARB, FARB);
    const bool vec16 = ((nz & 15) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15) == 0);   // warp-uniform

    uint32_t P[WPL], PC[WPL];
    int a_in[WPL], e_in[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
        const int z0 = (k * 32 + g) * 32;
        uint32_t p = 0;
        if (line_ok && z0 < nz) {
            if (vec16) {
                p = pack_nonzero_bytes16(__ldg(reinterpret_cast<const uint4*>(mp + z0)), sel);
                if (z0 + 16 < nz) p |= pack_nonzero_bytes16(__ldg(reinterpret_cast<const uint4*>(mp + z0 + 16)), sel) << 16;
            } else {
                for (int j = 0; j < 32 && z0 + j < nz; ++j) p |= (uint32_t)((__ldg(mp + z0 + j) & sel) != 0) << j;
            }
        }
        P[k] = p;
    }
    // run-start bits: polarity differs from the previous position
    uint32_t carry_top = __shfl_sync(0xffffffffu, P[0] & 1u, gbase);   // position 0 never starts a "new" run
    int last_run_start = -1;                    // carried across the WPL blocks
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
        const int w = k * 32 + g;
        uint32_t prevtop = __shfl_up_sync(0xffffffffu, P[k] >> 31, 1);
        if (g == 0) prevtop = carry_top;
        carry_top = __shfl_sync(0xffffffffu, P[k] >> 31, gbase + LPL - 1);
        uint32_t pc = P[k] ^ ((P[k] << 1) | prevtop);
        const int valid = nz - w * 32;
        if (valid < 32) pc &= (valid > 0) ? ((1u << valid) - 1u) : 0u;
        PC[k] = pc;
        const uint32_t bal = __ballot_sync(0xffffffffu, pc != 0) & gmask;
        const int myhi = pc ? (w * 32 + 31 - __clz((int)pc)) : -1;
        const uint32_t lower = bal & ((1u << lane) - 1u);
        const int src = lower ? (31 - __clz((int)lower)) : 0;
        const int from = __shfl_sync(0xffffffffu, myhi, src);
        a_in[k] = lower ? from : last_run_start;
        const int blk_last = __shfl_sync(0xffffffffu, myhi, bal ? (31 - __clz((int)bal)) : 0);
        if (bal) last_run_start = blk_last;
    }
    int next_run_start = nz;                    // carried backwards across the blocks
#pragma unroll
    for (int k = WPL - 1; k >= 0; --k) {
        const int w = k * 32 + g;
        const uint32_t pc = PC[k];
        const uint32_t bal = __ballot_sync(0xffffffffu, pc != 0) & gmask;
        const int mylo = pc ? (w * 32 + __ffs((int)pc) - 1) : nz;
        const uint32_t upper = (lane < 31) ? (bal & ~((2u << lane) - 1u)) : 0u;
        const int src = upper ? (__ffs((int)upper) - 1) : 0;
        const int from = __shfl_sync(0xffffffffu, mylo, src);
        e_in[k] = upper ? from : next_run_start;
        const int blk_first = __shfl_sync(0xffffffffu, mylo, bal ? (__ffs((int)bal) - 1) : 0);
        if (bal) next_run_start = blk_first;
    }
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
        const int z0 = (k * 32 + g) * 32;
        const uint32_t p = P[k], pc = PC[k];
        int dl = (a_in[k] >= 1) ? (z0 - a_in[k] + 1) : (cy.x < FARB ? z0 + cy.x : FARB);       // z - a + 1
        int e = (pc & ~1u) ? (z0 + __ffs((int)(pc & ~1u)) - 1) : e_in[k];                      // first run start after z0
        int dr = (e < nz) ? (e - z0) : (cy.y < FARB ? nz - 1 - z0 + cy.y : FARB);              // e - z
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if ((pc >> j) & 1u) {               // a run starts here
                dl = 1;
                const uint32_t rest = (j < 31) ? (pc >> (j + 1)) : 0u;
                e = rest ? (z0 + j + __ffs((int)rest)) : e_in[k];
                dr = (e < nz) ? (e - z0 - j) : (cy.y < FARB ? nz - 1 - z0 - j + cy.y : FARB);
            }
            const int d = dl < dr ? dl : dr;
            const int32_t v = (d >= FARB) ? SDF_INF : d * d;
            tile[lane * 33 + j] = ((p >> j) & 1u) ? -v : v;
            ++dl; --dr;
        }
        __syncwarp();
        // row r of the tile = word (r % LPL) of the warp's line (r / LPL)
        const int64_t line0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wib) * LNW;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const int64_t ln = line0 + r / LPL;
            const int z = (k * 32 + (r & (LPL - 1))) * 32 + lane;
            if (ln < nlines && z < nz) out[ln * nz + z] = tile[r * 33 + lane];
        }
        __syncwarp();
    }
}

// ===================================================================================== slab carries (multi-GPU)
// The grid is cut into slabs along z (one per GPU).  The binary pass is the only one that crosses the cut, and all
// it needs from the other slabs is, per z line, the distance from the slab face to the nearest voxel of the other
// polarity.  Each rank publishes per line  {polarity of its first voxel, length of the leading run,
// polarity of its last voxel, length of the trailing run}  (4 bytes); after one all-gather every rank resolves its
// own carries locally, also across slabs that are a single run.
__global__ void __launch_bounds__(256) slab_summary_kernel(const uint8_t* __restrict__ mask, int nz, int64_t nlines,
                                                           uint32_t* __restrict__ summary)
{
    const int lane = threadIdx.x & 31;
    const int64_t line = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (line >= nlines) return;
    const uint8_t* mp = mask + line * nz;
    const bool first = __ldg(mp) != 0, last = __ldg(mp + nz - 1) != 0;
    int lead = nz, trail = nz;                        // warp-wide min of the first / last position that differs
    for (int z = lane; z < nz; z += 32) {
        const bool f = __ldg(mp + z) != 0;
        if (f != first && z < lead) lead = z;
        if (f != last && (nz - 1 - z) < trail) trail = nz - 1 - z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lead = min(lead, __shfl_xor_sync(0xffffffffu, lead, o));
        trail = min(trail, __shfl_xor_sync(0xffffffffu, trail, o));
    }
    if (lane == 0)
        summary[line] = (uint32_t)lead | ((uint32_t)first << 15) | ((uint32_t)trail << 16) | ((uint32_t)last << 31);
}

// The same summary, fused with its exchange: the word of every line is stored straight into slot `rank` of the gather
// buffer of EVERY rank (dst.p[k] = rank k's buffer, peer memory mapped over NVLink: cudaDeviceEnablePeerAccess, CUDA IPC
// or symmetric memory), so step 2 of the protocol (the all-gather) costs no extra launch and no collective call.
struct PeerPtrs { uint32_t* p[SDF_MAX_PEERS]; };
__global__ void __launch_bounds__(256) slab_summary_scatter_kernel(const uint8_t* __restrict__ mask, int nz, int64_t nlines,
                                                                   PeerPtrs dst, int nranks, int rank)
{
    constexpr int LINES = 64;                         // lines per block: remote stores leave as whole 128-byte rows
    __shared__ uint32_t words[LINES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t line0 = (int64_t)blockIdx.x * LINES;
    for (int l = warp; l < LINES; l += 8) {
        const int64_t line = line0 + l;
        if (line >= nlines) break;
        const uint8_t* mp = mask + line * nz;
        const bool first = __ldg(mp) != 0, last = __ldg(mp + nz - 1) != 0;
        int lead = nz, trail = nz;
        for (int z = lane; z < nz; z += 32) {
            const bool f = __ldg(mp + z) != 0;
            if (f != first && z < lead) lead = z;
            if (f != last && (nz - 1 - z) < trail) trail = nz - 1 - z;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lead = min(lead, __shfl_xor_sync(0xffffffffu, lead, o));
            trail = min(trail, __shfl_xor_sync(0xffffffffu, trail, o));
        }
        if (lane == 0)
            words[l] = (uint32_t)lead | ((uint32_t)first << 15) | ((uint32_t)trail << 16) | ((uint32_t)last << 31);
    }
    __syncthreads();
    for (int k = warp; k < nranks; k += 8) {          // one warp per destination rank
        uint32_t* out = dst.p[k] + (int64_t)rank * nlines + line0;
        for (int l = lane; l < LINES && line0 + l < nlines; l += 32) out[l] = words[l];
    }
}

// all[r * nlines + line] = summary of rank r; widths[r] = slab thickness of rank r.
__global__ void __launch_bounds__(256) slab_carry_kernel(const uint32_t* __restrict__ all, const int* __restrict__ widths,
                                                         int nranks, int rank, int64_t nlines, int2* __restrict__ carries)
{
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= nlines) return;
    constexpr int FARB = 40000;
    const uint32_t me = all[(int64_t)rank * nlines + line];
    const bool myfirst = (me >> 15) & 1u, mylast = (me >> 31) & 1u;
    int cl = FARB, cr = FARB, acc = 0;
    for (int k = rank - 1; k >= 0; --k) {             // towards lower z
        const uint32_t s = all[(int64_t)k * nlines + line];
        const bool klast = (s >> 31) & 1u;
        const int ktrail = (int)((s >> 16) & 0x7fffu), w = widths[k];
        if (klast != myfirst) { cl = acc + 1; break; }
        if (ktrail < w) { cl = acc + ktrail + 1; break; }
        acc += w;
    }
    acc = 0;
    for (int k = rank + 1; k < nranks; ++k) {         // towards higher z
        const uint32_t s = all[(int64_t)k * nlines + line];
        const bool kfirst = (s >> 15) & 1u;
        const int klead = (int)(s & 0x7fffu), w = widths[k];
        if (kfirst != mylast) { cr = acc + 1; break; }
        if (klead < w) { cr = acc + klead + 1; break; }
        acc += w;
    }
    carries[line] = make_int2(cl < FARB ? cl : FARB, cr < FARB ? cr : FARB);
}

// ========================================================================================= K2/K3 band hull
// Envelope pass along a strided axis (see band_hull.cuh): tile = LPC adjacent columns x n rows; thread (c, b)
// owns band b (B rows) of column c; lanes of a warp = adjacent columns.  In place.  FINAL fuses the reference's
// combine loop (sdf_generation.hpp:246-268): (float)(sqrt((double)d2) * res) with the polarity sign, plus the
// per-polarity maxima of d2 for the (max, min) pair returned at :269.
__device__ __forceinline__ float sdf_from_signed_d2(int32_t s, double res)
{
    const int32_t m = s < 0 ? -s : s;
    const float f = (m >= SDF_INF) ? __int_as_float(0x7f800000) : (float)(sqrt((double)m) * res);
    return s < 0 ? -f : f;
}

// lut[m] = (float)(sqrt((double)m) * res) for m < n: the exact value the combine loop stores (sdf_generation.hpp:254-265),
// tabulated once per resolution so the final pass replaces ~25 FP64 issue slots per voxel by one cached 4-byte load.
__global__ void __launch_bounds__(256) sqrt_lut_kernel(float* __restrict__ lut, int n, double res)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n) lut[m] = (float)(sqrt((double)m) * res);
}

template <int LOGB, int LPC, int NBMAX, bool FINAL>   // NBMAX = most bands per line this instance is launched with
__global__ void __launch_bounds__(NBMAX * LPC, (NBMAX * LPC == 512) ? 3 : 1) band_strided_kernel(int32_t* __restrict__ g, int n, int nb,
                                                                 int64_t line_stride, int64_t ncols,
                                                                 int64_t outer_stride, double res, int* __restrict__ ext,
                                                                 long long* __restrict__ dbg,
                                                                 const float* __restrict__ lut, int lut_n,
                                                                 const __grid_constant__ CUtensorMap tmap, int tma_rows)
{
    using BH = BandHull<LOGB, LPC>;
    // optional phase clocks (diagnostics): thread 0 of the first 1024 CTAs stores clock64() at phase boundaries
    const unsigned dbg_lin = blockIdx.y * gridDim.x + blockIdx.x;
    const unsigned dbg_stride = (gridDim.x * gridDim.y + 1023u) / 1024u;            // sample ~1024 CTAs evenly
    const int dbg_cta = (int)(dbg_lin / dbg_stride);
    const bool dbg_on = dbg != nullptr && threadIdx.x == 0 && (dbg_lin % dbg_stride) == 0 && dbg_cta < 1024;
    if (dbg_on) dbg[dbg_cta * 8 + 0] = clock64();
    constexpr int B = BH::B;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long tma_bar;
    BandTile T;
    {
        uint32_t p = (uint32_t)__cvta_generic_to_shared(smem_raw);
        T.S = p;      p += (uint32_t)nb * B * LPC * 4;
        T.meta = p;   p += (uint32_t)nb * LPC * 4;
        T.pcA = p;    p += (uint32_t)nb * LPC * 2;
        T.pcE = p;    p += (uint32_t)nb * LPC * 2;
        T.anyw = p;
    }
    const int c = threadIdx.x % LPC, b = threadIdx.x / LPC;
    if (b < (nb + 31) / 32) bh_st32(T.anyw + 4u * (uint32_t)(b * LPC + c), 0u);   // ordered before build by the barrier below
    const int64_t col = (int64_t)blockIdx.x * LPC + c;
    const bool col_ok = col < ncols;
    int32_t* gc = g + (int64_t)blockIdx.y * outer_stride + col;
    const int base = b << LOGB;
    int nlevels = 0;
    while ((1 << nlevels) < nb) ++nlevels;

    typename BH::pcw_t Pw, PCw;
    if (tma_rows) {
        // Tile staging by TMA: 3-D tensor-map boxes (LPC columns x tma_rows rows x 1, whole bands) land directly in the bands'
        // stack slots (the slot layout [row][line] IS the dense box layout); rows / columns outside the grid arrive as
        // zeros and are never read.  One thread arms an mbarrier with the byte count and issues the copies; everybody
        // waits on the barrier's phase.
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)nb * B * LPC * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            const int c0 = (int)(blockIdx.x * LPC), o = (int)blockIdx.y;
            for (int r0 = 0; r0 < nb * B; r0 += tma_rows) {       // boxes of LPC columns x tma_rows rows (whole bands)
                const uint32_t dstp = T.S + (uint32_t)r0 * (LPC * 4u);
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             :: "r"(dstp), "l"(&tmap), "r"(c0), "r"(r0), "r"(o), "r"(bar) : "memory");
            }
        }
        uint32_t done = 0;
        int spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
            if (!done && ++spins > (1 << 22)) {               // never hang the device on a mis-programmed copy: flag the build
                atomicOr(ext + 2, 1);                         // as failed (ext[2], reported by sdf_b200_get_extrema) and go on
                break;                                        // with whatever the tile holds -- no index depends on tile data
            }
        }
        // polarity of the position just before the band, read from the staged tile before any stack overwrites it
        const int qp = b > 0 ? base - 1 : 0;
        const bool prevpol = col_ok ? ((int32_t)bh_ld32(T.S + 4u * (uint32_t)(qp * LPC + c)) < 0) : false;
        if (dbg_on) dbg[dbg_cta * 8 + 1] = clock64();
        __syncthreads();
        BH::build(T, c, b, col_ok ? n : 0, prevpol, Pw, PCw);
    } else {
        // stage the band's values into its own stack slots (the hull is then built in place)
        const bh_addr_t slot0 = T.S + 4u * (uint32_t)BH::sidx(b, 0, c);
        constexpr int U = 8;
#pragma unroll
        for (int i0 = 0; i0 < B; i0 += U) {
            int32_t tmp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + i0 + u;
                tmp[u] = (col_ok && q < n) ? gc[(int64_t)q * line_stride] : SDF_INF;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) bh_st32(slot0 + (uint32_t)((i0 + u) * (LPC * 4)), (uint32_t)tmp[u]);
        }
        bool prevpol = false;
        if (col_ok) {
            const int qp = b > 0 ? base - 1 : 0;
            prevpol = gc[(int64_t)qp * line_stride] < 0;
        }
        if (dbg_on) dbg[dbg_cta * 8 + 1] = clock64();
        __syncthreads();
        BH::build(T, c, b, col_ok ? n : 0, prevpol, Pw, PCw);
    }
    if (dbg_on) dbg[dbg_cta * 8 + 2] = clock64();
    // Seam round 0 joins bands 2i and 2i+1: with <= 16 lines per CTA both were built by the same warp, so a warp-level
    // barrier orders it after the build and one CTA-wide barrier is saved.
    constexpr bool ROUND0_IN_WARP = (2 * LPC <= 32);
    if (ROUND0_IN_WARP && nlevels > 0) {
        __syncwarp();
        BH::merge(T, c, b, nb, 0);
    }
    __syncthreads();
    if (dbg_on) dbg[dbg_cta * 8 + 3] = clock64();
    for (int lv = ROUND0_IN_WARP ? 1 : 0; lv < nlevels; ++lv) {
        BH::merge(T, c, b, nb, lv);
        __syncthreads();
    }
    if (dbg_on) dbg[dbg_cta * 8 + 4] = clock64();
    int32_t maxpos = 0, maxneg = 0;
    if (col_ok) {
        int32_t* dst = gc + (int64_t)base * line_stride;
        BH::query(T, c, b, nb, n, nlevels, Pw, PCw, [&](int i, int32_t v) {
            (void)i;
            if (FINAL) {
                const int32_t m = v < 0 ? -v : v;
                if (v < 0) maxneg = max(maxneg, m); else maxpos = max(maxpos, m);
                float f;
                if (m < lut_n) f = __ldg(lut + m);
                else f = (m >= SDF_INF) ? __int_as_float(0x7f800000) : (float)(sqrt((double)m) * res);
                *dst = __float_as_int(v < 0 ? -f : f);
            } else {
                *dst = v;
            }
            dst += line_stride;
        });
    }
    if (dbg_on) dbg[dbg_cta * 8 + 5] = clock64();
    if (FINAL) {
        // reduce over the lanes that exist: blockDim = nb * LPC need not be a multiple of 32 (e.g. n = 1500: 376 threads)
        const unsigned live = __activemask();
        maxpos = __reduce_max_sync(live, maxpos);
        maxneg = __reduce_max_sync(live, maxneg);
        if ((int)(threadIdx.x & 31) == __ffs((int)live) - 1) {
            if (maxpos > 0) atomicMax(ext + 0, maxpos);
            if (maxneg > 0) atomicMax(ext + 1, maxneg);
        }
    }
}

// ============================================================================================== virtual border
// add_virtual_border (reference sdf_generation.hpp:287-419): axes longer than one cell grow by a one-cell shell that is
// FILLED in the build that supplies the free-space values and FREE in the build that supplies the inside values.
// pad: inner mask (nx,ny,nz) -> padded mask (nx+2ox, ny+2oy, nz+2oz) with the shell set to `shell`;
// select: per real voxel  a >= 0 ? a : (b <= -0 ? b : 0)  from the two padded fields (:381-409).
__global__ void __launch_bounds__(256) border_pad_kernel(const uint8_t* __restrict__ inner, int64_t nx, int64_t ny, int64_t nz,
                                                         int ox, int oy, int oz, uint8_t shell, uint8_t* __restrict__ padded)
{
    const int64_t bx = nx + 2 * ox, by = ny + 2 * oy, bz = nz + 2 * oz;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bx * by * bz) return;
    const int64_t z = i % bz - oz, y = (i / bz) % by - oy, x = i / (bz * by) - ox;
    const bool in = x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz;
    padded[i] = in ? (uint8_t)(__ldg(inner + (x * ny + y) * nz + z) != 0) : shell;
}

__global__ void __launch_bounds__(256) border_select_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t nx,
                                                            int64_t ny, int64_t nz, int ox, int oy, int oz, float* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx * ny * nz) return;
    const int64_t z = i % nz, y = (i / nz) % ny, x = i / (nz * ny);
    const int64_t by = ny + 2 * oy, bz = nz + 2 * oz;
    const int64_t q = ((x + ox) * by + (y + oy)) * bz + (z + oz);
    const float fa = __ldg(a + q), fb = __ldg(b + q);
    out[i] = (fa >= 0.0f) ? fa : ((fb <= -0.0f) ? fb : 0.0f);
}

cudaError_t launch_border_pad(const uint8_t* d_inner, int64_t nx, int64_t ny, int64_t nz, int ox, int oy, int oz, int shell,
                              uint8_t* d_padded, cudaStream_t st)
{
    const int64_t n = (nx + 2 * ox) * (ny + 2 * oy) * (nz + 2 * oz);
    if (n <= 0) return cudaSuccess;
    border_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_inner, nx, ny, nz, ox, oy, oz, (uint8_t)(shell != 0), d_padded);
    return cudaGetLastError();
}

cudaError_t launch_border_select(const float* d_a, const float* d_b, int64_t nx, int64_t ny, int64_t nz, int ox, int oy, int oz,
                                 float* d_out, cudaStream_t st)
{
    const int64_t n = nx * ny * nz;
    if (n <= 0) return cudaSuccess;
    border_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, d_b, nx, ny, nz, ox, oy, oz, d_out);
    return cudaGetLastError();
}

// ============================================================================================== gradient field
// Grid-aligned gradient of a finished SDF, one thread per voxel: SignedDistanceField::GetGridAlignedGradient
// (/root/reference/include/sdf_tools/sdf.hpp:432-526) evaluated for every voxel as GetFullGradient does (:341-358).
// Arithmetic order of the reference, so results are bit-identical to its doubles: interior voxels subtract the two
// float neighbours IN FLOAT and scale by the double 1/(2 res) (:445-453); face voxels (enable_edge_gradients) widen both
// values to double first and scale by 1/((hi-lo) res) (:486-518); an axis of one cell gives 0.  Voxels for which the
// reference returns an empty vector (faces without edge gradients) get NaN in all three components.
template <typename OUT>
__global__ void __launch_bounds__(256) gradient_kernel(const float* __restrict__ sdf, int64_t nx, int64_t ny, int64_t nz,
                                                       double res, double inv_twice_res, int edges, OUT* __restrict__ grad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx * ny * nz) return;
    const int64_t z = i % nz, y = (i / nz) % ny, x = i / (nz * ny);
    const int64_t sx = ny * nz, sy = nz;
    double g[3];
    const bool interior = x > 0 && y > 0 && z > 0 && x < nx - 1 && y < ny - 1 && z < nz - 1;
    if (interior) {
        g[0] = __dmul_rn((double)__fsub_rn(__ldg(sdf + i + sx), __ldg(sdf + i - sx)), inv_twice_res);
        g[1] = __dmul_rn((double)__fsub_rn(__ldg(sdf + i + sy), __ldg(sdf + i - sy)), inv_twice_res);
        g[2] = __dmul_rn((double)__fsub_rn(__ldg(sdf + i + 1), __ldg(sdf + i - 1)), inv_twice_res);
    } else if (edges) {
        const int64_t pos[3] = {x, y, z}, dim[3] = {nx, ny, nz}, str[3] = {sx, sy, 1};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int64_t lo = pos[a] > 0 ? pos[a] - 1 : 0, hi = pos[a] < dim[a] - 1 ? pos[a] + 1 : dim[a] - 1;
            const double inc = __dmul_rn((double)(hi - lo), res);
            g[a] = 0.0;
            if (inc > 0.0) {
                const double vhi = (double)__ldg(sdf + i + (hi - pos[a]) * str[a]);
                const double vlo = (double)__ldg(sdf + i + (lo - pos[a]) * str[a]);
                g[a] = __dmul_rn(__dsub_rn(vhi, vlo), __ddiv_rn(1.0, inc));
            }
        }
    } else {
        g[0] = g[1] = g[2] = __longlong_as_double(0x7ff8000000000000LL);
    }
    grad[3 * i + 0] = (OUT)g[0];
    grad[3 * i + 1] = (OUT)g[1];
    grad[3 * i + 2] = (OUT)g[2];
}

cudaError_t launch_gradient(const float* d_sdf, int64_t nx, int64_t ny, int64_t nz, double res, int edges, int out_f64,
                            void* d_grad, cudaStream_t st)
{
    const int64_t n = nx * ny * nz;
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    const double inv2 = 1.0 / (2.0 * res);
    if (out_f64) gradient_kernel<double><<<blocks, 256, 0, st>>>(d_sdf, nx, ny, nz, res, inv2, edges, (double*)d_grad);
    else gradient_kernel<float><<<blocks, 256, 0, st>>>(d_sdf, nx, ny, nz, res, inv2, edges, (float*)d_grad);
    return cudaGetLastError();
}

// ================================================================================================ launchers
cudaError_t launch_pack_cells(const void* d_cells, int64_t n, int unknown_is_filled, uint8_t* d_mask, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int64_t threads = (n + 3) / 4;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    pack_cells_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float2*>(d_cells), n, unknown_is_filled, d_mask);
    return cudaGetLastError();
}

cudaError_t launch_pack_tagged(const void* d_cells16, int64_t n, int unknown_is_filled, const uint32_t* ids,
                               const uint32_t* kinds, int nq, uint8_t* d_mask, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    if (nq < 1 || nq > 8) return cudaErrorInvalidValue;
    TaggedQueries q;
    for (int k = 0; k < 8; ++k) { q.id[k] = k < nq ? ids[k] : 0u; q.kind[k] = k < nq ? kinds[k] : 0u; }
    q.n = nq;
    pack_tagged_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(d_cells16), n,
                                                                    unknown_is_filled, q, d_mask);
    return cudaGetLastError();
}

cudaError_t launch_sqrt_lut(float* d_lut, int n, double res, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    sqrt_lut_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_lut, n, res);
    return cudaGetLastError();
}

template <int LPL, int WPL>
static cudaError_t launch_binary_z_cfg(const uint8_t* m, int32_t* out, int nz, int64_t nlines, const int2* cy, uint32_t sel, cudaStream_t st)
{
    const int wpb = 8;
    const int64_t lines_per_block = (int64_t)wpb * (32 / LPL);
    binary_z_kernel<LPL, WPL><<<(unsigned)((nlines + lines_per_block - 1) / lines_per_block), wpb * 32, 0, st>>>(m, out, nz, nlines, cy, sel);
    return cudaGetLastError();
}

cudaError_t launch_binary_z(const uint8_t* d_mask, int32_t* d_out, int nz, int64_t nlines, const void* d_carries,
                            cudaStream_t st, unsigned mask_bits)
{
    const int2* cy = reinterpret_cast<const int2*>(d_carries);
    const uint32_t sel = (mask_bits & 0xffu) * 0x01010101u;
    if (nz <= 32) return launch_binary_z_cfg<1, 1>(d_mask, d_out, nz, nlines, cy, sel, st);
    if (nz <= 64) return launch_binary_z_cfg<2, 1>(d_mask, d_out, nz, nlines, cy, sel, st);
    if (nz <= 128) return launch_binary_z_cfg<4, 1>(d_mask, d_out, nz, nlines, cy, sel, st);
    if (nz <= 256) return launch_binary_z_cfg<8, 1>(d_mask, d_out, nz, nlines, cy, sel, st);
    if (nz <= 512) return launch_binary_z_cfg<16, 1>(d_mask, d_out, nz, nlines, cy, sel, st);
    if (nz <= 1024) return launch_binary_z_cfg<32, 1>(d_mask, d_out, nz, nlines, cy, sel, st);
    if (nz <= 2048) return launch_binary_z_cfg<32, 2>(d_mask, d_out, nz, nlines, cy, sel, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_slab_summary(const uint8_t* d_mask, int nz, int64_t nlines, uint32_t* d_summary, cudaStream_t st)
{
    const int wpb = 8;
    slab_summary_kernel<<<(unsigned)((nlines + wpb - 1) / wpb), wpb * 32, 0, st>>>(d_mask, nz, nlines, d_summary);
    return cudaGetLastError();
}

cudaError_t launch_slab_summary_scatter(const uint8_t* d_mask, int nz, int64_t nlines, void* const* peer_gather, int nranks,
                                        int rank, cudaStream_t st)
{
    if (nranks < 1 || nranks > SDF_MAX_PEERS) return cudaErrorInvalidValue;
    PeerPtrs dst;
    for (int k = 0; k < SDF_MAX_PEERS; ++k) dst.p[k] = k < nranks ? reinterpret_cast<uint32_t*>(peer_gather[k]) : nullptr;
    slab_summary_scatter_kernel<<<(unsigned)((nlines + 63) / 64), 256, 0, st>>>(d_mask, nz, nlines, dst, nranks, rank);
    return cudaGetLastError();
}

cudaError_t launch_slab_carry(const uint32_t* d_all, const int* d_widths, int nranks, int rank, int64_t nlines,
                              void* d_carries, cudaStream_t st)
{
    slab_carry_kernel<<<(unsigned)((nlines + 255) / 256), 256, 0, st>>>(d_all, d_widths, nranks, rank, nlines,
                                                                        reinterpret_cast<int2*>(d_carries));
    return cudaGetLastError();
}

typedef CUresult (*tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_fn tensor_map_encoder()
{
    static tensor_map_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tensor_map_encode_fn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

template <int LOGB, int LPC, int NBMAX, bool FINAL>
static cudaError_t launch_band_cfg(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                   int64_t outer_stride, double res, int* ext, long long* dbg, const float* lut, int lut_n,
                                   cudaStream_t st)
{
    constexpr int B = 1 << LOGB;
    const int nb = (n + B - 1) / B;
    if (nb > NBMAX) return cudaErrorInvalidValue;
    const size_t smem = (size_t)nb * B * LPC * 4 + (size_t)nb * LPC * 4 + (size_t)nb * LPC * 4 + (size_t)((nb + 31) / 32) * LPC * 4;
    auto kern = band_strided_kernel<LOGB, LPC, NBMAX, FINAL>;
    if (smem > 48 * 1024) {   // per-device attribute; cheap enough to set on every launch
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)((ncols + LPC - 1) / LPC), (unsigned)nouter);
    // 3-D tensor map of the pass's view of the field: (columns, rows along the axis, outer blocks), box = LPC x B x 1.
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int tma_rows = 0;                                 // rows per tensor-map box; 0 = LDG staging
    static int tma_knob = -1;                         // SDF_B200_TMA=0 forces the LDG staging path
    if (tma_knob < 0) { const char* e = getenv("SDF_B200_TMA"); tma_knob = (e && e[0] == '0') ? 0 : 1; }
    const int64_t ostride = nouter > 1 ? outer_stride : line_stride * (int64_t)n;
    if (tma_knob && tensor_map_encoder() && (line_stride % 4) == 0 && (ostride % 4) == 0 &&
        (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const cuuint64_t gdim[3] = {(cuuint64_t)ncols, (cuuint64_t)n, (cuuint64_t)nouter};
        const cuuint64_t gstr[2] = {(cuuint64_t)line_stride * 4u, (cuuint64_t)ostride * 4u};
        int bpb = 1;                                  // bands per box: the largest divisor of nb with <= 256 rows
        for (int d = 2; d * B <= 256; ++d) if (nb % d == 0) bpb = d;
        tma_rows = bpb * B;
        const cuuint32_t box[3] = {(cuuint32_t)LPC, (cuuint32_t)tma_rows, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, g, gdim, gstr, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tma_rows = 0;
    }
    kern<<<grid, nb * LPC, smem, st>>>(g, n, nb, line_stride, ncols, outer_stride, res, ext, dbg, lut, lut_n, tmap, tma_rows);
    return cudaGetLastError();
}

// Tile shapes: bands of 32 rows everywhere; lines per CTA chosen so a CTA has <= 512 threads and ~64 KB of stacks
// (3 CTAs per SM): n <= 256 -> 32 lines x <= 8 bands, n <= 1024 -> 16 lines x <= 32 bands, n <= 2048 -> 8 lines x <= 64 bands.
template <bool FINAL>
static cudaError_t launch_band_f(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                 int64_t outer_stride, double res, int* ext, long long* dbg, const float* lut, int lut_n,
                                 cudaStream_t st)
{
    if (n <= 256) return launch_band_cfg<5, 32, 8, FINAL>(g, n, line_stride, ncols, nouter, outer_stride, res, ext, dbg, lut, lut_n, st);
    if (n <= 1024) return launch_band_cfg<5, 16, 32, FINAL>(g, n, line_stride, ncols, nouter, outer_stride, res, ext, dbg, lut, lut_n, st);
    if (n <= 2048) return launch_band_cfg<5, 8, 64, FINAL>(g, n, line_stride, ncols, nouter, outer_stride, res, ext, dbg, lut, lut_n, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_band_strided(int32_t* g, int n, int64_t line_stride, int64_t ncols, int64_t nouter,
                                int64_t outer_stride, bool final_pass, double res, int* d_ext, long long* d_dbg,
                                const float* d_lut, int lut_n, cudaStream_t st)
{
    return final_pass ? launch_band_f<true>(g, n, line_stride, ncols, nouter, outer_stride, res, d_ext, d_dbg, d_lut, lut_n, st)
                      : launch_band_f<false>(g, n, line_stride, ncols, nouter, outer_stride, res, d_ext, d_dbg, d_lut, lut_n, st);
}

}  // namespace sdfb200
