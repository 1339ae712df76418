// sdf_tools_b200/csrc/line_hull.cuh
//
// One 1-D pass of the signed-fused exact Euclidean distance transform along a STRIDED axis, ONE THREAD PER LINE:
// the inner work of sdf_generation::BuildDistanceField (/root/reference/include/sdf_tools/sdf_generation.hpp:95-207,
// called twice at :242-243) for one axis.  The 32 lanes of a warp own 32 ADJACENT lines, so every row of the field is
// read and written as one 128-byte access; each thread runs the classic sequential lower-envelope algorithm
// (Felzenszwalb / Meijster) over its whole line: a build sweep that keeps the hull of the lifted points as a stack, then a
// query sweep that walks the stack.  No bands, no seam merges, no block-wide barriers, no shared memory.
//
// The stack does not fit on chip (up to n entries per line), so it lives in a global scratch: one block of n x 32 entries
// per warp, laid out so that a line's own pushes, pops and fetches stay inside the 16 bytes they just touched ("Stack
// layout" below).  Adjacent lines have similar hulls, so a warp's stack accesses land in a few cache lines that stay in
// L1 / the 126 MB L2; only the long-lived part of a hull spills to HBM.
//
// Signed-fused field: s > 0 <=> voxel FREE, s = d^2 to the nearest FILLED; s < 0 <=> voxel FILLED, -s = d^2 to the
// nearest FREE (partial over the axes done so far; |s| == SDF_INF: other polarity not seen yet).  Along a line the two
// transforms decouple into maximal RUNS [a,e) of equal polarity; inside a run the candidates are the run's finite
// positions (q, |s(q)|) plus the two BOUNDARY SITES (a-1, 0) and (e, 0) -- the nearest voxels of the other polarity.
// They are ordinary hull points: the left one is the bottom of the run's chain, the right one closes the chain (pops
// what it hides) when the run ends.  The query sweep then needs two candidate values and one comparison per position.
//
// Scratch contents.  The hull of run [a,e) occupies the line's stack slots a, a+1, .. (at most e-a of them: one per vertex,
// in position order), followed by a terminator when there is room; runs never share slots.  The polarity of every position
// goes to a bit field pol[word][line] (1/32 of the field), which the query sweep reads to find run ends and signs.
//
// Exactness: all integer; hull tests are cross-multiplied in int64 (|dY| <= 2^30 + 2^25 with the "nothing on the left"
// site, |dq| < 2^11).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SDF_HD __host__ __device__ __forceinline__
#else
#define SDF_HD inline
#endif

#ifndef SDF_INF
#define SDF_INF 0x40000000
#endif

namespace sdfb200 {

#if defined(LH_COUNT) && !defined(__CUDA_ARCH__)
#define LH_TICK(x) (++lh_counters[x])
static long long lh_counters[8];   // 0 sites, 1 finite, 2 pops, 3 runs, 4 query advances
#else
#define LH_TICK(x) ((void)0)
#endif

// Stack entries, three formats: 32 bits {position : 10, height : 22} for lines of <= 1024 positions whose heights stay
// below 2^22, 32 bits {position : 11, height : 21} for lines of <= 2048 positions with heights below 2^21, else 64 bits
// {position, height}.  The all-ones word is the terminator: a 32-bit format is only chosen when every height stays below
// its all-ones height field, so no vertex can look like it.
enum { LH_FMT_P10 = 0, LH_FMT_P11 = 1, LH_FMT_WIDE = 2 };
template <int FMT> struct LineEntry {               // the two 32-bit formats
    static constexpr int HBITS = FMT == LH_FMT_P10 ? 22 : 21;
    typedef uint32_t type;
    static constexpr uint32_t TERM = 0xffffffffu;
    static SDF_HD type pack(int pos, int32_t h) { return ((uint32_t)pos << HBITS) | (uint32_t)h; }
    static SDF_HD bool is_term(type e) { return e == TERM; }
    static SDF_HD type term() { return TERM; }
    static SDF_HD int pos(type e) { return (int)(e >> HBITS); }
    static SDF_HD int32_t h(type e) { return (int32_t)(e & ((1u << HBITS) - 1u)); }
};
template <> struct LineEntry<LH_FMT_WIDE> {
    typedef unsigned long long type;
    static SDF_HD type pack(int pos, int32_t h) { return ((unsigned long long)(uint32_t)pos << 32) | (uint32_t)h; }
    static SDF_HD bool is_term(type e) { return (e >> 32) == 0xffffffffull; }
    static SDF_HD type term() { return 0xffffffff00000000ull; }
    static SDF_HD int pos(type e) { return (int)(e >> 32); }
    static SDF_HD int32_t h(type e) { return (int32_t)(e & 0xffffffffull); }
};
// the narrowest format that holds lines of n positions with heights <= hmax
SDF_HD int lh_entry_format(int64_t n, int64_t hmax)
{
    if (n <= 1024 && hmax < (1 << 22) - 1) return LH_FMT_P10;
    if (n <= 2048 && hmax < (1 << 21) - 1) return LH_FMT_P11;
    return LH_FMT_WIDE;
}

// Stack layout.  The 32 lines of a warp share one block of n x 32 entries; slot s of lane l lives at
//     (s / C) * 32 * C  +  l * C  +  s % C            C = LH_CHUNK = 4
// i.e. C consecutive slots of ONE line are contiguous (16 bytes with 32-bit entries: a 32-byte sector belongs to two
// neighbouring lines, a 128-byte cache line to eight), and a warp whose lanes sit at similar depths touches four or so
// cache lines per access.  A line's own pushes, pops and fetches therefore keep hitting the sector they just used (the
// hull fetches of the query sweep walk a lane's slots in order), instead of sharing every sector with seven other lines
// that want it at different times, as the plain [slot][line] layout does.  Measured at 1024^3 (y / x pass, ms):
// C = 1: 4.36 / 7.77, C = 2: 4.20 / 7.12, C = 4: 4.08 / 6.87, C = 8: 4.14 / 7.05, C = 16: 4.24 / 7.25; DRAM traffic of the
// x pass 2.07 -> 1.63 x the algorithmic bytes, L1 hit rate 68 -> 81 %.  `stk` below is the lane's slot 0: block base + l * C.
#ifndef LH_U
#define LH_U 4
#endif
#ifndef LH_CHUNK_LOG
#define LH_CHUNK_LOG 2
#endif
constexpr int LH_CHUNK = 1 << LH_CHUNK_LOG;          // consecutive slots of one line that are contiguous in memory
template <typename T> SDF_HD T* lh_slot(T* stk, int s) { return stk + ((s & ~(LH_CHUNK - 1)) << 5) + (s & (LH_CHUNK - 1)); }
SDF_HD int64_t lh_block_entries(int n) { return (int64_t)32 * ((n + LH_CHUNK - 1) & ~(LH_CHUNK - 1)); }   // stack block of one warp
SDF_HD int lh_lane_offset(int lane) { return lane * LH_CHUNK; }

// t on or above the chord s->u of the lifted points  <=>  parabola t never wins
SDF_HD bool lh_dominated(int sp, int32_t sY, int tp, int32_t tY, int up, int32_t uY) {
    const long long lhs = (long long)(tY - sY) * (long long)(up - tp);
    const long long rhs = (long long)(uY - tY) * (long long)(tp - sp);
    return lhs >= rhs;
}

#if defined(__CUDA_ARCH__)
#define LH_LDCS(p) __ldcs(p)        // streamed once: evict-first, so the stack rows keep their place in L2
#define LH_CTZ(x) (__ffs((int)(x)) - 1)
#else
#define LH_LDCS(p) (*(p))
#define LH_CTZ(x) __builtin_ctz(x)
#endif

// One line.  in: element 0 of the line, consecutive positions `stride` elements apart; stk: the lane's slot 0 (above);
// pol: word 0 of the line's polarity bits, consecutive words `pstride` elements apart.  emit(m, neg) receives the
// squared distance m (<= SDF_INF) and the polarity of positions 0, 1, .. n-1 in order.
//
// Chain of a run [a,e), bottom to top:  B, v0, v1, ..  with B the left boundary site (a-1, 0) -- for a = 0 the site
// (-1, 2^30), which can never win and so stands for "nothing on the left".  Only the vertices v_i are stored (slot a+i).
// The build keeps the top three elements in registers (T, S unpacked; R = v_{d-3} as a packed entry): a pop refills R
// with a load whose result is not needed before the next pop.  The query sweep likewise holds the entry after `next`
// in a register, fetched one advance ahead.
//
// (Variants that were measured and dropped, see DESIGN.md: a shared-memory ring for the top of the stack, write-back
// instead of write-through pushes, an asynchronous-copy FIFO for the query sweep.  Each removed stalls or L2 traffic
// but cost more issue slots than it saved; this kernel is issue-bound.)
template <int FMT, typename Emit>
SDF_HD void line_hull_pass(const int32_t* __restrict__ in, int64_t stride, int n, typename LineEntry<FMT>::type* __restrict__ stk,
                           uint32_t* __restrict__ pol, int64_t pstride, Emit emit)
{
    typedef LineEntry<FMT> E;
    typedef typename E::type ent_t;
    constexpr int32_t BIG = 0x40000000;              // height of the "nothing on the left" boundary site
    constexpr int32_t NOSITE = 0x60000000;           // height of a non-existent next point: never wins
    if (n <= 0) return;
    // ------------------------------------------------------------------------------------------------ build sweep
    {
        int a = 0, d = 0;                            // current run start; number of vertices on its stack
        // left boundary site of the run: position a - 1, lifted height (a-1)^2, or 1 + BIG when there is nothing on the left
        int tp = -1; int32_t tY = 1 + BIG;           // T: top of the chain (starts as the boundary site)
        int sp = tp; int32_t sY = tY + 1;            // S: below T, or the none-sentinel (sp == tp: the hull test is false)
        ent_t R = 0;                                 // v_{d-3} when it exists
        int ts = 0;                                  // slot a + d (next push)
        bool prevpol = in[0] < 0;
        auto pop_against = [&](int qp, int32_t qY) {
            while (lh_dominated(sp, sY, tp, tY, qp, qY)) {
                LH_TICK(2);
                --d;
                --ts;
                tp = sp; tY = sY;
                if (d >= 2) {
                    sp = E::pos(R); sY = sp * sp + E::h(R);                  // v_{d-2}
                    if (d >= 3) R = *lh_slot(stk, ts - 3);                        // v_{d-3}
                } else if (d == 1) { sp = a - 1; sY = a > 0 ? sp * sp : 1 + BIG; }
                else { sp = tp; sY = tY + 1; }                               // T is the boundary site: nothing below
            }
        };
        auto row = [&](int y, int32_t v, bool p) {
            LH_TICK(0);
            if (p != prevpol) {                      // a run ends at y-1 and the next starts at y
                LH_TICK(3);
                pop_against(y, y * y);               // close against the right boundary site (y, 0)
                if (d < y - a) *lh_slot(stk, ts) = E::term();
                a = y; d = 0;
                tp = y - 1; tY = tp * tp; sp = tp; sY = tY + 1;
                ts = y;
                prevpol = p;
            }
            const int32_t h = p ? -v : v;
            if (h < SDF_INF) {
                LH_TICK(1);
                const int32_t Y = y * y + h;
                pop_against(y, Y);
                *lh_slot(stk, ts) = E::pack(y, h);
                ++ts;
                ++d;
                R = E::pack(sp, sY - sp * sp);       // (garbage when S is not a vertex: then never used)
                sp = tp; sY = tY;
                tp = y; tY = Y;
            }
        };
        constexpr int U = LH_U;                      // rows read ahead of the (data-dependent) stack work
        const int32_t* ip = in;
        uint32_t pw = 0;
        int32_t vnx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { vnx[u] = (u < n) ? LH_LDCS(ip) : 0; ip += stride; }
        int y0 = 0;
        for (; y0 + 2 * U <= n; y0 += U) {           // whole groups with a whole group to read ahead: no bounds checks
            int32_t vin[U];
            uint32_t bits = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) { vin[u] = vnx[u]; bits |= ((uint32_t)vin[u] >> 31) << u; }
#pragma unroll
            for (int u = 0; u < U; ++u) { vnx[u] = LH_LDCS(ip); ip += stride; }
#pragma unroll
            for (int u = 0; u < U; ++u) row(y0 + u, vin[u], (bits >> u) & 1u);
            pw |= bits << (y0 & 31);
            if ((y0 & 31) == 32 - U) { pol[(int64_t)(y0 >> 5) * pstride] = pw; pw = 0; }
        }
        for (; y0 < n; y0 += U) {                    // the last one or two groups
            int32_t vin[U];
#pragma unroll
            for (int u = 0; u < U; ++u) vin[u] = vnx[u];
#pragma unroll
            for (int u = 0; u < U; ++u) { vnx[u] = (y0 + U + u < n) ? LH_LDCS(ip) : 0; ip += stride; }
            for (int u = 0; u < U && y0 + u < n; ++u) {
                const bool p = vin[u] < 0;
                row(y0 + u, vin[u], p);
                pw |= (uint32_t)p << ((y0 + u) & 31);
            }
            if ((y0 & 31) == 32 - U || y0 + U >= n) { pol[(int64_t)(y0 >> 5) * pstride] = pw; pw = 0; }
        }
        if (d < n - a) *lh_slot(stk, ts) = E::term();  // the last run has no right boundary site
    }
    // ------------------------------------------------------------------------------------------------ query sweep
    {
        int e = 0;                                   // end of the current run (the first run starts at 0)
        int cpos = 0, npos = 0;                      // current / next point of the run's hull (vertex or boundary site)
        int32_t ch = BIG, nh = NOSITE;
        ent_t NE = E::term();                        // the stack entry after `next`, fetched ahead (terminator: none)
        int fs = 0;                                  // slot after NE's
        int left = 0;                                // slots of the run not fetched yet
        bool tail = true, neg = false;
        // next <- the point after it: the entry fetched ahead, then the right boundary site (e, 0), then nothing
        auto advance = [&]() {
            cpos = npos; ch = nh;
            if (!E::is_term(NE)) {
                npos = E::pos(NE); nh = E::h(NE);
                NE = E::term();
                if (left > 0) { NE = *lh_slot(stk, fs); ++fs; --left; }
            } else if (!tail) { tail = true; npos = e; nh = (e < n) ? 0 : NOSITE; }
            else nh = NOSITE;
        };
        const uint32_t* pp = pol;
        const int nw = (n + 31) >> 5;
        for (int w = 0; w < nw; ++w) {
            const uint32_t pw = *pp; pp += pstride;
            const int yend = (w + 1) * 32 < n ? (w + 1) * 32 : n;
            for (int y = w * 32; y < yend; ++y) {
                if (y == e) {                        // a run starts here: find its end, restart the walk
                    neg = (pw >> (y & 31)) & 1u;
                    uint32_t x = (neg ? ~pw : pw) & ~((2u << (y & 31)) - 1u);  // later positions of the other polarity
                    int w2 = w;
                    const uint32_t* p2 = pp;
                    while (!x && ++w2 < nw) { const uint32_t q = *p2; p2 += pstride; x = neg ? ~q : q; }
                    e = x ? (w2 << 5) + LH_CTZ(x) : n;
                    if (e > n) e = n;                // (the bits past the end of the line are zero)
                    fs = y;
                    left = e - y;
                    NE = *lh_slot(stk, fs); ++fs; --left;  // (left >= 1 here)
                    tail = false;
                    npos = y - 1; nh = y > 0 ? 0 : BIG;       // the left boundary site becomes `current` ...
                    advance();                       // ... and the first vertex (or the right boundary site) `next`
                }
                for (;;) {
                    const int dc = y - cpos, dn = y - npos;
                    const int32_t vc = dc * dc + ch, vn = dn * dn + nh;
                    if (vn > vc) {
                        emit(vc < SDF_INF ? vc : (int32_t)SDF_INF, neg);
                        break;
                    }
                    LH_TICK(4);
                    advance();
                }
            }
        }
    }
}

}  // namespace sdfb200
