// sdf_tools_b200/csrc/run_hull.cuh
//
// One 1-D pass of the signed-fused exact Euclidean distance transform along a STRIDED axis -- the inner work of
// sdf_generation::BuildDistanceField (/root/reference/include/sdf_tools/sdf_generation.hpp:95-207, called twice at
// :242-243) for one axis -- organised for the GPU: the 32 lanes of a warp work on ADJACENT LINES in lock step
// (coalesced rows), different warps of the CTA own different BANDS (32 consecutive positions) of those lines.
//
// Signed-fused field: s > 0 <=> voxel FREE, s = d^2 to the nearest FILLED; s < 0 <=> voxel FILLED, -s = d^2 to the
// nearest FREE (partial over the axes done so far; |s| == SDF_INF: other polarity not seen yet).  Along a line the
// two transforms decouple into maximal RUNS [a,e) of equal polarity.  Inside a run the candidates are the run's own
// finite positions (q, |s(q)|) plus the two BOUNDARY SITES (a-1, 0) and (e, 0) -- the nearest voxels of the other
// polarity -- and
//     out(p) = min over candidates (p-q)^2 + h(q)  =  lower envelope of parabolas
//            = walk over the lower convex hull of the lifted points (q, q^2 + h(q)).
// Boundary sites are ordinary hull points (always extreme, never dropped), so the per-position work of the final
// walk is two candidate values and one comparison.
//
// Representation.  The staged tile S[position][line] is READ-ONLY (raw signed values); the hull of every run is a
// bit mask of ALIVE positions (AW, one word per band and line); run starts are a second bit mask (PC); two summary
// words per line say which bands hold alive vertices (NE) / run starts (PCS), so "previous / next vertex" and "run
// bounds" are a couple of bit scans -- no packed stack entries, no slot moves, no per-band metadata walks.
//
// Per tile = LPC adjacent lines x n positions:
//   phase build   thread (line c, band b): sequential hull (Felzenszwalb stack) of the band's own positions in
//                 band-relative lifted coordinates (32-bit tests); the stack is the alive mask plus (top, top-second)
//                 in registers.  A run that starts inside the band opens with its left boundary site as the bottom;
//                 a run that ends inside the band (or exactly at its end) is closed against its right boundary site.
//   phase merge   log2(NB) rounds; round lv joins the hulls of band groups of 2^lv pairwise by walking the common
//                 tangent across the seam (only when a run crosses it), boundary sites included when they lie in the
//                 group.  Tasks (line, seam) are independent of the build mapping and are packed onto the first
//                 threads of the CTA.
//   phase query   thread (c, b) walks the finished hull for its band's positions (monotone minimiser) and emits
//                 the results row by row (coalesced across lines).
// Phases communicate through shared memory only and are race-free inside a phase, so the same functions run on the
// GPU (separated by __syncthreads) and, for tests without a GPU, on the CPU as loops over the threads
// (tests/emul/emul_run.cpp).
//
// Exactness: all integer.  |s| <= 3*2048^2 < 2^24; hull tests inside a band fit 32 bits (|dY| < 2^25, |dq| <= 33),
// tests across bands are cross-multiplied in int64.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SDF_HD __host__ __device__ __forceinline__
#define SDF_HD_COLD __host__ __device__ __noinline__
#else
#define SDF_HD inline
#define SDF_HD_COLD inline
#endif

#ifndef SDF_INF
#define SDF_INF 0x40000000
#endif

namespace sdfb200 {

// ------------------------------------------------------------------------------------------------ bit helpers
SDF_HD int rh_hibit(uint32_t m) {             // index of the highest set bit (m != 0)
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)m);
#else
    return 31 - __builtin_clz(m);
#endif
}
SDF_HD int rh_lobit(uint32_t m) {             // index of the lowest set bit (m != 0)
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
SDF_HD uint32_t rh_lowmask(int k) {           // bits [0, k), 0 <= k <= 32
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0), "r"(k));
    return r;
#else
    return k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
#endif
}
SDF_HD uint32_t rh_range(int lo, int hi) {    // bits [lo, hi), 0 <= lo, hi <= 32 (empty when hi <= lo)
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(lo), "r"(hi > lo ? hi - lo : 0));
    return r;
#else
    return hi > lo ? (rh_lowmask(hi) & ~rh_lowmask(lo)) : 0u;
#endif
}
SDF_HD uint32_t rh_shift_in_sign(uint32_t acc, int32_t v) {   // (acc << 1) | sign(v): one SHF on the GPU
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)v, acc, 1);
#else
    return (acc << 1) | ((uint32_t)v >> 31);
#endif
}
SDF_HD uint32_t rh_brev(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
SDF_HD int32_t rh_abs(int32_t v) { return v < 0 ? -v : v; }

// t on or above the chord s->u of the lifted points  <=>  parabola t never wins (int64: positions span the line)
SDF_HD bool rh_dominated(int sp, int32_t sY, int tp, int32_t tY, int up, int32_t uY) {
    const long long lhs = (long long)(tY - sY) * (long long)(up - tp);
    const long long rhs = (long long)(uY - tY) * (long long)(tp - sp);
    return lhs >= rhs;
}

#if defined(RH_COUNT) && !defined(__CUDA_ARCH__)
#define RH_TICK(x) (++rh_counters[x])
static long long rh_counters[16];   // 0 build sites, 1 finite, 2 pops, 3 run events, 4 merge tasks, 5 merge drops,
                                    // 6 query advances, 7 query left-descent steps, 8 merge tasks that walk
#else
#define RH_TICK(x) ((void)0)
#endif

// Shared-memory handles: on the GPU 32-bit shared-window byte addresses used with ld.shared / st.shared (a generic
// pointer would make the compiler re-derive the window base at every access); on the CPU ordinary addresses.
#if defined(__CUDA_ARCH__)
typedef uint32_t rh_addr_t;
SDF_HD uint32_t rh_ld32(rh_addr_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
SDF_HD uint32_t rh_ld32_ro(rh_addr_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }   // tile data: never written after staging
SDF_HD void rh_st32(rh_addr_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
SDF_HD void rh_or32(rh_addr_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
SDF_HD void rh_and32(rh_addr_t a, uint32_t v) { asm volatile("red.shared.and.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
#else
typedef uintptr_t rh_addr_t;
SDF_HD uint32_t rh_ld32(rh_addr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
SDF_HD uint32_t rh_ld32_ro(rh_addr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
SDF_HD void rh_st32(rh_addr_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(a) = v; }
SDF_HD void rh_or32(rh_addr_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(a) |= v; }      // emulation is sequential
SDF_HD void rh_and32(rh_addr_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(a) &= v; }
#endif

struct RunTile {
    rh_addr_t S;      // int32  [NB*32][LPC]  staged signed values, read-only
    rh_addr_t AW;     // uint32 [NB][LPC]     alive (hull vertex) bits of band b
    rh_addr_t PC;     // uint32 [NB][LPC]     run-start bits of band b (polarity differs from the previous position)
    rh_addr_t NE;     // uint32 [NW][LPC]     bit (b & 31) of word (b >> 5): AW[b] != 0          (zeroed before build)
    rh_addr_t PCS;    // uint32 [NW][LPC]     bit (b & 31) of word (b >> 5): PC[b] != 0          (zeroed before build)
};

template <int LPC, int NW>     // lines per tile; summary words per line (bands <= 32 * NW)
struct RunHull {
    static constexpr int B = 32;
    static constexpr uint32_t STEP = LPC * 4;                 // bytes between consecutive positions of one line
    static constexpr int32_t NOSITE = 0x60000000;             // height of a non-existent vertex: never wins
    static constexpr int NONE = -0x40000000;                  // position of a non-existent vertex

    struct Summ { uint32_t w[NW]; };

    static SDF_HD rh_addr_t s_addr(const RunTile& T, int pos, int c) { return T.S + 4u * (uint32_t)(pos * LPC + c); }
    static SDF_HD int32_t ldS(const RunTile& T, int pos, int c) { return (int32_t)rh_ld32_ro(s_addr(T, pos, c)); }
    static SDF_HD uint32_t ldAW(const RunTile& T, int b, int c) { return rh_ld32(T.AW + 4u * (uint32_t)(b * LPC + c)); }
    static SDF_HD void stAW(const RunTile& T, int b, int c, uint32_t v) { rh_st32(T.AW + 4u * (uint32_t)(b * LPC + c), v); }
    static SDF_HD uint32_t ldPC(const RunTile& T, int b, int c) { return rh_ld32(T.PC + 4u * (uint32_t)(b * LPC + c)); }
    static SDF_HD void stPC(const RunTile& T, int b, int c, uint32_t v) { rh_st32(T.PC + 4u * (uint32_t)(b * LPC + c), v); }
    static SDF_HD Summ ld_summ(rh_addr_t base, int c) {
        Summ s;
#pragma unroll
        for (int w = 0; w < NW; ++w) s.w[w] = rh_ld32(base + 4u * (uint32_t)(w * LPC + c));
        return s;
    }
    // lowest / highest set band index in [lo, hi); -1 if none
    static SDF_HD int next_band(const Summ& s, int lo, int hi) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t x = s.w[w] & rh_range(lo - 32 * w < 0 ? 0 : lo - 32 * w, hi - 32 * w > 32 ? 32 : hi - 32 * w);
            if (x) return 32 * w + rh_lobit(x);
        }
        return -1;
    }
    static SDF_HD int prev_band(const Summ& s, int lo, int hi) {
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) {
            const uint32_t x = s.w[w] & rh_range(lo - 32 * w < 0 ? 0 : lo - 32 * w, hi - 32 * w > 32 ? 32 : hi - 32 * w);
            if (x) return 32 * w + rh_hibit(x);
        }
        return -1;
    }
    // start of the run that contains position q (largest run start <= q; 0 if none)
    static SDF_HD int run_start(const RunTile& T, int c, const Summ& pcs, int q) {
        const int b = q >> 5;
        const uint32_t x = ldPC(T, b, c) & rh_lowmask((q & 31) + 1);
        if (x) return 32 * b + rh_hibit(x);
        const int pb = prev_band(pcs, 0, b);
        return pb < 0 ? 0 : 32 * pb + rh_hibit(ldPC(T, pb, c));
    }
    // first run start after the band `b` (bands > b); n if none
    static SDF_HD int run_start_after_band(const RunTile& T, int c, const Summ& pcs, int b, int n) {
        const int nbnd = next_band(pcs, b + 1, 32 * NW);
        return nbnd < 0 ? n : 32 * nbnd + rh_lobit(ldPC(T, nbnd, c));
    }

    // ---------------------------------------------------------------------------------------- vertex iterators
    // Alive vertices at positions [lo, start), descending.
    struct LIt {
        int band, lo; uint32_t w;
        SDF_HD void init(const RunTile& T, int c, int lo_, int start) {
            lo = lo_;
            if (start <= lo) { band = 0; w = 0; lo = 0x7fffffff; return; }           // empty range
            band = (start - 1) >> 5;
            w = ldAW(T, band, c) & rh_lowmask(((start - 1) & 31) + 1);
        }
        SDF_HD int next(const RunTile& T, int c, const Summ& ne) {
            if (!w) {
                if (lo == 0x7fffffff) return -1;
                const int pb = prev_band(ne, lo >> 5, band);
                if (pb < 0) { lo = 0x7fffffff; return -1; }
                band = pb; w = ldAW(T, pb, c);
                if (!w) { lo = 0x7fffffff; return -1; }                               // (summary is exact: not reached)
            }
            const int k = rh_hibit(w);
            w ^= 1u << k;
            const int pos = 32 * band + k;
            if (pos < lo) { w = 0; lo = 0x7fffffff; return -1; }
            return pos;
        }
    };
    // Alive vertices at positions [start, hi), ascending.
    struct RIt {
        int band, hi; uint32_t w;
        SDF_HD void init(const RunTile& T, int c, int start, int hi_) {
            hi = hi_;
            if (start >= hi) { band = 0; w = 0; hi = -1; return; }
            band = start >> 5;
            w = ldAW(T, band, c) & ~rh_lowmask(start & 31);
        }
        SDF_HD void init_word(int b, uint32_t word, int hi_) { band = b; w = word; hi = hi_; }
        SDF_HD int next(const RunTile& T, int c, const Summ& ne) {
            if (!w) {
                if (hi < 0) return -1;
                const int nbnd = next_band(ne, band + 1, (hi + 31) >> 5);
                if (nbnd < 0) { hi = -1; return -1; }
                band = nbnd; w = ldAW(T, nbnd, c);
                if (!w) { hi = -1; return -1; }
            }
            const int k = rh_lobit(w);
            w &= w - 1u;
            const int pos = 32 * band + k;
            if (pos >= hi) { w = 0; hi = -1; return -1; }
            return pos;
        }
    };

    // ------------------------------------------------------------------------------------------- phase build
    // Requires NE / PCS zeroed and a barrier before.  Returns the band's polarity bits (pw) and run-start bits (pcw).
    static SDF_HD void build(const RunTile& T, int c, int b, int n, uint32_t& pw_out, uint32_t& pcw_out) {
        const int base = b << 5;
        int nvalid = n - base;
        nvalid = nvalid < 0 ? 0 : (nvalid > B ? B : nvalid);
        uint32_t aw = 0, pcw = 0, pw = 0;
        if (nvalid > 0) {
            const rh_addr_t rd0 = s_addr(T, base, c);
            // polarity bits of the band (rows past the end of the line are staged zeros: masked off) and its run starts
            {
                uint32_t pwr = 0;
#pragma unroll
                for (int k = 0; k < B; ++k) pwr = rh_shift_in_sign(pwr, (int32_t)rh_ld32_ro(rd0 + (uint32_t)k * STEP));
                pw = rh_brev(pwr) & rh_lowmask(nvalid);
                const uint32_t prevpol = b > 0 ? (rh_ld32_ro(rd0 - STEP) >> 31) : (pw & 1u);
                pcw = (pw ^ ((pw << 1) | prevpol)) & rh_lowmask(nvalid);
            }
            // Chain state in band-relative lifted coordinates (i, i^2 + h): top t = (tp, tY); d = t - (vertex below t),
            // or the sentinel d = (0, -1) that keeps the test false while the chain has a single element.
            // bp / bY: the left boundary site of the current run when the run starts inside the band (bp = run start - 1).
            int tp = -1, dP = 0, bp = NONE;
            int32_t tY = 0, dY = -1, bY = 0;
            uint32_t botmask = 0xffffffffu;                                          // positions >= start of the current run
            // The chain is popped against the lifted point (pi, Y); vertices that it hides are removed from `aw`.
            auto pop_against = [&](int pi, int32_t Y, int32_t& A, int& Bq) {
                while (dY * Bq >= A * dP) {                                          // t on or above the chord s -> new point
                    RH_TICK(2);
                    aw ^= 1u << tp;                                                  // (the top is a vertex here)
                    tp -= dP; tY -= dY;                                              // t <- s
                    const uint32_t below = tp > 0 ? (aw & botmask & rh_lowmask(tp)) : 0u;
                    if (below) {
                        const int sp = rh_hibit(below);
                        const int32_t sY = sp * sp + rh_abs((int32_t)rh_ld32_ro(rd0 + (uint32_t)sp * STEP));
                        dP = tp - sp; dY = tY - sY;
                    } else if (bp != NONE && tp != bp) { dP = tp - bp; dY = tY - bY; }   // the boundary site is below t
                    else { dP = 0; dY = -1; }                                        // t is the bottom of the chain
                    A = Y - tY; Bq = pi - tp;
                }
            };
            int i = 0;
            uint32_t rest = pcw;
            for (;;) {
                if (rest & (1u << i)) {                                              // a run starts at i
                    RH_TICK(3);
                    rest ^= 1u << i;
                    { int32_t A = i * i - tY; int Bq = i - tp; pop_against(i, i * i, A, Bq); }   // close the previous run against (i, 0)
                    tp = i - 1; tY = (i - 1) * (i - 1); dP = 0; dY = -1;             // open with the boundary site (i-1, 0)
                    bp = tp; bY = tY; botmask = ~rh_lowmask(i);
                }
                const int seg_end = rest ? rh_lobit(rest) : nvalid;                  // positions [i, seg_end) belong to one run
                // leading positions without a value: nothing to do
                int32_t h = SDF_INF;
                for (; i < seg_end; ++i) {
                    RH_TICK(0);
                    h = rh_abs((int32_t)rh_ld32_ro(rd0 + (uint32_t)i * STEP));
                    if (h < SDF_INF) break;
                }
                if (i < seg_end) {
                    RH_TICK(1);
                    const int32_t Y0 = i * i + h;                                    // first vertex of the run inside the band
                    if (bp != NONE) { dP = i - bp; dY = Y0 - bY; }
                    tp = i; tY = Y0; aw |= 1u << i;
                    ++i;
#pragma unroll 2
                    for (; i < seg_end; ++i) {
                        RH_TICK(0);
                        const int32_t hh = rh_abs((int32_t)rh_ld32_ro(rd0 + (uint32_t)i * STEP));
                        const int32_t Y = i * i + hh;
                        int32_t A = Y - tY;
                        int Bq = i - tp;
                        if (hh < SDF_INF) {
                            RH_TICK(1);
                            pop_against(i, Y, A, Bq);
                            aw |= 1u << i; dP = Bq; dY = A; tp = i; tY = Y;
                        }
                    }
                }
                if (seg_end >= nvalid) break;
            }
            // a run that ends exactly at the end of the band is closed against its right boundary site (32, 0)
            if (nvalid == B && base + B < n) {
                const uint32_t nextpol = rh_ld32_ro(rd0 + (uint32_t)B * STEP) >> 31;
                if (nextpol != (pw >> 31)) { int32_t A = B * B - tY; int Bq = B - tp; pop_against(B, B * B, A, Bq); }
            }
        }
        stAW(T, b, c, aw);
        stPC(T, b, c, pcw);
        if (aw) rh_or32(T.NE + 4u * (uint32_t)((b >> 5) * LPC + c), 1u << (b & 31));
        if (pcw) rh_or32(T.PCS + 4u * (uint32_t)((b >> 5) * LPC + c), 1u << (b & 31));
        pw_out = pw; pcw_out = pcw;
    }

    // ------------------------------------------------------------------------------------------- phase merge
    // Clears the alive bits at positions [p0, p1) (all of them belong to dropped vertices).
    static SDF_HD_COLD void clear_range(const RunTile& T, int c, int p0, int p1) {
        if (p1 <= p0) return;
        const int b0 = p0 >> 5, b1 = (p1 - 1) >> 5;
        for (int bb = b0; bb <= b1; ++bb) {
            const int lo = bb == b0 ? (p0 & 31) : 0, hi = bb == b1 ? ((p1 - 1) & 31) + 1 : 32;
            uint32_t w = 0;
            if (lo > 0 || hi < 32) { w = ldAW(T, bb, c) & ~rh_range(lo, hi); }
            stAW(T, bb, c, w);
            if (!w) rh_and32(T.NE + 4u * (uint32_t)((bb >> 5) * LPC + c), ~(1u << (bb & 31)));
        }
    }

    struct Vtx { int p; int32_t Y; };     // p == NONE: no vertex

    // Seam merge when no run starts inside the two groups (nor right after them): both chains are plain lists of alive
    // vertices, walked through the band summary `ne` (bits of the left / right group) and the alive words.
    static SDF_HD void merge_plain(const RunTile& T, int c, int sp, uint32_t neL, uint32_t neR) {
        if (!neL || !neR) return;
        RH_TICK(8);
        int bL = rh_hibit(neL), bR = rh_lobit(neR);
        neL ^= 1u << bL; neR ^= 1u << bR;
        uint32_t wL = ldAW(T, bL, c), wR = ldAW(T, bR, c);
        auto fetchL = [&]() -> Vtx {                                       // next vertex to the left, NONE when exhausted
            if (!wL) {
                if (!neL) return Vtx{NONE, 0};
                bL = rh_hibit(neL); neL ^= 1u << bL; wL = ldAW(T, bL, c);
            }
            const int k = rh_hibit(wL);
            wL ^= 1u << k;
            const int p = 32 * bL + k;
            return Vtx{p, p * p + rh_abs(ldS(T, p, c))};
        };
        auto fetchR = [&]() -> Vtx {
            if (!wR) {
                if (!neR) return Vtx{NONE, 0};
                bR = rh_lobit(neR); neR ^= 1u << bR; wR = ldAW(T, bR, c);
            }
            const uint32_t lb = wR & (0u - wR);
            wR ^= lb;
            const int p = 32 * bR + rh_hibit(lb);
            return Vtx{p, p * p + rh_abs(ldS(T, p, c))};
        };
        Vtx L = fetchL(), R = fetchR();
        Vtx L2 = fetchL(), R2 = fetchR();
        bool dropL = false, dropR = false;
        for (;;) {
            bool changed = false;
            while (L2.p != NONE && rh_dominated(L2.p, L2.Y, L.p, L.Y, R.p, R.Y)) {
                RH_TICK(5);
                L = L2; L2 = fetchL(); dropL = true; changed = true;
            }
            while (R2.p != NONE && rh_dominated(L.p, L.Y, R.p, R.Y, R2.p, R2.Y)) {
                RH_TICK(5);
                R = R2; R2 = fetchR(); dropR = true; changed = true;
            }
            if (!changed) break;
        }
        if (dropL) clear_range(T, c, L.p + 1, sp);
        if (dropR) clear_range(T, c, sp, R.p);
    }

    // Round lv, seam at band b (b = (2j+1) << lv): joins the hulls of band groups [b-2^lv, b) and [b, b+2^lv).
    static SDF_HD void merge(const RunTile& T, int c, int b, int nb, int lv, int n) {
        const int g = 1 << lv;
        if (b >= nb) return;
        const uint32_t pc_b = ldPC(T, b, c);
        if (pc_b & 1u) return;                                   // a run starts exactly at the seam
        RH_TICK(4);
        const int glo = b - g, ghi = (b + g > nb) ? nb : b + g;
        const int sp = b << 5;
        const Summ ne = ld_summ(T.NE, c), pcs = ld_summ(T.PCS, c);
        if (NW == 1 && !(pcs.w[0] & rh_range(glo, ghi + 1))) {     // no run start in the groups or right after them
            merge_plain(T, c, sp, ne.w[0] & rh_range(glo, b), ne.w[0] & rh_range(b, ghi));
            return;
        }
        // the run that crosses the seam: [a, e)
        const int a = run_start(T, c, pcs, sp - 1);
        int e;
        { const uint32_t x = pc_b & ~1u; e = x ? sp + rh_lobit(x) : run_start_after_band(T, c, pcs, b, n); }
        const int lo = a > (glo << 5) ? a : (glo << 5);
        int hi = e; if (hi > (ghi << 5)) hi = ghi << 5; if (hi > n) hi = n;
        bool pl = a > 0 && a >= (glo << 5);                      // left boundary site (a-1, 0) belongs to the left chain
        bool pr = e < n && e <= (ghi << 5);                      // right boundary site (e, 0) belongs to the right chain
        LIt li; li.init(T, c, lo, sp);
        RIt ri; ri.init(T, c, sp, hi);
        auto fetchL = [&]() -> Vtx {
            const int p = li.next(T, c, ne);
            if (p >= 0) return Vtx{p, p * p + rh_abs(ldS(T, p, c))};
            if (pl) { pl = false; return Vtx{a - 1, (a - 1) * (a - 1)}; }
            return Vtx{NONE, 0};
        };
        auto fetchR = [&]() -> Vtx {
            const int p = ri.next(T, c, ne);
            if (p >= 0) return Vtx{p, p * p + rh_abs(ldS(T, p, c))};
            if (pr) { pr = false; return Vtx{e, e * e}; }
            return Vtx{NONE, 0};
        };
        Vtx L = fetchL();
        if (L.p == NONE) return;
        Vtx R = fetchR();
        if (R.p == NONE) return;
        RH_TICK(8);
        Vtx L2 = fetchL(), R2 = fetchR();
        bool dropL = false, dropR = false;
        for (;;) {
            bool changed = false;
            while (L2.p != NONE && rh_dominated(L2.p, L2.Y, L.p, L.Y, R.p, R.Y)) {
                RH_TICK(5);
                L = L2; L2 = fetchL(); dropL = true; changed = true;
            }
            while (R2.p != NONE && rh_dominated(L.p, L.Y, R.p, R.Y, R2.p, R2.Y)) {
                RH_TICK(5);
                R = R2; R2 = fetchR(); dropR = true; changed = true;
            }
            if (!changed) break;
        }
        if (dropL) clear_range(T, c, L.p + 1, sp);               // L a boundary site: L.p = a-1 -> clears [a, sp)
        if (dropR) clear_range(T, c, sp, R.p);                   // R a boundary site: R.p = e   -> clears [sp, e)
    }

    // ------------------------------------------------------------------------------------------- phase query
    // Emits out(p) for the band's positions through `emit(i, value)`, i = position inside the band, increasing.
    // One loop, one copy of the "next vertex" code: a position is re-evaluated after every advance, and a run start
    // makes current == next == the run's left boundary site, so the ordinary advance fetches the run's first vertex.
    template <typename Emit>
    static SDF_HD void query(const RunTile& T, int c, int b, int nb, int n, uint32_t pw, uint32_t pcw, Emit emit) {
        const int base = b << 5;
        if (base >= n) return;
        int nvalid = n - base;
        nvalid = nvalid > B ? B : nvalid;
        const Summ ne = ld_summ(T.NE, c), pcs = ld_summ(T.PCS, c);
        const uint32_t myaw = ldAW(T, b, c);
        const int e_after = run_start_after_band(T, c, pcs, b, n);          // first run start in later bands (n if none)
        int e;                                                              // end of the current run
        { const uint32_t x = pcw & ~1u; e = x ? base + rh_lobit(x) : e_after; }
        // current / next hull vertex; "none" is encoded in the heights
        int cpos = base, npos = base;
        int32_t ch = SDF_INF, nh = NOSITE;
        // vertices after `next`: alive bits `itw` of band `itb` still ahead, then later bands; tail: nothing follows
        int itb = b; uint32_t itw = 0; bool tail = true;
        if (!(pcw & 1u)) {                                                  // position `base` continues a run from the left
            const int a = base > 0 ? run_start(T, c, pcs, base - 1) : 0;
            // L: last vertex before base, R: first vertex at / after base (boundary sites included)
            LIt li; li.init(T, c, a, base);
            int Lp = li.next(T, c, ne); int32_t Lh = 0;
            if (Lp >= 0) Lh = rh_abs(ldS(T, Lp, c)); else if (a > 0) { Lp = a - 1; Lh = 0; } else Lp = NONE;
            int Rp; int32_t Rh = 0;
            { int pp = -1;
              if (myaw) pp = base + rh_lobit(myaw);
              else { const int nbnd = next_band(ne, b + 1, 32 * NW); if (nbnd >= 0) pp = 32 * nbnd + rh_lobit(ldAW(T, nbnd, c)); }
              if (pp >= 0 && pp < e) { Rp = pp; Rh = rh_abs(ldS(T, pp, c)); } else if (e < n) { Rp = e; Rh = 0; } else Rp = NONE; }
            bool takeR = Lp == NONE;
            if (!takeR && Rp != NONE) takeR = (base - Rp) * (base - Rp) + Rh <= (base - Lp) * (base - Lp) + Lh;
            if (takeR) {
                if (Rp != NONE) { cpos = Rp; ch = Rh; }
            } else {
                cpos = Lp; ch = Lh;
                while (cpos >= a) {                                         // local search towards the minimiser of `base`
                    int Pp = li.next(T, c, ne); int32_t Ph = 0;
                    if (Pp >= 0) Ph = rh_abs(ldS(T, Pp, c)); else if (a > 0) { Pp = a - 1; } else break;
                    const int dc = base - cpos, dp = base - Pp;
                    if (dp * dp + Ph < dc * dc + ch) { RH_TICK(7); cpos = Pp; ch = Ph; } else break;
                }
            }
            if (ch < SDF_INF && cpos < e) {                                 // a vertex that can have a successor
                npos = cpos; nh = ch;                                       // next == current: the loop fetches the successor
                tail = false;
                itb = (cpos + 1) >> 5;                                      // (cpos >= -1)
                itw = itb < nb ? (ldAW(T, itb, c) & ~rh_lowmask((cpos + 1) & 31)) : 0u;
            }
        }
        uint32_t rest = pcw;
        uint32_t sgn = 0u - (pw & 1u);                                      // polarity of the current run as a mask
        uint32_t bit = 1u;
        const uint32_t endbit = nvalid < B ? (1u << nvalid) : 0u;           // `bit` after the last position (0 after 32 shifts)
        int p = base;
        do {
            const int dc = p - cpos, dn = p - npos;
            const int32_t vc = dc * dc + ch, vn = dn * dn + nh;
            if ((vn <= vc) | ((rest & bit) != 0u)) {                        // rare: a run starts at p, or the next vertex takes over
                if (rest & bit) {                                           // run start: left boundary site (p-1, 0)
                    rest ^= bit;
                    e = rest ? base + rh_lobit(rest) : e_after;
                    cpos = npos = p - 1; ch = nh = 0;
                    tail = false; itb = b; itw = myaw & (0u - bit);
                    sgn = (pw & bit) ? 0xffffffffu : 0u;
                    continue;
                }
                RH_TICK(6);
                cpos = npos; ch = nh;                                       // advance, then fetch the successor
                nh = NOSITE;
                if (!tail) {
                    if (!itw) {
                        const int nbnd = next_band(ne, itb + 1, 32 * NW);
                        if (nbnd >= 0) { itb = nbnd; itw = ldAW(T, nbnd, c); }
                    }
                    int q = e;                                              // the run's right boundary site unless a vertex comes first
                    if (itw) { const uint32_t lb = itw & (0u - itw); q = 32 * itb + rh_hibit(lb); itw ^= lb; }
                    if (q < e) { npos = q; nh = rh_abs(ldS(T, q, c)); }
                    else { tail = true; if (e < n) { npos = e; nh = 0; } }
                }
                continue;
            }
            const int32_t best = vc < SDF_INF ? vc : SDF_INF;
            emit(p - base, (int32_t)(((uint32_t)best ^ sgn) - sgn));
            ++p;
            bit <<= 1;
        } while (bit != endbit);
    }
};

}  // namespace sdfb200
