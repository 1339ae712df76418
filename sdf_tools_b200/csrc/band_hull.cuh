// sdf_tools_b200/csrc/band_hull.cuh
//
// One 1-D pass of the signed-fused exact Euclidean distance transform along a STRIDED axis, organised so that
// the 32 lanes of a warp always work on 32 ADJACENT LINES in lock step (coalesced rows, no intra-line divergence),
// while different warps of the CTA own different BANDS (B consecutive positions) of those lines.
// It replaces the inner work of sdf_generation::BuildDistanceField
// (/root/reference/include/sdf_tools/sdf_generation.hpp:95-207, called twice at :242-243) for one axis.
//
// Signed-fused field: s > 0 <=> voxel FREE, s = d^2 to the nearest FILLED; s < 0 <=> voxel FILLED, -s = d^2 to the
// nearest FREE (partial over the axes done so far; |s| == SDF_INF: other polarity not seen yet).  Along a line the
// two transforms decouple into maximal RUNS of equal polarity:
//     p in run [a,e):  out(p) = min( (p-(a-1))^2 , (e-p)^2 , min_{q in [a,e)} (p-q)^2 + |s(q)| )
// The last term is the lower envelope of parabolas == lower convex hull of the lifted points (q, q^2 + |s(q)|).
//
// Per tile = LPC adjacent lines x n positions, thread (line c, band b):
//   phase build   sequential hull (Felzenszwalb stack) of the band's own positions, in band-relative lifted
//                 coordinates (i, i^2 + |s|) so that every test is a pair of 32-bit multiplies; the stack lives in
//                 shared memory at S[(b*B + k)*LPC + c], built IN PLACE over the band's staged values, as packed words
//                 {position-in-band : 6, first-of-run flag : 1, |s| : 24}.  Restarted at every polarity change.  Bands
//                 that hold a run start set their bit in the line's band mask (anyw).
//   phase merge   log2(NB) rounds; round lv joins band groups of 2^lv pairwise by walking the common tangent
//                 across the seam (only when a run continues across it); trims band stacks from their ends.  A walk
//                 that keeps dropping switches to jumps: whole bands, then a binary search inside the last band.
//   phase query   each thread walks the finished hull for its own band's positions (monotone minimiser),
//                 adds the two run-end terms (nearest run starts through the band mask) and writes the result row
//                 (coalesced across lines).
// Phases only communicate through the shared arrays and are race-free inside a phase, so the same functions run on
// the GPU (separated by __syncthreads; seam round 0 by __syncwarp when a warp holds both bands) and, for tests without
// a GPU, on the CPU as plain loops over (c, b).
//
// Exactness: all integer.  |s| <= 3*2048^2 < 2^24, lifted heights < 2^25; hull tests inside a band fit 32 bits
// (|dY| < 2^25, |dq| < 2^LOGB <= 64), tests across bands are cross-multiplied in int64.
#pragma once
#include <stdint.h>
#include <type_traits>

#if defined(__CUDACC__)
#define SDF_HD __host__ __device__ __forceinline__
#else
#define SDF_HD inline
#endif

#ifndef SDF_INF
#define SDF_INF 0x40000000
#endif

namespace sdfb200 {

SDF_HD int bh_hibit(uint32_t m) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)m);
#else
    return 31 - __builtin_clz(m);
#endif
}
SDF_HD int bh_lobit(uint32_t m) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
SDF_HD int bh_lobit64(uint64_t m) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}

SDF_HD int bh_hibit64(uint64_t m) {
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)m);
#else
    return 63 - __builtin_clzll(m);
#endif
}
// (acc << 1) | (sign bit of v): one SHF.L.W on the GPU
SDF_HD uint32_t bh_shift_in_sign(uint32_t acc, int32_t v) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)v, acc, 1);
#else
    return (acc << 1) | ((uint32_t)v >> 31);
#endif
}
SDF_HD uint32_t bh_brev(uint32_t x) {
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

// t on or above the chord s->u of the lifted points  <=>  parabola t never wins
SDF_HD bool bh_dominated(int sp, int32_t sY, int tp, int32_t tY, int up, int32_t uY) {
    const long long lhs = (long long)(tY - sY) * (long long)(up - tp);
    const long long rhs = (long long)(uY - tY) * (long long)(tp - sp);
    return lhs >= rhs;
}

#if defined(BH_COUNT) && !defined(__CUDA_ARCH__)
#define BH_TICK(x) (++bh_counters[x])
#define BH_STEP() (++bh_serial)
static long long bh_counters[8];
static long long bh_serial;               // dependent shared-memory round trips (latency proxy)
#else
#define BH_TICK(x) ((void)0)
#define BH_STEP() ((void)0)
#endif
constexpr int BH_BIG = 0x7fff;          // "no run boundary after": larger than any position

// Shared-memory arrays of one tile.  On the GPU the handles are 32-bit shared-window byte addresses accessed
// with ld.shared / st.shared (a generic pointer would make the compiler re-derive the window base from
// SR_CgaCtaId at every access); on the CPU (tests) they are ordinary addresses.
#if defined(__CUDA_ARCH__)
typedef uint32_t bh_addr_t;
SDF_HD uint32_t bh_ld32(bh_addr_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
SDF_HD void bh_st32(bh_addr_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
SDF_HD int bh_ld16(bh_addr_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
SDF_HD void bh_st16(bh_addr_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((short)v) : "memory"); }
SDF_HD void bh_or32(bh_addr_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
#else
typedef uintptr_t bh_addr_t;
SDF_HD uint32_t bh_ld32(bh_addr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
SDF_HD void bh_st32(bh_addr_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(a) = v; }
SDF_HD int bh_ld16(bh_addr_t a) { return *reinterpret_cast<const int16_t*>(a); }
SDF_HD void bh_st16(bh_addr_t a, int v) { *reinterpret_cast<int16_t*>(a) = (int16_t)v; }
SDF_HD void bh_or32(bh_addr_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(a) |= v; }     // emulation is sequential
#endif

struct BandTile {
    bh_addr_t S;      // uint32 [NB*B][LPC]  packed hull entries
    bh_addr_t meta;   // uint32 [NB][LPC]    lo:7 | hi:7 <<8 | tailPC <<16 | anyPC <<17 | headPC <<18
    bh_addr_t pcA;    // int16  [NB][LPC]    last run start inside band b   (-1 if none)
    bh_addr_t pcE;    // int16  [NB][LPC]    first run start inside band b  (BH_BIG if none)
    bh_addr_t anyw;   // uint32 [NW][LPC]    bit (b & 31) of word (b >> 5): band b holds a run start; zeroed before build
};

template <int LOGB, int LPC>
struct BandHull {
    static constexpr int B = 1 << LOGB;
    using pcw_t = typename std::conditional<(LOGB > 5), uint64_t, uint32_t>::type;

    // ---- packed hull entry ------------------------------------------------------------------------------
    // {position in band : 6 (bits 26..31), first vertex of its run : 1 (bit 25), |s| : 24 (bits 0..23)}
    static constexpr int REL_SHIFT = 26;
    static constexpr uint32_t RS_BIT = 1u << 25;
    static SDF_HD uint32_t pack(int rel, int32_t h, bool rs) { return (uint32_t)h | ((uint32_t)rel << REL_SHIFT) | (rs ? RS_BIT : 0u); }
    static SDF_HD int32_t e_h(uint32_t e) { return (int32_t)(e & 0xFFFFFFu); }
    static SDF_HD int e_rel(uint32_t e) { return (int)(e >> REL_SHIFT); }
    static SDF_HD bool e_rs(uint32_t e) { return (e & RS_BIT) != 0; }
    // ---- band meta ---------------------------------------------------------------------------------------
    static SDF_HD uint32_t mk_meta(int lo, int hi, bool tailpc, bool anypc, bool headpc) {
        return (uint32_t)lo | ((uint32_t)hi << 8) | ((uint32_t)tailpc << 16) | ((uint32_t)anypc << 17) | ((uint32_t)headpc << 18);
    }
    static SDF_HD int m_lo(uint32_t m) { return (int)(m & 127u); }
    static SDF_HD int m_hi(uint32_t m) { return (int)((m >> 8) & 127u); }
    static SDF_HD bool m_tail(uint32_t m) { return (m >> 16) & 1u; }
    static SDF_HD bool m_any(uint32_t m) { return (m >> 17) & 1u; }
    static SDF_HD bool m_head(uint32_t m) { return (m >> 18) & 1u; }

    static SDF_HD int sidx(int b, int k, int c) { return ((b << LOGB) + k) * LPC + c; }
    static SDF_HD int midx(int b, int c) { return b * LPC + c; }
    // A hull vertex is addressed by its SLOT u = b*B + k (band b, index k in that band's stack).
    static SDF_HD uint32_t ldU(const BandTile& T, int u, int c) { return bh_ld32(T.S + 4u * (uint32_t)(u * LPC + c)); }
    static SDF_HD uint32_t ldS(const BandTile& T, int b, int k, int c) { return ldU(T, (b << LOGB) + k, c); }
    static SDF_HD uint32_t ldM(const BandTile& T, int b, int c) { return bh_ld32(T.meta + 4u * (uint32_t)midx(b, c)); }
    static SDF_HD void stM(const BandTile& T, int b, int c, uint32_t v) { bh_st32(T.meta + 4u * (uint32_t)midx(b, c), v); }
    static SDF_HD int u_pos(int u, uint32_t e) { return (u & ~(B - 1)) + e_rel(e); }

    // First vertex of the run that enters band b0 from the left (no polarity change up to it), bands [b0, ghi):
    // returns its slot (and the packed entry / the meta word of its band) or -1.
    static SDF_HD int first_from(const BandTile& T, int c, int b0, int ghi, uint32_t& e, uint32_t& m) {
        for (int b2 = b0; b2 < ghi; ++b2) {
            BH_STEP();
            m = ldM(T, b2, c);
            if (m_hi(m) > m_lo(m)) {
                const int u = (b2 << LOGB) + m_lo(m);
                e = ldU(T, u, c);
                return e_rs(e) ? -1 : u;
            }
            if (m_any(m)) return -1;
        }
        return -1;
    }
    // Last vertex of the run that leaves band b0 to the right (no polarity change after it), bands [glo, b0].
    static SDF_HD int last_upto(const BandTile& T, int c, int b0, int glo, uint32_t& e, uint32_t& m) {
        for (int b2 = b0; b2 >= glo; --b2) {
            BH_STEP();
            m = ldM(T, b2, c);
            if (m_tail(m)) return -1;
            if (m_hi(m) > m_lo(m)) {
                const int u = (b2 << LOGB) + m_hi(m) - 1;
                e = ldU(T, u, c);
                return u;
            }
            if (m_any(m)) return -1;
        }
        return -1;
    }
    // Predecessor of vertex u (entry eu, band meta mu) inside the same run, not below band glo; -1 if none.
    static SDF_HD int prev_of(const BandTile& T, int c, int glo, int u, uint32_t eu, uint32_t mu, uint32_t& e, uint32_t& m) {
        if (e_rs(eu)) return -1;                              // u is the first vertex of its run
        if ((u & (B - 1)) > m_lo(mu)) { m = mu; e = ldU(T, u - 1, c); return u - 1; }
        return last_upto(T, c, (u >> LOGB) - 1, glo, e, m);
    }
    // Successor of vertex u (band meta mu) inside the same run, below band ghi; -1 if none.
    static SDF_HD int next_of(const BandTile& T, int c, int ghi, int u, uint32_t mu, uint32_t& e, uint32_t& m) {
        if ((u & (B - 1)) + 1 < m_hi(mu)) {
            m = mu; e = ldU(T, u + 1, c);
            return e_rs(e) ? -1 : u + 1;
        }
        if (m_tail(mu)) return -1;
        return first_from(T, c, (u >> LOGB) + 1, ghi, e, m);
    }

    // ------------------------------------------------------------------------------------------- phase build
    static SDF_HD int pcw_lo(pcw_t m) { return (LOGB > 5) ? bh_lobit64((uint64_t)m) : bh_lobit((uint32_t)m); }
    static SDF_HD int pcw_hi(pcw_t m) { return (LOGB > 5) ? bh_hibit64((uint64_t)m) : bh_hibit((uint32_t)m); }
    static SDF_HD pcw_t low_incl(int i) { return (((pcw_t)2) << i) - 1; }          // bits [0, i]  (i = width-1 -> all ones)

    // The band's raw signed values must already sit in its own stack slots (slot i = position b*B + i, SDF_INF for
    // positions >= n): the hull is built IN PLACE -- vertex k <= i overwrites slot k after slot i has been read.
    // prevpol = polarity of position b*B - 1 (for b == 0: polarity of position 0).
    static SDF_HD void build(const BandTile& T, int c, int b, int n, bool prevpol, pcw_t& Pw, pcw_t& PCw) {
        static_assert(LOGB <= 6, "32-bit hull tests inside a band need |dY| * |dq| < 2^31: dY < 2^25, dq < 2^LOGB");
        const int base = b << LOGB;
        int nvalid = n - base;
        nvalid = nvalid < 0 ? 0 : (nvalid > B ? B : nvalid);
        constexpr uint32_t STEP = LPC * 4;            // bytes between consecutive slots of this thread
        const bh_addr_t st = T.S + 4u * (uint32_t)sidx(b, 0, c);
        bh_addr_t rd = st;                            // slot of the position being read
        bh_addr_t top = st;                           // slot the next vertex goes to
        bh_addr_t bot = st;                           // slot of the first vertex of the current run
        pcw_t pwr = 0;                                // polarity bits, shifted in from the right (bit-reversed)
        int32_t vprev = prevpol ? -1 : 0;             // only its sign is used
        int32_t pcacc = 0;                            // sign set <=> a run boundary since the previous vertex
        // Hull in band-relative lifted coordinates (i, i^2 + h): differs from (q, q^2 + h) by an affine map, which
        // leaves every hull test unchanged.  State: top vertex t = (tp, tY) and d = t - s to the vertex below it, or the
        // sentinel d = (0, -1) that makes the test false while the run has fewer than two vertices (needs i > tp).
        int tp = -1, dP = 0;
        int32_t tY = 0, dY = -1;
#pragma unroll 2
        for (int i = 0; i < nvalid; ++i) {
            const int32_t v = (int32_t)bh_ld32(rd);
            rd += STEP;
            if (LOGB > 5) pwr = (pwr << 1) | (pcw_t)((uint32_t)v >> 31); else pwr = (pcw_t)bh_shift_in_sign((uint32_t)pwr, v);
            pcacc |= v ^ vprev;
            vprev = v;
            const int32_t h = v < 0 ? -v : v;
            if (h < SDF_INF) {
                if (pcacc < 0) { bot = top; dY = -1; dP = 0; }
                const int32_t Y = i * i + h;
                int32_t A = Y - tY;
                int Bq = i - tp;
                while (dY * Bq >= A * dP) {           // t on or above the chord s -> new point: pop it
                    BH_TICK(0);
                    top -= STEP;
                    tp -= dP; tY -= dY;               // t <- s
                    if (top - bot >= 2 * STEP) {
                        const uint32_t e = bh_ld32(top - 2 * STEP);
                        const int sp = e_rel(e);
                        dP = tp - sp; dY = tY - (sp * sp + e_h(e));
                    } else { dY = -1; dP = 0; }
                    A = Y - tY; Bq = i - tp;
                }
                bh_st32(top, (uint32_t)h | ((uint32_t)i << REL_SHIFT) | (((uint32_t)pcacc >> 6) & RS_BIT));
                if (top == bot) { dY = -1; dP = 0; } else { dY = A; dP = Bq; }   // first vertex of its run
                top += STEP;
                tp = i; tY = Y;
                pcacc = 0;
            }
        }
        // polarity / run-start words from the shifted-in signs: bit i of pw = polarity of position i
        pcw_t pw;
        if (LOGB > 5) {
            const uint64_t r = (uint64_t)pwr;
            pw = (pcw_t)(((uint64_t)bh_brev((uint32_t)r) << 32) | (uint64_t)bh_brev((uint32_t)(r >> 32)));
        } else {
            pw = (pcw_t)bh_brev((uint32_t)pwr);
        }
        pw = nvalid ? (pcw_t)(pw >> ((LOGB > 5 ? 64 : 32) - nvalid)) : (pcw_t)0;
        const pcw_t valid = nvalid ? low_incl(nvalid - 1) : (pcw_t)0;
        const pcw_t pcw = (pw ^ ((pcw_t)(pw << 1) | (pcw_t)(prevpol ? 1 : 0))) & valid;
        const int k = (int)((top - st) / STEP);
        const bool anypc = pcw != 0;
        const bool tailpc = pcacc < 0;               // a run boundary after the band's last vertex
        stM(T, b, c, mk_meta(0, k, tailpc, anypc, (pcw & 1) != 0));
        if (anypc) {
            bh_st16(T.pcA + 2u * (uint32_t)midx(b, c), base + pcw_hi(pcw));
            bh_st16(T.pcE + 2u * (uint32_t)midx(b, c), base + pcw_lo(pcw));
            bh_or32(T.anyw + 4u * (uint32_t)((b >> 5) * LPC + c), 1u << (b & 31));
        }
        Pw = pw; PCw = pcw;
    }

    // ------------------------------------------------------------------------------------------- phase merge
    // Round lv: thread (c,b) with b = (2i+1)<<lv joins groups [b-2^lv, b) and [b, b+2^lv).  Round 0 only touches bands
    // 2i and 2i+1, which the same warp built when 2*LPC <= 32: it then needs no CTA-wide barrier before it.
    static constexpr int JUMP_AFTER = 2;               // single steps before the tangent walk starts jumping
    static SDF_HD void merge(const BandTile& T, int c, int b, int nb, int lv) {
        const int g = 1 << lv;
        if ((b & (2 * g - 1)) != g) return;
        const uint32_t mR = ldM(T, b, c);
        if (m_head(mR)) return;                        // a run starts exactly at the seam
        const int glo = b - g;
        int ghi = b + g; if (ghi > nb) ghi = nb;
        // Two cursors walk the common tangent.  L / R = current bridge ends (slot, entry, band meta), L2 / R2 = their
        // neighbours further inside (-1: none in the same run).  Bands are trimmed in registers and written back when left.
        uint32_t eL = 0, mLm = 0, eR = 0, mRm = 0, eL2 = 0, mL2 = 0, eR2 = 0, mR2 = 0;
        int uL = -1, uR = -1, uL2 = -1, uR2 = -1;
        bool have = false;
        {   // common case: the two bands next to the seam hold the bridge end points and their inner neighbours
            const uint32_t mL = ldM(T, b - 1, c);
            const int loR = m_lo(mR), hiR = m_hi(mR), loL = m_lo(mL), hiL = m_hi(mL);
            if (hiL > loL && !m_tail(mL) && hiR > loR) {
                eL = ldS(T, b - 1, hiL - 1, c); eR = ldS(T, b, loR, c);
                eL2 = ldS(T, b - 1, hiL >= 2 ? hiL - 2 : 0, c);
                eR2 = ldS(T, b, loR + 1 < B ? loR + 1 : loR, c);
                if (e_rs(eR)) return;                  // the run ends before the first vertex right of the seam
                const bool simpleL2 = (hiL - 2 >= loL) && !e_rs(eL);
                const bool noL2 = e_rs(eL);
                const bool simpleR2 = loR + 1 < hiR;
                if ((simpleL2 || noL2) && (simpleR2 || m_tail(mR))) {
                    have = true;
                    uL = ((b - 1) << LOGB) + hiL - 1; mLm = mL; uL2 = simpleL2 ? uL - 1 : -1; mL2 = mL;
                    uR = (b << LOGB) + loR; mRm = mR; uR2 = (simpleR2 && !e_rs(eR2)) ? uR + 1 : -1; mR2 = mR;
                }
            }
        }
        if (!have) {
            uL = last_upto(T, c, b - 1, glo, eL, mLm);
            if (uL < 0) return;
            uR = first_from(T, c, b, ghi, eR, mRm);
            if (uR < 0) return;
            uL2 = prev_of(T, c, glo, uL, eL, mLm, eL2, mL2);
            uR2 = next_of(T, c, ghi, uR, mRm, eR2, mR2);
        }
        int Lp = u_pos(uL, eL), Rp = u_pos(uR, eR);
        int32_t LY = Lp * Lp + e_h(eL), RY = Rp * Rp + e_h(eR);
        bool dirtyL = false, dirtyR = false;
        int nL = 0, nR = 0;                                // vertices dropped so far on each side
        for (;;) {
            bool changed = false;
            while (uL2 >= 0) {
                const int L2p = u_pos(uL2, eL2);
                const int32_t L2Y = L2p * L2p + e_h(eL2);
                BH_STEP();
                if (!bh_dominated(L2p, L2Y, Lp, LY, Rp, RY)) break;
                dirtyL = true; changed = true;
                // L is dropped.  Dropped vertices form a SUFFIX of the left chain (the test is monotone along a convex
                // chain), so jump instead of stepping when L's band holds a single run: first the whole band (its
                // first vertex against its predecessor in an earlier band), else a binary search inside the band.
                const int kL = uL & (B - 1), loL = m_lo(mLm), bL = uL >> LOGB;
                if (++nL > JUMP_AFTER && !m_any(mLm) && kL - loL >= 2) {
                    uint32_t eP = 0, mP = 0;
                    const int uP = (bL > glo) ? last_upto(T, c, bL - 1, glo, eP, mP) : -1;
                    const uint32_t eF = ldS(T, bL, loL, c);
                    const int Fp = (bL << LOGB) + e_rel(eF);
                    const int32_t FY = Fp * Fp + e_h(eF);
                    bool whole = false;
                    if (uP >= 0) {
                        const int Pp = u_pos(uP, eP);
                        whole = bh_dominated(Pp, Pp * Pp + e_h(eP), Fp, FY, Rp, RY);
                    }
                    if (whole) {
                        BH_TICK(6);
                        stM(T, bL, c, (mLm & ~(127u << 8)) | ((uint32_t)loL << 8));      // band now empty
                        uL = uP; eL = eP; mLm = mP;
                    } else {
                        BH_TICK(7);
                        int a = loL, z = kL;                   // v_a survives, v_z is dropped
                        while (z - a > 1) {
                            const int mid = (a + z) >> 1; BH_STEP();
                            const uint32_t e0 = ldS(T, bL, mid - 1, c), e1 = ldS(T, bL, mid, c);
                            const int p0 = (bL << LOGB) + e_rel(e0), p1 = (bL << LOGB) + e_rel(e1);
                            if (bh_dominated(p0, p0 * p0 + e_h(e0), p1, p1 * p1 + e_h(e1), Rp, RY)) z = mid; else a = mid;
                        }
                        uL = (bL << LOGB) + a; eL = ldS(T, bL, a, c);
                    }
                    Lp = u_pos(uL, eL); LY = Lp * Lp + e_h(eL);
                } else {
                    BH_TICK(1 + (lv < 5 ? lv : 5));
                    if ((uL2 >> LOGB) != bL)                       // leaving a band: it is now empty
                        stM(T, bL, c, (mLm & ~(127u << 8)) | ((uint32_t)loL << 8));
                    uL = uL2; eL = eL2; mLm = mL2; Lp = L2p; LY = L2Y;
                }
                uL2 = prev_of(T, c, glo, uL, eL, mLm, eL2, mL2);
            }
            while (uR2 >= 0) {
                const int R2p = u_pos(uR2, eR2);
                const int32_t R2Y = R2p * R2p + e_h(eR2);
                BH_STEP();
                if (!bh_dominated(Lp, LY, Rp, RY, R2p, R2Y)) break;
                dirtyR = true; changed = true;
                // R is dropped; dropped vertices form a PREFIX of the right chain: same jumps, mirrored.
                const int kR = uR & (B - 1), hiR = m_hi(mRm), bR = uR >> LOGB;
                if (++nR > JUMP_AFTER && !m_any(mRm) && hiR - 1 - kR >= 2) {
                    uint32_t eN = 0, mN = 0;
                    const int uN = (bR + 1 < ghi) ? first_from(T, c, bR + 1, ghi, eN, mN) : -1;
                    const uint32_t eE = ldS(T, bR, hiR - 1, c);
                    const int Ep = (bR << LOGB) + e_rel(eE);
                    const int32_t EY = Ep * Ep + e_h(eE);
                    bool whole = false;
                    if (uN >= 0) {
                        const int Np = u_pos(uN, eN);
                        whole = bh_dominated(Lp, LY, Ep, EY, Np, Np * Np + e_h(eN));
                    }
                    if (whole) {
                        BH_TICK(6);
                        stM(T, bR, c, (mRm & ~127u) | (uint32_t)hiR);                      // band now empty
                        uR = uN; eR = eN; mRm = mN;
                    } else {
                        BH_TICK(7);
                        int a = kR, z = hiR - 1;               // v_a is dropped, v_z survives
                        while (z - a > 1) {
                            const int mid = (a + z) >> 1; BH_STEP();
                            const uint32_t e0 = ldS(T, bR, mid, c), e1 = ldS(T, bR, mid + 1, c);
                            const int p0 = (bR << LOGB) + e_rel(e0), p1 = (bR << LOGB) + e_rel(e1);
                            if (bh_dominated(Lp, LY, p0, p0 * p0 + e_h(e0), p1, p1 * p1 + e_h(e1))) a = mid; else z = mid;
                        }
                        uR = (bR << LOGB) + z; eR = ldS(T, bR, z, c);
                    }
                    Rp = u_pos(uR, eR); RY = Rp * Rp + e_h(eR);
                } else {
                    BH_TICK(1 + (lv < 5 ? lv : 5));
                    if ((uR2 >> LOGB) != bR)
                        stM(T, bR, c, (mRm & ~127u) | (uint32_t)hiR);
                    uR = uR2; eR = eR2; mRm = mR2; Rp = R2p; RY = R2Y;
                }
                uR2 = next_of(T, c, ghi, uR, mRm, eR2, mR2);
            }
            if (!changed) break;
        }
        if (dirtyL) stM(T, uL >> LOGB, c, (mLm & ~(127u << 8)) | ((uint32_t)((uL & (B - 1)) + 1) << 8));
        if (dirtyR) stM(T, uR >> LOGB, c, (mRm & ~127u) | (uint32_t)(uR & (B - 1)));
    }

    // ------------------------------------------------------------------------------------------- phase query
    // Emits out(p) for p in the band through `emit(i, value)`, i = position inside the band, in increasing i.
    static constexpr int32_t NOSITE = 0x60000000;       // height of a non-existent "next" vertex: never wins
    static constexpr int FAR = 40000;                    // distance to a non-existent zero site: FAR^2 > SDF_INF

    template <typename Emit>
    static SDF_HD void query(const BandTile& T, int c, int b, int nb, int n, int nlevels, pcw_t Pw, pcw_t PCw, Emit emit) {
        const int base = b << LOGB;
        if (base >= n) return;
        int nvalid = n - base;
        nvalid = nvalid > B ? B : nvalid;
        (void)nlevels;
        // nearest run start before / after the band, through the line's mask of bands that hold one
        int a0 = -1, e_after = BH_BIG;
        {
            int w = b >> 5;
            uint32_t m = bh_ld32(T.anyw + 4u * (uint32_t)(w * LPC + c)) & ((1u << (b & 31)) - 1u);
            while (!m && w > 0) { --w; m = bh_ld32(T.anyw + 4u * (uint32_t)(w * LPC + c)); }
            if (m) a0 = bh_ld16(T.pcA + 2u * (uint32_t)midx((w << 5) + bh_hibit(m), c));
            w = b >> 5;
            m = bh_ld32(T.anyw + 4u * (uint32_t)(w * LPC + c)) & ~((2u << (b & 31)) - 1u);
            const int nw = (nb + 31) >> 5;
            while (!m && w + 1 < nw) { ++w; m = bh_ld32(T.anyw + 4u * (uint32_t)(w * LPC + c)); }
            if (m) e_after = bh_ld16(T.pcE + 2u * (uint32_t)midx((w << 5) + bh_lobit(m), c));
        }
        const uint32_t mym = ldM(T, b, c);
        const int myhi = base + m_hi(mym);
        int us = base + m_lo(mym);                          // scan slot over this band's own vertices
        // current / next hull vertex as scalars; "none" is encoded in the heights.
        // cu = slot of cur, cm = meta of its band; nu/ne/nm = slot, entry, band meta of next.
        int cu = base, cpos = base, nu = -1, npos = 0;
        uint32_t cm = mym, nm = mym;
        int32_t ch = SDF_INF, nh = NOSITE;
        int e;
        {
            const pcw_t rest = PCw & ~(pcw_t)1;
            e = rest ? base + pcw_lo(rest) : e_after;
        }
        int dl = (a0 >= 1) ? base - a0 + 1 : FAR;           // p - a + 1: distance to the zero site before the run
        int dr = (e < n) ? e - base : FAR;                  // e - p:     distance to the zero site after the run
        if (!(PCw & 1)) {                                   // position `base` continues a run from the left
            uint32_t eLq = 0, mLq = 0, eRq = 0, mRq = 0;
            const int uLq = (b > 0) ? last_upto(T, c, b - 1, 0, eLq, mLq) : -1;
            const int uRq = first_from(T, c, b, nb, eRq, mRq);
            int u = -1; uint32_t eu = 0, mu = 0;
            if (uRq >= 0) {
                const int rp = u_pos(uRq, eRq);
                bool takeR = uLq < 0;
                if (!takeR) { const int lp = u_pos(uLq, eLq); takeR = (base - rp) * (base - rp) + e_h(eRq) <= (base - lp) * (base - lp) + e_h(eLq); }
                if (takeR) { u = uRq; eu = eRq; mu = mRq; }
            }
            if (u < 0 && uLq >= 0) {
                u = uLq; eu = eLq; mu = mLq;
                for (;;) {                                  // local search towards the minimiser of `base`
                    uint32_t ep = 0, mp = 0;
                    const int up = prev_of(T, c, 0, u, eu, mu, ep, mp);
                    if (up < 0) break;
                    const int dc = base - u_pos(u, eu), dp = base - u_pos(up, ep);
                    if (dp * dp + e_h(ep) < dc * dc + e_h(eu)) { u = up; eu = ep; mu = mp; } else break;
                }
            }
            if (u >= 0) {
                cu = u; cm = mu; cpos = u_pos(u, eu); ch = e_h(eu);
                uint32_t en = 0;
                nu = next_of(T, c, nb, cu, cm, en, nm);
                if (nu >= 0) { npos = u_pos(nu, en); nh = e_h(en); }
            }
        }
#pragma unroll 2
        for (int i = 0; i < nvalid; ++i) {
            const int p = base + i;
            if ((PCw >> i) & 1) {                           // a new run starts at p
                const pcw_t rest = (i + 1 < B) ? (PCw >> (i + 1)) : (pcw_t)0;
                e = rest ? p + 1 + pcw_lo(rest) : e_after;
                dl = 1;
                dr = (e < n) ? e - p : FAR;
                ch = SDF_INF; nh = NOSITE; cpos = p; nu = -1;
                int u = -1; uint32_t eu = 0, mu = mym;
                while (us < myhi && e_rel(ldU(T, us, c)) < i) ++us;            // vertices of earlier runs
                if (us < myhi) {
                    eu = ldU(T, us, c);
                    if (u_pos(us, eu) < e) u = us;
                } else if (!rest && e_after > ((b + 1) << LOGB) && b + 1 < nb) {
                    // the run continues into later bands and has no vertex here: its first vertex there
                    u = first_from(T, c, b + 1, nb, eu, mu);
                }
                if (u >= 0) {
                    cu = u; cm = mu; cpos = u_pos(u, eu); ch = e_h(eu);
                    uint32_t en = 0;
                    nu = next_of(T, c, nb, cu, cm, en, nm);
                    if (nu >= 0) { npos = u_pos(nu, en); nh = e_h(en); }
                }
            }
            int dc = p - cpos, dn = p - npos;
            int32_t vc = dc * dc + ch;
            int32_t vn = dn * dn + nh;
            while (vn <= vc) {                              // the next vertex takes over
                cu = nu; cm = nm; cpos = npos; ch = nh;
                vc = vn;
                nh = NOSITE;
                if ((cu & (B - 1)) + 1 < m_hi(cm)) {        // fast path: the following slot of the same band
                    const uint32_t en = ldU(T, cu + 1, c);
                    nu = cu + 1; npos = u_pos(nu, en);
                    if (!e_rs(en)) nh = e_h(en);
                } else {
                    uint32_t en = 0;
                    nu = next_of(T, c, nb, cu, cm, en, nm);
                    if (nu >= 0) { npos = u_pos(nu, en); nh = e_h(en); }
                }
                dn = p - npos;
                vn = dn * dn + nh;
            }
            int32_t best = vc < SDF_INF ? vc : SDF_INF;
            const int32_t zl = dl * dl, zr = dr * dr;
            best = zl < best ? zl : best;
            best = zr < best ? zr : best;
            ++dl; --dr;
            emit(i, ((Pw >> i) & 1) ? -best : best);
        }
    }
};

}  // namespace sdfb200
