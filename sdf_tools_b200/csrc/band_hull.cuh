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
//   phase build   sequential hull (Felzenszwalb stack) of the band's own positions; the stack lives in shared memory
//                 at S[(b*B + k)*LPC + c] -- bank == line, so lanes never conflict whatever their stack depths --
//                 as packed words {|s| : 24, position-in-band : 6, first-of-run flag : 1}.  Restarted at every
//                 polarity change.
//   phase merge   log2(NB) rounds; round lv joins band groups of 2^lv pairwise by walking the common tangent
//                 across the seam (only when a run continues across it); trims band stacks from their ends.
//                 The same rounds run a doubling scan that gives every band the nearest run boundary before/after it.
//   phase query   each thread walks the finished hull for its own band's positions (monotone minimiser),
//                 adds the two run-end terms and writes the result row (coalesced across lines).
// Phases only communicate through the shared arrays and are race-free inside a phase, so the same functions run on
// the GPU (separated by __syncthreads) and, for tests without a GPU, on the CPU as plain loops over (c, b).
//
// Exactness: all integer.  |s| <= 3*2048^2 < 2^24, lifted heights < 2^25, hull tests cross-multiplied in int64.
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

constexpr int BH_NONE = -1;
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
#else
typedef uintptr_t bh_addr_t;
SDF_HD uint32_t bh_ld32(bh_addr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
SDF_HD void bh_st32(bh_addr_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(a) = v; }
SDF_HD int bh_ld16(bh_addr_t a) { return *reinterpret_cast<const int16_t*>(a); }
SDF_HD void bh_st16(bh_addr_t a, int v) { *reinterpret_cast<int16_t*>(a) = (int16_t)v; }
#endif

struct BandTile {
    bh_addr_t S;      // uint32 [NB*B][LPC]  packed hull entries
    bh_addr_t meta;   // uint32 [NB][LPC]    lo:7 | hi:7 <<8 | tailPC <<16 | anyPC <<17 | headPC <<18
    bh_addr_t pcA;    // int16  [2][NB][LPC] last run start inside bands <= b (double-buffered doubling scan)
    bh_addr_t pcE;    // int16  [2][NB][LPC] first run start inside bands >= b
    int pcstride;     // NB*LPC: elements between the two scan buffers
};

template <int LOGB, int LPC>
struct BandHull {
    static constexpr int B = 1 << LOGB;
    using pcw_t = typename std::conditional<(LOGB > 5), uint64_t, uint32_t>::type;

    // ---- packed hull entry ------------------------------------------------------------------------------
    static SDF_HD uint32_t pack(int rel, int32_t h, bool rs) { return (uint32_t)h | ((uint32_t)rel << 24) | ((uint32_t)rs << 30); }
    static SDF_HD int32_t e_h(uint32_t e) { return (int32_t)(e & 0xFFFFFFu); }
    static SDF_HD int e_rel(uint32_t e) { return (int)((e >> 24) & 63u); }
    static SDF_HD bool e_rs(uint32_t e) { return (e >> 30) & 1u; }
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
    static SDF_HD uint32_t ldS(const BandTile& T, int b, int k, int c) { return bh_ld32(T.S + 4u * (uint32_t)sidx(b, k, c)); }
    static SDF_HD uint32_t ldM(const BandTile& T, int b, int c) { return bh_ld32(T.meta + 4u * (uint32_t)midx(b, c)); }
    static SDF_HD void stM(const BandTile& T, int b, int c, uint32_t v) { bh_st32(T.meta + 4u * (uint32_t)midx(b, c), v); }

    // A hull vertex reference: band, index inside the band stack, and its decoded position / height.
    struct Ref { int b, k, pos; int32_t h; };

    static SDF_HD Ref load_ref(const BandTile& T, int c, int b, int k) {
        const uint32_t e = ldS(T, b, k, c);
        Ref r; r.b = b; r.k = k; r.pos = (b << LOGB) + e_rel(e); r.h = e_h(e);
        return r;
    }

    // Predecessor of (b,k) inside the same run, not below band glo.  Returns false if there is none.
    static SDF_HD bool prev_entry(const BandTile& T, int c, int glo, const Ref& x, Ref& out) {
        const uint32_t e = ldS(T, x.b, x.k, c);
        if (e_rs(e)) return false;                       // x is the first vertex of its run
        const uint32_t m = ldM(T, x.b, c);
        if (x.k > m_lo(m)) { out = load_ref(T, c, x.b, x.k - 1); return true; }
        for (int b2 = x.b - 1; b2 >= glo; --b2) {
            const uint32_t m2 = ldM(T, b2, c);
            if (m_tail(m2)) return false;
            if (m_hi(m2) > m_lo(m2)) { out = load_ref(T, c, b2, m_hi(m2) - 1); return true; }
            if (m_any(m2)) return false;
        }
        return false;
    }

    // Successor of (b,k) inside the same run, below band ghi.
    static SDF_HD bool next_entry(const BandTile& T, int c, int ghi, const Ref& x, Ref& out) {
        const uint32_t m = ldM(T, x.b, c);
        if (x.k + 1 < m_hi(m)) {
            const uint32_t e = ldS(T, x.b, x.k + 1, c);
            if (e_rs(e)) return false;
            out.b = x.b; out.k = x.k + 1; out.pos = (x.b << LOGB) + e_rel(e); out.h = e_h(e);
            return true;
        }
        if (m_tail(m)) return false;
        return first_from(T, c, x.b + 1, ghi, out);
    }

    // First vertex of the run that enters band b0 from the left (no polarity change up to it), bands [b0, ghi).
    static SDF_HD bool first_from(const BandTile& T, int c, int b0, int ghi, Ref& out) {
        for (int b2 = b0; b2 < ghi; ++b2) {
            const uint32_t m2 = ldM(T, b2, c);
            if (m_hi(m2) > m_lo(m2)) {
                const uint32_t e = ldS(T, b2, m_lo(m2), c);
                if (e_rs(e)) return false;
                out.b = b2; out.k = m_lo(m2); out.pos = (b2 << LOGB) + e_rel(e); out.h = e_h(e);
                return true;
            }
            if (m_any(m2)) return false;
        }
        return false;
    }

    // Last vertex of the run that leaves band b0 to the right (no polarity change after it), bands [glo, b0].
    static SDF_HD bool last_upto(const BandTile& T, int c, int b0, int glo, Ref& out) {
        for (int b2 = b0; b2 >= glo; --b2) {
            const uint32_t m2 = ldM(T, b2, c);
            if (m_tail(m2)) return false;
            if (m_hi(m2) > m_lo(m2)) { out = load_ref(T, c, b2, m_hi(m2) - 1); return true; }
            if (m_any(m2)) return false;
        }
        return false;
    }

    // ------------------------------------------------------------------------------------------- phase build
    // in[i] = signed value at position b*B + i (SDF_INF for positions >= n); prevpol = polarity of position
    // b*B - 1 (for b == 0: polarity of position 0).  Outputs the band's polarity / run-start bit words.
    static SDF_HD int pcw_lo(pcw_t m) { return (LOGB > 5) ? bh_lobit64((uint64_t)m) : bh_lobit((uint32_t)m); }
    static SDF_HD int pcw_hi(pcw_t m) { return (LOGB > 5) ? bh_hibit64((uint64_t)m) : bh_hibit((uint32_t)m); }
    static SDF_HD pcw_t low_incl(int i) { return (((pcw_t)2) << i) - 1; }          // bits [0, i]  (i = width-1 -> all ones)

    static SDF_HD void build(const BandTile& T, int c, int b, int n, const int32_t* in, bool prevpol,
                             pcw_t& Pw, pcw_t& PCw) {
        const int base = b << LOGB;
        // polarity word: one funnel shift per element, bit-reversed at the end
        pcw_t pw;
        {
            uint32_t w0 = 0, w1 = 0;
#pragma unroll
            for (int i = 0; i < B; ++i) {
                if (i < 32) w0 = bh_shift_in_sign(w0, in[i]); else w1 = bh_shift_in_sign(w1, in[i]);
            }
            w0 = bh_brev(w0);
            if (B < 32) w0 >>= (32 - (B < 32 ? B : 0)) & 31;
            if constexpr (LOGB > 5) { w1 = bh_brev(w1); pw = (pcw_t)w0 | ((pcw_t)w1 << 32); } else { pw = w0; (void)w1; }
        }
        int nvalid = n - base;
        nvalid = nvalid < 0 ? 0 : (nvalid > B ? B : nvalid);
        const pcw_t lim = nvalid ? low_incl(nvalid - 1) : (pcw_t)0;
        pw &= lim;
        const pcw_t pcw = (pw ^ ((pw << 1) | (pcw_t)prevpol)) & lim;
        const bool anypc = pcw != 0;

        const bh_addr_t st = T.S + 4u * (uint32_t)sidx(b, 0, c);   // this thread's stack: slot k at st + k*LPC*4
        int k = 0, cnt = 0;                           // entries in the band / in the current run
        int tp = 0, sp = 0, dPts = 1;
        int32_t tY = 0, sY = 0, dYts = 0;
        pcw_t notbelow = ~(pcw_t)0;                   // positions after the previous vertex of this band
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const int32_t v = in[i];
            const int32_t h = v < 0 ? -v : v;
            if (h < SDF_INF) {
                const bool pcf = (pcw & notbelow & low_incl(i)) != 0;     // a run boundary since the previous vertex
                if (pcf) cnt = 0;
                const int q = base + i;
                const int32_t Y = q * q + h;
                int32_t A = Y - tY;
                int Bq = q - tp;
                while (cnt >= 2 && (long long)dYts * (long long)Bq >= (long long)A * (long long)dPts) {
                    --k; --cnt;
                    tp = sp; tY = sY;
                    if (cnt >= 2) {
                        const uint32_t e = bh_ld32(st + (uint32_t)((k - 2) * (LPC * 4)));
                        sp = base + e_rel(e);
                        sY = sp * sp + e_h(e);
                    }
                    dYts = tY - sY; dPts = tp - sp;
                    A = Y - tY; Bq = q - tp;
                }
                bh_st32(st + (uint32_t)(k * (LPC * 4)), (uint32_t)h | ((uint32_t)i << 24) | ((uint32_t)pcf << 30));
                ++k; ++cnt;
                sp = tp; sY = tY; tp = q; tY = Y;
                dYts = A; dPts = Bq;
                notbelow = ~low_incl(i);
            }
        }
        const bool tailpc = (pcw & notbelow) != 0;
        stM(T, b, c, mk_meta(0, k, tailpc, anypc, (pcw & 1) != 0));
        bh_st16(T.pcA + 2u * (uint32_t)midx(b, c), anypc ? base + pcw_hi(pcw) : -1);
        bh_st16(T.pcE + 2u * (uint32_t)midx(b, c), anypc ? base + pcw_lo(pcw) : BH_BIG);
        Pw = pw; PCw = pcw;
    }

    // ------------------------------------------------------------------------------------------- phase merge
    // Round lv: (a) thread (c,b) with b = (2i+1)<<lv joins groups [b-2^lv, b) and [b, b+2^lv);
    //           (b) every thread advances the run-boundary doubling scans by distance 2^lv (buffers lv&1 -> (lv+1)&1).
    static SDF_HD void merge(const BandTile& T, int c, int b, int nb, int lv) {
        // (b) scans: inclusive prefix-max of last run start, inclusive suffix-min of first run start
        {
            const int d = 1 << lv, src = lv & 1, dst = src ^ 1;
            const bh_addr_t As = T.pcA + 2u * (uint32_t)(src * T.pcstride), Ad = T.pcA + 2u * (uint32_t)(dst * T.pcstride);
            const bh_addr_t Es = T.pcE + 2u * (uint32_t)(src * T.pcstride), Ed = T.pcE + 2u * (uint32_t)(dst * T.pcstride);
            int a = bh_ld16(As + 2u * (uint32_t)midx(b, c));
            if (b - d >= 0) { const int o = bh_ld16(As + 2u * (uint32_t)midx(b - d, c)); a = o > a ? o : a; }
            bh_st16(Ad + 2u * (uint32_t)midx(b, c), a);
            int e = bh_ld16(Es + 2u * (uint32_t)midx(b, c));
            if (b + d < nb) { const int o = bh_ld16(Es + 2u * (uint32_t)midx(b + d, c)); e = o < e ? o : e; }
            bh_st16(Ed + 2u * (uint32_t)midx(b, c), e);
        }
        // (a) seam
        const int g = 1 << lv;
        if ((b & (2 * g - 1)) != g) return;
        if (m_head(ldM(T, b, c))) return;            // a run starts exactly at the seam
        const int glo = b - g;
        int ghi = b + g; if (ghi > nb) ghi = nb;
        Ref L, R, L2, R2;
        if (!last_upto(T, c, b - 1, glo, L)) return;
        if (!first_from(T, c, b, ghi, R)) return;
        bool hasL2 = prev_entry(T, c, glo, L, L2);
        bool hasR2 = next_entry(T, c, ghi, R, R2);
        int32_t LY = L.pos * L.pos + L.h, RY = R.pos * R.pos + R.h;
        for (;;) {
            bool changed = false;
            while (hasL2 && bh_dominated(L2.pos, L2.pos * L2.pos + L2.h, L.pos, LY, R.pos, RY)) {
                stM(T, L.b, c, ldM(T, L.b, c) - (1u << 8));  // drop L: trim its band from the back
                // bands between L2 and L that became empty stay empty; L2 becomes the tail
                L = L2; LY = L.pos * L.pos + L.h;
                hasL2 = prev_entry(T, c, glo, L, L2);
                changed = true;
            }
            while (hasR2 && bh_dominated(L.pos, LY, R.pos, RY, R2.pos, R2.pos * R2.pos + R2.h)) {
                stM(T, R.b, c, ldM(T, R.b, c) + 1u);         // drop R: trim its band from the front
                R = R2; RY = R.pos * R.pos + R.h;
                hasR2 = next_entry(T, c, ghi, R, R2);
                changed = true;
            }
            if (!changed) break;
        }
    }

    // ------------------------------------------------------------------------------------------- phase query
    // Emits out(p) for p in the band through `emit(i, value)`, i = position inside the band, in increasing i.
    static constexpr int32_t NOSITE = 0x60000000;       // height of a non-existent "next" vertex: never wins
    static constexpr int FAR = 40000;                    // distance to a non-existent zero site: FAR^2 > SDF_INF

    // successor of vertex (cb, ck) of the same run -> (nb_, nk, npos, nh); nh = NOSITE if there is none
    static SDF_HD void load_next(const BandTile& T, int c, int nbands, int cb, int ck, int& nb_, int& nk, int& npos, int32_t& nh) {
        const uint32_t m = ldM(T, cb, c);
        if (ck + 1 < m_hi(m)) {
            const uint32_t e = ldS(T, cb, ck + 1, c);
            nb_ = cb; nk = ck + 1; npos = (cb << LOGB) + e_rel(e);
            nh = e_rs(e) ? NOSITE : e_h(e);
            return;
        }
        Ref r;
        if (!m_tail(m) && first_from(T, c, cb + 1, nbands, r)) { nb_ = r.b; nk = r.k; npos = r.pos; nh = r.h; }
        else { nb_ = cb; nk = ck; npos = 0; nh = NOSITE; }
    }

    template <typename Emit>
    static SDF_HD void query(const BandTile& T, int c, int b, int nb, int n, int nlevels, pcw_t Pw, pcw_t PCw, Emit emit) {
        const int base = b << LOGB;
        if (base >= n) return;
        int nvalid = n - base;
        nvalid = nvalid > B ? B : nvalid;
        const int fin = nlevels & 1;                        // buffer holding the finished scans
        const int a0 = (b > 0) ? bh_ld16(T.pcA + 2u * (uint32_t)(fin * T.pcstride + midx(b - 1, c))) : -1;         // last run start before the band
        const int e_after = (b + 1 < nb) ? bh_ld16(T.pcE + 2u * (uint32_t)(fin * T.pcstride + midx(b + 1, c))) : BH_BIG;   // first run start after it
        const uint32_t mym = ldM(T, b, c);
        const int myhi = m_hi(mym);
        int ks = m_lo(mym);                                 // scan pointer over this band's own vertices
        // current / next hull vertex as scalars; "none" is encoded in the heights
        int cb = b, ck = 0, cpos = base, nb_ = b, nk = 0, npos = 0;
        int32_t ch = SDF_INF, nh = NOSITE;
        int e;
        {
            const pcw_t rest = PCw & ~(pcw_t)1;
            e = rest ? base + pcw_lo(rest) : e_after;
        }
        int dl = (a0 >= 1) ? base - a0 + 1 : FAR;           // p - a + 1: distance to the zero site before the run
        int dr = (e < n) ? e - base : FAR;                  // e - p:     distance to the zero site after the run
        if (!(PCw & 1)) {                                   // position `base` continues a run from the left
            Ref Lr, Rr;
            const bool hasL = (b > 0) && last_upto(T, c, b - 1, 0, Lr);
            const bool hasR = first_from(T, c, b, nb, Rr);
            bool have = false;
            Ref cur;
            if (hasR && (!hasL || (base - Rr.pos) * (base - Rr.pos) + Rr.h <= (base - Lr.pos) * (base - Lr.pos) + Lr.h)) {
                cur = Rr; have = true;
            } else if (hasL) {
                cur = Lr; have = true;
                for (;;) {                                  // local search towards the minimiser of `base`
                    Ref pr;
                    if (!prev_entry(T, c, 0, cur, pr)) break;
                    const int dc = base - cur.pos, dp = base - pr.pos;
                    if (dp * dp + pr.h < dc * dc + cur.h) cur = pr; else break;
                }
            }
            if (have) {
                cb = cur.b; ck = cur.k; cpos = cur.pos; ch = cur.h;
                load_next(T, c, nb, cb, ck, nb_, nk, npos, nh);
            }
        }
#pragma unroll 8
        for (int i = 0; i < B; ++i) {
            if (i >= nvalid) break;
            const int p = base + i;
            if ((PCw >> i) & 1) {                           // a new run starts at p
                const pcw_t rest = (i + 1 < B) ? (PCw >> (i + 1)) : (pcw_t)0;
                e = rest ? p + 1 + pcw_lo(rest) : e_after;
                dl = 1;
                dr = (e < n) ? e - p : FAR;
                ch = SDF_INF; nh = NOSITE; cpos = p;
                bool have = false;
                Ref r;
                while (ks < myhi && e_rel(ldS(T, b, ks, c)) < i) ++ks;      // vertices of earlier runs
                if (ks < myhi) {
                    r = load_ref(T, c, b, ks);
                    have = r.pos < e;
                } else if (!rest && e_after > ((b + 1) << LOGB)) {
                    // the run continues into later bands and has no vertex here: its first vertex there
                    have = (b + 1 < nb) && first_from(T, c, b + 1, nb, r);
                }
                if (have) {
                    cb = r.b; ck = r.k; cpos = r.pos; ch = r.h;
                    load_next(T, c, nb, cb, ck, nb_, nk, npos, nh);
                }
            }
            int dc = p - cpos, dn = p - npos;
            int32_t vc = dc * dc + ch;
            int32_t vn = dn * dn + nh;
            while (vn <= vc) {                              // the next vertex takes over
                cb = nb_; ck = nk; cpos = npos; ch = nh;
                vc = vn;
                load_next(T, c, nb, cb, ck, nb_, nk, npos, nh);
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
