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

// t on or above the chord s->u of the lifted points  <=>  parabola t never wins
SDF_HD bool bh_dominated(int sp, int32_t sY, int tp, int32_t tY, int up, int32_t uY) {
    const long long lhs = (long long)(tY - sY) * (long long)(up - tp);
    const long long rhs = (long long)(uY - tY) * (long long)(tp - sp);
    return lhs >= rhs;
}

constexpr int BH_NONE = -1;
constexpr int BH_BIG = 0x7fff;          // "no run boundary after": larger than any position

// Shared-memory arrays of one tile (pointers into the CTA's dynamic shared memory).
struct BandTile {
    uint32_t* S;      // [NB*B][LPC]  packed hull entries
    uint32_t* meta;   // [NB][LPC]    lo:7 | hi:7 <<8 | tailPC <<16 | anyPC <<17 | headPC <<18
    int16_t* pcA[2];  // [NB][LPC] x2 last run start inside bands <= b (double-buffered doubling scan)
    int16_t* pcE[2];  // [NB][LPC] x2 first run start inside bands >= b
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

    // A hull vertex reference: band, index inside the band stack, and its decoded position / height.
    struct Ref { int b, k, pos; int32_t h; };

    static SDF_HD Ref load_ref(const BandTile& T, int c, int b, int k) {
        const uint32_t e = T.S[sidx(b, k, c)];
        Ref r; r.b = b; r.k = k; r.pos = (b << LOGB) + e_rel(e); r.h = e_h(e);
        return r;
    }

    // Predecessor of (b,k) inside the same run, not below band glo.  Returns false if there is none.
    static SDF_HD bool prev_entry(const BandTile& T, int c, int glo, const Ref& x, Ref& out) {
        const uint32_t e = T.S[sidx(x.b, x.k, c)];
        if (e_rs(e)) return false;                       // x is the first vertex of its run
        const uint32_t m = T.meta[midx(x.b, c)];
        if (x.k > m_lo(m)) { out = load_ref(T, c, x.b, x.k - 1); return true; }
        for (int b2 = x.b - 1; b2 >= glo; --b2) {
            const uint32_t m2 = T.meta[midx(b2, c)];
            if (m_tail(m2)) return false;
            if (m_hi(m2) > m_lo(m2)) { out = load_ref(T, c, b2, m_hi(m2) - 1); return true; }
            if (m_any(m2)) return false;
        }
        return false;
    }

    // Successor of (b,k) inside the same run, below band ghi.
    static SDF_HD bool next_entry(const BandTile& T, int c, int ghi, const Ref& x, Ref& out) {
        const uint32_t m = T.meta[midx(x.b, c)];
        if (x.k + 1 < m_hi(m)) {
            const uint32_t e = T.S[sidx(x.b, x.k + 1, c)];
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
            const uint32_t m2 = T.meta[midx(b2, c)];
            if (m_hi(m2) > m_lo(m2)) {
                const uint32_t e = T.S[sidx(b2, m_lo(m2), c)];
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
            const uint32_t m2 = T.meta[midx(b2, c)];
            if (m_tail(m2)) return false;
            if (m_hi(m2) > m_lo(m2)) { out = load_ref(T, c, b2, m_hi(m2) - 1); return true; }
            if (m_any(m2)) return false;
        }
        return false;
    }

    // ------------------------------------------------------------------------------------------- phase build
    // in[i] = signed value at position b*B + i (i < B, positions >= n ignored); prevpol = polarity of position
    // b*B - 1 (for b == 0: polarity of position 0).  Outputs the band's polarity / run-start bit words.
    static SDF_HD void build(const BandTile& T, int c, int b, int n, const int32_t* in, bool prevpol,
                             pcw_t& Pw, pcw_t& PCw) {
        const int base = b << LOGB;
        int k = 0, runbase = 0;
        bool pcflag = false, anypc = false, headpc = false;
        int lastpc = -1, firstpc = BH_BIG;
        int tp = 0, sp = 0;
        int32_t tY = 0, sY = 0;
        pcw_t pw = 0, pcw = 0;
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const int q = base + i;
            if (q < n) {
                const int32_t v = in[i];
                const bool pol = v < 0;
                const int32_t h = pol ? -v : v;
                if (pol) pw |= (pcw_t)1 << i;
                if (pol != prevpol) {
                    prevpol = pol;
                    pcw |= (pcw_t)1 << i;
                    pcflag = true; anypc = true;
                    if (i == 0) headpc = true;
                    runbase = k;
                    lastpc = q;
                    if (firstpc == BH_BIG) firstpc = q;
                }
                if (h < SDF_INF) {
                    const int32_t Y = q * q + h;
                    while (k - runbase >= 2 && bh_dominated(sp, sY, tp, tY, q, Y)) {
                        --k;
                        tp = sp; tY = sY;
                        if (k - runbase >= 2) {
                            const uint32_t e = T.S[sidx(b, k - 2, c)];
                            sp = base + e_rel(e);
                            sY = sp * sp + e_h(e);
                        }
                    }
                    T.S[sidx(b, k, c)] = pack(i, h, pcflag);
                    pcflag = false;
                    sp = tp; sY = tY; tp = q; tY = Y;
                    ++k;
                }
            }
        }
        T.meta[midx(b, c)] = mk_meta(0, k, pcflag, anypc, headpc);
        T.pcA[0][midx(b, c)] = (int16_t)lastpc;
        T.pcE[0][midx(b, c)] = (int16_t)firstpc;
        Pw = pw; PCw = pcw;
    }

    // ------------------------------------------------------------------------------------------- phase merge
    // Round lv: (a) thread (c,b) with b = (2i+1)<<lv joins groups [b-2^lv, b) and [b, b+2^lv);
    //           (b) every thread advances the run-boundary doubling scans by distance 2^lv (buffers lv&1 -> (lv+1)&1).
    static SDF_HD void merge(const BandTile& T, int c, int b, int nb, int lv) {
        // (b) scans: inclusive prefix-max of last run start, inclusive suffix-min of first run start
        {
            const int d = 1 << lv, src = lv & 1, dst = src ^ 1;
            int16_t a = T.pcA[src][midx(b, c)];
            if (b - d >= 0) { const int16_t o = T.pcA[src][midx(b - d, c)]; a = o > a ? o : a; }
            T.pcA[dst][midx(b, c)] = a;
            int16_t e = T.pcE[src][midx(b, c)];
            if (b + d < nb) { const int16_t o = T.pcE[src][midx(b + d, c)]; e = o < e ? o : e; }
            T.pcE[dst][midx(b, c)] = e;
        }
        // (a) seam
        const int g = 1 << lv;
        if ((b & (2 * g - 1)) != g) return;
        if (m_head(T.meta[midx(b, c)])) return;            // a run starts exactly at the seam
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
                const int mi = midx(L.b, c);                // drop L: trim its band from the back
                T.meta[mi] = T.meta[mi] - (1u << 8);
                // bands between L2 and L that became empty stay empty; L2 becomes the tail
                L = L2; LY = L.pos * L.pos + L.h;
                hasL2 = prev_entry(T, c, glo, L, L2);
                changed = true;
            }
            while (hasR2 && bh_dominated(L.pos, LY, R.pos, RY, R2.pos, R2.pos * R2.pos + R2.h)) {
                const int mi = midx(R.b, c);                // drop R: trim its band from the front
                T.meta[mi] = T.meta[mi] + 1u;
                R = R2; RY = R.pos * R.pos + R.h;
                hasR2 = next_entry(T, c, ghi, R, R2);
                changed = true;
            }
            if (!changed) break;
        }
    }

    // ------------------------------------------------------------------------------------------- phase query
    // Emits out(p) for p in the band through `emit(i, value)`, i = position inside the band, in increasing i.
    template <typename Emit>
    static SDF_HD void query(const BandTile& T, int c, int b, int nb, int n, int nlevels, pcw_t Pw, pcw_t PCw, Emit emit) {
        const int base = b << LOGB;
        if (base >= n) return;
        const int fin = nlevels & 1;                        // buffer holding the finished scans
        // run context at the band start: a = last run start <= p (or -1), e = first run start > p (or >= n)
        int a = (b > 0) ? (int)T.pcA[fin][midx(b - 1, c)] : -1;
        const int e_after = (b + 1 < nb) ? (int)T.pcE[fin][midx(b + 1, c)] : BH_BIG;
        const uint32_t mym = T.meta[midx(b, c)];
        const int myhi = m_hi(mym);
        int ks = m_lo(mym);                                 // scan pointer over this band's own vertices
        bool hasCur = false, hasNxt = false;
        Ref cur, nxt;
        cur.b = cur.k = cur.pos = 0; cur.h = 0; nxt = cur;
        int e;
        {
            const pcw_t rest = PCw & ~(pcw_t)1;
            e = rest ? base + (LOGB > 5 ? bh_lobit64((uint64_t)rest) : bh_lobit((uint32_t)rest)) : e_after;
        }
        if (!(PCw & 1)) {                                   // position `base` continues a run from the left
            Ref Lr, Rr;
            const bool hasL = (b > 0) && last_upto(T, c, b - 1, 0, Lr);
            const bool hasR = first_from(T, c, b, nb, Rr);
            if (hasR && (!hasL || (base - Rr.pos) * (base - Rr.pos) + Rr.h <= (base - Lr.pos) * (base - Lr.pos) + Lr.h)) {
                cur = Rr; hasCur = true;
            } else if (hasL) {
                cur = Lr; hasCur = true;
                for (;;) {                                  // local search towards the minimiser of `base`
                    Ref pr;
                    if (!prev_entry(T, c, 0, cur, pr)) break;
                    const int dc = base - cur.pos, dp = base - pr.pos;
                    if (dp * dp + pr.h < dc * dc + cur.h) cur = pr; else break;
                }
            }
            if (hasCur) hasNxt = next_entry(T, c, nb, cur, nxt);
        }
#pragma unroll 4
        for (int i = 0; i < B; ++i) {
            const int p = base + i;
            if (p >= n) break;
            if ((PCw >> i) & 1) {                           // a new run starts at p
                a = p;
                const pcw_t rest = (i + 1 < B) ? (PCw >> (i + 1)) : (pcw_t)0;
                e = rest ? p + 1 + (LOGB > 5 ? bh_lobit64((uint64_t)rest) : bh_lobit((uint32_t)rest)) : e_after;
                hasCur = false; hasNxt = false;
                while (ks < myhi && e_rel(T.S[sidx(b, ks, c)]) < i) ++ks;      // vertices of earlier runs
                if (ks < myhi) {
                    const Ref r = load_ref(T, c, b, ks);
                    if (r.pos < e) { cur = r; hasCur = true; }
                } else if (!rest && e_after > ((b + 1) << LOGB)) {
                    // the run continues into later bands and has no vertex here: its first vertex there
                    hasCur = (b + 1 < nb) && first_from(T, c, b + 1, nb, cur);
                }
                if (hasCur) hasNxt = next_entry(T, c, nb, cur, nxt);
            }
            int32_t best = SDF_INF;
            if (hasCur) {
                for (;;) {
                    const int dc = p - cur.pos;
                    const int32_t vc = dc * dc + cur.h;
                    if (hasNxt) {
                        const int dn = p - nxt.pos;
                        if (dn * dn + nxt.h <= vc) {
                            cur = nxt;
                            hasNxt = next_entry(T, c, nb, cur, nxt);
                            continue;
                        }
                    }
                    best = vc;
                    break;
                }
            }
            if (a >= 1) { const int d = p - a + 1; const int32_t z = d * d; best = z < best ? z : best; }
            if (e < n)  { const int d = e - p;     const int32_t z = d * d; best = z < best ? z : best; }
            emit(i, ((Pw >> i) & 1) ? -best : best);
        }
    }
};

}  // namespace sdfb200
