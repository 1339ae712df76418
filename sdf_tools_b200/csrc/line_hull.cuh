// sdf_tools_b200/csrc/line_hull.cuh
//
// One 1-D pass of the signed-fused exact Euclidean distance transform, written as per-lane phases of a
// warp that owns ONE grid line.  It replaces the inner work of
//   sdf_generation::BuildDistanceField            /root/reference/include/sdf_tools/sdf_generation.hpp:95-207
// (called twice, :242-243) for one axis: the reference floods d^2 outwards through a bucket queue; here
// every line of the axis gets   out(p) = min_q (p-q)^2 + f(q)   exactly, for both polarities at once.
//
// Signed-fused field: s(q) > 0  <=> voxel q is FREE   and s = d^2 to the nearest FILLED
//                                            s(q) < 0  <=> voxel q is FILLED and -s = d^2 to the nearest FREE
// (partial d^2 over the axes already transformed; |s| == SDF_INF: nothing of the other polarity seen yet).
// Exactly one of the reference's two distances (:254-255) is non-zero at any voxel, so one int32 carries both.
//
// Along a line the two transforms decouple into maximal RUNS of equal polarity:
//   for p in run [a,b]:  out(p) = min( (p-(a-1))^2 , ((b+1)-p)^2 , min_{q in [a,b]} (p-q)^2 + |s(q)| )
// -- the voxels a-1 / b+1 have the other polarity (distance 0 in that field) and shadow everything behind them.
// The last term is the lower envelope of parabolas == lower convex hull of the lifted points (q, q^2+|s(q)|).
//
// Work split inside the warp (line length n <= 32*CHUNK): lane l owns positions [l*CHUNK, (l+1)*CHUNK).
//   phase 1  lane_build : sequential hull of the lane's own chunk; the stack is a BITMASK over the chunk
//                         (top / second-from-top cached in registers), restarted at every polarity change.
//   phase 2  lane_merge : log2(32) rounds; round k joins the hulls of 2^k-chunk groups pairwise by walking the
//                         common tangent across the seam (only if a run continues across it).
//   phase 3  lane_query : each lane walks the finished hull for its own positions (monotone minimiser), adds
//                         the two run-end terms, and keeps the results in registers.
// The phases only communicate through the line's shared arrays (val / M / PC) and are race-free inside a
// phase, so the same code runs (a) on the GPU with __syncwarp() between phases and (b) on the CPU with a plain
// loop over lanes -- tests/ use (b) to check the algorithm without a GPU.
//
// Exactness: everything is integer.  |s| <= 3*2048^2 < 2^24, lifted heights < 2^25, hull tests are
// cross-multiplied in int64.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SDF_HD __host__ __device__ __forceinline__
#else
#define SDF_HD inline
#endif

#define SDF_INF 0x40000000          // |s| sentinel: no voxel of the other polarity found so far
#define SDF_MAX_AXIS 2048           // 32 lanes x 64-bit chunk masks

namespace sdfb200 {

// ---------------------------------------------------------------------------------------------- bit helpers
template <typename MaskT> struct Bits;
template <> struct Bits<uint32_t> {
    static constexpr int N = 32;
    static SDF_HD int hi(uint32_t m) {   // index of highest set bit, m != 0
#if defined(__CUDA_ARCH__)
        return 31 - __clz((int)m);
#else
        return 31 - __builtin_clz(m);
#endif
    }
    static SDF_HD int lo(uint32_t m) {   // index of lowest set bit, m != 0
#if defined(__CUDA_ARCH__)
        return __ffs((int)m) - 1;
#else
        return __builtin_ctz(m);
#endif
    }
};
template <> struct Bits<uint64_t> {
    static constexpr int N = 64;
    static SDF_HD int hi(uint64_t m) {
#if defined(__CUDA_ARCH__)
        return 63 - __clzll((long long)m);
#else
        return 63 - __builtin_clzll(m);
#endif
    }
    static SDF_HD int lo(uint64_t m) {
#if defined(__CUDA_ARCH__)
        return __ffsll((long long)m) - 1;
#else
        return __builtin_ctzll(m);
#endif
    }
};

template <typename MaskT> SDF_HD MaskT bit_at(int k) { return (MaskT)1 << k; }
// bits [0,k)
template <typename MaskT> SDF_HD MaskT low_mask(int k) {
    return (k >= Bits<MaskT>::N) ? ~(MaskT)0 : (((MaskT)1 << k) - 1);
}

SDF_HD int32_t iabs32(int32_t v) { return v < 0 ? -v : v; }

// ------------------------------------------------------------------------------------------------ the line
// CHUNK = 2^LOGC positions per lane.  The line lives in shared memory in a "rotated chunk-transposed" order:
// position r = l*CHUNK + j  ->  word  j*32 + ((l + j) & 31).  Lane l's own chunk then sits in 32 different
// banks for any fixed j (conflict-free in the unrolled loops), and both tile loaders (row-of-columns for the
// strided axes, float4 along the line for the contiguous axis) store conflict-free.
template <int LOGC>
struct Line {
    static constexpr int CHUNK = 1 << LOGC;
    static constexpr int WORDS = 32 * CHUNK;
    static SDF_HD int idx(int r) {
        const int j = r & (CHUNK - 1), l = r >> LOGC;
        return j * 32 + ((l + j) & 31);
    }
};

// t lies on or above the chord s->u  (lifted points) <=> parabola t never wins strictly between s and u
SDF_HD bool hull_dominated(int sp, int32_t sY, int tp, int32_t tY, int up, int32_t uY) {
    const long long lhs = (long long)(tY - sY) * (long long)(up - tp);
    const long long rhs = (long long)(uY - tY) * (long long)(tp - sp);
    return lhs >= rhs;
}

template <int LOGC, typename MaskT>
struct LineHull {
    using L = Line<LOGC>;
    using B = Bits<MaskT>;
    static constexpr int CHUNK = L::CHUNK;

    static SDF_HD int32_t height(const int32_t* val, int pos) { return iabs32(val[L::idx(pos)]); }
    static SDF_HD int32_t lifted(const int32_t* val, int pos) { return pos * pos + height(val, pos); }

    // Highest hull site e with lo <= e < t inside the run that contains t (run start = highest PC bit <= t).
    static SDF_HD int prev_site(const MaskT* Mw, const MaskT* PCw, int t, int lo) {
        int w = t >> LOGC;
        const int j = t & (CHUNK - 1);
        MaskT pcm = PCw[w] & low_mask<MaskT>(j + 1);
        MaskT mm = Mw[w] & low_mask<MaskT>(j);
        for (;;) {
            if (pcm) {
                const int a = B::hi(pcm);
                mm &= ~low_mask<MaskT>(a);
                if (!mm) return -1;
                const int pos = (w << LOGC) + B::hi(mm);
                return pos >= lo ? pos : -1;
            }
            if (mm) {
                const int pos = (w << LOGC) + B::hi(mm);
                return pos >= lo ? pos : -1;
            }
            if ((w << LOGC) <= lo) return -1;
            --w;
            pcm = PCw[w];
            mm = Mw[w];
        }
    }

    // Lowest hull site e with start <= e < hi and no run boundary (PC bit) in [pcstart, e].
    static SDF_HD int next_site(const MaskT* Mw, const MaskT* PCw, int start, int pcstart, int hi) {
        if (start >= hi) return -1;
        int w = start >> LOGC;
        MaskT mm = Mw[w] & ~low_mask<MaskT>(start - (w << LOGC));
        MaskT pcm = PCw[w] & ~low_mask<MaskT>(pcstart - (w << LOGC));
        for (;;) {
            if (pcm) {
                mm &= low_mask<MaskT>(B::lo(pcm));
                if (!mm) return -1;
                const int pos = (w << LOGC) + B::lo(mm);
                return pos < hi ? pos : -1;
            }
            if (mm) {
                const int pos = (w << LOGC) + B::lo(mm);
                return pos < hi ? pos : -1;
            }
            ++w;
            if ((w << LOGC) >= hi) return -1;
            mm = Mw[w];
            pcm = PCw[w];
        }
    }

    // Highest run start (PC bit) at a position <= t, or -1.
    static SDF_HD int prev_pc(const MaskT* PCw, int t) {
        int w = t >> LOGC;
        MaskT pcm = PCw[w] & low_mask<MaskT>((t & (CHUNK - 1)) + 1);
        for (;;) {
            if (pcm) return (w << LOGC) + B::hi(pcm);
            if (w == 0) return -1;
            --w;
            pcm = PCw[w];
        }
    }
    // Lowest run start at a position >= s (s may equal 32*CHUNK), or n when there is none.
    static SDF_HD int next_pc(const MaskT* PCw, int s, int n) {
        if (s >= n) return n;
        int w = s >> LOGC;
        MaskT pcm = PCw[w] & ~low_mask<MaskT>(s - (w << LOGC));
        for (;;) {
            if (pcm) return (w << LOGC) + B::lo(pcm);
            ++w;
            if ((w << LOGC) >= n) return n;
            pcm = PCw[w];
        }
    }

    // ---------------------------------------------------------------------------------------------- phase 1
    // Returns the lane's polarity word (bit j = voxel l*CHUNK+j is FILLED).  Writes Mw[lane], PCw[lane].
    static SDF_HD MaskT lane_build(const int32_t* val, MaskT* Mw, MaskT* PCw, int n, int lane) {
        const int base = lane << LOGC;
        MaskT m_done = 0, m_run = 0, pcw = 0, pw = 0;
        if (base < n) {
            int cnt = 0;                       // hull sites of the current run inside this chunk
            int tp = 0, sp = 0;
            int32_t tY = 0, sY = 0;
            bool prevpol = val[L::idx(lane > 0 ? base - 1 : 0)] < 0;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const int q = base + j;
                if (q < n) {
                    const int32_t v = val[L::idx(q)];
                    const bool pol = v < 0;
                    const int32_t h = iabs32(v);
                    if (pol) pw |= bit_at<MaskT>(j);
                    if (pol != prevpol) {      // a new run starts at q
                        pcw |= bit_at<MaskT>(j);
                        m_done |= m_run;
                        m_run = 0;
                        cnt = 0;
                        prevpol = pol;
                    }
                    if (h < SDF_INF) {
                        const int32_t Y = q * q + h;
                        while (cnt >= 2 && hull_dominated(sp, sY, tp, tY, q, Y)) {
                            m_run &= ~bit_at<MaskT>(tp - base);
                            tp = sp;
                            tY = sY;
                            --cnt;
                            if (cnt >= 2) {
                                const MaskT below = m_run & low_mask<MaskT>(tp - base);
                                sp = base + B::hi(below);
                                sY = lifted(val, sp);
                            }
                        }
                        sp = tp;
                        sY = tY;
                        tp = q;
                        tY = Y;
                        ++cnt;
                        m_run |= bit_at<MaskT>(j);
                    }
                }
            }
        }
        Mw[lane] = m_done | m_run;
        PCw[lane] = pcw;
        return pw;
    }

    // ---------------------------------------------------------------------------------------------- phase 2
    // Round lv (0..4): lane i < (16 >> lv) joins chunk groups [cm-2^lv, cm) and [cm, cm+2^lv), cm = (2i+1)<<lv.
    static SDF_HD void lane_merge(const int32_t* val, MaskT* Mw, const MaskT* PCw, int n, int lane, int lv) {
        if (lane >= (16 >> lv)) return;
        const int cm = (2 * lane + 1) << lv;
        const int seam = cm << LOGC;
        if (seam >= n) return;
        if (PCw[cm] & 1) return;                       // a run starts exactly at the seam: nothing crosses it
        const int lo = (cm - (1 << lv)) << LOGC;
        int hi = (cm + (1 << lv)) << LOGC;
        if (hi > n) hi = n;
        int Lp = prev_site(Mw, PCw, seam, lo);
        if (Lp < 0) return;
        int Rp = next_site(Mw, PCw, seam, seam + 1, hi);
        if (Rp < 0) return;
        int32_t LY = lifted(val, Lp), RY = lifted(val, Rp);
        int L2 = prev_site(Mw, PCw, Lp, lo);
        int R2 = next_site(Mw, PCw, Rp + 1, Rp + 1, hi);
        int32_t L2Y = L2 >= 0 ? lifted(val, L2) : 0;
        int32_t R2Y = R2 >= 0 ? lifted(val, R2) : 0;
        for (;;) {
            bool changed = false;
            while (L2 >= 0 && hull_dominated(L2, L2Y, Lp, LY, Rp, RY)) {
                Mw[Lp >> LOGC] &= ~bit_at<MaskT>(Lp & (CHUNK - 1));
                Lp = L2;
                LY = L2Y;
                L2 = prev_site(Mw, PCw, Lp, lo);
                if (L2 >= 0) L2Y = lifted(val, L2);
                changed = true;
            }
            while (R2 >= 0 && hull_dominated(Lp, LY, Rp, RY, R2, R2Y)) {
                Mw[Rp >> LOGC] &= ~bit_at<MaskT>(Rp & (CHUNK - 1));
                Rp = R2;
                RY = R2Y;
                R2 = next_site(Mw, PCw, Rp + 1, Rp + 1, hi);
                if (R2 >= 0) R2Y = lifted(val, R2);
                changed = true;
            }
            if (!changed) break;
        }
    }

    // ---------------------------------------------------------------------------------------------- phase 3
    // out[j] = signed result for position lane*CHUNK + j  (undefined for positions >= n).
    static SDF_HD void lane_query(const int32_t* val, const MaskT* Mw, const MaskT* PCw, int n, int lane,
                                  MaskT pw, int32_t* out) {
        const int base = lane << LOGC;
        if (base >= n) return;
        // run context of the first position
        int a = prev_pc(PCw, base);                    // run start (>=1) or -1: opening zero site at a-1
        int e = next_pc(PCw, base + 1, n);             // next run start (closing zero site) or n
        // nearest hull site of this run at/before base, else the first one after it
        int cur;
        if ((Mw[lane] & 1) && true) cur = base; else cur = prev_site(Mw, PCw, base, 0);
        if (cur < 0) cur = next_site(Mw, PCw, base + 1, base + 1, n);
        int32_t ch = 0;
        int nxt = -1;
        int32_t nh = 0;
        if (cur >= 0) {
            ch = height(val, cur);
            // local search towards the minimiser of position `base` (values along the hull are unimodal)
            for (;;) {
                const int pr = prev_site(Mw, PCw, cur, 0);
                if (pr < 0) break;
                const int32_t ph = height(val, pr);
                const int dc = base - cur, dp = base - pr;
                if (dp * dp + ph < dc * dc + ch) { cur = pr; ch = ph; } else break;
            }
            nxt = next_site(Mw, PCw, cur + 1, cur + 1, n);
            if (nxt >= 0) nh = height(val, nxt);
        }
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
            const int q = base + j;
            if (q < n) {
                if (j > 0 && ((PCw[lane] >> j) & 1)) {     // a new run starts at q
                    a = q;
                    e = next_pc(PCw, q + 1, n);
                    cur = next_site(Mw, PCw, q, q + 1, n);
                    nxt = -1;
                    if (cur >= 0) {
                        ch = height(val, cur);
                        nxt = next_site(Mw, PCw, cur + 1, cur + 1, n);
                        if (nxt >= 0) nh = height(val, nxt);
                    }
                }
                int32_t best = SDF_INF;
                if (cur >= 0) {
                    for (;;) {
                        const int dc = q - cur;
                        const int32_t vc = dc * dc + ch;
                        if (nxt >= 0) {
                            const int dn = q - nxt;
                            const int32_t vn = dn * dn + nh;
                            if (vn <= vc) {
                                cur = nxt;
                                ch = nh;
                                nxt = next_site(Mw, PCw, cur + 1, cur + 1, n);
                                if (nxt >= 0) nh = height(val, nxt);
                                continue;
                            }
                        }
                        best = vc;
                        break;
                    }
                }
                if (a >= 1) { const int d = q - a + 1; const int32_t z = d * d; if (z < best) best = z; }
                if (e < n)  { const int d = e - q;     const int32_t z = d * d; if (z < best) best = z; }
                out[j] = ((pw >> j) & 1) ? -best : best;
            }
        }
    }
};

}  // namespace sdfb200
