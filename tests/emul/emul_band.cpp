// tests/emul/emul_band.cpp -- CPU emulation of the band-per-warp / line-per-lane phases in
// sdf_tools_b200/csrc/band_hull.cuh.  Test infrastructure: runs the SAME phase functions the CUDA kernel calls,
// as plain loops over the (line, band) threads of a tile (optionally in shuffled order, to expose races inside a phase).
#define BH_COUNT 1
#include "../../sdf_tools_b200/csrc/band_hull.cuh"
#include <vector>
#include <algorithm>
#include <cstring>

using namespace sdfb200;

static uint64_t rng_state = 1;
static long long merge_critical_path = 0;   // sum over tiles and rounds of the longest per-thread dependent chain
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(rng_state >> 33); }

// one strided pass in place: g[r*line_stride + c], r < n, c < ncols
template <int LOGB, int LPC>
static void pass_strided(int32_t* g, int n, int64_t line_stride, int64_t ncols, int shuffle)
{
    using BH = BandHull<LOGB, LPC>;
    constexpr int B = BH::B;
    const int nb = (n + B - 1) / B;
    int nlevels = 0; while ((1 << nlevels) < nb) ++nlevels;
    std::vector<uint32_t> S((size_t)nb * B * LPC), meta((size_t)nb * LPC);
    std::vector<int16_t> a01((size_t)nb * LPC), e01((size_t)nb * LPC);
    std::vector<uint32_t> anyw((size_t)((nb + 31) / 32) * LPC);
    BandTile T; T.S = (bh_addr_t)S.data(); T.meta = (bh_addr_t)meta.data(); T.pcA = (bh_addr_t)a01.data(); T.pcE = (bh_addr_t)e01.data(); T.anyw = (bh_addr_t)anyw.data();
    std::vector<typename BH::pcw_t> Pw((size_t)nb * LPC), PCw((size_t)nb * LPC);
    std::vector<int> order((size_t)nb * LPC);
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    auto shuf = [&]() { if (shuffle) for (size_t i = order.size() - 1; i > 0; i--) std::swap(order[i], order[rnd() % (i + 1)]); };
    std::vector<int32_t> outbuf((size_t)n * LPC);
    for (int64_t c0 = 0; c0 < ncols; c0 += LPC) {
        std::fill(anyw.begin(), anyw.end(), 0u);
        shuf();
        for (int t : order) {
            const int b = t / LPC, c = t % LPC;
            bool prevpol = false;
            for (int i = 0; i < B; i++) {
                const int q = b * B + i;
                S[(size_t)BH::sidx(b, i, c)] = (c0 + c < ncols && q < n) ? (uint32_t)g[(int64_t)q * line_stride + c0 + c] : (uint32_t)SDF_INF;
            }
            if (c0 + c < ncols) {
                const int qp = b > 0 ? b * B - 1 : 0;
                prevpol = (qp < n) ? g[(int64_t)qp * line_stride + c0 + c] < 0 : false;
            }
            BH::build(T, c, b, (c0 + c < ncols) ? n : 0, prevpol, Pw[t], PCw[t]);
        }
        for (int lv = 0; lv < nlevels; lv++) {
            shuf();
            long long worst = 0;
            for (int t : order) {
                const long long s0 = bh_serial;
                BH::merge(T, t % LPC, t / LPC, nb, lv);
                worst = std::max(worst, bh_serial - s0);
            }
            merge_critical_path += worst;              // the round ends when its slowest thread does
        }
        shuf();
        for (int t : order) {
            const int b = t / LPC, c = t % LPC;
            if (c0 + c >= ncols) continue;
            BH::query(T, c, b, nb, n, nlevels, Pw[t], PCw[t], [&](int i, int32_t v) { outbuf[(size_t)(b * B + i) * LPC + c] = v; });
        }
        for (int c = 0; c < LPC && c0 + c < ncols; c++)
            for (int r = 0; r < n; r++) g[(int64_t)r * line_stride + c0 + c] = outbuf[(size_t)r * LPC + c];
    }
}

static int dispatch(int32_t* g, int n, int64_t ls, int64_t ncols, int shuffle, int logb, int lpc)
{
#define CASE(LB, LP) if (logb == LB && lpc == LP) { pass_strided<LB, LP>(g, n, ls, ncols, shuffle); return 0; }
    CASE(3, 4) CASE(4, 4) CASE(5, 4) CASE(6, 4) CASE(5, 16) CASE(5, 32) CASE(5, 8) CASE(6, 8) CASE(4, 32) CASE(3, 32)
#undef CASE
    return -1;
}

extern "C" {

long long emul_band_merge_critical_path(int reset) { const long long v = merge_critical_path; if (reset) merge_critical_path = 0; return v; }

void emul_band_counters(long long* out, int reset) { for (int i = 0; i < 8; i++) { out[i] = bh_counters[i]; if (reset) bh_counters[i] = 0; } }

int emul_band_pass(int32_t* g, int n, int64_t line_stride, int64_t ncols, int logb, int lpc, int shuffle, uint64_t seed)
{
    rng_state = seed * 2 + 1;
    if (((n + (1 << logb) - 1) >> logb) > 64) return -2;
    return dispatch(g, n, line_stride, ncols, shuffle, logb, lpc);
}

// full signed-fused transform in the product's pass order: binary along z, hull passes along y then x
int emul_band_signed_d2(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int32_t* s, int logb, int lpc,
                        int shuffle, uint64_t seed)
{
    rng_state = seed * 2 + 1;
    for (int64_t l = 0; l < nx * ny; l++) {
        const uint8_t* m = mask + l * nz; int32_t* o = s + l * nz;
        int64_t a = -1;
        for (int64_t z = 0; z < nz; z++) {
            const bool pol = m[z] != 0;
            if (z > 0 && pol != (m[z - 1] != 0)) a = z;
            int64_t e = nz;
            for (int64_t k = z + 1; k < nz; k++) if ((m[k] != 0) != pol) { e = k; break; }
            int64_t best = SDF_INF;
            if (a >= 1) best = (z - a + 1) * (z - a + 1);
            if (e < nz && (e - z) * (e - z) < best) best = (e - z) * (e - z);
            o[z] = pol ? -(int32_t)best : (int32_t)best;
        }
    }
    for (int64_t x = 0; x < nx; x++)
        if (dispatch(s + x * ny * nz, (int)ny, nz, nz, shuffle, logb, lpc)) return -1;
    if (dispatch(s, (int)nx, ny * nz, ny * nz, shuffle, logb, lpc)) return -1;
    return 0;
}

}  // extern "C"
