// tests/emul/emul_run.cpp -- CPU emulation of the phases in sdf_tools_b200/csrc/run_hull.cuh.
// Test infrastructure: runs the SAME phase functions the CUDA kernel calls, as plain loops over the threads of a
// tile (optionally in shuffled order, to expose races inside a phase).
#define RH_COUNT 1
#include "../../sdf_tools_b200/csrc/run_hull.cuh"
#include <vector>
#include <algorithm>
#include <cstring>

using namespace sdfb200;

static uint64_t rng_state = 1;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(rng_state >> 33); }

// one strided pass in place: g[r*line_stride + c], r < n, c < ncols
template <int LPC, int NW>
static void pass_strided(int32_t* g, int n, int64_t line_stride, int64_t ncols, int shuffle)
{
    using RHT = RunHull<LPC, NW>;
    constexpr int B = 32;
    const int nb = (n + B - 1) / B;
    int nlevels = 0; while ((1 << nlevels) < nb) ++nlevels;
    std::vector<int32_t> S((size_t)(nb * B + 1) * LPC);
    std::vector<uint32_t> AW((size_t)nb * LPC), PC((size_t)nb * LPC), NE((size_t)NW * LPC), PCS((size_t)NW * LPC);
    RunTile T; T.S = (rh_addr_t)S.data(); T.AW = (rh_addr_t)AW.data(); T.PC = (rh_addr_t)PC.data();
    T.NE = (rh_addr_t)NE.data(); T.PCS = (rh_addr_t)PCS.data();
    std::vector<uint32_t> Pw((size_t)nb * LPC), PCw((size_t)nb * LPC);
    std::vector<int> order((size_t)nb * LPC);
    auto shuf = [&](size_t cnt) { order.resize(cnt); for (size_t i = 0; i < cnt; i++) order[i] = (int)i;
        if (shuffle) for (size_t i = cnt - 1; i > 0 && cnt > 1; i--) std::swap(order[i], order[rnd() % (i + 1)]); };
    std::vector<int32_t> outbuf((size_t)n * LPC);
    for (int64_t c0 = 0; c0 < ncols; c0 += LPC) {
        std::fill(NE.begin(), NE.end(), 0u); std::fill(PCS.begin(), PCS.end(), 0u);
        // staging: out-of-grid rows / columns arrive as zeros (what the TMA tile load does) and must never be used
        for (int r = 0; r < nb * B + 1; r++) for (int c = 0; c < LPC; c++)
            S[(size_t)r * LPC + c] = (c0 + c < ncols && r < n) ? g[(int64_t)r * line_stride + c0 + c] : 0;
        shuf((size_t)nb * LPC);
        for (int t : order) {
            const int b = t / LPC, c = t % LPC;
            RHT::build(T, c, b, (c0 + c < ncols) ? n : 0, Pw[t], PCw[t]);
        }
        for (int lv = 0; lv < nlevels; lv++) {
            const int gsz = 1 << lv;
            const int nseams = (nb - gsz + 2 * gsz - 1) / (2 * gsz);
            shuf((size_t)nseams * LPC);
            for (int t : order) {
                const int c = t % LPC, j = t / LPC;
                RHT::merge(T, c, (2 * j + 1) << lv, nb, lv, (c0 + c < ncols) ? n : 0);
            }
        }
        shuf((size_t)nb * LPC);
        for (int t : order) {
            const int b = t / LPC, c = t % LPC;
            if (c0 + c >= ncols) continue;
            RHT::query(T, c, b, nb, n, Pw[t], PCw[t], [&](int i, int32_t v) { outbuf[(size_t)(b * B + i) * LPC + c] = v; });
        }
        for (int c = 0; c < LPC && c0 + c < ncols; c++)
            for (int r = 0; r < n; r++) g[(int64_t)r * line_stride + c0 + c] = outbuf[(size_t)r * LPC + c];
    }
}

static int dispatch(int32_t* g, int n, int64_t ls, int64_t ncols, int shuffle, int lpc)
{
    const int nb = (n + 31) / 32;
#define CASE(LP) if (lpc == LP) { if (nb <= 32) pass_strided<LP, 1>(g, n, ls, ncols, shuffle); else pass_strided<LP, 2>(g, n, ls, ncols, shuffle); return 0; }
    CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
    return -1;
}

extern "C" {

void emul_run_counters(long long* out, int reset) { for (int i = 0; i < 16; i++) { out[i] = rh_counters[i]; if (reset) rh_counters[i] = 0; } }

int emul_run_pass(int32_t* g, int n, int64_t line_stride, int64_t ncols, int lpc, int shuffle, uint64_t seed)
{
    rng_state = seed * 2 + 1;
    if ((n + 31) / 32 > 64) return -2;
    return dispatch(g, n, line_stride, ncols, shuffle, lpc);
}

// full signed-fused transform in the product's pass order: binary along z, hull passes along y then x
int emul_run_signed_d2(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int32_t* s, int lpc, int shuffle, uint64_t seed)
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
        if (dispatch(s + x * ny * nz, (int)ny, nz, nz, shuffle, lpc)) return -1;
    if (dispatch(s, (int)nx, ny * nz, ny * nz, shuffle, lpc)) return -1;
    return 0;
}

}  // extern "C"
