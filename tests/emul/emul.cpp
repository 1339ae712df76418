// tests/emul/emul.cpp -- CPU emulation of the warp-per-line phases in sdf_tools_b200/csrc/line_hull.cuh.
// Test infrastructure: runs the SAME phase functions the CUDA kernels call, with a plain loop over the 32
// lanes (optionally in a shuffled order, to expose any cross-lane race inside a phase).  Lets the algorithm
// be checked against the oracle on machines without a GPU.
#include "../../sdf_tools_b200/csrc/line_hull.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>
#include <algorithm>

using namespace sdfb200;

static uint64_t rng_state = 1;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(rng_state >> 33); }

template <int LOGC, typename MaskT>
static void run_line(int32_t* line, int64_t stride, int n, int shuffle)
{
    using LH = LineHull<LOGC, MaskT>;
    using L = Line<LOGC>;
    constexpr int CHUNK = L::CHUNK;
    std::vector<int32_t> val(L::WORDS, 0);
    MaskT Mw[32], PCw[32], pw[32];
    for (int r = 0; r < n; r++) val[L::idx(r)] = line[r * stride];
    int order[32];
    for (int i = 0; i < 32; i++) order[i] = i;
    auto shuf = [&]() { if (shuffle) for (int i = 31; i > 0; i--) std::swap(order[i], order[rnd() % (i + 1)]); };
    shuf();
    for (int i = 0; i < 32; i++) pw[order[i]] = LH::lane_build(val.data(), Mw, PCw, n, order[i]);
    for (int lv = 0; lv < 5; lv++) {
        shuf();
        for (int i = 0; i < 32; i++) LH::lane_merge(val.data(), Mw, PCw, n, order[i], lv);
    }
    std::vector<int32_t> out(32 * CHUNK, 0);
    shuf();
    for (int i = 0; i < 32; i++) LH::lane_query(val.data(), Mw, PCw, n, order[i], pw[order[i]], &out[order[i] * CHUNK]);
    for (int r = 0; r < n; r++) line[r * stride] = out[r];
}

static int logc_for(int n) { int lc = 0; while ((32 << lc) < n) lc++; return lc; }

static int dispatch_line(int32_t* line, int64_t stride, int n, int shuffle, int logc)
{
    if (logc < 0) logc = logc_for(n);
    if ((32 << logc) < n) return -1;
    switch (logc) {
        case 0: run_line<0, uint32_t>(line, stride, n, shuffle); break;
        case 1: run_line<1, uint32_t>(line, stride, n, shuffle); break;
        case 2: run_line<2, uint32_t>(line, stride, n, shuffle); break;
        case 3: run_line<3, uint32_t>(line, stride, n, shuffle); break;
        case 4: run_line<4, uint32_t>(line, stride, n, shuffle); break;
        case 5: run_line<5, uint32_t>(line, stride, n, shuffle); break;
        case 6: run_line<6, uint64_t>(line, stride, n, shuffle); break;
        default: return -1;
    }
    return 0;
}

extern "C" {

// one signed line pass in place (logc < 0: smallest chunk that fits)
int emul_line_pass(int32_t* line, int n, int logc, int shuffle, uint64_t seed)
{
    rng_state = seed * 2 + 1;
    return dispatch_line(line, 1, n, shuffle, logc);
}

// full signed-fused 3-pass transform: binary scan along x, hull passes along y then z
int emul_signed_d2(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int32_t* s, int shuffle, uint64_t seed)
{
    rng_state = seed * 2 + 1;
    if (nx > SDF_MAX_AXIS || ny > SDF_MAX_AXIS || nz > SDF_MAX_AXIS) return -1;
    const int64_t cs = ny * nz;
    // pass 1 (binary, along x): distance to the nearest voxel of the other polarity in the x line
    for (int64_t c = 0; c < cs; c++) {
        int64_t last_change = -1;   // start of current run
        for (int64_t x = 0; x < nx; x++) {
            const bool pol = mask[x * cs + c] != 0;
            if (x > 0 && pol != (mask[(x - 1) * cs + c] != 0)) last_change = x;
            int64_t e = nx;
            for (int64_t k = x + 1; k < nx; k++) if ((mask[k * cs + c] != 0) != pol) { e = k; break; }
            int64_t best = SDF_INF;
            if (last_change >= 1) { int64_t d = x - last_change + 1; best = d * d; }
            if (e < nx) { int64_t d = e - x; if (d * d < best) best = d * d; }
            s[x * cs + c] = pol ? -(int32_t)best : (int32_t)best;
        }
    }
    // pass 2 along y
    for (int64_t x = 0; x < nx; x++) for (int64_t z = 0; z < nz; z++)
        if (dispatch_line(s + x * cs + z, nz, (int)ny, shuffle, -1)) return -2;
    // pass 3 along z
    for (int64_t l = 0; l < nx * ny; l++)
        if (dispatch_line(s + l * nz, 1, (int)nz, shuffle, -1)) return -3;
    return 0;
}

}  // extern "C"
