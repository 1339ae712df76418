// tests/emul/emul_line.cpp -- CPU run of the per-line function in sdf_tools_b200/csrc/line_hull.cuh (the SAME source the
// CUDA kernel compiles), one call per line.  Test infrastructure.
#define LH_COUNT 1
#include "../../sdf_tools_b200/csrc/line_hull.cuh"
#include <vector>

using namespace sdfb200;

template <int FMT>
static void pass_lines(const int32_t* in, int32_t* out, int n, int64_t line_stride, int64_t ncols)
{
    typedef typename LineEntry<FMT>::type ent_t;
    const int nw = (n + 31) / 32;
    const int64_t tiles = (ncols + 31) / 32;
    std::vector<ent_t> stk((size_t)tiles * lh_block_entries(n), (ent_t)0x12345678);   // stale garbage: the pass must not depend on it
    std::vector<uint32_t> pol((size_t)nw * ncols, 0xdeadbeefu);
    const int64_t tile_elems = lh_block_entries(n);
    for (int64_t c = 0; c < ncols; c++) {
        int32_t* dst = out + c;
        const int l = (int)(c & 31);
        line_hull_pass<FMT>(in + c, line_stride, n, stk.data() + (c >> 5) * tile_elems + lh_lane_offset(l), pol.data() + c, ncols,
                             [&](int32_t m, bool neg) { *dst = neg ? -m : m; dst += line_stride; });
    }
}

extern "C" {

void emul_line_counters(long long* o, int reset) { for (int i = 0; i < 8; i++) { o[i] = lh_counters[i]; if (reset) lh_counters[i] = 0; } }

// stk here is laid out [row][ncols] while the data is [row][line_stride]: use ncols == line_stride callers only
int emul_line_pass(const int32_t* in, int32_t* out, int n, int64_t line_stride, int64_t ncols, int fmt)
{
    if (ncols != line_stride) return -1;
    if (fmt == LH_FMT_WIDE) pass_lines<LH_FMT_WIDE>(in, out, n, line_stride, ncols);
    else if (fmt == LH_FMT_P11) pass_lines<LH_FMT_P11>(in, out, n, line_stride, ncols);
    else pass_lines<LH_FMT_P10>(in, out, n, line_stride, ncols);
    return 0;
}

int emul_line_signed_d2(const uint8_t* mask, int64_t nx, int64_t ny, int64_t nz, int32_t* s, int fmt)
{
    std::vector<int32_t> t((size_t)(nx * ny * nz));
    for (int64_t l = 0; l < nx * ny; l++) {
        const uint8_t* m = mask + l * nz; int32_t* o = t.data() + l * nz;
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
        emul_line_pass(t.data() + x * ny * nz, s + x * ny * nz, (int)ny, nz, nz, fmt);
    emul_line_pass(s, t.data(), (int)nx, ny * nz, ny * nz, fmt);
    for (size_t i = 0; i < t.size(); i++) s[i] = t[i];
    return 0;
}

}  // extern "C"

extern "C" int emul_line_entry_format(int64_t n, int64_t hmax) { return lh_entry_format(n, hmax); }
