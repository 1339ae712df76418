/* tests/emul/sqrt_seq.c -- CPU restatement of sdf_dist_f32's FP64 sequence (test infrastructure). */
#include <math.h>
static double seq(int m, float seed)
{
    const double x = (double)m, r = (double)seed;
    double g = x * r, h = 0.5 * r;
    const double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}
long sqrt_seq_count_bad(int m0, int m1)
{
    long bad = 0;
    for (int m = m0; m < m1; m++) {
        const float s = 1.0f / sqrtf((float)m);
        for (int k = -2; k <= 2; k++)
            if (seq(m, s * (1.0f + k * 2.4e-7f)) != sqrt((double)m)) bad++;
    }
    return bad;
}
