/* Exhaustive proof that the device's FMA form of glibc's logf kernel (pvc_analyze.cu::decibelsNormal) returns the host libm's
 * logf bit for bit on EVERY float in [0.5, 2) -- the only inputs fdlibm's log10f hands to logf (e_log10f.c reduces x to
 * m * 2^k with m in [0.5, 2)).  Three forms are compared on all 2^24 inputs:
 *   (a) glibc 2.39 e_logf.c as written, separate multiply / add roundings (what an x86-64 build without FMA executes)
 *   (b) the same with every a*b+c contracted to fma (what glibc's __logf_fma ifunc variant executes on FMA-capable CPUs)
 *   (c) the device form of round 1: 33-entry table indexed by (ix - OFF) >> 19 holding { invc * 2^-k, logc + k*ln2 }, so that
 *       r = fma(m, invc', -1) needs neither the reduced mantissa z nor the k*ln2 term; e_logf.c's polynomial in FMA form (6 DP ops)
 *   (d) the device form now: the same table and r, and y = logc' + r + r^2 (A2 + A1 r + A0 r^2) as one Horner chain in r,
 *       fma(fma(fma(fma(A0, r, A1), r, A2), r, 1), r, logc') (5 DP ops)
 * Build + run:  gcc -O2 -ffp-contract=off -o logf_fma_recipe logf_fma_recipe.c -lm && ./logf_fma_recipe
 * (tests/test_oracle.py::test_device_logf_fma_form_is_exhaustively_exact runs it.)                                            */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
typedef struct { double invc, logc; } E;
static const E tab[16] = {
    { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
    { 0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2 }, { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
    { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3 },
    { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
    { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1.0000000000000p+0, 0x0.0p+0 },
    { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 }, { 0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4 },
    { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 }, { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3 },
    { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 }, { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 } };
static const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2, Ln2 = 0x1.62e42fefa39efp-1;
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static float logf_glibc(float x, int usefma)
{
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0.f;
    uint32_t tmp = ix - 0x3f330000u;
    int i = (tmp >> 19) & 15, k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double z = (double)u2f(iz), invc = tab[i].invc, logc = tab[i].logc, r, y0, r2, y;
    if (usefma) { r = fma(z, invc, -1.0); y0 = fma((double)k, Ln2, logc); r2 = r * r; y = fma(A1, r, A2); y = fma(A0, r2, y); y = fma(y, r2, y0 + r); }
    else
    {
        volatile double t;
        t = z * invc; r = t - 1.0; t = (double)k * Ln2; y0 = logc + t; r2 = r * r; t = A1 * r; y = t + A2; t = A0 * r2; y = t + y;
        t = y * r2; y = t + (y0 + r);
    }
    return (float)y;
}

static E tab33[33];
static void build33(void)
{
    for (int idx = 0; idx < 33; ++idx)
    {
        int j = idx - 7, ti = j & 15, tk = j >> 4;                  /* arithmetic shift: -1, 0, 1 */
        volatile double kl = (double)tk * Ln2;                      /* exact: tk is -1, 0 or 1 */
        tab33[idx].invc = ldexp(tab[ti].invc, -tk);                 /* exact power-of-two scaling */
        tab33[idx].logc = tab[ti].logc + kl;                        /* the reference's y0, one rounding */
    }
}
static float logf_device(float m)
{
    uint32_t ix = f2u(m);
    int idx = ((int32_t)(ix - 0x3f330000u) >> 19) + 7;
    double md = (double)m, r = fma(md, tab33[idx].invc, -1.0), r2 = r * r;
    double y = fma(A1, r, A2);
    y = fma(A0, r2, y);
    y = fma(y, r2, tab33[idx].logc + r);
    return (float)y;
}

static float logf_device_horner(float m)
{
    uint32_t ix = f2u(m);
    int idx = ((int32_t)(ix - 0x3f330000u) >> 19) + 7;
    double md = (double)m, r = fma(md, tab33[idx].invc, -1.0);
    double q = fma(A0, r, A1);
    q = fma(q, r, A2);
    q = fma(q, r, 1.0);
    return (float)fma(q, r, tab33[idx].logc);
}

int main(void)
{
    long ab = 0, a_l = 0, b_l = 0, c_l = 0, d_l = 0, n = 0;
    build33();
    for (uint32_t u = 0x3f000000u; u < 0x40000000u; ++u, ++n)
    {
        const float x = u2f(u), a = logf_glibc(x, 0), b = logf_glibc(x, 1), c = logf_device(x), d = logf_device_horner(x), ref = logf(x);
        ab += f2u(a) != f2u(b); a_l += f2u(a) != f2u(ref); b_l += f2u(b) != f2u(ref);
        c_l += f2u(c) != f2u(ref) && !(u == 0x3f800000u && c == 0.f && ref == 0.f);
        d_l += f2u(d) != f2u(ref) && !(u == 0x3f800000u && d == 0.f && ref == 0.f);
    }
    printf("inputs %ld  a!=b %ld  a!=libm %ld  b!=libm %ld  device!=libm %ld  horner!=libm %ld\n", n, ab, a_l, b_l, c_l, d_l);
    return (ab | a_l | b_l | c_l | d_l) ? 1 : 0;
}
