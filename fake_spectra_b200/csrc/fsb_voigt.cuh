// Device Voigt profile H(a,u) = Re w(u + i a), a >= 0  (consumer: singleabs.h:56-61).
//
// voigt_exact : restatement of the reference's Faddeeva::w real part (Faddeeva.cpp:679-971,
//               relerr = DBL_EPSILON branch) — the parity anchor for every other strategy.
// voigt fast  : this library's own evaluation for the small damping parameters of real lines
//               (0 <= y <= kFastYMax): expansion of w about the real axis to order y^7,
//                   H(x,y) = U(x) Pe(s) + G(x) A(s) + B(s),   s = x^2,  U = exp(-s),
//                   G(x) = 1 - 2 x Dawson(x)   (193 piecewise degree-7 polynomials, |x| < 24, stored in
//                   mixed precision: 48 bytes per piece, see fsb_voigt_tables.h),
//               with Pe, A, B polynomials in s whose coefficients depend on y only (per-particle
//               constants).  Beyond |x| >= 12 the Gaussian has vanished and the profile is the
//               Taylor series in y of the damping wing, -y L' + y^3 L'''/6 - y^5 L^(5)/120 with
//               L = Im w on the real axis expanded in u = 1/x^2 (polynomials P1, P3, P5 below).
//               Agrees with voigt_exact to < 1e-11 relative on its domain (tests/test_gpu_parity).
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

#include "fsb_voigt_tables.h"

namespace fsb {

// exp(-a2 n^2), a2 = 0.268657157075235951582, n = 1..51, then 0 (generated with mpmath, 40 digits).
// Same role as the expa2n2 table of Faddeeva.cpp:622-675; the trailing zero also terminates the
// series loops once every term has underflowed.
__constant__ double c_expa2n2[52] = {
    7.64405281671221563e-1, 3.41424527166548425e-1, 8.91072646929412548e-2,
    1.35887299055460086e-2, 1.21085455253437481e-3, 6.30452613933449404e-5,
    1.91805156577114683e-6, 3.40969447714832381e-8, 3.54175089099469393e-10,
    2.14965079583260681e-12, 7.62368911833724355e-15, 1.57982797110681093e-17,
    1.91294189103582676e-20, 1.3534465676420534e-23, 5.59535712428588719e-27,
    1.35164257972401769e-30, 1.90784582843501168e-34, 1.5735192029144293e-38,
    7.58312432328032848e-43, 2.13536275438697082e-47, 3.5135206378719577e-52,
    3.37800830266396921e-57, 1.89769439468301001e-62, 6.22929926072668851e-68,
    1.19481172006938723e-73, 1.33908181133005952e-79, 8.76924303483223948e-86,
    3.35555576166254989e-92, 7.50264110688173025e-99, 9.80192200745410261e-106,
    7.48265412822268965e-113, 3.33770122566809428e-120, 8.69934598159861142e-128,
    1.32486951484088855e-135, 1.17898144201315251e-143, 6.13039120236180011e-152,
    1.862587859508221e-160, 3.30668408201432789e-169, 3.43017280887946242e-178,
    2.07915397775808218e-187, 7.3638454532398496e-197, 1.52394760394085743e-206,
    1.84281935046532101e-216, 1.30209553802992926e-226, 5.37588903521080535e-237,
    1.29689584599763147e-247, 1.82813078022866562e-258, 1.50576355348684241e-269,
    7.24692320799294216e-281, 2.03797051314726835e-292, 3.34880215927873805e-304,
    0.0};

__device__ __forceinline__ double sinc_of(double x, double sinx)  // Faddeeva.cpp:609-611
{
    return fabs(x) < 1e-4 ? 1 - 0.1666666666666666666667 * x * x : sinx / x;
}

// Re w(xin + i y) for y >= 0.  erfcx_y = erfcx(y) is a per-particle constant, hoisted by callers
// (the reference recomputes it per call, Faddeeva.cpp:905-907).
__device__ double voigt_exact(double xin, double y, double erfcx_y)
{
    if (xin == 0.0) return erfcx_y;          // :681-683
    if (y == 0.0) return exp(-xin * xin);    // :684-686
    const double a = 0.518321480430085929872;
    const double c = 0.329973702884629072537;
    const double a2 = 0.268657157075235951582;
    const double relerr = DBL_EPSILON;
    const double x = fabs(xin);
    double ret;
    double sum1 = 0, sum2 = 0, sum3 = 0, sum5 = 0;

    if (y > 7 || (x > 6 && (y > 0.1 || (x > 8 && y > 1e-10) || x > 28))) {  // :712-717
        const double ispi = 0.56418958354775628694807945156;
        if (x + y > 4000) {
            if (x + y > 1e7) {  // w(z) ~ i/sqrt(pi)/z
                if (x > y) {
                    const double yax = y / x;
                    const double denom = ispi / (x + yax * y);
                    return denom * yax;
                }
                if (isinf(y)) return isnan(x) ? nan("") : 0.0;
                const double xya = x / y;
                return ispi / (xya * x + y);
            }
            const double dr = x * x - y * y - 0.5, di = 2 * x * y;
            const double denom = ispi / (dr * dr + di * di);
            return denom * (x * di - y * dr);
        }
        // Laplace continued fraction with the fitted term count nu(z), :758-772
        double nu = floor(3.9 + 11.398 / (0.08254 * x + 0.1421 * y + 0.2023));
        double wr = x, wi = y;
        for (nu = 0.5 * (nu - 1); nu > 0.4; nu -= 0.5) {
            const double denom = nu / (wr * wr + wi * wi);
            wr = x - wr * denom;
            wi = y + wi * denom;
        }
        const double denom = ispi / (wr * wr + wi * wi);
        return denom * wi;
    } else if (x < 10) {  // Algorithm-916 style exponential sums, :816-922
        double prod2ax = 1, prodm2ax = 1;
        double expx2;
        if (isnan(y)) return y;
        if (x < 5e-4) {  // :828-851
            const double x2 = x * x;
            expx2 = 1 - x2 * (1 - 0.5 * x2);
            const double ax2 = 1.036642960860171859744 * x;
            const double exp2ax = 1 + ax2 * (1 + ax2 * (0.5 + 0.166666666666666666667 * ax2));
            const double expm2ax = 1 - ax2 * (1 - ax2 * (0.5 - 0.166666666666666666667 * ax2));
            for (int n = 1;; ++n) {
                const double coef = c_expa2n2[n - 1] * expx2 / (a2 * (n * n) + y * y);
                prod2ax *= exp2ax;
                prodm2ax *= expm2ax;
                sum1 += coef;
                sum2 += coef * prodm2ax;
                sum3 += coef * prod2ax;
                if (coef * prod2ax < relerr * sum3 || n >= 52) break;
            }
        } else {  // :852-867 — terminates on sum5 although only sum1..3 enter Re w
            expx2 = exp(-x * x);
            const double exp2ax = exp((2 * a) * x), expm2ax = 1 / exp2ax;
            for (int n = 1;; ++n) {
                const double coef = c_expa2n2[n - 1] * expx2 / (a2 * (n * n) + y * y);
                prod2ax *= exp2ax;
                prodm2ax *= expm2ax;
                sum1 += coef;
                sum2 += coef * prodm2ax;
                sum3 += coef * prod2ax;
                sum5 += (coef * prod2ax) * (a * n);
                if ((coef * prod2ax) * (a * n) < relerr * sum5 || n >= 52) break;
            }
        }
        const double expx2erfcxy = expx2 * erfcx_y;  // :905-907 (y > -6 always here)
        if (y > 5) {  // :908-912
            const double sinxy = sin(x * y);
            ret = (expx2erfcxy - c * y * sum1) * cos(2 * x * y) + (c * x * expx2) * sinxy * sinc_of(x * y, sinxy);
        } else {  // :913-921 (real part is even in x)
            const double sinxy = sin(x * y);
            const double cos2xy = cos(2 * x * y);
            const double coef1 = expx2erfcxy - c * y * sum1;
            const double coef2 = c * x * expx2;
            ret = coef1 * cos2xy + coef2 * sinxy * sinc_of(x * y, sinxy);
        }
    } else {  // x >= 10 with y <= 1e-10: :923-967
        if (isnan(x)) return x;
        if (isnan(y)) return y;
        ret = exp(-x * x);
        const double n0 = floor(x / a + 0.5);
        const double dx = a * n0 - x;
        sum3 = exp(-dx * dx) / (a2 * (n0 * n0) + y * y);
        sum5 = a * n0 * sum3;
        const double exp1 = exp(4 * a * dx);
        double exp1dn = 1;
        int dn;
        bool done = false;
        for (dn = 1; n0 - dn > 0; ++dn) {
            const double np = n0 + dn, nm = n0 - dn;
            double tp = exp(-(a * dn + dx) * (a * dn + dx));
            double tm = tp * (exp1dn *= exp1);
            tp /= (a2 * (np * np) + y * y);
            tm /= (a2 * (nm * nm) + y * y);
            sum3 += tp + tm;
            sum5 += a * (np * tp + nm * tm);
            if (a * (np * tp + nm * tm) < relerr * sum5) {
                done = true;
                break;
            }
        }
        while (!done) {
            const double np = n0 + dn++;
            const double tp = exp(-(a * dn + dx) * (a * dn + dx)) / (a2 * (np * np) + y * y);
            sum3 += tp;
            sum5 += a * np * tp;
            if (a * np * tp < relerr * sum5) done = true;
        }
    }
    return ret + (0.5 * c) * y * (sum2 + sum3);  // :968-970
}

// ---- fast path -----------------------------------------------------------------------------

constexpr double kFarXMin = 12.0;     // the damping-wing series is used from here on; the table reaches FSB_GTAB_XMAX = 24
constexpr double kFastYMax = 0.03;    // above: voigt_exact (error of the y^7 truncation < 3e-13 below)
constexpr double kFastYMin = 1e-30;   // below (and > 0): voigt_exact (Gaussian cut-off would pass exp underflow)

// FP32 fast path: degree-3 pieces on the same intervals, {c0, c1, c2, c3} per interval.
__device__ __align__(16) const float d_gtable32[4 * FSB_GTAB_NINT] = FSB_GTAB32_VALUES;

// y-dependent polynomial coefficients (derived with sympy from w' = -2zw + 2i/sqrt(pi)).
struct FastCoef {
    double pe[4];  // Pe(s): even orders y^0..y^6, multiplies U
    double a[4];   // A(s):  multiplies G
    double b[3];   // B(s)
    double xU2;    // x^2 beyond which exp(-x^2) is below 1e-14 of the Lorentzian wing
    double y;
};

__device__ __forceinline__ void fast_coefs(double y, FastCoef &c)
{
    const double isp = 0.56418958354775628694807945156;  // 1/sqrt(pi)
    const double y2 = y * y, y3 = y * y2, y5 = y3 * y2, y7 = y5 * y2;
    const double e = 6 + y2 * (6 + y2 * (3 + y2));
    const double f = 2 + y2 * (2 + y2);
    c.pe[0] = e / 6;
    c.pe[1] = -y2 * f;
    c.pe[2] = (2. / 3) * y2 * y2 * (1 + y2);
    c.pe[3] = -(4. / 45) * y2 * y2 * y2;
    c.a[0] = -isp * y * e / 3;
    c.a[1] = isp * (2. / 3) * y3 * f;
    c.a[2] = -isp * (4. / 15) * y5 * (1 + y2);
    c.a[3] = isp * (8. / 315) * y7;
    c.b[0] = isp * y3 * (70 + y2 * (49 + 19 * y2)) / 105;
    c.b[1] = -isp * 2 * y5 * (7 + 6 * y2) / 105;
    c.b[2] = isp * (4. / 315) * y7;
    c.xU2 = y > 0 ? 37.0 - log(y) : 1e300;
    c.y = y;
}

// ---- the G(x) table (fsb_voigt_tables.h) ------------------------------------------------------------------
// 193 intervals of width 1/8 centred on k/8, degree 7 in t = |x| - k/8, 48 bytes per interval: three 16-byte
// pieces {c0, c1}, {c2, (c3, c4)}, {(c5, c6), (c7, -)} with c3..c7 as floats.  The intervals are deliberately
// COARSE: the lanes of a quarter-warp hold adjacent pixels, a fraction of an interval apart, so their 16-byte
// loads hit the same or consecutive intervals, which the 48-byte stride (an odd multiple of 16 bytes) spreads
// over distinct banks.  A finer table of lower degree (32 bytes per interval at h = 1/64, tried) makes the lanes
// hit intervals several entries apart at an irregular stride: 2.5 wavefronts per quarter-warp instead of 1, whatever
// the slot permutation (profiles/README.md).
__device__ __align__(16) const unsigned long long d_gtable_words[FSB_GTAB_SIZE] = FSB_GTAB_WORDS;
#define d_gtable (reinterpret_cast<const double2 *>(d_gtable_words))
constexpr int kGtabPieces = FSB_GTAB_SIZE / 2;  // double2 pieces of the staged table

// Table index and offset of |x| < 24: k = rint(8|x|), t = |x| - k/8.
__device__ __forceinline__ void g_index(double ax, int &k, double &t)
{
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: low mantissa bits = rint(8|x|)
    const double m = fma(ax, FSB_GTAB_INV_DELTA, magic);
    k = __double2loint(m);
    t = fma(m - magic, -1.0 / FSB_GTAB_INV_DELTA, ax);
}

// Degree-7 polynomial of one interval at offset t from its three 16-byte pieces: the t^3..t^7 part runs in single
// precision (|t| <= 1/16 scales its rounding error by 2^-12 or less), the rest in double.
__device__ __forceinline__ double g_poly(double2 v0, double2 v1, double2 v2, double t)
{
    const float tf = (float) t;
    float hi = __int_as_float(__double2loint(v2.y));                 // c7
    hi = fmaf(hi, tf, __int_as_float(__double2hiint(v2.x)));         // c6
    hi = fmaf(hi, tf, __int_as_float(__double2loint(v2.x)));         // c5
    hi = fmaf(hi, tf, __int_as_float(__double2hiint(v1.y)));         // c4
    hi = fmaf(hi, tf, __int_as_float(__double2loint(v1.y)));         // c3
    return fma(fma(fma((double) hi, t, v1.x), t, v0.y), t, v0.x);
}

// G(|x|) for |x| < FSB_GTAB_XMAX from the table (tab may be the shared-memory copy or the global master).
__device__ __forceinline__ double g_table(double ax, const double2 *__restrict__ tab)
{
    int k;
    double t;
    g_index(ax, k, t);
    k = (int) min((unsigned) k, (unsigned) (FSB_GTAB_NINT - 1));
    const double2 *c = tab + 3 * k;
    return g_poly(c[0], c[1], c[2], t);
}

// Damping wing for |x| >= 12, u = 1/x^2:  H = (y/sqrt(pi)) u [P1(u) - (y^2 u) P3(u) + (y^2 u)^2 P5(u)],
// P1 = sum c_k (2k+1) u^k, P3 = sum c_k C(2k+3,3) u^k, P5 = sum c_k C(2k+5,5) u^k, c_k = (2k-1)!!/2^k.
// Truncation error < 3e-14 relative for y <= 0.03 (scripts/voigt_design.py).
__device__ __forceinline__ void far_polys(double u, double &p1, double &p3, double &p5)
{
    p1 = fma(fma(fma(fma(fma(fma(fma(fma(fma(1278767.724609375, u, 134607.12890625), u, 15836.1328125), u, 2111.484375), u, 324.84375), u, 59.0625), u, 13.125), u, 3.75), u, 1.5), u, 1.0);
    p3 = fma(fma(fma(fma(fma(fma(73901.953125, u, 8445.9375), u, 1082.8125), u, 157.5), u, 26.25), u, 5.0), u, 1.0);
    p5 = fma(10.5, u, 1.0);
}

// exp(x) for |x| <= 700 (arguments beyond are clamped): Cody-Waite reduction x = k ln2 + r, |r| <= ln2 / 2, the Taylor
// polynomial of degree 12 (remainder 1.7e-16 relative) and the exponent added into the high word; none of the library
// routine's special cases (NaN, overflow, denormal results): 24 instructions where exp() inlines 60.  Relative error
// below 4e-16.  The Gaussian factors of the march and of the per-particle set-up are the only callers.
__device__ __forceinline__ double fast_exp(double x)
{
    x = fmin(fmax(x, -700.0), 700.0);
    const double magic = 6755399441055744.0;  // 2^52 + 2^51: the sum's low word holds rint(x log2 e)
    const double t = fma(x, 1.4426950408889634074, magic);
    const int k = __double2loint(t);
    const double kd = t - magic;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 2.08767569878680989792e-09;  // 1/12!
    p = fma(p, r, 2.50521083854417187751e-08);
    p = fma(p, r, 2.75573192239858906526e-07);
    p = fma(p, r, 2.75573192239858906526e-06);
    p = fma(p, r, 2.48015873015873015873e-05);
    p = fma(p, r, 1.98412698412698412698e-04);
    p = fma(p, r, 1.38888888888888888889e-03);
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// 1/s for a normal positive s (the wing series has s = x^2 >= 256): hardware seed (relative error 2^-23) and
// two Newton steps, none of the library routine's special-case handling.  Within an ulp of 1/s.
__device__ __forceinline__ double fast_rcp(double s)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    double e = fma(-s, r, 1.0);
    r = fma(r, e, r);
    e = fma(-s, r, 1.0);
    return fma(r, e, r);
}

__device__ __forceinline__ double voigt_far(double s, double y)
{
    const double isp = 0.56418958354775628694807945156;
    const double u = 1.0 / s;
    double p1, p3, p5;
    far_polys(u, p1, p3, p5);
    const double v = y * y * u;
    return isp * y * u * fma(v, fma(v, p5, -p3), p1);
}

// One profile value with a known U = exp(-x^2) (or 0 where it is negligible).
__device__ __forceinline__ double voigt_fast_with_u(double ax, double s, double U, const FastCoef &c,
                                                    const double2 *__restrict__ tab)
{
    if (ax >= kFarXMin) return voigt_far(s, c.y);
    const double G = g_table(ax, tab);
    const double Pe = fma(fma(fma(c.pe[3], s, c.pe[2]), s, c.pe[1]), s, c.pe[0]);
    const double A = fma(fma(fma(c.a[3], s, c.a[2]), s, c.a[1]), s, c.a[0]);
    const double B = fma(fma(c.b[2], s, c.b[1]), s, c.b[0]);
    return fma(U, Pe, fma(G, A, B));
}

// Stand-alone evaluation (tests, and any caller without a shared U recurrence).
__device__ __forceinline__ double voigt_fast(double x, const FastCoef &c, const double2 *__restrict__ tab)
{
    const double ax = fabs(x), s = x * x;
    const double U = s < c.xU2 ? exp(-s) : 0.0;
    return voigt_fast_with_u(ax, s, U, c, tab);
}

// y == 0 (gamma = 0, spectra.py:669-672) takes the exact path, whose y == 0 branch is exp(-x^2).
__device__ __forceinline__ bool fast_domain(double y) { return y >= kFastYMin && y <= kFastYMax; }

}  // namespace fsb
