/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the fake_spectra sightline-interpolation
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (fake_spectra_b200) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below against the
 * unmodified reference compiled into oracle/_ref/libfsref.so (oracle/build.py), and
 * tests/test_oracle_kat.py checks it against the known-answer values held by the reference's
 * own test.cpp and Faddeeva self-test (values transcribed as data, committed under tests/golden).
 *
 * All "ref:" citations are relative to /root/reference/fake_spectra/.
 * Plain C11, strict IEEE (compiled with -ffp-contract=off, no -ffast-math).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define FSO_TOPHAT 0   /* ref: singleabs.h:9-12 */
#define FSO_CUBIC 1
#define FSO_VORONOI 2
#define FSO_QUINTIC 3
#define FSO_NGRID 8    /* ref: singleabs.h:8 */

/* ref: absorption.cpp:21-26, absorption.h:4 */
static const double K_SIGMA_T = 6.652458558e-25;
static const double K_BOLTZMANN = 1.3806504e-16;
static const double K_LIGHT = 2.99792458e10;
static const double K_PROTONMASS = 1.67262178e-24;
static const double K_PI = 3.14159265358979323846;

/* ------------------------------------------------------------------------------------------
 * Faddeeva: only Re w(x + i y) is consumed (ref: singleabs.h:56-61).
 * ---------------------------------------------------------------------------------------- */

/* Scaled complementary error function exp(y^2) erfc(y) for real y.
 * ref: Faddeeva.cpp:1421-1437 uses a Chebyshev table + continued fraction; this restatement uses
 * the defining identity with libm erfc for |y| < 10 and the Laplace continued fraction above
 * (both accurate to a few ulp there), and the reflection erfcx(-y) = 2 exp(y^2) - erfcx(y). */
static double fso_erfcx_pos(double y)
{
    if (y < 10.0)
        return exp(y * y) * erfc(y);
    /* erfcx(y) = (1/sqrt(pi)) / (y + (1/2)/(y + 1/(y + (3/2)/(y + ...)))) evaluated bottom-up */
    double t = y;
    for (int k = 60; k >= 1; --k)
        t = y + 0.5 * k / t;
    return 0.56418958354775628694807945156 / t;
}

double fso_erfcx(double y)
{
    if (y >= 0)
        return fso_erfcx_pos(y);
    if (y < -26.7)
        return HUGE_VAL;
    return 2.0 * exp(y * y) - fso_erfcx_pos(-y);
}

static double fso_sinc(double x, double sinx) /* ref: Faddeeva.cpp:609-611 */
{
    return fabs(x) < 1e-4 ? 1 - 0.1666666666666666666667 * x * x : sinx / x;
}

/* exp(-a2 n^2) for a2 = 0.26865..., n = 1..52.  ref: Faddeeva.cpp:622-675 holds these as a literal
 * table; here they are generated once (long double) and checked against the reference by the
 * w(z) parity tests. */
static double g_expa2n2[53];
static int g_expa2n2_ready = 0;
static void fso_init_tables(void)
{
    if (g_expa2n2_ready)
        return;
    #pragma omp critical(fso_tables)
    {
        if (!g_expa2n2_ready) {
            const long double a2 = 0.268657157075235951582L;
            for (int n = 1; n <= 52; ++n) {
                long double v = expl(-a2 * (long double) n * (long double) n);
                g_expa2n2[n - 1] = (n == 52) ? 0.0 : (double) v; /* ref: :674 last entry is 0 */
            }
            g_expa2n2[52] = 0.0;
            g_expa2n2_ready = 1;
        }
    }
}

/* Real part of w(z), z = xin + i y, machine-precision branch (relerr = DBL_EPSILON).
 * ref: Faddeeva.cpp:679-971. */
double fso_faddeeva_re(double xin, double y)
{
    if (xin == 0.0)
        return fso_erfcx(y); /* ref: :681-683 */
    if (y == 0.0)
        return exp(-xin * xin); /* ref: :684-686 */
    fso_init_tables();

    const double a = 0.518321480430085929872;  /* ref: :689-694 */
    const double c = 0.329973702884629072537;
    const double a2 = 0.268657157075235951582;
    const double relerr = DBL_EPSILON;
    const double x = fabs(xin), ya = fabs(y);
    double ret = 0.0;
    double sum1 = 0, sum2 = 0, sum3 = 0, sum5 = 0;

    if (ya > 7 || (x > 6 && (ya > 0.1 || (x > 8 && ya > 1e-10) || x > 28))) { /* ref: :712-717 */
        const double ispi = 0.56418958354775628694807945156;
        const double xs = y < 0 ? -xin : xin;
        if (x + ya > 4000) { /* ref: :733-757 */
            if (x + ya > 1e7) {
                if (x > ya) {
                    const double yax = ya / xs;
                    const double denom = ispi / (xs + yax * ya);
                    ret = denom * yax;
                } else if (isinf(ya)) {
                    return (isnan(x) || y < 0) ? NAN : 0.0;
                } else {
                    const double xya = xs / ya;
                    const double denom = ispi / (xya * xs + ya);
                    ret = denom;
                }
            } else {
                const double dr = xs * xs - ya * ya - 0.5, di = 2 * xs * ya;
                const double denom = ispi / (dr * dr + di * di);
                ret = denom * (xs * di - ya * dr);
            }
        } else { /* general continued fraction, ref: :758-772 */
            const double c0 = 3.9, c1 = 11.398, c2 = 0.08254, c3 = 0.1421, c4 = 0.2023;
            double nu = floor(c0 + c1 / (c2 * x + c3 * ya + c4));
            double wr = xs, wi = ya;
            for (nu = 0.5 * (nu - 1); nu > 0.4; nu -= 0.5) {
                const double denom = nu / (wr * wr + wi * wi);
                wr = xs - wr * denom;
                wi = ya + wi * denom;
            }
            const double denom = ispi / (wr * wr + wi * wi);
            ret = denom * wi;
        }
        if (y < 0) { /* ref: :773-778: 2 exp(-z^2) - w(-z), real part (cexp gives 0 when it underflows) */
            const double mag = exp((ya - xs) * (xs + ya));
            return mag == 0.0 ? -ret : 2.0 * mag * cos(2 * xs * y) - ret;
        }
        return ret;
    } else if (x < 10) { /* Algorithm-916-style series, ref: :816-922 */
        double prod2ax = 1, prodm2ax = 1;
        double expx2;
        if (isnan(y))
            return y;
        if (x < 5e-4) { /* ref: :828-851 */
            const double x2 = x * x;
            expx2 = 1 - x2 * (1 - 0.5 * x2);
            const double ax2 = 1.036642960860171859744 * x;
            const double exp2ax = 1 + ax2 * (1 + ax2 * (0.5 + 0.166666666666666666667 * ax2));
            const double expm2ax = 1 - ax2 * (1 - ax2 * (0.5 - 0.166666666666666666667 * ax2));
            for (int n = 1;; ++n) {
                const double coef = g_expa2n2[n - 1] * expx2 / (a2 * (n * n) + y * y);
                prod2ax *= exp2ax;
                prodm2ax *= expm2ax;
                sum1 += coef;
                sum2 += coef * prodm2ax;
                sum3 += coef * prod2ax;
                if (coef * prod2ax < relerr * sum3)
                    break;
            }
        } else { /* ref: :852-867; the loop stops on sum5 although only sum1..3 feed Re w */
            expx2 = exp(-x * x);
            const double exp2ax = exp((2 * a) * x), expm2ax = 1 / exp2ax;
            for (int n = 1;; ++n) {
                const double coef = g_expa2n2[n - 1] * expx2 / (a2 * (n * n) + y * y);
                prod2ax *= exp2ax;
                prodm2ax *= expm2ax;
                sum1 += coef;
                sum2 += coef * prodm2ax;
                sum3 += coef * prod2ax;
                sum5 += (coef * prod2ax) * (a * n);
                if ((coef * prod2ax) * (a * n) < relerr * sum5)
                    break;
            }
        }
        const double expx2erfcxy = y > -6 ? expx2 * fso_erfcx(y) : 2 * exp(y * y - x * x); /* ref: :905-907 */
        if (y > 5) { /* ref: :908-912 */
            const double sinxy = sin(x * y);
            ret = (expx2erfcxy - c * y * sum1) * cos(2 * x * y) + (c * x * expx2) * sinxy * fso_sinc(x * y, sinxy);
        } else { /* ref: :913-921 */
            const double xs = xin;
            const double sinxy = sin(xs * y);
            const double cos2xy = cos(2 * xs * y);
            const double coef1 = expx2erfcxy - c * y * sum1;
            const double coef2 = c * xs * expx2;
            ret = coef1 * cos2xy + coef2 * sinxy * fso_sinc(xs * y, sinxy);
        }
    } else { /* x >= 10 and |y| tiny: ref: :923-967 */
        if (isnan(x))
            return x;
        if (isnan(y))
            return y;
        ret = exp(-x * x);
        const double n0 = floor(x / a + 0.5);
        const double dx = a * n0 - x;
        sum3 = exp(-dx * dx) / (a2 * (n0 * n0) + y * y);
        sum5 = a * n0 * sum3;
        const double exp1 = exp(4 * a * dx);
        double exp1dn = 1;
        int dn, done = 0;
        for (dn = 1; n0 - dn > 0; ++dn) {
            const double np = n0 + dn, nm = n0 - dn;
            double tp = exp(-(a * dn + dx) * (a * dn + dx));
            double tm = tp * (exp1dn *= exp1);
            tp /= (a2 * (np * np) + y * y);
            tm /= (a2 * (nm * nm) + y * y);
            sum3 += tp + tm;
            sum5 += a * (np * tp + nm * tm);
            if (a * (np * tp + nm * tm) < relerr * sum5) {
                done = 1;
                break;
            }
        }
        while (!done) {
            const double np = n0 + dn++;
            const double tp = exp(-(a * dn + dx) * (a * dn + dx)) / (a2 * (np * np) + y * y);
            sum3 += tp;
            sum5 += a * np * tp;
            if (a * np * tp < relerr * sum5)
                done = 1;
        }
    }
    return ret + (0.5 * c) * y * (sum2 + sum3); /* ref: :968-970 */
}

void fso_faddeeva_re_many(const double *x, const double *y, double *out, long long n)
{
    fso_init_tables();
    #pragma omp parallel for
    for (long long i = 0; i < n; ++i)
        out[i] = fso_faddeeva_re(x[i], y[i]);
}

/* ------------------------------------------------------------------------------------------
 * SPH kernels and their line integrals.
 * ---------------------------------------------------------------------------------------- */

double fso_cubic_kernel(double q) /* ref: singleabs.h:17-26 */
{
    const double norm = 32. / 4 / K_PI;
    if (q >= 1)
        return 0;
    if (q < 0.5)
        return norm * (1 - 6 * q * q + 6 * q * q * q);
    return norm * (2 * pow(1. - q, 3));
}

double fso_quintic_kernel(double q) /* ref: singleabs.h:31-42 */
{
    const double norm = 9. / 40 / K_PI;
    if (q >= 1)
        return 0;
    if (q < (1. / 3))
        return norm * 6 * (11 - 90 * q * q + 405 * q * q * q * q - 405 * q * q * q * q * q);
    if (q < (2. / 3) && q >= (1. / 3))
        return norm * (pow(3. - 3 * q, 5) - 6 * pow(2. - 3 * q, 5));
    return norm * (243 * pow(1. - q, 5));
}

/* 8-interval trapezoid of K(sqrt(dr2+z^2)/smooth) clipped to +-zrange.
 * ref: absorption.cpp:53-74 (cubic), :76-97 (quintic). */
static double fso_spline_frac(double (*kern)(double), double zlow, double zhigh, double smooth, double dr2, double zrange)
{
    zlow = zlow > -zrange ? zlow : -zrange;
    zhigh = zhigh < zrange ? zhigh : zrange;
    if (zlow > zhigh)
        return 0;
    const double qlow = sqrt(dr2 + zlow * zlow) / smooth;
    double total = kern(qlow) / 2.;
    const double deltaz = (zhigh - zlow) / FSO_NGRID;
    for (int i = 1; i < FSO_NGRID; ++i) {
        const double zz = i * deltaz + zlow;
        const double q = sqrt(dr2 + zz * zz) / smooth;
        total += kern(q);
    }
    const double qhigh = sqrt(dr2 + zhigh * zhigh) / smooth;
    total += kern(qhigh) / 2.;
    return deltaz * total;
}

/* ref: absorption.cpp:136-148 dispatch; :109-115 tophat; :129-134 arepo. */
double fso_kern_frac(int kernel, double zlow, double zhigh, double smooth, double dr2, double zrange)
{
    if (kernel == FSO_CUBIC)
        return fso_spline_frac(fso_cubic_kernel, zlow, zhigh, smooth, dr2, zrange);
    if (kernel == FSO_QUINTIC)
        return fso_spline_frac(fso_quintic_kernel, zlow, zhigh, smooth, dr2, zrange);
    zlow = zlow > -zrange ? zlow : -zrange;
    zhigh = zhigh < zrange ? zhigh : zrange;
    const double len = (zhigh - zlow) > 0. ? (zhigh - zlow) : 0.;
    if (kernel == FSO_VORONOI)
        return len;
    return 3. / 4. / K_PI * len;
}

/* ------------------------------------------------------------------------------------------
 * Line constants and the per-particle accumulators.
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    double tautail, sigma_a, bfac, voigt_fac, velfac, vbox, atime;
    int kernel;
} fso_line;

/* ref: absorption.cpp:152-161 */
static fso_line fso_line_init(double lambda, double gamma, double fosc, double amumass, double velfac,
                              double box, double atime, int kernel, double tautail)
{
    fso_line L;
    L.tautail = tautail;
    L.sigma_a = sqrt(3.0 * K_PI * K_SIGMA_T / 8.0) * lambda * fosc;
    L.bfac = sqrt(2.0 * K_BOLTZMANN / (amumass * K_PROTONMASS)) / 1e5;
    L.voigt_fac = gamma * lambda / (4. * K_PI) / 1e5;
    L.velfac = velfac;
    L.vbox = box * velfac;
    L.atime = atime;
    L.kernel = kernel;
    return L;
}

/* ref: absorption.cpp:167-210 */
static void fso_add_colden(const fso_line *L, double *colden, int nbins, double dr2, float dens, float pos, float smooth)
{
    double pos1 = pos;
    if (L->kernel == FSO_VORONOI) {
        if (dr2 > 2 * L->vbox / L->velfac || smooth > 2 * L->vbox / L->velfac)
            return;
        pos1 = (dr2 + smooth) / 2.;
    } else {
        if (smooth * smooth - dr2 <= 0) /* float product, then double subtraction */
            return;
    }
    double zrange = sqrt(smooth * smooth - dr2);
    if (L->kernel == FSO_VORONOI)
        zrange = (smooth - dr2) / 2.;
    const double boxtokpc = L->vbox / nbins / L->velfac;
    const int zlow = (int) floor((pos1 - zrange) / boxtokpc);
    const int zhigh = (int) ceil((pos1 + zrange) / boxtokpc);
    for (int z = zlow; z <= zhigh; z++) {
        const double plow = boxtokpc * z - pos1;
        int j = z % nbins;
        if (j < 0)
            j += nbins;
        colden[j] += dens * fso_kern_frac(L->kernel, plow, plow + boxtokpc, smooth, dr2, zrange);
    }
}

/* ref: singleabs.h:63-175 (SingleAbsorber) */
typedef struct {
    double btherm, vdr2, vsmooth, aa, vhigh;
    int kernel;
} fso_absorber;

static fso_absorber fso_absorber_init(double btherm, double vdr2, double vsmooth, double aa, int kernel)
{
    fso_absorber A = {btherm, vdr2, vsmooth, aa, 0.0, kernel};
    A.vhigh = (vsmooth * vsmooth > vdr2) ? sqrt(vsmooth * vsmooth - vdr2) : 0; /* ref: :83 */
    if (kernel == FSO_VORONOI)                                                   /* ref: :85-89 */
        A.vhigh = (vdr2 > 0 && vsmooth > 0) ? (vsmooth - vdr2) / 2. : 0;
    return A;
}

/* ref: singleabs.h:143-167 */
static double fso_tau_kern_inner(const fso_absorber *A, double vouter)
{
    const double deltav = 2. * A->vhigh / FSO_NGRID;
    double total = 0;
    for (int i = 1; i < FSO_NGRID; ++i) {
        const double vv = i * deltav - A->vhigh;
        const double q = sqrt(A->vdr2 + vv * vv) / A->vsmooth;
        const double vdiff = vv - vouter;
        const double T0 = vdiff / A->btherm;
        double tbin = fso_faddeeva_re(T0, A->aa);
        if (A->kernel == FSO_CUBIC)
            tbin *= fso_cubic_kernel(q);
        else if (A->kernel == FSO_QUINTIC)
            tbin *= fso_quintic_kernel(q);
        else if (A->kernel == FSO_TOPHAT)
            tbin *= 3. / 4. / K_PI;
        total += tbin;
    }
    return deltav * total;
}

/* ref: singleabs.h:104-126 */
static double fso_tau_kern_outer(const fso_absorber *A, double vlow, double vhigh)
{
    if ((vhigh - vlow) < A->btherm / 2.)
        return fso_tau_kern_inner(A, (vhigh + vlow) / 2.);
    const int npoints = (int) (2 * ceil((vhigh - vlow) / (A->btherm / 2.) / 2) + 1.);
    double total = fso_tau_kern_inner(A, vlow) / 2.;
    const double deltav = (vhigh - vlow) / (npoints - 1);
    for (int i = 1; i < npoints - 1; ++i) {
        const double vv = i * deltav + vlow;
        total += fso_tau_kern_inner(A, vv);
    }
    total += fso_tau_kern_inner(A, vhigh) / 2.;
    return total / (npoints - 1);
}

/* ref: absorption.cpp:212-279 */
static void fso_add_tau(const fso_line *L, double *tau, int nbins, double dr2, float dens, float ppos, float pvel,
                        float temp, float smooth)
{
    double pos1 = ppos;
    const double btherm = L->bfac * sqrt((double) temp);
    if (L->kernel == FSO_VORONOI) {
        if (dr2 > 2 * L->vbox / L->velfac || smooth > 2 * L->vbox / L->velfac)
            return;
        pos1 = (dr2 + smooth) / 2.;
    } else {
        if (smooth * smooth - dr2 <= 0)
            return;
    }
    const double vel = L->velfac * pos1 + pvel;
    double val1 = L->velfac * dr2;
    if (L->kernel != FSO_VORONOI)
        val1 *= L->velfac;
    const fso_absorber A = fso_absorber_init(btherm, val1, L->velfac * smooth, L->voigt_fac / btherm, L->kernel);
    const double bintov = L->vbox / nbins;
    const double amp = L->sigma_a / sqrt(K_PI) * (K_LIGHT / 1e5 / btherm);
    const int zmax = (int) floor(vel / bintov);
    for (int z = zmax; z < zmax + nbins / 2; ++z) {
        const double vlow = z * bintov - vel;
        const double taulast = amp * dens * fso_tau_kern_outer(&A, vlow, vlow + bintov) / L->velfac;
        int j = z % nbins;
        if (j < 0)
            j += nbins;
        tau[j] += taulast;
        if (taulast < L->tautail)
            break;
    }
    for (int z = zmax - 1; z >= zmax - nbins / 2; --z) {
        const double vlow = z * bintov - vel;
        const double taulast = amp * dens * fso_tau_kern_outer(&A, vlow, vlow + bintov) / L->velfac;
        int j = z % nbins;
        if (j < 0)
            j += nbins;
        tau[j] += taulast;
        if (taulast < L->tautail)
            break;
    }
}

/* Single-particle entry points for the known-answer tests. */
void fso_add_colden_particle(double lambda, double gamma, double fosc, double amumass, double velfac, double box,
                             double atime, int kernel, double tautail, double *colden, int nbins, double dr2,
                             float dens, float ppos, float smooth)
{
    const fso_line L = fso_line_init(lambda, gamma, fosc, amumass, velfac, box, atime, kernel, tautail);
    fso_add_colden(&L, colden, nbins, dr2, dens, ppos, smooth);
}

void fso_add_tau_particle(double lambda, double gamma, double fosc, double amumass, double velfac, double box,
                          double atime, int kernel, double tautail, double *tau, int nbins, double dr2, float dens,
                          float ppos, float pvel, float temp, float smooth)
{
    const fso_line L = fso_line_init(lambda, gamma, fosc, amumass, velfac, box, atime, kernel, tautail);
    fso_add_tau(&L, tau, nbins, dr2, dens, ppos, pvel, temp, smooth);
}

double fso_tau_kern_outer_pub(double btherm, double vdr2, double vsmooth, double aa, int kernel, double vlow, double vhigh)
{
    const fso_absorber A = fso_absorber_init(btherm, vdr2, vsmooth, aa, kernel);
    return fso_tau_kern_outer(&A, vlow, vhigh);
}

/* ------------------------------------------------------------------------------------------
 * Sightline <-> particle search.
 * The reference keeps two std::multimap (ref: index_table.cpp:7-18): axis==1 lines keyed by
 * cofm[3i+1], axis 2/3 lines keyed by cofm[3i].  Here: two arrays sorted by key.
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    double key;
    int line;
} fso_ent;

typedef struct {
    const double *cofm;
    const int *axis;
    int nlos;
    double box;
    fso_ent *tab_x;  /* axis 2,3: key = x */
    int n_x;
    fso_ent *tab_yy; /* axis 1: key = y */
    int n_yy;
} fso_table;

static int fso_ent_cmp(const void *a, const void *b)
{
    const fso_ent *x = (const fso_ent *) a, *y = (const fso_ent *) b;
    if (x->key < y->key)
        return -1;
    if (x->key > y->key)
        return 1;
    return (x->line > y->line) - (x->line < y->line);
}

static fso_table fso_table_init(const double *cofm, const int *axis, int nlos, double box)
{
    fso_table T = {cofm, axis, nlos, box, NULL, 0, NULL, 0};
    T.tab_x = (fso_ent *) malloc(sizeof(fso_ent) * (size_t) (nlos > 0 ? nlos : 1));
    T.tab_yy = (fso_ent *) malloc(sizeof(fso_ent) * (size_t) (nlos > 0 ? nlos : 1));
    for (int i = 0; i < nlos; i++) {
        if (axis[i] == 1) {
            T.tab_yy[T.n_yy].key = cofm[3 * i + 1];
            T.tab_yy[T.n_yy++].line = i;
        } else {
            T.tab_x[T.n_x].key = cofm[3 * i];
            T.tab_x[T.n_x++].line = i;
        }
    }
    qsort(T.tab_x, (size_t) T.n_x, sizeof(fso_ent), fso_ent_cmp);
    qsort(T.tab_yy, (size_t) T.n_yy, sizeof(fso_ent), fso_ent_cmp);
    return T;
}

static void fso_table_free(fso_table *T)
{
    free(T->tab_x);
    free(T->tab_yy);
}

/* first entry with key >= v (std::multimap::lower_bound) */
static int fso_lower_bound(const fso_ent *tab, int n, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (tab[mid].key < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* ref: index_table.cpp:52-68 */
static int fso_second_close(double box, float second, double lproj2, float hh)
{
    float ffp = second + hh;
    if (ffp > box)
        if (lproj2 < ffp - box)
            return 1;
    float ffm = second - hh;
    if (ffm < 0)
        if (lproj2 > ffm + box)
            return 1;
    return (lproj2 > ffm && lproj2 < ffp);
}

/* ref: index_table.cpp:70-87 */
static double fso_calc_dr2(double box, double d1, double d2)
{
    double dr = fabs(d1);
    if (dr > 0.5 * box)
        dr = box - dr;
    double dr2 = dr * dr;
    dr = fabs(d2);
    if (dr > 0.5 * box)
        dr = box - dr;
    dr2 += (dr * dr);
    return dr2;
}

typedef void (*fso_hit_fn)(void *ctx, int line, double dr2);

/* ref: index_table.cpp:22-50 */
static void fso_scan_range(const fso_table *T, const fso_ent *tab, int lo, int hi, const float *pos, float hh,
                           float first, fso_hit_fn hit, void *ctx)
{
    for (int k = lo; k < hi; ++k) {
        const int iproc = tab[k].line;
        const int iaxis = T->axis[iproc];
        float second;
        double lproj2;
        if (iaxis == 3) {
            second = pos[1];
            lproj2 = T->cofm[3 * iproc + 1];
        } else {
            second = pos[2];
            lproj2 = T->cofm[3 * iproc + 2];
        }
        const double lproj = tab[k].key;
        if (fso_second_close(T->box, second, lproj2, hh)) {
            const double dr2 = fso_calc_dr2(T->box, first - lproj, second - lproj2);
            if (dr2 <= hh * hh) /* float product */
                hit(ctx, iproc, dr2);
        }
    }
}

/* ref: index_table.cpp:89-113 */
static void fso_nearby(const fso_table *T, float first, const fso_ent *tab, int n, const float *pos, float hh,
                       fso_hit_fn hit, void *ctx)
{
    float ffp = first + hh;
    if (ffp > T->box)
        ffp -= T->box;
    float ffm = first - hh;
    if (ffm < 0)
        ffm += T->box;
    const int low = fso_lower_bound(tab, n, ffm);
    const int high = fso_lower_bound(tab, n, ffp);
    if (ffm <= ffp) {
        fso_scan_range(T, tab, low, high, pos, hh, first, hit, ctx);
    } else {
        fso_scan_range(T, tab, 0, high, pos, hh, first, hit, ctx);
        fso_scan_range(T, tab, low, n, pos, hh, first, hit, ctx);
    }
}

/* ref: index_table.cpp:117-127 */
static void fso_near_lines_of(const fso_table *T, const float *pos, float hh, fso_hit_fn hit, void *ctx)
{
    if (T->n_x > 0)
        fso_nearby(T, pos[0], T->tab_x, T->n_x, pos, hh, hit, ctx);
    if (T->n_yy > 0)
        fso_nearby(T, pos[1], T->tab_yy, T->n_yy, pos, hh, hit, ctx);
}

typedef struct {
    long long *counts;
    const long long *offsets;
    long long *cursor;
    int *part;
    double *dr2;
    int ipart;
    int any;
} fso_fill_ctx;

static void fso_hit_count(void *vctx, int line, double dr2)
{
    (void) dr2;
    fso_fill_ctx *c = (fso_fill_ctx *) vctx;
    c->counts[line]++;
    c->any = 1;
}

static void fso_hit_fill(void *vctx, int line, double dr2)
{
    fso_fill_ctx *c = (fso_fill_ctx *) vctx;
    const long long slot = c->offsets[line] + c->cursor[line]++;
    c->part[slot] = c->ipart;
    c->dr2[slot] = dr2;
}

/* Per-line candidate lists (ascending particle index).  ref: index_table.cpp:130-150.
 * counts[nlos] is always filled; part/dr2 (CSR order, offsets = exclusive scan of counts) are
 * filled when non-NULL.  Returns the total number of pairs. */
long long fso_near_particles(const double *cofm, const int *axis, int nlos, double box, const float *pos,
                             const float *h, long long npart, long long *counts, int *part, double *dr2)
{
    fso_table T = fso_table_init(cofm, axis, nlos, box);
    fso_fill_ctx c;
    memset(&c, 0, sizeof(c));
    c.counts = counts;
    for (int i = 0; i < nlos; i++)
        counts[i] = 0;
    for (long long i = 0; i < npart; i++)
        fso_near_lines_of(&T, &pos[3 * i], h[i], fso_hit_count, &c);
    long long total = 0;
    for (int i = 0; i < nlos; i++)
        total += counts[i];
    if (part && dr2) {
        long long *offsets = (long long *) malloc(sizeof(long long) * (size_t) (nlos + 1));
        long long *cursor = (long long *) calloc((size_t) (nlos + 1), sizeof(long long));
        offsets[0] = 0;
        for (int i = 0; i < nlos; i++)
            offsets[i + 1] = offsets[i] + counts[i];
        c.offsets = offsets;
        c.cursor = cursor;
        c.part = part;
        c.dr2 = dr2;
        for (long long i = 0; i < npart; i++) {
            c.ipart = (int) i;
            fso_near_lines_of(&T, &pos[3 * i], h[i], fso_hit_fill, &c);
        }
        free(offsets);
        free(cursor);
    }
    fso_table_free(&T);
    return total;
}

/* Ascending indices of particles with at least one candidate line.  ref: py_module.cpp:63-89. */
long long fso_near_lines(double box, const float *pos, const float *h, long long npart, const int *axis,
                         const double *cofm, int nlos, int *out)
{
    fso_table T = fso_table_init(cofm, axis, nlos, box);
    long long *counts = (long long *) calloc((size_t) (nlos > 0 ? nlos : 1), sizeof(long long));
    long long n = 0;
    for (long long i = 0; i < npart; i++) {
        fso_fill_ctx c;
        memset(&c, 0, sizeof(c));
        c.counts = counts;
        fso_near_lines_of(&T, &pos[3 * i], h[i], fso_hit_count, &c);
        if (c.any) {
            if (out)
                out[n] = (int) i;
            ++n;
        }
    }
    free(counts);
    fso_table_free(&T);
    return n;
}

/* Voronoi cell extents along one sightline.  ref: index_table.cpp:152-223.
 * cand[ncells] = ascending candidate particle indices of this line; arr2[2*ncells] output.
 * Returns 0, or 1 when the reference would have hit its exit(1) guard (ref: :204-208). */
int fso_assign_cells(const double *cofm, const int *axis, double box, int line, const int *cand, int ncells,
                     const float *pos, float *arr2)
{
    for (int i = 0; i < 2 * ncells; ++i)
        arr2[i] = 3 * box;
    const int N = (int) (box / 0.1); /* RESO, ref: index_table.h:8 */
    const double reso = box / N;
    const int ax = axis[line];
    const double yp = cofm[3 * line + ax % 3], zp = cofm[3 * line + (ax + 1) % 3];
    for (int i = 0; i < N; ++i) {
        const double xp = (i + 0.5) * reso;
        double min_dist = box;
        int min_ind = 0;
        for (int ind = 0; ind < ncells; ++ind) {
            const int ip = cand[ind];
            double dx = fabs(pos[3 * ip + ax - 1] - xp);
            if (dx > box / 2.)
                dx = box - dx;
            double dy = fabs(pos[3 * ip + ax % 3] - yp);
            if (dy > box / 2.)
                dy = box - dy;
            double dz = fabs(pos[3 * ip + (ax + 1) % 3] - zp);
            if (dz > box / 2.)
                dz = box - dz;
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);
            if (dist < min_dist) {
                min_dist = dist;
                min_ind = ind;
            }
        }
        if (ncells == 0)
            break;
        if (arr2[2 * min_ind] < reso && xp > box / 2. + 0.5 * reso) {
            arr2[2 * min_ind] = xp;
            arr2[2 * min_ind + 1] += box;
            break;
        }
        if (xp > 1.5 * reso + arr2[2 * min_ind + 1])
            return 1;
        if (arr2[2 * min_ind] > 2 * box)
            arr2[2 * min_ind] = xp;
        arr2[2 * min_ind + 1] = xp;
    }
    for (int i = 0; i < ncells; ++i) {
        arr2[2 * i] -= 0.5 * reso;
        arr2[2 * i + 1] += 0.5 * reso;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Drivers.  ref: part_int.cpp:20-51 (tau), :53-84 (colden).  out[nlos*nbins] is accumulated into.
 * Returns 0, or 1 if a Voronoi invariant was violated.
 * ---------------------------------------------------------------------------------------- */
static int fso_compute(int do_tau, int nbins, double lambda, double gamma, double fosc, double amumass, double box,
                       double velfac, double atime, const double *cofm, const int *axis, int nlos, int kernel,
                       double tautail, double *out, const float *pos, const float *vel, const float *dens,
                       const float *temp, const float *h, long long npart)
{
    const fso_line L = fso_line_init(lambda, gamma, fosc, amumass, velfac, box, atime, kernel, tautail);
    long long *counts = (long long *) malloc(sizeof(long long) * (size_t) (nlos > 0 ? nlos : 1));
    long long total = fso_near_particles(cofm, axis, nlos, box, pos, h, npart, counts, NULL, NULL);
    int *part = (int *) malloc(sizeof(int) * (size_t) (total > 0 ? total : 1));
    double *dr2 = (double *) malloc(sizeof(double) * (size_t) (total > 0 ? total : 1));
    fso_near_particles(cofm, axis, nlos, box, pos, h, npart, counts, part, dr2);
    long long *offsets = (long long *) malloc(sizeof(long long) * (size_t) (nlos + 1));
    offsets[0] = 0;
    for (int i = 0; i < nlos; i++)
        offsets[i + 1] = offsets[i] + counts[i];
    int err = 0;
    fso_init_tables();
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < nlos; ++i) {
        const int ax = axis[i];
        double *row = &out[(size_t) i * (size_t) nbins];
        const long long o = offsets[i];
        const int nc = (int) counts[i];
        float *arr2 = NULL;
        if (kernel == FSO_VORONOI) {
            arr2 = (float *) malloc(sizeof(float) * (size_t) (2 * nc + 2));
            if (fso_assign_cells(cofm, axis, box, i, &part[o], nc, pos, arr2)) {
                #pragma omp atomic write
                err = 1;
                free(arr2);
                continue;
            }
        }
        for (int k = 0; k < nc; ++k) {
            const int ip = part[o + k];
            const float ppos = pos[3 * ip + ax - 1];
            const double d2 = (kernel == FSO_VORONOI) ? arr2[2 * k] : dr2[o + k];
            const float sm = (kernel == FSO_VORONOI) ? arr2[2 * k + 1] : h[ip];
            if (do_tau)
                fso_add_tau(&L, row, nbins, d2, dens[ip], ppos, vel[3 * ip + ax - 1], temp[ip], sm);
            else
                fso_add_colden(&L, row, nbins, d2, dens[ip], ppos, sm);
        }
        free(arr2);
    }
    free(offsets);
    free(dr2);
    free(part);
    free(counts);
    return err;
}

int fso_compute_tau(int nbins, double lambda, double gamma, double fosc, double amumass, double box, double velfac,
                    double atime, const double *cofm, const int *axis, int nlos, int kernel, double tautail,
                    double *tau, const float *pos, const float *vel, const float *dens, const float *temp,
                    const float *h, long long npart)
{
    return fso_compute(1, nbins, lambda, gamma, fosc, amumass, box, velfac, atime, cofm, axis, nlos, kernel, tautail,
                       tau, pos, vel, dens, temp, h, npart);
}

int fso_compute_colden(int nbins, double lambda, double gamma, double fosc, double amumass, double box, double velfac,
                       double atime, const double *cofm, const int *axis, int nlos, int kernel, double tautail,
                       double *colden, const float *pos, const float *dens, const float *h, long long npart)
{
    return fso_compute(0, nbins, lambda, gamma, fosc, amumass, box, velfac, atime, cofm, axis, nlos, kernel, tautail,
                       colden, pos, NULL, dens, NULL, h, npart);
}

int fso_omp_max_threads(void) { return omp_get_max_threads(); }
void fso_omp_set_threads(int n) { omp_set_num_threads(n); }

/* ------------------------------------------------------------------------------------------
 * Mean-flux rescaling.  ref: py_module.cpp:235-262 (get_mean_flux_scale):
 * Newton-Raphson on the scale s so that mean(exp(-s tau)) over the pixels with tau <= thresh equals
 * the desired mean flux.  Pinned by the reference's own known answers (tests/test_statistics.py:8-19,
 * restated in tests/test_oracle_stats.py).
 * ---------------------------------------------------------------------------------------- */
double fso_mean_flux_scale(const double *tau, double mean_flux_desired, long long nbins, double tol, double thresh,
                           int *iterations)
{
    double scale, newscale = 1;
    int it = 0;
    do {
        scale = newscale;
        double mean_flux = 0, tau_mean_flux = 0;
        long long nbins_used = 0;
        for (long long i = 0; i < nbins; i++) {
            if (tau[i] > thresh) continue;
            const double temp = exp(-scale * tau[i]);
            mean_flux += temp;
            tau_mean_flux += temp * tau[i];
            nbins_used++;
        }
        newscale = scale + (mean_flux - mean_flux_desired * nbins_used) / tau_mean_flux;
        if (newscale <= 0) newscale = 1e-10;
        ++it;
    } while (fabs(newscale - scale) > tol * newscale && it < 100000);
    if (iterations) *iterations = it;
    return newscale;
}
