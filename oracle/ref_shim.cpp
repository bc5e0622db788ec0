/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * extern "C" shim over the UNMODIFIED reference classes.  It is compiled by
 * oracle/build.py together with the reference's own absorption.cpp,
 * index_table.cpp, part_int.cpp and Faddeeva.cpp (read in place from
 * /root/reference/fake_spectra, never copied) into oracle/_ref/libfsref.so.
 * Every function here only constructs a reference object and calls one of its
 * public methods, so that Python (ctypes) can drive the real reference:
 *
 *   ParticleInterp            part_int.h:24-50
 *   IndexTable                index_table.h:10-52
 *   LineAbsorption            absorption.h:12-65
 *   SingleAbsorber / profile  singleabs.h:56-175
 *   Faddeeva::w               Faddeeva.h / Faddeeva.cpp:679
 */
#include <cstdint>
#include <cstring>
#include <complex>
#include <map>
#include <valarray>
#include <omp.h>

#include "part_int.h"
#include "singleabs.h"
#include "Faddeeva.h"

/* Defined (non-static) in absorption.cpp:76,109,129 but not declared in a header. */
double sph_quintic_kern_frac(double zlow, double zhigh, const double smooth, const double dr2, const double zrange);
double tophat_kern_frac(double zlow, double zhigh, const double smooth, const double dr2, const double zrange);
double arepo_kern_frac(double zlow, double zhigh, const double smooth, const double dr2, const double zrange);

extern "C" {

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { omp_set_num_threads(n); }

/* part_int.h:27 ctor + part_int.cpp:20 compute_tau.  tau accumulates into the caller's buffer. */
double ref_compute_tau(int nbins, double lambda, double gamma, double fosc, double amumass,
                       double box, double velfac, double atime, const double *cofm,
                       const int *axis, int nlos, int kernel, double tautail, double *tau,
                       const float *pos, const float *vel, const float *dens, const float *temp,
                       const float *h, long long npart)
{
    ParticleInterp pint(nbins, lambda, gamma, fosc, amumass, box, velfac, atime, cofm, axis, nlos, kernel, tautail);
    const double t0 = omp_get_wtime();
    pint.compute_tau(tau, pos, vel, dens, temp, h, npart);
    return omp_get_wtime() - t0;
}

/* part_int.cpp:53 compute_colden. */
double ref_compute_colden(int nbins, double lambda, double gamma, double fosc, double amumass,
                          double box, double velfac, double atime, const double *cofm,
                          const int *axis, int nlos, int kernel, double tautail, double *colden,
                          const float *pos, const float *dens, const float *h, long long npart)
{
    ParticleInterp pint(nbins, lambda, gamma, fosc, amumass, box, velfac, atime, cofm, axis, nlos, kernel, tautail);
    const double t0 = omp_get_wtime();
    pint.compute_colden(colden, pos, dens, h, npart);
    return omp_get_wtime() - t0;
}

/* index_table.cpp:130 get_near_particles.  Two-phase: counts[nlos] first, then the
 * flattened (ascending particle index per line) lists.  Returns the build time. */
double ref_near_particles(const double *cofm, const int *axis, int nlos, double box,
                          const float *pos, const float *h, long long npart,
                          long long *counts, int *part_out, double *dr2_out, long long cap)
{
    IndexTable tab(cofm, axis, nlos, box);
    const double t0 = omp_get_wtime();
    std::valarray< std::map<int, double> > near = tab.get_near_particles(pos, h, npart);
    const double dt = omp_get_wtime() - t0;
    long long off = 0;
    for (int i = 0; i < nlos; ++i) {
        counts[i] = (long long) near[i].size();
        for (std::map<int, double>::const_iterator it = near[i].begin(); it != near[i].end(); ++it) {
            if (part_out && off < cap) {
                part_out[off] = it->first;
                dr2_out[off] = it->second;
            }
            ++off;
        }
    }
    return dt;
}

/* index_table.cpp:117 get_near_lines for ONE particle; returns the number of lines found. */
int ref_get_near_lines(const double *cofm, const int *axis, int nlos, double box,
                       const float *pos3, float hh, int *line_out, double *dr2_out, int cap)
{
    IndexTable tab(cofm, axis, nlos, box);
    std::map<int, double> nearby = tab.get_near_lines(pos3, hh);
    int n = 0;
    for (std::map<int, double>::const_iterator it = nearby.begin(); it != nearby.end(); ++it, ++n) {
        if (n < cap) { line_out[n] = it->first; dr2_out[n] = it->second; }
    }
    return n;
}

/* The loop body of py_module.cpp:63-82 (Py_near_lines) driven through the same
 * IndexTable::get_near_lines; out receives ascending particle indices. */
long long ref_near_lines(double box, const float *pos, const float *h, long long npart,
                         const int *axis, const double *cofm, int nlos, int *out, long long cap)
{
    IndexTable tab(cofm, axis, nlos, box);
    unsigned char *flag = new unsigned char[npart > 0 ? npart : 1];
    #pragma omp parallel for
    for (long long i = 0; i < npart; i++) {
        std::map<int, double> nearby = tab.get_near_lines(&(pos[3*i]), h[i]);
        flag[i] = nearby.size() > 0;
    }
    long long n = 0;
    for (long long i = 0; i < npart; i++)
        if (flag[i]) { if (out && n < cap) out[n] = (int) i; ++n; }
    delete [] flag;
    return n;
}

/* index_table.cpp:152 assign_cells for one line (Voronoi).  Writes 2*Ncells floats. Returns Ncells. */
int ref_assign_cells(const double *cofm, const int *axis, int nlos, double box, int line,
                     const float *pos, const float *h, long long npart, float *arr_out, int cap)
{
    IndexTable tab(cofm, axis, nlos, box);
    std::valarray< std::map<int, double> > near = tab.get_near_particles(pos, h, npart);
    const int ncells = (int) near[line].size();
    float *arr = tab.assign_cells(line, near, pos);
    for (int i = 0; i < 2*ncells && i < cap; ++i) arr_out[i] = arr[i];
    delete [] arr;
    return ncells;
}

/* absorption.cpp:167 / :212 on a caller-provided row. */
void ref_add_colden_particle(double lambda, double gamma, double fosc, double amumass, double velfac,
                             double box, double atime, int kernel, double tautail,
                             double *colden, int nbins, double dr2, float dens, float ppos, float smooth)
{
    LineAbsorption la(lambda, gamma, fosc, amumass, velfac, box, atime, kernel, tautail);
    la.add_colden_particle(colden, nbins, dr2, dens, ppos, smooth);
}

void ref_add_tau_particle(double lambda, double gamma, double fosc, double amumass, double velfac,
                          double box, double atime, int kernel, double tautail,
                          double *tau, int nbins, double dr2, float dens, float ppos, float pvel,
                          float temp, float smooth)
{
    LineAbsorption la(lambda, gamma, fosc, amumass, velfac, box, atime, kernel, tautail);
    la.add_tau_particle(tau, nbins, dr2, dens, ppos, pvel, temp, smooth);
}

/* singleabs.h:17,31 */
double ref_sph_cubic_kernel(double q) { return sph_cubic_kernel(q); }
double ref_sph_quintic_kernel(double q) { return sph_quintic_kernel(q); }

/* absorption.cpp:53,76,109,129 selected by the kernel ids of singleabs.h:9-12 */
double ref_kern_frac(int kernel, double zlow, double zhigh, double smooth, double dr2, double zrange)
{
    if (kernel == SPH_CUBIC_SPLINE) return sph_cubic_kern_frac(zlow, zhigh, smooth, dr2, zrange);
    if (kernel == SPH_QUINTIC_SPLINE) return sph_quintic_kern_frac(zlow, zhigh, smooth, dr2, zrange);
    if (kernel == VORONOI_MESH) return arepo_kern_frac(zlow, zhigh, smooth, dr2, zrange);
    return tophat_kern_frac(zlow, zhigh, smooth, dr2, zrange);
}

/* singleabs.h:56 profile -> Faddeeva::w */
double ref_profile(double uu, double aa) { return profile(uu, aa); }

void ref_profile_many(const double *uu, const double *aa, double *out, long long n)
{
    #pragma omp parallel for
    for (long long i = 0; i < n; ++i) out[i] = profile(uu[i], aa[i]);
}

/* Faddeeva.cpp:679 full complex value (relerr = 0 -> machine precision) */
void ref_faddeeva_w(double x, double y, double *re, double *im)
{
    std::complex<double> r = Faddeeva::w(std::complex<double>(x, y), 0);
    *re = r.real(); *im = r.imag();
}

/* singleabs.h:81 ctor + :104 tau_kern_outer */
double ref_tau_kern_outer(double btherm, double vdr2, double vsmooth, double aa, int kernel,
                          double vlow, double vhigh)
{
    SingleAbsorber sa(btherm, vdr2, vsmooth, aa, kernel);
    return sa.tau_kern_outer(vlow, vhigh);
}

} /* extern "C" */
