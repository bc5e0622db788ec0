// Accumulation kernels: optical depth (replaces part_int.cpp:20-51 + absorption.cpp:212-279 +
// singleabs.h:63-175) and column density (part_int.cpp:53-84 + absorption.cpp:53-210).
//
// Work decomposition (both kernels): one warp per work item = (sightline, contiguous run of its
// candidate list).  Lanes are consecutive pixels of the current particle; per-particle constants
// are warp-uniform.  Each item owns its output row (the caller's row when a line is one item, a
// private scratch row otherwise), so accumulation needs no atomics and is bit-reproducible;
// scratch rows are summed in list order by k_reduce_rows.
#include <algorithm>

#include "fsb_common.cuh"
#include "fsb_scan.cuh"
#include "fsb_voigt.cuh"

namespace fsb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr double kSqrtPi = 1.77245385090551602729816748334;

// ---- SPH kernels: singleabs.h:17-42 ------------------------------------------------------------
__device__ __forceinline__ double cubic_kernel(double q)
{
    const double norm = 32. / 4 / kPi;
    if (q >= 1) return 0;
    if (q < 0.5) return norm * (1 - 6 * q * q + 6 * q * q * q);
    const double u = 1. - q;
    return norm * (2 * (u * u * u));
}

__device__ __forceinline__ double pow5(double u)
{
    const double u2 = u * u;
    return u2 * u2 * u;
}

__device__ __forceinline__ double quintic_kernel(double q)
{
    const double norm = 9. / 40 / kPi;
    if (q >= 1) return 0;
    if (q < (1. / 3)) return norm * 6 * (11 - 90 * q * q + 405 * q * q * q * q - 405 * q * q * q * q * q);
    if (q < (2. / 3)) return norm * (pow5(3. - 3 * q) - 6 * pow5(2. - 3 * q));
    return norm * (243 * pow5(1. - q));
}

template <int KERNEL>
__device__ __forceinline__ double sph_kernel(double q)
{
    if (KERNEL == FSB_KERNEL_CUBIC) return cubic_kernel(q);
    if (KERNEL == FSB_KERNEL_QUINTIC) return quintic_kernel(q);
    if (KERNEL == FSB_KERNEL_TOPHAT) return 3. / 4. / kPi;
    return 1.0;  // Voronoi: no kernel weight (singleabs.h:158-163)
}

// Line integral of the kernel over [zlow, zhigh] clipped to +-zrange: absorption.cpp:53-148.
template <int KERNEL>
__device__ __forceinline__ double kern_frac(double zlow, double zhigh, double smooth, double dr2, double zrange)
{
    zlow = fmax(zlow, -zrange);
    zhigh = fmin(zhigh, zrange);
    if (KERNEL == FSB_KERNEL_TOPHAT) return 3. / 4. / kPi * fmax(0., zhigh - zlow);
    if (KERNEL == FSB_KERNEL_VORONOI) return fmax(0., zhigh - zlow);
    if (zlow > zhigh) return 0;
    const double qlow = sqrt(dr2 + zlow * zlow) / smooth;
    double total = sph_kernel<KERNEL>(qlow) / 2.;
    const double deltaz = (zhigh - zlow) / kNGrid;
    #pragma unroll
    for (int i = 1; i < kNGrid; ++i) {
        const double zz = i * deltaz + zlow;
        const double q = sqrt(dr2 + zz * zz) / smooth;
        total += sph_kernel<KERNEL>(q);
    }
    const double qhigh = sqrt(dr2 + zhigh * zhigh) / smooth;
    total += sph_kernel<KERNEL>(qhigh) / 2.;
    return deltaz * total;
}

// ---- work items ---------------------------------------------------------------------------------
struct Items {
    const int32_t *item_start;  // [nlos+1] first item of each line (NULL: one item per line)
    int32_t seg_pairs;
};

// item -> (line, [kbeg, kend) in the pair arrays).  Returns false for items past the end.
__device__ __forceinline__ bool locate_item(const Items &it, const int64_t *__restrict__ offsets, int nlos, int item,
                                            int &line, int64_t &kbeg, int64_t &kend)
{
    if (it.item_start == nullptr) {
        if (item >= nlos) return false;
        line = item;
        kbeg = offsets[line];
        kend = offsets[line + 1];
        return kend > kbeg;
    }
    if (item >= it.item_start[nlos]) return false;
    int lo = 0, hi = nlos;  // last line with item_start <= item
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (it.item_start[mid] <= item) lo = mid;
        else hi = mid;
    }
    line = lo;
    const int seg = item - it.item_start[lo];
    kbeg = offsets[line] + (int64_t) seg * it.seg_pairs;
    kend = min(kbeg + (int64_t) it.seg_pairs, offsets[line + 1]);
    return kend > kbeg;
}

__global__ void k_items_per_line(const int64_t *__restrict__ offsets, int nlos, int seg_pairs, int32_t *__restrict__ nitems)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlos) return;
    const int64_t n = offsets[l + 1] - offsets[l];
    nitems[l] = (int32_t) ((n + seg_pairs - 1) / seg_pairs);
}

// out[w][line][j] += sum over the line's items (in list order) of scratch[w][item][j]
__global__ void k_reduce_rows(const int32_t *__restrict__ item_start, const double *__restrict__ scratch, int64_t scratch_stride,
                              double *__restrict__ out, int64_t out_stride, int nbins)
{
    const int line = blockIdx.x;
    const int w = blockIdx.z;
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= nbins) return;
    const int ibeg = item_start[line], iend = item_start[line + 1];
    if (iend == ibeg) return;
    double acc = 0;
    for (int it = ibeg; it < iend; ++it) acc += scratch[(int64_t) w * scratch_stride + (int64_t) it * nbins + j];
    out[(int64_t) w * out_stride + (int64_t) line * nbins + j] += acc;
}

__device__ __forceinline__ int wrap_bin(int z, int nbins)
{
    int j = z % nbins;
    if (j < 0) j += nbins;
    return j;
}

// ---- optical depth ------------------------------------------------------------------------------
//
// Kernel layout.  A CTA holds kTauWarps warps; the G(x) table of the fast Voigt path is staged once
// per CTA in shared memory, and every warp owns a slab of shared memory for the constants of the
// 32 particles it is currently working through.  Warps pull work items from a global counter
// (persistent CTAs), so long and short sightlines balance without host-side sorting.
//   per 32 particles : lane-parallel gather + per-particle setup (one particle per lane)
//   per particle     : warp-uniform constants read back from shared memory; lanes = pixels
//   per pixel        : 7-node kernel x Voigt quadrature (singleabs.h:143-167)

constexpr int kTauWarps = 4;
constexpr int kTauThreads = 32 * kTauWarps;

// Per-particle constants, one slot per lane of the owning warp (struct-of-arrays in shared memory).
enum PairField {
    F_VEL = 0,   // velfac*pos + pvel                                  absorption.cpp:234
    F_INVB,      // 1/btherm
    F_HALFB,     // btherm/2: sub-sampling threshold                    singleabs.h:110
    F_STEP,      // node spacing in units of btherm: (2 vhigh/8)/btherm
    F_XOFF,      // -vhigh/btherm
    F_CD,        // amp*dens/velfac * deltav
    F_Q,         // exp(-2 step^2): ratio of the Gaussian recurrence across nodes
    F_KW0,       // 7 kernel weights                                    singleabs.h:152-163
    F_PE0 = F_KW0 + 7,
    F_A0 = F_PE0 + 4,
    F_B0 = F_A0 + 4,
    F_XU2 = F_B0 + 3,
    F_Y,         // aa = voigt_fac/btherm
    F_ERFCX,     // erfcx(aa), exact mode only
    F_ZMAX,      // floor(vel/bintov) as a double
    F_MODE,      // 0 skip, 1 fast, 2 exact
    F_COUNT
};
constexpr int kPairSmemDoubles = F_COUNT * 32;

struct PairC {
    double vel, inv_b, half_b, step, xoff, cd, q;
    double kw[7];
    FastCoef fc;
    double erfcx_y;
    int zmax;
    int mode;
};

// 7-node sum at velocity offset vouter, fast Voigt.  Returns sum_i H(x_i) K_i (without deltav).
__device__ __forceinline__ double node_sum_fast(double vouter, const PairC &P, const double *__restrict__ tab, bool warp_needs_u)
{
    const double xb = fma(-vouter, P.inv_b, P.xoff);  // x of node i is xb + i*step
    double x[7], s[7], U[7];
    double smin = 1e300;
    #pragma unroll
    for (int i = 0; i < 7; ++i) {
        x[i] = fma((double) (i + 1), P.step, xb);
        s[i] = x[i] * x[i];
        smin = fmin(smin, s[i]);
    }
    #pragma unroll
    for (int i = 0; i < 7; ++i) U[i] = 0.0;
    if (warp_needs_u && smin < P.fc.xU2) {
        if (P.step <= 1.0) {
            // exp(-(x+step)^2) = exp(-x^2) exp(-(2x+step) step): two exponentials, then products.
            double u = exp(-s[0]);
            double r = exp(-fma(2.0, x[0], P.step) * P.step);
            U[0] = u;
            #pragma unroll
            for (int i = 1; i < 7; ++i) {
                u *= r;
                r *= P.q;
                U[i] = u;
            }
        } else {
            #pragma unroll
            for (int i = 0; i < 7; ++i) U[i] = s[i] < P.fc.xU2 ? exp(-s[i]) : 0.0;
        }
    }
    double total = 0;
    #pragma unroll
    for (int i = 0; i < 7; ++i) total = fma(voigt_fast_with_u(fabs(x[i]), s[i], U[i], P.fc, tab), P.kw[i], total);
    return total;
}

__device__ __noinline__ double node_sum_exact(double vouter, const PairC &P)
{
    const double xb = fma(-vouter, P.inv_b, P.xoff);
    double total = 0;
    #pragma unroll 1
    for (int i = 0; i < 7; ++i) total += voigt_exact(fma((double) (i + 1), P.step, xb), P.fc.y, P.erfcx_y) * P.kw[i];
    return total;
}

// Pixel average tau_kern_outer (singleabs.h:104-126) times amp*dens/velfac.
template <bool EXACT>
__device__ __forceinline__ double pixel_tau(double vlow, double bintov, const PairC &P, const double *__restrict__ tab,
                                            bool any_sub, bool warp_needs_u, unsigned &nvoigt)
{
    const double vhigh_px = __dadd_rn(vlow, bintov);
    const double width = vhigh_px - vlow;
    if (!any_sub || width < P.half_b) {
        nvoigt += 7;
        const double vmid = (vhigh_px + vlow) / 2.;
        return P.cd * (EXACT ? node_sum_exact(vmid, P) : node_sum_fast(vmid, P, tab, warp_needs_u));
    }
    const int npoints = (int) (2 * ceil(width / P.half_b / 2) + 1.);
    const double dv = width / (npoints - 1);
    double total = 0;
    for (int i = 0; i < npoints; ++i) {
        const double v = (i == 0) ? vlow : ((i == npoints - 1) ? vhigh_px : i * dv + vlow);
        const double wgt = (i == 0 || i == npoints - 1) ? 0.5 : 1.0;
        total += wgt * (EXACT ? node_sum_exact(v, P) : node_sum_fast(v, P, tab, true));
    }
    nvoigt += 7u * (unsigned) npoints;
    return P.cd * total / (npoints - 1);
}

// Outward pixel march of one particle (absorption.cpp:250-278): up from zmax, down from zmax-1,
// each direction adds pixels until (and including) the first with taulast < tautail.  While both
// directions are live each gets half the warp; afterwards all 32 lanes serve the remaining one.
template <bool EXACT>
__device__ __forceinline__ void march(const PairC &P, const double *__restrict__ tab, double *__restrict__ row, int nbins,
                                      double bintov, double tautail, int lane, unsigned &n_pix, unsigned &n_voigt,
                                      unsigned &n_iter)
{
    const int half = nbins / 2;
    int base_up = 0, base_dn = 0;
    bool live_up = half > 0, live_dn = half > 0;
    // pixel width >= btherm/2 anywhere?  (bintov is rounded differently per pixel by at most an ulp)
    const bool any_sub = !(bintov * (1 + 1e-12) < P.half_b);
    // furthest |x| at which the Gaussian still matters, in pixels from the particle (conservative)
    while (live_up || live_dn) {
        const bool both = live_up && live_dn;
        const int dir = both ? (lane >> 4) : (live_dn ? 1 : 0);
        const int sub = both ? (lane & 15) : lane;
        const int o = (dir ? base_dn : base_up) + sub;  // outward pixel index
        const bool mine = o < half;
        const int z = dir ? P.zmax - 1 - o : P.zmax + o;
        const double vlow = __dsub_rn(__dmul_rn((double) z, bintov), P.vel);
        // does any lane of this step still see the Gaussian core?  closest node distance in x:
        bool needs_u = false;
        if (!EXACT) {
            const double xc = fma(-(vlow + 0.5 * bintov), P.inv_b, P.xoff);  // node 0 (virtual) at pixel centre
            const double lo = xc + P.step, hi = fma(7.0, P.step, xc);       // node range [lo, hi]
            const double dmin = (lo > 0) ? lo : ((hi < 0) ? -hi : 0.0);
            needs_u = __any_sync(kFull, mine && dmin * dmin < P.fc.xU2);
        }
        double t = 0;
        if (mine) t = pixel_tau<EXACT>(vlow, bintov, P, tab, any_sub, needs_u, n_voigt);
        const unsigned stop = __ballot_sync(kFull, mine && (t < tautail));
        int first;  // sub-index of the first stopping lane of my direction
        if (both) {
            const unsigned sbits = dir ? (stop >> 16) : (stop & 0xffffu);
            first = sbits ? __ffs(sbits) - 1 : 16;
        } else {
            first = stop ? __ffs(stop) - 1 : 32;
        }
        if (mine && sub <= first) {
            row[wrap_bin(z, nbins)] += t;
            ++n_pix;
        }
        if (both) {
            base_up += 16;
            base_dn += 16;
            if ((stop & 0xffffu) || base_up >= half) live_up = false;
            if ((stop >> 16) || base_dn >= half) live_dn = false;
        } else if (dir) {
            base_dn += 32;
            if (stop || base_dn >= half) live_dn = false;
        } else {
            base_up += 32;
            if (stop || base_up >= half) live_up = false;
        }
        ++n_iter;
        __syncwarp();
    }
}

template <int KERNEL>
__global__ void __launch_bounds__(kTauThreads) k_tau(InterpConsts C, Items items, int n_items, int *__restrict__ next_item,
                                                     const int64_t *__restrict__ offsets, const int32_t *__restrict__ particle,
                                                     const double *__restrict__ dr2s, const int32_t *__restrict__ axis,
                                                     const float *__restrict__ pos, const float *__restrict__ vel,
                                                     const float *__restrict__ dens, const float *__restrict__ temp,
                                                     const float *__restrict__ hsml, const float *__restrict__ cells,
                                                     double *__restrict__ out, double *__restrict__ scratch,
                                                     unsigned long long *__restrict__ counters)
{
    extern __shared__ double smem[];
    double *tab = smem;                                            // [FSB_GTAB_SIZE]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *slab = smem + FSB_GTAB_SIZE + warp * kPairSmemDoubles;  // [F_COUNT][32]
    for (int i = threadIdx.x; i < FSB_GTAB_SIZE; i += kTauThreads) tab[i] = d_gtable[i];
    __syncthreads();

    const int nbins = C.nbins;
    const double bintov = C.bintov;
    const double sigma_a = C.line[0].sigma_a, voigt_fac = C.line[0].voigt_fac;
    const bool force_exact = C.voigt == FSB_VOIGT_EXACT;
    unsigned n_pix = 0, n_voigt = 0, n_iter = 0;
    unsigned long long n_pairs = 0;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1);
        item = __shfl_sync(kFull, item, 0);
        if (item >= n_items) break;
        int line;
        int64_t kbeg, kend;
        if (!locate_item(items, offsets, C.nlos, item, line, kbeg, kend)) continue;
        double *row = items.item_start ? scratch + (int64_t) item * nbins : out + (int64_t) line * nbins;
        const int ax = axis[line] - 1;
        n_pairs += (unsigned long long) (kend - kbeg);

        for (int64_t k0 = kbeg; k0 < kend; k0 += 32) {
            const int nb = (int) min((int64_t) 32, kend - k0);
            __syncwarp();
            if (lane < nb) {
                // one particle per lane: gather + constants (absorption.cpp:218-246, singleabs.h:81-90)
                const int64_t k = k0 + lane;
                const int64_t ip = particle[k];
                const float ppos = pos[3 * ip + ax], pvel = vel[3 * ip + ax];
                const float pdens = dens[ip], ptemp = temp[ip];
                double dr2;
                float smooth;
                if (KERNEL == FSB_KERNEL_VORONOI) {
                    dr2 = (double) cells[2 * k];
                    smooth = cells[2 * k + 1];
                } else {
                    dr2 = dr2s[k];
                    smooth = hsml[ip];
                }
                double pos1 = (double) ppos;
                int mode = 1;
                if (KERNEL == FSB_KERNEL_VORONOI) {
                    const double lim = 2 * C.vbox / C.velfac;
                    if (dr2 > lim || (double) smooth > lim) mode = 0;
                    pos1 = __dmul_rn(__dadd_rn(dr2, (double) smooth), 0.5);
                } else {
                    if (__dsub_rn((double) __fmul_rn(smooth, smooth), dr2) <= 0) mode = 0;
                }
                const double btherm = C.bfac * sqrt((double) ptemp);
                const double velp = __dadd_rn(__dmul_rn(C.velfac, pos1), (double) pvel);
                double vdr2 = C.velfac * dr2;
                if (KERNEL != FSB_KERNEL_VORONOI) vdr2 *= C.velfac;
                const double vsmooth = C.velfac * (double) smooth;
                const double inv_b = 1.0 / btherm;
                const double aa = voigt_fac * inv_b;
                const double amp = sigma_a / kSqrtPi * (kLight / 1e5 * inv_b);
                double vhigh = (vsmooth * vsmooth > vdr2) ? sqrt(vsmooth * vsmooth - vdr2) : 0;
                if (KERNEL == FSB_KERNEL_VORONOI) vhigh = (vdr2 > 0 && vsmooth > 0) ? (vsmooth - vdr2) / 2. : 0;
                const double deltav = 2. * vhigh / kNGrid;
                const double step = deltav * inv_b;
                if (mode && (force_exact || !fast_domain(aa))) mode = 2;
                slab[F_VEL * 32 + lane] = velp;
                slab[F_INVB * 32 + lane] = inv_b;
                slab[F_HALFB * 32 + lane] = btherm / 2.;
                slab[F_STEP * 32 + lane] = step;
                slab[F_XOFF * 32 + lane] = -vhigh * inv_b;
                slab[F_CD * 32 + lane] = amp * (double) pdens / C.velfac * deltav;
                slab[F_Q * 32 + lane] = exp(-2.0 * step * step);
                #pragma unroll
                for (int i = 1; i < kNGrid; ++i) {
                    const double vv = i * deltav - vhigh;
                    slab[(F_KW0 + i - 1) * 32 + lane] = sph_kernel<KERNEL>(sqrt(vdr2 + vv * vv) / vsmooth);
                }
                FastCoef fc;
                fast_coefs(aa, fc);
                #pragma unroll
                for (int i = 0; i < 4; ++i) slab[(F_PE0 + i) * 32 + lane] = fc.pe[i];
                #pragma unroll
                for (int i = 0; i < 4; ++i) slab[(F_A0 + i) * 32 + lane] = fc.a[i];
                #pragma unroll
                for (int i = 0; i < 3; ++i) slab[(F_B0 + i) * 32 + lane] = fc.b[i];
                slab[F_XU2 * 32 + lane] = fc.xU2;
                slab[F_Y * 32 + lane] = aa;
                slab[F_ERFCX * 32 + lane] = mode == 2 ? erfcx(aa) : 0.0;
                slab[F_ZMAX * 32 + lane] = floor(velp / bintov);
                slab[F_MODE * 32 + lane] = (double) mode;
            }
            __syncwarp();
            for (int b = 0; b < nb; ++b) {
                PairC P;
                P.mode = (int) slab[F_MODE * 32 + b];
                if (P.mode == 0) continue;
                P.vel = slab[F_VEL * 32 + b];
                P.inv_b = slab[F_INVB * 32 + b];
                P.half_b = slab[F_HALFB * 32 + b];
                P.step = slab[F_STEP * 32 + b];
                P.xoff = slab[F_XOFF * 32 + b];
                P.cd = slab[F_CD * 32 + b];
                P.q = slab[F_Q * 32 + b];
                #pragma unroll
                for (int i = 0; i < 7; ++i) P.kw[i] = slab[(F_KW0 + i) * 32 + b];
                #pragma unroll
                for (int i = 0; i < 4; ++i) P.fc.pe[i] = slab[(F_PE0 + i) * 32 + b];
                #pragma unroll
                for (int i = 0; i < 4; ++i) P.fc.a[i] = slab[(F_A0 + i) * 32 + b];
                #pragma unroll
                for (int i = 0; i < 3; ++i) P.fc.b[i] = slab[(F_B0 + i) * 32 + b];
                P.fc.xU2 = slab[F_XU2 * 32 + b];
                P.fc.y = slab[F_Y * 32 + b];
                P.erfcx_y = slab[F_ERFCX * 32 + b];
                P.zmax = (int) slab[F_ZMAX * 32 + b];
                if (P.mode == 2) march<true>(P, tab, row, nbins, bintov, C.tautail, lane, n_pix, n_voigt, n_iter);
                else march<false>(P, tab, row, nbins, bintov, C.tautail, lane, n_pix, n_voigt, n_iter);
            }
        }
    }
    if (counters) {
        unsigned long long pix = n_pix, vg = n_voigt;
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            pix += __shfl_down_sync(kFull, pix, d);
            vg += __shfl_down_sync(kFull, vg, d);
        }
        if (lane == 0) {
            atomicAdd(&counters[0], n_pairs);
            atomicAdd(&counters[1], pix);
            atomicAdd(&counters[2], vg);
            atomicAdd(&counters[3], 32ull * n_iter);
        }
    }
}

// ---- column density -----------------------------------------------------------------------------
template <int KERNEL>
__global__ void __launch_bounds__(32) k_colden(InterpConsts C, Items items, const int64_t *__restrict__ offsets,
                                               const int32_t *__restrict__ particle, const double *__restrict__ dr2s,
                                               const int32_t *__restrict__ axis, const float *__restrict__ pos,
                                               const float *__restrict__ dens, int64_t dens_stride,
                                               const float *__restrict__ hsml, const float *__restrict__ cells,
                                               double *__restrict__ out, int64_t out_stride, double *__restrict__ scratch,
                                               int64_t scratch_stride, unsigned long long *__restrict__ counters)
{
    const int lane = threadIdx.x;
    int line;
    int64_t kbeg, kend;
    if (!locate_item(items, offsets, C.nlos, blockIdx.x, line, kbeg, kend)) return;
    double *row = items.item_start ? scratch + (int64_t) blockIdx.x * C.nbins : out + (int64_t) line * C.nbins;
    const int64_t wstride = items.item_start ? scratch_stride : out_stride;
    const int ax = axis[line] - 1;
    const int nbins = C.nbins;
    const int nw = C.nlines;
    const int chunk = nbins < 32 ? nbins : 32;  // keep the pixels of one step distinct modulo nbins
    const double boxtokpc = C.boxtokpc;
    unsigned n_pix = 0;

    for (int64_t k = kbeg; k < kend; ++k) {
        const int64_t ip = particle[k];
        const float ppos = pos[3 * ip + ax];
        double dr2;
        float smooth;
        if (KERNEL == FSB_KERNEL_VORONOI) {
            dr2 = (double) cells[2 * k];
            smooth = cells[2 * k + 1];
        } else {
            dr2 = dr2s[k];
            smooth = hsml[ip];
        }
        // absorption.cpp:167-193
        double pos1 = (double) ppos;
        double zrange;
        if (KERNEL == FSB_KERNEL_VORONOI) {
            const double lim = 2 * C.vbox / C.velfac;
            if (dr2 > lim || (double) smooth > lim) continue;
            pos1 = __dmul_rn(__dadd_rn(dr2, (double) smooth), 0.5);
            zrange = __dmul_rn(__dsub_rn((double) smooth, dr2), 0.5);
        } else {
            const double arg = __dsub_rn((double) __fmul_rn(smooth, smooth), dr2);
            if (arg <= 0) continue;
            zrange = sqrt(arg);
        }
        const int zlow = (int) floor(__ddiv_rn(__dsub_rn(pos1, zrange), boxtokpc));
        const int zhigh = (int) ceil(__ddiv_rn(__dadd_rn(pos1, zrange), boxtokpc));
        for (int zb = zlow; zb <= zhigh; zb += chunk) {
            const int z = zb + lane;
            if (lane < chunk && z <= zhigh) {
                const double plow = __dsub_rn(__dmul_rn(boxtokpc, (double) z), pos1);
                const double frac = kern_frac<KERNEL>(plow, __dadd_rn(plow, boxtokpc), (double) smooth, dr2, zrange);
                const int j = wrap_bin(z, nbins);
                for (int w = 0; w < nw; ++w) row[(int64_t) w * wstride + j] += (double) dens[(int64_t) w * dens_stride + ip] * frac;
                ++n_pix;
            }
            __syncwarp();
        }
    }
    if (counters) {
        unsigned long long pix = n_pix;
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1) pix += __shfl_down_sync(kFull, pix, d);
        if (lane == 0) {
            atomicAdd(&counters[0], (unsigned long long) (kend - kbeg));
            atomicAdd(&counters[1], pix);
        }
    }
}

__global__ void k_voigt_profile(const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ out,
                                int64_t n, int voigt)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double yy = y[i];
    if (voigt == FSB_VOIGT_FAST && fast_domain(yy)) {
        FastCoef fc;
        fast_coefs(yy, fc);
        out[i] = voigt_fast(x[i], fc, d_gtable);
    } else {
        out[i] = voigt_exact(x[i], yy, erfcx(yy));
    }
}

// Work-item table for one launch.
struct ItemPlan {
    Scratch item_start, nitems, scratch_rows;
    Items items;
    int64_t n_items = 0;  // upper bound on the number of items (= grid size)
    bool segmented = false;
};

int plan_items(const fsb_index *idx, int seg_pairs_req, int nbins, int nrows_per_item, cudaStream_t stream, ItemPlan &plan)
{
    const int64_t target_items = 8192;  // ~ 148 SMs x 16+ resident warps, a few waves
    int seg = seg_pairs_req;
    if (seg <= 0) {
        if (idx->nlos >= target_items / 2 || idx->npairs == 0) seg = 0;  // enough sightlines: one item per line
        else seg = (int) std::max<int64_t>(32, (idx->npairs + target_items - 1) / target_items);
    }
    if (seg <= 0 || seg >= idx->max_list) {
        plan.items.item_start = nullptr;
        plan.items.seg_pairs = 0;
        plan.n_items = idx->nlos;
        plan.segmented = false;
        return FSB_OK;
    }
    plan.segmented = true;
    plan.items.seg_pairs = seg;
    plan.n_items = (int64_t) idx->nlos + idx->npairs / seg;
    const size_t nl = (size_t) std::max(idx->nlos, 1);
    FSB_TRY(plan.nitems.alloc(sizeof(int32_t) * (nl + 1), stream));
    FSB_TRY(plan.item_start.alloc(sizeof(int32_t) * (nl + 1), stream));
    count_launch(); k_items_per_line<<<(idx->nlos + 255) / 256, 256, 0, stream>>>(idx->offsets, idx->nlos, seg, plan.nitems.as<int32_t>());
    count_launch(); k_scan_single<int32_t, int32_t><<<1, 1024, 0, stream>>>(plan.nitems.as<int32_t>(), plan.item_start.as<int32_t>(), idx->nlos, nullptr);
    FSB_CUDA_TRY(cudaGetLastError());
    plan.items.item_start = plan.item_start.as<int32_t>();
    const size_t bytes = sizeof(double) * (size_t) plan.n_items * (size_t) nbins * (size_t) nrows_per_item;
    FSB_TRY(plan.scratch_rows.alloc(bytes, stream));
    FSB_CUDA_TRY(cudaMemsetAsync(plan.scratch_rows.ptr, 0, bytes, stream));
    return FSB_OK;
}

}  // namespace

int launch_tau(const fsb_index *idx, const InterpConsts &c, const float *pos, const float *vel, const float *dens,
               const float *temp, const float *h, const float *cells, double *out, fsb_counters *counters, int precision,
               cudaStream_t stream)
{
    (void) precision;
    if (idx->nlos == 0 || idx->npairs == 0) return FSB_OK;
    ItemPlan plan;
    FSB_TRY(plan_items(idx, c.seg_pairs, c.nbins, 1, stream, plan));
    Scratch next_item;
    FSB_TRY(next_item.alloc(sizeof(int), stream));
    FSB_CUDA_TRY(cudaMemsetAsync(next_item.ptr, 0, sizeof(int), stream));
    int dev = 0, sms = 0;
    FSB_CUDA_TRY(cudaGetDevice(&dev));
    FSB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = sizeof(double) * (size_t) (FSB_GTAB_SIZE + kTauWarps * kPairSmemDoubles);
    double *scratch = plan.scratch_rows.as<double>();
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(counters);
    const int n_items = (int) plan.n_items;
#define FSB_LAUNCH_TAU(K)                                                                                              \
    do {                                                                                                               \
        int per_sm = 0;                                                                                                \
        FSB_CUDA_TRY(cudaFuncSetAttribute(k_tau<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));         \
        FSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tau<K>, kTauThreads, smem));             \
        const int grid = std::max(1, std::min((n_items + kTauWarps - 1) / kTauWarps, sms * std::max(per_sm, 1)));      \
        count_launch(); k_tau<K><<<grid, kTauThreads, smem, stream>>>(c, plan.items, n_items, next_item.as<int>(), idx->offsets,       \
                                                      idx->particle, idx->dr2, idx->axis, pos, vel, dens, temp, h,    \
                                                      cells, out, scratch, ctr);                                       \
    } while (0)
    switch (c.kernel) {
    case FSB_KERNEL_TOPHAT: FSB_LAUNCH_TAU(FSB_KERNEL_TOPHAT); break;
    case FSB_KERNEL_CUBIC: FSB_LAUNCH_TAU(FSB_KERNEL_CUBIC); break;
    case FSB_KERNEL_VORONOI: FSB_LAUNCH_TAU(FSB_KERNEL_VORONOI); break;
    case FSB_KERNEL_QUINTIC: FSB_LAUNCH_TAU(FSB_KERNEL_QUINTIC); break;
    default: set_error("unknown kernel id %d", c.kernel); return FSB_EINVAL;
    }
#undef FSB_LAUNCH_TAU
    FSB_CUDA_TRY(cudaGetLastError());
    if (plan.segmented) {
        dim3 g(idx->nlos, (c.nbins + 255) / 256, 1);
        count_launch(); k_reduce_rows<<<g, 256, 0, stream>>>(plan.items.item_start, scratch, 0, out, 0, c.nbins);
        FSB_CUDA_TRY(cudaGetLastError());
    }
    return FSB_OK;
}

int launch_colden(const fsb_index *idx, const InterpConsts &c, const float *pos, const float *dens, int64_t dens_stride,
                  const float *h, const float *cells, double *out, fsb_counters *counters, cudaStream_t stream)
{
    if (idx->nlos == 0 || idx->npairs == 0) return FSB_OK;
    ItemPlan plan;
    FSB_TRY(plan_items(idx, c.seg_pairs, c.nbins, c.nlines, stream, plan));
    const unsigned grid = (unsigned) plan.n_items;
    double *scratch = plan.scratch_rows.as<double>();
    const int64_t out_stride = (int64_t) idx->nlos * c.nbins;
    const int64_t scratch_stride = plan.n_items * c.nbins;
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(counters);
#define FSB_LAUNCH_COLDEN(K)                                                                                          \
    count_launch(); k_colden<K><<<grid, 32, 0, stream>>>(c, plan.items, idx->offsets, idx->particle, idx->dr2, idx->axis, pos, dens,  \
                                         dens_stride, h, cells, out, out_stride, scratch, scratch_stride, ctr)
    switch (c.kernel) {
    case FSB_KERNEL_TOPHAT: FSB_LAUNCH_COLDEN(FSB_KERNEL_TOPHAT); break;
    case FSB_KERNEL_CUBIC: FSB_LAUNCH_COLDEN(FSB_KERNEL_CUBIC); break;
    case FSB_KERNEL_VORONOI: FSB_LAUNCH_COLDEN(FSB_KERNEL_VORONOI); break;
    case FSB_KERNEL_QUINTIC: FSB_LAUNCH_COLDEN(FSB_KERNEL_QUINTIC); break;
    default: set_error("unknown kernel id %d", c.kernel); return FSB_EINVAL;
    }
#undef FSB_LAUNCH_COLDEN
    FSB_CUDA_TRY(cudaGetLastError());
    if (plan.segmented) {
        dim3 g(idx->nlos, (c.nbins + 255) / 256, c.nlines);
        count_launch(); k_reduce_rows<<<g, 256, 0, stream>>>(plan.items.item_start, scratch, scratch_stride, out, out_stride, c.nbins);
        FSB_CUDA_TRY(cudaGetLastError());
    }
    return FSB_OK;
}

int launch_voigt(const double *x, const double *y, double *out, int64_t n, int voigt, cudaStream_t stream)
{
    if (n <= 0) return FSB_OK;
    count_launch(); k_voigt_profile<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(x, y, out, n, voigt);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

}  // namespace fsb
