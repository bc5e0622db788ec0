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

// Warp-uniform per-particle state: singleabs.h:81-90 (SingleAbsorber) + absorption.cpp:218-246.
struct Absorber {
    double vel;       // velfac*pos + pvel
    double inv_b;     // 1/btherm
    double half_b;    // btherm/2 (sub-sampling threshold, singleabs.h:110)
    double aa;        // voigt_fac/btherm
    double erfcx_aa;  // erfcx(aa)
    double coef;      // amp*dens/velfac
    double vhigh;     // kernel support in velocity units
    double deltav;    // 2*vhigh/8
    double kw[7];     // kernel weight of the 7 interior quadrature nodes
};

template <int VOIGT>
__device__ __forceinline__ double voigt_eval(double T0, const Absorber &A)
{
    return voigt_exact(T0, A.aa, A.erfcx_aa);
}

// tau at one velocity offset: 7-node kernel x Voigt sum, singleabs.h:143-167.
template <int VOIGT>
__device__ __forceinline__ double tau_inner(double vouter, const Absorber &A)
{
    double total = 0;
    #pragma unroll 1
    for (int i = 1; i < kNGrid; ++i) {
        const double vv = i * A.deltav - A.vhigh;
        const double T0 = (vv - vouter) * A.inv_b;
        total += voigt_eval<VOIGT>(T0, A) * A.kw[i - 1];
    }
    return A.deltav * total;
}

// pixel average: singleabs.h:104-126.  nvoigt counts profile evaluations.
template <int VOIGT>
__device__ __forceinline__ double tau_outer(double vlow, double vhigh_px, const Absorber &A, unsigned &nvoigt)
{
    const double width = vhigh_px - vlow;
    if (width < A.half_b) {
        nvoigt += 7;
        return tau_inner<VOIGT>((vhigh_px + vlow) / 2., A);
    }
    const int npoints = (int) (2 * ceil(width / A.half_b / 2) + 1.);
    double total = tau_inner<VOIGT>(vlow, A) / 2.;
    const double dv = width / (npoints - 1);
    for (int i = 1; i < npoints - 1; ++i) total += tau_inner<VOIGT>(i * dv + vlow, A);
    total += tau_inner<VOIGT>(vhigh_px, A) / 2.;
    nvoigt += 7u * (unsigned) npoints;
    return total / (npoints - 1);
}

template <int KERNEL, int VOIGT>
__global__ void __launch_bounds__(32) k_tau(InterpConsts C, Items items, const int64_t *__restrict__ offsets,
                                            const int32_t *__restrict__ particle, const double *__restrict__ dr2s,
                                            const int32_t *__restrict__ axis, const float *__restrict__ pos,
                                            const float *__restrict__ vel, const float *__restrict__ dens,
                                            const float *__restrict__ temp, const float *__restrict__ hsml,
                                            const float *__restrict__ cells, double *__restrict__ out,
                                            double *__restrict__ scratch, unsigned long long *__restrict__ counters)
{
    const int lane = threadIdx.x;
    int line;
    int64_t kbeg, kend;
    if (!locate_item(items, offsets, C.nlos, blockIdx.x, line, kbeg, kend)) return;
    double *row = items.item_start ? scratch + (int64_t) blockIdx.x * C.nbins : out + (int64_t) line * C.nbins;
    const int ax = axis[line] - 1;
    const int nbins = C.nbins, half = nbins / 2;
    const double bintov = C.bintov;
    const double sigma_a = C.line[0].sigma_a, voigt_fac = C.line[0].voigt_fac;
    unsigned n_pix = 0, n_voigt = 0, n_lanes = 0;

    for (int64_t k0 = kbeg; k0 < kend; k0 += 32) {
        // lane-parallel gather of up to 32 particles, then warp-uniform processing one at a time
        const int nb = (int) min((int64_t) 32, kend - k0);
        double my_dr2 = 0;
        float my_pos = 0, my_vel = 0, my_dens = 0, my_temp = 1, my_h = 0;
        if (lane < nb) {
            const int64_t k = k0 + lane;
            const int64_t ip = particle[k];
            my_pos = pos[3 * ip + ax];
            my_vel = vel[3 * ip + ax];
            my_dens = dens[ip];
            my_temp = temp[ip];
            if (KERNEL == FSB_KERNEL_VORONOI) {
                my_dr2 = (double) cells[2 * k];
                my_h = cells[2 * k + 1];
            } else {
                my_dr2 = dr2s[k];
                my_h = hsml[ip];
            }
        }
        for (int b = 0; b < nb; ++b) {
            const double dr2 = __shfl_sync(kFull, my_dr2, b);
            const float ppos = __shfl_sync(kFull, my_pos, b);
            const float pvel = __shfl_sync(kFull, my_vel, b);
            const float pdens = __shfl_sync(kFull, my_dens, b);
            const float ptemp = __shfl_sync(kFull, my_temp, b);
            const float smooth = __shfl_sync(kFull, my_h, b);

            // absorption.cpp:218-246
            double pos1 = (double) ppos;
            const double btherm = C.bfac * sqrt((double) ptemp);
            if (KERNEL == FSB_KERNEL_VORONOI) {
                const double lim = 2 * C.vbox / C.velfac;
                if (dr2 > lim || (double) smooth > lim) continue;
                pos1 = __dmul_rn(__dadd_rn(dr2, (double) smooth), 0.5);
            } else {
                if (__dsub_rn((double) __fmul_rn(smooth, smooth), dr2) <= 0) continue;
            }
            Absorber A;
            A.vel = __dadd_rn(__dmul_rn(C.velfac, pos1), (double) pvel);
            double vdr2 = C.velfac * dr2;
            if (KERNEL != FSB_KERNEL_VORONOI) vdr2 *= C.velfac;
            const double vsmooth = C.velfac * (double) smooth;
            A.inv_b = 1.0 / btherm;
            A.half_b = btherm / 2.;
            A.aa = voigt_fac / btherm;
            A.erfcx_aa = erfcx(A.aa);
            const double amp = sigma_a / kSqrtPi * (kLight / 1e5 / btherm);
            A.coef = amp * (double) pdens / C.velfac;
            // singleabs.h:83-89
            A.vhigh = (vsmooth * vsmooth > vdr2) ? sqrt(vsmooth * vsmooth - vdr2) : 0;
            if (KERNEL == FSB_KERNEL_VORONOI) A.vhigh = (vdr2 > 0 && vsmooth > 0) ? (vsmooth - vdr2) / 2. : 0;
            A.deltav = 2. * A.vhigh / kNGrid;
            #pragma unroll
            for (int i = 1; i < kNGrid; ++i) {
                const double vv = i * A.deltav - A.vhigh;
                A.kw[i - 1] = sph_kernel<KERNEL>(sqrt(vdr2 + vv * vv) / vsmooth);
            }
            const int zmax = (int) floor(A.vel / bintov);

            // Outward pixel march, absorption.cpp:250-278: up from zmax, down from zmax-1, each
            // direction adds pixels until (and including) the first with taulast < tautail.
            // While both directions are live each gets half the warp; afterwards all 32 lanes
            // serve the remaining one.
            int base_up = 0, base_dn = 0;
            bool live_up = half > 0, live_dn = half > 0;
            while (live_up || live_dn) {
                const bool both = live_up && live_dn;
                const int dir = both ? (lane >> 4) : (live_dn ? 1 : 0);
                const int sub = both ? (lane & 15) : lane;
                const int o = (dir ? base_dn : base_up) + sub;  // outward pixel index
                const bool mine = o < half;
                const int z = dir ? zmax - 1 - o : zmax + o;
                double t = 0;
                if (mine) {
                    const double vlow = __dsub_rn(__dmul_rn((double) z, bintov), A.vel);
                    t = A.coef * tau_outer<VOIGT>(vlow, __dadd_rn(vlow, bintov), A, n_voigt);
                }
                const unsigned stop = __ballot_sync(kFull, mine && (t < C.tautail));
                int first;  // sub-index of the first stopping lane of my direction
                if (both) {
                    const unsigned s = dir ? (stop >> 16) : (stop & 0xffffu);
                    first = s ? __ffs(s) - 1 : 16;
                } else {
                    first = stop ? __ffs(stop) - 1 : 32;
                }
                if (mine && sub <= first) {
                    const int j = wrap_bin(z, nbins);
                    row[j] += t;
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
                ++n_lanes;
                __syncwarp();
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
            atomicAdd(&counters[0], (unsigned long long) (kend - kbeg));
            atomicAdd(&counters[1], pix);
            atomicAdd(&counters[2], vg);
            atomicAdd(&counters[3], 32ull * n_lanes);
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
    out[i] = voigt_exact(x[i], yy, erfcx(yy));
    (void) voigt;
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
    const unsigned grid = (unsigned) plan.n_items;
    double *scratch = plan.scratch_rows.as<double>();
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(counters);
#define FSB_LAUNCH_TAU(K)                                                                                          \
    count_launch(); k_tau<K, FSB_VOIGT_EXACT><<<grid, 32, 0, stream>>>(c, plan.items, idx->offsets, idx->particle, idx->dr2, idx->axis, \
                                                        pos, vel, dens, temp, h, cells, out, scratch, ctr)
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
