// Snapshot fields -> inputs of the interpolation, on the device.
//
// The reference prepares the arrays of _Particle_Interpolate with O(N) numpy on the host, per segment and per
// species: select the particles near the sightlines, then gather positions, peculiar velocities (x sqrt(a),
// abstractsnapshot.py:114-119), smoothing lengths (SmoothingLength / 2 or Volume^(1/3), :253-282), temperatures from
// the internal energy (:121-154, floor of 1 K: spectra.py:585-589), the hydrogen number density from the code density
// (gas_properties.py:104-110), the neutral fraction with the Rahmati et al. (2013) value above the star-formation
// threshold (:116-146, eq. A8 at 1e4 K), the element's mass fraction and finally the species density
// den * rscale * mass_frac [* x_HI] / amu (spectra.py:593-615).  Here the raw fields of a segment stay in HBM and one
// gather kernel produces the five float32 arrays for the selected particles (fsb_near_lines gives the selection).
// Metal ions (spectra.py:598-611,637-664): particles without mass in the element are dropped first
// (fsb_prepare_select = _filter_particles), then the ion fraction comes from a Cloudy table of log10 fractions on a
// regular (log10 nH, log10 T) grid at the snapshot's redshift, interpolated like scipy.ndimage.map_coordinates does in
// CloudyTable.ion (convert_cloudy.py:167-200): cubic B-spline, mode "nearest".  The host filters the table once
// (scipy's own prefilter on the edge-padded table, cloudy.py); the kernel evaluates the 4 x 4 tensor-product spline.
//
// Arithmetic: every product the reference forms in float32 is formed in float32 here too (operands rounded like
// numpy rounds them); the transcendental part of the Rahmati formula is evaluated in double and rounded once, so
// results agree with the host path to float32 rounding (a few 1e-7 relative), not bit for bit.
#include <math.h>

#include "fsb_common.cuh"
#include "fsb_scan.cuh"

namespace fsb {

namespace {

// Rahmati et al. 2013, eqs. 13, 14, A3, A6, A8 at temperature T (the reference evaluates them at 1e4 K).
__device__ __forceinline__ double rahmati_neutral_fraction(double nH, double T, double gray_opac, double gamma_uvb, double f_bar)
{
    const double T4 = T / 1e4, G12 = gamma_uvb / 1e-12;
    const double nSSh = 6.73e-3 * pow(gray_opac / 2.49e-18, -2. / 3) * pow(T4, 0.17) * pow(G12, 2. / 3) * pow(f_bar / 0.17, -1. / 3);
    const double ratio = nH / nSSh;
    const double photo = (0.98 * pow(1 + pow(ratio, 1.64), -2.28) + 0.02 * pow(1 + ratio, -0.84)) * gamma_uvb;
    const double lamb = 315614. / T;
    const double alpha_A = 1.269e-13 * pow(lamb, 1.503) / pow(1 + pow(lamb / 0.522, 0.47), 1.923);
    const double lambda_T = 1.17e-10 * sqrt(T) * exp(-157809. / T) / (1 + sqrt(T / 1e5));
    const double A = alpha_A + lambda_T;
    const double B = 2 * alpha_A + photo / nH + lambda_T;
    return (B - sqrt(B * B - 4 * A * alpha_A)) / (2 * A);
}

// Cubic B-spline weights of the four coefficients around a coordinate with fractional part t.
__device__ __forceinline__ void bspline3(double t, double w[4])
{
    const double u = 1 - t;
    w[0] = u * u * u / 6;
    w[1] = (3 * t * t * t - 6 * t * t + 4) / 6;
    w[2] = (3 * u * u * u - 6 * u * u + 4) / 6;
    w[3] = t * t * t / 6;
}

// CloudyTable.ion for one particle: nH and T float32 as the reference holds them; returns the ion fraction.
__device__ __forceinline__ float ion_fraction(const fsb_ion_table &tb, float nH, float T)
{
    // spectra.py:649-663: clip to the table's bounds (assigned into float32 arrays), convert_cloudy.py:183: rho *= 0.774132
    T = fminf(fmaxf(T, tb.temp_lo), tb.temp_hi);
    nH = fminf(fmaxf(nH, tb.dens_lo), tb.dens_hi);
    nH = __fmul_rn(nH, tb.rho_factor);
    // grid coordinates: float32 log10, then double arithmetic (numpy: float32 array with float64 scalars).  numpy's
    // float32 log10 is off by an ulp for half of its arguments; the correctly rounded value is used here.
    double c0 = ((double) (float) log10((double) nH) - tb.dens0) * (double) (tb.nd - 1) / tb.dens_span;
    double c1 = ((double) (float) log10((double) T) - tb.temp0) * (double) (tb.nt - 1) / tb.temp_span;
    // mode "nearest": the table is edge-padded by `pad` cells and coordinates are used as they are inside the pad
    const int pad = tb.pad, n0 = tb.nd + 2 * pad, n1 = tb.nt + 2 * pad;
    c0 = fmin(fmax(c0, (double) -pad), (double) (tb.nd - 1 + pad)) + pad;
    c1 = fmin(fmax(c1, (double) -pad), (double) (tb.nt - 1 + pad)) + pad;
    const double f0 = floor(c0), f1 = floor(c1);
    double w0[4], w1[4];
    bspline3(c0 - f0, w0);
    bspline3(c1 - f1, w1);
    const int i0 = (int) f0 - 1, i1 = (int) f1 - 1;
    double acc = 0;
    #pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int r = min(max(i0 + a, 0), n0 - 1);
        double row = 0;
        #pragma unroll
        for (int b = 0; b < 4; ++b) row += w1[b] * tb.coef[(int64_t) r * n1 + min(max(i1 + b, 0), n1 - 1)];
        acc += w0[a] * row;
    }
    return (float) pow(10.0, acc);  // np.float32(10**ions)
}

// get_smooth_length (abstractsnapshot.py:253-282) for all particles of a segment
__global__ void __launch_bounds__(256) k_smooth_length(const float *__restrict__ a, const float *__restrict__ b, int64_t n, int mode,
                                                       float *__restrict__ hh)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = mode == 2 ? __fdiv_rn(a[i], b[i]) : a[i];
    hh[i] = mode == 0 ? __fmul_rn(v, 0.5f) : (float) pow((double) v, (double) (1.0f / 3.0f));
}

// _filter_particles (spectra.py:600): flag the selected particles whose element density (den * rscale) * mass_frac is > 0
__global__ void __launch_bounds__(1024) k_select_flags(fsb_prep cfg, const int32_t *__restrict__ index, int64_t m,
                                                       const float *__restrict__ density, const float *__restrict__ mass_frac,
                                                       int64_t mass_frac_stride, uint8_t *__restrict__ flag)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t p = index ? (int64_t) index[i] : i;
    float mf = mass_frac ? mass_frac[p * mass_frac_stride] : cfg.mass_frac_const;
    if (mf <= 0) mf = 0;
    flag[i] = __fmul_rn(__fmul_rn(__fmul_rn(density[p], cfg.dens_conv), cfg.rscale), mf) > 0;
}

__global__ void __launch_bounds__(256) k_prepare(fsb_prep cfg, fsb_ion_table ion, const int32_t *__restrict__ index, int64_t m,
                                                 const float *__restrict__ position, const float *__restrict__ velocity,
                                                 const float *__restrict__ density, const float *__restrict__ ienergy,
                                                 const float *__restrict__ nelec, const float *__restrict__ nh0,
                                                 const float *__restrict__ smoothing, const float *__restrict__ mass_frac,
                                                 int64_t mass_frac_stride, float *__restrict__ pos, float *__restrict__ vel,
                                                 float *__restrict__ elem_den, float *__restrict__ temp, float *__restrict__ hh)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t p = index ? (int64_t) index[i] : i;
    pos[3 * i] = position[3 * p], pos[3 * i + 1] = position[3 * p + 1], pos[3 * i + 2] = position[3 * p + 2];
    if (vel) {
        // vel *= np.sqrt(atime) (or vel /= atime) with a float64 scalar: numpy forms the result in double and rounds it
        // back to float32
        const double f = cfg.velocity_factor;
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = (double) velocity[3 * p + c];
            vel[3 * i + c] = (float) (cfg.velocity_divides ? v / f : v * f);
        }
    }
    hh[i] = smoothing[p];
    const float rho = density[p];
    const float den = __fmul_rn(rho, cfg.dens_conv);  // physical H atoms / cm^3
    float t_used = 0;
    if (temp || ion.coef) {
        // abstractsnapshot.py:121-154: mu = 4 / (hy (3 + 4 ne) + 1) in float32; then, with a unit system of Python floats,
        // ienergy * unit, mu * ienergy and the (gamma-1) mp / kB factor in float32 as well; with numpy float64 units
        // (headers read from files) those three products are double and the result is rounded to float32 once
        const float ne = nelec ? nelec[p] : cfg.nelec_const;
        const float mu = __fdiv_rn(4.0f, __fadd_rn(__fmul_rn(cfg.hy_mass, __fadd_rn(3.0f, __fmul_rn(4.0f, ne))), 1.0f));
        float t;
        if (cfg.temp_double) t = (float) (cfg.temp_factor * ((double) mu * ((double) ienergy[p] * cfg.unit_ienergy)));
        else t = __fmul_rn((float) cfg.temp_factor, __fmul_rn(mu, __fmul_rn(ienergy[p], (float) cfg.unit_ienergy)));
        t_used = t <= 0 ? 1.0f : t;
        if (temp) temp[i] = t_used;
    }
    float mf = mass_frac ? mass_frac[p * mass_frac_stride] : cfg.mass_frac_const;
    if (mf <= 0) mf = 0;
    float ed = __fmul_rn(__fmul_rn(den, cfg.rscale), mf);  // (den * rscale) * mass_frac, spectra.py:593
    if (cfg.neutral_hydrogen) {
        float x = nh0[p];
        if (cfg.sf_neutral && rho > cfg.dens_thresh_code)
            x = cfg.redshift_coverage ? (float) rahmati_neutral_fraction((double) den, 1e4, cfg.gray_opac, cfg.gamma_uvb, cfg.f_bar) : 1.0f;
        ed = __fmul_rn(ed, x);
    } else if (ion.coef) {
        ed = __fmul_rn(ed, ion_fraction(ion, den, t_used));
    }
    elem_den[i] = __fdiv_rn(ed, cfg.amumass);
}

}  // namespace

}  // namespace fsb

using namespace fsb;

extern "C" int fsb_prepare_particles(const fsb_prep *cfg, const int32_t *index, int64_t m, const float *position,
                                     const float *velocity, const float *density, const float *ienergy, const float *nelec,
                                     const float *nh0, const float *smoothing, const float *mass_frac,
                                     int64_t mass_frac_stride, const fsb_ion_table *ion, float *pos, float *vel, float *elem_den,
                                     float *temp, float *hh, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(cfg != nullptr && m >= 0, "bad arguments");
    if (m == 0) return FSB_OK;
    FSB_REQUIRE(position && density && smoothing && pos && elem_den && hh, "NULL array");
    FSB_REQUIRE((vel == nullptr) == (velocity == nullptr) || vel == nullptr, "velocity output without velocity input");
    FSB_REQUIRE((temp == nullptr && ion == nullptr) || ienergy != nullptr, "temperatures need the internal energy");
    fsb_ion_table tb = {};
    if (ion) {
        tb = *ion;
        FSB_REQUIRE(!cfg->neutral_hydrogen, "an ion table and the neutral-hydrogen route exclude each other");
        FSB_REQUIRE(tb.coef != nullptr && tb.nd >= 2 && tb.nt >= 2 && tb.pad >= 2, "bad ion table");
        FSB_REQUIRE(tb.dens_span > 0 && tb.temp_span > 0 && tb.dens_lo > 0 && tb.temp_lo > 0, "bad ion table grid");
    }
    FSB_REQUIRE(!cfg->neutral_hydrogen || nh0 != nullptr, "neutral hydrogen needs the snapshot's neutral fraction");
    FSB_REQUIRE(cfg->amumass > 0, "amumass must be positive");
    count_launch();
    k_prepare<<<(unsigned) ((m + 255) / 256), 256, 0, stream>>>(*cfg, tb, index, m, position, vel ? velocity : nullptr, density, ienergy, nelec, nh0,
                                                              smoothing, mass_frac, mass_frac_stride, pos, vel, elem_den, temp, hh);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

// _filter_particles of the metal-ion route: the entries of `index` (NULL = 0..m-1) whose element density is positive,
// in order.  out_index may alias nothing; *count is written after a stream synchronisation.
extern "C" int fsb_prepare_select(const fsb_prep *cfg, const int32_t *index, int64_t m, const float *density,
                                  const float *mass_frac, int64_t mass_frac_stride, int32_t *out_index, int64_t *count,
                                  void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(cfg != nullptr && count != nullptr && m >= 0, "bad arguments");
    *count = 0;
    if (m == 0) return FSB_OK;
    FSB_REQUIRE(m <= (int64_t) INT32_MAX, "m exceeds int32 particle indices");
    FSB_REQUIRE(density && out_index, "NULL array");
    Scratch flag, block_count, block_start;
    const int threads = 1024;
    const int64_t nblocks = (m + threads - 1) / threads;
    FSB_TRY(flag.alloc((size_t) m, stream));
    FSB_TRY(block_count.alloc(sizeof(int32_t) * (size_t) (nblocks + 1), stream));
    FSB_TRY(block_start.alloc(sizeof(int64_t) * (size_t) (nblocks + 1), stream));
    count_launch(); k_select_flags<<<(unsigned) nblocks, threads, 0, stream>>>(*cfg, index, m, density, mass_frac, mass_frac_stride, flag.as<uint8_t>());
    count_launch(); k_flag_block_counts<<<(unsigned) nblocks, threads, 0, stream>>>(flag.as<uint8_t>(), m, block_count.as<int32_t>());
    count_launch(); k_scan_single<int32_t, int64_t><<<1, 1024, 0, stream>>>(block_count.as<int32_t>(), block_start.as<int64_t>(), nblocks, nullptr);
    count_launch(); k_flag_compact<<<(unsigned) nblocks, threads, 0, stream>>>(flag.as<uint8_t>(), m, block_start.as<int64_t>(), index, out_index);
    FSB_CUDA_TRY(cudaGetLastError());
    FSB_CUDA_TRY(cudaMemcpyAsync(count, block_start.as<int64_t>() + nblocks, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaStreamSynchronize(stream));
    return FSB_OK;
}

// Kernel support radius of every particle (abstractsnapshot.py:253-282).  mode 0: a = SmoothingLength -> a / 2;
// mode 1: a = Volume -> a^(1/3); mode 2: a = Masses, b = Density -> (a / b)^(1/3).  DEVICE float32 arrays of n entries.
extern "C" int fsb_smoothing_lengths(const float *a, const float *b, int64_t n, int32_t mode, float *hh, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(n >= 0 && mode >= 0 && mode <= 2, "bad arguments");
    if (n == 0) return FSB_OK;
    FSB_REQUIRE(a && hh && (mode != 2 || b), "NULL array");
    count_launch();
    k_smooth_length<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(a, b, n, mode, hh);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}
