/* fsb200.h — C ABI of libfsb200.so: the B200-native (sm_100a) replacement for the native hot path
 * of sbird/fake_spectra.  Plain pointers and sizes only; no torch / Python types.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference's
 * fake_spectra/ directory).  Unless a name ends in _host, all array pointers are DEVICE pointers
 * on the current CUDA device and the call is asynchronous on `stream` (a cudaStream_t passed as
 * void*; NULL = the legacy default stream) except where noted.  The library never frees or
 * retains caller memory.  All functions return FSB_OK (0) or a negative FSB_E* code; the
 * message of the last error on the calling thread is available from fsb_last_error().
 * The library never calls exit()/abort() (the reference can: index_table.cpp:204-208).
 */
#ifndef FSB200_H
#define FSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSB_ABI_VERSION 1

#if defined(__GNUC__)
#define FSB_API __attribute__((visibility("default")))
#else
#define FSB_API
#endif

/* status codes */
#define FSB_OK 0
#define FSB_EINVAL (-1)   /* bad argument (shape, kernel id, NULL pointer, nbins <= 0 ...)            */
#define FSB_ECUDA (-2)    /* a CUDA runtime call or kernel failed; see fsb_last_error()              */
#define FSB_ENOMEM (-3)   /* device allocation failed                                                */
#define FSB_EVORONOI (-4) /* Voronoi cell ownership not contiguous (reference: exit(1))              */
#define FSB_ENODEV (-5)   /* no CUDA device / not an sm_100 device                                   */

/* kernel ids: singleabs.h:9-12 */
#define FSB_KERNEL_TOPHAT 0
#define FSB_KERNEL_CUBIC 1
#define FSB_KERNEL_VORONOI 2
#define FSB_KERNEL_QUINTIC 3

/* arithmetic of the tau kernel */
#define FSB_PRECISION_FP64 0 /* parity mode: <= 1e-10 relative to the reference C++               */
#define FSB_PRECISION_FP32 1 /* fast path: <= 1e-5 absolute on flux exp(-tau)                      */

/* Voigt evaluation strategy (FP64 only) */
#define FSB_VOIGT_FAST 0     /* this library's own small-y expansion, falls back to EXACT outside its domain */
#define FSB_VOIGT_EXACT 1    /* operation-for-operation restatement of Faddeeva.cpp:679-971      */

/* Scalars of one absorption line + spectrum geometry.  Same quantities, units and meaning as the
 * positional arguments of _Particle_Interpolate (py_module.cpp:103-115) and of the
 * ParticleInterp / LineAbsorption constructors (part_int.h:27, absorption.cpp:152-161). */
typedef struct fsb_params {
    int32_t nbins;      /* pixels per spectrum                                                    */
    int32_t kernel;     /* FSB_KERNEL_*                                                           */
    double box;         /* box size, comoving kpc/h                                               */
    double velfac;      /* km/s per comoving kpc/h                                                */
    double atime;       /* scale factor (carried for parity of the argument list; unused)         */
    double lambda_cm;   /* rest wavelength in cm                                                  */
    double gamma;       /* damping constant 1/s (0 => pure Gaussian, spectra.py:669-672)          */
    double fosc;        /* oscillator strength                                                    */
    double amumass;     /* ion mass in amu                                                        */
    double tautail;     /* per-particle tau below which the pixel march stops (spectra.py:135)    */
    int32_t precision;  /* FSB_PRECISION_*                                                        */
    int32_t voigt;      /* FSB_VOIGT_*                                                            */
    int32_t seg_pairs;  /* pairs per work item; 0 = choose from the problem size                  */
    int32_t reserved;
} fsb_params;

/* Counters a compute call can report (device-resident, uint64 each). */
typedef struct fsb_counters {
    uint64_t pairs;     /* candidate pairs visited                                                */
    uint64_t pixels;    /* pixel contributions added                                              */
    uint64_t voigt;     /* Voigt profile evaluations (singleabs.h:157 calls)                      */
    uint64_t lanes;     /* lane-slots spent in the pixel march (pixels + discarded lanes)         */
    uint64_t steps_near_u; /* warp march steps by route of the tau kernel: table + Gaussian        */
    uint64_t steps_near;   /*   table only (Gaussian negligible)                                   */
    uint64_t steps_far;    /*   damping-wing series, |x| >= 16                                     */
    uint64_t steps_mixed;  /*   straddling |x| = 16 (generic evaluation)                           */
    uint64_t steps_slow;   /*   exact Faddeeva or sub-sampled pixels                               */
    uint64_t reserved;
} fsb_counters;
#define FSB_N_COUNTERS 10

/* ---- library ------------------------------------------------------------------------------ */
FSB_API int fsb_abi_version(void);
FSB_API const char *fsb_strerror(int code);
FSB_API const char *fsb_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
FSB_API uint64_t fsb_kernel_launches(void);
/* Device properties the bench reports: SM count, clock (kHz), compute capability major/minor. */
FSB_API int fsb_device_info(int32_t *sm_count, int32_t *clock_khz, int32_t *cc_major, int32_t *cc_minor);

/* ---- candidate index (replaces IndexTable: index_table.h:10-52) ---------------------------- */
typedef struct fsb_index fsb_index;

/* IndexTable ctor (index_table.cpp:7-18) + get_near_particles (:130-150).
 * Builds, for every sightline, the ascending list of particles whose kernel support reaches it
 * (exact predicate of index_table.cpp:22-113) and the periodic squared impact parameter.
 * cofm[nlos*3] f64, axis[nlos] i32 (1-based), pos[npart*3] f32, h[npart] f32.
 * Synchronises `stream` once (the pair count sizes the lists).  Free with fsb_index_free. */
FSB_API int fsb_index_build(double box, const double *cofm, const int32_t *axis, int32_t nlos,
                    const float *pos, const float *h, int64_t npart, void *stream, fsb_index **out);
/* Same, with the list sizes handed in: counts[nlos] int32 (DEVICE) = the output of fsb_count_pairs for exactly
 * these sightlines against exactly these particles (e.g. summed over the ranks that each counted a slice of the
 * particles).  Skips the counting pass over the particles.  Wrong counts are undefined behaviour. */
FSB_API int fsb_index_build_counted(double box, const double *cofm, const int32_t *axis, int32_t nlos,
                            const float *pos, const float *h, int64_t npart, const int32_t *counts,
                            void *stream, fsb_index **out);
FSB_API int fsb_index_free(fsb_index *idx, void *stream);
/* sizes: number of sightlines, total candidate pairs, longest single list */
FSB_API int fsb_index_sizes(const fsb_index *idx, int32_t *nlos, int64_t *npairs, int64_t *max_list);
/* Copies the lists out (device to device): offsets[nlos+1] i64, particle[npairs] i32 ascending
 * within each line (the iteration order of std::map<int,double>, part_int.cpp:35), dr2[npairs] f64
 * (the map's values).  Any of the three may be NULL. */
FSB_API int fsb_index_export(const fsb_index *idx, int64_t *offsets, int32_t *particle, double *dr2, void *stream);

/* ---- accumulation (replaces ParticleInterp::compute_tau / compute_colden) ------------------ */
/* part_int.cpp:20-51.  tau[nlos*nbins] f64 row-major is ACCUMULATED into (callers zero it,
 * py_module.cpp:194).  vel[npart*3], dens/temp/h[npart] f32.  counters may be NULL. */
FSB_API int fsb_compute_tau(const fsb_index *idx, const fsb_params *p, const float *pos, const float *vel,
                    const float *dens, const float *temp, const float *h, double *tau,
                    fsb_counters *counters, void *stream);
/* Same geometry, several lines of one ion in ONE pass (Lya+Lyb ...): lines differ only in
 * lambda_cm, gamma, fosc (p[i].nbins/kernel/box/velfac/amumass/tautail must agree).
 * tau[nlines][nlos*nbins]. */
FSB_API int fsb_compute_tau_multi(const fsb_index *idx, const fsb_params *p, int32_t nlines, const float *pos,
                          const float *vel, const float *dens, const float *temp, const float *h,
                          double *tau, fsb_counters *counters, void *stream);
/* fsb_compute_tau_multi restricted to the sightlines [line_begin, line_end) of the index; tau is still the full
 * [nlines][nlos*nbins] array (only the rows of the range are touched).  Lets a caller overlap a collective on the
 * finished rows of one block with the computation of the next (particle-sharded mode: spectra.py:825-831 sums the
 * whole array after the fact).  One work row per sightline. */
FSB_API int fsb_compute_tau_multi_range(const fsb_index *idx, const fsb_params *p, int32_t nlines, int32_t line_begin,
                                int32_t line_end, const float *pos, const float *vel, const float *dens,
                                const float *temp, const float *h, double *tau, void *stream);

/* ---- sightline-sharded multi-GPU: rows pushed to the peers while the kernel runs -------------------------
 * One process per GPU, every rank interpolates a block of the sightlines (no reference counterpart: the reference
 * shards particles and Allreduces, spectra.py:825-831).  Every rank owns a FULL result array [nlines_total][numlos]
 * [nbins] in memory obtained from fsb_peer_alloc, and maps the other ranks' arrays with fsb_peer_open (CUDA IPC over
 * NVLink / NVSwitch).  fsb_compute_tau_multi_push is fsb_compute_tau_multi plus: the warp that completes a sightline's
 * row stores that row into every array listed in `push`, from inside the tau kernel, so the gather of the result rows
 * overlaps the computation and needs no collective (a barrier across the ranks afterwards makes the arrays complete). */
#define FSB_MAX_PEERS 16
typedef struct fsb_push {
    int32_t npeers;           /* number of destination arrays (the rank's own array may be one of them)            */
    int32_t reserved;
    int64_t line_stride;      /* elements between consecutive LINES in the destination arrays: numlos * nbins         */
    double *dest[FSB_MAX_PEERS]; /* per destination: address of [first line of this call][first sightline of this
                                 rank's block][0] in that array, valid in THIS process (own or fsb_peer_open'ed)     */
} fsb_push;
/* cudaMalloc'ed device memory that other processes of this node can map: handle[64] = cudaIpcMemHandle_t. */
FSB_API int fsb_peer_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle64);
FSB_API int fsb_peer_free(void *dev_ptr);
/* Maps another process's fsb_peer_alloc memory into this process (peer access is enabled as needed). */
FSB_API int fsb_peer_open(const unsigned char *handle64, void **dev_ptr);
FSB_API int fsb_peer_close(void *dev_ptr);
/* tau[nlines][nlos*nbins] as fsb_compute_tau_multi (the rank's working rows); one work row per sightline is forced. */
FSB_API int fsb_compute_tau_multi_push(const fsb_index *idx, const fsb_params *p, int32_t nlines, const float *pos,
                               const float *vel, const float *dens, const float *temp, const float *h,
                               double *tau, const fsb_push *push, void *stream);

/* part_int.cpp:53-84, with nweights density-like columns sharing one geometry pass
 * (spectra.py:945-1024 issues one colden call per weight).  dens[nweights][npart] f32,
 * colden[nweights][nlos*nbins] f64, accumulated into. */
FSB_API int fsb_compute_colden(const fsb_index *idx, const fsb_params *p, const float *pos, const float *dens,
                       int32_t nweights, const float *h, double *colden, fsb_counters *counters,
                       void *stream);

/* ---- one-shot boundary (replaces Py_Particle_Interpolation, py_module.cpp:103-233) --------- */
/* compute_tau != 0: tau; else column density.  out[nlos*nbins] f64 accumulated into.
 * vel/temp are ignored (may be NULL) when compute_tau == 0, as in spectra.py:570-571. */
FSB_API int fsb_particle_interpolate(int32_t compute_tau, const fsb_params *p, const float *pos, const float *vel,
                             const float *dens, const float *temp, const float *h, int64_t npart,
                             const int32_t *axis, const double *cofm, int32_t nlos, double *out,
                             void *stream);
/* Same with HOST pointers: copies inputs to the device, runs, copies out[nlos*nbins] back
 * (out is overwritten with the zero-initialised result, like the array the reference returns).
 * Synchronous.  This is the call the reference's Python binding would make. */
FSB_API int fsb_particle_interpolate_host(int32_t compute_tau, const fsb_params *p, const float *pos, const float *vel,
                                  const float *dens, const float *temp, const float *h, int64_t npart,
                                  const int32_t *axis, const double *cofm, int32_t nlos, double *out);

/* Several lines of one ion from one upload and one index (HOST pointers): out[nlines][nlos*nbins].
 * What Spectra.get_tau does for Lya then Lyb as two boundary calls, in one.  With compute_tau == 0,
 * nlines counts density-like weight columns instead: dens[nlines][npart], p[0] is used
 * (the three passes of Spectra.get_velocity, spectra.py:945-956, in one). */
FSB_API int fsb_particle_interpolate_multi_host(int32_t compute_tau, const fsb_params *p, int32_t nlines,
                                        const float *pos, const float *vel, const float *dens, const float *temp,
                                        const float *h, int64_t npart, const int32_t *axis, const double *cofm,
                                        int32_t nlos, double *out);
/* Optical depths of several IONS from one upload and one index (HOST pointers): what Spectra.get_tau does for H I, C IV,
 * Mg II ... as one boundary call per ion and line (spectra.py:801-831), each re-reading the particles and rebuilding the
 * index (part_int.cpp:22).  dens_columns[nions]: one HOST array of npart species densities per ion (used where they
 * lie); line_ion[i] (ascending) names the column of line i, p[i].amumass the ion's mass; out[nlines][nlos*nbins].  Not
 * for the Voronoi kernel. */
FSB_API int fsb_particle_interpolate_ions_host(const fsb_params *p, int32_t nlines, const int32_t *line_ion, int32_t nions,
                                       const float *pos, const float *vel, const float *const *dens_columns, const float *temp,
                                       const float *h, int64_t npart, const int32_t *axis, const double *cofm,
                                       int32_t nlos, double *out);

/* ---- particle filter (replaces Py_near_lines, py_module.cpp:25-99) ------------------------- */
/* Ascending indices of particles with at least one candidate sightline.  out_index must hold
 * npart int32 (worst case); *count (HOST) receives the number written.  Synchronises `stream`. */
FSB_API int fsb_near_lines(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                   const double *cofm, int32_t nlos, int32_t *out_index, int64_t *count, void *stream);
FSB_API int fsb_near_lines_host(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                        const double *cofm, int32_t nlos, int32_t *out_index, int64_t *count);

/* Candidate pairs per sightline (the sizes of the lists fsb_index_build would make) without building them:
 * the count pass that balances sightline blocks across GPUs.  counts[nlos] int32, overwritten; DEVICE
 * pointers + stream (synchronised once), or HOST pointers (synchronous).  Counts are additive over any split of
 * the particles: ranks may each count a slice and sum.  No reference counterpart (the reference shards particles). */
FSB_API int fsb_count_pairs(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                    const double *cofm, int32_t nlos, int32_t *counts, void *stream);
FSB_API int fsb_count_pairs_host(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                         const double *cofm, int32_t nlos, int32_t *counts);

/* ---- Voronoi cells (replaces IndexTable::assign_cells, index_table.cpp:152-223) ------------ */
/* cells[2*npairs] f32 in list order: (lo, hi) extent of each candidate's cell along its
 * sightline, 3*box sentinel when it owns nothing.  Returns FSB_EVORONOI (after finishing all
 * lines) where the reference would exit(1).  Synchronises `stream`. */
FSB_API int fsb_assign_cells(const fsb_index *idx, double box, const double *cofm, const int32_t *axis,
                     const float *pos, float *cells, void *stream);

/* ---- roofline denominators ---------------------------------------------------------------- */
/* Measured FMA throughput of the current device in TFLOP/s (2 flops per FMA), FP64 if fp64 != 0
 * else FP32: 8 independent chains per thread, best of several launches.  Synchronous. */
FSB_API int fsb_measure_fma_peak(int32_t fp64, double *tflops, void *stream);

/* ---- snapshot fields -> inputs of the interpolation, on the device -------------------------
 * Replaces the host numpy of Spectra._read_particle_data (spectra.py:550-617) with its helpers
 * AbstractSnapshot.get_peculiar_velocity / get_temp / get_smooth_length (abstractsnapshot.py:114-154,253-282) and
 * GasProperties.get_code_rhoH / get_reproc_HI (gas_properties.py:104-146) for the particles listed in `index`
 * (e.g. the output of fsb_near_lines; NULL = all m particles).  Inputs are the raw snapshot fields of one segment
 * (DEVICE, float32, Gadget-HDF5 names): Coordinates[n][3], Velocities[n][3], Density, InternalEnergy,
 * ElectronAbundance (NULL: cfg.nelec_const), NeutralHydrogenAbundance (needed when cfg.neutral_hydrogen),
 * the kernel support radii of all n particles (fsb_smoothing_lengths), and an optional mass-fraction column (element e of GFM_Metals[n][9]: pointer to
 * [0][e], stride 9; NULL: cfg.mass_frac_const).  Outputs (DEVICE, float32, m entries): pos[m][3], vel[m][3] (NULL to
 * skip), elem_den, temp (NULL to skip; values <= 0 become 1), hh.
 * Metal ions (spectra.py:598-611,637-664): fsb_prepare_select keeps the particles with mass in the element
 * (_filter_particles), and `ion` (may be NULL) is the Cloudy table of CloudyTable.ion (convert_cloudy.py:167-200) at the
 * snapshot's redshift for one (element, ion): log10 of the ion fraction on a regular (log10 nH, log10 T) grid,
 * interpolated with cubic B-splines in scipy.ndimage.map_coordinates' mode "nearest". */
typedef struct fsb_prep {
    float dens_conv;         /* code density -> physical H atoms / cm^3 (gas_properties.py:108)                 */
    float rscale;            /* cm per comoving kpc/h (spectra.py:210)                                          */
    float hy_mass;           /* hydrogen mass fraction of the temperature formula (0.76)                        */
    float nelec_const;       /* electron abundance when the snapshot has none                                   */
    float mass_frac_const;   /* element mass fraction when the snapshot has no metal table (0.76 / 0.24)        */
    float amumass;           /* ion mass in amu (1 for "Z")                                                     */
    float dens_thresh_code;  /* star-formation threshold in code density units (gas_properties.py:138)          */
    int32_t velocity_divides;/* peculiar velocity = Velocities / velocity_factor (MP-Gadget) instead of times it      */
    int32_t neutral_hydrogen;/* multiply by the (reprocessed) neutral fraction: H I                             */
    int32_t sf_neutral;      /* replace the neutral fraction above the threshold by the Rahmati value at 1e4 K  */
    int32_t redshift_coverage; /* the UVB table covers the redshift (else star-forming gas is fully neutral)    */
    double gray_opac, gamma_uvb, f_bar; /* Rahmati et al. 2013 parameters at this redshift                      */
    double unit_ienergy;     /* UnitInternalEnergy_in_cgs                                                       */
    double temp_factor;      /* (gamma - 1) m_p / k_B                                                           */
    int32_t temp_double;     /* the unit system holds numpy float64 values (headers read from files): numpy then forms
                                the temperature in double and the reference rounds it to float32 once; 0: Python
                                floats, every operation in float32 */
    int32_t reserved;
    double velocity_factor;  /* peculiar velocity = (float)(Velocities * sqrt(a)) for Gadget HDF5 (abstractsnapshot.py:
                                114-119), (float)(Velocities / a) for MP-Gadget (:398-405), formed in double */
} fsb_prep;
typedef struct fsb_ion_table {
    const double *coef;      /* DEVICE [nd + 2 pad][nt + 2 pad]: B-spline coefficients of the table padded by `pad` cells
                                of edge values per side (scipy.ndimage.spline_filter(order 3, mode "nearest") of it) */
    int32_t nd, nt, pad, reserved;
    double dens0, dens_span; /* log10 nH of the first grid point, last minus first                              */
    double temp0, temp_span; /* log10 T  of the first grid point, last minus first                              */
    float dens_lo, dens_hi;  /* nH and T are clipped to the table's bounds first (spectra.py:649-663)           */
    float temp_lo, temp_hi;
    float rho_factor;        /* gas density -> Cloudy hden, 0.774132 (convert_cloudy.py:183)                    */
    float reserved2;
} fsb_ion_table;
FSB_API int fsb_prepare_particles(const fsb_prep *cfg, const int32_t *index, int64_t m, const float *position,
                          const float *velocity, const float *density, const float *ienergy, const float *nelec,
                          const float *nh0, const float *smoothing, const float *mass_frac,
                          int64_t mass_frac_stride, const fsb_ion_table *ion, float *pos, float *vel, float *elem_den,
                          float *temp, float *hh, void *stream);
/* get_smooth_length (abstractsnapshot.py:253-282) for the n particles of a segment, DEVICE float32.  mode 0:
 * a = SmoothingLength -> a / 2 (Gadget); 1: a = Volume -> a^(1/3) (Arepo); 2: a = Masses, b = Density -> (a / b)^(1/3). */
FSB_API int fsb_smoothing_lengths(const float *a, const float *b, int64_t n, int32_t mode, float *hh, void *stream);
/* The entries of index[m] (NULL = 0..m-1) whose element density (Density * dens_conv * rscale) * mass fraction is
 * positive, in order (DEVICE int32 out_index[m]); *count (HOST) is valid on return (synchronises the stream). */
FSB_API int fsb_prepare_select(const fsb_prep *cfg, const int32_t *index, int64_t m, const float *density,
                          const float *mass_frac, int64_t mass_frac_stride, int32_t *out_index, int64_t *count,
                          void *stream);

/* ---- flux statistics on device-resident tau ----------------------------- */
/* Replaces get_mean_flux_scale / _rescale_mean_flux (py_module.cpp:235-282): the factor s with
 * mean(exp(-s tau)) = mean_flux_desired over the pixels with tau <= thresh, by the reference's Newton
 * iteration (same update, same clamp, same stopping rule |ds| <= tol s).  tau: n DEVICE doubles;
 * *scale_out and *iterations (may be NULL) are HOST.  n == 0 gives 0 (fluxstatistics.py:39-40).
 * Synchronises `stream` once per iteration. */
FSB_API int fsb_rescale_mean_flux(const double *tau, int64_t n, double mean_flux_desired, double tol, double thresh,
                          double *scale_out, int32_t *iterations, void *stream);
/* One evaluation of the sums of that iteration, over the pixels with tau <= thresh: sum exp(-scale tau),
 * sum tau exp(-scale tau) and their number (HOST outputs).  The mean flux of Spectra.get_mean_flux
 * (spectra.py:1272-1276) is sum_flux / used at scale 1.  Synchronises `stream`. */
FSB_API int fsb_flux_sums(const double *tau, int64_t n, double scale, double thresh, double *sum_flux, double *sum_tau_flux,
                  int64_t *used, void *stream);
/* out[r] = max_j a[r][j] for a DEVICE array of nrows x n doubles: the per-sightline maximum optical depth that
 * Spectra._filter_tau compares with tau_thresh (spectra.py:1258). */
FSB_API int fsb_row_max(const double *a, int64_t nrows, int64_t n, double *out, void *stream);
/* counts[nbins] (DEVICE, overwritten) = histogram of exp(-scale tau) on nbins equal bins of [0, 1] with
 * numpy.histogram's edge rules: the counts behind fluxstatistics.flux_pdf (fluxstatistics.py:43-52). */
FSB_API int fsb_flux_pdf(const double *tau, int64_t n, double scale, int32_t nbins, uint64_t *counts, void *stream);
/* out[i] = exp(-scale tau[i]) / mean_flux - 1 (fluxstatistics.py:100), DEVICE arrays. */
FSB_API int fsb_delta_flux(const double *tau, int64_t n, double scale, double mean_flux, double *out, void *stream);
/* The 1-D flux power of fluxstatistics.flux_power (fluxstatistics.py:74-108) without a library FFT: a two-level
 * direct Fourier sum in shared memory for any pixel count (csrc/fsb_stats.cu).  in: DEVICE [nspec][npix] doubles.
 * mode 0: x_s = exp(-scale in_s) / mean_flux - 1 formed on the fly; mode 1: x_s = in_s.  per_row == NULL:
 * power[k] += factor * sum_s |rfft(x_s)[k]|^2 (DEVICE [npix/2 + 1]); else per_row[s][k] = |rfft(x_s)[k]|^2 / npix^2
 * (fluxstatistics._powerspectrum, fluxstatistics.py:54-61) and power is not touched. */
FSB_API int fsb_flux_power(const double *in, int64_t nspec, int32_t npix, int32_t mode, double scale, double mean_flux,
                   double factor, double *power, double *per_row, void *stream);
/* power[k] += factor * sum_s (re^2 + im^2) of rfft_interleaved[s][k] (nspec x nk complex doubles, DEVICE):
 * the accumulation of fluxstatistics.py:54-61,102-104 for one batch of sightlines. */
FSB_API int fsb_power_accumulate(const double *rfft_interleaved, int64_t nspec, int32_t nk, double factor, double *power,
                         void *stream);

/* ---- Voigt profile, for tests ------------------------------------------------------------- */
/* out[i] = Re w(x[i] + i y[i]) (singleabs.h:56-61) with the strategy `voigt` (FSB_VOIGT_*). */
FSB_API int fsb_voigt_profile(const double *x, const double *y, double *out, int64_t n, int32_t voigt, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_H */
