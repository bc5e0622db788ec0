// Flux statistics on optical depths that are already resident in HBM: the immediate consumers of
// tau.
//   fsb_rescale_mean_flux : replaces get_mean_flux_scale (py_module.cpp:235-262), the Newton iteration
//                           that rescales tau to an observed mean flux
//   fsb_flux_pdf          : the histogram behind fluxstatistics.flux_pdf (fluxstatistics.py:43-52)
//   fsb_delta_flux        : exp(-scale tau)/mean_flux - 1, the input of the 1-D flux power
//                           (fluxstatistics.py:100)
//   fsb_power_accumulate  : sum over sightlines of |rfft|^2 (fluxstatistics.py:54-61, 102-104)
// All four are single streaming passes over tau: HBM-bound (8 bytes per pixel read; the power pass
// writes 8 more).  Reductions use a fixed CTA count and a fixed tree, so results are bit-reproducible.
#include <math.h>

#include "fsb_common.cuh"

namespace fsb {

namespace {

constexpr int kStatThreads = 256;
constexpr int kStatBlocksPerSM = 6;  // 256 threads x 40 registers: six CTAs are resident per SM
constexpr int kMaxStatBlocks = 148 * kStatBlocksPerSM * 2;  // room for larger parts

// exp(t) for t <= 0 (fluxes): t = (64 e + j) ln2/64 + r with |r| <= ln2/128, exp(t) = 2^e 2^(j/64) exp(r); 2^(j/64)
// from a 64-entry table staged in shared memory, exp(r) by its degree-5 Taylor polynomial (remainder 3e-17).
// About 10 FP64 instructions instead of the library routine's ~30, which would make these streaming passes
// FP64-bound instead of HBM-bound.  Within 1.5 ulp; results below 1e-300 flush to zero.
__device__ const double d_exp2_64[64] = {1.00000000000000000e+00, 1.01088928605170048e+00, 1.02189714865411663e+00, 1.03302487902122841e+00, 1.04427378242741375e+00, 1.05564517836055716e+00, 1.06714040067682370e+00, 1.07876079775711986e+00, 1.09050773266525769e+00, 1.10238258330784089e+00, 1.11438674259589243e+00, 1.12652161860824185e+00, 1.13878863475669156e+00, 1.15118922995298267e+00, 1.16372485877757748e+00, 1.17639699165028122e+00, 1.18920711500272103e+00, 1.20215673145270308e+00, 1.21524735998046896e+00, 1.22848053610687002e+00, 1.24185781207348400e+00, 1.25538075702469110e+00, 1.26905095719173322e+00, 1.28287001607877826e+00, 1.29683955465100964e+00, 1.31096121152476441e+00, 1.32523664315974132e+00, 1.33966752405330292e+00, 1.35425554693689265e+00, 1.36900242297459052e+00, 1.38390988196383202e+00, 1.39897967253831124e+00, 1.41421356237309515e+00, 1.42961333839197002e+00, 1.44518080697704665e+00, 1.46091779418064704e+00, 1.47682614593949935e+00, 1.49290772829126484e+00, 1.50916442759342284e+00, 1.52559815074453842e+00, 1.54221082540794074e+00, 1.55900440023783693e+00, 1.57598084510788650e+00, 1.59314215134226700e+00, 1.61049033194925428e+00, 1.62802742185734783e+00, 1.64575547815396495e+00, 1.66367658032673638e+00, 1.68179283050742900e+00, 1.70010635371852348e+00, 1.71861929812247793e+00, 1.73733383527370622e+00, 1.75625216037329945e+00, 1.77537649252652119e+00, 1.79470907500310717e+00, 1.81425217550039886e+00, 1.83400808640934243e+00, 1.85397912508338547e+00, 1.87416763411029996e+00, 1.89457598158696561e+00, 1.91520656139714740e+00, 1.93606179349229435e+00, 1.95714412417540018e+00, 1.97845602638795093e+00};

__device__ __forceinline__ void stage_exp_table(double *tab)
{
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tab[i] = d_exp2_64[i];
    __syncthreads();
}

__device__ __forceinline__ double exp_nonpos(double t, const double *__restrict__ tab)
{
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    double kd = fma(t, 9.23324826168936567683e+01, magic);
    const int ki = __double2loint(kd);
    kd -= magic;
    double r = fma(kd, -1.08304246950865490362e-02, t);
    r = fma(kd, -1.16259642343943700740e-12, r);
    double p = fma(r, 1.0 / 120, 1.0 / 24);
    p = fma(p, r, 1.0 / 6);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = p * tab[ki & 63];
    const double scaled = __hiloint2double(__double2hiint(v) + ((ki >> 6) << 20), __double2loint(v));
    return t < -690.0 ? 0.0 : (t > 0.0 ? exp(t) : scaled);  // t > 0 (negative tau) never happens for optical depths
}

struct FluxSums {
    double flux, tau_flux;     // sum exp(-s tau), sum tau exp(-s tau) over pixels with tau <= thresh
    unsigned long long used;   // number of such pixels
};

__device__ __forceinline__ void warp_reduce(double &a, double &b, unsigned long long &n)
{
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, d);
        b += __shfl_down_sync(0xffffffffu, b, d);
        n += __shfl_down_sync(0xffffffffu, n, d);
    }
}

// One Newton step's sums.  Each CTA owns a contiguous slab of pixels (coalesced 16-byte loads) and
// writes one partial; k_flux_sums_final adds the partials in index order.
__global__ void __launch_bounds__(kStatThreads, kStatBlocksPerSM)
k_flux_sums(const double *__restrict__ tau, int64_t n, double scale, double thresh, FluxSums *__restrict__ partial)
{
    __shared__ double etab[64];
    stage_exp_table(etab);
    double f = 0, tf = 0;
    unsigned long long used = 0;
    auto add = [&](double t) {
        if (t > thresh) return;
        const double e = exp_nonpos(-scale * t, etab);
        f += e;
        tf = fma(e, t, tf);
        ++used;
    };
    // grid-stride over 16-byte pairs, four loads in flight per thread; the element-to-thread map depends only on
    // (n, grid), so the partial sums are reproducible
    const bool aligned = (reinterpret_cast<uintptr_t>(tau) & 15u) == 0;
    const int64_t head = aligned ? 0 : (n > 0 ? 1 : 0);       // one scalar element brings the rest to 16 bytes
    const double2 *t2 = reinterpret_cast<const double2 *>(tau + head);
    const int64_t n2 = (n - head) / 2, stride = (int64_t) gridDim.x * kStatThreads;
    for (int64_t i = (int64_t) blockIdx.x * kStatThreads + threadIdx.x; i < n2; i += 4 * stride) {
        double2 v[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i + u * stride < n2 ? t2[i + u * stride] : make_double2(INFINITY, INFINITY);
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i + u * stride < n2) {
                add(v[u].x);
                add(v[u].y);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (head) add(tau[0]);
        if ((n - head) & 1) add(tau[n - 1]);
    }
    __shared__ double sf[kStatThreads / 32], stf[kStatThreads / 32];
    __shared__ unsigned long long su[kStatThreads / 32];
    warp_reduce(f, tf, used);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sf[warp] = f, stf[warp] = tf, su[warp] = used;
    __syncthreads();
    if (threadIdx.x == 0) {
        FluxSums s = {0, 0, 0};
        for (int w = 0; w < kStatThreads / 32; ++w) s.flux += sf[w], s.tau_flux += stf[w], s.used += su[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void k_flux_sums_final(const FluxSums *__restrict__ partial, int nblocks, FluxSums *__restrict__ total)
{
    if (threadIdx.x || blockIdx.x) return;
    FluxSums s = {0, 0, 0};
    for (int b = 0; b < nblocks; ++b) s.flux += partial[b].flux, s.tau_flux += partial[b].tau_flux, s.used += partial[b].used;
    *total = s;
}

// numpy.histogram's uniform-bin rule (numpy/lib/_histograms_impl.py, the branch for equal-width bins):
// index from the scaled value, then corrected against the actual edges k/nbins; x == last edge goes into
// the last bin; x outside [0, 1] is dropped.
__global__ void __launch_bounds__(kStatThreads)
k_flux_hist(const double *__restrict__ tau, int64_t n, double scale, int nbins, unsigned long long *__restrict__ counts)
{
    extern __shared__ unsigned int hist[];
    __shared__ double etab[64];
    for (int b = threadIdx.x; b < nbins; b += kStatThreads) hist[b] = 0;
    stage_exp_table(etab);
    const double nb = (double) nbins;
    // shared counters are 32-bit: flush every 2^31 / threads pixels per thread at the latest (never reached:
    // a CTA sees n / gridDim pixels)
    for (int64_t i = (int64_t) blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (int64_t) gridDim.x * kStatThreads) {
        const double x = exp_nonpos(-scale * tau[i], etab);
        if (!(x >= 0.0 && x <= 1.0)) continue;
        int k = (int) (x * nb / 1.0);
        if (k == nbins) --k;
        if (x < (double) k / nb) --k;
        else if (k != nbins - 1 && x >= (double) (k + 1) / nb) ++k;
        atomicAdd(&hist[k], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += kStatThreads)
        if (hist[b]) atomicAdd(&counts[b], (unsigned long long) hist[b]);
}

__global__ void __launch_bounds__(kStatThreads)
k_delta_flux(const double *__restrict__ tau, int64_t n, double scale, double inv_mean, double *__restrict__ out)
{
    __shared__ double etab[64];
    stage_exp_table(etab);
    for (int64_t i = (int64_t) blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (int64_t) gridDim.x * kStatThreads)
        out[i] = exp_nonpos(-scale * tau[i], etab) * inv_mean - 1.0;
}

// partial[y][k] = sum over the y-th slab of spectra of |F[s][k]|^2, F interleaved (re, im) as produced by an
// rfft over pixels: coalesced across k, spectra in order inside a slab; k_power_final adds the slabs in order
// (deterministic) and accumulates factor * sum into power[k].
constexpr int kPowerSlabs = 128;

__global__ void __launch_bounds__(kStatThreads)
k_power_partial(const double2 *__restrict__ f, int64_t nspec, int nk, double *__restrict__ partial)
{
    const int k = blockIdx.x * kStatThreads + threadIdx.x;
    if (k >= nk) return;
    const int64_t per = (nspec + gridDim.y - 1) / gridDim.y;
    const int64_t s0 = (int64_t) blockIdx.y * per, s1 = min(nspec, s0 + per);
    double acc = 0;
    for (int64_t s = s0; s < s1; ++s) {
        const double2 v = f[s * nk + k];
        acc += fma(v.x, v.x, v.y * v.y);
    }
    partial[(int64_t) blockIdx.y * nk + k] = acc;
}

__global__ void __launch_bounds__(kStatThreads)
k_power_final(const double *__restrict__ partial, int nslabs, int nk, double factor, double *__restrict__ power)
{
    const int k = blockIdx.x * kStatThreads + threadIdx.x;
    if (k >= nk) return;
    double acc = 0;
    for (int y = 0; y < nslabs; ++y) acc += partial[(int64_t) y * nk + k];
    power[k] += factor * acc;
}

// ---- the flux power spectrum without a library FFT ---------------------------------------------------------------
// |rfft(x)|^2 of every sightline (fluxstatistics.py:54-61) for ANY pixel count n (n = int(vmax / res) is whatever the box
// and the pixel width make it: 4460 = 2^2 5 223, 8921 = 11 811, 1115 = 5 223), as a two-level discrete Fourier
// transform entirely in shared memory: n = n1 n2 with n1 >= n2 the divisor pair of smallest sum,
//     j = j1 n2 + j2,  k = k1 + n1 k2:
//     X[k] = sum_j2 W_n2^(j2 k2) { W_n^(j2 k1) sum_j1 x[j1 n2 + j2] W_n1^(j1 k1) },
// stage A (the braces) is n direct sums of length n1 over the REAL input (only k1 <= n1/2 is summed, the rest follows by
// conjugation, two k1 per thread sharing the input value), stage B n/2 + 1 direct sums of length n2 that apply the
// twiddle W_n^(j2 k1) on the way.  Cost n (n1 / 2 + n2 / 2) complex multiply-adds per
// sightline -- O(n^1.5) for composite n, the plain O(n^2) sum for prime n -- all of it FP64 FMAs on operands in
// shared memory, no index permutation passes, no workspace in HBM: 7e11 FMAs for 1e5 sightlines of 8921 pixels
// (40 ms at the FP64 peak).  One persistent CTA per SM takes sightlines in turn: delta_F = exp(-s tau)/<F> - 1 is
// formed on the fly (the input of the transform never exists in HBM) and |X|^2 is added to the CTA's own row of
// partial sums, which k_power_final adds in a fixed order (deterministic).
// Roots of unity: tables built on the device (sincospi of reduced arguments: exact to double rounding); the n1 roots of
// stage A are staged in shared memory.
constexpr int kDftThreads = 1024;

__global__ void k_dft_tables(int n, int n1, int n2, double2 *__restrict__ w1, double2 *__restrict__ w2, double2 *__restrict__ tw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto root = [](long long m, int len) {  // exp(-2 pi i m / len), 0 <= m < len
        double sn, cs;
        sincospi(2.0 * (double) m / (double) len, &sn, &cs);
        return make_double2(cs, -sn);
    };
    if (i < n1) w1[i] = root(i, n1);
    if (i < n2) w2[i] = root(i, n2);
    if (i < n) {
        const int k1 = i / n2, j2 = i - k1 * n2;
        tw[i] = root(((long long) j2 * k1) % n, n);
    }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }

// mode 0: x = exp(-scale in) inv_mean - 1 (in = optical depths); mode 1: x = in.
// per_row == nullptr: partial[blockIdx.x][k] += |X_k|^2; else per_row[s][k] = |X_k|^2 / n^2.
// Shared memory: x[n] | I[(n1/2 + 1) n2] (the inner sums of stage A for k1 <= n1/2; the other half is their conjugate)
// | the n1 roots of unity of stage A.  `flags` bit 0: I lives in global memory (i_global, one slice per CTA), bit 1: the
// roots are read from global memory (very long prime lengths).
template <bool I_SMEM, bool W_SMEM>
__global__ void __launch_bounds__(kDftThreads, 1)
k_flux_power(const double *__restrict__ in, int64_t nspec, int n, int n1, int n2, int mode, double scale, double inv_mean,
             const double2 *__restrict__ w1, const double2 *__restrict__ w2, const double2 *__restrict__ tw,
             double *__restrict__ partial, double *__restrict__ per_row, double2 *__restrict__ i_global)
{
    extern __shared__ __align__(16) double dft_smem[];
    __shared__ double etab[64];
    stage_exp_table(etab);
    const int nk = n / 2 + 1, h1 = n1 / 2, ni = (h1 + 1) * n2;
    double *x = dft_smem;  // [n]
    // the two placements are template parameters so that the common case compiles to shared-memory loads (a pointer
    // that may be either address space makes every access a generic load: 25 % of the stall samples when it was one)
    double2 *const i_smem = reinterpret_cast<double2 *>(dft_smem + n + (n & 1));
    double2 *const w_smem = i_smem + (I_SMEM ? ni : 0);
    double2 *const I = I_SMEM ? i_smem : i_global + (int64_t) blockIdx.x * ni;
    if (W_SMEM)
        for (int i = threadIdx.x; i < n1; i += kDftThreads) w_smem[i] = w1[i];
    const double2 *const wr = W_SMEM ? w_smem : w1;
    const double inv_n2 = 1.0 / ((double) n * (double) n);
    for (int64_t s = blockIdx.x; s < nspec; s += gridDim.x) {
        const double *row = in + s * n;
        for (int i = threadIdx.x; i < n; i += kDftThreads) x[i] = mode == 0 ? exp_nonpos(-scale * row[i], etab) * inv_mean - 1.0 : row[i];
        __syncthreads();
        // stage A: thread <-> (pair of k1, j2), j2 fastest: consecutive lanes read consecutive x, one x per two outputs,
        // the roots are (nearly) broadcasts
        const int npair = (h1 + 2) / 2;
        for (int item = threadIdx.x; item < npair * n2; item += kDftThreads) {
            const int kp = item / n2, j2 = item - kp * n2;
            const int ka = 2 * kp, kb = min(2 * kp + 1, h1);
            double ar = 0, ai = 0, br = 0, bi = 0;
            int ia = 0, ib = 0;
            const double *xp = x + j2;
            #pragma unroll 2
            for (int j1 = 0; j1 < n1; ++j1, xp += n2) {
                const double v = *xp;
                const double2 a = wr[ia], b = wr[ib];
                ar = fma(v, a.x, ar), ai = fma(v, a.y, ai);
                br = fma(v, b.x, br), bi = fma(v, b.y, bi);
                ia += ka;
                ia -= ia >= n1 ? n1 : 0;
                ib += kb;
                ib -= ib >= n1 ? n1 : 0;
            }
            I[ka * n2 + j2] = make_double2(ar, ai);
            I[kb * n2 + j2] = make_double2(br, bi);
        }
        __syncthreads();
        // stage B: thread <-> output k = k1 + n1 k2 <= n / 2
        for (int k = threadIdx.x; k < nk; k += kDftThreads) {
            const int k2 = k / n1, k1 = k - k2 * n1;
            const bool mirror = k1 > h1;
            const double2 *ip = I + (mirror ? n1 - k1 : k1) * n2;
            const double2 *tp = tw + k1 * n2;
            double2 acc = make_double2(0, 0);
            int idx = 0;
            for (int j2 = 0; j2 < n2; ++j2) {
                double2 v = ip[j2];
                v.y = mirror ? -v.y : v.y;
                const double2 t = cmul(cmul(v, __ldg(tp + j2)), __ldg(w2 + idx));
                acc.x += t.x, acc.y += t.y;
                idx += k2;
                idx -= idx >= n2 ? n2 : 0;
            }
            const double p = fma(acc.x, acc.x, acc.y * acc.y);
            if (per_row) per_row[s * nk + k] = p * inv_n2;
            else partial[(int64_t) blockIdx.x * nk + k] += p;
        }
        __syncthreads();
    }
}

// One wave of resident CTAs (grid-stride kernels): SMs x kStatBlocksPerSM, fewer for small inputs.
int stat_grid(int64_t n)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + kStatThreads * 4 - 1) / (kStatThreads * 4);
    return (int) std::max<int64_t>(1, std::min<int64_t>(std::min(kMaxStatBlocks, sms * kStatBlocksPerSM), want));
}

}  // namespace

}  // namespace fsb

using namespace fsb;

extern "C" int fsb_rescale_mean_flux(const double *tau, int64_t n, double mean_flux_desired, double tol, double thresh,
                                     double *scale_out, int32_t *iterations, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(scale_out != nullptr, "scale_out is NULL");
    FSB_REQUIRE(n >= 0, "negative n");
    if (iterations) *iterations = 0;
    if (n == 0) {  // fluxstatistics.py:39-40
        *scale_out = 0;
        return FSB_OK;
    }
    FSB_REQUIRE(tau != nullptr, "tau is NULL");
    const int grid = stat_grid(n);
    Scratch partial, total;
    FSB_TRY(partial.alloc(sizeof(FluxSums) * (size_t) grid, stream));
    FSB_TRY(total.alloc(sizeof(FluxSums), stream));
    // Newton-Raphson exactly as py_module.cpp:237-261; the sums come from the device each iteration
    double scale, newscale = 1;
    int it = 0;
    do {
        scale = newscale;
        count_launch(); k_flux_sums<<<grid, kStatThreads, 0, stream>>>(tau, n, scale, thresh, partial.as<FluxSums>());
        count_launch(); k_flux_sums_final<<<1, 32, 0, stream>>>(partial.as<FluxSums>(), grid, total.as<FluxSums>());
        FSB_CUDA_TRY(cudaGetLastError());
        FluxSums s;
        FSB_CUDA_TRY(cudaMemcpyAsync(&s, total.ptr, sizeof(s), cudaMemcpyDeviceToHost, stream));
        FSB_CUDA_TRY(cudaStreamSynchronize(stream));
        newscale = scale + (s.flux - mean_flux_desired * (double) s.used) / s.tau_flux;
        if (newscale <= 0) newscale = 1e-10;
        if (++it >= 1000) {
            set_error("fsb_rescale_mean_flux: no convergence after %d iterations (scale %g)", it, newscale);
            return FSB_EINVAL;
        }
    } while (fabs(newscale - scale) > tol * newscale);
    *scale_out = newscale;
    if (iterations) *iterations = it;
    return FSB_OK;
}

extern "C" int fsb_flux_sums(const double *tau, int64_t n, double scale, double thresh, double *sum_flux, double *sum_tau_flux,
                             int64_t *used, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(n >= 0 && sum_flux && sum_tau_flux && used, "bad arguments");
    *sum_flux = *sum_tau_flux = 0;
    *used = 0;
    if (n == 0) return FSB_OK;
    FSB_REQUIRE(tau != nullptr, "tau is NULL");
    const int grid = stat_grid(n);
    Scratch partial, total;
    FSB_TRY(partial.alloc(sizeof(FluxSums) * (size_t) grid, stream));
    FSB_TRY(total.alloc(sizeof(FluxSums), stream));
    count_launch(); k_flux_sums<<<grid, kStatThreads, 0, stream>>>(tau, n, scale, thresh, partial.as<FluxSums>());
    count_launch(); k_flux_sums_final<<<1, 32, 0, stream>>>(partial.as<FluxSums>(), grid, total.as<FluxSums>());
    FSB_CUDA_TRY(cudaGetLastError());
    FluxSums s;
    FSB_CUDA_TRY(cudaMemcpyAsync(&s, total.ptr, sizeof(s), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaStreamSynchronize(stream));
    *sum_flux = s.flux;
    *sum_tau_flux = s.tau_flux;
    *used = (int64_t) s.used;
    return FSB_OK;
}

// Row maxima of a [nrows][n] array: one warp per row (the damped-absorber cut of Spectra._filter_tau, spectra.py:1258).
__global__ void __launch_bounds__(256) k_row_max(const double *__restrict__ a, int64_t nrows, int64_t n, double *__restrict__ out)
{
    const int64_t row = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const double *r = a + row * n;
    double m = -INFINITY;
    for (int64_t j = threadIdx.x & 31; j < n; j += 32) m = fmax(m, r[j]);
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) out[row] = m;
}

extern "C" int fsb_row_max(const double *a, int64_t nrows, int64_t n, double *out, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(nrows >= 0 && n >= 1, "bad sizes");
    if (nrows == 0) return FSB_OK;
    FSB_REQUIRE(a != nullptr && out != nullptr, "NULL array");
    count_launch(); k_row_max<<<(unsigned) ((nrows + 7) / 8), 256, 0, stream>>>(a, nrows, n, out);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

extern "C" int fsb_flux_pdf(const double *tau, int64_t n, double scale, int32_t nbins, uint64_t *counts, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(nbins >= 1 && nbins <= 8192, "nbins must be in 1..8192");
    FSB_REQUIRE(counts != nullptr && n >= 0, "bad arguments");
    FSB_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * (size_t) nbins, stream));
    if (n == 0) return FSB_OK;
    FSB_REQUIRE(tau != nullptr, "tau is NULL");
    count_launch();
    k_flux_hist<<<stat_grid(n), kStatThreads, sizeof(unsigned int) * (size_t) nbins, stream>>>(
        tau, n, scale, nbins, reinterpret_cast<unsigned long long *>(counts));
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

extern "C" int fsb_delta_flux(const double *tau, int64_t n, double scale, double mean_flux, double *out, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(n >= 0, "negative n");
    if (n == 0) return FSB_OK;
    FSB_REQUIRE(tau != nullptr && out != nullptr, "NULL array");
    count_launch(); k_delta_flux<<<stat_grid(n), kStatThreads, 0, stream>>>(tau, n, scale, 1.0 / mean_flux, out);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

extern "C" int fsb_power_accumulate(const double *rfft_interleaved, int64_t nspec, int32_t nk, double factor, double *power,
                                    void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(nspec >= 0 && nk >= 1, "bad sizes");
    if (nspec == 0) return FSB_OK;
    FSB_REQUIRE(rfft_interleaved != nullptr && power != nullptr, "NULL array");
    const int nslabs = (int) std::min<int64_t>(kPowerSlabs, nspec);
    Scratch partial;
    FSB_TRY(partial.alloc(sizeof(double) * (size_t) nslabs * (size_t) nk, stream));
    const dim3 grid((nk + kStatThreads - 1) / kStatThreads, nslabs);
    count_launch();
    k_power_partial<<<grid, kStatThreads, 0, stream>>>(reinterpret_cast<const double2 *>(rfft_interleaved), nspec, nk, partial.as<double>());
    count_launch();
    k_power_final<<<grid.x, kStatThreads, 0, stream>>>(partial.as<double>(), nslabs, nk, factor, power);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

// Divisor pair n = n1 n2, n1 >= n2, of smallest sum.
static void dft_factors(int n, int &n1, int &n2)
{
    n2 = 1;
    for (int d = 1; (int64_t) d * d <= n; ++d)
        if (n % d == 0) n2 = d;
    n1 = n / n2;
}

// The 1-D flux power of fluxstatistics.flux_power (fluxstatistics.py:74-108) in one call, no FFT library:
//   mode 0: power[k] += factor * sum_s |rfft(exp(-scale tau[s]) / mean_flux - 1)[k]|^2      (power: DEVICE [npix/2+1])
//   mode 1: the same sum for in[s] used as it is (already a flux contrast)
//   per_row (DEVICE [nspec][npix/2+1], may be NULL): instead of summing, |rfft(x_s)|^2 / npix^2 per sightline
//            (fluxstatistics._powerspectrum, fluxstatistics.py:54-61); power may then be NULL.
extern "C" int fsb_flux_power(const double *in, int64_t nspec, int32_t npix, int32_t mode, double scale, double mean_flux,
                              double factor, double *power, double *per_row, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(nspec >= 0 && npix >= 1 && (mode == 0 || mode == 1), "bad arguments");
    if (nspec == 0) return FSB_OK;
    FSB_REQUIRE(in != nullptr && (power != nullptr || per_row != nullptr), "NULL array");
    FSB_REQUIRE(mode == 1 || mean_flux > 0, "mean flux must be positive");
    int n1, n2;
    dft_factors(npix, n1, n2);
    const int nk = npix / 2 + 1;
    int dev = 0, sms = 148, smem_max = 0;
    FSB_CUDA_TRY(cudaGetDevice(&dev));
    FSB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FSB_CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int grid = (int) std::min<int64_t>(nspec, sms);
    // shared memory: x always; the inner sums and the stage-A roots when they fit
    const size_t budget = (size_t) smem_max - 1024;
    const size_t ni = (size_t) (n1 / 2 + 1) * (size_t) n2;
    size_t smem = sizeof(double) * (size_t) (npix + (npix & 1));
    FSB_REQUIRE(smem <= budget, "too many pixels per sightline for the shared-memory transform");
    int flags = 0;
    if (smem + sizeof(double2) * ni <= budget) smem += sizeof(double2) * ni;
    else flags |= 1;
    if (smem + sizeof(double2) * (size_t) n1 <= budget) smem += sizeof(double2) * (size_t) n1;
    else flags |= 2;
    Scratch w1, w2, tw, partial, yglob;
    FSB_TRY(w1.alloc(sizeof(double2) * (size_t) n1, stream));
    FSB_TRY(w2.alloc(sizeof(double2) * (size_t) n2, stream));
    FSB_TRY(tw.alloc(sizeof(double2) * (size_t) npix, stream));
    if (flags & 1) FSB_TRY(yglob.alloc(sizeof(double2) * ni * (size_t) grid, stream));
    if (!per_row) {
        FSB_TRY(partial.alloc(sizeof(double) * (size_t) grid * (size_t) nk, stream));
        FSB_CUDA_TRY(cudaMemsetAsync(partial.ptr, 0, sizeof(double) * (size_t) grid * (size_t) nk, stream));
    }
    count_launch(); k_dft_tables<<<(npix + 255) / 256, 256, 0, stream>>>(npix, n1, n2, w1.as<double2>(), w2.as<double2>(), tw.as<double2>());
    auto go = [&](auto kern) -> int {
        FSB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        count_launch();
        kern<<<grid, kDftThreads, smem, stream>>>(in, nspec, npix, n1, n2, mode, scale, mode == 0 ? 1.0 / mean_flux : 1.0, w1.as<double2>(),
                                                  w2.as<double2>(), tw.as<double2>(), partial.as<double>(), per_row, yglob.as<double2>());
        FSB_CUDA_TRY(cudaGetLastError());
        return FSB_OK;
    };
    if (flags == 0) FSB_TRY(go(k_flux_power<true, true>));
    else if (flags == 1) FSB_TRY(go(k_flux_power<false, true>));
    else if (flags == 2) FSB_TRY(go(k_flux_power<true, false>));
    else FSB_TRY(go(k_flux_power<false, false>));
    if (!per_row) {
        count_launch();
        k_power_final<<<(nk + kStatThreads - 1) / kStatThreads, kStatThreads, 0, stream>>>(partial.as<double>(), grid, nk, factor, power);
        FSB_CUDA_TRY(cudaGetLastError());
    }
    return FSB_OK;
}
