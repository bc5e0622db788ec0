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
constexpr int kStatBlocksPerSM = 8;
constexpr int kMaxStatBlocks = 148 * kStatBlocksPerSM * 2;  // room for larger parts

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
__global__ void __launch_bounds__(kStatThreads)
k_flux_sums(const double *__restrict__ tau, int64_t n, double scale, double thresh, FluxSums *__restrict__ partial)
{
    const int64_t per_block = ((n + gridDim.x - 1) / gridDim.x + 1) & ~1ll;  // even: keeps double2 alignment
    const int64_t beg = (int64_t) blockIdx.x * per_block, end = min(n, beg + per_block);
    double f = 0, tf = 0;
    unsigned long long used = 0;
    auto add = [&](double t) {
        if (t > thresh) return;
        const double e = exp(-scale * t);
        f += e;
        tf = fma(e, t, tf);
        ++used;
    };
    const bool aligned = (reinterpret_cast<uintptr_t>(tau) & 15u) == 0;
    if (aligned) {
        const double2 *t2 = reinterpret_cast<const double2 *>(tau);
        for (int64_t i = beg / 2 + threadIdx.x; 2 * i + 1 < end; i += kStatThreads) {
            const double2 v = t2[i];
            add(v.x);
            add(v.y);
        }
        if (threadIdx.x == 0 && ((end - beg) & 1) && end > beg) add(tau[end - 1]);
    } else {
        for (int64_t i = beg + threadIdx.x; i < end; i += kStatThreads) add(tau[i]);
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
    for (int b = threadIdx.x; b < nbins; b += kStatThreads) hist[b] = 0;
    __syncthreads();
    const double nb = (double) nbins;
    // shared counters are 32-bit: flush every 2^31 / threads pixels per thread at the latest (never reached:
    // a CTA sees n / gridDim pixels)
    for (int64_t i = (int64_t) blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (int64_t) gridDim.x * kStatThreads) {
        const double x = exp(-scale * tau[i]);
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
    for (int64_t i = (int64_t) blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (int64_t) gridDim.x * kStatThreads)
        out[i] = exp(-scale * tau[i]) * inv_mean - 1.0;
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
