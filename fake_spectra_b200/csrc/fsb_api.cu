// C-ABI glue of libfsb200 (include/fsb200.h): argument validation, derived constants, the
// one-shot and host-pointer entry points that mirror the reference's Python boundary
// (py_module.cpp:25-233).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>

#include "fsb_common.cuh"

namespace fsb {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Keep freed blocks in the stream-ordered pool across calls: with the default release threshold (0)
// every synchronisation hands the multi-GB scratch and index arrays back to the driver and the
// next call pays for fresh allocations.
int retain_pool_memory()
{
    static thread_local int done_for_device = -1;
    int dev = 0;
    FSB_CUDA_TRY(cudaGetDevice(&dev));
    if (done_for_device == dev) return FSB_OK;
    cudaMemPool_t pool;
    FSB_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ull;
    FSB_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done_for_device = dev;
    return FSB_OK;
}

int Scratch::alloc(size_t bytes, cudaStream_t s)
{
    stream = s;
    FSB_TRY(retain_pool_memory());
    if (bytes == 0) bytes = 8;
    cudaError_t e = cudaMallocAsync(&ptr, bytes, s);
    if (e != cudaSuccess) {
        ptr = nullptr;
        set_error("cudaMallocAsync(%zu bytes): %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? FSB_ENOMEM : FSB_ECUDA;
    }
    return FSB_OK;
}

Scratch::~Scratch()
{
    if (ptr) cudaFreeAsync(ptr, stream);
}

// LineAbsorption ctor: absorption.cpp:152-161.
static int make_consts(const fsb_index *idx, const fsb_params *p, int nlines, InterpConsts &c)
{
    FSB_REQUIRE(idx != nullptr && p != nullptr, "NULL index or params");
    FSB_REQUIRE(nlines >= 1 && nlines <= kMaxFused, "between 1 and 4 fused lines / weight columns per call");
    FSB_REQUIRE(p->nbins > 0, "nbins must be positive");
    FSB_REQUIRE(p->kernel >= 0 && p->kernel <= 3, "kernel id must be 0..3 (singleabs.h:9-12)");
    FSB_REQUIRE(p->box > 0 && p->velfac > 0, "box and velfac must be positive");
    FSB_REQUIRE(p->box == idx->box, "params.box differs from the box the index was built with");
    FSB_REQUIRE(p->amumass > 0, "amumass must be positive");
    FSB_REQUIRE(p->gamma >= 0, "gamma must be non-negative");
    c.nbins = p->nbins;
    c.kernel = p->kernel;
    c.nlos = idx->nlos;
    c.nlines = nlines;
    c.box = p->box;
    c.velfac = p->velfac;
    c.vbox = p->box * p->velfac;
    c.bfac = sqrt(2.0 * kBoltzmann / (p->amumass * kProtonMass)) / 1e5;
    c.tautail = p->tautail;
    c.bintov = c.vbox / p->nbins;                 // absorption.cpp:239
    c.boxtokpc = c.vbox / p->nbins / p->velfac;   // absorption.cpp:188
    c.voigt = p->voigt;
    c.seg_pairs = p->seg_pairs;
    c.line0 = 0;
    c.nrange = idx->nlos;
    return FSB_OK;
}

static void line_consts(const fsb_params &p, LineConsts &l)
{
    l.sigma_a = sqrt(3.0 * kPi * kSigmaT / 8.0) * p.lambda_cm * p.fosc;
    l.voigt_fac = p.gamma * p.lambda_cm / (4. * kPi) / 1e5;
}

}  // namespace fsb

using namespace fsb;

extern "C" int fsb_abi_version(void) { return FSB_ABI_VERSION; }

extern "C" const char *fsb_last_error(void) { return g_err; }

extern "C" uint64_t fsb_kernel_launches(void) { return (uint64_t) g_launches.load(); }

extern "C" const char *fsb_strerror(int code)
{
    switch (code) {
    case FSB_OK: return "ok";
    case FSB_EINVAL: return "invalid argument";
    case FSB_ECUDA: return "CUDA error";
    case FSB_ENOMEM: return "out of device memory";
    case FSB_EVORONOI: return "Voronoi cell ownership is not contiguous along a sightline";
    case FSB_ENODEV: return "no usable CUDA device";
    default: return "unknown error";
    }
}

extern "C" int fsb_device_info(int32_t *sm_count, int32_t *clock_khz, int32_t *cc_major, int32_t *cc_minor)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return FSB_ENODEV;
    }
    int v = 0;
    if (sm_count) { FSB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (clock_khz) { FSB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev)); *clock_khz = v; }
    if (cc_major) { FSB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { FSB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return FSB_OK;
}

namespace {
// host == NULL: results stay on the device.  Otherwise every group of fused lines is also delivered to
// host[line][nlos][nbins]: streamed out while its kernel runs when the launch supports it, copied after it otherwise.
int compute_tau_multi_impl(const fsb_index *idx, const fsb_params *p, int32_t nlines, const float *pos, const float *vel,
                           const float *dens, const float *temp, const float *h, double *tau, fsb_counters *counters,
                           cudaStream_t stream, double *host, cudaStream_t copy_stream, const fsb_push *push = nullptr,
                           int32_t line_begin = 0, int32_t line_end = -1)
{
    FSB_REQUIRE(nlines >= 1, "nlines must be >= 1");
    FSB_REQUIRE(idx != nullptr && p != nullptr, "NULL index or params");
    FSB_REQUIRE(idx->npairs == 0 || (pos && vel && dens && temp && h && tau), "NULL array");
    FSB_REQUIRE(p[0].kernel != FSB_KERNEL_VORONOI, "Voronoi tau goes through fsb_particle_interpolate (needs fsb_assign_cells)");
    for (int32_t i = 0; i < nlines; ++i) {
        FSB_REQUIRE(p[i].nbins == p[0].nbins && p[i].kernel == p[0].kernel && p[i].box == p[0].box &&
                    p[i].velfac == p[0].velfac && p[i].amumass == p[0].amumass && p[i].tautail == p[0].tautail &&
                    p[i].voigt == p[0].voigt && p[i].precision == p[0].precision,
                    "fused lines must share nbins, kernel, box, velfac, amumass, tautail, voigt and precision");
    }
    // lines of one ion share every node position and Gaussian: the kernel takes them in groups
    const int32_t group = tau_max_fused_lines();
    for (int32_t i0 = 0; i0 < nlines; i0 += group) {
        const int32_t n = std::min(group, nlines - i0);
        InterpConsts c;
        FSB_TRY(make_consts(idx, &p[i0], n, c));
        for (int32_t k = 0; k < n; ++k) line_consts(p[i0 + k], c.line[k]);
        if (line_end >= 0) {
            FSB_REQUIRE(line_begin >= 0 && line_begin <= line_end && line_end <= idx->nlos, "sightline range outside the index");
            c.line0 = line_begin;
            c.nrange = line_end - line_begin;
            if (c.nrange == 0) continue;
        }
        const size_t off = (size_t) i0 * (size_t) idx->nlos * (size_t) c.nbins;
        // Host destination and many sightlines: the pass is cut into sightline ranges, and the rows of a finished range
        // leave on the copy stream while the next range is computed.  (The ticketed dispatch finishes EVERY row of a
        // launch in its last phase, so a whole-index launch has nothing to hand to the copy engine until it ends: the
        // 7-14 GB of a C3 pass were then copied after the kernel.)  Ranges of at least 4096 sightlines keep the
        // persistent kernel's tail below 2 %; one work item per sightline inside a range, so rows are bit-identical.
        const int nranges = (host && line_end < 0 && !counters && !push) ? std::min(8, idx->nlos / 4096) : 0;
        if (nranges >= 2) {
            InterpConsts cr = c;
            cr.seg_pairs = 1 << 30;
            cudaEvent_t done;
            FSB_CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
            int rc = FSB_OK;
            for (int r = 0; r < nranges && rc == FSB_OK; ++r) {
                const int l0 = (int) ((int64_t) idx->nlos * r / nranges), l1 = (int) ((int64_t) idx->nlos * (r + 1) / nranges);
                cr.line0 = l0;
                cr.nrange = l1 - l0;
                rc = launch_tau(idx, cr, pos, vel, dens, temp, h, nullptr, tau + off, nullptr, p[i0].precision, stream, nullptr, nullptr);
                cudaError_t e = rc == FSB_OK ? cudaEventRecord(done, stream) : cudaSuccess;
                if (e == cudaSuccess && rc == FSB_OK) e = cudaStreamWaitEvent(copy_stream, done, 0);
                for (int32_t k = 0; k < n && e == cudaSuccess && rc == FSB_OK; ++k) {
                    const size_t o = off + ((size_t) k * (size_t) idx->nlos + (size_t) l0) * (size_t) c.nbins;
                    e = cudaMemcpyAsync(host + o, tau + o, sizeof(double) * (size_t) (l1 - l0) * (size_t) c.nbins, cudaMemcpyDeviceToHost, copy_stream);
                }
                if (e != cudaSuccess) {
                    set_error("row delivery: %s", cudaGetErrorString(e));
                    rc = FSB_ECUDA;
                }
            }
            cudaEventDestroy(done);
            FSB_TRY(rc);
            continue;
        }
        HostSink sink;
        sink.host = host ? host + off : nullptr;
        sink.copy_stream = copy_stream;
        fsb_push gpush;
        if (push) {  // this group's lines start i0 lines into the destination arrays
            gpush = *push;
            for (int q = 0; q < gpush.npeers; ++q) gpush.dest[q] += (int64_t) i0 * gpush.line_stride;
        }
        FSB_TRY(launch_tau(idx, c, pos, vel, dens, temp, h, nullptr, tau + off, counters, p[i0].precision, stream,
                           host ? &sink : nullptr, push ? &gpush : nullptr));
        if (host && !sink.streamed) {
            const size_t bytes = sizeof(double) * (size_t) n * (size_t) idx->nlos * (size_t) c.nbins;
            cudaEvent_t done;
            FSB_CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
            cudaError_t e = cudaEventRecord(done, stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(copy_stream, done, 0);
            if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(host + off, tau + off, bytes, cudaMemcpyDeviceToHost, copy_stream);
            cudaEventDestroy(done);
            FSB_CUDA_TRY(e);
        }
    }
    return FSB_OK;
}
}  // namespace

extern "C" int fsb_compute_tau_multi(const fsb_index *idx, const fsb_params *p, int32_t nlines, const float *pos,
                                     const float *vel, const float *dens, const float *temp, const float *h, double *tau,
                                     fsb_counters *counters, void *stream_v)
{
    return compute_tau_multi_impl(idx, p, nlines, pos, vel, dens, temp, h, tau, counters, static_cast<cudaStream_t>(stream_v),
                                  nullptr, nullptr);
}

extern "C" int fsb_compute_tau_multi_push(const fsb_index *idx, const fsb_params *p, int32_t nlines, const float *pos,
                                          const float *vel, const float *dens, const float *temp, const float *h, double *tau,
                                          const fsb_push *push, void *stream_v)
{
    FSB_REQUIRE(push != nullptr && push->npeers >= 1 && push->npeers <= FSB_MAX_PEERS, "push: between 1 and 16 destination arrays");
    FSB_REQUIRE(nlines >= 1 && p != nullptr, "nlines must be >= 1");
    for (int q = 0; q < push->npeers; ++q) FSB_REQUIRE(push->dest[q] != nullptr, "push: NULL destination");
    // one work row per sightline, whatever the caller's seg_pairs: the rows that are pushed are the final rows
    fsb_params local[kMaxFused];
    FSB_REQUIRE(nlines <= kMaxFused, "at most 4 lines per call");
    for (int32_t i = 0; i < nlines; ++i) {
        local[i] = p[i];
        local[i].seg_pairs = 1 << 30;
    }
    return compute_tau_multi_impl(idx, local, nlines, pos, vel, dens, temp, h, tau, nullptr, static_cast<cudaStream_t>(stream_v),
                                  nullptr, nullptr, push);
}

extern "C" int fsb_compute_tau_multi_range(const fsb_index *idx, const fsb_params *p, int32_t nlines, int32_t line_begin,
                                           int32_t line_end, const float *pos, const float *vel, const float *dens,
                                           const float *temp, const float *h, double *tau, void *stream_v)
{
    FSB_REQUIRE(nlines >= 1 && nlines <= kMaxFused && p != nullptr, "between 1 and 4 lines per call");
    fsb_params local[kMaxFused];
    for (int32_t i = 0; i < nlines; ++i) {
        local[i] = p[i];
        local[i].seg_pairs = 1 << 30;  // one work row per sightline
    }
    return compute_tau_multi_impl(idx, local, nlines, pos, vel, dens, temp, h, tau, nullptr, static_cast<cudaStream_t>(stream_v),
                                  nullptr, nullptr, nullptr, line_begin, line_end);
}

extern "C" int fsb_peer_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle64)
{
    FSB_REQUIRE(dev_ptr != nullptr && handle64 != nullptr && bytes > 0, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    *dev_ptr = nullptr;
    FSB_CUDA_TRY(cudaMalloc(dev_ptr, (size_t) bytes));
    cudaIpcMemHandle_t hnd;
    const cudaError_t e = cudaIpcGetMemHandle(&hnd, *dev_ptr);
    if (e != cudaSuccess) {
        cudaFree(*dev_ptr);
        *dev_ptr = nullptr;
        set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return FSB_ECUDA;
    }
    memcpy(handle64, &hnd, sizeof(hnd));
    return FSB_OK;
}

extern "C" int fsb_peer_free(void *dev_ptr)
{
    if (dev_ptr) FSB_CUDA_TRY(cudaFree(dev_ptr));
    return FSB_OK;
}

extern "C" int fsb_peer_open(const unsigned char *handle64, void **dev_ptr)
{
    FSB_REQUIRE(dev_ptr != nullptr && handle64 != nullptr, "bad arguments");
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, sizeof(hnd));
    *dev_ptr = nullptr;
    FSB_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    return FSB_OK;
}

extern "C" int fsb_peer_close(void *dev_ptr)
{
    if (dev_ptr) FSB_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return FSB_OK;
}

extern "C" int fsb_compute_tau(const fsb_index *idx, const fsb_params *p, const float *pos, const float *vel,
                               const float *dens, const float *temp, const float *h, double *tau, fsb_counters *counters,
                               void *stream)
{
    return fsb_compute_tau_multi(idx, p, 1, pos, vel, dens, temp, h, tau, counters, stream);
}

extern "C" int fsb_compute_colden(const fsb_index *idx, const fsb_params *p, const float *pos, const float *dens,
                                  int32_t nweights, const float *h, double *colden, fsb_counters *counters, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(idx != nullptr && p != nullptr, "NULL index or params");
    FSB_REQUIRE(nweights >= 1, "nweights must be >= 1");
    FSB_REQUIRE(idx->npairs == 0 || (pos && dens && h && colden), "NULL array");
    FSB_REQUIRE(p->kernel != FSB_KERNEL_VORONOI, "Voronoi colden goes through fsb_particle_interpolate (needs fsb_assign_cells)");
    for (int32_t w0 = 0; w0 < nweights; w0 += kMaxFused) {
        const int nw = std::min<int32_t>(kMaxFused, nweights - w0);
        InterpConsts c;
        FSB_TRY(make_consts(idx, p, nw, c));
        FSB_TRY(launch_colden(idx, c, pos, dens + (size_t) w0 * (size_t) idx->npart, idx->npart, h, nullptr,
                              colden + (size_t) w0 * (size_t) idx->nlos * (size_t) c.nbins, counters, stream));
    }
    return FSB_OK;
}

extern "C" int fsb_particle_interpolate(int32_t compute_tau, const fsb_params *p, const float *pos, const float *vel,
                                        const float *dens, const float *temp, const float *h, int64_t npart,
                                        const int32_t *axis, const double *cofm, int32_t nlos, double *out, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(p != nullptr, "params is NULL");
    FSB_REQUIRE(p->nbins > 0, "nbins must be positive");
    fsb_index *idx = nullptr;
    FSB_TRY(fsb_index_build(p->box, cofm, axis, nlos, pos, h, npart, stream, &idx));
    int rc = FSB_OK;
    if (p->kernel == FSB_KERNEL_VORONOI) {
        Scratch cells;
        rc = cells.alloc(sizeof(float) * 2 * (size_t) std::max<int64_t>(idx->npairs, 1), stream);
        if (rc == FSB_OK) rc = fsb_assign_cells(idx, p->box, cofm, axis, pos, cells.as<float>(), stream);
        if (rc == FSB_OK) {
            InterpConsts c;
            rc = make_consts(idx, p, 1, c);
            if (rc == FSB_OK) {
                line_consts(*p, c.line[0]);
                rc = compute_tau ? launch_tau(idx, c, pos, vel, dens, temp, h, cells.as<float>(), out, nullptr, p->precision, stream)
                                 : launch_colden(idx, c, pos, dens, npart, h, cells.as<float>(), out, nullptr, stream);
            }
        }
    } else if (compute_tau) {
        rc = fsb_compute_tau(idx, p, pos, vel, dens, temp, h, out, nullptr, stream);
    } else {
        rc = fsb_compute_colden(idx, p, pos, dens, 1, h, out, nullptr, stream);
    }
    fsb_index_free(idx, stream);
    return rc;
}

namespace {
struct DevBuf {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    ~DevBuf() { if (ptr) cudaFreeAsync(ptr, stream); }
    int upload(const void *host, size_t bytes, cudaStream_t s)
    {
        stream = s;
        if (bytes == 0) bytes = 8, host = nullptr;
        FSB_CUDA_TRY(cudaMallocAsync(&ptr, bytes, s));
        if (host) FSB_CUDA_TRY(cudaMemcpyAsync(ptr, host, bytes, cudaMemcpyHostToDevice, s));
        return FSB_OK;
    }
};
}  // namespace

namespace {
struct OwnedStream {
    cudaStream_t s = nullptr;
    int create()
    {
        FSB_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        return FSB_OK;
    }
    ~OwnedStream() { if (s) cudaStreamDestroy(s); }
};
}  // namespace

namespace {
// The host entries: one upload of the particle arrays and one candidate index serve every line.  line_ion == NULL: all
// lines belong to one ion (dens is one column; or, for column density, nlines weight columns).  Otherwise (tau only)
// dens_columns[nions] point at one host array of npart densities per ion (wherever the caller holds them: nothing is
// gathered on the host) and line_ion[i] names the column of line i; lines of one ion must be consecutive.
int interpolate_host(int32_t compute_tau, const fsb_params *p, int32_t nlines, const int32_t *line_ion, int32_t nions,
                     const float *pos, const float *vel, const float *dens, const float *const *dens_columns, const float *temp,
                     const float *h, int64_t npart, const int32_t *axis, const double *cofm, int32_t nlos, double *out)
{
    FSB_REQUIRE(p != nullptr && out != nullptr, "params/out NULL");
    FSB_REQUIRE(nlines >= 1, "nlines must be >= 1");
    if (line_ion) {
        FSB_REQUIRE(compute_tau && nions >= 1, "several ions: optical depths only");
        FSB_REQUIRE(p[0].kernel != FSB_KERNEL_VORONOI, "several ions in one call: not for the Voronoi kernel");
        for (int32_t i = 0; i < nlines; ++i) {
            FSB_REQUIRE(line_ion[i] >= 0 && line_ion[i] < nions, "line_ion out of range");
            FSB_REQUIRE(i == 0 || line_ion[i] >= line_ion[i - 1], "lines of one ion must be consecutive (ascending line_ion)");
            FSB_REQUIRE(p[i].nbins == p[0].nbins && p[i].box == p[0].box && p[i].kernel == p[0].kernel, "lines must share nbins, box and kernel");
        }
    }
    FSB_REQUIRE(nlos >= 0 && npart >= 0 && p[0].nbins > 0, "bad sizes");
    FSB_TRY(retain_pool_memory());
    // own streams (declared first: destroyed after the buffers that are freed on them): compute, and a copy
    // stream that carries finished rows to the host while the tau kernel is still running
    OwnedStream compute, copy;
    FSB_TRY(compute.create());
    FSB_TRY(copy.create());
    cudaStream_t s = compute.s;
    struct DrainGuard {  // no copy from or to the caller's buffers may outlive this call, on any path
        cudaStream_t a, b;
        ~DrainGuard()
        {
            cudaStreamSynchronize(a);
            cudaStreamSynchronize(b);
        }
    } drain{compute.s, copy.s};
    DevBuf dpos, dvel, ddens, dtemp, dh, daxis, dcofm, dout;
    const size_t np = (size_t) npart, nl = (size_t) nlos;
    // what the candidate search needs goes first on the compute stream; the rest of the particle data is uploaded on
    // the copy stream while the index is being built, and joined before the accumulation kernels
    FSB_TRY(dpos.upload(pos, sizeof(float) * 3 * np, s));
    FSB_TRY(dh.upload(h, sizeof(float) * np, s));
    FSB_TRY(daxis.upload(axis, sizeof(int32_t) * nl, s));
    FSB_TRY(dcofm.upload(cofm, sizeof(double) * 3 * nl, s));
    // column density: nlines counts weight columns, dens is [nlines][npart]
    if (line_ion) {
        FSB_TRY(ddens.upload(nullptr, sizeof(float) * np * (size_t) nions, copy.s));
        for (int32_t k = 0; k < nions && np > 0; ++k) {
            FSB_REQUIRE(dens_columns && dens_columns[k], "NULL density column");
            FSB_CUDA_TRY(cudaMemcpyAsync((float *) ddens.ptr + (size_t) k * np, dens_columns[k], sizeof(float) * np, cudaMemcpyHostToDevice, copy.s));
        }
    } else {
        FSB_TRY(ddens.upload(dens, sizeof(float) * np * (compute_tau ? 1 : (size_t) nlines), copy.s));
    }
    if (compute_tau) {
        FSB_REQUIRE(npart == 0 || (vel && temp), "vel/temp NULL with compute_tau");
        FSB_TRY(dvel.upload(vel, sizeof(float) * 3 * np, copy.s));
        FSB_TRY(dtemp.upload(temp, sizeof(float) * np, copy.s));
    }
    cudaEvent_t uploaded;  // completion of the copy-stream uploads; the compute stream joins it after the index build
    FSB_CUDA_TRY(cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming));
    struct EventGuard {
        cudaEvent_t e;
        ~EventGuard() { cudaEventDestroy(e); }
    } uploaded_guard{uploaded};
    FSB_CUDA_TRY(cudaEventRecord(uploaded, copy.s));
    const size_t row_bytes = sizeof(double) * nl * (size_t) p[0].nbins;
    const size_t out_bytes = row_bytes * (size_t) nlines;
    FSB_TRY(dout.upload(nullptr, out_bytes, s));
    FSB_CUDA_TRY(cudaMemsetAsync(dout.ptr, 0, std::max<size_t>(out_bytes, 8), s));
    bool delivered = false;
    int rc = FSB_OK;
    if (p[0].kernel == FSB_KERNEL_VORONOI) {
        FSB_CUDA_TRY(cudaStreamWaitEvent(s, uploaded, 0));
        for (int32_t i = 0; i < nlines && rc == FSB_OK; ++i)
            rc = fsb_particle_interpolate(compute_tau, &p[compute_tau ? i : 0], (const float *) dpos.ptr, (const float *) dvel.ptr,
                                          (const float *) ddens.ptr + (compute_tau ? 0 : (size_t) i * np),
                                          (const float *) dtemp.ptr, (const float *) dh.ptr, npart,
                                          (const int32_t *) daxis.ptr, (const double *) dcofm.ptr, nlos,
                                          (double *) dout.ptr + (size_t) i * nl * (size_t) p[0].nbins, s);
    } else {
        fsb_index *idx = nullptr;
        rc = fsb_index_build(p[0].box, (const double *) dcofm.ptr, (const int32_t *) daxis.ptr, nlos, (const float *) dpos.ptr,
                             (const float *) dh.ptr, npart, s, &idx);
        if (rc == FSB_OK) {
            const cudaError_t ew = cudaStreamWaitEvent(s, uploaded, 0);  // velocities, densities, temperatures are in
            if (ew != cudaSuccess) {
                set_error("cudaStreamWaitEvent: %s", cudaGetErrorString(ew));
                rc = FSB_ECUDA;
            }
        }
        if (rc == FSB_OK) {
            if (compute_tau) {
                // one pass per ion: its lines are consecutive, its density column follows line_ion
                for (int32_t l0 = 0; l0 < nlines && rc == FSB_OK;) {
                    int32_t l1 = nlines;
                    if (line_ion)
                        for (l1 = l0 + 1; l1 < nlines && line_ion[l1] == line_ion[l0]; ++l1) {}
                    const size_t off = (size_t) l0 * nl * (size_t) p[0].nbins;
                    rc = compute_tau_multi_impl(idx, p + l0, l1 - l0, (const float *) dpos.ptr, (const float *) dvel.ptr,
                                                (const float *) ddens.ptr + (line_ion ? (size_t) line_ion[l0] * np : 0),
                                                (const float *) dtemp.ptr, (const float *) dh.ptr, (double *) dout.ptr + off, nullptr, s,
                                                out + off, copy.s);
                    l0 = l1;
                }
                delivered = rc == FSB_OK;
            } else {
                rc = fsb_compute_colden(idx, p, (const float *) dpos.ptr, (const float *) ddens.ptr, nlines,
                                        (const float *) dh.ptr, (double *) dout.ptr, nullptr, s);
            }
            fsb_index_free(idx, s);
        }
    }
    if (rc == FSB_OK && !delivered && out_bytes) {
        const cudaError_t e = cudaMemcpyAsync(out, dout.ptr, out_bytes, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { set_error("cudaMemcpyAsync (result): %s", cudaGetErrorString(e)); rc = FSB_ECUDA; }
    }
    // both streams drain before the buffers go out of scope, also on the error path
    const cudaError_t e1 = cudaStreamSynchronize(s), e2 = cudaStreamSynchronize(copy.s);
    if (rc == FSB_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        set_error("stream synchronize: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = FSB_ECUDA;
    }
    return rc;
}
}  // namespace

extern "C" int fsb_particle_interpolate_multi_host(int32_t compute_tau, const fsb_params *p, int32_t nlines,
                                                   const float *pos, const float *vel, const float *dens,
                                                   const float *temp, const float *h, int64_t npart, const int32_t *axis,
                                                   const double *cofm, int32_t nlos, double *out)
{
    return interpolate_host(compute_tau, p, nlines, nullptr, 1, pos, vel, dens, nullptr, temp, h, npart, axis, cofm, nlos, out);
}

extern "C" int fsb_particle_interpolate_ions_host(const fsb_params *p, int32_t nlines, const int32_t *line_ion, int32_t nions,
                                                  const float *pos, const float *vel, const float *const *dens_columns,
                                                  const float *temp, const float *h, int64_t npart, const int32_t *axis,
                                                  const double *cofm, int32_t nlos, double *out)
{
    FSB_REQUIRE(line_ion != nullptr && dens_columns != nullptr, "line_ion / dens_columns NULL");
    return interpolate_host(1, p, nlines, line_ion, nions, pos, vel, nullptr, dens_columns, temp, h, npart, axis, cofm, nlos, out);
}

extern "C" int fsb_particle_interpolate_host(int32_t compute_tau, const fsb_params *p, const float *pos, const float *vel,
                                             const float *dens, const float *temp, const float *h, int64_t npart,
                                             const int32_t *axis, const double *cofm, int32_t nlos, double *out)
{
    return fsb_particle_interpolate_multi_host(compute_tau, p, 1, pos, vel, dens, temp, h, npart, axis, cofm, nlos, out);
}

extern "C" int fsb_near_lines_host(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                                   const double *cofm, int32_t nlos, int32_t *out_index, int64_t *count)
{
    FSB_REQUIRE(count != nullptr, "count is NULL");
    *count = 0;
    FSB_REQUIRE(nlos >= 0 && npart >= 0, "negative size");
    if (npart == 0 || nlos == 0) return FSB_OK;
    cudaStream_t s = nullptr;
    DevBuf dpos, dh, daxis, dcofm, dout;
    FSB_TRY(dpos.upload(pos, sizeof(float) * 3 * (size_t) npart, s));
    FSB_TRY(dh.upload(h, sizeof(float) * (size_t) npart, s));
    FSB_TRY(daxis.upload(axis, sizeof(int32_t) * (size_t) nlos, s));
    FSB_TRY(dcofm.upload(cofm, sizeof(double) * 3 * (size_t) nlos, s));
    FSB_TRY(dout.upload(nullptr, sizeof(int32_t) * (size_t) npart, s));
    FSB_TRY(fsb_near_lines(box, (const float *) dpos.ptr, (const float *) dh.ptr, npart, (const int32_t *) daxis.ptr,
                           (const double *) dcofm.ptr, nlos, (int32_t *) dout.ptr, count, s));
    if (*count > 0) FSB_CUDA_TRY(cudaMemcpy(out_index, dout.ptr, sizeof(int32_t) * (size_t) *count, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

extern "C" int fsb_count_pairs_host(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                                    const double *cofm, int32_t nlos, int32_t *counts)
{
    FSB_REQUIRE(nlos >= 0 && npart >= 0, "negative size");
    if (nlos == 0) return FSB_OK;
    FSB_REQUIRE(counts != nullptr, "counts is NULL");
    cudaStream_t s = nullptr;
    DevBuf dpos, dh, daxis, dcofm, dout;
    FSB_TRY(dpos.upload(pos, sizeof(float) * 3 * (size_t) npart, s));
    FSB_TRY(dh.upload(h, sizeof(float) * (size_t) npart, s));
    FSB_TRY(daxis.upload(axis, sizeof(int32_t) * (size_t) nlos, s));
    FSB_TRY(dcofm.upload(cofm, sizeof(double) * 3 * (size_t) nlos, s));
    FSB_TRY(dout.upload(nullptr, sizeof(int32_t) * (size_t) nlos, s));
    FSB_TRY(fsb_count_pairs(box, (const float *) dpos.ptr, (const float *) dh.ptr, npart, (const int32_t *) daxis.ptr,
                            (const double *) dcofm.ptr, nlos, (int32_t *) dout.ptr, s));
    FSB_CUDA_TRY(cudaMemcpy(counts, dout.ptr, sizeof(int32_t) * (size_t) nlos, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

extern "C" int fsb_voigt_profile(const double *x, const double *y, double *out, int64_t n, int32_t voigt, void *stream)
{
    FSB_REQUIRE(n >= 0, "negative n");
    FSB_REQUIRE(n == 0 || (x && y && out), "NULL array");
    return launch_voigt(x, y, out, n, voigt, static_cast<cudaStream_t>(stream));
}
