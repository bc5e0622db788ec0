// Shared definitions for libfsb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "fsb200.h"

namespace fsb {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch();  // every kernel launch of this library is counted (fsb_kernel_launches)
int retain_pool_memory();  // raise the release threshold of the device's default memory pool (once per device)

#define FSB_CUDA_TRY(expr)                                                                     \
    do {                                                                                       \
        cudaError_t fsb_err_ = (expr);                                                         \
        if (fsb_err_ != cudaSuccess) {                                                         \
            fsb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                      \
                           cudaGetErrorString(fsb_err_));                                      \
            return fsb_err_ == cudaErrorMemoryAllocation ? FSB_ENOMEM : FSB_ECUDA;             \
        }                                                                                      \
    } while (0)

#define FSB_TRY(expr)                                                                          \
    do {                                                                                       \
        int fsb_rc_ = (expr);                                                                  \
        if (fsb_rc_ != FSB_OK) return fsb_rc_;                                                 \
    } while (0)

#define FSB_REQUIRE(cond, msg)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            fsb::set_error("%s: %s", __func__, msg);                                           \
            return FSB_EINVAL;                                                                 \
        }                                                                                      \
    } while (0)

// Stream-ordered scratch allocation that releases itself on scope exit.
struct Scratch {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    int alloc(size_t bytes, cudaStream_t s);
    template <typename T> T *as() const { return static_cast<T *>(ptr); }
    ~Scratch();
};

// ---- physical constants: absorption.cpp:21-26, absorption.h:4 -------------------------------
constexpr double kSigmaT = 6.652458558e-25;
constexpr double kBoltzmann = 1.3806504e-16;
constexpr double kLight = 2.99792458e10;
constexpr double kProtonMass = 1.67262178e-24;
constexpr double kPi = 3.14159265358979323846;
constexpr int kNGrid = 8;     // singleabs.h:8
constexpr double kReso = 0.1; // index_table.h:8

// ---- the candidate index --------------------------------------------------------------------
}  // namespace fsb

// Opaque to the C ABI.  All pointers are device memory owned by the index.
struct fsb_index {
    int32_t nlos = 0;
    int64_t npart = 0;
    int64_t npairs = 0;
    int64_t max_list = 0;
    double box = 0;
    int64_t *offsets = nullptr;  // [nlos+1]
    int32_t *particle = nullptr; // [npairs] ascending within a line
    double *dr2 = nullptr;       // [npairs]
    int32_t *zorder = nullptr;   // [npairs] traversal order of the optical-depth pass: each list's positions (absolute
                                 // pair indices) ordered by the particle's place along the sightline, 64 bins, stable
    double *cofm = nullptr;      // [nlos*3] private copy (the accumulation needs axis / cofm)
    int32_t *axis = nullptr;     // [nlos]
};

namespace fsb {

// Derived per-line constants handed to the accumulation kernels.
struct LineConsts {
    double sigma_a;   // absorption.cpp:154
    double voigt_fac; // absorption.cpp:156
};

struct InterpConsts {
    int32_t nbins;
    int32_t kernel;
    int32_t nlos;
    int32_t nlines; // tau: number of fused lines; colden: number of weight columns
    double box, velfac, vbox, bfac, tautail, bintov, boxtokpc;
    LineConsts line[4];
    int32_t voigt;
    int32_t seg_pairs;
    int32_t line0, nrange;  // sightlines [line0, line0 + nrange) of the index are processed (tau; default: all)
};

constexpr int kMaxFused = 4;

// Host destination of a tau launch: finished sightline rows are copied out on copy_stream while the kernel
// is still running (the kernel raises one flag in pinned host memory per chunk of sightlines).
struct HostSink {
    double *host = nullptr;       // same layout as the device output of this launch: [nlines][nlos][nbins]
    cudaStream_t copy_stream = nullptr;
    bool streamed = false;        // out: the launch copied its rows; false = the caller still has to
};

// launches implemented in the .cu files
int launch_tau(const fsb_index *idx, const InterpConsts &c, const float *pos, const float *vel, const float *dens,
               const float *temp, const float *h, const float *cells, double *out, fsb_counters *counters,
               int precision, cudaStream_t stream, HostSink *sink = nullptr, const fsb_push *push = nullptr);
int launch_colden(const fsb_index *idx, const InterpConsts &c, const float *pos, const float *dens, int64_t dens_stride,
                  const float *h, const float *cells, double *out, fsb_counters *counters, cudaStream_t stream);
int tau_max_fused_lines();
int launch_voigt(const double *x, const double *y, double *out, int64_t n, int voigt, cudaStream_t stream);

}  // namespace fsb
