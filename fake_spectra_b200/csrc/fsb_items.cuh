// Shared pieces of the accumulation kernels (fsb_tau.cu, fsb_colden.cu): SPH kernels and their line
// integrals, and the work-item table that splits long sightlines over several warps.
//
// Work decomposition (both kernels): one warp per work item = (sightline, contiguous run of its
// candidate list).  Each item owns its output row (the caller's row when a line is one item, a
// private scratch row otherwise), so accumulation needs no atomics and is bit-reproducible;
// scratch rows are summed in list order by k_reduce_rows (the deterministic segmented reduction).
#pragma once

#include <algorithm>

#include "fsb_common.cuh"
#include "fsb_scan.cuh"

namespace fsb {

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
    const int32_t *item_start;  // [nlos+1] first item of each line (NULL: one output row per line, no scratch rows)
    int32_t seg_pairs;
    // Ticketed segments (tau only, item_start == NULL): a line's list is cut into runs of ticket_pairs candidates that
    // are handed out as separate items, run-major (run 0 of every line, then run 1, ...); run s of a line waits until
    // line_done[line] == s, adds into the line's own row, then publishes s + 1.  The row therefore receives its
    // particles in exactly the order of an unsegmented pass (results are bit-identical whatever the number of
    // sightlines, warps or GPUs), while the scheduling granularity is a run instead of a whole sightline.
    //
    // Dispatch order: phases by REMAINING runs, most first.  Phase k (k = 0 .. max_runs - 1) holds, for every line with at
    // least r = max_runs - k runs, its run number nruns(line) - r; inside a phase the lines that start with this run
    // come first.  A line's runs keep their order (remaining runs fall from phase to phase), lines with long lists start
    // early, and the last phase is the last run of EVERY line: no thinly populated tail of phases at the end.
    int32_t ticket_pairs;       // 0: one item per line
    int32_t max_runs;           // runs of the longest list
    int32_t *line_done;         // [nlos] zero-initialised
    const int32_t *phase_start; // [max_runs + 1] first item of each phase; [max_runs] = total number of items
    const int32_t *order;       // [nlos] lines sorted by number of runs, descending
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

__device__ __forceinline__ int wrap_bin(int z, int nbins)
{
    int j = z % nbins;
    if (j < 0) j += nbins;
    return j;
}

// Work-item table for one launch.
struct ItemPlan {
    Scratch item_start, nitems, scratch_rows, line_done, phase_start, order;
    Items items;
    int64_t n_items = 0;  // upper bound on the number of items (= grid size)
    bool segmented = false;
};

// Chooses the segment length (0 = one item per sightline) and, when segmenting, builds the item
// table and the zeroed scratch rows ([nrows_per_item][n_items][nbins] doubles).
int plan_items(const fsb_index *idx, int seg_pairs_req, int nbins, int nrows_per_item, cudaStream_t stream, ItemPlan &plan);

// Switches an unsegmented plan to ticketed runs of `ticket` candidates (a multiple of the kernels' particle batch).
int plan_tickets(const fsb_index *idx, int ticket, int line0, int nrange, cudaStream_t stream, ItemPlan &plan);

// out[w][line][j] += sum over the line's items (in list order) of scratch[w][item][j]
int reduce_items(const ItemPlan &plan, const fsb_index *idx, int nbins, int nrows_per_item, double *out, cudaStream_t stream);

}  // namespace fsb
