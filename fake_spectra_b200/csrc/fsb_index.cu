// Candidate index: which particles reach which sightline (replaces IndexTable, index_table.cpp).
//
// The reference sorts sightlines by one perpendicular coordinate in two std::multimap and walks a
// key range per particle.  Here sightlines are binned on a 2-D grid over the plane perpendicular
// to their axis (one grid per axis value 1,2,3), particles are streamed once per pass with
// coalesced loads, and every (particle, line) pair that survives the grid walk is tested with the
// EXACT predicate of index_table.cpp:22-113 (same float/double mix, no FMA contraction), so the
// candidate sets are bit-identical to the reference's.  Passes:
//   lines:  classify -> per-cell counts -> scan -> scatter (cell-sorted line table)
//   pairs:  count per line (atomics) -> scan to int64 offsets -> fill -> per-line sort by
//           particle index (restores std::map order and run-to-run determinism) + dr^2
#include <algorithm>
#include <vector>

#include "fsb_common.cuh"
#include "fsb_scan.cuh"

namespace fsb {

namespace {

constexpr int kMaxGrid = 1024;

struct AxisGrid {
    int32_t G;          // cells per side (0 when the group has no lines)
    int32_t cell_base;  // offset of this group's cells in cell_start
    double inv_cs;      // G / box
};

struct LineTable {
    AxisGrid grid[3];          // axis 1, 2, 3
    const int32_t *cell_start; // [total_cells + 1] -> slot range in the arrays below
    const int32_t *line_id;    // [nlos] sightline index, cell-sorted
    const double *key;         // [nlos] primary perpendicular coordinate
    const double *proj2;       // [nlos] secondary perpendicular coordinate
    double box;
};

// Perpendicular coordinates of a sightline: (primary key, secondary), index_table.cpp:10-15,29-40.
__device__ __forceinline__ void line_coords(const double *cofm, int l, int ax, double &key, double &proj2)
{
    if (ax == 1) {
        key = cofm[3 * l + 1];
        proj2 = cofm[3 * l + 2];
    } else if (ax == 3) {
        key = cofm[3 * l];
        proj2 = cofm[3 * l + 1];
    } else {
        key = cofm[3 * l];
        proj2 = cofm[3 * l + 2];
    }
}

// Monotone map coordinate -> cell; the same function bins lines and bounds particle walks, so a
// line inside a coordinate interval is always inside the corresponding cell interval.
__device__ __forceinline__ int cell_of(double v, double inv_cs, int G)
{
    const double c = floor(v * inv_cs);
    if (!(c > 0.0)) return 0;  // also NaN
    if (c >= (double) G) return G - 1;
    return (int) c;
}

__device__ __forceinline__ int group_of_axis(int ax) { return ax == 1 ? 0 : (ax == 2 ? 1 : 2); }

__global__ void k_line_cells(const double *__restrict__ cofm, const int32_t *__restrict__ axis, int nlos,
                             AxisGrid g0, AxisGrid g1, AxisGrid g2, int32_t *__restrict__ cell_of_line,
                             int32_t *__restrict__ cell_count, int32_t *__restrict__ bad_axis)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlos) return;
    int ax = axis[l];
    if (ax < 1 || ax > 3) {  // reported by the caller at its next synchronisation: {line + 1, axis}
        if (atomicCAS(&bad_axis[0], 0, l + 1) == 0) bad_axis[1] = ax;
        ax = 1;
    }
    const int grp = group_of_axis(ax);
    const AxisGrid g = grp == 0 ? g0 : (grp == 1 ? g1 : g2);
    double key, proj2;
    line_coords(cofm, l, ax, key, proj2);
    const int cell = g.cell_base + cell_of(key, g.inv_cs, g.G) * g.G + cell_of(proj2, g.inv_cs, g.G);
    cell_of_line[l] = cell;
    atomicAdd(&cell_count[cell], 1);
}

__global__ void k_line_scatter(const double *__restrict__ cofm, const int32_t *__restrict__ axis, int nlos,
                               const int32_t *__restrict__ cell_of_line, const int32_t *__restrict__ cell_start,
                               int32_t *__restrict__ cursor, int32_t *__restrict__ line_id, double *__restrict__ key,
                               double *__restrict__ proj2)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlos) return;
    const int cell = cell_of_line[l];
    const int slot = cell_start[cell] + atomicAdd(&cursor[cell], 1);
    double k, p2;
    const int ax = axis[l];
    line_coords(cofm, l, (ax < 1 || ax > 3) ? 1 : ax, k, p2);
    line_id[slot] = l;
    key[slot] = k;
    proj2[slot] = p2;
}

// One axis group of one particle: walks the grid cells its kernel square can reach and applies the exact
// predicate of index_table.cpp:22-113 to the sightlines binned there.
//
// Cells are enumerated from the UNWRAPPED square [first - h - eps, first + h + eps] x [second - h - eps,
// second + h + eps] (eps = 1e-6 box, far above the float rounding of first +- h that the exact predicate
// sees), taken modulo the grid: a superset of the cells that hold a line passing the predicate, whatever
// side of the periodic box the particle's reach falls on (lines outside [0, box] sit in the edge cells,
// where cell_of clamps them, and those are the cells a wrapped reach lands in).
// Exact predicate of one (particle, axis group): every constant the comparisons of index_table.cpp:22-113 need,
// derived once per particle and group, and only when the cell walk found a sightline to test.
struct PairTest {
    double box, first, second, dffm, dffp, dsfm, dsfp, wrap_hi, wrap_lo, h2;
    bool wrapped, hi_wrap, lo_wrap;
    __device__ __forceinline__ void init(double box_, float first_f, float second_f, float h, double h2_)
    {
        box = box_;
        h2 = h2_;
        first = (double) first_f;
        second = (double) second_f;
        // B1, index_table.cpp:89-113: float add, wrap in double then round to float.
        float ffp = __fadd_rn(first_f, h);
        if ((double) ffp > box) ffp = __double2float_rn(__dsub_rn((double) ffp, box));
        float ffm = __fsub_rn(first_f, h);
        if (ffm < 0) ffm = __double2float_rn(__dadd_rn((double) ffm, box));
        wrapped = !(ffm <= ffp);
        dffm = (double) ffm;
        dffp = (double) ffp;
        // B2, index_table.cpp:52-68: wrap arithmetic stays in double here.
        const float sfp = __fadd_rn(second_f, h);
        const float sfm = __fsub_rn(second_f, h);
        dsfp = (double) sfp;
        dsfm = (double) sfm;
        hi_wrap = dsfp > box;
        lo_wrap = sfm < 0;
        wrap_hi = __dsub_rn(dsfp, box);  // lproj2 < sfp - box
        wrap_lo = __dadd_rn(dsfm, box);  // lproj2 > sfm + box
    }
    __device__ __forceinline__ bool operator()(double key, double lp2) const
    {
        // B1: lower_bound on both ends -> [ffm, ffp)
        const bool in1 = !wrapped ? (key >= dffm && key < dffp) : (key < dffp || key >= dffm);
        if (!in1) return false;
        bool in2 = false;
        if (hi_wrap && lp2 < wrap_hi) in2 = true;
        else if (lo_wrap && lp2 > wrap_lo) in2 = true;
        else in2 = (lp2 > dsfm && lp2 < dsfp);
        if (!in2) return false;
        // B3, index_table.cpp:70-87: separately rounded products and sum
        double d1 = fabs(__dsub_rn(first, key));
        if (d1 > 0.5 * box) d1 = __dsub_rn(box, d1);
        double d2 = fabs(__dsub_rn(second, lp2));
        if (d2 > 0.5 * box) d2 = __dsub_rn(box, d2);
        const double dr2 = __dadd_rn(__dmul_rn(d1, d1), __dmul_rn(d2, d2));
        return dr2 <= h2;
    }
};

// One axis group of one particle: walks the grid cells its kernel square can reach and applies the exact
// predicate to the sightlines binned there.
//
// Cells are enumerated from the UNWRAPPED square [first - h, first + h] x [second - h, second + h] widened by
// 1e-3 of a cell (far above the float rounding of first +- h that the exact predicate sees, and above the rounding of
// this single-precision cell arithmetic), taken modulo the grid: a superset of the cells that hold a line passing
// the predicate, whatever side of the periodic box the particle's reach falls on (lines outside [0, box] sit in the
// edge cells, where cell_of clamps them, and those are the cells a wrapped reach lands in).  With few sightlines
// most reached cells are empty and the particle costs two loads per grid row.
template <int MODE, int GRP>
__device__ __forceinline__ void pairs_in_group(const LineTable &T, const AxisGrid g, float px, float py, float pz, float h,
                                               int64_t p, int32_t *__restrict__ count,
                                               const int64_t *__restrict__ offsets, int32_t *__restrict__ particle, bool &any)
{
    const int G = g.G;
    if (G == 0) return;
    // axis 1: (y,z); axis 2: (x,z); axis 3: (x,y)   (index_table.cpp:29-40,120-125)
    const float first = GRP == 0 ? py : px;
    const float second = GRP == 2 ? py : pz;
    const float inv = (float) g.inv_cs, fG = (float) G;
    const float fr0 = floorf(fmaf(first - h, inv, -1e-3f)), fr1 = floorf(fmaf(first + h, inv, 1e-3f));
    const float fc0 = floorf(fmaf(second - h, inv, -1e-3f)), fc1 = floorf(fmaf(second + h, inv, 1e-3f));
    if (!(fr1 >= fr0) || !(fc1 >= fc0)) return;  // NaN coordinates reach nothing
    const int nrow = (int) fminf(fr1 - fr0 + 1.0f, fG), ncol = (int) fminf(fc1 - fc0 + 1.0f, fG);
    // first row / column modulo G (the common case needs one correction; anything farther out takes the division)
    int row = (int) fmaxf(fminf(fr0, 3.0e8f), -3.0e8f);
    row = (row >= -G && row < 2 * G) ? (row < 0 ? row + G : (row >= G ? row - G : row)) : ((row % G) + G) % G;
    int col0 = (int) fmaxf(fminf(fc0, 3.0e8f), -3.0e8f);
    col0 = (col0 >= -G && col0 < 2 * G) ? (col0 < 0 ? col0 + G : (col0 >= G ? col0 - G : col0)) : ((col0 % G) + G) % G;
    // the reach in columns: [col0, col0 + ncol) modulo G = one span, or two when it crosses the edge
    const int span1_hi = min(col0 + ncol, G) - 1, span2_hi = col0 + ncol - G - 1;  // span 2 = [0, span2_hi] when >= 0
    PairTest test;
    bool have_test = false;
    for (int rr = 0; rr < nrow; ++rr) {
        const int32_t *cs = T.cell_start + g.cell_base + row * G;
        row = row + 1 == G ? 0 : row + 1;
        #pragma unroll 1
        for (int sp = 0; sp < 2; ++sp) {
            if (sp == 1 && span2_hi < 0) break;
            const int beg = cs[sp == 0 ? col0 : 0];
            const int end = cs[(sp == 0 ? span1_hi : span2_hi) + 1];
            if (end > beg && !have_test) {
                test.init(T.box, first, second, h, (double) __fmul_rn(h, h));  // float product: index_table.cpp:45
                have_test = true;
            }
            for (int s = beg; s < end; ++s) {
                if (!test(T.key[s], T.proj2[s])) continue;
                if (MODE == 0) {
                    atomicAdd(&count[T.line_id[s]], 1);
                } else if (MODE == 1) {
                    const int l = T.line_id[s];
                    const int slot = atomicAdd(&count[l], 1);
                    particle[offsets[l] + slot] = (int32_t) p;
                } else {
                    any = true;
                    return;
                }
            }
        }
    }
}

// ---- bulk asynchronous copies (TMA, 1-D form) with an mbarrier in shared memory ---------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred done;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n\t"
        "@!done bra WAIT_%=;\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// MODE 0: count pairs per line.  MODE 1: fill the lists.  MODE 2: flag particles with >= 1 line.
//
// One CTA per tile of 256 particles.  The tile's positions ([256][3] floats, 3 KB) and smoothing lengths (1 KB) are
// contiguous in global memory: one elected thread fetches them with two bulk asynchronous copies (cp.async.bulk, the
// 1-D TMA path) into shared memory and arms an mbarrier with the byte count; the threads then pick their particle
// out of shared memory (a thread reading its own three floats from global memory would touch every sector three
// times, and staging through registers costs eight load/store instructions per thread).  The four CTAs resident on an
// SM overlap each other's copy with the sightline-grid walk, which is a chain of dependent loads.  A persistent variant
// with a two-stage ring in one CTA was built and dropped: the loop state pushed the walk over 64 registers (spills, or a
// resident CTA less).  The last tile, when its byte counts are not multiples of 16, and unaligned base pointers take
// plain coalesced loads.
constexpr int kPairTile = 256;

template <int MODE>
__global__ void __launch_bounds__(kPairTile) k_pairs(LineTable T, const float *__restrict__ pos, const float *__restrict__ hh,
                                                     int64_t npart, int32_t *__restrict__ count,
                                                     const int64_t *__restrict__ offsets, int32_t *__restrict__ particle,
                                                     uint8_t *__restrict__ flag)
{
    __shared__ __align__(128) float s_pos[3 * kPairTile];
    __shared__ __align__(128) float s_h[kPairTile];
    __shared__ __align__(8) uint64_t s_bar;
    const int64_t p0 = (int64_t) blockIdx.x * kPairTile, p = p0 + threadIdx.x;
    const int nhere = (int) min((int64_t) kPairTile, npart - p0);
    const bool bulk = nhere == kPairTile && ((reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(hh)) & 15u) == 0;
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&s_bar, 16 * kPairTile);
            bulk_load(s_pos, pos + 3 * p0, 12 * kPairTile, &s_bar);
            bulk_load(s_h, hh + p0, 4 * kPairTile, &s_bar);
        }
        __syncthreads();  // the barrier is initialised before anyone polls it
        mbar_wait(&s_bar, 0);
    } else {
        for (int i = threadIdx.x; i < 3 * nhere; i += kPairTile) s_pos[i] = pos[3 * p0 + i];
        if ((int) threadIdx.x < nhere) s_h[threadIdx.x] = hh[p];
        __syncthreads();
    }
    if ((int) threadIdx.x >= nhere) return;
    const float px = s_pos[3 * threadIdx.x], py = s_pos[3 * threadIdx.x + 1], pz = s_pos[3 * threadIdx.x + 2];
    const float h = s_h[threadIdx.x];
    bool any = false;
    pairs_in_group<MODE, 0>(T, T.grid[0], px, py, pz, h, p, count, offsets, particle, any);
    if (!(MODE == 2 && any)) pairs_in_group<MODE, 1>(T, T.grid[1], px, py, pz, h, p, count, offsets, particle, any);
    if (!(MODE == 2 && any)) pairs_in_group<MODE, 2>(T, T.grid[2], px, py, pz, h, p, count, offsets, particle, any);
    if (MODE == 2) flag[p] = any ? 1 : 0;
}

static unsigned pair_grid(int64_t npart) { return (unsigned) ((npart + kPairTile - 1) / kPairTile); }

// Squared periodic impact parameter of (particle, line): index_table.cpp:44,70-87.
__device__ __forceinline__ double pair_dr2(const float *__restrict__ pos, int64_t p, const double *__restrict__ cofm,
                                           int l, int ax, double box)
{
    double key, lp2;
    line_coords(cofm, l, ax, key, lp2);
    const float first = ax == 1 ? pos[3 * p + 1] : pos[3 * p];
    const float second = ax == 3 ? pos[3 * p + 1] : pos[3 * p + 2];
    double d1 = fabs(__dsub_rn((double) first, key));
    if (d1 > 0.5 * box) d1 = __dsub_rn(box, d1);
    double d2 = fabs(__dsub_rn((double) second, lp2));
    if (d2 > 0.5 * box) d2 = __dsub_rn(box, d2);
    return __dadd_rn(__dmul_rn(d1, d1), __dmul_rn(d2, d2));
}

// One CTA per sightline: bitonic sort of its particle list in shared memory (ascending particle
// index = std::map iteration order, part_int.cpp:35), then dr^2 for each entry.
//
// Then the traversal order of the optical-depth pass (zorder): the list's entries binned by the particle's
// coordinate along the sightline (kZBins bins), stably, i.e. ascending particle index inside a bin.  Consecutive
// particles of that order update overlapping pixel windows of the output row, which then stay in L1/L2 (in
// particle-index order every particle lands at a random place of a 71 KB row and 2368 concurrent rows thrash
// the L2: 326 GB of DRAM traffic per C2 launch).  A stable counting sort with thread-private counters: each of
// the 256 threads owns a contiguous chunk of the list, counts its entries per bin, one scan over (bin, thread)
// gives every thread its first slot per bin.  Deterministic, no atomics.
constexpr int kZBins = 64;

__global__ void __launch_bounds__(256) k_sort_lists(const int64_t *__restrict__ offsets, int32_t *__restrict__ particle,
                                                    double *__restrict__ dr2, int32_t *__restrict__ zorder,
                                                    const float *__restrict__ pos,
                                                    const double *__restrict__ cofm, const int32_t *__restrict__ axis,
                                                    double box, int cap /* power of two >= longest in-smem list */)
{
    extern __shared__ int32_t s_key[];
    __shared__ unsigned short s_cnt[kZBins * 256];
    __shared__ int s_tot[kZBins];
    static_assert(kZBins == 64, "the scan of the bin totals assumes 64 bins");
    const int l = blockIdx.x;
    const int64_t beg = offsets[l];
    const int n = (int) (offsets[l + 1] - beg);
    if (n == 0) return;
    if (n <= cap) {
        int m = 1;
        while (m < n) m <<= 1;
        for (int i = threadIdx.x; i < m; i += blockDim.x) s_key[i] = i < n ? particle[beg + i] : INT32_MAX;
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < m; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const int32_t a = s_key[i], b = s_key[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) {
                            s_key[i] = b;
                            s_key[ixj] = a;
                        }
                    }
                }
                __syncthreads();
            }
        const int ax = axis[l];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int32_t p = s_key[i];
            particle[beg + i] = p;
            dr2[beg + i] = pair_dr2(pos, p, cofm, l, ax, box);
        }
        // ---- zorder: stable counting sort of the entries by bin of the coordinate along the sightline
        const int t = threadIdx.x, lane = t & 31, w = t >> 5;
        const int chunk = (n + 255) / 256, i0 = min(n, t * chunk), i1 = min(n, i0 + chunk);
        const double to_bin = (double) kZBins / box;
        unsigned char *s_bin = reinterpret_cast<unsigned char *>(s_key + cap);  // [cap] bytes after the keys
        {
            uint32_t *z = reinterpret_cast<uint32_t *>(s_cnt);
            for (int e = t; e < kZBins * 256 / 2; e += 256) z[e] = 0;
        }
        __syncthreads();
        for (int i = i0; i < i1; ++i) {  // thread-private column t of the (bin, thread) counter matrix
            const int bb = min(kZBins - 1, max(0, (int) ((double) pos[3 * (int64_t) s_key[i] + (ax - 1)] * to_bin)));
            s_bin[i] = (unsigned char) bb;
            ++s_cnt[bb * 256 + t];
        }
        __syncthreads();
        // per bin: exclusive scan over the 256 threads (warp w takes bins w, w + 8, ...; a lane holds 8 threads' counts)
        for (int bb = w; bb < kZBins; bb += 8) {
            unsigned short *row = s_cnt + bb * 256 + lane * 8;
            int v[8], sum = 0;
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = row[j];
                sum += v[j];
            }
            int incl = sum;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            int excl = incl - sum;
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                row[j] = (unsigned short) excl;
                excl += v[j];
            }
            if (lane == 31) s_tot[bb] = incl;
        }
        __syncthreads();
        if (t < 32) {  // exclusive scan of the 64 bin totals
            const int a = s_tot[2 * t], b2 = s_tot[2 * t + 1];
            int incl = a + b2;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (t >= d) incl += up;
            }
            s_tot[2 * t] = incl - a - b2;
            s_tot[2 * t + 1] = incl - b2;
        }
        __syncthreads();
        for (int i = i0; i < i1; ++i) {
            const int bb = s_bin[i];
            const int slot = s_tot[bb] + s_cnt[bb * 256 + t]++;
            zorder[beg + slot] = (int32_t) (beg + i);
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) zorder[beg + i] = (int32_t) (beg + i);  // long list: list order
    }
}

// Lists longer than the shared-memory capacity (a single line through > 32768 particles; rare):
// one CTA runs the same network in global memory over a copy padded to a power of two.
__global__ void __launch_bounds__(1024) k_bitonic_global(int32_t *__restrict__ a, int64_t m)
{
    for (int64_t k = 2; k <= m; k <<= 1)
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
                const int64_t ixj = i ^ j;
                if (ixj > i) {
                    const int32_t x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
}

__global__ void k_dr2_of_list(const int32_t *__restrict__ particle, double *__restrict__ dr2, int64_t n,
                              const float *__restrict__ pos, const double *__restrict__ cofm,
                              const int32_t *__restrict__ axis, int line, double box)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dr2[i] = pair_dr2(pos, particle[i], cofm, line, axis[line], box);
}

struct BuiltTable {
    Scratch cell_start, line_id, key, proj2, bad_axis;
    LineTable T;
};

// Bin the sightlines of each axis group on its perpendicular grid.  The grid is sized by the PARTICLES' reach
// (cells of about two mean particle spacings: a particle then reaches a handful of cells whatever the number of
// sightlines, and with few sightlines most of those cells are empty), but never coarser than one sightline per
// cell.  No host synchronisation: an axis outside 1..3 is recorded in bad_axis (device) for check_axes().
int build_line_table(double box, const double *cofm, const int32_t *axis, int32_t nlos, int64_t npart, cudaStream_t stream,
                     BuiltTable &bt)
{
    int G = (int) llround(cbrt((double) std::max<int64_t>(npart, 1)) / 2.0);
    G = std::max(G, (int) ceil(sqrt((double) std::max(nlos, 1))));
    G = std::max(1, std::min(G, kMaxGrid));
    int32_t total_cells = 0;
    for (int g = 0; g < 3; ++g) {
        AxisGrid &ag = bt.T.grid[g];
        ag.G = G;
        ag.cell_base = total_cells;
        ag.inv_cs = (double) G / box;
        total_cells += G * G;
    }
    Scratch cell_of_line, cell_count, cursor;
    FSB_TRY(cell_of_line.alloc(sizeof(int32_t) * (size_t) std::max(nlos, 1), stream));
    FSB_TRY(cell_count.alloc(sizeof(int32_t) * (size_t) (total_cells + 1), stream));
    FSB_TRY(cursor.alloc(sizeof(int32_t) * (size_t) (total_cells + 1), stream));
    FSB_TRY(bt.cell_start.alloc(sizeof(int32_t) * (size_t) (total_cells + 2), stream));
    FSB_TRY(bt.line_id.alloc(sizeof(int32_t) * (size_t) std::max(nlos, 1), stream));
    FSB_TRY(bt.key.alloc(sizeof(double) * (size_t) std::max(nlos, 1), stream));
    FSB_TRY(bt.proj2.alloc(sizeof(double) * (size_t) std::max(nlos, 1), stream));
    FSB_TRY(bt.bad_axis.alloc(sizeof(int32_t) * 2, stream));
    FSB_CUDA_TRY(cudaMemsetAsync(bt.bad_axis.ptr, 0, sizeof(int32_t) * 2, stream));
    FSB_CUDA_TRY(cudaMemsetAsync(cell_count.ptr, 0, sizeof(int32_t) * (size_t) (total_cells + 1), stream));
    FSB_CUDA_TRY(cudaMemsetAsync(cursor.ptr, 0, sizeof(int32_t) * (size_t) (total_cells + 1), stream));
    if (nlos > 0) {
        const int threads = 256, blocks = (nlos + threads - 1) / threads;
        count_launch(); k_line_cells<<<blocks, threads, 0, stream>>>(cofm, axis, nlos, bt.T.grid[0], bt.T.grid[1], bt.T.grid[2],
                                                     cell_of_line.as<int32_t>(), cell_count.as<int32_t>(), bt.bad_axis.as<int32_t>());
        count_launch(); k_scan_single<int32_t, int32_t><<<1, 1024, 0, stream>>>(cell_count.as<int32_t>(), bt.cell_start.as<int32_t>(),
                                                                 total_cells, nullptr);
        count_launch(); k_line_scatter<<<blocks, threads, 0, stream>>>(cofm, axis, nlos, cell_of_line.as<int32_t>(),
                                                       bt.cell_start.as<int32_t>(), cursor.as<int32_t>(),
                                                       bt.line_id.as<int32_t>(), bt.key.as<double>(), bt.proj2.as<double>());
        FSB_CUDA_TRY(cudaGetLastError());
    } else {
        FSB_CUDA_TRY(cudaMemsetAsync(bt.cell_start.ptr, 0, sizeof(int32_t) * (size_t) (total_cells + 2), stream));
    }
    bt.T.cell_start = bt.cell_start.as<int32_t>();
    bt.T.line_id = bt.line_id.as<int32_t>();
    bt.T.key = bt.key.as<double>();
    bt.T.proj2 = bt.proj2.as<double>();
    bt.T.box = box;
    return FSB_OK;
}

// After the caller's synchronisation point: turns a recorded bad axis into FSB_EINVAL.
int check_axes(const int32_t h_bad[2])
{
    if (h_bad[0] != 0) {
        set_error("axis[%d] = %d: sightline axes are 1-based, 1..3 (spectra.py:681-683)", h_bad[0] - 1, h_bad[1]);
        return FSB_EINVAL;
    }
    return FSB_OK;
}

}  // namespace

}  // namespace fsb

using namespace fsb;

static int index_build_impl(fsb_index *idx, double box, const double *cofm, const int32_t *axis, int32_t nlos,
                            const float *pos, const float *h, int64_t npart, const int32_t *counts_in, cudaStream_t stream)
{
    const size_t nl = (size_t) std::max(nlos, 1);
    FSB_TRY(retain_pool_memory());
    FSB_CUDA_TRY(cudaMallocAsync(&idx->offsets, sizeof(int64_t) * (nl + 1), stream));
    FSB_CUDA_TRY(cudaMallocAsync(&idx->cofm, sizeof(double) * 3 * nl, stream));
    FSB_CUDA_TRY(cudaMallocAsync(&idx->axis, sizeof(int32_t) * nl, stream));
    if (nlos > 0) {
        FSB_CUDA_TRY(cudaMemcpyAsync(idx->cofm, cofm, sizeof(double) * 3 * (size_t) nlos, cudaMemcpyDeviceToDevice, stream));
        FSB_CUDA_TRY(cudaMemcpyAsync(idx->axis, axis, sizeof(int32_t) * (size_t) nlos, cudaMemcpyDeviceToDevice, stream));
    }

    BuiltTable bt;
    FSB_TRY(build_line_table(box, cofm, axis, nlos, npart, stream, bt));

    Scratch count, max_list;
    FSB_TRY(count.alloc(sizeof(int32_t) * (nl + 1), stream));
    FSB_TRY(max_list.alloc(sizeof(int64_t), stream));
    FSB_CUDA_TRY(cudaMemsetAsync(count.ptr, 0, sizeof(int32_t) * (nl + 1), stream));
    const int threads = 256;
    const unsigned pblocks = pair_grid(npart);
    // list sizes: counted here, or handed in by a caller that already ran fsb_count_pairs on these sightlines
    if (counts_in) {
        if (nlos > 0) FSB_CUDA_TRY(cudaMemcpyAsync(count.ptr, counts_in, sizeof(int32_t) * (size_t) nlos, cudaMemcpyDeviceToDevice, stream));
    } else if (npart > 0 && nlos > 0) {
        count_launch(); k_pairs<0><<<pblocks, threads, 0, stream>>>(bt.T, pos, h, npart, count.as<int32_t>(), nullptr, nullptr, nullptr);
        FSB_CUDA_TRY(cudaGetLastError());
    }
    count_launch(); k_scan_single<int32_t, int64_t><<<1, 1024, 0, stream>>>(count.as<int32_t>(), idx->offsets, nlos, max_list.as<int64_t>());
    FSB_CUDA_TRY(cudaGetLastError());
    int64_t h_total = 0, h_max = 0;
    int32_t h_bad[2] = {0, 0};
    FSB_CUDA_TRY(cudaMemcpyAsync(&h_total, idx->offsets + nlos, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaMemcpyAsync(&h_max, max_list.ptr, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaMemcpyAsync(h_bad, bt.bad_axis.ptr, sizeof(h_bad), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaStreamSynchronize(stream));
    FSB_TRY(check_axes(h_bad));
    idx->npairs = h_total;
    idx->max_list = h_max;

    const size_t np = (size_t) std::max<int64_t>(idx->npairs, 1);
    FSB_CUDA_TRY(cudaMallocAsync(&idx->particle, sizeof(int32_t) * np, stream));
    FSB_CUDA_TRY(cudaMallocAsync(&idx->dr2, sizeof(double) * np, stream));
    FSB_REQUIRE(idx->npairs <= (int64_t) INT32_MAX, "more than 2^31 candidate pairs in one index: split the sightlines");
    FSB_CUDA_TRY(cudaMallocAsync(&idx->zorder, sizeof(int32_t) * np, stream));
    if (idx->npairs == 0) return FSB_OK;

    FSB_CUDA_TRY(cudaMemsetAsync(count.ptr, 0, sizeof(int32_t) * (nl + 1), stream));
    count_launch(); k_pairs<1><<<pblocks, threads, 0, stream>>>(bt.T, pos, h, npart, count.as<int32_t>(), idx->offsets, idx->particle, nullptr);
    FSB_CUDA_TRY(cudaGetLastError());
    // in-smem sort capacity: next power of two of the longest list, at most 32768 entries (128 KB)
    int cap = 32;
    while (cap < idx->max_list && cap < 32768) cap <<= 1;
    const size_t smem = sizeof(int32_t) * (size_t) cap + (size_t) cap;  // sort keys + one bin byte per entry
    FSB_CUDA_TRY(cudaFuncSetAttribute(k_sort_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    count_launch(); k_sort_lists<<<nlos, 256, smem, stream>>>(idx->offsets, idx->particle, idx->dr2, idx->zorder, pos, idx->cofm, idx->axis, box, cap);
    FSB_CUDA_TRY(cudaGetLastError());
    if (idx->max_list > cap) {
        std::vector<int64_t> h_off((size_t) nlos + 1);
        FSB_CUDA_TRY(cudaMemcpyAsync(h_off.data(), idx->offsets, sizeof(int64_t) * ((size_t) nlos + 1), cudaMemcpyDeviceToHost, stream));
        FSB_CUDA_TRY(cudaStreamSynchronize(stream));
        for (int32_t l = 0; l < nlos; ++l) {
            const int64_t n = h_off[l + 1] - h_off[l];
            if (n <= cap) continue;
            int64_t m = 1;
            while (m < n) m <<= 1;
            Scratch pad;
            FSB_TRY(pad.alloc(sizeof(int32_t) * (size_t) m, stream));
            FSB_CUDA_TRY(cudaMemsetAsync(pad.ptr, 0x7f, sizeof(int32_t) * (size_t) m, stream));  // 0x7f7f7f7f > any index
            FSB_CUDA_TRY(cudaMemcpyAsync(pad.ptr, idx->particle + h_off[l], sizeof(int32_t) * (size_t) n, cudaMemcpyDeviceToDevice, stream));
            count_launch(); k_bitonic_global<<<1, 1024, 0, stream>>>(pad.as<int32_t>(), m);
            FSB_CUDA_TRY(cudaMemcpyAsync(idx->particle + h_off[l], pad.ptr, sizeof(int32_t) * (size_t) n, cudaMemcpyDeviceToDevice, stream));
            count_launch(); k_dr2_of_list<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(idx->particle + h_off[l], idx->dr2 + h_off[l], n, pos,
                                                                           idx->cofm, idx->axis, l, box);
            FSB_CUDA_TRY(cudaGetLastError());
        }
    }
    return FSB_OK;
}

extern "C" int fsb_index_build_counted(double box, const double *cofm, const int32_t *axis, int32_t nlos, const float *pos,
                                       const float *h, int64_t npart, const int32_t *counts, void *stream_v, fsb_index **out)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    FSB_REQUIRE(nlos >= 0 && npart >= 0, "negative size");
    FSB_REQUIRE(npart <= (int64_t) INT32_MAX, "npart exceeds int32 particle indices (index_table.cpp:145 uses int)");
    FSB_REQUIRE(box > 0, "box must be positive");
    FSB_REQUIRE(nlos == 0 || (cofm && axis), "cofm/axis NULL");
    FSB_REQUIRE(npart == 0 || (pos && h), "pos/h NULL");
    fsb_index *idx = new fsb_index();
    idx->nlos = nlos;
    idx->npart = npart;
    idx->box = box;
    const int rc = index_build_impl(idx, box, cofm, axis, nlos, pos, h, npart, counts, stream);
    if (rc != FSB_OK) {
        fsb_index_free(idx, stream);
        return rc;
    }
    *out = idx;
    return FSB_OK;
}

extern "C" int fsb_index_build(double box, const double *cofm, const int32_t *axis, int32_t nlos, const float *pos,
                               const float *h, int64_t npart, void *stream_v, fsb_index **out)
{
    return fsb_index_build_counted(box, cofm, axis, nlos, pos, h, npart, nullptr, stream_v, out);
}

extern "C" int fsb_index_free(fsb_index *idx, void *stream_v)
{
    if (!idx) return FSB_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    void *ptrs[] = {idx->offsets, idx->particle, idx->dr2, idx->zorder, idx->cofm, idx->axis};
    for (void *p : ptrs)
        if (p) cudaFreeAsync(p, stream);
    delete idx;
    return FSB_OK;
}

extern "C" int fsb_index_sizes(const fsb_index *idx, int32_t *nlos, int64_t *npairs, int64_t *max_list)
{
    FSB_REQUIRE(idx != nullptr, "index is NULL");
    if (nlos) *nlos = idx->nlos;
    if (npairs) *npairs = idx->npairs;
    if (max_list) *max_list = idx->max_list;
    return FSB_OK;
}

extern "C" int fsb_index_export(const fsb_index *idx, int64_t *offsets, int32_t *particle, double *dr2, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(idx != nullptr, "index is NULL");
    if (offsets)
        FSB_CUDA_TRY(cudaMemcpyAsync(offsets, idx->offsets, sizeof(int64_t) * ((size_t) idx->nlos + 1), cudaMemcpyDeviceToDevice, stream));
    if (particle && idx->npairs > 0)
        FSB_CUDA_TRY(cudaMemcpyAsync(particle, idx->particle, sizeof(int32_t) * (size_t) idx->npairs, cudaMemcpyDeviceToDevice, stream));
    if (dr2 && idx->npairs > 0)
        FSB_CUDA_TRY(cudaMemcpyAsync(dr2, idx->dr2, sizeof(double) * (size_t) idx->npairs, cudaMemcpyDeviceToDevice, stream));
    return FSB_OK;
}

extern "C" int fsb_near_lines(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                              const double *cofm, int32_t nlos, int32_t *out_index, int64_t *count, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(count != nullptr, "count is NULL");
    *count = 0;
    FSB_REQUIRE(nlos >= 0 && npart >= 0, "negative size");
    FSB_REQUIRE(npart <= (int64_t) INT32_MAX, "npart exceeds int32 particle indices");
    FSB_REQUIRE(box > 0, "box must be positive");
    if (npart == 0 || nlos == 0) return FSB_OK;
    FSB_REQUIRE(pos && h && axis && cofm && out_index, "NULL array");
    BuiltTable bt;
    FSB_TRY(build_line_table(box, cofm, axis, nlos, npart, stream, bt));
    Scratch flag, block_count, block_start;
    const int threads = 1024;
    const int64_t nblocks = (npart + threads - 1) / threads;
    FSB_TRY(flag.alloc((size_t) npart, stream));
    FSB_TRY(block_count.alloc(sizeof(int32_t) * (size_t) (nblocks + 1), stream));
    FSB_TRY(block_start.alloc(sizeof(int64_t) * (size_t) (nblocks + 1), stream));
    count_launch(); k_pairs<2><<<pair_grid(npart), kPairTile, 0, stream>>>(bt.T, pos, h, npart, nullptr, nullptr, nullptr, flag.as<uint8_t>());
    count_launch(); k_flag_block_counts<<<(unsigned) nblocks, threads, 0, stream>>>(flag.as<uint8_t>(), npart, block_count.as<int32_t>());
    count_launch(); k_scan_single<int32_t, int64_t><<<1, 1024, 0, stream>>>(block_count.as<int32_t>(), block_start.as<int64_t>(), nblocks, nullptr);
    count_launch(); k_flag_compact<<<(unsigned) nblocks, threads, 0, stream>>>(flag.as<uint8_t>(), npart, block_start.as<int64_t>(), nullptr, out_index);
    FSB_CUDA_TRY(cudaGetLastError());
    int32_t h_bad[2] = {0, 0};
    FSB_CUDA_TRY(cudaMemcpyAsync(h_bad, bt.bad_axis.ptr, sizeof(h_bad), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaMemcpyAsync(count, block_start.as<int64_t>() + nblocks, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaStreamSynchronize(stream));
    FSB_TRY(check_axes(h_bad));
    return FSB_OK;
}

// Candidate pairs per sightline without building the lists: the "cheap count pass" that balances sightline
// blocks across GPUs.  counts[nlos] int32, DEVICE, overwritten.
extern "C" int fsb_count_pairs(double box, const float *pos, const float *h, int64_t npart, const int32_t *axis,
                               const double *cofm, int32_t nlos, int32_t *counts, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(nlos >= 0 && npart >= 0, "negative size");
    FSB_REQUIRE(npart <= (int64_t) INT32_MAX, "npart exceeds int32 particle indices");
    FSB_REQUIRE(box > 0, "box must be positive");
    if (nlos == 0) return FSB_OK;
    FSB_REQUIRE(counts != nullptr && axis && cofm, "NULL array");
    FSB_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t) nlos, stream));
    if (npart == 0) return FSB_OK;
    FSB_REQUIRE(pos && h, "NULL array");
    BuiltTable bt;
    FSB_TRY(build_line_table(box, cofm, axis, nlos, npart, stream, bt));
    count_launch(); k_pairs<0><<<pair_grid(npart), kPairTile, 0, stream>>>(bt.T, pos, h, npart, counts, nullptr, nullptr, nullptr);
    FSB_CUDA_TRY(cudaGetLastError());
    // the line table is released when this function returns: finish the pass first (also reports a bad axis)
    int32_t h_bad[2] = {0, 0};
    FSB_CUDA_TRY(cudaMemcpyAsync(h_bad, bt.bad_axis.ptr, sizeof(h_bad), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaStreamSynchronize(stream));
    return check_axes(h_bad);
}
