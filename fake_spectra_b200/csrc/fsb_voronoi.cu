// Voronoi cell extents along each sightline (kernel id 2): replaces IndexTable::assign_cells,
// index_table.cpp:152-223.
//
// The reference marches N = int(box/0.1) points along the line, finds the nearest candidate for
// each (strict <, lowest candidate wins ties) and grows [lo, hi] of the owning cell point by point,
// with a wrap rule that ends the march early and an exit(1) when a cell's ownership is not
// contiguous.  Here the march is split in two data-parallel passes:
//   k_owners : nearest candidate of every march point (candidates tiled through shared memory,
//              several points per thread, same distance arithmetic without FMA contraction);
//   k_extents: i* = first point past box/2 owned again by the owner of point 0 (the reference's
//              break, :196-201); per-cell first/last owned point among points < i* from run
//              boundaries; more than one run for a cell = the reference's exit(1) (:204-208).
#include <algorithm>
#include <vector>

#include "fsb_common.cuh"

namespace fsb {
namespace {

constexpr int kOwnThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kCandTile = 512;

__global__ void __launch_bounds__(kOwnThreads) k_owners(const int64_t *__restrict__ offsets, const int32_t *__restrict__ particle,
                                                        const double *__restrict__ cofm, const int32_t *__restrict__ axis,
                                                        const float *__restrict__ pos, double box, int npts, double reso,
                                                        int line0, int32_t *__restrict__ owner /* [lines_in_batch][npts] */)
{
    __shared__ double s_x[kCandTile], s_dy2[kCandTile], s_dz2[kCandTile];
    const int line = line0 + blockIdx.y;
    const int64_t beg = offsets[line];
    const int ncells = (int) (offsets[line + 1] - beg);
    const int ax = axis[line];
    // index_table.cpp:168: yp = cofm[3l + ax%3], zp = cofm[3l + (ax+1)%3]
    const double yp = cofm[3 * line + ax % 3], zp = cofm[3 * line + (ax + 1) % 3];
    const double halfbox = box / 2.;

    double xp[kPtsPerThread], best[kPtsPerThread];
    int best_i[kPtsPerThread];
    #pragma unroll
    for (int k = 0; k < kPtsPerThread; ++k) {
        const int i = (blockIdx.x * kPtsPerThread + k) * kOwnThreads + threadIdx.x;
        xp[k] = ((double) i + 0.5) * reso;
        best[k] = __dmul_rn(box, box);  // min_dist = boxsize (:176); no periodic distance reaches it
        best_i[k] = 0;
    }
    for (int c0 = 0; c0 < ncells; c0 += kCandTile) {
        const int nt = min(kCandTile, ncells - c0);
        __syncthreads();
        for (int c = threadIdx.x; c < nt; c += kOwnThreads) {
            const int64_t ip = particle[beg + c0 + c];
            s_x[c] = (double) pos[3 * ip + ax - 1];
            double dy = fabs(__dsub_rn((double) pos[3 * ip + ax % 3], yp));
            if (dy > halfbox) dy = __dsub_rn(box, dy);
            double dz = fabs(__dsub_rn((double) pos[3 * ip + (ax + 1) % 3], zp));
            if (dz > halfbox) dz = __dsub_rn(box, dz);
            s_dy2[c] = __dmul_rn(dy, dy);
            s_dz2[c] = __dmul_rn(dz, dz);
        }
        __syncthreads();
        for (int c = 0; c < nt; ++c) {
            const double cx = s_x[c], dy2 = s_dy2[c], dz2 = s_dz2[c];
            #pragma unroll
            for (int k = 0; k < kPtsPerThread; ++k) {
                double dx = fabs(__dsub_rn(cx, xp[k]));
                if (dx > halfbox) dx = __dsub_rn(box, dx);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), dy2), dz2);
                if (d2 < best[k]) {
                    // The reference compares sqrt(d2) with strict <; two squared distances a few
                    // ulp apart can round to the same sqrt, in which case the earlier one stays.
                    if (d2 < best[k] * (1.0 - 1e-15) || sqrt(d2) < sqrt(best[k])) {
                        best[k] = d2;
                        best_i[k] = c0 + c;
                    }
                }
            }
        }
    }
    #pragma unroll
    for (int k = 0; k < kPtsPerThread; ++k) {
        const int i = (blockIdx.x * kPtsPerThread + k) * kOwnThreads + threadIdx.x;
        if (i < npts) owner[(int64_t) blockIdx.y * npts + i] = best_i[k];
    }
}

// One CTA per sightline of the batch.
__global__ void __launch_bounds__(1024) k_extents(const int64_t *__restrict__ offsets, double box, int npts, double reso,
                                                  int line0, const int32_t *__restrict__ owner,
                                                  int32_t *__restrict__ first_pt, int32_t *__restrict__ last_pt,
                                                  int32_t *__restrict__ nruns, float *__restrict__ cells,
                                                  int32_t *__restrict__ error_flag)
{
    __shared__ int s_istar;
    const int line = line0 + blockIdx.x;
    const int64_t beg = offsets[line];
    const int ncells = (int) (offsets[line + 1] - beg);
    if (ncells == 0) return;
    const int32_t *own = owner + (int64_t) blockIdx.x * npts;
    int32_t *first = first_pt + beg, *last = last_pt + beg, *runs = nruns + beg;
    for (int c = threadIdx.x; c < ncells; c += blockDim.x) {
        first[c] = INT32_MAX;
        last[c] = -1;
        runs[c] = 0;
    }
    if (threadIdx.x == 0) s_istar = npts;
    __syncthreads();
    const int own0 = own[0];
    const double thresh = box / 2. + 0.5 * reso;
    for (int i = threadIdx.x; i < npts; i += blockDim.x) {
        const double xp = ((double) i + 0.5) * reso;
        if (own[i] == own0 && xp > thresh) atomicMin(&s_istar, i);
    }
    __syncthreads();
    const int istar = s_istar;
    for (int i = threadIdx.x; i < istar; i += blockDim.x) {
        const int c = own[i];
        if (i == 0 || own[i - 1] != c) {
            atomicAdd(&runs[c], 1);
            atomicMin(&first[c], i);
        }
        if (i == istar - 1 || own[i + 1] != c) atomicMax(&last[c], i);
    }
    __syncthreads();
    const float sentinel = (float) (3 * box);
    for (int c = threadIdx.x; c < ncells; c += blockDim.x) {
        float lo = sentinel, hi = sentinel;
        if (runs[c] > 0) {
            lo = (float) (((double) first[c] + 0.5) * reso);
            hi = (float) (((double) last[c] + 0.5) * reso);
            if (runs[c] > 1) atomicExch(error_flag, 1);  // the reference's exit(1), :204-208
        }
        if (c == own0 && istar < npts) {  // wrap rule, :196-201
            lo = (float) (((double) istar + 0.5) * reso);
            hi = (float) ((double) hi + box);
        }
        cells[2 * (beg + c)] = (float) ((double) lo - 0.5 * reso);       // :217-220
        cells[2 * (beg + c) + 1] = (float) ((double) hi + 0.5 * reso);
    }
}

}  // namespace
}  // namespace fsb

using namespace fsb;

extern "C" int fsb_assign_cells(const fsb_index *idx, double box, const double *cofm, const int32_t *axis,
                                const float *pos, float *cells, void *stream_v)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    FSB_REQUIRE(idx != nullptr, "index is NULL");
    FSB_REQUIRE(box > 0 && box == idx->box, "box differs from the box the index was built with");
    (void) cofm;
    (void) axis;  // the index holds its own copies
    if (idx->nlos == 0 || idx->npairs == 0) return FSB_OK;
    FSB_REQUIRE(pos && cells, "NULL array");
    const int npts = (int) (box / kReso);  // :162
    FSB_REQUIRE(npts > 0, "box smaller than the march resolution");
    const double reso = box / npts;
    // bound the owner scratch to ~1 GiB
    const int64_t max_lines = std::max<int64_t>(1, (int64_t) (1 << 28) / npts);
    const int batch = (int) std::min<int64_t>(std::min<int64_t>(idx->nlos, max_lines), 65535);
    Scratch owner, first, last, runs, err;
    FSB_TRY(owner.alloc(sizeof(int32_t) * (size_t) batch * (size_t) npts, stream));
    FSB_TRY(first.alloc(sizeof(int32_t) * (size_t) idx->npairs, stream));
    FSB_TRY(last.alloc(sizeof(int32_t) * (size_t) idx->npairs, stream));
    FSB_TRY(runs.alloc(sizeof(int32_t) * (size_t) idx->npairs, stream));
    FSB_TRY(err.alloc(sizeof(int32_t), stream));
    FSB_CUDA_TRY(cudaMemsetAsync(err.ptr, 0, sizeof(int32_t), stream));
    const int pts_per_block = kOwnThreads * kPtsPerThread;
    for (int l0 = 0; l0 < idx->nlos; l0 += batch) {
        const int nl = std::min(batch, idx->nlos - l0);
        dim3 grid((npts + pts_per_block - 1) / pts_per_block, nl, 1);
        count_launch(); k_owners<<<grid, kOwnThreads, 0, stream>>>(idx->offsets, idx->particle, idx->cofm, idx->axis, pos, box, npts, reso, l0,
                                                   owner.as<int32_t>());
        count_launch(); k_extents<<<nl, 1024, 0, stream>>>(idx->offsets, box, npts, reso, l0, owner.as<int32_t>(), first.as<int32_t>(),
                                           last.as<int32_t>(), runs.as<int32_t>(), cells, err.as<int32_t>());
        FSB_CUDA_TRY(cudaGetLastError());
    }
    int32_t h_err = 0;
    FSB_CUDA_TRY(cudaMemcpyAsync(&h_err, err.ptr, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    FSB_CUDA_TRY(cudaStreamSynchronize(stream));
    if (h_err) {
        set_error("Voronoi cell ownership is not contiguous along at least one sightline "
                  "(the reference exits here: index_table.cpp:204-208)");
        return FSB_EVORONOI;
    }
    return FSB_OK;
}
