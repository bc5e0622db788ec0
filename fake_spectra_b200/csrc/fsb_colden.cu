// Column density accumulation (replaces part_int.cpp:53-84 + absorption.cpp:53-210).
// One warp per work item; lanes are consecutive pixels of the current particle; K weight columns
// share one geometry pass (the kernel fraction does not depend on the weight: absorption.cpp:208).
#include "fsb_items.cuh"

namespace fsb {

namespace {

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

}  // namespace

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
    FSB_TRY(reduce_items(plan, idx, c.nbins, c.nlines, out, stream));
    return FSB_OK;
}

}  // namespace fsb
