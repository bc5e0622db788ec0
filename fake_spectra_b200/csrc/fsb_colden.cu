// Column density accumulation (replaces part_int.cpp:53-84 + absorption.cpp:53-210).
// One warp per work item; lanes are consecutive pixels of the current particle; K weight columns
// share one geometry pass (the kernel fraction does not depend on the weight: absorption.cpp:208).
#include "fsb_items.cuh"

namespace fsb {

namespace {

// ---- column density -----------------------------------------------------------------------------
constexpr int kColdenWarps = 4;

// sqrt(x) for x >= 0: hardware reciprocal-square-root seed and two coupled Newton steps (within an ulp; the
// library routine's special-case handling is not needed here), 0 for x == 0.
__device__ __forceinline__ double fast_sqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    return x > 0 ? g : 0.0;
}

// kern_frac (absorption.cpp:53-148) with q = sqrt(dr2 + z^2) * (1/smooth): one multiplication per node instead
// of a division (differs from the reference's quotient by an ulp of q).
template <int KERNEL>
__device__ __forceinline__ double kern_frac_fast(double zlow, double zhigh, double inv_smooth, double dr2, double zrange)
{
    zlow = fmax(zlow, -zrange);
    zhigh = fmin(zhigh, zrange);
    if (KERNEL == FSB_KERNEL_TOPHAT) return 3. / 4. / kPi * fmax(0., zhigh - zlow);
    if (KERNEL == FSB_KERNEL_VORONOI) return fmax(0., zhigh - zlow);
    if (zlow > zhigh) return 0;
    const double deltaz = (zhigh - zlow) / kNGrid;
    double total = sph_kernel<KERNEL>(fast_sqrt(fma(zlow, zlow, dr2)) * inv_smooth) / 2.;
    #pragma unroll
    for (int i = 1; i < kNGrid; ++i) {
        const double zz = fma((double) i, deltaz, zlow);
        total += sph_kernel<KERNEL>(fast_sqrt(fma(zz, zz, dr2)) * inv_smooth);
    }
    total += sph_kernel<KERNEL>(fast_sqrt(fma(zhigh, zhigh, dr2)) * inv_smooth) / 2.;
    return deltaz * total;
}

// One warp per work item.  Per batch of 32 candidates each lane gathers one particle and derives its pixel
// range (32 gathers in flight instead of a dependent chain per particle); then, particle by particle, the
// warp's lanes take consecutive pixels, the particle's constants arriving by shuffle.
template <int KERNEL>
__global__ void __launch_bounds__(32 * kColdenWarps)
k_colden(InterpConsts C, Items items, int n_items, const int64_t *__restrict__ offsets,
         const int32_t *__restrict__ particle, const double *__restrict__ dr2s, const int32_t *__restrict__ axis,
         const float *__restrict__ pos, const float *__restrict__ dens, int64_t dens_stride,
         const float *__restrict__ hsml, const float *__restrict__ cells, double *__restrict__ out, int64_t out_stride,
         double *__restrict__ scratch, int64_t scratch_stride, unsigned long long *__restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kColdenWarps + (threadIdx.x >> 5);
    if (item >= n_items) return;
    int line;
    int64_t kbeg, kend;
    if (!locate_item(items, offsets, C.nlos, item, line, kbeg, kend)) return;
    double *row = items.item_start ? scratch + (int64_t) item * C.nbins : out + (int64_t) line * C.nbins;
    const int64_t wstride = items.item_start ? scratch_stride : out_stride;
    const int ax = axis[line] - 1;
    const int nbins = C.nbins;
    const int nw = C.nlines;
    const int chunk = nbins < 32 ? nbins : 32;  // keep the pixels of one step distinct modulo nbins
    const double boxtokpc = C.boxtokpc;
    unsigned n_pix = 0;

    for (int64_t k0 = kbeg; k0 < kend; k0 += 32) {
        const int nb = (int) min((int64_t) 32, kend - k0);
        // ---- my particle of the batch: absorption.cpp:167-193
        double my_pos1 = 0, my_zrange = 0, my_dr2 = 0, my_inv = 0;
        float my_dens[kMaxFused] = {0, 0, 0, 0};
        int my_zlow = 0, my_zhigh = -1;  // empty range = skip
        if (lane < nb) {
            const int64_t k = k0 + lane;
            const int64_t ip = particle[k];
            const float ppos = pos[3 * ip + ax];
            float smooth;
            if (KERNEL == FSB_KERNEL_VORONOI) {
                my_dr2 = (double) cells[2 * k];
                smooth = cells[2 * k + 1];
            } else {
                my_dr2 = dr2s[k];
                smooth = hsml[ip];
            }
            for (int w = 0; w < nw; ++w) my_dens[w] = dens[(int64_t) w * dens_stride + ip];
            my_pos1 = (double) ppos;
            bool ok = true;
            if (KERNEL == FSB_KERNEL_VORONOI) {
                const double lim = 2 * C.vbox / C.velfac;
                ok = !(my_dr2 > lim || (double) smooth > lim);
                my_pos1 = __dmul_rn(__dadd_rn(my_dr2, (double) smooth), 0.5);
                my_zrange = __dmul_rn(__dsub_rn((double) smooth, my_dr2), 0.5);
            } else {
                const double arg = __dsub_rn((double) __fmul_rn(smooth, smooth), my_dr2);
                ok = arg > 0;
                my_zrange = ok ? sqrt(arg) : 0.0;
            }
            my_inv = 1.0 / (double) smooth;
            if (ok) {
                my_zlow = (int) floor(__ddiv_rn(__dsub_rn(my_pos1, my_zrange), boxtokpc));
                my_zhigh = (int) ceil(__ddiv_rn(__dadd_rn(my_pos1, my_zrange), boxtokpc));
            }
        }
        // ---- the batch, particle by particle, lanes = pixels
        for (int b = 0; b < nb; ++b) {
            const int zlow = __shfl_sync(kFull, my_zlow, b), zhigh = __shfl_sync(kFull, my_zhigh, b);
            if (zhigh < zlow) continue;
            const double pos1 = __shfl_sync(kFull, my_pos1, b), zrange = __shfl_sync(kFull, my_zrange, b);
            const double dr2 = __shfl_sync(kFull, my_dr2, b), inv_smooth = __shfl_sync(kFull, my_inv, b);
            float pd[kMaxFused];
            #pragma unroll
            for (int w = 0; w < kMaxFused; ++w) pd[w] = w < nw ? __shfl_sync(kFull, my_dens[w], b) : 0.f;
            for (int zb = zlow; zb <= zhigh; zb += chunk) {
                const int z = zb + lane;
                if (lane < chunk && z <= zhigh) {
                    const double plow = __dsub_rn(__dmul_rn(boxtokpc, (double) z), pos1);
                    const double frac = kern_frac_fast<KERNEL>(plow, __dadd_rn(plow, boxtokpc), inv_smooth, dr2, zrange);
                    const int j = wrap_bin(z, nbins);
                    #pragma unroll
                    for (int w = 0; w < kMaxFused; ++w)
                        if (w < nw) row[(int64_t) w * wstride + j] += (double) pd[w] * frac;
                    ++n_pix;
                }
                __syncwarp();
            }
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
    const int n_items = (int) plan.n_items;
    const unsigned grid = (unsigned) ((n_items + kColdenWarps - 1) / kColdenWarps);
    double *scratch = plan.scratch_rows.as<double>();
    const int64_t out_stride = (int64_t) idx->nlos * c.nbins;
    const int64_t scratch_stride = plan.n_items * c.nbins;
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(counters);
#define FSB_LAUNCH_COLDEN(K)                                                                                          \
    count_launch(); k_colden<K><<<grid, 32 * kColdenWarps, 0, stream>>>(c, plan.items, n_items, idx->offsets, idx->particle, idx->dr2, idx->axis, pos, dens,  \
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
