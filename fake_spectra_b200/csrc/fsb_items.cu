// Work-item table shared by the accumulation kernels (see fsb_items.cuh).
#include "fsb_items.cuh"

namespace fsb {

namespace {

__global__ void k_items_per_line(const int64_t *__restrict__ offsets, int nlos, int seg_pairs, int32_t *__restrict__ nitems)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlos) return;
    const int64_t n = offsets[l + 1] - offsets[l];
    nitems[l] = (int32_t) ((n + seg_pairs - 1) / seg_pairs);
}

__global__ void k_reduce_rows(const int32_t *__restrict__ item_start, const double *__restrict__ scratch, int64_t scratch_stride,
                              double *__restrict__ out, int64_t out_stride, int nbins)
{
    const int line = blockIdx.x;
    const int w = blockIdx.z;
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= nbins) return;
    const int ibeg = item_start[line], iend = item_start[line + 1];
    if (iend == ibeg) return;
    double acc = 0;
    for (int it = ibeg; it < iend; ++it) acc += scratch[(int64_t) w * scratch_stride + (int64_t) it * nbins + j];
    out[(int64_t) w * out_stride + (int64_t) line * nbins + j] += acc;
}

}  // namespace

int plan_items(const fsb_index *idx, int seg_pairs_req, int nbins, int nrows_per_item, cudaStream_t stream, ItemPlan &plan)
{
    const int64_t target_items = 8192;  // ~ 148 SMs x 16+ resident warps, a few waves
    int seg = seg_pairs_req;
    if (seg <= 0) {
        if (idx->nlos >= target_items / 2 || idx->npairs == 0) seg = 0;  // enough sightlines: one item per line
        else seg = (int) std::max<int64_t>(32, (idx->npairs + target_items - 1) / target_items);
    }
    if (seg <= 0 || seg >= idx->max_list) {
        plan.items.item_start = nullptr;
        plan.items.seg_pairs = 0;
        plan.items.ticket_pairs = 0;
        plan.items.max_runs = 0;
        plan.items.line_done = nullptr;
        plan.items.phase_start = nullptr;
        plan.items.order = nullptr;
        plan.n_items = idx->nlos;
        plan.segmented = false;
        return FSB_OK;
    }
    plan.items.ticket_pairs = 0;
    plan.items.max_runs = 0;
    plan.items.line_done = nullptr;
    plan.items.phase_start = nullptr;
    plan.items.order = nullptr;
    plan.segmented = true;
    plan.items.seg_pairs = seg;
    plan.n_items = (int64_t) idx->nlos + idx->npairs / seg;
    const size_t nl = (size_t) std::max(idx->nlos, 1);
    FSB_TRY(plan.nitems.alloc(sizeof(int32_t) * (nl + 1), stream));
    FSB_TRY(plan.item_start.alloc(sizeof(int32_t) * (nl + 1), stream));
    count_launch(); k_items_per_line<<<(idx->nlos + 255) / 256, 256, 0, stream>>>(idx->offsets, idx->nlos, seg, plan.nitems.as<int32_t>());
    count_launch(); k_scan_single<int32_t, int32_t><<<1, 1024, 0, stream>>>(plan.nitems.as<int32_t>(), plan.item_start.as<int32_t>(), idx->nlos, nullptr);
    FSB_CUDA_TRY(cudaGetLastError());
    plan.items.item_start = plan.item_start.as<int32_t>();
    const size_t bytes = sizeof(double) * (size_t) plan.n_items * (size_t) nbins * (size_t) nrows_per_item;
    FSB_TRY(plan.scratch_rows.alloc(bytes, stream));
    FSB_CUDA_TRY(cudaMemsetAsync(plan.scratch_rows.ptr, 0, bytes, stream));
    return FSB_OK;
}

namespace {

// hist[r] = number of lines with exactly r runs (r = 0 .. max_runs)
__global__ void k_runs_hist(const int64_t *__restrict__ offsets, int line0, int nrange, int ticket, int32_t *__restrict__ hist)
{
    const int l = line0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= line0 + nrange) return;
    const int64_t n = offsets[l + 1] - offsets[l];
    atomicAdd(&hist[(int) ((n + ticket - 1) / ticket)], 1);
}

// One CTA.  ge[r] = lines with >= r runs; phase k holds ge[max_runs - k] items; bin_start[r] = first slot of the lines with
// exactly r runs in the descending order.
__global__ void k_runs_plan(const int32_t *__restrict__ hist, int max_runs, int32_t *__restrict__ phase_start,
                            int32_t *__restrict__ bin_start)
{
    if (threadIdx.x != 0) return;
    int ge = 0, items = 0, slot = 0;
    for (int k = 0; k < max_runs; ++k) {  // r = max_runs - k, descending
        const int r = max_runs - k;
        bin_start[r] = slot;
        slot += hist[r];
        ge += hist[r];
        phase_start[k] = items;
        items += ge;
    }
    phase_start[max_runs] = items;
    bin_start[0] = slot;
}

__global__ void k_runs_order(const int64_t *__restrict__ offsets, int line0, int nrange, int ticket, int32_t *__restrict__ bin_start,
                             int32_t *__restrict__ order)
{
    const int l = line0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= line0 + nrange) return;
    const int64_t n = offsets[l + 1] - offsets[l];
    const int r = (int) ((n + ticket - 1) / ticket);
    order[atomicAdd(&bin_start[r], 1)] = l;  // any order inside a bin: every line's own additions stay sequential
}

}  // namespace

// Ticketed runs (fsb_items.cuh): builds the dispatch order on the device, no host synchronisation.
int plan_tickets(const fsb_index *idx, int ticket, int line0, int nrange, cudaStream_t stream, ItemPlan &plan)
{
    const int64_t max_runs = (idx->max_list + ticket - 1) / ticket;  // of the whole index: an upper bound for the range
    if (max_runs < 2 || max_runs > (1 << 20) || max_runs * (int64_t) nrange > (int64_t) INT32_MAX || nrange <= 0) return FSB_OK;  // whole lists
    const int nlos = idx->nlos;
    FSB_TRY(plan.line_done.alloc(sizeof(int32_t) * (size_t) nlos, stream));
    FSB_TRY(plan.phase_start.alloc(sizeof(int32_t) * (size_t) (max_runs + 1), stream));
    FSB_TRY(plan.order.alloc(sizeof(int32_t) * (size_t) nlos, stream));
    Scratch hist, bin_start;
    FSB_TRY(hist.alloc(sizeof(int32_t) * (size_t) (max_runs + 1), stream));
    FSB_TRY(bin_start.alloc(sizeof(int32_t) * (size_t) (max_runs + 1), stream));
    FSB_CUDA_TRY(cudaMemsetAsync(plan.line_done.ptr, 0, sizeof(int32_t) * (size_t) nlos, stream));
    FSB_CUDA_TRY(cudaMemsetAsync(hist.ptr, 0, sizeof(int32_t) * (size_t) (max_runs + 1), stream));
    const int blocks = (nrange + 255) / 256;
    count_launch(); k_runs_hist<<<blocks, 256, 0, stream>>>(idx->offsets, line0, nrange, ticket, hist.as<int32_t>());
    count_launch(); k_runs_plan<<<1, 32, 0, stream>>>(hist.as<int32_t>(), (int) max_runs, plan.phase_start.as<int32_t>(), bin_start.as<int32_t>());
    count_launch(); k_runs_order<<<blocks, 256, 0, stream>>>(idx->offsets, line0, nrange, ticket, bin_start.as<int32_t>(), plan.order.as<int32_t>());
    FSB_CUDA_TRY(cudaGetLastError());
    plan.items.ticket_pairs = ticket;
    plan.items.max_runs = (int32_t) max_runs;
    plan.items.line_done = plan.line_done.as<int32_t>();
    plan.items.phase_start = plan.phase_start.as<int32_t>();
    plan.items.order = plan.order.as<int32_t>();
    plan.n_items = max_runs * (int64_t) nrange;  // upper bound; the kernel stops at phase_start[max_runs]
    return FSB_OK;
}

int reduce_items(const ItemPlan &plan, const fsb_index *idx, int nbins, int nrows_per_item, double *out, cudaStream_t stream)
{
    if (!plan.segmented) return FSB_OK;
    dim3 g(idx->nlos, (nbins + 255) / 256, nrows_per_item);
    count_launch(); k_reduce_rows<<<g, 256, 0, stream>>>(plan.items.item_start, plan.scratch_rows.as<double>(), plan.n_items * (int64_t) nbins, out,
                                         (int64_t) idx->nlos * nbins, nbins);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

}  // namespace fsb
