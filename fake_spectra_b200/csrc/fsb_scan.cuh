// Small device-wide scan used for sightline / cell / work-item tables (sizes up to a few million).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fsb {

// Single-CTA exclusive scan (n up to a few million: sightlines, cells, particle blocks).
// out has n+1 entries; also reports the maximum input element.
template <typename TIn, typename TOut>
__global__ void k_scan_single(const TIn *__restrict__ in, TOut *__restrict__ out, int64_t n, TOut *__restrict__ max_out)
{
    __shared__ TOut warp_tot[32];
    __shared__ TOut carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    TOut vmax = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + tid;
        const TOut v = i < n ? (TOut) in[i] : (TOut) 0;
        vmax = v > vmax ? v : vmax;
        TOut incl = v;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const TOut up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            TOut w = lane < (blockDim.x >> 5) ? warp_tot[lane] : (TOut) 0;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const TOut up = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += up;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const TOut carry = carry_s;
        const TOut before = carry + (wid ? warp_tot[wid - 1] : (TOut) 0) + incl - v;
        if (i < n) out[i] = before;
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = before + v;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
    if (max_out) {
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const TOut o = __shfl_down_sync(0xffffffffu, vmax, d);
            vmax = o > vmax ? o : vmax;
        }
        __syncthreads();
        if (lane == 0) warp_tot[wid] = vmax;
        __syncthreads();
        if (tid == 0) {
            TOut m = 0;
            for (int w = 0; w < (int) (blockDim.x >> 5); ++w) m = warp_tot[w] > m ? warp_tot[w] : m;
            *max_out = m;
        }
    }
}

// Ordered flag compaction (near_lines, particle selection): per-block popcounts, single-CTA scan of the block counts,
// then an ordered scatter of the flagged positions (or of map[position]).
static __global__ void __launch_bounds__(1024) k_flag_block_counts(const uint8_t *__restrict__ flag, int64_t npart,
                                                            int32_t *__restrict__ block_count)
{
    __shared__ int warp_cnt[32];
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const bool f = p < npart && flag[p];
    const unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += warp_cnt[w];
        block_count[blockIdx.x] = t;
    }
}

static __global__ void __launch_bounds__(1024) k_flag_compact(const uint8_t *__restrict__ flag, int64_t npart,
                                                       const int64_t *__restrict__ block_start,
                                                       const int32_t *__restrict__ map, int32_t *__restrict__ out)
{
    __shared__ int warp_cnt[32];
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const bool f = p < npart && flag[p];
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < wid; ++w) before += warp_cnt[w];
    if (f) out[block_start[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u))] = map ? map[p] : (int32_t) p;
}

}  // namespace fsb
