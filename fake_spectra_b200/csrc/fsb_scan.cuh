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

}  // namespace fsb
