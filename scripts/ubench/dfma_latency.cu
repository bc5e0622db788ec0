// Micro-benchmark: FP64 FMA throughput per SM sub-partition as a function of (warps per SMSP, independent
// chains per thread).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[ILP];
    #pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int u = 0; u < 16; ++u) {
            #pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
        }
    }
    double s = 0;
    #pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int warps_per_smsp, int sms)
{
    double *out;
    cudaMalloc(&out, 64);
    const int threads = 32 * 4 * warps_per_smsp;  // one CTA per SM, warps spread over the 4 SMSPs
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ILP><<<sms, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double inst_per_warp = (double) iters * 16 * ILP;
    // cycles per DFMA per SMSP = cycles / (inst_per_warp * warps_per_smsp)
    printf("warps/SMSP %2d ILP %d : %.2f cycles per warp-DFMA per SMSP (%.1f%% of 2-cycle peak), chain latency <= %.1f cycles\n",
           warps_per_smsp, ILP, cycles / (inst_per_warp * warps_per_smsp), 200.0 * inst_per_warp * warps_per_smsp / cycles,
           cycles / (iters * 16.0));
    cudaFree(out);
}

int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {1, 2, 4, 8}) {
        run<1>(w, sms);
        run<2>(w, sms);
        run<4>(w, sms);
        run<7>(w, sms);
        run<8>(w, sms);
    }
    return 0;
}
