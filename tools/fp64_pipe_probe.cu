// fp64_pipe_probe.cu -- how many warps x how much ILP does the B200 FP64 pipe need?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe_probe fp64_pipe_probe.cu
// Prints DFMA throughput (fraction of 64 FMA/clk/SM at the measured clock) for a grid of
// (warps per SM, independent chains per thread), plus a dependent-chain latency estimate.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, int iters, double m, double c) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
double run(int warps_per_sm, int sms, double *d) {
    int iters = 20000 / ILP;
    dim3 grid(sms), block(32 * warps_per_sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<grid, block>>>(d, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<ILP><<<grid, block>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma_per_s = (double)sms * warps_per_sm * 32 * iters * 16.0 * ILP / (ms * 1e-3);
    return fma_per_s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *d;
    cudaMalloc(&d, 8);
    int sms = p.multiProcessorCount;
    double peak = run<8>(32, sms, d);          // reference: lots of warps
    printf("SMs %d, saturated DFMA rate %.3e FMA/s (%.2f TFLOP/s)\n", sms, peak, 2 * peak / 1e12);
    int ws[] = {4, 8, 12, 16, 20, 24, 32};
    printf("warps/SM  ILP1   ILP2   ILP4   ILP8  (fraction of saturated rate)\n");
    for (int w : ws) {
        printf("%7d  %.3f  %.3f  %.3f  %.3f\n", w, run<1>(w, sms, d) / peak, run<2>(w, sms, d) / peak,
               run<4>(w, sms, d) / peak, run<8>(w, sms, d) / peak);
    }
    // one warp per SMSP, ILP 1: cycles per dependent DFMA = latency
    double r1 = run<1>(4, sms, d);
    double clk = peak / (64.0 * sms);
    printf("dependent DFMA latency ~ %.1f cycles (clock %.0f MHz from the saturated rate)\n",
           (double)sms * 4 * 32 / r1 * clk, clk / 1e6);
    return 0;
}
