// issue_port_probe.cu -- do non-FP64 instructions issue "for free" in the second cycle of a DFMA,
// or does every instruction take an issue slot away from the FP64 pipe?
// Loop body: 8 independent DFMA chains + M independent integer ops; 4 warps per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_port_probe issue_port_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int M>
__global__ void __launch_bounds__(128) k(double *out, int iters, double m, double c, unsigned seed) {
    double a[8];
    unsigned v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = seed + i * 7919u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
#pragma unroll
            for (int i = 0; i < M; ++i) v[i % 16] = (v[i % 16] ^ (v[(i + 5) % 16] >> 3)) + 0x9e3779b9u;   // LOP3/SHF/IADD mix
        }
    }
    double s = 0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) t ^= v[i];
    if (s == 123.456 || t == 12345u) out[0] = s + t;
}

template <int M>
void run(int sms, double *d) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<M><<<sms * 4, 128>>>(d, iters, 0.999999, 1e-9, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<M><<<sms * 4, 128>>>(d, iters, 0.999999, 1e-9, 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double dfma = (double)sms * 4 * 128 * iters * 32.0 / (ms * 1e-3);
    printf("M = %2d integer-op groups per 8 DFMA: %.3e DFMA thread-instr/s (%.1f%% of 148*64*1.965e9), %.2f ms\n", M, dfma,
           100.0 * dfma / (148.0 * 64 * 1.965e9), ms);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *d;
    cudaMalloc(&d, 8);
    int sms = p.multiProcessorCount;
    run<0>(sms, d); run<1>(sms, d); run<2>(sms, d); run<4>(sms, d); run<8>(sms, d); run<16>(sms, d);
    return 0;
}
