// quartic_loop_probe.cu -- the scan kernel's inner loop in isolation: how close to the FP64
// pipe's rate does "P[j] *= quartic(D[j]; e1..e4 from shared memory)" run, for J = 16 grid
// points per lane, as a function of resident warps and of how many chains are interleaved?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o quartic_loop_probe quartic_loop_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int J, int U>
__device__ __forceinline__ void mul_quartic(double (&P)[J], const double (&D)[J], double e1, double e2,
                                            double e3, double e4) {
#pragma unroll
    for (int j0 = 0; j0 < J; j0 += U) {
        double q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = fma(D[j0 + u], e4, e3);
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = fma(D[j0 + u], q[u], e2);
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = fma(D[j0 + u], q[u], e1);
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = fma(D[j0 + u], q[u], 1.0);
#pragma unroll
        for (int u = 0; u < U; ++u) P[j0 + u] *= q[u];
    }
}

template <int U, int MINB>
__global__ void __launch_bounds__(128, MINB) k(double *out, const double *in, int n_grp, int reps) {
    __shared__ __align__(16) double s_poly[4][16][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double D[16], P[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { D[j] = in[lane + 32 * j]; P[j] = 1.0; }
    if (lane < 16)
        for (int q = 0; q < 4; ++q) s_poly[warp][lane][q] = in[512 + lane * 4 + q];
    __syncwarp();
    for (int r = 0; r < reps; ++r) {
        for (int gi = 0; gi < n_grp; ++gi) {
            const double2 e12 = *reinterpret_cast<const double2 *>(&s_poly[warp][gi][0]);
            const double2 e34 = *reinterpret_cast<const double2 *>(&s_poly[warp][gi][2]);
            mul_quartic<16, U>(P, D, e12.x, e12.y, e34.x, e34.y);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += P[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same loop with the next group's coefficients loaded before the current group is evaluated
template <int U, int MINB>
__global__ void __launch_bounds__(128, MINB) kp(double *out, const double *in, int n_grp, int reps) {
    __shared__ __align__(16) double s_poly[4][17][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double D[16], P[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { D[j] = in[lane + 32 * j]; P[j] = 1.0; }
    if (lane < 16)
        for (int q = 0; q < 4; ++q) s_poly[warp][lane][q] = in[512 + lane * 4 + q];
    if (lane == 16)
        for (int q = 0; q < 4; ++q) s_poly[warp][16][q] = 0.0;
    __syncwarp();
    for (int r = 0; r < reps; ++r) {
        double2 e12 = *reinterpret_cast<const double2 *>(&s_poly[warp][0][0]);
        double2 e34 = *reinterpret_cast<const double2 *>(&s_poly[warp][0][2]);
        for (int gi = 0; gi < n_grp; ++gi) {
            const double2 n12 = *reinterpret_cast<const double2 *>(&s_poly[warp][gi + 1][0]);
            const double2 n34 = *reinterpret_cast<const double2 *>(&s_poly[warp][gi + 1][2]);
            mul_quartic<16, U>(P, D, e12.x, e12.y, e34.x, e34.y);
            e12 = n12; e34 = n34;
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += P[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int U, int MINB>
double runp(int ctas_per_sm, int sms, double *out, double *in, int n_grp) {
    const int reps = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kp<U, MINB><<<sms * ctas_per_sm, 128>>>(out, in, n_grp, reps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kp<U, MINB><<<sms * ctas_per_sm, 128>>>(out, in, n_grp, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return (double)sms * ctas_per_sm * 128 * reps * (double)n_grp * 80.0 / (ms * 1e-3);
}

template <int U, int MINB>
double run(int ctas_per_sm, int sms, double *out, double *in, int n_grp) {
    const int reps = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<U, MINB><<<sms * ctas_per_sm, 128>>>(out, in, n_grp, reps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<U, MINB><<<sms * ctas_per_sm, 128>>>(out, in, n_grp, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // FP64 instructions per thread: reps * n_grp * 16 * 5
    return (double)sms * ctas_per_sm * 128 * reps * (double)n_grp * 80.0 / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out, *in;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 128);
    cudaMalloc(&in, sizeof(double) * 1024);
    double h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 1e-3 * ((i * 37) % 101) / 101.0;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    // reference rate: FP64 instr/s at 64 lanes/clk/SM needs the clock; use cudaDevAttrClockRate as a guide
    printf("FP64 thread-instr/s (x1e12), J=16; nominal 148*64*1.965e9 = %.3f\n", 148 * 64 * 1.965e9 / 1e12);
    printf("CTAs/SM(4 warps each) n_grp   U=2     U=4     U=8    U=16   | coefficients prefetched: U=4  U=8\n");
    for (int c = 1; c <= 4; ++c)
        for (int n_grp : {2, 8, 16}) {
            printf("%8d %12d  %.3f  %.3f  %.3f  %.3f  | %.3f  %.3f\n", c, n_grp,
                   run<2, 4>(c, sms, out, in, n_grp) / 1e12, run<4, 4>(c, sms, out, in, n_grp) / 1e12,
                   run<8, 4>(c, sms, out, in, n_grp) / 1e12, run<16, 3>(c, sms, out, in, n_grp) / 1e12,
                   runp<4, 4>(c, sms, out, in, n_grp) / 1e12, runp<8, 4>(c, sms, out, in, n_grp) / 1e12);
        }
    return 0;
}
