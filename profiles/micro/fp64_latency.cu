// Dependent-issue latency and per-SM throughput of the fp64 instructions the flight kernel is made of
// (B200, sm_100a): nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITER 4096
template <int OP, int ILP>
__global__ void chain(double *out, double a, double b, long long *cycles)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_ITER / 8; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (OP == 0) x[k] = fma(x[k], b, a);
                else if (OP == 1) x[k] = x[k] * b;
                else if (OP == 2) x[k] = x[k] + b;
                else if (OP == 3) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[k])); x[k] = y; }
                else if (OP == 4) x[k] = (x[k] > b) ? a : x[k] + 1.0;        // DSETP + select / predicated
                else if (OP == 5) { float f = (float)x[k]; f = fmaf(f, (float)b, (float)a); x[k] = f; }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int OP, int ILP>
void run(const char *name, int threads)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
    chain<OP, ILP><<<1, threads>>>(out, 1.0000001, 0.9999999, cyc);
    chain<OP, ILP><<<1, threads>>>(out, 1.0000001, 0.9999999, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s ILP %d, %4d threads/SM: %.2f cycles per op per warp, %.2f cycles per warp-instruction on the SM sub-partition\n", name, ILP,
           threads, (double)h / N_ITER, (double)h / N_ITER / ILP / ((threads + 127) / 128));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0, 1>("DFMA", 32); run<0, 2>("DFMA", 32); run<0, 4>("DFMA", 32); run<0, 8>("DFMA", 32);
    run<0, 1>("DFMA", 128); run<0, 1>("DFMA", 512); run<0, 4>("DFMA", 512); run<0, 1>("DFMA", 1024);
    run<1, 1>("DMUL", 32); run<2, 1>("DADD", 32); run<3, 1>("MUFU.RSQ64", 32); run<3, 4>("MUFU.RSQ64", 512);
    run<4, 1>("DSETP+sel", 32); run<5, 1>("F2F+FFMA", 32);
    return 0;
}
