// Can another instruction issue in the cycle after a DFMA?  (B200, sm_100a)
// Each warp runs NF independent DFMA chains and NI independent integer chains (LOP3 / IMAD / FSEL-free),
// with enough chains and warps that no dependency is waited for; the cycles per loop iteration on one SM
// sub-partition tell the two models apart:
//   A  an fp64 instruction blocks the sub-partition's issue port for two cycles:  W * (2 NF + NI)
//   B  it only occupies the fp64 pipe for two cycles:                             W * max(2 NF, NF + NI)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITER 2048
template <int NF, int NI, int IOP>
__global__ void mix(double *out, double a, double b, unsigned c, long long *cycles)
{
    double x[NF > 0 ? NF : 1];
    unsigned y[NI > 0 ? NI : 1];
#pragma unroll
    for (int k = 0; k < NF; ++k) x[k] = a + k + threadIdx.x * 1e-9;
#pragma unroll
    for (int k = 0; k < NI; ++k) y[k] = c + k + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_ITER / 4; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < (NF > NI ? NF : NI); ++k) {
                if (k < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[k]) : "d"(b), "d"(a));
                if (k < NI) {
                    if (IOP == 0) asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[k]) : "r"(c));                 // LOP3
                    else if (IOP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(y[k]) : "r"(c));     // IMAD
                    else if (IOP == 2) asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(y[k]));           // SHF
                    else if (IOP == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(c));          // IMAD.HI
                    else if (IOP == 4) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(y[k]), "r"(c)); y[k] = (unsigned)(w >> 32) ^ (unsigned)w; }   // IMAD.WIDE.U32 (+ LOP3)
                    else if (IOP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(c));            // IADD3 / IMAD.IADD
                    else if (IOP == 6) asm volatile("{.reg .pred q; setp.lt.u32 q, %0, %1; selp.u32 %0, %0, %1, q;}" : "+r"(y[k]) : "r"(c));   // ISETP + SEL
                    else if (IOP == 7) asm volatile("popc.b32 %0, %0;" : "+r"(y[k]));                         // POPC
                    else if (IOP == 8) asm volatile("{.reg .u64 w; mad.wide.u32 w, %0, 8, %1; cvt.u32.u64 %0, w;}" : "+r"(y[k]) : "l"((unsigned long long)c));   // IMAD.WIDE.U32 with 64-bit addend
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += x[k];
#pragma unroll
    for (int k = 0; k < NI; ++k) s += y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int NF, int NI, int IOP>
void run(const char *name, int threads)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 2048 * sizeof(double)); cudaMalloc(&cyc, 8);
    mix<NF, NI, IOP><<<1, threads>>>(out, 1.0000001, 0.9999999, 0x9E3779B9u, cyc);
    mix<NF, NI, IOP><<<1, threads>>>(out, 1.0000001, 0.9999999, 0x9E3779B9u, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int w = (threads + 127) / 128;
    const double per_iter = (double)h / N_ITER;
    printf("%-8s NF %d NI %d, %d warps/sub-partition: %.2f cycles per iteration per sub-partition | model A (port blocked) %d, model B (pipe only) %d | %.2f instr/cycle\n",
           name, NF, NI, w, per_iter, w * (2 * NF + NI), w * (2 * NF > NF + NI ? 2 * NF : NF + NI), w * (NF + NI) / per_iter);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<4, 0, 0>("DFMA", 512); run<0, 4, 0>("LOP3", 512); run<0, 4, 1>("IMAD", 512); run<0, 4, 3>("IMAD.HI", 512); run<0, 4, 2>("SHF", 512);
    run<4, 4, 0>("DFMA+LOP3", 512); run<4, 8, 0>("DFMA+LOP3", 512); run<2, 6, 0>("DFMA+LOP3", 512);
    run<4, 4, 1>("DFMA+IMAD", 512); run<4, 4, 2>("DFMA+SHF", 512); run<4, 4, 3>("DFMA+IMAD.HI", 512);
    run<4, 4, 0>("DFMA+LOP3", 128); run<4, 4, 0>("DFMA+LOP3", 256);
    run<0, 4, 4>("WIDE+LOP3", 512); run<0, 4, 5>("ADD", 512); run<0, 4, 6>("ISETP+SEL", 512); run<0, 4, 7>("POPC", 512); run<0, 4, 8>("WIDE.ADDR", 512);
    run<4, 4, 4>("DFMA+WIDE", 512); run<4, 4, 8>("DFMA+WADDR", 512);
    return 0;
}
