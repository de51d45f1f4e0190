// Instruction-fetch limits on B200 (sm_100a): a loop whose body is N independent integer instructions
// (issue-bound if fetched from a cache that keeps up), footprint N x 16 bytes, run by W warps per SM
// sub-partition.  cycles per instruction per warp vs footprint shows the instruction-cache levels.
#include <cstdio>
#include <cuda_runtime.h>
#define I4(a, b, c, d) asm volatile("mad.lo.u32 %0, %0, %4, %5;\n\tmad.lo.u32 %1, %1, %4, %5;\n\tmad.lo.u32 %2, %2, %4, %5;\n\tmad.lo.u32 %3, %3, %4, %5;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"(e1), "r"(e2));
#define I16(a, b, c, d) I4(a, b, c, d) I4(a, b, c, d) I4(a, b, c, d) I4(a, b, c, d)
#define I64(a, b, c, d) I16(a, b, c, d) I16(a, b, c, d) I16(a, b, c, d) I16(a, b, c, d)
#define I256(a, b, c, d) I64(a, b, c, d) I64(a, b, c, d) I64(a, b, c, d) I64(a, b, c, d)
#define I1024(a, b, c, d) I256(a, b, c, d) I256(a, b, c, d) I256(a, b, c, d) I256(a, b, c, d)
template <int KB>
__global__ void body(unsigned *out, int iters, long long *cycles, unsigned e1, unsigned e2)
{
    unsigned a = threadIdx.x, b = 1, c = 2, d = 3;
    // stagger the warps so that they are at different places of the body
    const int skew = (threadIdx.x >> 5) * 37;
    for (int s = 0; s < skew; ++s) { I4(a, b, c, d) }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (KB >= 1) { I64(a, b, c, d) }
        if (KB >= 2) { I64(a, b, c, d) }
        if (KB >= 4) { I64(a, b, c, d) I64(a, b, c, d) }
        if (KB >= 8) { I256(a, b, c, d) }
        if (KB >= 16) { I256(a, b, c, d) I256(a, b, c, d) }
        if (KB >= 24) { I256(a, b, c, d) I256(a, b, c, d) }
        if (KB >= 32) { I256(a, b, c, d) I256(a, b, c, d) }
        if (KB >= 48) { I1024(a, b, c, d) }
        if (KB >= 64) { I1024(a, b, c, d) }
        if (KB >= 128) { I1024(a, b, c, d) I1024(a, b, c, d) I1024(a, b, c, d) I1024(a, b, c, d) }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int KB>
void run(int threads)
{
    unsigned *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int n_inst = (KB >= 1 ? 64 : 0) + (KB >= 2 ? 64 : 0) + (KB >= 4 ? 128 : 0) + (KB >= 8 ? 256 : 0) + (KB >= 16 ? 512 : 0) +
                       (KB >= 24 ? 512 : 0) + (KB >= 32 ? 512 : 0) + (KB >= 48 ? 1024 : 0) + (KB >= 64 ? 1024 : 0) + (KB >= 128 ? 4096 : 0);
    const int iters = 2000000 / n_inst + 1;
    body<KB><<<148, threads>>>(out, iters, cyc, 1664525u, 1013904223u);
    body<KB><<<148, threads>>>(out, iters, cyc, 1664525u, 1013904223u);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)iters * n_inst);
    printf("footprint %5.1f KB (%5d instr), %4d threads/SM (%d warps/sub-partition): %.2f cycles per instruction per warp, %.2f issue cycles per instruction\n",
           n_inst * 16 / 1024.0, n_inst, threads, threads / 128, per, per / (threads / 128));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int threads : {128, 512, 640}) {
        run<1>(threads); run<4>(threads); run<8>(threads); run<16>(threads); run<24>(threads); run<32>(threads);
        run<48>(threads); run<64>(threads); run<128>(threads);
    }
    return 0;
}
