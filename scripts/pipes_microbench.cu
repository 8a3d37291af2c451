// Microbenchmark: do the half-rate FP64 pipe and the ALU pipe of sm_100 issue side by side?  Each variant runs a loop whose
// body is NF independent f64 add/mul, NA integer ALU ops (LOP3 / ISETP+SEL / VIMNMX) and NM IMAD ops; 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NA, int NM>
__global__ void __launch_bounds__(256, 2) k(double *out, int *iout, int iters, double a, int b)
{
    double f[8];
    int x[8], m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = a + threadIdx.x + i; x[i] = b + threadIdx.x * (i + 1); m[i] = b ^ (threadIdx.x + i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i & 7] = (i & 1) ? f[i & 7] * a : f[i & 7] + a;
#pragma unroll
        for (int i = 0; i < NA; ++i) x[i & 7] = min(x[i & 7] ^ b, x[(i + 1) & 7]);       // LOP3 + VIMNMX: two ALU ops
#pragma unroll
        for (int i = 0; i < NM; ++i) m[i & 7] = m[i & 7] * b + it;                          // IMAD
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += f[i]; t += x[i] + m[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int NF, int NA, int NM>
void run(const char *name)
{
    double *o; int *io;
    cudaMalloc(&o, 296 * 256 * 8); cudaMalloc(&io, 296 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<NF, NA, NM><<<296, 256>>>(o, io, iters, 1.0000001, 3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // cycles per loop iteration per SM sub-partition (4 warps each)
    const double cyc = best * 1e-3 * clk * 1e3 / iters;
    printf("%-28s NF=%2d NA(x2)=%2d NM=%2d : %.3f ms, %.1f cycles per iteration per scheduler (4 warps) -> %.2f cycles per warp-iteration\n",
           name, NF, NA, NM, best, cyc, cyc / 4);
    cudaFree(o); cudaFree(io);
}
int main()
{
    run<8, 0, 0>("f64 only");
    run<0, 3, 0>("alu only (6 ops)");
    run<0, 0, 6>("imad only (6 ops)");
    run<8, 3, 0>("f64 + 6 alu");
    run<8, 0, 6>("f64 + 6 imad");
    run<8, 3, 6>("f64 + 6 alu + 6 imad");
    run<6, 3, 0>("6 f64 + 6 alu");
    run<8, 6, 0>("f64 + 12 alu");
    return 0;
}
