// pipe_rates.cu -- issue-rate microbenchmarks for the instruction classes the SRLA kernels lean on
// (IMAD, IDP.2A/4A, VIADDMNMX, SHF, DADD/DMUL/DFMA).  Prints lanes/clk/SM for each.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates tools/pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CH 8
template <int OP> __device__ __forceinline__ void step(int (&x)[CH], double (&d)[CH], int a, int b, double fa, double fb)
{
    #pragma unroll
    for (int i = 0; i < CH; ++i) {
        if (OP == 0) { x[i] = x[i] * a + b; }                                   // IMAD
        if (OP == 1) { x[i] = __dp2a_lo(x[i], a, x[i]); }                       // IDP.2A
        if (OP == 2) { x[i] = __dp4a(x[i], a, x[i]); }                          // IDP.4A
        if (OP == 3) { x[i] = __viaddmax_s32_relu(x[i], a, b); }                // VIADDMNMX.RELU
        if (OP == 4) { x[i] = __viaddmax_s16x2_relu(x[i], a, b); }              // VIADDMNMX.S16x2.RELU
        if (OP == 5) { x[i] = __funnelshift_r(x[i], a, b); }                    // SHF
        if (OP == 6) { x[i] = (x[i] + a) ^ b; }                                 // IADD + LOP3 (2 alu ops)
        if (OP == 7) { d[i] = d[i] + fa; }                                      // DADD
        if (OP == 8) { d[i] = d[i] * fa; }                                      // DMUL
        if (OP == 9) { d[i] = __fma_rn(d[i], fa, fb); }                         // DFMA
        if (OP == 10) { x[i] = max(x[i] - a, 0) >> (b & 31); }                  // sub, max, shift
        if (OP == 11) { x[i] = x[i] >> (a & 31); }                              // SHF.R.S32
        if (OP == 12) { d[i] = d[i] + (double)(x[i]); x[i] += a; }              // I2F.F64 + DADD + IADD
        if (OP == 13) { d[i] = d[i] + __hiloint2double(0x43300000, x[i] ^ 0x80000000); x[i] += a; }  // LOP3 + DADD + IADD
        if (OP == 14) { x[i] = (int)(((long long)x[i] * a) >> 7) + b; }          // IMAD.WIDE + shift
    }
}
template <int OP> __global__ void __launch_bounds__(256) bench(int *out, int a, int b, double fa, double fb)
{
    int x[CH]; double d[CH];
    #pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x + i; d[i] = threadIdx.x * 1e-3 + i; }
    for (int it = 0; it < ITERS; ++it) { step<OP>(x, d, a, b, fa, fb); }
    int s = 0; double t = 0;
    #pragma unroll
    for (int i = 0; i < CH; ++i) { s += x[i]; t += d[i]; }
    if (s == b * 977 || t == fb * 3.0) { out[0] = s; }
}
template <int OP> void run(const char *name, int ops_per_step)
{
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int *o; cudaMalloc(&o, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = pr.multiProcessorCount * 8;
    bench<OP><<<grid, 256>>>(o, 3, 5, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    bench<OP><<<grid, 256>>>(o, 3, 5, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double ops = (double)grid * 256 * ITERS * CH * ops_per_step;
    const double per_s = ops / (ms * 1e-3);
    printf("%-28s %8.3f ms  %7.2f Tops/s  %6.1f lanes/clk/SM (at %d MHz max clock)\n", name, ms, per_s / 1e12,
           per_s / pr.multiProcessorCount / (khz * 1e3), khz / 1000);
    cudaFree(o);
}
int main()
{
    run<0>("IMAD", 1); run<1>("IDP.2A", 1); run<2>("IDP.4A", 1); run<3>("VIADDMNMX.RELU", 1); run<4>("VIADDMNMX.S16x2.RELU", 1);
    run<5>("SHF (funnel)", 1); run<6>("IADD+LOP3", 2); run<7>("DADD", 1); run<8>("DMUL", 1); run<9>("DFMA", 1);
    run<10>("sub+max+shr", 3); run<11>("SHF.R.S32", 1);
    run<12>("I2F.F64+DADD+IADD", 1); run<13>("LOP3+DADD+IADD (magic cvt)", 1); run<14>("IMAD.WIDE path", 1);
    return 0;
}
