#include <cstdio>
#include <cuda_runtime.h>
template<int OP> __global__ void lat(double *out, long long *clk, double a, double b, int ia){
    double x = threadIdx.x * 1e-3 + 1.0; int xi = threadIdx.x + 1;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 1024; ++i) {
        #pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (OP==0) x = x + a;
            if (OP==1) x = x * a;
            if (OP==2) x = x / a;
            if (OP==3) x = sqrt(x) + a;
            if (OP==4) x = log(x) + b;
            if (OP==5) xi = xi * ia + 1;
            if (OP==6) xi = __dp2a_lo(xi, ia, xi);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { clk[0] = t1 - t0; }
    out[threadIdx.x] = x + xi;
}
template<int OP> void run(const char *n){ double *o; long long *c, h; cudaMalloc(&o, 256*8); cudaMalloc(&c, 8);
    lat<OP><<<1,32>>>(o,c,1.0000001,3.0,3); lat<OP><<<1,32>>>(o,c,1.0000001,3.0,3); cudaMemcpy(&h,c,8,cudaMemcpyDeviceToHost);
    printf("%-10s %.1f clk per dependent op\n", n, h/16384.0); }
int main(){ run<0>("DADD"); run<1>("DMUL"); run<2>("DDIV"); run<3>("DSQRT+add"); run<4>("log+add"); run<5>("IMAD"); run<6>("IDP.2A"); return 0; }
