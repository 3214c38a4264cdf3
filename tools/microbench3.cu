// Scratch: dependent-chain latencies (cycles) of the instructions on the LM critical path, one warp per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP>
__global__ void k(double* out, double b, float a) {
    double d = threadIdx.x * 1e-3 + b;
    float f = threadIdx.x * 1e-3f + a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) d = fma(d, b, b);
        if (OP == 1) d = d + b;
        if (OP == 2) f = fmaf(f, a, a);
        if (OP == 3) { asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d)); }
        if (OP == 4) { f = (float)d; d = (double)f; }           // F2F down + up
        if (OP == 5) { d = __shfl_xor_sync(0xffffffffu, d, 1); }
        if (OP == 6) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f)); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (double)(t1 - t0) / N;
    if (d == 123.456 && f == 1.f) out[0] = d;
}
template <int OP> void run(const char* name, double* out) {
    k<OP><<<1, 32>>>(out, 1.0000001, 1.0001f);
    cudaDeviceSynchronize();
    double h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.1f cycles per dependent op\n", name, h);
}
int main() {
    double* out; cudaMalloc(&out, 1024);
    run<0>("DFMA", out); run<1>("DADD", out); run<2>("FFMA", out); run<3>("MUFU.RCP64H", out);
    run<4>("F2F f64->f32->f64 (pair)", out); run<5>("SHFL 64-bit (2 SHFL)", out); run<6>("MUFU.RCP f32", out);
    return 0;
}
