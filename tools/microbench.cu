// Scratch microbenchmark: per-SM issue rates of the instructions the PnP kernel leans on (B200).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void k(float* out, float a, double b) {
    float f0 = threadIdx.x * 1e-3f + a, f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    double d0 = threadIdx.x * 1e-3 + b, d1 = d0 + 1., d2 = d0 + 2., d3 = d0 + 3.;
    for (int i = 0; i < ITERS; ++i) {
        if (OP == 0) { f0 = fmaf(f0, a, f1); f1 = fmaf(f1, a, f2); f2 = fmaf(f2, a, f3); f3 = fmaf(f3, a, f0); }
        if (OP == 1) { d0 = fma(d0, b, d1); d1 = fma(d1, b, d2); d2 = fma(d2, b, d3); d3 = fma(d3, b, d0); }
        if (OP == 2) { d0 += (double)f0; d1 += (double)f1; d2 += (double)f2; d3 += (double)f3; f0 += a; f1 += a; f2 += a; f3 += a; }
        if (OP == 3) { asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d0)); asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d1));
                       asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d2)); asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d3)); }
        if (OP == 4) { f0 += (float)d0; f1 += (float)d1; f2 += (float)d2; f3 += (float)d3; d0 += b; d1 += b; d2 += b; d3 += b; }
        if (OP == 5) { d0 = d0 + d1; d1 = d1 + d2; d2 = d2 + d3; d3 = d3 + d0; }
        if (OP == 6) { f0 = __shfl_xor_sync(0xffffffffu, f0, 1); f1 = __shfl_xor_sync(0xffffffffu, f1, 2); f2 = __shfl_xor_sync(0xffffffffu, f2, 4); f3 = __shfl_xor_sync(0xffffffffu, f3, 8); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + (float)(d0 + d1 + d2 + d3);
}
template <int OP> void run(const char* name, int opsPerIter, float* out) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 512, 1024}) {
        k<OP><<<sms, threads>>>(out, 1.0001f, 1.0000001);
        cudaEventRecord(e0);
        k<OP><<<sms, threads>>>(out, 1.0001f, 1.0000001);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * threads * ITERS * opsPerIter;
        printf("%-24s threads/SM %4d: %8.1f Gop/s  = %6.1f lanes/clk/SM @%d MHz nominal (%.3f ms)\n", name, threads, ops / ms / 1e6,
               ops / ms / 1e3 / sms / clk, clk / 1000, ms);
    }
}
int main() {
    float* out; cudaMalloc(&out, 1 << 22);
    run<0>("FFMA", 4, out); run<1>("DFMA", 4, out); run<5>("DADD", 4, out); run<2>("F2F.F64.F32 (+DADD,FADD)", 4, out);
    run<4>("F2F.F32.F64 (+FADD,DADD)", 4, out); run<3>("MUFU.RCP64H", 4, out); run<6>("SHFL", 4, out);
    return 0;
}
