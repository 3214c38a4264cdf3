// Scratch: do XU conversions (F2F.F64.F32) overlap with FFMA / DFMA from the same and other warps?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ double f2d_alu(float f) {
    const unsigned b = __float_as_uint(f);
    const unsigned hi = (((b >> 3) & 0x0fffffffu) + 0x38000000u) | (b & 0x80000000u);
    return __hiloint2double((int)hi, (int)(b << 29));
}
template <int MODE>
__global__ void k(float* out, float a, double b) {
    float f0 = threadIdx.x * 1e-3f + a, f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    float g0 = f0 * 2, g1 = f1 * 2, g2 = f2 * 2, g3 = f3 * 2, g4 = f0 * 3, g5 = f1 * 3, g6 = f2 * 3, g7 = f3 * 3;
    double d0 = threadIdx.x * 1e-3 + b, d1 = d0 + 1., d2 = d0 + 2., d3 = d0 + 3.;
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0 || MODE == 2) {  // 4 x F2F feeding DADD
            d0 += (double)f0; d1 += (double)f1; d2 += (double)f2; d3 += (double)f3;
            f0 += a; f1 += a; f2 += a; f3 += a;
        }
        if (MODE == 1 || MODE == 2) {  // 16 independent FFMA
            g0 = fmaf(g0, a, g1); g1 = fmaf(g1, a, g2); g2 = fmaf(g2, a, g3); g3 = fmaf(g3, a, g4);
            g4 = fmaf(g4, a, g5); g5 = fmaf(g5, a, g6); g6 = fmaf(g6, a, g7); g7 = fmaf(g7, a, g0);
            g0 = fmaf(g0, a, g2); g1 = fmaf(g1, a, g3); g2 = fmaf(g2, a, g4); g3 = fmaf(g3, a, g5);
            g4 = fmaf(g4, a, g6); g5 = fmaf(g5, a, g7); g6 = fmaf(g6, a, g0); g7 = fmaf(g7, a, g1);
        }
        if (MODE == 3) {  // 4 x ALU-emulated conversion feeding DADD
            d0 += f2d_alu(f0); d1 += f2d_alu(f1); d2 += f2d_alu(f2); d3 += f2d_alu(f3);
            f0 += a; f1 += a; f2 += a; f3 += a;
        }
        if (MODE == 4) {  // ALU conversion + 16 FFMA
            d0 += f2d_alu(f0); d1 += f2d_alu(f1); d2 += f2d_alu(f2); d3 += f2d_alu(f3);
            f0 += a; f1 += a; f2 += a; f3 += a;
            g0 = fmaf(g0, a, g1); g1 = fmaf(g1, a, g2); g2 = fmaf(g2, a, g3); g3 = fmaf(g3, a, g4);
            g4 = fmaf(g4, a, g5); g5 = fmaf(g5, a, g6); g6 = fmaf(g6, a, g7); g7 = fmaf(g7, a, g0);
            g0 = fmaf(g0, a, g2); g1 = fmaf(g1, a, g3); g2 = fmaf(g2, a, g4); g3 = fmaf(g3, a, g5);
            g4 = fmaf(g4, a, g6); g5 = fmaf(g5, a, g7); g6 = fmaf(g6, a, g0); g7 = fmaf(g7, a, g1);
        }
        if (MODE == 5) {  // 8 DFMA + 16 FFMA
            d0 = fma(d0, b, d1); d1 = fma(d1, b, d2); d2 = fma(d2, b, d3); d3 = fma(d3, b, d0);
            d0 = fma(d0, b, d2); d1 = fma(d1, b, d3); d2 = fma(d2, b, d0); d3 = fma(d3, b, d1);
            g0 = fmaf(g0, a, g1); g1 = fmaf(g1, a, g2); g2 = fmaf(g2, a, g3); g3 = fmaf(g3, a, g4);
            g4 = fmaf(g4, a, g5); g5 = fmaf(g5, a, g6); g6 = fmaf(g6, a, g7); g7 = fmaf(g7, a, g0);
            g0 = fmaf(g0, a, g2); g1 = fmaf(g1, a, g3); g2 = fmaf(g2, a, g4); g3 = fmaf(g3, a, g5);
            g4 = fmaf(g4, a, g6); g5 = fmaf(g5, a, g7); g6 = fmaf(g6, a, g0); g7 = fmaf(g7, a, g1);
        }
        if (MODE == 6) {  // 8 DFMA only
            d0 = fma(d0, b, d1); d1 = fma(d1, b, d2); d2 = fma(d2, b, d3); d3 = fma(d3, b, d0);
            d0 = fma(d0, b, d2); d1 = fma(d1, b, d3); d2 = fma(d2, b, d0); d3 = fma(d3, b, d1);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + g0 + g1 + g2 + g3 + g4 + g5 + g6 + g7 + (float)(d0 + d1 + d2 + d3);
}
template <int MODE> void run(const char* name, float* out) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 320, 1024}) {
        k<MODE><<<sms, threads>>>(out, 1.0001f, 1.0000001);
        cudaEventRecord(e0);
        k<MODE><<<sms, threads>>>(out, 1.0001f, 1.0000001);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // cycles per loop iteration per warp-slot, assuming 1.9 GHz
        printf("%-34s threads/SM %4d: %.3f ms  -> %.1f cycles/iter (per SMSP: %.1f warps)\n", name, threads, ms,
               ms * 1e-3 * 1.9e9 / ITERS, threads / 128.0);
    }
}
int main() {
    float* out; cudaMalloc(&out, 1 << 22);
    run<0>("4 F2F(+4 DADD,4 FADD)", out); run<1>("16 FFMA", out); run<2>("4 F2F + 16 FFMA", out);
    run<3>("4 ALU-cvt(+4 DADD,4 FADD)", out); run<4>("4 ALU-cvt + 16 FFMA", out);
    run<6>("8 DFMA", out); run<5>("8 DFMA + 16 FFMA", out);
    return 0;
}
