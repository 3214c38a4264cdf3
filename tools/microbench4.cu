// FFMA2 (fma.rn.f32x2) vs FFMA: throughput at several occupancies and dependent-chain latency (B200).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void k(float* out, float a) {
    float f0 = threadIdx.x * 1e-3f + a, f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    float2 p0 = make_float2(f0, f1), p1 = make_float2(f2, f3), p2 = make_float2(f1, f2), p3 = make_float2(f3, f0);
    const float2 aa = make_float2(a, a * 0.999f);
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
        if (OP == 0) { f0 = fmaf(f0, a, f1); f1 = fmaf(f1, a, f2); f2 = fmaf(f2, a, f3); f3 = fmaf(f3, a, f0); }
        if (OP == 1) { p0 = __ffma2_rn(p0, aa, p1); p1 = __ffma2_rn(p1, aa, p2); p2 = __ffma2_rn(p2, aa, p3); p3 = __ffma2_rn(p3, aa, p0); }
        if (OP == 2) { f0 = fmaf(f0, a, f0); }                    // dependent chain FFMA
        if (OP == 3) { p0 = __ffma2_rn(p0, aa, p0); }              // dependent chain FFMA2
        if (OP == 4) { p0 = __fadd2_rn(p0, p1); p1 = __fmul2_rn(p1, aa); p2 = __fadd2_rn(p2, p3); p3 = __fmul2_rn(p3, aa); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (float)(t1 - t0) / ITERS;
}
template <int OP> void run(const char* name, int instrPerIter, float* out) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {32, 128, 256, 512, 1024}) {
        k<OP><<<sms, threads>>>(out, 1.0001f);
        cudaEventRecord(e0);
        k<OP><<<sms, threads>>>(out, 1.0001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms, cyc; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&cyc, out + (1 << 20), 4, cudaMemcpyDeviceToHost);
        double winstr = (double)sms * (threads / 32) * ITERS * instrPerIter;
        printf("%-28s warps/SM %2d: %6.2f warp-instr/clk/SM (clock64: %.2f cycles/iter of %d instr)\n", name, threads / 32,
               winstr / (cyc * ITERS) / sms, cyc, instrPerIter);
    }
}
int main() {
    float* out; cudaMalloc(&out, (1 << 22) + 64);
    run<0>("FFMA x4 indep", 4, out); run<1>("FFMA2 x4 indep", 4, out); run<2>("FFMA dependent", 1, out);
    run<3>("FFMA2 dependent", 1, out); run<4>("FADD2/FMUL2 x4", 4, out);
    return 0;
}
