// Issue-rate microbenchmark (B200): what does one FFMA2 cost next to FFMA / FMUL / FADD with three distinct register
// operands, and can other pipes issue in the shadow of a packed instruction?  16 warps per SM (4 per scheduler).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
template <int OP>
__global__ void k(float* out, const float* in, float s) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = in[i & 255];
    __syncthreads();
    const int t = threadIdx.x;
    float a0 = in[t & 255], a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, b0 = a0 * .5f, b1 = a1 * .5f, b2 = a2 * .5f, b3 = a3 * .5f;
    float c0 = in[(t + 1) & 255], c1 = in[(t + 2) & 255], c2 = in[(t + 3) & 255], c3 = in[(t + 4) & 255], m = in[(t + 5) & 255];
    float2 A0 = f2(a0, a1), A1 = f2(a2, a3), A2 = f2(b0, b1), A3 = f2(b2, b3), B0 = f2(b0, a1), B1 = f2(b1, a2), B2 = f2(b2, a3), B3 = f2(b3, a0);
    float2 C0 = f2(c0, c1), C1 = f2(c1, c2), C2 = f2(c2, c3), C3 = f2(c3, c0);
    const float* sp = sm + (t & 31) * 2;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
        if (OP == 0) { c0 = fmaf(a0, b0, c0); c1 = fmaf(a1, b1, c1); c2 = fmaf(a2, b2, c2); c3 = fmaf(a3, b3, c3); }           // FFMA 3-reg
        if (OP == 1) { C0 = __ffma2_rn(A0, B0, C0); C1 = __ffma2_rn(A1, B1, C1); C2 = __ffma2_rn(A2, B2, C2); C3 = __ffma2_rn(A3, B3, C3); }
        if (OP == 2) { c0 = a0 * c0; c1 = a1 * c1; c2 = a2 * c2; c3 = a3 * c3; }                                             // FMUL 2-reg
        if (OP == 3) { c0 = a0 + c0; c1 = a1 + c1; c2 = a2 + c2; c3 = a3 + c3; }                                             // FADD 2-reg
        if (OP == 4) { C0 = __ffma2_rn(A0, f2(s, s), C0); C1 = __ffma2_rn(A1, f2(s, s), C1); C2 = __ffma2_rn(A2, f2(s, s), C2); C3 = __ffma2_rn(A3, f2(s, s), C3); }
        if (OP == 5) {  // FFMA2 + LDS.64 1:1
            C0 = __ffma2_rn(A0, B0, C0); float2 l0 = *(const float2*)(sp + ((i * 64) & 1023)); C1 = __ffma2_rn(A1, B1, C1); float2 l1 = *(const float2*)(sp + ((i * 64 + 64) & 1023));
            m += l0.x + l1.y;
        }
        if (OP == 6) { C0 = __ffma2_rn(A0, B0, C0); m = fmaxf(m, a0); C1 = __ffma2_rn(A1, B1, C1); m = fminf(m, b1); C2 = __ffma2_rn(A2, B2, C2); m = fmaxf(m, a2); C3 = __ffma2_rn(A3, B3, C3); m = fminf(m, b3);
                       a0 += 1.f; }   // FFMA2 + FMNMX (ALU) 1:1 (+1 FADD)
        if (OP == 7) { c0 = fmaf(a0, b0, c0); m = fmaxf(m, a0); c1 = fmaf(a1, b1, c1); m = fminf(m, b1); c2 = fmaf(a2, b2, c2); m = fmaxf(m, a2); c3 = fmaf(a3, b3, c3); m = fminf(m, b3); a0 += 1.f; }
        if (OP == 8) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c1)); C2 = __ffma2_rn(A2, B2, C2); C3 = __ffma2_rn(A3, B3, C3); }  // 2 MUFU + 2 FFMA2
        if (OP == 9) { C0 = __fmul2_rn(A0, C0); C1 = __fadd2_rn(A1, C1); C2 = __fmul2_rn(A2, C2); C3 = __fadd2_rn(A3, C3); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + t] = c0 + c1 + c2 + c3 + C0.x + C0.y + C1.x + C1.y + C2.x + C2.y + C3.x + C3.y + m;
    if (t == 0 && blockIdx.x == 0) out[1 << 20] = (float)(t1 - t0) / ITERS;
}
template <int OP> void run(const char* name, int instr, float* out, float* in) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int threads : {128, 512}) {
        k<OP><<<sms, threads>>>(out, in, 1.0001f);
        k<OP><<<sms, threads>>>(out, in, 1.0001f);
        cudaDeviceSynchronize();
        float cyc; cudaMemcpy(&cyc, out + (1 << 20), 4, cudaMemcpyDeviceToHost);
        printf("%-34s warps/SMSP %d: %6.2f cycles per iteration of %d instr => %.2f issue cycles per instr per SMSP\n", name, threads / 128, cyc, instr,
               cyc / instr / (threads / 128));
    }
}
int main() {
    float *out, *in; cudaMalloc(&out, (1 << 22) + 64); cudaMalloc(&in, 4096); cudaMemset(in, 0x3c, 4096);
    run<0>("FFMA 3 regs x4", 4, out, in); run<1>("FFMA2 3 regs x4", 4, out, in); run<2>("FMUL 2 regs x4", 4, out, in); run<3>("FADD 2 regs x4", 4, out, in);
    run<4>("FFMA2 scalar-broadcast x4", 4, out, in); run<9>("FMUL2/FADD2 x4", 4, out, in); run<5>("2 FFMA2 + 2 LDS.64 (+2 FADD)", 6, out, in);
    run<6>("4 FFMA2 + 4 FMNMX + FADD", 9, out, in); run<7>("4 FFMA + 4 FMNMX + FADD", 9, out, in); run<8>("2 MUFU.RCP + 2 FFMA2", 4, out, in);
    return 0;
}
