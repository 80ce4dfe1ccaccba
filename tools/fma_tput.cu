// microbenchmark: issue rate of FFMA / FHFMA (fma.rn.f32.f16) / HFMA2 per SM
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a[16]; unsigned h[16];
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; h[i] = 0x3c003c00u + i; }
    const unsigned w = 0x38003800u + threadIdx.x; const float wf = 1.0001f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], wf, 0.5f + i);
            if (MODE == 1) asm volatile("{.reg .f16 al,ah,bl,bh; mov.b32 {al,ah}, %1; mov.b32 {bl,bh}, %2; fma.rn.f32.f16 %0, al, bl, %0;}" : "+f"(a[i]) : "r"(h[i]), "r"(w));
            if (MODE == 2) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(w));
            if (MODE == 3) a[i] = fmaf(a[i], a[(i + 1) & 15], a[(i + 2) & 15]);   // 3 distinct regs
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
    const int iters = 4096;
    const char* names[4] = {"FFMA (reg,imm)", "FHFMA f32+=f16*f16", "HFMA2", "FFMA 3 regs"};
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int m = 0; m < 4; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
                if (m == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
                if (m == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
                if (m == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            double instr = (double)iters * 16 * warps;   // warp-instructions per SM
            printf("%-22s warps/SM %2d : %.3f warp-instr/clk/SM (%.2f per SMSP)\n", names[m], warps, instr / *cyc, instr / *cyc / 4);
        }
    return 0;
}
